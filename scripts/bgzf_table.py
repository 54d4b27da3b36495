"""Offline generator of the fixed deflate code tables used by the BGZF stage (mimick_b200/csrc/mmk_bgzf_table.h).

    python tests/golden/make_bgzf_table.py  # builds the fitting corpus, prints the statistics, rewrites the header

The BGZF kernel codes every block with ONE pre-agreed "dynamic Huffman" table (so the block header is a constant
bit string and no per-block histogram / tree construction runs on the device).  The table is fitted here to the token
stream the kernel's FASTQ-aware parse produces on FASTQ text of every barcode format (the corpus is produced by
tests/golden/make_bgzf_table.py; this module is a development tool and test-side model, not part of the product path).

`tokenize()` below is the specification of that parse (the CUDA kernel k_bgzf follows it token for token) and
`deflate_block()` a bit-exact Python model of the encoder; tests use them to cross-check compressed sizes.
"""
from __future__ import annotations

import heapq
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MAX_BITS = 12            # 5 literals of one lane fit a 64-bit register
SUPPORTED = [9, 10] + list(range(32, 127))
MIN_MATCH = 4
MAX_MATCH = 258


# ---------------------------------------------------------------------------------------------- deflate alphabets
def length_symbol(L):
    """match length 3..258 -> (symbol, n_extra, extra)"""
    if L == 258:
        return 285, 0, 0
    l = L - 3
    if l < 8:
        return 257 + l, 0, 0
    k = l.bit_length() - 1
    e = k - 2
    return 257 + 4 * e + 4 + ((l - (1 << k)) >> e), e, (l - (1 << k)) & ((1 << e) - 1)


def dist_symbol(D):
    """match distance 1..32768 -> (symbol, n_extra, extra)"""
    d = D - 1
    if d < 4:
        return d, 0, 0
    k = d.bit_length() - 1
    e = k - 1
    return 2 * k + ((d >> e) & 1), e, d & ((1 << e) - 1)


def split_match(total):
    """cut a long match into pieces of 3..258"""
    out = []
    rem = total
    while rem > 0:
        take = min(MAX_MATCH, rem)
        if 0 < rem - take < 3:
            take = rem - 3
        out.append(take)
        rem -= take
    return out


# ---------------------------------------------------------------------------------------------- the parse
def tokenize(block: bytes):
    """FASTQ-aware LZ parse of one block of whole 4-line records.
    Returns a list of ('lit', byte) / ('match', length, distance), or None when the block is not 4-line text."""
    nl = [i for i, b in enumerate(block) if b == 10]
    if not block or block[-1] != 10 or len(nl) % 4:
        return None
    ls = [0] + [i + 1 for i in nl]
    toks = []
    for r in range(len(nl) // 4):
        h0, s0, p0, q0, e0 = ls[4 * r:4 * r + 5]
        P = S = 0
        if r:
            hp0, sp0, pp0 = ls[4 * r - 4], ls[4 * r - 3], ls[4 * r - 2]
            lh, lhp = s0 - h0, sp0 - hp0
            m = min(lh, lhp, MAX_MATCH)
            while P < m and block[h0 + P] == block[hp0 + P]:
                P += 1
            if P < MIN_MATCH:
                P = 0
            m = min(min(lh, lhp) - P, MAX_MATCH)
            while S < m and block[s0 - 1 - S] == block[sp0 - 1 - S]:
                S += 1
            if S < MIN_MATCH:
                S = 0
        if P:
            toks.append(("match", P, h0 - hp0))
        toks += [("lit", b) for b in block[h0 + P:s0 - S]]
        if S:
            toks.append(("match", S, s0 - sp0))
        toks += [("lit", b) for b in block[s0:p0]]
        if r and block[p0:e0] == block[pp0:h0]:
            toks += [("match", n, p0 - pp0) for n in split_match(e0 - p0)]
        else:
            toks += [("lit", b) for b in block[p0:q0]]
            q = block[q0:e0 - 1]
            if len(q) >= MIN_MATCH and q == q[:1] * len(q):
                toks.append(("lit", q[0]))
                toks += [("match", n, 1) for n in split_match(len(q) - 1)]
            else:
                toks += [("lit", b) for b in q]
            toks.append(("lit", 10))
    return toks


# ---------------------------------------------------------------------------------------------- code construction
def limited_lengths(freq, limit):
    """Length-limited Huffman code lengths (package-merge). freq: dict symbol -> weight (> 0)."""
    syms = sorted(freq, key=lambda s: (freq[s], s))
    if len(syms) == 1:
        return {syms[0]: 1}
    leaves = [(freq[s], [s]) for s in syms]
    packages = list(leaves)
    for _ in range(limit - 1):
        paired = [(packages[i][0] + packages[i + 1][0], packages[i][1] + packages[i + 1][1])
                  for i in range(0, len(packages) - 1, 2)]
        packages = sorted(leaves + paired, key=lambda t: t[0])
    lengths = {s: 0 for s in syms}
    for w, members in packages[:2 * len(syms) - 2]:
        for s in members:
            lengths[s] += 1
    assert abs(sum(2.0 ** -l for l in lengths.values()) - 1.0) < 1e-12
    return lengths


def canonical(lengths, n):
    """RFC 1951 canonical codes. lengths: list of n code lengths -> list of (code, len)."""
    bl_count = [0] * 16
    for l in lengths:
        bl_count[l] += 1
    bl_count[0] = 0
    code = 0
    next_code = [0] * 16
    for bits in range(1, 16):
        code = (code + bl_count[bits - 1]) << 1
        next_code[bits] = code
    out = []
    for s in range(n):
        l = lengths[s]
        if l:
            out.append((next_code[l], l))
            next_code[l] += 1
        else:
            out.append((0, 0))
    return out


class BitWriter:
    def __init__(self):
        self.acc = 0
        self.n = 0

    def bits(self, value, nbits):          # LSB first
        self.acc |= (value & ((1 << nbits) - 1)) << self.n
        self.n += nbits

    def code(self, code, nbits):           # Huffman codes go MSB first
        for i in range(nbits - 1, -1, -1):
            self.bits((code >> i) & 1, 1)

    def tobytes(self):
        return self.acc.to_bytes((self.n + 7) // 8, "little")


CL_ORDER = [16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15]


def dynamic_header(ll_len, d_len):
    """BFINAL=1, BTYPE=2 and the code-length section for the given literal/length and distance code lengths."""
    seq = list(ll_len) + list(d_len)
    rle = []
    i = 0
    while i < len(seq):
        v = seq[i]
        j = i
        while j < len(seq) and seq[j] == v:
            j += 1
        run = j - i
        if v == 0:
            while run >= 11:
                t = min(run, 138)
                rle.append((18, t - 11, 7))
                run -= t
            if run >= 3:
                rle.append((17, run - 3, 3))
                run = 0
            rle += [(0, 0, 0)] * run
        else:
            rle.append((v, 0, 0))
            run -= 1
            while run >= 3:
                t = min(run, 6)
                rle.append((16, t - 3, 2))
                run -= t
            rle += [(v, 0, 0)] * run
        i = j
    f = {}
    for s, _, _ in rle:
        f[s] = f.get(s, 0) + 1
    cl = limited_lengths(f, 7)
    cl_len = [cl.get(s, 0) for s in range(19)]
    cl_code = canonical(cl_len, 19)
    hclen = 19
    while hclen > 4 and cl_len[CL_ORDER[hclen - 1]] == 0:
        hclen -= 1
    w = BitWriter()
    w.bits(1, 1)
    w.bits(2, 2)
    w.bits(len(ll_len) - 257, 5)
    w.bits(len(d_len) - 1, 5)
    w.bits(hclen - 4, 4)
    for k in range(hclen):
        w.bits(cl_len[CL_ORDER[k]], 3)
    for s, ev, eb in rle:
        w.code(*cl_code[s])
        if eb:
            w.bits(ev, eb)
    return w


def deflate_block(block: bytes, ll_len, d_len, header: BitWriter | None = None):
    """Bit-exact model of the kernel's coded block: header + tokens + end-of-block. Returns bytes or None."""
    toks = tokenize(block)
    if toks is None:
        return None
    ll = canonical(ll_len, len(ll_len))
    dd = canonical(d_len, len(d_len))
    w = BitWriter()
    h = header or dynamic_header(ll_len, d_len)
    w.bits(h.acc, h.n)
    for t in toks:
        if t[0] == "lit":
            if ll[t[1]][1] == 0:
                return None
            w.code(*ll[t[1]])
        else:
            s, e, x = length_symbol(t[1])
            w.code(*ll[s])
            w.bits(x, e)
            s, e, x = dist_symbol(t[2])
            w.code(*dd[s])
            w.bits(x, e)
    w.code(*ll[256])
    return w.tobytes()


# ---------------------------------------------------------------------------------------------- fitting
def fit(texts):
    """texts: list of (format name, R1 text, R2 text).  Fits the code lengths to the token stream of 32-record blocks,
    checks the resulting coder against zlib, prints the statistics and rewrites mmk_bgzf_table.h.
    The corpus comes from tests/golden/make_bgzf_table.py (which runs the CPU oracle — test infrastructure)."""
    ll_f = {s: 0 for s in SUPPORTED + list(range(256, 286))}
    d_f = {s: 0 for s in range(30)}
    total_in = 0
    blocks = []
    for fmt, r1, r2 in texts:
        for txt in (r1, r2):
            recs = txt.split(b"\n")
            lines = recs[:-1]
            for b0 in range(0, len(lines) // 4, 32):
                blk = b"\n".join(lines[4 * b0:4 * (b0 + 32)]) + b"\n"
                blocks.append((fmt, blk))
    # weight haplotagging (the headline workload) like all others together
    for fmt, blk in blocks:
        wgt = 5 if fmt == "haplotagging" else 1
        total_in += len(blk)
        for t in tokenize(blk):
            if t[0] == "lit":
                ll_f[t[1]] += wgt
            else:
                ll_f[length_symbol(t[1])[0]] += wgt
                d_f[dist_symbol(t[2])[0]] += wgt
        ll_f[256] += wgt
    floor_ll = max(1, sum(ll_f.values()) >> 16)
    floor_d = max(1, sum(d_f.values()) >> 12)
    ll_w = {s: max(f, floor_ll) for s, f in ll_f.items()}
    d_w = {s: max(f, floor_d) for s, f in d_f.items()}
    ll_l = limited_lengths(ll_w, MAX_BITS)
    d_l = limited_lengths(d_w, MAX_BITS)
    ll_len = [ll_l.get(s, 0) for s in range(286)]
    d_len = [d_l.get(s, 0) for s in range(30)]
    hdr = dynamic_header(ll_len, d_len)
    print("literal/length code lengths of the frequent symbols:")
    for s in sorted(ll_f, key=lambda s: -ll_f[s])[:40]:
        name = repr(chr(s)) if s < 256 else ("EOB" if s == 256 else f"len{s}")
        print(f"  {name:>8} f={ll_f[s]:>9} bits={ll_len[s]}")
    print("distance code lengths:", d_len)
    print(f"header: {hdr.n} bits ({(hdr.n + 7) // 8} bytes)")
    # verify with zlib on every format and report the ratio
    by_fmt = {}
    for fmt, blk in blocks[::7]:
        z = deflate_block(blk, ll_len, d_len, hdr)
        assert zlib.decompress(z, wbits=-15) == blk
        a = by_fmt.setdefault(fmt, [0, 0, 0])
        a[0] += len(blk)
        a[1] += len(z) + 26
        a[2] += len(zlib.compress(blk, 6))
    for fmt, (a, b, c) in by_fmt.items():
        print(f"  {fmt:>24}: ratio {a / b:.2f} (zlib -6 on the same blocks: {a / c:.2f})")

    out = os.path.join(ROOT, "mimick_b200", "csrc", "mmk_bgzf_table.h")
    hb = hdr.tobytes()
    with open(out, "w") as fh:
        fh.write("// Generated by tests/golden/make_bgzf_table.py (scripts/bgzf_table.py:fit) — do not edit.\n")
        fh.write("// Fixed deflate code of the BGZF stage: code lengths of the literal/length and distance alphabets and the\n")
        fh.write("// constant block prefix (BFINAL=1, BTYPE=2, code-length section) as a little-endian bit string.\n")
        fh.write("#pragma once\n#include <stdint.h>\n")
        fh.write(f"#define MMK_BGZF_MAX_BITS {MAX_BITS}\n")
        fh.write(f"#define MMK_BGZF_HDR_BITS {hdr.n}\n")
        fh.write("static const uint8_t MMK_BGZF_LL_LEN[286] = {" + ",".join(map(str, ll_len)) + "};\n")
        fh.write("static const uint8_t MMK_BGZF_D_LEN[30] = {" + ",".join(map(str, d_len)) + "};\n")
        fh.write(f"static const uint8_t MMK_BGZF_HDR[{len(hb)}] = {{" + ",".join(map(str, hb)) + "};\n")
    print("wrote", out)




if __name__ == "__main__":
    raise SystemExit("run tests/golden/make_bgzf_table.py: it builds the fitting corpus and calls fit()")
