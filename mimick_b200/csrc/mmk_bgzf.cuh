// mmk_bgzf.cuh — BGZF (blocked gzip) coding of emitted FASTQ text on the device.
//
// The reference writes every FASTQ through BGZFCompressorStream (src/Main.jl:42-43,111-112); its files are
// `<prefix>.R{1,2}.fq.gz`.  Here one CTA turns one block of whole records (the tiles the emit kernel wrote) into
// one BGZF member: 18-byte gzip header with the BC extra field, a single deflate block, CRC-32 and ISIZE.
//
// Deflate side.  Every block uses the same pre-agreed dynamic Huffman code (mmk_bgzf_table.h, fitted offline by
// scripts/bgzf_table.py), so the block prefix is a constant bit string.  The LZ parse is FASTQ-aware instead of
// hash-based — scripts/bgzf_table.py:tokenize() is its specification:
//   header line   common prefix and common suffix with the previous record's header -> two matches, rest literals
//   sequence line literals (A/C/G/T cost 2-3 bits)
//   "+" and quality lines: identical to the previous record's -> one match; else a run (literal + distance-1 match)
// Every match is verified by comparing bytes, so any 4-line text is coded correctly; text that is not 4-line
// records, bytes without a code, or a block that would grow are written as a stored deflate block.
// CRC-32: per-lane slicing-by-4 CRC over word chunks aligned to the END of the block (leading zeros leave a zero
// state unchanged); every lane's value is moved to the end by one carry-less multiplication with x^(8·bytes after
// its chunk) mod P and the lanes are folded with XOR.
//
// Work decomposition.  ONE WARP codes one block and ONE LANE codes one record: header comparison, token choice and
// bit packing are serial byte loops per record, so a warp instruction advances 32 records at once and nothing needs
// a CTA barrier.  A record's bits are first packed into the shared-memory bytes of its own quality line (nobody
// reads those again once the token choice is made), then — after a warp scan of the bit counts — shifted to their
// final position in the member (quality lines too short for that — short reads — switch the warp to a direct mode:
// count, scan, write in place).  Only quality lines that are a run of one character are coded (this library writes
// no other kind); anything else falls back to a stored block.
#pragma once
#include "mmk_internal.h"

constexpr uint32_t BZ_WARPS = 6;            // blocks per CTA
constexpr uint32_t BZ_THREADS = BZ_WARPS * 32;
constexpr uint32_t BZ_GZ_HDR = 18;
constexpr uint32_t BZ_MIN_MATCH = 4, BZ_MAX_MATCH = 258;
// words of the table buffer: literal/length codes, distance codes, CRC-32 slicing tables, per-lane CRC shifts
constexpr uint32_t BZ_T_LL = 0, BZ_T_DD = 286, BZ_T_CRC = 286 + 30, BZ_T_SHIFT = BZ_T_CRC + 1024, BZ_T_HDR = BGZF_TAB_CODES;
static_assert(BZ_T_SHIFT + 32 == BGZF_TAB_CODES, "table layout");

__device__ __forceinline__ uint32_t bz_gf_mul(uint32_t a, uint32_t b) {  // a·b mod P, reflected bit order
    uint32_t p = 0;
#pragma unroll 8
    for (int i = 0; i < 32; ++i) {
        p ^= (a & 0x80000000u) ? b : 0u;
        a <<= 1;
        b = (b >> 1) ^ ((b & 1u) ? 0xEDB88320u : 0u);
    }
    return p;
}
__device__ __forceinline__ uint32_t bz_crc_word(const uint32_t *t, uint32_t x) {  // slicing-by-4 step, x = state ^ word
    return t[768 + (x & 0xffu)] ^ t[512 + ((x >> 8) & 0xffu)] ^ t[256 + ((x >> 16) & 0xffu)] ^ t[x >> 24];
}

struct BzSmem {
    const uint32_t *ll;     // [286] reversed code | len << 16
    const uint32_t *dd;     // [30]
    const uint32_t *crc_t;  // [4][256] slicing-by-4 tables
    // per warp
    uint32_t *in_w;         // 16-byte aligned staging buffer of the text (quality lines double as bit staging)
    uint8_t *in;            // (uint8_t *)in_w + phase: first byte of the text
    uint32_t phase;
    uint32_t *out;          // zeroed words: the member
    uint32_t *ls;           // line starts
    uint32_t *rec;          // per record: quality class, then bit count
    uint32_t *rec2;         // per record: token choice, then bit position
};

__device__ __forceinline__ void bz_or_bits(uint32_t *buf, uint32_t pos, uint64_t acc, uint32_t nb) {
    if (!nb) return;
    const uint32_t w = pos >> 5, s = pos & 31;
    const uint64_t lo = acc << s;
    atomicOr(&buf[w], (uint32_t)lo);
    const uint32_t w1 = (uint32_t)(lo >> 32);
    if (w1) atomicOr(&buf[w + 1], w1);
    const uint32_t hi = s ? (uint32_t)(acc >> (64 - s)) : 0u;
    if (hi) atomicOr(&buf[w + 2], hi);
}

// Lane-private bit writer into whole words [w, end)
struct BzWriter {
    uint32_t *w, *w0, *end;
    uint32_t acc, nb;
    bool bad;
    __device__ __forceinline__ void init(uint32_t *begin, uint32_t *stop) { w = w0 = begin, end = stop, acc = 0, nb = 0, bad = false; }
    __device__ __forceinline__ void put(uint32_t v, uint32_t n) {  // n <= 32 bits of v, nb < 32
        acc |= v << nb;
        const uint32_t room = 32 - nb;  // 1..32 bits went into the current word
        nb += n;
        if (nb >= 32) {
            if (w < end) *w = acc;
            else bad = true;
            ++w;
            nb -= 32;
            acc = __funnelshift_rc(v, 0u, room);  // the rest of v (0 when room == 32)
        }
    }
    __device__ __forceinline__ uint32_t finish() {  // returns the bits written
        const uint32_t bits = (uint32_t)(w - w0) * 32 + nb;
        if (nb) {
            if (w < end) *w = acc;
            else bad = true;
        }
        return bits;
    }
};

// Same interface, counting only (sizing pass of the direct mode)
struct BzCounter {
    uint32_t bits;
    bool bad;
    __device__ __forceinline__ void put(uint32_t, uint32_t n) { bits += n; }
};
// Same interface, writing at an arbitrary bit position of the zeroed member buffer: words are OR-ed in because the
// first and the last one are shared with the neighbouring records
struct BzDirect {
    uint32_t *w;
    uint32_t acc, nb;
    bool bad;
    __device__ __forceinline__ void init(uint32_t *buf, uint32_t bitpos) { w = buf + (bitpos >> 5), nb = bitpos & 31u, acc = 0, bad = false; }
    __device__ __forceinline__ void put(uint32_t v, uint32_t n) {
        acc |= v << nb;
        const uint32_t room = 32 - nb;
        nb += n;
        if (nb >= 32) {
            atomicOr(w, acc);
            ++w;
            nb -= 32;
            acc = __funnelshift_rc(v, 0u, room);
        }
    }
    __device__ __forceinline__ void finish() {
        if (nb) atomicOr(w, acc);
    }
};

template <class Wr>
__device__ __forceinline__ void bz_put_literals(const BzSmem &S, Wr &bw, uint32_t a, uint32_t b) {
    uint32_t j = a;
#pragma unroll 4
    for (; j + 2 <= b; j += 2) {  // two literals (<= 24 bits) per put
        const uint32_t e0 = S.ll[S.in[j]], e1 = S.ll[S.in[j + 1]];
        const uint32_t l0 = e0 >> 16, l1 = e1 >> 16;
        if (!l0 || !l1) bw.bad = true;
        bw.put((e0 & 0xffffu) | ((e1 & 0xffffu) << l0), l0 + l1);
    }
    if (j < b) {
        const uint32_t e = S.ll[S.in[j]];
        if (!(e >> 16)) bw.bad = true;
        bw.put(e & 0xffffu, e >> 16);
    }
}
// One match token (length 3..258, distance 1..32768)
template <class Wr>
__device__ __forceinline__ void bz_put_match(const BzSmem &S, Wr &bw, uint32_t len, uint32_t dist) {
    uint32_t sym, eb, ev;
    const uint32_t l = len - 3;
    if (len == 258) {
        sym = 285, eb = 0, ev = 0;
    } else if (l < 8) {
        sym = 257 + l, eb = 0, ev = 0;
    } else {
        const uint32_t k = 31 - __clz(l);
        eb = k - 2;
        sym = 257 + 4 * eb + 4 + ((l - (1u << k)) >> eb);
        ev = (l - (1u << k)) & ((1u << eb) - 1);
    }
    const uint32_t e1 = S.ll[sym];
    bw.put((e1 & 0xffffu) | (ev << (e1 >> 16)), (e1 >> 16) + eb);
    uint32_t dsym, deb, dev;
    const uint32_t d = dist - 1;
    if (d < 4) {
        dsym = d, deb = 0, dev = 0;
    } else {
        const uint32_t k = 31 - __clz(d);
        deb = k - 1;
        dsym = 2 * k + ((d >> deb) & 1);
        dev = d & ((1u << deb) - 1);
    }
    const uint32_t e2 = S.dd[dsym];
    bw.put((e2 & 0xffffu) | (dev << (e2 >> 16)), (e2 >> 16) + deb);
}
// A match longer than 258 is cut into pieces of 3..258 (scripts/bgzf_table.py:split_match)
template <class Wr>
__device__ __forceinline__ void bz_put_long_match(const BzSmem &S, Wr &bw, uint32_t total, uint32_t dist) {
    uint32_t rem = total;
    while (rem) {
        uint32_t take = min(BZ_MAX_MATCH, rem);
        if (rem - take > 0 && rem - take < 3) take = rem - 3;
        bz_put_match(S, bw, take, dist);
        rem -= take;
    }
}

// The tokens of record r, in stream order, given the choices of pass 2b (scripts/bgzf_table.py:tokenize).
template <class Wr>
__device__ __forceinline__ void bz_record_tokens(const BzSmem &S, uint32_t r, uint32_t choice, Wr &bw) {
    const uint32_t h0 = S.ls[4 * r], s0 = S.ls[4 * r + 1], p0 = S.ls[4 * r + 2], q0 = S.ls[4 * r + 3], e0 = S.ls[4 * r + 4];
    const uint32_t P = choice & 0x1ffu, Sx = (choice >> 9) & 0x1ffu, same = (choice >> 18) & 1u, qchar = (choice >> 19) & 0xffu;
    if (P) bz_put_match(S, bw, P, h0 - S.ls[4 * r - 4]);
    bz_put_literals(S, bw, h0 + P, s0 - Sx);
    if (Sx) bz_put_match(S, bw, Sx, s0 - S.ls[4 * r - 3]);
    bz_put_literals(S, bw, s0, p0);
    if (same) {
        bz_put_long_match(S, bw, e0 - p0, p0 - S.ls[4 * r - 2]);
    } else {
        bz_put_literals(S, bw, p0, q0);
        const uint32_t eq = S.ll[qchar], en = S.ll[10];
        if (!(eq >> 16) || !(en >> 16)) bw.bad = true;
        bw.put(eq & 0xffffu, eq >> 16);
        bz_put_long_match(S, bw, e0 - 1 - q0 - 1, 1);
        bw.put(en & 0xffffu, en >> 16);
    }
}

// Quality line in[q0, q0 + nq): (is a run of one character) | first character << 1 | length << 9.  Word-wise.
__device__ __forceinline__ uint32_t bz_qual_class(const BzSmem &S, uint32_t q0, uint32_t nq) {
    if (!nq) return 0;
    const uint32_t c = S.in[q0];
    const uint32_t splat = c * 0x01010101u;
    const uint32_t lo = S.phase + q0, hi = lo + nq - 1;  // absolute byte offsets of the first and last byte
    uint32_t diff = 0;
    for (uint32_t w = lo >> 2; w <= (hi >> 2); ++w) {
        uint32_t x = S.in_w[w] ^ splat;
        if (w == (lo >> 2)) x &= 0xffffffffu << (8 * (lo & 3));
        if (w == (hi >> 2)) x &= 0xffffffffu >> (8 * (3 - (hi & 3)));
        diff |= x;
    }
    return (diff ? 0u : 1u) | (c << 1) | (nq << 9);
}

// Output slot of block b whose text starts at chunk-relative offset g0.  A member codes up to in_cap bytes of text
// into at most that many + 31 bytes, so with 2^slot_shift <= in_cap / 32 the text offset + offset >> slot_shift +
// 64 bytes per block (rounded to 16) leaves every block its worst case, however many members it needs.
__device__ __forceinline__ uint64_t bz_slot_off(uint64_t g0, uint32_t b, uint32_t shift) {
    return (g0 + (g0 >> shift) + 64ull * b + 15) & ~15ull;
}

// gzip header + trailer bytes of a member straight to (possibly unaligned) memory
__device__ __forceinline__ void bz_frame(uint8_t *o, uint32_t deflate_bytes, uint32_t crc, uint32_t n) {
    const uint32_t total = BZ_GZ_HDR + deflate_bytes + 8;
    const uint8_t gz[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
    for (int i = 0; i < 16; ++i) o[i] = gz[i];
    o[16] = (uint8_t)(total - 1), o[17] = (uint8_t)((total - 1) >> 8);
    uint8_t *t = o + BZ_GZ_HDR + deflate_bytes;
    t[0] = (uint8_t)crc, t[1] = (uint8_t)(crc >> 8), t[2] = (uint8_t)(crc >> 16), t[3] = (uint8_t)(crc >> 24);
    t[4] = (uint8_t)n, t[5] = (uint8_t)(n >> 8), t[6] = (uint8_t)(n >> 16), t[7] = (uint8_t)(n >> 24);
}

// One BGZF member from text[0, n) (n <= in_cap) into dst, by one warp; returns its size.
// `may_code` false forces a stored block.
__device__ __forceinline__ uint32_t bz_member(const BgzfArgs &a, BzSmem S, const uint8_t *text, uint32_t n, bool may_code,
                                              uint8_t *dst) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t FULL = 0xffffffffu;
    const uint32_t phase = (uint32_t)((uintptr_t)text & 15);
    S.phase = phase;
    S.in = (uint8_t *)S.in_w + phase;
    const uint32_t out_words = a.out_cap / 4;
    for (uint32_t i = lane; i < out_words / 4; i += 32) ((uint4 *)S.out)[i] = make_uint4(0, 0, 0, 0);
    {
        const uint4 *src = (const uint4 *)(text - phase);
        const uint32_t nv = (phase + n + 15) / 16;
        for (uint32_t i = lane; i < nv; i += 32) ((uint4 *)S.in_w)[i] = src[i];
    }
    __syncwarp();
    const uint8_t *in = S.in;

    // pass 1, word-wise: CRC of the lane's chunk, its newline count and the newline positions (a queue of up to
    // eight 16-bit offsets in registers).  The message is taken from the word holding its first byte — the `pad` bytes
    // before it are zeroed, which leaves a zero CRC state unchanged — and chunks are aligned to the END of the whole
    // words, so the per-lane shift constants do not depend on n.  The 0xFFFFFFFF preset is folded into the first four
    // message bytes (complemented in place for the duration of the pass).
    const uint32_t pad = phase & 3, nv = n + pad;
    uint32_t *W = S.in_w + (phase >> 2);
    if (lane == 0) {
        for (uint32_t k = 0; k < pad; ++k) ((uint8_t *)W)[k] = 0;
        if (n >= 4)
            for (uint32_t k = 0; k < 4; ++k) S.in[k] ^= 0xffu;
    }
    __syncwarp();
    const int wf = (int)(nv >> 2);  // whole words
    const int CW = (int)a.crc_words;
    const int w_lo = wf - (int)(32 - lane) * CW, w_hi = w_lo + CW;
    uint32_t crc = 0, nl = 0;
    unsigned long long q_lo = 0, q_hi = 0;
    auto push = [&](uint32_t pos) {
        q_hi = (q_hi << 16) | (q_lo >> 48);
        q_lo = (q_lo << 16) | pos;
        ++nl;
    };
#pragma unroll 2
    for (int w = max(w_lo, 0); w < w_hi; ++w) {
        const uint32_t x = W[w];
        uint32_t m = __vcmpeq4(x, 0x0a0a0a0au);
        while (m) {  // about one word in twenty
            const uint32_t b = (__ffs(m) - 1) >> 3;
            push((uint32_t)w * 4 + b + 1 - pad);
            m &= ~(0xffu << (8 * b));
        }
        crc = bz_crc_word(S.crc_t, crc ^ x);
    }
    const uint32_t tail0 = (uint32_t)wf * 4;  // trailing bytes (virtual offsets) belong to the last lane
    if (lane == 31)
        for (uint32_t v = tail0; v < nv; ++v)
            if (in[v - pad] == 10) push(v + 1 - pad);
    crc = bz_gf_mul(crc, a.tab[BZ_T_SHIFT + lane]);  // moved to the end of the whole words ...
    crc = __reduce_xor_sync(FULL, crc);              // ... and folded
    if (n >= 4) {
        for (uint32_t v = tail0; v < nv; ++v) crc = S.crc_t[(crc ^ in[v - pad]) & 0xffu] ^ (crc >> 8);
    } else {  // too short for the folded preset (last piece of an oversized block): plain bytewise CRC
        crc = 0xffffffffu;
        for (uint32_t i = 0; i < n; ++i) crc = S.crc_t[(crc ^ in[i]) & 0xffu] ^ (crc >> 8);
    }
    crc = ~crc;
    __syncwarp();
    if (lane == 0 && n >= 4)
        for (uint32_t k = 0; k < 4; ++k) S.in[k] ^= 0xffu;

    // newline ranks -> line starts
    uint32_t inc = nl;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(FULL, inc, o);
        if ((int)lane >= o) inc += y;
    }
    const uint32_t n_lines = __shfl_sync(FULL, inc, 31);
    const uint32_t n_rec = n_lines >> 2;
    bool ok = may_code && n >= 8 && (n_lines & 3) == 0 && n_rec <= a.max_records && in[n - 1] == 10;
    bool staged = true;
    uint32_t total_bits = 0;
    if (ok) {
        uint32_t before = inc - nl;
        if (lane == 0) S.ls[0] = 0;
        if (!__any_sync(FULL, nl > 8)) {
            for (uint32_t i = 0; i < nl; ++i) {  // oldest first
                const uint32_t slot = nl - 1 - i;
                const unsigned long long q = slot < 4 ? q_lo : q_hi;
                S.ls[++before] = (uint32_t)(q >> (16 * (slot & 3))) & 0xffffu;
            }
        } else {  // a lane met more newlines than its queue holds (very short lines): scan again
            for (int w = max(w_lo, 0); w < w_hi; ++w) {
                uint32_t m = __vcmpeq4(W[w], 0x0a0a0a0au);
                while (m) {
                    const uint32_t b = (__ffs(m) - 1) >> 3;
                    S.ls[++before] = (uint32_t)w * 4 + b + 1 - pad;
                    m &= ~(0xffu << (8 * b));
                }
            }
            if (lane == 31)
                for (uint32_t v = tail0; v < nv; ++v)
                    if (in[v - pad] == 10) S.ls[++before] = v + 1 - pad;
        }
        __syncwarp();

        // pass 2a: quality class of every record
        for (uint32_t r = lane; r < n_rec; r += 32) {
            const uint32_t q0 = S.ls[4 * r + 3];
            S.rec[r] = bz_qual_class(S, q0, S.ls[4 * r + 4] - 1 - q0);
        }
        __syncwarp();
        // pass 2b: token choice — everything that looks at the previous record
        bool bad = false;
        auto qc_of = [&](uint32_t r) { return S.rec[r]; };
        for (uint32_t r0 = 0; r0 < n_rec; r0 += 32) {
            const uint32_t r = r0 + lane;
            uint32_t choice = 0;
            if (r < n_rec) {
                const uint32_t h0 = S.ls[4 * r], s0 = S.ls[4 * r + 1], p0 = S.ls[4 * r + 2], q0 = S.ls[4 * r + 3], e0 = S.ls[4 * r + 4];
                const uint32_t qc = S.rec[r];
                uint32_t P = 0, Sx = 0, same = 0;
                if (r) {
                    const uint32_t hp0 = S.ls[4 * r - 4], sp0 = S.ls[4 * r - 3], pp0 = S.ls[4 * r - 2], qp0 = S.ls[4 * r - 1];
                    const uint32_t mh = min(s0 - h0, sp0 - hp0);
                    const uint32_t mp = min(mh, BZ_MAX_MATCH);
                    while (P < mp && in[h0 + P] == in[hp0 + P]) ++P;
                    if (P < BZ_MIN_MATCH) P = 0;
                    const uint32_t ms = min(mh - P, BZ_MAX_MATCH);
                    while (Sx < ms && in[s0 - 1 - Sx] == in[sp0 - 1 - Sx]) ++Sx;
                    if (Sx < BZ_MIN_MATCH) Sx = 0;
                    if (h0 - pp0 == e0 - p0 && q0 - p0 == qp0 - pp0 && qc == S.rec[r - 1]) {
                        same = 1;
                        for (uint32_t i = 0; i < q0 - p0; ++i) same &= in[p0 + i] == in[pp0 + i];
                    }
                }
                // the bits are staged in the quality line, which is then no longer text: it must be a run
                if (!(qc & 1) || (qc >> 9) < BZ_MIN_MATCH) bad = true;
                choice = P | (Sx << 9) | (same << 18);
            }
            if (r < n_rec) S.rec2[r] = choice | (qc_of(r) & 0x1feu) << 18;  // the run character in bits 19..26
        }
        __syncwarp();
        // pass 2c: tokens of the lane's records.  Staged mode packs them into the words of the record's own quality
        // line (no sizing pass); when a quality line looks too short for that (short reads), direct mode sizes the
        // records first and writes them at their final positions after the scan.
        for (uint32_t r = lane; r < n_rec; r += 32) {
            const uint32_t choice = S.rec2[r];
            const uint32_t h0 = S.ls[4 * r], s0 = S.ls[4 * r + 1], p0 = S.ls[4 * r + 2], q0 = S.ls[4 * r + 3], e0 = S.ls[4 * r + 4];
            const uint32_t lits = (s0 - h0) - (choice & 0x1ffu) - ((choice >> 9) & 0x1ffu);
            const uint32_t est = 8u * lits + 3u * (p0 - s0) + 128u;   // digits cost <= 8 bits, bases <= 3, three matches
            const int cap = 32 * ((int)((phase + e0 - 1) >> 2) - (int)((phase + q0 + 3) >> 2));
            if ((int)est > cap) staged = false;
        }
        staged = !__any_sync(FULL, !staged);
        for (uint32_t r = lane; r < n_rec && !bad; r += 32) {
            const uint32_t choice = S.rec2[r];
            if (staged) {
                BzWriter bw;
                bw.init(S.in_w + ((phase + S.ls[4 * r + 3] + 3) >> 2), S.in_w + ((phase + S.ls[4 * r + 4] - 1) >> 2));
                bz_record_tokens(S, r, choice, bw);
                S.rec[r] = bw.finish();
                bad |= bw.bad;
            } else {
                BzCounter cw;
                cw.bits = 0, cw.bad = false;
                bz_record_tokens(S, r, choice, cw);
                S.rec[r] = cw.bits;
                bad |= cw.bad;
            }
        }
        ok = !__any_sync(FULL, bad);
    }
    const uint32_t base_bit = BZ_GZ_HDR * 8 + a.hdr_bits;
    const uint32_t eob = S.ll[256];
    uint32_t deflate_bytes = n + 5;
    if (ok) {
        // pass 3: scan of the bit counts, records shifted to their final positions
        uint32_t carry = base_bit;
        for (uint32_t r0 = 0; r0 < n_rec; r0 += 32) {
            const uint32_t r = r0 + lane;
            const uint32_t v = r < n_rec ? S.rec[r] : 0;
            uint32_t x = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(FULL, x, o);
                if ((int)lane >= o) x += y;
            }
            if (r < n_rec) {
                if (staged) S.rec2[r] = carry + x - v;  // staged: rec = bits, rec2 = position
                else S.rec[r] = carry + x - v;          // direct: rec = position, rec2 keeps the token choice
            }
            carry += __shfl_sync(FULL, x, 31);
        }
        total_bits = carry;
        const uint32_t end_bit = total_bits + (eob >> 16);
        const uint32_t coded_bytes = (end_bit + 7) / 8 - BZ_GZ_HDR;
        ok = coded_bytes < n + 5 && BZ_GZ_HDR + coded_bytes + 8 + 16 <= a.out_cap;
        if (ok) {
            deflate_bytes = coded_bytes;
            const uint32_t *hdr = a.tab + BZ_T_HDR;
            for (uint32_t i = lane; i * 32 < a.hdr_bits; i += 32)
                bz_or_bits(S.out, BZ_GZ_HDR * 8 + i * 32, hdr[i], min(32u, a.hdr_bits - i * 32));
            for (uint32_t r = lane; r < n_rec && !staged; r += 32) {
                BzDirect dw;
                dw.init(S.out, S.rec[r]);
                bz_record_tokens(S, r, S.rec2[r], dw);
                dw.finish();
            }
            for (uint32_t r = lane; r < n_rec && staged; r += 32) {
                const uint32_t *src = S.in_w + ((phase + S.ls[4 * r + 3] + 3) >> 2);
                const uint32_t d0 = S.rec2[r], nb = S.rec[r];
                for (uint32_t b = 0; b < nb; b += 32) {
                    uint32_t v = src[b >> 5];
                    const uint32_t take = min(32u, nb - b);
                    if (take < 32) v &= (1u << take) - 1;
                    bz_or_bits(S.out, d0 + b, v, take);
                }
            }
            if (lane == 0) bz_or_bits(S.out, total_bits, eob & 0xffffu, eob >> 16);
        }
    }
    const uint32_t total = BZ_GZ_HDR + deflate_bytes + 8;
    __syncwarp();
    if (ok) {
        if (lane == 0) bz_frame((uint8_t *)S.out, deflate_bytes, crc, n);
        __syncwarp();
        const uint4 *src = (const uint4 *)S.out;  // dst is 16-byte aligned: coded blocks are single members
        uint4 *d4 = (uint4 *)dst;
        for (uint32_t i = lane; i < (total + 15) / 16; i += 32) d4[i] = src[i];
    } else {  // stored block, text taken from global memory again (the staging copy may hold coded bits by now)
        if (lane == 0) {
            bz_frame(dst, deflate_bytes, crc, n);
            uint8_t *o = dst + BZ_GZ_HDR;
            o[0] = 1;
            o[1] = (uint8_t)n, o[2] = (uint8_t)(n >> 8);
            o[3] = (uint8_t)~n, o[4] = (uint8_t)(~n >> 8);
        }
        for (uint32_t i = lane; i < n; i += 32) dst[BZ_GZ_HDR + 5 + i] = text[i];
    }
    __syncwarp();
    return total;
}

__global__ void __launch_bounds__(BZ_THREADS) k_bgzf(BgzfArgs a, uint32_t n_blocks) {
    extern __shared__ __align__(16) uint8_t bz_smem[];
    const uint32_t tid = threadIdx.x, warp = tid >> 5;
    // carve shared memory: tables for the CTA, then one region per warp
    BzSmem S;
    uint32_t *tab = (uint32_t *)bz_smem;
    S.ll = tab + BZ_T_LL, S.dd = tab + BZ_T_DD, S.crc_t = tab + BZ_T_CRC;
    for (uint32_t i = tid; i < BZ_T_SHIFT; i += BZ_THREADS) tab[i] = a.tab[i];
    uint32_t off = (BZ_T_SHIFT * 4 + 15) & ~15u;
    off += warp * a.warp_smem;
    S.in_w = (uint32_t *)(bz_smem + off), off += a.in_cap + 32;
    S.out = (uint32_t *)(bz_smem + off), off += a.out_cap;
    S.ls = (uint32_t *)(bz_smem + off), off += (4 * a.max_records + 4) * 4;
    S.rec = (uint32_t *)(bz_smem + off), off += (a.max_records + 4) * 4;
    S.rec2 = (uint32_t *)(bz_smem + off);
    S.in = nullptr;
    S.phase = 0;
    __syncthreads();
    const uint32_t b = blockIdx.x * BZ_WARPS + warp;
    if (b >= n_blocks) return;

    const uint32_t tb = a.tile0 + b * a.tiles_per_block;
    const uint32_t te = min(tb + a.tiles_per_block, a.tile0 + a.n_tiles);
    const uint64_t g0 = a.tile_off[tb] - a.chunk_off, g1 = a.tile_off[te] - a.chunk_off;
    const uint64_t n = g1 - g0;
    uint8_t *slot = a.slots + bz_slot_off(g0, b, a.slot_shift);
    // Normally the block is one member.  Records longer than the tile bound (the reference's literal circular wrap,
    // Breakpoints.jl:111-121) can make a tile arbitrarily large: it is then cut into stored members of in_cap bytes.
    uint32_t total = 0;
    if (n <= a.in_cap) {
        total = bz_member(a, S, a.text + g0, (uint32_t)n, true, slot);
    } else {
        for (uint64_t p = 0; p < n; p += a.in_cap) {
            const uint32_t np = (uint32_t)min((uint64_t)a.in_cap, n - p);
            total += bz_member(a, S, a.text + g0 + p, np, false, slot + total);
        }
    }
    if ((tid & 31) == 0) a.sizes[b] = total;
}

// Packs the coded blocks back to back: block b goes to dst[offs[b] ...).  One warp per block.
// *total (pinned host memory, written over the bus: a copy-engine transfer would queue behind the bulk copy of the
// previous chunk) receives the size of the packed chunk.
__global__ void __launch_bounds__(256) k_bgzf_compact(BgzfArgs a, uint32_t n_blocks, const uint64_t *__restrict__ offs,
                                                      uint8_t *__restrict__ dst, unsigned long long *total) {
    const uint32_t b = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *total = offs[n_blocks];
        __threadfence_system();
    }
    if (b >= n_blocks) return;
    const uint32_t tb = a.tile0 + b * a.tiles_per_block;
    const uint8_t *src = a.slots + bz_slot_off(a.tile_off[tb] - a.chunk_off, b, a.slot_shift);
    const uint32_t n = a.sizes[b];
    uint8_t *d = dst + offs[b];
    // head bytes up to 4-byte alignment of the destination, then realigned words
    const uint32_t head = min(n, (uint32_t)((4 - ((uintptr_t)d & 3)) & 3));
    if (lane < head) d[lane] = src[lane];
    const uint32_t nw = (n - head) / 4;
    const uint32_t sh = head * 8;
    const uint32_t *sw = (const uint32_t *)src;  // word i of the destination = bytes head + 4 i .. of src
    uint32_t *dw = (uint32_t *)(d + head);
    for (uint32_t i = lane; i < nw; i += 32) {
        const uint32_t lo = sw[i], hi = sw[i + 1];
        dw[i] = sh ? __funnelshift_r(lo, hi, sh) : lo;
    }
    const uint32_t done = head + nw * 4;
    if (lane < n - done) d[done + lane] = src[done + lane];
}
