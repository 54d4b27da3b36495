"""Shared test helpers: seeded synthetic inputs and the GPU-vs-oracle comparison."""
import gzip
import os

import numpy as np


def random_bases(rng, n, alphabet=b"ACGT"):
    return np.frombuffer(alphabet, dtype=np.uint8)[rng.integers(0, len(alphabet), size=n)].tobytes()


def write_fasta(path, records, width=80, compress=None):
    compress = path.endswith(".gz") if compress is None else compress
    op = gzip.open if compress else open
    with op(path, "wb") as fh:
        for name, seq in records:
            fh.write(b">" + name.encode() + b"\n")
            for i in range(0, len(seq), width):
                fh.write(seq[i:i + width] + b"\n")


def make_ci_haplotypes(tmpdir, seed=7, big=200000):
    """Seeded twin of the reference's test/create_test_data.py: two contigs of `big` random bases and a
    1000-base contig followed by NNNN, written as hap1.fa.gz / hap2.fa.gz."""
    rng = np.random.default_rng(seed)
    paths = []
    for i in (1, 2):
        recs = [("Contig1", random_bases(rng, big, b"ATCG")), ("Contig2", random_bases(rng, big, b"ATCG")),
                ("Contig3", random_bases(rng, 1000, b"ATCG") + b"NNNN")]
        p = os.path.join(str(tmpdir), f"hap{i}.fa.gz")
        write_fasta(p, recs, width=10**9)
        paths.append(p)
    return paths


def first_diff(a: bytes, b: bytes):
    n = min(len(a), len(b))
    if a[:n] == b[:n]:
        return None if len(a) == len(b) else n
    x = np.frombuffer(a[:n], dtype=np.uint8)
    y = np.frombuffer(b[:n], dtype=np.uint8)
    return int(np.flatnonzero(x != y)[0])


def assert_same_bytes(got: bytes, want: bytes, what: str):
    d = first_diff(got, want)
    if d is not None:
        lo = max(0, d - 200)
        raise AssertionError(
            f"{what}: first difference at byte {d} (len got {len(got)}, want {len(want)})\n"
            f"--- got ---\n{got[lo:d + 120]!r}\n--- want ---\n{want[lo:d + 120]!r}")


def bgzf_members(z: bytes):
    """Walks a BGZF stream; yields (offset, member size, ISIZE, CRC32) and checks every fixed header field."""
    import struct
    pos = 0
    while pos < len(z):
        assert z[pos:pos + 4] == b"\x1f\x8b\x08\x04", f"gzip magic / FEXTRA at {pos}"
        xlen = struct.unpack_from("<H", z, pos + 10)[0]
        assert xlen == 6 and z[pos + 12:pos + 16] == b"BC\x02\x00", f"BC subfield at {pos}"
        bsize = struct.unpack_from("<H", z, pos + 16)[0] + 1
        crc, isize = struct.unpack_from("<II", z, pos + bsize - 8)
        assert isize <= 65536
        yield pos, bsize, isize, crc
        pos += bsize
    assert pos == len(z)


def bgzf_inflate(z: bytes) -> bytes:
    """Member by member with zlib (checks CRC-32 and ISIZE of every member)."""
    import zlib
    out = []
    for off, size, isize, crc in bgzf_members(z):
        raw = zlib.decompress(z[off + 18:off + size - 8], wbits=-15)
        assert len(raw) == isize and zlib.crc32(raw) == crc, f"member at {off}: CRC/ISIZE mismatch"
        out.append(raw)
    return b"".join(out)


# ------------------------------------------------------------------------------------------
# full-size parity: the oracle in digest mode against the CUDA path streamed through the C ABI
# ------------------------------------------------------------------------------------------
def oracle_digest(fmt, schemas, seed, sample, window, max_pairs=0, **kw):
    """Runs the oracle in digest mode.  `schemas`: [(name, haplotype, numpy uint8 array of upper-case symbols)] in
    schema order, or 4-tuples with an explicit reads_required (VCF mode inherits the master's).  The arrays are
    borrowed, not copied."""
    import oracle as O
    keys = ("coverage", "n_molecules", "mol_coverage", "mol_length", "insert_length", "insert_stdev", "read_length",
            "singletons", "circular", "attempts", "strict_wrap", "error_rate", "indel_error")
    o = O.Oracle(fmt, **{k: kw[k] for k in keys if k in kw})
    o.set_philox(seed, sample)
    L = O.lib()
    keep = []
    for sc in schemas:
        name, hap, arr = sc[:3]
        arr = np.ascontiguousarray(arr, dtype=np.uint8)
        keep.append(arr)
        req = sc[3] if len(sc) > 3 else int(L.orc_reads_required(arr.ctypes.data, arr.size, o.coverage, o.read_length[0], o.read_length[1]))
        assert req >= 0
        o.add_schema_borrowed(name, hap, arr.ctypes.data, arr.size, req)
    o.set_digest(window)
    if max_pairs:
        o.set_max_pairs(max_pairs)
    res = o.run()
    o.close()
    return res


def stream_crc_at(sim, cps):
    """Streams the placed shard through mmk_emit_stream (host ring, chunk by chunk) and returns, for every oracle
    checkpoint the shard reaches, the running CRC-32 of R1 and R2 at the checkpoint's byte offsets, plus the shard's
    total bytes."""
    import ctypes as C
    import zlib
    slot_bytes, _ = sim.chunk_layout()
    slots = [np.empty(slot_bytes, dtype=np.uint8) for _ in range(4)]
    state = [dict(crc=0, pos=0, k=0, out=[]), dict(crc=0, pos=0, k=0, out=[])]
    marks = (cps["bytes1"].tolist(), cps["bytes2"].tolist())

    def feed(st, mk, addr, n):
        mv = memoryview((C.c_char * n).from_address(addr)).cast("B")
        o = 0
        while st["k"] < len(mk) and mk[st["k"]] <= st["pos"] + (n - o):
            cut = mk[st["k"]] - st["pos"]
            st["crc"] = zlib.crc32(mv[o:o + cut], st["crc"])
            st["pos"] += cut
            o += cut
            st["out"].append(st["crc"])
            st["k"] += 1
        if o < n:
            st["crc"] = zlib.crc32(mv[o:], st["crc"])
            st["pos"] += n - o

    def consume(ch, p1, p2):
        assert ch.r1_offset == state[0]["pos"] and ch.r2_offset == state[1]["pos"]
        feed(state[0], marks[0], p1, ch.r1_bytes)
        feed(state[1], marks[1], p2, ch.r2_bytes)

    sim.emit_stream([slots[0].ctypes.data, slots[1].ctypes.data], [slots[2].ctypes.data, slots[3].ctypes.data], slot_bytes, consume)
    return state[0]["out"], state[1]["out"], state[0]["pos"], state[1]["pos"]


def assert_digest_parity(sim, plan, place, orc, what, prefix=False):
    """GPU text (streamed) == oracle text at every checkpoint; plan and inserts element-wise.  `prefix`: the oracle
    stopped early (max_pairs), so it is compared as a prefix of the shard."""
    cps = orc.checkpoints
    n_or = int(cps["pairs"][-1])
    if not prefix:
        assert plan.n_pairs == n_or and place.n_pairs == n_or, (what, plan.n_pairs, n_or)
        assert place.r1_bytes == int(cps["bytes1"][-1]) and place.r2_bytes == int(cps["bytes2"][-1]), what
    else:
        assert place.first_pair == 0 and place.n_pairs >= n_or, (what, place.n_pairs, n_or)
    c1, c2, n1, n2 = stream_crc_at(sim, cps)
    assert len(c1) == len(cps) and len(c2) == len(cps), (what, len(c1), len(cps))
    bad = [k for k in range(len(cps)) if c1[k] != int(cps["crc1"][k]) or c2[k] != int(cps["crc2"][k])]
    assert not bad, f"{what}: text differs from the oracle in checkpoint window {bad[0]} (pairs <= {int(cps['pairs'][bad[0]])})"
    if not prefix:
        assert n1 == place.r1_bytes and n2 == place.r2_bytes
    om = orc.molecules
    gm = sim.molecules(0, len(om))
    for f in ("group", "slot", "schema", "start", "molsize", "n_frags", "first_read", "barcode_index"):
        assert np.array_equal(gm[f], om[f]), f"{what}: molecule field {f} differs"
    st, en = sim.inserts(0, n_or)
    assert np.array_equal(st, orc.insert_start) and np.array_equal(en, orc.insert_stop), f"{what}: insert breakpoints differ"
    return len(cps)
