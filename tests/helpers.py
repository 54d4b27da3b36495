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
