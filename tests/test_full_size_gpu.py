"""BASELINE.json configurations at their FULL sizes, checked through size-independent properties
(the oracle would take minutes to hours here): every record of sampled chunks parses, read sequences are the
reverse complement / forward copy of the haplotype at the positions the header states, inserts of a molecule
are ordered and inside the molecule, barcodes count up, coverage targets are met with bounded overshoot,
byte counts add up, and shards chain."""
import os
import re
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu

COMP = bytes.maketrans(b"ACGT", b"TGCA")
HDR = {
    "stlfr": re.compile(rb"^@(\S+?):(\d+)\|(\d+):(\d+)\|(\d+):(\d+)#(\d+_\d+_\d+) ([12]):N:0:ATGACA$"),
    "haplotagging": re.compile(rb"^@(\S+?):(\d+)\|(\d+):(\d+)\|(\d+):(\d+)/([12])\tBX:Z:(A\d\dC\d\dB\d\dD\d\d)$"),
    "10x": re.compile(rb"^@(\S+?):(\d+)\|(\d+):(\d+)\|(\d+):(\d+)\|([ATGC]{16}) ([12]):N:0:ATGACA$"),
}


def check_chunk(fmt, r1, r2, haps, r1_len, r2_len, circular_n=None):
    """haps: {(name, hap): bytes}. Returns (#pairs, list of (mol_key, bstart, bend, barcode))."""
    l1, l2 = r1.split(b"\n"), r2.split(b"\n")
    assert l1[-1] == b"" and l2[-1] == b"" and len(l1) == len(l2) and (len(l1) - 1) % 4 == 0
    n = (len(l1) - 1) // 4
    out = []
    for i in range(n):
        h1, s1, p1, q1 = l1[4 * i: 4 * i + 4]
        h2, s2, p2, q2 = l2[4 * i: 4 * i + 4]
        m1, m2 = HDR[fmt].match(h1), HDR[fmt].match(h2)
        assert m1 and m2, (h1, h2)
        g1, g2 = m1.groups(), m2.groups()
        if fmt == "haplotagging":
            name, hap, ms, me, bs, be, rd1, bc = g1
            assert g2[:6] == g1[:6] and g2[7] == bc and rd1 == b"1" and g2[6] == b"2"
        else:
            name, hap, ms, me, bs, be, bc, rd1 = g1
            assert g2[:7] == g1[:7] and rd1 == b"1" and g2[7] == b"2"
        ms, me, bs, be = int(ms), int(me), int(bs), int(be)
        assert p1 == b"+" and p2 == b"+" and ms <= bs <= be <= me and be - bs + 1 >= 200
        seq = haps[(name.decode(), int(hap))]
        nn = len(seq)
        inl = bc if fmt == "10x" else b""
        assert len(q1) == len(s1) == r1_len + len(inl) and len(q2) == len(s2) == r2_len and set(q1) == {73} and set(q2) == {73}
        if circular_n is None:
            w1, w2 = seq[bs - 1: bs - 1 + r1_len], seq[be - r2_len: be]
        else:  # strict (modular) wrap
            idx1 = (np.arange(bs - 1, bs - 1 + r1_len) % nn)
            idx2 = (np.arange(be - r2_len, be) % nn)
            arr = np.frombuffer(seq, dtype=np.uint8)
            w1, w2 = arr[idx1].tobytes(), arr[idx2].tobytes()
        assert s1 == inl + w1.translate(COMP)[::-1], (h1,)
        assert s2 == w2, (h2,)
        out.append(((name, hap, ms, me), bs, be, bc))
    return n, out


def check_bgzf_chunk(coded: bytes, text: bytes):
    import zlib
    from helpers import bgzf_members
    members = list(bgzf_members(coded))
    assert sum(m[2] for m in members) == len(text)
    pos = 0
    for i, (off, size, isize, crc) in enumerate(members):
        if i < 300 or i >= len(members) - 300 or i % 997 == 0:
            raw = zlib.decompress(coded[off + 18:off + size - 8], wbits=-15)
            assert len(raw) == isize and zlib.crc32(raw) == crc and raw == text[pos:pos + isize], f"member {i}"
        pos += isize


def run_config(wl, max_chunk_mib, sample_chunks):
    import bench
    import mimick_b200 as M
    desc, _, nhap, fmt, kw = bench.WORKLOADS[wl]
    simkw = dict(bench.COMMON_SIM)
    simkw.update(kw)
    genome = bench.make_genome(wl, 1.0, "cuda:0", pin=False)
    haps = {(nm, h): t.numpy().tobytes() for nm, h, t in genome}
    with M.Simulator(fmt, seed=42, device=0, max_chunk_bytes=max_chunk_mib << 20, **simkw) as sim:
        for nm, hap, h in genome:
            sim.add_contig(nm, (h.data_ptr(), h.numel()), hap)
        plan = sim.plan()
        place = sim.place()
        schemas = sim.schemas()
        slot_bytes, n_chunks = sim.chunk_layout()
        got_pairs, b1, b2, checked, k = 0, 0, 0, [], 0
        texts = {}
        while True:
            ch = sim.emit_chunk()
            got_pairs += ch.n_pairs
            assert ch.r1_offset == b1 and ch.r2_offset == b2 and ch.first_pair == got_pairs - ch.n_pairs
            b1 += ch.r1_bytes
            b2 += ch.r2_bytes
            if k in sample_chunks or ch.done:
                r1, r2 = sim.copy_chunk(ch)
                # chunks hold whole records: check a window of records from both ends of the chunk
                r1, r2 = r1.tobytes(), r2.tobytes()
                assert r1.startswith(b"@") and r1.endswith(b"\n")
                head1 = r1[: 1 << 20]
                head1 = head1[: head1.rfind(b"\n@") + 1]
                nrec = head1.count(b"\n") // 4
                cut2 = 0
                for _ in range(nrec * 4):
                    cut2 = r2.index(b"\n", cut2) + 1
                circ = simkw.get("circular")
                n, recs = check_chunk(fmt, head1, r2[:cut2], haps, 150, 150, circular_n=True if circ else None)
                checked.append((n, recs))
                texts[k] = (r1, r2)
            k += 1
            if ch.done:
                break
        assert k == n_chunks and got_pairs == plan.n_pairs == place.n_pairs and b1 == place.r1_bytes and b2 == place.r2_bytes
        # the same shard through the BGZF stage: coded offsets chain, every member of the sampled chunks has a valid
        # frame whose ISIZEs add up to the text chunk, and a sample of members inflates to exactly that text
        sim.rewind()
        z1 = z2 = k = 0
        while True:
            ch = sim.emit_chunk_bgzf()
            assert ch.r1_offset == z1 and ch.r2_offset == z2
            z1 += ch.r1_bytes
            z2 += ch.r2_bytes
            if k in texts:
                for coded, text in zip(sim.copy_chunk(ch), texts[k]):
                    check_bgzf_chunk(coded.tobytes(), text)
            k += 1
            if ch.done:
                break
        assert k == n_chunks and 4.5 < (b1 + b2) / (z1 + z2) < 7.0
        mols = sim.molecules(0, min(plan.n_molecules, 200000))
    return plan, place, schemas, checked, mols


def common_checks(plan, schemas, mols, coverage, checked):
    for s in schemas:  # reads_required (ProcessFASTA.jl:61) reached, overshoot bounded by one barcode group
        assert s["reads_required"] == int((coverage * s["length"] / 150) / 2)
        assert s["reads_required"] <= s["reads_current"] < s["reads_required"] + 5 * 2000
    assert np.array_equal(mols["barcode_index"], np.arange(len(mols)))       # one barcode per molecule, counting up
    assert np.all(np.diff(mols["group"]) >= 0) and np.all(mols["n_frags"] >= 1)
    assert np.array_equal(mols["first_read"][1:], np.cumsum(mols["n_frags"])[:-1])
    for n, recs in checked:
        assert n > 100
        by_mol = {}
        for key, bs, be, bc in recs:
            by_mol.setdefault((key, bc), []).append((bs, be))
        for v in by_mol.values():   # inserts of a molecule: ordered, non-overlapping
            assert all(v[i][1] < v[i + 1][0] for i in range(len(v) - 1))


def test_config2_full_size_circular_stlfr():
    """BASELINE configs[1]: 5 Mbp circular, stLFR, 1 haplotype, 50x (strict wrap) — 833 333 pairs."""
    plan, place, schemas, checked, mols = run_config("c2", 64, {0, 3, 7})
    assert schemas[0]["reads_required"] == 833333 and plan.n_pairs >= 833333
    common_checks(plan, schemas, mols, 50, checked)
    assert np.all(mols["molsize"] >= 4000)


def test_config3_full_size_diploid_10x():
    """BASELINE configs[2]: 250 Mbp diploid, 10X, 30x, molecule length 50 kbp — 50 M pairs, ~40 GB of FASTQ."""
    plan, place, schemas, checked, mols = run_config("c3", 2048, {0, 5})
    assert [s["reads_required"] for s in schemas] == [25_000_000, 25_000_000]
    assert 50_000_000 <= plan.n_pairs < 50_010_000 and place.r1_bytes > place.r2_bytes  # barcode inline on R1
    common_checks(plan, schemas, mols, 30, checked)
    assert abs(mols["molsize"].mean() - 53000) < 1500 and abs(mols["n_frags"].mean() - 22.2) < 1.0  # SURVEY §8 Monte-Carlo figures


def test_config5_full_size_sharded():
    """BASELINE configs[4]: 3 Gbp diploid, haplotagging, 30x, 10 % singletons, as shard 3 of 8."""
    import bench
    import mimick_b200 as M
    desc, _, nhap, fmt, kw = bench.WORKLOADS["c5"]
    simkw = dict(bench.COMMON_SIM)
    simkw.update(kw)
    genome = bench.make_genome("c5", 1.0, "cuda:0", pin=False)
    haps = None
    with M.Simulator(fmt, seed=42, device=0, max_chunk_bytes=1024 << 20, **simkw) as sim:
        for nm, hap, h in genome:
            sim.add_contig(nm, (h.data_ptr(), h.numel()), hap)
        plan = sim.plan()
        assert 600_000_000 <= plan.n_pairs < 600_100_000 and plan.n_epochs == 48
        assert abs(plan.n_pairs / plan.n_molecules - 30.8) < 0.5   # 30.8 inserts per molecule with 10 % singletons
        bounds = []
        for rank in (2, 3):
            sim.set_shard(rank, 8)
            place = sim.place()
            bounds.append((place.first_pair, place.n_pairs, place.first_molecule, place.n_molecules))
            assert abs(place.n_pairs - plan.n_pairs / 8) < 10000
        assert bounds[0][0] + bounds[0][1] == bounds[1][0] and bounds[0][2] + bounds[0][3] == bounds[1][2]  # shards chain
        ch = sim.emit_chunk()
        r1, r2 = sim.copy_chunk(ch)
        r1 = r1.tobytes()[: 1 << 19]
        r1 = r1[: r1.rfind(b"\n@") + 1]
        nrec = r1.count(b"\n") // 4
        r2b = r2.tobytes()
        cut = 0
        for _ in range(nrec * 4):
            cut = r2b.index(b"\n", cut) + 1
        haps = {(nm, h): t.numpy().tobytes() for nm, h, t in genome}
        n, recs = check_chunk(fmt, r1, r2b[:cut], haps, 150, 150)
        assert n > 500
        mols = sim.molecules(place.first_molecule, 1000)
        assert mols["barcode_index"][0] == place.first_molecule and mols["first_read"][0] == place.first_pair
        sing = sim.molecules(0, 500000)["n_frags"] == 1
        assert abs(sing.mean() - 0.1) < 0.01


def test_config4_full_size_vcf_sample_tellseq():
    """BASELINE configs[3], one of its 8 samples (one sample per GPU): 100 Mbp master contig, 100 k SNPs and 10 k
    indels (1-20 bp, VCF-anchored), genotypes Bernoulli(0.3) per haplotype, TELLseq, 20x.  The two haplotypes built on
    the device equal a numpy construction (SNPs painted, indel pieces spliced), reads_required is inherited from the
    master (ProcessVariants.jl:157-185), and sampled records are windows of those haplotypes."""
    import mimick_b200 as M
    HDR["tellseq"] = re.compile(rb"^@(\S+?):(\d+)\|(\d+):(\d+)\|(\d+):(\d+):([ATGC]{18}) ([12]):N:0:ATGACA$")
    rng = np.random.default_rng(1004)
    n = 100_000_000
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    master = acgt[rng.integers(0, 4, size=n)]
    n_var = 110_000
    cells = np.sort(rng.choice(n // 64 - 2, size=n_var, replace=False)) + 1
    pos = cells.astype(np.int64) * 64 + rng.integers(0, 16, size=n_var)          # 1-based, >= 48 apart
    is_indel = np.zeros(n_var, dtype=bool)
    is_indel[rng.choice(n_var, size=10_000, replace=False)] = True
    ent_pos, ent_type, ent_reflen, ent_alt, ent_hap = [], [], [], [], []
    per_hap = {1: [], 2: []}
    snp_pos = {1: [], 2: []}
    snp_alt = {1: [], 2: []}
    for i in range(n_var):
        p = int(pos[i])
        ref = master[p - 1]
        if is_indel[i]:
            k = int(rng.integers(1, 21))
            if rng.random() < 0.5:   # deletion: REF = k + 1 bases, ALT = the anchor base
                ref_len, alt = k + 1, bytes([ref])
            else:                    # insertion: REF = anchor, ALT = anchor + k bases
                ref_len, alt = 1, bytes([ref]) + acgt[rng.integers(0, 4, size=k)].tobytes()
            typ = 1
        else:
            ref_len, alt, typ = 1, bytes([acgt[(np.searchsorted(acgt, ref) + int(rng.integers(1, 4))) % 4]]), 0
        for hap in (1, 2):
            if rng.random() < 0.3:
                ent_pos.append(p), ent_type.append(typ), ent_reflen.append(ref_len), ent_alt.append(alt), ent_hap.append(hap)
                if typ:
                    per_hap[hap].append((p, ref_len, alt))
                else:
                    snp_pos[hap].append(p - 1), snp_alt[hap].append(alt[0])
    want = {}
    for hap in (1, 2):
        seq = master.copy()
        seq[np.array(snp_pos[hap], dtype=np.int64)] = np.array(snp_alt[hap], dtype=np.uint8)
        raw = seq.tobytes()
        out, prev = [], 0
        for p, ref_len, alt in per_hap[hap]:
            out.append(raw[prev:p - 1])
            out.append(alt)
            prev = p - 1 + ref_len
        out.append(raw[prev:])
        want[hap] = b"".join(out)
    simkw = dict(coverage=20, n_molecules=2, mol_coverage=0.2, mol_length=80000, insert_length=500, insert_stdev=50,
                 read_length=[150, 150], singletons=0.0, attempts=300)
    with M.Simulator("tellseq", seed=42, device=0, max_chunk_bytes=1024 << 20, **simkw) as sim:
        sim.add_contig("chr1", master.tobytes(), 1)
        sim.set_sample_variants(3, 2, [0], np.zeros(len(ent_pos), dtype=np.int32), ent_pos, ent_type, ent_reflen, ent_alt, ent_hap)
        schemas = sim.schemas()
        assert [s["reads_required"] for s in schemas] == [6_666_666, 6_666_666] and [s["haplotype"] for s in schemas] == [1, 2]
        for i, hap in enumerate((1, 2)):
            got = sim.schema_bases(i)
            assert len(got) == len(want[hap]) == schemas[i]["length"]
            assert got == want[hap], f"haplotype {hap} differs from the numpy construction"
        plan = sim.plan()
        place = sim.place()
        assert 13_333_332 <= plan.n_pairs < 13_343_332 and plan.n_epochs == 2
        ch = sim.emit_chunk()
        r1, r2 = sim.copy_chunk(ch)
        r1 = r1.tobytes()[: 1 << 20]
        r1 = r1[: r1.rfind(b"\n@") + 1]
        nrec = r1.count(b"\n") // 4
        r2b = r2.tobytes()
        cut = 0
        for _ in range(nrec * 4):
            cut = r2b.index(b"\n", cut) + 1
        nchk, recs = check_chunk("tellseq", r1, r2b[:cut], {("chr1", 1): want[1], ("chr1", 2): want[2]}, 150, 150)
        assert nchk > 1000 and {k[0][1] for k in recs} == {b"1", b"2"}
