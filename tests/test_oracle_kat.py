"""Pins the CPU oracle (oracle/mimick_oracle.c) to everything the reference offers that can be checked
without running Julia: the literal format examples of its docs, hand-derived known answers on its own
fixtures, Random123's Philox known-answer vectors, and the analytic laws of the
distributions it samples.  The reference itself ships no golden outputs ("parity unpinned")."""
import ctypes as C
import os
import re

import numpy as np
import pytest
from scipy import stats

import oracle as O
from helpers import make_ci_haplotypes, random_bases, write_fasta


def barcode(fmt, k):
    buf = C.create_string_buffer(64)
    n = O.lib().orc_barcode(O.FMT[fmt], k, buf)
    return None if n < 0 else buf.value.decode()


def test_philox_known_answers():
    L = O.lib()
    for ctr, key, want in [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
                           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
                           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
                            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]:
        c = np.array(ctr, dtype=np.uint32)
        k = np.array(key, dtype=np.uint32)
        o = np.zeros(4, dtype=np.uint32)
        L.orc_philox(c.ctypes.data, k.ctypes.data, o.ctypes.data)
        assert tuple(int(x) for x in o) == want


def test_deterministic_log_exp_accuracy():
    L = O.lib()
    rng = np.random.default_rng(0)
    x = np.exp(rng.uniform(-700, 700, 20000))
    got = np.array([L.orc_det_log(float(v)) for v in x])
    assert np.max(np.abs(got - np.log(x)) / np.maximum(np.abs(np.log(x)), 1e-300)) < 1e-15
    y = rng.uniform(-700, 700, 20000)
    got = np.array([L.orc_det_exp(float(v)) for v in y])
    assert np.max(np.abs(got - np.exp(y)) / np.exp(y)) < 1e-15


def test_barcode_enumerators_match_docs_and_iterator_order():
    # shapes from docs/algorithm.md:40-46; order = Iterators.product, first factor fastest (Barcodes.jl:12-34)
    assert barcode("haplotagging", 0) == "A01C01B01D01"
    assert barcode("haplotagging", 1) == "A02C01B01D01"
    assert barcode("haplotagging", 95) == "A96C01B01D01"
    assert barcode("haplotagging", 96) == "A01C02B01D01"
    assert barcode("haplotagging", 96 ** 2) == "A01C01B02D01"
    assert barcode("haplotagging", 96 ** 3) == "A01C01B01D02"
    assert barcode("haplotagging", 96 ** 4 - 1) == "A96C96B96D96"
    assert barcode("haplotagging", 96 ** 4) is None  # format_barcode(::Nothing) → error, Barcodes.jl:75
    assert re.fullmatch(r"A\d\dC\d\dB\d\dD\d\d", barcode("haplotagging", 4242))  # like "A43C21B01D93"
    assert barcode("stlfr", 0) == "1_1_1"
    assert barcode("stlfr", 1999) == "2000_1_1"
    assert barcode("stlfr", 2000) == "1_2_1"
    assert barcode("stlfr", 4_000_000) == "1_1_2"
    assert barcode("stlfr", 8_000_000_000) is None
    assert barcode("tellseq", 0) == "A" * 18
    assert barcode("tellseq", 1) == "T" + "A" * 17   # alphabet order A,T,G,C; leftmost fastest
    assert barcode("tellseq", 2) == "G" + "A" * 17
    assert barcode("tellseq", 3) == "C" + "A" * 17
    assert barcode("tellseq", 4) == "AT" + "A" * 16
    assert len(barcode("tenx", 12345)) == 16 and set(barcode("tenx", 12345)) <= set("ATGC")
    assert barcode("tenx", 4 ** 16) is None


def test_interperet_format_table():
    # Common.jl:56-74
    assert O.interperet_format("haplotagging") == ("haplotagging", "haplotagging")
    assert O.interperet_format("stlfr") == ("stlfr", "stlfr")
    assert O.interperet_format("tenx") == ("tenx", "tenx")
    assert O.interperet_format("10x") == ("tenx", "tenx")
    assert O.interperet_format("standard") == ("tellseq", "standard")
    assert O.interperet_format("standard:haplotagging") == ("haplotagging", "standard")
    assert O.interperet_format("standard:stlfr") == ("stlfr", "standard")
    with pytest.raises(O.OracleError):
        O.interperet_format("nanopore")


def test_reads_required_on_reference_fixture(ref_fasta):
    # -g 5 on the 1.5 Mbp contig → 25 000 pairs per haplotype
    (name, seq), = O.read_fasta(ref_fasta)
    assert name == "chr1" and len(seq) == 1_500_000 and seq.count(b"N") == 0
    o = O.Oracle("haplotagging", coverage=5)
    s, req = o.prepare_contig(seq, False, 0, 1)
    assert req == 25000
    o2 = O.Oracle("haplotagging", coverage=30, read_length=(150, 108))
    assert o2.prepare_contig(seq, False, 0, 1)[1] == int((30 * 1_500_000 / ((150 + 108) // 2)) / 2)
    with pytest.raises(O.OracleError) as e:  # fewer than round(0.95 * 300) unambiguous bases
        o.prepare_contig(b"ACGT" * 50 + b"N" * 200, False, 0, 1)
    assert e.value.code == 5


def test_snp_painting_of_reference_vcf(ref_fasta, ref_vcf):
    # per-sample-haplotype ALT counts and the first variants of tsk_0 / tsk_1
    vcf = O.read_vcf(ref_vcf)
    assert vcf.contigs == ["chr1"] and vcf.ploidy == 2 and len(vcf.records) == 13705
    want = [(2069, 2123), (1997, 2134), (2046, 2110), (2102, 2057), (2071, 2082), (2116, 2150), (2082, 2117)]
    master = O.read_fasta(ref_fasta)[0][1]
    for si, (n1, n2) in enumerate(want):
        v = O.sample_variants(vcf, si)
        assert (len(v["chr1_1"]), len(v["chr1_2"])) == (n1, n2)
    v1 = O.sample_variants(vcf, 1)["chr1_1"]
    assert [(p, r, a) for p, t, r, a in v1[:2]] == [(187, "G", "A"), (192, "G", "A")]
    v0 = O.sample_variants(vcf, 0)["chr1_1"]
    assert v0[0][0] == 192
    hap = O.Oracle.mutate(master, v1)
    assert len(hap) == len(master) and hap[186:187] == b"A" and hap[191:192] == b"A" and master[186:187] == b"G"
    diff = np.flatnonzero(np.frombuffer(hap, np.uint8) != np.frombuffer(master, np.uint8))
    assert len(diff) == 1997  # every ALT differs from REF in the fixture


def test_indel_splice_semantics():
    # mutate!: all SNPs first, then non-SNPs in reverse file order on the current sequence (ProcessVariants.jl:170-176)
    ref = b"AAAACCCCGGGGTTTT"
    var = [(2, 0, "A", "T"), (5, 1, "CCC", "C"), (9, 1, "G", "GAT"), (13, 1, "TT", "AC")]
    assert O.Oracle.mutate(ref, var) == b"ATAACCGATGGGACTT"
    # REF span at the very end and an insertion at position 1
    assert O.Oracle.mutate(b"ACGT", [(1, 1, "A", "TTA"), (4, 1, "T", "")]) == b"TTACG"


def test_record_text_known_answers(ref_fasta, tmp_path):
    """Hand-derived records (SURVEY §8c ii): every format's header for one molecule, R1 = revcomp of the
    insert's left r1 bases, R2 = the insert's right r2 bases (Breakpoints.jl:99-100, FormatFASTQ.jl)."""
    seq = O.read_fasta(ref_fasta)[0][1]
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    formats = {
        "haplotagging": (b"/1\tBX:Z:A01C01B01D01", b"/2\tBX:Z:A01C01B01D01"),
        "stlfr": (b"#1_1_1 1:N:0:ATGACA", b"#1_1_1 2:N:0:ATGACA"),
        "tellseq": (b":AAAAAAAAAAAAAAAAAA 1:N:0:ATGACA", b":AAAAAAAAAAAAAAAAAA 2:N:0:ATGACA"),
        "10x": (b"|AAAAAAAAAAAAAAAA 1:N:0:ATGACA", b"|AAAAAAAAAAAAAAAA 2:N:0:ATGACA"),
        "standard": (b"/1\tVX:i:1\tBX:Z:AAAAAAAAAAAAAAAAAA", b"/2\tVX:i:1\tBX:Z:AAAAAAAAAAAAAAAAAA"),
        "standard:stlfr": (b"/1\tVX:i:1\tBX:Z:1_1_1", b"/2\tVX:i:1\tBX:Z:1_1_1"),
    }
    for fmt, (t1, t2) in formats.items():
        res = O.mimick_single([ref_fasta], fmt, seed=3, coverage=0.05, attempts=300, singletons=0.0)
        m = res.molecules[0]
        bs, be = int(res.insert_start[0]), int(res.insert_stop[0])
        head = b"@chr1:1|%d:%d|%d:%d" % (m["start"], m["start"] + m["molsize"], bs, be)
        l1 = res.r1.split(b"\n")[:4]
        l2 = res.r2.split(b"\n")[:4]
        assert l1[0] == head + t1 and l2[0] == head + t2
        want_r1 = seq[bs - 1:bs - 1 + 150].translate(comp)[::-1]
        want_r2 = seq[be - 150:be]
        bc = b"AAAAAAAAAAAAAAAA" if fmt == "10x" else b""
        assert l1[1] == bc + want_r1 and l2[1] == want_r2
        assert l1[2] == b"+" and l1[3] == b"I" * (150 + len(bc)) and l2[3] == b"I" * 150


def test_structural_invariants(tmp_path):
    fastas = make_ci_haplotypes(tmp_path, seed=5, big=40000)
    res = O.mimick_single(fastas, "haplotagging", seed=9, coverage=10, attempts=300, singletons=0.2)
    m = res.molecules
    assert np.all(np.diff(m["barcode_index"]) == 1) and m["barcode_index"][0] == 0  # one barcode per molecule (Main.jl:51)
    for mol in m[:400]:
        a, b = mol["first_read"], mol["first_read"] + mol["n_frags"]
        st, en = res.insert_start[a:b], res.insert_stop[a:b]
        assert np.all(st[1:] > en[:-1])                       # ordered, non-overlapping
        assert st[0] >= mol["start"] and en[-1] <= mol["start"] + mol["molsize"]   # inside position = start:start+L
        assert np.all(en - st + 1 >= 200)                     # truncated(Normal, lower = 200)
    for cur, req in zip(res.reads_current, res.reads_required):
        assert cur >= req                                      # every schema reaches its target ...
    recs = res.r1.split(b"\n")
    assert len(recs) == 4 * len(res.insert_start) + 1 and recs[-1] == b""
    assert all(r.startswith(b"@") for r in recs[0:-1:4]) and all(r == b"+" for r in recs[2:-1:4])


def test_truncated_exponential_law():
    L = O.lib()
    n = 200000
    out = np.zeros(n)
    # rejection branch (tp > 0.25): molecule length, Exponential(80000) truncated below at 4000
    L.orc_sample_trunc_exp(1, 80000.0, 4000.0, 0, 0.0, n, out.ctypes.data)
    assert out.min() >= 4000
    assert stats.kstest(out - 4000, stats.expon(scale=80000).cdf).pvalue > 1e-3  # memoryless
    # both bounds: n_molecules Exponential(2) on [1, 5]
    L.orc_sample_trunc_exp(2, 2.0, 1.0, 1, 5.0, n, out.ctypes.data)
    d = stats.truncexpon(b=(5 - 1) / 2.0, loc=1, scale=2.0)
    assert out.min() >= 1 and out.max() <= 5 and stats.kstest(out, d.cdf).pvalue > 1e-3
    # quantile branch (tp <= 0.25): the tiny-contig case truncated(Exponential(2), 2.0, 2.008)
    info = np.zeros(4)
    L.orc_trunc_exp_info(2.0, 2.0, 1, 2.008, info.ctypes.data)
    assert info[0] == 0 and 1e-8 < info[3] < 0.25 and abs(info[3] - (np.exp(-1) - np.exp(-1.004))) < 1e-12
    L.orc_sample_trunc_exp(3, 2.0, 2.0, 1, 2.008, n, out.ctypes.data)
    d = stats.truncexpon(b=0.008 / 2.0, loc=2.0, scale=2.0)
    assert out.min() >= 2.0 and out.max() <= 2.008 and stats.kstest(out, d.cdf).pvalue > 1e-3
    # lower > upper is the reference's error site (Breakpoints.jl:191)
    L.orc_trunc_exp_info(2.0, 2.0, 1, 1.4, info.ctypes.data)
    assert info[0] == 7


def test_insert_size_law():
    L = O.lib()
    n = 200000
    out = np.zeros(n)
    L.orc_sample_insert(5, 500, 50, n, out.ctypes.data)
    assert out.min() >= 200 and stats.kstest(out, stats.norm(500, 50).cdf).pvalue > 1e-3
    L.orc_sample_insert(6, 300, 100, n, out.ctypes.data)   # truncation at 200 is 1 sigma below the mean
    d = stats.truncnorm(a=(200 - 300) / 100, b=np.inf, loc=300, scale=100)
    assert out.min() >= 200 and stats.kstest(out, d.cdf).pvalue > 1e-3
    # kept mass <= 0.3: Distributions.jl's randnt switches to Robert's exponential proposal (-i 150 -d 50: lb = 1, tp = 0.159)
    L.orc_sample_insert(7, 150, 50, n, out.ctypes.data)
    d = stats.truncnorm(a=1.0, b=np.inf, loc=150, scale=50)
    assert out.min() >= 200 and stats.kstest(out, d.cdf).pvalue > 1e-3
    L.orc_sample_insert(8, 100, 20, n, out.ctypes.data)     # lb = 5: tp = 2.9e-7
    d = stats.truncnorm(a=5.0, b=np.inf, loc=100, scale=20)
    assert out.min() >= 200 and stats.kstest(out, d.cdf).pvalue > 1e-3
    L.orc_sample_insert(9, 210, 200, n, out.ctypes.data)    # tp = 0.52: rejection with half the normals rejected
    d = stats.truncnorm(a=-0.05, b=np.inf, loc=210, scale=200)
    assert out.min() >= 200 and stats.kstest(out, d.cdf).pvalue > 1e-3


def test_seeds_share_no_insert_variates(tmp_path):
    """Nothing is mixed into the Philox key: runs whose seeds differ in any bit (0/1, bit 32, the old
    round/insert XOR pattern) consume disjoint insert-length and gap variates."""
    rng = np.random.default_rng(3)
    p = str(tmp_path / "g.fa")
    write_fasta(p, [("a", random_bases(rng, 400_000))])
    seen = {}
    for seed in (0, 1, 2, 3, 1 << 32, (1 << 32) | 5, 7 ^ 5):
        res = O.mimick_single([p], "haplotagging", seed=seed, coverage=4, attempts=300, singletons=0.0, record=True)
        v = res.variates
        sel = v[(v["tag"] == O.P_INSERT) | (v["tag"] == O.P_GAP)]
        assert len(sel) > 5000
        seen[seed] = set(sel["bits"][sel["tag"] == O.P_INSERT].tolist())
    seeds = list(seen)
    for i in range(len(seeds)):
        for j in range(i + 1, len(seeds)):
            assert not (seen[seeds[i]] & seen[seeds[j]]), (seeds[i], seeds[j])


def test_free_running_distributions(tmp_path):
    """Free-running (Philox) mode against the reference's laws: molecule length, inserts per molecule,
    singleton fraction, schema choice (north_star: KS / chi-square checks)."""
    rng = np.random.default_rng(1)
    p = str(tmp_path / "g.fa")
    write_fasta(p, [("a", random_bases(rng, 3_000_000)), ("b", random_bases(rng, 3_000_000))])
    res = O.mimick_single([p], "stlfr", seed=12, coverage=6, attempts=300, singletons=0.25)
    m = res.molecules
    # molecule length: Exponential(80000) | >= 4000, redrawn while > contig (3 Mbp: negligible), then trunc
    assert stats.kstest(m["molsize"] - 4000 + rng.uniform(0, 1, len(m)), stats.expon(scale=80000).cdf).pvalue > 1e-3
    single = m["n_frags"] == 1
    assert abs(single.mean() - 0.25) < 4 * np.sqrt(0.25 * 0.75 / len(m))
    # targets: uniform over the two live schemas until one completes
    first = m[: len(m) // 3]
    cnt = np.bincount(first["schema"], minlength=2)
    assert stats.chisquare(cnt).pvalue > 1e-3
    # insert length = trunc(Normal(500, 50))
    lens = (res.insert_stop - res.insert_start + 1).astype(float)
    assert abs(lens.mean() + 0.5 - 500) < 1.0 and abs(lens.std() - 50) < 1.0
    # gaps are uniform on the remaining leeway: the first gap of a molecule is U{0..leeway}
    nf = m["n_frags"].astype(np.int64)
    ok = ~single
    tot = np.add.reduceat(res.insert_stop - res.insert_start + 1, m["first_read"])
    leeway = m["molsize"] + 1 - tot
    u = (res.insert_start[m["first_read"]] - m["start"] + rng.uniform(0, 1, len(m))) / (leeway + 1)
    assert stats.kstest(u[ok], "uniform").pvalue > 1e-3


# ------------------------------------------------------------------------------------------
# sequencing-error model (not in the reference: reads there are verbatim copies with constant 'I')
# ------------------------------------------------------------------------------------------
def _reads(txt):
    lines = txt.split(b"\n")
    return lines[1::4], lines[3::4]


def test_error_model_off_is_the_reference_text(tmp_path):
    """All rates zero = the parity default: identical bytes to a run that never heard of the model."""
    rng = np.random.default_rng(5)
    p = os.path.join(str(tmp_path), "e.fa")
    write_fasta(p, [("chrE", random_bases(rng, 120_000))])
    kw = dict(coverage=8, attempts=300, singletons=0.2)
    a = O.mimick_single([p], "haplotagging", seed=3, **kw)
    b = O.mimick_single([p], "haplotagging", seed=3, error_rate=(0.0, 0.0), indel_error=(0.0, 0.0), **kw)
    assert a.r1 == b.r1 and a.r2 == b.r2
    assert set(_reads(a.r1)[1]) == {b"I" * 150}


def test_error_model_substitution_rate_and_quality_law(tmp_path):
    """Substitutions only: the per-cycle mismatch rate against the error-free run follows
    e(c) = e0 + (e1 - e0) c / (S - 1) (chi-square over 10 cycle bins, alpha = 1e-4), a substituted base never equals the
    template, and the quality of a cycle is 33 + clamp(round(-10 log10 e(c)) + jitter - 10 [substituted], 2, 41) with
    jitter in {-2, -1, 0, 0, 0, 0, 1, 2} equally likely (chi-square on the jitter histogram, alpha = 1e-4)."""
    from scipy import stats
    rng = np.random.default_rng(6)
    p = os.path.join(str(tmp_path), "e.fa")
    write_fasta(p, [("chrE", random_bases(rng, 400_000))])
    kw = dict(coverage=60, attempts=300, singletons=0.2)
    e0, e1, S = 0.002, 0.03, 150
    clean = O.mimick_single([p], "haplotagging", seed=11, **kw)
    noisy = O.mimick_single([p], "haplotagging", seed=11, error_rate=(e0, e1), **kw)
    for rd, (ct, nt) in enumerate(((clean.r1, noisy.r1), (clean.r2, noisy.r2)), start=1):
        cs, _ = _reads(ct)
        ns, nq = _reads(nt)
        assert len(cs) == len(ns) > 50_000
        A = np.frombuffer(b"".join(cs), dtype=np.uint8).reshape(-1, S)
        B = np.frombuffer(b"".join(ns), dtype=np.uint8).reshape(-1, S)
        Q = np.frombuffer(b"".join(nq), dtype=np.uint8).reshape(-1, S).astype(int) - 33
        mism = A != B
        n = A.shape[0]
        exp_rate = e0 + (e1 - e0) * np.arange(S) / (S - 1)
        # thresholds are floor(e * 2^24) / 2^24: indistinguishable from e at this sample size
        obs = mism.sum(axis=0).reshape(10, 15).sum(axis=1)
        exp = (exp_rate * n).reshape(10, 15).sum(axis=1)
        chi = ((obs - exp) ** 2 / exp).sum()
        assert chi < stats.chi2.ppf(1 - 1e-4, 10), (rd, chi, obs, exp)
        # quality law
        qb = np.clip(np.floor(-10 * np.log10(np.maximum(exp_rate, 1e-4)) + 0.5).astype(int), 2, 40)
        jit = Q - qb[None, :] + 10 * mism
        inner = (qb[None, :] - 12 >= 2) & (qb[None, :] + 2 <= 41)  # clamps cannot bite
        vals, counts = np.unique(jit[np.broadcast_to(inner, jit.shape)], return_counts=True)
        assert set(vals.tolist()) <= {-2, -1, 0, 1, 2}
        h = dict(zip(vals.tolist(), counts.tolist()))
        tot = sum(h.values())
        want = {-2: 1 / 8, -1: 1 / 8, 0: 4 / 8, 1: 1 / 8, 2: 1 / 8}
        chi = sum((h.get(k, 0) - tot * w) ** 2 / (tot * w) for k, w in want.items())
        assert chi < stats.chi2.ppf(1 - 1e-4, 4), (rd, chi, h)
        assert Q.min() >= 2 and Q.max() <= 41


def test_error_model_indels_shift_the_template(tmp_path):
    """Insertions and deletions only: read length stays S, a read without an event equals the clean read, the share of
    touched reads is 1 - (1 - pi - pd)^S within 4 sigma, and after the first event the read follows the template
    shifted by the net indel length."""
    rng = np.random.default_rng(7)
    p = os.path.join(str(tmp_path), "e.fa")
    write_fasta(p, [("chrE", random_bases(rng, 300_000))])
    kw = dict(coverage=40, attempts=300, singletons=0.2)
    pi, pd, S = 0.002, 0.003, 150
    clean = O.mimick_single([p], "stlfr", seed=12, **kw)
    noisy = O.mimick_single([p], "stlfr", seed=12, indel_error=(pi, pd), **kw)
    cs, _ = _reads(clean.r2)
    ns, nq = _reads(noisy.r2)
    assert all(len(x) == S for x in ns) and all(len(x) == S for x in nq)
    touched = sum(1 for a, b in zip(cs, ns) if a != b)
    n = len(cs)
    # a read can be touched without changing (an insertion of the base that follows, ...): the observed share is a bit lower
    p_any = 1 - (1 - pi - pd) ** S
    assert 0.6 * p_any * n < touched < p_any * n + 4 * (n * p_any * (1 - p_any)) ** 0.5
    # quality drops by 10 exactly where something was inserted (no substitutions here)
    ins_marks = sum(q.count(bytes([33 + 23 - 10 + j])) for q in nq[:2000] for j in (-2, -1, 0, 1, 2))
    assert ins_marks > 0


# ------------------------------------------------------------------------------------------
# A second, independent restatement of the text half (barcode enumerators, extraction, the five record formats),
# written in Python straight from the Julia sources, run over EVERY record of small oracle runs: the C oracle's
# plan (molecules, insert breakpoints) goes in, the whole R1 / R2 text must come out.
# ------------------------------------------------------------------------------------------
_COMP = bytes.maketrans(b"ACGTMRWSYKVHDBN", b"TGCAKYWSRMBDHVN")   # BioSequences complement of the 4-bit alphabet


def _julia_product(k, factors):
    """k-th element (0-based) of Base.Iterators.product(factors...): the FIRST factor varies fastest (Barcodes.jl:12-34)."""
    out = []
    for f in factors:
        out.append(f[k % len(f)])
        k //= len(f)
    assert k == 0, "barcodes exhausted"
    return out


def _py_barcode(style, k):
    if style == "haplotagging":     # product('A', 1:96, 'C', 1:96, 'B', 1:96, 'D', 1:96) -> A%02dC%02dB%02dD%02d (Barcodes.jl:53-56)
        a, c, b, d = _julia_product(k, [range(1, 97)] * 4)
        return b"A%02dC%02dB%02dD%02d" % (a, c, b, d)
    if style == "stlfr":            # product(1:2000, 1:2000, 1:2000) joined by "_" (Barcodes.jl:64)
        return b"_".join(b"%d" % x for x in _julia_product(k, [range(1, 2001)] * 3))
    n = 18 if style == "tellseq" else 16   # product(("ATGC" for i in 1:n)...) joined (Barcodes.jl:72-74)
    return bytes(_julia_product(k, [b"ATGC"] * n))


def _py_circular_index(seq, start, stop):
    """circular_index (Breakpoints.jl:111-121), literally: 1-based inclusive range; past the end it is
    seq[start:end] * seq[begin:(stop % n)] (an empty range when start > n, as Julia indexes it)."""
    n = len(seq)
    if stop > n:
        return seq[start - 1:] + seq[:stop % n]
    return seq[start - 1:stop]


def _py_records(fq_format, bc_style, schemas, res, r1_len, r2_len):
    """format_R1 / format_R2 (FormatFASTQ.jl:8-137) over extract_sequences! (Breakpoints.jl:95-109)."""
    tails = {
        "stlfr": (b"#%s 1:N:0:ATGACA", b"#%s 2:N:0:ATGACA"),
        "haplotagging": (b"/1\tBX:Z:%s", b"/2\tBX:Z:%s"),
        "tellseq": (b":%s 1:N:0:ATGACA", b":%s 2:N:0:ATGACA"),
        "tenx": (b"|%s 1:N:0:ATGACA", b"|%s 2:N:0:ATGACA"),
        "standard": (b"/1\tVX:i:1\tBX:Z:%s", b"/2\tVX:i:1\tBX:Z:%s"),
    }[fq_format]
    out1, out2 = [], []
    for m in res.molecules:
        name, hap, seq = schemas[int(m["schema"])]
        bc = _py_barcode(bc_style, int(m["barcode_index"]))
        start, size = int(m["start"]), int(m["molsize"])
        header = b"@%s:%d|%d:%d|" % (name.encode(), hap, start, start + size)
        a, b = int(m["first_read"]), int(m["first_read"]) + int(m["n_frags"])
        reads = []
        for bs, be in zip(res.insert_start[a:b].tolist(), res.insert_stop[a:b].tolist()):
            r1 = _py_circular_index(seq, bs, bs + r1_len - 1).translate(_COMP)[::-1]      # circular_index_R2: reverse complement
            r2 = _py_circular_index(seq, be - r2_len + 1, be)                               # circular_index_R1: forward
            reads.append((bs, be, r1, r2))
        inline = bc if fq_format == "tenx" else b""
        q1 = b"I" * (len(reads[0][2]) + len(inline))     # "I"^length(first read of the molecule) (+ barcode for 10x R1)
        q2 = b"I" * len(reads[0][3])
        for bs, be, r1, r2 in reads:
            out1.append(header + b"%d:%d" % (bs, be) + tails[0] % bc + b"\n" + inline + r1 + b"\n+\n" + q1 + b"\n")
            out2.append(header + b"%d:%d" % (bs, be) + tails[1] % bc + b"\n" + r2 + b"\n+\n" + q2 + b"\n")
    return b"".join(out1), b"".join(out2)


@pytest.mark.parametrize("fmt,fq,bc", [
    ("haplotagging", "haplotagging", "haplotagging"), ("stlfr", "stlfr", "stlfr"), ("tellseq", "tellseq", "tellseq"),
    ("10x", "tenx", "tenx"), ("standard", "standard", "tellseq"), ("standard:haplotagging", "standard", "haplotagging"),
    ("standard:stlfr", "standard", "stlfr"),
])
def test_every_record_rederived_in_python(tmp_path, fmt, fq, bc):
    rng = np.random.default_rng(101)
    p1 = os.path.join(str(tmp_path), "a.fa")
    p2 = os.path.join(str(tmp_path), "b.fa")
    recs1 = [("chrA", random_bases(rng, 60_000)), ("scaffold_2", random_bases(rng, 9_000, b"ACGTRYKMSWN"))]
    recs2 = [("chrA", random_bases(rng, 60_100)), ("scaffold_2", random_bases(rng, 9_050))]
    write_fasta(p1, recs1)
    write_fasta(p2, recs2)
    schemas = [(n, 1, s) for n, s in recs1] + [(n, 2, s) for n, s in recs2]
    res = O.mimick_single([p1, p2], fmt, seed=12, coverage=6, attempts=300, singletons=0.3, read_length=(150, 120),
                          mol_length=20000)
    assert len(res.molecules) > 40 and len(res.insert_start) > 400
    want1, want2 = _py_records(fq, bc, schemas, res, 150, 120)
    assert res.r1 == want1 and res.r2 == want2


@pytest.mark.parametrize("fmt,fq,bc", [("stlfr", "stlfr", "stlfr"), ("10x", "tenx", "tenx")])
def test_every_record_rederived_in_python_circular_literal_wrap(tmp_path, fmt, fq, bc):
    """circular=true with the reference's literal circular_index: molecules start anywhere on the contig, reads
    that pass the end become seq[start:end] * seq[1:stop % n] — longer or shorter than nominal — and the quality line keeps the
    length of the molecule's first read."""
    rng = np.random.default_rng(102)
    p = os.path.join(str(tmp_path), "c.fa")
    recs = [("plasmid", random_bases(rng, 30_000))]
    write_fasta(p, recs)
    res = O.mimick_single([p], fmt, seed=5, coverage=20, attempts=300, singletons=0.2, circular=True, mol_length=12000)
    want1, want2 = _py_records(fq, bc, [("plasmid", 1, recs[0][1])], res, 150, 150)
    assert res.r1 == want1 and res.r2 == want2
    lens = {len(x) for x in res.r1.split(b"\n")[1::4]}
    assert len(lens) > 1, "no read crossed the end of the contig: the wrap rule was not exercised"


# ------------------------------------------------------------------------------------------
# The plan half, the same way: the Julia loop (Main.jl:44-59) restated in Python with every rand() replaced by the next
# variate of the stream the oracle RECORDED (the replay contract of SURVEY §8a-bis).  The interpreter checks every tag, so
# a draw taken in another order, a missing redraw or a different truncation branch fails at the first variate.
# ------------------------------------------------------------------------------------------
class _Replay:
    TAGS = dict(NMOL=2, TARGET=3, MOLSIZE=4, SINGLETON=5, NFRAGS=6, INSERT=7, START=8, GAP=9)

    def __init__(self, variates):
        self.v, self.k = variates, 0

    def _pop(self, tag):
        e = self.v[self.k]
        assert int(e["tag"]) == self.TAGS[tag], f"variate {self.k}: stream has tag {int(e['tag'])}, the Julia order wants {tag}"
        self.k += 1
        return e

    def real(self, tag):            # rand() / randexp() / randn(): the double itself
        return float(np.array([self._pop(tag)["bits"]], dtype="<u8").view("<f8")[0])

    def index(self, tag, n):        # rand(a:b) as an offset in [0, n)
        e = self._pop(tag)
        r = int(np.array([e["bits"]], dtype="<u8").view("<i8")[0])
        assert int(e["aux"]) == (n & 0xffffffff) and 0 <= r < n, (tag, n, r)
        return r


def _trunc_exp_rand(rp, tag, theta, lower, upper=None):
    """rand(truncated(Exponential(theta), lower, upper)) — Distributions.jl src/truncate.jl: rejection on theta * randexp() when
    the kept mass exceeds 0.25, else the quantile of lcdf + rand() * tp (the third branch, tp <= sqrt(eps), is not reachable here)."""
    import math
    cdf = lambda x: -math.expm1(-max(x, 0.0) / theta)
    lcdf = cdf(lower)
    tp = (cdf(upper) if upper is not None else 1.0) - lcdf
    if tp > 0.25:
        while True:
            r = theta * rp.real(tag)
            if r >= lower and (upper is None or r <= upper):
                return r
    assert tp > 1.5e-8
    return -theta * math.log1p(-(lcdf + rp.real(tag) * tp))


def _py_replay_plan(variates, schemas, *, n_molecules, mol_length, mol_coverage, insert_length, insert_stdev, singletons,
                    attempts, circular):
    """schemas: [(sequence length, reads_required)] in first-seen order.  Returns (molecules, inserts)."""
    import math
    rp = _Replay(variates)
    lowerbound = max(3000, int(mol_length * mol_coverage * 0.25)) if mol_coverage < 1 else mol_coverage * (insert_length + 0.75 * insert_stdev)
    assert (200 - insert_length) / insert_stdev < -0.6      # truncated(Normal, lower = 200) keeps > 0.3 of the mass: plain rejection
    current = [0] * len(schemas)
    live = list(range(len(schemas)))                         # keys(schema) in first-seen order (stated deviation i)
    mols, ins = [], []
    while live:                                              # while !isempty(schema)            Main.jl:44
        if n_molecules > 0:                                  # get_n_molecules                   Common.jl:89-91, Structs.jl:63-67
            x = _trunc_exp_rand(rp, "NMOL", float(n_molecules), 1.0, 2.5 * n_molecules)
            n_mol = int(round(x))                            # round(x, digits = 0): ties to even, like Python's round
        else:
            n_mol = abs(n_molecules)
        targets = [live[rp.index("TARGET", len(live))] for _ in range(n_mol)]     # rand(candidates, n_mol)  Main.jl:47
        for t in targets:
            seq_len = schemas[t][0]
            molsize, tries = seq_len + 1, 0                  # get_molecule_size                 Breakpoints.jl:149-163
            while molsize > seq_len and tries < 30:
                tries += 1
                molsize = _trunc_exp_rand(rp, "MOLSIZE", float(mol_length), float(lowerbound))
            molsize = int(min(molsize, seq_len))
            if singletons > 0.0 and rp.real("SINGLETON") <= singletons:           # calculate_n_frags    Breakpoints.jl:171-196
                n = 1
            else:
                max_possible = molsize / insert_length
                if max_possible == 1:
                    n = 1
                else:
                    _n = max(2.0, molsize * mol_coverage / insert_length) if mol_coverage < 1 else float(mol_coverage)
                    n = int(_trunc_exp_rand(rp, "NFRAGS", _n, 2.0, max_possible))

            def draw_sizes():                                # rand(params.insert_size, n_reads): randn() until >= the lower bound
                out = []
                for _ in range(n):
                    while True:
                        z = rp.real("INSERT")
                        if not z < (200 - insert_length) / insert_stdev:
                            out.append(insert_length + insert_stdev * z)
                            break
                return out
            sizes, tries = draw_sizes(), 0                   # calculate_insert_sizes            Breakpoints.jl:205-218
            while sum(sizes) > molsize and tries < attempts:     # float sum, left to right (stated deviation iii)
                tries += 1
                sizes = draw_sizes()
            assert not sum(sizes) > molsize
            lens = [int(s) for s in sizes]                   # trunc.(Int, read_sizes)
            adjusted_end = seq_len if circular else max(seq_len - molsize, 1)     # get_sequences        Breakpoints.jl:235-242
            start = 1 + rp.index("START", adjusted_end)
            remaining = (molsize + 1) - sum(lens)            # length(start:start+molsize) - sum     seq_breakpoints!  :21-35
            assert remaining >= 0
            cursor = start
            first = len(ins)
            for ln in lens:
                gap = rp.index("GAP", remaining + 1)
                s0 = cursor + gap
                ins.append((s0, s0 + ln - 1))
                cursor = s0 + ln
                remaining -= gap
            mols.append((t, start, molsize, len(lens), first))
            current[t] += len(lens)                          # tracker                            Main.jl:54
        live = [t for t in live if current[t] < schemas[t][1]]    # filter!(is_incomplete, schema)    Main.jl:58
    assert rp.k == len(variates), "the oracle drew variates the Julia order does not account for"
    return mols, ins, current


@pytest.mark.parametrize("kw", [
    dict(coverage=8, singletons=0.3, n_molecules=2, mol_length=20000),
    dict(coverage=5, singletons=0.0, n_molecules=-3, mol_length=30000, mol_coverage=0.5),      # fixed molecules per barcode
    dict(coverage=6, singletons=0.7, n_molecules=3, mol_length=8000, circular=True),
    dict(coverage=4, singletons=0.1, n_molecules=2, mol_length=20000, insert_length=400, insert_stdev=150),   # redraws at the 200 bound
])
def test_plan_rederived_in_python_from_the_recorded_variates(tmp_path, kw):
    rng = np.random.default_rng(33)
    p1 = os.path.join(str(tmp_path), "a.fa")
    p2 = os.path.join(str(tmp_path), "b.fa")
    # c3 is shorter than most molecules (30 draws, then the cap: Breakpoints.jl:155-161); c4 is the CI fixture's 1 004-base contig:
    # truncated(Exponential(2), 2, 2.008) keeps 1.5e-3 of the mass, so n_frags comes from the quantile branch
    recs1 = [("c1", random_bases(rng, 90_000)), ("c2", random_bases(rng, 25_000)), ("c3", random_bases(rng, 6_000)),
             ("c4", random_bases(rng, 1_004))]
    recs2 = [("c1", random_bases(rng, 90_500)), ("c2", random_bases(rng, 24_000)), ("c3", random_bases(rng, 6_100)),
             ("c4", random_bases(rng, 1_004))]
    write_fasta(p1, recs1)
    write_fasta(p2, recs2)
    res = O.mimick_single([p1, p2], "haplotagging", seed=77, attempts=300, record=True, **kw)
    par = dict(n_molecules=2, mol_length=80000, mol_coverage=0.2, insert_length=500, insert_stdev=50, singletons=0.0,
               attempts=300, circular=False)
    par.update({k: v for k, v in kw.items() if k in par})
    lens = [len(s) for _, s in recs1] + [len(s) for _, s in recs2]
    schemas = list(zip(lens, [int(r) for r in res.reads_required]))
    mols, ins, current = _py_replay_plan(res.variates, schemas, **par)
    m = res.molecules
    assert len(mols) == len(m) and len(mols) > 30
    got = np.array(mols, dtype=np.int64)
    for col, f in enumerate(("schema", "start", "molsize", "n_frags", "first_read")):
        assert np.array_equal(got[:, col], m[f].astype(np.int64)), f
    got_ins = np.array(ins, dtype=np.int64)
    assert np.array_equal(got_ins[:, 0], res.insert_start) and np.array_equal(got_ins[:, 1], res.insert_stop)
    assert current == [int(c) for c in res.reads_current]
    assert len({int(x) for x in m["schema"]}) == 8            # every schema was live and drawn


def _py_mutate(master: bytes, variants):
    """build_sample_schema's inner loop (ProcessVariants.jl:169-176) on a Python list: every SNP first (seq[pos] = alt[1]), then the
    non-SNPs in REVERSE file order, each spliceinto!(seq, pos:pos+len(ref)-1, alt) on the sequence as it then is."""
    seq = bytearray(master)
    for pos, typ, ref, alt in variants:
        if typ == 0:
            seq[pos - 1] = ord(alt[0])
    for pos, typ, ref, alt in reversed([v for v in variants if v[1] != 0]):
        seq[pos - 1:pos - 1 + len(ref)] = alt.encode()
    return bytes(seq)


@pytest.mark.parametrize("seed,sorted_sites", [(1, True), (2, True), (3, False)])
def test_mutate_rederived_in_python(seed, sorted_sites):
    """Random SNP / insertion / deletion / MNP sets — in file order as a VCF has them, and shuffled with overlapping REF spans,
    where the reference's reverse-order splice still has a literal meaning — against the oracle's mutate."""
    rng = np.random.default_rng(seed)
    master = random_bases(rng, 20_000)
    pos = np.sort(rng.choice(np.arange(2, 19_900), size=600, replace=False)) if sorted_sites else rng.integers(2, 19_900, size=600)
    variants = []
    for p in pos.tolist():
        kind = rng.integers(0, 4)
        if kind == 0:
            variants.append((p, 0, chr(master[p - 1]), "ACGT"[int(rng.integers(0, 4))]))
        elif kind == 1:      # insertion: REF = anchor, ALT = anchor + k bases
            variants.append((p, 1, chr(master[p - 1]), chr(master[p - 1]) + random_bases(rng, int(rng.integers(1, 12))).decode()))
        elif kind == 2:      # deletion: REF = anchor + k bases, ALT = anchor
            k = int(rng.integers(1, 12))
            variants.append((p, 1, master[p - 1:p + k].decode(), chr(master[p - 1])))
        else:                # MNP / complex: both longer than one base
            k = int(rng.integers(2, 6))
            variants.append((p, 1, master[p - 1:p - 1 + k].decode(), random_bases(rng, int(rng.integers(2, 6))).decode()))
    if sorted_sites:         # keep REF spans inside the still-untouched prefix, as a normalised VCF does
        keep, last_end = [], 0
        for v in variants:
            if v[0] > last_end:
                keep.append(v)
                last_end = v[0] + len(v[2]) - 1
        variants = keep
    assert O.Oracle.mutate(master, variants) == _py_mutate(master, variants)


def test_oracle_output_is_frozen():
    """The oracle's bytes for ten seeded cases (all formats, circular literal / strict, fixed molecules per barcode, N masking,
    Robert's truncated-Normal branch, the error model, the reference's FASTA + VCF fixtures) equal the committed digests:
    a change of oracle/ or of the shared include/mmk_rng.h that moves the output fails HERE, on the CPU — the GPU parity tests
    cannot see it, both sides share that header.  tests/golden/make_oracle_digests.py regenerates the file after a deliberate change."""
    import importlib.util
    import json
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_oracle_digests", os.path.join(here, "make_oracle_digests.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    want = json.load(open(os.path.join(here, "oracle_digests.json")))
    got = mod.compute()
    assert sorted(got) == sorted(want)
    for name in sorted(want):
        assert got[name] == want[name], f"{name}: the oracle's output moved ({[k for k in want[name] if got[name][k] != want[name][k]]})"
