"""Parity of the CUDA path (through the C ABI) against the CPU oracle — the first gate.

Every case runs the same seeded inputs through libmimick_b200.so and oracle/mimick_oracle.c and
requires byte-identical R1/R2 text, identical plans, and that the oracle, fed ONLY the variates the
kernels report having consumed (replay mode, no RNG on the CPU side), reproduces the same bytes.
Bar: bit-exact (integer / byte work; the few FP64 expressions are single-rounded on both sides).
"""
import os

import numpy as np
import pytest

import oracle as O
from helpers import assert_same_bytes, bgzf_inflate, bgzf_members, make_ci_haplotypes, random_bases, write_fasta

pytestmark = pytest.mark.gpu

CLI = dict(attempts=300, singletons=0.0)  # CLI defaults that differ from the Julia ones (mimick/__main__.py:46,51)


def gpu_single(fastas, fmt, seed, mask=False, variates=True, **kw):
    import mimick_b200 as M
    with M.Simulator(fmt, seed=seed, mask=mask, **kw) as sim:
        for h, p in enumerate(fastas, start=1):
            sim.load_fasta(p, h)
        plan = sim.plan()
        place = sim.place()
        r1, r2 = sim.emit_to_host()
        out = dict(r1=r1.tobytes(), r2=r2.tobytes(), mol=sim.molecules(), ins=sim.inserts(), plan=plan, place=place,
                   schemas=sim.schemas())
        # the BGZF stage on the same shard: blocked gzip that inflates to exactly the text above
        sim.rewind()
        z1, z2 = sim.emit_bgzf_to_host()
        out["z1"], out["z2"] = z1, z2
        assert_same_bytes(bgzf_inflate(z1), out["r1"], f"{fmt} BGZF R1")
        assert_same_bytes(bgzf_inflate(z2), out["r2"], f"{fmt} BGZF R2")
        if variates:
            out["var"] = np.concatenate([sim.mask_variates(), sim.variates()])
        return out


def compare(gpu, orc, what):
    assert_same_bytes(gpu["r1"], orc.r1, what + " R1")
    assert_same_bytes(gpu["r2"], orc.r2, what + " R2")
    gm, om = gpu["mol"], orc.molecules
    assert len(gm) == len(om)
    for f in ("group", "slot", "schema", "start", "molsize", "n_frags", "first_read", "barcode_index"):
        assert np.array_equal(gm[f], om[f]), f"{what}: molecule field {f}"
    assert np.array_equal(gpu["ins"][0], orc.insert_start)
    assert np.array_equal(gpu["ins"][1], orc.insert_stop)
    assert gpu["plan"].n_groups == orc.n_groups
    assert [s["reads_current"] for s in gpu["schemas"]] == orc.reads_current
    assert [s["reads_required"] for s in gpu["schemas"]] == orc.reads_required


def check_single(fastas, fmt, seed, mask=False, **kw):
    gpu = gpu_single(fastas, fmt, seed, mask=mask, **kw)
    okw = {k: v for k, v in kw.items() if k not in ("force_generic", "max_chunk_bytes", "plan_batch_groups")}
    orc = O.mimick_single(fastas, fmt, seed=seed, mask=mask, record=True, **okw)
    compare(gpu, orc, f"{fmt} philox")
    # the variates the kernels say they consumed == the variates the oracle consumed, bit for bit
    assert orc.variates is not None and len(orc.variates) == len(gpu["var"])
    assert np.array_equal(orc.variates["tag"], gpu["var"]["tag"])
    assert np.array_equal(orc.variates["bits"], gpu["var"]["bits"])
    # replay: the oracle sees no seed, only the GPU's stream
    rep = O.mimick_single(fastas, fmt, seed=0xDEAD, mask=mask, replay=gpu["var"], **okw)
    compare(gpu, rep, f"{fmt} replay")
    return gpu, orc


# ------------------------------------------------------------------------------------------
def test_device_philox_and_math_are_bit_exact():
    import mimick_b200 as M
    L = O.lib()
    with M.Simulator("haplotagging", seed=1) as sim:
        for ctr, key, want in [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
                               ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
                               ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
                                (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]:
            assert tuple(int(x) for x in sim.kat_philox(ctr, key)) == want
        rng = np.random.default_rng(3)
        x = np.exp(rng.uniform(-700, 700, 200000))
        got = sim.kat_math(0, x)
        want = np.array([L.orc_det_log(float(v)) for v in x[:20000]])
        assert np.array_equal(got[:20000].view(np.uint64), want.view(np.uint64))
        assert np.max(np.abs(got - np.log(x)) / np.maximum(np.abs(np.log(x)), 1e-300)) < 1e-15
        y = rng.uniform(-700, 700, 20000)
        got = sim.kat_math(1, y)
        want = np.array([L.orc_det_exp(float(v)) for v in y])
        assert np.array_equal(got.view(np.uint64), want.view(np.uint64))


@pytest.mark.parametrize("fmt", ["haplotagging", "stlfr", "tellseq", "10x", "standard:stlfr", "standard", "standard:haplotagging"])
@pytest.mark.parametrize("mask", [False, True])
def test_ci_single_sample(tmp_path, fmt, mask):
    """The reference's CI matrix (.github/workflows/tests.yml:47-57): two generated haplotypes with a
    1000-base contig + NNNN, -g 10, CLI defaults. Exercises the 4-bit path, N masking, the 30-attempt
    molecule fallback, the quantile branch of the truncated exponential and the 1-base wrap."""
    fastas = make_ci_haplotypes(tmp_path, seed=11, big=60000)
    check_single(fastas, fmt, seed=5, mask=mask, coverage=10, **CLI)


def test_c1_multisample(ref_fasta, ref_vcf):
    """BASELINE config 1: test/fasta.fa.gz + test/test.vcf.gz, haplotagging, -g 5, all 7 samples."""
    import mimick_b200 as M
    orc = O.mimick_multi(ref_fasta, ref_vcf, "haplotagging", seed=42, coverage=5, record=True, **CLI)
    vcf = M.Vcf(ref_vcf)
    assert vcf.samples == [f"tsk_{i}" for i in range(7)] and vcf.ploidy == 2 and vcf.n_records == 13705
    replay = {}
    with M.Simulator("haplotagging", seed=42, coverage=5, **CLI) as sim:
        sim.load_fasta(ref_fasta, 1)
        for si, name in enumerate(vcf.samples):
            sim.apply_vcf_sample(vcf, si)
            plan = sim.plan()
            place = sim.place()
            assert place.n_pairs == plan.n_pairs
            r1, r2 = sim.emit_to_host()
            gpu = dict(r1=r1.tobytes(), r2=r2.tobytes(), mol=sim.molecules(), ins=sim.inserts(), plan=plan,
                       schemas=sim.schemas())
            compare(gpu, orc[name], name)
            assert [s["reads_required"] for s in gpu["schemas"]] == [25000, 25000]
            replay[name] = sim.variates()
            if si < 2:  # haplotype construction itself
                vs = O.sample_variants(O.read_vcf(ref_vcf), si)
                master = O.read_fasta(ref_fasta)[0][1].upper()
                for h in (1, 2):
                    assert sim.schema_bases(h - 1) == O.Oracle.mutate(master, vs[f"chr1_{h}"])
    rep = O.mimick_multi(ref_fasta, ref_vcf, "haplotagging", seed=1, coverage=5, replay=replay, samples=["tsk_0", "tsk_3"], **CLI)
    for name in ("tsk_0", "tsk_3"):
        assert rep[name].r1 == orc[name].r1 and rep[name].r2 == orc[name].r2


@pytest.mark.parametrize("strict", [False, True])
def test_circular(tmp_path, strict):
    """circular=true: molecule starts anywhere; literal mode reproduces the reference's wrap
    (reads that START past the contig end become hap[1 : stop % n]), strict mode wraps
    modulo n."""
    rng = np.random.default_rng(2)
    p = os.path.join(str(tmp_path), "circ.fa")
    write_fasta(p, [("plasmid", random_bases(rng, 150000))])
    gpu, orc = check_single([p], "stlfr", seed=9, coverage=40, circular=True, strict_wrap=strict, mol_length=40000, **CLI)
    lens = {len(l) for l in gpu["r1"].split(b"\n")[1::4]}
    if strict:
        assert lens == {150}
    else:
        assert max(lens) > 150  # the D5 giant reads are really there


@pytest.mark.parametrize("kw", [
    dict(singletons=0.35, attempts=300),
    dict(n_molecules=-3, singletons=0.0, attempts=50),
    dict(n_molecules=5, mol_coverage=6.0, singletons=0.1, attempts=300),
    dict(read_length=(132, 150), insert_length=350, insert_stdev=80, singletons=0.0, attempts=300),
    dict(read_length=(150, 108), mol_length=20000, mol_coverage=0.5, singletons=0.0, attempts=300),
    dict(mol_length=300, singletons=0.0, attempts=300),   # molecule-length law in its quantile branch (tp = e^-10)
    dict(read_length=(250, 250), insert_length=600, insert_stdev=150, singletons=0.0, attempts=300),
    dict(read_length=(192, 161), insert_length=450, insert_stdev=60, singletons=0.0, attempts=300),   # the fast path's longest reads
    dict(read_length=(17, 191), insert_length=300, insert_stdev=30, singletons=0.2, attempts=300),    # ... and its shortest
    # truncated(Normal(ins, sd), lower = 200) outside plain rejection (src/Structs.jl:68; the CLI admits -i 0.., mimick/__main__.py:44):
    # (the truncated mean exceeds insert_length there, so molecules whose insert count comes out near molsize / insert_length
    # cannot fit — in the reference too; mol_coverage 0.02 keeps the counts far from that bound)
    dict(insert_length=150, insert_stdev=50, read_length=(100, 100), mol_coverage=0.02, singletons=0.0, attempts=300),   # kept mass 0.16: Robert's exponential proposal
    dict(insert_length=210, insert_stdev=200, read_length=(100, 100), mol_coverage=0.02, singletons=0.0, attempts=300),  # kept mass 0.52: rejection, half the normals rejected
    dict(insert_length=100, insert_stdev=20, read_length=(75, 75), mol_coverage=0.02, singletons=0.0, attempts=300),     # 5 sigma into the tail
])
def test_parameter_regions(tmp_path, kw):
    rng = np.random.default_rng(4)
    p1 = os.path.join(str(tmp_path), "a.fa")
    p2 = os.path.join(str(tmp_path), "b.fa.gz")
    write_fasta(p1, [("chrA", random_bases(rng, 90000)), ("chrB", random_bases(rng, 30000))])
    write_fasta(p2, [("chrA", random_bases(rng, 90100)), ("chrB", random_bases(rng, 29000))])
    check_single([p1, p2], "tellseq", seed=77, coverage=8, **kw)


def test_unfittable_inserts_fail_like_the_reference(tmp_path):
    """-i 150 -d 50 at the default molecule coverage: the truncated insert law has mean 226 > 150, so a molecule whose insert
    count lands near molsize / 150 can never fit: "FAILED TO FIND PRACTICAL READ SIZES" (Breakpoints.jl:214-216) after
    `attempts` whole-vector redraws — same status, same molecule as the oracle."""
    import mimick_b200 as M
    rng = np.random.default_rng(4)
    p = os.path.join(str(tmp_path), "a.fa")
    write_fasta(p, [("chrA", random_bases(rng, 200000))])
    kw = dict(insert_length=150, insert_stdev=50, read_length=(100, 100), singletons=0.0, attempts=5, coverage=8)
    with pytest.raises(O.OracleError) as eo:
        O.mimick_single([p], "tellseq", seed=77, **kw)
    assert eo.value.code == 3
    with pytest.raises(M.MimickError) as eg:
        gpu_single([p], "tellseq", 77, variates=False, **kw)
    assert eg.value.status == 3


def test_iupac_ambiguity_codes_4bit(tmp_path):
    """IUPAC ambiguity codes, N runs, gaps and lower case: the 4-bit reference path (fast tiles convert
    non-ACGT symbols one by one; complement of an ambiguity code is its bit-reversed mask)."""
    rng = np.random.default_rng(10)
    seqs = []
    for L in (70000, 50000):
        a = np.frombuffer(random_bases(rng, L), dtype=np.uint8).copy()
        idx = rng.integers(0, L, size=L // 40)
        a[idx] = np.frombuffer(b"MRWSYKVHDBN-", dtype=np.uint8)[rng.integers(0, 12, size=len(idx))]
        a[1000:1400] = ord("N")
        low = rng.integers(0, L, size=L // 10)
        a[low] = np.where((a[low] >= 65) & (a[low] <= 90), a[low] + 32, a[low])
        seqs.append(a.tobytes())
    p = os.path.join(str(tmp_path), "amb.fa")
    write_fasta(p, [("ctgA", seqs[0]), ("ctgB", seqs[1])])
    gpu, orc = check_single([p], "haplotagging", seed=21, coverage=12, **CLI)
    assert b"N" in gpu["r1"] and b"R" in gpu["r2"] and b"-" in gpu["r1"]
    forced = gpu_single([p], "haplotagging", 21, coverage=12, variates=False, force_generic=True, **CLI)
    assert forced["r1"] == gpu["r1"] and forced["r2"] == gpu["r2"]
    masked = gpu_single([p], "haplotagging", 21, mask=True, coverage=12, variates=False, **CLI)
    assert b"N" not in masked["r1"].replace(b"1:N:0", b"")  # mask_N! replaces exactly the N's (other codes stay)


def test_many_schemas_many_epochs(tmp_path):
    """40 contigs × 2 haplotypes of very different lengths: dozens of pruning epochs, small batches."""
    rng = np.random.default_rng(8)
    recs1 = [(f"ctg{i}", random_bases(rng, int(6000 + 3000 * i))) for i in range(40)]
    recs2 = [(n, random_bases(rng, len(s) + 17)) for n, s in recs1]
    p1, p2 = os.path.join(str(tmp_path), "h1.fa"), os.path.join(str(tmp_path), "h2.fa")
    write_fasta(p1, recs1)
    write_fasta(p2, recs2)
    gpu, orc = check_single([p1, p2], "standard:haplotagging", seed=3, coverage=12, plan_batch_groups=512, **CLI)
    assert gpu["plan"].n_epochs >= 20


def test_generic_path_equals_fast_path(tmp_path):
    fastas = make_ci_haplotypes(tmp_path, seed=21, big=50000)
    a = gpu_single(fastas, "10x", 6, coverage=6, variates=False, **CLI)
    b = gpu_single(fastas, "10x", 6, coverage=6, variates=False, force_generic=True, **CLI)
    assert a["r1"] == b["r1"] and a["r2"] == b["r2"]


def test_sharding_and_chunking(tmp_path):
    import mimick_b200 as M
    fastas = make_ci_haplotypes(tmp_path, seed=31, big=80000)
    whole = gpu_single(fastas, "haplotagging", 12, coverage=9, variates=False, **CLI)
    parts1, parts2, pairs = [], [], 0
    for rank in range(3):
        with M.Simulator("haplotagging", seed=12, coverage=9, max_chunk_bytes=1 << 20, **CLI) as sim:
            for h, p in enumerate(fastas, start=1):
                sim.load_fasta(p, h)
            sim.plan()
            sim.set_shard(rank, 3)
            place = sim.place()
            pairs += place.n_pairs
            n_chunks = 0
            while True:
                ch = sim.emit_chunk()
                if ch.n_pairs:
                    a1, a2 = sim.copy_chunk(ch)
                    parts1.append(a1.tobytes())
                    parts2.append(a2.tobytes())
                    n_chunks += 1
                if ch.done:
                    break
            assert n_chunks > 1
            r1, r2 = sim.emit_to_host()  # pipelined path over the same ring
            assert r1.tobytes() == b"".join(parts1)[-len(r1):]
    assert pairs == whole["plan"].n_pairs
    assert b"".join(parts1) == whole["r1"] and b"".join(parts2) == whole["r2"]


def _write_vcf(path, contig, length, samples, records):
    with open(path, "w") as fh:
        fh.write("##fileformat=VCFv4.2\n")
        fh.write(f"##contig=<ID={contig},length={length}>\n")
        fh.write('##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">\n')
        fh.write("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(samples) + "\n")
        for pos, ref, alts, gts in records:
            fh.write(f"{contig}\t{pos}\t.\t{ref}\t{','.join(alts)}\t.\tPASS\t.\tGT\t" + "\t".join(gts) + "\n")


def test_indels_and_multiallelic(tmp_path):
    """The reference's fixture has no indel; synthesise SNPs, insertions, deletions, MNPs and
    multi-allelic sites and check haplotype construction + FASTQ against the oracle's literal splices."""
    import mimick_b200 as M
    rng = np.random.default_rng(5)
    L = 120000
    ref = random_bases(rng, L)
    fa = os.path.join(str(tmp_path), "ref.fa")
    write_fasta(fa, [("chr7", ref)])
    recs, pos = [], 50
    samples = ["s0", "s1", "s2"]
    while pos < L - 100:
        kind = rng.integers(0, 5)
        r = ref[pos - 1:pos].decode()
        alt1 = "ACGT"[("ACGT".index(r) + 1) % 4]
        if kind == 0:
            ref_a, alts = r, [alt1]
        elif kind == 1:
            ref_a, alts = r, [r + random_bases(rng, int(rng.integers(1, 20))).decode()]
        elif kind == 2:
            dl = int(rng.integers(1, 20))
            ref_a, alts = ref[pos - 1:pos + dl].decode(), [r]
        elif kind == 3:
            ref_a, alts = r, [alt1, r + "GG"]  # typed :indel for every allele because one ALT is long
        else:
            ref_a, alts = ref[pos - 1:pos + 2].decode(), [random_bases(rng, 3).decode()]
        gts = ["|".join(str(int(rng.integers(0, len(alts) + 1))) for _ in range(2)) for _ in samples]
        recs.append((pos, ref_a, alts, gts))
        pos += len(ref_a) + int(rng.integers(1, 400))
    vcfp = os.path.join(str(tmp_path), "v.vcf")
    _write_vcf(vcfp, "chr7", L, samples, recs)
    orc = O.mimick_multi(fa, vcfp, "tellseq", seed=19, coverage=6, **CLI)
    vcf = M.Vcf(vcfp)
    with M.Simulator("tellseq", seed=19, coverage=6, **CLI) as sim:
        sim.load_fasta(fa, 1)
        for si, name in enumerate(samples):
            sim.apply_vcf_sample(vcf, si)
            vs = O.sample_variants(O.read_vcf(vcfp), si)
            for h in (1, 2):
                want = O.Oracle.mutate(ref, vs[f"chr7_{h}"])
                assert sim.schema_bases(h - 1) == want, f"{name} hap {h}"
            plan = sim.plan()
            sim.place()
            r1, r2 = sim.emit_to_host()
            assert_same_bytes(r1.tobytes(), orc[name].r1, name + " R1")
            assert_same_bytes(r2.tobytes(), orc[name].r2, name + " R2")


def test_error_sites(tmp_path):
    import mimick_b200 as M
    rng = np.random.default_rng(6)
    with pytest.raises(M.MimickError) as e:
        M.Simulator("nanopore")
    assert e.value.status == 2
    with pytest.raises(M.MimickError) as e:
        M.mimick("x.fa", "calls.BCF", "stlfr")
    assert e.value.status == 6
    with M.Simulator("stlfr", seed=1, **CLI) as sim:
        with pytest.raises(M.MimickError) as e:
            sim.add_contig("tiny", b"ACGT" * 50 + b"N" * 100)  # 200 unambiguous < round(0.95 * 300)
        assert e.value.status == 5
        with pytest.raises(M.MimickError) as e:
            sim.add_contig("bad", b"ACGT" * 100 + b"J" + b"ACGT" * 100)
        assert e.value.status == 8
        with pytest.raises(M.MimickError) as e:
            sim.load_fasta(os.path.join(str(tmp_path), "missing.fa"))
        assert e.value.status == 9
        # a contig shorter than two inserts: truncated(Exponential, 2.0, 700/500) → lower > upper
        sim.add_contig("short", random_bases(rng, 700))
        with pytest.raises(M.MimickError) as e:
            sim.plan()
        assert e.value.status == 7
    with pytest.raises(O.OracleError) as eo:
        p = os.path.join(str(tmp_path), "short.fa")
        write_fasta(p, [("short", random_bases(np.random.default_rng(6), 700))])
        O.mimick_single([p], "stlfr", seed=1, **CLI)
    assert eo.value.code == 7
    # insert sizes that do not fit: two ~1000±50 bp inserts in a 2004-base molecule with no redraw allowed
    # fail for about half of the molecules (float sum > molsize)
    p = os.path.join(str(tmp_path), "c.fa")
    write_fasta(p, [("c", random_bases(rng, 2004))])
    with M.Simulator("stlfr", seed=1, insert_length=1000, insert_stdev=50, attempts=0, singletons=0.0, mol_coverage=2.0) as sim:
        sim.load_fasta(p)
        sim.plan()
        with pytest.raises(M.MimickError) as e:
            sim.place()
        assert e.value.status == 3


def test_mimick_function_writes_the_reference_file_layout(tmp_path, ref_fasta, ref_vcf):
    """The public function with the reference's two methods and keywords (Main.jl:24, :88): file naming
    `<prefix>.R{1,2}.fq.gz` and `<outdir>/<sample>.R{1,2}.fq.gz` (BGZF, as the reference's writers), contents
    identical to the oracle's after inflation; `compress=False` gives the plain text."""
    import gzip

    import mimick_b200 as M
    fastas = make_ci_haplotypes(tmp_path, seed=41, big=30000)
    prefix = os.path.join(str(tmp_path), "single", "SIM")
    out = M.mimick(fastas, "stlfr", prefix=prefix, coverage=6, seed=5, mask=True, compress=False, **CLI)
    orc = O.mimick_single(fastas, "stlfr", seed=5, mask=True, coverage=6, **CLI)
    assert open(prefix + ".R1.fq", "rb").read() == orc.r1 and open(prefix + ".R2.fq", "rb").read() == orc.r2
    assert out[0].n_pairs == len(orc.insert_start)
    out = M.mimick(fastas, "stlfr", prefix=prefix, coverage=6, seed=5, mask=True, **CLI)
    assert out[0].r1_path == prefix + ".R1.fq.gz"
    z = open(prefix + ".R1.fq.gz", "rb").read()
    assert z.endswith(M.bgzf_eof()) and gzip.decompress(z) == orc.r1
    assert gzip.decompress(open(prefix + ".R2.fq.gz", "rb").read()) == orc.r2
    # two "GPUs" on one device: the parts concatenate to the same files
    for rank in range(2):
        M.mimick(fastas, "stlfr", prefix=prefix + "_sh", coverage=6, seed=5, mask=True, rank=rank, world=2, **CLI)
    cat1 = b"".join(open(f"{prefix}_sh.R1.fq.gz.part{r}", "rb").read() for r in range(2))
    assert cat1.endswith(M.bgzf_eof()) and cat1.count(M.bgzf_eof()) == 1 and gzip.decompress(cat1) == orc.r1
    outdir = os.path.join(str(tmp_path), "multi")
    outs = M.mimick(ref_fasta, ref_vcf, "standard:haplotagging", outdir=outdir, coverage=0.5, seed=9, rank=1, world=4, **CLI)
    assert [o.name for o in outs] == ["tsk_1", "tsk_5"]   # samples round-robin over ranks
    want = O.mimick_multi(ref_fasta, ref_vcf, "standard:haplotagging", seed=9, coverage=0.5, samples=["tsk_1", "tsk_5"], **CLI)
    for o in outs:
        assert gzip.decompress(open(o.r1_path, "rb").read()) == want[o.name].r1
        assert gzip.decompress(open(o.r2_path, "rb").read()) == want[o.name].r2
        assert os.path.basename(o.r1_path) == o.name + ".R1.fq.gz"
    with pytest.raises(TypeError):
        M.mimick(fastas, "stlfr", gzip=True)


def test_bgzf_members_and_ratio(tmp_path):
    """Structure of the coded stream (SAM spec 4.1: gzip member with the BC extra field, BSIZE, CRC-32, ISIZE), several
    chunks through the ring, and the coding gain on the headline format: every member inflates on its own, sizes add
    up, and haplotagging text shrinks at least 4.5x (A/C/G/T at 2-3 bits, headers and quality lines as matches)."""
    import zlib

    import mimick_b200 as M
    rng = np.random.default_rng(77)
    p = os.path.join(str(tmp_path), "a.fa")
    write_fasta(p, [("chr1", random_bases(rng, 3_000_000))])
    with M.Simulator("haplotagging", seed=3, coverage=2.0, max_chunk_bytes=3 << 20, **CLI) as sim:
        sim.load_fasta(p)
        sim.plan()
        place = sim.place()
        assert sim.chunk_layout()[1] >= 3
        r1, r2 = (x.tobytes() for x in sim.emit_to_host())
        sim.rewind()
        z1, z2 = sim.emit_bgzf_to_host()
        # chunk-at-a-time entry point gives the same coded bytes
        sim.rewind()
        parts = []
        while True:
            ch = sim.emit_chunk_bgzf()
            if ch.n_pairs == 0 and ch.done:
                break
            a, b = sim.copy_chunk(ch)
            parts.append((a.tobytes(), b.tobytes(), ch.r1_offset))
            if ch.done:
                break
        assert b"".join(x[0] for x in parts) == z1 and b"".join(x[1] for x in parts) == z2
        assert [x[2] for x in parts] == list(np.cumsum([0] + [len(x[0]) for x in parts[:-1]]))
    # the kernel's token choice and bit packing are specified by scripts/bgzf_table.py (tokenize / deflate_block, a
    # bit-exact Python model built on the same code lengths): the first members must equal the model byte for byte
    import re
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "scripts"))
    import bgzf_table as T
    txt = open(os.path.join(root, "mimick_b200", "csrc", "mmk_bgzf_table.h")).read()
    ll_len = [int(x) for x in re.search(r"LL_LEN\[286\] = \{([^}]*)\}", txt).group(1).split(",")]
    d_len = [int(x) for x in re.search(r"D_LEN\[30\] = \{([^}]*)\}", txt).group(1).split(",")]
    for z, raw in ((z1, r1), (z2, r2)):
        pos = 0
        for i, (off, size, isize, crc) in enumerate(bgzf_members(z)):
            if i >= 12:
                break
            assert z[off + 18:off + size - 8] == T.deflate_block(raw[pos:pos + isize], ll_len, d_len), f"member {i} differs from the model"
            pos += isize
    for z, raw in ((z1, r1), (z2, r2)):
        pos = 0
        for off, size, isize, crc in bgzf_members(z):
            payload = zlib.decompress(z[off + 18:off + size - 8], wbits=-15)
            assert len(payload) == isize and zlib.crc32(payload) == crc
            assert payload == raw[pos:pos + isize]
            assert payload.count(b"\n") == 4 * 32 or pos + isize == len(raw)   # one tile of 32 records per member
            pos += isize
        assert pos == len(raw)
        assert len(raw) / len(z) > 4.5, len(raw) / len(z)
    assert place.r1_bytes == len(r1)


def test_command_line_runs_like_the_reference_cli(tmp_path):
    """`python -m mimick_b200` with the reference's options (mimick/__main__.py): exit code 0, the two .fq.gz files of the
    single-sample mode, contents equal to the oracle run with the same parameters; a Julia-side error (here: a contig
    too short for the reads) gives exit code 1 like the reference."""
    import gzip
    import subprocess
    import sys
    fastas = make_ci_haplotypes(tmp_path, seed=21, big=40000)
    prefix = os.path.join(str(tmp_path), "cli", "S")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "mimick_b200", "-o", prefix, "-s", "11", "-f", "stlfr", "-g", "4", "-l", "150,108", "-q", "-m"] + fastas
    r = subprocess.run(cmd, cwd=root, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    orc = O.mimick_single(fastas, "stlfr", seed=11, mask=True, coverage=4, read_length=[150, 108], **CLI)
    assert gzip.decompress(open(prefix + ".R1.fq.gz", "rb").read()) == orc.r1
    assert gzip.decompress(open(prefix + ".R2.fq.gz", "rb").read()) == orc.r2
    tiny = os.path.join(str(tmp_path), "tiny.fa")
    write_fasta(tiny, [("t", b"ACGT" * 20)])
    r = subprocess.run([sys.executable, "-m", "mimick_b200", "-o", prefix + "2", tiny], cwd=root, capture_output=True, text=True)
    assert r.returncode == 1 and "ERROR" in r.stderr


def test_contig_symbols_may_already_be_on_the_device(tmp_path):
    """mmk_add_contig takes host or device memory: a multi-GPU caller uploads each contig once and replicates it over
    NVLink (bench.py does); the output must not depend on where the symbols came from."""
    import torch

    import mimick_b200 as M
    rng = np.random.default_rng(12)
    seqs = [("c1", random_bases(rng, 300000)), ("c2", random_bases(rng, 200000, b"ACGTN"))]
    outs = []
    for on_device in (False, True):
        with M.Simulator("tellseq", seed=4, coverage=3, mask=True, **CLI) as sim:
            keep = []
            for name, seq in seqs:
                t = torch.frombuffer(bytearray(seq), dtype=torch.uint8)
                if on_device:
                    t = t.cuda()
                keep.append(t)
                sim.add_contig(name, (t.data_ptr(), t.numel()), 1)
            outs.append(sim.run())
    assert outs[0] == outs[1]
    p = os.path.join(str(tmp_path), "a.fa")
    write_fasta(p, seqs)
    orc = O.mimick_single([p], "tellseq", seed=4, mask=True, coverage=3, **CLI)
    assert outs[0][0] == orc.r1 and outs[0][1] == orc.r2


def test_bgzf_fallbacks_and_consumer_errors(tmp_path):
    """BGZF edge cases: (i) text with a byte the fixed code has no symbol for (a contig name with a non-ASCII
    character) must come out as stored members that still inflate to the oracle's text; (ii) very short reads
    (more newlines per lane than the register queue holds) still code correctly; (iii) a consumer that fails
    aborts the stream with its status."""
    import mimick_b200 as M
    rng = np.random.default_rng(31)
    p = os.path.join(str(tmp_path), "u.fa")
    write_fasta(p, [("chré", random_bases(rng, 120000))])
    gpu = gpu_single([p], "haplotagging", 6, coverage=4, variates=False, **CLI)   # checks inflate(BGZF) == text inside
    orc = O.mimick_single([p], "haplotagging", seed=6, coverage=4, **CLI)
    assert gpu["r1"] == orc.r1 and gpu["r2"] == orc.r2 and "chré".encode() in orc.r1[:40]
    stored = sum(1 for off, size, isize, crc in bgzf_members(gpu["z1"]) if size == isize + 18 + 5 + 8)
    assert stored == sum(1 for _ in bgzf_members(gpu["z1"]))          # every member is a stored block
    q = os.path.join(str(tmp_path), "s.fa")
    write_fasta(q, [("c", random_bases(rng, 90000))])
    short = dict(read_length=[16, 20], insert_length=260, insert_stdev=10, mol_length=5000)
    gpu = gpu_single([q], "stlfr", 8, coverage=3, variates=False, **short, **CLI)
    orc = O.mimick_single([q], "stlfr", seed=8, coverage=3, **short, **CLI)
    assert gpu["r1"] == orc.r1 and gpu["r2"] == orc.r2 and len(gpu["r1"]) / len(gpu["z1"]) > 2.0
    with M.Simulator("stlfr", seed=8, coverage=3, **short, **CLI) as sim:
        sim.load_fasta(q)
        sim.plan()
        sim.place()
        cap = sim.bgzf_bound()
        slots = [np.empty(cap, dtype=np.uint8) for _ in range(4)]

        def boom(ch, p1, p2):
            raise RuntimeError("disk full")
        with pytest.raises(M.MimickError):
            sim.emit_stream([slots[0].ctypes.data, slots[1].ctypes.data], [slots[2].ctypes.data, slots[3].ctypes.data], cap, boom, bgzf=True)
        with pytest.raises(M.MimickError):   # slots smaller than the bound are refused before anything runs
            sim.emit_stream([slots[0].ctypes.data, slots[1].ctypes.data], [slots[2].ctypes.data, slots[3].ctypes.data], 1024, None, bgzf=True)


# ------------------------------------------------------------------------------------------
# placement: the fused kernel (mmk_place.cuh), its long-molecule side kernel, the legacy three-kernel path
# ------------------------------------------------------------------------------------------
def test_device_insert_length_table():
    """mmk_kat_insert_len: the kernels' table lookup (with its exact fallback) against the oracle's evaluation of the
    definition, on random draws and on both edges of the guard band of every threshold."""
    import ctypes as C
    import mimick_b200 as M
    from mimick_b200 import _lib
    L = _lib.load()
    for mu, sd in [(500, 50), (350, 80), (600, 150), (210, 200)]:
        en, n, base, ulo, uhi, guard = (C.c_uint32() for _ in range(6))
        thr = np.zeros(4096, dtype=np.uint32)
        assert L.mmk_insert_table(mu, sd, C.byref(en), C.byref(n), C.byref(base), C.byref(ulo), C.byref(uhi), C.byref(guard),
                                  thr.ctypes.data, thr.size) == 0 and en.value == 1
        t = thr[:n.value].astype(np.int64)
        g = guard.value
        rng = np.random.default_rng(mu + sd)
        edges = np.concatenate([t + d for d in (-g - 2, -g - 1, -g, -1, 0, 1, g, g + 1, g + 2)])
        edges = edges[(edges >= 0) & (edges < (1 << 32))].astype(np.uint64)
        bits = np.concatenate([rng.integers(0, 1 << 64, size=1_000_000, dtype=np.uint64), edges << np.uint64(32),
                               (edges << np.uint64(32)) | np.uint64(0xffffffff)])
        with M.Simulator("haplotagging", seed=1, insert_length=mu, insert_stdev=sd) as sim:
            got, slow = sim.kat_insert_len(bits)
        want = O.insert_len_from_bits(mu, sd, bits)
        assert np.array_equal(got, want), (mu, sd, int(np.flatnonzero(got != want)[0]))
        from scipy import stats
        rejected = stats.norm.cdf((200 - mu) / sd)   # draws the truncation at 200 rejects always take the exact path
        assert slow[:1_000_000].mean() < rejected + 0.01 and slow[1_000_000:].mean() > 0.3   # the fallback is rare, except where we aimed at it


def _env(**kv):
    import contextlib

    @contextlib.contextmanager
    def cm():
        old = {k: os.environ.get(k) for k in kv}
        os.environ.update({k: str(v) for k, v in kv.items()})
        try:
            yield
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
    return cm()


@pytest.mark.parametrize("kw", [
    dict(coverage=15, singletons=0.1),
    dict(coverage=15, singletons=1.0),                                  # every molecule one pair
    dict(coverage=20, mol_length=2_000_000, mol_coverage=0.5, singletons=0.0),   # molecules of 1000s of pairs
    dict(coverage=15, mol_length=400_000, mol_coverage=0.6, singletons=0.3, n_molecules=4),   # a mix of short and long molecules
])
def test_placement_paths_agree_with_the_oracle(tmp_path, kw):
    """The default placement path (one CTA per block of molecules, big molecules through their own kernel) and the
    three-kernel path give the oracle's bytes, plan and inserts."""
    rng = np.random.default_rng(14)
    p1 = os.path.join(str(tmp_path), "a.fa")
    p2 = os.path.join(str(tmp_path), "b.fa")
    write_fasta(p1, [("chrA", random_bases(rng, 5_000_000)), ("chrB", random_bases(rng, 700_000))])
    write_fasta(p2, [("chrA", random_bases(rng, 5_000_100)), ("chrB", random_bases(rng, 699_000))])
    okw = dict(kw, attempts=300)
    gpu, orc = check_single([p1, p2], "haplotagging", seed=31, **okw)
    nf = gpu["mol"]["n_frags"]
    if kw.get("mol_length", 0) >= 400_000:
        assert (nf > 1024).sum() > 3 and (nf <= 1024).sum() > 3   # both sides of PB_BIG (mmk_place.cuh)
    for env in (dict(MMK_PLACE_MODE="block"), dict(MMK_PLACE_MODE="legacy")):
        with _env(**env):
            alt = gpu_single([p1, p2], "haplotagging", 31, variates=False, **okw)
        assert alt["r1"] == gpu["r1"] and alt["r2"] == gpu["r2"], env
        assert np.array_equal(alt["ins"][0], gpu["ins"][0]) and np.array_equal(alt["ins"][1], gpu["ins"][1]), env


# ------------------------------------------------------------------------------------------
# sequencing-error model: optional, default off; byte parity with the oracle's definition
# ------------------------------------------------------------------------------------------
ERR = dict(error_rate=(0.002, 0.02), indel_error=(0.001, 0.0015))


def _err_case(fastas, fmt, seed, **kw):
    gpu = gpu_single(fastas, fmt, seed, variates=False, **kw)
    orc = O.mimick_single(fastas, fmt, seed=seed, **{k: v for k, v in kw.items() if k != "force_generic"})
    compare(gpu, orc, f"{fmt} error model")
    return gpu, orc


@pytest.mark.parametrize("fmt", ["haplotagging", "10x", "standard:stlfr"])
def test_error_model_matches_the_oracle(tmp_path, fmt):
    """Fast path (k_emit_w's error variant): substitutions, insertions, deletions and per-cycle qualities equal the
    oracle's erroneous_read byte for byte; the text really differs from the error-free run."""
    rng = np.random.default_rng(21)
    p1 = os.path.join(str(tmp_path), "a.fa")
    p2 = os.path.join(str(tmp_path), "b.fa")
    write_fasta(p1, [("chrA", random_bases(rng, 400_000)), ("chrB", random_bases(rng, 90_000))])
    write_fasta(p2, [("chrA", random_bases(rng, 400_100)), ("chrB", random_bases(rng, 89_000))])
    kw = dict(coverage=12, attempts=300, singletons=0.2)
    gpu, _ = _err_case([p1, p2], fmt, 17, **kw, **ERR)
    clean = gpu_single([p1, p2], fmt, 17, variates=False, **kw)
    assert len(clean["r1"]) == len(gpu["r1"]) and clean["r1"] != gpu["r1"]
    quals = set(gpu["r2"].split(b"\n")[3::4])
    assert len(quals) > 100  # qualities vary per cycle
    # the general writer applies the same model
    alt = gpu_single([p1, p2], fmt, 17, variates=False, force_generic=True, **kw, **ERR)
    assert alt["r1"] == gpu["r1"] and alt["r2"] == gpu["r2"]


@pytest.mark.parametrize("strict", [False, True])
def test_error_model_circular_and_iupac(tmp_path, strict):
    """Reads near the contig ends (clamped, or modular with strict_wrap), tiles that take the general writer, and a
    4-bit reference with IUPAC symbols (substitutions leave them alone)."""
    rng = np.random.default_rng(22)
    p = os.path.join(str(tmp_path), "circ.fa")
    seq = bytearray(random_bases(rng, 120_000))
    for i in rng.integers(0, len(seq), size=300):
        seq[i] = b"RYKMSWN"[i % 7]
    write_fasta(p, [("plasmid", bytes(seq))])
    _err_case([p], "stlfr", 9, coverage=30, circular=True, strict_wrap=strict, mol_length=30000, attempts=300, singletons=0.0, **ERR)


def test_error_rates_are_validated():
    import mimick_b200 as M
    with pytest.raises(M.MimickError):
        M.Simulator("haplotagging", seed=1, error_rate=(0.4, 0.4), indel_error=(0.1, 0.1))
    with pytest.raises(M.MimickError):
        M.Simulator("haplotagging", seed=1, error_rate=(-0.1, 0.0))


def test_long_contig_names(tmp_path):
    """Assembly-style contig names: 120 characters (records go through the general writer: header prefix > 64 bytes) next
    to a short one; the reference has no limit, the library's is 160."""
    import mimick_b200 as M
    rng = np.random.default_rng(31)
    p = os.path.join(str(tmp_path), "long.fa")
    long_name = "scaffold_" + "x" * 99 + "_pilon_pilon"
    assert len(long_name) == 120
    write_fasta(p, [(long_name + " some description", random_bases(rng, 60_000)), ("ctg2", random_bases(rng, 50_000))])
    gpu, _ = check_single([p], "tellseq", seed=4, coverage=6, **CLI)
    assert (b"@" + long_name.encode() + b":1|") in gpu["r1"]
    with M.Simulator("tellseq", seed=4, coverage=6, **CLI) as sim:    # the schema view gives the name in full
        sim.load_fasta(p)
        assert [s["name"] for s in sim.schemas()] == [long_name, "ctg2"]
    with M.Simulator("tellseq", seed=1) as sim:
        with pytest.raises(M.MimickError):
            sim.add_contig("y" * 161, b"ACGT" * 1000, 1)
