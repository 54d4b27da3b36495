"""Byte parity with the oracle at BASELINE.json's FULL sizes (configs[1..4]).

The text of these runs (0.6-40 GB) does not fit beside the oracle's in host memory, so the oracle runs in digest
mode (oracle/mimick_oracle.c: a running CRC-32 of R1 and of R2, sampled at the first barcode-group boundary after
every `window` read pairs, the text discarded; molecules and insert breakpoints kept), and the CUDA path is streamed
through the C ABI (mmk_emit_stream, host ring) with the same running CRC-32 taken at the same byte offsets.  Equal
CRCs at every checkpoint = equal bytes window by window; molecules (group, slot, schema, start, size, inserts,
barcode) and all insert breakpoints are compared element-wise.  Reference loop: src/Main.jl:44-59, :113-126.

The oracle is sequential (0.6 M read pairs/s): the oracle runs of all selected tests of this module start side by side
on host threads when the test session begins and the module runs last (tests/conftest.py), so they finish beside the
other GPU tests instead of in front of them.
"""
import os
import sys
import threading

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from helpers import assert_digest_parity, oracle_digest

pytestmark = pytest.mark.gpu


class Background:
    """the oracle run beside the GPU work (ctypes releases the GIL)"""

    def __init__(self, fn, *a, **kw):
        self.res = self.exc = None

        def work():
            try:
                self.res = fn(*a, **kw)
            except BaseException as e:  # noqa: BLE001 - re-raised in join()
                self.exc = e
        self.t = threading.Thread(target=work, daemon=True)
        self.t.start()

    def join(self):
        self.t.join()
        if self.exc is not None:
            raise self.exc
        return self.res


def bench_workload(wl, **override):
    import bench
    desc, _, nhap, fmt, kw = bench.WORKLOADS[wl]
    simkw = dict(bench.COMMON_SIM)
    simkw.update(kw)
    simkw.update(override)
    genome = bench.make_genome(wl, 1.0, "cuda:0", pin=False)
    return fmt, simkw, genome


def start_single(wl, window, max_pairs=0, genome=None, **override):
    """genome of the bench workload + its oracle run started in the background"""
    fmt, simkw, g = bench_workload(wl, **override)
    if genome is not None:
        g = genome
    schemas = [(nm, hap, h.numpy()) for nm, hap, h in g]
    ofmt = "tenx" if fmt == "10x" else fmt
    bg = Background(oracle_digest, ofmt, schemas, 42, 0, window, max_pairs=max_pairs, **simkw)
    return dict(wl=wl, fmt=fmt, simkw=simkw, genome=g, bg=bg, prefix=bool(max_pairs))


def finish_single(job, chunk_mib, shard=None):
    import mimick_b200 as M
    with M.Simulator(job["fmt"], seed=42, device=0, max_chunk_bytes=chunk_mib << 20, **job["simkw"]) as sim:
        for nm, hap, h in job["genome"]:
            sim.add_contig(nm, (h.data_ptr(), h.numel()), hap)
        plan = sim.plan()
        if shard:
            sim.set_shard(*shard)
        place = sim.place()
        orc = job["bg"].join()
        n = assert_digest_parity(sim, plan, place, orc, job["wl"], prefix=job["prefix"])
        got = {s["name"] + "_" + str(s["haplotype"]): s["reads_current"] for s in sim.schemas()}
    return plan, place, orc, n, got


C4_SIMKW = dict(coverage=20, n_molecules=2, mol_coverage=0.2, mol_length=80000, insert_length=500, insert_stdev=50,
                read_length=(150, 150), singletons=0.0, attempts=300)


def start_c4_sample():
    import oracle as O
    rng = np.random.default_rng(1004)
    master, per_hap, soa = make_c4_sample(rng, 100_000_000, 110_000, 10_000)
    mbytes = master.tobytes()

    def oracle_side():
        haps = [np.frombuffer(O.Oracle.mutate(mbytes, per_hap[h]), dtype=np.uint8) for h in (1, 2)]
        req = int(O.lib().orc_reads_required(master.ctypes.data, master.size, 20.0, 150, 150))   # inherited from the master
        res = oracle_digest("tellseq", [("chr1", 1, haps[0], req), ("chr1", 2, haps[1], req)], 42, 3, 500_000, **C4_SIMKW)
        return haps, res
    return dict(master=master, mbytes=mbytes, soa=soa, bg=Background(oracle_side))


JOBS = {}


def start_jobs(selected):
    """Starts, longest first, the oracle run of every selected test of this module; tests join their own."""
    jobs = {}
    if "test_c3_full_size_oracle_parity" in selected:
        jobs["c3"] = start_single("c3", 1_000_000)
    if "test_c4_sample_full_size_oracle_parity" in selected:
        jobs["c4"] = start_c4_sample()
    if "test_c5_prefix_oracle_parity" in selected:
        jobs["c5"] = start_single("c5", 500_000, max_pairs=5_000_000)
    c2 = None
    for strict in (True, False):
        if f"test_c2_full_size_oracle_parity[{strict}]" in selected:
            jobs["c2", strict] = start_single("c2", 50_000, genome=c2, strict_wrap=strict)
            c2 = jobs["c2", strict]["genome"]
    return jobs


@pytest.fixture(scope="module")
def oracles(request):
    """conftest.py starts the jobs with the session and runs this module last, so the oracle works beside the other
    GPU tests; run on its own, the module starts them here."""
    if not JOBS:
        JOBS.update(start_jobs({it.name for it in request.session.items if str(it.path) == str(request.path)}))
    return JOBS


@pytest.mark.parametrize("strict", [True, False])
def test_c2_full_size_oracle_parity(oracles, strict):
    """configs[1]: 5 Mbp circular, stLFR, 1 haplotype, 50x — 833 333+ pairs; strict (modular) wrap as benchmarked, and
    the reference's literal circular_index (reads past the end become hap[1 : stop % n])."""
    plan, place, orc, n, _ = finish_single(oracles["c2", strict], 64)
    assert plan.n_pairs >= 833_333 and n >= 16
    if not strict:
        assert place.r1_bytes + place.r2_bytes > 1.2 * 833_333 * 700   # the giant D5 reads are really there


def test_c3_full_size_oracle_parity(oracles):
    """configs[2]: 250 Mbp diploid, 10X, 30x, molecule length 50 kbp — 50 M pairs, 39.8 GB of FASTQ, 2 pruning epochs."""
    plan, place, orc, n, cur = finish_single(oracles["c3"], 1024)
    assert 50_000_000 <= plan.n_pairs < 50_010_000 and n >= 50 and plan.n_epochs == 2
    assert [int(c) for c in orc.reads_current] == [cur["chr1_1"], cur["chr1_2"]]


def test_c5_prefix_oracle_parity(oracles):
    """configs[4]: 3 Gbp diploid, haplotagging, 30x, 10 % singletons — the first 5 M+ read pairs of the run (the oracle
    stopped at the first group boundary after 5 M pairs) against the head of shard 0 of 64 of the full 600 M-pair plan."""
    plan, place, orc, n, _ = finish_single(oracles["c5"], 512, shard=(0, 64))
    assert plan.n_pairs >= 600_000_000 and int(orc.checkpoints["pairs"][-1]) >= 5_000_000 and n >= 10


def make_c4_sample(rng, n, n_var, n_indel, p_alt=0.3):
    """One sample of configs[3]: SNPs and indels of 1-20 bp (VCF-anchored), genotypes Bernoulli(p_alt) per haplotype.
    Returns (master, per-haplotype variant lists for the oracle [(pos, type, ref, alt)], SoA for the C ABI)."""
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    master = acgt[rng.integers(0, 4, size=n)]
    cells = np.sort(rng.choice(n // 64 - 2, size=n_var, replace=False)) + 1
    pos = cells.astype(np.int64) * 64 + rng.integers(0, 16, size=n_var)          # 1-based, >= 48 apart
    is_indel = np.zeros(n_var, dtype=bool)
    is_indel[rng.choice(n_var, size=n_indel, replace=False)] = True
    soa = dict(pos=[], type=[], ref_len=[], alt=[], hap=[])
    per_hap = {1: [], 2: []}
    for i in range(n_var):
        p = int(pos[i])
        ref = master[p - 1]
        if is_indel[i]:
            k = int(rng.integers(1, 21))
            if rng.random() < 0.5:   # deletion: REF = k + 1 bases, ALT = the anchor base
                refs, alt = master[p - 1:p + k].tobytes(), bytes([ref])
            else:                    # insertion: REF = anchor, ALT = anchor + k bases
                refs, alt = bytes([ref]), bytes([ref]) + acgt[rng.integers(0, 4, size=k)].tobytes()
            typ = 1
        else:
            refs, alt, typ = bytes([ref]), bytes([acgt[(np.searchsorted(acgt, ref) + int(rng.integers(1, 4))) % 4]]), 0
        for hap in (1, 2):
            if rng.random() < p_alt:
                soa["pos"].append(p), soa["type"].append(typ), soa["ref_len"].append(len(refs)), soa["alt"].append(alt), soa["hap"].append(hap)
                per_hap[hap].append((p, typ, refs.decode(), alt.decode()))
    return master, per_hap, soa


def test_c4_sample_full_size_oracle_parity(oracles):
    """configs[3], one of its 8 samples at full size: 100 Mbp master, 100 k SNPs + 10 k indels, TELLseq, 20x — 13.3 M
    pairs.  Haplotypes: the oracle's mutate (all SNPs, then spliceinto! in reverse file order, ProcessVariants.jl:143-155)
    against the device's segmented gather; then the simulation byte for byte (sample stream 3)."""
    import mimick_b200 as M
    job = oracles["c4"]
    soa = job["soa"]
    with M.Simulator("tellseq", seed=42, device=0, max_chunk_bytes=512 << 20, **C4_SIMKW) as sim:
        sim.add_contig("chr1", job["mbytes"], 1)
        sim.set_sample_variants(3, 2, [0], np.zeros(len(soa["pos"]), dtype=np.int32), soa["pos"], soa["type"], soa["ref_len"],
                                soa["alt"], soa["hap"])
        plan = sim.plan()
        place = sim.place()
        haps, orc = job["bg"].join()
        for i in (0, 1):
            assert sim.schema_bases(i) == haps[i].tobytes(), f"haplotype {i + 1} differs from the oracle's mutate"
        n = assert_digest_parity(sim, plan, place, orc, "c4 sample 3")
    assert 13_333_332 <= plan.n_pairs < 13_343_332 and n >= 26
