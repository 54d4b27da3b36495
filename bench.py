#!/usr/bin/env python
"""Benchmark of the simulation hot path (BASELINE.json: read pairs/s and FASTQ GB/s).

    python bench.py --gpus N --steps K --warmup W                # CUDA arm
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU arm (oracle on the host cores)

A "step" is one full pass of the hot path over the workload: barcode-group planning, read-pair
placement and FASTQ emission (north_star subsystems 2-5) with the packed haplotypes already
resident in HBM.  `value` is whole-job read pairs/s for that device-resident pass; `e2e` is the
same job through the C ABI with HOST buffers: FASTA symbols uploaded from pinned memory, FASTQ
text copied back to a pinned host ring, every step.

Default workload = BASELINE.json configs[4] (the configuration the metric is quoted on): 3 Gbp
synthetic diploid, haplotagging, 30x, 10 % singletons.  One process per GPU; with N > 1 the
molecule range is sharded across ranks (total work fixed → "strong" scaling) and the only
collective is an all-gather of per-rank byte counts / read ids (torch.distributed, NCCL).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# GRCh38 primary-assembly chromosome lengths, scaled to the requested total
GRCH38 = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717,
          133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285,
          58617616, 64444167, 46709983, 50818468, 156040895, 57227415]
CHR_NAMES = [f"chr{i}" for i in range(1, 23)] + ["chrX", "chrY"]

WORKLOADS = {
    # name: (description, contig lengths, n_haplotypes, format, sim kwargs)
    "c5": ("3 Gbp synthetic human-scale diploid (24 contigs, GRCh38 length ratios), haplotagging, 30x, 10% singletons",
           None, 2, "haplotagging", dict(coverage=30, singletons=0.1, mol_length=80000)),
    "c3": ("250 Mbp synthetic chr1-sized diploid, 10X, 30x, molecule length 50 kbp",
           [250_000_000], 2, "10x", dict(coverage=30, singletons=0.0, mol_length=50000)),
    "c2": ("5 Mbp synthetic circular prokaryote genome, stLFR, 1 haplotype, 50x (strict wrap)",
           [5_000_000], 1, "stlfr", dict(coverage=50, singletons=0.0, mol_length=80000, circular=True, strict_wrap=True)),
    # BASELINE configs[3]: one FASTA + one VCF, every sample an independent simulation (src/Main.jl:104-130); samples go
    # round-robin to ranks.  nhap = 1 here means "one master FASTA"; each sample builds two haplotypes from it.
    "c4": ("100 Mbp synthetic FASTA + VCF with 8 samples (1e-3 SNPs, 1e-4 indels of 1-20 bp, genotypes Bernoulli(0.3) per haplotype), "
           "TELLseq, 20x per sample, sharded by sample",
           [100_000_000], 1, "tellseq", dict(coverage=20, singletons=0.0, mol_length=80000)),
}
C4_SAMPLES = 8
COMMON_SIM = dict(n_molecules=2, mol_coverage=0.2, insert_length=500, insert_stdev=50, read_length=(150, 150),
                  attempts=300)


def contig_lengths(name, scale):
    desc, lens, nhap, fmt, kw = WORKLOADS[name]
    if lens is None:
        tot = sum(GRCH38)
        lens = [int(round(x * 3.0e9 / tot)) for x in GRCH38]
    lens = [max(20000, int(x * scale)) for x in lens]
    names = CHR_NAMES[: len(lens)] if len(lens) > 1 else ["chr1"]
    return names, lens


def make_genome(name, scale, device, pin):
    """i.i.d. uniform ACGT haplotype 1; haplotype 2 = haplotype 1 with i.i.d. SNPs at 1e-3.
    Returns [(contig name, haplotype index, uint8 host tensor)] in the reference's schema order
    (FASTA 1 contigs, then FASTA 2 contigs)."""
    import torch
    desc, _, nhap, fmt, kw = WORKLOADS[name]
    names, lens = contig_lengths(name, scale)
    g = torch.Generator(device=device)
    g.manual_seed(1005)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    haps = [[], []]
    for nm, L in zip(names, lens):
        code = torch.randint(0, 4, (L,), dtype=torch.uint8, device=device, generator=g)
        h1 = lut[code.long()] if L < (1 << 28) else torch.cat([lut[c.long()] for c in code.split(1 << 27)])
        haps[0].append((nm, h1))
        if nhap == 2:
            snp = torch.rand((L,), device=device, generator=g) < 1e-3
            shift = torch.randint(1, 4, (L,), dtype=torch.uint8, device=device, generator=g)
            code2 = torch.where(snp, (code + shift) % 4, code)
            h2 = lut[code2.long()] if L < (1 << 28) else torch.cat([lut[c.long()] for c in code2.split(1 << 27)])
            haps[1].append((nm, h2))
            del snp, shift, code2
        del code
    out = []
    for hi in range(nhap):
        for nm, t in haps[hi]:
            h = torch.empty(t.shape, dtype=torch.uint8, pin_memory=pin)
            h.copy_(t)
            out.append((nm, hi + 1, h))
    del haps
    if device != "cpu":
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
    return out


def make_c4_variants(master, n_samples=C4_SAMPLES, seed=1004, snp_rate=1e-3, indel_rate=1e-4, p_alt=0.3):
    """Synthetic phased multi-sample VCF content for configs[3], as the SoA mmk_set_sample_variants takes (one entry per
    record and haplotype slot with a non-zero allele, file order).  Sites are shared by all samples, >= 48 bases apart;
    indels are VCF-anchored (deletion: REF = anchor + k bases, ALT = anchor; insertion: REF = anchor, ALT = anchor + k
    random bases; k in 1..20).  Returns (per-sample dict of numpy arrays, shared ALT blob)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    n = master.size
    n_snp, n_ind = int(n * snp_rate), int(n * indel_rate)
    n_var = n_snp + n_ind
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    cells = np.sort(rng.choice(n // 64 - 2, size=n_var, replace=False)) + 1
    pos = cells.astype(np.int64) * 64 + rng.integers(0, 16, size=n_var)          # 1-based
    is_indel = np.zeros(n_var, dtype=bool)
    is_indel[rng.choice(n_var, size=n_ind, replace=False)] = True
    k = rng.integers(1, 21, size=n_var)
    is_del = rng.random(n_var) < 0.5
    ref = master[pos - 1]
    code = np.searchsorted(acgt, ref)
    snp_alt = acgt[(code + rng.integers(1, 4, size=n_var)) % 4]
    ref_len = np.where(is_indel & is_del, k + 1, 1).astype(np.int32)
    alt_len = np.where(is_indel & ~is_del, k + 1, 1).astype(np.int32)
    alt_off = np.zeros(n_var, dtype=np.int64)
    alt_off[1:] = np.cumsum(alt_len[:-1])
    blob = acgt[rng.integers(0, 4, size=int(alt_len.sum()))].copy()               # insertion tails; first bytes fixed below
    blob[alt_off] = np.where(is_indel, ref, snp_alt)
    typ = is_indel.astype(np.uint8)
    out = []
    for _ in range(n_samples):
        g = rng.random((n_var, 2)) < p_alt                                        # [site, haplotype slot]
        idx = np.flatnonzero(g.reshape(-1))
        v, h = idx // 2, (idx % 2 + 1).astype(np.uint8)
        out.append(dict(contig=np.zeros(len(v), dtype=np.int32), pos=pos[v].copy(), type=typ[v].copy(), ref_len=ref_len[v].copy(),
                        alt_off=alt_off[v].copy(), alt_len=alt_len[v].copy(), hap=h))
    return out, blob.tobytes()


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md's clocks line).  Sampled in-process
    through NVML every few milliseconds, so that even a 100 ms timed region (8 ranks x few steps) gets samples; the
    `nvidia-smi -lms` loop of the recipe is the fallback when NVML cannot be opened."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, index, pci_bus_id=None, period_s=0.004):
        self.index = index
        self.pci = pci_bus_id
        self.period = period_s
        self.proc = None
        self.lines = []
        self.samples = []          # (sm_mhz, reason bits)
        self.sm_max = None
        self.nv = None
        self.handle = None
        self.run = False
        self.thread = None
        self.how = None

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.handle = nv.nvmlDeviceGetHandleByPciBusId(self.pci.encode() if hasattr(self.pci, "encode") else self.pci) \
                if self.pci else nv.nvmlDeviceGetHandleByIndex(self.index)
            self.sm_max = float(nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM))
            self._sample_nvml(nv)              # proves the calls work before the timed region depends on them
            self.samples.clear()
            self.nv = nv
            self.run = True
            self.how = "nvml"
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nv = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.pci or self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.how = "nvidia-smi"
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _sample_nvml(self, nv):
        sm = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
        try:
            bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
        except Exception:
            bits = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
        self.samples.append((sm, bits))

    def _loop(self):
        while self.run:
            try:
                self._sample_nvml(self.nv)
            except Exception:
                pass
            time.sleep(self.period)

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        sm, mx, reasons = [], [], set()
        if self.nv is not None:
            self.run = False
            self.thread.join(timeout=2)
            for clk, bits in self.samples:
                sm.append(clk)
                for bit, nm in self.REASONS:
                    if bits & bit:
                        reasons.add(nm)
            mx = [self.sm_max]
            try:
                self.nv.nvmlShutdown()
            except Exception:
                pass
        elif self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
            for ln in self.lines:
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for nm, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "how": self.how}
        sm.sort()
        # median over the samples taken under load (upper half of the clock samples)
        load = sm[len(sm) // 2:]
        return {"sm_mhz": load[len(load) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "how": self.how}


def ncu_traffic():
    """DRAM bytes (read + write) of one k_emit launch from the committed `ncu --set full` capture, with the
    launch's algorithmic bytes for comparison (profiles/r02_final_k_emit_w_ncu_full.csv); None if absent."""
    p = os.path.join(ROOT, "profiles", "r02_k_emit_w_ncu_full.csv")
    if not os.path.exists(p):
        return None, None
    vals = {}
    for ln in open(p):
        f = ln.strip().split(",")
        if len(f) >= 3 and f[0].startswith("dram__bytes_"):
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(f[-1], 1)
            vals[f[0]] = float(f[1]) * mult
    if len(vals) < 2:
        return None, None
    return sum(vals.values()), "profiles/r02_final_k_emit_w_ncu_full.csv (one 2 GiB-chunk launch of k_emit_w<2,0>: 4.34 GB of FASTQ + 0.45 GB gathered)"


def peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle on the host cores
# ------------------------------------------------------------------------------------------
def cpu_run(genome, wl, pairs_per_thread, threads, seed=42):
    """Runs `threads` independent simulations of the workload (the reference's own multi-threading
    shape is one task per sample, Main.jl:104), each stopped after `pairs_per_thread` read pairs,
    sharing one copy of the genome. Returns (total pairs, total FASTQ bytes, wall seconds)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    desc, _, nhap, fmt, kw = WORKLOADS[wl]
    okw = dict(COMMON_SIM)
    okw.update({k: v for k, v in kw.items()})
    cov = okw["coverage"]
    r1, r2 = okw["read_length"]
    sims = []
    schemas = list(genome)
    if wl == "c4":   # a sample = two haplotypes of the master; the bounded CPU sample simulates on the unmutated master
        schemas = [(genome[0][0], 1, genome[0][2]), (genome[0][0], 2, genome[0][2])]   # (mutate! is O(len) per indel on the CPU)
    for t in range(threads):
        o = O.Oracle("tenx" if fmt == "10x" else fmt, **okw)
        o.set_philox(seed, t)
        for nm, hap, h in schemas:
            n = h.numel()
            req = int((cov * n / ((r1 + r2) // 2)) / 2)
            o.add_schema_borrowed(nm, hap, h.data_ptr(), n, req)
        sims.append(o)
    res = [None] * threads

    def work(i):
        res[i] = sims[i].run_counting(pairs_per_thread)

    t0 = time.perf_counter()
    ths = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    dt = time.perf_counter() - t0
    for o in sims:
        o.close()
    return sum(r[0] for r in res), sum(r[1] for r in res), dt


METRIC = "read pairs/sec"  # BASELINE.json metric; both arms report it on the same workload


def reference_arm(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    genome = make_genome(args.workload, args.scale, dev, pin=False)
    threads = args.cpu_threads or min(os.cpu_count() or 1, 64)
    if args.workload == "c4":
        threads = min(threads, C4_SAMPLES)   # the reference's threading shape in VCF mode: one task per sample (src/Main.jl:104)
    per = args.cpu_pairs or 600_000
    for _ in range(args.warmup):
        cpu_run(genome, args.workload, max(per // 10, 1000), threads)
    pairs = byts = 0
    secs = 0.0
    for _ in range(args.steps):
        p, b, dt = cpu_run(genome, args.workload, per, threads)
        pairs += p
        byts += b
        secs += dt
    v = pairs / secs
    desc = WORKLOADS[args.workload][0]
    line = {
        "impl": "reference", "metric": METRIC, "value": v,
        "unit": "read pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * secs / max(args.steps, 1), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}" + (f" (scale {args.scale})" if args.scale != 1.0 else "")},
        "fastq_gb_per_s": byts / secs / 1e9,
        "output": "FASTQ text kept in host memory, gzip off (the reference would also BGZF-compress it on the CPU)",
        "cpu_baseline": {"value": v, "unit": "read pairs/s", "cores": threads, "kind": "port",
                         "sample": f"{threads} independent simulations of the workload (one per host thread, the reference's "
                                   f"task-per-sample shape), each stopped after {per} read pairs; Julia is not installed, so "
                                   "this is the C restatement of the reference (oracle/), text kept in memory, gzip off"},
        "e2e": {"value": v, "unit": "read pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------
# CUDA arm
# ------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(index):
    """Pin this process to the CPUs of the NUMA node the GPU hangs off, so that the pinned host buffers it allocates
    (first touch) and the threads that consume them sit next to the GPU's PCIe root.  Only meaningful on multi-node
    hosts: returns what was found and whether anything was bound (the boxes of this pool are single-node VMs whose
    PCI devices report node -1 — nothing to bind, and the report says so instead of a silent null)."""
    info = {"bound": False, "gpu_node": None, "host_nodes": None, "why": ""}
    try:
        import torch
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")]
        info["host_nodes"] = len(nodes)
        p = torch.cuda.get_device_properties(index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bdf}"
        node = int(open(f"{base}/numa_node").read())
        info["gpu_node"] = node
        if len(nodes) < 2 or node < 0:
            info["why"] = "single NUMA node (or the GPU reports no node): nothing to bind"
            return info
        cpus = set()
        for part in open(f"{base}/local_cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            info["bound"] = True
            info["why"] = f"bound to the {len(cpus)} CPUs of node {node}"
        else:
            info["why"] = "no allowed CPU on the GPU's node"
    except Exception as e:  # noqa: BLE001
        info["why"] = f"topology unreadable: {type(e).__name__}"
    return info



def cuda_arm(args):
    import torch
    import torch.distributed as dist
    import mimick_b200 as M
    from mimick_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the CUDA arm has no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    numa = bind_to_gpu_numa_node(local)   # before any pinned allocation: host rings next to the GPU's PCIe root
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(dev))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    desc, _, nhap, fmt, kw = WORKLOADS[args.workload]
    simkw = dict(COMMON_SIM)
    simkw.update(kw)
    genome = make_genome(args.workload, args.scale, dev, pin=True)
    h2d_bytes = sum(h.numel() for _, _, h in genome)
    L = _lib.load()

    sim = M.Simulator(fmt, seed=42, device=local, max_chunk_bytes=args.chunk_mib << 20, **simkw)
    stream = torch.cuda.ExternalStream(sim.cuda_stream, device=dev)

    # Units of work of this rank.  Single-sample workloads: ONE job, this rank's molecule range of the one simulation.
    # c4 (FASTA + VCF): one job per sample of this rank (samples round-robin over ranks, src/Main.jl:104); each job builds the
    # sample's haplotypes from the resident master (north_star subsystem 1), then plans, places and emits the whole sample.
    is_vcf = args.workload == "c4"
    if is_vcf:
        from mimick_b200.dist import samples_for_rank
        variants, alt_blob = make_c4_variants(genome[0][2].numpy())
        jobs = samples_for_rank(C4_SAMPLES, rank, world)
        n_master = genome[0][2].numel()
    else:
        jobs = [None]

    dev_syms = {}

    def load_inputs():
        """Every FASTA symbol from pinned host memory into the context.  One rank: straight H2D per contig.  Several
        ranks: contig i crosses PCIe once, on rank i % world, and is replicated to the other GPUs by an NCCL broadcast
        over NVLink; every rank then hands the device copy to mmk_add_contig.  Returns this rank's H2D bytes."""
        if world == 1:
            for nm, hap, h in genome:
                sim.add_contig(nm, (h.data_ptr(), h.numel()), hap)
            return h2d_bytes
        mine = 0
        for i, (nm, hap, h) in enumerate(genome):   # staging buffers live on the default stream, outside the context's
            if i not in dev_syms:
                dev_syms[i] = torch.empty(h.numel(), dtype=torch.uint8, device=dev)
        works = []
        for i, (nm, hap, h) in enumerate(genome):    # torch's own streams: H2D of this rank's contigs, then the broadcasts
            if i % world == rank:
                dev_syms[i].copy_(h, non_blocking=True)
                mine += h.numel()
            works.append(dist.broadcast(dev_syms[i], src=i % world, async_op=True))
        with torch.cuda.stream(stream):
            for i, (nm, hap, h) in enumerate(genome):
                works[i].wait()          # the context's stream waits for the broadcast, the host does not
                sim.add_contig(nm, (dev_syms[i].data_ptr(), h.numel()), hap)
        return mine

    def exchange(tot):
        """the path's only collective: per-rank (bytes_R1, bytes_R2, n_pairs, first_read_id) over NCCL"""
        if is_vcf:   # samples are separate files: ranks only need each other's totals, nothing chains
            t = torch.tensor(list(tot), dtype=torch.int64, device=dev)
            if world == 1:
                return [tuple(tot)]
            out = torch.empty(world * 4, dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(out, t)
            return [tuple(r) for r in out.cpu().view(world, 4).tolist()]
        lay = M.exchange_counts(M.RankCounts(*tot), device=dev)
        return [(c.r1_bytes, c.r2_bytes, c.n_pairs, c.first_pair) for c in lay.counts]

    state = {}

    def run_jobs(emit):
        """plan + place (+ haplotype construction) of every job of this rank, `emit(place)` for its text; returns this rank's
        (bytes_R1, bytes_R2, n_pairs, first_read_id) and adds the phase times to `state`."""
        tot = [0, 0, 0, 0]
        algo = mols = 0
        for job in jobs:
            if is_vcf:
                t0 = time.perf_counter()
                sim.set_sample_variants_raw(job, 2, [0], variants[job], alt_blob)   # synchronous
                state["hap_ms"] = state.get("hap_ms", 0.0) + 1e3 * (time.perf_counter() - t0)
                state["hap_bytes"] = n_master // 4 * 2 + sum(sc["length"] for sc in sim.schemas()) // 4 + len(variants[job]["pos"]) * 30
            plan = sim.plan()
            sim.set_shard(*((0, 1) if is_vcf else (rank, world)))
            place = sim.place()
            emit(place)
            tot[0] += place.r1_bytes
            tot[1] += place.r2_bytes
            tot[2] += place.n_pairs
            tot[3] = 0 if is_vcf else place.first_pair
            algo += place.r1_bytes + place.r2_bytes + place.gather_bytes
            mols += plan.n_molecules
            state["plan_ms"] = state.get("plan_ms", 0.0) + plan.host_ms
            state["place_ms"] = state.get("place_ms", 0.0) + place.host_ms
            state.update(plan=plan, place=place)
        state.update(algo=algo, mols=mols)
        return tot

    def step_resident():
        ms = [0.0]

        def emit(place):
            ms[0] += sim.emit_all()
        tot = run_jobs(emit)
        state["all"] = exchange(tot)
        return ms[0]

    # ---- device-resident pass: `value` ------------------------------------------------------
    load_inputs()
    for _ in range(args.warmup):
        step_resident()
    state["plan_ms"] = state["place_ms"] = state["hap_ms"] = 0.0
    _p = torch.cuda.get_device_properties(local)
    sampler = ClockSampler(local, f"{_p.pci_domain_id:08x}:{_p.pci_bus_id:02x}:{_p.pci_device_id:02x}.0")
    if rank == 0:
        sampler.start()
    barrier()
    l0 = L.mmk_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record(stream)
    emit_ms = 0.0
    for _ in range(args.steps):
        emit_ms += step_resident()
    ev1.record(stream)
    barrier()
    wall = time.perf_counter() - t0
    launches = L.mmk_launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = ev0.elapsed_time(ev1)
    # host-side phase times of the timed steps (the end-to-end passes below reuse `state`)
    phase_resident = {"plan": state["plan_ms"] / args.steps, "place": state["place_ms"] / args.steps, "emit": emit_ms / args.steps}
    tmax = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_per_step = float(tmax.item()) / args.steps
    total_pairs = sum(c[2] for c in state["all"])
    total_bytes = sum(c[0] + c[1] for c in state["all"])
    value = total_pairs / (ms_per_step / 1e3)
    algo_bytes = state["algo"]  # this rank, per step: FASTQ bytes written + packed reference bytes gathered
    hap_ms_per_sample = state["hap_ms"] / args.steps / max(len(jobs), 1) if is_vcf else None
    achieved = algo_bytes * args.steps / (emit_ms / 1e3) / 1e9
    peak, peak_src = peak_hbm()
    traffic, traffic_src = ncu_traffic()

    # ---- end to end through the C ABI with host buffers: `e2e` --------------------------------
    # The reference's product is <prefix>.R{1,2}.fq.gz (BGZF); `e2e` delivers exactly that to host memory: FASTA
    # symbols in from pinned memory, BGZF members out to a pinned host ring, coded on the device.  `e2e_text`
    # is the same with uncompressed text crossing PCIe instead.
    def make_e2e(bgzf):
        slot_bytes = sim.bgzf_bound() if bgzf else sim.chunk_layout()[0]
        if is_vcf:   # every sample goes through the same host ring; their bounds differ by a few members per chunk
            slot_bytes = int(slot_bytes * 1.05) + (1 << 20)
        ring = [torch.empty(slot_bytes, dtype=torch.uint8, pin_memory=True) for _ in range(4)]
        d2h = {"bytes": 0, "step": 0, "h2d": 0}

        def consume(ch, p1, p2):
            d2h["step"] += ch.r1_bytes + ch.r2_bytes

        phases = {"load": 0.0, "haplotypes": 0.0, "plan": 0.0, "place": 0.0, "stream": 0.0}

        def step():
            t0 = time.perf_counter()
            sim.clear_contigs()
            d2h["h2d"] = load_inputs()         # H2D of every FASTA symbol from pinned memory
            t1 = time.perf_counter()
            d2h["step"] = 0
            for k in ("plan_ms", "place_ms", "hap_ms"):
                state[k] = 0.0

            def emit(place):
                sim.emit_stream([ring[0].data_ptr(), ring[1].data_ptr()], [ring[2].data_ptr(), ring[3].data_ptr()], slot_bytes,
                                consume, bgzf=bgzf)
            tot = run_jobs(emit)
            exchange(tot)
            t2 = time.perf_counter()
            d2h["bytes"] = d2h["step"]
            phases["load"] += 1e3 * (t1 - t0)
            phases["plan"] += state["plan_ms"]
            phases["place"] += state["place_ms"]
            phases["haplotypes"] += state["hap_ms"]
            phases["stream"] += 1e3 * (t2 - t1) - state["plan_ms"] - state["place_ms"] - state["hap_ms"]
        return step, d2h, slot_bytes, phases

    def run_e2e(bgzf, steps, warmup):
        step, d2h, slot_bytes, phases = make_e2e(bgzf)
        for _ in range(warmup):
            step()
        for k in phases:
            phases[k] = 0.0
        barrier()
        t0 = time.perf_counter()
        ev0.record(stream)
        for _ in range(steps):
            step()
        ev1.record(stream)
        barrier()
        e_wall = time.perf_counter() - t0
        # the stream's events bracket the region; the copy stream is joined by the synchronisation at the end of
        # every emit_stream call, so the wall clock between the barriers is the conservative figure
        e_ms = max(ev0.elapsed_time(ev1), 1e3 * e_wall)
        tm = torch.tensor([e_ms], dtype=torch.float64, device=dev)
        nb = torch.tensor([d2h["bytes"]], dtype=torch.float64, device=dev)
        nh = torch.tensor([d2h["h2d"]], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            dist.all_reduce(nb, op=dist.ReduceOp.SUM)
            dist.all_reduce(nh, op=dist.ReduceOp.SUM)
        e_step = float(tm.item()) / steps
        return {"value": total_pairs / (e_step / 1e3), "unit": "read pairs/s", "h2d_bytes_per_step": int(nh.item()),
                "input_replication": "one H2D per contig + NCCL broadcast over NVLink" if world > 1 else "none (one GPU)",
                "d2h_bytes_per_step": int(nb.item()), "ms_per_step": e_step, "steps": steps, "warmup": warmup,
                "output": "BGZF members (*.fq.gz as the reference writes), coded on the device" if bgzf else "FASTQ text",
                "fastq_gb_per_s": total_bytes / (e_step / 1e3) / 1e9, "pcie_gb_per_s": nb.item() / world / (e_step / 1e3) / 1e9,
                "compression_ratio": total_bytes / max(nb.item(), 1.0) if bgzf else 1.0,
                "host_ring_bytes": int(4 * slot_bytes),
                "host_phase_ms_per_step": {k: v / steps for k, v in phases.items()}}

    e2e = e2e_text = None
    if args.e2e_steps > 0:
        e2e = run_e2e(True, args.e2e_steps, args.e2e_warmup)
        if args.e2e_text_steps > 0:
            e2e_text = run_e2e(False, args.e2e_text_steps, 1)
    torch.cuda.synchronize()
    dev_syms.clear()       # before the context (and with it the stream these buffers were used on) goes away
    sim.close()

    # ---- CPU baseline beside it (rank 0, N = 1 only) -------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = args.cpu_threads or min(os.cpu_count() or 1, 64)
        if is_vcf:
            threads = min(threads, C4_SAMPLES)
        per = args.cpu_pairs or 4_000_000
        p1, b1, dt1 = cpu_run(genome, args.workload, max(per // 2, 1000), 1)
        p, b, dt = cpu_run(genome, args.workload, per, threads)
        cpu = {"value": p / dt, "unit": "read pairs/s", "cores": threads, "kind": "port",
               "reference_threading_shape": {
                   "value": p1 / dt1, "unit": "read pairs/s", "cores": 1,
                   "sample": f"ONE simulation thread (the reference's single-sample loop, src/Main.jl:44-59, is sequential; its "
                             f"other threads only compress), stopped after {max(per // 2, 1000)} read pairs; {dt1:.1f} s"},
               "sample": f"{threads} independent simulations (one per host thread, the reference's task-per-sample shape), "
                         f"each stopped after {per} read pairs of the same workload; C restatement of the reference "
                         f"(Julia absent), text to memory, gzip off; {dt:.1f} s",
               "fastq_gb_per_s": b / dt / 1e9}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "read pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {desc}" + (f" (scale {args.scale})" if args.scale != 1.0 else ""),
                       "read_pairs": int(total_pairs),
                       "molecules": int(state["mols"]) if is_vcf else int(state["plan"].n_molecules),
                       "fastq_bytes": int(total_bytes),
                       "sharding": (f"{C4_SAMPLES} samples round-robin over {world} rank(s), {len(jobs)} on rank 0; haplotypes built "
                                    "from the resident master inside every step") if is_vcf
                       else f"molecule ranges over {world} rank(s), replicated planning",
                       "l2": "outputs (GBs per step) far exceed the 126 MB L2; no flush needed",
                       "output_ring": f"2 x 2 x {args.chunk_mib} MiB device buffers, reused chunk after chunk"},
            "fastq_gb_per_s": total_bytes / (ms_per_step / 1e3) / 1e9,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "kernel": "k_emit_w", "peak_source": peak_src,
                         "algorithmic_bytes_per_step": int(algo_bytes), "kernel_ms_per_step": emit_ms / args.steps,
                         "kernel_share_of_step": (emit_ms / args.steps) / ms_per_step},
            "cpu_baseline": cpu,
            "e2e": e2e,
            "e2e_text": e2e_text,
            "gpu_launches": int(launches),
            "host_numa": numa,
            "clocks": clocks,
            "wall_s_timed_region": wall,
            "plan": {"groups": int(state["plan"].n_groups), "epochs": int(state["plan"].n_epochs),
                     "batches": int(state["plan"].n_batches)},
            "haplotype_construction": None if not is_vcf else {
                "ms_per_sample": hap_ms_per_sample, "ms_per_step_rank0": hap_ms_per_sample * len(jobs),
                "algorithmic_bytes_per_sample": int(state["hap_bytes"]),
                "bytes": "packed master read once per haplotype + packed haplotypes written + 30 B per variant entry (SURVEY 8d)",
                "achieved_gb_per_s": state["hap_bytes"] / (hap_ms_per_sample / 1e3) / 1e9,
                "frac_of_hbm_peak": state["hap_bytes"] / (hap_ms_per_sample / 1e3) / 1e9 / peak,
                "note": "wall time of mmk_set_sample_variants (host piece list + H2D of the variant arrays + device kernels)"},
            "phase_ms_per_step": phase_resident,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS))
    ap.add_argument("--scale", type=float, default=1.0, help="shrink every contig (testing only; the reported config says so)")
    ap.add_argument("--chunk-mib", type=int, default=2048)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-warmup", type=int, default=3)
    ap.add_argument("--e2e-text-steps", type=int, default=1, help="extra e2e pass with uncompressed text over PCIe (0 = skip)")
    ap.add_argument("--cpu-threads", type=int, default=0)
    ap.add_argument("--cpu-pairs", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    return cuda_arm(args)


if __name__ == "__main__":
    sys.exit(main())
