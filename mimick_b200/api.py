"""Host-side mirror of the reference's Julia API on top of the C ABI.

`mimick(...)` keeps the two methods and the keyword surface of
mimick/MimickLinkedReads.jl/src/Main.jl:24 (single sample, N haplotype FASTA files) and :88 (one
FASTA + VCF, every sample), including their defaults (singletons=0.70, attempts=25) and output
naming (`<prefix>.R{1,2}.fq`, `<outdir>/<sample>.R{1,2}.fq`).  Differences, all deliberate:

* output is uncompressed FASTQ text (`.fq`); BGZF compression is outside this stage;
* `seed < 0` draws a seed from the OS, which `Simulator.seed` reports (the reference leaves the
  task-local RNG unseeded);
* extra keywords select the device, the shard of a multi-GPU run and the wrap mode.

`Simulator` exposes the stages (load → plan → place → emit) for callers that keep the FASTQ in
device memory.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from . import _lib
from .dist import samples_for_rank
from ._lib import Chunk, MimickError, MolRecord, Params, PlaceInfo, PlanInfo, SchemaInfo

VARIATE_DTYPE = np.dtype([("tag", "<u4"), ("aux", "<u4"), ("bits", "<u8")])
MOL_DTYPE = np.dtype([("group", "<i8"), ("slot", "<i4"), ("schema", "<i4"), ("start", "<i8"), ("molsize", "<i8"),
                      ("n_frags", "<i8"), ("first_read", "<i8"), ("barcode_index", "<i8")])


def interperet_format(fmt: str):
    """src/Common.jl:56-74 → (barcode_format, out_format) as enum ints."""
    L = _lib.load()
    bc, fq = C.c_int32(), C.c_int32()
    rc = L.mmk_interpret_format(fmt.encode(), C.byref(bc), C.byref(fq))
    if rc:
        raise MimickError(rc, f"Barcode format {fmt} is not recognized as one of: haplotagging, stlfr, tellseq, tenx")
    return bc.value, fq.value


class Vcf:
    """One-pass VCF reader handle (mmk_vcf_*)."""

    def __init__(self, path: str):
        L = _lib.load()
        self._h = C.c_void_p()
        err = C.create_string_buffer(512)
        rc = L.mmk_vcf_open(os.fspath(path).encode(), C.byref(self._h), err, len(err))
        if rc:
            raise MimickError(rc, err.value.decode())
        self.samples = [L.mmk_vcf_sample_name(self._h, i).decode() for i in range(L.mmk_vcf_n_samples(self._h))]
        self.contigs = [L.mmk_vcf_contig_name(self._h, i).decode() for i in range(L.mmk_vcf_n_contigs(self._h))]
        self.ploidy = L.mmk_vcf_ploidy(self._h)
        self.n_records = L.mmk_vcf_n_records(self._h)

    def close(self):
        if self._h:
            _lib.load().mmk_vcf_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Simulator:
    """A simulation context bound to one CUDA device (mmk_ctx)."""

    def __init__(self, format="standard:haplotagging", *, coverage=30, n_molecules=2, mol_coverage=0.2,
                 mol_length=80000, insert_length=500, insert_stdev=50, read_length=(150, 150), singletons=0.70,
                 circular=False, attempts=25, seed=-1, mask=False, device=0, strict_wrap=False,
                 force_generic=False, max_chunk_bytes=0, plan_batch_groups=0, error_rate=(0.0, 0.0), indel_error=(0.0, 0.0)):
        """`error_rate` = substitution probability at the first and the last cycle of a read (linear in between),
        `indel_error` = insertion and deletion probability per cycle: the optional sequencing-error model (not in the
        reference, whose reads are verbatim copies with constant quality; all zero = off, the parity default)."""
        L = _lib.load()
        self._L = L
        bc, fq = interperet_format(format)
        read_length = list(read_length)
        if len(read_length) == 1:  # "Using it for both R1 and R2" (Main.jl:30-32, without its vector-of-vectors bug)
            read_length = read_length * 2
        p = Params()
        L.mmk_params_default(C.byref(p))
        p.device = int(device)
        p.fq_format, p.bc_format = fq, bc
        p.coverage = float(coverage)
        p.n_molecules = int(n_molecules)
        p.attempts = int(attempts)
        p.mol_coverage = float(mol_coverage)
        p.mol_length = int(mol_length)
        p.insert_length = int(insert_length)
        p.insert_stdev = int(insert_stdev)
        p.length_r1, p.length_r2 = int(read_length[0]), int(read_length[1])
        p.singletons = float(singletons)
        p.circular = int(bool(circular))
        p.mask = int(bool(mask))
        p.seed = int(seed)
        p.strict_wrap = int(bool(strict_wrap))
        p.force_generic = int(bool(force_generic))
        p.max_chunk_bytes = int(max_chunk_bytes)
        p.plan_batch_groups = int(plan_batch_groups)
        p.err_sub_start, p.err_sub_end = float(error_rate[0]), float(error_rate[1])
        p.err_ins, p.err_del = float(indel_error[0]), float(indel_error[1])
        self.params = p
        self._h = C.c_void_p()
        rc = L.mmk_create(C.byref(p), C.byref(self._h))
        if rc:
            msg = L.mmk_last_error(self._h).decode() if self._h else "mmk_create failed"
            if self._h:
                L.mmk_destroy(self._h)
                self._h = C.c_void_p()
            raise MimickError(rc, msg)
        self.plan_info = None
        self.place_info = None

    # ---- plumbing -------------------------------------------------------------------------
    def _ck(self, rc):
        if rc:
            raise MimickError(rc, self._L.mmk_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None):
            self._L.mmk_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def seed(self) -> int:
        return int(self._L.mmk_seed(self._h))

    @property
    def cuda_stream(self) -> int:
        return int(self._L.mmk_cuda_stream(self._h) or 0)

    # ---- inputs ---------------------------------------------------------------------------
    def add_contig(self, name: str, bases, haplotype: int = 1):
        """`bases`: bytes / bytearray / numpy uint8 array / (address, length) tuple of host memory."""
        if isinstance(bases, tuple):
            addr, n = bases
        else:
            arr = bases if isinstance(bases, np.ndarray) else np.frombuffer(bases, dtype=np.uint8)
            arr = np.ascontiguousarray(arr, dtype=np.uint8)
            addr, n = arr.ctypes.data, arr.size
        self._ck(self._L.mmk_add_contig(self._h, name.encode(), C.c_void_p(addr), n, int(haplotype)))

    def clear_contigs(self):
        self._ck(self._L.mmk_clear_contigs(self._h))

    def load_fasta(self, path, haplotype: int = 1):
        self._ck(self._L.mmk_load_fasta(self._h, os.fspath(path).encode(), int(haplotype)))

    def apply_vcf_sample(self, vcf: Vcf, sample_index: int):
        self._ck(self._L.mmk_apply_vcf_sample(self._h, vcf._h, int(sample_index)))

    def set_sample_variants(self, sample_stream, ploidy, vcf_contigs, contig, pos, type, ref_len, alts, hap):
        """Raw SoA entry point; `alts` is a list of bytes."""
        vc = np.ascontiguousarray(vcf_contigs, dtype=np.int32)
        contig = np.ascontiguousarray(contig, dtype=np.int32)
        pos = np.ascontiguousarray(pos, dtype=np.int64)
        typ = np.ascontiguousarray(type, dtype=np.uint8)
        ref_len = np.ascontiguousarray(ref_len, dtype=np.int32)
        hap = np.ascontiguousarray(hap, dtype=np.uint8)
        alt_len = np.array([len(a) for a in alts], dtype=np.int32)
        alt_off = np.zeros(len(alts), dtype=np.int64)
        if len(alts):
            alt_off[1:] = np.cumsum(alt_len[:-1])
        blob = b"".join(alts)
        self._ck(self._L.mmk_set_sample_variants(
            self._h, int(sample_stream), int(ploidy), vc.ctypes.data, len(vc), len(pos), contig.ctypes.data,
            pos.ctypes.data, typ.ctypes.data, ref_len.ctypes.data, alt_off.ctypes.data, alt_len.ctypes.data,
            hap.ctypes.data, blob, len(blob)))

    def set_sample_variants_raw(self, sample_stream, ploidy, vcf_contigs, v, blob: bytes):
        """Same, from ready-made arrays: `v` holds contig (i32), pos (i64), type (u8), ref_len (i32), alt_off (i64), alt_len (i32),
        hap (u8); `blob` the ALT bytes the offsets point into."""
        vc = np.ascontiguousarray(vcf_contigs, dtype=np.int32)
        self._ck(self._L.mmk_set_sample_variants(
            self._h, int(sample_stream), int(ploidy), vc.ctypes.data, len(vc), len(v["pos"]), v["contig"].ctypes.data,
            v["pos"].ctypes.data, v["type"].ctypes.data, v["ref_len"].ctypes.data, v["alt_off"].ctypes.data,
            v["alt_len"].ctypes.data, v["hap"].ctypes.data, blob, len(blob)))

    def schemas(self):
        out = []
        for i in range(self._L.mmk_n_schemas(self._h)):
            s = SchemaInfo()
            self._ck(self._L.mmk_schema_get(self._h, i, C.byref(s)))
            name = C.create_string_buffer(256)     # the info struct keeps 63 characters of the name; this is all of it
            self._ck(self._L.mmk_schema_name(self._h, i, name, len(name)))
            out.append(dict(name=name.value.decode(), haplotype=s.haplotype, contig=s.contig, length=s.length,
                            reads_required=s.reads_required, reads_current=s.reads_current))
        return out

    def schema_bases(self, i: int) -> bytes:
        s = SchemaInfo()
        self._ck(self._L.mmk_schema_get(self._h, i, C.byref(s)))
        # length may change once the schema set is finalised; ask again afterwards
        buf = C.create_string_buffer(max(int(s.length), 1) + 64)
        self._ck(self._L.mmk_schema_bases(self._h, i, buf, len(buf)))
        self._ck(self._L.mmk_schema_get(self._h, i, C.byref(s)))
        return buf.raw[: s.length]

    # ---- stages ---------------------------------------------------------------------------
    def plan(self) -> PlanInfo:
        info = PlanInfo()
        self._ck(self._L.mmk_plan(self._h, C.byref(info)))
        self.plan_info = info
        return info

    def set_shard(self, rank: int, world: int):
        self._ck(self._L.mmk_set_shard(self._h, int(rank), int(world)))

    def place(self) -> PlaceInfo:
        info = PlaceInfo()
        self._ck(self._L.mmk_place(self._h, C.byref(info)))
        self.place_info = info
        return info

    def emit_chunk(self) -> Chunk:
        ch = Chunk()
        self._ck(self._L.mmk_emit_chunk(self._h, C.byref(ch)))
        return ch

    def emit_chunk_bgzf(self) -> Chunk:
        """Next chunk as BGZF members (device pointers and sizes describe the coded bytes)."""
        ch = Chunk()
        self._ck(self._L.mmk_emit_chunk_bgzf(self._h, C.byref(ch)))
        return ch

    def bgzf_bound(self) -> int:
        n = C.c_uint64()
        self._ck(self._L.mmk_bgzf_bound(self._h, C.byref(n)))
        return int(n.value)

    def rewind(self):
        self._ck(self._L.mmk_rewind(self._h))

    def copy_chunk(self, ch: Chunk):
        r1 = np.empty(ch.r1_bytes, dtype=np.uint8)
        r2 = np.empty(ch.r2_bytes, dtype=np.uint8)
        self._ck(self._L.mmk_copy_chunk(self._h, C.byref(ch), r1.ctypes.data, r2.ctypes.data))
        return r1, r2

    def emit_to_host(self, r1=None, r2=None):
        """Whole shard into host memory. `r1`/`r2`: optional (address, capacity) of caller buffers
        (e.g. pinned); otherwise numpy arrays are allocated and returned."""
        pi = self.place_info
        if r1 is None:
            a1 = np.empty(pi.r1_bytes, dtype=np.uint8)
            a2 = np.empty(pi.r2_bytes, dtype=np.uint8)
            self._ck(self._L.mmk_emit_to_host(self._h, a1.ctypes.data, a1.size, a2.ctypes.data, a2.size))
            return a1, a2
        self._ck(self._L.mmk_emit_to_host(self._h, C.c_void_p(r1[0]), r1[1], C.c_void_p(r2[0]), r2[1]))
        return None

    def emit_all(self) -> float:
        """Device-resident emission of the whole shard into the output ring; returns kernel ms."""
        ms = C.c_float()
        self._ck(self._L.mmk_emit_all(self._h, C.byref(ms)))
        return float(ms.value)

    def chunk_layout(self):
        sb, n = C.c_uint64(), C.c_int32()
        self._ck(self._L.mmk_chunk_layout(self._h, C.byref(sb), C.byref(n)))
        return int(sb.value), int(n.value)

    def emit_stream(self, r1_slots, r2_slots, slot_cap, consumer=None, bgzf=False):
        """Whole shard through two host slots per read (addresses; pinned memory recommended).
        `consumer(chunk, r1_addr, r2_addr)` is called per chunk once its bytes are in host memory.
        `bgzf=True` delivers BGZF members instead of text (slot_cap >= bgzf_bound())."""
        a1 = (C.c_void_p * 2)(*r1_slots)
        a2 = (C.c_void_p * 2)(*r2_slots)
        cb = None
        if consumer is not None:
            def _cb(user, ch, p1, p2):
                try:
                    consumer(ch.contents, p1, p2)
                    return 0
                except Exception:  # never unwind through C
                    import traceback
                    traceback.print_exc()
                    return 1
            cb = _lib.CHUNK_CB(_cb)
        fn = self._L.mmk_emit_stream_bgzf if bgzf else self._L.mmk_emit_stream
        self._ck(fn(self._h, a1, a2, int(slot_cap), C.cast(cb, C.c_void_p) if cb else None, None))

    def emit_bgzf_to_host(self):
        """Whole shard as BGZF members (no end-of-file member); returns (R1, R2) bytes."""
        cap = self.bgzf_bound()
        slots = [np.empty(cap, dtype=np.uint8) for _ in range(4)]
        parts = ([], [])

        def consume(ch, p1, p2):
            parts[0].append(C.string_at(p1, ch.r1_bytes))
            parts[1].append(C.string_at(p2, ch.r2_bytes))
        self.emit_stream([slots[0].ctypes.data, slots[1].ctypes.data], [slots[2].ctypes.data, slots[3].ctypes.data],
                         cap, consume, bgzf=True)
        return b"".join(parts[0]), b"".join(parts[1])

    def run(self):
        """plan → place → emit everything to host; returns (R1 bytes, R2 bytes)."""
        self.plan()
        self.place()
        a1, a2 = self.emit_to_host()
        return a1.tobytes(), a2.tobytes()

    # ---- introspection --------------------------------------------------------------------
    def molecules(self, first=0, n=None) -> np.ndarray:
        if n is None:
            n = self.plan_info.n_molecules - first
        out = np.zeros(n, dtype=MOL_DTYPE)
        assert MOL_DTYPE.itemsize == C.sizeof(MolRecord)
        self._ck(self._L.mmk_dump_molecules(self._h, first, n, out.ctypes.data))
        return out

    def inserts(self, first=0, n=None):
        if n is None:
            n = self.place_info.n_pairs - first
        st = np.zeros(n, dtype=np.int64)
        en = np.zeros(n, dtype=np.int64)
        self._ck(self._L.mmk_dump_inserts(self._h, first, n, st.ctypes.data, en.ctypes.data))
        return st, en

    def variates(self) -> np.ndarray:
        n = C.c_uint64()
        self._ck(self._L.mmk_dump_variates(self._h, None, 0, C.byref(n)))
        out = np.zeros(n.value, dtype=VARIATE_DTYPE)
        self._ck(self._L.mmk_dump_variates(self._h, out.ctypes.data, out.size, C.byref(n)))
        return out

    def mask_variates(self) -> np.ndarray:
        n = C.c_uint64()
        self._ck(self._L.mmk_dump_mask_variates(self._h, None, 0, C.byref(n)))
        out = np.zeros(n.value, dtype=VARIATE_DTYPE)
        if n.value:
            self._ck(self._L.mmk_dump_mask_variates(self._h, out.ctypes.data, out.size, C.byref(n)))
        return out

    def kat_philox(self, ctr, key):
        c = np.array(ctr, dtype=np.uint32)
        k = np.array(key, dtype=np.uint32)
        o = np.zeros(4, dtype=np.uint32)
        self._ck(self._L.mmk_kat_philox(self._h, c.ctypes.data, k.ctypes.data, o.ctypes.data))
        return o

    def kat_insert_len(self, bits: np.ndarray):
        """Insert lengths the placement kernels derive from raw 64-bit draws; returns (len, took_exact_path)."""
        bits = np.ascontiguousarray(bits, dtype=np.uint64)
        ln = np.zeros(bits.size, dtype=np.uint32)
        slow = np.zeros(bits.size, dtype=np.uint8)
        self._ck(self._L.mmk_kat_insert_len(self._h, bits.ctypes.data, ln.ctypes.data, slow.ctypes.data, bits.size))
        return ln, slow

    def kat_math(self, kind: int, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float64)
        o = np.zeros_like(x)
        self._ck(self._L.mmk_kat_math(self._h, kind, x.ctypes.data, o.ctypes.data, x.size))
        return o


# ------------------------------------------------------------------------------------------
# the reference's public function
# ------------------------------------------------------------------------------------------
_SIM_KW = ("coverage", "n_molecules", "mol_coverage", "mol_length", "insert_length", "insert_stdev", "read_length",
           "singletons", "circular", "attempts", "seed", "mask", "device", "strict_wrap", "max_chunk_bytes", "error_rate", "indel_error")


@dataclass
class SampleOutput:
    name: str
    r1_path: str
    r2_path: str
    n_pairs: int
    n_molecules: int
    r1_bytes: int
    r2_bytes: int
    seed: int


def bgzf_eof() -> bytes:
    buf = C.create_string_buffer(28)
    n = _lib.load().mmk_bgzf_eof(buf, 28)
    assert n == 28
    return buf.raw


def _write_shard(sim: Simulator, r1_path: str, r2_path: str, mode: str, bgzf: bool = False, eof: bool = True):
    """mmk_emit_files: the library streams the placed shard through its own pinned host ring into the two files (file writes of
    chunk k overlap the generation and copy-out of chunk k+1).  `bgzf`: BGZF members coded on the device; `eof` appends the
    end-of-file member."""
    n = (C.c_uint64 * 2)()
    sim._ck(sim._L.mmk_emit_files(sim._h, os.fspath(r1_path).encode(), os.fspath(r2_path).encode(), int(bool(bgzf)), int(bool(eof)),
                                  int("a" in mode), n))
    return int(n[0]), int(n[1])


# device chunk of the file-writing path: the pinned host ring is 4 x this (text) or ~4 x this / 5 (BGZF); 8 GiB chunks are a bench setting
FILE_CHUNK_BYTES = 256 << 20


def mimick(*args, prefix: str = "simulated/SIM", outdir: str = "simulated", quiet: bool = False,
           rank: int = 0, world: int = 1, compress: bool = True, **kw):
    """mimick(fasta::Vector{String}, format; ...)            — src/Main.jl:24
    mimick(fasta::String, vcf::String, format; outdir, ...) — src/Main.jl:88

    `rank`/`world` shard one run over several GPUs: every rank plans the same schedule and
    writes its own contiguous part to `<...>.R1.fq.part<rank>` (concatenate in rank order), or —
    in VCF mode — takes samples `rank, rank+world, ...`.
    `compress` (default, as the reference): files are BGZF `*.fq.gz`, coded on the device; the parts of a
    sharded run are whole BGZF members and only the last rank's part carries the end-of-file member.
    `compress=False` writes plain `*.fq`.
    """
    unknown = set(kw) - set(_SIM_KW)
    if unknown:
        raise TypeError(f"mimick() got unexpected keyword arguments {sorted(unknown)}")
    if len(args) == 2:
        fasta, fmt = args
        vcf = None
        if isinstance(fasta, (str, os.PathLike)):
            fasta = [fasta]
    elif len(args) == 3:
        fasta, vcf, fmt = args
    else:
        raise TypeError("mimick(fasta_list, format, ...) or mimick(fasta, vcf, format, ...)")
    L = _lib.load()
    outs = []
    kw.setdefault("max_chunk_bytes", FILE_CHUNK_BYTES)
    if vcf is None:
        with Simulator(fmt, **kw) as sim:
            for hap, path in enumerate(fasta, start=1):
                sim.load_fasta(path, hap)
            d = os.path.dirname(prefix)
            if d:
                os.makedirs(d, exist_ok=True)
            plan = sim.plan()
            sim.set_shard(rank, world)
            place = sim.place()
            suffix = "" if world == 1 else f".part{rank}"
            ext = "fq.gz" if compress else "fq"
            r1p, r2p = f"{prefix}.R1.{ext}{suffix}", f"{prefix}.R2.{ext}{suffix}"
            _write_shard(sim, r1p, r2p, "wb", bgzf=compress, eof=rank == world - 1)
            outs.append(SampleOutput("", r1p, r2p, place.n_pairs, place.n_molecules, place.r1_bytes, place.r2_bytes, sim.seed))
        return outs
    rc = L.mmk_check_variant_path(os.fspath(vcf).encode())
    if rc:
        raise MimickError(rc, "BCF file type is not yet supported")
    v = Vcf(vcf)
    os.makedirs(outdir, exist_ok=True)
    with Simulator(fmt, **kw) as sim:
        sim.load_fasta(fasta, 1)  # master schema once, masking included (Main.jl:92)
        for si in samples_for_rank(len(v.samples), rank, world):  # Threads.@threads over samples → samples over GPUs
            sim.apply_vcf_sample(v, si)
            plan = sim.plan()
            sim.set_shard(0, 1)
            place = sim.place()
            ext = "fq.gz" if compress else "fq"
            r1p = os.path.join(outdir, f"{v.samples[si]}.R1.{ext}")
            r2p = os.path.join(outdir, f"{v.samples[si]}.R2.{ext}")
            _write_shard(sim, r1p, r2p, "wb", bgzf=compress)
            outs.append(SampleOutput(v.samples[si], r1p, r2p, place.n_pairs, place.n_molecules, place.r1_bytes,
                                     place.r2_bytes, sim.seed))
    v.close()
    return outs
