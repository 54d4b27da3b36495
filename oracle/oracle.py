"""ctypes front end of the CPU checker (oracle/mimick_oracle.c) — TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  It restates, on top of the C restatement, the host-side steps of the reference:

* interperet_format            mimick/MimickLinkedReads.jl/src/Common.jl:56-74
* FASTA reading / setup_schema mimick/MimickLinkedReads.jl/src/ProcessFASTA.jl:45-95
* VCF reading / variants       mimick/MimickLinkedReads.jl/src/ProcessVariants.jl:10-20,42-46,61-106
* build_sample_schema          mimick/MimickLinkedReads.jl/src/ProcessVariants.jl:157-185
* the two drivers              mimick/MimickLinkedReads.jl/src/Main.jl:24-62, 88-131

The text parsers here are deliberately independent of the product's C++ parsers so that the
parity tests also cross-check those.
"""
from __future__ import annotations

import ctypes as C
import gzip
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FMT = {"stlfr": 0, "haplotagging": 1, "tellseq": 2, "tenx": 3, "standard": 4}

# purposes (include/mmk_rng.h)
P_MASK, P_NMOL, P_TARGET, P_MOLSIZE, P_SINGLETON, P_NFRAGS, P_INSERT, P_START, P_GAP = range(1, 10)

VARIATE_DTYPE = np.dtype([("tag", "<u4"), ("aux", "<u4"), ("bits", "<u8")])


class OrcParams(C.Structure):
    _fields_ = [
        ("insert_length", C.c_int32),
        ("insert_stdev", C.c_int32),
        ("length_r1", C.c_int32),
        ("length_r2", C.c_int32),
        ("n_molecules", C.c_int32),
        ("attempts", C.c_int32),
        ("mol_length", C.c_int64),
        ("mol_coverage", C.c_double),
        ("singletons", C.c_double),
        ("circular", C.c_int32),
        ("fq_format", C.c_int32),
        ("bc_format", C.c_int32),
        ("strict_wrap", C.c_int32),
        ("pad", C.c_int32),
        ("err_sub_start", C.c_double), ("err_sub_end", C.c_double), ("err_ins", C.c_double), ("err_del", C.c_double),
    ]


class OrcMol(C.Structure):
    _fields_ = [
        ("group", C.c_int64),
        ("slot", C.c_int32),
        ("schema", C.c_int32),
        ("start", C.c_int64),
        ("molsize", C.c_int64),
        ("n_frags", C.c_int64),
        ("first_read", C.c_int64),
        ("barcode_index", C.c_int64),
    ]


CHECKPOINT_DTYPE = np.dtype([("pairs", "<i8"), ("bytes1", "<u8"), ("bytes2", "<u8"), ("crc1", "<u4"), ("crc2", "<u4"),
                             ("pad", "<u4", (2,))])

MOL_DTYPE = np.dtype(
    [("group", "<i8"), ("slot", "<i4"), ("schema", "<i4"), ("start", "<i8"), ("molsize", "<i8"),
     ("n_frags", "<i8"), ("first_read", "<i8"), ("barcode_index", "<i8")]
)


class OracleError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"oracle error {code}: {msg}")
        self.code = code


def build(force: bool = False) -> str:
    """Compile oracle/libmimick_oracle.so with the committed Makefile."""
    so = os.path.join(_HERE, "libmimick_oracle.so")
    src = os.path.join(_HERE, "mimick_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "mmk_rng.h")
    stale = (not os.path.exists(so)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(so) for p in (src, hdr)
    )
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "libmimick_oracle.so"], check=True, capture_output=True)
    return so


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    L = C.CDLL(build())
    vp, i64, i32, u64, u32, dbl, sz = C.c_void_p, C.c_int64, C.c_int32, C.c_uint64, C.c_uint32, C.c_double, C.c_size_t
    L.orc_create.restype = vp
    L.orc_create.argtypes = [C.POINTER(OrcParams)]
    L.orc_destroy.argtypes = [vp]
    L.orc_error.argtypes = [vp]
    L.orc_error.restype = i32
    L.orc_error_message.argtypes = [vp]
    L.orc_error_message.restype = C.c_char_p
    L.orc_add_schema.argtypes = [vp, C.c_char_p, i32, C.c_char_p, i64, i64]
    L.orc_add_schema_borrowed.argtypes = [vp, C.c_char_p, i32, vp, i64, i64]
    L.orc_set_discard_output.argtypes = [vp, i32]
    L.orc_total_pairs.argtypes = [vp]
    L.orc_total_pairs.restype = i64
    L.orc_total_bytes.argtypes = [vp]
    L.orc_total_bytes.restype = u64
    L.orc_set_philox.argtypes = [vp, u64, u32]
    L.orc_set_replay.argtypes = [vp, vp, sz]
    L.orc_set_record.argtypes = [vp, i32]
    L.orc_set_max_pairs.argtypes = [vp, i64]
    L.orc_set_digest.argtypes = [vp, i64]
    L.orc_checkpoints.argtypes = [vp, C.POINTER(sz)]
    L.orc_checkpoints.restype = vp
    L.orc_normalize_bases.argtypes = [vp, i64]
    L.orc_normalize_bases.restype = i64
    L.orc_mask_n.argtypes = [vp, vp, i64, u32, u32]
    L.orc_mask_n.restype = i64
    L.orc_reads_required.argtypes = [vp, i64, dbl, i32, i32]
    L.orc_reads_required.restype = i64
    L.orc_mutate.argtypes = [vp, i64, i64, vp, vp, vp, vp, C.POINTER(vp), C.POINTER(i64)]
    L.orc_mutate.restype = i32
    L.orc_free.argtypes = [vp]
    L.orc_barcode.argtypes = [i32, u64, C.c_char_p]
    L.orc_barcode.restype = i32
    L.orc_run.argtypes = [vp]
    L.orc_run.restype = i32
    for name in ("orc_r1", "orc_r2"):
        f = getattr(L, name)
        f.argtypes = [vp, C.POINTER(sz)]
        f.restype = vp
    for name in ("orc_n_groups", "orc_n_molecules", "orc_n_pairs"):
        f = getattr(L, name)
        f.argtypes = [vp]
        f.restype = i64
    L.orc_molecules.argtypes = [vp]
    L.orc_molecules.restype = vp
    L.orc_insert_starts.argtypes = [vp]
    L.orc_insert_starts.restype = vp
    L.orc_insert_stops.argtypes = [vp]
    L.orc_insert_stops.restype = vp
    L.orc_recorded.argtypes = [vp, C.POINTER(sz)]
    L.orc_recorded.restype = vp
    L.orc_replay_consumed.argtypes = [vp]
    L.orc_replay_consumed.restype = sz
    L.orc_schema_reads.argtypes = [vp, i32, C.POINTER(i64)]
    L.orc_schema_reads.restype = i64
    L.orc_lowerbound.argtypes = [vp]
    L.orc_lowerbound.restype = dbl
    L.orc_philox.argtypes = [vp, vp, vp]
    L.orc_det_log.argtypes = [dbl]
    L.orc_det_log.restype = dbl
    L.orc_det_exp.argtypes = [dbl]
    L.orc_det_exp.restype = dbl
    L.orc_trunc_exp_info.argtypes = [dbl, dbl, i32, dbl, vp]
    L.orc_sample_trunc_exp.argtypes = [u64, dbl, dbl, i32, dbl, i64, vp]
    L.orc_sample_insert.argtypes = [u64, i32, i32, i64, vp]
    L.orc_insert_len_from_bits.argtypes = [i32, i32, vp, vp, i64]
    _LIB = L
    return L


def insert_len_from_bits(mu: int, sd: int, bits: np.ndarray) -> np.ndarray:
    """trunc(Int, mu + sd * z(bits)) per raw 64-bit draw, 0 where the truncation at 200 rejects it."""
    bits = np.ascontiguousarray(bits, dtype=np.uint64)
    out = np.zeros(bits.size, dtype=np.uint32)
    lib().orc_insert_len_from_bits(int(mu), int(sd), bits.ctypes.data, out.ctypes.data, bits.size)
    return out


# ------------------------------------------------------------------------------------------
# host-side restatements
# ------------------------------------------------------------------------------------------
def interperet_format(fmt: str):
    """Common.jl:56-74 → (barcode_format, out_format). `10x` is normalised to `tenx` by the CLI
    (mimick/__main__.py:73)."""
    if fmt == "10x":
        fmt = "tenx"
    if "standard" in fmt:
        out = "standard"
        if "stlfr" in fmt:
            bc = "stlfr"
        elif "haplotagging" in fmt:
            bc = "haplotagging"
        else:
            bc = "tellseq"
    elif fmt in ("stlfr", "haplotagging", "tellseq", "tenx"):
        out = bc = fmt
    else:
        raise OracleError(2, f"Barcode format {fmt} is not recognized as one of: haplotagging, stlfr, tellseq, tenx")
    return bc, out


def _open_text(path):
    """Common.jl:6-40 safe_read: gzip, BGZF (a multi-member gzip) or plain."""
    with open(path, "rb") as fh:
        magic = fh.read(2)
    if magic == b"\x1f\x8b":
        return gzip.open(path, "rb")
    return open(path, "rb")


def read_fasta(path):
    """FASTX.FASTAReader: identifier = header up to the first whitespace; sequence lines joined."""
    out, name, chunks = [], None, []
    with _open_text(path) as fh:
        for line in fh:
            line = line.rstrip(b"\r\n")
            if line.startswith(b">"):
                if name is not None:
                    out.append((name, b"".join(chunks)))
                name = line[1:].split()[0].decode() if len(line) > 1 else ""
                chunks = []
            elif name is not None:
                chunks.append(line.strip())
    if name is not None:
        out.append((name, b"".join(chunks)))
    return out


@dataclass
class Vcf:
    contigs: list
    samples: list
    ploidy: int
    # per record: chrom, pos, ref, alts(list), gts(list of str per sample)
    records: list = field(default_factory=list)


def read_vcf(path) -> Vcf:
    contigs, samples, records = [], [], []
    with _open_text(path) as fh:
        for raw in fh:
            line = raw.rstrip(b"\r\n").decode()
            if line.startswith("##"):
                if line.startswith("##contig=<"):
                    body = line[len("##contig=<"):-1]
                    for kv in body.split(","):
                        if kv.startswith("ID="):
                            contigs.append(kv[3:])
                continue
            if line.startswith("#CHROM"):
                samples = line.split("\t")[9:]
                continue
            if not line:
                continue
            f = line.split("\t")
            fmt_keys = f[8].split(":")
            gi = fmt_keys.index("GT")
            gts = [s.split(":")[gi] for s in f[9:]]
            records.append((f[0], int(f[1]), f[3], f[4].split(","), gts))
    # get_ploidy_and_contigs: ploidy from sample 1 of the first record (ProcessVariants.jl:10-20)
    ploidy = len(records[0][4][0].replace("|", "/").split("/")) if records else 0
    return Vcf(contigs, samples, ploidy, records)


def sample_variants(vcf: Vcf, sample_number: int):
    """get_sample_variants, ProcessVariants.jl:80-106 → {"contig_h": [(pos, type, ref, alt), ...]}
    in creation order (contigs in header order × haplotype 1..ploidy)."""
    out = {}
    for c in vcf.contigs:
        for h in range(1, vcf.ploidy + 1):
            out[f"{c}_{h}"] = []
    for chrom, pos, ref, alts, gts in vcf.records:
        vtype = 0 if (len(ref) == 1 and all(len(a) == 1 for a in alts)) else 1  # vartype :61-69
        gt = gts[sample_number]
        alleles = [int(x) for x in gt.replace("|", "/").split("/")]  # "." raises like parse(UInt8, ".")
        for i, a in enumerate(alleles, start=1):
            if a != 0:
                out[f"{chrom}_{i}"].append((pos, vtype, ref, alts[a - 1]))
    return out


# ------------------------------------------------------------------------------------------
# the simulator object
# ------------------------------------------------------------------------------------------
@dataclass
class SimResult:
    r1: bytes
    r2: bytes
    molecules: np.ndarray
    insert_start: np.ndarray
    insert_stop: np.ndarray
    n_groups: int
    reads_current: list
    reads_required: list
    variates: np.ndarray | None = None
    checkpoints: np.ndarray | None = None   # digest mode: running CRC-32 of R1 / R2 at group boundaries


class Oracle:
    def __init__(self, fmt="standard:haplotagging", coverage=30, n_molecules=2, mol_coverage=0.2,
                 mol_length=80000, insert_length=500, insert_stdev=50, read_length=(150, 150),
                 singletons=0.0, circular=False, attempts=25, strict_wrap=False, error_rate=(0.0, 0.0), indel_error=(0.0, 0.0)):
        L = lib()
        bc, out = interperet_format(fmt)
        self.coverage = float(coverage)
        self.read_length = tuple(read_length)
        self.par = OrcParams(int(insert_length), int(insert_stdev), int(read_length[0]), int(read_length[1]),
                             int(n_molecules), int(attempts), int(mol_length), float(mol_coverage),
                             float(singletons), int(bool(circular)), FMT[out], FMT[bc], int(bool(strict_wrap)), 0,
                             float(error_rate[0]), float(error_rate[1]), float(indel_error[0]), float(indel_error[1]))
        self.h = L.orc_create(C.byref(self.par))
        self._keep = []
        self.n_schema = 0
        self._check()

    def _check(self):
        L = lib()
        code = L.orc_error(self.h)
        if code:
            raise OracleError(code, L.orc_error_message(self.h).decode())

    def close(self):
        if self.h:
            lib().orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_philox(self, seed, sample=0):
        lib().orc_set_philox(self.h, int(seed), int(sample))

    def set_replay(self, stream: np.ndarray):
        stream = np.ascontiguousarray(stream, dtype=VARIATE_DTYPE)
        self._keep.append(stream)
        lib().orc_set_replay(self.h, stream.ctypes.data, len(stream))

    def set_record(self, on=True):
        lib().orc_set_record(self.h, int(on))

    def set_max_pairs(self, n):
        lib().orc_set_max_pairs(self.h, int(n))

    def set_digest(self, window_pairs):
        """Digest mode: the text is folded into running CRC-32s (checkpoint every `window_pairs` read pairs, at the
        next barcode-group boundary) and discarded; molecules and inserts are kept.  For full-size parity runs."""
        lib().orc_set_digest(self.h, int(window_pairs))

    # ---- schema construction ------------------------------------------------------------
    def prepare_contig(self, seq: bytes, mask: bool, contig_index: int, hap_index: int):
        """process_fasta body, ProcessFASTA.jl:51-62: normalise, optional mask_N!, reads_required."""
        L = lib()
        buf = C.create_string_buffer(seq, len(seq))
        bad = L.orc_normalize_bases(C.addressof(buf), len(seq))
        if bad >= 0:
            raise OracleError(8, f"invalid base at {bad}")
        if mask:
            L.orc_mask_n(self.h, C.addressof(buf), len(seq), contig_index, hap_index)
            self._check()
        req = L.orc_reads_required(C.addressof(buf), len(seq), self.coverage, self.read_length[0], self.read_length[1])
        if req < 0:
            raise OracleError(-req, "Contigs must have at least 95% of the requested read length sum as unambiguous (non-N) bases")
        return buf.raw, int(req)

    def add_schema(self, chrom: str, haplotype: int, seq: bytes, reads_required: int):
        lib().orc_add_schema(self.h, chrom.encode(), int(haplotype), seq, len(seq), int(reads_required))
        self.n_schema += 1

    def add_schema_borrowed(self, chrom: str, haplotype: int, address: int, length: int, reads_required: int):
        """No copy: `address` must point to upper-case IUPAC bytes that outlive the run."""
        lib().orc_add_schema_borrowed(self.h, chrom.encode(), int(haplotype), C.c_void_p(address), int(length), int(reads_required))
        self.n_schema += 1

    def run_counting(self, max_pairs: int):
        """Bench mode: run until `max_pairs`, discarding the text; returns (pairs, fastq_bytes)."""
        L = lib()
        L.orc_set_discard_output(self.h, 1)
        L.orc_set_max_pairs(self.h, int(max_pairs))
        rc = L.orc_run(self.h)
        if rc:
            raise OracleError(rc, L.orc_error_message(self.h).decode())
        return int(L.orc_total_pairs(self.h)), int(L.orc_total_bytes(self.h))

    @staticmethod
    def mutate(master: bytes, variants):
        """build_sample_schema inner body, ProcessVariants.jl:170-176."""
        L = lib()
        n = len(variants)
        pos = np.array([v[0] for v in variants], dtype=np.int64)
        typ = np.array([v[1] for v in variants], dtype=np.int32)
        rlen = np.array([len(v[2]) for v in variants], dtype=np.int32)
        alts = [v[3].upper().encode() for v in variants]
        arr = (C.c_char_p * max(n, 1))(*alts) if n else (C.c_char_p * 1)()
        out = C.c_void_p()
        out_len = C.c_int64()
        rc = L.orc_mutate(master, len(master), n, pos.ctypes.data, typ.ctypes.data, rlen.ctypes.data,
                          C.cast(arr, C.c_void_p), C.byref(out), C.byref(out_len))
        if rc:
            raise OracleError(rc, "mutate failed")
        res = C.string_at(out, out_len.value)
        L.orc_free(out)
        return res

    # ---- run ---------------------------------------------------------------------------------
    def run(self) -> SimResult:
        L = lib()
        rc = L.orc_run(self.h)
        if rc:
            raise OracleError(rc, L.orc_error_message(self.h).decode())
        n = C.c_size_t()
        p = L.orc_r1(self.h, C.byref(n))
        r1 = C.string_at(p, n.value) if n.value else b""
        p = L.orc_r2(self.h, C.byref(n))
        r2 = C.string_at(p, n.value) if n.value else b""
        nm = L.orc_n_molecules(self.h)
        npairs = L.orc_n_pairs(self.h)
        mols = np.frombuffer(C.string_at(L.orc_molecules(self.h), nm * MOL_DTYPE.itemsize), dtype=MOL_DTYPE).copy() if nm else np.zeros(0, MOL_DTYPE)
        st = np.frombuffer(C.string_at(L.orc_insert_starts(self.h), npairs * 8), dtype=np.int64).copy() if npairs else np.zeros(0, np.int64)
        en = np.frombuffer(C.string_at(L.orc_insert_stops(self.h), npairs * 8), dtype=np.int64).copy() if npairs else np.zeros(0, np.int64)
        cur, req = [], []
        for i in range(self.n_schema):
            r = C.c_int64()
            cur.append(int(L.orc_schema_reads(self.h, i, C.byref(r))))
            req.append(int(r.value))
        p = L.orc_recorded(self.h, C.byref(n))
        var = np.frombuffer(C.string_at(p, n.value * 16), dtype=VARIATE_DTYPE).copy() if n.value else None
        p = L.orc_checkpoints(self.h, C.byref(n))
        cps = np.frombuffer(C.string_at(p, n.value * CHECKPOINT_DTYPE.itemsize), dtype=CHECKPOINT_DTYPE).copy() if n.value else None
        return SimResult(r1, r2, mols, st, en, int(L.orc_n_groups(self.h)), cur, req, var, cps)


# ------------------------------------------------------------------------------------------
# the two reference drivers, restated
# ------------------------------------------------------------------------------------------
def _common_kwargs(kw):
    keys = ("coverage", "n_molecules", "mol_coverage", "mol_length", "insert_length", "insert_stdev",
            "read_length", "singletons", "circular", "attempts", "strict_wrap", "error_rate", "indel_error")
    return {k: kw[k] for k in keys if k in kw}


def mimick_single(fasta_files, fmt, seed=0, mask=False, replay=None, record=False, max_pairs=0, **kw) -> SimResult:
    """Main.jl:24-62 — N FASTA files are N haplotypes of one sample. Schemas are kept in
    first-seen order: FASTA argument order, then contig order within the file."""
    o = Oracle(fmt, **_common_kwargs(kw))
    o.set_philox(seed, 0)
    if replay is not None:
        o.set_replay(replay)
    o.set_record(record)
    for hi, path in enumerate(fasta_files, start=1):
        for ci, (name, seq) in enumerate(read_fasta(path)):
            s, req = o.prepare_contig(seq, mask, ci, hi)
            o.add_schema(name, hi, s, req)
    if max_pairs:
        o.set_max_pairs(max_pairs)
    res = o.run()
    o.close()
    return res


def mimick_multi(fasta, vcf_path, fmt, seed=0, mask=False, samples=None, replay=None, record=False, **kw):
    """Main.jl:88-131 — one FASTA + VCF; every sample is an independent simulation with its own
    barcode enumerator. Returns {sample_name: SimResult}. Sample `i` (0-based VCF column order)
    draws from Philox sample-stream `i`."""
    vcf = read_vcf(vcf_path)
    contigs = read_fasta(fasta)
    out = {}
    # master schema once, including masking (Main.jl:92). Masking consumes variates of sample-stream 0.
    om = Oracle(fmt, **_common_kwargs(kw))
    om.set_philox(seed, 0)
    if replay is not None and "__master__" in replay:
        om.set_replay(replay["__master__"])
    master = {}
    for ci, (name, seq) in enumerate(contigs):
        s, req = om.prepare_contig(seq, mask, ci, 1)
        master[name] = (s, req)
    om.close()
    for si, sname in enumerate(vcf.samples):
        if samples is not None and sname not in samples and si not in samples:
            continue
        o = Oracle(fmt, **_common_kwargs(kw))
        o.set_philox(seed, si)
        if replay is not None:
            o.set_replay(replay[sname])
        o.set_record(record)
        variants = sample_variants(vcf, si)
        # build_sample_schema: only contigs named in the VCF header get schemas (:161-162)
        for key, var in variants.items():
            chrom, hap = key.rsplit("_", 1)
            if chrom not in master:
                raise OracleError(10, f"VCF contig {chrom} not in FASTA")
            mseq, req = master[chrom]
            o.add_schema(chrom, int(hap), Oracle.mutate(mseq, var), req)
        out[sname] = o.run()
        o.close()
    return out
