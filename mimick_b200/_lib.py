"""ctypes binding of include/mimick_b200.h (the C ABI). Fails loudly when the CUDA library is
missing: there is no CPU path in this package."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmimick_b200.so")

FMT_STLFR, FMT_HAPLOTAGGING, FMT_TELLSEQ, FMT_TENX, FMT_STANDARD = range(5)
STATUS_NAMES = {
    0: "MMK_OK", 1: "MMK_ERR_INVALID_ARG", 2: "MMK_ERR_BAD_FORMAT", 3: "MMK_ERR_INSERTS_UNFITTABLE",
    4: "MMK_ERR_BARCODES_EXHAUSTED", 5: "MMK_ERR_CONTIG_TOO_AMBIGUOUS", 6: "MMK_ERR_BCF_UNSUPPORTED",
    7: "MMK_ERR_TRUNCATION_BOUNDS", 8: "MMK_ERR_BAD_BASE", 9: "MMK_ERR_IO", 10: "MMK_ERR_VCF",
    11: "MMK_ERR_BOUNDS", 12: "MMK_ERR_UNSUPPORTED", 13: "MMK_ERR_CUDA", 14: "MMK_ERR_NOMEM",
    15: "MMK_ERR_REPLAY", 16: "MMK_ERR_STATE",
}


class MimickError(RuntimeError):
    """Raised for every non-zero status of the C ABI (the reference raises Julia `error(...)`)."""

    def __init__(self, status: int, message: str):
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {message}")
        self.status = status
        self.message = message


class Params(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("device", C.c_int32), ("fq_format", C.c_int32), ("bc_format", C.c_int32),
        ("coverage", C.c_double), ("n_molecules", C.c_int32), ("attempts", C.c_int32), ("mol_coverage", C.c_double),
        ("mol_length", C.c_int64), ("insert_length", C.c_int32), ("insert_stdev", C.c_int32),
        ("length_r1", C.c_int32), ("length_r2", C.c_int32), ("singletons", C.c_double), ("circular", C.c_int32),
        ("mask", C.c_int32), ("seed", C.c_int64), ("strict_wrap", C.c_int32), ("force_generic", C.c_int32),
        ("max_chunk_bytes", C.c_uint64), ("plan_batch_groups", C.c_int64),
        ("err_sub_start", C.c_double), ("err_sub_end", C.c_double), ("err_ins", C.c_double), ("err_del", C.c_double),
    ]


class SchemaInfo(C.Structure):
    _fields_ = [("length", C.c_int64), ("reads_required", C.c_int64), ("reads_current", C.c_int64),
                ("haplotype", C.c_int32), ("contig", C.c_int32), ("name", C.c_char * 64)]


class PlanInfo(C.Structure):
    _fields_ = [("n_groups", C.c_int64), ("n_molecules", C.c_int64), ("n_pairs", C.c_int64),
                ("n_epochs", C.c_int64), ("n_batches", C.c_int64), ("host_ms", C.c_double)]


class PlaceInfo(C.Structure):
    _fields_ = [("first_molecule", C.c_int64), ("n_molecules", C.c_int64), ("first_pair", C.c_int64),
                ("n_pairs", C.c_int64), ("r1_bytes", C.c_uint64), ("r2_bytes", C.c_uint64), ("gather_bytes", C.c_uint64),
                ("host_ms", C.c_double), ("phase_ms", C.c_double * 4)]


class Chunk(C.Structure):
    _fields_ = [("r1_dev", C.c_void_p), ("r2_dev", C.c_void_p), ("r1_bytes", C.c_uint64), ("r2_bytes", C.c_uint64),
                ("r1_offset", C.c_uint64), ("r2_offset", C.c_uint64), ("first_pair", C.c_int64), ("n_pairs", C.c_int64),
                ("done", C.c_int32), ("slot", C.c_int32), ("kernel_ms", C.c_float)]


class MolRecord(C.Structure):
    _fields_ = [("group", C.c_int64), ("slot", C.c_int32), ("schema", C.c_int32), ("start", C.c_int64),
                ("molsize", C.c_int64), ("n_frags", C.c_int64), ("first_read", C.c_int64), ("barcode_index", C.c_int64)]


# every symbol include/mimick_b200.h declares (tests check the .so exports exactly these)
EXPORTS = [
    "mmk_params_default", "mmk_interpret_format", "mmk_check_variant_path", "mmk_abi_version", "mmk_strerror",
    "mmk_last_error", "mmk_create", "mmk_destroy", "mmk_seed", "mmk_cuda_stream", "mmk_add_contig", "mmk_load_fasta",
    "mmk_n_contigs", "mmk_set_sample_variants", "mmk_vcf_open", "mmk_vcf_close", "mmk_vcf_n_samples",
    "mmk_vcf_sample_name", "mmk_vcf_ploidy", "mmk_vcf_n_contigs", "mmk_vcf_contig_name", "mmk_vcf_n_records",
    "mmk_apply_vcf_sample", "mmk_n_schemas", "mmk_schema_get", "mmk_schema_name", "mmk_schema_bases", "mmk_plan", "mmk_set_shard",
    "mmk_place", "mmk_emit_chunk", "mmk_rewind", "mmk_copy_chunk", "mmk_emit_to_host", "mmk_dump_molecules",
    "mmk_dump_inserts", "mmk_dump_variates", "mmk_dump_mask_variates", "mmk_kat_philox", "mmk_kat_math",
    "mmk_emit_all", "mmk_emit_stream", "mmk_chunk_layout", "mmk_launch_count", "mmk_clear_contigs",
    "mmk_bgzf_bound", "mmk_emit_chunk_bgzf", "mmk_emit_stream_bgzf", "mmk_bgzf_eof", "mmk_bgzf_table",
    "mmk_kat_insert_len", "mmk_insert_table", "mmk_emit_files",
]

CHUNK_CB = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.POINTER(Chunk), C.c_void_p, C.c_void_p)

_LIB = None


def load():
    """dlopen the in-tree library and declare prototypes. Raises if it was not built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing. Build it with `python mimick_b200/build.py` (needs nvcc). "
            "mimick_b200 has no CPU fallback."
        )
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, u32, u64, cp = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_char_p
    P = C.POINTER

    def sig(name, res, *args):
        f = getattr(L, name)
        f.restype = res
        f.argtypes = list(args)

    sig("mmk_params_default", None, P(Params))
    sig("mmk_interpret_format", i32, cp, P(i32), P(i32))
    sig("mmk_check_variant_path", i32, cp)
    sig("mmk_abi_version", i32)
    sig("mmk_strerror", cp, i32)
    sig("mmk_last_error", cp, vp)
    sig("mmk_create", i32, P(Params), P(vp))
    sig("mmk_destroy", None, vp)
    sig("mmk_seed", u64, vp)
    sig("mmk_cuda_stream", vp, vp)
    sig("mmk_add_contig", i32, vp, cp, vp, i64, i32)
    sig("mmk_load_fasta", i32, vp, cp, i32)
    sig("mmk_n_contigs", i32, vp)
    sig("mmk_clear_contigs", i32, vp)
    sig("mmk_set_sample_variants", i32, vp, u32, i32, vp, i32, i64, vp, vp, vp, vp, vp, vp, vp, vp, i64)
    sig("mmk_vcf_open", i32, cp, P(vp), C.c_char_p, C.c_size_t)
    sig("mmk_vcf_close", None, vp)
    sig("mmk_vcf_n_samples", i32, vp)
    sig("mmk_vcf_sample_name", cp, vp, i32)
    sig("mmk_vcf_ploidy", i32, vp)
    sig("mmk_vcf_n_contigs", i32, vp)
    sig("mmk_vcf_contig_name", cp, vp, i32)
    sig("mmk_vcf_n_records", i64, vp)
    sig("mmk_apply_vcf_sample", i32, vp, vp, i32)
    sig("mmk_n_schemas", i32, vp)
    sig("mmk_schema_get", i32, vp, i32, P(SchemaInfo))
    sig("mmk_schema_name", i32, vp, i32, vp, i64)
    sig("mmk_schema_bases", i32, vp, i32, vp, i64)
    sig("mmk_plan", i32, vp, P(PlanInfo))
    sig("mmk_set_shard", i32, vp, i32, i32)
    sig("mmk_place", i32, vp, P(PlaceInfo))
    sig("mmk_emit_chunk", i32, vp, P(Chunk))
    sig("mmk_rewind", i32, vp)
    sig("mmk_copy_chunk", i32, vp, P(Chunk), vp, vp)
    sig("mmk_emit_to_host", i32, vp, vp, u64, vp, u64)
    sig("mmk_emit_all", i32, vp, P(C.c_float))
    sig("mmk_emit_stream", i32, vp, P(vp), P(vp), u64, vp, vp)
    sig("mmk_chunk_layout", i32, vp, P(u64), P(i32))
    sig("mmk_bgzf_bound", i32, vp, P(u64))
    sig("mmk_emit_chunk_bgzf", i32, vp, P(Chunk))
    sig("mmk_emit_stream_bgzf", i32, vp, P(vp), P(vp), u64, vp, vp)
    sig("mmk_bgzf_eof", i32, vp, C.c_uint32)
    sig("mmk_bgzf_table", i32, vp, vp, vp, C.c_uint32, P(C.c_uint32))
    sig("mmk_launch_count", u64)
    sig("mmk_dump_molecules", i32, vp, i64, i64, vp)
    sig("mmk_dump_inserts", i32, vp, i64, i64, vp, vp)
    sig("mmk_dump_variates", i32, vp, vp, u64, P(u64))
    sig("mmk_dump_mask_variates", i32, vp, vp, u64, P(u64))
    sig("mmk_kat_philox", i32, vp, vp, vp, vp)
    sig("mmk_kat_math", i32, vp, i32, vp, vp, i64)
    sig("mmk_kat_insert_len", i32, vp, vp, vp, vp, i64)
    sig("mmk_emit_files", i32, vp, cp, cp, i32, i32, i32, vp)
    sig("mmk_insert_table", i32, i32, i32, P(u32), P(u32), P(u32), P(u32), P(u32), P(u32), vp, u32)
    _LIB = L
    return L
