/* mimick_b200.h — C ABI of the B200-native simulation core for pdimens/mimick.
 *
 * This is the drop-in boundary for the simulation stage of the reference: everything
 * `mimick(...)` does between "FASTA/VCF are readable" and "FASTQ text exists"
 * (mimick/MimickLinkedReads.jl/src/Main.jl:24-62 and :88-131).  The reference has no FFI; a
 * maintainer would bind these entry points from Julia with `ccall` (INTEGRATION.md shows the
 * stub; julia/MimickB200.jl ships it).  Plain pointers and sizes only; no C++ or torch types.
 *
 * Conventions
 *   - every function returns an int32 status (MMK_OK or MMK_ERR_*); nothing throws or exits;
 *     mmk_last_error(ctx) gives the message (the reference's error(...) sites are mapped to
 *     distinct codes, cited below);
 *   - the caller owns host inputs (copied during the call); the library owns all device memory
 *     until mmk_destroy; device output pointers stay valid until the next mmk_emit_chunk that
 *     reuses the same ring slot, or mmk_reset / mmk_destroy;
 *   - a context is bound to one CUDA device and one stream and is not re-entrant; distinct
 *     contexts are independent (one per GPU / per sample shard);
 *   - the library REQUIRES a CUDA device: there is no CPU fallback.
 */
#ifndef MIMICK_B200_H
#define MIMICK_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMK_ABI_VERSION 1

/* ---- status codes ---------------------------------------------------------------------- */
enum mmk_status {
    MMK_OK = 0,
    MMK_ERR_INVALID_ARG = 1,
    MMK_ERR_BAD_FORMAT = 2,           /* src/Common.jl:71  "Barcode format ... is not recognized" */
    MMK_ERR_INSERTS_UNFITTABLE = 3,   /* src/Breakpoints.jl:214-216 "FAILED TO FIND PRACTICAL READ SIZES" */
    MMK_ERR_BARCODES_EXHAUSTED = 4,   /* src/Barcodes.jl:75 "no more barcodes available" */
    MMK_ERR_CONTIG_TOO_AMBIGUOUS = 5, /* src/ProcessFASTA.jl:58-60 "<95% ... unambiguous" */
    MMK_ERR_BCF_UNSUPPORTED = 6,      /* src/Main.jl:95-97 "BCF file type is not yet supported" */
    MMK_ERR_TRUNCATION_BOUNDS = 7,    /* truncated(Exponential, 2.0, molsize/insert) with lower > upper, src/Breakpoints.jl:191 */
    MMK_ERR_BAD_BASE = 8,             /* symbol LongDNA{4} cannot encode (FASTA or ALT allele) */
    MMK_ERR_IO = 9,                   /* src/Common.jl:7-9 "... does not exist" and read failures */
    MMK_ERR_VCF = 10,                 /* missing GT, contig not in FASTA/header, malformed record */
    MMK_ERR_BOUNDS = 11,              /* a read window the reference would index out of bounds */
    MMK_ERR_UNSUPPORTED = 12,         /* parameter region not implemented (stated in DESIGN.md) */
    MMK_ERR_CUDA = 13,
    MMK_ERR_NOMEM = 14,
    MMK_ERR_REPLAY = 15,
    MMK_ERR_STATE = 16                /* calls out of order */
};

/* output / barcode styles: src/Common.jl:56-74, src/Barcodes.jl, src/FormatFASTQ.jl */
enum mmk_format {
    MMK_FMT_STLFR = 0,
    MMK_FMT_HAPLOTAGGING = 1,
    MMK_FMT_TELLSEQ = 2,
    MMK_FMT_TENX = 3,
    MMK_FMT_STANDARD = 4 /* output style only */
};

/* ---- parameters: the keyword surface of mimick(...) (src/Main.jl:24, :88) ----------------- */
typedef struct mmk_params {
    uint32_t struct_size;   /* sizeof(mmk_params), for ABI evolution */
    int32_t device;         /* CUDA device ordinal */
    int32_t fq_format;      /* enum mmk_format: record style (interperet_format → out_format) */
    int32_t bc_format;      /* enum mmk_format: barcode enumerator */
    double coverage;        /* coverage      (-g)  per (contig, haplotype) read target */
    int32_t n_molecules;    /* n_molecules   (-N)  >0 mean of truncated Exponential, <0 fixed */
    int32_t attempts;       /* attempts      (-A)  redraws of the insert-size vector */
    double mol_coverage;    /* mol_coverage  (-C) */
    int64_t mol_length;     /* mol_length    (-L) */
    int32_t insert_length;  /* insert_length (-i) */
    int32_t insert_stdev;   /* insert_stdev  (-d) */
    int32_t length_r1;      /* read_length[1] (-l) */
    int32_t length_r2;      /* read_length[2] */
    double singletons;      /* singletons    (-S) */
    int32_t circular;       /* circular      (-c) */
    int32_t mask;           /* mask          (-m) */
    int64_t seed;           /* seed          (-s)  <0: take a seed from the OS (reported by mmk_seed) */
    int32_t strict_wrap;    /* 0: reproduce the reference's circular_index literally;
                               1: modular wrap, every read has its nominal length */
    int32_t force_generic;  /* 1: emit every record through the general byte-wise path (testing) */
    uint64_t max_chunk_bytes; /* capacity of each device output buffer (R1 and R2 each); 0 = default */
    int64_t plan_batch_groups; /* barcode groups planned per speculative batch; 0 = default */
    /* Sequencing-error / quality model.  The reference has none (docs/algorithm.md:11: reads are verbatim copies,
     * quality constant 'I', src/FormatFASTQ.jl:14); with all four rates 0 (the default) the output is the reference's.
     * DESIGN.md "Error model" specifies the draws; every read of nominal length gets, per cycle c of its S cycles:
     *   substitution with probability err_sub_start + (err_sub_end - err_sub_start) c / (S - 1),
     *   insertion / deletion with probabilities err_ins / err_del, and a Phred quality drawn around the cycle's error rate. */
    double err_sub_start, err_sub_end, err_ins, err_del;
} mmk_params;

typedef struct mmk_ctx mmk_ctx;

/* Julia defaults of src/Main.jl:24 (singletons 0.70, attempts 25, format "standard:haplotagging"
 * as in the CLI, device 0, seed -1). */
void mmk_params_default(mmk_params *p);
/* interperet_format, src/Common.jl:56-74; also accepts the CLI spelling "10x" (mimick/__main__.py:73). */
int32_t mmk_interpret_format(const char *fmt, int32_t *bc_format, int32_t *fq_format);
/* endswith(lowercase(vcf), "bcf") → MMK_ERR_BCF_UNSUPPORTED, src/Main.jl:95-97 */
int32_t mmk_check_variant_path(const char *path);

int32_t mmk_abi_version(void);
const char *mmk_strerror(int32_t status);
const char *mmk_last_error(const mmk_ctx *ctx);

/* SimParams constructor, src/Structs.jl:54-70; binds the CUDA device and creates the stream. */
int32_t mmk_create(const mmk_params *params, mmk_ctx **out);
void mmk_destroy(mmk_ctx *ctx);
uint64_t mmk_seed(const mmk_ctx *ctx);
/* cudaStream_t the context enqueues on (for callers that time with their own events). */
void *mmk_cuda_stream(const mmk_ctx *ctx);

/* ---- inputs ------------------------------------------------------------------------------
 * process_fasta / setup_schema, src/ProcessFASTA.jl:45-95.  `bases` is ASCII IUPAC (either
 * case), `len` symbols, in host memory or in device memory of the context's GPU (unified
 * addressing tells them apart; a multi-GPU caller can upload each contig once and replicate it
 * over NVLink); `haplotype` is the 1-based FASTA file index.  The contig
 * is uploaded, validated, optionally N-masked (mask_N!, :29-36) and counted on the device;
 * reads_required follows :47-61.  Contigs become schemas in the order they are added. */
int32_t mmk_add_contig(mmk_ctx *ctx, const char *name, const char *bases, int64_t len, int32_t haplotype);
/* FASTX reader equivalent: gzip / BGZF / plain (safe_read, src/Common.jl:6-40); one
 * mmk_add_contig per record. */
int32_t mmk_load_fasta(mmk_ctx *ctx, const char *path, int32_t haplotype);
int32_t mmk_n_contigs(const mmk_ctx *ctx);
/* forgets every contig, schema and plan (device buffers are kept for reuse): a new run on the same context */
int32_t mmk_clear_contigs(mmk_ctx *ctx);

/* Multi-sample mode (src/Main.jl:88-131): contigs added so far are the master; each call
 * replaces the schema set by this sample's haplotypes (build_sample_schema,
 * src/ProcessVariants.jl:157-185: all SNPs, then indels in reverse file order) and resets the
 * barcode counter and coverage trackers.  One entry per (record, haplotype slot with a
 * non-zero allele) — the `Mutation`s of get_sample_variants, :80-106 — in file order.
 *   vcf_contigs[n_vcf_contigs]  master contig index of each ##contig header line, header order
 *   contig[i], pos[i] (1-based), type[i] (0 :snp, 1 :indel per vartype :61-69), ref_len[i],
 *   alt_off[i]/alt_len[i] into alt_bytes, hap[i] (1-based haplotype slot)
 *   sample_stream: Philox sample-stream id (the 0-based VCF sample column). */
int32_t mmk_set_sample_variants(mmk_ctx *ctx, uint32_t sample_stream, int32_t ploidy,
                                const int32_t *vcf_contigs, int32_t n_vcf_contigs, int64_t n_mut,
                                const int32_t *contig, const int64_t *pos, const uint8_t *type,
                                const int32_t *ref_len, const int64_t *alt_off, const int32_t *alt_len,
                                const uint8_t *hap, const char *alt_bytes, int64_t alt_bytes_len);

/* VCF text reader (VariantCallFormat equivalent for the fields the reference uses): one pass
 * over a gzip / BGZF / plain VCF collects header contigs, sample names, ploidy
 * (get_ploidy_and_contigs, src/ProcessVariants.jl:10-20) and every sample's mutations. */
typedef struct mmk_vcf mmk_vcf;
int32_t mmk_vcf_open(const char *path, mmk_vcf **out, char *errbuf, size_t errbuf_len);
void mmk_vcf_close(mmk_vcf *v);
int32_t mmk_vcf_n_samples(const mmk_vcf *v);
const char *mmk_vcf_sample_name(const mmk_vcf *v, int32_t i);
int32_t mmk_vcf_ploidy(const mmk_vcf *v);
int32_t mmk_vcf_n_contigs(const mmk_vcf *v);
const char *mmk_vcf_contig_name(const mmk_vcf *v, int32_t i);
int64_t mmk_vcf_n_records(const mmk_vcf *v);
/* resolves header contigs against ctx's master contigs and calls mmk_set_sample_variants */
int32_t mmk_apply_vcf_sample(mmk_ctx *ctx, const mmk_vcf *v, int32_t sample_index);

/* schemas = (contig, haplotype) units the loop samples from */
int32_t mmk_n_schemas(const mmk_ctx *ctx);
typedef struct mmk_schema_info {
    int64_t length;          /* haplotype length after variants */
    int64_t reads_required;  /* src/ProcessFASTA.jl:61 (inherited from the master in VCF mode) */
    int64_t reads_current;   /* after mmk_plan */
    int32_t haplotype;
    int32_t contig;          /* index of the master contig */
    char name[64];
} mmk_schema_info;
int32_t mmk_schema_get(const mmk_ctx *ctx, int32_t i, mmk_schema_info *out);
/* the schema's chrom (Schema.chrom, src/Structs.jl:78-84) in full, NUL-terminated: mmk_schema_info.name holds only its first
 * 63 characters, contig names may have 160.  MMK_ERR_INVALID_ARG when `cap` is too small. */
int32_t mmk_schema_name(const mmk_ctx *ctx, int32_t i, char *out, int64_t cap);
/* copies schema i back as uppercase ASCII (parity check of haplotype construction) */
int32_t mmk_schema_bases(mmk_ctx *ctx, int32_t i, char *out, int64_t cap);

/* ---- planning ------------------------------------------------------------------------------
 * The barcode-group loop of src/Main.jl:44-59: molecules per barcode, target picks, molecule
 * size, number of inserts, coverage tracking and schema pruning — for the whole run, on the
 * device, in speculative batches cut at each pruning point.  Every rank of a sharded run plans
 * the same (cheap) schedule. */
typedef struct mmk_plan_info {
    int64_t n_groups;
    int64_t n_molecules;     /* = barcodes consumed (one per molecule, src/Main.jl:51) */
    int64_t n_pairs;
    int64_t n_epochs;        /* pruning points */
    int64_t n_batches;
    double host_ms;          /* wall time of the call (kernels + control round trips) */
} mmk_plan_info;
int32_t mmk_plan(mmk_ctx *ctx, mmk_plan_info *out);

/* Restrict this context's emission to shard `rank` of `world`: a contiguous molecule range
 * holding ~1/world of the read pairs.  Concatenating the shards' outputs in rank order gives
 * the unsharded files. Call after mmk_plan, before mmk_place. Default: (0, 1). */
int32_t mmk_set_shard(mmk_ctx *ctx, int32_t rank, int32_t world);

/* Read-pair placement for this shard: insert sizes, molecule start, gap chain
 * (src/Breakpoints.jl:205-218, :227-251, :21-35), per-molecule header text and barcodes
 * (src/Barcodes.jl), record lengths and the byte offsets of every tile of records. */
typedef struct mmk_place_info {
    int64_t first_molecule, n_molecules;
    int64_t first_pair, n_pairs;
    uint64_t r1_bytes, r2_bytes; /* total FASTQ bytes of this shard */
    uint64_t gather_bytes;       /* packed reference bytes the extraction reads (algorithmic) */
    double host_ms;              /* wall time of the call */
    double phase_ms[4];          /* device time: placement, tile sums + scans, chunk table, (unused) */
} mmk_place_info;
int32_t mmk_place(mmk_ctx *ctx, mmk_place_info *out);

/* ---- emission ------------------------------------------------------------------------------
 * extract_sequences! + format_R1/format_R2 (src/Breakpoints.jl:95-140, src/FormatFASTQ.jl):
 * fills the next device buffer pair with as many whole records as fit and synchronises. */
typedef struct mmk_chunk {
    void *r1_dev, *r2_dev;       /* device pointers */
    uint64_t r1_bytes, r2_bytes;
    uint64_t r1_offset, r2_offset; /* byte offset of this chunk inside the shard's R1 / R2 text */
    int64_t first_pair, n_pairs; /* shard-relative */
    int32_t done;                /* 1 when this was the last chunk of the shard */
    int32_t slot;
    float kernel_ms;             /* device time of the emit kernel(s) for this chunk (CUDA events) */
} mmk_chunk;
int32_t mmk_emit_chunk(mmk_ctx *ctx, mmk_chunk *out);
/* rewinds emission to the first chunk of the shard (plan and placement are kept) */
int32_t mmk_rewind(mmk_ctx *ctx);
/* device → host copy of a chunk (synchronous) */
int32_t mmk_copy_chunk(mmk_ctx *ctx, const mmk_chunk *chunk, void *r1_host, void *r2_host);
/* Whole shard to host memory: emission of chunk k+1 overlaps the copy of chunk k.
 * r1_host / r2_host need mmk_place_info.r1_bytes / r2_bytes bytes. */
int32_t mmk_emit_to_host(mmk_ctx *ctx, void *r1_host, uint64_t r1_cap, void *r2_host, uint64_t r2_cap);

/* Whole shard, device-resident: every chunk is generated back to back into the output ring (later
 * chunks overwrite earlier ones when the shard is larger than the ring — "device-resident FASTQ"
 * for a 446 GB job means exactly this).  No host copies, one synchronisation at the end.
 * *kernel_ms = device time from the first to the last emit kernel (CUDA events on the stream). */
int32_t mmk_emit_all(mmk_ctx *ctx, float *kernel_ms);

/* Whole shard through a bounded host ring: chunk k is copied into r1_host[k % 2] / r2_host[k % 2]
 * (each `slot_cap` bytes, >= the context's chunk size; pinned memory recommended) while chunk k+1 is
 * generated.  `cb`, if not NULL, is called once per chunk, in order, when its bytes are in host
 * memory; the slot is reused after it returns.  A non-zero return from `cb` aborts with that status. */
typedef int32_t (*mmk_chunk_cb)(void *user, const mmk_chunk *chunk, const void *r1_host, const void *r2_host);
int32_t mmk_emit_stream(mmk_ctx *ctx, void *const r1_host[2], void *const r2_host[2], uint64_t slot_cap,
                        mmk_chunk_cb cb, void *user);

/* ---- BGZF output -----------------------------------------------------------------------------
 * The reference writes <prefix>.R{1,2}.fq.gz through BGZFCompressorStream (src/Main.jl:42-43,
 * 111-112).  These entry points deliver the same text as blocked gzip: every chunk is a whole number
 * of BGZF members (one per tile of records, coded on the device), so chunks — and the shards of
 * several ranks — concatenate into a valid file; the writer appends mmk_bgzf_eof() once at the end.
 * In the mmk_chunk of these calls r*_dev / r*_bytes / r*_offset describe the CODED bytes.
 *
 * mmk_bgzf_bound       host (or device) capacity one coded chunk of one stream may need
 * mmk_emit_chunk_bgzf  like mmk_emit_chunk; the device pointers stay valid until the next emit call
 * mmk_emit_stream_bgzf like mmk_emit_stream (slot_cap >= mmk_bgzf_bound)
 * mmk_bgzf_eof         writes the 28-byte BGZF end-of-file member, returns 28
 * mmk_bgzf_table       the fixed deflate code (reversed code | length << 16 per symbol) and the constant
 *                      block prefix bit string; host-only, for tests */
int32_t mmk_bgzf_bound(mmk_ctx *ctx, uint64_t *slot_bytes);
int32_t mmk_emit_chunk_bgzf(mmk_ctx *ctx, mmk_chunk *out);
int32_t mmk_emit_stream_bgzf(mmk_ctx *ctx, void *const r1_host[2], void *const r2_host[2], uint64_t slot_cap,
                             mmk_chunk_cb cb, void *user);
int32_t mmk_bgzf_eof(void *out, uint32_t cap);
int32_t mmk_bgzf_table(uint32_t *litlen_codes, uint32_t *dist_codes, uint8_t *prefix, uint32_t prefix_cap,
                       uint32_t *prefix_bits);
/* Whole shard into two files, as the reference's writers do (src/Main.jl:42-43,111-112): BGZF members coded on the
 * device when `bgzf` != 0 (then `append_eof` != 0 appends the end-of-file member: the last rank of a sharded run), plain
 * FASTQ text otherwise.  `append` != 0 opens the files for appending.  The library owns a pinned host ring; the writes of
 * chunk k (one thread per file) overlap the generation and copy-out of chunk k + 1.  bytes_written[2]: R1, R2 (may be NULL). */
int32_t mmk_emit_files(mmk_ctx *ctx, const char *r1_path, const char *r2_path, int32_t bgzf, int32_t append_eof, int32_t append,
                       uint64_t bytes_written[2]);
/* bytes each output slot holds for this shard (after mmk_place) and the number of chunks */
int32_t mmk_chunk_layout(mmk_ctx *ctx, uint64_t *slot_bytes, int32_t *n_chunks);

/* number of kernels this library has launched in this process so far (all contexts) */
uint64_t mmk_launch_count(void);

/* ---- introspection for the parity tests ------------------------------------------------------ */
typedef struct mmk_mol_record { /* one per planned molecule, run order */
    int64_t group;
    int32_t slot;
    int32_t schema;
    int64_t start;       /* 1-based; 0 until mmk_place covered the molecule */
    int64_t molsize;
    int64_t n_frags;
    int64_t first_read;  /* global pair index of the molecule's first insert */
    int64_t barcode_index;
} mmk_mol_record;
/* copies molecules [first, first+n) of the plan */
int32_t mmk_dump_molecules(mmk_ctx *ctx, int64_t first, int64_t n, mmk_mol_record *out);
/* insert breakpoints (1-based inclusive) of shard-relative pairs [first, first+n) */
int32_t mmk_dump_inserts(mmk_ctx *ctx, int64_t first, int64_t n, int64_t *start, int64_t *stop);

/* The variates the kernels consumed, serialised in the reference's draw order
 * (SURVEY.md §8a-bis): for the oracle's replay mode. Entries are 16 bytes. */
typedef struct mmk_variate {
    uint32_t tag;  /* enum mmk_purpose (mmk_rng.h) */
    uint32_t aux;  /* integer draws: low 32 bits of the range size */
    union {
        double d;
        int64_t i;
    } v;
} mmk_variate;
/* sizes first (out == NULL), then fills. Covers the whole plan (all molecules), so call it on an
 * unsharded context. N-masking variates are reported by mmk_dump_mask_variates. */
int32_t mmk_dump_variates(mmk_ctx *ctx, mmk_variate *out, uint64_t cap, uint64_t *n);
int32_t mmk_dump_mask_variates(mmk_ctx *ctx, mmk_variate *out, uint64_t cap, uint64_t *n);

/* Known-answer hooks: evaluate device code on small inputs. */
int32_t mmk_kat_philox(mmk_ctx *ctx, const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* kind 0: mmk_log, 1: mmk_exp, 2: bits→exp variate, 3: polar normal from stream (seed=x bits) */
int32_t mmk_kat_math(mmk_ctx *ctx, int32_t kind, const double *in, double *out, int64_t n);
/* Insert lengths the placement kernels derive from raw 64-bit draws (threshold table with its exact fallback):
 * len[i] = trunc(Int, mu + sd * z(bits[i])), 0 where the truncation at 200 rejects the draw; slow[i] = 1 where
 * the double-precision fallback ran (near a threshold, outside the table, table disabled). */
/* Host-only: the threshold table mmk_create builds for Normal(insert_length, insert_stdev) truncated at 200 (needs no
 * device).  thr[k] = top 32 of the 52 draw bits of the smallest draw whose length is base_len + k + 1; the table answers
 * for top-32 draws in [u_lo, u_hi] farther than `guard` from every threshold.  *enabled = 0: no table (exact path only). */
int32_t mmk_insert_table(int32_t insert_length, int32_t insert_stdev, uint32_t *enabled, uint32_t *n_thr, uint32_t *base_len,
                         uint32_t *u_lo, uint32_t *u_hi, uint32_t *guard, uint32_t *thr, uint32_t thr_cap);
int32_t mmk_kat_insert_len(mmk_ctx *ctx, const uint64_t *bits, uint32_t *len, uint8_t *slow, int64_t n);

#ifdef __cplusplus
}
#endif
#endif /* MIMICK_B200_H */
