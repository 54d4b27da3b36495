// mmk_host.cpp — the context behind the C ABI (include/mimick_b200.h): parameter freezing,
// schema bookkeeping, the epoch loop of the planner, sharding, chunked emission.
//
// Reference anchors (mimick/MimickLinkedReads.jl/src): Main.jl:24-62, :88-131 (drivers),
// Structs.jl:54-70 (SimParams), ProcessFASTA.jl:45-95 (schemas), ProcessVariants.jl:157-185
// (sample haplotypes).  All sampling, extraction and formatting happens in mmk_device.cu; this
// file only sequences kernels and moves small control values.
#include <math.h>
#include <errno.h>
#include <fcntl.h>
#include <stdarg.h>
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <map>
#include <random>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "mmk_host.h"
#include "mmk_internal.h"

using namespace mmkdev;

// ------------------------------------------------------------------------------------------
// device buffer helper
// ------------------------------------------------------------------------------------------
template <class T>
struct DBuf {
    T *p = nullptr;
    size_t cap = 0;  // elements
    ~DBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    // grow to at least n elements; keeps the first `keep` elements
    cudaError_t ensure(size_t n, size_t keep = 0, cudaStream_t st = 0) {
        if (n <= cap) return cudaSuccess;
        size_t nc = std::max(n, cap + cap / 2);
        T *q = nullptr;
        cudaError_t e = cudaMalloc((void **)&q, nc * sizeof(T));
        if (e != cudaSuccess) return e;
        if (p && keep) {
            e = cudaMemcpyAsync(q, p, keep * sizeof(T), cudaMemcpyDeviceToDevice, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        }
        if (p) cudaFree(p);
        p = q;
        cap = nc;
        return e;
    }
};

struct Contig {
    std::string name;
    int32_t haplotype = 1;
    int64_t len = 0;
    char *d_ascii = nullptr;  // upper-case, masked; kept while variants may still be applied
    int64_t reads_required = 0;
    uint64_t n_N = 0, n_nonACGT = 0;
    std::vector<mmk_variate> mask_variates;
};
struct Schema {
    int32_t contig = 0;
    int32_t haplotype = 1;
    int64_t len = 0;
    int64_t required = 0, current = 0;
    uint64_t word_off = 0;
    std::string name;
};

struct mmk_ctx {
    mmk_params par;
    SimConst sc;
    uint64_t seed = 0;
    cudaStream_t st = nullptr, st_copy = nullptr;
    cudaMemPool_t sym_pool = nullptr;  // contig symbols come and go per run: stream-ordered pool that keeps its memory
    std::string err;

    std::vector<Contig> contigs;
    std::vector<Schema> schemas;
    bool schemas_ready = false;  // arena packed + SchemaDev uploaded
    bool multi_mode = false;
    int bpb = 2;
    DBuf<uint32_t> d_hap;
    DBuf<SchemaDev> d_sch;
    DBuf<char> d_names;  // contig names of the schemas, back to back (SchemaDev::name points here)
    ContigStats *d_stats = nullptr;
    ContigStats *h_stats = nullptr;  // pinned

    // plan
    bool planned = false;
    DBuf<uint64_t> p_group;
    DBuf<uint32_t> p_slot;
    DBuf<uint16_t> p_schema;
    DBuf<uint16_t> p_epoch;
    DBuf<uint32_t> p_size;
    DBuf<uint32_t> p_nfrags;
    DBuf<uint8_t> p_status;
    DBuf<uint64_t> d_read_off;  // [n_mol + 1]
    // planner scratch, kept between calls
    DBuf<uint32_t> d_gnmol;     // group table: molecules per barcode group
    DBuf<uint64_t> d_goff;      // group table: first molecule of every group
    uint64_t g_tab = 0;         // groups covered by the table
    DBuf<uint16_t> d_live, d_lpos;
    DBuf<int64_t> d_need, d_required, d_current;
    DBuf<unsigned long long> d_hist, d_first_err, d_sums;
    DBuf<uint32_t> d_epoch_nlive;
    DBuf<PlanState> d_pstate;   // the planner's control block (mmk_plan.cuh)
    PlanState *h_pstate = nullptr;  // pinned mirror, read once per round of batches
    DBuf<uint64_t> d_scan_tmp;
    // placement / chunking scratch, kept between calls
    DBuf<uint64_t> s_bounds, s_b1, s_b2;
    DBuf<unsigned long long> s_err, s_gather;
    DBuf<uint32_t> s_sum1, s_sum2, s_cb, s_cn;
    uint64_t n_mol = 0, n_pairs = 0, n_groups = 0;
    std::vector<uint32_t> epoch_nlive;
    mmk_plan_info plan_info{};

    // shard + placement
    int32_t rank = 0, world = 1;
    bool placed = false;
    uint64_t mol0 = 0, s_nmol = 0, pair0 = 0, s_npairs = 0;
    DBuf<uint2> d_ins;       // insert plan: written by the three-kernel placement; derived from the descriptors on demand otherwise
    bool ins_valid = false;
    DBuf<MolRec> d_mol;
    DBuf<uint8_t> d_text;
    DBuf<uint64_t> d_roff_l;
    DBuf<uint64_t> d_tile_off1, d_tile_off2;
    DBuf<uint32_t> d_tile_mol0;
    DBuf<uint4> d_fdesc;
    DBuf<uint8_t> d_tile_flags;
    DBuf<uint8_t> d_qfill;        // 2 * stage_cap bytes of 'I': source of the emit kernel's stage fill (bulk copy)
    // sequencing-error model (optional): per-cycle substitution thresholds and base qualities, R1 cycles then R2
    bool err_on = false;
    uint32_t err_tins = 0, err_tdel = 0;
    DBuf<uint32_t> d_err_thr;
    DBuf<uint8_t> d_err_qb;
    DBuf<uint8_t> d_tmp, d_tmp2;  // placement scratch: (gap bits, insert length) per pair of one segment (two alternate)
    // haplotype construction scratch, kept between samples (a 100 Mbp master: cudaMalloc + cudaFree of these cost more than the kernels)
    DBuf<char> v_alt, v_tmp, v_snp_alt;
    DBuf<int64_t> v_snp_pos, v_dst, v_src;
    DBuf<uint8_t> v_kind;
    DBuf<InsTable> d_instab;  // insert-length threshold table (built by mmk_create)
    InsTable *h_instab = nullptr;
    DBuf<uint32_t> d_weight, d_blk, d_big_list, d_big_count;  // block placement: molecule weights, block table, big molecules
    DBuf<uint64_t> d_woff;                                    // ... weight prefix
    DBuf<uint32_t> d_acc_i;   // placement scratch: sum of insert lengths per molecule
    uint32_t text_stride = 0, tile_pairs = 64, n_tiles = 0, stage_cap = 0;
    int32_t fast_ok = 0;
    mmk_place_info place_info{};

    // emission
    uint8_t *out[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};  // [slot][read]
    uint64_t out_cap = 0;
    uint32_t tile_cursor = 0;  // index into chunk_tiles of the next chunk
    int slot = 0;
    std::vector<uint32_t> chunk_tiles;          // [n_chunks + 1] tile boundaries
    std::vector<uint64_t> chunk_off1, chunk_off2;  // [n_chunks + 1] byte offsets at those boundaries
    cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_emit[2] = {nullptr, nullptr}, ev_copy[2] = {nullptr, nullptr};
    cudaEvent_t ev_draw[2] = {nullptr, nullptr}, ev_chain[2] = {nullptr, nullptr};  // placement: draws / chain per scratch buffer

    // BGZF stage
    uint32_t tile_bytes_bound = 0;   // upper bound of one tile's text in one stream
    DBuf<uint32_t> d_bz_tab;         // codes, CRC table, block prefix
    uint32_t bz_hdr_bits = 0;
    uint32_t bz_tiles_per_block = 1, bz_in_cap = 0, bz_max_blocks = 0, bz_slot_shift = 1, bz_crc_words = 1;
    uint64_t bz_out_cap = 0;
    DBuf<uint8_t> d_bz_slots;
    DBuf<uint32_t> d_bz_sizes;
    DBuf<uint64_t> d_bz_offs;
    DBuf<uint8_t> bz_out[2][2];      // [slot][read] packed BGZF bytes of a chunk
    void *h_ring[4] = {nullptr, nullptr, nullptr, nullptr};  // pinned host ring of mmk_emit_files
    uint64_t h_ring_cap = 0;
    uint64_t *h_bz_tot = nullptr;    // pinned [slot][read]
    uint64_t bz_cum[2] = {0, 0};     // BGZF bytes handed out so far per stream
    bool bz_ready = false;
};

// contig symbol buffers (uploaded ASCII, released once packed): stream-ordered allocations from a pool that keeps
// freed memory, so repeated runs on one context do not pay cudaMalloc / cudaFree (~1 ms each at these sizes)
static cudaError_t sym_alloc(mmk_ctx *c, char **p, size_t bytes) {
    if (c->sym_pool) return cudaMallocFromPoolAsync((void **)p, bytes, c->sym_pool, c->st);
    return cudaMalloc((void **)p, bytes);
}
static void sym_free(mmk_ctx *c, char *p) {
    if (!p) return;
    if (c->sym_pool) cudaFreeAsync(p, c->st);
    else cudaFree(p);
}

// ------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------
static int32_t fail(mmk_ctx *c, int32_t code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    return code;
}
#define CU(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? MMK_ERR_NOMEM : MMK_ERR_CUDA, "%s: %s", #call, \
                        cudaGetErrorString(e_));                                                        \
    } while (0)

extern "C" const char *mmk_strerror(int32_t s) {
    switch (s) {
    case MMK_OK: return "ok";
    case MMK_ERR_INVALID_ARG: return "invalid argument";
    case MMK_ERR_BAD_FORMAT: return "barcode format not recognized";
    case MMK_ERR_INSERTS_UNFITTABLE: return "failed to find practical read sizes less than molecule";
    case MMK_ERR_BARCODES_EXHAUSTED: return "no more barcodes available for the selected barcode type";
    case MMK_ERR_CONTIG_TOO_AMBIGUOUS: return "contig lacks 95% of the read length sum as unambiguous bases";
    case MMK_ERR_BCF_UNSUPPORTED: return "BCF file type is not yet supported";
    case MMK_ERR_TRUNCATION_BOUNDS: return "molecule shorter than two inserts (truncation lower > upper)";
    case MMK_ERR_BAD_BASE: return "symbol outside the DNA alphabet";
    case MMK_ERR_IO: return "file does not exist or cannot be read";
    case MMK_ERR_VCF: return "malformed or inconsistent VCF";
    case MMK_ERR_BOUNDS: return "read window out of bounds";
    case MMK_ERR_UNSUPPORTED: return "parameter region not supported";
    case MMK_ERR_CUDA: return "CUDA error";
    case MMK_ERR_NOMEM: return "out of memory";
    case MMK_ERR_REPLAY: return "replay stream mismatch";
    case MMK_ERR_STATE: return "call out of order";
    default: return "unknown status";
    }
}
extern "C" const char *mmk_last_error(const mmk_ctx *c) { return c ? c->err.c_str() : "null context"; }
extern "C" int32_t mmk_abi_version(void) { return MMK_ABI_VERSION; }

// ------------------------------------------------------------------------------------------
// parameters
// ------------------------------------------------------------------------------------------
extern "C" void mmk_params_default(mmk_params *p) {
    memset(p, 0, sizeof *p);
    p->struct_size = sizeof *p;
    p->device = 0;
    p->fq_format = MMK_FMT_STANDARD;
    p->bc_format = MMK_FMT_HAPLOTAGGING;
    p->coverage = 30.0;
    p->n_molecules = 2;
    p->attempts = 25;
    p->mol_coverage = 0.2;
    p->mol_length = 80000;
    p->insert_length = 500;
    p->insert_stdev = 50;
    p->length_r1 = 150;
    p->length_r2 = 150;
    p->singletons = 0.70;
    p->circular = 0;
    p->mask = 0;
    p->seed = -1;
    p->strict_wrap = 0;
    p->force_generic = 0;
    p->max_chunk_bytes = 0;
    p->plan_batch_groups = 0;
}

extern "C" int32_t mmk_interpret_format(const char *fmt_c, int32_t *bc, int32_t *fq) {
    if (!fmt_c || !bc || !fq) return MMK_ERR_INVALID_ARG;
    std::string fmt(fmt_c);
    if (fmt == "10x") fmt = "tenx";  // CLI normalisation, mimick/__main__.py:73
    auto has = [&](const char *s) { return fmt.find(s) != std::string::npos; };
    if (has("standard")) {
        *fq = MMK_FMT_STANDARD;
        if (has("stlfr")) *bc = MMK_FMT_STLFR;
        else if (has("haplotagging")) *bc = MMK_FMT_HAPLOTAGGING;
        else *bc = MMK_FMT_TELLSEQ;
        return MMK_OK;
    }
    int v = fmt == "stlfr" ? MMK_FMT_STLFR : fmt == "haplotagging" ? MMK_FMT_HAPLOTAGGING : fmt == "tellseq" ? MMK_FMT_TELLSEQ
            : fmt == "tenx" ? MMK_FMT_TENX : -1;
    if (v < 0) return MMK_ERR_BAD_FORMAT;
    *bc = *fq = v;
    return MMK_OK;
}

extern "C" int32_t mmk_check_variant_path(const char *path) {
    if (!path) return MMK_ERR_INVALID_ARG;
    std::string s(path);
    std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return (char)tolower(c); });
    if (s.size() >= 3 && s.compare(s.size() - 3, 3, "bcf") == 0) return MMK_ERR_BCF_UNSUPPORTED;
    return MMK_OK;
}

// truncated(Exponential(theta), lower[, upper]) — host twin of trunc_exp_make in mmk_sampling.cuh
static bool host_trunc_exp(TruncExp &d, double theta, double lower, bool has_upper, double upper) {
    memset(&d, 0, sizeof d);
    d.theta = theta;
    d.lower = lower;
    d.upper = upper;
    d.has_upper = has_upper;
    if (has_upper && !(lower <= upper)) return false;
    double zl = MMK_DIV(lower > 0.0 ? lower : 0.0, theta);
    d.loglcdf = mmk_log1mexp(-zl);
    d.lcdf = mmk_exp(d.loglcdf);
    if (has_upper) {
        double zu = MMK_DIV(upper > 0.0 ? upper : 0.0, theta);
        d.logtp = mmk_logsubexp(d.loglcdf, mmk_log1mexp(-zu));
    } else
        d.logtp = mmk_log1mexp(d.loglcdf);
    d.tp = mmk_exp(d.logtp);
    return true;
}

static uint64_t barcode_capacity(int bc) {
    switch (bc) {
    case MMK_FMT_HAPLOTAGGING: return 96ull * 96 * 96 * 96;
    case MMK_FMT_STLFR: return 8000000000ull;
    case MMK_FMT_TELLSEQ: return 1ull << 36;
    case MMK_FMT_TENX: return 1ull << 32;
    default: return 0;
    }
}

// ------------------------------------------------------------------------------------------
// insert-length threshold table (mmk_place.cuh explains the scheme)
// ------------------------------------------------------------------------------------------
static const uint32_t INS_GUARD_HOST = 16;  // must equal INS_GUARD in mmk_place.cuh
// trunc(Int, mu + sd z(u)) of the 52-bit draw u, 0 when the truncation at 200 rejects it (z < lb)
static uint32_t exact_insert_len(const SimConst &sc, uint64_t u52) {
    const double z = mmk_bits_to_normal(u52 << 12);
    if (z < sc.ins_lb) return 0;
    const double x = MMK_ADD(sc.ins_mu, MMK_MUL(sc.ins_sd, z));
    uint32_t li = x > 0 ? (uint32_t)(long long)x : 0u;
    return li > 0x00ffffffu ? 0x00ffffffu : li;
}
static bool table_insert_len(const InsTable &t, uint32_t u32, uint32_t &len) {  // host twin of ins_len_fast
    const uint32_t c = u32 >> (32 - INS_CELL_BITS);
    uint32_t lo = t.cell[c], hi = t.cell[c + 1];
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (t.thr[mid] <= u32) lo = mid + 1; else hi = mid;
    }
    len = t.base_len + lo;
    const bool below_ok = lo == 0 || u32 - t.thr[lo - 1] > INS_GUARD_HOST;
    const bool above_ok = lo == t.n_thr || t.thr[lo] - u32 > INS_GUARD_HOST;
    return below_ok && above_ok && u32 >= t.u_lo && u32 <= t.u_hi;
}
template <class Pred>
static uint64_t first_u52(uint64_t lo, uint64_t hi, Pred pred) {  // smallest u in [lo, hi] with pred(u), hi + 1 if none (pred monotone)
    if (!pred(hi)) return hi + 1;
    while (lo < hi) {
        const uint64_t mid = lo + ((hi - lo) >> 1);
        if (pred(mid)) hi = mid; else lo = mid + 1;
    }
    return lo;
}
static void build_insert_table(const SimConst &sc, InsTable &t) {
    memset(&t, 0, sizeof t);
    if (sc.ins_robert || !(sc.ins_sd > 0.0)) return;  // the table tabulates the rejection sampler's first (accepted) normal
    const uint64_t U = (1ull << 52) - 1;
    const double zmax = std::min(7.9, 0.48 * INS_MAX_THR / sc.ins_sd);
    const double zmin = std::max(-zmax, sc.ins_lb);
    if (!(zmin < zmax)) return;
    const uint64_t ua = first_u52(0, U, [&](uint64_t u) { return mmk_bits_to_normal(u << 12) >= zmin; });
    const uint64_t ub = first_u52(0, U, [&](uint64_t u) { return mmk_bits_to_normal(u << 12) > zmax; }) - 1;
    if (ua > U || ub <= ua || (ub >> 20) < (ua >> 20) + 4 * INS_GUARD_HOST + 8) return;
    const uint32_t base = exact_insert_len(sc, ua), top = exact_insert_len(sc, ub);
    if (base == 0 || top < base || top - base > (uint32_t)INS_MAX_THR) return;
    t.base_len = base;
    t.n_thr = top - base;
    for (uint32_t k = 0; k < t.n_thr; ++k) {
        const uint32_t L = base + k + 1;
        const uint64_t T = first_u52(ua, ub, [&](uint64_t u) { return exact_insert_len(sc, u) >= L; });
        t.thr[k] = (uint32_t)(T >> 20);
    }
    t.u_lo = (uint32_t)(ua >> 20) + 1 + INS_GUARD_HOST;
    t.u_hi = (uint32_t)(ub >> 20) - 1 - INS_GUARD_HOST;
    {
        uint32_t k = 0;
        for (uint32_t c = 0; c <= (uint32_t)INS_CELLS; ++c) {
            const uint64_t edge = (uint64_t)c << (32 - INS_CELL_BITS);
            while (k < t.n_thr && (uint64_t)t.thr[k] < edge) ++k;
            t.cell[c] = (uint16_t)k;
        }
    }
    t.enabled = 1;
    // self-check against the definition: both edges of the guard band of every threshold, the range ends, and a sweep
    auto agrees = [&](uint32_t u32) {
        uint32_t fast;
        if (!table_insert_len(t, u32, fast)) return true;  // exact path on the device
        return fast == exact_insert_len(sc, (uint64_t)u32 << 20) && fast == exact_insert_len(sc, ((uint64_t)u32 << 20) | 0xfffffu);
    };
    bool ok = true;
    for (uint32_t k = 0; k < t.n_thr && ok; ++k)
        for (uint32_t d = INS_GUARD_HOST + 1; d <= INS_GUARD_HOST + 2 && ok; ++d) ok = agrees(t.thr[k] - d) && agrees(t.thr[k] + d);
    ok = ok && agrees(t.u_lo) && agrees(t.u_hi);
    uint64_t x = 0x9E3779B97F4A7C15ull;
    for (int i = 0; i < 20000 && ok; ++i) {
        x ^= x << 13, x ^= x >> 7, x ^= x << 17;
        ok = agrees((uint32_t)(x >> 32));
    }
    if (!ok) t.enabled = 0;  // never observed; the exact path is always correct
}

extern "C" int32_t mmk_create(const mmk_params *p, mmk_ctx **out) {
    if (!p || !out) return MMK_ERR_INVALID_ARG;
    *out = nullptr;
    mmk_ctx *ctx = new mmk_ctx();
    *out = ctx;  // returned even on failure so the caller can read the message, then mmk_destroy
    mmk_params_default(&ctx->par);
    memcpy(&ctx->par, p, std::min<size_t>(p->struct_size ? p->struct_size : sizeof *p, sizeof *p));
    const mmk_params &q = ctx->par;
    if (q.fq_format < 0 || q.fq_format > MMK_FMT_STANDARD || q.bc_format < 0 || q.bc_format > MMK_FMT_TENX)
        return fail(ctx, MMK_ERR_BAD_FORMAT, "format enum out of range");
    if (q.fq_format != MMK_FMT_STANDARD && q.fq_format != q.bc_format)
        return fail(ctx, MMK_ERR_BAD_FORMAT, "barcode style must equal the output style unless the output style is standard");
    if (q.n_molecules == 0) return fail(ctx, MMK_ERR_INVALID_ARG, "n_molecules = 0 never terminates in the reference");
    if (q.length_r1 < 1 || q.length_r2 < 1 || q.insert_length < 1 || q.insert_stdev < 0 || q.mol_length < 1 ||
        !(q.mol_coverage > 0) || q.attempts < 0 || !(q.coverage >= 0) || q.singletons < 0)
        return fail(ctx, MMK_ERR_INVALID_ARG, "parameter out of range");
    if (q.length_r1 > 100000 || q.length_r2 > 100000) return fail(ctx, MMK_ERR_UNSUPPORTED, "read length > 100000");

    SimConst &sc = ctx->sc;
    memset(&sc, 0, sizeof sc);
    if (q.seed >= 0) ctx->seed = (uint64_t)q.seed;
    else {
        std::random_device rd;
        ctx->seed = (((uint64_t)rd() << 32) | rd()) >> 1;
    }
    sc.seed = ctx->seed;
    sc.sample = 0;
    // SimParams, Structs.jl:57-69
    double lowerbound;
    if (q.mol_coverage < 1.0) {
        double t = MMK_MUL(MMK_MUL((double)q.mol_length, q.mol_coverage), 0.25);
        int64_t ti = (int64_t)t;
        lowerbound = (double)(ti > 3000 ? ti : 3000);
    } else
        lowerbound = MMK_MUL(q.mol_coverage, MMK_ADD((double)q.insert_length, MMK_MUL(0.75, (double)q.insert_stdev)));
    host_trunc_exp(sc.mollen, (double)q.mol_length, lowerbound, false, 0.0);
    if (q.n_molecules > 0) {
        sc.nmol_fixed = 0;
        if (!host_trunc_exp(sc.nmol, (double)q.n_molecules, 1.0, true, MMK_MUL(2.5, (double)q.n_molecules)))
            return fail(ctx, MMK_ERR_TRUNCATION_BOUNDS, "n_molecules truncation bounds");
    } else
        sc.nmol_fixed = -q.n_molecules;
    sc.ins_mu = (double)q.insert_length;
    sc.ins_sd = (double)q.insert_stdev;
    sc.ins_lb = MMK_DIV(MMK_SUB(200.0, sc.ins_mu), sc.ins_sd);
    {   // the kept mass tp = 1 - cdf(lb) picks the sampler (Distributions.jl src/truncated/normal.jl: tp > 0.3 rejection)
        const double tp = 0.5 * erfc(sc.ins_lb / sqrt(2.0));
        if (sc.ins_lb != sc.ins_lb || sc.ins_lb == INFINITY)  // sd = 0 and mu <= 200
            return fail(ctx, MMK_ERR_UNSUPPORTED,
                        "insert-size truncation at 200 keeps no mass of Normal(%d, %d): the reference's sampler never returns",
                        q.insert_length, q.insert_stdev);
        sc.ins_robert = tp > 0.3 ? 0 : 1;
        sc.ins_a = MMK_DIV(MMK_ADD(sc.ins_lb, MMK_SQRT(MMK_ADD(MMK_MUL(sc.ins_lb, sc.ins_lb), 4.0))), 2.0);
        sc.ins_theta = MMK_DIV(1.0, sc.ins_a);
    }
    if ((double)q.insert_length + 13.0 * (double)q.insert_stdev >= 16777215.0)
        return fail(ctx, MMK_ERR_UNSUPPORTED, "insert lengths of 2^24 bases or more");
    sc.singletons = q.singletons;
    sc.mol_coverage = q.mol_coverage;
    sc.insert_length = q.insert_length;
    sc.attempts = q.attempts;
    sc.r1 = q.length_r1;
    sc.r2 = q.length_r2;
    sc.circular = q.circular ? 1 : 0;
    sc.strict = q.strict_wrap ? 1 : 0;
    sc.fq_format = q.fq_format;
    sc.bc_format = q.bc_format;

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(ctx, MMK_ERR_CUDA, "no CUDA device: %s (this library has no CPU path)", cudaGetErrorString(e));
    if (q.device < 0 || q.device >= ndev) return fail(ctx, MMK_ERR_INVALID_ARG, "device %d out of range", q.device);
    CU(cudaSetDevice(q.device));
    CU(cudaStreamCreateWithFlags(&ctx->st, cudaStreamNonBlocking));
    {   // second stream at the highest priority: in placement it runs the latency-bound chain kernel beside the draws
        int lo = 0, hi = 0;
        CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CU(cudaStreamCreateWithPriority(&ctx->st_copy, cudaStreamNonBlocking, hi));
    }
    {
        cudaMemPoolProps props;
        memset(&props, 0, sizeof props);
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = q.device;
        if (cudaMemPoolCreate(&ctx->sym_pool, &props) == cudaSuccess) {
            uint64_t keep = 16ull << 30;  // freed symbol buffers stay with the context up to this much
            cudaMemPoolSetAttribute(ctx->sym_pool, cudaMemPoolAttrReleaseThreshold, &keep);
        } else {
            ctx->sym_pool = nullptr;  // plain cudaMalloc / cudaFree
            cudaGetLastError();
        }
    }
    CU(cudaEventCreate(&ctx->ev_a));
    CU(cudaEventCreate(&ctx->ev_b));
    for (int i = 0; i < 2; ++i) {
        CU(cudaEventCreateWithFlags(&ctx->ev_emit[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ctx->ev_copy[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ctx->ev_draw[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ctx->ev_chain[i], cudaEventDisableTiming));
    }
    CU(cudaMalloc((void **)&ctx->d_stats, sizeof(ContigStats)));
    CU(cudaMallocHost((void **)&ctx->h_stats, sizeof(ContigStats)));
    ctx->err_on = q.err_sub_start > 0.0 || q.err_sub_end > 0.0 || q.err_ins > 0.0 || q.err_del > 0.0;
    if (q.err_sub_start < 0 || q.err_sub_end < 0 || q.err_ins < 0 || q.err_del < 0 ||
        (ctx->err_on && std::max(q.err_sub_start, q.err_sub_end) + q.err_ins + q.err_del > 0.5))
        return fail(ctx, MMK_ERR_INVALID_ARG, "error rates must be >= 0 and sum to at most 0.5");
    if (ctx->err_on) {  // same expressions as the oracle (oracle/mimick_oracle.c orc_create): thresholds scaled by 2^24, Phred of the cycle's rate
        const int S1 = q.length_r1, S2 = q.length_r2;
        std::vector<uint32_t> thr((size_t)S1 + S2);
        std::vector<uint8_t> qb((size_t)S1 + S2);
        for (int r = 0; r < 2; ++r) {
            const int S = r ? S2 : S1, o = r ? S1 : 0;
            for (int c = 0; c < S; ++c) {
                const double e = S > 1 ? q.err_sub_start + (q.err_sub_end - q.err_sub_start) * (double)c / (double)(S - 1) : q.err_sub_start;
                thr[(size_t)o + c] = (uint32_t)floor(e * MMK_ERR_SCALE);
                double et = e + q.err_ins + q.err_del;
                if (et < 1e-4) et = 1e-4;
                const int ph = (int)floor(-10.0 * log10(et) + 0.5);
                qb[(size_t)o + c] = (uint8_t)(ph < 2 ? 2 : ph > 40 ? 40 : ph);
            }
        }
        ctx->err_tins = (uint32_t)floor(q.err_ins * MMK_ERR_SCALE);
        ctx->err_tdel = (uint32_t)floor(q.err_del * MMK_ERR_SCALE);
        CU(ctx->d_err_thr.ensure(thr.size()));
        CU(ctx->d_err_qb.ensure(qb.size()));
        CU(cudaMemcpy(ctx->d_err_thr.p, thr.data(), thr.size() * 4, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(ctx->d_err_qb.p, qb.data(), qb.size(), cudaMemcpyHostToDevice));
    }
    ctx->h_instab = new InsTable();
    build_insert_table(sc, *ctx->h_instab);
    CU(ctx->d_instab.ensure(1));
    CU(cudaMemcpyAsync(ctx->d_instab.p, ctx->h_instab, sizeof(InsTable), cudaMemcpyHostToDevice, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    return MMK_OK;
}

static void free_outputs(mmk_ctx *c) {
    for (int s = 0; s < 2; ++s)
        for (int r = 0; r < 2; ++r)
            if (c->out[s][r]) {
                cudaFree(c->out[s][r]);
                c->out[s][r] = nullptr;
            }
    c->out_cap = 0;
}

extern "C" void mmk_destroy(mmk_ctx *c) {
    if (!c) return;
    if (c->st) {
        cudaSetDevice(c->par.device);
        cudaStreamSynchronize(c->st);
        if (c->st_copy) cudaStreamSynchronize(c->st_copy);  // placement and the streamed emission also queue work there
    }
    for (auto &k : c->contigs) sym_free(c, k.d_ascii);
    if (c->st) cudaStreamSynchronize(c->st);
    if (c->sym_pool) cudaMemPoolDestroy(c->sym_pool);
    free_outputs(c);
    if (c->d_stats) cudaFree(c->d_stats);
    if (c->h_stats) cudaFreeHost(c->h_stats);
    if (c->h_pstate) cudaFreeHost(c->h_pstate);
    if (c->h_bz_tot) cudaFreeHost(c->h_bz_tot);
    for (auto &p : c->h_ring)
        if (p) cudaFreeHost(p);
    delete c->h_instab;
    if (c->ev_a) cudaEventDestroy(c->ev_a);
    if (c->ev_b) cudaEventDestroy(c->ev_b);
    for (int i = 0; i < 2; ++i) {
        if (c->ev_emit[i]) cudaEventDestroy(c->ev_emit[i]);
        if (c->ev_copy[i]) cudaEventDestroy(c->ev_copy[i]);
        if (c->ev_draw[i]) cudaEventDestroy(c->ev_draw[i]);
        if (c->ev_chain[i]) cudaEventDestroy(c->ev_chain[i]);
    }
    if (c->st_copy) cudaStreamDestroy(c->st_copy);
    if (c->st) cudaStreamDestroy(c->st);
    delete c;
}
extern "C" uint64_t mmk_seed(const mmk_ctx *c) { return c ? c->seed : 0; }
extern "C" void *mmk_cuda_stream(const mmk_ctx *c) { return c ? (void *)c->st : nullptr; }

// ------------------------------------------------------------------------------------------
// inputs
// ------------------------------------------------------------------------------------------
static void invalidate_plan(mmk_ctx *c) {
    c->planned = false;
    c->placed = false;
    c->schemas_ready = false;
    c->schemas.clear();
    c->tile_cursor = 0;
}

extern "C" int32_t mmk_add_contig(mmk_ctx *ctx, const char *name, const char *bases, int64_t len, int32_t haplotype) {
    if (!ctx || !name || (!bases && len > 0) || len < 0) return fail(ctx, MMK_ERR_INVALID_ARG, "mmk_add_contig: bad argument");
    if (!ctx->st) return fail(ctx, MMK_ERR_STATE, "context was not created successfully");
    if (ctx->multi_mode) return fail(ctx, MMK_ERR_STATE, "contigs cannot be added after variants were applied");
    if (len >= (1ll << 31)) return fail(ctx, MMK_ERR_UNSUPPORTED, "contig %s has >= 2^31 bases", name);
    if (strlen(name) > 160) return fail(ctx, MMK_ERR_UNSUPPORTED, "contig name longer than 160 characters: %s", name);
    if (haplotype < 0 || haplotype > 65535) return fail(ctx, MMK_ERR_INVALID_ARG, "haplotype index out of range");
    if (ctx->contigs.size() >= 65535) return fail(ctx, MMK_ERR_UNSUPPORTED, "more than 65535 (contig, haplotype) units");
    CU(cudaSetDevice(ctx->par.device));
    Contig k;
    k.name = name;
    k.haplotype = haplotype;
    k.len = len;
    CU(sym_alloc(ctx, &k.d_ascii, (size_t)len + 64));
    ctx->contigs.push_back(k);
    Contig &c = ctx->contigs.back();
    const mmk_params &q = ctx->par;
    auto drop = [&]() {
        sym_free(ctx, c.d_ascii);
        ctx->contigs.pop_back();
    };
    // `bases` may be host memory (pageable or pinned) or device memory, e.g. symbols another rank broadcast over NVLink
    cudaError_t e = cudaMemcpyAsync(c.d_ascii, bases, (size_t)len, cudaMemcpyDefault, ctx->st);
    ContigStats zero{~0ull, 0, 0};
    if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->d_stats, &zero, sizeof zero, cudaMemcpyHostToDevice, ctx->st);
    if (e == cudaSuccess) {
        normalize_ascii(c.d_ascii, len, ctx->d_stats, ctx->st);
        e = cudaMemcpyAsync(ctx->h_stats, ctx->d_stats, sizeof zero, cudaMemcpyDeviceToHost, ctx->st);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->st);
    if (e != cudaSuccess) {
        drop();
        return fail(ctx, MMK_ERR_CUDA, "upload of contig failed: %s", cudaGetErrorString(e));
    }
    if (ctx->h_stats->first_bad != ~0ull) {
        unsigned long long b = ctx->h_stats->first_bad;
        drop();
        return fail(ctx, MMK_ERR_BAD_BASE, "contig %s: symbol at position %llu is not in the DNA alphabet", name, b + 1);
    }
    c.n_N = ctx->h_stats->n_N;
    c.n_nonACGT = ctx->h_stats->n_nonACGT;
    if (q.mask && c.n_N > 0) {  // mask_N!, ProcessFASTA.jl:29-36,53-55
        uint32_t ci = 0;        // contig index within its FASTA file (= haplotype)
        for (size_t i = 0; i + 1 < ctx->contigs.size(); ++i) ci += ctx->contigs[i].haplotype == haplotype;
        {  // keep the picks for mmk_dump_mask_variates (they are few: one per N)
            DBuf<uint32_t> flags;
            DBuf<uint64_t> off, tmp;
            DBuf<mmk_variate> outv;
            CU(flags.ensure((size_t)len));
            CU(off.ensure((size_t)len + 1));
            CU(tmp.ensure((size_t)len / 1024 + 8));
            CU(outv.ensure((size_t)c.n_N));
            dump_mask(c.d_ascii, len, ctx->seed, ci, (uint32_t)haplotype, flags.p, off.p, tmp.p, outv.p, 0, ctx->st);
            dump_mask(c.d_ascii, len, ctx->seed, ci, (uint32_t)haplotype, flags.p, off.p, tmp.p, outv.p, 1, ctx->st);
            c.mask_variates.resize((size_t)c.n_N);
            CU(cudaMemcpyAsync(c.mask_variates.data(), outv.p, (size_t)c.n_N * sizeof(mmk_variate), cudaMemcpyDeviceToHost, ctx->st));
            CU(cudaStreamSynchronize(ctx->st));
        }
        mask_n(c.d_ascii, len, ctx->seed, ci, (uint32_t)haplotype, ctx->st);
        zero = ContigStats{~0ull, 0, 0};
        CU(cudaMemcpyAsync(ctx->d_stats, &zero, sizeof zero, cudaMemcpyHostToDevice, ctx->st));
        count_bases(c.d_ascii, len, ctx->d_stats, ctx->st);
        CU(cudaMemcpyAsync(ctx->h_stats, ctx->d_stats, sizeof zero, cudaMemcpyDeviceToHost, ctx->st));
        CU(cudaStreamSynchronize(ctx->st));
        c.n_N = ctx->h_stats->n_N;
        c.n_nonACGT = ctx->h_stats->n_nonACGT;
    }
    // process_fasta, ProcessFASTA.jl:47-61
    int64_t mean_readlen = (int64_t)(q.length_r1 + q.length_r2) / 2;
    int64_t cutoff = (int64_t)nearbyint((double)(q.length_r1 + q.length_r2) * 0.95);
    int64_t normalized = len - (int64_t)c.n_N;
    if (normalized < cutoff) {
        drop();
        return fail(ctx, MMK_ERR_CONTIG_TOO_AMBIGUOUS,
                    "Contigs must have at least 95%% of the requested read length sum as unambiguous (non-N) bases: %s has %lld, needs %lld",
                    name, (long long)normalized, (long long)cutoff);
    }
    if (mean_readlen < 1) {
        drop();
        return fail(ctx, MMK_ERR_INVALID_ARG, "mean read length is zero");
    }
    double x = MMK_DIV(MMK_DIV(MMK_MUL(q.coverage, (double)normalized), (double)mean_readlen), 2.0);
    c.reads_required = (int64_t)x;
    invalidate_plan(ctx);
    return MMK_OK;
}

extern "C" int32_t mmk_load_fasta(mmk_ctx *ctx, const char *path, int32_t haplotype) {
    if (!ctx || !path) return fail(ctx, MMK_ERR_INVALID_ARG, "mmk_load_fasta: bad argument");
    std::vector<mmkio::FastaRecord> recs;
    std::string err;
    int rc = mmkio::read_fasta(path, recs, err);
    if (rc) return fail(ctx, rc, "%s", err.c_str());
    for (auto &r : recs) {
        rc = mmk_add_contig(ctx, r.name.c_str(), r.seq.data(), (int64_t)r.seq.size(), haplotype);
        if (rc) return rc;
    }
    return MMK_OK;
}
extern "C" int32_t mmk_n_contigs(const mmk_ctx *c) { return c ? (int32_t)c->contigs.size() : 0; }
extern "C" int32_t mmk_clear_contigs(mmk_ctx *ctx) {
    if (!ctx) return MMK_ERR_INVALID_ARG;
    if (ctx->st) {
        CU(cudaSetDevice(ctx->par.device));
        CU(cudaStreamSynchronize(ctx->st));
    }
    for (auto &k : ctx->contigs) sym_free(ctx, k.d_ascii);
    ctx->contigs.clear();
    ctx->multi_mode = false;
    ctx->sc.sample = 0;
    invalidate_plan(ctx);
    return MMK_OK;
}

// upload SchemaDev table
static int32_t upload_schemas(mmk_ctx *ctx) {
    std::vector<SchemaDev> h(ctx->schemas.size());
    std::string names;
    for (const Schema &s : ctx->schemas) names += s.name;
    CU(ctx->d_names.ensure(names.size() + 1));
    if (!names.empty()) CU(cudaMemcpyAsync(ctx->d_names.p, names.data(), names.size(), cudaMemcpyHostToDevice, ctx->st));
    size_t off = 0;
    for (size_t i = 0; i < h.size(); ++i) {
        const Schema &s = ctx->schemas[i];
        memset(&h[i], 0, sizeof(SchemaDev));
        h[i].word_off = s.word_off;
        h[i].len = (uint32_t)s.len;
        h[i].haplotype = (uint16_t)s.haplotype;
        h[i].name_len = (uint16_t)s.name.size();
        h[i].name = ctx->d_names.p + off;
        off += s.name.size();
    }
    CU(ctx->d_sch.ensure(h.size() ? h.size() : 1));
    if (!h.empty()) CU(cudaMemcpyAsync(ctx->d_sch.p, h.data(), h.size() * sizeof(SchemaDev), cudaMemcpyHostToDevice, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));  // `h` and `names` are host temporaries
    return MMK_OK;
}

static inline uint64_t words_for(int64_t len, int bpb) {
    uint64_t per = 32 / bpb;
    uint64_t w = ((uint64_t)len + per - 1) / per + 48;  // pad words: the emit kernel's 16/32-word window loads over-read
    return (w + 3) & ~3ull;                            // 16-byte aligned starts
}

// single-sample mode: schemas are the contigs, in the order they were added (setup_schema)
static int32_t finalize_single(mmk_ctx *ctx) {
    if (ctx->schemas_ready) return MMK_OK;
    if (!ctx->schemas.empty() && ctx->schemas.size() == ctx->contigs.size() && ctx->d_hap.p && !ctx->contigs[0].d_ascii) {
        ctx->schemas_ready = true;  // arena from an earlier plan is still valid (no contig was added since)
        return MMK_OK;
    }
    if (ctx->contigs.empty()) return fail(ctx, MMK_ERR_STATE, "no contigs loaded");
    ctx->bpb = 2;
    for (auto &c : ctx->contigs)
        if (c.n_nonACGT) ctx->bpb = 4;
    ctx->schemas.clear();
    uint64_t off = 0;
    for (size_t i = 0; i < ctx->contigs.size(); ++i) {
        Contig &c = ctx->contigs[i];
        if (!c.d_ascii) return fail(ctx, MMK_ERR_STATE, "contig symbols were already released");
        Schema s;
        s.contig = (int32_t)i;
        s.haplotype = c.haplotype;
        s.len = c.len;
        s.required = c.reads_required;
        s.word_off = off;
        s.name = c.name;
        off += words_for(c.len, ctx->bpb);
        ctx->schemas.push_back(s);
    }
    CU(ctx->d_hap.ensure(off ? off : 4));
    CU(cudaMemsetAsync(ctx->d_hap.p, 0, off * 4, ctx->st));
    ContigStats zero{~0ull, 0, 0};
    CU(cudaMemcpyAsync(ctx->d_stats, &zero, sizeof zero, cudaMemcpyHostToDevice, ctx->st));
    DBuf<int64_t> piece;  // dst = 0, src = 0
    DBuf<uint8_t> kind;
    CU(piece.ensure(2));
    CU(kind.ensure(1));
    CU(cudaMemsetAsync(piece.p, 0, 16, ctx->st));
    CU(cudaMemsetAsync(kind.p, 0, 1, ctx->st));
    for (size_t i = 0; i < ctx->contigs.size(); ++i) {
        Contig &c = ctx->contigs[i];
        if (ctx->bpb == 2)  // whole contigs of upper-case ACGT: 16 symbols per thread, no piece search
            pack_plain2(c.d_ascii, c.len, ctx->d_hap.p + ctx->schemas[i].word_off, ctx->d_stats, ctx->st);
        else
            pack_pieces(c.d_ascii, c.d_ascii, piece.p, piece.p + 1, kind.p, 1, c.len, ctx->d_hap.p + ctx->schemas[i].word_off,
                        ctx->bpb, ctx->d_stats, ctx->st);
    }
    CU(cudaMemcpyAsync(ctx->h_stats, ctx->d_stats, sizeof zero, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    if (ctx->h_stats->first_bad != ~0ull) return fail(ctx, MMK_ERR_BAD_BASE, "unencodable symbol while packing");
    int32_t rc = upload_schemas(ctx);
    if (rc) return rc;
    // the packed arena is now the only copy a single-sample run needs
    for (auto &c : ctx->contigs) {
        sym_free(ctx, c.d_ascii);
        c.d_ascii = nullptr;
    }
    ctx->schemas_ready = true;
    return MMK_OK;
}

static bool valid_iupac(char &c) {
    if (c >= 'a' && c <= 'z') c = (char)(c - 32);
    return c == '-' || (c != 0 && strchr("ACMGRSVTWYHKDBN", c) != nullptr);
}

struct Piece {
    uint8_t kind;  // 0 master (SNP-painted copy), 1 alt bytes
    int64_t src, len;
};

extern "C" int32_t mmk_set_sample_variants(mmk_ctx *ctx, uint32_t sample_stream, int32_t ploidy, const int32_t *vcf_contigs,
                                           int32_t n_vcf_contigs, int64_t n_mut, const int32_t *contig, const int64_t *pos,
                                           const uint8_t *type, const int32_t *ref_len, const int64_t *alt_off,
                                           const int32_t *alt_len, const uint8_t *hap, const char *alt_bytes,
                                           int64_t alt_bytes_len) {
    if (!ctx || ploidy < 1 || ploidy > 255 || n_vcf_contigs < 0 || n_mut < 0 || (n_vcf_contigs && !vcf_contigs))
        return fail(ctx, MMK_ERR_INVALID_ARG, "mmk_set_sample_variants: bad argument");
    if (ctx->contigs.empty()) return fail(ctx, MMK_ERR_STATE, "no master contigs loaded");
    if ((int64_t)n_vcf_contigs * ploidy > 65535) return fail(ctx, MMK_ERR_UNSUPPORTED, "more than 65535 (contig, haplotype) units");
    CU(cudaSetDevice(ctx->par.device));
    for (auto &c : ctx->contigs)
        if (!c.d_ascii) return fail(ctx, MMK_ERR_STATE, "master symbols were released by a single-sample run");
    ctx->multi_mode = true;
    invalidate_plan(ctx);
    ctx->sc.sample = sample_stream & 0xffffu;

    // group mutations per (vcf contig slot, haplotype), file order kept
    const int nC = (int)ctx->contigs.size();
    std::vector<int> slot_of(nC, -1);
    for (int i = 0; i < n_vcf_contigs; ++i) {
        if (vcf_contigs[i] < 0 || vcf_contigs[i] >= nC) return fail(ctx, MMK_ERR_VCF, "VCF contig index out of range");
        slot_of[vcf_contigs[i]] = i;
    }
    std::vector<std::vector<int64_t>> by((size_t)n_vcf_contigs * ploidy);
    std::string alts(alt_bytes ? alt_bytes : "", (size_t)alt_bytes_len);
    bool alt_ambiguous = false;
    for (int64_t i = 0; i < n_mut; ++i) {
        if (contig[i] < 0 || contig[i] >= nC || slot_of[contig[i]] < 0)
            return fail(ctx, MMK_ERR_VCF, "record %lld names a contig absent from the VCF header", (long long)i);
        if (hap[i] < 1 || hap[i] > ploidy) return fail(ctx, MMK_ERR_VCF, "record %lld: haplotype slot %d exceeds ploidy %d", (long long)i, hap[i], ploidy);
        if (alt_off[i] < 0 || alt_len[i] < 0 || alt_off[i] + alt_len[i] > alt_bytes_len) return fail(ctx, MMK_ERR_INVALID_ARG, "alt slice out of range");
        for (int32_t k = 0; k < alt_len[i]; ++k) {
            char &ch = alts[(size_t)alt_off[i] + k];
            if (!valid_iupac(ch)) return fail(ctx, MMK_ERR_BAD_BASE, "ALT allele of record %lld has a symbol outside the DNA alphabet", (long long)i);
            if (!(ch == 'A' || ch == 'C' || ch == 'G' || ch == 'T')) alt_ambiguous = true;
        }
        if (type[i] == 0 && alt_len[i] < 1) return fail(ctx, MMK_ERR_VCF, "empty SNP allele");
        by[(size_t)slot_of[contig[i]] * ploidy + (hap[i] - 1)].push_back(i);
    }
    ctx->bpb = alt_ambiguous ? 4 : 2;
    for (auto &c : ctx->contigs)
        if (c.n_nonACGT) ctx->bpb = 4;

    // per schema: SNP table (last write wins per position) and the piece list after all splices
    struct Build {
        std::vector<int64_t> snp_pos;
        std::string snp_alt;
        std::vector<Piece> pieces;
        int64_t out_len;
    };
    std::vector<Build> builds(by.size());
    ctx->schemas.clear();
    uint64_t woff = 0;
    for (int ci = 0; ci < n_vcf_contigs; ++ci) {
        const Contig &mc = ctx->contigs[vcf_contigs[ci]];
        for (int h = 0; h < ploidy; ++h) {
            Build &b = builds[(size_t)ci * ploidy + h];
            std::vector<int64_t> indels;
            bool sorted_unique = true;  // a VCF is sorted: one SNP per position, ascending
            for (int64_t i : by[(size_t)ci * ploidy + h]) {
                if (type[i] == 0) {
                    if (pos[i] < 1 || pos[i] > mc.len) return fail(ctx, MMK_ERR_BOUNDS, "SNP position %lld outside contig %s", (long long)pos[i], mc.name.c_str());
                    if (!b.snp_pos.empty() && pos[i] <= b.snp_pos.back()) sorted_unique = false;
                    b.snp_pos.push_back(pos[i]);
                    b.snp_alt.push_back(alts[(size_t)alt_off[i]]);
                } else
                    indels.push_back(i);
            }
            if (!sorted_unique) {  // several SNPs on one position: the last in file order wins (seq[pos] = alt, one after the other)
                std::vector<size_t> ord(b.snp_pos.size());
                for (size_t k = 0; k < ord.size(); ++k) ord[k] = k;
                std::stable_sort(ord.begin(), ord.end(), [&](size_t x, size_t y) { return b.snp_pos[x] < b.snp_pos[y]; });
                std::vector<int64_t> up;
                std::string ua;
                for (size_t k = 0; k < ord.size(); ++k) {
                    if (k + 1 < ord.size() && b.snp_pos[ord[k + 1]] == b.snp_pos[ord[k]]) continue;  // a later record overwrites this one
                    up.push_back(b.snp_pos[ord[k]]);
                    ua.push_back(b.snp_alt[ord[k]]);
                }
                b.snp_pos.swap(up);
                b.snp_alt.swap(ua);
            }
            // spliceinto!(seq, pos:pos+len(ref)-1, alt) in REVERSE file order on the current sequence.
            // The list is kept back to front (rev.back() is the FIRST piece).  A VCF is sorted by position, so in
            // reverse order nearly every splice falls inside the still untouched master prefix [0, L0): that is three
            // pushes.  Anything else (unsorted or overlapping records) replays the splice on the whole list.
            std::vector<Piece> rev;
            rev.push_back(Piece{0, 0, mc.len});
            int64_t cur_len = mc.len;
            for (size_t r = indels.size(); r-- > 0;) {
                int64_t i = indels[r];
                int64_t a = pos[i] - 1, e = a + ref_len[i];  // [a, e) 0-based in current coordinates
                if (a < 0 || e > cur_len || ref_len[i] < 0) return fail(ctx, MMK_ERR_BOUNDS, "indel at %lld outside contig %s", (long long)pos[i], mc.name.c_str());
                const Piece head = rev.back();
                if (head.kind == 0 && head.src == 0 && e <= head.len) {
                    rev.pop_back();
                    if (head.len > e) rev.push_back(Piece{0, e, head.len - e});
                    if (alt_len[i] > 0) rev.push_back(Piece{1, alt_off[i], alt_len[i]});
                    rev.push_back(Piece{0, 0, a});  // possibly empty: keeps the "untouched prefix first" form
                } else {
                    std::vector<Piece> np;
                    np.reserve(rev.size() + 2);
                    int64_t o = 0;
                    bool inserted = false;
                    for (size_t k = rev.size(); k-- > 0;) {
                        const Piece &p = rev[k];
                        int64_t lo = o, hi = o + p.len;
                        o = hi;
                        if (hi <= a) {  // entirely before
                            np.push_back(p);
                            continue;
                        }
                        if (lo < a) np.push_back(Piece{p.kind, p.src, a - lo});  // head survives
                        if (!inserted) {
                            if (alt_len[i] > 0) np.push_back(Piece{1, alt_off[i], alt_len[i]});
                            inserted = true;
                        }
                        if (hi > e) {  // tail survives
                            int64_t cut = std::max(e, lo) - lo;
                            np.push_back(Piece{p.kind, p.src + cut, p.len - cut});
                        }
                    }
                    if (!inserted && alt_len[i] > 0) np.push_back(Piece{1, alt_off[i], alt_len[i]});  // a == cur_len
                    rev.assign(np.rbegin(), np.rend());
                }
                cur_len += alt_len[i] - ref_len[i];
            }
            std::vector<Piece> &pc = b.pieces;
            pc.assign(rev.rbegin(), rev.rend());
            b.out_len = cur_len;
            if (cur_len >= (1ll << 31)) return fail(ctx, MMK_ERR_UNSUPPORTED, "haplotype has >= 2^31 bases");
            Schema s;
            s.contig = vcf_contigs[ci];
            s.haplotype = h + 1;  // parse(Int, split(key, "_")[end])
            s.len = cur_len;
            s.required = mc.reads_required;  // inherited from the master (ProcessVariants.jl:181)
            s.word_off = woff;
            s.name = mc.name;
            woff += words_for(cur_len, ctx->bpb);
            ctx->schemas.push_back(s);
        }
    }
    CU(ctx->d_hap.ensure(woff ? woff : 4));
    CU(cudaMemsetAsync(ctx->d_hap.p, 0, woff * 4, ctx->st));
    DBuf<char> &d_alt = ctx->v_alt, &d_tmp = ctx->v_tmp, &d_snp_alt = ctx->v_snp_alt;
    DBuf<int64_t> &d_snp_pos = ctx->v_snp_pos, &d_dst = ctx->v_dst, &d_src = ctx->v_src;
    DBuf<uint8_t> &d_kind = ctx->v_kind;
    CU(d_alt.ensure(alts.size() + 16));
    if (!alts.empty()) CU(cudaMemcpyAsync(d_alt.p, alts.data(), alts.size(), cudaMemcpyHostToDevice, ctx->st));
    ContigStats zero{~0ull, 0, 0};
    CU(cudaMemcpyAsync(ctx->d_stats, &zero, sizeof zero, cudaMemcpyHostToDevice, ctx->st));
    std::vector<int64_t> h_dst, h_src;
    std::vector<uint8_t> h_kind;
    for (size_t si = 0; si < builds.size(); ++si) {
        Build &b = builds[si];
        const Schema &s = ctx->schemas[si];
        const Contig &mc = ctx->contigs[s.contig];
        const char *src_ascii = mc.d_ascii;
        if (!b.snp_pos.empty()) {  // paint SNPs on a private copy of the master
            CU(d_tmp.ensure((size_t)mc.len + 16));
            CU(cudaMemcpyAsync(d_tmp.p, mc.d_ascii, (size_t)mc.len, cudaMemcpyDeviceToDevice, ctx->st));
            CU(d_snp_pos.ensure(b.snp_pos.size()));
            CU(d_snp_alt.ensure(b.snp_alt.size()));
            CU(cudaMemcpyAsync(d_snp_pos.p, b.snp_pos.data(), b.snp_pos.size() * 8, cudaMemcpyHostToDevice, ctx->st));
            CU(cudaMemcpyAsync(d_snp_alt.p, b.snp_alt.data(), b.snp_alt.size(), cudaMemcpyHostToDevice, ctx->st));
            scatter_snps(d_tmp.p, d_snp_pos.p, d_snp_alt.p, (int64_t)b.snp_pos.size(), ctx->st);
            src_ascii = d_tmp.p;
        }
        h_dst.clear();
        h_src.clear();
        h_kind.clear();
        int64_t o = 0;
        for (const Piece &p : b.pieces) {
            if (p.len <= 0) continue;
            h_dst.push_back(o);
            h_src.push_back(p.src);
            h_kind.push_back(p.kind);
            o += p.len;
        }
        if (h_dst.empty()) continue;  // empty haplotype
        CU(d_dst.ensure(h_dst.size()));
        CU(d_src.ensure(h_src.size()));
        CU(d_kind.ensure(h_kind.size()));
        CU(cudaMemcpyAsync(d_dst.p, h_dst.data(), h_dst.size() * 8, cudaMemcpyHostToDevice, ctx->st));
        CU(cudaMemcpyAsync(d_src.p, h_src.data(), h_src.size() * 8, cudaMemcpyHostToDevice, ctx->st));
        CU(cudaMemcpyAsync(d_kind.p, h_kind.data(), h_kind.size(), cudaMemcpyHostToDevice, ctx->st));
        pack_pieces(src_ascii, d_alt.p, d_dst.p, d_src.p, d_kind.p, (int32_t)h_dst.size(), b.out_len, ctx->d_hap.p + s.word_off,
                    ctx->bpb, ctx->d_stats, ctx->st);
        // (copies from pageable host memory return once the source has been staged: the vectors may be reused)
    }
    CU(cudaMemcpyAsync(ctx->h_stats, ctx->d_stats, sizeof zero, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    if (ctx->h_stats->first_bad != ~0ull) return fail(ctx, MMK_ERR_BAD_BASE, "unencodable symbol while packing a haplotype");
    int32_t rc = upload_schemas(ctx);
    if (rc) return rc;
    ctx->schemas_ready = true;
    return MMK_OK;
}

// ---- VCF handle ---------------------------------------------------------------------------
extern "C" int32_t mmk_vcf_open(const char *path, mmk_vcf **out, char *errbuf, size_t errbuf_len) {
    if (!path || !out) return MMK_ERR_INVALID_ARG;
    *out = nullptr;
    int32_t rc = mmk_check_variant_path(path);
    std::string err;
    if (rc) err = "BCF file type is not yet supported";
    mmk_vcf *v = nullptr;
    if (!rc) {
        v = new mmk_vcf();
        rc = mmkio::read_vcf(path, v->d, err);
    }
    if (rc) {
        if (errbuf && errbuf_len) snprintf(errbuf, errbuf_len, "%s", err.c_str());
        delete v;
        return rc;
    }
    *out = v;
    return MMK_OK;
}
extern "C" void mmk_vcf_close(mmk_vcf *v) { delete v; }
extern "C" int32_t mmk_vcf_n_samples(const mmk_vcf *v) { return v ? (int32_t)v->d.samples.size() : 0; }
extern "C" const char *mmk_vcf_sample_name(const mmk_vcf *v, int32_t i) {
    return (v && i >= 0 && i < (int32_t)v->d.samples.size()) ? v->d.samples[i].c_str() : nullptr;
}
extern "C" int32_t mmk_vcf_ploidy(const mmk_vcf *v) { return v ? v->d.ploidy : 0; }
extern "C" int32_t mmk_vcf_n_contigs(const mmk_vcf *v) { return v ? (int32_t)v->d.contigs.size() : 0; }
extern "C" const char *mmk_vcf_contig_name(const mmk_vcf *v, int32_t i) {
    return (v && i >= 0 && i < (int32_t)v->d.contigs.size()) ? v->d.contigs[i].c_str() : nullptr;
}
extern "C" int64_t mmk_vcf_n_records(const mmk_vcf *v) { return v ? (int64_t)v->d.records.size() : 0; }

extern "C" int32_t mmk_apply_vcf_sample(mmk_ctx *ctx, const mmk_vcf *v, int32_t sample) {
    if (!ctx || !v) return fail(ctx, MMK_ERR_INVALID_ARG, "mmk_apply_vcf_sample: bad argument");
    const mmkio::VcfData &d = v->d;
    if (sample < 0 || sample >= (int32_t)d.samples.size()) return fail(ctx, MMK_ERR_INVALID_ARG, "sample index out of range");
    if (d.ploidy < 1) return fail(ctx, MMK_ERR_VCF, "VCF has no records to take the ploidy from");
    std::unordered_map<std::string, int> master;
    for (size_t i = 0; i < ctx->contigs.size(); ++i)
        if (!master.count(ctx->contigs[i].name)) master[ctx->contigs[i].name] = (int)i;
    std::vector<int32_t> vc;
    std::unordered_map<std::string, int> in_header;
    for (auto &n : d.contigs) {
        auto it = master.find(n);
        if (it == master.end()) return fail(ctx, MMK_ERR_VCF, "VCF header contig %s is not in the FASTA", n.c_str());
        in_header[n] = it->second;
        vc.push_back(it->second);
    }
    std::vector<int32_t> contig, ref_len, alt_len;
    std::vector<int64_t> pos, alt_off;
    std::vector<uint8_t> type, hap;
    std::string alt_bytes;
    for (auto &r : d.records) {
        size_t off = 0;
        for (int s = 0; s < sample; ++s) off += r.ploidy_of[s];
        int np = r.ploidy_of[sample];
        for (int a = 0; a < np; ++a) {
            uint8_t al = r.alleles[off + a];
            if (!al) continue;
            auto it = in_header.find(r.chrom);
            if (it == in_header.end()) return fail(ctx, MMK_ERR_VCF, "record contig %s has no ##contig header line", r.chrom.c_str());
            if (a + 1 > d.ploidy) return fail(ctx, MMK_ERR_VCF, "genotype at %s:%lld has more alleles than the ploidy", r.chrom.c_str(), (long long)r.pos);
            const std::string &alt = r.alts[al - 1];
            contig.push_back(it->second);
            pos.push_back(r.pos);
            type.push_back((uint8_t)r.type);
            ref_len.push_back((int32_t)r.ref.size());
            alt_off.push_back((int64_t)alt_bytes.size());
            alt_len.push_back((int32_t)alt.size());
            hap.push_back((uint8_t)(a + 1));
            alt_bytes += alt;
        }
    }
    return mmk_set_sample_variants(ctx, (uint32_t)sample, d.ploidy, vc.data(), (int32_t)vc.size(), (int64_t)pos.size(),
                                   contig.data(), pos.data(), type.data(), ref_len.data(), alt_off.data(), alt_len.data(),
                                   hap.data(), alt_bytes.data(), (int64_t)alt_bytes.size());
}

// ---- schema queries ---------------------------------------------------------------------------
extern "C" int32_t mmk_n_schemas(const mmk_ctx *c) {
    if (!c) return 0;
    if (!c->multi_mode && !c->schemas_ready) return (int32_t)c->contigs.size();
    return (int32_t)c->schemas.size();
}
extern "C" int32_t mmk_schema_get(const mmk_ctx *c, int32_t i, mmk_schema_info *out) {
    if (!c || !out) return MMK_ERR_INVALID_ARG;
    memset(out, 0, sizeof *out);
    if (!c->multi_mode && !c->schemas_ready) {
        if (i < 0 || i >= (int32_t)c->contigs.size()) return MMK_ERR_INVALID_ARG;
        const Contig &k = c->contigs[i];
        out->length = k.len;
        out->reads_required = k.reads_required;
        out->haplotype = k.haplotype;
        out->contig = i;
        snprintf(out->name, sizeof out->name, "%s", k.name.c_str());
        return MMK_OK;
    }
    if (i < 0 || i >= (int32_t)c->schemas.size()) return MMK_ERR_INVALID_ARG;
    const Schema &s = c->schemas[i];
    out->length = s.len;
    out->reads_required = s.required;
    out->reads_current = s.current;
    out->haplotype = s.haplotype;
    out->contig = s.contig;
    snprintf(out->name, sizeof out->name, "%s", s.name.c_str());
    return MMK_OK;
}
extern "C" int32_t mmk_schema_name(const mmk_ctx *c, int32_t i, char *out, int64_t cap) {
    if (!c || !out || cap < 1) return MMK_ERR_INVALID_ARG;
    const std::string *name = nullptr;
    if (!c->multi_mode && !c->schemas_ready) {
        if (i < 0 || i >= (int32_t)c->contigs.size()) return MMK_ERR_INVALID_ARG;
        name = &c->contigs[i].name;
    } else {
        if (i < 0 || i >= (int32_t)c->schemas.size()) return MMK_ERR_INVALID_ARG;
        name = &c->schemas[i].name;
    }
    if ((int64_t)name->size() + 1 > cap) return MMK_ERR_INVALID_ARG;
    memcpy(out, name->c_str(), name->size() + 1);
    return MMK_OK;
}
extern "C" int32_t mmk_schema_bases(mmk_ctx *ctx, int32_t i, char *out, int64_t cap) {
    if (!ctx || !out) return fail(ctx, MMK_ERR_INVALID_ARG, "mmk_schema_bases: bad argument");
    CU(cudaSetDevice(ctx->par.device));
    if (!ctx->multi_mode) {
        int32_t rc = finalize_single(ctx);
        if (rc) return rc;
    }
    if (!ctx->schemas_ready) return fail(ctx, MMK_ERR_STATE, "schemas not built");
    if (i < 0 || i >= (int32_t)ctx->schemas.size()) return fail(ctx, MMK_ERR_INVALID_ARG, "schema index out of range");
    const Schema &s = ctx->schemas[i];
    if (cap < s.len) return fail(ctx, MMK_ERR_INVALID_ARG, "buffer too small");
    DBuf<char> tmp;
    CU(tmp.ensure((size_t)s.len + 1));
    unpack(ctx->d_hap.p + s.word_off, s.len, ctx->bpb, tmp.p, ctx->st);
    CU(cudaMemcpyAsync(out, tmp.p, (size_t)s.len, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    return MMK_OK;
}

// ------------------------------------------------------------------------------------------
// planning: the barcode-group loop of Main.jl:44-59 in speculative batches
// ------------------------------------------------------------------------------------------
static PlanDev plan_dev(mmk_ctx *c) {
    PlanDev p;
    p.group = c->p_group.p;
    p.slot = c->p_slot.p;
    p.schema = c->p_schema.p;
    p.epoch = c->p_epoch.p;
    p.size = c->p_size.p;
    p.nfrags = c->p_nfrags.p;
    p.status = c->p_status.p;
    return p;
}
static int32_t plan_reserve(mmk_ctx *ctx, uint64_t n, uint64_t keep) {
    CU(ctx->p_group.ensure(n, keep, ctx->st));
    CU(ctx->p_slot.ensure(n, keep, ctx->st));
    CU(ctx->p_schema.ensure(n, keep, ctx->st));
    CU(ctx->p_epoch.ensure(n, keep, ctx->st));
    CU(ctx->p_size.ensure(n, keep, ctx->st));
    CU(ctx->p_nfrags.ensure(n, keep, ctx->st));
    CU(ctx->p_status.ensure(n, keep, ctx->st));
    return MMK_OK;
}

static inline double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// grow the group table to cover groups [0, g_need) and the plan arrays to hold their molecules
static int32_t ensure_group_table(mmk_ctx *ctx, uint64_t g_need, uint64_t g_keep, uint64_t keep_mol) {
    if (g_need <= ctx->g_tab) return MMK_OK;
    uint64_t g_new = std::max<uint64_t>(g_need, ctx->g_tab + ctx->g_tab / 2);
    g_new = (g_new + 1023) & ~1023ull;
    CU(ctx->d_gnmol.ensure(g_new, ctx->g_tab, ctx->st));
    CU(ctx->d_goff.ensure(g_new + 1, ctx->g_tab + 1, ctx->st));
    CU(ctx->d_scan_tmp.ensure((g_new - ctx->g_tab) / 1024 + 8));
    plan_extend_table(ctx->sc, ctx->g_tab, g_new, ctx->d_gnmol.p, ctx->d_goff.p, ctx->d_scan_tmp.p, ctx->st);
    uint64_t total = 0;
    CU(cudaMemcpyAsync(&total, ctx->d_goff.p + g_new, 8, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    CU(cudaGetLastError());
    ctx->g_tab = g_new;
    int32_t rc = plan_reserve(ctx, total + 1, keep_mol);
    if (rc) return rc;
    // a reallocation keeps only the committed molecules [0, keep_mol) = groups [0, g_keep): label the rest again
    plan_fill_slots(g_keep, g_new, ctx->d_gnmol.p, ctx->d_goff.p, plan_dev(ctx), ctx->st);
    return MMK_OK;
}

extern "C" int32_t mmk_plan(mmk_ctx *ctx, mmk_plan_info *out) {
    if (!ctx) return MMK_ERR_INVALID_ARG;
    const double t_begin = now_ms();
    if (!ctx->st) return fail(ctx, MMK_ERR_STATE, "context was not created successfully");
    CU(cudaSetDevice(ctx->par.device));
    if (!ctx->multi_mode) {
        int32_t rc = finalize_single(ctx);
        if (rc) return rc;
    }
    if (!ctx->schemas_ready) return fail(ctx, MMK_ERR_STATE, "schemas not built");
    const uint32_t K = (uint32_t)ctx->schemas.size();
    if (K == 0) return fail(ctx, MMK_ERR_STATE, "no schemas");
    for (auto &s : ctx->schemas) s.current = 0;
    ctx->placed = false;
    ctx->planned = false;
    ctx->tile_cursor = 0;
    ctx->epoch_nlive.clear();
    ctx->n_mol = 0;

    const uint32_t max_nmol = ctx->sc.nmol_fixed > 0 ? (uint32_t)ctx->sc.nmol_fixed
                                                     : (uint32_t)nearbyint(2.5 * (double)ctx->par.n_molecules) + 1;
    uint64_t Bmax = ctx->par.plan_batch_groups > 0 ? (uint64_t)ctx->par.plan_batch_groups : (1ull << 20);
    {   // the insert histogram holds one row per live schema and one column per 512 molecules of a batch: bound it to 256 MB
        const uint64_t max_cols = std::max<uint64_t>((256ull << 20) / 8 / K, 64);
        Bmax = std::max<uint64_t>(1, std::min<uint64_t>(Bmax, max_cols * 512 / max_nmol));
    }
    const uint64_t Bmin = std::min<uint64_t>(Bmax, 8192);
    const uint32_t hstride = plan_hist_chunks(Bmax * max_nmol) + 1;

    // control block, trackers, scratch
    CU(ctx->d_live.ensure(K));
    CU(ctx->d_lpos.ensure(65536));
    CU(ctx->d_need.ensure(K));
    CU(ctx->d_required.ensure(K));
    CU(ctx->d_current.ensure(K));
    CU(ctx->d_sums.ensure(K));
    CU(ctx->d_epoch_nlive.ensure(65536));
    CU(ctx->d_first_err.ensure(1));
    CU(ctx->d_pstate.ensure(1));
    CU(ctx->d_hist.ensure((size_t)hstride * K));
    if (!ctx->h_pstate) CU(cudaMallocHost((void **)&ctx->h_pstate, sizeof(PlanState)));
    CU(cudaMemsetAsync(ctx->d_hist.p, 0, (size_t)hstride * K * sizeof(unsigned long long), ctx->st));
    CU(cudaMemsetAsync(ctx->d_first_err.p, 0xff, 8, ctx->st));
    std::vector<int64_t> required(K);
    for (uint32_t i = 0; i < K; ++i) required[i] = ctx->schemas[i].required;
    CU(cudaMemcpyAsync(ctx->d_required.p, required.data(), (size_t)K * 8, cudaMemcpyHostToDevice, ctx->st));
    PlanArrays pa;
    pa.live = ctx->d_live.p;
    pa.lpos = ctx->d_lpos.p;
    pa.required = ctx->d_required.p;
    pa.current = ctx->d_current.p;
    pa.need = ctx->d_need.p;
    pa.sums = ctx->d_sums.p;
    pa.hist = ctx->d_hist.p;
    pa.epoch_nlive = ctx->d_epoch_nlive.p;

    // the group table is rebuilt by every call (no state is carried between runs)
    ctx->g_tab = 0;
    CU(ctx->d_goff.ensure(1));
    CU(cudaMemsetAsync(ctx->d_goff.p, 0, 8, ctx->st));
    {
        int32_t rc = ensure_group_table(ctx, Bmin, 0, 0);
        if (rc) return rc;
    }
    PlanState &hs = *ctx->h_pstate;
    memset(&hs, 0, sizeof hs);
    hs.B = Bmin;
    hs.Bmin = Bmin;
    hs.Bmax = Bmax;
    hs.g_tab = ctx->g_tab;
    hs.barcode_cap = barcode_capacity(ctx->par.bc_format);
    hs.cut = ~0ull;
    hs.n_live = K;
    hs.hstride = hstride;
    CU(cudaMemcpyAsync(ctx->d_pstate.p, &hs, sizeof hs, cudaMemcpyHostToDevice, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));  // `required` and the staged state are host temporaries
    plan_init(ctx->d_pstate.p, pa, K, ctx->st);

    // rounds of batches: the first batch alone (it measures the insert rate that sizes the group table), then 16 at a time
    int round_batches = 1;
    for (int round = 0;; ++round) {
        if (round > 100000) return fail(ctx, MMK_ERR_CUDA, "planner does not terminate (internal error)");
        plan_batches(ctx->sc, ctx->d_sch.p, ctx->d_pstate.p, pa, ctx->d_goff.p, plan_dev(ctx), ctx->d_first_err.p, round_batches, ctx->st);
        CU(cudaMemcpyAsync(&hs, ctx->d_pstate.p, sizeof hs, cudaMemcpyDeviceToHost, ctx->st));
        CU(cudaStreamSynchronize(ctx->st));
        CU(cudaGetLastError());
        if (hs.error_code)
            return fail(ctx, (int32_t)hs.error_code, hs.error_code == MMK_ERR_BARCODES_EXHAUSTED
                                                         ? "There are no more barcodes available for the selected barcode type (molecule %llu)."
                                                         : "molecule %llu: %s",
                        (unsigned long long)hs.error_mol, mmk_strerror((int32_t)hs.error_code));
        if (hs.done) break;
        // table for the expected rest of the run in one extension (5 % margin); a run that outlives it asks again
        uint64_t want = hs.g0 + hs.B;
        if (hs.rate > 0.0) want = std::max<uint64_t>(want, hs.g0 + (uint64_t)((double)hs.remaining_need / hs.rate * 1.05) + 4096);
        if (hs.need_table || want > ctx->g_tab) {
            int32_t rc = ensure_group_table(ctx, want, hs.g0, hs.n_mol);
            if (rc) return rc;
            uint64_t patch[1] = {ctx->g_tab};
            uint32_t zero = 0;
            CU(cudaMemcpyAsync((char *)ctx->d_pstate.p + offsetof(PlanState, g_tab), patch, 8, cudaMemcpyHostToDevice, ctx->st));
            CU(cudaMemcpyAsync((char *)ctx->d_pstate.p + offsetof(PlanState, need_table), &zero, 4, cudaMemcpyHostToDevice, ctx->st));
            if (hs.need_table) {  // the device stopped with nothing left to plan: give the next batch its size back
                uint64_t nb = std::min<uint64_t>(std::max<uint64_t>(hs.B, Bmin), ctx->g_tab - hs.g0);
                CU(cudaMemcpyAsync((char *)ctx->d_pstate.p + offsetof(PlanState, B), &nb, 8, cudaMemcpyHostToDevice, ctx->st));
            }
            CU(cudaStreamSynchronize(ctx->st));
        }
        round_batches = 16;
    }
    const uint64_t n_mol = hs.n_mol;
    ctx->n_mol = n_mol;
    ctx->n_groups = hs.g0;
    {   // trackers and the live-set size of every epoch back to the host
        std::vector<int64_t> cur(K);
        ctx->epoch_nlive.resize((size_t)hs.epoch + 1);
        CU(cudaMemcpyAsync(cur.data(), ctx->d_current.p, (size_t)K * 8, cudaMemcpyDeviceToHost, ctx->st));
        CU(cudaMemcpyAsync(ctx->epoch_nlive.data(), ctx->d_epoch_nlive.p, ctx->epoch_nlive.size() * 4, cudaMemcpyDeviceToHost, ctx->st));
        CU(cudaStreamSynchronize(ctx->st));
        for (uint32_t i = 0; i < K; ++i) ctx->schemas[i].current = cur[i];
    }
    // first pair of every molecule
    CU(ctx->d_read_off.ensure(n_mol + 1));
    CU(ctx->d_scan_tmp.ensure((size_t)n_mol / 1024 + 8));
    exclusive_scan_u32(ctx->p_nfrags.p, ctx->d_read_off.p, n_mol, ctx->d_scan_tmp.p, ctx->st);
    uint64_t total = 0;
    CU(cudaMemcpyAsync(&total, ctx->d_read_off.p + n_mol, 8, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    ctx->n_pairs = total;
    ctx->planned = true;
    ctx->plan_info.n_groups = (int64_t)hs.g0;
    ctx->plan_info.n_molecules = (int64_t)n_mol;
    ctx->plan_info.n_pairs = (int64_t)total;
    ctx->plan_info.n_epochs = (int64_t)hs.n_epochs;
    ctx->plan_info.n_batches = (int64_t)hs.n_batches;
    ctx->plan_info.host_ms = now_ms() - t_begin;
    if (out) *out = ctx->plan_info;
    return MMK_OK;
}

extern "C" int32_t mmk_set_shard(mmk_ctx *ctx, int32_t rank, int32_t world) {
    if (!ctx || world < 1 || world > 1023 || rank < 0 || rank >= world) return fail(ctx, MMK_ERR_INVALID_ARG, "bad shard");
    ctx->rank = rank;
    ctx->world = world;
    ctx->placed = false;
    ctx->tile_cursor = 0;
    return MMK_OK;
}

// ------------------------------------------------------------------------------------------
// placement + tile offsets for this shard
// ------------------------------------------------------------------------------------------
static const bool PLACE_BLOCK_DEFAULT = true;  // MMK_PLACE_MODE=block|legacy overrides (tests run both)

static EmitArgs base_args(mmk_ctx *c) {
    EmitArgs a;
    memset(&a, 0, sizeof a);
    a.hap = c->d_hap.p;
    a.sch = c->d_sch.p;
    a.read_off = c->d_roff_l.p;
    a.mol = c->d_mol.p;
    a.text = c->d_text.p;
    a.ins = c->d_ins.p;
    a.tile_mol0 = c->d_tile_mol0.p;
    a.fdesc = c->d_fdesc.p;
    a.tile_flags = c->d_tile_flags.p;
    a.stage_cap = c->stage_cap;
    a.fast_ok = c->fast_ok;
    a.qfill = c->d_qfill.p;
    a.err_on = c->err_on ? 1 : 0;
    a.err_tins = c->err_tins;
    a.err_tdel = c->err_tdel;
    a.err_thr = c->d_err_thr.p;
    a.err_qb = c->d_err_qb.p;
    a.seed = c->sc.seed;
    a.sample = c->sc.sample;
    a.pair_base = c->pair0;
    a.tile_off1 = c->d_tile_off1.p;
    a.tile_off2 = c->d_tile_off2.p;
    a.n_pairs = c->s_npairs;
    a.n_mol = (uint32_t)c->s_nmol;
    a.tile_pairs = c->tile_pairs;
    a.text_stride = c->text_stride;
    a.r1 = c->par.length_r1;
    a.r2 = c->par.length_r2;
    a.fq_format = c->par.fq_format;
    a.bclen_inline = c->par.fq_format == MMK_FMT_TENX ? 16 : 0;
    a.strict = c->par.strict_wrap ? 1 : 0;
    a.bpb = c->bpb;
    a.force_generic = c->par.force_generic;
    return a;
}

static int32_t prepare_chunks(mmk_ctx *ctx);

extern "C" int32_t mmk_place(mmk_ctx *ctx, mmk_place_info *out) {
    if (!ctx) return MMK_ERR_INVALID_ARG;
    if (!ctx->planned) return fail(ctx, MMK_ERR_STATE, "mmk_plan has not run");
    CU(cudaSetDevice(ctx->par.device));
    const double t_begin = now_ms();
    double t_phase = t_begin;
    // shard bounds
    DBuf<uint64_t> &d_bounds = ctx->s_bounds;
    CU(d_bounds.ensure((size_t)std::max(ctx->world, 1024) + 1));
    shard_bounds(ctx->d_read_off.p, ctx->n_mol, ctx->n_pairs, ctx->world, d_bounds.p, ctx->st);
    std::vector<uint64_t> bounds((size_t)ctx->world + 1);
    CU(cudaMemcpyAsync(bounds.data(), d_bounds.p, bounds.size() * 8, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    ctx->mol0 = bounds[ctx->rank];
    ctx->s_nmol = bounds[ctx->rank + 1] - bounds[ctx->rank];
    if (ctx->s_nmol >= (1ull << 32) - 2) return fail(ctx, MMK_ERR_UNSUPPORTED, "more than 2^32 molecules in one shard");
    uint64_t ro[2] = {0, 0};
    CU(cudaMemcpyAsync(&ro[0], ctx->d_read_off.p + ctx->mol0, 8, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaMemcpyAsync(&ro[1], ctx->d_read_off.p + ctx->mol0 + ctx->s_nmol, 8, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    ctx->pair0 = ro[0];
    ctx->s_npairs = ro[1] - ro[0];

    // header text stride
    size_t max_name = 0;
    int max_hap = 1;
    for (auto &s : ctx->schemas) {
        max_name = std::max(max_name, s.name.size());
        max_hap = std::max(max_hap, s.haplotype);
    }
    size_t hap_digits = std::to_string(max_hap).size();
    size_t max_prefix = 1 + max_name + 1 + hap_digits + 1 + 10 + 1 + 10 + 1;
    if (max_prefix > 200) return fail(ctx, MMK_ERR_UNSUPPORTED, "header prefix longer than 200 bytes");
    ctx->text_stride = (uint32_t)((max_prefix + 36 + 15) & ~15ull);
    {   // tile geometry: the fast path stages one tile of R1 and of R2 text in shared memory
        const mmk_params &q = ctx->par;
        uint32_t inl = q.fq_format == MMK_FMT_TENX ? 16u : 0u;
        uint32_t max_hdr = (uint32_t)max_prefix + 10 + 1 + 10 + 34;
        uint32_t max_rec = max_hdr + inl + (uint32_t)std::max(q.length_r1, q.length_r2) * 2 + inl + 4;
        ctx->fast_ok = q.length_r1 >= 16 && q.length_r1 <= 192 && q.length_r2 >= 16 && q.length_r2 <= 192 && !q.force_generic &&
                       ctx->d_hap.cap < (1ull << 32);  // the fast path indexes the arena with 32-bit word offsets
        uint32_t tp = 32;
        if (const char *e = getenv("MMK_TILE_PAIRS")) tp = (uint32_t)atoi(e);
        tp = std::max(8u, std::min(64u, tp));
        while (tp & (tp - 1)) tp &= tp - 1;  // power of two: kernels turn pair index → tile with a shift
        while (tp > 8 && (uint64_t)tp * max_rec + 32 > 40 * 1024) tp >>= 1;
        ctx->tile_pairs = tp;
        ctx->tile_bytes_bound = tp * max_rec;
        ctx->bz_ready = false;
        ctx->stage_cap = (uint32_t)(((uint64_t)tp * max_rec + 32 + 15) & ~15ull);
        if (!ctx->fast_ok) ctx->stage_cap = 16;
        CU(ctx->d_qfill.ensure(2 * (size_t)ctx->stage_cap));
        CU(cudaMemsetAsync(ctx->d_qfill.p, 'I', 2 * (size_t)ctx->stage_cap, ctx->st));
    }

    CU(ctx->d_mol.ensure(ctx->s_nmol + 1));
    CU(ctx->d_text.ensure((size_t)(ctx->s_nmol + 1) * ctx->text_stride));
    CU(ctx->d_roff_l.ensure(ctx->s_nmol + 2));
    DBuf<unsigned long long> &d_err = ctx->s_err;
    CU(d_err.ensure(3));
    unsigned long long init[3] = {~0ull, 0, 0};
    CU(cudaMemcpyAsync(d_err.p, init, sizeof init, cudaMemcpyHostToDevice, ctx->st));
    ctx->n_tiles = (uint32_t)((ctx->s_npairs + ctx->tile_pairs - 1) / ctx->tile_pairs);
    CU(ctx->d_tile_mol0.ensure((size_t)ctx->n_tiles + 1));
    // Block placement (mmk_place.cuh) needs 32-pair tiles and insert lengths that fit 16 bits; MMK_PLACE_MODE=legacy keeps the
    // three-kernel path (draws → chain → tile preparation through global scratch), which also serves the other cases.
    const bool block_ok = ctx->tile_pairs == 32 && place_block_supports(ctx->sc);
    bool fused = block_ok && PLACE_BLOCK_DEFAULT;
    if (const char *e = getenv("MMK_PLACE_MODE")) fused = block_ok && strcmp(e, "legacy") != 0;
    if (fused) place_local_roff(ctx->mol0, (uint32_t)ctx->s_nmol, ctx->d_read_off.p, ctx->d_roff_l.p, ctx->st);
    else place_prepare(ctx->mol0, (uint32_t)ctx->s_nmol, ctx->d_read_off.p, ctx->d_roff_l.p, ctx->d_tile_mol0.p, ctx->tile_pairs, ctx->st);
    // per-tile outputs of the placement, launched per segment below
    DBuf<uint32_t> &d_sum1 = ctx->s_sum1, &d_sum2 = ctx->s_sum2;
    CU(d_sum1.ensure((size_t)ctx->n_tiles + 1));
    CU(d_sum2.ensure((size_t)ctx->n_tiles + 1));
    CU(ctx->d_scan_tmp.ensure((size_t)ctx->n_tiles / 1024 + 8));
    CU(ctx->d_tile_off1.ensure((size_t)ctx->n_tiles + 1));
    CU(ctx->d_tile_off2.ensure((size_t)ctx->n_tiles + 1));
    DBuf<unsigned long long> &d_gather = ctx->s_gather;
    CU(d_gather.ensure(256));
    CU(cudaMemsetAsync(d_gather.p, 0, 256 * 8, ctx->st));
    CU(ctx->d_fdesc.ensure(ctx->s_npairs + 1));
    CU(ctx->d_tile_flags.ensure((size_t)ctx->n_tiles + 1));
    if (!fused) CU(ctx->d_ins.ensure(ctx->s_npairs + 1));  // the three-kernel path writes the insert plan; before base_args takes the pointer
    const EmitArgs targs = base_args(ctx);
    bool text_pending = false;  // block placement: k_mol_text is in flight on the second stream (ev_chain[0] marks its end)
    auto run_fused = [&]() -> int32_t {
        const uint32_t nm = (uint32_t)ctx->s_nmol;
        CU(ctx->d_weight.ensure((size_t)nm + 1));
        CU(ctx->d_woff.ensure((size_t)nm + 2));
        CU(ctx->d_scan_tmp.ensure((size_t)std::max<uint64_t>(nm, ctx->n_tiles) / 1024 + 8));
        CU(ctx->d_blk.ensure((size_t)place_block_count(ctx->s_npairs, nm) + 2));
        CU(ctx->d_big_list.ensure((size_t)place_big_cap(ctx->s_npairs) + 1));
        CU(ctx->d_big_count.ensure(1));
        place_block(ctx->sc, ctx->d_sch.p, plan_dev(ctx), ctx->mol0, nm, ctx->s_npairs, ctx->n_tiles, ctx->d_roff_l.p,
                    ctx->d_tile_mol0.p, ctx->d_instab.p, ctx->d_mol.p, ctx->d_text.p, ctx->text_stride, ctx->d_fdesc.p,
                    d_sum1.p, d_sum2.p, ctx->d_tile_flags.p, d_gather.p, d_err.p, ctx->d_weight.p, ctx->d_woff.p, ctx->d_scan_tmp.p,
                    ctx->d_blk.p, ctx->d_big_list.p, ctx->d_big_count.p, ctx->stage_cap, ctx->bpb, ctx->st, ctx->st_copy, ctx->ev_draw[0],
                    ctx->ev_chain[0]);
        text_pending = ctx->s_nmol != 0 && ctx->s_npairs != 0;
        return MMK_OK;
    };
    auto run_legacy = [&]() -> int32_t
    {   // Segments of at most SEG pairs bound the scratch of the flat per-pair draws.  The draws of segment k + 1
        // (compute-bound) run on the main stream while the chain of segment k (latency-bound) runs on the second
        // stream; two scratch buffers alternate.
        uint64_t SEG = 100ull << 20;  // larger segments have shorter tails in the chain kernel; 100 M pairs = 1.6 GB of scratch each
        if (const char *e = getenv("MMK_PLACE_SEG_MPAIRS")) SEG = (uint64_t)std::max(1, atoi(e)) << 20;
        const bool overlap = !(getenv("MMK_PLACE_OVERLAP") && atoi(getenv("MMK_PLACE_OVERLAP")) == 0);
        const int n_seg = (int)std::max<uint64_t>(1, (ctx->s_npairs + SEG - 1) / SEG);
        if (n_seg > 1023) return fail(ctx, MMK_ERR_UNSUPPORTED, "shard too large: more than 2^34 read pairs");
        std::vector<uint64_t> sb((size_t)n_seg + 1), sp((size_t)n_seg + 1);
        if (ctx->s_nmol) {
            shard_bounds(ctx->d_roff_l.p, ctx->s_nmol, ctx->s_npairs, n_seg, d_bounds.p, ctx->st);
            if ((size_t)n_seg + 1 > d_bounds.cap) return fail(ctx, MMK_ERR_NOMEM, "segment table");
            CU(cudaMemcpyAsync(sb.data(), d_bounds.p, sb.size() * 8, cudaMemcpyDeviceToHost, ctx->st));
            CU(cudaStreamSynchronize(ctx->st));
            uint64_t max_pairs = 0;
            for (int k = 0; k <= n_seg; ++k)
                CU(cudaMemcpyAsync(&sp[k], ctx->d_roff_l.p + sb[k], 8, cudaMemcpyDeviceToHost, ctx->st));
            CU(cudaStreamSynchronize(ctx->st));
            for (int k = 0; k < n_seg; ++k) max_pairs = std::max(max_pairs, sp[k + 1] - sp[k]);
            CU(ctx->d_tmp.ensure(place_tmp_bytes(max_pairs)));
            if (overlap && n_seg > 1) CU(ctx->d_tmp2.ensure(place_tmp_bytes(max_pairs)));
            CU(ctx->d_acc_i.ensure(ctx->s_nmol + 1));
            CU(cudaMemsetAsync(ctx->d_acc_i.p, 0, (ctx->s_nmol + 1) * 4, ctx->st));
            const bool two = overlap && n_seg > 1;
            cudaStream_t s_draw = ctx->st, s_chain = two ? ctx->st_copy : ctx->st;
            if (two) {  // the second stream starts after everything queued so far (error slot, prepare kernels)
                CU(cudaEventRecord(ctx->ev_draw[0], ctx->st));
                CU(cudaStreamWaitEvent(s_chain, ctx->ev_draw[0], 0));
            }
            // tiles whose pairs are all placed: record lengths, descriptors
            uint32_t tiles_done = 0;
            auto prepare_upto = [&](int k_done, cudaStream_t st) {  // segments 0..k_done are placed before this point of `st`
                const uint32_t t_hi = k_done + 1 == n_seg ? ctx->n_tiles : (uint32_t)(sp[k_done + 1] / ctx->tile_pairs);
                tile_prepare(targs, tiles_done, t_hi, d_sum1.p, d_sum2.p, ctx->d_fdesc.p, ctx->d_tile_flags.p, d_gather.p, st);
                tiles_done = std::max(tiles_done, t_hi);
            };
            for (int k = 0; k < n_seg; ++k) {
                const int b = two ? (k & 1) : 0;
                void *tmp = b ? (void *)ctx->d_tmp2.p : (void *)ctx->d_tmp.p;
                if (two && k >= 2) CU(cudaStreamWaitEvent(s_draw, ctx->ev_chain[b], 0));  // chain k - 2 done: scratch b free again
                place_segment_draws(ctx->sc, plan_dev(ctx), ctx->mol0, (uint32_t)ctx->s_nmol, sp[k], sp[k + 1], ctx->d_roff_l.p,
                                    ctx->d_tile_mol0.p, ctx->n_tiles, ctx->tile_pairs, tmp, ctx->d_acc_i.p, ctx->d_instab.p, s_draw);
                if (two) {
                    CU(cudaEventRecord(ctx->ev_draw[b], s_draw));
                    CU(cudaStreamWaitEvent(s_chain, ctx->ev_draw[b], 0));
                }
                place_segment_chain(ctx->sc, ctx->d_sch.p, plan_dev(ctx), ctx->mol0, (uint32_t)sb[k], (uint32_t)sb[k + 1], sp[k],
                                    ctx->d_roff_l.p, tmp, ctx->d_acc_i.p, ctx->d_ins.p, ctx->d_mol.p, ctx->d_text.p,
                                    ctx->text_stride, d_err.p, s_chain);
                if (two) {
                    CU(cudaEventRecord(ctx->ev_chain[b], s_chain));
                    prepare_upto(k, s_chain);  // memory-bound; behind the chain it overlaps the next segment's draws
                }
            }
            if (two) {  // join: the main stream continues after everything queued on the second stream
                CU(cudaEventRecord(ctx->ev_draw[0], s_chain));
                CU(cudaStreamWaitEvent(ctx->st, ctx->ev_draw[0], 0));
            }
            prepare_upto(n_seg - 1, ctx->st);  // what is left (everything when there is one stream)
        }
        return MMK_OK;
    };
    unsigned long long herr[3];
    for (;;) {
        int32_t rc = fused ? run_fused() : run_legacy();
        if (rc) return rc;
        ctx->ins_valid = !fused;
        CU(cudaMemcpyAsync(herr, d_err.p, sizeof herr, cudaMemcpyDeviceToHost, ctx->st));
        CU(cudaStreamSynchronize(ctx->st));
        CU(cudaGetLastError());
        break;
    }
    if (herr[0] != ~0ull) {
        if (text_pending) cudaStreamWaitEvent(ctx->st, ctx->ev_chain[0], 0);
        uint8_t code = 0;
        CU(cudaMemcpy(&code, ctx->p_status.p + herr[0], 1, cudaMemcpyDeviceToHost));
        return fail(ctx, code, "molecule %llu: %s", herr[0], mmk_strerror(code));
    }
    ctx->place_info.phase_ms[0] = now_ms() - t_phase;
    t_phase = now_ms();
    // tile offsets (the per-tile record lengths were computed segment by segment above)
    DBuf<uint64_t> &d_tmp = ctx->d_scan_tmp;
    exclusive_scan_u32(d_sum1.p, ctx->d_tile_off1.p, ctx->n_tiles, d_tmp.p, ctx->st);
    exclusive_scan_u32(d_sum2.p, ctx->d_tile_off2.p, ctx->n_tiles, d_tmp.p, ctx->st);
    uint64_t tot[3] = {0, 0, 0};
    CU(cudaMemcpyAsync(&tot[0], ctx->d_tile_off1.p + ctx->n_tiles, 8, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaMemcpyAsync(&tot[1], ctx->d_tile_off2.p + ctx->n_tiles, 8, cudaMemcpyDeviceToHost, ctx->st));
    unsigned long long h_gather[256];
    CU(cudaMemcpyAsync(h_gather, d_gather.p, sizeof h_gather, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    for (int k = 0; k < 256; ++k) tot[2] += h_gather[k];
    CU(cudaGetLastError());
    ctx->place_info.first_molecule = (int64_t)ctx->mol0;
    ctx->place_info.n_molecules = (int64_t)ctx->s_nmol;
    ctx->place_info.first_pair = (int64_t)ctx->pair0;
    ctx->place_info.n_pairs = (int64_t)ctx->s_npairs;
    ctx->place_info.r1_bytes = tot[0];
    ctx->place_info.r2_bytes = tot[1];
    ctx->place_info.gather_bytes = tot[2];
    ctx->tile_cursor = 0;
    ctx->slot = 0;
    ctx->place_info.phase_ms[1] = now_ms() - t_phase;
    t_phase = now_ms();
    int32_t rc_chunks = prepare_chunks(ctx);
    if (text_pending) CU(cudaStreamWaitEvent(ctx->st, ctx->ev_chain[0], 0));  // whatever follows on the main stream sees the header text
    if (rc_chunks) return rc_chunks;
    ctx->place_info.phase_ms[2] = now_ms() - t_phase;
    ctx->place_info.host_ms = now_ms() - t_begin;
    ctx->placed = true;
    if (out) *out = ctx->place_info;
    return MMK_OK;
}

// ------------------------------------------------------------------------------------------
// emission
// ------------------------------------------------------------------------------------------
static uint64_t wanted_slot_bytes(mmk_ctx *ctx) {
    uint64_t want = ctx->par.max_chunk_bytes ? ctx->par.max_chunk_bytes : (8ull << 30);
    uint64_t need = std::max(ctx->place_info.r1_bytes, ctx->place_info.r2_bytes);
    want = std::min(want, need);
    want = std::max<uint64_t>(want, 1 << 20);
    return (want + 255) & ~255ull;
}

// output ring + chunk table for the placed shard
static int32_t prepare_chunks(mmk_ctx *ctx) {
    uint64_t want = wanted_slot_bytes(ctx);
    uint64_t need = std::max(ctx->place_info.r1_bytes, ctx->place_info.r2_bytes);
    int slots = need <= want ? 1 : 2;  // a single chunk needs no ring
    bool have = ctx->out_cap >= want && ctx->out[0][0] && (slots == 1 || ctx->out[1][0]);
    if (!have) {
        free_outputs(ctx);
        for (int s = 0; s < slots; ++s)
            for (int r = 0; r < 2; ++r) CU(cudaMalloc((void **)&ctx->out[s][r], want + 256));
        ctx->out_cap = want;
    }
    const uint32_t max_chunks = 1 << 16;
    DBuf<uint32_t> &d_bounds = ctx->s_cb, &d_n = ctx->s_cn;
    DBuf<uint64_t> &d_b1 = ctx->s_b1, &d_b2 = ctx->s_b2;
    CU(d_bounds.ensure(max_chunks + 1));
    CU(d_b1.ensure(max_chunks + 1));
    CU(d_b2.ensure(max_chunks + 1));
    CU(d_n.ensure(1));
    ctx->chunk_tiles.assign(1, 0);
    ctx->chunk_off1.assign(1, 0);
    ctx->chunk_off2.assign(1, 0);
    if (ctx->n_tiles == 0) return MMK_OK;
    chunk_table(ctx->d_tile_off1.p, ctx->d_tile_off2.p, ctx->n_tiles, ctx->out_cap, d_bounds.p, d_b1.p, d_b2.p, max_chunks,
                d_n.p, ctx->st);
    uint32_t n = 0;
    CU(cudaMemcpyAsync(&n, d_n.p, 4, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    if (n == 0xffffffffu) return fail(ctx, MMK_ERR_INVALID_ARG, "max_chunk_bytes is smaller than one tile of records");
    if (n == 0xfffffffeu) return fail(ctx, MMK_ERR_INVALID_ARG, "max_chunk_bytes gives more than 65536 chunks");
    ctx->chunk_tiles.resize(n + 1);
    ctx->chunk_off1.resize(n + 1);
    ctx->chunk_off2.resize(n + 1);
    CU(cudaMemcpyAsync(ctx->chunk_tiles.data(), d_bounds.p, (n + 1) * 4, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaMemcpyAsync(ctx->chunk_off1.data(), d_b1.p, (n + 1) * 8, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaMemcpyAsync(ctx->chunk_off2.data(), d_b2.p, (n + 1) * 8, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    return MMK_OK;
}

// enqueue the emit kernel of chunk k into ring slot `slot` (no synchronisation)
static int32_t launch_chunk(mmk_ctx *ctx, uint32_t k, int slot, mmk_chunk *out) {
    const uint32_t t0 = ctx->chunk_tiles[k], t1 = ctx->chunk_tiles[k + 1];
    EmitArgs a = base_args(ctx);
    a.out1 = ctx->out[slot][0];
    a.out2 = ctx->out[slot][1];
    a.chunk_off1 = ctx->chunk_off1[k];
    a.chunk_off2 = ctx->chunk_off2[k];
    a.tile0 = t0;
    a.n_tiles = t1 - t0;
    emit(a, ctx->st);
    if (out) {
        memset(out, 0, sizeof *out);
        out->r1_dev = a.out1;
        out->r2_dev = a.out2;
        out->r1_bytes = ctx->chunk_off1[k + 1] - ctx->chunk_off1[k];
        out->r2_bytes = ctx->chunk_off2[k + 1] - ctx->chunk_off2[k];
        out->r1_offset = ctx->chunk_off1[k];
        out->r2_offset = ctx->chunk_off2[k];
        out->first_pair = (int64_t)t0 * ctx->tile_pairs;
        uint64_t last = std::min<uint64_t>((uint64_t)t1 * ctx->tile_pairs, ctx->s_npairs);
        out->n_pairs = (int64_t)(last - (uint64_t)out->first_pair);
        out->slot = slot;
        out->done = (k + 2 >= ctx->chunk_tiles.size());
    }
    return MMK_OK;
}
static inline uint32_t n_chunks_of(const mmk_ctx *c) { return (uint32_t)c->chunk_tiles.size() - 1; }

extern "C" int32_t mmk_chunk_layout(mmk_ctx *ctx, uint64_t *slot_bytes, int32_t *n_chunks) {
    if (!ctx) return MMK_ERR_INVALID_ARG;
    if (!ctx->placed) return fail(ctx, MMK_ERR_STATE, "mmk_place has not run");
    if (slot_bytes) *slot_bytes = ctx->out_cap;
    if (n_chunks) *n_chunks = (int32_t)n_chunks_of(ctx);
    return MMK_OK;
}

extern "C" int32_t mmk_emit_chunk(mmk_ctx *ctx, mmk_chunk *out) {
    if (!ctx || !out) return MMK_ERR_INVALID_ARG;
    if (!ctx->placed) return fail(ctx, MMK_ERR_STATE, "mmk_place has not run");
    CU(cudaSetDevice(ctx->par.device));
    memset(out, 0, sizeof *out);
    if (ctx->tile_cursor >= n_chunks_of(ctx)) {
        out->done = 1;
        return MMK_OK;
    }
    int slot = ctx->out[1][0] ? ctx->slot : 0;
    CU(cudaEventRecord(ctx->ev_a, ctx->st));
    int32_t rc = launch_chunk(ctx, ctx->tile_cursor, slot, out);
    if (rc) return rc;
    CU(cudaEventRecord(ctx->ev_b, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    CU(cudaGetLastError());
    CU(cudaEventElapsedTime(&out->kernel_ms, ctx->ev_a, ctx->ev_b));
    ctx->tile_cursor++;
    ctx->slot ^= 1;
    return MMK_OK;
}
extern "C" int32_t mmk_rewind(mmk_ctx *ctx) {
    if (!ctx) return MMK_ERR_INVALID_ARG;
    ctx->tile_cursor = 0;
    ctx->slot = 0;
    ctx->bz_cum[0] = ctx->bz_cum[1] = 0;
    return MMK_OK;
}
extern "C" int32_t mmk_copy_chunk(mmk_ctx *ctx, const mmk_chunk *ch, void *r1_host, void *r2_host) {
    if (!ctx || !ch) return MMK_ERR_INVALID_ARG;
    CU(cudaSetDevice(ctx->par.device));
    if (r1_host && ch->r1_bytes) CU(cudaMemcpyAsync(r1_host, ch->r1_dev, ch->r1_bytes, cudaMemcpyDeviceToHost, ctx->st));
    if (r2_host && ch->r2_bytes) CU(cudaMemcpyAsync(r2_host, ch->r2_dev, ch->r2_bytes, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    return MMK_OK;
}

extern "C" int32_t mmk_emit_all(mmk_ctx *ctx, float *kernel_ms) {
    if (!ctx) return MMK_ERR_INVALID_ARG;
    if (!ctx->placed) return fail(ctx, MMK_ERR_STATE, "mmk_place has not run");
    CU(cudaSetDevice(ctx->par.device));
    const uint32_t n = n_chunks_of(ctx);
    CU(cudaEventRecord(ctx->ev_a, ctx->st));
    for (uint32_t k = 0; k < n; ++k) {
        int32_t rc = launch_chunk(ctx, k, ctx->out[1][0] ? (int)(k & 1) : 0, nullptr);
        if (rc) return rc;
    }
    CU(cudaEventRecord(ctx->ev_b, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    CU(cudaGetLastError());
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, ctx->ev_a, ctx->ev_b));
    if (kernel_ms) *kernel_ms = ms;
    return MMK_OK;
}

// Pipelined whole-shard emission: kernel of chunk k+1 overlaps the device→host copy of chunk k.
// dst(k, r) gives the host address for read r of chunk k; `after(k)` runs once chunk k is in host memory.
template <class Dst, class After>
static int32_t emit_pipeline(mmk_ctx *ctx, Dst dst, After after) {
    const uint32_t n = n_chunks_of(ctx);
    const bool ring = ctx->out[1][0] != nullptr;
    std::vector<mmk_chunk> chunks(n);
    for (uint32_t k = 0; k < n; ++k) {
        int slot = ring ? (int)(k & 1) : 0;
        // the device slot is free once the copy of chunk k-2 (same slot) finished
        if (k >= 2 || !ring) CU(cudaStreamWaitEvent(ctx->st, ctx->ev_copy[slot], 0));
        int32_t rc = launch_chunk(ctx, k, slot, &chunks[k]);
        if (rc) return rc;
        CU(cudaEventRecord(ctx->ev_emit[slot], ctx->st));
        // host slot of chunk k is reused from chunk k-2: its consumer must have returned
        if (k >= 1) {
            CU(cudaEventSynchronize(ctx->ev_copy[(ring ? (k - 1) & 1 : 0)]));
            rc = after(k - 1, chunks[k - 1]);
            if (rc) return rc;
        }
        CU(cudaStreamWaitEvent(ctx->st_copy, ctx->ev_emit[slot], 0));
        const mmk_chunk &ch = chunks[k];
        if (ch.r1_bytes) CU(cudaMemcpyAsync(dst(k, 0, ch), ch.r1_dev, ch.r1_bytes, cudaMemcpyDeviceToHost, ctx->st_copy));
        if (ch.r2_bytes) CU(cudaMemcpyAsync(dst(k, 1, ch), ch.r2_dev, ch.r2_bytes, cudaMemcpyDeviceToHost, ctx->st_copy));
        CU(cudaEventRecord(ctx->ev_copy[slot], ctx->st_copy));
    }
    CU(cudaStreamSynchronize(ctx->st_copy));
    CU(cudaStreamSynchronize(ctx->st));
    CU(cudaGetLastError());
    if (n >= 1) {
        int32_t rc = after(n - 1, chunks[n - 1]);
        if (rc) return rc;
    }
    return MMK_OK;
}

extern "C" int32_t mmk_emit_to_host(mmk_ctx *ctx, void *r1_host, uint64_t r1_cap, void *r2_host, uint64_t r2_cap) {
    if (!ctx || !r1_host || !r2_host) return MMK_ERR_INVALID_ARG;
    if (!ctx->placed) return fail(ctx, MMK_ERR_STATE, "mmk_place has not run");
    if (r1_cap < ctx->place_info.r1_bytes || r2_cap < ctx->place_info.r2_bytes) return fail(ctx, MMK_ERR_INVALID_ARG, "host buffers too small");
    CU(cudaSetDevice(ctx->par.device));
    return emit_pipeline(
        ctx, [&](uint32_t, int r, const mmk_chunk &ch) { return r ? (char *)r2_host + ch.r2_offset : (char *)r1_host + ch.r1_offset; },
        [&](uint32_t, const mmk_chunk &) { return (int32_t)MMK_OK; });
}

extern "C" int32_t mmk_emit_stream(mmk_ctx *ctx, void *const r1_host[2], void *const r2_host[2], uint64_t slot_cap,
                                   mmk_chunk_cb cb, void *user) {
    if (!ctx || !r1_host || !r2_host || !r1_host[0] || !r1_host[1] || !r2_host[0] || !r2_host[1]) return MMK_ERR_INVALID_ARG;
    if (!ctx->placed) return fail(ctx, MMK_ERR_STATE, "mmk_place has not run");
    if (slot_cap < std::min<uint64_t>(ctx->out_cap, std::max(ctx->place_info.r1_bytes, ctx->place_info.r2_bytes)))
        return fail(ctx, MMK_ERR_INVALID_ARG, "host slots are smaller than the device chunk size");
    CU(cudaSetDevice(ctx->par.device));
    return emit_pipeline(
        ctx, [&](uint32_t k, int r, const mmk_chunk &) { return r ? (char *)r2_host[k & 1] : (char *)r1_host[k & 1]; },
        [&](uint32_t k, const mmk_chunk &ch) {
            if (!cb) return (int32_t)MMK_OK;
            int32_t rc = cb(user, &ch, r1_host[k & 1], r2_host[k & 1]);
            if (rc) return fail(ctx, rc, "chunk consumer returned %d", rc);
            return (int32_t)MMK_OK;
        });
}

// ------------------------------------------------------------------------------------------
// BGZF output (the reference's writers are BGZFCompressorStream, src/Main.jl:42-43,111-112)
// ------------------------------------------------------------------------------------------
#include "mmk_bgzf_table.h"

static uint32_t bit_reverse(uint32_t v, int n) {
    uint32_t r = 0;
    for (int i = 0; i < n; ++i) r |= ((v >> i) & 1u) << (n - 1 - i);
    return r;
}
// canonical Huffman codes (RFC 1951 3.2.2) of `n` code lengths, bit-reversed for LSB-first packing: code | len << 16
static void canonical_codes(const uint8_t *len, int n, uint32_t *out) {
    uint32_t count[16] = {0}, next[16] = {0};
    for (int i = 0; i < n; ++i) count[len[i]]++;
    count[0] = 0;
    uint32_t code = 0;
    for (int b = 1; b < 16; ++b) {
        code = (code + count[b - 1]) << 1;
        next[b] = code;
    }
    for (int i = 0; i < n; ++i) out[i] = len[i] ? (bit_reverse(next[len[i]]++, len[i]) | ((uint32_t)len[i] << 16)) : 0u;
}
static uint32_t gf_mul(uint32_t a, uint32_t b) {
    uint32_t p = 0;
    for (int i = 0; i < 32; ++i) {
        if (a & 0x80000000u) p ^= b;
        a <<= 1;
        b = (b >> 1) ^ ((b & 1u) ? 0xEDB88320u : 0u);
    }
    return p;
}
// crc_words: 32-bit words per thread of the kernel's CRC pass (0: leave the per-thread shifts out)
static std::vector<uint32_t> bgzf_table_words(uint32_t crc_words) {
    std::vector<uint32_t> t(BGZF_TAB_CODES + (MMK_BGZF_HDR_BITS + 31) / 32, 0u);
    canonical_codes(MMK_BGZF_LL_LEN, 286, t.data());
    canonical_codes(MMK_BGZF_D_LEN, 30, t.data() + 286);
    uint32_t *crc = t.data() + 286 + 30;  // slicing-by-4: table k = CRC of a byte followed by k zero bytes
    for (uint32_t i = 0; i < 256; ++i) {
        uint32_t c = i;
        for (int k = 0; k < 8; ++k) c = (c >> 1) ^ ((c & 1u) ? 0xEDB88320u : 0u);
        crc[i] = c;
    }
    for (int k = 1; k < 4; ++k)
        for (uint32_t i = 0; i < 256; ++i) crc[k * 256 + i] = (crc[(k - 1) * 256 + i] >> 8) ^ crc[crc[(k - 1) * 256 + i] & 0xffu];
    if (crc_words) {  // lane t of T: x^(32 * crc_words * (T - 1 - t)) mod P
        const uint32_t T = 32;
        uint32_t step = 0x80000000u;  // x^0
        for (uint32_t i = 0; i < 32 * crc_words; ++i) step = (step >> 1) ^ ((step & 1u) ? 0xEDB88320u : 0u);
        uint32_t v = 0x80000000u;
        uint32_t *shift = crc + 1024;
        for (uint32_t k = 0; k < T; ++k) {
            shift[T - 1 - k] = v;
            v = gf_mul(v, step);
        }
    }
    for (size_t i = 0; i < sizeof MMK_BGZF_HDR; ++i) t[BGZF_TAB_CODES + i / 4] |= (uint32_t)MMK_BGZF_HDR[i] << (8 * (i % 4));
    return t;
}

extern "C" int32_t mmk_bgzf_table(uint32_t *litlen_codes, uint32_t *dist_codes, uint8_t *prefix, uint32_t prefix_cap,
                                  uint32_t *prefix_bits) {
    std::vector<uint32_t> t = bgzf_table_words(0);
    if (litlen_codes) memcpy(litlen_codes, t.data(), 286 * 4);
    if (dist_codes) memcpy(dist_codes, t.data() + 286, 30 * 4);
    if (prefix) {
        if (prefix_cap < sizeof MMK_BGZF_HDR) return MMK_ERR_INVALID_ARG;
        memcpy(prefix, MMK_BGZF_HDR, sizeof MMK_BGZF_HDR);
    }
    if (prefix_bits) *prefix_bits = MMK_BGZF_HDR_BITS;
    return MMK_OK;
}

extern "C" int32_t mmk_bgzf_eof(void *out, uint32_t cap) {
    static const uint8_t eof[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 27, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (!out || cap < sizeof eof) return MMK_ERR_INVALID_ARG;
    memcpy(out, eof, sizeof eof);
    return (int32_t)sizeof eof;
}

// geometry and buffers of the BGZF stage for the placed shard
static int32_t bgzf_prepare(mmk_ctx *ctx) {
    if (ctx->bz_ready) return MMK_OK;
    if (!ctx->h_bz_tot) CU(cudaMallocHost((void **)&ctx->h_bz_tot, 4 * sizeof(uint64_t)));
    const uint32_t bound = std::max(ctx->tile_bytes_bound, 64u);
    if (bound > 65280u) return fail(ctx, MMK_ERR_UNSUPPORTED, "one tile of records exceeds a BGZF block");
    uint32_t G = 1;
    if (const char *e = getenv("MMK_BGZF_TILES")) G = (uint32_t)std::max(1, atoi(e));
    G = std::min(G, 65280u / bound);
    G = std::max(1u, std::min(G, 64u / ctx->tile_pairs));  // a lane codes at most two records
    ctx->bz_tiles_per_block = G;
    ctx->bz_in_cap = (G * bound + 15) & ~15u;
    {
        BgzfArgs g;
        memset(&g, 0, sizeof g);
        g.in_cap = ctx->bz_in_cap;
        g.max_records = G * ctx->tile_pairs;
        bgzf_geometry(g);
        ctx->bz_crc_words = g.crc_words;
        std::vector<uint32_t> t = bgzf_table_words(ctx->bz_crc_words);
        CU(ctx->d_bz_tab.ensure(t.size()));
        CU(cudaMemcpyAsync(ctx->d_bz_tab.p, t.data(), t.size() * 4, cudaMemcpyHostToDevice, ctx->st));
        CU(cudaStreamSynchronize(ctx->st));
        ctx->bz_hdr_bits = MMK_BGZF_HDR_BITS;
    }
    uint32_t max_tiles = 0;
    for (size_t k = 0; k + 1 < ctx->chunk_tiles.size(); ++k) max_tiles = std::max(max_tiles, ctx->chunk_tiles[k + 1] - ctx->chunk_tiles[k]);
    ctx->bz_max_blocks = (max_tiles + G - 1) / G;
    ctx->bz_slot_shift = 0;
    while ((64u << ctx->bz_slot_shift) <= ctx->bz_in_cap) ++ctx->bz_slot_shift;  // largest shift with 2^shift <= in_cap / 32
    ctx->bz_out_cap = bgzf_slots_bytes(ctx->out_cap, ctx->bz_max_blocks, ctx->bz_slot_shift);
    CU(ctx->d_bz_slots.ensure(ctx->bz_out_cap));
    CU(ctx->d_bz_sizes.ensure((size_t)ctx->bz_max_blocks + 1));
    CU(ctx->d_bz_offs.ensure((size_t)ctx->bz_max_blocks + 2));
    CU(ctx->d_scan_tmp.ensure((size_t)ctx->bz_max_blocks / 1024 + 8));
    const int slots = ctx->out[1][0] ? 2 : 1;
    for (int s = 0; s < slots; ++s)
        for (int r = 0; r < 2; ++r) CU(ctx->bz_out[s][r].ensure(ctx->bz_out_cap));
    ctx->bz_ready = true;
    return MMK_OK;
}

// enqueue: text of chunk k (already being written into out[slot] on the stream) -> bz_out[slot][*], totals -> h_bz_tot
static int32_t bgzf_launch(mmk_ctx *ctx, uint32_t k, int slot) {
    const uint32_t t0 = ctx->chunk_tiles[k], t1 = ctx->chunk_tiles[k + 1];
    const uint32_t G = ctx->bz_tiles_per_block;
    const uint32_t nblk = (t1 - t0 + G - 1) / G;
    BgzfArgs a;
    memset(&a, 0, sizeof a);
    a.tile0 = t0;
    a.n_tiles = t1 - t0;
    a.tiles_per_block = G;
    a.in_cap = ctx->bz_in_cap;
    a.max_records = G * ctx->tile_pairs;
    bgzf_geometry(a);
    a.tab = ctx->d_bz_tab.p;
    a.hdr_bits = ctx->bz_hdr_bits;
    a.slots = ctx->d_bz_slots.p;
    a.slot_shift = ctx->bz_slot_shift;
    a.sizes = ctx->d_bz_sizes.p;
    for (int r = 0; r < 2; ++r) {
        a.text = ctx->out[slot][r];
        a.tile_off = r ? ctx->d_tile_off2.p : ctx->d_tile_off1.p;
        a.chunk_off = r ? ctx->chunk_off2[k] : ctx->chunk_off1[k];
        bgzf_encode(a, nblk, ctx->st);
        exclusive_scan_u32(ctx->d_bz_sizes.p, ctx->d_bz_offs.p, nblk, ctx->d_scan_tmp.p, ctx->st);
        ctx->h_bz_tot[slot * 2 + r] = 0;
        bgzf_compact(a, ctx->d_bz_offs.p, ctx->bz_out[slot][r].p, nblk, (unsigned long long *)&ctx->h_bz_tot[slot * 2 + r], ctx->st);
    }
    return MMK_OK;
}
static void bgzf_fill(mmk_ctx *ctx, int slot, mmk_chunk *ch) {  // after the totals arrived
    ch->r1_dev = ctx->bz_out[slot][0].p;
    ch->r2_dev = ctx->bz_out[slot][1].p;
    ch->r1_bytes = ctx->h_bz_tot[slot * 2 + 0];
    ch->r2_bytes = ctx->h_bz_tot[slot * 2 + 1];
    ch->r1_offset = ctx->bz_cum[0];
    ch->r2_offset = ctx->bz_cum[1];
    ctx->bz_cum[0] += ch->r1_bytes;
    ctx->bz_cum[1] += ch->r2_bytes;
}

extern "C" int32_t mmk_bgzf_bound(mmk_ctx *ctx, uint64_t *slot_bytes) {
    if (!ctx || !slot_bytes) return MMK_ERR_INVALID_ARG;
    if (!ctx->placed) return fail(ctx, MMK_ERR_STATE, "mmk_place has not run");
    CU(cudaSetDevice(ctx->par.device));
    int32_t rc = bgzf_prepare(ctx);
    if (rc) return rc;
    *slot_bytes = ctx->bz_out_cap;
    return MMK_OK;
}

extern "C" int32_t mmk_emit_chunk_bgzf(mmk_ctx *ctx, mmk_chunk *out) {
    if (!ctx || !out) return MMK_ERR_INVALID_ARG;
    if (!ctx->placed) return fail(ctx, MMK_ERR_STATE, "mmk_place has not run");
    CU(cudaSetDevice(ctx->par.device));
    int32_t rc = bgzf_prepare(ctx);
    if (rc) return rc;
    memset(out, 0, sizeof *out);
    if (ctx->tile_cursor >= n_chunks_of(ctx)) {
        out->done = 1;
        return MMK_OK;
    }
    int slot = ctx->out[1][0] ? ctx->slot : 0;
    CU(cudaEventRecord(ctx->ev_a, ctx->st));
    rc = launch_chunk(ctx, ctx->tile_cursor, slot, out);
    if (rc) return rc;
    rc = bgzf_launch(ctx, ctx->tile_cursor, slot);
    if (rc) return rc;
    CU(cudaEventRecord(ctx->ev_b, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    CU(cudaGetLastError());
    CU(cudaEventElapsedTime(&out->kernel_ms, ctx->ev_a, ctx->ev_b));
    bgzf_fill(ctx, slot, out);
    ctx->tile_cursor++;
    ctx->slot ^= 1;
    return MMK_OK;
}

extern "C" int32_t mmk_emit_stream_bgzf(mmk_ctx *ctx, void *const r1_host[2], void *const r2_host[2], uint64_t slot_cap,
                                        mmk_chunk_cb cb, void *user) {
    if (!ctx || !r1_host || !r2_host || !r1_host[0] || !r1_host[1] || !r2_host[0] || !r2_host[1]) return MMK_ERR_INVALID_ARG;
    if (!ctx->placed) return fail(ctx, MMK_ERR_STATE, "mmk_place has not run");
    CU(cudaSetDevice(ctx->par.device));
    int32_t rc = bgzf_prepare(ctx);
    if (rc) return rc;
    if (slot_cap < ctx->bz_out_cap) return fail(ctx, MMK_ERR_INVALID_ARG, "host slots are smaller than mmk_bgzf_bound");
    ctx->bz_cum[0] = ctx->bz_cum[1] = 0;
    const uint32_t n = n_chunks_of(ctx);
    const bool ring = ctx->out[1][0] != nullptr;
    std::vector<mmk_chunk> chunks(n);
    auto slot_of = [&](uint32_t k) { return ring ? (int)(k & 1) : 0; };
    // chunk j: its coded sizes are known once ev_emit fired; then the copy is queued on the copy stream
    auto start_copy = [&](uint32_t j) -> int32_t {
        const int sl = slot_of(j);
        CU(cudaEventSynchronize(ctx->ev_emit[sl]));
        bgzf_fill(ctx, sl, &chunks[j]);
        const mmk_chunk &ch = chunks[j];
        if (ch.r1_bytes) CU(cudaMemcpyAsync(r1_host[j & 1], ch.r1_dev, ch.r1_bytes, cudaMemcpyDeviceToHost, ctx->st_copy));
        if (ch.r2_bytes) CU(cudaMemcpyAsync(r2_host[j & 1], ch.r2_dev, ch.r2_bytes, cudaMemcpyDeviceToHost, ctx->st_copy));
        CU(cudaEventRecord(ctx->ev_copy[sl], ctx->st_copy));
        return MMK_OK;
    };
    auto finish = [&](uint32_t j) -> int32_t {
        CU(cudaEventSynchronize(ctx->ev_copy[slot_of(j)]));
        if (!cb) return MMK_OK;
        int32_t r = cb(user, &chunks[j], r1_host[j & 1], r2_host[j & 1]);
        if (r) return fail(ctx, r, "chunk consumer returned %d", r);
        return MMK_OK;
    };
    for (uint32_t k = 0; k < n; ++k) {
        const int slot = slot_of(k);
        if (!ring && k >= 1) {  // one device slot: strictly serial
            if ((rc = start_copy(k - 1))) return rc;
            if ((rc = finish(k - 1))) return rc;
        }
        // the coded buffers of this slot are free once the copy of chunk k-2 finished
        if (ring && k >= 2) CU(cudaStreamWaitEvent(ctx->st, ctx->ev_copy[slot], 0));
        if ((rc = launch_chunk(ctx, k, slot, &chunks[k]))) return rc;
        if ((rc = bgzf_launch(ctx, k, slot))) return rc;
        CU(cudaEventRecord(ctx->ev_emit[slot], ctx->st));
        if (ring && k >= 1) {
            // host slot (k-1)&1 was handed back by finish(k-3) one iteration ago; the consumer of chunk k-2 then
            // runs while chunk k-1 crosses PCIe and chunk k is being generated
            if ((rc = start_copy(k - 1))) return rc;
            if (k >= 2 && (rc = finish(k - 2))) return rc;
        }
    }
    if (n >= 1) {
        if ((rc = start_copy(n - 1))) return rc;
        if (ring && n >= 2 && (rc = finish(n - 2))) return rc;
        if ((rc = finish(n - 1))) return rc;
    }
    CU(cudaStreamSynchronize(ctx->st_copy));
    CU(cudaStreamSynchronize(ctx->st));
    CU(cudaGetLastError());
    return MMK_OK;
}

// ------------------------------------------------------------------------------------------
// files
// ------------------------------------------------------------------------------------------
namespace {
struct FileSink {
    int fd[2] = {-1, -1};
    uint64_t written[2] = {0, 0};
    int err = 0;
};
bool write_all(int fd, const char *p, uint64_t n) {
    while (n) {
        const ssize_t w = ::write(fd, p, (size_t)std::min<uint64_t>(n, 1u << 30));
        if (w < 0) {
            if (errno == EINTR) continue;
            return false;
        }
        p += w;
        n -= (uint64_t)w;
    }
    return true;
}
int32_t file_sink_cb(void *user, const mmk_chunk *ch, const void *r1, const void *r2) {
    FileSink *s = (FileSink *)user;
    bool ok2 = true;
    std::thread t2([&] { ok2 = write_all(s->fd[1], (const char *)r2, ch->r2_bytes); });  // R2 beside R1
    const bool ok1 = write_all(s->fd[0], (const char *)r1, ch->r1_bytes);
    t2.join();
    if (!ok1 || !ok2) {
        s->err = errno ? errno : EIO;
        return MMK_ERR_IO;
    }
    s->written[0] += ch->r1_bytes;
    s->written[1] += ch->r2_bytes;
    return MMK_OK;
}
}  // namespace

extern "C" int32_t mmk_emit_files(mmk_ctx *ctx, const char *r1_path, const char *r2_path, int32_t bgzf, int32_t append_eof,
                                  int32_t append, uint64_t bytes_written[2]) {
    if (!ctx || !r1_path || !r2_path) return MMK_ERR_INVALID_ARG;
    if (!ctx->placed) return fail(ctx, MMK_ERR_STATE, "mmk_place has not run");
    CU(cudaSetDevice(ctx->par.device));
    uint64_t slot = 0;
    if (bgzf) {
        int32_t rc = mmk_bgzf_bound(ctx, &slot);
        if (rc) return rc;
    } else
        slot = ctx->out_cap;
    slot = std::max<uint64_t>(slot, 4096);
    if (ctx->h_ring_cap < slot) {  // pinned ring, kept with the context
        for (auto &p : ctx->h_ring) {
            if (p) cudaFreeHost(p);
            p = nullptr;
        }
        ctx->h_ring_cap = 0;
        for (auto &p : ctx->h_ring) CU(cudaMallocHost((void **)&p, slot));
        ctx->h_ring_cap = slot;
    }
    FileSink sink;
    const int flags = O_WRONLY | O_CREAT | (append ? O_APPEND : O_TRUNC);
    sink.fd[0] = ::open(r1_path, flags, 0644);
    sink.fd[1] = ::open(r2_path, flags, 0644);
    if (sink.fd[0] < 0 || sink.fd[1] < 0) {
        const int e = errno;
        if (sink.fd[0] >= 0) ::close(sink.fd[0]);
        if (sink.fd[1] >= 0) ::close(sink.fd[1]);
        return fail(ctx, MMK_ERR_IO, "cannot open %s / %s for writing: %s", r1_path, r2_path, strerror(e));
    }
    void *const s1[2] = {ctx->h_ring[0], ctx->h_ring[1]}, *const s2[2] = {ctx->h_ring[2], ctx->h_ring[3]};
    int32_t rc = bgzf ? mmk_emit_stream_bgzf(ctx, s1, s2, slot, file_sink_cb, &sink) : mmk_emit_stream(ctx, s1, s2, slot, file_sink_cb, &sink);
    if (!rc && bgzf && append_eof) {
        uint8_t eof[28];
        mmk_bgzf_eof(eof, sizeof eof);
        for (int r = 0; r < 2; ++r) {
            if (!write_all(sink.fd[r], (const char *)eof, sizeof eof)) rc = MMK_ERR_IO;
            else sink.written[r] += sizeof eof;
        }
    }
    const bool c1 = ::close(sink.fd[0]) == 0, c2 = ::close(sink.fd[1]) == 0;
    if (rc == MMK_ERR_IO || (!rc && !(c1 && c2))) return fail(ctx, MMK_ERR_IO, "write to %s / %s failed: %s", r1_path, r2_path, strerror(sink.err ? sink.err : errno));
    if (bytes_written) {
        bytes_written[0] = sink.written[0];
        bytes_written[1] = sink.written[1];
    }
    return rc;
}

extern "C" uint64_t mmk_launch_count(void) { return (uint64_t)launch_count(); }

// ------------------------------------------------------------------------------------------
// introspection
// ------------------------------------------------------------------------------------------
extern "C" int32_t mmk_dump_molecules(mmk_ctx *ctx, int64_t first, int64_t n, mmk_mol_record *out) {
    if (!ctx || !out || first < 0 || n < 0) return MMK_ERR_INVALID_ARG;
    if (!ctx->planned) return fail(ctx, MMK_ERR_STATE, "mmk_plan has not run");
    if ((uint64_t)(first + n) > ctx->n_mol) return fail(ctx, MMK_ERR_INVALID_ARG, "molecule range out of bounds");
    CU(cudaSetDevice(ctx->par.device));
    std::vector<uint64_t> g(n), ro(n);
    std::vector<uint32_t> sl(n), sz(n), nf(n);
    std::vector<uint16_t> sc(n);
    if (n) {
        CU(cudaMemcpy(g.data(), ctx->p_group.p + first, n * 8, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(ro.data(), ctx->d_read_off.p + first, n * 8, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(sl.data(), ctx->p_slot.p + first, n * 4, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(sz.data(), ctx->p_size.p + first, n * 4, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(nf.data(), ctx->p_nfrags.p + first, n * 4, cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(sc.data(), ctx->p_schema.p + first, n * 2, cudaMemcpyDeviceToHost));
    }
    std::vector<MolRec> mr;
    int64_t lo = 0, hi = 0;  // overlap with the placed shard
    if (ctx->placed) {
        lo = std::max<int64_t>(first, (int64_t)ctx->mol0);
        hi = std::min<int64_t>(first + n, (int64_t)(ctx->mol0 + ctx->s_nmol));
        if (hi > lo) {
            mr.resize(hi - lo);
            CU(cudaMemcpy(mr.data(), ctx->d_mol.p + (lo - (int64_t)ctx->mol0), (hi - lo) * sizeof(MolRec), cudaMemcpyDeviceToHost));
        }
    }
    for (int64_t i = 0; i < n; ++i) {
        mmk_mol_record &o = out[i];
        o.group = (int64_t)g[i];
        o.slot = (int32_t)sl[i];
        o.schema = sc[i];
        o.molsize = sz[i];
        o.n_frags = nf[i];
        o.first_read = (int64_t)ro[i];
        o.barcode_index = first + i;
        int64_t gi = first + i;
        o.start = (gi >= lo && gi < hi) ? mr[gi - lo].start : 0;
    }
    return MMK_OK;
}

// the insert plan of the placed shard, derived from the pair descriptors when the block placement ran
static int32_t ensure_ins(mmk_ctx *ctx) {
    if (ctx->ins_valid) return MMK_OK;
    CU(ctx->d_ins.ensure(ctx->s_npairs + 1));
    ins_from_fdesc(ctx->d_fdesc.p, ctx->d_mol.p, ctx->d_tile_mol0.p, ctx->s_npairs, ctx->d_ins.p, ctx->st);
    CU(cudaStreamSynchronize(ctx->st));
    CU(cudaGetLastError());
    ctx->ins_valid = true;
    return MMK_OK;
}

extern "C" int32_t mmk_dump_inserts(mmk_ctx *ctx, int64_t first, int64_t n, int64_t *start, int64_t *stop) {
    if (!ctx || !start || !stop || first < 0 || n < 0) return MMK_ERR_INVALID_ARG;
    if (!ctx->placed) return fail(ctx, MMK_ERR_STATE, "mmk_place has not run");
    if ((uint64_t)(first + n) > ctx->s_npairs) return fail(ctx, MMK_ERR_INVALID_ARG, "pair range out of bounds");
    CU(cudaSetDevice(ctx->par.device));
    {
        int32_t rc = ensure_ins(ctx);
        if (rc) return rc;
    }
    std::vector<uint2> in(n);
    std::vector<uint64_t> ro(ctx->s_nmol + 1);
    std::vector<MolRec> mr(ctx->s_nmol);
    if (n) CU(cudaMemcpy(in.data(), ctx->d_ins.p + first, n * sizeof(uint2), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(ro.data(), ctx->d_roff_l.p, ro.size() * 8, cudaMemcpyDeviceToHost));
    if (ctx->s_nmol) CU(cudaMemcpy(mr.data(), ctx->d_mol.p, mr.size() * sizeof(MolRec), cudaMemcpyDeviceToHost));
    size_t m = std::upper_bound(ro.begin(), ro.end(), (uint64_t)first) - ro.begin();
    m = m ? m - 1 : 0;
    for (int64_t i = 0; i < n; ++i) {
        uint64_t p = (uint64_t)(first + i);
        while (m + 1 < ro.size() && ro[m + 1] <= p) ++m;
        start[i] = mr[m].start + in[i].x;
        stop[i] = start[i] + (in[i].y & 0x00ffffffu) - 1;
    }
    return MMK_OK;
}

extern "C" int32_t mmk_dump_variates(mmk_ctx *ctx, mmk_variate *out, uint64_t cap, uint64_t *n_out) {
    if (!ctx || !n_out) return MMK_ERR_INVALID_ARG;
    if (!ctx->placed || ctx->world != 1) return fail(ctx, MMK_ERR_STATE, "mmk_dump_variates needs mmk_place on an unsharded context");
    CU(cudaSetDevice(ctx->par.device));
    {
        int32_t rc = ensure_ins(ctx);
        if (rc) return rc;
    }
    DBuf<uint32_t> d_counts, d_nlive;
    DBuf<uint64_t> d_offsets, d_tmp;
    CU(d_counts.ensure(ctx->n_mol + 1));
    CU(d_offsets.ensure(ctx->n_mol + 2));
    CU(d_tmp.ensure(ctx->n_mol / 1024 + 8));
    CU(d_nlive.ensure(ctx->epoch_nlive.size() + 1));
    CU(cudaMemcpyAsync(d_nlive.p, ctx->epoch_nlive.data(), ctx->epoch_nlive.size() * 4, cudaMemcpyHostToDevice, ctx->st));
    dump_variates(ctx->sc, ctx->d_sch.p, plan_dev(ctx), ctx->n_mol, ctx->d_read_off.p, ctx->d_ins.p, d_nlive.p,
                  ctx->d_tile_mol0.p, ctx->tile_pairs, d_counts.p, d_offsets.p, d_tmp.p, nullptr, 0, ctx->st);
    uint64_t total = 0;
    CU(cudaMemcpyAsync(&total, d_offsets.p + ctx->n_mol, 8, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    CU(cudaGetLastError());
    *n_out = total;
    if (!out) return MMK_OK;
    if (cap < total) return fail(ctx, MMK_ERR_INVALID_ARG, "variate buffer too small");
    DBuf<mmk_variate> d_out;
    CU(d_out.ensure(total + 1));
    dump_variates(ctx->sc, ctx->d_sch.p, plan_dev(ctx), ctx->n_mol, ctx->d_read_off.p, ctx->d_ins.p, d_nlive.p,
                  ctx->d_tile_mol0.p, ctx->tile_pairs, d_counts.p, d_offsets.p, d_tmp.p, d_out.p, 1, ctx->st);
    CU(cudaMemcpyAsync(out, d_out.p, total * sizeof(mmk_variate), cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    CU(cudaGetLastError());
    return MMK_OK;
}

extern "C" int32_t mmk_dump_mask_variates(mmk_ctx *ctx, mmk_variate *out, uint64_t cap, uint64_t *n_out) {
    if (!ctx || !n_out) return MMK_ERR_INVALID_ARG;
    uint64_t total = 0;
    for (auto &c : ctx->contigs) total += c.mask_variates.size();
    *n_out = total;
    if (!out) return MMK_OK;
    if (cap < total) return fail(ctx, MMK_ERR_INVALID_ARG, "variate buffer too small");
    uint64_t o = 0;
    for (auto &c : ctx->contigs) {
        if (!c.mask_variates.empty()) memcpy(out + o, c.mask_variates.data(), c.mask_variates.size() * sizeof(mmk_variate));
        o += c.mask_variates.size();
    }
    return MMK_OK;
}

extern "C" int32_t mmk_kat_philox(mmk_ctx *ctx, const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    if (!ctx || !ctr || !key || !out) return MMK_ERR_INVALID_ARG;
    CU(cudaSetDevice(ctx->par.device));
    DBuf<uint32_t> d;
    CU(d.ensure(10));
    uint32_t in[6] = {ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]};
    CU(cudaMemcpyAsync(d.p, in, sizeof in, cudaMemcpyHostToDevice, ctx->st));
    kat_philox(d.p, d.p + 6, ctx->st);
    CU(cudaMemcpyAsync(out, d.p + 6, 16, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    CU(cudaGetLastError());
    return MMK_OK;
}
extern "C" int32_t mmk_kat_math(mmk_ctx *ctx, int32_t kind, const double *in, double *out, int64_t n) {
    if (!ctx || !in || !out || n < 0) return MMK_ERR_INVALID_ARG;
    CU(cudaSetDevice(ctx->par.device));
    DBuf<double> d;
    CU(d.ensure((size_t)2 * n + 2));
    CU(cudaMemcpyAsync(d.p, in, n * 8, cudaMemcpyHostToDevice, ctx->st));
    kat_math(kind, d.p, d.p + n, n, ctx->st);
    CU(cudaMemcpyAsync(out, d.p + n, n * 8, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    CU(cudaGetLastError());
    return MMK_OK;
}
extern "C" int32_t mmk_kat_insert_len(mmk_ctx *ctx, const uint64_t *bits, uint32_t *len, uint8_t *slow, int64_t n) {
    if (!ctx || !bits || !len || !slow || n < 0) return MMK_ERR_INVALID_ARG;
    CU(cudaSetDevice(ctx->par.device));
    DBuf<uint64_t> d_bits;
    DBuf<uint32_t> d_len;
    DBuf<uint8_t> d_slow;
    CU(d_bits.ensure((size_t)n + 1));
    CU(d_len.ensure((size_t)n + 1));
    CU(d_slow.ensure((size_t)n + 1));
    CU(cudaMemcpyAsync(d_bits.p, bits, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->st));
    kat_insert_len(ctx->sc, ctx->d_instab.p, d_bits.p, d_len.p, d_slow.p, n, ctx->st);
    CU(cudaMemcpyAsync(len, d_len.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaMemcpyAsync(slow, d_slow.p, (size_t)n, cudaMemcpyDeviceToHost, ctx->st));
    CU(cudaStreamSynchronize(ctx->st));
    CU(cudaGetLastError());
    return MMK_OK;
}
extern "C" int32_t mmk_insert_table(int32_t insert_length, int32_t insert_stdev, uint32_t *enabled, uint32_t *n_thr, uint32_t *base_len,
                                    uint32_t *u_lo, uint32_t *u_hi, uint32_t *guard, uint32_t *thr, uint32_t thr_cap) {
    if (!enabled || !n_thr || !base_len || !u_lo || !u_hi || !guard) return MMK_ERR_INVALID_ARG;
    SimConst sc;
    memset(&sc, 0, sizeof sc);
    sc.ins_mu = (double)insert_length;
    sc.ins_sd = (double)insert_stdev;
    sc.ins_lb = MMK_DIV(MMK_SUB(200.0, sc.ins_mu), sc.ins_sd);
    sc.ins_robert = (0.5 * erfc(sc.ins_lb / sqrt(2.0)) > 0.3) ? 0 : 1;
    InsTable *t = new InsTable();
    build_insert_table(sc, *t);
    *enabled = t->enabled;
    *n_thr = t->n_thr;
    *base_len = t->base_len;
    *u_lo = t->u_lo;
    *u_hi = t->u_hi;
    *guard = INS_GUARD_HOST;
    int32_t rc = MMK_OK;
    if (thr) {
        if (thr_cap < t->n_thr) rc = MMK_ERR_INVALID_ARG;
        else memcpy(thr, t->thr, (size_t)t->n_thr * 4);
    }
    delete t;
    return rc;
}
