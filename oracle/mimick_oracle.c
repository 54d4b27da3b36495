/* mimick_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A sequential CPU restatement of the simulation core of pdimens/mimick
 * (the .jl files under mimick/MimickLinkedReads.jl/src), used only as the checker for the CUDA path:
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 * Nothing under mimick_b200/ links, imports or executes this file.
 *
 * PARITY UNPINNED: the reference ships no golden vectors, no unit tests and no seeded
 * outputs (test/runtests.jl:4-6 is an empty stub; CI checks exit codes only), and Julia is
 * not installable in the build container, so this restatement could not be run against the
 * reference itself.  It is pinned instead to (i) the literal format examples in the
 * reference docs, (ii) hand-derived known answers on the reference's own fixtures
 * (tests/test_oracle_kat.py) and (iii) analytic distribution laws.  Third-party arithmetic
 * the reference leans on (Distributions.jl >= 0.25.125 `truncated`/`rand`, BioSequences
 * >= 3.5 slicing/reverse_complement!/spliceinto!) is restated from its published algorithm,
 * recalled rather than vendored — see DESIGN.md §Oracle.
 *
 * Every random number is drawn through `variate source` callbacks so the same code runs
 *   (a) REPLAY: from a recorded tagged stream (what the GPU consumed, dumped by
 *       mmk_dump_variates), touching no RNG at all; or
 *   (b) PHILOX: free-running from include/mmk_rng.h with the same stream addressing as the
 *       kernels.
 *
 * Julia line references use the shorthand src/X.jl = mimick/MimickLinkedReads.jl/src/X.jl.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <zlib.h> /* crc32: digest mode only */

#include "mmk_rng.h"

/* error codes: numerically identical to include/mimick_b200.h (tests assert this) */
enum {
    ORC_OK = 0,
    ORC_ERR_INVALID_ARG = 1,
    ORC_ERR_BAD_FORMAT = 2,          /* src/Common.jl:71 */
    ORC_ERR_INSERTS_UNFITTABLE = 3,  /* src/Breakpoints.jl:214-216 */
    ORC_ERR_BARCODES_EXHAUSTED = 4,  /* src/Barcodes.jl:75 */
    ORC_ERR_CONTIG_TOO_AMBIGUOUS = 5,/* src/ProcessFASTA.jl:58-60 */
    ORC_ERR_BCF_UNSUPPORTED = 6,     /* src/Main.jl:95-97 */
    ORC_ERR_TRUNCATION_BOUNDS = 7,   /* truncated(...; lower > upper), src/Breakpoints.jl:191 */
    ORC_ERR_BAD_BASE = 8,
    ORC_ERR_IO = 9,
    ORC_ERR_VCF = 10,
    ORC_ERR_BOUNDS = 11,             /* Julia BoundsError on a read window */
    ORC_ERR_UNSUPPORTED = 12,
    ORC_ERR_CUDA = 13,
    ORC_ERR_NOMEM = 14,
    ORC_ERR_REPLAY = 15,
    ORC_ERR_STATE = 16
};

enum { ORC_FMT_STLFR = 0, ORC_FMT_HAPLOTAGGING = 1, ORC_FMT_TELLSEQ = 2, ORC_FMT_TENX = 3, ORC_FMT_STANDARD = 4 };

typedef struct {
    uint32_t tag; /* enum mmk_purpose */
    uint32_t aux; /* randint: low 32 bits of the range size n (consistency check); else 0 */
    union {
        double d;
        int64_t i;
    } v;
} orc_variate;

typedef struct {
    int32_t insert_length;   /* -i  src/Structs.jl:68-69 */
    int32_t insert_stdev;    /* -d */
    int32_t length_r1;       /* -l */
    int32_t length_r2;
    int32_t n_molecules;     /* -N */
    int32_t attempts;        /* -A */
    int64_t mol_length;      /* -L */
    double mol_coverage;     /* -C */
    double singletons;       /* -S */
    int32_t circular;        /* -c */
    int32_t fq_format;       /* ORC_FMT_* (output style)   src/Common.jl:56-74 */
    int32_t bc_format;       /* ORC_FMT_* (barcode style) */
    int32_t strict_wrap;     /* 0 = literal reference wrap (D5), 1 = modular wrap */
    int32_t pad;
    double err_sub_start, err_sub_end, err_ins, err_del; /* sequencing-error model (not in the reference; all 0 = off) */
} orc_params;

typedef struct {
    double theta, lower, upper;
    int has_upper;
    double loglcdf, lcdf, logucdf, ucdf, tp, logtp;
} trunc_exp;

typedef struct {
    char *chrom;
    int haplotype;
    char *seq; /* uppercase IUPAC, 0-based storage of a 1-based sequence */
    int borrowed; /* seq is owned by the caller (large benchmarks share one copy between threads) */
    int64_t len;
    int64_t reads_required, reads_current;
} orc_schema;

typedef struct {
    uint8_t *p;
    size_t n, cap;
} bytebuf;

typedef struct {
    int64_t group;
    int32_t slot;
    int32_t schema;
    int64_t start;   /* 1-based molecule start */
    int64_t molsize;
    int64_t n_frags;
    int64_t first_read;
    int64_t barcode_index;
} orc_mol;

/* Digest mode (full-size parity runs, where the text would not fit in host memory): a running CRC-32
 * of each stream, sampled at the first barcode-group boundary at or after every `window` read pairs. */
typedef struct {
    int64_t pairs;          /* read pairs emitted so far */
    uint64_t bytes1, bytes2; /* R1 / R2 bytes emitted so far */
    uint32_t crc1, crc2;    /* CRC-32 (zlib) of R1[0:bytes1], R2[0:bytes2] */
    uint32_t pad[2];
} orc_checkpoint;

typedef struct orc_sim {
    orc_params par;
    /* derived: src/Structs.jl:54-70 */
    double lowerbound;
    trunc_exp d_mollen, d_nmol;
    int nmol_fixed;
    double ins_mu, ins_sd, ins_lb; /* truncated Normal, standardised lower bound */
    double ins_tp;                 /* kept mass 1 - cdf(lb): picks the sampler of randnt */

    orc_schema *sch;
    int n_sch, cap_sch;

    /* variate source */
    int mode; /* 0 replay, 1 philox */
    const orc_variate *rp;
    size_t rp_n, rp_pos;
    uint64_t seed;
    uint32_t sample;
    mmk_stream st[16];
    int st_init[16];
    uint64_t cur_g;
    uint32_t cur_m;
    uint32_t cur_round, cur_insert; /* insert-length normals: one Philox stream per (round, insert) */
    /* optional record of what was consumed */
    orc_variate *rec;
    size_t rec_n, rec_cap;
    int record;

    /* outputs */
    bytebuf r1, r2;
    orc_mol *mols;
    size_t n_mols, cap_mols;
    int64_t *ins_start; /* 1-based inclusive */
    int64_t *ins_stop;
    size_t n_ins, cap_ins;
    int64_t n_groups;
    int64_t max_pairs; /* optional stop (bench sampling): 0 = run to completion */
    int discard;       /* bench: recycle the output buffers after every group, count bytes only */
    int64_t digest_window; /* > 0: digest mode (text discarded after every group, molecules and inserts kept) */
    int64_t next_cp;
    uint32_t crc1, crc2;
    orc_checkpoint *cps;
    size_t n_cps, cap_cps;
    uint64_t bytes_r1, bytes_r2;
    int64_t total_pairs;
    /* sequencing-error model: per-cycle substitution thresholds and base qualities of R1 / R2 (DESIGN.md "Error model") */
    int err_on;
    uint32_t *err_thr[2];
    uint8_t *err_qb[2];
    uint32_t err_tins, err_tdel;
    int err;
    char errmsg[256];
} orc_sim;

/* ---------- small utilities ------------------------------------------------------------ */
static int bb_reserve(bytebuf *b, size_t extra) {
    if (b->n + extra <= b->cap) return 0;
    size_t nc = b->cap ? b->cap : 1 << 16;
    while (nc < b->n + extra) nc *= 2;
    uint8_t *q = (uint8_t *)realloc(b->p, nc);
    if (!q) return -1;
    b->p = q;
    b->cap = nc;
    return 0;
}
static void bb_put(bytebuf *b, const void *s, size_t n) {
    if (bb_reserve(b, n)) return;
    memcpy(b->p + b->n, s, n);
    b->n += n;
}
static void bb_fill(bytebuf *b, int c, size_t n) {
    if (bb_reserve(b, n)) return;
    memset(b->p + b->n, c, n);
    b->n += n;
}
static int fail(orc_sim *s, int code, const char *msg) {
    if (!s->err) {
        s->err = code;
        snprintf(s->errmsg, sizeof s->errmsg, "%s", msg);
    }
    return code;
}

/* ---------- variate source ------------------------------------------------------------- */
static void rec_push(orc_sim *s, uint32_t tag, uint32_t aux, double d, int64_t i, int is_int) {
    if (!s->record) return;
    if (s->rec_n == s->rec_cap) {
        s->rec_cap = s->rec_cap ? s->rec_cap * 2 : 1 << 16;
        s->rec = (orc_variate *)realloc(s->rec, s->rec_cap * sizeof(orc_variate));
    }
    orc_variate *v = &s->rec[s->rec_n++];
    v->tag = tag;
    v->aux = aux;
    if (is_int) v->v.i = i; else v->v.d = d;
}
static mmk_stream *stream_for(orc_sim *s, uint32_t purpose) {
    if (!s->st_init[purpose]) {
        uint64_t g = s->cur_g;
        uint32_t m = s->cur_m;
        if (purpose == MMK_P_NMOL || purpose == MMK_P_TARGET) m = 0;
        if (purpose == MMK_P_INSERT) /* one stream per (redraw round, insert): include/mmk_rng.h "Insert streams" */
            mmk_insert_stream_init(&s->st[purpose], mmk_round_key(s->seed, s->cur_round), s->sample, g, m, s->cur_insert);
        else
            mmk_stream_init(&s->st[purpose], s->seed, s->sample, purpose, g, m);
        s->st_init[purpose] = 1;
    }
    return &s->st[purpose];
}
static void src_begin_group(orc_sim *s, uint64_t g) {
    s->cur_g = g;
    s->cur_m = 0;
    memset(s->st_init, 0, sizeof s->st_init);
}
static void src_begin_molecule(orc_sim *s, uint32_t m) {
    s->cur_m = m;
    for (int p = MMK_P_MOLSIZE; p <= MMK_P_GAP; ++p) s->st_init[p] = 0;
}
static void src_begin_insert(orc_sim *s, uint32_t round, uint32_t i) {
    s->cur_round = round;
    s->cur_insert = i;
    s->st_init[MMK_P_INSERT] = 0;
}
static const orc_variate *rp_pop(orc_sim *s, uint32_t tag) {
    if (s->rp_pos >= s->rp_n) {
        fail(s, ORC_ERR_REPLAY, "replay stream exhausted");
        return NULL;
    }
    const orc_variate *v = &s->rp[s->rp_pos];
    if (v->tag != tag) {
        char m[128];
        snprintf(m, sizeof m, "replay tag mismatch at %zu: stream has %u, logic wants %u", s->rp_pos, v->tag, tag);
        fail(s, ORC_ERR_REPLAY, m);
        return NULL;
    }
    s->rp_pos++;
    return v;
}
static double src_u01(orc_sim *s, uint32_t purpose) {
    double d;
    if (s->mode == 0) {
        const orc_variate *v = rp_pop(s, purpose);
        d = v ? v->v.d : 0.5;
    } else
        d = mmk_next_u01(stream_for(s, purpose));
    rec_push(s, purpose, 0, d, 0, 0);
    return d;
}
static double src_randexp(orc_sim *s, uint32_t purpose) {
    double d;
    if (s->mode == 0) {
        const orc_variate *v = rp_pop(s, purpose);
        d = v ? v->v.d : 1.0;
    } else
        d = mmk_next_exp(stream_for(s, purpose));
    rec_push(s, purpose, 0, d, 0, 0);
    return d;
}
/* variates of the insert-length sampler (tag INSERT): kind 0 randn, 1 randexp, 2 rand */
static double src_insert_variate(orc_sim *s, int kind) {
    double d;
    if (s->mode == 0) {
        const orc_variate *v = rp_pop(s, MMK_P_INSERT);
        d = v ? v->v.d : (kind == 1 ? 1.0 : kind == 2 ? 0.5 : 0.0);
    } else {
        uint64_t bits = mmk_insert_next64(stream_for(s, MMK_P_INSERT));
        d = kind == 0 ? mmk_bits_to_normal(bits) : kind == 1 ? mmk_bits_to_exp(bits) : mmk_bits_to_u01(bits);
    }
    rec_push(s, MMK_P_INSERT, 0, d, 0, 0);
    return d;
}
/* gap before insert i, uniform in [0, n): Philox mode reads draw 1 of the insert's round-0 stream */
static int64_t src_gap(orc_sim *s, uint32_t i, uint64_t n) {
    int64_t r;
    if (s->mode == 0) {
        const orc_variate *v = rp_pop(s, MMK_P_GAP);
        r = v ? v->v.i : 0;
        if (v && (v->aux != (uint32_t)n || r < 0 || (uint64_t)r >= n)) {
            fail(s, ORC_ERR_REPLAY, "replay gap variate outside the range the logic asked for");
            r = 0;
        }
    } else {
        mmk_stream st;
        mmk_insert_stream_init(&st, s->seed, s->sample, s->cur_g, s->cur_m, i);
        r = (int64_t)mmk_bits_to_index(mmk_stream_at(&st, MMK_INSERT_GAP_DRAW), n);
    }
    rec_push(s, MMK_P_GAP, (uint32_t)n, 0, r, 1);
    return r;
}
/* uniform integer in [0, n) */
static int64_t src_index(orc_sim *s, uint32_t purpose, uint64_t n) {
    int64_t r;
    if (s->mode == 0) {
        const orc_variate *v = rp_pop(s, purpose);
        r = v ? v->v.i : 0;
        if (v && (v->aux != (uint32_t)n || r < 0 || (uint64_t)r >= n)) {
            fail(s, ORC_ERR_REPLAY, "replay integer variate outside the range the logic asked for");
            r = 0;
        }
    } else
        r = (int64_t)mmk_next_index(stream_for(s, purpose), n);
    rec_push(s, purpose, (uint32_t)n, 0, r, 1);
    return r;
}

/* ---------- Distributions.jl restated --------------------------------------------------- */
/* Exponential(θ): logcdf(x) = log1mexp(-max(x,0)/θ); quantile(p) = -θ·log1p(-p);
 * invlogcdf(lp) = -θ·log1mexp(lp); rand = θ·randexp().
 * truncated(d, l, u): loglcdf = logcdf(l), lcdf = exp(loglcdf), logucdf = logcdf(u),
 * ucdf = exp(logucdf), logtp = logsubexp(loglcdf, logucdf), tp = exp(logtp);
 * lower-only: logtp = log1mexp(loglcdf). */
static double exp_logcdf(double theta, double x) {
    double z = MMK_DIV(x > 0.0 ? x : 0.0, theta);
    return mmk_log1mexp(-z);
}
static int trunc_exp_make(trunc_exp *d, double theta, double lower, int has_upper, double upper) {
    d->theta = theta;
    d->lower = lower;
    d->upper = upper;
    d->has_upper = has_upper;
    if (has_upper && !(lower <= upper)) return ORC_ERR_TRUNCATION_BOUNDS;
    d->loglcdf = exp_logcdf(theta, lower);
    d->lcdf = mmk_exp(d->loglcdf);
    if (has_upper) {
        d->logucdf = exp_logcdf(theta, upper);
        d->ucdf = mmk_exp(d->logucdf);
        d->logtp = mmk_logsubexp(d->loglcdf, d->logucdf);
    } else {
        d->logucdf = 0.0;
        d->ucdf = 1.0;
        d->logtp = mmk_log1mexp(d->loglcdf);
    }
    d->tp = mmk_exp(d->logtp);
    return 0;
}
/* rand(::Truncated): rejection if tp > 0.25, quantile if tp > sqrt(eps), else invlogcdf */
static double trunc_exp_rand(orc_sim *s, const trunc_exp *d, uint32_t purpose) {
    if (d->tp > 0.25) {
        for (;;) {
            double r = MMK_MUL(src_randexp(s, purpose), d->theta);
            if (r >= d->lower && (!d->has_upper || r <= d->upper)) return r;
            if (s->err) return d->lower;
        }
    } else if (d->tp > 1.4901161193847656e-08) {
        double p = MMK_ADD(d->lcdf, MMK_MUL(src_u01(s, purpose), d->tp));
        return -MMK_MUL(mmk_log1p(-p), d->theta);
    } else {
        double lp = mmk_logaddexp(d->loglcdf, MMK_SUB(d->logtp, src_randexp(s, purpose)));
        return -MMK_MUL(mmk_log1mexp(lp), d->theta);
    }
}
/* rand(::Truncated{Normal}) — Distributions.jl src/truncated/normal.jl: z = randnt(lb, ub, tp); mu + sigma z.
 * Only a lower bound here (src/Structs.jl:68), so ub = Inf:
 *   tp > 0.3: r = randn() until lb <= r;
 *   else (then lb > 0, and `span > 2/(lb + sqrt(lb^2+4)) exp(...)` holds for an infinite span):
 *     Robert (1995) exponential proposal: a = (lb + sqrt(lb^2 + 4))/2; repeat r = rand(Exponential(1/a)) + lb,
 *     u = rand() until u < exp(-0.5 (r - a)^2) (and r < ub). */
static double trunc_normal_rand(orc_sim *s) {
    if (s->ins_tp > 0.3) {
        for (;;) {
            double z = src_insert_variate(s, 0);
            if (!(z < s->ins_lb)) return MMK_ADD(s->ins_mu, MMK_MUL(s->ins_sd, z));
            if (s->err) return s->ins_mu;
        }
    }
    const double lb = s->ins_lb;
    const double a = MMK_DIV(MMK_ADD(lb, MMK_SQRT(MMK_ADD(MMK_MUL(lb, lb), 4.0))), 2.0);
    const double theta = MMK_DIV(1.0, a);
    for (;;) {
        double r = MMK_ADD(MMK_MUL(theta, src_insert_variate(s, 1)), lb);
        double u = src_insert_variate(s, 2);
        double t = MMK_SUB(r, a);
        if (u < mmk_exp(MMK_MUL(-0.5, MMK_MUL(t, t)))) return MMK_ADD(s->ins_mu, MMK_MUL(s->ins_sd, r));
        if (s->err) return s->ins_mu;
    }
}
static double round_half_even(double x) { return nearbyint(x); } /* default FE_TONEAREST */

/* ---------- public: lifecycle ---------------------------------------------------------- */
orc_sim *orc_create(const orc_params *p) {
    orc_sim *s = (orc_sim *)calloc(1, sizeof(orc_sim));
    if (!s) return NULL;
    s->par = *p;
    /* src/Structs.jl:57-61 */
    if (p->mol_coverage < 1.0) {
        double t = MMK_MUL(MMK_MUL((double)p->mol_length, p->mol_coverage), 0.25);
        int64_t ti = (int64_t)t;
        s->lowerbound = (double)(ti > 3000 ? ti : 3000);
    } else {
        s->lowerbound = MMK_MUL(p->mol_coverage, MMK_ADD((double)p->insert_length, MMK_MUL(0.75, (double)p->insert_stdev)));
    }
    /* :62 */
    trunc_exp_make(&s->d_mollen, (double)p->mol_length, s->lowerbound, 0, 0.0);
    /* :63-67 */
    if (p->n_molecules > 0) {
        s->nmol_fixed = 0;
        int e = trunc_exp_make(&s->d_nmol, (double)p->n_molecules, 1.0, 1, MMK_MUL(2.5, (double)p->n_molecules));
        if (e) fail(s, e, "n_molecules truncation bounds");
    } else {
        s->nmol_fixed = p->n_molecules < 0 ? -p->n_molecules : 0;
    }
    /* :68 truncated(Normal(ins, sd), lower = 200) */
    s->ins_mu = (double)p->insert_length;
    s->ins_sd = (double)p->insert_stdev;
    s->ins_lb = MMK_DIV(MMK_SUB(200.0, s->ins_mu), s->ins_sd);
    s->ins_tp = 0.5 * erfc(s->ins_lb / sqrt(2.0)); /* used only to pick the sampler (tp > 0.3) */
    if (s->ins_lb != s->ins_lb || s->ins_lb == INFINITY) /* sd = 0 and mu <= 200: randnt never returns */
        fail(s, ORC_ERR_UNSUPPORTED, "insert-size truncation at 200 keeps no mass of the normal law");
    s->mode = 1;
    s->seed = 0;
    s->err_on = p->err_sub_start > 0.0 || p->err_sub_end > 0.0 || p->err_ins > 0.0 || p->err_del > 0.0;
    if (s->err_on) {
        if (p->err_sub_start < 0 || p->err_sub_end < 0 || p->err_ins < 0 || p->err_del < 0 ||
            (p->err_sub_start > p->err_sub_end ? p->err_sub_start : p->err_sub_end) + p->err_ins + p->err_del > 0.5)
            fail(s, ORC_ERR_INVALID_ARG, "error rates must be >= 0 and sum to at most 0.5");
        for (int r = 0; r < 2; ++r) {
            const int S = r ? p->length_r2 : p->length_r1;
            s->err_thr[r] = (uint32_t *)malloc((size_t)S * 4);
            s->err_qb[r] = (uint8_t *)malloc((size_t)S);
            for (int c = 0; c < S; ++c) {
                const double e = S > 1 ? p->err_sub_start + (p->err_sub_end - p->err_sub_start) * (double)c / (double)(S - 1) : p->err_sub_start;
                s->err_thr[r][c] = (uint32_t)floor(e * MMK_ERR_SCALE);
                double et = e + p->err_ins + p->err_del;
                if (et < 1e-4) et = 1e-4;
                int q = (int)floor(-10.0 * log10(et) + 0.5);
                s->err_qb[r][c] = (uint8_t)(q < 2 ? 2 : q > 40 ? 40 : q);
            }
        }
        s->err_tins = (uint32_t)floor(p->err_ins * MMK_ERR_SCALE);
        s->err_tdel = (uint32_t)floor(p->err_del * MMK_ERR_SCALE);
    }
    return s;
}
void orc_destroy(orc_sim *s) {
    if (!s) return;
    for (int i = 0; i < s->n_sch; ++i) {
        free(s->sch[i].chrom);
        if (!s->sch[i].borrowed) free(s->sch[i].seq);
    }
    free(s->sch);
    free(s->r1.p);
    free(s->r2.p);
    free(s->mols);
    free(s->ins_start);
    free(s->ins_stop);
    free(s->rec);
    free(s->cps);
    for (int r = 0; r < 2; ++r) { free(s->err_thr[r]); free(s->err_qb[r]); }
    free(s);
}
int orc_error(const orc_sim *s) { return s->err; }
const char *orc_error_message(const orc_sim *s) { return s->errmsg; }

int orc_add_schema(orc_sim *s, const char *chrom, int haplotype, const char *seq, int64_t len, int64_t reads_required) {
    if (s->n_sch == s->cap_sch) {
        s->cap_sch = s->cap_sch ? 2 * s->cap_sch : 8;
        s->sch = (orc_schema *)realloc(s->sch, (size_t)s->cap_sch * sizeof(orc_schema));
    }
    orc_schema *c = &s->sch[s->n_sch++];
    c->chrom = strdup(chrom);
    c->haplotype = haplotype;
    c->borrowed = 0;
    c->seq = (char *)malloc((size_t)len + 1);
    memcpy(c->seq, seq, (size_t)len);
    c->seq[len] = 0;
    c->len = len;
    c->reads_required = reads_required;
    c->reads_current = 0;
    return 0;
}
/* same, without copying: `seq` must stay valid and already be upper-case IUPAC */
int orc_add_schema_borrowed(orc_sim *s, const char *chrom, int haplotype, const char *seq, int64_t len, int64_t reads_required) {
    if (s->n_sch == s->cap_sch) {
        s->cap_sch = s->cap_sch ? 2 * s->cap_sch : 8;
        s->sch = (orc_schema *)realloc(s->sch, (size_t)s->cap_sch * sizeof(orc_schema));
    }
    orc_schema *c = &s->sch[s->n_sch++];
    c->chrom = strdup(chrom);
    c->haplotype = haplotype;
    c->borrowed = 1;
    c->seq = (char *)seq;
    c->len = len;
    c->reads_required = reads_required;
    c->reads_current = 0;
    return 0;
}
/* keep only byte counts, not the FASTQ text (long benchmark samples) */
void orc_set_discard_output(orc_sim *s, int on) { s->discard = on; }
void orc_set_philox(orc_sim *s, uint64_t seed, uint32_t sample) {
    s->mode = 1;
    s->seed = seed;
    s->sample = sample;
}
void orc_set_replay(orc_sim *s, const orc_variate *stream, size_t n) {
    s->mode = 0;
    s->rp = stream;
    s->rp_n = n;
    s->rp_pos = 0;
}
void orc_set_record(orc_sim *s, int on) { s->record = on; }
void orc_set_max_pairs(orc_sim *s, int64_t n) { s->max_pairs = n; }
void orc_set_digest(orc_sim *s, int64_t window_pairs) {
    s->digest_window = window_pairs;
    s->next_cp = window_pairs;
}
static void push_checkpoint(orc_sim *s, int64_t pairs) {
    if (s->n_cps == s->cap_cps) {
        s->cap_cps = s->cap_cps ? 2 * s->cap_cps : 64;
        s->cps = (orc_checkpoint *)realloc(s->cps, s->cap_cps * sizeof(orc_checkpoint));
    }
    orc_checkpoint *c = &s->cps[s->n_cps++];
    memset(c, 0, sizeof *c);
    c->pairs = pairs;
    c->bytes1 = s->bytes_r1;
    c->bytes2 = s->bytes_r2;
    c->crc1 = s->crc1;
    c->crc2 = s->crc2;
}
const orc_checkpoint *orc_checkpoints(const orc_sim *s, size_t *n) { *n = s->n_cps; return s->cps; }

/* ---------- src/ProcessFASTA.jl --------------------------------------------------------- */
static const char IUPAC[] = "ACMGRSVTWYHKDBN";
/* Normalise one FASTA contig to uppercase IUPAC in place (LongDNA{4} accepts either case and
 * prints uppercase; '-' is the gap symbol). Returns index of first bad char or -1. */
int64_t orc_normalize_bases(char *seq, int64_t len) {
    for (int64_t i = 0; i < len; ++i) {
        char c = seq[i];
        if (c >= 'a' && c <= 'z') c = (char)(c - 32);
        if (c != '-' && (c == 0 || !strchr(IUPAC, c))) return i;
        seq[i] = c;
    }
    return -1;
}
/* mask_N!, src/ProcessFASTA.jl:29-36: every N <- rand((A,C,G,T)) in ascending position.
 * Variate addressing: purpose MASK, g = contig index in file, m = haplotype (file) index,
 * draw index = 0-based position. */
int64_t orc_mask_n(orc_sim *s, char *seq, int64_t len, uint32_t contig_index, uint32_t hap_index) {
    static const char ACGT[4] = {'A', 'C', 'G', 'T'};
    int64_t n = 0;
    mmk_stream st;
    mmk_stream_init(&st, s->seed, s->sample, MMK_P_MASK, contig_index, hap_index);
    for (int64_t i = 0; i < len; ++i) {
        if (seq[i] != 'N') continue;
        int64_t k;
        if (s->mode == 0) {
            const orc_variate *v = rp_pop(s, MMK_P_MASK);
            k = v ? v->v.i : 0;
        } else {
            k = (int64_t)mmk_bits_to_index(mmk_stream_at(&st, (uint32_t)i), 4);
        }
        rec_push(s, MMK_P_MASK, 4, 0, k, 1);
        seq[i] = ACGT[k & 3];
        ++n;
    }
    return n;
}
/* process_fasta, src/ProcessFASTA.jl:47-61 (after optional masking). Returns reads_required or
 * -ORC_ERR_CONTIG_TOO_AMBIGUOUS. */
int64_t orc_reads_required(const char *seq, int64_t len, double coverage, int32_t r1, int32_t r2) {
    int64_t mean_readlen = (int64_t)(r1 + r2) / 2;                    /* :47  sum ÷ 2 */
    int64_t cutoff = (int64_t)round_half_even((double)(r1 + r2) * 0.95); /* :48 */
    int64_t nN = 0;
    for (int64_t i = 0; i < len; ++i) nN += seq[i] == 'N';
    int64_t normalized = len - nN; /* :57 */
    if (normalized < cutoff) return -ORC_ERR_CONTIG_TOO_AMBIGUOUS; /* :58-60 */
    double x = MMK_DIV(MMK_DIV(MMK_MUL(coverage, (double)normalized), (double)mean_readlen), 2.0); /* :61 */
    return (int64_t)x;
}

/* ---------- src/ProcessVariants.jl:143-176 ---------------------------------------------- */
/* Apply one sample-haplotype's variants to a copy of the master contig: all SNPs in file
 * order (seq[pos] = alt[1]), then all non-SNPs in REVERSE file order via
 * spliceinto!(seq, pos:pos+len(ref)-1, alt). `types[i]` is 0 for :snp, 1 for :indel
 * (vartype, :61-69, is evaluated by the caller per VCF record over *all* ALTs).
 * Returns a malloc'd sequence and its length through out params. */
int orc_mutate(const char *master, int64_t len, int64_t n_var, const int64_t *pos, const int32_t *types,
               const int32_t *ref_len, const char *const *alt, char **out_seq, int64_t *out_len) {
    size_t cap = (size_t)len + 64;
    char *q = (char *)malloc(cap);
    if (!q) return ORC_ERR_NOMEM;
    memcpy(q, master, (size_t)len);
    int64_t n = len;
    for (int64_t i = 0; i < n_var; ++i)
        if (types[i] == 0) {
            if (pos[i] < 1 || pos[i] > n) { free(q); return ORC_ERR_BOUNDS; }
            q[pos[i] - 1] = alt[i][0];
        }
    for (int64_t i = n_var - 1; i >= 0; --i)
        if (types[i] != 0) {
            int64_t a = pos[i], b = pos[i] + ref_len[i] - 1; /* 1-based inclusive span to replace */
            if (a < 1 || b > n) { free(q); return ORC_ERR_BOUNDS; }
            int64_t al = (int64_t)strlen(alt[i]);
            int64_t nn = n - ref_len[i] + al;
            if ((size_t)nn + 1 > cap) {
                cap = (size_t)nn * 2 + 64;
                q = (char *)realloc(q, cap);
            }
            memmove(q + (a - 1) + al, q + b, (size_t)(n - b));
            memcpy(q + (a - 1), alt[i], (size_t)al);
            n = nn;
        }
    *out_seq = q;
    *out_len = n;
    return 0;
}
void orc_free(void *p) { free(p); }

/* ---------- src/Barcodes.jl -------------------------------------------------------------- */
/* Barcode for 0-based index k of the deterministic enumerator (first factor fastest).
 * Returns length, or -1 when the enumerator is exhausted (format_barcode(::Nothing), :75). */
int orc_barcode(int bc_format, uint64_t k, char *out) {
    switch (bc_format) {
    case ORC_FMT_HAPLOTAGGING: { /* :12-16, :53-56  product('A',1:96,'C',1:96,'B',1:96,'D',1:96) */
        if (k >= 96ull * 96 * 96 * 96) return -1;
        int a = (int)(k % 96), c = (int)(k / 96 % 96), b = (int)(k / 9216 % 96), d = (int)(k / 884736);
        return sprintf(out, "A%02dC%02dB%02dD%02d", a + 1, c + 1, b + 1, d + 1);
    }
    case ORC_FMT_STLFR: { /* :18-22, :64  product(1:2000,1:2000,1:2000), join "_" */
        if (k >= 8000000000ull) return -1;
        return sprintf(out, "%d_%d_%d", (int)(k % 2000) + 1, (int)(k / 2000 % 2000) + 1, (int)(k / 4000000) + 1);
    }
    case ORC_FMT_TELLSEQ:
    case ORC_FMT_TENX: { /* :24-34, :73-74  product(("ATGC" for i in 1:n)...) */
        int n = bc_format == ORC_FMT_TELLSEQ ? 18 : 16;
        if (k >> (2 * n)) return -1;
        for (int t = 0; t < n; ++t) out[t] = "ATGC"[(k >> (2 * t)) & 3];
        out[n] = 0;
        return n;
    }
    default:
        return -1;
    }
}

/* ---------- src/Breakpoints.jl ----------------------------------------------------------- */
/* get_molecule_size, :149-163 */
static int64_t get_molecule_size(orc_sim *s, int64_t seq_len) {
    double ml = (double)(seq_len + 1);
    int attempts = 0;
    while (ml > (double)seq_len && attempts < 30) {
        attempts++;
        ml = trunc_exp_rand(s, &s->d_mollen, MMK_P_MOLSIZE);
        if (s->err) break;
    }
    if (ml > (double)seq_len) ml = (double)seq_len;
    return (int64_t)ml;
}
/* calculate_n_frags, :171-196 */
static int64_t calculate_n_frags(orc_sim *s, int64_t molsize) {
    const orc_params *p = &s->par;
    double N;
    if (p->singletons > 0.0 && src_u01(s, MMK_P_SINGLETON) <= p->singletons) {
        N = 1.0;
    } else {
        double max_possible = MMK_DIV((double)molsize, (double)p->insert_length);
        if (max_possible == 1.0) {
            N = 1.0;
        } else {
            double _n;
            if (p->mol_coverage < 1.0) {
                double exact = MMK_DIV(MMK_MUL((double)molsize, p->mol_coverage), (double)p->insert_length);
                _n = exact > 2.0 ? exact : 2.0;
            } else
                _n = p->mol_coverage;
            trunc_exp d;
            int e = trunc_exp_make(&d, _n, 2.0, 1, max_possible);
            if (e) {
                fail(s, e, "molecule shorter than two inserts: truncated(Exponential, 2.0, molsize/insert) has lower > upper");
                return 1;
            }
            N = trunc_exp_rand(s, &d, MMK_P_NFRAGS);
        }
    }
    return (int64_t)N;
}
/* calculate_insert_sizes, :205-218 — float sum taken left to right */
static int64_t calculate_insert_sizes(orc_sim *s, int64_t molsize, int64_t **lens_io, size_t *cap_io) {
    int64_t n = calculate_n_frags(s, molsize);
    if (s->err) return 0;
    if ((size_t)n > *cap_io) {
        *cap_io = (size_t)n * 2;
        *lens_io = (int64_t *)realloc(*lens_io, *cap_io * sizeof(int64_t));
    }
    double *tmp = (double *)malloc((size_t)(n > 0 ? n : 1) * sizeof(double));
    int attempts = 0;
    double sum;
    for (;;) {
        sum = 0.0;
        for (int64_t i = 0; i < n; ++i) {
            src_begin_insert(s, (uint32_t)attempts, (uint32_t)i);
            tmp[i] = trunc_normal_rand(s);
            sum = MMK_ADD(sum, tmp[i]);
        }
        if (s->err) break;
        if (sum > (double)molsize && attempts < s->par.attempts) {
            attempts++;
            continue;
        }
        break;
    }
    if (!s->err && sum > (double)molsize) fail(s, ORC_ERR_INSERTS_UNFITTABLE, "FAILED TO FIND PRACTICAL READ SIZES LESS THAN MOLECULE");
    for (int64_t i = 0; i < n; ++i) (*lens_io)[i] = (int64_t)tmp[i];
    free(tmp);
    return n;
}

/* complement of an uppercase IUPAC symbol (BioSequences.complement) */
static char comp(char c) {
    switch (c) {
    case 'A': return 'T'; case 'T': return 'A'; case 'C': return 'G'; case 'G': return 'C';
    case 'M': return 'K'; case 'K': return 'M'; case 'R': return 'Y'; case 'Y': return 'R';
    case 'V': return 'B'; case 'B': return 'V'; case 'H': return 'D'; case 'D': return 'H';
    default: return c; /* W S N - */
    }
}
/* circular_index, :111-121 (literal) or modular wrap (strict). Appends sub-sequence of the
 * 1-based inclusive window [a,b] to `out`; returns its length or -1 on a Julia BoundsError. */
static int64_t circular_index(const orc_schema *c, int64_t a, int64_t b, int strict, bytebuf *out) {
    int64_t n = c->len;
    size_t n0 = out->n;
    if (strict) {
        for (int64_t p = a; p <= b; ++p) {
            int64_t q = ((p - 1) % n + n) % n;
            bb_put(out, &c->seq[q], 1);
        }
        return (int64_t)(out->n - n0);
    }
    if (b > n) {
        /* seq[a:end] * seq[begin:(b % n)] */
        if (a < 1) return -1;
        if (a <= n) bb_put(out, c->seq + (a - 1), (size_t)(n - a + 1)); /* a > n: empty range, no error */
        int64_t r = b % n;
        if (r > 0) bb_put(out, c->seq, (size_t)r);
    } else {
        if (a < 1 || b < a - 1) return -1;
        if (b >= a) bb_put(out, c->seq + (a - 1), (size_t)(b - a + 1));
    }
    return (int64_t)(out->n - n0);
}

static int dec(char *o, int64_t v) { return sprintf(o, "%lld", (long long)v); }

/* Sequencing-error model (NOT in the reference: docs/algorithm.md:11 lists it as future work; specified in DESIGN.md
 * "Error model").  One read of nominal length S: cycle c takes one 32-bit draw (mmk_error_draw: seed, sample, run-wide
 * pair index, read, cycle).  Bits 0-23 u: u < thr_sub[c] substitution; next t_ins insertion; next t_del deletion (thresholds
 * are probabilities scaled by 2^24).  Bits 24-25: replacement selector, bits 26-28: quality jitter.  The template is walked from the read's first base
 * (R2: a2 upwards; R1: bs + r1 - 1 downwards, complemented): an insertion emits ACGT[sel] and consumes no template base; a
 * deletion skips one template base first; positions past the contig ends are clamped (strict circular wrap: modular).
 * A substitution replaces an A/C/G/T by ACGT[(code + 1 + sel % 3) % 4] (other IUPAC symbols stay).  Quality of cycle c:
 * 33 + clamp(qbase[c] + jitter[-2,-1,0,0,0,0,1,2] - 10 [substituted or inserted], 2, 41). */
static void erroneous_read(orc_sim *s, const orc_schema *c, int rd, int64_t first_pos, uint64_t pair, char *seq_out, char *qual_out) {
    static const int JIT[8] = {-2, -1, 0, 0, 0, 0, 1, 2};
    static const char ACGT[4] = {'A', 'C', 'G', 'T'};
    const int S = rd == 1 ? s->par.length_r1 : s->par.length_r2;
    const uint32_t *thr = s->err_thr[rd - 1];
    const uint8_t *qb = s->err_qb[rd - 1];
    const int64_t n = c->len;
    int64_t r = 0; /* template bases consumed */
    for (int cyc = 0; cyc < S; ++cyc) {
        const uint32_t d = mmk_error_draw(s->seed, s->sample, pair, (uint32_t)rd, (uint32_t)cyc);
        const uint32_t u = MMK_ERR_U(d), sel = MMK_ERR_SEL(d), jit = MMK_ERR_JIT(d);
        const int is_sub = u < thr[cyc];
        const int is_ins = !is_sub && (u - thr[cyc]) < s->err_tins;
        const int is_del = !is_sub && !is_ins && (u - thr[cyc] - s->err_tins) < s->err_tdel;
        char b;
        if (is_ins) {
            b = ACGT[sel];
        } else {
            if (is_del) r++;
            int64_t pos = rd == 1 ? first_pos - r : first_pos + r; /* 1-based */
            if (s->par.strict_wrap) pos = ((pos - 1) % n + n) % n + 1;
            else pos = pos < 1 ? 1 : pos > n ? n : pos;
            b = c->seq[pos - 1];
            if (rd == 1) b = comp(b);
            r++;
            if (is_sub) {
                const char *at = b ? strchr("ACGT", b) : NULL;
                if (at) b = ACGT[((int)(at - "ACGT") + 1 + (int)(sel % 3u)) % 4];
            }
        }
        seq_out[cyc] = b;
        int q = (int)qb[cyc] + JIT[jit] - ((is_sub || is_ins) ? 10 : 0);
        q = q < 2 ? 2 : q > 41 ? 41 : q;
        qual_out[cyc] = (char)(33 + q);
    }
}

/* One molecule: get_sequences (:227-251) + seq_breakpoints! (:21-35) + extract_sequences!
 * (:95-109) + format_R1/format_R2 (src/FormatFASTQ.jl). */
static void simulate_molecule(orc_sim *s, int schema_idx, int64_t molsize, const int64_t *lens, int64_t n_frags,
                              const char *barcode, int bclen, bytebuf *t1, bytebuf *t2, uint64_t pair_base) {
    const orc_params *p = &s->par;
    orc_schema *c = &s->sch[schema_idx];
    int64_t seq_len = c->len;
    /* :235 */
    int64_t adjusted_end = p->circular ? seq_len : (seq_len - molsize > 1 ? seq_len - molsize : 1);
    /* :240-242 */
    int64_t start_pos = 1 + src_index(s, MMK_P_START, (uint64_t)adjusted_end);
    int64_t end_pos = start_pos + molsize;
    /* seq_breakpoints! :22-33 */
    int64_t sum = 0;
    for (int64_t i = 0; i < n_frags; ++i) sum += lens[i];
    int64_t leeway = (end_pos - start_pos + 1) - sum;
    int64_t cursor = start_pos, remaining = leeway;
    if (s->n_ins + (size_t)n_frags > s->cap_ins) {
        s->cap_ins = (s->n_ins + (size_t)n_frags) * 2;
        s->ins_start = (int64_t *)realloc(s->ins_start, s->cap_ins * sizeof(int64_t));
        s->ins_stop = (int64_t *)realloc(s->ins_stop, s->cap_ins * sizeof(int64_t));
    }
    int64_t *bs = s->ins_start + s->n_ins, *be = s->ins_stop + s->n_ins;
    for (int64_t i = 0; i < n_frags; ++i) {
        int64_t gap = src_gap(s, (uint32_t)i, (uint64_t)remaining + 1);
        int64_t st = cursor + gap;
        bs[i] = st;
        be[i] = st + lens[i] - 1;
        cursor = st + lens[i];
        remaining -= gap;
    }
    s->n_ins += (size_t)n_frags;
    if (s->err) return;

    /* extract_sequences! :95-109 (N filter is dead code: count(==("N"), ::String) == 0) */
    t1->n = 0; /* R1 strings back to back */
    t2->n = 0;
    int64_t *l1 = (int64_t *)malloc((size_t)n_frags * 2 * sizeof(int64_t)), *l2 = l1 + n_frags;
    for (int64_t i = 0; i < n_frags; ++i) {
        /* R1 = reverse_complement(circular_index(bp.start : bp.start + r1 - 1)) */
        size_t o = t1->n;
        l1[i] = circular_index(c, bs[i], bs[i] + p->length_r1 - 1, p->strict_wrap, t1);
        if (l1[i] < 0) { fail(s, ORC_ERR_BOUNDS, "R1 window out of bounds"); free(l1); return; }
        for (int64_t x = 0, y = l1[i] - 1; x <= y; ++x, --y) {
            char cx = comp((char)t1->p[o + x]), cy = comp((char)t1->p[o + y]);
            t1->p[o + x] = (uint8_t)cy;
            t1->p[o + y] = (uint8_t)cx;
        }
        /* R2 = circular_index(bp.stop - r2 + 1 : bp.stop) */
        l2[i] = circular_index(c, be[i] - p->length_r2 + 1, be[i], p->strict_wrap, t2);
        if (l2[i] < 0) { fail(s, ORC_ERR_BOUNDS, "R2 window out of bounds"); free(l1); return; }
    }

    /* format_R1 / format_R2 */
    char hdr[512], num[32];
    int hl = snprintf(hdr, sizeof hdr, "@%s:%d|%lld:%lld|", c->chrom, c->haplotype, (long long)start_pos, (long long)end_pos);
    int64_t q1 = l1[0], q2 = l2[0]; /* _qual = "I"^length(first read of the molecule) */
    if (p->fq_format == ORC_FMT_TENX) q1 += bclen;
    size_t o1 = 0, o2 = 0;
    for (int64_t i = 0; i < n_frags; ++i) {
        for (int rd = 1; rd <= 2; ++rd) {
            bytebuf *out = rd == 1 ? &s->r1 : &s->r2;
            bb_put(out, hdr, (size_t)hl);
            bb_put(out, num, (size_t)dec(num, bs[i]));
            bb_put(out, ":", 1);
            bb_put(out, num, (size_t)dec(num, be[i]));
            char rc = (char)('0' + rd);
            switch (p->fq_format) {
            case ORC_FMT_STLFR: /* FormatFASTQ.jl:16,85 */
                bb_put(out, "#", 1); bb_put(out, barcode, (size_t)bclen); bb_put(out, " ", 1); bb_put(out, &rc, 1); bb_put(out, ":N:0:ATGACA\n", 12);
                break;
            case ORC_FMT_HAPLOTAGGING: /* :28,97 */
                bb_put(out, "/", 1); bb_put(out, &rc, 1); bb_put(out, "\tBX:Z:", 6); bb_put(out, barcode, (size_t)bclen); bb_put(out, "\n", 1);
                break;
            case ORC_FMT_TELLSEQ: /* :40,109 */
                bb_put(out, ":", 1); bb_put(out, barcode, (size_t)bclen); bb_put(out, " ", 1); bb_put(out, &rc, 1); bb_put(out, ":N:0:ATGACA\n", 12);
                break;
            case ORC_FMT_TENX: /* :52,121 */
                bb_put(out, "|", 1); bb_put(out, barcode, (size_t)bclen); bb_put(out, " ", 1); bb_put(out, &rc, 1); bb_put(out, ":N:0:ATGACA\n", 12);
                if (rd == 1) bb_put(out, barcode, (size_t)bclen);
                break;
            default: /* standard :64,133 */
                bb_put(out, "/", 1); bb_put(out, &rc, 1); bb_put(out, "\tVX:i:1\tBX:Z:", 13); bb_put(out, barcode, (size_t)bclen); bb_put(out, "\n", 1);
                break;
            }
            const int64_t li = rd == 1 ? l1[i] : l2[i], qi = rd == 1 ? q1 : q2;
            const int S = rd == 1 ? p->length_r1 : p->length_r2;
            const int inl = (rd == 1 && p->fq_format == ORC_FMT_TENX) ? bclen : 0;
            if (s->err_on && li == S && qi == S + inl) { /* the model covers reads of nominal length (others: the literal-wrap artefacts) */
                char *es = (char *)malloc((size_t)S * 2), *eq = es + S;
                erroneous_read(s, c, rd, rd == 1 ? bs[i] + p->length_r1 - 1 : be[i] - p->length_r2 + 1, pair_base + (uint64_t)i, es, eq);
                bb_put(out, es, (size_t)S);
                bb_put(out, "\n+\n", 3);
                bb_fill(out, 'I', (size_t)inl); /* the inline barcode is not sequenced template: constant quality */
                bb_put(out, eq, (size_t)S);
                bb_put(out, "\n", 1);
                free(es);
            } else {
                bb_put(out, (rd == 1 ? t1->p + o1 : t2->p + o2), (size_t)li);
                bb_put(out, "\n+\n", 3);
                bb_fill(out, 'I', (size_t)qi);
                bb_put(out, "\n", 1);
            }
            if (rd == 1) o1 += (size_t)l1[i]; else o2 += (size_t)l2[i];
        }
    }
    free(l1);
    /* bookkeeping for plan-level checks */
    orc_mol *m = &s->mols[s->n_mols - 1];
    m->start = start_pos;
}

/* ---------- src/Main.jl:44-59 / :113-126 -------------------------------------------------- */
int orc_run(orc_sim *s) {
    if (s->err) return s->err;
    int K = s->n_sch;
    int *live = (int *)malloc((size_t)(K > 0 ? K : 1) * sizeof(int));
    int n_live = K;
    for (int i = 0; i < K; ++i) live[i] = i;
    int64_t *lens = NULL;
    size_t lens_cap = 0;
    int *targets = NULL;
    size_t targets_cap = 0;
    bytebuf t1 = {0}, t2 = {0};
    uint64_t bc_counter = 0;
    int64_t total_pairs = 0;
    uint64_t g = 0;
    while (n_live > 0 && !s->err) { /* while !isempty(schema) */
        src_begin_group(s, g);
        /* get_n_molecules, src/Common.jl:89-91 */
        int64_t n_mol;
        if (s->par.n_molecules > 0) {
            double x = trunc_exp_rand(s, &s->d_nmol, MMK_P_NMOL);
            n_mol = (int64_t)round_half_even(x);
        } else
            n_mol = s->nmol_fixed;
        if (s->err) break;
        /* rand(candidates, n_mol), src/Main.jl:47 — uniform index into the live list kept in
         * first-seen order (Julia samples a Dict KeySet; order there is hash-layout dependent) */
        if ((size_t)n_mol > targets_cap) {
            targets_cap = (size_t)n_mol * 2;
            targets = (int *)realloc(targets, targets_cap * sizeof(int));
        }
        for (int64_t j = 0; j < n_mol; ++j) targets[j] = live[src_index(s, MMK_P_TARGET, (uint64_t)n_live)];
        for (int64_t j = 0; j < n_mol && !s->err; ++j) {
            int target = targets[j];
            src_begin_molecule(s, (uint32_t)j);
            int64_t molsize = get_molecule_size(s, s->sch[target].len);     /* :49 */
            if (s->err) break;
            int64_t nf = calculate_insert_sizes(s, molsize, &lens, &lens_cap); /* :50 */
            if (s->err) break;
            char bc[32];
            int bclen = orc_barcode(s->par.bc_format, bc_counter, bc);       /* :51 get_next! */
            if (bclen < 0) { fail(s, ORC_ERR_BARCODES_EXHAUSTED, "There are no more barcodes available for the selected barcode type."); break; }
            if (s->n_mols == s->cap_mols) {
                s->cap_mols = s->cap_mols ? s->cap_mols * 2 : 1024;
                s->mols = (orc_mol *)realloc(s->mols, s->cap_mols * sizeof(orc_mol));
            }
            orc_mol *m = &s->mols[s->n_mols++];
            m->group = (int64_t)g; m->slot = (int32_t)j; m->schema = target; m->molsize = molsize;
            m->n_frags = nf; m->first_read = (int64_t)s->n_ins; m->barcode_index = (int64_t)bc_counter; m->start = 0;
            bc_counter++;
            simulate_molecule(s, target, molsize, lens, nf, bc, bclen, &t1, &t2, (uint64_t)total_pairs); /* :51-53 */
            s->sch[target].reads_current += nf;                                /* :54 */
            total_pairs += nf;
        }
        /* filter!(is_incomplete, schema), :58 */
        int w = 0;
        for (int i = 0; i < n_live; ++i)
            if (s->sch[live[i]].reads_current < s->sch[live[i]].reads_required) live[w++] = live[i];
        n_live = w;
        g++;
        if (s->discard) {
            s->bytes_r1 += s->r1.n; s->bytes_r2 += s->r2.n;
            s->r1.n = 0; s->r2.n = 0; s->n_ins = 0; s->n_mols = 0;
        } else if (s->digest_window > 0) {
            for (size_t o = 0; o < s->r1.n; o += 1u << 30) s->crc1 = (uint32_t)crc32(s->crc1, s->r1.p + o, (uInt)(s->r1.n - o < (1u << 30) ? s->r1.n - o : (1u << 30)));
            for (size_t o = 0; o < s->r2.n; o += 1u << 30) s->crc2 = (uint32_t)crc32(s->crc2, s->r2.p + o, (uInt)(s->r2.n - o < (1u << 30) ? s->r2.n - o : (1u << 30)));
            s->bytes_r1 += s->r1.n; s->bytes_r2 += s->r2.n;
            s->r1.n = 0; s->r2.n = 0;
            if (total_pairs >= s->next_cp) {
                push_checkpoint(s, total_pairs);
                s->next_cp = (total_pairs / s->digest_window + 1) * s->digest_window;
            }
        }
        s->total_pairs = total_pairs;
        if (s->max_pairs > 0 && total_pairs >= s->max_pairs) break;
    }
    if (s->digest_window > 0 && !s->err && (s->n_cps == 0 || s->cps[s->n_cps - 1].pairs != total_pairs)) push_checkpoint(s, total_pairs);
    s->n_groups = (int64_t)g;
    free(live); free(lens); free(targets); free(t1.p); free(t2.p);
    return s->err;
}

/* ---------- accessors -------------------------------------------------------------------- */
const uint8_t *orc_r1(const orc_sim *s, size_t *n) { *n = s->r1.n; return s->r1.p; }
const uint8_t *orc_r2(const orc_sim *s, size_t *n) { *n = s->r2.n; return s->r2.p; }
int64_t orc_n_groups(const orc_sim *s) { return s->n_groups; }
int64_t orc_total_pairs(const orc_sim *s) { return s->total_pairs; }
uint64_t orc_total_bytes(const orc_sim *s) { return (s->discard || s->digest_window > 0) ? s->bytes_r1 + s->bytes_r2 : s->r1.n + s->r2.n; }
int64_t orc_n_molecules(const orc_sim *s) { return (int64_t)s->n_mols; }
int64_t orc_n_pairs(const orc_sim *s) { return (int64_t)s->n_ins; }
const orc_mol *orc_molecules(const orc_sim *s) { return s->mols; }
const int64_t *orc_insert_starts(const orc_sim *s) { return s->ins_start; }
const int64_t *orc_insert_stops(const orc_sim *s) { return s->ins_stop; }
const orc_variate *orc_recorded(const orc_sim *s, size_t *n) { *n = s->rec_n; return s->rec; }
size_t orc_replay_consumed(const orc_sim *s) { return s->rp_pos; }
int64_t orc_schema_reads(const orc_sim *s, int i, int64_t *required) { *required = s->sch[i].reads_required; return s->sch[i].reads_current; }
double orc_lowerbound(const orc_sim *s) { return s->lowerbound; }

/* ---------- known-answer helpers exposed to tests --------------------------------------- */
void orc_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { mmk_philox4x32_10(ctr, key, out); }
double orc_det_log(double x) { return mmk_log(x); }
double orc_det_exp(double x) { return mmk_exp(x); }
void orc_trunc_exp_info(double theta, double lower, int has_upper, double upper, double out[4]) {
    trunc_exp d;
    int e = trunc_exp_make(&d, theta, lower, has_upper, upper);
    out[0] = e; out[1] = d.lcdf; out[2] = d.ucdf; out[3] = d.tp;
}
/* draw n values from a truncated exponential / the insert normal with the Philox source */
void orc_sample_trunc_exp(uint64_t seed, double theta, double lower, int has_upper, double upper, int64_t n, double *out) {
    orc_params p; memset(&p, 0, sizeof p); p.insert_length = 500; p.insert_stdev = 50; p.mol_length = 80000; p.mol_coverage = 0.2; p.n_molecules = 2;
    orc_sim *s = orc_create(&p);
    orc_set_philox(s, seed, 0);
    trunc_exp d; trunc_exp_make(&d, theta, lower, has_upper, upper);
    for (int64_t i = 0; i < n; ++i) { src_begin_group(s, (uint64_t)i); out[i] = trunc_exp_rand(s, &d, MMK_P_NFRAGS); }
    orc_destroy(s);
}
/* the definition the device's threshold table must reproduce: trunc(Int, mu + sd * z(bits)), 0 when z < lb */
void orc_insert_len_from_bits(int32_t mu, int32_t sd, const uint64_t *bits, uint32_t *out, int64_t n) {
    const double dmu = (double)mu, dsd = (double)sd, lb = MMK_DIV(MMK_SUB(200.0, dmu), dsd);
    for (int64_t i = 0; i < n; ++i) {
        const double z = mmk_bits_to_normal(bits[i]);
        out[i] = z < lb ? 0u : (uint32_t)(int64_t)MMK_ADD(dmu, MMK_MUL(dsd, z));
    }
}
void orc_sample_insert(uint64_t seed, int32_t mu, int32_t sd, int64_t n, double *out) {
    orc_params p; memset(&p, 0, sizeof p); p.insert_length = mu; p.insert_stdev = sd; p.mol_length = 80000; p.mol_coverage = 0.2; p.n_molecules = 2;
    orc_sim *s = orc_create(&p);
    orc_set_philox(s, seed, 0);
    for (int64_t i = 0; i < n; ++i) { src_begin_group(s, (uint64_t)i); src_begin_insert(s, 0, (uint32_t)(i & 0xffff)); out[i] = trunc_normal_rand(s); }
    orc_destroy(s);
}
