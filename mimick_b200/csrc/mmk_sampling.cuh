// mmk_sampling.cuh — device-side sampling of one barcode group / one molecule.
//
// Behavioural contract (reference = mimick/MimickLinkedReads.jl/src):
//   draw_nmol            get_n_molecules            Common.jl:89-91, Structs.jl:63-67
//   molecule_shape       get_molecule_size          Breakpoints.jl:149-163
//                        calculate_n_frags          Breakpoints.jl:171-196
//   molecule_place       calculate_insert_sizes     Breakpoints.jl:205-218
//                        get_sequences (start)      Breakpoints.jl:235-242
//                        seq_breakpoints!           Breakpoints.jl:21-35
// Every function is templated on a recorder so the same code path can also report the variates
// it consumed, in the reference's draw order, for the oracle's replay mode.
#pragma once
#include "mmk_internal.h"

struct NoRec {
    __device__ __forceinline__ void put(uint32_t, uint32_t, uint64_t) {}
};
struct CountRec {
    uint32_t n = 0;
    __device__ __forceinline__ void put(uint32_t, uint32_t, uint64_t) { ++n; }
};
struct WriteRec {
    mmk_variate *p;
    __device__ __forceinline__ void put(uint32_t tag, uint32_t aux, uint64_t bits) {
        p->tag = tag;
        p->aux = aux;
        p->v.i = (int64_t)bits;
        ++p;
    }
};

// Build truncated(Exponential(theta), lower, upper) the way Distributions.jl does.
__device__ __forceinline__ bool trunc_exp_make(TruncExp &d, double theta, double lower, bool has_upper, double upper) {
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
        double logucdf = mmk_log1mexp(-zu);
        d.logtp = mmk_logsubexp(d.loglcdf, logucdf);
    } else {
        d.logtp = mmk_log1mexp(d.loglcdf);
    }
    d.tp = mmk_exp(d.logtp);
    return true;
}

// rand(::Truncated{Exponential}): rejection when tp > 0.25, inverse cdf when tp > sqrt(eps),
// log-space inverse otherwise.
template <class R>
__device__ __forceinline__ double trunc_exp_rand(mmk_stream &s, const TruncExp &d, uint32_t tag, R &rec) {
    if (d.tp > 0.25) {
        for (;;) {
            double e = mmk_next_exp(&s);
            rec.put(tag, 0, mmk_d2u(e));
            double r = MMK_MUL(e, d.theta);
            if (r >= d.lower && (!d.has_upper || r <= d.upper)) return r;
        }
    } else if (d.tp > 1.4901161193847656e-08) {
        double u = mmk_next_u01(&s);
        rec.put(tag, 0, mmk_d2u(u));
        double p = MMK_ADD(d.lcdf, MMK_MUL(u, d.tp));
        return -MMK_MUL(mmk_log1p(-p), d.theta);
    } else {
        double e = mmk_next_exp(&s);
        rec.put(tag, 0, mmk_d2u(e));
        double lp = mmk_logaddexp(d.loglcdf, MMK_SUB(d.logtp, e));
        return -MMK_MUL(mmk_log1mexp(lp), d.theta);
    }
}

// round to nearest, ties to even (Julia round(x, digits = 0)); x >= 0 here
__device__ __forceinline__ double round_half_even(double x) { return rint(x); }

template <class R>
__device__ __forceinline__ uint32_t draw_nmol(const SimConst &sc, uint64_t g, R &rec) {
    if (sc.nmol_fixed > 0) return (uint32_t)sc.nmol_fixed;
    mmk_stream s;
    mmk_stream_init(&s, sc.seed, sc.sample, MMK_P_NMOL, g, 0);
    double x = trunc_exp_rand(s, sc.nmol, MMK_P_NMOL, rec);
    return (uint32_t)(long long)round_half_even(x);
}

// Target pick of slot m in group g: uniform index into the live list (Main.jl:47).
__device__ __forceinline__ uint32_t draw_target(const SimConst &sc, uint64_t g, uint32_t m, uint32_t n_live) {
    mmk_stream s;
    mmk_stream_init(&s, sc.seed, sc.sample, MMK_P_TARGET, g, 0);
    return (uint32_t)mmk_bits_to_index(mmk_stream_at(&s, m), n_live);
}

// Molecule length and number of inserts. Returns status (0 ok).
template <class R>
__device__ __forceinline__ uint32_t molecule_shape(const SimConst &sc, uint64_t g, uint32_t m, uint32_t seq_len,
                                                   uint32_t &molsize, uint32_t &nfrags, R &rec) {
    mmk_stream s;
    mmk_stream_init(&s, sc.seed, sc.sample, MMK_P_MOLSIZE, g, m);
    double ml = (double)seq_len + 1.0;
    int attempts = 0;
    while (ml > (double)seq_len && attempts < 30) {
        ++attempts;
        ml = trunc_exp_rand(s, sc.mollen, MMK_P_MOLSIZE, rec);
    }
    if (ml > (double)seq_len) ml = (double)seq_len;
    molsize = (uint32_t)(long long)ml;

    double N;
    bool single = false;
    if (sc.singletons > 0.0) {
        mmk_stream_init(&s, sc.seed, sc.sample, MMK_P_SINGLETON, g, m);
        double u = mmk_next_u01(&s);
        rec.put(MMK_P_SINGLETON, 0, mmk_d2u(u));
        single = u <= sc.singletons;
    }
    uint32_t status = 0;
    if (single) {
        N = 1.0;
    } else {
        double max_possible = MMK_DIV((double)molsize, (double)sc.insert_length);
        if (max_possible == 1.0) {
            N = 1.0;
        } else {
            double n_mean;
            if (sc.mol_coverage < 1.0) {
                double exact = MMK_DIV(MMK_MUL((double)molsize, sc.mol_coverage), (double)sc.insert_length);
                n_mean = exact > 2.0 ? exact : 2.0;
            } else
                n_mean = sc.mol_coverage;
            TruncExp d;
            // Distributions.jl picks the sampler from the kept mass tp (> 0.25: rejection).  tp is
            // exp(-2/n) - exp(-max/n) up to rounding; unless that is within 1e-9 of 0.25 the decision is
            // the same as with the log-space tp of `truncated`, and the rejection sampler needs no cdf.
            const double tp_quick = MMK_SUB(mmk_exp(-MMK_DIV(2.0, n_mean)), mmk_exp(-MMK_DIV(max_possible, n_mean)));
            if (2.0 <= max_possible && tp_quick > 0.25 + 1e-9) {
                d.theta = n_mean;
                d.lower = 2.0;
                d.upper = max_possible;
                d.has_upper = 1;
                d.tp = tp_quick;  // only compared with 0.25 from here on
                mmk_stream_init(&s, sc.seed, sc.sample, MMK_P_NFRAGS, g, m);
                N = trunc_exp_rand(s, d, MMK_P_NFRAGS, rec);
            } else if (!trunc_exp_make(d, n_mean, 2.0, true, max_possible)) {
                status = MMK_ERR_TRUNCATION_BOUNDS;
                N = 1.0;
            } else {
                mmk_stream_init(&s, sc.seed, sc.sample, MMK_P_NFRAGS, g, m);
                N = trunc_exp_rand(s, d, MMK_P_NFRAGS, rec);
            }
        }
    }
    nfrags = (uint32_t)(long long)N;
    return status;
}

// One insert length: rand(truncated(Normal(mu, sd), lower = 200)) (Breakpoints.jl:207; Distributions.jl
// src/truncated/normal.jl `randnt`, lower bound only, ub = Inf):
//   tp > 0.3   z = randn() until z >= lb                                     (plain rejection)
//   tp <= 0.3  Robert (1995), exponential proposal (lb > 0 follows from tp <= 0.3; the branch test
//              `span > ...` is always true for an infinite span): a = (lb + sqrt(lb^2 + 4)) / 2;
//              r = randexp() / a + lb, u = rand() until u < exp(-(r - a)^2 / 2)
//   result     mu + sd * z
// `s` is the insert's own stream (mmk_insert_stream_init), positioned at its first variate, so any thread
// can make the draw; the replay order (insert 0's variates, then insert 1's, ...) is the reference's.
template <class R>
__device__ __forceinline__ double draw_insert_from(const SimConst &sc, mmk_stream &s, R &rec) {
    if (sc.ins_robert) {
        for (;;) {
            const double e = mmk_bits_to_exp(mmk_insert_next64(&s));
            rec.put(MMK_P_INSERT, 0, mmk_d2u(e));
            const double r = MMK_ADD(MMK_MUL(sc.ins_theta, e), sc.ins_lb);
            const double u = mmk_bits_to_u01(mmk_insert_next64(&s));
            rec.put(MMK_P_INSERT, 0, mmk_d2u(u));
            const double t = MMK_SUB(r, sc.ins_a);
            if (u < mmk_exp(MMK_MUL(-0.5, MMK_MUL(t, t)))) return MMK_ADD(sc.ins_mu, MMK_MUL(sc.ins_sd, r));
        }
    }
    for (;;) {
        const double z = mmk_bits_to_normal(mmk_insert_next64(&s));
        rec.put(MMK_P_INSERT, 0, mmk_d2u(z));
        if (!(z < sc.ins_lb)) return MMK_ADD(sc.ins_mu, MMK_MUL(sc.ins_sd, z));
    }
}
// `key` = mmk_round_key(seed, redraw round)
template <class R>
__device__ __forceinline__ double draw_insert(const SimConst &sc, uint64_t key, uint64_t g, uint32_t m, uint32_t i, R &rec) {
    mmk_stream s;
    mmk_insert_stream_init(&s, key, sc.sample, g, m, i);
    return draw_insert_from(sc, s, rec);
}
// the 64 bits of insert i's gap draw: draw 1 of its round-0 stream
__device__ __forceinline__ uint64_t gap_bits(const SimConst &sc, uint64_t g, uint32_t m, uint32_t i) {
    mmk_stream s;
    mmk_insert_stream_init(&s, sc.seed, sc.sample, g, m, i);
    return mmk_stream_at(&s, MMK_INSERT_GAP_DRAW);
}
__device__ __forceinline__ uint32_t insert_len_of(double x) {
    uint32_t li = (uint32_t)(long long)x;
    return li > 0x00ffffffu ? 0x00ffffffu : li;  // unreachable: mmk_create bounds insert_length + 13 sd below 2^24
}

// Insert lengths, molecule start, gap chain.  Plan entry of insert i:
//   ins[i].x = start_i - molecule start
//   ins[i].y = length (low 24 bits) | (molecule index - first molecule of the insert's tile) << 24
// The tile-relative molecule index lets the emit kernel find a pair's molecule without searching.
#define MMK_INS_LEN(y) ((y) & 0x00ffffffu)
#define MMK_INS_DM(y) ((y) >> 24)
struct TileRef {
    const uint32_t *tile_mol0;  // first (shard-relative) molecule of every tile
    uint64_t pair0;             // shard-relative index of this molecule's first pair
    uint32_t mol;               // shard-relative molecule index
    uint32_t tile_pairs;
};
template <class R>
__device__ __forceinline__ uint32_t molecule_place(const SimConst &sc, uint64_t g, uint32_t m, uint32_t seq_len,
                                                   uint32_t molsize, uint32_t nfrags, uint2 *ins, const TileRef &tr,
                                                   int64_t &start, R &rec) {
    mmk_stream s;
    int attempts = 0;
    double sum;
    uint64_t isum;
    for (;;) {
        sum = 0.0;
        isum = 0;
        const uint64_t key = mmk_round_key(sc.seed, (uint32_t)attempts);
        for (uint32_t i = 0; i < nfrags; ++i) {
            double x = draw_insert(sc, key, g, m, i, rec);
            sum = MMK_ADD(sum, x);
            uint32_t li = insert_len_of(x);
            ins[i].y = li;
            isum += li;
        }
        if (sum > (double)molsize && attempts < sc.attempts) {
            ++attempts;
            continue;
        }
        break;
    }
    uint32_t status = 0;
    if (sum > (double)molsize) status = MMK_ERR_INSERTS_UNFITTABLE;

    // start_pos = rand(1:adjusted_end)
    uint64_t adjusted_end = sc.circular ? (uint64_t)seq_len : (seq_len > molsize + 1u ? (uint64_t)(seq_len - molsize) : 1ull);
    mmk_stream_init(&s, sc.seed, sc.sample, MMK_P_START, g, m);
    uint64_t sidx = mmk_next_index(&s, adjusted_end);
    rec.put(MMK_P_START, (uint32_t)adjusted_end, sidx);
    start = (int64_t)sidx + 1;
    if (status) {
        for (uint32_t i = 0; i < nfrags; ++i) ins[i].x = 0;
        return status;
    }
    // gap chain
    uint64_t remaining = (uint64_t)molsize + 1 - isum;
    uint64_t cursor = 0;  // relative to start
    for (uint32_t i = 0; i < nfrags; ++i) {
        uint64_t gap = mmk_bits_to_index(gap_bits(sc, g, m, i), remaining + 1);
        rec.put(MMK_P_GAP, (uint32_t)(remaining + 1), gap);
        uint64_t st = cursor + gap;
        uint32_t li = MMK_INS_LEN(ins[i].y);
        uint32_t dm = tr.mol - tr.tile_mol0[(tr.pair0 + i) >> (31 - __clz(tr.tile_pairs))];  // power-of-two tiles
        ins[i] = make_uint2((uint32_t)st, li | (dm << 24));
        cursor = st + li;
        remaining -= gap;
    }
    return 0;
}
