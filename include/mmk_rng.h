/* mmk_rng.h — counter-based variate streams shared by the sm_100a kernels and the
 * CPU checker.
 *
 * Every random number in a simulation is a pure function of
 *     (seed, sample, purpose, group g, slot m, draw index)
 * through Philox4x32-10 (Salmon et al., SC'11; constants as in Random123).  There is no
 * sequential generator state anywhere, so a molecule can be sampled by any thread of any
 * GPU and still consume exactly the variates the reference's loop would have consumed at
 * that point of its sequential stream (SURVEY.md §8a-bis lists the order, taken from
 * mimick/MimickLinkedReads.jl/src/Main.jl:44-59 and src/Breakpoints.jl).
 *
 * The transforms from 64 random bits to u01 / randexp / randn / randint use only IEEE-754
 * add, sub, mul, div and sqrt, each rounded once (no FMA contraction: __dmul_rn & co. on the
 * device, -ffp-contract=off on the host), plus a polynomial log/exp written in those
 * operations.  Host and device therefore produce bit-identical doubles, and replay parity
 * tests compare sampling *logic*, not two libm implementations.
 *
 * Plain C99 / CUDA C++; no dependencies.
 */
#ifndef MMK_RNG_H
#define MMK_RNG_H

#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define MMK_HD __host__ __device__ __forceinline__
#else
#define MMK_HD static inline
#endif

/* ---- single-rounding arithmetic ------------------------------------------------------- */
#if defined(__CUDA_ARCH__)
#define MMK_MUL(a, b) __dmul_rn((a), (b))
#define MMK_ADD(a, b) __dadd_rn((a), (b))
#define MMK_SUB(a, b) __dsub_rn((a), (b))
#define MMK_DIV(a, b) __ddiv_rn((a), (b))
#define MMK_SQRT(a) __dsqrt_rn((a))
#else
#include <math.h>
#define MMK_MUL(a, b) ((double)(a) * (double)(b))
#define MMK_ADD(a, b) ((double)(a) + (double)(b))
#define MMK_SUB(a, b) ((double)(a) - (double)(b))
#define MMK_DIV(a, b) ((double)(a) / (double)(b))
#define MMK_SQRT(a) sqrt((double)(a))
#endif

MMK_HD uint64_t mmk_d2u(double d) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t u;
    memcpy(&u, &d, 8);
    return u;
#endif
}
MMK_HD double mmk_u2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double d;
    memcpy(&d, &u, 8);
    return d;
#endif
}
MMK_HD uint64_t mmk_mulhi64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((unsigned __int128)a * (unsigned __int128)b) >> 64);
#endif
}
MMK_HD uint32_t mmk_mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

/* ---- deterministic elementary functions ---------------------------------------------- */
#define MMK_LN2_HI 6.93147180369123816490e-01 /* 0x3fe62e42fee00000 */
#define MMK_LN2_LO 1.90821492927058770002e-10 /* 0x3dea39ef35793c76 */
#define MMK_INV_LN2 1.44269504088896338700e+00
#define MMK_SQRT2 1.41421356237309514547e+00
#define MMK_INF (mmk_u2d(0x7ff0000000000000ull))

/* log(x) for finite normal x > 0.  |rel err| ~ 2e-16 (checked against libm in tests). */
MMK_HD double mmk_log(double x) {
    uint64_t b = mmk_d2u(x);
    int e = (int)(b >> 52) - 1023;
    double m = mmk_u2d((b & 0x000fffffffffffffull) | 0x3ff0000000000000ull); /* [1,2) */
    if (m > MMK_SQRT2) {
        m = MMK_MUL(m, 0.5);
        e += 1;
    }
    double f = MMK_SUB(m, 1.0);
    double s = MMK_DIV(f, MMK_ADD(2.0, f));
    double z = MMK_MUL(s, s);
    /* 2*atanh(s) = 2s * sum_k z^k/(2k+1); |s| <= 0.1716 so 12 terms reach 1e-19 */
    double p = 1.0 / 25.0;
    p = MMK_ADD(MMK_MUL(p, z), 1.0 / 23.0);
    p = MMK_ADD(MMK_MUL(p, z), 1.0 / 21.0);
    p = MMK_ADD(MMK_MUL(p, z), 1.0 / 19.0);
    p = MMK_ADD(MMK_MUL(p, z), 1.0 / 17.0);
    p = MMK_ADD(MMK_MUL(p, z), 1.0 / 15.0);
    p = MMK_ADD(MMK_MUL(p, z), 1.0 / 13.0);
    p = MMK_ADD(MMK_MUL(p, z), 1.0 / 11.0);
    p = MMK_ADD(MMK_MUL(p, z), 1.0 / 9.0);
    p = MMK_ADD(MMK_MUL(p, z), 1.0 / 7.0);
    p = MMK_ADD(MMK_MUL(p, z), 1.0 / 5.0);
    p = MMK_ADD(MMK_MUL(p, z), 1.0 / 3.0);
    /* log(m) = 2s + 2s*z*p, kept as two terms so the small correction is added last */
    double two_s = MMK_ADD(s, s);
    double corr = MMK_MUL(MMK_MUL(two_s, z), p);
    double de = (double)e;
    double lo = MMK_ADD(MMK_MUL(de, MMK_LN2_LO), corr);
    return MMK_ADD(MMK_MUL(de, MMK_LN2_HI), MMK_ADD(two_s, lo));
}

/* exp(x); returns 0 below the normal range and +inf above it (deterministic flush). */
MMK_HD double mmk_exp(double x) {
    if (x != x) return x;
    if (x > 709.0) return MMK_INF;
    if (x < -708.0) return 0.0;
    double t = MMK_MUL(x, MMK_INV_LN2);
    long long k = (long long)(t >= 0.0 ? MMK_ADD(t, 0.5) : MMK_SUB(t, 0.5));
    double dk = (double)k;
    double r = MMK_SUB(MMK_SUB(x, MMK_MUL(dk, MMK_LN2_HI)), MMK_MUL(dk, MMK_LN2_LO));
    /* Taylor to degree 14 on |r| <= 0.3466 */
    double p = 1.0 / 87178291200.0;
    p = MMK_ADD(MMK_MUL(p, r), 1.0 / 6227020800.0);
    p = MMK_ADD(MMK_MUL(p, r), 1.0 / 479001600.0);
    p = MMK_ADD(MMK_MUL(p, r), 1.0 / 39916800.0);
    p = MMK_ADD(MMK_MUL(p, r), 1.0 / 3628800.0);
    p = MMK_ADD(MMK_MUL(p, r), 1.0 / 362880.0);
    p = MMK_ADD(MMK_MUL(p, r), 1.0 / 40320.0);
    p = MMK_ADD(MMK_MUL(p, r), 1.0 / 5040.0);
    p = MMK_ADD(MMK_MUL(p, r), 1.0 / 720.0);
    p = MMK_ADD(MMK_MUL(p, r), 1.0 / 120.0);
    p = MMK_ADD(MMK_MUL(p, r), 1.0 / 24.0);
    p = MMK_ADD(MMK_MUL(p, r), 1.0 / 6.0);
    p = MMK_ADD(MMK_MUL(p, r), 0.5);
    p = MMK_ADD(MMK_MUL(p, r), 1.0);
    p = MMK_ADD(MMK_MUL(p, r), 1.0);
    /* scale by 2^k (k in [-1022, 1023] given the input clamp) */
    return MMK_MUL(p, mmk_u2d((uint64_t)(k + 1023) << 52));
}

/* log1p / expm1 by Kahan's compensation on top of mmk_log / mmk_exp. */
MMK_HD double mmk_log1p(double x) {
    double u = MMK_ADD(1.0, x);
    if (u == 1.0) return x;
    if (u == 0.0) return -MMK_INF;
    return MMK_MUL(mmk_log(u), MMK_DIV(x, MMK_SUB(u, 1.0)));
}
MMK_HD double mmk_expm1(double x) {
    double u = mmk_exp(x);
    if (u == 1.0) return x;
    double um1 = MMK_SUB(u, 1.0);
    if (um1 == -1.0) return -1.0;
    if (u == MMK_INF) return u;
    return MMK_MUL(um1, MMK_DIV(x, mmk_log(u)));
}
/* log(1 - exp(x)) for x <= 0 (LogExpFunctions.log1mexp split at -ln 2) */
MMK_HD double mmk_log1mexp(double x) {
    if (x == 0.0) return -MMK_INF;
    if (x < -0.6931471805599453) return mmk_log1p(-mmk_exp(x));
    return mmk_log(-mmk_expm1(x));
}
/* log(|exp(a) - exp(b)|) */
MMK_HD double mmk_logsubexp(double a, double b) {
    double mx = a > b ? a : b;
    double d = a > b ? MMK_SUB(b, a) : MMK_SUB(a, b); /* -|a-b| */
    if (a == b) return -MMK_INF;
    if (mx == -MMK_INF) return -MMK_INF;
    return MMK_ADD(mx, mmk_log1mexp(d));
}
/* log(exp(a) + exp(b)) */
MMK_HD double mmk_logaddexp(double a, double b) {
    double mx = a > b ? a : b;
    double mn = a > b ? b : a;
    if (mn == -MMK_INF) return mx;
    return MMK_ADD(mx, mmk_log1p(mmk_exp(MMK_SUB(mn, mx))));
}

/* ---- Philox4x32-10 --------------------------------------------------------------------- */
#define MMK_PHILOX_M0 0xD2511F53u
#define MMK_PHILOX_M1 0xCD9E8D57u
#define MMK_PHILOX_W0 0x9E3779B9u
#define MMK_PHILOX_W1 0xBB67AE85u

MMK_HD void mmk_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = mmk_mulhi32(MMK_PHILOX_M0, c0), lo0 = MMK_PHILOX_M0 * c0;
        uint32_t hi1 = mmk_mulhi32(MMK_PHILOX_M1, c2), lo1 = MMK_PHILOX_M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0;
        c1 = lo1;
        c2 = n2;
        c3 = lo0;
        k0 += MMK_PHILOX_W0;
        k1 += MMK_PHILOX_W1;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}

/* ---- stream addressing ----------------------------------------------------------------
 * key     = (seed low 32, seed high 32) — always the user's seed, except in the redraw rounds of the
 *           insert-length vector (below); nothing is ever XORed into it, so runs with different seeds
 *           share no stream
 * counter = (c0 + (block << sh), purpose | sample<<8 | g[39:32]<<24, g[31:0], m)
 * One Philox block yields two 64-bit draws: draw 2b is words (0,1), draw 2b+1 is words (2,3).
 * Ordinary streams have c0 = 0, sh = 0: the block index is the whole first counter word.
 *
 * Insert streams (purpose MMK_P_INSERT) — one per (redraw round r, insert i) of a molecule, so that any
 * thread can make the draws of any read pair:
 *     c0 = i (inserts per molecule < 2^24: each is >= 200 bases of a contig < 2^31), sh = 24
 *     key = seed in round 0; in round r >= 1 the round key mmk_round_key(seed, r): words (0,1) of
 *           Philox(key = seed; counter = (r, MMK_P_ROUNDKEY, 0, 0))
 *     draw 0      first variate of the insert-length sampler
 *     draw 1      (round 0 only) the 64 bits of the insert's gap draw (Breakpoints.jl:28): length and gap
 *                 of a pair come from ONE Philox block
 *     draw 2, ... further variates of the length sampler (rejected proposals)
 * (after 256 blocks the block index wraps onto the insert index bits: unreachable, the samplers accept
 * with probability > 0.3 per proposal.)
 */
enum mmk_purpose {
    MMK_P_MASK = 1,     /* N masking: g = contig index, m = haplotype file index, draw = position */
    MMK_P_NMOL = 2,     /* group: molecules-per-barcode draw(s)                   Common.jl:89-91 */
    MMK_P_TARGET = 3,   /* group: draw j = target pick of slot j                  Main.jl:47      */
    MMK_P_MOLSIZE = 4,  /* molecule: molecule length draw(s)                      Breakpoints.jl:155-158 */
    MMK_P_SINGLETON = 5,/* molecule: singleton test                               Breakpoints.jl:173 */
    MMK_P_NFRAGS = 6,   /* molecule: number of inserts draw(s)                    Breakpoints.jl:191-192 */
    MMK_P_INSERT = 7,   /* insert: length sampler variates + gap draw             Breakpoints.jl:207,211 */
    MMK_P_START = 8,    /* molecule: start position                               Breakpoints.jl:240 */
    MMK_P_GAP = 9,      /* (tag of the replay stream only) gap before insert i    Breakpoints.jl:28 */
    MMK_P_ROUNDKEY = 10,/* key derivation of redraw round r >= 1 */
    MMK_P_ERROR = 11    /* read cycle: event / replacement / quality draw of the sequencing-error model */
};

typedef struct mmk_stream {
    uint32_t key[2];
    uint32_t c0, sh;
    uint32_t c1, c2, c3;
    uint32_t next; /* index of the next 64-bit draw */
    uint32_t blk;  /* Philox block currently held in buf (0xffffffff = none) */
    uint32_t buf[4];
} mmk_stream;

MMK_HD void mmk_stream_init(mmk_stream *s, uint64_t seed, uint32_t sample, uint32_t purpose,
                            uint64_t g, uint32_t m) {
    s->key[0] = (uint32_t)seed;
    s->key[1] = (uint32_t)(seed >> 32);
    s->c0 = 0;
    s->sh = 0;
    s->c1 = (purpose & 0xffu) | ((sample & 0xffffu) << 8) | ((uint32_t)((g >> 32) & 0xffu) << 24);
    s->c2 = (uint32_t)g;
    s->c3 = m;
    s->next = 0;
    s->blk = 0xffffffffu;
}
MMK_HD void mmk_stream_seek(mmk_stream *s, uint32_t draw) { s->next = draw; }
MMK_HD void mmk_stream_block(const mmk_stream *s, uint32_t block, uint32_t out[4]) {
    uint32_t ctr[4] = {s->c0 + (block << s->sh), s->c1, s->c2, s->c3};
    mmk_philox4x32_10(ctr, s->key, out);
}
MMK_HD uint64_t mmk_stream_at(const mmk_stream *s, uint32_t draw) {
    uint32_t o[4];
    mmk_stream_block(s, draw >> 1, o);
    return (draw & 1u) ? (((uint64_t)o[3] << 32) | o[2]) : (((uint64_t)o[1] << 32) | o[0]);
}
MMK_HD uint64_t mmk_stream_next64(mmk_stream *s) {
    uint32_t d = s->next++;
    if ((d >> 1) != s->blk) {
        mmk_stream_block(s, d >> 1, s->buf);
        s->blk = d >> 1;
    }
    return (d & 1u) ? (((uint64_t)s->buf[3] << 32) | s->buf[2])
                    : (((uint64_t)s->buf[1] << 32) | s->buf[0]);
}

/* key of redraw round r (r >= 1) of the insert-length vector; round 0 uses the seed itself */
MMK_HD uint64_t mmk_round_key(uint64_t seed, uint32_t round) {
    if (round == 0) return seed;
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t ctr[4] = {round, MMK_P_ROUNDKEY, 0u, 0u};
    uint32_t o[4];
    mmk_philox4x32_10(ctr, key, o);
    return ((uint64_t)o[1] << 32) | o[0];
}
/* stream of insert i of molecule (g, m) under `key` = mmk_round_key(seed, round) */
MMK_HD void mmk_insert_stream_init(mmk_stream *s, uint64_t key, uint32_t sample, uint64_t g, uint32_t m, uint32_t i) {
    mmk_stream_init(s, key, sample, MMK_P_INSERT, g, m);
    s->c0 = i;
    s->sh = 24;
}
/* next variate of the length sampler: draws 0, 2, 3, 4, ... (draw 1 is the gap draw) */
MMK_HD uint64_t mmk_insert_next64(mmk_stream *s) {
    uint64_t v = mmk_stream_next64(s);
    if (s->next == 1) s->next = 2;
    return v;
}
#define MMK_INSERT_GAP_DRAW 1u

/* Sequencing-error model (optional, default off): one 32-bit draw per (read pair, read, cycle), four cycles per block.
 *   key = seed; counter = (cycle >> 2 | (read - 1) << 24, MMK_P_ERROR | sample << 8, pair[31:0], pair[63:32]);
 *   cycle c takes word c & 3 of the block.  `pair` is the run-wide index of the read pair (it fixes barcode, molecule
 *   and insert), `read` 1 or 2, cycle < 2^26.
 * Bits 0-23: the event uniform (thresholds are probabilities scaled by 2^24); bits 24-25: replacement selector;
 * bits 26-28: quality jitter. */
MMK_HD void mmk_error_block(uint64_t seed, uint32_t sample, uint64_t pair, uint32_t read, uint32_t block, uint32_t o[4]) {
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t ctr[4] = {block | ((read - 1u) << 24), (uint32_t)MMK_P_ERROR | ((sample & 0xffffu) << 8), (uint32_t)pair, (uint32_t)(pair >> 32)};
    mmk_philox4x32_10(ctr, key, o);
}
MMK_HD uint32_t mmk_error_draw(uint64_t seed, uint32_t sample, uint64_t pair, uint32_t read, uint32_t cycle) {
    uint32_t o[4];
    mmk_error_block(seed, sample, pair, read, cycle >> 2, o);
    return o[cycle & 3u];
}
#define MMK_ERR_U(w) ((w) & 0xffffffu)
#define MMK_ERR_SEL(w) (((w) >> 24) & 3u)
#define MMK_ERR_JIT(w) (((w) >> 26) & 7u)
#define MMK_ERR_SCALE 16777216.0 /* 2^24 */

/* ---- variates ---------------------------------------------------------------------------- */
/* u01 in [0,1): 53 random bits */
MMK_HD double mmk_bits_to_u01(uint64_t x) { return MMK_MUL((double)(x >> 11), 1.1102230246251565e-16); }
/* standard exponential by inversion: -log(1-u), 1-u is exact and in (0,1] */
MMK_HD double mmk_bits_to_exp(uint64_t x) {
    double v = -mmk_log(MMK_SUB(1.0, mmk_bits_to_u01(x)));
    return v == 0.0 ? 0.0 : v; /* canonical +0 */
}
/* integer in [0, n): multiply-high (bias <= n / 2^64) */
MMK_HD uint64_t mmk_bits_to_index(uint64_t x, uint64_t n) { return mmk_mulhi64(x, n); }

MMK_HD double mmk_next_u01(mmk_stream *s) { return mmk_bits_to_u01(mmk_stream_next64(s)); }
MMK_HD double mmk_next_exp(mmk_stream *s) { return mmk_bits_to_exp(mmk_stream_next64(s)); }
MMK_HD uint64_t mmk_next_index(mmk_stream *s, uint64_t n) {
    return mmk_bits_to_index(mmk_stream_next64(s), n);
}
/* standard normal by inversion: z = Phi^-1(p), p = (k + 1/2) 2^-52 from the top 52 bits (0 < p < 1,
 * both p and 1 - p exact).  Phi^-1 is P. J. Acklam's rational approximation (relative error < 1.15e-9,
 * |Phi(z) - p| < 3e-10 — checked against scipy in tests), evaluated with single-rounded operations so the
 * host and the device agree bit for bit.  One draw per normal, no rejection; 95 % of the draws need no
 * transcendental at all. */
MMK_HD double mmk_bits_to_normal(uint64_t x) {
    const double p = MMK_MUL(MMK_ADD((double)(x >> 12), 0.5), 2.220446049250313e-16); /* 2^-52 */
    const double q0 = MMK_SUB(p, 0.5);
    if (q0 >= -0.47575 && q0 <= 0.47575) {
        const double r = MMK_MUL(q0, q0);
        double n = -3.969683028665376e+01;
        n = MMK_ADD(MMK_MUL(n, r), 2.209460984245205e+02);
        n = MMK_ADD(MMK_MUL(n, r), -2.759285104469687e+02);
        n = MMK_ADD(MMK_MUL(n, r), 1.383577518672690e+02);
        n = MMK_ADD(MMK_MUL(n, r), -3.066479806614716e+01);
        n = MMK_ADD(MMK_MUL(n, r), 2.506628277459239e+00);
        double d = -5.447609879822406e+01;
        d = MMK_ADD(MMK_MUL(d, r), 1.615858368580409e+02);
        d = MMK_ADD(MMK_MUL(d, r), -1.556989798598866e+02);
        d = MMK_ADD(MMK_MUL(d, r), 6.680131188771972e+01);
        d = MMK_ADD(MMK_MUL(d, r), -1.328068155288572e+01);
        d = MMK_ADD(MMK_MUL(d, r), 1.0);
        return MMK_DIV(MMK_MUL(n, q0), d);
    }
    const int upper = q0 > 0.0;
    const double t = upper ? MMK_SUB(1.0, p) : p;
    const double q = MMK_SQRT(MMK_MUL(-2.0, mmk_log(t)));
    double n = -7.784894002430293e-03;
    n = MMK_ADD(MMK_MUL(n, q), -3.223964580411365e-01);
    n = MMK_ADD(MMK_MUL(n, q), -2.400758277161838e+00);
    n = MMK_ADD(MMK_MUL(n, q), -2.549732539343734e+00);
    n = MMK_ADD(MMK_MUL(n, q), 4.374664141464968e+00);
    n = MMK_ADD(MMK_MUL(n, q), 2.938163982698783e+00);
    double d = 7.784695709041462e-03;
    d = MMK_ADD(MMK_MUL(d, q), 3.224671290700398e-01);
    d = MMK_ADD(MMK_MUL(d, q), 2.445134137142996e+00);
    d = MMK_ADD(MMK_MUL(d, q), 3.754408661907416e+00);
    d = MMK_ADD(MMK_MUL(d, q), 1.0);
    const double z = MMK_DIV(n, d);
    return upper ? -z : z;
}
MMK_HD double mmk_next_normal(mmk_stream *s) { return mmk_bits_to_normal(mmk_stream_next64(s)); }

#endif /* MMK_RNG_H */
