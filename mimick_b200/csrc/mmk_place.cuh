// mmk_place.cuh — read-pair placement in one pass over the shard.
//
// Reference behaviour: calculate_insert_sizes (Breakpoints.jl:205-218), get_sequences (:227-251),
// seq_breakpoints! (:21-35), and the record lengths of format_R1 / format_R2 (FormatFASTQ.jl).
//
// k_place_block — one CTA places a block of TB tiles (32 read pairs each):
//   A. one thread per read pair of every molecule that touches the block: ONE Philox block gives the insert
//      length (through a threshold table: no floating point) and the 64 bits of the gap draw; both stay in
//      shared memory; the integer length sum of every molecule is accumulated with shared-memory atomics;
//   B. one thread per molecule: redraw rule on the sum, start draw, the serial gap chain over shared memory;
//   C. one warp per tile: record lengths, intra-tile scan, the 16-byte pair descriptors the emit kernel reads,
//      tile totals and the fast/general flag.
// Molecules that straddle a block boundary are placed by both blocks (identical results; the block holding the
// molecule's first pair writes its record).  Molecules with more than BIG pairs would not fit the shared-memory
// window: k_place_big places them first, one CTA each, and the blocks read their inserts back from global memory.
//
// The insert-length table.  trunc(mu + sd * Phi^-1(p(u))) is a non-decreasing step function of the 52-bit uniform
// u up to rounding noise, so "which integer" is a search among ~16 sd thresholds.  The table keeps the top 32 bits of
// each threshold; draws within 2^24 units of a threshold (or outside the tabulated range, or rejected by the
// truncation at 200) take the exact double-precision path — 1e-5 of the draws — so the result is bit-identical
// to the definition in mmk_rng.h whatever the rounding noise does near a step (the noise, and the 2e-9 mismatch of the two
// branches of the approximation where they meet, are far smaller than the guard band).
#pragma once
#include "mmk_internal.h"
#include "mmk_sampling.cuh"

namespace mmkdev {

constexpr uint32_t INS_GUARD = 16;  // buckets of 2^20 units either side of a threshold that take the exact path

__device__ __forceinline__ bool ins_len_fast(const InsTable &t, uint32_t u32, uint32_t &len) {
    const uint32_t c = u32 >> (32 - INS_CELL_BITS);
    uint32_t lo = t.cell[c], hi = t.cell[c + 1];
    while (lo < hi) {  // #thresholds <= u32; the cell bounds leave 0-2 candidates except in the far tails
        const uint32_t mid = (lo + hi) >> 1;
        if (t.thr[mid] <= u32) lo = mid + 1; else hi = mid;
    }
    len = t.base_len + lo;
    const bool below_ok = lo == 0 || u32 - t.thr[lo - 1] > INS_GUARD;
    const bool above_ok = lo == t.n_thr || t.thr[lo] - u32 > INS_GUARD;
    return below_ok && above_ok && u32 >= t.u_lo && u32 <= t.u_hi;
}

// length and gap bits of insert i of molecule (g, m), round 0: one Philox block
__device__ __forceinline__ uint32_t pair_draw(const SimConst &sc, const InsTable &tab, uint64_t g, uint32_t m, uint32_t i,
                                              uint32_t &bits_lo, uint32_t &bits_hi) {
    mmk_stream is;
    mmk_insert_stream_init(&is, sc.seed, sc.sample, g, m, i);
    mmk_stream_block(&is, 0, is.buf);
    is.blk = 0;
    bits_lo = is.buf[2];
    bits_hi = is.buf[3];
    uint32_t len;
    if (tab.enabled && ins_len_fast(tab, is.buf[1], len)) return len;
    NoRec rec;
    return insert_len_of(draw_insert_from(sc, is, rec));
}
// length only, any round (`key` = mmk_round_key(seed, round))
__device__ __forceinline__ uint32_t pair_len(const SimConst &sc, const InsTable &tab, uint64_t key, uint64_t g, uint32_t m, uint32_t i) {
    mmk_stream is;
    mmk_insert_stream_init(&is, key, sc.sample, g, m, i);
    mmk_stream_block(&is, 0, is.buf);
    is.blk = 0;
    uint32_t len;
    if (tab.enabled && ins_len_fast(tab, is.buf[1], len)) return len;
    NoRec rec;
    return insert_len_of(draw_insert_from(sc, is, rec));
}

// Does the float sum of the molecule's insert lengths exceed the molecule (Breakpoints.jl:209)?  The sum lies in
// [isum, isum + n): decided from the integers unless they straddle the molecule size (then: the left-to-right sum).
__device__ __noinline__ bool float_sum_over(const SimConst &sc, uint64_t key, uint64_t g, uint32_t m, uint32_t nfrags, uint32_t molsize) {
    NoRec rec;
    double sum = 0.0;
    for (uint32_t k = 0; k < nfrags; ++k) sum = MMK_ADD(sum, draw_insert(sc, key, g, m, k, rec));
    return sum > (double)molsize;
}
__device__ __forceinline__ bool sum_over(const SimConst &sc, uint64_t key, uint64_t g, uint32_t m, uint32_t nfrags, uint32_t molsize,
                                         uint64_t isum) {
    if (isum >= (uint64_t)molsize + 2u) return true;
    if (isum + nfrags + 1u <= (uint64_t)molsize) return false;
    return float_sum_over(sc, key, g, m, nfrags, molsize);
}

// lengths of the header pieces every record of a molecule shares: "@chrom:hap|start:end|" and the format's tail
__device__ __forceinline__ uint32_t barcode_len(int bc_format, uint64_t k) {
    switch (bc_format) {
    case MMK_FMT_HAPLOTAGGING: return 12u;
    case MMK_FMT_STLFR:
        return ndigits_u32((uint32_t)(k % 2000) + 1u) + ndigits_u32((uint32_t)(k / 2000 % 2000) + 1u) + ndigits_u32((uint32_t)(k / 4000000) + 1u) + 2u;
    case MMK_FMT_TELLSEQ: return 18u;
    default: return 16u;
    }
}
__device__ __forceinline__ void header_lens(const SimConst &sc, const SchemaDev &S, int64_t start, uint32_t molsize, uint64_t gi,
                                            uint32_t &P, uint32_t &SUF) {
    P = 1u + S.name_len + 1u + ndigits_u32(S.haplotype) + 1u + ndigits_u32((uint32_t)start) + 1u +
        ndigits_u32((uint32_t)(start + molsize)) + 1u;
    const uint32_t bl = barcode_len(sc.bc_format, gi);
    switch (sc.fq_format) {
    case MMK_FMT_HAPLOTAGGING: SUF = 8u + bl + 1u; break;
    case MMK_FMT_STANDARD: SUF = 15u + bl + 1u; break;
    default: SUF = 1u + bl + 14u; break;  // stlfr '#', tellseq ':', tenx '|', then " 1:N:0:ATGACA\n"
    }
}

// quality-string lengths: those of the molecule's FIRST reads (FormatFASTQ.jl:14)
__device__ __forceinline__ void first_read_quals(const SimConst &sc, int64_t start, uint32_t st0, uint32_t len0, uint32_t n, uint32_t &q1,
                                                 uint32_t &q2) {
    const int64_t bs = start + st0, be = bs + len0 - 1;
    const int64_t a2 = be - sc.r2 + 1;
    q1 = window_len(bs, bs + sc.r1 - 1, n, sc.strict);
    q2 = (a2 >= 1 || sc.strict) ? window_len(a2, be, n, sc.strict) : 0u;
    if (sc.fq_format == MMK_FMT_TENX) q1 += 16u;
}

struct PlaceArgs {
    SimConst sc;
    const SchemaDev *sch;
    PlanDev plan;
    uint64_t mol0;              // first molecule of the shard in the plan
    uint32_t n_mol;             // molecules in the shard
    uint32_t n_tiles;
    uint64_t n_pairs;           // pairs in the shard
    const uint64_t *roff;       // [n_mol + 1] shard-relative first pair of every molecule
    const uint32_t *tile_mol0;  // [n_tiles] molecule holding the first pair of every tile
    const InsTable *tab;
    const uint32_t *big_list;   // shard molecules with more than `big` pairs
    const uint32_t *big_count;
    uint2 *ins;
    MolRec *mol;
    uint4 *fdesc;
    uint32_t *sum1, *sum2;
    uint8_t *tile_flags;
    unsigned long long *gather;  // [256]
    unsigned long long *err;     // [0]: first molecule (plan index) with a non-zero status
    uint32_t stage_cap;
    int32_t bpb;
    uint32_t n_blocks;
};

constexpr uint32_t PLB_THREADS = 256;
template <int TB>
struct PlbGeom {
    static constexpr uint32_t PB = TB * 32u;                   // pairs per block
    static constexpr uint32_t BIG = PB / 4u;                   // larger molecules go through k_place_big
    static constexpr uint32_t SPT = ((PB + 2u * BIG + PLB_THREADS - 1u) / PLB_THREADS) | 1u;  // slots per thread, odd: conflict-free
    static constexpr uint32_t SLOTS = SPT * PLB_THREADS;
    static constexpr uint32_t MAXM = PB + 2u;                  // molecules touching a block (each has >= 1 pair)
    static constexpr size_t SMEM = sizeof(InsTable) + (size_t)SLOTS * 14u + (size_t)(MAXM + 2u) * 12u + 64u;
};

// packed per-molecule word of phases B -> C
#define MPK(schema, P, SUF, qreg, big) ((schema) | ((P) << 16) | ((SUF) << 24) | ((qreg) << 30) | ((big) << 31))
#define MPK_SCHEMA(w) ((w) & 0xffffu)
#define MPK_P(w) (((w) >> 16) & 0xffu)
#define MPK_SUF(w) (((w) >> 24) & 0x3fu)
#define MPK_QREG(w) (((w) >> 30) & 1u)
#define MPK_BIG(w) ((w) >> 31)

__device__ __forceinline__ void stage_table(InsTable *dst, const InsTable *src) {
    const uint4 *s = reinterpret_cast<const uint4 *>(src);
    uint4 *d = reinterpret_cast<uint4 *>(dst);
    for (uint32_t i = threadIdx.x; i < sizeof(InsTable) / 16; i += blockDim.x) d[i] = s[i];
}

// One molecule (whose draws of round 0 are in the shared-memory slots s_len / s_lo / s_hi at [s0, s0 + nfrags)):
// redraw rule, start, chain.  Leaves the insert offsets in s_lo.  Returns the status.
__device__ __forceinline__ uint32_t chain_in_smem(const SimConst &sc, const InsTable &tab, uint64_t g, uint32_t m, uint32_t n_hap, uint32_t molsize,
                                                  uint32_t nfrags, uint64_t isum, uint32_t *s_lo, const uint32_t *s_hi, uint32_t *s_len,
                                                  int64_t &start) {
    bool over = sum_over(sc, sc.seed, g, m, nfrags, molsize, isum);
    for (int attempts = 1; over && attempts <= sc.attempts; ++attempts) {  // whole-vector redraws (Breakpoints.jl:209-212): rare
        const uint64_t key = mmk_round_key(sc.seed, (uint32_t)attempts);
        isum = 0;
        for (uint32_t k = 0; k < nfrags; ++k) {
            const uint32_t li = pair_len(sc, tab, key, g, m, k);
            s_len[k] = li;
            isum += li;
        }
        over = sum_over(sc, key, g, m, nfrags, molsize, isum);
    }
    // start_pos = rand(1:adjusted_end)
    const uint64_t adjusted_end = sc.circular ? (uint64_t)n_hap : (n_hap > molsize + 1u ? (uint64_t)(n_hap - molsize) : 1ull);
    mmk_stream s;
    mmk_stream_init(&s, sc.seed, sc.sample, MMK_P_START, g, m);
    start = (int64_t)mmk_next_index(&s, adjusted_end) + 1;
    if (over) return MMK_ERR_INSERTS_UNFITTABLE;
    uint32_t status = 0;
    uint64_t remaining = (uint64_t)molsize + 1 - isum;
    uint32_t cursor = 0;
    // R2 window bstop - r2 + 1 : bstop must start at >= 1 (Julia BoundsError otherwise): st + len >= r2 + 1 - start
    const uint32_t need_end = (!sc.strict && start <= (int64_t)sc.r2) ? (uint32_t)((int64_t)sc.r2 + 1 - start) : 0u;
    uint32_t lo = s_lo[0], hi = s_hi[0], len = s_len[0];
    for (uint32_t k = 0; k < nfrags; ++k) {
        const uint32_t kn = k + 1 < nfrags ? k + 1 : k;  // next step's operands do not depend on the chain
        const uint32_t nlo = s_lo[kn], nhi = s_hi[kn], nlen = s_len[kn];
        const uint32_t gap = (uint32_t)mmk_bits_to_index(((uint64_t)hi << 32) | lo, remaining + 1);
        const uint32_t st = cursor + gap;
        s_lo[k] = st;
        cursor = st + len;
        remaining -= gap;
        if (st + len < need_end) status = MMK_ERR_BOUNDS;
        lo = nlo;
        hi = nhi;
        len = nlen;
    }
    return status;
}

template <int TB>
__global__ void __launch_bounds__(PLB_THREADS) k_place_block(const __grid_constant__ PlaceArgs a) {
    using G = PlbGeom<TB>;
    extern __shared__ __align__(16) uint8_t plb_smem[];
    InsTable *tab = reinterpret_cast<InsTable *>(plb_smem);
    uint32_t *s_lo = reinterpret_cast<uint32_t *>(plb_smem + sizeof(InsTable));
    uint32_t *s_hi = s_lo + G::SLOTS;
    uint32_t *s_len = s_hi + G::SLOTS;
    int32_t *s_mroff = reinterpret_cast<int32_t *>(s_len + G::SLOTS);  // [MAXM + 2] first slot of every molecule (may be < 0)
    uint32_t *s_msum = reinterpret_cast<uint32_t *>(s_mroff + G::MAXM + 2);  // length sum, later the molecule start
    uint32_t *s_mpk = s_msum + G::MAXM + 2;
    uint16_t *s_mid = reinterpret_cast<uint16_t *>(s_mpk + G::MAXM + 2);     // [SLOTS] molecule (block-local) of every slot
    __shared__ uint32_t s_wmax[PLB_THREADS / 32];
    const SimConst &sc = a.sc;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    stage_table(tab, a.tab);
    for (uint32_t b = blockIdx.x; b < a.n_blocks; b += gridDim.x) {
        __syncthreads();  // previous block's phase C is done with the shared arrays (and the table is staged)
        const uint32_t t_lo = b * TB, t_hi = min(t_lo + TB, a.n_tiles);
        const uint64_t P0 = (uint64_t)t_lo * 32u, P1 = min((uint64_t)t_hi * 32u, a.n_pairs);
        const uint32_t m_lo = a.tile_mol0[t_lo];
        const uint32_t m_nx = t_hi < a.n_tiles ? a.tile_mol0[t_hi] : a.n_mol - 1u;
        uint32_t NM = m_nx - m_lo + 1u;
        const uint64_t R0 = a.roff[m_lo], R1 = a.roff[m_lo + 1u];
        const bool first_big = R1 - R0 > G::BIG;
        const uint64_t W0 = first_big ? R1 : R0;  // first pair of the shared-memory window
        for (uint32_t j = tid; j <= NM; j += PLB_THREADS) s_mroff[j] = (int32_t)(int64_t)(a.roff[m_lo + j] - W0);
        __syncthreads();
        if ((int64_t)s_mroff[NM - 1u] >= (int64_t)(P1 - W0)) --NM;  // the molecule holding pair P1 starts there: not ours
        const bool last_big = (uint32_t)(s_mroff[NM] - s_mroff[NM - 1u]) > G::BIG;
        const int32_t w1 = last_big ? s_mroff[NM - 1u] : s_mroff[NM];
        const uint32_t nslots = w1 > 0 ? (uint32_t)w1 : 0u;  // <= PB + 2 BIG <= SLOTS
        for (uint32_t s = tid; s < nslots; s += PLB_THREADS) s_mid[s] = 0;
        for (uint32_t j = tid; j < NM; j += PLB_THREADS) s_msum[j] = 0;
        __syncthreads();
        for (uint32_t j = tid; j < NM; j += PLB_THREADS) {  // head flags: the molecule index at its first slot
            const int32_t s0 = s_mroff[j];
            if (s0 >= 0 && (uint32_t)s0 < nslots) s_mid[s0] = (uint16_t)j;
        }
        __syncthreads();
        // ---- phase A: max-scan of the head flags gives every slot its molecule; then the draws -----------------
        const uint32_t sbase = tid * G::SPT;
        uint32_t mid[G::SPT];
        {
            uint32_t run = 0;
#pragma unroll
            for (uint32_t k = 0; k < G::SPT; ++k) {
                const uint32_t s = sbase + k;
                const uint32_t v = s < nslots ? s_mid[s] : 0u;
                run = max(run, v);
                mid[k] = run;
            }
            uint32_t inc = run;
            for (int o = 1; o < 32; o <<= 1) inc = max(inc, __shfl_up_sync(0xffffffffu, inc, o));
            if (lane == 31) s_wmax[warp] = inc;
            uint32_t prev = __shfl_up_sync(0xffffffffu, inc, 1);
            if (lane == 0) prev = 0;
            __syncthreads();
            for (uint32_t w = 0; w < warp; ++w) prev = max(prev, s_wmax[w]);
#pragma unroll
            for (uint32_t k = 0; k < G::SPT; ++k) mid[k] = max(mid[k], prev);
        }
        {
            uint32_t jprev = 0xffffffffu, acc = 0, ms = 0;
            uint64_t g = 0;
            bool jbig = false;
#pragma unroll
            for (uint32_t k = 0; k < G::SPT; ++k) {
                const uint32_t s = sbase + k;
                if (s >= nslots) break;
                const uint32_t j = mid[k];
                if (j != jprev) {
                    if (acc) atomicAdd(&s_msum[jprev], acc);
                    acc = 0;
                    jprev = j;
                    jbig = (uint32_t)(s_mroff[j + 1u] - s_mroff[j]) > G::BIG;
                    g = a.plan.group[a.mol0 + m_lo + j];
                    ms = a.plan.slot[a.mol0 + m_lo + j];
                }
                s_mid[s] = (uint16_t)j;
                if (jbig) continue;  // placed by k_place_big
                uint32_t blo, bhi;
                const uint32_t li = pair_draw(sc, *tab, g, ms, (uint32_t)((int32_t)s - s_mroff[j]), blo, bhi);
                s_lo[s] = blo;
                s_hi[s] = bhi;
                s_len[s] = li;
                acc += li;
            }
            if (acc) atomicAdd(&s_msum[jprev], acc);
        }
        __syncthreads();
        // ---- phase B: one thread per molecule --------------------------------------------------------------------
        for (uint32_t j = tid; j < NM; j += PLB_THREADS) {
            const uint32_t ml = m_lo + j;  // shard-relative molecule
            const uint64_t gi = a.mol0 + ml;
            const uint32_t nfrags = (uint32_t)(s_mroff[j + 1u] - s_mroff[j]);
            if (nfrags > G::BIG) {  // already placed: fetch what phase C needs
                const MolRec r = a.mol[ml];
                const uint32_t inl = sc.fq_format == MMK_FMT_TENX ? 16u : 0u;
                const uint32_t qreg = r.qual1 == (uint32_t)sc.r1 + inl && r.qual2 == (uint32_t)sc.r2;
                s_msum[j] = (uint32_t)r.start;
                s_mpk[j] = MPK((uint32_t)r.schema, (uint32_t)r.prefix_len, (uint32_t)r.suffix_len, qreg, 1u);
                continue;
            }
            const uint64_t g = a.plan.group[gi];
            const uint32_t ms = a.plan.slot[gi];
            const uint32_t t = a.plan.schema[gi];
            const uint32_t molsize = a.plan.size[gi];
            const SchemaDev &S = a.sch[t];
            const uint32_t n = S.len;
            const int32_t s0 = s_mroff[j];
            int64_t start;
            const uint32_t status = chain_in_smem(sc, *tab, g, ms, n, molsize, nfrags, (uint64_t)s_msum[j], s_lo + s0, s_hi + s0, s_len + s0, start);
            uint32_t q1 = 0, q2 = 0, P, SUF;
            if (!status) first_read_quals(sc, start, s_lo[s0], s_len[s0], n, q1, q2);
            else if (sc.fq_format == MMK_FMT_TENX) q1 = 16u;
            header_lens(sc, S, start, molsize, gi, P, SUF);
            const uint32_t inl = sc.fq_format == MMK_FMT_TENX ? 16u : 0u;
            const uint32_t qreg = q1 == (uint32_t)sc.r1 + inl && q2 == (uint32_t)sc.r2;
            s_msum[j] = (uint32_t)start;
            s_mpk[j] = MPK(t, P, SUF, qreg, 0u);
            if ((int64_t)s0 + (int64_t)W0 >= (int64_t)P0) {  // this block holds the molecule's first pair: it owns the record
                if (status) {
                    atomicMin(&a.err[0], (unsigned long long)gi);
                    a.plan.status[gi] = (uint8_t)status;
                }
                MolRec r;
                r.start = start;
                r.size = molsize;
                r.schema = (uint16_t)t;
                r.prefix_len = (uint8_t)P;
                r.suffix_len = (uint8_t)SUF;
                r.qual1 = q1;
                r.qual2 = q2;
                a.mol[ml] = r;
            }
        }
        __syncthreads();
        // ---- phase C: one warp per tile ------------------------------------------------------------------------------
        unsigned long long gsum = 0;
        for (uint32_t tile = t_lo + warp; tile < t_hi; tile += PLB_THREADS / 32) {
            const uint64_t p = (uint64_t)tile * 32u + lane;
            const bool valid = p < P1;
            uint32_t len1 = 0, len2 = 0, sq1 = 0, sq2 = 0, regular = 1;
            uint4 fd = make_uint4(0, 0, 0, 0);
            uint2 in = make_uint2(0, 0);
            bool write_ins = false;
            if (valid) {
                const int64_t s = (int64_t)(p - W0);  // p >= W0 - BIG ... may be negative (first molecule big)
                uint32_t j;
                if (s < 0) j = 0;
                else if (s >= (int64_t)nslots) j = NM - 1u;
                else j = s_mid[s];
                const uint32_t pk = s_mpk[j];
                const uint32_t ml = m_lo + j;
                const uint32_t dm = ml - a.tile_mol0[tile];
                const int64_t start = (int64_t)s_msum[j];
                uint32_t st, li;
                if (MPK_BIG(pk)) {
                    const uint2 q = a.ins[p];
                    st = q.x;
                    li = MMK_INS_LEN(q.y);
                } else {
                    st = s_lo[s];
                    li = s_len[s];
                    write_ins = true;
                }
                in = make_uint2(st, li | (dm << 24));
                const SchemaDev &S = a.sch[MPK_SCHEMA(pk)];
                const uint32_t n = S.len;
                const uint32_t bstart = (uint32_t)start + st, bend = bstart + li - 1u;
                const uint32_t P = MPK_P(pk), SUF = MPK_SUF(pk);
                const uint32_t d1 = ndigits_u32(bstart), d2 = ndigits_u32(bend);
                const uint32_t hdr = P + d1 + 1u + d2 + SUF;
                const uint32_t inl = sc.fq_format == MMK_FMT_TENX ? 16u : 0u;
                const int64_t a2 = (int64_t)bend - sc.r2 + 1;
                if (bend <= n && bstart + (uint32_t)sc.r1 - 1u <= n) {  // both windows inside the haplotype: nominal lengths
                    sq1 = (uint32_t)sc.r1;
                    sq2 = (uint32_t)sc.r2;
                } else {
                    sq1 = window_len(bstart, (int64_t)bstart + sc.r1 - 1, n, sc.strict);
                    sq2 = window_len(a2, bend, n, sc.strict);
                }
                uint32_t q1 = (uint32_t)sc.r1 + inl, q2 = (uint32_t)sc.r2;
                if (!MPK_QREG(pk)) {
                    if (MPK_BIG(pk)) {
                        q1 = a.mol[ml].qual1;
                        q2 = a.mol[ml].qual2;
                    } else {  // from the molecule's first insert, still in shared memory
                        const int32_t s0 = s_mroff[j];
                        first_read_quals(sc, start, s_lo[s0], s_len[s0], n, q1, q2);
                    }
                }
                len1 = hdr + inl + sq1 + 3u + q1 + 1u;
                len2 = hdr + sq2 + 3u + q2 + 1u;
                regular = sq1 == (uint32_t)sc.r1 && sq2 == (uint32_t)sc.r2 && q1 == (uint32_t)sc.r1 + inl && q2 == (uint32_t)sc.r2 &&
                          (uint64_t)bstart + sc.r1 - 1 <= n && bend <= n && a2 >= 1 && P <= 64u && SUF <= 40u;
                fd = make_uint4(bstart, in.y, 0u, (P & 0x7fu) | (SUF << 7) | (MPK_SCHEMA(pk) << 13));
            }
            uint32_t i1 = len1, i2 = len2, gb = (sq1 * a.bpb + 7) / 8 + (sq2 * a.bpb + 7) / 8;
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y1 = __shfl_up_sync(0xffffffffu, i1, o), y2 = __shfl_up_sync(0xffffffffu, i2, o);
                if (lane >= (uint32_t)o) {
                    i1 += y1;
                    i2 += y2;
                }
            }
            gb = __reduce_add_sync(0xffffffffu, gb);
            gsum += gb;
            const bool irregular = __any_sync(0xffffffffu, !regular);
            const uint32_t tot1 = __shfl_sync(0xffffffffu, i1, 31), tot2 = __shfl_sync(0xffffffffu, i2, 31);
            if (valid) {
                fd.z = ((i1 - len1) & 0xffffu) | ((i2 - len2) << 16);
                a.fdesc[p] = fd;
                if (write_ins) a.ins[p] = in;
            }
            if (lane == 0) {
                a.sum1[tile] = tot1;
                a.sum2[tile] = tot2;
                a.tile_flags[tile] = (uint8_t)(!irregular && tot1 + 48u <= a.stage_cap && tot2 + 48u <= a.stage_cap && tot1 < 65536u && tot2 < 65536u);
            }
        }
        if (lane == 0 && gsum) atomicAdd(&a.gather[(b * 8u + warp) & 255u], gsum);
    }
}

// One CTA per molecule with more than BIG pairs: lengths and gap bits in chunks through shared memory, the chain by
// one thread.  Writes the molecule's inserts and its record.
constexpr uint32_t BIGC = 1024;  // pairs per chunk
__global__ void __launch_bounds__(PLB_THREADS) k_place_big(const __grid_constant__ PlaceArgs a) {
    __shared__ InsTable tab;
    __shared__ uint32_t s_lo[BIGC], s_hi[BIGC], s_len[BIGC];
    __shared__ unsigned long long s_red[PLB_THREADS / 32];
    __shared__ unsigned long long s_isum;
    __shared__ uint32_t s_flag;
    __shared__ unsigned long long s_state[3];  // remaining, cursor, status
    const SimConst &sc = a.sc;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    stage_table(&tab, a.tab);
    __syncthreads();
    const uint32_t count = *a.big_count;
    for (uint32_t bi = blockIdx.x; bi < count; bi += gridDim.x) {
        const uint32_t ml = a.big_list[bi];
        const uint64_t gi = a.mol0 + ml;
        const uint64_t ro = a.roff[ml];
        const uint32_t nfrags = (uint32_t)(a.roff[ml + 1u] - ro);
        const uint64_t g = a.plan.group[gi];
        const uint32_t ms = a.plan.slot[gi];
        const uint32_t t = a.plan.schema[gi];
        const uint32_t molsize = a.plan.size[gi];
        const SchemaDev &S = a.sch[t];
        const uint32_t n = S.len;
        // the round whose length vector fits (Breakpoints.jl:209-212)
        uint64_t key = sc.seed, isum = 0;
        bool over = true;
        int round = 0;
        for (int attempts = 0; attempts <= sc.attempts; ++attempts) {
            round = attempts;
            key = mmk_round_key(sc.seed, (uint32_t)attempts);
            unsigned long long part = 0;
            for (uint32_t k = tid; k < nfrags; k += PLB_THREADS) part += pair_len(sc, tab, key, g, ms, k);
            for (int o = 16; o; o >>= 1) part += __shfl_down_sync(0xffffffffu, part, o);
            __syncthreads();  // s_red / s_flag of the previous round are consumed
            if (lane == 0) s_red[warp] = part;
            __syncthreads();
            if (tid == 0) {
                unsigned long long tot = 0;
                for (uint32_t w = 0; w < PLB_THREADS / 32; ++w) tot += s_red[w];
                s_isum = tot;
                s_flag = sum_over(sc, key, g, ms, nfrags, molsize, tot) ? 1u : 0u;
            }
            __syncthreads();
            isum = s_isum;
            over = s_flag != 0u;
            if (!over) break;
        }
        const uint64_t adjusted_end = sc.circular ? (uint64_t)n : (n > molsize + 1u ? (uint64_t)(n - molsize) : 1ull);
        mmk_stream s;
        mmk_stream_init(&s, sc.seed, sc.sample, MMK_P_START, g, ms);
        const int64_t start = (int64_t)mmk_next_index(&s, adjusted_end) + 1;
        uint32_t status = over ? (uint32_t)MMK_ERR_INSERTS_UNFITTABLE : 0u;
        uint32_t q1 = 0, q2 = 0;
        if (!status) {
            const uint32_t need_end = (!sc.strict && start <= (int64_t)sc.r2) ? (uint32_t)((int64_t)sc.r2 + 1 - start) : 0u;
            if (tid == 0) {
                s_state[0] = (uint64_t)molsize + 1 - isum;
                s_state[1] = 0;
                s_state[2] = 0;
            }
            for (uint32_t c0 = 0; c0 < nfrags; c0 += BIGC) {
                const uint32_t cn = min(BIGC, nfrags - c0);
                __syncthreads();  // previous chunk written out
                for (uint32_t k = tid; k < cn; k += PLB_THREADS) {
                    uint32_t blo, bhi;
                    uint32_t li = pair_draw(sc, tab, g, ms, c0 + k, blo, bhi);
                    if (round) li = pair_len(sc, tab, key, g, ms, c0 + k);  // lengths of the accepted round, gap bits of round 0
                    s_lo[k] = blo;
                    s_hi[k] = bhi;
                    s_len[k] = li;
                }
                __syncthreads();
                if (tid == 0) {
                    uint64_t remaining = s_state[0];
                    uint32_t cursor = (uint32_t)s_state[1];
                    uint32_t st_bad = (uint32_t)s_state[2];
                    for (uint32_t k = 0; k < cn; ++k) {
                        const uint32_t gap = (uint32_t)mmk_bits_to_index(((uint64_t)s_hi[k] << 32) | s_lo[k], remaining + 1);
                        const uint32_t st = cursor + gap;
                        s_lo[k] = st;
                        cursor = st + s_len[k];
                        remaining -= gap;
                        if (st + s_len[k] < need_end) st_bad = MMK_ERR_BOUNDS;
                    }
                    s_state[0] = remaining;
                    s_state[1] = cursor;
                    s_state[2] = st_bad;
                }
                __syncthreads();
                if (c0 == 0) first_read_quals(sc, start, s_lo[0], s_len[0], n, q1, q2);
                for (uint32_t k = tid; k < cn; k += PLB_THREADS) {
                    const uint64_t p = ro + c0 + k;
                    const uint32_t dm = ml - a.tile_mol0[p >> 5];
                    a.ins[p] = make_uint2(s_lo[k], s_len[k] | (dm << 24));
                }
            }
            __syncthreads();
            status = (uint32_t)s_state[2];
            if (status) q1 = q2 = 0;
        }
        if (status && sc.fq_format == MMK_FMT_TENX) q1 = 16u;
        if (tid == 0) {
            if (status) {
                atomicMin(&a.err[0], (unsigned long long)gi);
                a.plan.status[gi] = (uint8_t)status;
            }
            uint32_t P, SUF;
            header_lens(sc, S, start, molsize, gi, P, SUF);
            MolRec r;
            r.start = start;
            r.size = molsize;
            r.schema = (uint16_t)t;
            r.prefix_len = (uint8_t)P;
            r.suffix_len = (uint8_t)SUF;
            r.qual1 = q1;
            r.qual2 = q2;
            a.mol[ml] = r;
        }
        __syncthreads();
    }
}

// first molecule of every tile, and the list of molecules k_place_big takes
__global__ void k_tile_mol0_big(const uint64_t *__restrict__ roff_l, uint32_t n_mol, uint32_t tile_pairs, uint32_t *tile_mol0, uint32_t big,
                                uint32_t *big_list, uint32_t *big_count, uint32_t big_cap) {
    uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_mol) return;
    uint64_t lo = roff_l[m], hi = roff_l[m + 1];  // pairs [lo, hi)
    if (hi == lo) return;
    const uint32_t sh = 31 - __clz(tile_pairs);  // tile_pairs is a power of two
    for (uint64_t t = (lo + tile_pairs - 1) >> sh; (t << sh) < hi; ++t) tile_mol0[t] = m;
    if (big && hi - lo > big) {
        const uint32_t k = atomicAdd(big_count, 1u);
        if (k < big_cap) big_list[k] = m;
    }
}

}  // namespace mmkdev
