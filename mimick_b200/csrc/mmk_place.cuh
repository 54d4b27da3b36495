// mmk_place.cuh — read-pair placement of a shard, block by block, in shared memory.
//
// Reference behaviour: calculate_insert_sizes (Breakpoints.jl:205-218), get_sequences (:227-251),
// seq_breakpoints! (:21-35), and the record lengths of format_R1 / format_R2 (FormatFASTQ.jl).
//
// A block is a run of consecutive molecules of bounded weight (pairs + PB_W per molecule: at most PB_CAP pairs and
// PB_MAXM molecules, found by k_place_blocks with a search over the weight prefix).  One CTA per block (k_place_block):
//   A. one thread per read pair: ONE Philox block gives the insert length (through a threshold table: no floating
//      point) and the 64 bits of the gap draw; both stay in shared memory;
//   B. one thread per molecule: integer length sum, redraw rule (Breakpoints.jl:209-212), start draw (:240), the
//      serial gap chain over shared memory (:27-33), header-piece lengths, the molecule record;
//   C. one warp per 32-pair tile (aligned to the tiles of the emit kernel): breakpoints, record lengths, the 16-byte
//      pair descriptors the emit kernel reads, record bytes per tile and stream, the tile's irregular flag.
// Nothing per pair goes through global scratch, and the only per-pair output is the 16-byte descriptor (the insert plan
// `ins` of the three-kernel path is derived from it on demand: k_ins_from_fdesc).  Molecules with more than PB_BIG pairs do not fit a block's window:
// k_place_big places them first, one CTA each, and phase C reads their inserts back from global memory.
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
// the table path alone: false when the draw needs the exact (double-precision) definition — 1e-5 of the draws
__device__ __forceinline__ bool pair_draw_fast(const SimConst &sc, const InsTable &tab, uint64_t g, uint32_t m, uint32_t i, uint32_t &len,
                                               uint32_t &bits_lo, uint32_t &bits_hi) {
    mmk_stream is;
    mmk_insert_stream_init(&is, sc.seed, sc.sample, g, m, i);
    uint32_t o[4];
    mmk_stream_block(&is, 0, o);
    bits_lo = o[2];
    bits_hi = o[3];
    return tab.enabled && ins_len_fast(tab, o[1], len);
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

constexpr uint32_t PB_THREADS = 256;
constexpr uint32_t PB_W = 8;          // weight of a molecule besides its pairs (bounds the molecules of a block)
constexpr uint32_t PB_BIG = 1024;     // molecules with more pairs go through k_place_big
constexpr uint32_t PB_EXACT = 0xffffu; // phase A: length to be drawn by the exact definition (lengths are < 65000: place_block_supports)
#ifndef MMK_PB_TARGET
#define MMK_PB_TARGET 2496
#endif
constexpr uint32_t PB_TARGET = MMK_PB_TARGET;                 // weight of a block
constexpr uint32_t PB_CAP = PB_TARGET + PB_BIG;               // pairs in shared memory: < TARGET + W + BIG - W * molecules
constexpr uint32_t PB_MAXM = (PB_TARGET + PB_BIG) / PB_W + 2; // molecules of a block: each weighs >= W
constexpr uint32_t PB_SORT_MAX = 160;  // blocks with more molecules run phase B in molecule order
#ifndef MMK_PB_SCAN
#define MMK_PB_SCAN 0  // 1: phase B stops a molecule's chain when its leeway is used up; the offsets after that point come from a block-wide prefix sum of the lengths (all threads) instead of the molecule thread's serial loop.  Measured (C5, alternating builds on one box, 63 GPU tests green with it): placement 17.53 ms against 17.79 ms — the serial tail is not what the other warps wait for; kept as a switch
#endif
constexpr size_t PB_SMEM = sizeof(InsTable) + (size_t)PB_CAP * 12u + (size_t)(PB_MAXM + 1u) * 16u + (size_t)PB_SORT_MAX * 2u + 16u +
                           (MMK_PB_SCAN ? (size_t)(PB_MAXM + 1u) * 2u + 2u : 0u);

struct PlaceArgs {
    SimConst sc;
    const SchemaDev *sch;
    PlanDev plan;
    uint64_t mol0;              // first molecule of the shard in the plan
    uint32_t n_mol;             // molecules in the shard
    uint32_t n_tiles;
    uint64_t n_pairs;           // pairs in the shard
    const uint64_t *roff;       // [n_mol + 1] shard-relative first pair of every molecule
    const uint64_t *woff;       // [n_mol + 1] weight prefix: PB_W per molecule + its pairs unless it is big
    const uint32_t *blk;        // [n_blocks + 1] first molecule of every block
    uint32_t n_blocks;
    const uint32_t *tile_mol0;  // [n_tiles] molecule holding the first pair of every tile
    const InsTable *tab;
    const uint32_t *big_list;   // shard molecules with more than PB_BIG pairs
    const uint32_t *big_count;
    uint32_t big_cap;
    MolRec *mol;
    uint4 *fdesc;
    uint32_t *sum1, *sum2;       // [n_tiles] record bytes per tile and stream; zero on entry
    uint8_t *tile_flags;         // [n_tiles] zero on entry; 1 = the tile holds a pair the fast emit path cannot take
    unsigned long long *gather;  // [256]
    unsigned long long *err;     // [0]: first molecule (plan index) with a non-zero status
    int32_t bpb;
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

// R2 window bstop - r2 + 1 : bstop must start at >= 1 (Julia BoundsError otherwise): insert end offset >= r2 + 1 - start
__device__ __forceinline__ uint32_t need_end_of(const SimConst &sc, int64_t start) {
    return (!sc.strict && start <= (int64_t)sc.r2) ? (uint32_t)((int64_t)sc.r2 + 1 - start) : 0u;
}

// per molecule: weight for the block partition, the tiles whose first pair it holds, the list of big molecules
__global__ void k_mol_weights(const uint64_t *__restrict__ roff_l, uint32_t n_mol, uint32_t *tile_mol0, uint32_t *weight, uint32_t *big_list,
                              uint32_t *big_count, uint32_t big_cap) {
    const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_mol) return;
    const uint64_t lo = roff_l[m], hi = roff_l[m + 1];  // pairs [lo, hi)
    const uint32_t n = (uint32_t)(hi - lo);
    weight[m] = PB_W + (n > PB_BIG ? 0u : n);
    for (uint64_t t = (lo + 31u) >> 5; (t << 5) < hi; ++t) tile_mol0[t] = m;
    if (n > PB_BIG) {
        const uint32_t k = atomicAdd(big_count, 1u);
        if (k < big_cap) big_list[k] = m;
    }
}
// blk[b] = first molecule whose weight prefix is >= b * PB_TARGET (blocks past the total weight are empty)
__global__ void k_place_blocks(const uint64_t *__restrict__ woff, uint32_t n_mol, uint32_t n_blocks, uint32_t *blk) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > n_blocks) return;
    const uint64_t target = (uint64_t)b * PB_TARGET;
    uint32_t lo = 0, hi = n_mol;
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (woff[mid] >= target) hi = mid; else lo = mid + 1;
    }
    blk[b] = b == n_blocks ? n_mol : lo;
}

// whole-vector redraws of one molecule's lengths in shared memory while the sum exceeds the molecule
// (Breakpoints.jl:209-212); true if it still does after `attempts` rounds.  Rare: kept out of line.
__device__ __noinline__ bool redraw_in_smem(const SimConst &sc, const InsTable &tab, uint64_t g, uint32_t m, uint32_t nfrags, uint32_t molsize,
                                            uint16_t *s_len, uint64_t &isum) {
    bool over = true;
    for (int attempts = 1; over && attempts <= sc.attempts; ++attempts) {
        const uint64_t key = mmk_round_key(sc.seed, (uint32_t)attempts);
        isum = 0;
        for (uint32_t k = 0; k < nfrags; ++k) {
            const uint32_t li = pair_len(sc, tab, key, g, m, k);
            s_len[k] = (uint16_t)li;
            isum += li;
        }
        over = sum_over(sc, key, g, m, nfrags, molsize, isum);
    }
    return over;
}

__global__ void __launch_bounds__(PB_THREADS, 4) k_place_block(const __grid_constant__ PlaceArgs a) {
    extern __shared__ __align__(16) uint8_t pb_smem[];
    InsTable *tab = reinterpret_cast<InsTable *>(pb_smem);
    uint2 *s_bits = reinterpret_cast<uint2 *>(pb_smem + sizeof(InsTable));  // gap bits; after phase B: .x = insert offset in the molecule
    uint32_t *s_ro = reinterpret_cast<uint32_t *>(s_bits + PB_CAP);          // [nm + 1] first pair of every molecule, block-relative
    uint32_t *s_so = s_ro + PB_MAXM + 1;                                     // [nm + 1] first shared-memory slot (big molecules have none)
    uint32_t *s_start = s_so + PB_MAXM + 1;                                  // molecule start
    uint32_t *s_meta = s_start + PB_MAXM + 1;                                // MPK word
    uint16_t *s_len = reinterpret_cast<uint16_t *>(s_meta + PB_MAXM + 1);    // [PB_CAP]
    uint16_t *s_mj = s_len + PB_CAP;                                         // [PB_CAP] block-local molecule of every slot
    uint16_t *s_perm = s_mj + PB_CAP;                                        // [PB_SORT_MAX] molecules by decreasing insert count; after phase B: warp totals of the scan
#if MMK_PB_SCAN
    uint16_t *s_kstop = s_perm + PB_SORT_MAX;                                // [nm] inserts placed by the chain (the rest: prefix sum)
#endif
    const SimConst &sc = a.sc;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t m_lo = a.blk[blockIdx.x], m_hi = a.blk[blockIdx.x + 1u];
    if (m_hi <= m_lo) return;
    const uint32_t nm = m_hi - m_lo;  // <= PB_MAXM
    stage_table(tab, a.tab);
    const uint64_t R0 = a.roff[m_lo];
    const uint64_t W0 = a.woff[m_lo];
    for (uint32_t j = tid; j <= nm; j += PB_THREADS) {
        s_ro[j] = (uint32_t)(a.roff[m_lo + j] - R0);
        s_so[j] = (uint32_t)(a.woff[m_lo + j] - W0) - PB_W * j;
        s_start[j] = 0;  // phase A: integer sum of the molecule's insert lengths
    }
    __syncthreads();
    const uint32_t np = s_ro[nm];        // pairs of the block, big molecules included
    const uint32_t ns = s_so[nm];        // ... in shared memory (big molecules have no slots)
    const bool any_big = ns != np;       // otherwise slot == block-relative pair
    // ---- phase A: the draws; a thread takes `spt` consecutive slots (odd: conflict-free shared-memory stores) ---------
    {
        const uint32_t spt = ((ns + PB_THREADS - 1u) / PB_THREADS) | 1u;
        const uint32_t s_begin = tid * spt, s_end = min(s_begin + spt, ns);
        if (s_begin < ns) {
            uint32_t lo = 0, hi = nm - 1u;
            while (lo < hi) {  // largest j with s_so[j] <= s_begin (molecules without slots share their successor's: skipped)
                const uint32_t mid = (lo + hi + 1u) >> 1;
                if (s_so[mid] <= s_begin) lo = mid; else hi = mid - 1u;
            }
            uint32_t j = lo, next = s_so[j + 1u], first = s_so[j], acc = 0;
            uint64_t g = a.plan.group[a.mol0 + m_lo + j];
            uint32_t ms = a.plan.slot[a.mol0 + m_lo + j];
            for (uint32_t slot = s_begin; slot < s_end; ++slot) {
                if (slot >= next) {  // the next molecule that has slots
                    if (acc) atomicAdd(&s_start[j], acc);
                    acc = 0;
                    do {
                        ++j;
                        first = next;
                        next = s_so[j + 1u];
                    } while (slot >= next);
                    g = a.plan.group[a.mol0 + m_lo + j];
                    ms = a.plan.slot[a.mol0 + m_lo + j];
                }
                uint32_t blo, bhi, li;
                if (pair_draw_fast(sc, *tab, g, ms, slot - first, li, blo, bhi)) acc += li;
                else li = PB_EXACT;
                s_bits[slot] = make_uint2(blo, bhi);
                s_len[slot] = (uint16_t)li;
                s_mj[slot] = (uint16_t)j;
            }
            if (acc) atomicAdd(&s_start[j], acc);
            for (uint32_t slot = s_begin; slot < s_end; ++slot) {  // draws off the table: the exact definition (1e-5 of them)
                if (s_len[slot] != PB_EXACT) continue;
                const uint32_t jj = s_mj[slot];
                const uint64_t gi = a.mol0 + m_lo + jj;
                NoRec rec;
                const uint32_t li = insert_len_of(draw_insert(sc, sc.seed, a.plan.group[gi], a.plan.slot[gi], slot - s_so[jj], rec));
                s_len[slot] = (uint16_t)li;
                atomicAdd(&s_start[jj], li);
            }
        }
    }
    // molecules by decreasing insert count, so that the lanes of a warp chain alike (the counts are exponentially distributed)
    const bool sorted = nm <= PB_SORT_MAX;
    if (sorted && tid < nm) {
        const uint32_t n_mine = s_ro[tid + 1u] - s_ro[tid];
        uint32_t rank = 0;
        for (uint32_t i = 0; i < nm; ++i) {
            const uint32_t n_i = s_ro[i + 1u] - s_ro[i];
            rank += (n_i > n_mine) || (n_i == n_mine && i < tid);
        }
        s_perm[rank] = (uint16_t)tid;
    }
    __syncthreads();
    // ---- phase B: one thread per molecule ----------------------------------------------------------------------------
    const uint32_t inl = sc.fq_format == MMK_FMT_TENX ? 16u : 0u;
    for (uint32_t jt = tid; jt < nm; jt += PB_THREADS) {
        const uint32_t j = sorted ? s_perm[jt] : jt;
        const uint32_t ml = m_lo + j;  // shard-relative molecule
        const uint64_t gi = a.mol0 + ml;
        const uint32_t nfrags = s_ro[j + 1u] - s_ro[j];
        if (nfrags > PB_BIG) {  // already placed: fetch what phase C needs
            const MolRec r = a.mol[ml];
            const uint32_t qreg = r.qual1 == (uint32_t)sc.r1 + inl && r.qual2 == (uint32_t)sc.r2;
            s_start[j] = (uint32_t)r.start;
            s_meta[j] = MPK((uint32_t)r.schema, (uint32_t)r.prefix_len, (uint32_t)r.suffix_len, qreg, 1u);
            continue;
        }
        const uint64_t g = a.plan.group[gi];
        const uint32_t ms = a.plan.slot[gi];
        const uint32_t t = a.plan.schema[gi];
        const uint32_t molsize = a.plan.size[gi];
        const SchemaDev &S = a.sch[t];
        const uint32_t n_hap = S.len;
        const uint32_t s0 = s_so[j];
        uint64_t isum = s_start[j];
        bool over = sum_over(sc, sc.seed, g, ms, nfrags, molsize, isum);
        if (over) over = redraw_in_smem(sc, *tab, g, ms, nfrags, molsize, s_len + s0, isum);
        // start_pos = rand(1:adjusted_end) (Breakpoints.jl:240)
        const uint64_t adjusted_end = sc.circular ? (uint64_t)n_hap : (n_hap > molsize + 1u ? (uint64_t)(n_hap - molsize) : 1ull);
        mmk_stream st;
        mmk_stream_init(&st, sc.seed, sc.sample, MMK_P_START, g, ms);
        const uint32_t start = (uint32_t)mmk_next_index(&st, adjusted_end) + 1u;
        uint32_t status = over ? (uint32_t)MMK_ERR_INSERTS_UNFITTABLE : 0u;
        uint32_t q1 = inl, q2 = 0;
        if (!status) {
            const uint32_t need_end = need_end_of(sc, (int64_t)start);
            uint32_t remaining = (uint32_t)((uint64_t)molsize + 1u - isum), cursor = 0;
            uint2 bits = s_bits[s0];
            uint32_t len = s_len[s0];
            uint32_t k = 0;
            // Every gap is uniform on 0..remaining, so the leeway roughly halves per insert and is used up after ~log2(leeway)
            // of them: from there on every gap is mulhi64(bits, 1) = 0 and the offsets are a running sum of the lengths.
            for (; k < nfrags && remaining != 0u; ++k) {
                const uint32_t kn = k + 1u < nfrags ? k + 1u : k;  // next step's operands do not depend on the chain
                const uint2 nbits = s_bits[s0 + kn];
                const uint32_t nlen = s_len[s0 + kn];
                // gap = mulhi64(bits, remaining + 1) with a 32-bit multiplier: two 32 x 32 -> 64 products
                const uint32_t n1 = remaining + 1u;  // <= molsize + 1 < 2^32
                const uint64_t lo64 = (uint64_t)bits.x * n1;
                const uint64_t hi64 = (uint64_t)bits.y * n1 + (lo64 >> 32);
                const uint32_t gap = (uint32_t)(hi64 >> 32);
                const uint32_t so = cursor + gap;
                s_bits[s0 + k].x = so;
                cursor = so + len;
                remaining -= gap;
                if (so + len < need_end) status = MMK_ERR_BOUNDS;
                bits = nbits;
                len = nlen;
            }
#if MMK_PB_SCAN
            // no leeway left: inserts k.. lie back to back from `cursor` on — their offsets are a prefix sum of the lengths, taken by the
            // whole block after this phase.  so + len grows with the insert index, so the bound on the R2 window is decided by insert 0.
            s_kstop[j] = (uint16_t)k;
            if (k == 0u && len < need_end) status = MMK_ERR_BOUNDS;
            if (!status) first_read_quals(sc, (int64_t)start, k ? s_bits[s0].x : 0u, s_len[s0], n_hap, q1, q2);
#else
            for (; k < nfrags; ++k) {  // no leeway left: back to back
                const uint32_t l = s_len[s0 + k];
                s_bits[s0 + k].x = cursor;
                cursor += l;
                if (cursor < need_end) status = MMK_ERR_BOUNDS;
            }
            if (!status) first_read_quals(sc, (int64_t)start, s_bits[s0].x, s_len[s0], n_hap, q1, q2);
#endif
        }
#if MMK_PB_SCAN
        else s_kstop[j] = 0;
#endif
        uint32_t P, SUF;
        header_lens(sc, S, (int64_t)start, molsize, gi, P, SUF);
        const uint32_t qreg = q1 == (uint32_t)sc.r1 + inl && q2 == (uint32_t)sc.r2;
        s_start[j] = start;
        s_meta[j] = MPK(t, P, SUF, qreg, 0u);
        if (status) {
            atomicMin(&a.err[0], (unsigned long long)gi);
            a.plan.status[gi] = (uint8_t)status;
        }
        MolRec r;
        r.start = (int64_t)start;
        r.size = molsize;
        r.schema = (uint16_t)t;
        r.prefix_len = (uint8_t)P;
        r.suffix_len = (uint8_t)SUF;
        r.qual1 = q1;
        r.qual2 = q2;
        a.mol[ml] = r;
    }
    __syncthreads();
#if MMK_PB_SCAN
    // ---- exclusive prefix sum of the (final) lengths over the block's slots, into the high word of the dead gap bits ----------
    {
        const uint32_t spt = (ns + PB_THREADS - 1u) / PB_THREADS;
        const uint32_t b0 = min(tid * spt, ns), b1 = min(b0 + spt, ns);
        uint32_t tsum = 0;
        for (uint32_t slot = b0; slot < b1; ++slot) tsum += s_len[slot];
        uint32_t inc = tsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
            if ((int)lane >= o) inc += y;
        }
        uint32_t *s_wtot = reinterpret_cast<uint32_t *>(s_perm);  // the permutation is no longer needed
        if (lane == 31u) s_wtot[warp] = inc;
        __syncthreads();
        uint32_t run = inc - tsum;
        for (uint32_t w = 0; w < warp; ++w) run += s_wtot[w];
        for (uint32_t slot = b0; slot < b1; ++slot) {
            s_bits[slot].y = run;
            run += s_len[slot];
        }
    }
    __syncthreads();
#endif
    // ---- phase C: one warp per tile of the emit kernel ---------------------------------------------------------------
    const uint64_t P1 = R0 + np;
    const uint32_t t_lo = (uint32_t)(R0 >> 5), t_hi = (uint32_t)((P1 + 31u) >> 5);
    const uint32_t gb_nominal = (uint32_t)((sc.r1 * a.bpb + 7) / 8 + (sc.r2 * a.bpb + 7) / 8);
    unsigned long long gsum = 0;
    for (uint32_t tile = t_lo + warp; tile < t_hi; tile += PB_THREADS / 32u) {
        const uint64_t p = (uint64_t)tile * 32u + lane;
        const bool valid = p >= R0 && p < P1;
        uint32_t len1 = 0, len2 = 0, gb = 0, irr = 0;
        if (valid) {
            const uint32_t i = (uint32_t)(p - R0);
            uint32_t j;
            if (!any_big) j = s_mj[i];
            else {
                uint32_t lo = 0, hi = nm - 1u;
                while (lo < hi) {
                    const uint32_t mid = (lo + hi + 1u) >> 1;
                    if (s_ro[mid] <= i) lo = mid; else hi = mid - 1u;
                }
                j = lo;
            }
            const uint32_t pk = s_meta[j];
            const uint32_t ml = m_lo + j;
            const uint32_t dm = ml - a.tile_mol0[tile];
            const uint32_t start = s_start[j];
            uint32_t so, li;
            if (MPK_BIG(pk)) {  // k_place_big left (insert start, length) in the descriptor's first two words
                const uint4 q = a.fdesc[p];
                so = q.x - start;
                li = MMK_INS_LEN(q.y);
            } else {
                const uint32_t slot = s_so[j] + (i - s_ro[j]);
                li = s_len[slot];
#if MMK_PB_SCAN
                const uint32_t ks = s_kstop[j], k = i - s_ro[j];
                if (k < ks) so = s_bits[slot].x;
                else {  // back to back after the chain's last insert
                    const uint32_t bs = s_so[j] + ks;
                    const uint32_t cur = ks ? s_bits[bs - 1u].x + s_len[bs - 1u] : 0u;
                    so = cur + (s_bits[slot].y - s_bits[bs].y);
                }
#else
                so = s_bits[slot].x;
#endif
            }
            const SchemaDev &S = a.sch[MPK_SCHEMA(pk)];
            const uint32_t n_hap = S.len;
            const uint32_t bstart = start + so, bend = bstart + li - 1u;
            const uint32_t P = MPK_P(pk), SUF = MPK_SUF(pk);
            const uint32_t hdr = P + ndigits_u32(bstart) + 1u + ndigits_u32(bend) + SUF;
            uint32_t sq1 = (uint32_t)sc.r1, sq2 = (uint32_t)sc.r2;
            const bool inside = bend <= n_hap && bstart + sq1 - 1u <= n_hap;  // both windows inside the haplotype: nominal lengths
            const bool a2ok = bend >= sq2;                                    // the R2 window starts at >= 1
            if (!inside) {
                sq1 = window_len(bstart, (int64_t)bstart + sc.r1 - 1, n_hap, sc.strict);
                sq2 = window_len((int64_t)bend - sc.r2 + 1, bend, n_hap, sc.strict);
                gb = (sq1 * a.bpb + 7) / 8 + (sq2 * a.bpb + 7) / 8;
            } else
                gb = gb_nominal;
            uint32_t q1 = (uint32_t)sc.r1 + inl, q2 = (uint32_t)sc.r2;
            if (!MPK_QREG(pk)) {  // quality lengths of the molecule's first reads (FormatFASTQ.jl:14): from its record
                const MolRec r = a.mol[ml];
                q1 = r.qual1;
                q2 = r.qual2;
            }
            len1 = hdr + inl + sq1 + 3u + q1 + 1u;
            len2 = hdr + sq2 + 3u + q2 + 1u;
            irr = !(inside && a2ok && MPK_QREG(pk) && P <= 64u && SUF <= 40u);
            a.fdesc[p] = make_uint4(bstart, li | (dm << 24), (uint32_t)S.word_off, (P & 0x7fu) | (SUF << 7) | (MPK_SCHEMA(pk) << 13));
        }
        const uint32_t tot1 = __reduce_add_sync(0xffffffffu, len1), tot2 = __reduce_add_sync(0xffffffffu, len2);
        gsum += __reduce_add_sync(0xffffffffu, gb);
        const bool irregular = __any_sync(0xffffffffu, irr != 0u);
        if (lane == 0) {
            const bool whole = (uint64_t)tile * 32u >= R0 && ((uint64_t)tile * 32u + 32u <= P1 || P1 == a.n_pairs);
            if (whole) {
                a.sum1[tile] = tot1;
                a.sum2[tile] = tot2;
            } else {  // a tile shared with the neighbouring block
                atomicAdd(&a.sum1[tile], tot1);
                atomicAdd(&a.sum2[tile], tot2);
            }
            if (irregular) a.tile_flags[tile] = 1;
        }
    }
    if (lane == 0 && gsum) atomicAdd(&a.gather[(blockIdx.x * 8u + warp) & 255u], gsum);
}

// One CTA per molecule with more than PB_BIG pairs: lengths and gap bits in chunks through shared memory, the chain by
// one thread.  Writes the molecule's inserts and its record.
constexpr uint32_t BIGC = 1024;  // pairs per chunk
__global__ void __launch_bounds__(PB_THREADS) k_place_big(const __grid_constant__ PlaceArgs a) {
    __shared__ InsTable tab;
    __shared__ __align__(16) uint32_t s_lo[2 * BIGC];  // gap bits: low words, then high words (s_hi); doubles in the float-sum pass
    uint32_t *const s_hi = s_lo + BIGC;
    __shared__ uint32_t s_len[BIGC];
    __shared__ double s_sum;
    __shared__ unsigned long long s_red[PB_THREADS / 32];
    __shared__ unsigned long long s_isum;
    __shared__ uint32_t s_flag;
    __shared__ unsigned long long s_state[3];  // remaining, cursor, status
    const SimConst &sc = a.sc;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t count = min(*a.big_count, a.big_cap);
    if (blockIdx.x >= count) return;
    stage_table(&tab, a.tab);
    __syncthreads();
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
            for (uint32_t k = tid; k < nfrags; k += PB_THREADS) part += pair_len(sc, tab, key, g, ms, k);
            for (int o = 16; o; o >>= 1) part += __shfl_down_sync(0xffffffffu, part, o);
            __syncthreads();  // s_red / s_flag of the previous round are consumed
            if (lane == 0) s_red[warp] = part;
            __syncthreads();
            if (tid == 0) {
                unsigned long long tot = 0;
                for (uint32_t w = 0; w < PB_THREADS / 32; ++w) tot += s_red[w];
                s_isum = tot;
                // 0 fits, 1 too long, 2 the integer bounds straddle the molecule: the float sum decides (sum_over)
                s_flag = tot >= (unsigned long long)molsize + 2u ? 1u : (tot + nfrags + 1u <= (unsigned long long)molsize ? 0u : 2u);
            }
            __syncthreads();
            isum = s_isum;
            if (s_flag == 2u) {
                // the left-to-right float sum of the oracle: the draws in parallel (double precision, 1024 at a time through
                // the shared-memory words of the chain below), the additions in order by one thread
                double *s_x = reinterpret_cast<double *>(s_lo);  // s_lo and s_hi are contiguous: 1024 doubles
                for (uint32_t c0 = 0; c0 < nfrags; c0 += BIGC) {
                    const uint32_t cn = min(BIGC, nfrags - c0);
                    __syncthreads();
                    for (uint32_t k = tid; k < cn; k += PB_THREADS) {
                        NoRec rec;
                        s_x[k] = draw_insert(sc, key, g, ms, c0 + k, rec);
                    }
                    __syncthreads();
                    if (tid == 0) {
                        double sum = c0 ? s_sum : 0.0;
                        for (uint32_t k = 0; k < cn; ++k) sum = MMK_ADD(sum, s_x[k]);
                        s_sum = sum;
                    }
                }
                __syncthreads();
                if (tid == 0) s_flag = s_sum > (double)molsize ? 1u : 0u;
                __syncthreads();
            }
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
            const uint32_t need_end = need_end_of(sc, start);
            if (tid == 0) {
                s_state[0] = (uint64_t)molsize + 1 - isum;
                s_state[1] = 0;
                s_state[2] = 0;
            }
            for (uint32_t c0 = 0; c0 < nfrags; c0 += BIGC) {
                const uint32_t cn = min(BIGC, nfrags - c0);
                __syncthreads();  // previous chunk written out
                for (uint32_t k = tid; k < cn; k += PB_THREADS) {
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
                for (uint32_t k = tid; k < cn; k += PB_THREADS)  // k_place_block completes the descriptor
                    a.fdesc[ro + c0 + k] = make_uint4((uint32_t)start + s_lo[k], s_len[k], 0u, 0u);
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

// (insert start - molecule start, length | dm << 24) per pair from the descriptors: for mmk_dump_inserts / mmk_dump_variates
__global__ void k_ins_from_fdesc(const uint4 *__restrict__ fdesc, const MolRec *__restrict__ mol, const uint32_t *__restrict__ tile_mol0,
                                 uint64_t n_pairs, uint2 *ins) {
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    const uint4 q = fdesc[p];
    const uint32_t m = tile_mol0[p >> 5] + MMK_INS_DM(q.y);
    ins[p] = make_uint2(q.x - (uint32_t)mol[m].start, q.y);
}

// tile_flags: 1 = irregular (from k_place_block) -> 1 = every pair regular and the tile fits the emit kernel's stage
__global__ void k_tile_flags(const uint32_t *__restrict__ sum1, const uint32_t *__restrict__ sum2, uint32_t n_tiles, uint32_t stage_cap,
                             uint8_t *tile_flags) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_tiles) return;
    const uint32_t tot1 = sum1[i], tot2 = sum2[i];
    tile_flags[i] = (uint8_t)(!tile_flags[i] && tot1 + 48u <= stage_cap && tot2 + 48u <= stage_cap && tot1 < 65536u && tot2 < 65536u);
}

}  // namespace mmkdev
