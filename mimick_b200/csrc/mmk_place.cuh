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
    uint2 *ins;
    MolRec *mol;
    uint4 *fdesc;
    uint32_t *sum1, *sum2;       // [n_tiles] record bytes per tile and stream; zero on entry
    uint8_t *tile_flags;         // [n_tiles] zero on entry; 1 = the tile holds a pair the fast emit path cannot take
    unsigned long long *gather;  // [256]
    unsigned long long *err;     // [0]: first molecule (plan index) with a non-zero status
    unsigned long long *msum;    // [n_mol] k_mol_sum -> k_mol_chain: length sum | accepted round << 32 | PLM_OVER
    int32_t bpb;
    uint32_t n_warps;            // warps of the launch; warp w places an equal share of the pairs
};

__device__ __forceinline__ void stage_table(InsTable *dst, const InsTable *src) {
    const uint4 *s = reinterpret_cast<const uint4 *>(src);
    uint4 *d = reinterpret_cast<uint4 *>(dst);
    for (uint32_t i = threadIdx.x; i < sizeof(InsTable) / 16; i += blockDim.x) d[i] = s[i];
}

// R2 window bstop - r2 + 1 : bstop must start at >= 1 (Julia BoundsError otherwise): insert end offset >= r2 + 1 - start
__device__ __forceinline__ uint32_t need_end_of(const SimConst &sc, int64_t start) {
    return (!sc.strict && start <= (int64_t)sc.r2) ? (uint32_t)((int64_t)sc.r2 + 1 - start) : 0u;
}

// Placement of a shard, one LANE per molecule, in two kernels.
// A warp owns a contiguous run of molecules holding 1/n_warps of the shard's pairs; its lanes pull molecules from that
// run one after the other, so all 32 lanes stay busy whatever the molecules' insert counts (they are exponentially
// distributed: a lane per molecule without refill leaves ~6 of 32 lanes active).
//   k_mol_sum    insert k of the lane's molecule: one Philox block -> length (threshold table); integer sum of the
//                  lengths; at the end the redraw rule of Breakpoints.jl:209-212 (whole-vector redraws under the next
//                  round key while the sum exceeds the molecule).  Leaves (sum, accepted round, fits?) per molecule.
//   k_mol_chain  start draw (Breakpoints.jl:240), header-piece lengths, then insert k: length and the 64 bits of the
//                  gap draw (the Philox block again); gap = mulhi64(bits, leeway + 1) (Breakpoints.jl:27-33);
//                  the pair's breakpoints, record lengths and the 16-byte descriptor for the emit kernel; record bytes
//                  accumulated per tile and flushed with one atomic per (molecule, tile) run.
// No shared memory beyond the 5 KB table, no scratch besides 8 bytes per molecule.
constexpr uint32_t PLM_THREADS = 128;
constexpr unsigned long long PLM_OVER = 1ull << 63;  // msum: the length vector never fitted (Breakpoints.jl:214-216)

// molecules [first with roff >= share * gw, first with roff >= share * (gw + 1)): lanes 0 and 1 search the two bounds
__device__ __forceinline__ void warp_molecule_run(const PlaceArgs &a, uint32_t gw, uint32_t lane, uint32_t &m_begin, uint32_t &m_end) {
    uint32_t bnd = 0;
    if (lane < 2) {
        const uint32_t w = gw + lane;
        if (w >= a.n_warps) bnd = a.n_mol;
        else {
            const uint64_t target = (uint64_t)(((unsigned __int128)a.n_pairs * w) / a.n_warps);
            uint32_t lo = 0, hi = a.n_mol;
            while (lo < hi) {
                const uint32_t mid = lo + ((hi - lo) >> 1);
                if (a.roff[mid] >= target) hi = mid; else lo = mid + 1;
            }
            bnd = lo;
        }
    }
    m_begin = __shfl_sync(0xffffffffu, bnd, 0);
    m_end = __shfl_sync(0xffffffffu, bnd, 1);
}

__global__ void __launch_bounds__(PLM_THREADS, 8) k_mol_sum(const __grid_constant__ PlaceArgs a) {
    __shared__ InsTable tab;
    stage_table(&tab, a.tab);
    __syncthreads();
    const SimConst &sc = a.sc;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t gw = blockIdx.x * (PLM_THREADS / 32u) + (threadIdx.x >> 5);
    if (gw >= a.n_warps) return;
    uint32_t m_next, m_end;
    warp_molecule_run(a, gw, lane, m_next, m_end);
    const uint32_t lt_mask = (1u << lane) - 1u;
    bool idle = true, done = false;
    uint32_t ml = 0, ms = 0, n = 0, k = 0, molsize = 0, round = 0;
    uint64_t g = 0, ro = 0, key = sc.seed, isum = 0;
    for (;;) {
        const unsigned need = __ballot_sync(0xffffffffu, idle && !done);
        if (need) {
            if (idle && !done) {
                const uint32_t idx = m_next + __popc(need & lt_mask);
                if (idx < m_end) {
                    ml = idx;
                    const uint64_t gi = a.mol0 + ml;
                    g = a.plan.group[gi];
                    ms = a.plan.slot[gi];
                    molsize = a.plan.size[gi];
                    ro = a.roff[ml];
                    n = (uint32_t)(a.roff[ml + 1u] - ro);
                    k = 0;
                    isum = 0;
                    round = 0;
                    key = sc.seed;
                    idle = false;
                } else
                    done = true;
            }
            m_next += __popc(need);
        }
        if (__all_sync(0xffffffffu, done)) break;
        if (done) continue;
        mmk_stream is;
        mmk_insert_stream_init(&is, key, sc.sample, g, ms, k);
        mmk_stream_block(&is, 0, is.buf);
        is.blk = 0;
        uint32_t li;
        if (!(tab.enabled && ins_len_fast(tab, is.buf[1], li))) {
            NoRec rec;
            li = insert_len_of(draw_insert_from(sc, is, rec));
        }
        isum += li;
        if (++k < n) continue;
        // the vector is drawn: does it fit (Breakpoints.jl:209-212)?
        const bool over = sum_over(sc, key, g, ms, n, molsize, isum);
        if (over && round < (uint32_t)sc.attempts) {  // whole-vector redraw under the next round's key
            ++round;
            key = mmk_round_key(sc.seed, round);
            k = 0;
            isum = 0;
            continue;
        }
        a.msum[ml] = (over ? PLM_OVER : 0ull) | ((unsigned long long)round << 32) | (unsigned long long)(uint32_t)isum;
        idle = true;
    }
}

__global__ void __launch_bounds__(PLM_THREADS, 5) k_mol_chain(const __grid_constant__ PlaceArgs a) {
    __shared__ InsTable tab;
    stage_table(&tab, a.tab);
    __syncthreads();
    const SimConst &sc = a.sc;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t gw = blockIdx.x * (PLM_THREADS / 32u) + (threadIdx.x >> 5);
    if (gw >= a.n_warps) return;
    uint32_t m_next, m_end;
    warp_molecule_run(a, gw, lane, m_next, m_end);
    const uint32_t inl = sc.fq_format == MMK_FMT_TENX ? 16u : 0u;
    const uint32_t lt_mask = (1u << lane) - 1u;
    bool idle = true, done = false;
    // the molecule in hand
    uint32_t ml = 0, ms = 0, n = 0, k = 0, t = 0, molsize = 0, n_hap = 0, word_off = 0, round = 0;
    uint64_t g = 0, ro = 0, key = sc.seed;
    uint32_t start = 0, remaining = 0, cursor = 0, need_end = 0, P = 0, SUF = 0, q1 = 0, q2 = 0, qreg = 0, status = 0;
    // per-tile run
    uint32_t cur_tile = 0, tm0 = 0, acc1 = 0, acc2 = 0, irr = 0;
    unsigned long long gacc = 0;
    const uint32_t gb_nominal = (uint32_t)((sc.r1 * a.bpb + 7) / 8 + (sc.r2 * a.bpb + 7) / 8);
    for (;;) {
        const unsigned need = __ballot_sync(0xffffffffu, idle && !done);
        if (need) {
            if (idle && !done) {
                const uint32_t idx = m_next + __popc(need & lt_mask);
                if (idx < m_end) {
                    ml = idx;
                    const uint64_t gi = a.mol0 + ml;
                    g = a.plan.group[gi];
                    ms = a.plan.slot[gi];
                    t = a.plan.schema[gi];
                    molsize = a.plan.size[gi];
                    ro = a.roff[ml];
                    n = (uint32_t)(a.roff[ml + 1u] - ro);
                    n_hap = a.sch[t].len;
                    word_off = (uint32_t)a.sch[t].word_off;
                    const unsigned long long sm = a.msum[ml];
                    round = (uint32_t)(sm >> 32) & 0x7fffffffu;
                    key = mmk_round_key(sc.seed, round);
                    // start_pos = rand(1:adjusted_end) (Breakpoints.jl:240)
                    const uint64_t adjusted_end = sc.circular ? (uint64_t)n_hap : (n_hap > molsize + 1u ? (uint64_t)(n_hap - molsize) : 1ull);
                    mmk_stream st;
                    mmk_stream_init(&st, sc.seed, sc.sample, MMK_P_START, g, ms);
                    start = (uint32_t)mmk_next_index(&st, adjusted_end) + 1u;
                    header_lens(sc, a.sch[t], (int64_t)start, molsize, gi, P, SUF);
                    need_end = need_end_of(sc, (int64_t)start);
                    status = (sm & PLM_OVER) ? (uint32_t)MMK_ERR_INSERTS_UNFITTABLE : 0u;
                    q1 = inl;
                    q2 = 0;
                    qreg = 0;
                    remaining = (uint32_t)((uint64_t)molsize + 1 - (uint32_t)sm);
                    cursor = 0;
                    k = 0;
                    idle = false;
                    if (status) {  // Breakpoints.jl:214-216: nothing to chain
                        atomicMin(&a.err[0], (unsigned long long)gi);
                        a.plan.status[gi] = (uint8_t)status;
                        MolRec r;
                        r.start = (int64_t)start;
                        r.size = molsize;
                        r.schema = (uint16_t)t;
                        r.prefix_len = (uint8_t)P;
                        r.suffix_len = (uint8_t)SUF;
                        r.qual1 = q1;
                        r.qual2 = q2;
                        a.mol[ml] = r;
                        idle = true;  // next molecule on the next pass of the loop
                    }
                } else
                    done = true;
            }
            m_next += __popc(need);
        }
        if (__all_sync(0xffffffffu, done)) break;
        if (done || idle) continue;
        const uint64_t p = ro + k;
        uint32_t blo, bhi, li;
        {   // the insert's Philox block again (carrying 12 bytes per pair through memory instead measured 40 % slower)
            mmk_stream is;
            mmk_insert_stream_init(&is, key, sc.sample, g, ms, k);
            mmk_stream_block(&is, 0, is.buf);
            is.blk = 0;
            blo = is.buf[2];
            bhi = is.buf[3];
            if (!(tab.enabled && ins_len_fast(tab, is.buf[1], li))) {
                NoRec rec;
                li = insert_len_of(draw_insert_from(sc, is, rec));
            }
            if (round) {  // lengths of the accepted redraw round; the gap bits are always those of round 0 (rare)
                mmk_stream gs;
                mmk_insert_stream_init(&gs, sc.seed, sc.sample, g, ms, k);
                uint32_t o[4];
                mmk_stream_block(&gs, 0, o);
                blo = o[2];
                bhi = o[3];
            }
        }
        const uint32_t tile = (uint32_t)(p >> 5);
        if (k == 0 || tile != cur_tile) {
            if (k != 0) {  // the run in the previous tile
                atomicAdd(&a.sum1[cur_tile], acc1);
                atomicAdd(&a.sum2[cur_tile], acc2);
                if (irr) a.tile_flags[cur_tile] = 1;
            }
            cur_tile = tile;
            tm0 = a.tile_mol0[tile];
            acc1 = acc2 = irr = 0;
        }
        const uint32_t gap = (uint32_t)mmk_bits_to_index(((uint64_t)bhi << 32) | blo, (uint64_t)remaining + 1);
        const uint32_t st = cursor + gap;
        cursor = st + li;
        remaining -= gap;
        if (st + li < need_end) status = MMK_ERR_BOUNDS;
        if (k == 0) {  // quality lengths follow the molecule's FIRST reads (FormatFASTQ.jl:14)
            first_read_quals(sc, (int64_t)start, st, li, n_hap, q1, q2);
            qreg = q1 == (uint32_t)sc.r1 + inl && q2 == (uint32_t)sc.r2;
        }
        const uint32_t dm = ml - tm0;
        const uint32_t bstart = start + st, bend = bstart + li - 1u;
        const uint32_t hdr = P + ndigits_u32(bstart) + 1u + ndigits_u32(bend) + SUF;
        uint32_t sq1 = (uint32_t)sc.r1, sq2 = (uint32_t)sc.r2;
        const bool inside = bend <= n_hap && bstart + sq1 - 1u <= n_hap;  // both windows inside the haplotype: nominal lengths
        const bool a2ok = bend >= sq2;                                   // the R2 window starts at >= 1
        if (!inside) {
            sq1 = window_len(bstart, (int64_t)bstart + sc.r1 - 1, n_hap, sc.strict);
            sq2 = window_len((int64_t)bend - sc.r2 + 1, bend, n_hap, sc.strict);
            gacc += (sq1 * a.bpb + 7) / 8 + (sq2 * a.bpb + 7) / 8;
        } else
            gacc += gb_nominal;
        acc1 += hdr + inl + sq1 + 3u + q1 + 1u;
        acc2 += hdr + sq2 + 3u + q2 + 1u;
        irr |= !(inside && a2ok && qreg && P <= 64u && SUF <= 40u);
        a.ins[p] = make_uint2(st, li | (dm << 24));
        a.fdesc[p] = make_uint4(bstart, li | (dm << 24), word_off, (P & 0x7fu) | (SUF << 7) | (t << 13));
        if (++k < n) continue;
        // the molecule is placed
        atomicAdd(&a.sum1[cur_tile], acc1);
        atomicAdd(&a.sum2[cur_tile], acc2);
        if (irr) a.tile_flags[cur_tile] = 1;
        if (status) {
            atomicMin(&a.err[0], (unsigned long long)(a.mol0 + ml));
            a.plan.status[a.mol0 + ml] = (uint8_t)status;
        }
        MolRec r;
        r.start = (int64_t)start;
        r.size = molsize;
        r.schema = (uint16_t)t;
        r.prefix_len = (uint8_t)P;
        r.suffix_len = (uint8_t)SUF;
        r.qual1 = status ? inl : q1;
        r.qual2 = status ? 0u : q2;
        a.mol[ml] = r;
        idle = true;
    }
    if (gacc) atomicAdd(&a.gather[gw & 255u], gacc);
}

// tile_flags: 1 = irregular (from k_place_mol) -> 1 = every pair regular and the tile fits the emit kernel's stage
__global__ void k_tile_flags(const uint32_t *__restrict__ sum1, const uint32_t *__restrict__ sum2, uint32_t n_tiles, uint32_t stage_cap,
                             uint8_t *tile_flags) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_tiles) return;
    const uint32_t tot1 = sum1[i], tot2 = sum2[i];
    tile_flags[i] = (uint8_t)(!tile_flags[i] && tot1 + 48u <= stage_cap && tot2 + 48u <= stage_cap && tot1 < 65536u && tot2 < 65536u);
}

}  // namespace mmkdev
