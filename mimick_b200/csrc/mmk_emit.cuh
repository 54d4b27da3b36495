// mmk_emit.cuh — records: per-pair decode, the general byte-wise writer, the staged fast path.
//
// Reference behaviour: extract_sequences! / circular_index (Breakpoints.jl:95-140) and
// format_R1 / format_R2 (FormatFASTQ.jl).  One CTA emits one tile of `tile_pairs` consecutive
// read pairs into the R1 and the R2 stream.
//
// Fast path (every pair of the tile is "regular": nominal read lengths, no wrap):
//   1. the tile's R1 and R2 text are built in shared memory at the SAME 16-byte phase as their
//      place in the output stream, so the copy-out is aligned 16-byte stores of the stream;
//   2. the stage is pre-filled with 'I' (quality strings cost nothing more);
//   3. per pair, one warp: each half-warp expands its read from the 2-bit reference — one
//      32-bit word per lane, shuffles + funnel shift + PRMT lookup — as full 16-byte stores that
//      may spill garbage up to 15 bytes either side of the sequence;
//   4. header bytes are written afterwards, one thread per (pair, header piece), and the
//      "\n+\n" / first quality bytes right after the sequence: write order — not masks — resolves
//      every overlap.
// Anything else (4-bit reference, literal circular wrap, reads outside 16..192, huge headers,
// force_generic) goes through write_record_generic, straight to global memory.
#pragma once
#include "mmk_internal.h"
#include "mmk_sampling.cuh"

namespace mmkdev {

constexpr int EMIT_MAX_TP = 64;  // pairs per tile <= threads / 4 (one thread per header piece)

struct PairDesc {
    uint64_t word_off;   // packed haplotype start (32-bit words)
    uint32_t mol;        // shard-relative molecule index
    uint64_t pair;       // shard-relative pair index
    uint32_t bstart, bend;   // 1-based inclusive insert breakpoints (contigs < 2^31 bases)
    uint32_t hdr;        // header line length incl. '\n' (without an inline barcode)
    uint32_t sq1, sq2;   // sequence lengths
    uint32_t len1, len2; // record lengths
    uint32_t q1, q2;     // quality lengths
    uint32_t n;          // haplotype length
    uint8_t d1, d2;      // decimal widths of bstart / bend
    uint8_t P, SUF;      // header prefix / suffix lengths
    uint32_t regular;
};

__device__ __forceinline__ uint32_t window_len(int64_t a, int64_t b, uint32_t n, int strict);

// FD: from the packed descriptor (x = bstart, y = length | dm << 24; every emit kernel) instead of the insert plan (k_tile_prepare,
// which builds the descriptors of the three-kernel placement)
template <bool FD>
__device__ __forceinline__ void decode_pair(const EmitArgs &a, uint64_t p, uint32_t tile_m0, PairDesc &d) {
    uint2 in;
    if (FD) {
        const uint4 q = a.fdesc[p];
        in = make_uint2(q.x, q.y);
    } else
        in = a.ins[p];
    uint32_t m = tile_m0 + MMK_INS_DM(in.y);
    MolRec mr = a.mol[m];
    const SchemaDev &S = a.sch[mr.schema];
    d.word_off = S.word_off;
    d.n = S.len;
    d.mol = m;
    d.pair = p;
    d.bstart = FD ? in.x : (uint32_t)mr.start + in.x;
    d.bend = d.bstart + MMK_INS_LEN(in.y) - 1;
    d.d1 = (uint8_t)ndigits_u32(d.bstart);
    d.d2 = (uint8_t)ndigits_u32(d.bend);
    d.P = mr.prefix_len;
    d.SUF = mr.suffix_len;
    d.hdr = (uint32_t)d.P + d.d1 + 1 + d.d2 + d.SUF;
    int64_t a2 = (int64_t)d.bend - a.r2 + 1;
    if (d.bend <= d.n && d.bstart + (uint32_t)a.r1 - 1u <= d.n) {  // both windows inside the haplotype: nominal lengths
        d.sq1 = (uint32_t)a.r1;
        d.sq2 = (uint32_t)a.r2;
    } else {
        d.sq1 = window_len(d.bstart, (int64_t)d.bstart + a.r1 - 1, d.n, a.strict);
        d.sq2 = window_len(a2, d.bend, d.n, a.strict);
    }
    d.q1 = mr.qual1;
    d.q2 = mr.qual2;
    d.len1 = d.hdr + a.bclen_inline + d.sq1 + 3 + d.q1 + 1;
    d.len2 = d.hdr + d.sq2 + 3 + d.q2 + 1;
    d.regular = d.sq1 == (uint32_t)a.r1 && d.sq2 == (uint32_t)a.r2 && d.q1 == (uint32_t)(a.r1 + a.bclen_inline) &&
                d.q2 == (uint32_t)a.r2 && (uint64_t)d.bstart + a.r1 - 1 <= d.n && d.bend <= d.n && a2 >= 1 && d.P <= 64 &&
                d.SUF <= 40;
}

// 0-based haplotype position of base j of circular_index(seq, a:b), forward orientation.
__device__ __forceinline__ uint32_t window_pos(int64_t a, int64_t b, uint32_t n, int strict, uint32_t j) {
    if (strict) {
        int64_t q = a - 1 + (int64_t)j;  // may be negative when the insert is shorter than R2 near position 1
        if (q < 0) q += (int64_t)n * ((-q) / (int64_t)n + 1);
        return (uint32_t)((uint64_t)q % n);
    }
    if (b <= (int64_t)n) return (uint32_t)(a - 1 + j);
    uint32_t part1 = a <= (int64_t)n ? (uint32_t)((int64_t)n - a + 1) : 0u;
    return j < part1 ? (uint32_t)(a - 1 + j) : j - part1;
}
template <int BPB>
__device__ __forceinline__ uint32_t fetch_code(const uint32_t *__restrict__ w, uint32_t pos) {
    constexpr uint32_t PER = 32 / BPB;
    return (__ldg(&w[pos / PER]) >> (BPB * (pos % PER))) & ((1u << BPB) - 1u);
}

// Sequencing-error model (optional, off by default; the reference has none: docs/algorithm.md:11 lists it as future
// work).  Definition: oracle/mimick_oracle.c erroneous_read.  One warp per read of nominal length: a lane takes 4
// consecutive cycles of a round of 128 (one Philox block, mmk_error_block); an insertion consumes no template base, a
// deletion two, anything else one, so the template offset of a lane's first cycle is a warp scan away.
// put(cycle, base character, quality character) stores the two bytes.
template <int BPB, class Put>
__device__ __forceinline__ void err_read_warp(const EmitArgs &a, const uint32_t *__restrict__ hw, uint32_t n, uint32_t rd, int64_t first_pos,
                                              uint64_t pair, uint32_t lane, Put put) {
    const uint32_t S = rd == 1u ? (uint32_t)a.r1 : (uint32_t)a.r2;
    const uint32_t *thr = a.err_thr + (rd == 1u ? 0u : (uint32_t)a.r1);
    const uint8_t *qb = a.err_qb + (rd == 1u ? 0u : (uint32_t)a.r1);
    uint32_t consumed = 0;  // template bases consumed by the cycles of earlier rounds
    for (uint32_t c0 = 0; c0 < S; c0 += 128u) {
        const uint32_t cb = c0 + 4u * lane;
        uint32_t o[4];
        mmk_error_block(a.seed, a.sample, pair, rd, cb >> 2, o);
        uint32_t ev = 0, cons = 0;  // 2 bits per cycle: 0 none, 1 substitution, 2 insertion, 3 deletion
#pragma unroll
        for (uint32_t k = 0; k < 4; ++k) {
            if (cb + k < S) {
                const uint32_t u = MMK_ERR_U(o[k]), t = __ldg(thr + cb + k);
                const uint32_t e = u < t ? 1u : (u - t < a.err_tins ? 2u : (u - t - a.err_tins < a.err_tdel ? 3u : 0u));
                ev |= e << (2u * k);
                cons += e == 2u ? 0u : (e == 3u ? 2u : 1u);
            }
        }
        uint32_t inc = cons;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= (uint32_t)d) inc += y;
        }
        uint32_t r = consumed + inc - cons;
        consumed += __shfl_sync(0xffffffffu, inc, 31);
#pragma unroll
        for (uint32_t k = 0; k < 4; ++k) {
            const uint32_t c = cb + k;
            if (c >= S) break;
            const uint32_t e = (ev >> (2u * k)) & 3u, sel = MMK_ERR_SEL(o[k]), jit = MMK_ERR_JIT(o[k]);
            uint32_t ch;
            if (e == 2u) {
                ch = (0x54474341u >> (8u * sel)) & 0xffu;  // "ACGT"[sel]
            } else {
                if (e == 3u) ++r;
                int64_t pos = rd == 1u ? first_pos - (int64_t)r : first_pos + (int64_t)r;  // 1-based
                if (a.strict) pos = ((pos - 1) % (int64_t)n + (int64_t)n) % (int64_t)n + 1;
                else pos = pos < 1 ? 1 : (pos > (int64_t)n ? (int64_t)n : pos);
                const uint32_t code = fetch_code<BPB>(hw, (uint32_t)(pos - 1));
                ++r;
                uint32_t idx;  // index in "ACGT", 4 = another IUPAC symbol
                if (BPB == 2) {
                    idx = rd == 1u ? 3u - code : code;
                    ch = (0x54474341u >> (8u * idx)) & 0xffu;
                } else {
                    const uint32_t c4 = rd == 1u ? comp4(code) : code;
                    ch = (uint32_t)(uint8_t)code4_to_ascii(c4);
                    idx = c4 == 1u ? 0u : (c4 == 2u ? 1u : (c4 == 4u ? 2u : (c4 == 8u ? 3u : 4u)));
                }
                if (e == 1u && idx < 4u) ch = (0x54474341u >> (8u * ((idx + 1u + sel % 3u) & 3u))) & 0xffu;
            }
            int q = (int)__ldg(qb + c) + (int)((0x43222210u >> (4u * jit)) & 15u) - 2 - ((e == 1u || e == 2u) ? 10 : 0);
            q = q < 2 ? 2 : (q > 41 ? 41 : q);
            put(c, ch, (uint32_t)(33 + q));
        }
    }
}

// General byte-wise record writer: any read length, wrap mode, quality length. One warp per record.
template <int BPB>
__device__ void write_record_generic(const EmitArgs &a, const PairDesc &d, int rd, uint8_t *out, uint32_t lane) {
    const uint8_t *tx = a.text + (uint64_t)d.mol * a.text_stride;
    const uint32_t *hw = a.hap + d.word_off;
    const uint32_t n = d.n;
    const uint32_t P = d.P, SUF = d.SUF;
    const uint32_t rd_pos = (a.fq_format == MMK_FMT_HAPLOTAGGING || a.fq_format == MMK_FMT_STANDARD) ? 1u : SUF - 13u;
    const uint32_t inl = rd == 1 ? (uint32_t)a.bclen_inline : 0u;
    const uint32_t sq = rd == 1 ? d.sq1 : d.sq2;
    const uint32_t q = rd == 1 ? d.q1 : d.q2;
    const uint32_t len = rd == 1 ? d.len1 : d.len2;
    const int64_t wa = rd == 1 ? (int64_t)d.bstart : (int64_t)d.bend - a.r2 + 1;
    const int64_t wb = rd == 1 ? (int64_t)d.bstart + a.r1 - 1 : (int64_t)d.bend;
    for (uint32_t k0 = lane; k0 < len; k0 += 32) {
        uint32_t k = k0;
        uint8_t c;
        if (k < P) c = tx[k];
        else if ((k -= P) < d.d1) c = (uint8_t)('0' + (d.bstart / c_pow10[d.d1 - 1 - k]) % 10);
        else if ((k -= d.d1) == 0) c = ':';
        else if ((k -= 1) < d.d2) c = (uint8_t)('0' + (d.bend / c_pow10[d.d2 - 1 - k]) % 10);
        else if ((k -= d.d2) < SUF) c = (k == rd_pos) ? (uint8_t)('0' + rd) : tx[P + k];
        else if ((k -= SUF) < inl) c = tx[P + 1 + k];
        else if ((k -= inl) < sq) {
            uint32_t j = rd == 1 ? sq - 1 - k : k;
            uint32_t code = fetch_code<BPB>(hw, window_pos(wa, wb, n, a.strict, j));
            if (BPB == 2) c = (uint8_t)(rd == 1 ? "TGCA"[code] : "ACGT"[code]);
            else c = (uint8_t)code4_to_ascii(rd == 1 ? comp4(code) : code);
        } else if ((k -= sq) < 3) c = (uint8_t)(k == 1 ? '+' : '\n');
        else if ((k -= 3) < q) c = 'I';
        else c = '\n';
        out[k0] = c;
    }
    if (a.err_on && sq == (rd == 1 ? (uint32_t)a.r1 : (uint32_t)a.r2) && q == sq + inl) {  // nominal read: the error model rewrites bases and qualities
        __syncwarp();
        const uint32_t hdr_all = P + d.d1 + 1u + d.d2 + SUF + inl;
        uint8_t *seq = out + hdr_all, *qual = out + hdr_all + sq + 3u + inl;
        err_read_warp<BPB>(a, hw, n, (uint32_t)rd, rd == 1 ? wb : wa, a.pair_base + d.pair, lane, [&](uint32_t c, uint32_t ch, uint32_t qc) {
            seq[c] = (uint8_t)ch;
            qual[c] = (uint8_t)qc;
        });
    }
}

// ------------------------------------------------------------------------------------------
// tile preparation: one decode of every pair → record lengths per tile (byte offsets after a scan),
// a packed 16-byte descriptor per pair for the emit kernel, and the tile's fast/general flag
// ------------------------------------------------------------------------------------------
// x: bstart   y: insert length | dm << 24   w: P | SUF << 7 | schema << 13
// z: 32-pair tiles: first arena word of the pair's haplotype (the emit kernel scans the record lengths itself and
//    needs no dependent load for the window address); other tile sizes: sb1 | sb2 << 16 (tile-relative record offsets)
__global__ void __launch_bounds__(256) k_tile_prepare(EmitArgs a, uint32_t tile_lo, uint32_t n_tiles, uint32_t *sum1, uint32_t *sum2,
                                                      uint4 *fdesc, uint8_t *tile_flags, unsigned long long *gather) {
    // 256 threads = 256 / tile_pairs tiles; tile_pairs is a power of two in 8..64
    __shared__ uint32_t s_w[8][2];
    __shared__ uint32_t s_irr[8];
    const uint32_t TP = a.tile_pairs, tsh = 31 - __clz(TP);
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t tile = tile_lo + blockIdx.x * (256u >> tsh) + (threadIdx.x >> tsh), it = threadIdx.x & (TP - 1u);  // tiles [tile_lo, n_tiles)
    const bool tile_ok = tile < n_tiles;
    const uint64_t p0 = (uint64_t)tile << tsh;
    const uint32_t np = tile_ok ? (uint32_t)((a.n_pairs - p0) < TP ? (a.n_pairs - p0) : TP) : 0u;
    PairDesc d;
    d.regular = 1;
    d.len1 = d.len2 = d.sq1 = d.sq2 = 0;
    uint32_t schema = 0, iny = 0;
    if (it < np) {
        const uint32_t m0 = a.tile_mol0[tile];
        decode_pair<false>(a, p0 + it, m0, d);
        schema = a.mol[d.mol].schema;
        iny = a.ins[p0 + it].y;
    }
    // inclusive scans inside the tile: segments of min(TP, 32) lanes, two warps combined for TP = 64
    const uint32_t W = TP < 32u ? TP : 32u, il = lane & (W - 1u);
    uint32_t i1 = d.len1, i2 = d.len2, gb = (d.sq1 * a.bpb + 7) / 8 + (d.sq2 * a.bpb + 7) / 8;
    for (uint32_t o = 1; o < W; o <<= 1) {
        uint32_t y1 = __shfl_up_sync(0xffffffffu, i1, o), y2 = __shfl_up_sync(0xffffffffu, i2, o);
        if (il >= o) {
            i1 += y1;
            i2 += y2;
        }
    }
    for (int o = 16; o; o >>= 1) gb += __shfl_down_sync(0xffffffffu, gb, o);
    // (accounting only) spread the per-warp totals over 256 counters: same-address atomics serialise in L2
    if (lane == 0 && gb) atomicAdd(&gather[(blockIdx.x * 8u + warp) & 255u], (unsigned long long)gb);
    const unsigned irr = __ballot_sync(0xffffffffu, !d.regular);
    uint32_t tot1 = __shfl_sync(0xffffffffu, i1, (int)((lane & ~(W - 1u)) + W - 1u));
    uint32_t tot2 = __shfl_sync(0xffffffffu, i2, (int)((lane & ~(W - 1u)) + W - 1u));
    uint32_t off1 = i1 - d.len1, off2 = i2 - d.len2;
    bool irregular = (irr & (W == 32u ? 0xffffffffu : (((1u << W) - 1u) << (lane & ~(W - 1u))))) != 0u;
    if (TP == 64u) {
        if (lane == 31) {
            s_w[warp][0] = i1;
            s_w[warp][1] = i2;
            s_irr[warp] = irr;
        }
        __syncthreads();
        const uint32_t w0 = warp & ~1u;
        if (warp & 1u) {
            off1 += s_w[w0][0];
            off2 += s_w[w0][1];
        }
        tot1 = s_w[w0][0] + s_w[w0 + 1][0];
        tot2 = s_w[w0][1] + s_w[w0 + 1][1];
        irregular = (s_irr[w0] | s_irr[w0 + 1]) != 0u;
    }
    if (it < np)
        fdesc[p0 + it] = make_uint4(d.bstart, iny, TP == 32u ? (uint32_t)d.word_off : ((off1 & 0xffffu) | (off2 << 16)),
                                    (uint32_t)d.P | ((uint32_t)d.SUF << 7) | (schema << 13));
    if (tile_ok && it == 0) {
        sum1[tile] = tot1;
        sum2[tile] = tot2;
        tile_flags[tile] = (uint8_t)(!irregular && tot1 + 48u <= a.stage_cap && tot2 + 48u <= a.stage_cap && tot1 < 65536u &&
                                     tot2 < 65536u);
    }
}

// ------------------------------------------------------------------------------------------
// fast path pieces
// ------------------------------------------------------------------------------------------
// 16 bases (32 bits, base i in bits 2i..2i+1) → 16 ASCII bytes through a 4-entry byte LUT.
// `bits & 0x33333333` already has the even bases as PRMT selectors (one nibble each), `bits >> 2`
// the odd ones; four lookups and four interleaving PRMTs give the 16 bytes in order.
__device__ __forceinline__ uint4 expand16(uint32_t bits, uint32_t lut) {
    const uint32_t ev = bits & 0x33333333u, od = (bits >> 2) & 0x33333333u;
    const uint32_t e_lo = __byte_perm(lut, 0, ev), e_hi = __byte_perm(lut, 0, ev >> 16);  // bases 0,2,4,6 / 8,10,12,14
    const uint32_t o_lo = __byte_perm(lut, 0, od), o_hi = __byte_perm(lut, 0, od >> 16);  // bases 1,3,5,7 / 9,11,13,15
    uint4 v;
    v.x = __byte_perm(e_lo, o_lo, 0x5140);
    v.y = __byte_perm(e_lo, o_lo, 0x7362);
    v.z = __byte_perm(e_hi, o_hi, 0x5140);
    v.w = __byte_perm(e_hi, o_hi, 0x7362);
    return v;
}

// PRMT as the hardware has it: bit 3 of a selector nibble asks for sign replication and must be 0 here (__byte_perm masks every
// selector with 0x7777 to give its documented behaviour; the callers below already have clean selectors)
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

// shared memory through 32-bit window addresses (computed once per kernel)
__device__ __forceinline__ void sts128(uint32_t addr, const uint4 &v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void sts8(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}

// Per-pair parameters of the fast path, computed once by the pair's decode thread and read back
// with two 16-byte shared-memory loads.
struct FastDesc {
    uint32_t sb1, sb2;       // stage byte offsets of the R1 / R2 record (stage2 offsets include stage_cap)
    uint32_t bstart, bend;
    uint32_t wbase1, wbase2; // first packed word each half-warp loads (arena word index)
    uint32_t mol;
    uint32_t pk;             // hdr | P << 8 | SUF << 16 | d1 << 22 | d2 << 26   (hdr <= 255, P <= 64, SUF <= 40)
};
static_assert(sizeof(FastDesc) == 32, "FastDesc layout");
// kept in shared memory as two uint4 arrays (16-byte stride: conflict-free writes, broadcast reads)
__device__ __forceinline__ FastDesc load_fast_desc(const uint4 *fa, const uint2 *fb, uint32_t t) {  // header phase (no window words)
    const uint4 x = fa[t];
    const uint2 y = fb[t];
    FastDesc f;
    f.sb1 = x.x; f.sb2 = x.y; f.bstart = x.z; f.bend = x.w;
    f.wbase1 = f.wbase2 = 0; f.mol = y.x; f.pk = y.y;
    return f;
}
#ifndef MMK_EMIT_WPF
#define MMK_EMIT_WPF 0  // 1: warp 1 of k_emit_w prefetches the reference windows of the tile one wave ahead into L2 (measured: 7 % slower — its loads sit on the path to the first barrier)
#endif
// look-ahead distances in tiles: one wave of CTAs is 148 SMs x 8 resident CTAs
#ifndef MMK_EMIT_AHEAD
#define MMK_EMIT_AHEAD 1184
#endif
constexpr uint32_t EMIT_AHEAD_WIN = MMK_EMIT_AHEAD, EMIT_AHEAD_DESC = MMK_EMIT_WPF ? 2u * MMK_EMIT_AHEAD : MMK_EMIT_AHEAD;
#ifndef MMK_EMIT_EARLY
#define MMK_EMIT_EARLY 1  // k_emit_w: window loads issued by every warp before the tile's descriptors are unpacked
#endif
#ifndef MMK_EMIT_ILP
#define MMK_EMIT_ILP 8  // read pairs a warp of k_emit_w expands together (1, 2, 4 or 8)
#endif
#ifndef MMK_PREFETCH_MODE
#define MMK_PREFETCH_MODE 1
#endif
__device__ __forceinline__ void prefetch_l1(const void *p) {
#if MMK_PREFETCH_MODE == 1
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#elif MMK_PREFETCH_MODE == 2
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#endif
}
#define FD_HDR(pk) ((pk) & 0xffu)
#define FD_P(pk) (((pk) >> 8) & 0xffu)
#define FD_SUF(pk) (((pk) >> 16) & 0x3fu)
#define FD_D1(pk) (((pk) >> 22) & 0xfu)
#define FD_D2(pk) (((pk) >> 26) & 0xfu)

// unpack the 16-byte descriptor the placement wrote (x: bstart, y: len | dm << 24, z: see k_tile_prepare,
// w: P | SUF << 7 | schema << 13) into the emit kernel's working form.  ZWORD: z is the haplotype's first arena word and
// the caller fills sb1 / sb2 from its own scan of the record lengths.
template <bool ZWORD>
__device__ __forceinline__ void make_fast_desc(const EmitArgs &a, const uint4 &q, uint32_t tile_m0, uint32_t base1, uint32_t base2,
                                               FastDesc &f) {
    const uint32_t bstart = q.x, bend = q.x + MMK_INS_LEN(q.y) - 1u;
    const uint32_t P = q.w & 0x7fu, SUF = (q.w >> 7) & 0x3fu, schema = q.w >> 13;
    const uint32_t d1 = ndigits_u32(bstart), d2 = ndigits_u32(bend);
    const uint32_t word_off = ZWORD ? q.z : (uint32_t)a.sch[schema].word_off;
    f.sb1 = ZWORD ? base1 : base1 + (q.z & 0xffffu);
    f.sb2 = ZWORD ? base2 : base2 + (q.z >> 16);
    f.bstart = bstart;
    f.bend = bend;
    const uint32_t ws1 = bstart - 1u, ws2 = bend - (uint32_t)a.r2;  // 0-based window starts
    const uint32_t wsh = a.bpb == 2 ? 4u : 3u;                        // 16 or 8 bases per packed word
    f.wbase1 = word_off + (ws1 >= 15u ? (ws1 - 15u) >> wsh : 0u);
    f.wbase2 = word_off + (ws2 >= 15u ? (ws2 - 15u) >> wsh : 0u);
    f.mol = tile_m0 + MMK_INS_DM(q.y);
    f.pk = (P + d1 + 1u + d2 + SUF) | (P << 8) | (SUF << 16) | (d1 << 22) | (d2 << 26);
}

// Phase 1 — one read pair, one warp (half-warp 0: R1, reverse complement; half-warp 1: R2, forward):
// sequence chunks, then the 15 bytes after the sequence + the final newline (the stage already holds the quality
// character everywhere).  All 16-byte stores are aligned to the stream; what they spill on is rewritten later.
__device__ __forceinline__ void emit_pair_seq(const EmitArgs &a, const uint2 rdesc, uint32_t W, uint32_t smem, uint32_t lane) {
    const uint32_t half = lane >> 4, hl = lane & 15u;
    const uint32_t inl = half ? 0u : (uint32_t)a.bclen_inline;
    const uint32_t S = half ? (uint32_t)a.r2 : (uint32_t)a.r1;
    const uint32_t seq_off = rdesc.x & 0xfffffu;   // stage byte offset of the read's first base
    const int rel = (int)(rdesc.x >> 20);           // window start relative to the first loaded word
    const uint32_t seq_end = seq_off + S;
    const uint32_t rec_end = seq_end + 3u + S + inl + 1u;
    {
        const uint32_t c = (seq_off >> 4) + hl;
        const int k0 = (int)(16u * c) - (int)seq_off;                     // first base index of this chunk (>= -15)
        const int q = half ? rel + k0 : rel + (int)S - 16 - k0;          // in bases from the first loaded word
        const int j = q >> 4;
        const uint32_t sh = ((uint32_t)q & 15u) * 2u;
        uint32_t lo = __shfl_sync(0xffffffffu, W, (int)((lane & 16u) | ((uint32_t)j & 15u)));
        uint32_t hi = __shfl_sync(0xffffffffu, W, (int)((lane & 16u) | ((uint32_t)(j + 1) & 15u)));
        if (j < 0) lo = 0;
        uint32_t bits = __funnelshift_r(lo, hi, sh);
        // reverse complement: brev mirrors the 16 bases and also swaps the two bits of every base; the lookup is
        // indexed by the swapped code (00 A -> 'T', 10 = swapped C -> 'G', 01 = swapped G -> 'C', 11 T -> 'A')
        if (!half) bits = __brev(bits);
        const uint4 v = expand16(bits, half ? 0x54474341u /* "ACGT" */ : 0x41474354u /* "TCGA" */);
        if (16u * c < seq_end) sts128(smem + 16u * c, v);
    }
    __syncwarp();
    {
        const uint32_t addr = hl < 15 ? seq_end + hl : rec_end - 1u;
        const uint32_t val = (hl == 0 || hl == 2 || hl == 15) ? '\n' : (hl == 1 ? '+' : 'I');
        sts8(smem + addr, val);
    }
}

// 8 bases of a 4-bit word (BioSequences one-hot codes A=1 C=2 G=4 T=8) → 8 ASCII bytes; false if any of
// the 8 symbols is not one of A, C, G, T (IUPAC ambiguity or gap: the caller converts those one by one).
__device__ __forceinline__ bool expand8_onehot(uint32_t w, uint32_t &lo, uint32_t &hi) {
    // popcount per nibble must be 1 everywhere
    const uint32_t pc = w - ((w >> 1) & 0x77777777u) - ((w >> 2) & 0x33333333u) - ((w >> 3) & 0x11111111u);
    // one-hot nibble → index 0..3: (n >> 1) - (n >> 3) gives 0, 1, 2, 3 for 1, 2, 4, 8 (no borrow between nibbles)
    const uint32_t idx = ((w >> 1) & 0x77777777u) - ((w >> 3) & 0x11111111u);
    lo = __byte_perm(0x54474341u /* "ACGT" */, 0, idx);
    hi = __byte_perm(0x54474341u, 0, idx >> 16);
    return pc == 0x11111111u;
}
__device__ __forceinline__ uint32_t expand4_any(uint32_t nib16) {  // 4 symbols, any IUPAC code
    uint32_t r = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) r |= (uint32_t)(uint8_t)code4_to_ascii((nib16 >> (4 * b)) & 15u) << (8 * b);
    return r;
}

// Phase 1 for a 4-bit reference: same structure as emit_pair_seq; each lane of a half-warp holds two packed
// words (W0: word hl, W1: word hl + 16), a 16-base chunk is cut from three of them. Reverse complement of
// 8 one-hot symbols is a plain bit reversal of the word (reversing a nibble complements it).
__device__ __forceinline__ void emit_pair_seq4(const EmitArgs &a, const uint2 rdesc, uint32_t W0, uint32_t W1, uint32_t smem,
                                               uint32_t lane) {
    const uint32_t half = lane >> 4, hl = lane & 15u;
    const uint32_t inl = half ? 0u : (uint32_t)a.bclen_inline;
    const uint32_t S = half ? (uint32_t)a.r2 : (uint32_t)a.r1;
    const uint32_t seq_off = rdesc.x & 0xfffffu;
    const int rel = (int)(rdesc.x >> 20);  // window start relative to the first loaded word
    const uint32_t seq_end = seq_off + S;
    const uint32_t rec_end = seq_end + 3u + S + inl + 1u;
    {
        const uint32_t c = (seq_off >> 4) + hl;
        const int k0 = (int)(16u * c) - (int)seq_off;
        const int q = half ? rel + k0 : rel + (int)S - 16 - k0;  // first of 16 consecutive haplotype bases, from the first word
        const int j = q >> 3;
        const uint32_t sh = ((uint32_t)q & 7u) * 4u;
        uint32_t w[3];
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            const int jj = j + t;
            const uint32_t x0 = __shfl_sync(0xffffffffu, W0, (int)((lane & 16u) | ((uint32_t)jj & 15u)));
            const uint32_t x1 = __shfl_sync(0xffffffffu, W1, (int)((lane & 16u) | ((uint32_t)jj & 15u)));
            w[t] = jj < 0 ? 0u : (jj < 16 ? x0 : x1);
        }
        uint32_t b0 = __funnelshift_r(w[0], w[1], sh), b1 = __funnelshift_r(w[1], w[2], sh);  // bases 0-7, 8-15 of the run
        if (!half) {  // reverse complement: reverse the 16 symbols and complement each
            const uint32_t t0 = __brev(b1), t1 = __brev(b0);
            b0 = t0;
            b1 = t1;
        }
        uint4 v;
        const bool ok0 = expand8_onehot(b0, v.x, v.y), ok1 = expand8_onehot(b1, v.z, v.w);
        if (!(ok0 && ok1)) {  // rare: ambiguity codes (reads outside the read are masked by the later stores)
            v.x = expand4_any(b0);
            v.y = expand4_any(b0 >> 16);
            v.z = expand4_any(b1);
            v.w = expand4_any(b1 >> 16);
        }
        if (16u * c < seq_end) sts128(smem + 16u * c, v);
    }
    __syncwarp();
    {
        const uint32_t addr = hl < 15 ? seq_end + hl : rec_end - 1u;
        const uint32_t val = (hl == 0 || hl == 2 || hl == 15) ? '\n' : (hl == 1 ? '+' : 'I');
        sts8(smem + addr, val);
    }
}

// Phase 2 — header bytes, one thread per (pair, piece), straight into both stages. Records are
// ~93 words apart, so the 32 pairs of a warp hit 32 different banks.
__device__ __forceinline__ void emit_header_piece(const EmitArgs &a, const FastDesc &f, uint32_t smem, uint32_t role,
                                                  uint32_t rd_off) {
    const uint32_t P = FD_P(f.pk), SUF = FD_SUF(f.pk), d1 = FD_D1(f.pk), d2 = FD_D2(f.pk);
    const uint8_t *tx = a.text + (uint64_t)f.mol * a.text_stride;
    const uint32_t s1 = smem + f.sb1, s2 = smem + f.sb2;
    if (role == 0) {  // "@chrom:hap|start:end|"
        const uint32_t *tw = reinterpret_cast<const uint32_t *>(tx);
        for (uint32_t w = 0; 4u * w < P; ++w) {
            uint32_t x = __ldg(tw + w);
#pragma unroll
            for (uint32_t b = 0; b < 4; ++b) {
                if (4u * w + b < P) {
                    sts8(s1 + 4u * w + b, x);
                    sts8(s2 + 4u * w + b, x);
                }
                x >>= 8;
            }
        }
    } else if (role == 3) {  // format tail, read digit patched for R2; inline barcode for 10x R1
        const uint32_t o = P + d1 + 1u + d2;
        for (uint32_t i = 0; i < SUF; ++i) {
            const uint32_t ch = __ldg(tx + P + i);
            sts8(s1 + o + i, ch);
            sts8(s2 + o + i, i == rd_off ? (uint32_t)'2' : ch);
        }
        for (uint32_t i = 0; i < (uint32_t)a.bclen_inline; ++i) sts8(s1 + o + SUF + i, __ldg(tx + P + 1u + i));
    } else {  // decimal insert breakpoints and the ':' between them
        uint32_t v = role == 1 ? f.bstart : f.bend;
        const uint32_t D = role == 1 ? d1 : d2;
        const uint32_t o = role == 1 ? P : P + d1 + 1u;
        for (uint32_t k = D; k-- > 0;) {
            const uint32_t q = __umulhi(v, 0xCCCCCCCDu) >> 3;
            const uint32_t ch = '0' + (v - 10u * q);
            sts8(s1 + o + k, ch);
            sts8(s2 + o + k, ch);
            v = q;
        }
        if (role == 1) {
            sts8(s1 + P + d1, ':');
            sts8(s2 + P + d1, ':');
        }
    }
}

// Phase 2 as k_emit_w runs it — one warp per piece, one pair per lane, the pieces of about equal length: 0 the prefix,
// 1 both decimal numbers, 2 the first 16 bytes of the format tail, 3 the rest of it and the inline barcode.  Text is read as
// aligned 32-bit words and shifted into place (the molecule's text line is 16-byte aligned, its tail starts at byte P).
#ifndef MMK_HDR_WORDS
#define MMK_HDR_WORDS 0  // 1: header text goes into the stages as aligned 32-bit words (byte stores only up to the first and after the last whole word); 0: byte stores.  Measured equal (emit 94.4 / 95.2 ms against 94.9 / 95.0 ms, C5, same box): the header stores' bank conflicts are not what bounds the kernel
#endif
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
// n bytes of text from global memory into both stages, all three at any byte alignment.  The bytes are exactly [0, n) of each
// destination (the other header pieces are written by other warps at the same time: nothing may spill).  v_j = source bytes
// [4j, 4j + 4); a destination whose first aligned word starts h bytes in takes funnel(v_j, v_j+1, 8 h) as that word.
// Reads at most 11 bytes past the text wanted (the text array ends with a spare line).
__device__ __forceinline__ void copy_text_words(const uint8_t *src, uint32_t n, uint32_t d1, uint32_t d2) {
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 3u);
    const uint32_t *w = reinterpret_cast<const uint32_t *>(src - mis);
    uint32_t x0 = __ldg(w);
    uint32_t x1 = __ldg(++w);
    uint32_t vp = __funnelshift_r(x0, x1, 8u * mis);
    x0 = x1;
    const uint32_t h1 = (4u - (d1 & 3u)) & 3u, h2 = (4u - (d2 & 3u)) & 3u;
#pragma unroll
    for (uint32_t b = 0; b < 3; ++b) {
        if (b < h1 && b < n) sts8(d1 + b, vp >> (8u * b));
        if (b < h2 && b < n) sts8(d2 + b, vp >> (8u * b));
    }
    for (uint32_t o = 0; o < n; o += 4u) {
        x1 = __ldg(++w);
        const uint32_t v = __funnelshift_r(x0, x1, 8u * mis);
        x0 = x1;
        const uint32_t y1 = __funnelshift_r(vp, v, 8u * h1), y2 = __funnelshift_r(vp, v, 8u * h2);
        const uint32_t o1 = o + h1, o2 = o + h2;
        if (o1 + 4u <= n) sts32(d1 + o1, y1);
        else {
#pragma unroll
            for (uint32_t b = 0; b < 3; ++b)
                if (o1 + b < n) sts8(d1 + o1 + b, y1 >> (8u * b));
        }
        if (o2 + 4u <= n) sts32(d2 + o2, y2);
        else {
#pragma unroll
            for (uint32_t b = 0; b < 3; ++b)
                if (o2 + b < n) sts8(d2 + o2 + b, y2 >> (8u * b));
        }
        vp = v;
    }
}
__device__ __forceinline__ void copy_text_bytes(const uint8_t *src, uint32_t n, uint32_t d1, uint32_t d2) {
#if MMK_HDR_WORDS
    copy_text_words(src, n, d1, d2);
    return;
#endif
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 3u);
    const uint32_t *w = reinterpret_cast<const uint32_t *>(src - mis);
    uint32_t x0 = __ldg(w);
    for (uint32_t i = 0; i < n; i += 4u) {
        const uint32_t x1 = __ldg(++w);  // at most one word past the bytes wanted: inside the text array (it has a spare line)
        uint32_t v = __funnelshift_r(x0, x1, 8u * mis);
        x0 = x1;
#pragma unroll
        for (uint32_t b = 0; b < 4; ++b) {
            if (i + b < n) {
                sts8(d1 + i + b, v);
                sts8(d2 + i + b, v);
            }
            v >>= 8;
        }
    }
}
__device__ __forceinline__ void emit_header_piece_w(const EmitArgs &a, const FastDesc &f, uint32_t smem, uint32_t role, uint32_t rd_off) {
    const uint32_t P = FD_P(f.pk), SUF = FD_SUF(f.pk), d1 = FD_D1(f.pk), d2 = FD_D2(f.pk);
    const uint8_t *tx = a.text + (uint64_t)f.mol * a.text_stride;
    const uint32_t s1 = smem + f.sb1, s2 = smem + f.sb2;
    if (role == 0) {  // "@chrom:hap|start:end|"
        copy_text_bytes(tx, P, s1, s2);
    } else if (role == 1) {  // decimal insert breakpoints and the ':' between them
        uint32_t v = f.bend;
        uint32_t o = P + d1 + 1u;
        for (uint32_t k = d2; k-- > 0;) {
            const uint32_t q = __umulhi(v, 0xCCCCCCCDu) >> 3;
            const uint32_t ch = '0' + (v - 10u * q);
            sts8(s1 + o + k, ch);
            sts8(s2 + o + k, ch);
            v = q;
        }
        v = f.bstart;
        o = P;
        for (uint32_t k = d1; k-- > 0;) {
            const uint32_t q = __umulhi(v, 0xCCCCCCCDu) >> 3;
            const uint32_t ch = '0' + (v - 10u * q);
            sts8(s1 + o + k, ch);
            sts8(s2 + o + k, ch);
            v = q;
        }
        sts8(s1 + P + d1, ':');
        sts8(s2 + P + d1, ':');
    } else {  // format tail: bytes [0, 16) or [16, SUF); the read digit patched for R2; inline barcode for 10x R1
        const uint32_t o = P + d1 + 1u + d2;
        const uint32_t lo = role == 2 ? 0u : 16u, hi = role == 2 ? (SUF < 16u ? SUF : 16u) : SUF;
        if (hi > lo) copy_text_bytes(tx + P + lo, hi - lo, s1 + o + lo, s2 + o + lo);
        if (rd_off >= lo && rd_off < hi) sts8(s2 + o + rd_off, (uint32_t)'2');
        if (role == 3)
            for (uint32_t i = 0; i < (uint32_t)a.bclen_inline; ++i) sts8(s1 + o + SUF + i, __ldg(tx + P + 1u + i));
    }
}

// stage → global.  The 16-byte aligned interior of the tile goes out as ONE bulk asynchronous copy
// (TMA engine, cp.async.bulk shared::cta → global) issued by thread 0; the partial chunks at the two
// ends, which neighbouring tiles share, are written byte-wise.
__device__ __forceinline__ void bulk_store(uint8_t *gdst, uint32_t ssrc, uint32_t nbytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(nbytes) : "memory");
}
__device__ __forceinline__ void copy_out_edges(uint32_t stage, uint32_t A, uint32_t bytes, uint8_t *gbase, uint32_t tid) {
    const uint32_t end = A + bytes;
    const uint32_t c0 = (A + 15u) >> 4, c1 = end >> 4;  // full chunks [c0, c1)
    if (tid < 16) {  // head bytes [A, min(16 c0, end))
        const uint32_t b = A + tid;
        if (b < 16u * c0 && b < end) gbase[b] = (uint8_t)lds8(stage + b);
    } else if (tid < 32) {  // tail bytes [16 c1, end) when the tile reaches past its first chunk
        const uint32_t b = 16u * c1 + (tid - 16u);
        if (c1 >= c0 && b < end) gbase[b] = (uint8_t)lds8(stage + b);
    }
}
__device__ __forceinline__ void copy_out_bulk(uint32_t stage, uint32_t A, uint32_t bytes, uint8_t *gbase) {
    const uint32_t end = A + bytes;
    const uint32_t c0 = (A + 15u) >> 4, c1 = end >> 4;
    if (c1 > c0) bulk_store(gbase + 16u * c0, stage + 16u * c0, 16u * (c1 - c0));
}

template <int BPB, int THREADS>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS) k_emit(const __grid_constant__ EmitArgs a) {
    constexpr uint32_t EMIT_WARPS = THREADS / 32;
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr uint32_t MAXP = THREADS / 4;                   // pairs per tile this instantiation serves
    __shared__ uint4 s_fa[MAXP];                             // per pair, for the header phase: sb1, sb2, bstart, bend
    __shared__ uint2 s_fb[MAXP];                             // ... mol, pk
    __shared__ uint2 s_rd[2 * MAXP];                         // per read, for the sequence phase: seq_off | rel << 20, first word
    __shared__ uint32_t s_off1[MAXP + 1], s_off2[MAXP + 1];
    const uint32_t tile = a.tile0 + blockIdx.x;
    const uint64_t p0 = (uint64_t)tile * a.tile_pairs;  // tile_pairs is a power of two
    const uint32_t np = (uint32_t)((a.n_pairs - p0) < a.tile_pairs ? (a.n_pairs - p0) : a.tile_pairs);
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // every load of the prologue is independent of the others: one memory latency, not a chain
    const uint32_t tile_m0 = a.tile_mol0[tile];
    const uint64_t t1 = a.tile_off1[tile], t2 = a.tile_off2[tile];
    const uint32_t bytes1 = (uint32_t)(a.tile_off1[tile + 1] - t1), bytes2 = (uint32_t)(a.tile_off2[tile + 1] - t2);
    const bool fast = a.fast_ok && a.tile_flags[tile];
    // warm L2 for a tile that a later wave of CTAs will run (descriptors are read exactly once)
    if (threadIdx.x < 4 && blockIdx.x + 1184u < a.n_tiles)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(a.fdesc + (p0 + 1184ull * a.tile_pairs)) + 128u * threadIdx.x));
    {   // the quality character over both stages while the loads above are in flight: sequence chunks, separators and
        // headers overwrite everything that is not quality (writing only the quality lines, per read, measured 6 % slower)
        const uint4 fill = make_uint4(0x49494949u, 0x49494949u, 0x49494949u, 0x49494949u);
        uint4 *st = reinterpret_cast<uint4 *>(smem);
        const uint32_t n16 = a.stage_cap / 8u;  // 16-byte units of both stages
#pragma unroll 4
        for (uint32_t i = threadIdx.x; i < n16; i += THREADS) st[i] = fill;
    }
    const uint64_t g1 = t1 - a.chunk_off1, g2 = t2 - a.chunk_off2;
    const uint32_t A1 = (uint32_t)(g1 & 15u), A2 = (uint32_t)(g2 & 15u);
    if (!fast) {
        // general path: every record byte-wise, straight to global memory (warp per pair)
        PairDesc d;
        d.len1 = d.len2 = 0;
        if (threadIdx.x < np) decode_pair<true>(a, p0 + threadIdx.x, tile_m0, d);
        if (threadIdx.x < MAXP) {
            uint32_t i1 = d.len1, i2 = d.len2;
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t y1 = __shfl_up_sync(0xffffffffu, i1, o), y2 = __shfl_up_sync(0xffffffffu, i2, o);
                if (lane >= o) {
                    i1 += y1;
                    i2 += y2;
                }
            }
            s_off1[threadIdx.x + 1] = i1;
            s_off2[threadIdx.x + 1] = i2;
            if (threadIdx.x == 0) s_off1[0] = s_off2[0] = 0;
        }
        __syncthreads();
        const uint32_t base1 = s_off1[32], base2 = s_off2[32];
        __syncthreads();
        if (threadIdx.x < np) {  // exclusive offsets of the pair's two records inside the tile
            s_off1[threadIdx.x + 1] -= d.len1 - (threadIdx.x >= 32 ? base1 : 0u);
            s_off2[threadIdx.x + 1] -= d.len2 - (threadIdx.x >= 32 ? base2 : 0u);
        }
        __syncthreads();
        uint8_t *o1 = a.out1 + g1, *o2 = a.out2 + g2;
        for (uint32_t t = warp; t < np; t += EMIT_WARPS) {
            PairDesc e;
            decode_pair<true>(a, p0 + t, tile_m0, e);
            write_record_generic<BPB>(a, e, 1, o1 + s_off1[t + 1], lane);
            write_record_generic<BPB>(a, e, 2, o2 + s_off2[t + 1], lane);
        }
        return;
    }
    // stage1 = smem[0, stage_cap), stage2 = smem[stage_cap, 2 stage_cap): record offsets index smem directly
    constexpr bool ZWORD = THREADS == 128;  // 32-pair tiles: one warp holds the tile's pairs and scans their record lengths
    if (ZWORD ? threadIdx.x < 32u : threadIdx.x < np) {
        const bool have = threadIdx.x < np;
        FastDesc f;
        make_fast_desc<ZWORD>(a, have ? a.fdesc[p0 + threadIdx.x] : make_uint4(1u, 1u, 0u, 0u), tile_m0, A1, a.stage_cap + A2, f);
        if (ZWORD) {  // record offsets in the tile: exclusive scan of (len1 | len2 << 16); both totals are < 2^16 (tile flag)
            const uint32_t hdr = FD_HDR(f.pk);
            const uint32_t lens = have ? ((hdr + 2u * (uint32_t)a.bclen_inline + 2u * (uint32_t)a.r1 + 4u) | ((hdr + 2u * (uint32_t)a.r2 + 4u) << 16)) : 0u;
            uint32_t inc = lens;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= (uint32_t)o) inc += y;
            }
            f.sb1 += (inc - lens) & 0xffffu;
            f.sb2 += (inc - lens) >> 16;
        }
        if (have) {
            s_fa[threadIdx.x] = make_uint4(f.sb1, f.sb2, f.bstart, f.bend);
            s_fb[threadIdx.x] = make_uint2(f.mol, f.pk);
            {   // what phase 1 needs of each read, so that its lanes load 8 bytes and select nothing
                constexpr uint32_t WM = BPB == 2 ? 15u : 7u;
                const uint32_t ws1 = f.bstart - 1u, ws2 = f.bend - (uint32_t)a.r2;
                const uint32_t rel1 = ws1 >= 15u ? ((ws1 - 15u) & WM) + 15u : ws1, rel2 = ws2 >= 15u ? ((ws2 - 15u) & WM) + 15u : ws2;
                s_rd[2 * threadIdx.x] = make_uint2((f.sb1 + FD_HDR(f.pk) + (uint32_t)a.bclen_inline) | (rel1 << 20), f.wbase1);
                s_rd[2 * threadIdx.x + 1] = make_uint2((f.sb2 + FD_HDR(f.pk)) | (rel2 << 20), f.wbase2);
            }
            // start the loads phases 1 and 2 will need
            prefetch_l1(a.hap + f.wbase1);
            prefetch_l1(a.hap + f.wbase1 + (BPB == 2 ? 15 : 31));
            prefetch_l1(a.hap + f.wbase2);
            prefetch_l1(a.hap + f.wbase2 + (BPB == 2 ? 15 : 31));
            prefetch_l1(a.text + (uint64_t)f.mol * a.text_stride);
        }
    }
    const uint32_t sm = (uint32_t)__cvta_generic_to_shared(smem);
    __syncthreads();
    {   // the reference words of pair t + EMIT_WARPS are loaded while pair t is expanded
        const uint32_t half = lane >> 4, hl = lane & 15u;
        const uint32_t *hap_lane = a.hap + hl;
        uint2 rd = s_rd[2 * (warp < np ? warp : 0) + half];
        uint32_t W0 = __ldg(hap_lane + rd.y), W1 = BPB == 4 ? __ldg(hap_lane + rd.y + 16) : 0u;
        for (uint32_t t = warp; t < np; t += EMIT_WARPS) {
            const uint32_t tn = t + EMIT_WARPS < np ? t + EMIT_WARPS : t;
            const uint2 rn = s_rd[2 * tn + half];
            const uint32_t W0n = __ldg(hap_lane + rn.y), W1n = BPB == 4 ? __ldg(hap_lane + rn.y + 16) : 0u;
            if (BPB == 2) emit_pair_seq(a, rd, W0, sm, lane);
            else emit_pair_seq4(a, rd, W0, W1, sm, lane);
            rd = rn;
            W0 = W0n;
            W1 = W1n;
        }
    }
    __syncthreads();
    {
        const uint32_t role = threadIdx.x / (THREADS / 4), t = threadIdx.x % (THREADS / 4);  // 4 pieces x up to THREADS/4 pairs
        if (t < np) {
            const FastDesc f = load_fast_desc(s_fa, s_fb, t);
            const uint32_t rd_off = (a.fq_format == MMK_FMT_HAPLOTAGGING || a.fq_format == MMK_FMT_STANDARD) ? 1u : FD_SUF(f.pk) - 13u;
            emit_header_piece(a, f, sm, role, rd_off);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes → visible to the bulk copy
    __syncthreads();
    uint8_t *gb1 = a.out1 + (g1 - A1), *gb2 = a.out2 + (g2 - A2);
    if (threadIdx.x == 0) {
        copy_out_bulk(sm, A1, bytes1, gb1);
        copy_out_bulk(sm + a.stage_cap, A2, bytes2, gb2);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    if (warp == 1) copy_out_edges(sm, A1, bytes1, gb1, lane);
    if (warp == 2) copy_out_edges(sm + a.stage_cap, A2, bytes2, gb2, lane);
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // stage must outlive the copy
}


// ------------------------------------------------------------------------------------------
// k_emit_w — the fast path of 32-pair tiles with the stage filled by the copy engine (the default emit kernel)
// ------------------------------------------------------------------------------------------
// Same phases as k_emit, three differences:
//   * the quality character arrives in both stages by ONE bulk asynchronous copy (TMA: cp.async.bulk global -> shared from
//     an L2-resident buffer of 'I', completion on an mbarrier), issued by thread 0 before anything else: no fill loop, a
//     third of the shared-memory store wavefronts gone;
//   * a warp has the reference words of all its 8 pairs in flight before its first pass (2-bit references);
//   * ERR: the optional sequencing-error model rewrites the bases and qualities of every read (err_read_warp) — a template
//     variant, so the plain kernel keeps its instruction count.
// Measured (profiles/r02_summary.md): the same time per step as k_emit; variants with autonomous warps, per-warp copy-out,
// two pairs per pass, look-ahead L2 prefetch and persistent CTAs with two stage sets were all equal or slower.
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t sdst, const void *gsrc, uint32_t nbytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sdst), "l"(gsrc), "r"(nbytes),
                 "r"(mbar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MMK_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MMK_DONE_%=;\n"
        "bra MMK_WAIT_%=;\n"
        "MMK_DONE_%=:\n"
        "}\n" ::"r"(mbar),
        "r"(parity)
        : "memory");
}

template <int BPB, bool ERR>
__global__ void __launch_bounds__(128, 8) k_emit_w(const __grid_constant__ EmitArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr bool HOT = BPB == 2 && !ERR;  // the plain 2-bit kernel: per-read records in the form its sequence loop consumes
    __shared__ uint4 s_fa[32];
    __shared__ uint2 s_fb[32];
    __shared__ uint2 s_rd[HOT ? 1 : 64];
    __shared__ uint4 s_rq[HOT ? 64 : 1];  // HOT: first word, address of the read's first 16-byte chunk, base shift, chunk limit
    __shared__ uint2 s_er[ERR ? 32 : 1];  // error model: first arena word and length of the pair's haplotype
    __shared__ uint32_t s_off1[33], s_off2[33];
    __shared__ __align__(8) unsigned long long s_mbar;
    const uint32_t sm = (uint32_t)__cvta_generic_to_shared(smem), mb = (uint32_t)__cvta_generic_to_shared(&s_mbar);
    // the next chunk's kernel (another ring slot, the same placement) may fill the SMs this grid's last wave leaves idle
    asm volatile("griddepcontrol.launch_dependents;");
    if (threadIdx.x == 0) {  // the quality character over both stages, by the copy engine
        mbar_init(mb, 1u);
        mbar_expect_tx(mb, 2u * a.stage_cap);
        bulk_load(sm, a.qfill, 2u * a.stage_cap, mb);
    }
    const uint32_t tile = a.tile0 + blockIdx.x;
    const uint64_t p0 = (uint64_t)tile * 32u;
    const uint32_t np = (uint32_t)((a.n_pairs - p0) < 32u ? (a.n_pairs - p0) : 32u);
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // every load of the prologue is independent of the others: one memory latency, not a chain
    const bool fast = a.fast_ok && a.tile_flags[tile];
    const uint64_t t1 = a.tile_off1[tile], t2 = a.tile_off2[tile];
    const uint32_t bytes1 = (uint32_t)(a.tile_off1[tile + 1] - t1), bytes2 = (uint32_t)(a.tile_off2[tile + 1] - t2);
    const uint64_t g1 = t1 - a.chunk_off1, g2 = t2 - a.chunk_off2;
    const uint32_t A1 = (uint32_t)(g1 & 15u), A2 = (uint32_t)(g2 & 15u);
    uint32_t W[8];  // HOT: the first packed words of the warp's 8 pairs (lane hl of a half-warp: word hl of the read's window)
    if (HOT && MMK_EMIT_EARLY && fast) {
        // every warp starts the loads of its own pairs' windows from the packed descriptors, without waiting for warp 0's pass
        // over the tile: lanes 0-7 take the R1 windows of pairs warp + 4 i, lanes 8-15 the R2 windows, a shuffle hands them out
        uint32_t wb = 0u;
        const uint32_t t = warp + 4u * (lane & 7u);
        if (lane < 16u && t < np) {
            const uint4 qe = a.fdesc[p0 + t];
            const uint32_t ws = lane < 8u ? qe.x - 1u : qe.x + MMK_INS_LEN(qe.y) - 1u - (uint32_t)a.r2;  // 0-based window start
            wb = qe.z + (ws >> 4);
        }
        const uint32_t hl = lane & 15u, half = lane >> 4;
        const bool ldw = hl <= (((half ? (uint32_t)a.r2 : (uint32_t)a.r1) + 14u) >> 4);  // words 0 .. (15 + S - 1) / 16 hold the window
#pragma unroll
        for (uint32_t i = 0; i < 8; ++i) {
            const uint32_t b = __shfl_sync(0xffffffffu, wb, (int)(i + 8u * half));
            W[i] = ldw ? __ldg(a.hap + (b + hl)) : 0u;
        }
    }
    if (warp == 0) {  // descriptors of the whole tile
        const uint32_t tile_m0 = a.tile_mol0[tile];
        const bool have = lane < np;
        const uint4 q = have ? a.fdesc[p0 + lane] : make_uint4(1u, 1u, 0u, 0u);
        if (lane < 8 && blockIdx.x + EMIT_AHEAD_DESC < a.n_tiles) {  // warm L2 for a tile that a later wave of CTAs will run: its 512 bytes
            // of descriptors, and the lines of the four per-tile arrays (the 16 or more tiles that share such a line start within
            // microseconds of each other: without this every one of them waits for the same DRAM access)
            const uint32_t tn = tile + EMIT_AHEAD_DESC;
            const char *pf = lane < 4   ? reinterpret_cast<const char *>(a.fdesc + (p0 + (uint64_t)EMIT_AHEAD_DESC * 32u)) + 128u * lane
                             : lane == 4 ? reinterpret_cast<const char *>(a.tile_off1 + tn + 1)
                             : lane == 5 ? reinterpret_cast<const char *>(a.tile_off2 + tn + 1)
                             : lane == 6 ? reinterpret_cast<const char *>(a.tile_flags + tn)
                                         : reinterpret_cast<const char *>(a.tile_mol0 + tn);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
        }
        if (fast) {
            // stage1 = smem[0, stage_cap), stage2 = smem[stage_cap, 2 stage_cap): record offsets index smem directly
            FastDesc f;
            make_fast_desc<true>(a, q, tile_m0, A1, a.stage_cap + A2, f);
            {   // record offsets in the tile: exclusive scan of (len1 | len2 << 16); both totals are < 2^16 (tile flag)
                const uint32_t hdr = FD_HDR(f.pk);
                const uint32_t lens = have ? ((hdr + 2u * (uint32_t)a.bclen_inline + 2u * (uint32_t)a.r1 + 4u) | ((hdr + 2u * (uint32_t)a.r2 + 4u) << 16)) : 0u;
                uint32_t inc = lens;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= (uint32_t)o) inc += y;
                }
                f.sb1 += (inc - lens) & 0xffffu;
                f.sb2 += (inc - lens) >> 16;
            }
            if (have) {
                s_fa[lane] = make_uint4(f.sb1, f.sb2, f.bstart, f.bend);
                s_fb[lane] = make_uint2(f.mol, f.pk);
                constexpr uint32_t WM = BPB == 2 ? 15u : 7u;
                const uint32_t ws1 = f.bstart - 1u, ws2 = f.bend - (uint32_t)a.r2;
                const uint32_t rel1 = ws1 >= 15u ? ((ws1 - 15u) & WM) + 15u : ws1, rel2 = ws2 >= 15u ? ((ws2 - 15u) & WM) + 15u : ws2;
                const uint32_t so1 = f.sb1 + FD_HDR(f.pk) + (uint32_t)a.bclen_inline, so2 = f.sb2 + FD_HDR(f.pk);  // stage offsets of the first bases
                if (HOT) {
                    // chunk c of a read (16 bytes of the stage, lane hl = c) starts at base k0 = 16 hl - (so & 15) of the read and takes
                    // the 16 haplotype bases from q = rel + k0 (forward) / rel + S - 16 - k0 (reverse complement) past the first loaded word:
                    // z = q at hl = 0, w = the first 16 hl that lies past the sequence
                    // The first loaded word is the one holding the window's first base (no lead: chunks that start before the
                    // sequence shift in another lane's word, for bytes that are rewritten anyway), and a half-warp loads only the
                    // words a window can reach: 11 instead of 16 for 150-base reads, 2.2 instead of 2.9 32-byte sectors per read.
                    f.wbase1 = q.z + (ws1 >> 4);
                    f.wbase2 = q.z + (ws2 >> 4);
                    const uint32_t r1_ = ws1 & 15u, r2_ = ws2 & 15u;
                    s_rq[2 * lane] = make_uint4(f.wbase1, sm + (so1 & ~15u), r1_ + (uint32_t)a.r1 - 16u + (so1 & 15u), (so1 & 15u) + (uint32_t)a.r1);
                    s_rq[2 * lane + 1] = make_uint4(f.wbase2, sm + (so2 & ~15u), r2_ - (so2 & 15u), (so2 & 15u) + (uint32_t)a.r2);
                } else {
                    s_rd[2 * lane] = make_uint2(so1 | (rel1 << 20), f.wbase1);
                    s_rd[2 * lane + 1] = make_uint2(so2 | (rel2 << 20), f.wbase2);
                }
                if (ERR) s_er[lane] = make_uint2(q.z, a.sch[q.w >> 13].len);
            } else if (HOT) {  // a short last tile: the word loads of the missing pairs stay inside the arena
                s_rq[2 * lane] = s_rq[2 * lane + 1] = make_uint4(0u, 0u, 0u, 0u);
            }
            {   // start the loads the next phases will need — one request per distinct 128-byte line: the last word of a window
                // is often in the line of the first, and neighbouring pairs share their molecule's header text
                const uint32_t WL1 = HOT ? ((uint32_t)a.r1 + 14u) >> 4 : (BPB == 2 ? 15u : 31u);  // last word a half-warp loads
                const uint32_t WL2 = HOT ? ((uint32_t)a.r2 + 14u) >> 4 : (BPB == 2 ? 15u : 31u);
                const uint32_t mol_prev = __shfl_up_sync(0xffffffffu, f.mol, 1);
                if (have) {
                    if (!(HOT && MMK_EMIT_EARLY)) {
                        prefetch_l1(a.hap + f.wbase1);
                        if ((f.wbase1 ^ (f.wbase1 + WL1)) >> 5) prefetch_l1(a.hap + f.wbase1 + WL1);
                        prefetch_l1(a.hap + f.wbase2);
                        if ((f.wbase2 ^ (f.wbase2 + WL2)) >> 5) prefetch_l1(a.hap + f.wbase2 + WL2);
                    }
                    if (lane == 0 || f.mol != mol_prev) prefetch_l1(a.text + (uint64_t)f.mol * a.text_stride);
                }
            }
        }
    }
    if (MMK_EMIT_WPF && BPB == 2 && warp == 1 && blockIdx.x + EMIT_AHEAD_WIN < a.n_tiles) {
        // while warp 0 unpacks this tile: the reference windows of the tile one wave ahead, into L2 (its descriptors were brought
        // there by the tile two waves back), so that the CTA that emits it waits an L2 hit, not a DRAM access, for its words
        const uint32_t tn = tile + EMIT_AHEAD_WIN;
        const uint64_t pn = (uint64_t)tn * 32u + lane;
        if (a.tile_flags[tn] && pn < a.n_pairs) {  // fast tile: every window lies inside its haplotype
            const uint4 qn = a.fdesc[pn];
            const uint32_t ws1 = qn.x - 1u, ws2 = qn.x + MMK_INS_LEN(qn.y) - 1u - (uint32_t)a.r2;
            const uint32_t *w1 = a.hap + (qn.z + (ws1 >> 4)), *w2 = a.hap + (qn.z + (ws2 >> 4));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(w1));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(w1 + (((uint32_t)a.r1 + 14u) >> 4)));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(w2));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(w2 + (((uint32_t)a.r2 + 14u) >> 4)));
        }
    }
    __syncthreads();  // descriptors and the mbarrier are published
    if (!fast) {
        // general path: every record byte-wise, straight to global memory (warp per pair)
        const uint32_t tile_m0 = a.tile_mol0[tile];
        PairDesc d;
        d.len1 = d.len2 = 0;
        if (threadIdx.x < np) decode_pair<true>(a, p0 + threadIdx.x, tile_m0, d);
        if (threadIdx.x < 32) {
            uint32_t i1 = d.len1, i2 = d.len2;
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t y1 = __shfl_up_sync(0xffffffffu, i1, o), y2 = __shfl_up_sync(0xffffffffu, i2, o);
                if (lane >= o) {
                    i1 += y1;
                    i2 += y2;
                }
            }
            s_off1[threadIdx.x + 1] = i1 - d.len1;  // exclusive offsets of the pair's two records inside the tile
            s_off2[threadIdx.x + 1] = i2 - d.len2;
        }
        __syncthreads();
        uint8_t *o1 = a.out1 + g1, *o2 = a.out2 + g2;
        for (uint32_t t = warp; t < np; t += 4) {
            PairDesc e;
            decode_pair<true>(a, p0 + t, tile_m0, e);
            write_record_generic<BPB>(a, e, 1, o1 + s_off1[t + 1], lane);
            write_record_generic<BPB>(a, e, 2, o2 + s_off2[t + 1], lane);
        }
        mbar_wait(mb, 0u);  // the stage must outlive the copy into it
        return;
    }
    {   // phase 1: warp per pair (t = warp, warp + 4, ...), half-warp per read
        const uint32_t half = lane >> 4, hl = lane & 15u;
        const uint32_t *hap_lane = a.hap + hl;
        if constexpr (HOT) {  // the words of all the warp's pairs in flight at once, then the passes
            // what depends on the lane alone, kept in registers (the asm statements stop the compiler from recomputing it per pass)
            uint32_t hl16 = 16u * hl, qs = half ? hl16 : 0u - hl16;
            // R2 (forward): "ACGT"[code], bytes in base order.  R1 (reverse complement): "TGCA"[code], and the 16 bases leave in
            // reverse order — the two halves of the word swapped (rsh) and the interleaving byte selectors mirrored (selA, selB)
            uint32_t lut = half ? 0x54474341u : 0x41434754u, rsh = half ? 0u : 16u;
            uint32_t selA = half ? 0x5140u : 0x2637u, selB = half ? 0x7362u : 0x0415u;
            // the 15 bytes after the sequence ("\n+\n" and the first quality characters) and the record's final newline
            uint32_t tval = (hl == 0u || hl == 2u || hl == 15u) ? (uint32_t)'\n' : (hl == 1u ? (uint32_t)'+' : (uint32_t)'I');
            uint32_t tdelta = hl < 15u ? hl : 3u + (half ? (uint32_t)a.r2 : (uint32_t)a.r1 + (uint32_t)a.bclen_inline);
            asm volatile("" : "+r"(hl16), "+r"(qs), "+r"(lut), "+r"(rsh), "+r"(selA), "+r"(selB), "+r"(tval), "+r"(tdelta));
            const uint4 *rq = s_rq + 2u * warp + half;  // pair warp + 4 i: rq[8 i]
            const bool ldw = hl <= (((half ? (uint32_t)a.r2 : (uint32_t)a.r1) + 14u) >> 4);  // words 0 .. (15 + S - 1) / 16 hold the window
            if (!MMK_EMIT_EARLY) {
#pragma unroll
                for (uint32_t i = 0; i < 8; ++i) W[i] = ldw ? __ldg(a.hap + (rq[8u * i].x + hl)) : 0u;
            }
            mbar_wait(mb, 0u);  // the quality fill has landed
            // 16 bytes of sequence text for this lane's chunk of the read described by R
            auto chunk = [&](const uint4 R, const uint32_t Wi) -> uint4 {
                const int q = (int)(R.z + qs);
                // words j and j + 1 of the half-warp (width 16: the source lane is taken modulo 16 inside the half).  Lanes whose q is
                // negative or past the loaded words hold bases outside the read: what they write is rewritten later
                const uint32_t lo = __shfl_sync(0xffffffffu, Wi, q >> 4, 16);
                const uint32_t hi = __shfl_sync(0xffffffffu, Wi, (q >> 4) + 1, 16);
                uint32_t bits = __funnelshift_r(lo, hi, 2u * (uint32_t)q);  // the shift is taken modulo 32: 2 (q mod 16)
                bits = __funnelshift_l(bits, bits, rsh);
                uint32_t ev = bits & 0x33333333u, od = (bits >> 2) & 0x33333333u;  // PRMT selectors: one base per nibble
                asm volatile("" : "+r"(ev), "+r"(od));
                const uint32_t e_lo = prmt(lut, 0u, ev), e_hi = prmt(lut, 0u, ev >> 16);
                const uint32_t o_lo = prmt(lut, 0u, od), o_hi = prmt(lut, 0u, od >> 16);
                return make_uint4(prmt(e_lo, o_lo, selA), prmt(e_lo, o_lo, selB), prmt(e_hi, o_hi, selA), prmt(e_hi, o_hi, selB));
            };
            if (np == 32u) {  // MMK_EMIT_ILP pairs at a time: their dependent chains (LDS -> SHFL -> PRMT -> STS) interleave
#pragma unroll
                for (uint32_t i = 0; i < 8; i += MMK_EMIT_ILP) {
                    uint4 R[MMK_EMIT_ILP], v[MMK_EMIT_ILP];
#pragma unroll
                    for (uint32_t g = 0; g < MMK_EMIT_ILP; ++g) R[g] = rq[8u * (i + g)];
#pragma unroll
                    for (uint32_t g = 0; g < MMK_EMIT_ILP; ++g) v[g] = chunk(R[g], W[i + g]);
#pragma unroll
                    for (uint32_t g = 0; g < MMK_EMIT_ILP; ++g)
                        if (hl16 < R[g].w) sts128(R[g].y + hl16, v[g]);
                    __syncwarp();
#pragma unroll
                    for (uint32_t g = 0; g < MMK_EMIT_ILP; ++g) sts8(R[g].y + R[g].w + tdelta, tval);
                }
            } else {
#pragma unroll
                for (uint32_t i = 0; i < 8; ++i)
                    if (warp + 4u * i < np) {
                        const uint4 R = rq[8u * i];
                        const uint4 v = chunk(R, W[i]);
                        if (hl16 < R.w) sts128(R.y + hl16, v);
                        __syncwarp();
                        sts8(R.y + R.w + tdelta, tval);
                    }
            }
        } else {  // the words of the next pair are loaded while this one is expanded
            uint2 rd = s_rd[2 * (warp < np ? warp : 0) + half];
            uint32_t W0 = __ldg(hap_lane + rd.y), W1 = BPB == 4 ? __ldg(hap_lane + rd.y + 16) : 0u;
            mbar_wait(mb, 0u);
            for (uint32_t t = warp; t < np; t += 4) {
                const uint32_t tn = t + 4u < np ? t + 4u : t;
                const uint2 rn = s_rd[2 * tn + half];
                const uint32_t W0n = __ldg(hap_lane + rn.y), W1n = BPB == 4 ? __ldg(hap_lane + rn.y + 16) : 0u;
                if (BPB == 2) emit_pair_seq(a, rd, W0, sm, lane);
                else emit_pair_seq4(a, rd, W0, W1, sm, lane);
                if (ERR) {  // bases and qualities again, through the error model (every pair of a fast tile has nominal reads)
                    __syncwarp();
                    const uint4 fa = s_fa[t];
                    const uint2 er = s_er[t];
                    const uint32_t so1 = sm + (s_rd[2 * t].x & 0xfffffu), so2 = sm + (s_rd[2 * t + 1].x & 0xfffffu);
                    const uint32_t qo1 = so1 + (uint32_t)a.r1 + 3u + (uint32_t)a.bclen_inline, qo2 = so2 + (uint32_t)a.r2 + 3u;
                    err_read_warp<BPB>(a, a.hap + er.x, er.y, 1u, (int64_t)fa.z + a.r1 - 1, a.pair_base + p0 + t, lane, [&](uint32_t c, uint32_t ch, uint32_t qc) {
                        sts8(so1 + c, ch);
                        sts8(qo1 + c, qc);
                    });
                    err_read_warp<BPB>(a, a.hap + er.x, er.y, 2u, (int64_t)fa.w - a.r2 + 1, a.pair_base + p0 + t, lane, [&](uint32_t c, uint32_t ch, uint32_t qc) {
                        sts8(so2 + c, ch);
                        sts8(qo2 + c, qc);
                    });
                }
                rd = rn;
                W0 = W0n;
                W1 = W1n;
            }
        }
    }
    __syncthreads();
    if (lane < np) {  // phase 2: one header piece per warp, one pair per lane
        const FastDesc f = load_fast_desc(s_fa, s_fb, lane);
        const uint32_t rd_off = (a.fq_format == MMK_FMT_HAPLOTAGGING || a.fq_format == MMK_FMT_STANDARD) ? 1u : FD_SUF(f.pk) - 13u;
        emit_header_piece_w(a, f, sm, warp, rd_off);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes → visible to the bulk copy
    __syncthreads();
    uint8_t *gb1 = a.out1 + (g1 - A1), *gb2 = a.out2 + (g2 - A2);
    if (threadIdx.x == 0) {
        copy_out_bulk(sm, A1, bytes1, gb1);
        copy_out_bulk(sm + a.stage_cap, A2, bytes2, gb2);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    if (warp == 1) copy_out_edges(sm, A1, bytes1, gb1, lane);
    if (warp == 2) copy_out_edges(sm + a.stage_cap, A2, bytes2, gb2, lane);
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // stage must outlive the copy
}

}  // namespace mmkdev
