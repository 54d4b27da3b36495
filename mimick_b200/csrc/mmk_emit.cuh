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
//   4. the header (built once per pair in a per-warp scratch, copied with a funnel shift to the
//      unaligned record start) and the "\n+\n" / first quality bytes are written AFTER the
//      sequence, so write order — not masks — resolves the overlaps.
// Anything else (4-bit reference, literal circular wrap, reads outside 16..192, huge headers,
// force_generic) goes through write_record_generic, straight to global memory.
#pragma once
#include "mmk_internal.h"
#include "mmk_sampling.cuh"

namespace mmkdev {

constexpr int EMIT_THREADS = 256;
constexpr int EMIT_WARPS = EMIT_THREADS / 32;
constexpr int EMIT_MAX_TP = 64;
constexpr int EMIT_SCR = 192;  // per-warp header scratch bytes

struct PairDesc {
    uint64_t word_off;   // packed haplotype start (32-bit words)
    uint32_t mol;        // shard-relative molecule index
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

__device__ __forceinline__ void decode_pair(const EmitArgs &a, uint64_t p, uint32_t tile_m0, PairDesc &d) {
    uint2 in = a.ins[p];
    uint32_t m = tile_m0 + MMK_INS_DM(in.y);
    MolRec mr = a.mol[m];
    const SchemaDev &S = a.sch[mr.schema];
    d.word_off = S.word_off;
    d.n = S.len;
    d.mol = m;
    d.bstart = (uint32_t)(mr.start + in.x);
    d.bend = d.bstart + MMK_INS_LEN(in.y) - 1;
    d.d1 = (uint8_t)ndigits_u32(d.bstart);
    d.d2 = (uint8_t)ndigits_u32(d.bend);
    d.P = mr.prefix_len;
    d.SUF = mr.suffix_len;
    d.hdr = (uint32_t)d.P + d.d1 + 1 + d.d2 + d.SUF;
    int64_t a2 = (int64_t)d.bend - a.r2 + 1;
    d.sq1 = window_len(d.bstart, (int64_t)d.bstart + a.r1 - 1, d.n, a.strict);
    d.sq2 = window_len(a2, d.bend, d.n, a.strict);
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
}

// ------------------------------------------------------------------------------------------
// tile sums: record lengths of every tile → byte offsets (after a scan)
// ------------------------------------------------------------------------------------------
__global__ void k_tile_sums(EmitArgs a, uint32_t n_tiles, uint32_t *sum1, uint32_t *sum2, unsigned long long *gather) {
    __shared__ uint32_t s_acc[3];
    uint32_t tile = blockIdx.x;
    if (tile >= n_tiles) return;
    uint64_t p0 = (uint64_t)tile * a.tile_pairs;
    uint32_t np = (uint32_t)((a.n_pairs - p0) < a.tile_pairs ? (a.n_pairs - p0) : a.tile_pairs);
    if (threadIdx.x == 0) s_acc[0] = s_acc[1] = s_acc[2] = 0;
    __syncthreads();
    uint32_t l1 = 0, l2 = 0, gb = 0;
    if (threadIdx.x < np) {
        PairDesc d;
        decode_pair(a, p0 + threadIdx.x, a.tile_mol0[tile], d);
        l1 = d.len1;
        l2 = d.len2;
        gb = (d.sq1 * a.bpb + 7) / 8 + (d.sq2 * a.bpb + 7) / 8;
    }
    for (int o = 16; o; o >>= 1) {
        l1 += __shfl_down_sync(0xffffffffu, l1, o);
        l2 += __shfl_down_sync(0xffffffffu, l2, o);
        gb += __shfl_down_sync(0xffffffffu, gb, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&s_acc[0], l1);
        atomicAdd(&s_acc[1], l2);
        atomicAdd(&s_acc[2], gb);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        sum1[tile] = s_acc[0];
        sum2[tile] = s_acc[1];
        atomicAdd(gather, (unsigned long long)s_acc[2]);
    }
}

// ------------------------------------------------------------------------------------------
// fast path pieces
// ------------------------------------------------------------------------------------------
// ceil(2^64 / 10^(9-k)) for k = 0..8; k = 9 (divide by 1) is special-cased
__constant__ unsigned long long c_magic10[10] = {18446744074ull, 184467440738ull, 1844674407371ull, 18446744073710ull,
                                                 184467440737096ull, 1844674407370956ull, 18446744073709552ull,
                                                 184467440737095517ull, 1844674407370955162ull, 0ull};

// 16 bases (32 bits, base i in bits 2i..2i+1) → 16 ASCII bytes through a 4-entry byte LUT
__device__ __forceinline__ uint4 expand16(uint32_t bits, uint32_t lut) {
    uint4 v;
    uint32_t b, t;
    b = bits & 0xffu;
    t = (b | (b << 4)) & 0x0f0fu;
    t = (t | (t << 2)) & 0x3333u;
    v.x = __byte_perm(lut, 0, t);
    b = (bits >> 8) & 0xffu;
    t = (b | (b << 4)) & 0x0f0fu;
    t = (t | (t << 2)) & 0x3333u;
    v.y = __byte_perm(lut, 0, t);
    b = (bits >> 16) & 0xffu;
    t = (b | (b << 4)) & 0x0f0fu;
    t = (t | (t << 2)) & 0x3333u;
    v.z = __byte_perm(lut, 0, t);
    b = bits >> 24;
    t = (b | (b << 4)) & 0x0f0fu;
    t = (t | (t << 2)) & 0x3333u;
    v.w = __byte_perm(lut, 0, t);
    return v;
}

// Per-pair parameters of the fast path, computed once by the pair's decode thread and read by the
// emitting warp with three 16-byte shared-memory loads.
struct __align__(16) FastDesc {
    uint32_t sb1, sb2;       // stage byte offsets of the R1 / R2 record
    uint32_t bstart, bend;
    uint64_t wbase1, wbase2; // first packed word each half-warp loads (haplotype base + window word)
    uint32_t mol;
    uint16_t hdr;            // header length incl. '\n'
    uint8_t P, SUF;
    uint8_t d1, d2;
    uint8_t rel1, rel2;      // window start relative to the first loaded word, in bases (0..30)
    uint32_t pad;
};
static_assert(sizeof(FastDesc) == 48, "FastDesc layout");

__device__ __forceinline__ void make_fast_desc(const EmitArgs &a, const PairDesc &d, uint32_t sb1, uint32_t sb2, FastDesc &f) {
    f.sb1 = sb1;
    f.sb2 = sb2;
    f.bstart = d.bstart;
    f.bend = d.bend;
    const uint32_t ws1 = d.bstart - 1u, ws2 = d.bend - (uint32_t)a.r2;  // 0-based window starts
    const uint32_t wb1 = ws1 >= 15u ? (ws1 - 15u) >> 4 : 0u, wb2 = ws2 >= 15u ? (ws2 - 15u) >> 4 : 0u;
    f.wbase1 = d.word_off + wb1;
    f.wbase2 = d.word_off + wb2;
    f.rel1 = (uint8_t)(ws1 - 16u * wb1);
    f.rel2 = (uint8_t)(ws2 - 16u * wb2);
    f.mol = d.mol;
    f.hdr = (uint16_t)d.hdr;
    f.P = d.P;
    f.SUF = d.SUF;
    f.d1 = d.d1;
    f.d2 = d.d2;
    f.pad = 0;
}

// One read pair, one warp, 2-bit reference, regular pair.
__device__ __forceinline__ void emit_pair_fast(const EmitArgs &a, const FastDesc &f, uint8_t *stage1, uint8_t *stage2,
                                               uint8_t *scr, uint32_t lane, unsigned long long magic, uint32_t rd_off) {
    const uint32_t half = lane >> 4, hl = lane & 15u;
    const uint32_t inl = (uint32_t)a.bclen_inline;
    const uint32_t hdr = f.hdr;
    const uint32_t S = half ? (uint32_t)a.r2 : (uint32_t)a.r1;
    uint8_t *stage = half ? stage2 : stage1;
    const uint32_t seq_off = (half ? f.sb2 : f.sb1 + inl) + hdr;
    // ---- sequences: half-warp 0 → R1 (reverse complement), half-warp 1 → R2 (forward) --------
    {
        const uint32_t W = __ldg(a.hap + (half ? f.wbase2 : f.wbase1) + hl);
        const int rel = half ? f.rel2 : f.rel1;
        const uint32_t c = (seq_off >> 4) + hl;
        const int k0 = (int)(16u * c) - (int)seq_off;                      // first base index of this chunk (>= -15)
        const int q = half ? rel + k0 : rel + (int)S - 16 - k0;           // relative to the first loaded word
        const int j = q >> 4;
        const uint32_t sh = ((uint32_t)q & 15u) * 2u;
        uint32_t lo = __shfl_sync(0xffffffffu, W, (int)((lane & 16u) | ((uint32_t)j & 15u)));
        uint32_t hi = __shfl_sync(0xffffffffu, W, (int)((lane & 16u) | ((uint32_t)(j + 1) & 15u)));
        if (j < 0) lo = 0;
        uint32_t bits = __funnelshift_r(lo, hi, sh);
        if (!half) {
            bits = __brev(bits);
            bits = ((bits >> 1) & 0x55555555u) | ((bits & 0x55555555u) << 1);
        }
        const uint4 v = expand16(bits, half ? 0x54474341u /* "ACGT" */ : 0x41434754u /* "TGCA" */);
        if (16u * c < seq_off + S) *reinterpret_cast<uint4 *>(stage + 16u * c) = v;
    }
    // ---- header scratch: prefix words, digits, suffix bytes -----------------------------------
    const uint8_t *tx = a.text + (uint64_t)f.mol * a.text_stride;
    const uint32_t P = f.P, SUF = f.SUF, d1 = f.d1, d2 = f.d2;
    if (lane < 16) reinterpret_cast<uint32_t *>(scr)[lane] = __ldg(reinterpret_cast<const uint32_t *>(tx) + lane);  // P <= 64
    __syncwarp();
    {
        const bool first = lane < 10;
        const uint32_t v = first ? f.bstart : f.bend;
        const uint32_t qd = (lane == 9 || lane >= 19) ? v : (uint32_t)__umul64hi((unsigned long long)v, magic);
        const uint32_t dig = qd - 10u * (__umulhi(qd, 0xCCCCCCCDu) >> 3);
        // digit `lane` (most significant first, 10 wide) lands at base + lane - (10 - D)
        const int pos = first ? (int)(P + d1 + lane) - 10 : (int)(P + d1 + 1u + d2 + lane) - 20;
        const int lowest = first ? (int)P : (int)(P + d1 + 1u);
        if (lane < 20 && pos >= lowest) scr[pos] = (uint8_t)('0' + dig);
        if (lane == 20) scr[P + d1] = ':';
        const uint32_t o = P + d1 + 1u + d2;
        for (uint32_t jj = lane; jj < SUF + inl; jj += 32) scr[o + jj] = __ldg(tx + P + (jj < SUF ? jj : jj - SUF + 1u));
    }
    __syncwarp();
    // header bytes to both records (consecutive bytes: conflict-free shared-memory traffic)
    {
        const uint32_t T1 = hdr + inl;
        for (uint32_t i = lane; i < T1; i += 32) {
            const uint8_t ch = scr[i];
            stage1[f.sb1 + i] = ch;
            if (i < hdr) stage2[f.sb2 + i] = (i == rd_off + P + d1 + 1u + d2) ? (uint8_t)'2' : ch;
        }
    }
    // ---- bytes after each sequence: "\n+\n", the quality bytes the sequence stores spilled on,
    //      the final newline ---------------------------------------------------------------------
    {
        const uint32_t seq_end = seq_off + S;
        const uint32_t rec_end = seq_end + 3u + S + (half ? 0u : inl) + 1u;
        const uint32_t addr = hl < 15 ? seq_end + hl : rec_end - 1u;
        const uint8_t val = (hl == 0 || hl == 2 || hl == 15) ? '\n' : (hl == 1 ? '+' : 'I');
        stage[addr] = val;
    }
    __syncwarp();
}

// stage → global: aligned 16-byte stores of the stream; the two partial chunks byte-wise
__device__ __forceinline__ void copy_out(const uint8_t *stage, uint32_t A, uint32_t bytes, uint8_t *gbase) {
    const uint32_t n16 = (A + bytes + 15u) >> 4;
    for (uint32_t j = threadIdx.x; j < n16; j += EMIT_THREADS) {
        const uint32_t lo = 16u * j, hi = lo + 16u;
        if (lo >= A && hi <= A + bytes) {
            __stcs(reinterpret_cast<uint4 *>(gbase + lo), *reinterpret_cast<const uint4 *>(stage + lo));
        } else {
            const uint32_t b0 = lo > A ? lo : A, b1 = hi < A + bytes ? hi : A + bytes;
            for (uint32_t b = b0; b < b1; ++b) gbase[b] = stage[b];
        }
    }
}

template <int BPB>
__global__ void __launch_bounds__(EMIT_THREADS, 4) k_emit(const __grid_constant__ EmitArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ PairDesc s_d[EMIT_MAX_TP];
    __shared__ FastDesc s_f[EMIT_MAX_TP];
    __shared__ uint32_t s_off1[EMIT_MAX_TP + 1], s_off2[EMIT_MAX_TP + 1];
    const uint32_t tile = a.tile0 + blockIdx.x;
    const uint64_t p0 = (uint64_t)tile * a.tile_pairs;
    const uint32_t np = (uint32_t)((a.n_pairs - p0) < a.tile_pairs ? (a.n_pairs - p0) : a.tile_pairs);
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t regular = 1;
    if (threadIdx.x < np) {
        PairDesc d;
        decode_pair(a, p0 + threadIdx.x, a.tile_mol0[tile], d);
        s_d[threadIdx.x] = d;
        regular = d.regular;
    }
    const int all_regular = __syncthreads_and((int)regular);
    // exclusive scan of record lengths within the tile (np <= 64: two items per lane of warp 0)
    if (warp == 0) {
        uint32_t a1 = lane * 2 < np ? s_d[lane * 2].len1 : 0, b1 = lane * 2 + 1 < np ? s_d[lane * 2 + 1].len1 : 0;
        uint32_t a2 = lane * 2 < np ? s_d[lane * 2].len2 : 0, b2 = lane * 2 + 1 < np ? s_d[lane * 2 + 1].len2 : 0;
        uint32_t x1 = a1 + b1, x2 = a2 + b2;
        uint32_t i1 = x1, i2 = x2;
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y1 = __shfl_up_sync(0xffffffffu, i1, o), y2 = __shfl_up_sync(0xffffffffu, i2, o);
            if (lane >= o) {
                i1 += y1;
                i2 += y2;
            }
        }
        uint32_t e1 = i1 - x1, e2 = i2 - x2;
        s_off1[lane * 2] = e1;
        s_off2[lane * 2] = e2;
        s_off1[lane * 2 + 1] = e1 + a1;
        s_off2[lane * 2 + 1] = e2 + a2;
        if (lane == 31) {
            s_off1[64] = i1;
            s_off2[64] = i2;
        }
    }
    __syncthreads();
    const uint64_t g1 = a.tile_off1[tile] - a.chunk_off1, g2 = a.tile_off2[tile] - a.chunk_off2;
    const uint32_t bytes1 = s_off1[np], bytes2 = s_off2[np];
    const uint32_t A1 = (uint32_t)(g1 & 15u), A2 = (uint32_t)(g2 & 15u);
    const bool fast = BPB == 2 && a.fast_ok && all_regular && A1 + bytes1 + 16u <= a.stage_cap && A2 + bytes2 + 16u <= a.stage_cap;
    if (!fast) {
        uint8_t *o1 = a.out1 + g1, *o2 = a.out2 + g2;
        for (uint32_t t = warp; t < np; t += EMIT_WARPS) {
            PairDesc d = s_d[t];
            write_record_generic<BPB>(a, d, 1, o1 + s_off1[t], lane);
            write_record_generic<BPB>(a, d, 2, o2 + s_off2[t], lane);
        }
        return;
    }
    uint8_t *stage1 = smem, *stage2 = smem + a.stage_cap, *scr = smem + 2u * a.stage_cap + warp * EMIT_SCR;
    {   // quality bytes for free: pre-fill both stages with 'I'
        const uint4 fill = make_uint4(0x49494949u, 0x49494949u, 0x49494949u, 0x49494949u);
        const uint32_t n1 = (A1 + bytes1 + 15u) >> 4, n2 = (A2 + bytes2 + 15u) >> 4;
        for (uint32_t j = threadIdx.x; j < n1; j += EMIT_THREADS) reinterpret_cast<uint4 *>(stage1)[j] = fill;
        for (uint32_t j = threadIdx.x; j < n2; j += EMIT_THREADS) reinterpret_cast<uint4 *>(stage2)[j] = fill;
    }
    if (threadIdx.x < np) make_fast_desc(a, s_d[threadIdx.x], A1 + s_off1[threadIdx.x], A2 + s_off2[threadIdx.x], s_f[threadIdx.x]);
    const uint32_t kk = lane < 10 ? lane : (lane < 20 ? lane - 10u : 9u);
    const unsigned long long magic = c_magic10[kk];
    const uint32_t rd_off = (a.fq_format == MMK_FMT_HAPLOTAGGING || a.fq_format == MMK_FMT_STANDARD) ? 1u : 0x80000000u;
    __syncthreads();
    for (uint32_t t = warp; t < np; t += EMIT_WARPS) {
        const FastDesc f = s_f[t];
        const uint32_t ro = rd_off == 1u ? 1u : (uint32_t)f.SUF - 13u;
        emit_pair_fast(a, f, stage1, stage2, scr, lane, magic, ro);
    }
    __syncthreads();
    copy_out(stage1, A1, bytes1, a.out1 + (g1 - A1));
    copy_out(stage2, A2, bytes2, a.out2 + (g2 - A2));
}

}  // namespace mmkdev
