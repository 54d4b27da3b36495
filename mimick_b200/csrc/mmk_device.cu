// mmk_device.cu — sm_100a kernels of the simulation core and their launchers.
//
// Kernel inventory (DESIGN.md §5 describes each and the roofline that bounds it):
//   k_normalize / k_mask_n / k_count        FASTA symbols → validated upper-case, N masking        ProcessFASTA.jl:29-36,52-57
//   k_scatter_snps / k_pack_pieces /        haplotype construction: SNP scatter + segmented gather  ProcessVariants.jl:143-176
//   k_pack_plain2                           of indel pieces, packed to 2 or 4 bits per base
//   k_group_nmol / k_fill_slots             group table: molecules per barcode group, their labels  Common.jl:89-91, Main.jl:44-47
//   k_plan_molecules                        target, size and insert count of every molecule         Main.jl:47, Breakpoints.jl:149-196
//   k_plan_cut / k_plan_sums                coverage tracker + pruning point of a batch             Main.jl:54,58
//   k_insert_draws / k_place_chain /        insert sizes, start, gap chain, header text             Breakpoints.jl:205-251,21-35
//   k_mol_text
//   k_tile_prepare / k_chunk_table          record lengths per tile → byte offsets, chunk bounds     FormatFASTQ.jl
//   k_emit (mmk_emit.cuh)                   extraction, reverse complement, FASTQ text              Breakpoints.jl:95-140, FormatFASTQ.jl
//   k_bgzf / k_bgzf_compact (mmk_bgzf.cuh)  FASTQ text → BGZF members                               Main.jl:42-43,111-112
#include <algorithm>
#include <cstddef>
#include <cstdlib>
#include <cstdio>

#include "mmk_internal.h"
#include "mmk_sampling.cuh"

namespace mmkdev {

static unsigned long long g_launches = 0;  // host-side count of kernel launches (single-threaded per context)
unsigned long long launch_count() { return g_launches; }
#define LAUNCH(k, grid, block, smem, st, ...)    \
    do {                                         \
        ++g_launches;                            \
        k<<<(grid), (block), (smem), (st)>>>(__VA_ARGS__); \
    } while (0)

// Launch that lets the grid start while the previous kernel of the stream is draining — but only behind a kernel that says so
// (griddepcontrol.launch_dependents); behind any other kernel, copy or event it is an ordinary launch.  The emit kernels of
// consecutive chunks write different ring slots and read the same placement: nothing orders them but the stream.
template <class K, class... A>
static void launch_overlapped(K kernel, unsigned grid, unsigned block, size_t smem, cudaStream_t st, A... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    ++g_launches;
    cudaLaunchKernelEx(&cfg, kernel, args...);
}

// ------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ndigits_u32(uint32_t v) {
    if (v >= 100000u) {
        if (v >= 10000000u) return v >= 1000000000u ? 10u : (v >= 100000000u ? 9u : 8u);
        return v >= 1000000u ? 7u : 6u;
    }
    if (v >= 1000u) return v >= 10000u ? 5u : 4u;
    return v >= 100u ? 3u : (v >= 10u ? 2u : 1u);
}
__constant__ uint32_t c_pow10[10] = {1u, 10u, 100u, 1000u, 10000u, 100000u, 1000000u, 10000000u, 100000000u, 1000000000u};

// 4-bit code (BioSequences bit mask A=1 C=2 G=4 T=8) of an ASCII symbol, 0xff if invalid.
__device__ __forceinline__ uint32_t ascii_to_code4(uint32_t c) {
    c &= 0xdfu;  // upper-case letters; '-' (0x2d) → 0x0d handled below
    switch (c) {
    case 'A': return 1; case 'C': return 2; case 'G': return 4; case 'T': return 8;
    case 'M': return 3; case 'R': return 5; case 'W': return 9; case 'S': return 6;
    case 'Y': return 10; case 'K': return 12; case 'V': return 7; case 'H': return 11;
    case 'D': return 13; case 'B': return 14; case 'N': return 15;
    default: return 0xffu;
    }
}
__device__ __forceinline__ char code4_to_ascii(uint32_t k) { return "-ACMGRSVTWYHKDBN"[k & 15u]; }
__device__ __forceinline__ uint32_t code4_to_code2(uint32_t k) { return (k == 2u) ? 1u : (k == 4u) ? 2u : (k == 8u) ? 3u : 0u; }
__device__ __forceinline__ uint32_t comp4(uint32_t k) { return (__brev(k) >> 28) & 15u; }

// ------------------------------------------------------------------------------------------
// FASTA symbols
// ------------------------------------------------------------------------------------------
// all four bytes of w are upper-case A, C, G or T
__device__ __forceinline__ bool word_is_acgt(uint32_t w) {
    const uint32_t m = __vcmpeq4(w, 0x41414141u) | __vcmpeq4(w, 0x43434343u) | __vcmpeq4(w, 0x47474747u) | __vcmpeq4(w, 0x54545454u);
    return m == 0xffffffffu;
}
__device__ __forceinline__ void normalize_one(char *bases, int64_t i, unsigned long long &nN, unsigned long long &nX,
                                              unsigned long long &bad) {
    uint32_t c = (uint8_t)bases[i];
    uint32_t k = (c == '-') ? 0u : ascii_to_code4(c);
    if (k == 0xffu || ((c & 0xdfu) < 'A' && c != '-')) {
        if ((unsigned long long)i < bad) bad = (unsigned long long)i;
        return;
    }
    bases[i] = code4_to_ascii(k);
    nN += (k == 15u);
    nX += !(k == 1u || k == 2u || k == 4u || k == 8u);
}
// Upper-cases, validates and counts.  16 symbols per thread and iteration; runs that are already upper-case ACGT
// (nearly all of a real genome) are only read.  `bases` is 16-byte aligned.
__global__ void k_normalize(char *bases, int64_t len, ContigStats *stats) {
    const int64_t n16 = len >> 4;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    unsigned long long nN = 0, nX = 0, bad = ~0ull;
    const uint4 *v = reinterpret_cast<const uint4 *>(bases);
    for (; i < n16; i += stride) {
        const uint4 x = v[i];
        if (word_is_acgt(x.x) && word_is_acgt(x.y) && word_is_acgt(x.z) && word_is_acgt(x.w)) continue;
        for (int j = 0; j < 16; ++j) normalize_one(bases, 16 * i + j, nN, nX, bad);
    }
    {   // the last len % 16 symbols
        const int64_t t = (n16 << 4) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (t < len) normalize_one(bases, t, nN, nX, bad);
    }
    for (int o = 16; o; o >>= 1) {
        nN += __shfl_down_sync(0xffffffffu, nN, o);
        nX += __shfl_down_sync(0xffffffffu, nX, o);
        unsigned long long b2 = __shfl_down_sync(0xffffffffu, bad, o);
        bad = b2 < bad ? b2 : bad;
    }
    if ((threadIdx.x & 31) == 0) {
        if (nN) atomicAdd(&stats->n_N, nN);
        if (nX) atomicAdd(&stats->n_nonACGT, nX);
        if (bad != ~0ull) atomicMin(&stats->first_bad, bad);
    }
}

__global__ void k_count(const char *bases, int64_t len, ContigStats *stats) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    unsigned long long nN = 0, nX = 0;
    for (; i < len; i += stride) {
        char c = bases[i];
        nN += (c == 'N');
        nX += !(c == 'A' || c == 'C' || c == 'G' || c == 'T');
    }
    for (int o = 16; o; o >>= 1) {
        nN += __shfl_down_sync(0xffffffffu, nN, o);
        nX += __shfl_down_sync(0xffffffffu, nX, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (nN) atomicAdd(&stats->n_N, nN);
        if (nX) atomicAdd(&stats->n_nonACGT, nX);
    }
}

// mask_N!: every N ← (A,C,G,T)[pick(4)], the pick addressed by (contig, haplotype file, position)
__global__ void k_mask_n(char *bases, int64_t len, uint64_t seed, uint32_t contig, uint32_t hap) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    mmk_stream s;
    mmk_stream_init(&s, seed, 0, MMK_P_MASK, contig, hap);
    for (; i < len; i += stride) {
        if (bases[i] != 'N') continue;
        uint32_t k = (uint32_t)mmk_bits_to_index(mmk_stream_at(&s, (uint32_t)i), 4);
        bases[i] = "ACGT"[k];
    }
}

__global__ void k_scatter_snps(char *bases, const int64_t *pos, const char *alt, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) bases[pos[i] - 1] = alt[i];
}

// Segmented gather + pack. One thread per output word.
template <int BPB>
__global__ void k_pack_pieces(const char *__restrict__ master, const char *__restrict__ alt,
                              const int64_t *__restrict__ dst, const int64_t *__restrict__ src,
                              const uint8_t *__restrict__ kind, int32_t n_pieces, int64_t out_len,
                              uint32_t *__restrict__ words, ContigStats *stats) {
    constexpr int PER = 32 / BPB;
    int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t o = w * PER;
    if (o >= out_len) return;
    // piece containing o
    int32_t lo = 0, hi = n_pieces - 1;
    while (lo < hi) {
        int32_t mid = (lo + hi + 1) >> 1;
        if (dst[mid] <= o) lo = mid; else hi = mid - 1;
    }
    int32_t p = lo;
    int64_t pdst = dst[p], psrc = src[p];
    int64_t pend = (p + 1 < n_pieces) ? dst[p + 1] : out_len;
    const char *base = kind[p] ? alt : master;
    uint32_t word = 0;
    bool bad = false;
#pragma unroll 1
    for (int j = 0; j < PER && o + j < out_len; ++j) {
        int64_t q = o + j;
        while (q >= pend) {
            ++p;
            pdst = dst[p];
            psrc = src[p];
            pend = (p + 1 < n_pieces) ? dst[p + 1] : out_len;
            base = kind[p] ? alt : master;
        }
        uint32_t c = (uint8_t)base[psrc + (q - pdst)];
        uint32_t k4 = (c == '-') ? 0u : ascii_to_code4(c);
        if (BPB == 2) {
            bad |= !(k4 == 1u || k4 == 2u || k4 == 4u || k4 == 8u);
            word |= code4_to_code2(k4) << (2 * j);
        } else {
            bad |= (k4 == 0xffu);
            word |= (k4 & 15u) << (4 * j);
        }
    }
    words[w] = word;
    if (bad) atomicMin(&stats->first_bad, (unsigned long long)o);
}

// One unmodified contig, 2 bits per base: 16 upper-case ACGT symbols (one 16-byte load) -> one word.
// 'A' 0x41, 'C' 0x43, 'G' 0x47, 'T' 0x54: t = (c >> 1) & 3 gives 0, 1, 3, 2 and t ^ (t >> 1) the codes 0, 1, 2, 3;
// the multiplication gathers the four 2-bit fields of a word into its top byte.
__device__ __forceinline__ uint32_t pack4_acgt(uint32_t w) {
    const uint32_t t = (w >> 1) & 0x03030303u;
    const uint32_t c = t ^ ((t >> 1) & 0x01010101u);
    return (c * 0x01041040u) >> 24;
}
__global__ void k_pack_plain2(const char *__restrict__ ascii, int64_t len, uint32_t *__restrict__ words, ContigStats *stats) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t o = w * 16;
    if (o >= len) return;
    if (o + 16 <= len) {
        const uint4 x = reinterpret_cast<const uint4 *>(ascii)[w];
        if (!(word_is_acgt(x.x) && word_is_acgt(x.y) && word_is_acgt(x.z) && word_is_acgt(x.w)))
            atomicMin(&stats->first_bad, (unsigned long long)o);
        words[w] = pack4_acgt(x.x) | (pack4_acgt(x.y) << 8) | (pack4_acgt(x.z) << 16) | (pack4_acgt(x.w) << 24);
        return;
    }
    uint32_t word = 0;
    bool bad = false;
    for (int j = 0; o + j < len; ++j) {
        const uint32_t c = (uint8_t)ascii[o + j];
        const uint32_t k4 = ascii_to_code4(c);
        bad |= !(k4 == 1u || k4 == 2u || k4 == 4u || k4 == 8u);
        word |= code4_to_code2(k4) << (2 * j);
    }
    words[w] = word;
    if (bad) atomicMin(&stats->first_bad, (unsigned long long)o);
}

template <int BPB>
__global__ void k_unpack(const uint32_t *__restrict__ words, int64_t len, char *__restrict__ out) {
    constexpr int PER = 32 / BPB;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len) return;
    uint32_t w = words[i / PER];
    uint32_t k = (w >> (BPB * (i % PER))) & ((1u << BPB) - 1u);
    out[i] = BPB == 2 ? "ACGT"[k] : code4_to_ascii(k);
}

static inline int grid_for(int64_t n, int block, int cap = 148 * 16) {
    int64_t g = (n + block - 1) / block;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (int)g;
}

void normalize_ascii(char *d, int64_t len, ContigStats *st, cudaStream_t s) {
    LAUNCH(k_normalize, grid_for(len, 256), 256, 0, s, d, len, st);
}
void count_bases(const char *d, int64_t len, ContigStats *st, cudaStream_t s) {
    LAUNCH(k_count, grid_for(len, 256), 256, 0, s, d, len, st);
}
void mask_n(char *d, int64_t len, uint64_t seed, uint32_t contig, uint32_t hap, cudaStream_t s) {
    LAUNCH(k_mask_n, grid_for(len, 256), 256, 0, s, d, len, seed, contig, hap);
}
void scatter_snps(char *d, const int64_t *pos, const char *alt, int64_t n, cudaStream_t s) {
    if (n > 0) LAUNCH(k_scatter_snps, (unsigned)((n + 255) / 256), 256, 0, s, d, pos, alt, n);
}
void pack_pieces(const char *master, const char *alt, const int64_t *dst, const int64_t *src, const uint8_t *kind,
                 int32_t n_pieces, int64_t out_len, uint32_t *words, int bpb, ContigStats *st, cudaStream_t s) {
    if (out_len <= 0) return;
    int per = 32 / bpb;
    int64_t nw = (out_len + per - 1) / per;
    unsigned g = (unsigned)((nw + 255) / 256);
    if (bpb == 2)
        LAUNCH((k_pack_pieces<2>), g, 256, 0, s, master, alt, dst, src, kind, n_pieces, out_len, words, st);
    else
        LAUNCH((k_pack_pieces<4>), g, 256, 0, s, master, alt, dst, src, kind, n_pieces, out_len, words, st);
}
// one whole contig as it is (upper-case ACGT only), 2 bits per base; `ascii` 16-byte aligned
void pack_plain2(const char *ascii, int64_t len, uint32_t *words, ContigStats *st, cudaStream_t s) {
    if (len <= 0) return;
    const int64_t nw = (len + 15) / 16;
    LAUNCH(k_pack_plain2, (unsigned)((nw + 255) / 256), 256, 0, s, ascii, len, words, st);
}
void unpack(const uint32_t *words, int64_t len, int bpb, char *out, cudaStream_t s) {
    if (len <= 0) return;
    unsigned g = (unsigned)((len + 255) / 256);
    if (bpb == 2) LAUNCH((k_unpack<2>), g, 256, 0, s, words, len, out);
    else LAUNCH((k_unpack<4>), g, 256, 0, s, words, len, out);
}

// ------------------------------------------------------------------------------------------
// exclusive scan u32 → u64 (reduce / spine / apply; 1024 items per block)
// ------------------------------------------------------------------------------------------
constexpr int SCAN_T = 256, SCAN_I = 4, SCAN_B = SCAN_T * SCAN_I;

__device__ __forceinline__ uint64_t block_exclusive_u64(uint64_t v, uint64_t *s_warp, uint64_t &total) {
    // inclusive warp scan
    uint64_t x = v;
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) {
        uint64_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) s_warp[wid] = x;
    __syncthreads();
    if (wid == 0) {
        uint64_t w = lane < (blockDim.x >> 5) ? s_warp[lane] : 0;
        uint64_t ws = w;
        for (int o = 1; o < 32; o <<= 1) {
            uint64_t y = __shfl_up_sync(0xffffffffu, ws, o);
            if (lane >= o) ws += y;
        }
        s_warp[lane] = ws - w;                  // exclusive per warp
        if (lane == 31) s_warp[32] = ws;        // block total
    }
    __syncthreads();
    uint64_t res = s_warp[wid] + x - v;
    total = s_warp[32];
    __syncthreads();
    return res;
}

__global__ void k_scan_reduce(const uint32_t *in, uint64_t n, uint64_t *block_sums) {
    __shared__ uint64_t s_warp[33];
    uint64_t base = (uint64_t)blockIdx.x * SCAN_B + (uint64_t)threadIdx.x * SCAN_I;
    uint64_t v = 0;
#pragma unroll
    for (int j = 0; j < SCAN_I; ++j)
        if (base + j < n) v += in[base + j];
    uint64_t total;
    block_exclusive_u64(v, s_warp, total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}
__global__ void k_scan_spine(uint64_t *block_sums, uint64_t nb) {
    __shared__ uint64_t s_warp[33];
    uint64_t carry = 0;
    for (uint64_t base = 0; base < nb; base += blockDim.x) {
        uint64_t i = base + threadIdx.x;
        uint64_t v = i < nb ? block_sums[i] : 0;
        uint64_t total;
        uint64_t ex = block_exclusive_u64(v, s_warp, total);
        if (i < nb) block_sums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) block_sums[nb] = carry;
}
__global__ void k_scan_apply(const uint32_t *in, uint64_t n, const uint64_t *block_sums, uint64_t *out, const uint64_t *base_ptr) {
    __shared__ uint64_t s_warp[33];
    uint64_t base = (uint64_t)blockIdx.x * SCAN_B + (uint64_t)threadIdx.x * SCAN_I;
    uint32_t it[SCAN_I];
    uint64_t v = 0;
#pragma unroll
    for (int j = 0; j < SCAN_I; ++j) {
        it[j] = base + j < n ? in[base + j] : 0;
        v += it[j];
    }
    uint64_t total;
    const uint64_t b0 = base_ptr ? *base_ptr : 0ull;  // out[0] may alias *base_ptr: read before any write below
    __syncthreads();
    uint64_t ex = block_exclusive_u64(v, s_warp, total) + block_sums[blockIdx.x] + b0;
#pragma unroll
    for (int j = 0; j < SCAN_I; ++j) {
        if (base + j < n) out[base + j] = ex;
        ex += it[j];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) out[n] = block_sums[gridDim.x] + b0;
}
// out[i] = base + sum(in[0..i)), out[n] = base + total; base = *d_base (device) or 0
void exclusive_scan_u32(const uint32_t *in, uint64_t *out, uint64_t n, uint64_t *tmp, cudaStream_t s, const uint64_t *d_base) {
    if (n == 0) {
        if (!d_base) cudaMemsetAsync(out, 0, 8, s);
        else if (out != d_base) cudaMemcpyAsync(out, d_base, 8, cudaMemcpyDeviceToDevice, s);
        return;
    }
    unsigned nb = (unsigned)((n + SCAN_B - 1) / SCAN_B);
    LAUNCH(k_scan_reduce, nb, SCAN_T, 0, s, in, n, tmp);
    LAUNCH(k_scan_spine, 1, 1024, 0, s, tmp, nb);
    LAUNCH(k_scan_apply, nb, SCAN_T, 0, s, in, n, tmp, out, d_base);
}

// ------------------------------------------------------------------------------------------
// planning: barcode groups → molecules, coverage tracker, pruning point
// ------------------------------------------------------------------------------------------
__global__ void k_group_nmol(SimConst sc, uint64_t g0, uint32_t n_groups, uint32_t *nmol) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_groups) return;
    NoRec rec;
    nmol[i] = draw_nmol(sc, g0 + i, rec);
}

// group table, second half: (group, slot) of every molecule — independent of the live set
__global__ void k_fill_slots(uint64_t g_lo, uint32_t n_groups, const uint32_t *__restrict__ nmol, const uint64_t *__restrict__ off,
                             PlanDev plan) {
    uint32_t gl = blockIdx.x * blockDim.x + threadIdx.x;
    if (gl >= n_groups) return;
    const uint64_t g = g_lo + gl, o = off[g];
    const uint32_t n = nmol[g];
    for (uint32_t m = 0; m < n; ++m) {
        plan.group[o + m] = g;
        plan.slot[o + m] = m;
    }
}


}  // namespace mmkdev
#include "mmk_plan.cuh"
namespace mmkdev {

uint32_t plan_hist_chunks(uint64_t max_molecules) { return (uint32_t)((max_molecules + HCH - 1) / HCH); }

// group table: n_mol of groups [g_lo, g_hi) and their molecule offsets (off[g_lo] must already be valid)
void plan_extend_table(const SimConst &sc, uint64_t g_lo, uint64_t g_hi, uint32_t *d_nmol, uint64_t *d_off, uint64_t *d_scan_tmp,
                       cudaStream_t st) {
    if (g_hi <= g_lo) return;
    const uint32_t n = (uint32_t)(g_hi - g_lo);
    LAUNCH(k_group_nmol, (n + 255) / 256, 256, 0, st, sc, g_lo, n, d_nmol + g_lo);
    exclusive_scan_u32(d_nmol + g_lo, d_off + g_lo, n, d_scan_tmp, st, d_off + g_lo);
}
// after the plan arrays were sized for off[g_hi] molecules
void plan_fill_slots(uint64_t g_lo, uint64_t g_hi, const uint32_t *d_nmol, const uint64_t *d_off, PlanDev plan, cudaStream_t st) {
    if (g_hi <= g_lo) return;
    const uint32_t n = (uint32_t)(g_hi - g_lo);
    LAUNCH(k_fill_slots, (n + 255) / 256, 256, 0, st, g_lo, n, d_nmol, d_off, plan);
}

void plan_init(PlanState *d_st, const PlanArrays &pa, uint32_t K, cudaStream_t st) {
    LAUNCH(k_plan_init, (K + 255) / 256, 256, 0, st, d_st, pa, K);
}
void plan_batches(const SimConst &sc, const SchemaDev *d_sch, PlanState *d_st, const PlanArrays &pa, const uint64_t *d_off, PlanDev plan,
                  unsigned long long *d_first_err, int n_batches, cudaStream_t st) {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    const unsigned grid = (unsigned)sms * 8u;  // 2048 threads per SM: every batch is one wave, grid-stride inside
    for (int k = 0; k < n_batches; ++k) {
        LAUNCH(k_plan_molecules, grid, PLAN_THREADS, 0, st, sc, d_sch, d_st, pa, d_off, plan, d_first_err);
        LAUNCH(k_plan_cut, (unsigned)sms, PLAN_THREADS, 0, st, d_st, pa, plan.schema, plan.nfrags, plan.group, d_off);
        LAUNCH(k_plan_close, (unsigned)sms, PLAN_THREADS, 0, st, d_st, pa, plan.schema, plan.nfrags, d_off, plan.status, d_first_err);
    }
}

// ------------------------------------------------------------------------------------------
// sharding
// ------------------------------------------------------------------------------------------
__global__ void k_shard_bounds(const uint64_t *read_off, uint64_t n_mol, uint64_t n_pairs, int32_t world, uint64_t *bounds) {
    int w = threadIdx.x;
    if (w > world) return;
    if (w == world) {
        bounds[w] = n_mol;
        return;
    }
    uint64_t target = n_pairs / (uint64_t)world * (uint64_t)w + n_pairs % (uint64_t)world * (uint64_t)w / (uint64_t)world;
    uint64_t lo = 0, hi = n_mol;  // smallest m with read_off[m] >= target
    while (lo < hi) {
        uint64_t mid = (lo + hi) >> 1;
        if (read_off[mid] >= target) hi = mid; else lo = mid + 1;
    }
    bounds[w] = lo;
}
void shard_bounds(const uint64_t *d_read_off, uint64_t n_mol, uint64_t n_pairs, int32_t world, uint64_t *d_bounds,
                  cudaStream_t st) {
    LAUNCH(k_shard_bounds, 1, 1024, 0, st, d_read_off, n_mol, n_pairs, world, d_bounds);
}

// ------------------------------------------------------------------------------------------
// placement
// ------------------------------------------------------------------------------------------
// Length of circular_index(seq, a:b) (Breakpoints.jl:111-121) for a sequence of n bases.
__device__ __forceinline__ uint32_t window_len(int64_t a, int64_t b, uint32_t n, int strict) {
    if (strict || b <= (int64_t)n) return (uint32_t)(b - a + 1);
    uint32_t part1 = a <= (int64_t)n ? (uint32_t)((int64_t)n - a + 1) : 0u;
    return part1 + (uint32_t)b % n;  // b < 2^32: contigs are < 2^31 bases
}

__device__ __forceinline__ uint8_t *put_dec(uint8_t *p, uint64_t v) {
    char tmp[20];
    int n = 0;
    do {
        tmp[n++] = (char)('0' + v % 10);
        v /= 10;
    } while (v);
    while (n) *p++ = (uint8_t)tmp[--n];
    return p;
}
__device__ __forceinline__ uint8_t *put_str(uint8_t *p, const char *s) {
    while (*s) *p++ = (uint8_t)*s++;
    return p;
}
__device__ __forceinline__ uint8_t *put_barcode(uint8_t *p, int bc_format, uint64_t k) {
    switch (bc_format) {
    case MMK_FMT_HAPLOTAGGING: {
        uint32_t v[4] = {(uint32_t)(k % 96), (uint32_t)(k / 96 % 96), (uint32_t)(k / 9216 % 96), (uint32_t)(k / 884736)};
        const char L[4] = {'A', 'C', 'B', 'D'};
        for (int i = 0; i < 4; ++i) {
            uint32_t x = v[i] + 1;
            *p++ = (uint8_t)L[i];
            *p++ = (uint8_t)('0' + x / 10);
            *p++ = (uint8_t)('0' + x % 10);
        }
        return p;
    }
    case MMK_FMT_STLFR:
        p = put_dec(p, k % 2000 + 1);
        *p++ = '_';
        p = put_dec(p, k / 2000 % 2000 + 1);
        *p++ = '_';
        return put_dec(p, k / 4000000 + 1);
    default: {
        int n = bc_format == MMK_FMT_TELLSEQ ? 18 : 16;
        for (int t = 0; t < n; ++t) *p++ = (uint8_t)"ATGC"[(k >> (2 * t)) & 3];
        return p;
    }
    }
}

// Read-window checks the reference would fail on (BoundsError) and the quality-string lengths,
// which follow the molecule's FIRST reads (FormatFASTQ.jl:14: "I"^length(read_sequences.first[1])).
// __noinline__: cicc 12.9 crashes when this loop is inlined behind molecule_place.
__device__ __noinline__ uint32_t check_windows(const uint2 *ins, uint32_t nfrags, int64_t start, uint32_t n, int r1, int r2,
                                               int strict, uint32_t &q1, uint32_t &q2) {
    uint32_t status = 0;
    for (uint32_t k = 0; k < nfrags; ++k) {
        int64_t bs = start + ins[k].x, be = bs + MMK_INS_LEN(ins[k].y) - 1;
        int64_t a2 = be - r2 + 1;
        if (!strict && a2 < 1) status = MMK_ERR_BOUNDS;
        if (k == 0) {
            q1 = window_len(bs, bs + r1 - 1, n, strict);
            q2 = (a2 >= 1 || strict) ? window_len(a2, be, n, strict) : 0u;
        }
    }
    return status;
}

// shard-relative first pair of every shard molecule, and the first molecule of every tile
__global__ void k_local_roff(const uint64_t *__restrict__ roff_g, uint64_t mol0, uint32_t n_mol, uint64_t *roff_l) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n_mol) roff_l[i] = roff_g[mol0 + i] - roff_g[mol0];
}
__global__ void k_tile_mol0(const uint64_t *__restrict__ roff_l, uint32_t n_mol, uint32_t tile_pairs, uint32_t *tile_mol0) {
    uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n_mol) return;
    uint64_t lo = roff_l[m], hi = roff_l[m + 1];  // pairs [lo, hi)
    if (hi == lo) return;
    const uint32_t sh = 31 - __clz(tile_pairs);  // tile_pairs is a power of two
    for (uint64_t t = (lo + tile_pairs - 1) >> sh; (t << sh) < hi; ++t) tile_mol0[t] = m;
}

}  // namespace mmkdev
#include "mmk_place.cuh"  // threshold-table draws (pair_draw), sum_over, the fused placement kernels
namespace mmkdev {

// Placement, step 1 — one thread per read pair: the insert length of redraw round 0 and the 64
// random bits of the pair's gap draw.  Everything that needs a transcendental or a Philox block is
// here, flat over pairs, so it runs without divergence.  Per molecule it also accumulates the integer
// sum of the lengths, one atomic per warp run.
struct __align__(16) PairTmp {
    uint64_t bits;  // gap draw: gap = mulhi64(bits, remaining + 1)
    uint32_t len;   // trunc(Int, insert length)
    uint32_t dm;    // molecule index - first molecule of the pair's tile
};
__global__ void __launch_bounds__(256) k_insert_draws(const __grid_constant__ SimConst sc, PlanDev plan, uint64_t mol0, uint32_t n_mol,
                                                      const uint64_t *__restrict__ roff_l, const uint32_t *__restrict__ tile_mol0, uint32_t n_tiles,
                                                      uint32_t tile_pairs, uint64_t pair_lo, uint64_t pair_hi, PairTmp *__restrict__ tmp,
                                                      uint32_t *acc_i, const InsTable *__restrict__ gtab) {
    __shared__ InsTable tab;
    stage_table(&tab, gtab);
    __syncthreads();
    const uint64_t p_raw = pair_lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = p_raw < pair_hi;
    const uint64_t p = valid ? p_raw : pair_hi - 1;
    const uint32_t tile = (uint32_t)(p >> (31 - __clz(tile_pairs)));  // tile_pairs is a power of two
    const uint32_t tm0 = tile_mol0[tile];
    uint32_t lo = tm0, hi = tile + 1 < n_tiles ? tile_mol0[tile + 1] : n_mol - 1;
    while (lo < hi) {  // largest m with roff_l[m] <= p
        uint32_t mid = (lo + hi + 1) >> 1;
        if (roff_l[mid] <= p) lo = mid; else hi = mid - 1;
    }
    const uint32_t i = (uint32_t)(p - roff_l[lo]);
    const uint64_t g = plan.group[mol0 + lo];
    const uint32_t m = plan.slot[mol0 + lo];
    // ONE Philox block per pair: words (0,1) feed the length sampler (threshold table), words (2,3) are the gap draw
    uint32_t blo, bhi;
    const uint32_t li = pair_draw(sc, tab, g, m, i, blo, bhi);
    if (valid) {
        PairTmp t;
        t.bits = ((uint64_t)bhi << 32) | blo;
        t.len = li;
        t.dm = lo - tm0;
        tmp[p - pair_lo] = t;
    }
    // lanes of the same molecule are contiguous: one atomic per run
    const unsigned peers = __match_any_sync(0xffffffffu, valid ? lo : 0xffffffffu);
    const uint32_t s_len = __reduce_add_sync(peers, li);
    if (valid && (threadIdx.x & 31u) == (uint32_t)(__ffs(peers) - 1)) atomicAdd(&acc_i[lo], s_len);
}

// Placement, step 2 — one thread per molecule: the redraw rule on the float sum of the insert lengths
// (Breakpoints.jl:209-212), the start draw (:240), the serial gap chain (:27-33), the window checks, the
// header text and the per-molecule record.  `sum > molsize` is decided from the fixed-point bounds of
// step 1; only when those straddle the molecule size is the left-to-right float sum (the oracle's
// definition) evaluated.
// whole-vector redraws while the float sum exceeds the molecule (Breakpoints.jl:209-212); true if it still does
__device__ __noinline__ bool redraw_rounds(const SimConst &sc, uint64_t g, uint32_t m, uint32_t nfrags, uint32_t molsize,
                                           PairTmp *tp, uint64_t &isum) {
    NoRec rec;
    bool over = true;
    int attempts = 0;
    while (over && attempts < sc.attempts) {
        ++attempts;
        double sum = 0.0;
        isum = 0;
        const uint64_t key = mmk_round_key(sc.seed, (uint32_t)attempts);
        for (uint32_t k = 0; k < nfrags; ++k) {
            double x = draw_insert(sc, key, g, m, k, rec);
            sum = MMK_ADD(sum, x);
            const uint32_t li = insert_len_of(x);
            tp[k].len = li;
            isum += li;
        }
        over = sum > (double)molsize;
    }
    return over;
}
__device__ __forceinline__ void place_chain_one(const SimConst &sc, const SchemaDev *__restrict__ sch, const PlanDev &plan,
                                                uint64_t mol0, uint32_t i, const uint64_t *__restrict__ roff_l, uint64_t pair_lo,
                                                PairTmp *tmp, const uint32_t *__restrict__ acc_i, uint2 *ins, MolRec *mol,
                                                unsigned long long *err) {
    uint64_t gi = mol0 + i;
    uint64_t ro = roff_l[i];
    uint64_t g = plan.group[gi];
    uint32_t m = plan.slot[gi];
    uint32_t t = plan.schema[gi];
    uint32_t molsize = plan.size[gi], nfrags = plan.nfrags[gi];
    const SchemaDev &S = sch[t];
    const uint32_t n = S.len;
    PairTmp *tp = tmp + (ro - pair_lo);
    uint2 *in = ins + ro;
    uint64_t isum = acc_i[i];
    bool over = sum_over(sc, sc.seed, g, m, nfrags, molsize, isum);  // from the integer sum unless it straddles the molecule size
    if (over) over = redraw_rounds(sc, g, m, nfrags, molsize, tp, isum);  // rare: kept out of line
    uint32_t status = over ? (uint32_t)MMK_ERR_INSERTS_UNFITTABLE : 0u;
    // start_pos = rand(1:adjusted_end)
    uint64_t adjusted_end = sc.circular ? (uint64_t)n : (n > molsize + 1u ? (uint64_t)(n - molsize) : 1ull);
    mmk_stream s;
    mmk_stream_init(&s, sc.seed, sc.sample, MMK_P_START, g, m);
    int64_t start = (int64_t)mmk_next_index(&s, adjusted_end) + 1;
    uint32_t q1 = 0, q2 = 0;
    if (!status) {
        uint64_t remaining = (uint64_t)molsize + 1 - isum, cursor = 0;
        // R2 window bstop - r2 + 1 : bstop must start at >= 1 (Julia BoundsError otherwise): st + len >= r2 + 1 - start
        const uint32_t need_end = (!sc.strict && start <= (int64_t)sc.r2) ? (uint32_t)((int64_t)sc.r2 + 1 - start) : 0u;
        // the loop is latency-bound on its own loads: keep two 128-byte lines of the scratch in flight
        asm volatile("prefetch.global.L1 [%0];" ::"l"(tp));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(tp + 8));
        PairTmp e = tp[0];
        for (uint32_t k = 0; k < nfrags; ++k) {
            if ((k & 7u) == 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(tp + k + 16));
            const PairTmp nxt = tp[k + 1 < nfrags ? k + 1 : k];
            const uint32_t gap = (uint32_t)mmk_bits_to_index(e.bits, remaining + 1);
            const uint32_t st = (uint32_t)cursor + gap;
            in[k] = make_uint2(st, e.len | (e.dm << 24));
            cursor = st + e.len;
            remaining -= gap;
            if (st + e.len < need_end) status = MMK_ERR_BOUNDS;
            e = nxt;
        }
        // quality lengths follow the molecule's FIRST reads (FormatFASTQ.jl:14)
        if (!status) (void)check_windows(in, 1, start, n, sc.r1, sc.r2, sc.strict, q1, q2);
    }
    if (sc.fq_format == MMK_FMT_TENX) q1 += 16;
    if (status) {
        atomicMin(&err[0], (unsigned long long)gi);
        plan.status[gi] = (uint8_t)status;
    }
    MolRec r;
    r.start = start;
    r.size = molsize;
    r.schema = (uint16_t)t;
    r.prefix_len = 0;  // k_mol_text fills the two lengths
    r.suffix_len = 0;
    r.qual1 = q1;
    r.qual2 = q2;
    mol[i] = r;
}

// Placement, step 3 — one thread per molecule: the header text every record of the molecule shares.  A thread builds
// its text in its own shared-memory row (row stride = text_stride + 4 bytes: an odd number of words, so the byte stores of
// a warp spread over the banks); the block then copies its rows — contiguous in global memory — with 16-byte stores.
constexpr uint32_t TEXT_THREADS = 128;
__global__ void __launch_bounds__(TEXT_THREADS) k_mol_text(SimConst sc, const SchemaDev *__restrict__ sch, uint64_t mol0, uint32_t m_lo, uint32_t m_hi,
                                                           MolRec *mol, uint8_t *text, uint32_t text_stride) {
    extern __shared__ __align__(16) uint8_t s_rows[];
    __shared__ uint32_t s_nb[TEXT_THREADS];
    const uint32_t row_stride = text_stride + 4u;
    const uint32_t i0 = m_lo + blockIdx.x * TEXT_THREADS, i = i0 + threadIdx.x;
    s_nb[threadIdx.x] = 0;
    if (i < m_hi) {
        const uint64_t gi = mol0 + i;
        MolRec r = mol[i];
        const SchemaDev &S = sch[r.schema];
        const int64_t start = r.start;
        const uint32_t molsize = r.size;
        // "@chrom:hap|start:end|" then the format's tail with read digit '1'
        uint8_t *buf = s_rows + (size_t)threadIdx.x * row_stride;
        uint8_t *p = buf;
        *p++ = '@';
        for (int k = 0; k < S.name_len; ++k) *p++ = (uint8_t)S.name[k];
        *p++ = ':';
        p = put_dec(p, S.haplotype);
        *p++ = '|';
        p = put_dec(p, (uint64_t)start);
        *p++ = ':';
        p = put_dec(p, (uint64_t)(start + molsize));
        *p++ = '|';
        const uint32_t plen = (uint32_t)(p - buf);
        uint8_t *sp = p;
        switch (sc.fq_format) {
        case MMK_FMT_STLFR: *p++ = '#'; p = put_barcode(p, sc.bc_format, gi); p = put_str(p, " 1:N:0:ATGACA\n"); break;
        case MMK_FMT_HAPLOTAGGING: p = put_str(p, "/1\tBX:Z:"); p = put_barcode(p, sc.bc_format, gi); *p++ = '\n'; break;
        case MMK_FMT_TELLSEQ: *p++ = ':'; p = put_barcode(p, sc.bc_format, gi); p = put_str(p, " 1:N:0:ATGACA\n"); break;
        case MMK_FMT_TENX: *p++ = '|'; p = put_barcode(p, sc.bc_format, gi); p = put_str(p, " 1:N:0:ATGACA\n"); break;
        default: p = put_str(p, "/1\tVX:i:1\tBX:Z:"); p = put_barcode(p, sc.bc_format, gi); *p++ = '\n'; break;
        }
        s_nb[threadIdx.x] = (uint32_t)(p - buf);
        r.prefix_len = (uint8_t)plen;
        r.suffix_len = (uint8_t)(p - sp);
        mol[i] = r;
    }
    __syncthreads();
    // the block's rows are contiguous in `text`: 16-byte chunk c of the block lies in row 16 c / text_stride (a multiple of 16)
    const uint32_t chunks_per_row = text_stride >> 4, n_rows = min(TEXT_THREADS, m_hi - i0);
    uint4 *dst = reinterpret_cast<uint4 *>(text + (uint64_t)i0 * text_stride);
    for (uint32_t c = threadIdx.x; c < n_rows * chunks_per_row; c += TEXT_THREADS) {
        const uint32_t row = c / chunks_per_row, off = (c - row * chunks_per_row) << 4;
        if (off >= s_nb[row]) continue;  // past the text: those bytes are never read
        const uint32_t *src = reinterpret_cast<const uint32_t *>(s_rows + (size_t)row * row_stride + off);
        dst[c] = make_uint4(src[0], src[1], src[2], src[3]);
    }
}

void place_prepare(uint64_t mol0, uint32_t n_mol, const uint64_t *d_read_off_global, uint64_t *d_read_off_local,
                   uint32_t *d_tile_mol0, uint32_t tile_pairs, cudaStream_t st) {
    LAUNCH(k_local_roff, (n_mol + 1 + 255) / 256, 256, 0, st, d_read_off_global, mol0, n_mol, d_read_off_local);
    if (n_mol == 0) return;
    LAUNCH(k_tile_mol0, (n_mol + 255) / 256, 256, 0, st, d_read_off_local, n_mol, tile_pairs, d_tile_mol0);
}

// One CTA places PLACE_BLOCK consecutive molecules.  It first orders them by insert count (counting sort
// in shared memory), so that the lanes of a warp loop alike — with the raw order only ~6 of 32 lanes are
// active on average (insert counts are exponentially distributed) — while keeping the block's memory
// traffic local.
constexpr uint32_t PLACE_BLOCK = 1024, SORT_BINS = 1024;
__global__ void __launch_bounds__(256) k_place_chain(SimConst sc, const SchemaDev *__restrict__ sch, PlanDev plan, uint64_t mol0,
                                                     uint32_t m_lo, uint32_t m_hi, const uint64_t *__restrict__ roff_l,
                                                     uint64_t pair_lo, PairTmp *tmp, const uint32_t *__restrict__ acc_i, uint2 *ins, MolRec *mol,
                                                     unsigned long long *err) {
    __shared__ uint32_t s_bins[SORT_BINS];
    __shared__ uint32_t s_warp[9];
    __shared__ uint16_t s_perm[PLACE_BLOCK];
    const uint32_t base = m_lo + blockIdx.x * PLACE_BLOCK;
    const uint32_t cnt = m_hi - base < PLACE_BLOCK ? m_hi - base : PLACE_BLOCK;
    const uint32_t *nf = plan.nfrags + mol0 + base;
    for (uint32_t b = threadIdx.x; b < SORT_BINS; b += 256) s_bins[b] = 0;
    __syncthreads();
    for (uint32_t j = threadIdx.x; j < cnt; j += 256) atomicAdd(&s_bins[nf[j] < SORT_BINS ? nf[j] : SORT_BINS - 1], 1u);
    __syncthreads();
    {   // exclusive scan of the 1024 bins: 4 per thread
        const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        uint32_t v[4], tsum = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            v[q] = s_bins[threadIdx.x * 4 + q];
            tsum += v[q];
        }
        uint32_t inc = tsum;
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += y;
        }
        if (lane == 31) s_warp[warp + 1] = inc;
        if (threadIdx.x == 0) s_warp[0] = 0;
        __syncthreads();
        if (threadIdx.x == 0)
            for (int w = 1; w <= 8; ++w) s_warp[w] += s_warp[w - 1];
        __syncthreads();
        uint32_t ex = inc - tsum + s_warp[warp];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            s_bins[threadIdx.x * 4 + q] = ex;
            ex += v[q];
        }
    }
    __syncthreads();
    for (uint32_t j = threadIdx.x; j < cnt; j += 256) s_perm[atomicAdd(&s_bins[nf[j] < SORT_BINS ? nf[j] : SORT_BINS - 1], 1u)] = (uint16_t)j;
    __syncthreads();
    for (uint32_t j = threadIdx.x; j < cnt; j += 256)
        place_chain_one(sc, sch, plan, mol0, base + s_perm[j], roff_l, pair_lo, tmp, acc_i, ins, mol, err);
}

// one segment: pairs [pair_lo, pair_hi) = the pairs of molecules [m_lo, m_hi)
size_t place_tmp_bytes(uint64_t max_pairs) { return (size_t)(max_pairs + 1) * sizeof(PairTmp); }
// st_draws: the flat per-pair draws; st_chain: the per-molecule chain and the header text (the caller orders the two
// streams with events when they differ: the chain of a segment needs its draws, the draws of segment k + 2 the
// scratch of segment k)
void place_segment_draws(const SimConst &sc, PlanDev plan, uint64_t mol0, uint32_t n_mol, uint64_t pair_lo, uint64_t pair_hi,
                         const uint64_t *d_read_off_local, const uint32_t *d_tile_mol0, uint32_t n_tiles, uint32_t tile_pairs,
                         void *d_tmp, uint32_t *d_acc_i, const InsTable *d_tab, cudaStream_t st) {
    const uint64_t np = pair_hi - pair_lo;
    if (np) LAUNCH(k_insert_draws, (unsigned)((np + 255) / 256), 256, 0, st, sc, plan, mol0, n_mol, d_read_off_local, d_tile_mol0,
                   n_tiles, tile_pairs, pair_lo, pair_hi, (PairTmp *)d_tmp, d_acc_i, d_tab);
}
void place_segment_chain(const SimConst &sc, const SchemaDev *d_sch, PlanDev plan, uint64_t mol0, uint32_t m_lo, uint32_t m_hi,
                         uint64_t pair_lo, const uint64_t *d_read_off_local, void *d_tmp, uint32_t *d_acc_i, uint2 *d_ins, MolRec *d_mol, uint8_t *d_text, uint32_t text_stride,
                         unsigned long long *d_err, cudaStream_t st) {
    if (m_hi <= m_lo) return;
    const uint32_t nm = m_hi - m_lo;
    LAUNCH(k_place_chain, (nm + PLACE_BLOCK - 1) / PLACE_BLOCK, 256, 0, st, sc, d_sch, plan, mol0, m_lo, m_hi, d_read_off_local,
           pair_lo, (PairTmp *)d_tmp, d_acc_i, d_ins, d_mol, d_err);
    LAUNCH(k_mol_text, (nm + TEXT_THREADS - 1) / TEXT_THREADS, TEXT_THREADS, (size_t)TEXT_THREADS * (text_stride + 4u), st, sc, d_sch, mol0, m_lo, m_hi, d_mol, d_text, text_stride);
}

// ------------------------------------------------------------------------------------------
// block placement: launcher (kernels in mmk_place.cuh, included above)
// ------------------------------------------------------------------------------------------
void place_local_roff(uint64_t mol0, uint32_t n_mol, const uint64_t *d_read_off_global, uint64_t *d_read_off_local, cudaStream_t st) {
    LAUNCH(k_local_roff, (n_mol + 1 + 255) / 256, 256, 0, st, d_read_off_global, mol0, n_mol, d_read_off_local);
}
uint32_t place_block_count(uint64_t n_pairs, uint32_t n_mol) { return (uint32_t)((n_pairs + (uint64_t)PB_W * n_mol) / PB_TARGET + 1); }
uint32_t place_big_cap(uint64_t n_pairs) { return (uint32_t)(n_pairs / PB_BIG + 1); }
bool place_block_supports(const SimConst &sc) { return (double)sc.insert_length + 9.0 * sc.ins_sd < 65000.0; }

void place_block(const SimConst &sc, const SchemaDev *d_sch, PlanDev plan, uint64_t mol0, uint32_t n_mol, uint64_t n_pairs,
                 uint32_t n_tiles, const uint64_t *d_roff_l, uint32_t *d_tile_mol0, const InsTable *d_tab, MolRec *d_mol,
                 uint8_t *d_text, uint32_t text_stride, uint4 *d_fdesc, uint32_t *d_sum1, uint32_t *d_sum2, uint8_t *d_tile_flags,
                 unsigned long long *d_gather, unsigned long long *d_err, uint32_t *d_weight, uint64_t *d_woff, uint64_t *d_scan_tmp,
                 uint32_t *d_blk, uint32_t *d_big_list, uint32_t *d_big_count, uint32_t stage_cap, int bpb, cudaStream_t st, cudaStream_t st_text,
                 cudaEvent_t ev_placed, cudaEvent_t ev_text) {
    if (n_mol == 0 || n_pairs == 0) return;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(k_place_block, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PB_SMEM);
        attr_set = true;
    }
    const uint32_t n_blocks = place_block_count(n_pairs, n_mol), big_cap = place_big_cap(n_pairs);
    cudaMemsetAsync(d_sum1, 0, (size_t)n_tiles * 4, st);
    cudaMemsetAsync(d_sum2, 0, (size_t)n_tiles * 4, st);
    cudaMemsetAsync(d_tile_flags, 0, (size_t)n_tiles, st);
    cudaMemsetAsync(d_big_count, 0, 4, st);
    LAUNCH(k_mol_weights, (n_mol + 255) / 256, 256, 0, st, d_roff_l, n_mol, d_tile_mol0, d_weight, d_big_list, d_big_count, big_cap);
    exclusive_scan_u32(d_weight, d_woff, n_mol, d_scan_tmp, st);
    LAUNCH(k_place_blocks, (n_blocks + 1 + 255) / 256, 256, 0, st, d_woff, n_mol, n_blocks, d_blk);
    PlaceArgs a;
    a.sc = sc;
    a.sch = d_sch;
    a.plan = plan;
    a.mol0 = mol0;
    a.n_mol = n_mol;
    a.n_tiles = n_tiles;
    a.n_pairs = n_pairs;
    a.roff = d_roff_l;
    a.woff = d_woff;
    a.blk = d_blk;
    a.n_blocks = n_blocks;
    a.tile_mol0 = d_tile_mol0;
    a.tab = d_tab;
    a.big_list = d_big_list;
    a.big_count = d_big_count;
    a.big_cap = big_cap;
    a.mol = d_mol;
    a.fdesc = d_fdesc;
    a.sum1 = d_sum1;
    a.sum2 = d_sum2;
    a.tile_flags = d_tile_flags;
    a.gather = d_gather;
    a.err = d_err;
    a.bpb = bpb;
    LAUNCH(k_place_big, std::min<uint32_t>(big_cap, 592u), PB_THREADS, 0, st, a);  // exits at once when there are no big molecules
    LAUNCH(k_place_block, n_blocks, PB_THREADS, PB_SMEM, st, a);
    // the header text (memory-bound, 1.5 ms for 20 M molecules) runs on the second stream, beside the tile scans, the chunk table
    // and their host round trips on the main stream; the caller makes the main stream wait for ev_text before anything reads the text
    cudaEventRecord(ev_placed, st);
    cudaStreamWaitEvent(st_text, ev_placed, 0);
    LAUNCH(k_mol_text, (n_mol + TEXT_THREADS - 1) / TEXT_THREADS, TEXT_THREADS, (size_t)TEXT_THREADS * (text_stride + 4u), st_text, sc, d_sch, mol0, 0u, n_mol, d_mol, d_text, text_stride);
    cudaEventRecord(ev_text, st_text);
    LAUNCH(k_tile_flags, (n_tiles + 255) / 256, 256, 0, st, d_sum1, d_sum2, n_tiles, stage_cap, d_tile_flags);
}

void ins_from_fdesc(const uint4 *d_fdesc, const MolRec *d_mol, const uint32_t *d_tile_mol0, uint64_t n_pairs, uint2 *d_ins, cudaStream_t st) {
    if (n_pairs) LAUNCH(k_ins_from_fdesc, (unsigned)((n_pairs + 255) / 256), 256, 0, st, d_fdesc, d_mol, d_tile_mol0, n_pairs, d_ins);
}

__global__ void k_kat_insert_len(const __grid_constant__ SimConst sc, const InsTable *__restrict__ gtab, const uint64_t *__restrict__ bits,
                                 uint32_t *len, uint8_t *slow, int64_t n) {
    __shared__ InsTable tab;
    stage_table(&tab, gtab);
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t x = bits[i];
    uint32_t l = 0;
    uint8_t sl = 0;
    if (!(tab.enabled && ins_len_fast(tab, (uint32_t)(x >> 32), l))) {  // the exact definition; 0 = rejected by the truncation
        sl = 1;
        const double z = mmk_bits_to_normal(x);
        l = (sc.ins_robert || z < sc.ins_lb) ? 0u : insert_len_of(MMK_ADD(sc.ins_mu, MMK_MUL(sc.ins_sd, z)));
    }
    len[i] = l;
    slow[i] = sl;
}
void kat_insert_len(const SimConst &sc, const InsTable *d_tab, const uint64_t *d_bits, uint32_t *d_len, uint8_t *d_slow, int64_t n, cudaStream_t st) {
    if (n > 0) LAUNCH(k_kat_insert_len, (unsigned)((n + 255) / 256), 256, 0, st, sc, d_tab, d_bits, d_len, d_slow, n);
}

// ------------------------------------------------------------------------------------------
// records (decode, generic writer, fast path, tile sums, emit kernel): mmk_emit.cuh
// ------------------------------------------------------------------------------------------
}  // namespace mmkdev
#include "mmk_emit.cuh"
namespace mmkdev {

void tile_prepare(const EmitArgs &a, uint32_t tile_lo, uint32_t tile_hi, uint32_t *d_sum1, uint32_t *d_sum2, uint4 *d_fdesc,
                  uint8_t *d_tile_flags, unsigned long long *d_gather, cudaStream_t st) {
    if (tile_hi <= tile_lo) return;
    const uint32_t per_cta = 256u / a.tile_pairs;
    LAUNCH(k_tile_prepare, (tile_hi - tile_lo + per_cta - 1) / per_cta, 256, 0, st, a, tile_lo, tile_hi, d_sum1, d_sum2, d_fdesc,
           d_tile_flags, d_gather);
}

// Chunk table: tile boundaries so that each chunk's R1 and R2 bytes both fit `cap`.
// bounds[0] = 0, bounds[k+1] = end tile of chunk k; *n_out = number of chunks (0xffffffff if a
// single tile exceeds cap).
// One warp: 32-ary search (each lane probes one pivot) instead of 25 dependent loads per boundary.
__device__ __forceinline__ uint32_t warp_upper_tile(const uint64_t *off, uint32_t lo, uint32_t hi, uint64_t limit, uint32_t lane) {
    // largest t in [lo, hi] with off[t] <= limit, given off[lo] <= limit
    while (hi > lo) {
        const uint32_t span = hi - lo;
        const uint32_t step = (span + 31u) / 32u;
        const uint32_t t = lo + step * (lane + 1u);
        const bool ok = t <= hi && off[t] <= limit;
        const unsigned b = __ballot_sync(0xffffffffu, ok);
        const uint32_t n_ok = __popc(b);  // pivots are increasing, so ok lanes form a prefix
        const uint32_t nlo = lo + step * n_ok;
        uint32_t nhi = lo + step * (n_ok + 1u) - 1u;
        if (nhi > hi) nhi = hi;
        lo = nlo;
        hi = nhi;
    }
    return lo;
}
__global__ void __launch_bounds__(64) k_chunk_table(const uint64_t *off1, const uint64_t *off2, uint32_t n_tiles, uint64_t cap, uint32_t *bounds,
                                                    uint64_t *boff1, uint64_t *boff2, uint32_t max_chunks, uint32_t *n_out) {
    // warp 0 searches the R1 offsets, warp 1 the R2 offsets (each boundary is a chain of dependent probes: the two run side by side)
    __shared__ uint32_t s_e[2];
    const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    const uint64_t *off = w ? off2 : off1;
    uint32_t t0 = 0, k = 0;
    if (threadIdx.x == 0) {
        bounds[0] = 0;
        boff1[0] = off1[0];
        boff2[0] = off2[0];
    }
    while (t0 < n_tiles && k < max_chunks) {
        const uint32_t e = warp_upper_tile(off, t0, n_tiles, off[t0] + cap, lane);
        if (lane == 0) s_e[w] = e;
        __syncthreads();
        const uint32_t best = s_e[0] < s_e[1] ? s_e[0] : s_e[1];
        __syncthreads();
        if (best <= t0) {
            if (threadIdx.x == 0) *n_out = 0xffffffffu;
            return;
        }
        ++k;
        if (threadIdx.x == 0) {
            bounds[k] = best;
            boff1[k] = off1[best];
            boff2[k] = off2[best];
        }
        t0 = best;
    }
    if (threadIdx.x == 0) *n_out = (t0 < n_tiles) ? 0xfffffffeu : k;
}
void chunk_table(const uint64_t *d_off1, const uint64_t *d_off2, uint32_t n_tiles, uint64_t cap, uint32_t *d_bounds,
                 uint64_t *d_boff1, uint64_t *d_boff2, uint32_t max_chunks, uint32_t *d_n, cudaStream_t st) {
    LAUNCH(k_chunk_table, 1, 64, 0, st, d_off1, d_off2, n_tiles, cap, d_bounds, d_boff1, d_boff2, max_chunks, d_n);
}


void emit(const EmitArgs &a, cudaStream_t st) {
    if (a.n_tiles == 0) return;
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    size_t smem = 2 * (size_t)a.stage_cap;
    if (const char *e = getenv("MMK_EMIT_PAD")) smem += (size_t)atoi(e);  // occupancy probe (profiling only)
    if (!attr_set[dev & 63]) {
        cudaFuncSetAttribute(k_emit<2, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        cudaFuncSetAttribute(k_emit<4, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        cudaFuncSetAttribute(k_emit<2, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        cudaFuncSetAttribute(k_emit<4, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        cudaFuncSetAttribute(k_emit_w<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        cudaFuncSetAttribute(k_emit_w<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        cudaFuncSetAttribute(k_emit_w<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        cudaFuncSetAttribute(k_emit_w<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        attr_set[dev & 63] = true;
    }
    // 32-pair tiles: k_emit_w (stage fill by bulk copy); MMK_EMIT_MODE=cta keeps k_emit (fill by shared-memory stores)
    static const bool fill_by_copy = !(getenv("MMK_EMIT_MODE") && strcmp(getenv("MMK_EMIT_MODE"), "cta") == 0);
    if (a.tile_pairs == 32 && a.qfill && (fill_by_copy || a.err_on)) {
        if (a.err_on) {  // the error model's variant: the plain kernels keep their instruction count
            if (a.bpb == 2) launch_overlapped(k_emit_w<2, true>, a.n_tiles, 128, smem, st, a);
            else launch_overlapped(k_emit_w<4, true>, a.n_tiles, 128, smem, st, a);
        } else if (a.bpb == 2) launch_overlapped(k_emit_w<2, false>, a.n_tiles, 128, smem, st, a);
        else launch_overlapped(k_emit_w<4, false>, a.n_tiles, 128, smem, st, a);
        return;
    }
    EmitArgs b = a;
    if (a.err_on) b.fast_ok = 0;  // other tile sizes: every record through the general writer, which applies the model
    // one thread per (pair, header piece): 64-pair tiles run 256 threads, 32-pair tiles 128
    if (a.tile_pairs > 32) {
        if (a.bpb == 2) LAUNCH((k_emit<2, 256>), a.n_tiles, 256, smem, st, b);
        else LAUNCH((k_emit<4, 256>), a.n_tiles, 256, smem, st, b);
    } else {
        if (a.bpb == 2) LAUNCH((k_emit<2, 128>), a.n_tiles, 128, smem, st, b);
        else LAUNCH((k_emit<4, 128>), a.n_tiles, 128, smem, st, b);
    }
}

// ------------------------------------------------------------------------------------------
// BGZF stage
// ------------------------------------------------------------------------------------------
#include "mmk_bgzf.cuh"
void bgzf_geometry(BgzfArgs &a) {
    a.out_cap = (a.in_cap / 3 + 64 + 15) & ~15u;  // members that do not shrink 3x are stored
    a.crc_words = (a.in_cap / 4 + 2 + 31) / 32;
    size_t w = ((size_t)a.in_cap + 32) + a.out_cap + (4 * (size_t)a.max_records + 4) * 4 + 2 * ((size_t)a.max_records + 4) * 4;
    a.warp_smem = (uint32_t)((w + 15) & ~(size_t)15);
}
void bgzf_encode(const BgzfArgs &a, uint32_t n_blocks, cudaStream_t st) {
    if (!n_blocks) return;
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_set[dev & 63]) {
        cudaFuncSetAttribute(k_bgzf, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        attr_set[dev & 63] = true;
    }
    const size_t smem = ((BZ_T_SHIFT * 4 + 15) & ~(size_t)15) + (size_t)BZ_WARPS * a.warp_smem;
    LAUNCH(k_bgzf, (n_blocks + BZ_WARPS - 1) / BZ_WARPS, BZ_THREADS, smem, st, a, n_blocks);
}
void bgzf_compact(const BgzfArgs &a, const uint64_t *offs, uint8_t *dst, uint32_t n_blocks, unsigned long long *host_total,
                  cudaStream_t st) {
    if (!n_blocks) return;
    LAUNCH(k_bgzf_compact, (n_blocks + 7) / 8, 256, 0, st, a, n_blocks, offs, dst, host_total);
}
uint64_t bgzf_slots_bytes(uint64_t text_bytes, uint32_t n_blocks, uint32_t slot_shift) {
    return text_bytes + (text_bytes >> slot_shift) + 64ull * n_blocks + 512;
}

// ------------------------------------------------------------------------------------------
// variate dump (replay stream for the oracle)
// ------------------------------------------------------------------------------------------
template <class R>
__device__ __forceinline__ void dump_molecule(const SimConst &sc, const SchemaDev *sch, const PlanDev &plan, uint64_t i,
                                              const uint64_t *roff, uint2 *ins, const uint32_t *epoch_nlive,
                                              const uint32_t *tile_mol0, uint32_t tile_pairs, R &rec) {
    uint64_t g = plan.group[i];
    uint32_t m = plan.slot[i];
    uint32_t seq_len = sch[plan.schema[i]].len;
    if (m == 0) {
        uint32_t n = draw_nmol(sc, g, rec);
        uint32_t n_live = epoch_nlive[plan.epoch[i]];
        for (uint32_t j = 0; j < n; ++j) rec.put(MMK_P_TARGET, n_live, draw_target(sc, g, j, n_live));
    }
    uint32_t molsize, nfrags;
    molecule_shape(sc, g, m, seq_len, molsize, nfrags, rec);
    int64_t start;
    TileRef tr{tile_mol0, roff[i], (uint32_t)i, tile_pairs};
    molecule_place(sc, g, m, seq_len, molsize, nfrags, ins + roff[i], tr, start, rec);
}
__global__ void k_dump_count(SimConst sc, const SchemaDev *sch, PlanDev plan, uint64_t n_mol, const uint64_t *roff, uint2 *ins,
                             const uint32_t *epoch_nlive, const uint32_t *tile_mol0, uint32_t tile_pairs, uint32_t *counts) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_mol) return;
    CountRec rec;
    dump_molecule(sc, sch, plan, i, roff, ins, epoch_nlive, tile_mol0, tile_pairs, rec);
    counts[i] = rec.n;
}
__global__ void k_dump_write(SimConst sc, const SchemaDev *sch, PlanDev plan, uint64_t n_mol, const uint64_t *roff, uint2 *ins,
                             const uint32_t *epoch_nlive, const uint32_t *tile_mol0, uint32_t tile_pairs,
                             const uint64_t *offsets, mmk_variate *out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_mol) return;
    WriteRec rec{out + offsets[i]};
    dump_molecule(sc, sch, plan, i, roff, ins, epoch_nlive, tile_mol0, tile_pairs, rec);
}
void dump_variates(const SimConst &sc, const SchemaDev *d_sch, PlanDev plan, uint64_t n_mol, const uint64_t *d_read_off,
                   const uint2 *d_ins, const uint32_t *d_epoch_nlive, const uint32_t *d_tile_mol0, uint32_t tile_pairs,
                   uint32_t *d_counts, uint64_t *d_offsets, uint64_t *d_scan_tmp, mmk_variate *d_out, int pass,
                   cudaStream_t st) {
    if (n_mol == 0) return;
    unsigned g = (unsigned)((n_mol + 127) / 128);
    if (pass == 0) {
        LAUNCH(k_dump_count, g, 128, 0, st, sc, d_sch, plan, n_mol, d_read_off, (uint2 *)d_ins, d_epoch_nlive, d_tile_mol0, tile_pairs, d_counts);
        exclusive_scan_u32(d_counts, d_offsets, n_mol, d_scan_tmp, st);
    } else {
        LAUNCH(k_dump_write, g, 128, 0, st, sc, d_sch, plan, n_mol, d_read_off, (uint2 *)d_ins, d_epoch_nlive, d_tile_mol0, tile_pairs, d_offsets, d_out);
    }
}

__global__ void k_mask_flags(const char *raw, int64_t len, uint32_t *flags) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < len) flags[i] = raw[i] == 'N';
}
__global__ void k_mask_write(const char *raw, int64_t len, uint64_t seed, uint32_t contig, uint32_t hap, const uint64_t *off,
                             mmk_variate *out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= len || raw[i] != 'N') return;
    mmk_stream s;
    mmk_stream_init(&s, seed, 0, MMK_P_MASK, contig, hap);
    mmk_variate v;
    v.tag = MMK_P_MASK;
    v.aux = 4;
    v.v.i = (int64_t)mmk_bits_to_index(mmk_stream_at(&s, (uint32_t)i), 4);
    out[off[i]] = v;
}
void dump_mask(const char *d_raw_upper, int64_t len, uint64_t seed, uint32_t contig, uint32_t hap, uint32_t *d_flags,
               uint64_t *d_off, uint64_t *d_scan_tmp, mmk_variate *d_out, int pass, cudaStream_t st) {
    if (len <= 0) return;
    unsigned g = (unsigned)((len + 255) / 256);
    if (pass == 0) {
        LAUNCH(k_mask_flags, g, 256, 0, st, d_raw_upper, len, d_flags);
        exclusive_scan_u32(d_flags, d_off, (uint64_t)len, d_scan_tmp, st);
    } else {
        LAUNCH(k_mask_write, g, 256, 0, st, d_raw_upper, len, seed, contig, hap, d_off, d_out);
    }
}

// ------------------------------------------------------------------------------------------
// known-answer hooks
// ------------------------------------------------------------------------------------------
__global__ void k_kat_philox(const uint32_t *in6, uint32_t *out4) { mmk_philox4x32_10(in6, in6 + 4, out4); }
void kat_philox(const uint32_t *d_in6, uint32_t *d_out4, cudaStream_t st) { LAUNCH(k_kat_philox, 1, 1, 0, st, d_in6, d_out4); }

__global__ void k_kat_math(int kind, const double *in, double *out, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x = in[i];
    if (kind == 0) out[i] = mmk_log(x);
    else if (kind == 1) out[i] = mmk_exp(x);
    else if (kind == 2) out[i] = mmk_bits_to_exp(mmk_d2u(x));
    else {
        mmk_stream s;
        mmk_insert_stream_init(&s, mmk_d2u(x), 0, (uint64_t)i, 0, 0);
        out[i] = mmk_next_normal(&s);
    }
}
void kat_math(int kind, const double *d_in, double *d_out, int64_t n, cudaStream_t st) {
    if (n > 0) LAUNCH(k_kat_math, (unsigned)((n + 127) / 128), 128, 0, st, kind, d_in, d_out, n);
}

}  // namespace mmkdev
