// mmk_bgzf.cuh — BGZF (blocked gzip) coding of emitted FASTQ text on the device.
//
// The reference writes every FASTQ through BGZFCompressorStream (src/Main.jl:42-43,111-112); its files are
// `<prefix>.R{1,2}.fq.gz`.  Here one CTA turns one block of whole records (the tiles the emit kernel wrote) into
// one BGZF member: 18-byte gzip header with the BC extra field, a single deflate block, CRC-32 and ISIZE.
//
// Deflate side.  Every block uses the same pre-agreed dynamic Huffman code (mmk_bgzf_table.h, fitted offline by
// scripts/bgzf_table.py), so the block prefix is a constant bit string.  The LZ parse is FASTQ-aware instead of
// hash-based — scripts/bgzf_table.py:tokenize() is its specification:
//   header line   common prefix and common suffix with the previous record's header -> two matches, rest literals
//   sequence line literals (A/C/G/T cost 2-3 bits)
//   "+" and quality lines: identical to the previous record's -> one match; else a run (literal + distance-1 match)
// Every match is verified by comparing bytes, so any 4-line text is coded correctly; text that is not 4-line
// records, bytes without a code, or a block that would grow are written as a stored deflate block.
// CRC-32: per-thread table-driven CRC over chunks aligned to the END of the block (leading zeros leave a zero
// state unchanged), combined pairwise with carry-less multiplications by x^(8·len) mod P.
#pragma once
#include "mmk_internal.h"

constexpr uint32_t BZ_THREADS = 256;
constexpr uint32_t BZ_GZ_HDR = 18;
constexpr uint32_t BZ_MIN_MATCH = 4, BZ_MAX_MATCH = 258;

__device__ __forceinline__ uint32_t bz_gf_mul(uint32_t a, uint32_t b) {  // a·b mod P, reflected bit order
    uint32_t p = 0;
#pragma unroll 4
    for (int i = 0; i < 32; ++i) {
        p ^= (a & 0x80000000u) ? b : 0u;
        a <<= 1;
        b = (b >> 1) ^ ((b & 1u) ? 0xEDB88320u : 0u);
    }
    return p;
}

struct BzSmem {
    uint32_t *ll;      // [286] reversed code | len << 16
    uint32_t *dd;      // [30]
    uint32_t *crc_t;   // [256]
    uint32_t *out;     // zeroed words
    uint8_t *in;       // staged text (same 16-byte phase as in global memory)
    uint32_t *ls;      // line starts
    uint32_t *rec_pos; // bit position of every record's first token
};

__device__ __forceinline__ void bz_put_bits(uint32_t *out, uint32_t pos, uint64_t acc, uint32_t nb) {
    if (!nb) return;
    const uint32_t w = pos >> 5, s = pos & 31;
    const uint64_t lo = acc << s;
    atomicOr(&out[w], (uint32_t)lo);
    const uint32_t w1 = (uint32_t)(lo >> 32);
    if (w1) atomicOr(&out[w + 1], w1);
    const uint32_t hi = s ? (uint32_t)(acc >> (64 - s)) : 0u;
    if (hi) atomicOr(&out[w + 2], hi);
}

// warp-cooperative: number of leading positions where a[i] == b[i], i < m
__device__ __forceinline__ uint32_t bz_common_prefix(const uint8_t *a, const uint8_t *b, uint32_t m, uint32_t lane) {
    for (uint32_t i0 = 0; i0 < m; i0 += 32) {
        const uint32_t i = i0 + lane;
        const bool ne = i < m ? a[i] != b[i] : true;
        const uint32_t bal = __ballot_sync(0xffffffffu, ne);
        if (bal) return min(i0 + (uint32_t)__ffs(bal) - 1u, m);
    }
    return m;
}
// ... trailing positions where a_end[-1-i] == b_end[-1-i]
__device__ __forceinline__ uint32_t bz_common_suffix(const uint8_t *a_end, const uint8_t *b_end, uint32_t m, uint32_t lane) {
    for (uint32_t i0 = 0; i0 < m; i0 += 32) {
        const uint32_t i = i0 + lane;
        const bool ne = i < m ? a_end[-1 - (int)i] != b_end[-1 - (int)i] : true;
        const uint32_t bal = __ballot_sync(0xffffffffu, ne);
        if (bal) return min(i0 + (uint32_t)__ffs(bal) - 1u, m);
    }
    return m;
}

// Literals in[a, b): returns the bits they take; EMIT writes them at `pos`.  A byte without a code sets *bad.
template <bool EMIT>
__device__ __forceinline__ uint32_t bz_literals(const BzSmem &S, uint32_t a, uint32_t b, uint32_t pos, uint32_t lane, uint32_t *bad) {
    uint32_t total = 0;
    for (uint32_t c = a; c < b; c += 160) {
        uint64_t acc = 0;
        uint32_t nb = 0;
        const uint32_t j0 = c + lane * 5;
#pragma unroll
        for (uint32_t k = 0; k < 5; ++k) {
            const uint32_t j = j0 + k;
            if (j < b) {
                const uint32_t e = S.ll[S.in[j]];
                const uint32_t l = e >> 16;
                if (!l) *bad = 1;
                if (EMIT) acc |= (uint64_t)(e & 0xffffu) << nb;
                nb += l;
            }
        }
        if (EMIT) {
            uint32_t inc = nb;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
                if ((int)lane >= o) inc += y;
            }
            bz_put_bits(S.out, pos + total + inc - nb, acc, nb);
            total += __shfl_sync(0xffffffffu, inc, 31);
        } else
            total += __reduce_add_sync(0xffffffffu, nb);
    }
    return total;
}

// One match token (length 3..258, distance 1..32768); lane 0 writes.
template <bool EMIT>
__device__ __forceinline__ uint32_t bz_match(const BzSmem &S, uint32_t len, uint32_t dist, uint32_t pos, uint32_t lane) {
    uint32_t sym, eb, ev;
    const uint32_t l = len - 3;
    if (len == 258) {
        sym = 285, eb = 0, ev = 0;
    } else if (l < 8) {
        sym = 257 + l, eb = 0, ev = 0;
    } else {
        const uint32_t k = 31 - __clz(l);
        eb = k - 2;
        sym = 257 + 4 * eb + 4 + ((l - (1u << k)) >> eb);
        ev = (l - (1u << k)) & ((1u << eb) - 1);
    }
    const uint32_t e1 = S.ll[sym];
    uint32_t dsym, deb, dev;
    const uint32_t d = dist - 1;
    if (d < 4) {
        dsym = d, deb = 0, dev = 0;
    } else {
        const uint32_t k = 31 - __clz(d);
        deb = k - 1;
        dsym = 2 * k + ((d >> deb) & 1);
        dev = d & ((1u << deb) - 1);
    }
    const uint32_t e2 = S.dd[dsym];
    const uint32_t n1 = e1 >> 16, n2 = e2 >> 16;
    if (EMIT && lane == 0) {
        uint64_t acc = e1 & 0xffffu;
        uint32_t nb = n1;
        acc |= (uint64_t)ev << nb;
        nb += eb;
        acc |= (uint64_t)(e2 & 0xffffu) << nb;
        nb += n2;
        acc |= (uint64_t)dev << nb;
        nb += deb;
        bz_put_bits(S.out, pos, acc, nb);
    }
    return n1 + eb + n2 + deb;
}

// A match longer than 258 is cut into pieces of 3..258 (scripts/bgzf_table.py:split_match)
template <bool EMIT>
__device__ __forceinline__ uint32_t bz_long_match(const BzSmem &S, uint32_t total, uint32_t dist, uint32_t pos, uint32_t lane) {
    uint32_t bits = 0, rem = total;
    while (rem) {
        uint32_t take = min(BZ_MAX_MATCH, rem);
        if (rem - take > 0 && rem - take < 3) take = rem - 3;
        bits += bz_match<EMIT>(S, take, dist, pos + bits, lane);
        rem -= take;
    }
    return bits;
}

// All tokens of record r (warp-cooperative); returns their bits.
template <bool EMIT>
__device__ __forceinline__ uint32_t bz_record(const BzSmem &S, uint32_t r, uint32_t pos, uint32_t lane, uint32_t *bad) {
    const uint32_t h0 = S.ls[4 * r], s0 = S.ls[4 * r + 1], p0 = S.ls[4 * r + 2], q0 = S.ls[4 * r + 3], e0 = S.ls[4 * r + 4];
    const uint8_t *in = S.in;
    uint32_t bits = 0;
    uint32_t P = 0, Sx = 0, hp0 = 0, sp0 = 0, pp0 = 0;
    if (r) {
        hp0 = S.ls[4 * r - 4], sp0 = S.ls[4 * r - 3], pp0 = S.ls[4 * r - 2];
        const uint32_t lh = s0 - h0, lhp = sp0 - hp0;
        const uint32_t mh = min(lh, lhp);
        P = bz_common_prefix(in + h0, in + hp0, min(mh, BZ_MAX_MATCH), lane);
        if (P < BZ_MIN_MATCH) P = 0;
        Sx = bz_common_suffix(in + s0, in + sp0, min(mh - P, BZ_MAX_MATCH), lane);
        if (Sx < BZ_MIN_MATCH) Sx = 0;
    }
    if (P) bits += bz_match<EMIT>(S, P, h0 - hp0, pos + bits, lane);
    bits += bz_literals<EMIT>(S, h0 + P, s0 - Sx, pos + bits, lane, bad);
    if (Sx) bits += bz_match<EMIT>(S, Sx, s0 - sp0, pos + bits, lane);
    bits += bz_literals<EMIT>(S, s0, p0, pos + bits, lane, bad);
    const uint32_t lq = e0 - p0;
    bool same = false;
    if (r && h0 - pp0 == lq) same = bz_common_prefix(in + p0, in + pp0, lq, lane) == lq;
    if (same) {
        bits += bz_long_match<EMIT>(S, lq, p0 - pp0, pos + bits, lane);
    } else {
        bits += bz_literals<EMIT>(S, p0, q0, pos + bits, lane, bad);
        const uint32_t nq = e0 - 1 - q0;
        bool run = false;
        if (nq >= BZ_MIN_MATCH) run = bz_common_prefix(in + q0, in + q0 + 1, nq - 1, lane) == nq - 1;
        if (run) {
            bits += bz_literals<EMIT>(S, q0, q0 + 1, pos + bits, lane, bad);
            bits += bz_long_match<EMIT>(S, nq - 1, 1, pos + bits, lane);
            bits += bz_literals<EMIT>(S, e0 - 1, e0, pos + bits, lane, bad);
        } else
            bits += bz_literals<EMIT>(S, q0, e0, pos + bits, lane, bad);
    }
    return bits;
}

// Output slot of block b whose text starts at chunk-relative offset g0.  A member codes up to in_cap bytes of text
// into at most that many + 31 bytes, so with 2^slot_shift <= in_cap / 32 the text offset + offset >> slot_shift +
// 64 bytes per block (rounded to 16) leaves every block its worst case, however many members it needs.
__device__ __forceinline__ uint64_t bz_slot_off(uint64_t g0, uint32_t b, uint32_t shift) {
    return (g0 + (g0 >> shift) + 64ull * b + 15) & ~15ull;
}

// One BGZF member from text[0, n) (n <= in_cap) into dst; returns its size.  `may_code` false forces a stored block.
__device__ __forceinline__ uint32_t bz_member(const BgzfArgs &a, const BzSmem &S0, uint8_t *in_base, uint32_t out_words,
                                              const uint8_t *text, uint32_t n, bool may_code, uint8_t *dst) {
    __shared__ uint32_t s_warp[BZ_THREADS / 32];
    __shared__ uint32_t s_bad, s_total_bits, s_crc;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr uint32_t NW = BZ_THREADS / 32;
    BzSmem S = S0;
    const uint32_t phase = (uint32_t)((uintptr_t)text & 15);
    S.in = in_base + phase;
    for (uint32_t i = tid; i < out_words; i += BZ_THREADS) S.out[i] = 0;
    {
        const uint4 *src = (const uint4 *)(text - phase);
        const uint32_t nv = (phase + n + 15) / 16;
        for (uint32_t i = tid; i < nv; i += BZ_THREADS) ((uint4 *)in_base)[i] = src[i];
    }
    if (tid == 0) s_bad = 0;
    __syncthreads();
    const uint8_t *in = S.in;

    // pass 1: CRC of the thread's chunk (chunks are aligned to the end of the block) and its newline count
    const int C = (int)a.crc_chunk;
    const int c_lo = (int)n - (int)(BZ_THREADS - tid) * C, c_hi = c_lo + C;
    uint32_t crc = 0, nl = 0;
    for (int i = max(c_lo, 0); i < c_hi; ++i) {
        uint32_t b = in[i];
        nl += b == 10;
        if (i < 4) b ^= 0xffu;  // the 0xFFFFFFFF preset folded into the first four bytes
        crc = S.crc_t[(crc ^ b) & 0xffu] ^ (crc >> 8);
    }
    // newline ranks
    uint32_t inc = nl;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
        if ((int)lane >= o) inc += y;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t before = inc - nl, n_lines = 0;
#pragma unroll
    for (uint32_t w = 0; w < NW; ++w) {
        const uint32_t v = s_warp[w];
        if (w < warp) before += v;
        n_lines += v;
    }
    const uint32_t n_rec = n_lines >> 2;
    const bool text_ok = may_code && n >= 4 && (n_lines & 3) == 0 && n_rec <= a.max_records && in[n - 1] == 10;
    if (text_ok) {
        if (tid == 0) S.ls[0] = 0;
        for (int i = max(c_lo, 0); i < c_hi; ++i)
            if (in[i] == 10) S.ls[++before] = (uint32_t)i + 1;
    }
    // CRC tree: v[t] covers chunks [t, t + 2^k); right neighbour is 2^k chunks long
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const uint32_t right = __shfl_down_sync(0xffffffffu, crc, 1u << k);
        crc = bz_gf_mul(crc, a.crc_shift[k]) ^ right;
    }
    __syncthreads();  // s_warp reuse, ls complete
    if (lane == 0) s_warp[warp] = crc;
    __syncthreads();
    if (warp == 0) {
        uint32_t v = lane < NW ? s_warp[lane] : 0;
#pragma unroll
        for (int k = 5; k < 8; ++k) {
            const uint32_t right = __shfl_down_sync(0xffffffffu, v, 1u << (k - 5));
            v = bz_gf_mul(v, a.crc_shift[k]) ^ right;
        }
        if (lane == 0) s_crc = n >= 4 ? ~v : 0u;
    }
    if (n < 4 && tid == 0) {  // too short for the folded preset: plain bytewise CRC
        uint32_t c = 0xffffffffu;
        for (uint32_t i = 0; i < n; ++i) c = S.crc_t[(c ^ in[i]) & 0xffu] ^ (c >> 8);
        s_crc = ~c;
    }

    // pass 2: bits of every record
    uint32_t bad = 0;
    if (text_ok) {
        for (uint32_t r = warp; r < n_rec; r += NW) {
            const uint32_t bits = bz_record<false>(S, r, 0, lane, &bad);
            if (lane == 0) S.rec_pos[r] = bits;
        }
    }
    if (bad) s_bad = 1;
    __syncthreads();
    const uint32_t base_bit = BZ_GZ_HDR * 8 + a.hdr_bits;
    if (warp == 0 && text_ok) {  // exclusive scan of the record sizes
        uint32_t carry = base_bit;
        for (uint32_t r0 = 0; r0 < n_rec; r0 += 32) {
            const uint32_t r = r0 + lane;
            const uint32_t v = r < n_rec ? S.rec_pos[r] : 0;
            uint32_t x = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
                if ((int)lane >= o) x += y;
            }
            if (r < n_rec) S.rec_pos[r] = carry + x - v;
            carry += __shfl_sync(0xffffffffu, x, 31);
        }
        if (lane == 0) s_total_bits = carry;
    }
    __syncthreads();
    const uint32_t eob = S.ll[256];
    const uint32_t total_bits = text_ok ? s_total_bits : 0u;
    const uint32_t end_bit = total_bits + (eob >> 16);
    const uint32_t coded_bytes = (end_bit + 7) / 8 - BZ_GZ_HDR;
    const bool coded = text_ok && !s_bad && coded_bytes < n + 5;
    uint32_t deflate_bytes;
    if (coded) {
        deflate_bytes = coded_bytes;
        // constant block prefix
        const uint32_t *hdr = a.tab + BGZF_TAB_CODES;
        for (uint32_t i = tid; i * 32 < a.hdr_bits; i += BZ_THREADS)
            bz_put_bits(S.out, BZ_GZ_HDR * 8 + i * 32, hdr[i], min(32u, a.hdr_bits - i * 32));
        for (uint32_t r = warp; r < n_rec; r += NW) bz_record<true>(S, r, S.rec_pos[r], lane, &bad);
        if (tid == 0) bz_put_bits(S.out, total_bits, eob & 0xffffu, eob >> 16);
    } else {
        deflate_bytes = n + 5;
        uint8_t *o = (uint8_t *)S.out + BZ_GZ_HDR;
        if (tid == 0) {
            o[0] = 1;
            o[1] = (uint8_t)n, o[2] = (uint8_t)(n >> 8);
            o[3] = (uint8_t)~n, o[4] = (uint8_t)(~n >> 8);
        }
        for (uint32_t i = tid; i < n; i += BZ_THREADS) o[5 + i] = in[i];
    }
    const uint32_t total = BZ_GZ_HDR + deflate_bytes + 8;
    __syncthreads();
    if (tid == 0) {
        uint8_t *o = (uint8_t *)S.out;
        const uint8_t gz[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
        for (int i = 0; i < 16; ++i) o[i] = gz[i];
        o[16] = (uint8_t)(total - 1), o[17] = (uint8_t)((total - 1) >> 8);
        uint8_t *t = o + BZ_GZ_HDR + deflate_bytes;
        const uint32_t c = s_crc;
        t[0] = (uint8_t)c, t[1] = (uint8_t)(c >> 8), t[2] = (uint8_t)(c >> 16), t[3] = (uint8_t)(c >> 24);
        t[4] = (uint8_t)n, t[5] = (uint8_t)(n >> 8), t[6] = (uint8_t)(n >> 16), t[7] = (uint8_t)(n >> 24);
    }
    __syncthreads();
    if (((uintptr_t)dst & 15) == 0) {
        const uint4 *src = (const uint4 *)S.out;
        uint4 *d4 = (uint4 *)dst;
        for (uint32_t i = tid; i < (total + 15) / 16; i += BZ_THREADS) d4[i] = src[i];
    } else {  // later members of an oversized block
        const uint8_t *src = (const uint8_t *)S.out;
        for (uint32_t i = tid; i < total; i += BZ_THREADS) dst[i] = src[i];
    }
    __syncthreads();
    return total;
}

__global__ void __launch_bounds__(BZ_THREADS) k_bgzf(BgzfArgs a) {
    extern __shared__ __align__(16) uint8_t bz_smem[];
    const uint32_t tid = threadIdx.x;
    const uint32_t tb = a.tile0 + blockIdx.x * a.tiles_per_block;
    const uint32_t te = min(tb + a.tiles_per_block, a.tile0 + a.n_tiles);
    const uint64_t g0 = a.tile_off[tb] - a.chunk_off, g1 = a.tile_off[te] - a.chunk_off;
    const uint64_t n = g1 - g0;
    uint8_t *slot = a.slots + bz_slot_off(g0, blockIdx.x, a.slot_shift);

    // carve shared memory
    BzSmem S;
    uint32_t off = 0;
    S.ll = (uint32_t *)(bz_smem + off), off += 286 * 4;
    S.dd = (uint32_t *)(bz_smem + off), off += 30 * 4;
    S.crc_t = (uint32_t *)(bz_smem + off), off += 256 * 4;
    S.ls = (uint32_t *)(bz_smem + off), off += (4 * a.max_records + 4) * 4;
    S.rec_pos = (uint32_t *)(bz_smem + off), off += (a.max_records + 4) * 4;
    off = (off + 15) & ~15u;
    const uint32_t out_words = (a.in_cap + 64) / 4;
    S.out = (uint32_t *)(bz_smem + off), off += out_words * 4;
    uint8_t *in_base = bz_smem + off;
    S.in = in_base;
    for (uint32_t i = tid; i < BGZF_TAB_CODES; i += BZ_THREADS) S.ll[i] = a.tab[i];  // ll, dd, crc_t are contiguous

    // Normally the block is one member.  Records longer than the tile bound (the reference's literal circular wrap,
    // Breakpoints.jl:111-121) can make a tile arbitrarily large: it is then cut into stored members of in_cap bytes.
    uint32_t total = 0;
    if (n <= a.in_cap) {
        total = bz_member(a, S, in_base, out_words, a.text + g0, (uint32_t)n, true, slot);
    } else {
        for (uint64_t p = 0; p < n; p += a.in_cap) {
            const uint32_t np = (uint32_t)min((uint64_t)a.in_cap, n - p);
            total += bz_member(a, S, in_base, out_words, a.text + g0 + p, np, false, slot + total);
        }
    }
    if (tid == 0) a.sizes[blockIdx.x] = total;
}

// Packs the coded blocks back to back: block b goes to dst[offs[b] ...).
__global__ void __launch_bounds__(256) k_bgzf_compact(BgzfArgs a, const uint64_t *__restrict__ offs, uint8_t *__restrict__ dst) {
    const uint32_t tb = a.tile0 + blockIdx.x * a.tiles_per_block;
    const uint8_t *src = a.slots + bz_slot_off(a.tile_off[tb] - a.chunk_off, blockIdx.x, a.slot_shift);
    const uint32_t n = a.sizes[blockIdx.x];
    uint8_t *d = dst + offs[blockIdx.x];
    // head bytes up to 4-byte alignment of the destination, then realigned words
    const uint32_t head = min(n, (uint32_t)((4 - ((uintptr_t)d & 3)) & 3));
    if (threadIdx.x < head) d[threadIdx.x] = src[threadIdx.x];
    const uint32_t nw = (n - head) / 4;
    const uint32_t sh = head * 8;
    const uint32_t *sw = (const uint32_t *)src;  // word i of the destination = bytes head + 4 i .. of src
    uint32_t *dw = (uint32_t *)(d + head);
    for (uint32_t i = threadIdx.x; i < nw; i += 256) {
        const uint32_t lo = sw[i], hi = sw[i + 1];
        dw[i] = sh ? __funnelshift_r(lo, hi, sh) : lo;
    }
    const uint32_t done = head + nw * 4;
    if (threadIdx.x < n - done) d[done + threadIdx.x] = src[done + threadIdx.x];
}
