// mmk_internal.h — types shared by the host driver (mmk_host.cpp) and the kernels (mmk_device.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "mimick_b200.h"
#include "mmk_rng.h"

// Truncated Exponential as Distributions.jl builds it (truncated.jl; restated in DESIGN.md §3).
struct TruncExp {
    double theta, lower, upper;
    int32_t has_upper, pad;
    double loglcdf, lcdf, tp, logtp;
};

// Everything the sampling kernels need, frozen at mmk_create (SimParams, src/Structs.jl:54-70).
struct SimConst {
    uint64_t seed;
    uint32_t sample;
    int32_t nmol_fixed;  // >0: fixed molecules per barcode group; 0: draw from `nmol`
    TruncExp mollen;     // truncated(Exponential(mol_length), lower = lowerbound)
    TruncExp nmol;       // truncated(Exponential(N), 1, 2.5N)
    double ins_mu, ins_sd, ins_lb;  // truncated(Normal(mu, sd), lower = 200), lb standardised
    double ins_a, ins_theta;        // Robert's exponential proposal (kept mass <= 0.3): a = (lb + sqrt(lb^2 + 4)) / 2, theta = 1 / a
    int32_t ins_robert, pad0;       // 1: kept mass tp <= 0.3 (Distributions.jl randnt switches samplers there)
    double singletons, mol_coverage;
    int32_t insert_length, attempts;
    int32_t r1, r2;
    int32_t circular, strict;
    int32_t fq_format, bc_format;
};

// Insert-length table (mmk_place.cuh): trunc(mu + sd * Phi^-1) of the top 32 bits of a draw by search among thresholds.
constexpr int INS_CELL_BITS = 9, INS_CELLS = 1 << INS_CELL_BITS, INS_MAX_THR = 1024;
struct InsTable {
    uint32_t enabled;   // 0: every draw takes the double-precision path
    uint32_t n_thr;     // thresholds in thr[]
    uint32_t base_len;  // length of a draw below thr[0]
    uint32_t u_lo, u_hi;  // table valid for top-32 draws in [u_lo, u_hi] (accepted by the truncation, inside the tabulated range)
    uint32_t pad[3];
    uint32_t thr[INS_MAX_THR];     // top 32 of the 52 bits of the smallest draw whose length is base_len + k + 1, ascending
    uint16_t cell[INS_CELLS + 8];  // cell[c] = #thresholds below c << (32 - INS_CELL_BITS); cell[INS_CELLS] = n_thr
};
static_assert(sizeof(InsTable) % 16 == 0, "InsTable is staged with 16-byte copies");

// One (contig, haplotype) unit as the kernels see it.
struct SchemaDev {
    uint64_t word_off;  // first 32-bit word of the packed haplotype in the arena
    const char *name;   // contig name (device memory, not terminated; the names of all schemas share one buffer)
    uint32_t len;       // bases
    uint16_t haplotype;
    uint16_t name_len;
};
static_assert(sizeof(SchemaDev) == 24, "SchemaDev layout");

// Per-molecule record written by the placement kernel, read by the emit kernel.
struct MolRec {
    int64_t start;      // 1-based molecule start
    uint32_t size;      // molecule length (position = start : start+size)
    uint16_t schema;
    uint8_t prefix_len; // "@chrom:hap|start:end|"
    uint8_t suffix_len; // format-specific tail of the header line incl. '\n', read digit = '1'
    uint32_t qual1, qual2;  // quality-string lengths (length of the molecule's first R1 / R2)
};
static_assert(sizeof(MolRec) == 24, "MolRec layout");

// Plan arrays (SoA, device). Index = run-order molecule index.
struct PlanDev {
    uint64_t *group;
    uint32_t *slot;
    uint16_t *schema;
    uint16_t *epoch;
    uint32_t *size;
    uint32_t *nfrags;
    uint8_t *status;
};

// Control block of the planner, in device memory (mmk_plan.cuh); the host reads it back once per round of batches.
struct PlanState {
    uint64_t g0;              // first barcode group of the next batch
    uint64_t n_mol;           // molecules committed so far
    uint64_t g_tab;           // groups the group table covers (host-maintained)
    uint64_t B;               // groups in the next batch
    uint64_t Bmin, Bmax;
    uint64_t barcode_cap;     // barcodes the enumerator can produce (Barcodes.jl:75)
    uint64_t remaining_need;  // inserts still missing over all live schemas
    uint64_t error_mol;
    double rate;              // inserts per barcode group, running estimate
    unsigned long long cut;   // scratch: cut group of the batch in flight (atomicMin), ~0 = none
    uint32_t n_live, epoch, done, need_table, error_code, n_batches, n_epochs, warps_done, hstride, pad;
};
struct PlanArrays {
    uint16_t *live;            // [K] live schemas in first-seen order
    uint16_t *lpos;            // [65536] schema -> position in `live`
    const int64_t *required;   // [K] by schema
    int64_t *current;          // [K] by schema
    int64_t *need;             // [K] by live position: required - current
    unsigned long long *sums;  // [K] by live position: scratch of k_plan_close
    unsigned long long *hist;  // [n_live][hstride] inserts per (live schema, chunk); all-zero between batches
    uint32_t *epoch_nlive;     // [65536] live schemas in every pruning epoch (replay needs the range of each target draw)
};

struct EmitArgs {
    const uint32_t *hap;
    const SchemaDev *sch;
    const uint64_t *read_off;  // [n_mol_shard + 1] shard-relative first pair of each shard molecule
    const MolRec *mol;
    const uint8_t *text;
    const uint2 *ins;          // (start - mol.start, length) per shard pair
    const uint32_t *tile_mol0; // [n_tiles] first shard molecule of every tile
    const uint4 *fdesc;        // [n_pairs] packed per-pair descriptors written by k_tile_prepare
    const uint8_t *tile_flags; // [n_tiles] 1: every pair regular and the tile fits the stage (fast path)
    const uint64_t *tile_off1; // [n_tiles + 1] shard-relative byte offsets
    const uint64_t *tile_off2;
    uint8_t *out1, *out2;
    uint64_t chunk_off1, chunk_off2;  // byte offset of this chunk in the shard text
    uint64_t n_pairs;          // pairs in shard
    uint32_t n_mol;            // molecules in shard
    uint32_t tile0, n_tiles;
    uint32_t tile_pairs;
    uint32_t text_stride;
    int32_t r1, r2;
    int32_t fq_format, bclen_inline;  // bclen_inline = 16 for tenx else 0
    int32_t strict, bpb;
    int32_t force_generic;
    uint32_t stage_cap;        // bytes of shared-memory stage per stream (fast path)
    int32_t fast_ok;           // host-level eligibility of the fast path (read lengths, format)
    const uint8_t *qfill;      // >= 2 * stage_cap bytes of the quality character (source of the stage fill by bulk copy); null: k_emit
    // sequencing-error model (optional, off by default; not in the reference — DESIGN.md "Error model")
    int32_t err_on;
    uint32_t err_tins, err_tdel;  // insertion / deletion probability scaled by 2^24
    const uint32_t *err_thr;      // [r1 + r2] substitution thresholds per cycle (R1 cycles, then R2), scaled by 2^24
    const uint8_t *err_qb;        // [r1 + r2] base Phred quality per cycle
    uint64_t seed, pair_base;     // pair_base: run-wide index of the shard's first pair
    uint32_t sample;
};

#ifdef __cplusplus
extern "C++" {
#endif
// ---- launchers implemented in mmk_device.cu (all asynchronous on `st`) ---------------------
namespace mmkdev {

struct ContigStats {
    unsigned long long first_bad;  // index of first symbol outside the DNA alphabet, or ~0
    unsigned long long n_N;
    unsigned long long n_nonACGT;
};
void normalize_ascii(char *d_bases, int64_t len, ContigStats *d_stats, cudaStream_t st);
void mask_n(char *d_bases, int64_t len, uint64_t seed, uint32_t contig, uint32_t hap, cudaStream_t st);
void count_bases(const char *d_bases, int64_t len, ContigStats *d_stats, cudaStream_t st);
void scatter_snps(char *d_bases, const int64_t *d_pos, const char *d_alt, int64_t n, cudaStream_t st);
// segmented gather + pack: output base o comes from piece p = last piece with dst[p] <= o:
// src_kind 0 → d_master[src[p] + (o - dst[p])], 1 → d_alt[src[p] + (o - dst[p])]
void pack_plain2(const char *ascii, int64_t len, uint32_t *words, ContigStats *st, cudaStream_t s);
void pack_pieces(const char *d_master, const char *d_alt, const int64_t *d_dst, const int64_t *d_src,
                 const uint8_t *d_kind, int32_t n_pieces, int64_t out_len, uint32_t *d_words, int bpb,
                 ContigStats *d_stats, cudaStream_t st);
void unpack(const uint32_t *d_words, int64_t len, int bpb, char *d_out, cudaStream_t st);

void exclusive_scan_u32(const uint32_t *d_in, uint64_t *d_out /*[n+1]*/, uint64_t n, uint64_t *d_tmp /*[>= n/1024+2]*/,
                        cudaStream_t st, const uint64_t *d_base = nullptr);

// Planning. The group table (n_mol per barcode group, molecule offsets) does not depend on the live
// set, so molecule indices are final whatever the pruning epochs turn out to be.
uint32_t plan_hist_chunks(uint64_t max_molecules);
void plan_extend_table(const SimConst &sc, uint64_t g_lo, uint64_t g_hi, uint32_t *d_nmol, uint64_t *d_off, uint64_t *d_scan_tmp,
                       cudaStream_t st);
void plan_fill_slots(uint64_t g_lo, uint64_t g_hi, const uint32_t *d_nmol, const uint64_t *d_off, PlanDev plan, cudaStream_t st);
void plan_init(PlanState *d_st, const PlanArrays &pa, uint32_t K, cudaStream_t st);
// enqueues `n_batches` batches back to back; each reads its parameters from *d_st (a batch after the end of the plan is three empty launches)
void plan_batches(const SimConst &sc, const SchemaDev *d_sch, PlanState *d_st, const PlanArrays &pa, const uint64_t *d_off, PlanDev plan,
                  unsigned long long *d_first_err, int n_batches, cudaStream_t st);

void shard_bounds(const uint64_t *d_read_off, uint64_t n_mol, uint64_t n_pairs, int32_t world, uint64_t *d_bounds /*[world+1]*/,
                  cudaStream_t st);

void place_prepare(uint64_t mol0, uint32_t n_mol, const uint64_t *d_read_off_global, uint64_t *d_read_off_local,
                   uint32_t *d_tile_mol0, uint32_t tile_pairs, cudaStream_t st);
size_t place_tmp_bytes(uint64_t max_pairs);
// d_acc_i: per shard molecule, zeroed by the caller before the first segment
void place_segment_draws(const SimConst &sc, PlanDev plan, uint64_t mol0, uint32_t n_mol, uint64_t pair_lo, uint64_t pair_hi,
                         const uint64_t *d_read_off_local, const uint32_t *d_tile_mol0, uint32_t n_tiles, uint32_t tile_pairs,
                         void *d_tmp, uint32_t *d_acc_i, const InsTable *d_tab, cudaStream_t st);
void place_segment_chain(const SimConst &sc, const SchemaDev *d_sch, PlanDev plan, uint64_t mol0, uint32_t m_lo, uint32_t m_hi,
                         uint64_t pair_lo, const uint64_t *d_read_off_local, void *d_tmp, uint32_t *d_acc_i, uint2 *d_ins, MolRec *d_mol, uint8_t *d_text, uint32_t text_stride,
                         unsigned long long *d_err, cudaStream_t st);

// block placement (mmk_place.cuh): needs 32-pair tiles and insert lengths below 2^16
uint32_t place_block_count(uint64_t n_pairs, uint32_t n_mol);  // entries of d_blk - 1
uint32_t place_big_cap(uint64_t n_pairs);                      // entries of d_big_list
bool place_block_supports(const SimConst &sc);
void place_block(const SimConst &sc, const SchemaDev *d_sch, PlanDev plan, uint64_t mol0, uint32_t n_mol, uint64_t n_pairs,
                 uint32_t n_tiles, const uint64_t *d_roff_l, uint32_t *d_tile_mol0, const InsTable *d_tab, MolRec *d_mol,
                 uint8_t *d_text, uint32_t text_stride, uint4 *d_fdesc, uint32_t *d_sum1, uint32_t *d_sum2, uint8_t *d_tile_flags,
                 unsigned long long *d_gather, unsigned long long *d_err, uint32_t *d_weight, uint64_t *d_woff, uint64_t *d_scan_tmp,
                 uint32_t *d_blk, uint32_t *d_big_list, uint32_t *d_big_count, uint32_t stage_cap, int bpb, cudaStream_t st, cudaStream_t st_text,
                 cudaEvent_t ev_placed, cudaEvent_t ev_text);
// the insert plan of a block-placed shard (the block placement keeps only the descriptors): for the dump entry points
void ins_from_fdesc(const uint4 *d_fdesc, const MolRec *d_mol, const uint32_t *d_tile_mol0, uint64_t n_pairs, uint2 *d_ins, cudaStream_t st);
void place_local_roff(uint64_t mol0, uint32_t n_mol, const uint64_t *d_read_off_global, uint64_t *d_read_off_local, cudaStream_t st);
// KAT: lengths the kernels derive from raw 64-bit draws (table path with its exact fallback); slow[i] = 1 where the fallback ran
void kat_insert_len(const SimConst &sc, const InsTable *d_tab, const uint64_t *d_bits, uint32_t *d_len, uint8_t *d_slow, int64_t n, cudaStream_t st);

void tile_prepare(const EmitArgs &a, uint32_t tile_lo, uint32_t tile_hi, uint32_t *d_sum1, uint32_t *d_sum2, uint4 *d_fdesc,
                  uint8_t *d_tile_flags, unsigned long long *d_gather, cudaStream_t st);
void emit(const EmitArgs &a, cudaStream_t st);
void chunk_table(const uint64_t *d_off1, const uint64_t *d_off2, uint32_t n_tiles, uint64_t cap, uint32_t *d_bounds,
                 uint64_t *d_boff1, uint64_t *d_boff2, uint32_t max_chunks, uint32_t *d_n, cudaStream_t st);
unsigned long long launch_count();

void dump_variates(const SimConst &sc, const SchemaDev *d_sch, PlanDev plan, uint64_t n_mol, const uint64_t *d_read_off,
                   const uint2 *d_ins, const uint32_t *d_epoch_nlive, const uint32_t *d_tile_mol0, uint32_t tile_pairs,
                   uint32_t *d_counts, uint64_t *d_offsets, uint64_t *d_scan_tmp, mmk_variate *d_out, int pass,
                   cudaStream_t st);
void dump_mask(const char *d_raw_upper, int64_t len, uint64_t seed, uint32_t contig, uint32_t hap, uint32_t *d_flags,
               uint64_t *d_off, uint64_t *d_scan_tmp, mmk_variate *d_out, int pass, cudaStream_t st);

void kat_philox(const uint32_t *d_in6, uint32_t *d_out4, cudaStream_t st);
void kat_math(int kind, const double *d_in, double *d_out, int64_t n, cudaStream_t st);

}  // namespace mmkdev
#ifdef __cplusplus
}

// ------------------------------------------------------------------------------------------
// BGZF stage (mmk_bgzf.cuh)
// ------------------------------------------------------------------------------------------
struct BgzfArgs {
    const uint8_t *text;       // device text of the chunk (stream offset chunk_off sits at text[0])
    const uint64_t *tile_off;  // shard-wide stream offset of every tile
    uint64_t chunk_off;
    uint32_t tile0, n_tiles;   // tiles of the chunk
    uint32_t tiles_per_block;
    uint32_t in_cap;           // staging capacity in bytes (multiple of 16, <= 65280)
    uint32_t max_records;      // records per block the line table can hold
    uint32_t out_cap;          // bytes of the coded-member buffer (multiple of 16); larger members are stored
    uint32_t warp_smem;        // shared-memory bytes of one warp's region (multiple of 16)
    uint32_t crc_words;        // 32-bit words per lane in the CRC pass (32 * crc_words * 4 >= in_cap + 4)
    const uint32_t *tab;       // codes, CRC-32 slicing tables, per-lane CRC shifts x^(32 * crc_words * (31 - lane)), block prefix
    uint32_t hdr_bits;
    uint8_t *slots;            // slot area; block b writes at bz_slot_off(its text offset, b, slot_shift)
    uint32_t slot_shift;       // 2^slot_shift <= in_cap / 32
    uint32_t *sizes;           // coded size of every block
};

namespace mmkdev {
constexpr uint32_t BGZF_TAB_CODES = 286 + 30 + 4 * 256 + 32;  // words before the block prefix in BgzfArgs::tab
void bgzf_geometry(BgzfArgs &a);  // fills out_cap, warp_smem, crc_words from in_cap and max_records
void bgzf_encode(const BgzfArgs &a, uint32_t n_blocks, cudaStream_t st);
void bgzf_compact(const BgzfArgs &a, const uint64_t *offs, uint8_t *dst, uint32_t n_blocks, unsigned long long *host_total,
                  cudaStream_t st);  // host_total: pinned host memory
uint64_t bgzf_slots_bytes(uint64_t text_bytes, uint32_t n_blocks, uint32_t slot_shift);  // capacity of the slot area = bound of the coded chunk
}  // namespace mmkdev

#endif
