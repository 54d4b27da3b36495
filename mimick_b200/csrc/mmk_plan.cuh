// mmk_plan.cuh — the barcode-group loop of Main.jl:44-59 with its control on the device.
//
// The reference draws targets from the live schema set, adds every molecule's insert count to its schema's tracker
// (Main.jl:54) and prunes finished schemas at the end of each barcode group (:58).  The live set is constant between
// pruning points, so groups are planned speculatively in batches and each batch is cut at the first group at whose end a
// schema reaches its target; the groups after the cut are planned again against the smaller live set.
//
// A batch is three kernels, and everything a batch needs to know (first group, size, live list, per-schema need, epoch)
// is read from a control block in device memory that the last warp of the previous batch wrote: the host enqueues
// batches back to back and looks at the control block once per round, not once per batch.
//   k_plan_molecules  one thread per molecule of the batch (grid-stride): target, size, insert count (Main.jl:47,
//                     Breakpoints.jl:149-196); per-(schema, chunk of 512 molecules) insert histogram
//   k_plan_cut        one warp per live schema: the group in which its cumulative insert count reaches its need;
//                     the earliest over all schemas is the cut
//   k_plan_close      one warp per live schema: inserts received from the kept molecules; the last warp to finish
//                     updates the trackers, prunes the live list and sizes the next batch
#pragma once
#include "mmk_internal.h"
#include "mmk_sampling.cuh"

namespace mmkdev {

constexpr uint32_t HCH = 512;  // molecules per histogram chunk
constexpr uint32_t PLAN_THREADS = 256, PLAN_WARPS = PLAN_THREADS / 32;

__device__ __forceinline__ bool plan_idle(const PlanState *st) { return st->done || st->need_table || st->error_code; }

__global__ void __launch_bounds__(PLAN_THREADS) k_plan_molecules(const __grid_constant__ SimConst sc, const SchemaDev *__restrict__ sch,
                                                                 const PlanState *__restrict__ st, PlanArrays pa,
                                                                 const uint64_t *__restrict__ off, PlanDev plan, unsigned long long *first_err) {
    if (plan_idle(st)) return;
    const uint64_t g0 = st->g0;
    const uint32_t n_live = st->n_live, hstride = st->hstride;
    const uint16_t epoch = (uint16_t)st->epoch;
    const uint64_t base = off[g0], n_mol = off[g0 + st->B] - base;
    for (uint64_t j = (uint64_t)blockIdx.x * PLAN_THREADS + threadIdx.x; j < n_mol; j += (uint64_t)gridDim.x * PLAN_THREADS) {
        const uint64_t i = base + j;
        const uint64_t g = plan.group[i];
        const uint32_t m = plan.slot[i];
        NoRec rec;
        const uint32_t t = pa.live[draw_target(sc, g, m, n_live)];
        uint32_t molsize, nfrags;
        const uint32_t status = molecule_shape(sc, g, m, sch[t].len, molsize, nfrags, rec);
        plan.schema[i] = (uint16_t)t;
        plan.epoch[i] = epoch;
        plan.size[i] = molsize;
        plan.nfrags[i] = nfrags;
        plan.status[i] = (uint8_t)status;
        atomicAdd(&pa.hist[(uint64_t)pa.lpos[t] * hstride + (uint32_t)(j / HCH)], (unsigned long long)nfrags);
        if (status) atomicMin(first_err, (unsigned long long)j);
    }
}

__global__ void __launch_bounds__(PLAN_THREADS) k_plan_cut(PlanState *st, PlanArrays pa, const uint16_t *__restrict__ schema,
                                                           const uint32_t *__restrict__ nfrags, const uint64_t *__restrict__ group,
                                                           const uint64_t *__restrict__ off) {
    if (plan_idle(st)) return;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t g0 = st->g0;
    const uint32_t n_live = st->n_live, hstride = st->hstride;
    const uint64_t base = off[g0];
    const uint64_t n_mol = off[g0 + st->B] - base;
    const uint32_t n_chunks = (uint32_t)((n_mol + HCH - 1) / HCH);
    schema += base;
    nfrags += base;
    group += base;
    constexpr uint32_t PER = HCH / 32;
    for (uint32_t s = blockIdx.x * PLAN_WARPS + (threadIdx.x >> 5); s < n_live; s += gridDim.x * PLAN_WARPS) {
        const long long want = pa.need[s];
        unsigned long long cg = ~0ull;
        if (want <= 0) {
            cg = g0;  // already complete: pruned at the end of the batch's first group
        } else {
            const uint16_t sid = pa.live[s];
            unsigned long long cum = 0;
            uint32_t found = 0xffffffffu;
            // four rounds of loads in flight: the scan below is a chain of L2 latencies otherwise
            const unsigned long long *hs = pa.hist + (uint64_t)s * hstride;
            unsigned long long pre[4];
#pragma unroll
            for (uint32_t u = 0; u < 4; ++u) pre[u] = 32 * u + lane < n_chunks ? hs[32 * u + lane] : 0ull;
            for (uint32_t c0 = 0; c0 < n_chunks && found == 0xffffffffu; c0 += 32) {
                const unsigned long long v = pre[0];
                pre[0] = pre[1];
                pre[1] = pre[2];
                pre[2] = pre[3];
                pre[3] = c0 + 128 + lane < n_chunks ? hs[c0 + 128 + lane] : 0ull;
                unsigned long long inc = v;
                for (int o = 1; o < 32; o <<= 1) {
                    unsigned long long y = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= (uint32_t)o) inc += y;
                }
                const unsigned ball = __ballot_sync(0xffffffffu, (long long)(cum + inc) >= want);
                if (ball) {
                    const int l = __ffs(ball) - 1;
                    cum += __shfl_sync(0xffffffffu, inc - v, l);
                    found = c0 + l;
                } else
                    cum += __shfl_sync(0xffffffffu, inc, 31);
            }
            if (found != 0xffffffffu) {  // the molecule inside the chunk: each lane owns PER consecutive molecules
                const uint64_t lo = (uint64_t)found * HCH + (uint64_t)lane * PER;
                unsigned long long part = 0;
                for (uint32_t j = 0; j < PER; ++j) {
                    const uint64_t i = lo + j;
                    if (i < n_mol && schema[i] == sid) part += nfrags[i];
                }
                unsigned long long inc = part;
                for (int o = 1; o < 32; o <<= 1) {
                    unsigned long long y = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= (uint32_t)o) inc += y;
                }
                const unsigned ball = __ballot_sync(0xffffffffu, (long long)(cum + inc) >= want);
                const int l = __ffs(ball) - 1;  // non-zero by construction
                if ((int)lane == l) {
                    unsigned long long run = cum + inc - part;
                    for (uint32_t j = 0; j < PER; ++j) {
                        const uint64_t i = lo + j;
                        if (i < n_mol && schema[i] == sid) {
                            run += nfrags[i];
                            if ((long long)run >= want) {
                                cg = group[i];
                                break;
                            }
                        }
                    }
                }
                cg = __shfl_sync(0xffffffffu, cg, l);
            }
        }
        if (lane == 0 && cg != ~0ull) atomicMin(&st->cut, cg);
    }
}

// the state after a batch: trackers, live list, next batch (one warp)
__device__ void plan_finish_batch(PlanState *st, PlanArrays pa, const uint64_t *__restrict__ off, const uint8_t *__restrict__ status,
                                  unsigned long long *first_err, uint32_t lane) {
    const uint64_t g0 = st->g0, B = st->B;
    const uint32_t n_live = st->n_live;
    const uint64_t base = off[g0], n_mol = off[g0 + B] - base;
    const unsigned long long cutv = st->cut;
    const uint64_t nc = (cutv == ~0ull) ? n_mol : off[cutv + 1] - base;
    const unsigned long long fe = *first_err;
    unsigned long long batch_inserts = 0;
    for (uint32_t i = lane; i < n_live; i += 32) {
        const unsigned long long v = pa.sums[i];
        pa.current[pa.live[i]] += (long long)v;
        batch_inserts += v;
    }
    for (int o = 16; o; o >>= 1) batch_inserts += __shfl_xor_sync(0xffffffffu, batch_inserts, o);
    __syncwarp();
    uint32_t new_live = n_live;
    if (cutv != ~0ull) {  // filter!(is_incomplete, schema), Main.jl:58 — order kept
        new_live = 0;
        for (uint32_t i0 = 0; i0 < n_live; i0 += 32) {
            const uint32_t i = i0 + lane;
            uint16_t s = 0;
            bool keep = false;
            if (i < n_live) {
                s = pa.live[i];
                keep = pa.current[s] < pa.required[s];
            }
            const unsigned ball = __ballot_sync(0xffffffffu, keep);
            __syncwarp();
            if (keep) {
                const uint32_t pos = new_live + __popc(ball & ((1u << lane) - 1u));
                pa.live[pos] = s;  // pos <= i: never overwrites an entry not yet read (reads of this round are done)
                pa.lpos[s] = (uint16_t)pos;
            }
            new_live += __popc(ball);
            __syncwarp();
        }
    }
    long long min_need = 0x7fffffffffffffffll;
    unsigned long long remaining = 0;
    for (uint32_t i = lane; i < new_live; i += 32) {
        const uint16_t s = pa.live[i];
        const long long nd = pa.required[s] - pa.current[s];
        pa.need[i] = nd;
        min_need = nd < min_need ? nd : min_need;
        remaining += nd > 0 ? (unsigned long long)nd : 0ull;
    }
    for (int o = 16; o; o >>= 1) {
        const long long y = __shfl_xor_sync(0xffffffffu, min_need, o);
        min_need = y < min_need ? y : min_need;
        remaining += __shfl_xor_sync(0xffffffffu, remaining, o);
    }
    if (lane == 0) {
        const uint64_t groups_done = (cutv == ~0ull) ? B : (cutv - g0 + 1);
        st->n_batches += 1;
        if (fe < nc) {
            st->error_code = status[base + fe];
            st->error_mol = st->n_mol + fe;
        }
        st->n_mol += nc;
        st->g0 = g0 + groups_done;
        if (groups_done > 0 && batch_inserts > 0) st->rate = (double)batch_inserts / (double)groups_done;
        st->remaining_need = remaining;
        if (cutv != ~0ull) {
            st->n_epochs += 1;
            st->n_live = new_live;
            if (new_live > 0) {
                st->epoch += 1;
                if (st->epoch >= 65535u && !st->error_code) st->error_code = MMK_ERR_UNSUPPORTED;  // epochs are stored in 16 bits
                else pa.epoch_nlive[st->epoch] = new_live;
            }
        }
        if (new_live == 0) st->done = 1;
        if (st->n_mol > st->barcode_cap && !st->error_code) {
            st->error_code = MMK_ERR_BARCODES_EXHAUSTED;  // Barcodes.jl:75
            st->error_mol = st->barcode_cap;
        }
        // size the next batch to reach the next pruning point with a small overshoot
        uint64_t nB = st->B;
        if (st->rate > 0.0 && min_need > 0) {
            const double est = (double)min_need * (double)new_live / st->rate * 1.03 + 2048.0;
            nB = est >= (double)st->Bmax ? st->Bmax : (uint64_t)est;
            if (nB < st->Bmin) nB = st->Bmin;
        } else if (min_need <= 0)
            nB = st->Bmin;
        if (!st->done) {
            if (st->g0 + nB > st->g_tab) {  // the group table ends here: plan what it covers, or ask the host for more
                if (st->g_tab > st->g0) nB = st->g_tab - st->g0;
                else st->need_table = 1;
            }
            st->B = nB;
        }
        st->cut = ~0ull;
        *first_err = ~0ull;
        st->warps_done = 0;
    }
}

__global__ void __launch_bounds__(PLAN_THREADS) k_plan_close(PlanState *st, PlanArrays pa, const uint16_t *__restrict__ schema,
                                                             const uint32_t *__restrict__ nfrags, const uint64_t *__restrict__ off,
                                                             const uint8_t *__restrict__ status, unsigned long long *first_err) {
    if (plan_idle(st)) return;
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t g0 = st->g0;
    const uint32_t n_live = st->n_live, hstride = st->hstride;
    const uint64_t base = off[g0];
    const uint64_t n_mol = off[g0 + st->B] - base;
    const uint32_t n_chunks = (uint32_t)((n_mol + HCH - 1) / HCH);
    const unsigned long long cutv = st->cut;
    const uint64_t nc = (cutv == ~0ull) ? n_mol : off[cutv + 1] - base;
    const uint16_t *sc_b = schema + base;
    const uint32_t *nf_b = nfrags + base;
    // inserts of every live schema among the kept molecules: whole chunks from the histogram, the last from the molecules
    const uint32_t full = (uint32_t)(nc / HCH);
    for (uint32_t s = blockIdx.x * PLAN_WARPS + (threadIdx.x >> 5); s < n_live; s += gridDim.x * PLAN_WARPS) {
        const uint16_t sid = pa.live[s];
        unsigned long long *hs = pa.hist + (uint64_t)s * hstride;
        unsigned long long acc = 0;
        for (uint32_t c = lane; c < full; c += 32) acc += hs[c];
        for (uint64_t i = (uint64_t)full * HCH + lane; i < nc; i += 32)
            if (sc_b[i] == sid) acc += nf_b[i];
        for (int o = 16; o; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
        // leave the histogram all-zero for the next batch (this batch dirtied n_chunks columns of n_live rows)
        for (uint32_t c = lane; c < n_chunks; c += 32) hs[c] = 0ull;
        uint32_t ticket = 0;
        __threadfence();
        __syncwarp();
        if (lane == 0) {
            pa.sums[s] = acc;
            __threadfence();
            ticket = atomicAdd(&st->warps_done, 1u);
        }
        ticket = __shfl_sync(0xffffffffu, ticket, 0);
        if (ticket == n_live - 1u) {  // every other schema's sum is visible: close the batch
            __threadfence();
            plan_finish_batch(st, pa, off, status, first_err, lane);
        }
    }
}

// first batch of a plan: live = all schemas, trackers at zero
__global__ void k_plan_init(PlanState *st, PlanArrays pa, uint32_t K) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < K) {
        pa.live[i] = (uint16_t)i;
        pa.lpos[i] = (uint16_t)i;
        pa.current[i] = 0;
        pa.need[i] = pa.required[i];
    }
    if (i == 0) pa.epoch_nlive[0] = K;
}

}  // namespace mmkdev
