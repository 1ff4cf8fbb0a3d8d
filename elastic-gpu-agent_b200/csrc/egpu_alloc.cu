// egpu_alloc.cu — best-fit fractional-GPU allocation on B200 (sm_100a) + its C ABI.
//
// Product path.  There is no CPU fallback in this file and nothing here
// includes or links oracle/: when no CUDA device is usable every entry point
// returns EGPU_ERR_NO_DEVICE.
//
// Reference slot this fills: baseDevicePlugin.GetPreferredAllocation, an empty
// stub in elastic-ai/elastic-gpu-agent (pkg/plugins/base.go:94-96); units from
// pkg/common/const.go:4 and pkg/plugins/gpushare.go:24-33,159-168.  The decision
// rule is the builder-defined spec of DESIGN.md §2 (the reference has none).
#include "egpu_kernels.cuh"

#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

#include "../../include/egpu_alloc.h"

namespace egpu {

// =============================================================================
// Snapshot scan
// =============================================================================
//
// Why "sorted" is the fast formulation.  For a fixed table the best-fit device
// of request (c, m) minimises (fc-c, fm-m, d) over feasible d, which is the same
// as minimising (fc, fm, d): the request cancels out of the comparison.  So the
// answer is the FIRST feasible device in the table sorted by (fc, fm, d).  Each
// CTA sorts the <= 64 table rows once (rank sort in shared memory), every thread
// keeps the packed sorted rows in registers, and per (request, device) pair the
// work is: one subtract (both feasibility tests at once, see kGuards), one LOP3
// producing "sorted position, or a value >= 2^18 if infeasible", and half a
// 3-input unsigned min (VIMNMX3).  The chosen position maps back to the device
// index through a shared-memory tile.

template <int DT, int THREADS>
struct SnapSmem {
    int32_t sFc[kMaxD];                           // table tile (grid variant; re-sort scratch)
    int32_t sFm[kMaxD];
    int32_t sPosDev[kMaxD];
    unsigned long long sWarpAcc[THREADS / 32][2 * DT];
    int sLast;
    int32_t sDevTile[THREADS / 32][DT + 8];           // warp-private: sorted position -> device, [DT] = -1
    unsigned long long hist[THREADS / 32][DT + 1][32];  // lane-private demand sums; row 0 = "no device"
};

// Re-derive the sorted view of the table (DevState::sorted_k / sorted_dev /
// dev_packed).  Called by every thread of ONE CTA after thread d < D has put the
// new row d into sFc[d] / sFm[d].  Rank sort: position = rows ordering before.
__device__ __forceinline__ void resort_table_cta(DevState* st, int D, int32_t* sFc, int32_t* sFm,
                                                 int32_t* sPosDev, int tid) {
    const int nt = blockDim.x;
    __syncthreads();
    for (int d = tid; d < kMaxD; d += nt) sPosDev[d] = -1;
    __syncthreads();
    for (int d = tid; d < kMaxD; d += nt) {
        if (d < D) {
            const int32_t fc = sFc[d], fm = sFm[d];
            const uint32_t mine = (static_cast<uint32_t>(fc) << 24) | (static_cast<uint32_t>(fm) << 6) | d;
            int pos = 0;
            for (int k = 0; k < D; ++k) {
                const uint32_t other = (static_cast<uint32_t>(sFc[k]) << 24) | (static_cast<uint32_t>(sFm[k]) << 6) | k;
                pos += other < mine;
            }
            st->sorted_k[pos] = pack_table_word(fc, fm) | (static_cast<uint32_t>(pos) & 31u);
            st->sorted_dev[pos] = d;
            sPosDev[pos] = d;
        } else {  // positions >= D are never produced by a rank
            st->sorted_k[d] = kPadWord;
            st->sorted_dev[d] = -1;
        }
    }
    __syncthreads();
    if (tid == 0) {
        unsigned long long packed = 0;
        for (int j = 0; j < 8; ++j)
            packed |= static_cast<unsigned long long>(static_cast<uint32_t>(sPosDev[j]) & 0xffu) << (8 * j);
        st->dev_packed = packed;
    }
}

// First feasible sorted position; >= DT when there is none.
// Per (request, row) pair: one subtract (IMAD.IADD or IADD3, ptxas balances the FMA and
// ALU pipes), one 3-input LOP3 ((t ^ G) & M, both masks in registers), half a VIMNMX3.
template <int N>
__device__ __forceinline__ uint32_t first_feasible32(const uint32_t* K, uint32_t q, uint32_t gx, uint32_t gm) {
    uint32_t best = kNoCand;
#pragma unroll
    for (int j = 0; j < N; j += 2) {
        const uint32_t c0 = ((K[j] - q) ^ gx) & gm;
        const uint32_t c1 = ((K[j + 1] - q) ^ gx) & gm;
        best = __vimin3_u32(best, c0, c1);
    }
    return best;  // 0..N-1, or >= 32
}
template <int DT>
__device__ __forceinline__ uint32_t first_feasible(const uint32_t (&K)[DT], uint32_t q, uint32_t gx, uint32_t gm) {
    if constexpr (DT <= 32) {
        return min(first_feasible32<DT>(K, q, gx, gm), static_cast<uint32_t>(DT));
    } else {  // positions are stored mod 32: two halves
        const uint32_t lo = first_feasible32<32>(K, q, gx, gm);
        const uint32_t hi = first_feasible32<DT - 32>(K + 32, q, gx, gm);
        return lo < 32u ? lo : (hi < 32u ? 32u + hi : static_cast<uint32_t>(DT));
    }
}

template <int DT, int THREADS>
__device__ __forceinline__ void hist_zero(SnapSmem<DT, THREADS>& s, int warp, int lane) {
    // lane-private: each lane clears exactly the words it will use -> no barrier
#pragma unroll
    for (int d = 0; d <= DT; ++d) s.hist[warp][d][lane] = 0ull;
}

template <int DT, int THREADS>
__device__ __forceinline__ void hist_add(SnapSmem<DT, THREADS>& s, int warp, int lane, int32_t idx,
                                         int32_t core, int32_t mem) {
    // unconditional: infeasible rows (idx = -1) land in the dummy row 0, whose
    // content is never read (it may hold garbage from out-of-domain requests)
    s.hist[warp][idx + 1][lane] +=
        (static_cast<unsigned long long>(static_cast<uint32_t>(core)) << kAccShift) |
        static_cast<unsigned long long>(static_cast<uint32_t>(mem));
}

// Second half of every snapshot epilogue.  `wacc` holds per-warp demand sums in shared
// memory: core sum of device d of warp w at wacc[w * wstride + core_off + d], mem sum at
// [... + mem_off + d].  Called by all threads after a __syncthreads().  Publishes the CTA's
// sums with one red.global.add.u64 per device, takes an arrival ticket, and the last CTA
// writes delta / table', optionally commits (and re-sorts) the table and resets the slot.
template <int WARPS>
__device__ __forceinline__ void epilogue_publish(const unsigned long long* wacc, int wstride, int core_off, int mem_off,
                                                 int32_t* sFc, int32_t* sFm, int32_t* sPosDev, int* sLast,
                                                 DevState* st, int D, long long* __restrict__ delta_out,
                                                 int32_t* __restrict__ table_out, int flags, unsigned long long slot_step) {
    // slot_step: bits 0..7 = epilogue slot of this launch; bits 8.. = step + 1 when the demand
    // vector must also be pushed to the peers' exchange buffers (0 = single GPU)
    DevState::EpiSlot& ep = st->epi[slot_step & 0xffu];
    const unsigned long long push = slot_step >> 8;
    const int tid = threadIdx.x;
    if (tid < 2 * D) {
        const int j = tid < D ? core_off + tid : mem_off + (tid - D);
        unsigned long long tot = 0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) tot += wacc[w * wstride + j];
        if (tot) atomicAdd(&ep.acc[tid < D ? tid : kMaxD + (tid - D)], tot);
        __threadfence();  // only the threads that published sums need to order them before the ticket
    }
    __syncthreads();
    if (tid == 0) {
        const unsigned int ticket = atomicAdd(&ep.ticket, 1u);
        *sLast = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (!*sLast) return;
    __threadfence();
    const bool fin = (flags & kFlagFinalize) != 0;
    const bool commit = fin && (flags & kFlagCommit);
    if (fin && tid < D) {
        volatile unsigned long long* acc = ep.acc;
        const long long dc = static_cast<long long>(acc[tid]);
        const long long dm = static_cast<long long>(acc[kMaxD + tid]);
        acc[tid] = 0ull;
        acc[kMaxD + tid] = 0ull;
        const long long nc = static_cast<long long>(st->free_core[tid]) - dc;
        const long long nm = static_cast<long long>(st->free_mem[tid]) - dm;
        const int32_t over = (nc < 0 || nm < 0) ? 1 : 0;
        if (delta_out) {
            delta_out[tid] = dc;
            delta_out[D + tid] = dm;
        }
        if (push) {  // fused exchange: this rank's vector straight into every rank's buffer
            const int world = st->peer.world, me = st->peer.rank;
            const int xs = static_cast<int>((push - 1) % kXchgSlots);
            for (int p = 0; p < world; ++p) {
                XchgRow& row = st->peer.buf[p]->slot[xs][me];
                row.delta[tid] = dc;
                row.delta[D + tid] = dm;
            }
            __threadfence_system();
        }
        if (table_out) {
            table_out[tid] = sat_i32(nc);
            table_out[D + tid] = sat_i32(nm);
            table_out[2 * D + tid] = over;
        }
        if (commit) {
            // the committed table stays inside the spec's domain: negative
            // leftovers clamp to 0 and the oversubscription flag is sticky
            const int32_t cc = nc < 0 ? 0 : static_cast<int32_t>(nc);
            const int32_t cm = nm < 0 ? 0 : static_cast<int32_t>(nm);
            st->free_core[tid] = cc;
            st->free_mem[tid] = cm;
            st->oversub[tid] |= over;
            sFc[tid] = cc;
            sFm[tid] = cm;
        }
    }
    if (commit) resort_table_cta(st, D, sFc, sFm, sPosDev, tid);
    if (push) {
        __syncthreads();  // every delta store above is fenced; now raise the flags
        const int world = st->peer.world, me = st->peer.rank;
        if (tid < world) {
            unsigned long long* f = &st->peer.buf[tid]->slot[(push - 1) % kXchgSlots][me].flag;
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(push) : "memory");
        }
    }
    if (tid == 0) ep.ticket = 0u;
}

// Demand sums -> global running sums -> (last CTA) delta / table' publication.
template <int DT, int THREADS>
__device__ __forceinline__ void snapshot_epilogue(SnapSmem<DT, THREADS>& s, DevState* st, int D,
                                                  long long* __restrict__ delta_out,
                                                  int32_t* __restrict__ table_out, int flags, unsigned long long slot_step) {
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    __syncwarp();
    for (int d = 0; d < D; ++d) {
        const unsigned long long v = s.hist[warp][d + 1][lane];
        const uint32_t c = static_cast<uint32_t>(v >> kAccShift);
        const uint32_t ml = static_cast<uint32_t>(v) & 0x7FFFFu;
        const uint32_t mh = static_cast<uint32_t>(v >> 19) & 0x7FFFFu;
        const uint32_t sc = __reduce_add_sync(0xffffffffu, c);
        const uint32_t sl = __reduce_add_sync(0xffffffffu, ml);
        const uint32_t sh = __reduce_add_sync(0xffffffffu, mh);
        if (lane == 0) {
            s.sWarpAcc[warp][d] = sc;
            s.sWarpAcc[warp][DT + d] = static_cast<unsigned long long>(sl) + (static_cast<unsigned long long>(sh) << 19);
        }
    }
    __syncthreads();
    epilogue_publish<THREADS / 32>(&s.sWarpAcc[0][0], 2 * DT, 0, DT, s.sFc, s.sFm, s.sPosDev, &s.sLast, st, D, delta_out,
                                   table_out, flags, slot_step);
}

template <int DT, int THREADS>
__global__ void __launch_bounds__(THREADS)
bestfit_sorted_kernel(DevState* __restrict__ st, const int32_t* __restrict__ req_core,
                      const int32_t* __restrict__ req_mem, long long R, int32_t* __restrict__ out_idx,
                      long long* __restrict__ delta_out, int32_t* __restrict__ table_out, int flags, unsigned long long slot_step) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    auto& s = *reinterpret_cast<SnapSmem<DT, THREADS>*>(smem_raw);
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const bool late = (flags & kFlagLateWait) != 0;
    if (!late) pdl_wait();  // predecessor may have produced our inputs or changed the table
    if (flags & kFlagEarlyTrigger) pdl_trigger();

    const long long nvec = R >> 2;
    const long long stride = static_cast<long long>(gridDim.x) * THREADS;
    long long v = static_cast<long long>(blockIdx.x) * THREADS + tid;

    // issue the first tile's loads before anything else: the request stream is
    // the only HBM traffic that matters
    int4 c0 = make_int4(0, 0, 0, 0), m0 = c0, c1 = c0, m1 = c0;
    bool has0 = v < nvec, has1 = (v + stride) < nvec;
    if (has0) {
        c0 = ld_stream_v4(req_core + 4 * v);
        m0 = ld_stream_v4(req_mem + 4 * v);
    }
    if (has1) {
        c1 = ld_stream_v4(req_core + 4 * (v + stride));
        m1 = ld_stream_v4(req_mem + 4 * (v + stride));
    }

    // sorted table rows: uniform loads straight into registers, no barrier
    const int D = st->D;
    uint32_t K[DT];
#pragma unroll
    for (int j = 0; j < DT; j += 4) {
        const uint4 k4 = *reinterpret_cast<const uint4*>(&st->sorted_k[j]);
        K[j] = k4.x; K[j + 1] = k4.y; K[j + 2] = k4.z; K[j + 3] = k4.w;
    }
    const uint32_t gx = st->cand_xor, gm = st->cand_mask;
    // warp-private tile of the position -> device map: only a warp-level barrier
    int32_t* tile = s.sDevTile[warp];
    for (int j = lane; j <= DT; j += 32) tile[j] = j < DT ? st->sorted_dev[j] : -1;
    hist_zero<DT, THREADS>(s, warp, lane);
    __syncwarp();

    auto decide = [&](int32_t core, int32_t mem) -> int32_t {
        const uint32_t best = first_feasible<DT>(K, pack_request_word(core, mem), gx, gm);
        const int32_t idx = tile[best];  // best <= DT; tile[DT] = -1
        hist_add<DT, THREADS>(s, warp, lane, idx, core, mem);
        return idx;
    };
    auto decide4 = [&](const int4& c, const int4& m) -> int4 {
        int4 r;
        r.x = decide(c.x, m.x);
        r.y = decide(c.y, m.y);
        r.z = decide(c.z, m.z);
        r.w = decide(c.w, m.w);
        return r;
    };

    while (has0) {
        const long long vn = v + 2 * stride;
        const bool nhas0 = vn < nvec, nhas1 = (vn + stride) < nvec;
        int4 nc0 = make_int4(0, 0, 0, 0), nm0 = nc0, nc1 = nc0, nm1 = nc0;
        if (nhas0) {
            nc0 = ld_stream_v4(req_core + 4 * vn);
            nm0 = ld_stream_v4(req_mem + 4 * vn);
        }
        if (nhas1) {
            nc1 = ld_stream_v4(req_core + 4 * (vn + stride));
            nm1 = ld_stream_v4(req_mem + 4 * (vn + stride));
        }
        st_stream_v4(out_idx + 4 * v, decide4(c0, m0));
        if (has1) st_stream_v4(out_idx + 4 * (v + stride), decide4(c1, m1));
        v = vn;
        has0 = nhas0;
        has1 = nhas1;
        c0 = nc0; m0 = nm0; c1 = nc1; m1 = nm1;
    }
    // ragged tail: R % 4 rows, scalar
    if (blockIdx.x == 0 && tid < static_cast<int>(R & 3)) {
        const long long r = (nvec << 2) + tid;
        out_idx[r] = decide(req_core[r], req_mem[r]);
    }
    snapshot_epilogue<DT, THREADS>(s, st, D, delta_out, table_out, flags, slot_step);
    if (late) pdl_wait();  // do not complete before the predecessor has: keeps stream order transitive
}

// Packed wire format (include/egpu_alloc.h: egpu_bestfit_batch_packed): one uint32 per request
// (core << 18 | mem, anything >= 2^25 = "no valid request") and one int8 per decision - 5 bytes
// per decision instead of 12.  Same scan, same epilogue.
template <int DT, int THREADS>
__global__ void __launch_bounds__(THREADS)
bestfit_sorted_packed_kernel(DevState* __restrict__ st, const uint32_t* __restrict__ req, long long R,
                             signed char* __restrict__ out_idx8, long long* __restrict__ delta_out,
                             int32_t* __restrict__ table_out, int flags, unsigned long long slot_step) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    auto& s = *reinterpret_cast<SnapSmem<DT, THREADS>*>(smem_raw);
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const bool late = (flags & kFlagLateWait) != 0;
    if (!late) pdl_wait();
    if (flags & kFlagEarlyTrigger) pdl_trigger();

    // A warp takes chunks of 512 requests (2 KiB in, 512 B out): four fully coalesced 128-bit
    // loads per lane (lane-contiguous, 512 B per instruction) and four coalesced 32-bit stores.
    const long long nchunk = R >> 9;
    const long long wstride = static_cast<long long>(gridDim.x) * (THREADS / 32);
    long long ch = static_cast<long long>(blockIdx.x) * (THREADS / 32) + warp;
    uint4 p[4];
    bool has = ch < nchunk;
    auto load_chunk = [&](long long ci, uint4 (&dst)[4]) {
        const uint4* src = reinterpret_cast<const uint4*>(req) + 128 * ci + lane;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            int4 t = ld_stream_v4(reinterpret_cast<const int32_t*>(src + 32 * u));
            dst[u] = make_uint4(t.x, t.y, t.z, t.w);
        }
    };
    if (has) load_chunk(ch, p);

    const int D = st->D;
    uint32_t K[DT];
#pragma unroll
    for (int j = 0; j < DT; j += 4) {
        const uint4 k4 = *reinterpret_cast<const uint4*>(&st->sorted_k[j]);
        K[j] = k4.x; K[j + 1] = k4.y; K[j + 2] = k4.z; K[j + 3] = k4.w;
    }
    const uint32_t gx = st->cand_xor, gm = st->cand_mask;
    int32_t* tile = s.sDevTile[warp];
    for (int j = lane; j <= DT; j += 32) tile[j] = j < DT ? st->sorted_dev[j] : -1;
    hist_zero<DT, THREADS>(s, warp, lane);
    __syncwarp();

    auto decide = [&](uint32_t pw) -> uint32_t {
        // core << 18 | mem  ->  core << 24 | mem << 5; out-of-format words fail every guard
        const uint32_t q = (pw >> 25) ? (127u << 24) : (((pw & ~0x3FFFFu) << 6) | ((pw & 0x3FFFFu) << 5));
        const uint32_t best = first_feasible<DT>(K, q, gx, gm);
        const int32_t idx = tile[best];
        hist_add<DT, THREADS>(s, warp, lane, idx, static_cast<int32_t>((pw >> 18) & 127u), static_cast<int32_t>(pw & 0x3FFFFu));
        return static_cast<uint32_t>(idx) & 0xffu;
    };
    auto decide4 = [&](const uint4& v) -> uint32_t {
        return decide(v.x) | (decide(v.y) << 8) | (decide(v.z) << 16) | (decide(v.w) << 24);
    };
    while (has) {
        const long long cn = ch + wstride;
        const bool nhas = cn < nchunk;
        uint4 np[4];
        if (nhas) load_chunk(cn, np);
        uint32_t* out32 = reinterpret_cast<uint32_t*>(out_idx8) + 128 * ch + lane;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t r4 = decide4(p[u]);
            asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" ::"l"(out32 + 32 * u), "r"(r4) : "memory");
        }
        ch = cn;
        has = nhas;
#pragma unroll
        for (int u = 0; u < 4; ++u) p[u] = np[u];
    }
    // ragged tail: R % 512 rows, scalar, spread over the first CTA
    if (blockIdx.x == 0) {
        for (long long r = (nchunk << 9) + tid; r < R; r += THREADS) out_idx8[r] = static_cast<signed char>(decide(req[r]));
    }
    snapshot_epilogue<DT, THREADS>(s, st, D, delta_out, table_out, flags, slot_step);
    if (late) pdl_wait();
}

// The north-star's literal formulation: every (device, request) pair is scored
// with the spec's packed key (lc << 24 | lm << 6 | d) against a shared-memory
// tile of the table and the row is reduced with a running min.  Kept as an
// independent second device implementation (tests compare the two) and as the
// baseline the sorted variant is measured against.
template <int DT, int THREADS>
__global__ void __launch_bounds__(THREADS)
bestfit_grid_kernel(DevState* __restrict__ st, const int32_t* __restrict__ req_core,
                    const int32_t* __restrict__ req_mem, long long R, int32_t* __restrict__ out_idx,
                    long long* __restrict__ delta_out, int32_t* __restrict__ table_out, int flags, unsigned long long slot_step) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    auto& s = *reinterpret_cast<SnapSmem<DT, THREADS>*>(smem_raw);
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    pdl_wait();  // never triggers early: the literal variant keeps plain stream semantics
    const int D = st->D;
    hist_zero<DT, THREADS>(s, warp, lane);
    if (tid < D) {
        s.sFc[tid] = st->free_core[tid];
        s.sFm[tid] = st->free_mem[tid];
    }
    __syncthreads();

    auto decide = [&](int32_t core, int32_t mem) -> int32_t {
        int32_t best = 0x7fffffff;
        const bool valid = (core | mem) >= 0;
        for (int d = 0; d < D; ++d) {
            const int32_t lc = s.sFc[d] - core;
            const int32_t lm = s.sFm[d] - mem;
            const int32_t key = (lc << 24) | (lm << 6) | d;
            best = (valid && (lc | lm) >= 0) ? min(best, key) : best;
        }
        const int32_t idx = best == 0x7fffffff ? -1 : (best & 63);
        hist_add<DT, THREADS>(s, warp, lane, idx, core, mem);
        return idx;
    };

    const long long nvec = R >> 2;
    const long long stride = static_cast<long long>(gridDim.x) * THREADS;
    for (long long v = static_cast<long long>(blockIdx.x) * THREADS + tid; v < nvec; v += stride) {
        const int4 c = ld_stream_v4(req_core + 4 * v);
        const int4 m = ld_stream_v4(req_mem + 4 * v);
        int4 r;
        r.x = decide(c.x, m.x);
        r.y = decide(c.y, m.y);
        r.z = decide(c.z, m.z);
        r.w = decide(c.w, m.w);
        st_stream_v4(out_idx + 4 * v, r);
    }
    if (blockIdx.x == 0 && tid < static_cast<int>(R & 3)) {
        const long long r = (nvec << 2) + tid;
        out_idx[r] = decide(req_core[r], req_mem[r]);
    }
    snapshot_epilogue<DT, THREADS>(s, st, D, delta_out, table_out, flags, slot_step);
}

// =============================================================================
// Lookup-table scan for large D (EGPU_VARIANT_LUT; AUTO picks it for D > 16)
// =============================================================================
//
// The register-resident scan above costs 3.5 instructions per (request, device) pair:
// fine for D = 8 (HBM-bound), ALU-bound by 4x at D = 64.  The lookup form needs three
// shared-memory reads and ~20 instructions per request whatever D is.

// Builds DevLut from the sorted view in DevState.  One CTA; runs after every table change.
__global__ void __launch_bounds__(256)
lut_build_kernel(const DevState* __restrict__ st, DevLut* __restrict__ lut) {
    __shared__ uint32_t sFm[kMaxD], sFcs[kMaxD], sV[kMaxD];
    __shared__ int sFirst[kMaxD], sRidx[kMaxD], sNv;
    const int tid = threadIdx.x;
    const int D = st->D;
    if (tid < kMaxD) {
        const uint32_t k = tid < D ? st->sorted_k[tid] : 0u;
        sFm[tid] = (k >> 5) & 0x3FFFFu;
        sFcs[tid] = (k >> 24) & 0x7Fu;
        sV[tid] = 0xFFFFFFFFu;
    }
    __syncthreads();
    if (tid < D) {  // first occurrence of its fm value?
        int first = 1;
        for (int k = 0; k < tid; ++k) first &= (sFm[k] != sFm[tid]);
        sFirst[tid] = first;
    }
    __syncthreads();
    if (tid < D) {  // ridx = number of distinct values below mine
        int r = 0;
        for (int k = 0; k < D; ++k) r += (sFirst[k] && sFm[k] < sFm[tid]);
        sRidx[tid] = r;
        sV[r] = sFm[tid];
    }
    if (tid == 0) {
        int nv = 0;
        for (int k = 0; k < D; ++k) nv += sFirst[k];
        sNv = nv;
        lut->nv = nv;
    }
    __syncthreads();
    const int nv = sNv;
    if (tid < kMaxD) lut->v[tid] = sV[tid];
    if (tid < 128) {  // start[c] = first sorted position with fc >= c
        int n = 0;
        for (int k = 0; k < D; ++k) n += (sFcs[k] < static_cast<uint32_t>(tid));
        lut->start[tid] = static_cast<uint8_t>(n);
    }
    if (tid < kLutStride) {  // column r of a[][]: walk the suffixes from the back
        const int r = tid;
        uint8_t cur = 0xFF;
        for (int srow = kMaxD; srow >= 0; --srow) {
            if (srow < D && sRidx[srow] >= r) cur = static_cast<uint8_t>(st->sorted_dev[srow]);
            if (srow > D) cur = 0xFF;
            lut->a[srow * kLutStride + r] = cur;
        }
    }
    for (int b = tid; b < kLutBuckets; b += blockDim.x) {
        const uint32_t lo_v = static_cast<uint32_t>(b) << 6, hi_v = lo_v + 64u;
        int lo = 0, hi = 0;
        for (int k = 0; k < nv; ++k) {
            lo += (sV[k] < lo_v);
            hi += (sV[k] < hi_v);
        }
        lut->bucket[b] = static_cast<uint16_t>(lo | ((hi - lo) << 8));
    }
}

// Demand sums of the lookup scan: lane-private like SnapSmem::hist, but SHARE lanes share
// one accumulator (SHARE = 1, 2 or 4) and take turns, SHARE phases per update.  D = 64
// with SHARE = 1 costs 16.6 KB per warp, which caps an SM at 8-12 warps; sharing trades
// a few issue slots (the scan is nowhere near ALU-bound) for occupancy.
template <int THREADS, int SHARE>
struct LutSmem {
    DevLut lut;
    unsigned long long sWarpAcc[THREADS / 32][2 * kMaxD];
    int32_t sFc[kMaxD], sFm[kMaxD], sPosDev[kMaxD];
    int sLast;
    alignas(16) unsigned long long hist[THREADS / 32][kMaxD + 1][32 / SHARE];  // zeroed with 128-bit stores
};

template <int THREADS, int SHARE>
__global__ void __launch_bounds__(THREADS)
bestfit_lut_kernel(DevState* __restrict__ st, const int32_t* __restrict__ req_core,
                   const int32_t* __restrict__ req_mem, long long R, int32_t* __restrict__ out_idx,
                   long long* __restrict__ delta_out, int32_t* __restrict__ table_out, int flags, unsigned long long slot_step,
                   const DevLut* __restrict__ glut) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    auto& sm = *reinterpret_cast<LutSmem<THREADS, SHARE>*>(smem_raw);
    constexpr int LW = 32 / SHARE;  // accumulator columns per warp
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int col = lane & (LW - 1);
    const int phase = lane / LW;
    const bool late = (flags & kFlagLateWait) != 0;
    if (!late) pdl_wait();
    if (flags & kFlagEarlyTrigger) pdl_trigger();

    const long long nvec = R >> 2;
    const long long stride = static_cast<long long>(gridDim.x) * THREADS;
    long long v = static_cast<long long>(blockIdx.x) * THREADS + tid;
    int4 c0 = make_int4(0, 0, 0, 0), m0 = c0, c1 = c0, m1 = c0;
    bool has0 = v < nvec, has1 = (v + stride) < nvec;
    if (has0) {
        c0 = ld_stream_v4(req_core + 4 * v);
        m0 = ld_stream_v4(req_mem + 4 * v);
    }
    if (has1) {
        c1 = ld_stream_v4(req_core + 4 * (v + stride));
        m1 = ld_stream_v4(req_mem + 4 * (v + stride));
    }
    // shared-memory tile of the lookup tables (12.8 KB, L2-resident source)
    {
        const uint4* src = reinterpret_cast<const uint4*>(glut);
        uint4* dst = reinterpret_cast<uint4*>(&sm.lut);
        for (int i = tid; i < static_cast<int>(sizeof(DevLut) / 16); i += THREADS) dst[i] = src[i];
    }
    const int D = st->D;
    {   // zero this warp's accumulators with 128-bit stores
        uint4* hz = reinterpret_cast<uint4*>(&sm.hist[warp][0][0]);
        constexpr int n16 = (kMaxD + 1) * LW * 8 / 16;
        for (int i = lane; i < n16; i += 32) hz[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();
    const DevLut& L = sm.lut;

    // lookups for one request: no data-dependent loop on the common path
    auto lookup = [&](int32_t core, int32_t mem) -> int32_t {
        const uint32_t c = min(static_cast<uint32_t>(core), 127u);
        const uint32_t m = min(static_cast<uint32_t>(mem), 1u << 18);
        const uint32_t srow = L.start[c];
        const uint32_t e = L.bucket[m >> 6];
        const uint32_t lo = e & 0xffu, n = e >> 8;
        uint32_t rank = lo + ((n != 0u) & (L.v[lo & 63u] < m));
        for (uint32_t i = 1; i < n; ++i) rank += (L.v[lo + i] < m);  // rare: several distinct fm in one 64 MiB bucket
        return static_cast<int32_t>(static_cast<int8_t>(L.a[srow * kLutStride + rank]));
    };
    // Demand sums for the four requests of one vector at once.  Requests of this thread that
    // chose the same device are merged first (the later one is redirected to the dummy row
    // with nothing to add), so the four read-modify-writes are independent and can be issued
    // as four loads, four adds, four stores per phase instead of four dependent chains.
    auto accumulate4 = [&](const int4& r, const int4& c, const int4& m) {
        auto val = [](int32_t core, int32_t mem) {
            return (static_cast<unsigned long long>(static_cast<uint32_t>(core)) << kAccShift) |
                   static_cast<unsigned long long>(static_cast<uint32_t>(mem));
        };
        int32_t i0 = r.x, i1 = r.y, i2 = r.z, i3 = r.w;
        unsigned long long v0 = val(c.x, m.x), v1 = val(c.y, m.y), v2 = val(c.z, m.z), v3 = val(c.w, m.w);
        if (i0 < 0) v0 = 0;  // infeasible rows carry out-of-domain values: keep the dummy row harmless
        if (i1 < 0) v1 = 0;
        if (i2 < 0) v2 = 0;
        if (i3 < 0) v3 = 0;
        if (i3 == i2) { v2 += v3; v3 = 0; i3 = -1; }
        if (i3 == i1) { v1 += v3; v3 = 0; i3 = -1; }
        if (i3 == i0) { v0 += v3; v3 = 0; i3 = -1; }
        if (i2 == i1) { v1 += v2; v2 = 0; i2 = -1; }
        if (i2 == i0) { v0 += v2; v2 = 0; i2 = -1; }
        if (i1 == i0) { v0 += v1; v1 = 0; i1 = -1; }
        unsigned long long* h0 = &sm.hist[warp][i0 + 1][col];
        unsigned long long* h1 = &sm.hist[warp][i1 + 1][col];
        unsigned long long* h2 = &sm.hist[warp][i2 + 1][col];
        unsigned long long* h3 = &sm.hist[warp][i3 + 1][col];
#pragma unroll
        for (int p = 0; p < SHARE; ++p) {
            if (SHARE == 1 || phase == p) {
                const unsigned long long a0 = *h0, a1 = *h1, a2 = *h2, a3 = *h3;
                *h0 = a0 + v0;
                *h1 = a1 + v1;
                *h2 = a2 + v2;
                *h3 = a3 + v3;  // several redirected requests may all hit the dummy row: its content is never read
            }
            if (SHARE > 1) __syncwarp();
        }
    };
    auto decide4 = [&](const int4& c, const int4& m) -> int4 {
        int4 r;
        r.x = lookup(c.x, m.x);
        r.y = lookup(c.y, m.y);
        r.z = lookup(c.z, m.z);
        r.w = lookup(c.w, m.w);
        accumulate4(r, c, m);
        return r;
    };
    while (__any_sync(0xffffffffu, has0)) {  // warp-uniform trip count: accumulate() synchronises the warp
        const long long vn = v + 2 * stride;
        const bool nhas0 = vn < nvec, nhas1 = (vn + stride) < nvec;
        int4 nc0 = make_int4(0, 0, 0, 0), nm0 = nc0, nc1 = nc0, nm1 = nc0;
        if (nhas0) {
            nc0 = ld_stream_v4(req_core + 4 * vn);
            nm0 = ld_stream_v4(req_mem + 4 * vn);
        }
        if (nhas1) {
            nc1 = ld_stream_v4(req_core + 4 * (vn + stride));
            nm1 = ld_stream_v4(req_mem + 4 * (vn + stride));
        }
        // lanes past the end carry core = mem = -1: infeasible, lands in the dummy row
        if (!has0) { c0 = make_int4(-1, -1, -1, -1); m0 = c0; }
        if (!has1) { c1 = make_int4(-1, -1, -1, -1); m1 = c1; }
        const int4 r0 = decide4(c0, m0);
        const int4 r1 = decide4(c1, m1);
        if (has0) st_stream_v4(out_idx + 4 * v, r0);
        if (has1) st_stream_v4(out_idx + 4 * (v + stride), r1);
        v = vn;
        has0 = nhas0;
        has1 = nhas1;
        c0 = nc0; m0 = nm0; c1 = nc1; m1 = nm1;
    }
    if (blockIdx.x == 0 && warp == 0) {  // ragged tail: R % 4 rows; whole warp takes part in accumulate()
        const bool mine = lane < static_cast<int>(R & 3);
        const long long r = (nvec << 2) + lane;
        const int32_t c = mine ? req_core[r] : -1, m = mine ? req_mem[r] : -1;
        const int32_t idx = lookup(c, m);
        accumulate4(make_int4(idx, -1, -1, -1), make_int4(c, 0, 0, 0), make_int4(m, 0, 0, 0));
        if (mine) out_idx[r] = idx;
    }
    // warp sums -> sWarpAcc, transposed: lane L adds up the LW columns of devices L and L + 32
    // (rotated start column: conflict-free), instead of three warp reductions per device
    __syncwarp();
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int d = lane + 32 * half;
        unsigned long long sc = 0, smem_sum = 0;
        if (d < D) {
#pragma unroll 4
            for (int k = 0; k < LW; ++k) {
                const unsigned long long hv = sm.hist[warp][d + 1][(k + lane) & (LW - 1)];
                sc += hv >> kAccShift;
                smem_sum += hv & ((1ull << kAccShift) - 1ull);
            }
        }
        sm.sWarpAcc[warp][d] = sc;
        sm.sWarpAcc[warp][kMaxD + d] = smem_sum;
    }
    __syncthreads();
    epilogue_publish<THREADS / 32>(&sm.sWarpAcc[0][0], 2 * kMaxD, 0, kMaxD, sm.sFc, sm.sFm, sm.sPosDev, &sm.sLast, st, D,
                                   delta_out, table_out, flags, slot_step);
    if (late) pdl_wait();
}

// Multi-GPU step 2: table' = table - sum over ranks of their demand vectors.
__global__ void __launch_bounds__(kMaxD)
apply_deltas_kernel(DevState* __restrict__ st, const long long* __restrict__ deltas,
                    int G, int32_t* __restrict__ table_out, int commit) {
    __shared__ int32_t sFc[kMaxD], sFm[kMaxD], sPosDev[kMaxD];
    const int D = st->D;
    const int d = threadIdx.x;
    if (d < D) {
        long long dc = 0, dm = 0;
        for (int g = 0; g < G; ++g) {
            dc += deltas[static_cast<long long>(g) * 2 * D + d];
            dm += deltas[static_cast<long long>(g) * 2 * D + D + d];
        }
        const long long nc = static_cast<long long>(st->free_core[d]) - dc;
        const long long nm = static_cast<long long>(st->free_mem[d]) - dm;
        const int32_t over = (nc < 0 || nm < 0) ? 1 : 0;
        if (table_out) {
            table_out[d] = sat_i32(nc);
            table_out[D + d] = sat_i32(nm);
            table_out[2 * D + d] = over;
        }
        if (commit) {
            const int32_t cc = nc < 0 ? 0 : static_cast<int32_t>(nc);
            const int32_t cm = nm < 0 ? 0 : static_cast<int32_t>(nm);
            st->free_core[d] = cc;
            st->free_mem[d] = cm;
            st->oversub[d] |= over;
            sFc[d] = cc;
            sFm[d] = cm;
        }
    }
    if (commit) resort_table_cta(st, D, sFc, sFm, sPosDev, d);
}

// Multi-GPU step 2, peer-memory form: wait until every rank's demand vector of `step` has
// landed in THIS rank's exchange buffer, then apply their sum.  One CTA.  The spin gives up
// after ~2 s (a rank died): DevState::peer_timeout records it and the table is left alone.
struct ApplyOuts {
    int32_t* table_out[8];
};

__global__ void __launch_bounds__(kMaxD)
apply_peers_kernel(DevState* __restrict__ st, unsigned long long first_step_plus1, int nsteps, ApplyOuts outs, int commit) {
    __shared__ int32_t sFc[kMaxD], sFm[kMaxD], sPosDev[kMaxD];
    __shared__ int sOk;
    const int D = st->D;
    const int d = threadIdx.x;
    const int world = st->peer.world, me = st->peer.rank;
    // running table across the steps of this launch (only installed when commit is set)
    long long cur_c = d < D ? st->free_core[d] : 0, cur_m = d < D ? st->free_mem[d] : 0;
    int32_t sticky = 0;
    for (int k = 0; k < nsteps; ++k) {
        const unsigned long long step_plus1 = first_step_plus1 + k;
        XchgRow* rows = st->peer.buf[me]->slot[(step_plus1 - 1) % kXchgSlots];
        if (d == 0) sOk = 1;
        __syncthreads();
        if (d < world) {
            unsigned long long f = 0;
            const long long t0 = clock64();
            for (;;) {
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(f) : "l"(&rows[d].flag) : "memory");
                if (f == step_plus1) break;
                if (clock64() - t0 > 4000000000ll) {
                    sOk = 0;
                    break;
                }
                __nanosleep(100);
            }
        }
        __syncthreads();
        if (!sOk) {
            if (d == 0) st->peer_timeout = step_plus1;
            return;
        }
        long long dc = 0, dm = 0;
        if (d < D) {
            for (int g = 0; g < world; ++g) {
                dc += rows[g].delta[d];
                dm += rows[g].delta[D + d];
            }
        }
        __syncthreads();
        // consume the flags: a replayed CUDA graph pushes the same step numbers again, and a
        // stale flag must not look like the new one.  (The slot is not written again before
        // this rank has applied 16 more steps, see the header.)
        if (d < world) rows[d].flag = 0ull;
        if (d < D) {
            const long long nc = cur_c - dc, nm = cur_m - dm;
            const int32_t over = (nc < 0 || nm < 0) ? 1 : 0;
            if (outs.table_out[k]) {
                outs.table_out[k][d] = sat_i32(nc);
                outs.table_out[k][D + d] = sat_i32(nm);
                outs.table_out[k][2 * D + d] = over;
            }
            if (commit) {  // the next step of this launch is applied on top of this one
                cur_c = nc < 0 ? 0 : nc;
                cur_m = nm < 0 ? 0 : nm;
                sticky |= over;
            }
        }
    }
    if (commit) {
        if (d < D) {
            st->free_core[d] = static_cast<int32_t>(cur_c);
            st->free_mem[d] = static_cast<int32_t>(cur_m);
            st->oversub[d] |= sticky;
            sFc[d] = static_cast<int32_t>(cur_c);
            sFm[d] = static_cast<int32_t>(cur_m);
        }
        resort_table_cta(st, D, sFc, sFm, sPosDev, d);
    }
}

// =============================================================================
// Synthetic request generator (same counter RNG as synth.py)
// =============================================================================
__device__ __forceinline__ unsigned long long mix64(unsigned long long seed, unsigned long long stream,
                                                    unsigned long long i) {
    unsigned long long z = seed * 0x9E3779B97F4A7C15ull + stream * 0xD1B54A32D192ED03ull + i;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ int32_t uniform_i32(unsigned long long seed, unsigned long long stream,
                                               unsigned long long i, int lo, int hi) {
    const unsigned long long z = mix64(seed, stream, i);
    const unsigned long long n = static_cast<unsigned long long>(hi - lo + 1);
    return lo + static_cast<int32_t>(((z >> 32) * n) >> 32);
}

__global__ void synth_requests_kernel(int dist, unsigned long long seed, long long first_row, long long R,
                                      int32_t* __restrict__ req_core, int32_t* __restrict__ req_mem) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long k = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; k < R; k += stride) {
        const unsigned long long r = static_cast<unsigned long long>(first_row + k);
        int32_t core, mem;
        if (dist == 2) {
            const int ci = uniform_i32(seed, 2, r, 0, 5);
            const int mi = uniform_i32(seed, 3, r, 0, 6);
            core = ci == 0 ? 5 : ci == 1 ? 10 : ci == 2 ? 20 : ci == 3 ? 25 : ci == 4 ? 50 : 100;
            mem = 256 << mi;
        } else {
            core = uniform_i32(seed, 2, r, 1, 100);
            mem = uniform_i32(seed, 3, r, 1, dist == 3 ? 65536 : 24576);
            if ((r & 15ull) == 15ull) {
                if (((r >> 4) & 1ull) == 0ull) core = 101;
                else mem = 183359 + 1;
            }
        }
        req_core[k] = core;
        req_mem[k] = mem;
    }
}

// =============================================================================
// Sequential mode: one warp.  D <= 8: table in registers (replay8_kernel, below the general
// one); otherwise lane = device (two per lane when D > 32)
// =============================================================================
//
// Request k sees the table after k-1: a serial dependence chain, so there is no
// bandwidth roofline here — the figure of merit is cycles per event.  The warp
// loads 32 events at a time (coalesced), broadcasts them one by one with
// shuffles, scores the current table with one packed key per lane and reduces
// with CREDUX.MIN (__reduce_min_sync).  `live` (device currently held by each
// ALLOC event, -1 otherwise) sits in shared memory when it fits, else in HBM;
// only lane 0 touches it, so program order gives consistency.
constexpr int kReplaySmemEvents = 200 * 1024;

__global__ void __launch_bounds__(32)
replay_kernel(DevState* __restrict__ st, const int32_t* __restrict__ kind, const int32_t* __restrict__ ev_a,
              const int32_t* __restrict__ ev_b, long long E, int32_t* __restrict__ out_idx,
              signed char* __restrict__ live_global) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    signed char* live = (E <= kReplaySmemEvents) ? reinterpret_cast<signed char*>(smem_raw) : live_global;
    const int lane = threadIdx.x;
    const int D = st->D;
    const int d0 = lane, d1 = lane + 32;
    int32_t fc0 = d0 < D ? st->free_core[d0] : -1;
    int32_t fm0 = d0 < D ? st->free_mem[d0] : -1;
    int32_t fc1 = d1 < D ? st->free_core[d1] : -1;
    int32_t fm1 = d1 < D ? st->free_mem[d1] : -1;

    for (long long base = 0; base < E; base += 32) {
        const long long i = base + lane;
        int32_t k = -1, a = 0, b = 0, ta = 0, tb = 0;
        bool tvalid = false;
        if (i < E) {
            k = kind[i];
            a = ev_a[i];
            b = ev_b[i];
            if (k == 1 && a >= 0 && a < i) {  // gather the released event's request now
                tvalid = kind[a] == 0;
                ta = ev_a[a];
                tb = ev_b[a];
            }
        }
        int32_t my_out = -1;
        const int n = (E - base) < 32 ? static_cast<int>(E - base) : 32;
        for (int j = 0; j < n; ++j) {
            const int32_t kj = __shfl_sync(0xffffffffu, k, j);
            const int32_t aj = __shfl_sync(0xffffffffu, a, j);
            const int32_t bj = __shfl_sync(0xffffffffu, b, j);
            int32_t res = -1;
            if (kj == 0) {
                const int32_t lc0 = fc0 - aj, lm0 = fm0 - bj;
                const int32_t lc1 = fc1 - aj, lm1 = fm1 - bj;
                const bool valid = (aj | bj) >= 0;
                int32_t key = 0x7fffffff;
                if (valid && (lc0 | lm0) >= 0 && fc0 >= 0) key = (lc0 << 24) | (lm0 << 6) | d0;
                if (valid && (lc1 | lm1) >= 0 && fc1 >= 0) key = min(key, (lc1 << 24) | (lm1 << 6) | d1);
                const int32_t best = __reduce_min_sync(0xffffffffu, key);
                if (best != 0x7fffffff) {
                    res = best & 63;
                    if (res == d0) { fc0 -= aj; fm0 -= bj; }
                    if (res == d1) { fc1 -= aj; fm1 -= bj; }
                }
                if (lane == 0) live[base + j] = static_cast<signed char>(res);
            } else {
                const bool tv = __shfl_sync(0xffffffffu, static_cast<int>(tvalid), j) != 0;
                const int32_t taj = __shfl_sync(0xffffffffu, ta, j);
                const int32_t tbj = __shfl_sync(0xffffffffu, tb, j);
                int32_t dev = -1;
                if (lane == 0) {
                    live[base + j] = -1;
                    if (kj == 1 && tv) {
                        dev = live[aj];
                        live[aj] = -1;
                    }
                }
                dev = __shfl_sync(0xffffffffu, dev, 0);
                if (dev >= 0) {
                    if (dev == d0) { fc0 += taj; fm0 += tbj; }
                    if (dev == d1) { fc1 += taj; fm1 += tbj; }
                }
                res = dev;
            }
            if (lane == j) my_out = res;
        }
        if (i < E) out_idx[i] = my_out;
    }
    __shared__ int32_t sFc[kMaxD], sFm[kMaxD], sPosDev[kMaxD];
    if (d0 < D) { st->free_core[d0] = fc0; st->free_mem[d0] = fm0; sFc[d0] = fc0; sFm[d0] = fm0; }
    if (d1 < D) { st->free_core[d1] = fc1; st->free_mem[d1] = fm1; sFc[d1] = fc1; sFm[d1] = fm1; }
    resort_table_cta(st, D, sFc, sFm, sPosDev, lane);
}

// Sequential mode for D <= 8: the whole table lives in the registers of every lane as packed
// compare words (guard | free_core | guard | free_mem | device).  K[d] - Q is at once the
// feasibility test (both guards survive), the ordering key of the spec ((lc, lm, d) with the
// guards as constant top bits) and the updated table word of the chosen device — so an ALLOC
// is 8 subtracts, 8 guard tests, a 3-input-min tree and 8 selects, with no cross-lane
// traffic on the dependence chain.  All lanes compute the same thing; lane 0 keeps `live`
// and the outputs.  Events are held 32 at a time in registers (lane j = event j of the chunk),
// broadcast with shuffles; the next chunk is prefetched while the current one is processed.
__global__ void __launch_bounds__(32)
replay8_kernel(DevState* __restrict__ st, const int32_t* __restrict__ kind, const int32_t* __restrict__ ev_a,
               const int32_t* __restrict__ ev_b, long long E, int32_t* __restrict__ out_idx,
               signed char* __restrict__ live_global) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    signed char* live = (E <= kReplaySmemEvents) ? reinterpret_cast<signed char*>(smem_raw) : live_global;
    const int lane = threadIdx.x;
    const int D = st->D;
    uint32_t K[8];
#pragma unroll
    for (int d = 0; d < 8; ++d)
        K[d] = d < D ? (pack_table_word(st->free_core[d], st->free_mem[d]) | static_cast<uint32_t>(d)) : kPadWord;

    auto fetch = [&](long long i, int32_t& k, uint32_t& q, uint32_t& qt, int32_t& t) {
        k = -1; q = 0; qt = 0; t = -1;
        if (i < E) {
            k = kind[i];
            const int32_t a = ev_a[i], b = ev_b[i];
            q = pack_request_word(a, b);
            if (k == 1 && a >= 0 && a < i && kind[a] == 0) {  // FREE of an earlier ALLOC: fetch its request now
                t = a;
                qt = pack_request_word(ev_a[a], ev_b[a]);
            }
        }
    };
    int32_t nk; uint32_t nq, nqt; int32_t nt;
    fetch(lane, nk, nq, nqt, nt);
    for (long long base = 0; base < E; base += 32) {
        // this chunk's events stay in registers (lane j holds event base + j) and are broadcast
        // with shuffles, which do not sit on the dependence chain; `live` is lane 0's alone
        const int32_t ck = nk, ct = nt;
        const uint32_t cq = nq, cqt = nqt;
        fetch(base + 32 + lane, nk, nq, nqt, nt);  // prefetch the next chunk
        const int n = (E - base) < 32 ? static_cast<int>(E - base) : 32;
        int32_t my_out = -1;
#pragma unroll 4
        for (int j = 0; j < n; ++j) {
            const int32_t kj = __shfl_sync(0xffffffffu, ck, j);
            const uint32_t q = __shfl_sync(0xffffffffu, cq, j);
            int32_t res = -1;
            if (kj == 0) {
                uint32_t w[8], key[8];
#pragma unroll
                for (int d = 0; d < 8; ++d) {
                    w[d] = K[d] - q;
                    key[d] = ((w[d] & kGuards) == kGuards) ? w[d] : 0xFFFFFFFFu;
                }
                const uint32_t best = __vimin3_u32(__vimin3_u32(key[0], key[1], key[2]), __vimin3_u32(key[3], key[4], key[5]),
                                                   min(key[6], key[7]));
                if (best != 0xFFFFFFFFu) {
                    res = static_cast<int32_t>(best & 31u);
#pragma unroll
                    for (int d = 0; d < 8; ++d) K[d] = (w[d] == best) ? w[d] : K[d];
                }
                if (lane == 0) live[base + j] = static_cast<signed char>(res);
            } else {
                const int32_t t = __shfl_sync(0xffffffffu, ct, j);
                const uint32_t qt = __shfl_sync(0xffffffffu, cqt, j);
                int32_t dev = -1;
                if (lane == 0) {
                    live[base + j] = -1;
                    if (t >= 0) {
                        dev = live[t];
                        live[t] = -1;
                    }
                }
                dev = __shfl_sync(0xffffffffu, dev, 0);
                if (dev >= 0) {
#pragma unroll
                    for (int d = 0; d < 8; ++d) K[d] = (d == dev) ? K[d] + qt : K[d];
                    res = dev;
                }
            }
            if (lane == j) my_out = res;
        }
        if (base + lane < E) out_idx[base + lane] = my_out;
    }
    __shared__ int32_t sFc[kMaxD], sFm[kMaxD], sPosDev[kMaxD];
    if (lane < D) {
        uint32_t k = 0;
#pragma unroll
        for (int d = 0; d < 8; ++d) k = (d == lane) ? K[d] : k;
        const int32_t fc = static_cast<int32_t>((k >> 24) & 0x7Fu), fm = static_cast<int32_t>((k >> 5) & 0x3FFFFu);
        st->free_core[lane] = fc;
        st->free_mem[lane] = fm;
        sFc[lane] = fc;
        sFm[lane] = fm;
    }
    resort_table_cta(st, D, sFc, sFm, sPosDev, lane);
}

}  // namespace egpu

// =============================================================================
// Host side: context + C ABI
// =============================================================================
using namespace egpu;

#include "egpu_ctx.h"

namespace {

template <int DT, int THREADS>
SnapLaunch make_launch(bool grid_variant) {
    SnapLaunch l;
    l.fn = grid_variant ? bestfit_grid_kernel<DT, THREADS> : bestfit_sorted_kernel<DT, THREADS>;
    l.threads = THREADS;
    l.smem = sizeof(SnapSmem<DT, THREADS>);
    l.ctas_per_sm = 0;
    return l;
}

SnapLaunch pick_launch(int D, bool grid_variant) {
    if (D <= 8) return make_launch<8, 256>(grid_variant);
    if (D <= 16) return make_launch<16, 256>(grid_variant);
    if (D <= 32) return make_launch<32, 256>(grid_variant);
    return make_launch<64, 128>(grid_variant);
}

// host copy of resort_table_cta: sorted view of a freshly set table
void fill_sorted(DevState& h) {
    const int D = h.D;
    for (int j = 0; j < kMaxD; ++j) {
        h.sorted_k[j] = kPadWord;
        h.sorted_dev[j] = -1;
    }
    unsigned long long packed = ~0ull;
    for (int d = 0; d < D; ++d) {
        const uint32_t mine = (static_cast<uint32_t>(h.free_core[d]) << 24) | (static_cast<uint32_t>(h.free_mem[d]) << 6) | d;
        int pos = 0;
        for (int k = 0; k < D; ++k) {
            const uint32_t other = (static_cast<uint32_t>(h.free_core[k]) << 24) | (static_cast<uint32_t>(h.free_mem[k]) << 6) | k;
            pos += other < mine;
        }
        h.sorted_k[pos] = pack_table_word(h.free_core[d], h.free_mem[d]) | (static_cast<uint32_t>(pos) & 31u);
        h.sorted_dev[pos] = d;
        if (pos < 8) packed = (packed & ~(0xffull << (8 * pos))) | (static_cast<unsigned long long>(d) << (8 * pos));
    }
    h.dev_packed = packed;
}

// user_flags: EGPU_F_COMMIT | EGPU_F_INPUTS_READY.  finalize = 0 only for the
// chunked host pipeline (accumulate demand sums across launches).
int launch_snapshot(egpu_ctx* ctx, const int32_t* d_rc, const int32_t* d_rm, int64_t R, int32_t* d_idx,
                    long long* d_delta, int32_t* d_table_out, int user_flags, bool finalize, cudaStream_t s,
                    int rpt_hint = 0, unsigned long long push_step_plus1 = 0) {
    const bool grid_variant = ctx->variant == EGPU_VARIANT_GRID;
    const bool lut_variant = ctx->variant == EGPU_VARIANT_LUT || (ctx->variant == EGPU_VARIANT_AUTO && ctx->D > 16);
    const int bucket = ctx->D <= 8 ? 0 : ctx->D <= 16 ? 1 : ctx->D <= 32 ? 2 : 3;
    SnapLaunch& l = ctx->snap[grid_variant ? 1 : (lut_variant ? 2 : 0)][bucket];
    if (l.ctas_per_sm == 0) {  // first use on this context: opt in to the shared-memory size, ask occupancy
        int per_sm = 0;
        if (lut_variant) {
            const int share = ctx->lut_share;
            if (ctx->lut_threads == 256) {
                l.lut_fn = share == 2 ? bestfit_lut_kernel<256, 2> : share == 8 ? bestfit_lut_kernel<256, 8> : bestfit_lut_kernel<256, 4>;
                l.threads = 256;
                l.smem = share == 2 ? sizeof(LutSmem<256, 2>) : share == 8 ? sizeof(LutSmem<256, 8>) : sizeof(LutSmem<256, 4>);
            } else {
                l.lut_fn = share == 1 ? bestfit_lut_kernel<128, 1> : share == 2 ? bestfit_lut_kernel<128, 2>
                           : share == 4 ? bestfit_lut_kernel<128, 4> : bestfit_lut_kernel<128, 8>;
                l.threads = 128;
                l.smem = share == 1 ? sizeof(LutSmem<128, 1>) : share == 2 ? sizeof(LutSmem<128, 2>)
                         : share == 4 ? sizeof(LutSmem<128, 4>) : sizeof(LutSmem<128, 8>);
            }
            EGPU_CUDA(ctx, cudaFuncSetAttribute(l.lut_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(l.smem)));
            EGPU_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, l.lut_fn, l.threads, l.smem));
        } else {
            l = pick_launch(ctx->D, grid_variant);
            EGPU_CUDA(ctx, cudaFuncSetAttribute(l.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(l.smem)));
            EGPU_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, l.fn, l.threads, l.smem));
        }
        l.ctas_per_sm = per_sm < 1 ? 1 : per_sm;
    }
    if (lut_variant && ctx->lut_dirty) {  // refresh the lookup tables on the launching stream
        lut_build_kernel<<<1, 256, 0, s>>>(ctx->d_state, ctx->d_lut);
        EGPU_CUDA(ctx, cudaGetLastError());
        ctx->launches += 1;
        ctx->lut_dirty = false;
        ctx->prev_is_scan = false;
    }
    int flags = (finalize ? kFlagFinalize : 0) | ((user_flags & EGPU_F_COMMIT) ? kFlagCommit : 0);
    // Programmatic dependent launch.  Every scan carries the PDL attribute, so the
    // hardware may schedule it while its predecessor is still running.  A fully
    // ordered launch waits (griddepcontrol.wait) before it touches anything.  A
    // pipelined launch (kFlagLateWait) runs scan and epilogue at once — its epilogue
    // state is its own slot — and waits only before exiting.  That is allowed when
    //  - the caller vouches its inputs were complete before the previous launch on
    //    this stream (EGPU_F_INPUTS_READY),
    //  - the previous launch was a scan of this context on the same stream that
    //    does not rewrite the table, and this one is a plain finalising scan,
    //  - its outputs (indices, demand sums, table') are disjoint from the outputs
    //    of every launch since the last fully ordered one, and
    //  - fewer than pipe_group launches have been issued since then, which bounds
    //    the launches in flight to pipe_group + 1 < kEpiSlots.
    egpu_ctx::Range mine[3] = {
        {reinterpret_cast<uintptr_t>(d_idx), reinterpret_cast<uintptr_t>(d_idx) + static_cast<uintptr_t>(R) * sizeof(int32_t)},
        {reinterpret_cast<uintptr_t>(d_delta), reinterpret_cast<uintptr_t>(d_delta) + (d_delta ? sizeof(long long) * 2 * ctx->D : 0)},
        {reinterpret_cast<uintptr_t>(d_table_out), reinterpret_cast<uintptr_t>(d_table_out) + (d_table_out ? sizeof(int32_t) * 3 * ctx->D : 0)}};
    bool pipelined = !grid_variant && finalize && (user_flags & EGPU_F_INPUTS_READY) && ctx->prev_is_scan &&
                     !ctx->prev_changes_table && ctx->prev_stream == s && ctx->group_len > 0 &&
                     ctx->group_len < ctx->pipe_group;
    for (int i = 0; pipelined && i < 3 * ctx->group_len; ++i)
        for (int k = 0; k < 3; ++k)
            if (mine[k].lo < ctx->group_out[i].hi && ctx->group_out[i].lo < mine[k].hi) pipelined = false;
    if (pipelined) flags |= kFlagLateWait;
    else ctx->group_len = 0;
    // Early trigger (griddepcontrol.launch_dependents before the work is done) only helps when
    // the next launch is another scan of a pipelined stream, so only those launches do it.
    // (Checked on B200, scripts/probes/pdl_event_probe.cu and scripts/eager_probe.py: events
    // and ordinary kernels enqueued after a PDL launch still wait for its completion.)
    if (user_flags & EGPU_F_INPUTS_READY) flags |= kFlagEarlyTrigger;
    const unsigned long long slot = (finalize ? (ctx->seq % kEpiSlots) : static_cast<unsigned long long>(kEpiSlots)) |
                                    (push_step_plus1 << 8);

    // Grid: one resident wave at most.  A lone launch wants every SM pulling at once
    // (8 rows per thread, one trip); launches of a pipelined stream overlap each
    // other, so a smaller grid with more rows per thread (48) costs fewer CTA
    // launches, fewer atomics and leaves room for the neighbours — measured best on
    // B200 at R = 1M.  The zero-copy path passes its own hint (see egpu_bestfit_batch).
    const int64_t nvec = R >> 2;
    int rpt = 8;
    if (user_flags & EGPU_F_INPUTS_READY) rpt = (lut_variant || ctx->D <= 16) ? 48 : 8;  // measured, scripts/tune_*.sh
    else if (lut_variant) rpt = 32;  // the lookup scan has a 13 KB per-CTA table tile to amortise
    if (rpt_hint > 0) rpt = rpt_hint;
    if (ctx->rows_per_thread > 0) rpt = ctx->rows_per_thread;
    if (grid_variant) rpt = 4;
    const int64_t per_cta = static_cast<int64_t>(l.threads) * ((rpt + 3) / 4);
    int64_t want = (nvec + per_cta - 1) / per_cta;
    int per_sm = l.ctas_per_sm;
    if (ctx->ctas_per_sm_cap > 0 && ctx->ctas_per_sm_cap < per_sm) per_sm = ctx->ctas_per_sm_cap;
    const int64_t cap = static_cast<int64_t>(ctx->sm_count) * per_sm;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    // lane-private sums hold 2^19 rows per lane (kAccShift): keep rows/thread below that
    const int64_t rows_per_thread = R / (want * l.threads) + 8;
    if (rows_per_thread >= (1ll << 19)) return EGPU_ERR_INVALID;

    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(static_cast<unsigned>(want));
    cfg.blockDim = dim3(static_cast<unsigned>(l.threads));
    cfg.dynamicSmemBytes = l.smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (lut_variant)
        EGPU_CUDA(ctx, cudaLaunchKernelEx(&cfg, l.lut_fn, ctx->d_state, d_rc, d_rm, static_cast<long long>(R), d_idx,
                                          d_delta, d_table_out, flags, slot, static_cast<const DevLut*>(ctx->d_lut)));
    else
        EGPU_CUDA(ctx, cudaLaunchKernelEx(&cfg, l.fn, ctx->d_state, d_rc, d_rm, static_cast<long long>(R), d_idx,
                                          d_delta, d_table_out, flags, slot));
    if (flags & kFlagCommit) ctx->lut_dirty = true;
    ctx->launches += 1;
    ctx->seq += 1;
    for (int k = 0; k < 3; ++k) ctx->group_out[3 * ctx->group_len + k] = mine[k];
    ctx->group_len += 1;
    ctx->prev_is_scan = finalize;
    ctx->prev_changes_table = (flags & kFlagCommit) != 0;
    ctx->prev_stream = s;
    return EGPU_OK;
}

using PackedKernel = void (*)(DevState*, const uint32_t*, long long, signed char*, long long*, int32_t*, int, unsigned long long);

// Packed-format scan: always a fully ordered launch (it serves the synchronous host path).
int launch_packed(egpu_ctx* ctx, const uint32_t* d_req, int64_t R, signed char* d_idx8, long long* d_delta,
                  int32_t* d_table_out, int user_flags, cudaStream_t s, int rows_per_thread) {
    const int bucket = ctx->D <= 8 ? 0 : ctx->D <= 16 ? 1 : ctx->D <= 32 ? 2 : 3;
    static const PackedKernel fns[4] = {bestfit_sorted_packed_kernel<8, 256>, bestfit_sorted_packed_kernel<16, 256>,
                                        bestfit_sorted_packed_kernel<32, 256>, bestfit_sorted_packed_kernel<64, 128>};
    static const int threads[4] = {256, 256, 256, 128};
    static const size_t smem[4] = {sizeof(SnapSmem<8, 256>), sizeof(SnapSmem<16, 256>), sizeof(SnapSmem<32, 256>),
                                   sizeof(SnapSmem<64, 128>)};
    int& per_sm = ctx->packed_ctas_per_sm[bucket];
    if (per_sm == 0) {
        EGPU_CUDA(ctx, cudaFuncSetAttribute(fns[bucket], cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem[bucket])));
        int n = 0;
        EGPU_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fns[bucket], threads[bucket], smem[bucket]));
        per_sm = n < 1 ? 1 : n;
    }
    const int64_t nchunk = R >> 9;  // 512 requests per warp trip
    const int64_t per_cta = static_cast<int64_t>(threads[bucket] / 32) * ((rows_per_thread + 15) / 16);
    int64_t want = (nchunk + per_cta - 1) / per_cta;
    const int64_t cap = static_cast<int64_t>(ctx->sm_count) * per_sm;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    if (R / (want * threads[bucket]) + 16 >= (1ll << 19)) return EGPU_ERR_INVALID;
    const int flags = kFlagFinalize | ((user_flags & EGPU_F_COMMIT) ? kFlagCommit : 0);
    const unsigned long long slot = ctx->seq % kEpiSlots;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(static_cast<unsigned>(want));
    cfg.blockDim = dim3(static_cast<unsigned>(threads[bucket]));
    cfg.dynamicSmemBytes = smem[bucket];
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    EGPU_CUDA(ctx, cudaLaunchKernelEx(&cfg, fns[bucket], ctx->d_state, d_req, static_cast<long long>(R), d_idx8, d_delta,
                                      d_table_out, flags, slot));
    ctx->launches += 1;
    ctx->seq += 1;
    ctx->group_len = 0;
    ctx->prev_is_scan = false;  // the int32 scans do not pipeline behind this one
    if (flags & kFlagCommit) ctx->lut_dirty = true;
    return EGPU_OK;
}

int ensure_staging(egpu_ctx* ctx, int64_t rows) {
    if (rows <= ctx->d_cap_rows) return EGPU_OK;
    int64_t cap = ctx->d_cap_rows ? ctx->d_cap_rows : 1024;
    while (cap < rows) cap *= 2;
    if (ctx->d_req_core) cudaFree(ctx->d_req_core);
    if (ctx->d_req_mem) cudaFree(ctx->d_req_mem);
    if (ctx->d_idx) cudaFree(ctx->d_idx);
    ctx->d_req_core = ctx->d_req_mem = ctx->d_idx = nullptr;
    ctx->d_cap_rows = 0;
    EGPU_CUDA(ctx, cudaMalloc(&ctx->d_req_core, sizeof(int32_t) * cap));
    EGPU_CUDA(ctx, cudaMalloc(&ctx->d_req_mem, sizeof(int32_t) * cap));
    EGPU_CUDA(ctx, cudaMalloc(&ctx->d_idx, sizeof(int32_t) * cap));
    ctx->d_cap_rows = cap;
    return EGPU_OK;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// device-visible alias of a pinned (mapped) host allocation, nullptr for anything else
void* mapped_alias(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return nullptr;
    }
    return (a.type == cudaMemoryTypeHost) ? a.devicePointer : nullptr;
}

}  // namespace

extern "C" {

int egpu_abi_version(void) { return 1000; }

const char* egpu_strerror(int code) {
    switch (code) {
        case EGPU_OK: return "ok";
        case EGPU_ERR_INVALID: return "invalid argument";
        case EGPU_ERR_NO_DEVICE: return "no usable CUDA device (this library has no CPU fallback)";
        case EGPU_ERR_CUDA: return "CUDA runtime error";
        case EGPU_ERR_NOMEM: return "out of memory";
        case EGPU_ERR_NO_TABLE: return "capacity table not set";
        case EGPU_ERR_STATE: return "call not valid in the current state";
        case EGPU_ERR_PARSE: return "malformed device id";
        case EGPU_ERR_UNSAT: return "preferred allocation cannot be satisfied";
        default: return "unknown error";
    }
}

int egpu_ctx_create(int cuda_device, egpu_ctx** out) {
    if (!out) return EGPU_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        (void)cudaGetLastError();
        return EGPU_ERR_NO_DEVICE;
    }
    if (cuda_device < 0 || cuda_device >= n) return EGPU_ERR_INVALID;
    egpu_ctx* ctx = new (std::nothrow) egpu_ctx();
    if (!ctx) return EGPU_ERR_NOMEM;
    ctx->dev = cuda_device;
    int rc = [&]() -> int {
        EGPU_CUDA(ctx, cudaSetDevice(cuda_device));
        cudaDeviceProp prop;
        EGPU_CUDA(ctx, cudaGetDeviceProperties(&prop, cuda_device));
        ctx->sm_count = prop.multiProcessorCount;
        if (const char* e = std::getenv("EGPU_CTAS_PER_SM")) ctx->ctas_per_sm_cap = std::atoi(e);
        if (const char* e = std::getenv("EGPU_ROWS_PER_THREAD")) ctx->rows_per_thread = std::atoi(e);
        if (const char* e = std::getenv("EGPU_REPLAY_GENERAL")) ctx->replay_general = std::atoi(e) != 0;
        if (const char* e = std::getenv("EGPU_LUT_THREADS")) ctx->lut_threads = std::atoi(e) == 128 ? 128 : 256;
        if (const char* e = std::getenv("EGPU_LUT_SHARE")) {
            const int v = std::atoi(e);
            ctx->lut_share = (v == 1 || v == 2 || v == 8) ? v : 4;
        }
        if (const char* e = std::getenv("EGPU_PIPE_GROUP")) {
            const int g = std::atoi(e);
            ctx->pipe_group = g < 1 ? 1 : (g > kPipeGroupMax ? kPipeGroupMax : g);
        }
        EGPU_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        EGPU_CUDA(ctx, cudaMalloc(&ctx->d_state, sizeof(DevState)));
        EGPU_CUDA(ctx, cudaMemsetAsync(ctx->d_state, 0, sizeof(DevState), ctx->stream));
        EGPU_CUDA(ctx, cudaMalloc(&ctx->d_lut, sizeof(DevLut)));
        EGPU_CUDA(ctx, cudaMalloc(&ctx->d_xchg, sizeof(XchgBuf)));
        EGPU_CUDA(ctx, cudaMemsetAsync(ctx->d_xchg, 0, sizeof(XchgBuf), ctx->stream));
        EGPU_CUDA(ctx, cudaMalloc(&ctx->d_delta, sizeof(long long) * 2 * kMaxD));
        EGPU_CUDA(ctx, cudaMalloc(&ctx->d_table_out, sizeof(int32_t) * 3 * kMaxD));
        EGPU_CUDA(ctx, cudaMallocHost(&ctx->h_delta, sizeof(long long) * 2 * kMaxD));
        EGPU_CUDA(ctx, cudaHostGetDevicePointer(reinterpret_cast<void**>(&ctx->h_delta_dev), ctx->h_delta, 0));
        if (const char* e = std::getenv("EGPU_NO_ZERO_COPY")) ctx->no_zero_copy = std::atoi(e) != 0;
        EGPU_CUDA(ctx, cudaMallocHost(&ctx->h_table, sizeof(int32_t) * 3 * kMaxD));
        EGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return EGPU_OK;
    }();
    if (rc != EGPU_OK) {
        egpu_ctx_destroy(ctx);
        return rc;
    }
    *out = ctx;
    return EGPU_OK;
}

void egpu_ctx_destroy(egpu_ctx* ctx) {
    if (!ctx) return;
    if (ctx->dev >= 0) cudaSetDevice(ctx->dev);
    if (ctx->stream) {
        cudaStreamSynchronize(ctx->stream);
        cudaStreamDestroy(ctx->stream);
    }
    for (int r = 0; r < kMaxRanks; ++r)
        if (ctx->peer_open[r]) cudaIpcCloseMemHandle(ctx->peer_open[r]);
    cudaFree(ctx->d_xchg);
    cudaFree(ctx->arena);
    cudaFree(ctx->d_state);
    cudaFree(ctx->d_lut);
    cudaFree(ctx->d_req_core);
    cudaFree(ctx->d_req_mem);
    cudaFree(ctx->d_idx);
    cudaFree(ctx->d_delta);
    cudaFree(ctx->d_table_out);
    if (ctx->h_delta) cudaFreeHost(ctx->h_delta);
    if (ctx->h_table) cudaFreeHost(ctx->h_table);
    (void)cudaGetLastError();
    delete ctx;
}

const char* egpu_last_error(egpu_ctx* ctx) { return ctx ? ctx->last_err : ""; }
int egpu_backend(egpu_ctx* ctx) { return ctx ? 1 : EGPU_ERR_INVALID; }
int64_t egpu_launch_count(egpu_ctx* ctx) {
    if (!ctx) return 0;
    std::lock_guard<std::mutex> g(ctx->mu);
    return ctx->launches;
}

int egpu_set_variant(egpu_ctx* ctx, int variant) {
    if (!ctx || variant < EGPU_VARIANT_AUTO || variant > EGPU_VARIANT_LUT) return EGPU_ERR_INVALID;
    std::lock_guard<std::mutex> g(ctx->mu);
    ctx->variant = variant;
    return EGPU_OK;
}

int egpu_table_set(egpu_ctx* ctx, const int32_t* free_core, const int32_t* free_mem, int32_t D) {
    if (!ctx || !free_core || !free_mem || D < 1 || D > EGPU_MAX_DEVICES) return EGPU_ERR_INVALID;
    for (int d = 0; d < D; ++d) {
        if (free_core[d] < 0 || free_core[d] > EGPU_CORE_MAX) return EGPU_ERR_INVALID;
        if (free_mem[d] < 0 || free_mem[d] > EGPU_MEM_MAX) return EGPU_ERR_INVALID;
    }
    std::lock_guard<std::mutex> g(ctx->mu);
    EGPU_CUDA(ctx, cudaSetDevice(ctx->dev));
    DevState h;
    std::memset(&h, 0, sizeof h);
    std::memcpy(h.free_core, free_core, sizeof(int32_t) * D);
    std::memcpy(h.free_mem, free_mem, sizeof(int32_t) * D);
    h.D = D;
    h.cand_xor = kGuards;
    h.cand_mask = kCandMask;
    fill_sorted(h);
    // pageable source: the copy is staged before the call returns
    // only the table part: the epilogue slots that follow stay as the device left them
    EGPU_CUDA(ctx, cudaMemcpyAsync(ctx->d_state, &h, offsetof(DevState, epi), cudaMemcpyHostToDevice, ctx->stream));
    EGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->D = D;
    ctx->has_table = true;
    ctx->prev_is_scan = false;
    ctx->lut_dirty = true;
    return EGPU_OK;
}

int egpu_table_size(egpu_ctx* ctx) {
    if (!ctx) return EGPU_ERR_INVALID;
    std::lock_guard<std::mutex> g(ctx->mu);
    return ctx->has_table ? ctx->D : EGPU_ERR_NO_TABLE;
}

int egpu_table_get(egpu_ctx* ctx, int32_t* free_core, int32_t* free_mem, int32_t* oversub) {
    if (!ctx || !free_core || !free_mem) return EGPU_ERR_INVALID;
    std::lock_guard<std::mutex> g(ctx->mu);
    if (!ctx->has_table) return EGPU_ERR_NO_TABLE;
    EGPU_CUDA(ctx, cudaSetDevice(ctx->dev));
    EGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    EGPU_CUDA(ctx, cudaMemcpy(ctx->h_table, ctx->d_state, sizeof(int32_t) * 3 * kMaxD, cudaMemcpyDeviceToHost));
    std::memcpy(free_core, ctx->h_table, sizeof(int32_t) * ctx->D);
    std::memcpy(free_mem, ctx->h_table + kMaxD, sizeof(int32_t) * ctx->D);
    if (oversub) std::memcpy(oversub, ctx->h_table + 2 * kMaxD, sizeof(int32_t) * ctx->D);
    return EGPU_OK;
}

int egpu_host_alloc(egpu_ctx* ctx, void** out, int64_t bytes) {
    if (!ctx || !out || bytes <= 0) return EGPU_ERR_INVALID;
    std::lock_guard<std::mutex> g(ctx->mu);
    EGPU_CUDA(ctx, cudaSetDevice(ctx->dev));
    EGPU_CUDA(ctx, cudaMallocHost(out, static_cast<size_t>(bytes)));
    return EGPU_OK;
}

void egpu_host_free(egpu_ctx* ctx, void* p) {
    if (!ctx || !p) return;
    std::lock_guard<std::mutex> g(ctx->mu);
    cudaSetDevice(ctx->dev);
    cudaFreeHost(p);
    (void)cudaGetLastError();
}

int egpu_bestfit_batch_dev(egpu_ctx* ctx, const int32_t* d_req_core, const int32_t* d_req_mem, int64_t R,
                           int32_t* d_out_idx, int64_t* d_delta, int32_t* d_table_out, int flags,
                           void* stream) {
    if (!ctx || R < 0) return EGPU_ERR_INVALID;
    if (R > 0 && (!d_req_core || !d_req_mem || !d_out_idx)) return EGPU_ERR_INVALID;
    if (!aligned16(d_req_core) || !aligned16(d_req_mem) || !aligned16(d_out_idx)) return EGPU_ERR_INVALID;
    std::lock_guard<std::mutex> g(ctx->mu);
    if (!ctx->has_table) return EGPU_ERR_NO_TABLE;
    EGPU_CUDA(ctx, cudaSetDevice(ctx->dev));
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    return launch_snapshot(ctx, d_req_core, d_req_mem, R, d_out_idx, reinterpret_cast<long long*>(d_delta),
                           d_table_out, flags, true, s);
}

int egpu_bestfit_batch(egpu_ctx* ctx, const int32_t* req_core, const int32_t* req_mem, int64_t R,
                       int32_t* out_idx, int64_t* out_delta_core, int64_t* out_delta_mem, int commit) {
    if (!ctx || R < 0) return EGPU_ERR_INVALID;
    if (R > 0 && (!req_core || !req_mem || !out_idx)) return EGPU_ERR_INVALID;
    std::lock_guard<std::mutex> g(ctx->mu);
    if (!ctx->has_table) return EGPU_ERR_NO_TABLE;
    EGPU_CUDA(ctx, cudaSetDevice(ctx->dev));
    cudaStream_t s = ctx->stream;
    const int D = ctx->D;
    int rc;
    // Zero-copy: when all three caller arrays are pinned (egpu_host_alloc or
    // cudaHostRegister) and 16-byte aligned, the scan reads the requests and writes
    // the indices straight across PCIe — one launch, no staging in HBM, reads and
    // writes overlap on the full-duplex link.  Demand sums land in pinned memory too.
    const int32_t* zc = R > 0 ? static_cast<const int32_t*>(mapped_alias(req_core)) : nullptr;
    const int32_t* zm = R > 0 ? static_cast<const int32_t*>(mapped_alias(req_mem)) : nullptr;
    int32_t* zi = R > 0 ? static_cast<int32_t*>(mapped_alias(out_idx)) : nullptr;
    if (zc && zm && zi && aligned16(zc) && aligned16(zm) && aligned16(zi) && !ctx->no_zero_copy) {
        // 64 rows per thread: few CTAs, many trips, so reads of later rows and writes of
        // earlier ones are on the link at the same time (PCIe is full duplex)
        rc = launch_snapshot(ctx, zc, zm, R, zi, ctx->h_delta_dev, nullptr, commit ? EGPU_F_COMMIT : 0, true, s, 64);
        if (rc != EGPU_OK) return rc;
        EGPU_CUDA(ctx, cudaStreamSynchronize(s));
    } else {
        rc = ensure_staging(ctx, R > 0 ? R : 1);
        if (rc != EGPU_OK) return rc;
        if (R > 0) {
            EGPU_CUDA(ctx, cudaMemcpyAsync(ctx->d_req_core, req_core, sizeof(int32_t) * R, cudaMemcpyHostToDevice, s));
            EGPU_CUDA(ctx, cudaMemcpyAsync(ctx->d_req_mem, req_mem, sizeof(int32_t) * R, cudaMemcpyHostToDevice, s));
        }
        rc = launch_snapshot(ctx, ctx->d_req_core, ctx->d_req_mem, R, ctx->d_idx, ctx->d_delta, nullptr,
                             commit ? EGPU_F_COMMIT : 0, true, s);
        if (rc != EGPU_OK) return rc;
        if (R > 0) EGPU_CUDA(ctx, cudaMemcpyAsync(out_idx, ctx->d_idx, sizeof(int32_t) * R, cudaMemcpyDeviceToHost, s));
        EGPU_CUDA(ctx, cudaMemcpyAsync(ctx->h_delta, ctx->d_delta, sizeof(long long) * 2 * D, cudaMemcpyDeviceToHost, s));
        EGPU_CUDA(ctx, cudaStreamSynchronize(s));
    }
    if (out_delta_core) std::memcpy(out_delta_core, ctx->h_delta, sizeof(int64_t) * D);
    if (out_delta_mem) std::memcpy(out_delta_mem, ctx->h_delta + D, sizeof(int64_t) * D);
    return EGPU_OK;
}

int egpu_peer_export(egpu_ctx* ctx, void* handle_out) {
    if (!ctx || !handle_out) return EGPU_ERR_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == EGPU_IPC_HANDLE_BYTES, "IPC handle size");
    std::lock_guard<std::mutex> g(ctx->mu);
    EGPU_CUDA(ctx, cudaSetDevice(ctx->dev));
    cudaIpcMemHandle_t h;
    EGPU_CUDA(ctx, cudaIpcGetMemHandle(&h, ctx->d_xchg));
    std::memcpy(handle_out, &h, sizeof h);
    return EGPU_OK;
}

int egpu_peer_attach(egpu_ctx* ctx, int rank, int world, const void* handles) {
    if (!ctx || !handles || world < 1 || world > kMaxRanks || rank < 0 || rank >= world) return EGPU_ERR_INVALID;
    std::lock_guard<std::mutex> g(ctx->mu);
    if (ctx->attached) return EGPU_ERR_STATE;
    EGPU_CUDA(ctx, cudaSetDevice(ctx->dev));
    PeerCfg cfg;
    std::memset(&cfg, 0, sizeof cfg);
    cfg.world = world;
    cfg.rank = rank;
    for (int r = 0; r < world; ++r) {
        if (r == rank) {
            cfg.buf[r] = ctx->d_xchg;
            continue;
        }
        cudaIpcMemHandle_t h;
        std::memcpy(&h, static_cast<const char*>(handles) + static_cast<size_t>(r) * sizeof h, sizeof h);
        void* p = nullptr;
        EGPU_CUDA(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        ctx->peer_open[r] = p;
        cfg.buf[r] = static_cast<XchgBuf*>(p);
    }
    EGPU_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    EGPU_CUDA(ctx, cudaMemcpy(reinterpret_cast<char*>(ctx->d_state) + offsetof(DevState, peer), &cfg, sizeof cfg,
                              cudaMemcpyHostToDevice));
    ctx->world = world;
    ctx->rank = rank;
    ctx->attached = true;
    return EGPU_OK;
}

int egpu_peer_detach(egpu_ctx* ctx) {
    if (!ctx) return EGPU_ERR_INVALID;
    std::lock_guard<std::mutex> g(ctx->mu);
    if (!ctx->attached) return EGPU_OK;
    EGPU_CUDA(ctx, cudaSetDevice(ctx->dev));
    EGPU_CUDA(ctx, cudaDeviceSynchronize());
    for (int r = 0; r < kMaxRanks; ++r) {
        if (ctx->peer_open[r]) cudaIpcCloseMemHandle(ctx->peer_open[r]);
        ctx->peer_open[r] = nullptr;
    }
    ctx->attached = false;
    ctx->world = 1;
    ctx->rank = 0;
    return EGPU_OK;
}

int egpu_bestfit_batch_shard_dev(egpu_ctx* ctx, const int32_t* d_req_core, const int32_t* d_req_mem, int64_t R,
                                 int32_t* d_out_idx, int64_t* d_delta, int flags, uint64_t step, void* stream) {
    if (!ctx || R < 0 || (flags & EGPU_F_COMMIT)) return EGPU_ERR_INVALID;  // the commit happens in apply_peers
    if (R > 0 && (!d_req_core || !d_req_mem || !d_out_idx)) return EGPU_ERR_INVALID;
    if (!aligned16(d_req_core) || !aligned16(d_req_mem) || !aligned16(d_out_idx)) return EGPU_ERR_INVALID;
    std::lock_guard<std::mutex> g(ctx->mu);
    if (!ctx->has_table) return EGPU_ERR_NO_TABLE;
    if (!ctx->attached) return EGPU_ERR_STATE;
    EGPU_CUDA(ctx, cudaSetDevice(ctx->dev));
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    return launch_snapshot(ctx, d_req_core, d_req_mem, R, d_out_idx, reinterpret_cast<long long*>(d_delta), nullptr,
                           flags, true, s, 0, step + 1);
}

int egpu_table_apply_peers_multi_dev(egpu_ctx* ctx, uint64_t first_step, int nsteps, int32_t* const* d_table_outs,
                                     int commit, void* stream) {
    if (!ctx || nsteps < 1 || nsteps > 8) return EGPU_ERR_INVALID;
    std::lock_guard<std::mutex> g(ctx->mu);
    if (!ctx->has_table) return EGPU_ERR_NO_TABLE;
    if (!ctx->attached) return EGPU_ERR_STATE;
    EGPU_CUDA(ctx, cudaSetDevice(ctx->dev));
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    if (commit) {
        ctx->lut_dirty = true;
        ctx->prev_is_scan = false;  // the next scan must see the new table
    }
    ApplyOuts outs;
    for (int k = 0; k < 8; ++k) outs.table_out[k] = (d_table_outs && k < nsteps) ? d_table_outs[k] : nullptr;
    apply_peers_kernel<<<1, kMaxD, 0, s>>>(ctx->d_state, first_step + 1, nsteps, outs, commit);
    EGPU_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return EGPU_OK;
}

int egpu_table_apply_peers_dev(egpu_ctx* ctx, uint64_t step, int32_t* d_table_out, int commit, void* stream) {
    int32_t* outs[1] = {d_table_out};
    return egpu_table_apply_peers_multi_dev(ctx, step, 1, outs, commit, stream);
}

int64_t egpu_peer_last_timeout(egpu_ctx* ctx) {
    if (!ctx) return EGPU_ERR_INVALID;
    std::lock_guard<std::mutex> g(ctx->mu);
    if (cudaSetDevice(ctx->dev) != cudaSuccess) return EGPU_ERR_CUDA;
    unsigned long long v = 0;
    if (cudaMemcpy(&v, reinterpret_cast<char*>(ctx->d_state) + offsetof(DevState, peer_timeout), sizeof v,
                   cudaMemcpyDeviceToHost) != cudaSuccess) {
        (void)cudaGetLastError();
        return EGPU_ERR_CUDA;
    }
    return static_cast<int64_t>(v);
}

int egpu_bestfit_batch_packed_dev(egpu_ctx* ctx, const uint32_t* d_req_packed, int64_t R, int8_t* d_out_idx8,
                                  int64_t* d_delta, int32_t* d_table_out, int flags, void* stream) {
    if (!ctx || R < 0) return EGPU_ERR_INVALID;
    if (R > 0 && (!d_req_packed || !d_out_idx8)) return EGPU_ERR_INVALID;
    if (!aligned16(d_req_packed) || !aligned16(d_out_idx8)) return EGPU_ERR_INVALID;
    std::lock_guard<std::mutex> g(ctx->mu);
    if (!ctx->has_table) return EGPU_ERR_NO_TABLE;
    EGPU_CUDA(ctx, cudaSetDevice(ctx->dev));
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    return launch_packed(ctx, d_req_packed, R, reinterpret_cast<signed char*>(d_out_idx8), reinterpret_cast<long long*>(d_delta),
                         d_table_out, flags, s, 16);
}

int egpu_bestfit_batch_packed(egpu_ctx* ctx, const uint32_t* req_packed, int64_t R, int8_t* out_idx8,
                              int64_t* out_delta_core, int64_t* out_delta_mem, int commit) {
    if (!ctx || R < 0) return EGPU_ERR_INVALID;
    if (R > 0 && (!req_packed || !out_idx8)) return EGPU_ERR_INVALID;
    std::lock_guard<std::mutex> g(ctx->mu);
    if (!ctx->has_table) return EGPU_ERR_NO_TABLE;
    EGPU_CUDA(ctx, cudaSetDevice(ctx->dev));
    cudaStream_t s = ctx->stream;
    const int D = ctx->D;
    int rc;
    const uint32_t* zr = R > 0 ? static_cast<const uint32_t*>(mapped_alias(req_packed)) : nullptr;
    signed char* zi = R > 0 ? static_cast<signed char*>(mapped_alias(out_idx8)) : nullptr;
    if (zr && zi && aligned16(zr) && aligned16(zi) && !ctx->no_zero_copy) {  // zero-copy across PCIe, see egpu_bestfit_batch
        rc = launch_packed(ctx, zr, R, zi, ctx->h_delta_dev, nullptr, commit ? EGPU_F_COMMIT : 0, s, 128);
        if (rc != EGPU_OK) return rc;
        EGPU_CUDA(ctx, cudaStreamSynchronize(s));
    } else {
        rc = ensure_staging(ctx, R > 0 ? R : 1);  // d_req_core holds the packed words, d_idx the bytes
        if (rc != EGPU_OK) return rc;
        if (R > 0) EGPU_CUDA(ctx, cudaMemcpyAsync(ctx->d_req_core, req_packed, sizeof(uint32_t) * R, cudaMemcpyHostToDevice, s));
        rc = launch_packed(ctx, reinterpret_cast<const uint32_t*>(ctx->d_req_core), R, reinterpret_cast<signed char*>(ctx->d_idx),
                           ctx->d_delta, nullptr, commit ? EGPU_F_COMMIT : 0, s, 16);
        if (rc != EGPU_OK) return rc;
        if (R > 0) EGPU_CUDA(ctx, cudaMemcpyAsync(out_idx8, ctx->d_idx, static_cast<size_t>(R), cudaMemcpyDeviceToHost, s));
        EGPU_CUDA(ctx, cudaMemcpyAsync(ctx->h_delta, ctx->d_delta, sizeof(long long) * 2 * D, cudaMemcpyDeviceToHost, s));
        EGPU_CUDA(ctx, cudaStreamSynchronize(s));
    }
    if (out_delta_core) std::memcpy(out_delta_core, ctx->h_delta, sizeof(int64_t) * D);
    if (out_delta_mem) std::memcpy(out_delta_mem, ctx->h_delta + D, sizeof(int64_t) * D);
    return EGPU_OK;
}

int egpu_table_apply_deltas_dev(egpu_ctx* ctx, const int64_t* d_deltas, int G, int32_t* d_table_out,
                                int commit, void* stream) {
    if (!ctx || !d_deltas || G < 1) return EGPU_ERR_INVALID;
    std::lock_guard<std::mutex> g(ctx->mu);
    if (!ctx->has_table) return EGPU_ERR_NO_TABLE;
    EGPU_CUDA(ctx, cudaSetDevice(ctx->dev));
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    ctx->prev_is_scan = false;
    if (commit) ctx->lut_dirty = true;
    apply_deltas_kernel<<<1, kMaxD, 0, s>>>(ctx->d_state, reinterpret_cast<const long long*>(d_deltas), G,
                                            d_table_out, commit);
    EGPU_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return EGPU_OK;
}

int egpu_synth_requests_dev(egpu_ctx* ctx, int dist, uint64_t seed, int64_t first_row, int64_t R,
                            int32_t* d_req_core, int32_t* d_req_mem, void* stream) {
    if (!ctx || R < 0 || first_row < 0 || (dist != 2 && dist != 3 && dist != 4)) return EGPU_ERR_INVALID;
    if (R == 0) return EGPU_OK;
    if (!d_req_core || !d_req_mem) return EGPU_ERR_INVALID;
    std::lock_guard<std::mutex> g(ctx->mu);
    EGPU_CUDA(ctx, cudaSetDevice(ctx->dev));
    cudaStream_t s = stream ? static_cast<cudaStream_t>(stream) : ctx->stream;
    int64_t blocks = (R + 255) / 256;
    const int64_t cap = static_cast<int64_t>(ctx->sm_count) * 8;
    if (blocks > cap) blocks = cap;
    ctx->prev_is_scan = false;
    synth_requests_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(dist, seed, first_row, R, d_req_core, d_req_mem);
    EGPU_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return EGPU_OK;
}

int egpu_replay(egpu_ctx* ctx, const int32_t* kind, const int32_t* a, const int32_t* b, int64_t E,
                int32_t* out_idx) {
    if (!ctx || E < 0 || E > 0x7fffffffll) return EGPU_ERR_INVALID;
    if (E > 0 && (!kind || !a || !b || !out_idx)) return EGPU_ERR_INVALID;
    std::lock_guard<std::mutex> g(ctx->mu);
    if (!ctx->has_table) return EGPU_ERR_NO_TABLE;
    if (E == 0) return EGPU_OK;
    EGPU_CUDA(ctx, cudaSetDevice(ctx->dev));
    // staging comes from the context's grow-only arena (three inputs, one output, the `live`
    // map when it does not fit in shared memory): no cudaMalloc/cudaFree on the call path
    cudaStream_t s = ctx->stream;
    const size_t ebytes = (sizeof(int32_t) * static_cast<size_t>(E) + 255) & ~static_cast<size_t>(255);
    const size_t lbytes = E > kReplaySmemEvents ? ((static_cast<size_t>(E) + 255) & ~static_cast<size_t>(255)) : 0;
    const size_t need = 4 * ebytes + lbytes;
    if (need > ctx->arena_cap) {
        if (ctx->arena) cudaFree(ctx->arena);
        ctx->arena = nullptr;
        ctx->arena_cap = 0;
        EGPU_CUDA(ctx, cudaMalloc(&ctx->arena, need + need / 4));
        ctx->arena_cap = need + need / 4;
    }
    char* base = static_cast<char*>(ctx->arena);
    int32_t* d_kind = reinterpret_cast<int32_t*>(base);
    int32_t* d_a = reinterpret_cast<int32_t*>(base + ebytes);
    int32_t* d_b = reinterpret_cast<int32_t*>(base + 2 * ebytes);
    int32_t* d_out = reinterpret_cast<int32_t*>(base + 3 * ebytes);
    signed char* d_live = lbytes ? reinterpret_cast<signed char*>(base + 4 * ebytes) : nullptr;
    size_t smem = 0;
    if (E <= kReplaySmemEvents) {
        smem = static_cast<size_t>((E + 15) & ~15ll);
        if (!ctx->replay_configured) {
            EGPU_CUDA(ctx, cudaFuncSetAttribute(replay_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kReplaySmemEvents));
            EGPU_CUDA(ctx, cudaFuncSetAttribute(replay8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kReplaySmemEvents));
            ctx->replay_configured = true;
        }
    }
    EGPU_CUDA(ctx, cudaMemcpyAsync(d_kind, kind, sizeof(int32_t) * E, cudaMemcpyHostToDevice, s));
    EGPU_CUDA(ctx, cudaMemcpyAsync(d_a, a, sizeof(int32_t) * E, cudaMemcpyHostToDevice, s));
    EGPU_CUDA(ctx, cudaMemcpyAsync(d_b, b, sizeof(int32_t) * E, cudaMemcpyHostToDevice, s));
    ctx->prev_is_scan = false;
    ctx->lut_dirty = true;
    if (ctx->D <= 8 && !ctx->replay_general)
        replay8_kernel<<<1, 32, smem, s>>>(ctx->d_state, d_kind, d_a, d_b, E, d_out, d_live);
    else
        replay_kernel<<<1, 32, smem, s>>>(ctx->d_state, d_kind, d_a, d_b, E, d_out, d_live);
    EGPU_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    EGPU_CUDA(ctx, cudaMemcpyAsync(out_idx, d_out, sizeof(int32_t) * E, cudaMemcpyDeviceToHost, s));
    EGPU_CUDA(ctx, cudaStreamSynchronize(s));
    return EGPU_OK;
}

}  // extern "C"
