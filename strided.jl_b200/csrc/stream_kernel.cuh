// stream_kernel.cuh -- sm_100a kernel: streamed complete reduction (common.hpp "StreamParams", reduce_stream.hpp).
//
//   producer (one elected thread of warp 8): per chunk, waits for the stage's EMPTY barrier, arms the FULL barrier with
//            the chunk's byte count and issues one cp.async.bulk (1-D bulk copy global -> shared, mbarrier complete_tx)
//            per input: `nstage` chunks per CTA are in flight from the first cycle on            (SASS: UBLKCP, SYNCS)
//   consumers (8 warps): wait on the FULL barrier, fold the chunk from shared memory with 128-bit loads into four private
//            accumulators per thread, release the stage (one arrive per warp)
//   epilogue: warp butterfly + fixed-order fold of the 8 warp results -> CTA partial -> partials[blockIdx];
//            ONE atom.add.acq_rel.gpu per CTA on the arrival counter (no __threadfence by 256 threads, no extra
//            barriers); the last-arriving CTA folds the CTA partials in CTA order, exchanges the value across GPUs when
//            the call is collective (peer_ll_allreduce) and stores op(initop(out), total).
// Measured on the per-GPU share of BASELINE config 5 (4096 x 4096 Float64 -> scalar): see DESIGN.md section 4.
#pragma once
#include "tma_kernel.cuh"
#include "reduce_stream.hpp"

namespace sb {

__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

constexpr int STREAM_MAXSTAGE = 8;
constexpr int STREAM_THREADS = THREADS + 32;

template <class AT, int RC, int NIN>
__global__ void __launch_bounds__(STREAM_THREADS, 1) reduce_stream_kernel(const __grid_constant__ StreamArgs P)
{
    const StreamParams &S = P.S;
    extern __shared__ unsigned char sb_stream_smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STREAM_MAXSTAGE];
    __shared__ __align__(8) uint64_t empty_bar[STREAM_MAXSTAGE];
    __shared__ __align__(16) unsigned char fold_raw[(THREADS / 32) * STREAM_MAXOUT * sizeof(AT)]; // [warp][output] (dense mode uses 2 x 8 entries)
    __shared__ unsigned int is_last, s_epoch;
    AT *fold = reinterpret_cast<AT *>(fold_raw);
    unsigned char *ring = sb_stream_smem_raw + ((0u - smem_u32(sb_stream_smem_raw)) & 127u);
    const uint32_t ring_u32 = smem_u32(ring);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NS = S.nstage;
    pdl_launch_dependents();
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1);
            mbar_init(smem_u32(&empty_bar[s]), THREADS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
#ifdef SB_STREAM_CONSUMER_NOWAIT // (experiment, tools/ubench_stream_product.cu) only the producer waits for the previous grid
    if (warp == THREADS / 32) pdl_wait();
#else
    pdl_wait(); // operands, output, partials and the arrival counter may all be in use by the previous kernel
#endif
    const int64_t nchunks = S.nchunks;
    const uint32_t grid = gridDim.x;
    const int nout = S.nout;
    if (warp == THREADS / 32) {
        // ---------------- producer ----------------
        if (lane == 0) {
            int stage = 0;
            uint32_t parity = 1; // a fresh barrier passes a wait on parity 1: every stage starts out empty
            const int nruns_p = S.inter_g > 0 ? 1 : nout; // interleaved mode: ONE run carries all outputs
            for (int o = 0; o < nruns_p; ++o) {
                const unsigned char *src[NIN];
#pragma unroll
                for (int k = 0; k < NIN; ++k) src[k] = P.base[(k < S.nin ? k : 0) + 1] + stream_out_offset(S, o, k < S.nin ? k : 0);
                for (int64_t c = blockIdx.x; c < nchunks; c += grid) {
                    mbar_wait(smem_u32(&empty_bar[stage]), parity);
                    const int64_t off = c * (int64_t)S.chunk_bytes;
                    const int64_t left = S.vec_bytes - off;
                    const uint32_t nb = (uint32_t)(left < (int64_t)S.chunk_bytes ? left : (int64_t)S.chunk_bytes);
                    const uint32_t fb = smem_u32(&full_bar[stage]);
                    mbar_expect_tx(fb, nb * (uint32_t)S.nin);
                    const uint32_t dst = ring_u32 + (uint32_t)(stage * S.stage_bytes);
#pragma unroll
                    for (int k = 0; k < NIN; ++k)
                        if (k < S.nin) bulk_load_1d(dst + (uint32_t)(k * S.chunk_bytes), src[k] + off, nb, fb);
                    if (++stage == NS) {
                        stage = 0;
                        parity ^= 1u;
                    }
                }
            }
        }
        return;
    }
    // ---------------- consumers ----------------
    constexpr int V = StreamVec<AT>::V;
    int stage = 0;
    uint32_t parity = 0;
    const int nruns = S.inter_g > 0 ? 1 : nout; // interleaved mode: ONE run carries all outputs
    for (int o = 0; o < nruns; ++o) {
        AT acc[STREAM_ACC][V];
#pragma unroll
        for (int q = 0; q < STREAM_ACC; ++q)
#pragma unroll
            for (int u = 0; u < V; ++u) acc[q][u] = red_neutral<AT>(P.op);
        for (int64_t c = blockIdx.x; c < nchunks; c += grid) {
            const int64_t left = S.vec_bytes - c * (int64_t)S.chunk_bytes;
            const int nv = (int)((left < (int64_t)S.chunk_bytes ? left : (int64_t)S.chunk_bytes) >> 4);
            mbar_wait(smem_u32(&full_bar[stage]), parity);
            stream_chunk<AT, RC, NIN>(P, S, ring + (size_t)stage * S.stage_bytes, nv, tid, acc);
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&empty_bar[stage]));
            if (++stage == NS) {
                stage = 0;
                parity ^= 1u;
            }
        }
        if (S.inter_g > 0) {
            // interleaved outputs: lane u of thread t carries output (t mod G) * V + u.  Warp butterfly over the lanes of
            // equal class (xor masks >= G), one row of [warp][output] in shared memory, then thread o folds the 8 warps.
            const int G = S.inter_g;
#pragma unroll
            for (int u = 0; u < V; ++u) {
                AT pu = stream_thread_lane_total<AT>(P, acc, u);
                for (int m = 16; m >= G; m >>= 1) pu = red_apply<AT>(P.op, pu, shfl_xor_any(pu, m));
                if (lane < G) fold[warp * STREAM_MAXOUT + lane * V + u] = pu;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory");
            if (tid < nout) {
                AT q = fold[tid];
#pragma unroll
                for (int w = 1; w < THREADS / 32; ++w) q = red_apply<AT>(P.op, q, fold[w * STREAM_MAXOUT + tid]);
                reinterpret_cast<AT *>(P.scratch)[(size_t)tid * grid + blockIdx.x] = q; // this CTA's partial of output tid
            }
            // (the partials were written by threads 0..nout-1; the barrier orders them before thread 0's acq_rel arrival
            //  below, whose release is cumulative over everything that happens-before it)
            asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory");
            break;
        }
        AT p = stream_thread_total<AT>(P, acc);
        if (blockIdx.x == 0 && tid == 0) p = stream_rest<AT, RC, NIN>(P, S, o, p);
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) p = red_apply<AT>(P.op, p, shfl_xor_any(p, m));
        AT *fb = fold + (o & 1) * (THREADS / 32); // two buffers: thread 0 reads output o's while the others fill o + 1's
        if (lane == 0) fb[warp] = p;
        asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory"); // consumers only: the producer warp has retired
        if (tid == 0) {
            AT q = fb[0];
#pragma unroll
            for (int w = 1; w < THREADS / 32; ++w) q = red_apply<AT>(P.op, q, fb[w]);
            reinterpret_cast<AT *>(P.scratch)[(size_t)o * grid + blockIdx.x] = q; // this CTA's partial of output o
        }
    }
    if (tid == 0) {
        unsigned int old = grid - 1u;
        if (grid > 1) asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(P.counters) : "memory");
        is_last = (old == grid - 1u) ? 1u : 0u;
        if (is_last) {
            if (grid > 1) *P.counters = 0u; // re-arm for the next launch (stream-ordered: nobody else touches it before)
            uint32_t epoch = 0u;
            if (P.peer.world > 1) { // collective call: the call number lives on the device (graph-replay safe)
                epoch = __ldcg(P.peer.epoch_ptr) + 1u;
                __stcg(P.peer.epoch_ptr, epoch);
            }
            s_epoch = epoch;
        }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory");
    if (!is_last) return;
    // the last-arriving CTA: warp w folds the CTA partials of outputs w, w + 8, ... in CTA order (lane l takes l, l + 32, ...)
    const AT *sc = reinterpret_cast<const AT *>(P.scratch);
    for (int o = warp; o < nout; o += THREADS / 32) {
        AT r = red_neutral<AT>(P.op);
        for (uint32_t i0 = (uint32_t)lane; i0 < grid; i0 += 256) { // eight independent L2 loads in flight per lane, folded in CTA order
            AT v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = (i0 + 32u * u < grid) ? load_partial(sc + (size_t)o * grid + i0 + 32u * u) : red_neutral<AT>(P.op);
#pragma unroll
            for (int u = 0; u < 8; ++u) r = red_apply<AT>(P.op, r, v[u]);
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) r = red_apply<AT>(P.op, r, shfl_xor_any(r, m));
        if (lane == 0) {
            if constexpr (sizeof(AT) <= 8) {
                if (P.peer.world > 1) r = peer_ll_exchange<AT, StreamArgs>(P, o, r, s_epoch); // one value per rank crosses NVLink, folded in rank order
            }
            stream_store<AT>(P, S, o, r);
        }
    }
}

struct StreamEntry {
    KernelKey key; // (ct, recipe, nin); ept and uniform unused
    cudaError_t (*launch)(const StreamArgs &, int grid, size_t smem, cudaStream_t);
    const void *func;
};

template <class AT, int RC, int NIN> struct StreamLaunch {
    static cudaError_t launch(const StreamArgs &P, int grid, size_t smem, cudaStream_t s)
    {
        auto k = reduce_stream_kernel<AT, RC, NIN>;
        cudaError_t e = ensure_dynamic_smem((const void *)k, smem);
        if (e != cudaSuccess) return e;
        return launch_pdl(k, grid, STREAM_THREADS, smem, s, P);
    }
    static const void *func() { return (const void *)reduce_stream_kernel<AT, RC, NIN>; }
};

#define SB_STREAM_ENTRY(CT, DT, RC, NIN)                                                                             \
    StreamEntry { KernelKey{DT, RC, NIN, 0, 1}, &StreamLaunch<CT, RC, NIN>::launch, StreamLaunch<CT, RC, NIN>::func() }

const StreamEntry *stream_table(int *n);
const StreamEntry *find_stream_kernel(const KernelKey &k);

} // namespace sb
