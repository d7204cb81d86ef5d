// tma_kernel.cuh -- sm_100a kernel: TMA-fed, mbarrier-pipelined strided map (the permutedims / transpose /
// `(A .+ A') ./ 2` fast path).  SASS evidence: UTMALDG (cp.async.bulk.tensor) + SYNCS (mbarrier).
//
// Pipeline (warp-specialised, see map_tma_kernel below): a producer warp keeps `nstage` tiles in flight -- per tile it
//   waits for the stage's EMPTY barrier, arms the FULL barrier with the stage's byte count and lets one lane per
//   (operand, box) issue the cp.async.bulk.tensor; eight consumer warps wait on the FULL barrier, read their EPT elements
//   per operand from shared memory in OUTPUT order (XOR-swizzled addresses), evaluate f, store with streaming stores and
//   release the stage (one arrive per warp).  No CTA-wide barrier in the loop.  Every kernel executes
//   griddepcontrol.launch_dependents / griddepcontrol.wait (PDL, common.hpp).
#pragma once
#include "kernels.cuh"
#include "tma_tile.hpp"
#include <cuda.h>

namespace sb {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) __trap(); // a lost TMA transaction must not hang the GPU
    }
}

template <int RANK> __device__ __forceinline__ void tma_load(uint32_t dst, const CUtensorMap *map, uint32_t bar, const int32_t *c);
template <> __device__ __forceinline__ void tma_load<1>(uint32_t dst, const CUtensorMap *map, uint32_t bar, const int32_t *c)
{
    asm volatile("cp.async.bulk.tensor.1d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3}], [%2];" ::"r"(dst),
                 "l"(map), "r"(bar), "r"(c[0])
                 : "memory");
}
template <> __device__ __forceinline__ void tma_load<2>(uint32_t dst, const CUtensorMap *map, uint32_t bar, const int32_t *c)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
                 "l"(map), "r"(bar), "r"(c[0]), "r"(c[1])
                 : "memory");
}
template <> __device__ __forceinline__ void tma_load<3>(uint32_t dst, const CUtensorMap *map, uint32_t bar, const int32_t *c)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
                 "l"(map), "r"(bar), "r"(c[0]), "r"(c[1]), "r"(c[2])
                 : "memory");
}
template <> __device__ __forceinline__ void tma_load<4>(uint32_t dst, const CUtensorMap *map, uint32_t bar, const int32_t *c)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
                 "l"(map), "r"(bar), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3])
                 : "memory");
}
template <> __device__ __forceinline__ void tma_load<5>(uint32_t dst, const CUtensorMap *map, uint32_t bar, const int32_t *c)
{
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
                 "l"(map), "r"(bar), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4])
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// one lane of the producer warp = one box of one operand: decode the tile, build the box coordinates, issue
__device__ __forceinline__ void tma_issue_box(const MapParams &P, const TmaOperand &o, const CUtensorMap *map, int q, uint32_t id,
                                              uint32_t dst, uint32_t bar)
{
    int32_t crd[TMA_MAXRANK] = {0, 0, 0, 0, 0};
#pragma unroll
    for (int d = 0; d < TMA_MAXRANK; ++d) { // canonical ndim <= TMA_MAXRANK on this path
        if (d < P.ndim) {
            uint32_t qq, c;
            fast_divmod(P.tdiv[d], id, qq, c);
            id = qq;
            const int32_t origin = (int32_t)map_tile_origin(P, d, c);
#pragma unroll
            for (int i = 0; i < TMA_MAXRANK; ++i)
                if (o.cdim[i] == d && i < o.rank) crd[i] = origin;
        }
    }
    crd[0] += q * o.inner_step;
    switch (o.rank) {
    case 1: tma_load<1>(dst, map, bar, crd); break;
    case 2: tma_load<2>(dst, map, bar, crd); break;
    case 3: tma_load<3>(dst, map, bar, crd); break;
    case 4: tma_load<4>(dst, map, bar, crd); break;
    default: tma_load<5>(dst, map, bar, crd); break;
    }
}

constexpr int TMA_MAXSTAGE = 8;
constexpr int TMA_THREADS = THREADS + 32;
#ifndef SB_TMA_MINB
#define SB_TMA_MINB 3
#endif // 8 consumer warps + 1 producer warp

// Warp-specialised pipeline:
//   producer warp : for every tile of this CTA, wait until the stage is free (empty barrier), arm the full
//                   barrier with the stage's byte count, then each lane issues the cp.async.bulk.tensor of "its"
//                   (operand, box);
//   consumer warps: wait on the full barrier, read the tile from shared memory in OUTPUT order, evaluate f, store,
//                   then release the stage (one arrive per warp on the empty barrier).
// No CTA-wide barrier in the loop: the producer runs up to `nstage` tiles ahead of the slowest consumer warp.
//
// GROUP: several problems of ONE plan (same dims, strides, eltypes and program; different base pointers) in one launch --
// sb_mapreduce_batch's answer to launch- and latency-bound sizes (1000^2, 32^4: one wave of tiles each).  Global position
// g = problem * ntiles + tile; the problem selects the tensor maps and the output base, everything else is the plan's.
constexpr int TMA_GROUP_MAX = 16; // problems per launch
constexpr int TMA_GROUP_MAXIN = 2;
struct TmaGroup {
    int32_t nprob;
    int32_t pad_[15];
    unsigned char *out[TMA_GROUP_MAX];                              // output base of problem p (view offset included)
    alignas(64) CUtensorMap maps[TMA_GROUP_MAX][TMA_GROUP_MAXIN];   // tensor maps of problem p's inputs
};

template <class CT, int RC, int NIN, int EPT, bool GROUP>
__device__ __forceinline__ void map_tma_body(const MapParams &P, const TmaParams &T, const CUtensorMap *m0, const CUtensorMap *m1,
                                             const CUtensorMap *m2, const CUtensorMap *m3, const TmaGroup *G)
{
    extern __shared__ unsigned char sb_tma_smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[TMA_MAXSTAGE];
    __shared__ __align__(8) uint64_t empty_bar[TMA_MAXSTAGE];
    // stage ring, 1024-byte aligned (the 128-byte swizzle pattern is a function of address bits [4,10))
    unsigned char *ring = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(sb_tma_smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t ring_u32 = smem_u32(ring);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int S = T.nstage;
    pdl_launch_dependents();
    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1);
            mbar_init(smem_u32(&empty_bar[s]), THREADS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t ntiles1 = (uint32_t)P.ntiles;                                   // tiles of one problem
    const uint32_t ntiles = GROUP ? ntiles1 * (uint32_t)G->nprob : ntiles1;        // positions of this launch
    const uint32_t grid = gridDim.x;
    if (warp == THREADS / 32) {
        // ---------------- producer warp ----------------
        int bk = -1, bq = 0; // this lane's (operand, box)
        {
            int c = 0;
            for (int k = 0; k < T.nin; ++k)
                for (int q = 0; q < T.op[k].nbox; ++q, ++c)
                    if (c == lane) {
                        bk = k;
                        bq = q;
                    }
        }
        const CUtensorMap *map = bk == 0 ? m0 : bk == 1 ? m1 : bk == 2 ? m2 : m3;
        const uint32_t box_dst = bk >= 0 ? (uint32_t)T.op[bk].smem_off + (uint32_t)(bq * T.op[bk].box_bytes) : 0u;
        int my_cdim[TMA_MAXRANK] = {0, 0, 0, 0, 0};
        int my_rank = 1, my_inner_step = 0;
        if (bk >= 0) {
#pragma unroll
            for (int i = 0; i < TMA_MAXRANK; ++i) my_cdim[i] = (i < T.op[bk].rank) ? (int)T.op[bk].cdim[i] : 0;
            my_rank = T.op[bk].rank;
            my_inner_step = T.op[bk].inner_step;
        }
        int stage = 0;
        uint32_t parity = 1; // a fresh barrier passes a wait on parity 1: every stage starts out empty
        // The tile records were written once, at plan creation: they are fetched BEFORE griddepcontrol.wait (and one tile
        // ahead afterwards), so that no dependent global load sits between the wait and the first TMA issue -- for the
        // one-wave configs (1000^2, 32^4) that round trip was ~0.6 us of a ~4 us launch.
        // The tile records were written once, at plan creation: the FIRST one is fetched before griddepcontrol.wait, so that
        // no dependent global load sits between the wait and the first TMA issue -- for the one-wave configs (1000^2, 32^4)
        // that round trip was ~0.6 us of a ~4 us launch.  Later records are loaded where they are used, AFTER the stage
        // has been freed: fetching them earlier (one tile ahead, or before the empty-barrier wait) makes the producer issue
        // sooner and costs config 2 10 % (47.6 vs 42.7 us, profiles/r02_tma_prefetch_bisect.txt) -- the alias-aware tile
        // order relies on the A and A' tiles of neighbouring CTAs meeting in L2, and that pacing is part of it.
        TileDesc td_first = {};
        if (P.tile_desc && blockIdx.x < ntiles) td_first = P.tile_desc[GROUP ? blockIdx.x % ntiles1 : blockIdx.x];
        pdl_wait(); // the operands may be the previous kernel's output
        for (uint32_t gpos = blockIdx.x; gpos < ntiles; gpos += grid) {
            uint32_t pos = gpos;
            if (GROUP) {
                const uint32_t prob = gpos / ntiles1;
                pos = gpos - prob * ntiles1;
                if (bk >= 0) map = &G->maps[prob][bk < TMA_GROUP_MAXIN ? bk : 0];
            }
            mbar_wait(smem_u32(&empty_bar[stage]), parity);
            const uint32_t fb = smem_u32(&full_bar[stage]);
            if (lane == 0) mbar_expect_tx(fb, (uint32_t)T.stage_bytes);
            __syncwarp();
            if (P.uniform & 0x200) { // diagnostic (SB_DEBUG=noload): complete the barrier without moving data
                if (lane == 0) asm volatile("mbarrier.complete_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(fb), "r"((uint32_t)T.stage_bytes) : "memory");
            } else if (bk >= 0) {
                const uint32_t dst = ring_u32 + (uint32_t)(stage * T.stage_bytes) + box_dst;
                if (P.tile_desc) { // precomputed tile record: coordinates are a per-lane permutation of the origins
                    TileDesc td = td_first;
                    if (gpos != blockIdx.x) td = P.tile_desc[pos];
                    int32_t crd[TMA_MAXRANK];
#pragma unroll
                    for (int i = 0; i < TMA_MAXRANK; ++i) {
                        const int cd = my_cdim[i];
                        crd[i] = cd == 0 ? td.origin[0] : cd == 1 ? td.origin[1] : cd == 2 ? td.origin[2] : cd == 3 ? td.origin[3] : td.origin[4];
                    }
                    crd[0] += bq * my_inner_step;
                    switch (my_rank) {
                    case 1: tma_load<1>(dst, map, fb, crd); break;
                    case 2: tma_load<2>(dst, map, fb, crd); break;
                    case 3: tma_load<3>(dst, map, fb, crd); break;
                    case 4: tma_load<4>(dst, map, fb, crd); break;
                    default: tma_load<5>(dst, map, fb, crd); break;
                    }
                } else {
                    const uint32_t id = P.tile_order ? (uint32_t)P.tile_order[pos] : pos;
                    tma_issue_box(P, T.op[bk], map, bq, id, dst, fb);
                }
            }
            if (++stage == S) {
                stage = 0;
                parity ^= 1u;
            }
        }
    } else {
        // ---------------- consumer warps ----------------
        const int t = tid;
        MapThread<1> th0;
        map_thread_init<1>(P, t, th0);
        TmaThread<NIN> th;
        tma_thread_init<NIN>(P, T, t, th);
        int stage = 0;
        uint32_t parity = 0;
        TileDesc td_first = {}; // (first record fetched before the wait, see the producer)
        if (P.tile_desc && blockIdx.x < ntiles) td_first = P.tile_desc[GROUP ? blockIdx.x % ntiles1 : blockIdx.x];
        pdl_wait(); // the output may still be read or written by the previous kernel
        for (uint32_t gpos = blockIdx.x; gpos < ntiles; gpos += grid) {
            uint32_t pos = gpos;
            unsigned char *obase = P.base[0];
            if (GROUP) {
                const uint32_t prob = gpos / ntiles1;
                pos = gpos - prob * ntiles1;
                obase = G->out[prob];
            }
            MapTile<1> tl;
            if (P.tile_desc) {
                TileDesc td = td_first;
                if (gpos != blockIdx.x) td = P.tile_desc[pos];
                tl.id = td.id_full & 0x7fffffffu;
                tl.full = (td.id_full >> 31) != 0;
                tl.ptr[0] = obase + (td.out_off + th0.g_toff[0]);
            } else {
                map_tile_init<1>(P, th0, pos, tl);
                if (GROUP) tl.ptr[0] = obase + (tl.ptr[0] - P.base[0]);
            }
            mbar_wait(smem_u32(&full_bar[stage]), parity);
            tma_consume<CT, RC, NIN, EPT>(P, T, th, th0, tl, t, ring + (size_t)stage * T.stage_bytes);
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&empty_bar[stage])); // this warp is done with the stage
            if (++stage == S) {
                stage = 0;
                parity ^= 1u;
            }
        }
    }
}

template <class CT, int RC, int NIN, int EPT>
__global__ void __launch_bounds__(TMA_THREADS, SB_TMA_MINB)
map_tma_kernel(const __grid_constant__ MapParams P, const __grid_constant__ TmaParams T, const __grid_constant__ CUtensorMap m0,
               const __grid_constant__ CUtensorMap m1, const __grid_constant__ CUtensorMap m2, const __grid_constant__ CUtensorMap m3)
{
    map_tma_body<CT, RC, NIN, EPT, false>(P, T, &m0, &m1, &m2, &m3, nullptr);
}

template <class CT, int RC, int NIN, int EPT>
__global__ void __launch_bounds__(TMA_THREADS, SB_TMA_MINB)
map_tma_group_kernel(const __grid_constant__ MapParams P, const __grid_constant__ TmaParams T, const __grid_constant__ TmaGroup G)
{
    static_assert(NIN <= TMA_GROUP_MAXIN, "grouped launches carry two tensor maps per problem");
    map_tma_body<CT, RC, NIN, EPT, true>(P, T, nullptr, nullptr, nullptr, nullptr, &G);
}

struct TmaEntry {
    KernelKey key;
    cudaError_t (*launch)(const MapParams &, const TmaParams &, const CUtensorMap *, int grid, size_t smem, cudaStream_t);
    cudaError_t (*occupancy)(int *nblocks, size_t smem);
    const void *func;
};

template <class CT, int RC, int NIN, int EPT> struct TmaLaunch {
    static cudaError_t launch(const MapParams &P, const TmaParams &T, const CUtensorMap *maps, int grid, size_t smem, cudaStream_t s)
    {
        auto k = map_tma_kernel<CT, RC, NIN, EPT>;
        cudaError_t e = ensure_dynamic_smem((const void *)k, smem);
        if (e != cudaSuccess) return e;
        return launch_pdl(k, grid, TMA_THREADS, smem, s, P, T, maps[0], maps[1], maps[2], maps[3]);
    }
    static cudaError_t occupancy(int *nb, size_t smem)
    {
        auto k = map_tma_kernel<CT, RC, NIN, EPT>;
        cudaError_t e = ensure_dynamic_smem((const void *)k, smem);
        if (e != cudaSuccess) return e;
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(nb, k, TMA_THREADS, smem);
    }
    static const void *func() { return (const void *)map_tma_kernel<CT, RC, NIN, EPT>; }
};

#define SB_TMA_ENTRY(CT, DT, RC, NIN, EPT)                                                                           \
    TmaEntry { KernelKey{DT, RC, NIN, EPT, 1}, &TmaLaunch<CT, RC, NIN, EPT>::launch, &TmaLaunch<CT, RC, NIN, EPT>::occupancy, \
               TmaLaunch<CT, RC, NIN, EPT>::func() }

const TmaEntry *tma_table(int *n);
const TmaEntry *find_tma_kernel(const KernelKey &k);

// grouped launches (kernels_tma_group.cu): the same kernels for <= TMA_GROUP_MAXIN inputs
struct TmaGroupEntry {
    KernelKey key;
    cudaError_t (*launch)(const MapParams &, const TmaParams &, const TmaGroup &, int grid, size_t smem, cudaStream_t);
    cudaError_t (*occupancy)(int *nblocks, size_t smem);
    const void *func;
};

template <class CT, int RC, int NIN, int EPT> struct TmaGroupLaunch {
    static cudaError_t launch(const MapParams &P, const TmaParams &T, const TmaGroup &G, int grid, size_t smem, cudaStream_t s)
    {
        auto k = map_tma_group_kernel<CT, RC, NIN, EPT>;
        cudaError_t e = ensure_dynamic_smem((const void *)k, smem);
        if (e != cudaSuccess) return e;
        return launch_pdl(k, grid, TMA_THREADS, smem, s, P, T, G);
    }
    static cudaError_t occupancy(int *nb, size_t smem)
    {
        auto k = map_tma_group_kernel<CT, RC, NIN, EPT>;
        cudaError_t e = ensure_dynamic_smem((const void *)k, smem);
        if (e != cudaSuccess) return e;
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(nb, k, TMA_THREADS, smem);
    }
    static const void *func() { return (const void *)map_tma_group_kernel<CT, RC, NIN, EPT>; }
};

#define SB_TMA_GROUP_ENTRY(CT, DT, RC, NIN, EPT)                                                                     \
    TmaGroupEntry { KernelKey{DT, RC, NIN, EPT, 1}, &TmaGroupLaunch<CT, RC, NIN, EPT>::launch,                       \
                    &TmaGroupLaunch<CT, RC, NIN, EPT>::occupancy, TmaGroupLaunch<CT, RC, NIN, EPT>::func() }

const TmaGroupEntry *find_tma_group_kernel(const KernelKey &k);

} // namespace sb
