// kernels.cuh -- __global__ wrappers (sm_100a) around the tile bodies + the launch registry.
#pragma once
#include "kernel_bodies.cuh"
#include "planner.hpp"
#include <cuda_runtime.h>
#include <cstdlib>
#include <utility>

namespace sb {

// launch with programmatic stream serialisation (see common.hpp "PDL"); SB_NO_PDL=1 turns the attribute off
// The environment is read when a context is created (and again on sb_ctx_reload_env), never on the launch path.
struct EnvCache {
    bool pdl = true;                       // SB_NO_PDL
    bool fused_peer = true;                // SB_NO_FUSED_PEER
    long long jit_min_elements = 1 << 18;  // SB_JIT_MIN_ELEMENTS
    bool jit_sync = false;                 // SB_JIT_SYNC: block on the NVRTC compile instead of compiling in the background
    bool no_group = false;                 // SB_NO_GROUP: sb_mapreduce_batch never merges same-plan calls into one launch
    long long plan_table_mb = 512;         // SB_PLAN_TABLE_MB: device memory the plan tables of one context may hold before the cache is dropped
};
EnvCache &env_cache();  // abi.cu
void env_reload();      // abi.cu
inline bool pdl_enabled() { return env_cache().pdl; }

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is sticky per (device, function): set it once, and again only when a
// plan needs more than was ever requested (it used to run on every launch).
cudaError_t ensure_dynamic_smem(const void *func, size_t smem); // abi.cu
template <class... KArgs, class... Args>
cudaError_t launch_pdl(void (*k)(KArgs...), int grid, int block, size_t smem, cudaStream_t s, Args &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, k, std::forward<Args>(args)...);
}

// ---- statically instantiated kernels (the bodies live in kernel_bodies.cuh) ------------------------------------
template <class CT, int RC, int NIN, int EPT, bool UNIFORM>
__global__ void __launch_bounds__(THREADS, MinBlocks<CT, NIN, EPT>::value) map_tile_kernel(const __grid_constant__ MapParams P)
{
    map_tile_body<CT, RC, NIN, EPT, UNIFORM>(P);
}
template <class CT, int RC, int NIN, int EPT, bool UNIFORM>
__global__ void __launch_bounds__(THREADS, MinBlocks<CT, NIN, EPT>::value) map_tile_group_kernel(const __grid_constant__ MapParams P, const __grid_constant__ MapGroup G)
{
    map_tile_body_impl<CT, RC, NIN, EPT, UNIFORM, true>(P, &G);
}
template <class AT, int RC, int NIN, int EPT, bool UNIFORM>
__global__ void __launch_bounds__(THREADS, MinBlocks<AT, NIN, EPT>::value) reduce_tile_kernel(const __grid_constant__ ReduceParams P)
{
    reduce_tile_body<AT, RC, NIN, EPT, UNIFORM>(P);
}
template <class AT, bool UNIFORM> __global__ void __launch_bounds__(THREADS) reduce_finalize_kernel(const __grid_constant__ ReduceParams P)
{
    reduce_finalize_body<AT, UNIFORM>(P);
}

// ---- registry ---------------------------------------------------------------------------------------------
struct MapEntry {
    KernelKey key;
    cudaError_t (*launch)(const MapParams &, int grid, size_t smem, cudaStream_t);
    cudaError_t (*occupancy)(int *nblocks, size_t smem);
    const void *func;
};
struct ReduceEntry {
    KernelKey key;
    cudaError_t (*launch)(const ReduceParams &, int grid, size_t smem, cudaStream_t);
    cudaError_t (*finalize)(const ReduceParams &, int grid, cudaStream_t);
    cudaError_t (*occupancy)(int *nblocks, size_t smem);
    const void *func;
};

template <class CT, int RC, int NIN, int EPT, bool U> struct MapLaunch {
    static cudaError_t launch(const MapParams &P, int grid, size_t smem, cudaStream_t s)
    {
        auto k = map_tile_kernel<CT, RC, NIN, EPT, U>;
        if (smem > 48 * 1024) {
            cudaError_t e = ensure_dynamic_smem((const void *)k, smem);
            if (e != cudaSuccess) return e;
        }
        return launch_pdl(k, grid, THREADS, smem, s, P);
    }
    static cudaError_t occupancy(int *nb, size_t smem)
    {
        auto k = map_tile_kernel<CT, RC, NIN, EPT, U>;
        if (smem > 48 * 1024) {
            cudaError_t e = ensure_dynamic_smem((const void *)k, smem);
            if (e != cudaSuccess) return e;
        }
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(nb, k, THREADS, smem);
    }
    static const void *func() { return (const void *)map_tile_kernel<CT, RC, NIN, EPT, U>; }
};

// grouped launches of the LSU map kernel (kernels_map_group.cu: a small set of recipes)
struct MapGroupEntry {
    KernelKey key;
    cudaError_t (*launch)(const MapParams &, const MapGroup &, int grid, size_t smem, cudaStream_t);
    cudaError_t (*occupancy)(int *nblocks, size_t smem);
    const void *func;
};
template <class CT, int RC, int NIN, int EPT, bool U> struct MapGroupLaunch {
    static cudaError_t launch(const MapParams &P, const MapGroup &G, int grid, size_t smem, cudaStream_t s)
    {
        auto k = map_tile_group_kernel<CT, RC, NIN, EPT, U>;
        if (smem > 48 * 1024) {
            cudaError_t e = ensure_dynamic_smem((const void *)k, smem);
            if (e != cudaSuccess) return e;
        }
        return launch_pdl(k, grid, THREADS, smem, s, P, G);
    }
    static cudaError_t occupancy(int *nb, size_t smem)
    {
        auto k = map_tile_group_kernel<CT, RC, NIN, EPT, U>;
        if (smem > 48 * 1024) {
            cudaError_t e = ensure_dynamic_smem((const void *)k, smem);
            if (e != cudaSuccess) return e;
        }
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(nb, k, THREADS, smem);
    }
    static const void *func() { return (const void *)map_tile_group_kernel<CT, RC, NIN, EPT, U>; }
};
#define SB_MAP_GROUP_ENTRY(CT, DT, RC, NIN, EPT)                                                                     \
    MapGroupEntry { KernelKey{DT, RC, NIN, EPT, 1}, &MapGroupLaunch<CT, RC, NIN, EPT, true>::launch,                 \
                    &MapGroupLaunch<CT, RC, NIN, EPT, true>::occupancy, MapGroupLaunch<CT, RC, NIN, EPT, true>::func() }
const MapGroupEntry *find_map_group_kernel(const KernelKey &k);

template <class AT, int RC, int NIN, int EPT, bool U> struct ReduceLaunch {
    static cudaError_t launch(const ReduceParams &P, int grid, size_t smem, cudaStream_t s)
    {
        auto k = reduce_tile_kernel<AT, RC, NIN, EPT, U>;
        if (smem > 48 * 1024) {
            cudaError_t e = ensure_dynamic_smem((const void *)k, smem);
            if (e != cudaSuccess) return e;
        }
        return launch_pdl(k, grid, THREADS, smem, s, P);
    }
    static cudaError_t finalize(const ReduceParams &P, int grid, cudaStream_t s)
    {
        return launch_pdl(reduce_finalize_kernel<AT, U>, grid, THREADS, 0, s, P);
    }
    static cudaError_t occupancy(int *nb, size_t smem)
    {
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(nb, reduce_tile_kernel<AT, RC, NIN, EPT, U>, THREADS, smem);
    }
    static const void *func() { return (const void *)reduce_tile_kernel<AT, RC, NIN, EPT, U>; }
};

#define SB_MAP_ENTRY(CT, DT, RC, NIN, EPT, U)                                                                        \
    MapEntry { KernelKey{DT, RC, NIN, EPT, U}, &MapLaunch<CT, RC, NIN, EPT, (U) != 0>::launch,                       \
               &MapLaunch<CT, RC, NIN, EPT, (U) != 0>::occupancy, MapLaunch<CT, RC, NIN, EPT, (U) != 0>::func() }
#define SB_RED_ENTRY(CT, DT, RC, NIN, EPT, U)                                                                        \
    ReduceEntry { KernelKey{DT, RC, NIN, EPT, U}, &ReduceLaunch<CT, RC, NIN, EPT, (U) != 0>::launch,                 \
                  &ReduceLaunch<CT, RC, NIN, EPT, (U) != 0>::finalize,                                               \
                  &ReduceLaunch<CT, RC, NIN, EPT, (U) != 0>::occupancy, ReduceLaunch<CT, RC, NIN, EPT, (U) != 0>::func() }

// one table per compute type, each in its own translation unit (parallel nvcc)
const MapEntry *map_table_f32(int *n);
const MapEntry *map_table_f64(int *n);
const MapEntry *map_table_c32(int *n);
const MapEntry *map_table_c64(int *n);
const ReduceEntry *reduce_table_f32(int *n);
const ReduceEntry *reduce_table_f64(int *n);
const ReduceEntry *reduce_table_c32(int *n);
const ReduceEntry *reduce_table_c64(int *n);

const MapEntry *find_map_kernel(const KernelKey &k);
const ReduceEntry *find_reduce_kernel(const KernelKey &k);

} // namespace sb
