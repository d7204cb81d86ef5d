// kernel_bodies.cuh -- device bodies of the tile kernels (no host includes: this header is also compiled by NVRTC for
// the run-time specialised element functions, csrc/jit.cpp).
#pragma once
#include "map_tile.hpp"
#include "reduce_tile.hpp"

namespace sb {

// ---- map ------------------------------------------------------------------------------------------------
// Persistent CTAs: grid = min(ntiles, SMs x resident CTAs); each CTA walks tiles pos = blockIdx.x + i*grid,
// so neighbouring CTAs work on neighbouring tiles at the same time (DRAM page / L2 locality, and aliased
// operands such as A and A' meet in L2).
// resident CTAs per SM the register allocator is asked to allow (value registers = NIN*EPT words of CT)
template <class CT, int NIN, int EPT> struct MinBlocks {
    static constexpr int words = NIN * EPT * (int)(sizeof(CT) / 4);
    static constexpr int value = words <= 32 ? 4 : (words <= 64 ? 2 : 1);
};

template <class CT, int RC, int NIN, int EPT, bool UNIFORM> __device__ __forceinline__ void map_tile_body(const MapParams &P)
{
    extern __shared__ __align__(16) unsigned char sb_smem_raw[];
    const int t = threadIdx.x;
    pdl_launch_dependents();
    MapThread<NIN + 1> th;
    map_thread_init<NIN + 1>(P, t, th);
    const bool staged = P.nstaged > 0;
    pdl_wait(); // first global access below
    const uint32_t ntiles = (uint32_t)P.ntiles;
    for (uint32_t pos = blockIdx.x; pos < ntiles; pos += gridDim.x) {
        MapTile<NIN + 1> tl;
        map_tile_init<NIN + 1>(P, th, pos, tl);
        CT v[NIN][EPT];
        map_phase1<CT, NIN, EPT, UNIFORM>(P, th, tl, t, v, sb_smem_raw);
        if (staged) __syncthreads();
        map_phase2<CT, RC, NIN, EPT, UNIFORM>(P, th, tl, t, v, sb_smem_raw);
        if (staged) __syncthreads();
    }
}

// ---- reduce ---------------------------------------------------------------------------------------------
template <class T> __device__ __forceinline__ T shfl_xor_any(T v, int mask)
{
    constexpr int W = sizeof(T) / 4;
    union {
        T v;
        uint32_t w[W];
    } a, b;
    a.v = v;
#pragma unroll
    for (int i = 0; i < W; ++i) b.w[i] = __shfl_xor_sync(0xffffffffu, a.w[i], mask);
    return b.v;
}

template <class AT, int RC, int NIN, int EPT, bool UNIFORM> __device__ __forceinline__ void reduce_tile_body(const ReduceParams &P)
{
    extern __shared__ __align__(16) unsigned char sb_smem_raw[];
    AT *smem = reinterpret_cast<AT *>(sb_smem_raw);
    const int t = threadIdx.x;
    const uint32_t bid = blockIdx.x;
    pdl_launch_dependents();
    pdl_wait(); // (the previous launch may be a reduction re-arming the same arrival counters)
    red_accumulate<AT, RC, NIN, EPT, UNIFORM>(P, bid, t, smem);
    __syncthreads();
    if (P.warp_per_output && P.nout_tile == 1) {
        // a single output per CTA: all eight warps fold the THREADS*EPT accumulators (layout [0][r]), fixed order
        const int warp = t >> 5, lane = t & 31;
        AT p = red_neutral<AT>(P.op);
        for (int r = t; r < P.nred_tile; r += THREADS) p = red_apply<AT>(P.op, p, smem[r]);
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) p = red_apply<AT>(P.op, p, shfl_xor_any(p, m));
        __syncthreads();
        if (lane == 0) smem[warp] = p;
        __syncthreads();
        if (t == 0) {
            AT q = smem[0];
            for (int w = 1; w < THREADS / 32; ++w) q = red_apply<AT>(P.op, q, smem[w]);
            red_finish<AT, UNIFORM>(P, bid, 0, q);
        }
    } else if (P.warp_per_output) {
        const int warp = t >> 5, lane = t & 31;
        for (int o = warp; o < P.nout_tile; o += THREADS / 32) {
            AT p = red_lane_partial<AT>(P, smem, o, lane);
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) p = red_apply<AT>(P.op, p, shfl_xor_any(p, m));
            if (lane == 0) red_finish<AT, UNIFORM>(P, bid, o, p);
        }
    } else {
        for (int o = t; o < P.nout_tile; o += THREADS) red_finish<AT, UNIFORM>(P, bid, o, red_thread_partial<AT>(P, smem, o));
    }
    if (P.nsplit > 1) {
        // Fused finalize: the last split-CTA of an output tile to arrive folds that tile's partials in a fixed order
        // (deterministic, no floating-point atomics) and applies op(initop(out), .) -- saves a second launch.
        __shared__ unsigned int sb_is_last;
        __threadfence(); // this CTA's partials are visible device-wide before it counts itself in
        __syncthreads();
        uint32_t split, out_tile;
        fast_divmod(P.outdiv, bid, split, out_tile);
        if (t == 0) sb_is_last = (atomicAdd(P.counters + out_tile, 1u) == (unsigned)P.nsplit - 1u) ? 1u : 0u;
        __syncthreads();
        if (sb_is_last) {
            __threadfence();
            const int warp = t >> 5, lane = t & 31;
            if (P.nout_tile == 1) {
                // one output (complete reductions; config 5 per GPU): ALL eight warps fold -- thread t takes the splits
                // t, t+256, ... (eight L2 loads in flight), warp butterfly, then the eight warp results in warp order.
                // Same fixed order on every run; with one warp the 586 partials of a 128 MiB shard took three dependent
                // L2 round trips.
                const AT *sc = reinterpret_cast<const AT *>(P.scratch);
                const int64_t stride = P.nouttiles;
                AT p = red_neutral<AT>(P.op);
                for (int s0 = t; s0 < P.nsplit; s0 += 8 * THREADS) {
                    AT v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        v[u] = (s0 + THREADS * u < P.nsplit) ? load_partial(sc + (int64_t)(s0 + THREADS * u) * stride + out_tile) : red_neutral<AT>(P.op);
#pragma unroll
                    for (int u = 0; u < 8; ++u) p = red_apply<AT>(P.op, p, v[u]);
                }
#pragma unroll
                for (int m = 16; m >= 1; m >>= 1) p = red_apply<AT>(P.op, p, shfl_xor_any(p, m));
                __syncthreads(); // the accumulators in `smem` have been consumed by red_finish above
                if (lane == 0) smem[warp] = p;
                __syncthreads();
                if (t == 0) {
                    AT q = smem[0];
                    for (int w = 1; w < THREADS / 32; ++w) q = red_apply<AT>(P.op, q, smem[w]);
                    red_finalize_store<AT, UNIFORM>(P, (int64_t)out_tile, q);
                }
            } else {
                for (int o = warp; o < P.nout_tile; o += THREADS / 32) {
                    const int64_t out_idx = (int64_t)out_tile * P.nout_tile + o;
                    AT p = red_finalize_lane<AT>(P, out_idx, lane);
#pragma unroll
                    for (int m = 16; m >= 1; m >>= 1) p = red_apply<AT>(P.op, p, shfl_xor_any(p, m));
                    if (lane == 0) red_finalize_store<AT, UNIFORM>(P, out_idx, p);
                }
            }
            if (t == 0) P.counters[out_tile] = 0u; // re-arm for the next launch
        }
    }
}

template <class AT, bool UNIFORM> __device__ __forceinline__ void reduce_finalize_body(const ReduceParams &P)
{
    const int64_t out_idx = (int64_t)blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    pdl_wait();
    if (out_idx >= P.nouttiles * (int64_t)P.nout_tile) return; // warp-uniform
    AT p = red_finalize_lane<AT>(P, out_idx, lane);
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) p = red_apply<AT>(P.op, p, shfl_xor_any(p, m));
    if (lane == 0) red_finalize_store<AT, UNIFORM>(P, out_idx, p);
}

} // namespace sb
