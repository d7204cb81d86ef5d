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
    // (4-byte eltypes with <= 8 value words: the dense-copy kernels need 5 CTAs per SM -- 172 vs 189 us for a 2^27 Float32 copy
    //  when the per-tile record registers pushed them to 52; the 8-byte EPT = 4 kernels are FASTER at 4 CTAs per SM: config 3
    //  warm 3.21 vs 3.59 us)
    static constexpr int value = (words <= 8 && sizeof(CT) == 4) ? 5 : (words <= 32 ? 4 : (words <= 64 ? 2 : 1));
};

template <class CT, int RC, int NIN, int EPT, bool UNIFORM, bool GROUP>
__device__ __forceinline__ void map_tile_body_impl(const MapParams &P, const MapGroup *G)
{
    extern __shared__ __align__(16) unsigned char sb_smem_raw[];
    const int t = threadIdx.x;
    pdl_launch_dependents();
    MapThread<NIN + 1> th;
    map_thread_init<NIN + 1>(P, t, th);
    const bool staged = P.nstaged > 0;
    const uint32_t ntiles1 = (uint32_t)P.ntiles;                              // tiles of one problem
    const uint32_t ntiles = GROUP ? ntiles1 * (uint32_t)G->nprob : ntiles1;  // positions of this launch
    // per-tile records (MapParams::lsu_desc) are written once, at plan creation: the first one is fetched before
    // griddepcontrol.wait, the next one while the current tile is in flight
    const bool pre = P.lsu_desc != nullptr && P.lsu_prefetch != 0;
    int64_t rec[NIN + 2];
    if (pre && blockIdx.x < ntiles) map_tile_record<NIN + 1>(P, GROUP ? blockIdx.x % ntiles1 : blockIdx.x, rec);
    pdl_wait(); // first access to operand memory below
    for (uint32_t gpos = blockIdx.x; gpos < ntiles; gpos += gridDim.x) {
        uint32_t pos = gpos, prob = 0;
        if (GROUP) {
            prob = gpos / ntiles1;
            pos = gpos - prob * ntiles1;
        }
        MapTile<NIN + 1> tl;
        if (pre) {
            map_tile_from_record<NIN + 1>(P, th, rec, tl);
            const uint32_t nx = gpos + gridDim.x;
            if (nx < ntiles) map_tile_record<NIN + 1>(P, GROUP ? nx % ntiles1 : nx, rec);
        } else {
            map_tile_init<NIN + 1>(P, th, pos, tl);
        }
        if (GROUP) {
#pragma unroll
            for (int k = 0; k <= NIN; ++k)
                if (k < P.nops) tl.ptr[k] += G->delta[prob][k];
        }
        CT v[NIN][EPT];
        map_phase1<CT, NIN, EPT, UNIFORM>(P, th, tl, t, v, sb_smem_raw);
        if (staged) __syncthreads();
        map_phase2<CT, RC, NIN, EPT, UNIFORM>(P, th, tl, t, v, sb_smem_raw);
        if (staged) __syncthreads();
    }
}
template <class CT, int RC, int NIN, int EPT, bool UNIFORM> __device__ __forceinline__ void map_tile_body(const MapParams &P)
{
    map_tile_body_impl<CT, RC, NIN, EPT, UNIFORM, false>(P, nullptr);
}

// ---- fused exchange across GPUs (peer group) ---------------------------------------------------------------
// The thread that holds the FINAL local value of output `o` (single output tile plans) exchanges it with all ranks
// before the store: it pushes the value into slot [epoch parity][rank][logical index] of every rank's buffer as two
// 8-byte stores {32-bit half, epoch} (atomic over NVLink: no flag, fence or barrier), polls its own copies until every
// rank's halves carry the epoch, and folds them in rank order (bit-identical on all ranks).  Same wire format as the
// stand-alone peer_allreduce_kernel of abi.cu, so ranks may mix the two paths.
SB_D bool peer_logical_index(const ReduceParams &P, int o, int &idx)
{
    int64_t lin = 0, mul = 1;
    bool ok = true;
    int coord[MAXD];
    for (int d = 0; d < P.nkept; ++d) coord[d] = 0;
    for (int i = 0; i < P.kept_order.n; ++i) coord[P.tdim[P.kept_order.td[i]]] = field_of(P.kept_order, i, o);
    for (int d = 0; d < P.nkept; ++d) { // (single output tile: the tile origin is 0)
        ok = ok && coord[d] < P.dims[d];
        lin += coord[d] * mul;
        mul *= P.dims[d];
    }
    idx = (int)lin;
    return ok && lin < PEER_MAX_OUT;
}
// Executed by ALL threads of the one CTA that performs the exchange (CTA-uniform call site): the epoch of this
// collective call = device counter + 1; thread 0 advances the counter once everybody has read it.
SB_D uint32_t peer_epoch_begin(const PeerLink &L)
{
#if defined(__CUDA_ARCH__)
    if (L.world <= 1) return 0u;
    const uint32_t e = __ldcg(L.epoch_ptr) + 1u;
    __syncthreads();
    if (threadIdx.x == 0) __stcg(L.epoch_ptr, e);
    return e;
#else
    (void)L;
    return 0u;
#endif
}
template <class AT, class PT> SB_D AT peer_ll_exchange(const PT &P, int idx, AT p, uint32_t epoch);
template <class AT> SB_D AT peer_ll_allreduce(const ReduceParams &P, int o, AT p, uint32_t epoch)
{
    int idx;
    if (!peer_logical_index(P, o, idx)) return p; // padding lane of the output tile: nothing is stored for it
    return peer_ll_exchange<AT, ReduceParams>(P, idx, p, epoch);
}
// `idx`: logical output index on the wire (kept dims in ascending |output stride|, first dim fastest).
// PT: ReduceParams or StreamArgs (members `peer` and `op`).
template <class AT, class PT> SB_D AT peer_ll_exchange(const PT &P, int idx, AT p, uint32_t epoch)
{
#if defined(__CUDA_ARCH__)
    static_assert(sizeof(AT) <= 8, "the low-latency exchange carries 4- and 8-byte elements");
    union {
        AT v;
        uint32_t w[2];
    } u;
    u.w[1] = 0u;
    u.v = p;
    const size_t par = (size_t)(epoch & 1u) * PEER_MAX_WORLD;
    const size_t off = PEER_DATA_OFF + ((par + (size_t)P.peer.rank) * PEER_MAX_OUT + (size_t)idx) * PEER_SLOT;
    for (int g = 0; g < P.peer.world; ++g) {
        unsigned char *dst = P.peer.buf[g] + off;
        asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(dst), "r"(u.w[0]), "r"(epoch) : "memory");
        asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(dst + 8), "r"(u.w[1]), "r"(epoch) : "memory");
    }
    AT tot = p;
    const long long t0 = clock64();
    for (int g = 0; g < P.peer.world; ++g) {
        const unsigned char *src = P.peer.buf[P.peer.rank] + PEER_DATA_OFF + ((par + (size_t)g) * PEER_MAX_OUT + (size_t)idx) * PEER_SLOT;
        uint32_t a0, e0, a1, e1;
        for (;;) {
            asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(a0), "=r"(e0) : "l"(src) : "memory");
            asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(a1), "=r"(e1) : "l"(src + 8) : "memory");
            if (e0 == epoch && e1 == epoch) break;
            if (clock64() - t0 > P.peer.timeout_cycles) { // a missing peer must not hang the GPU -- and must not kill the
                *reinterpret_cast<volatile uint32_t *>(P.peer.err_flag) = 1u; // context either: flag it, the host reports it
                return p;
            }
        }
        u.w[0] = a0;
        u.w[1] = a1;
        tot = g == 0 ? u.v : red_apply<AT>(P.op, tot, u.v);
    }
    return tot;
#else
    (void)P;
    (void)idx;
    (void)epoch;
    return p;
#endif
}
template <class AT> SB_D AT peer_maybe(const ReduceParams &P, int o, AT p, uint32_t epoch)
{
    if constexpr (sizeof(AT) <= 8) {
        if (P.peer.world > 1) return peer_ll_allreduce<AT>(P, o, p, epoch);
    }
    (void)epoch;
    return p;
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
    uint32_t epoch = 0u; // (fused exchange: single output tile plans only, so exactly one CTA exchanges)
    if (P.nsplit == 1) epoch = peer_epoch_begin(P.peer);
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
            if (P.nsplit == 1) q = peer_maybe<AT>(P, 0, q, epoch);
            red_finish<AT, UNIFORM>(P, bid, 0, q);
        }
    } else if (P.warp_per_output) {
        const int warp = t >> 5, lane = t & 31;
        for (int o = warp; o < P.nout_tile; o += THREADS / 32) {
            AT p = red_lane_partial<AT>(P, smem, o, lane);
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) p = red_apply<AT>(P.op, p, shfl_xor_any(p, m));
            if (lane == 0) red_finish<AT, UNIFORM>(P, bid, o, P.nsplit == 1 ? peer_maybe<AT>(P, o, p, epoch) : p);
        }
    } else {
        for (int o = t; o < P.nout_tile; o += THREADS) {
            const AT p = red_thread_partial<AT>(P, smem, o);
            red_finish<AT, UNIFORM>(P, bid, o, P.nsplit == 1 ? peer_maybe<AT>(P, o, p, epoch) : p);
        }
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
            epoch = peer_epoch_begin(P.peer);
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
                    red_finalize_store<AT, UNIFORM>(P, (int64_t)out_tile, peer_maybe<AT>(P, 0, q, epoch));
                }
            } else if (P.nout_tile >= 32 && P.peer.world <= 1) {
                // many outputs per tile (`sum(A; dims=2)` of a column-major matrix): one THREAD per output, splits folded in
                // split order, eight coalesced L2 loads in flight per thread.  The warp-per-output loop below pays one
                // dependent L2 round trip per EIGHT outputs: 2048 outputs x 293 splits took ~1 ms in the last CTA.
                const AT *sc = reinterpret_cast<const AT *>(P.scratch);
                const int64_t stride = P.nouttiles * (int64_t)P.nout_tile;
                for (int o = t; o < P.nout_tile; o += THREADS) {
                    const int64_t out_idx = (int64_t)out_tile * P.nout_tile + o;
                    AT p = red_neutral<AT>(P.op);
                    for (int s0 = 0; s0 < P.nsplit; s0 += 8) {
                        AT v[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) v[u] = (s0 + u < P.nsplit) ? load_partial(sc + (int64_t)(s0 + u) * stride + out_idx) : red_neutral<AT>(P.op);
#pragma unroll
                        for (int u = 0; u < 8; ++u) p = red_apply<AT>(P.op, p, v[u]);
                    }
                    red_finalize_store<AT, UNIFORM>(P, out_idx, p);
                }
            } else {
                for (int o = warp; o < P.nout_tile; o += THREADS / 32) {
                    const int64_t out_idx = (int64_t)out_tile * P.nout_tile + o;
                    AT p = red_finalize_lane<AT>(P, out_idx, lane);
#pragma unroll
                    for (int m = 16; m >= 1; m >>= 1) p = red_apply<AT>(P.op, p, shfl_xor_any(p, m));
                    if (lane == 0) red_finalize_store<AT, UNIFORM>(P, out_idx, peer_maybe<AT>(P, o, p, epoch));
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
