// map_tile.hpp -- body of the N-D tiled strided map kernel (host/device neutral; the __global__ wrapper is
// in kernels.cuh, a CPU thread-grid emulation for tests is in tests/emul/).
//
// One tile = a power-of-two box of THREADS*EPT elements.  Per tile, per thread:
//   phase 1: issue ALL global loads of the tile (every input, each in its own fastest-stride order so a
//            warp reads contiguous runs), then write the "staged" inputs -- those whose order differs from
//            the output's -- to padded shared-memory buffers;
//   barrier
//   phase 2: read the staged values back in OUTPUT order, evaluate f, store coalesced along the output's
//            fastest dim.
// This is the GPU replacement of the reference's blocked loop nest (src/mapreduce.jl:229-425, map mode
// `A1[I1] = f(A2[I2], ...)`, :311): its cache blocks (`_computeblocks`, :463-500) become shared-memory
// tiles, its loop-order heuristic (`_mapreduce_order!`, :119-139) becomes per-operand load orders.
//
// Cost discipline (profiles/r01_v0 -> r01_v1): everything per tile is O(ndim + nops) 32-bit work held in
// registers (magic-number division, fully unrolled loops guarded by uniform predicates, no local arrays);
// per element it is one 64-bit add per access (all functionals are pre-scaled to bytes).
#pragma once
#include "functors.hpp"

namespace sb {

template <int NOPS> struct MapThread {
    int64_t g_toff[NOPS]; // global BYTE offset contributed by t, operand k (its load order)
    int32_t w_toff[NOPS]; // staging-buffer write byte offset contributed by t (own order)
    int32_t r_toff[NOPS]; // staging-buffer read byte offset contributed by t (output order)
    uint32_t c_toff[NOPS]; // packed tile coordinates contributed by t (operand k's load order), for edge masks
};

template <int NOPS> struct MapTile {
    const unsigned char *ptr[NOPS]; // operand base + byte offset of the tile origin
    uint32_t id;
    bool full;
};

template <int NOPS> SB_HD void map_thread_init(const MapParams &P, int t, MapThread<NOPS> &th)
{
#pragma unroll
    for (int k = 0; k < NOPS; ++k) {
        int64_t g = 0;
        int32_t w = 0, r = 0;
        uint32_t cc = 0;
        if (k < P.nops) {
            const OrderTab &o = P.order[k];
            const int lt = lin_t(t, P.vbits);
            for (int i = 0; i < o.n; ++i) {
                const int f = field_of(o, i, lt);
                g += (int64_t)f * P.g_tstr[k][i];
                w += f * P.w_tstr[k][i];
                cc += (uint32_t)f * P.c_tstr[k][i];
            }
            const OrderTab &oo = P.order[0];
            for (int i = 0; i < oo.n; ++i) r += field_of(oo, i, lt) * P.r_tstr[k][i];
        }
        th.g_toff[k] = g;
        th.w_toff[k] = w;
        th.r_toff[k] = r;
        th.c_toff[k] = cc;
    }
}

// per-tile record of launch position `pos` (MapParams::lsu_desc): nops + 1 words
template <int NOPS> SB_HD void map_tile_record(const MapParams &P, uint32_t pos, int64_t (&r)[NOPS + 1])
{
    const int64_t *src = P.lsu_desc + (size_t)pos * (size_t)(P.nops + 1);
#pragma unroll
    for (int k = 0; k <= NOPS; ++k) r[k] = k <= P.nops ? src[k] : 0;
}
template <int NOPS> SB_HD void map_tile_from_record(const MapParams &P, const MapThread<NOPS> &th, const int64_t (&r)[NOPS + 1], MapTile<NOPS> &tl)
{
    const uint32_t w = (uint32_t)r[0];
    tl.id = (uint32_t)((uint64_t)r[0] >> 32); // with a record, `id` carries the tile's packed edge mask (map_tile_rem, precomputed)
    tl.full = (w >> 31) != 0;
#pragma unroll
    for (int k = 0; k < NOPS; ++k) tl.ptr[k] = P.base[k < P.nops ? k : 0] + (th.g_toff[k] + (k < P.nops ? r[1 + k] : 0));
}

template <int NOPS> SB_HD void map_tile_init(const MapParams &P, const MapThread<NOPS> &th, uint32_t pos, MapTile<NOPS> &tl)
{
    if (P.lsu_desc) { // precomputed record (planner.cpp, "per-tile records for the LSU kernel"): same values as the decode below
        int64_t r[NOPS + 1];
        map_tile_record<NOPS>(P, pos, r);
        map_tile_from_record<NOPS>(P, th, r, tl);
        return;
    }
    uint32_t id = P.tile_order ? (uint32_t)P.tile_order[pos] : pos;
    tl.id = id;
    int64_t off[NOPS];
#pragma unroll
    for (int k = 0; k < NOPS; ++k) off[k] = th.g_toff[k];
    bool full = true;
#pragma unroll
    for (int d = 0; d < MAXD; ++d) {
        if (d < P.ndim) {
            uint32_t q, c;
            fast_divmod(P.tdiv[d], id, q, c);
            id = q;
            const bool shifted = P.shift_last && P.excess[d] != 0 && (int32_t)c == P.ntile[d] - 1; // last tile pulled back: full
            full = full && (shifted || (int32_t)c < P.nfull[d]);
#pragma unroll
            for (int k = 0; k < NOPS; ++k)
                if (k < P.nops) {
                    off[k] += (int64_t)c * P.tstep[k][d];
                    if (shifted) off[k] -= (int64_t)P.excess[d] * P.strides[k][d] * dtype_size(P.dtype[k]);
                }
        }
    }
#pragma unroll
    for (int k = 0; k < NOPS; ++k) tl.ptr[k] = P.base[k < P.nops ? k : 0] + off[k];
    tl.full = full;
}

// Packed remaining extents (R | guard) of tile `id` -- only needed on edge tiles.
SB_HD uint32_t map_tile_rem(const MapParams &P, uint32_t id)
{
    int64_t origin[MAXD];
    for (int d = 0; d < P.ndim; ++d) {
        uint32_t q, c;
        fast_divmod(P.tdiv[d], id, q, c);
        id = q;
        origin[d] = map_tile_origin(P, d, c);
    }
    int32_t rem[MAXTD];
    for (int i = 0; i < P.ntd; ++i) {
        int64_t r = P.dims[P.tdim[i]] - origin[P.tdim[i]];
        if (r > P.tile_b[P.tdim[i]]) r = P.tile_b[P.tdim[i]]; // (balanced tiles use fewer coordinates than the box holds)
        rem[i] = (int32_t)r;
    }
    return pack_rem(rem, P.ntd, P.cpos, P.cbits, P.guard);
}

// is element (t, j) of operand k's traversal inside the array?  (only evaluated on edge tiles)
template <int NOPS> SB_HD bool map_valid(const MapParams &P, const MapThread<NOPS> &th, uint32_t rg, int k, int j)
{
    return packed_valid(rg, th.c_toff[k] + P.c_joff[k][j], P.guard);
}

// 16-byte group of V consecutive elements of a traversal (V = 16 / sizeof(CT); V == 1: no vector path)
template <class CT> struct VecOf {
    static constexpr int V = sizeof(CT) >= 16 ? 1 : (int)(16 / sizeof(CT));
    struct alignas(16) type {
        CT e[V];
    };
};
template <class CT> SB_HD void store_vec16(unsigned char *p, const typename VecOf<CT>::type &x)
{
#if defined(__CUDA_ARCH__)
    __stcs(reinterpret_cast<float4 *>(p), *reinterpret_cast<const float4 *>(&x)); // streaming: the output is not re-read
#else
    *reinterpret_cast<typename VecOf<CT>::type *>(p) = x;
#endif
}

// Phase 1.  v[k-1][j] receives input k's element (t, j) in input k's LOAD order.
template <class CT, int NIN, int EPT, bool UNIFORM>
SB_HD void map_phase1(const MapParams &P, const MapThread<NIN + 1> &th, const MapTile<NIN + 1> &tl, int t,
                      CT (&v)[NIN][EPT], unsigned char *smem)
{
    constexpr int V = VecOf<CT>::V;
    constexpr bool VEC = UNIFORM && V > 1 && (EPT % V == 0);
    using vec_t = typename VecOf<CT>::type;
    if (tl.full && !P.umask) {
#pragma unroll
        for (int k = 1; k <= NIN; ++k) {
            if (k < P.nops) {
                if (VEC && P.gvec[k]) { // 128-bit loads: V elements contiguous along this operand's traversal
#pragma unroll
                    for (int j = 0; j < EPT; j += V) {
                        const vec_t x = *reinterpret_cast<const vec_t *>(tl.ptr[k] + P.g_joff[k][j]);
#pragma unroll
                        for (int u = 0; u < V; ++u) v[k - 1][j + u] = x.e[u];
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < EPT; ++j) v[k - 1][j] = load_elem<CT, UNIFORM>(tl.ptr[k] + P.g_joff[k][j], P.dtype[k], P.conj[k]);
                }
            } else {
#pragma unroll
                for (int j = 0; j < EPT; ++j) v[k - 1][j] = make<CT>(0.0, 0.0);
            }
        }
    } else {
        // masked tile: an edge tile (mask from the tile's remaining extents) or a balanced tile (plan-constant mask; whole
        // 16-byte groups are valid or not, the planner keeps the balanced extents multiples of V)
        const uint32_t rg = tl.full ? P.urg : (P.lsu_desc ? tl.id : map_tile_rem(P, tl.id));
#pragma unroll
        for (int k = 1; k <= NIN; ++k) {
            if (VEC && tl.full && k < P.nops && P.gvec[k]) {
#pragma unroll
                for (int j = 0; j < EPT; j += V) {
                    vec_t x;
#pragma unroll
                    for (int u = 0; u < V; ++u) x.e[u] = make<CT>(0.0, 0.0);
                    if (map_valid(P, th, rg, k, j)) x = *reinterpret_cast<const vec_t *>(tl.ptr[k] + P.g_joff[k][j]);
#pragma unroll
                    for (int u = 0; u < V; ++u) v[k - 1][j + u] = x.e[u];
                }
            } else {
#pragma unroll
                for (int j = 0; j < EPT; ++j) {
                    CT x = make<CT>(0.0, 0.0);
                    if (k < P.nops && map_valid(P, th, rg, k, j)) x = load_elem<CT, UNIFORM>(tl.ptr[k] + P.g_joff[k][j], P.dtype[k], P.conj[k]);
                    v[k - 1][j] = x;
                }
            }
        }
    }
#pragma unroll
    for (int k = 1; k <= NIN; ++k) {
        if (k < P.nops && P.staged[k]) {
            unsigned char *s = smem + P.smem_off[k] + th.w_toff[k];
            if (VEC && P.svec[k]) {
#pragma unroll
                for (int j = 0; j < EPT; j += V) {
                    vec_t x;
#pragma unroll
                    for (int u = 0; u < V; ++u) x.e[u] = v[k - 1][j + u];
                    *reinterpret_cast<vec_t *>(s + P.w_joff[k][j]) = x;
                }
            } else {
#pragma unroll
                for (int j = 0; j < EPT; ++j) *reinterpret_cast<CT *>(s + P.w_joff[k][j]) = v[k - 1][j];
            }
        }
    }
}

// Phase 2 (after the barrier).
template <class CT, int RC, int NIN, int EPT, bool UNIFORM>
SB_HD void map_phase2(const MapParams &P, const MapThread<NIN + 1> &th, const MapTile<NIN + 1> &tl, int t,
                      CT (&v)[NIN][EPT], const unsigned char *smem)
{
#pragma unroll
    for (int k = 1; k <= NIN; ++k) {
        if (k < P.nops && P.staged[k]) {
            const unsigned char *s = smem + P.smem_off[k] + th.r_toff[k];
#pragma unroll
            for (int j = 0; j < EPT; ++j) v[k - 1][j] = *reinterpret_cast<const CT *>(s + P.r_joff[k][j]);
        }
    }
    ElemFn<CT, RC> fn;
    unsigned char *ob = const_cast<unsigned char *>(tl.ptr[0]);
    constexpr int V = VecOf<CT>::V;
    constexpr bool VEC = UNIFORM && V > 1 && (EPT % V == 0);
    const bool um = P.umask != 0;
    if (tl.full && VEC && P.gvec[0]) {
#pragma unroll
        for (int j = 0; j < EPT; j += V) {
            typename VecOf<CT>::type x;
#pragma unroll
            for (int u = 0; u < V; ++u) {
                CT a[NIN];
#pragma unroll
                for (int k = 0; k < NIN; ++k) a[k] = v[k][j + u];
                x.e[u] = fn.template eval<NIN>(P.prog, a);
            }
            if (!um || map_valid(P, th, P.urg, 0, j)) store_vec16<CT>(ob + P.g_joff[0][j], x);
        }
    } else if (tl.full && !um) {
#pragma unroll
        for (int j = 0; j < EPT; ++j) {
            CT a[NIN];
#pragma unroll
            for (int k = 0; k < NIN; ++k) a[k] = v[k][j];
            store_elem<CT, UNIFORM>(ob + P.g_joff[0][j], P.dtype[0], P.conj[0], fn.template eval<NIN>(P.prog, a));
        }
    } else {
        const uint32_t rg = tl.full ? P.urg : (P.lsu_desc ? tl.id : map_tile_rem(P, tl.id));
#pragma unroll
        for (int j = 0; j < EPT; ++j) {
            CT a[NIN];
#pragma unroll
            for (int k = 0; k < NIN; ++k) a[k] = v[k][j];
            const CT r = fn.template eval<NIN>(P.prog, a);
            if (map_valid(P, th, rg, 0, j)) store_elem<CT, UNIFORM>(ob + P.g_joff[0][j], P.dtype[0], P.conj[0], r);
        }
    }
}

} // namespace sb
