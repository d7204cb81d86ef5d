// map_tile.hpp -- body of the N-D tiled strided map kernel (host/device neutral; the __global__ wrapper is
// in kernels_map.cuh, a CPU thread-grid emulation for tests is in tests/emul/).
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
#pragma once
#include "functors.hpp"

namespace sb {

template <int NOPS> struct MapThread {
    int64_t g_toff[NOPS]; // global element offset contributed by t, operand k (its load order)
    int32_t w_toff[NOPS]; // staging-buffer write slot contributed by t (own order)
    int32_t r_toff[NOPS]; // staging-buffer read slot contributed by t (output order)
};

struct MapTile {
    int64_t base[MAXO]; // element offset of the tile origin, per operand
    int32_t rem[MAXTD]; // remaining extent per tile-dim slot (>= tile extent for interior tiles)
    bool full;
};

template <int NOPS> SB_HD void map_thread_init(const MapParams &P, int t, MapThread<NOPS> &th)
{
#pragma unroll
    for (int k = 0; k < NOPS; ++k) {
        if (k >= P.nops) {
            th.g_toff[k] = 0;
            th.w_toff[k] = 0;
            th.r_toff[k] = 0;
            continue;
        }
        const OrderTab &o = P.order[k];
        int64_t g = 0;
        int32_t w = 0;
        for (int i = 0; i < o.n; ++i) {
            const int f = field_of(o, i, t);
            g += (int64_t)f * P.g_tstr[k][i];
            w += f * P.w_tstr[k][i];
        }
        th.g_toff[k] = g;
        th.w_toff[k] = w;
        int32_t r = 0;
        const OrderTab &oo = P.order[0];
        for (int i = 0; i < oo.n; ++i) r += field_of(oo, i, t) * P.r_tstr[k][i];
        th.r_toff[k] = r;
    }
}

SB_HD void map_tile_init(const MapParams &P, int64_t pos, MapTile &tl)
{
    int64_t id = P.tile_order ? (int64_t)P.tile_order[pos] : pos;
    for (int k = 0; k < MAXO; ++k) tl.base[k] = 0;
    int32_t origin[MAXD];
    for (int d = 0; d < P.ndim; ++d) {
        const int64_t q = id / P.ntile[d];
        const int32_t c = (int32_t)(id - q * P.ntile[d]);
        id = q;
        origin[d] = c * P.tile_b[d];
        for (int k = 0; k < P.nops; ++k) tl.base[k] += (int64_t)origin[d] * P.strides[k][d];
    }
    bool full = true;
    for (int i = 0; i < P.ntd; ++i) {
        const int d = P.tdim[i];
        const int64_t r = P.dims[d] - origin[d];
        tl.rem[i] = r > 0x7fffffff ? 0x7fffffff : (int32_t)r;
        full = full && (r >= P.tile_b[d]);
    }
    tl.full = full;
}

// is element (t, j) of operand k's traversal inside the array?  (only evaluated on edge tiles)
SB_HD bool map_valid(const MapParams &P, const MapTile &tl, int k, int t, int j)
{
    const OrderTab &o = P.order[k];
    bool ok = true;
    for (int i = 0; i < o.n; ++i) {
        const int c = field_of(o, i, t) + (int)P.jfield[k][j][i];
        ok = ok && (c < tl.rem[o.td[i]]);
    }
    return ok;
}

// Phase 1.  v[k-1][j] receives input k's element (t, j) in input k's LOAD order.
template <class CT, int NIN, int EPT, bool UNIFORM>
SB_HD void map_phase1(const MapParams &P, const MapThread<NIN + 1> &th, const MapTile &tl, int t, CT (&v)[NIN][EPT],
                      CT *smem)
{
#pragma unroll
    for (int k = 1; k <= NIN; ++k) {
        if (k >= P.nops) { // unused input slot of a wider instantiation
#pragma unroll
            for (int j = 0; j < EPT; ++j) v[k - 1][j] = make<CT>(0.0, 0.0);
            continue;
        }
        const unsigned char *b = P.base[k];
        const int64_t o0 = tl.base[k] + th.g_toff[k];
        if (tl.full) {
#pragma unroll
            for (int j = 0; j < EPT; ++j)
                v[k - 1][j] = load_elem<CT, UNIFORM>(b, o0 + P.g_joff[k][j], P.dtype[k], P.conj[k]);
        } else {
#pragma unroll
            for (int j = 0; j < EPT; ++j) {
                CT x = make<CT>(0.0, 0.0);
                if (map_valid(P, tl, k, t, j)) x = load_elem<CT, UNIFORM>(b, o0 + P.g_joff[k][j], P.dtype[k], P.conj[k]);
                v[k - 1][j] = x;
            }
        }
    }
#pragma unroll
    for (int k = 1; k <= NIN; ++k) {
        if (k >= P.nops || !P.staged[k]) continue;
        CT *s = smem + P.smem_off[k];
#pragma unroll
        for (int j = 0; j < EPT; ++j) s[th.w_toff[k] + P.w_joff[k][j]] = v[k - 1][j];
    }
}

// Phase 2 (after the barrier).
template <class CT, int RC, int NIN, int EPT, bool UNIFORM>
SB_HD void map_phase2(const MapParams &P, const MapThread<NIN + 1> &th, const MapTile &tl, int t, CT (&v)[NIN][EPT],
                      const CT *smem)
{
#pragma unroll
    for (int k = 1; k <= NIN; ++k) {
        if (k >= P.nops || !P.staged[k]) continue;
        const CT *s = smem + P.smem_off[k];
#pragma unroll
        for (int j = 0; j < EPT; ++j) v[k - 1][j] = s[th.r_toff[k] + P.r_joff[k][j]];
    }
    ElemFn<CT, RC> fn;
    unsigned char *ob = P.base[0];
    const int64_t o0 = tl.base[0] + th.g_toff[0];
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
        CT a[NIN];
#pragma unroll
        for (int k = 0; k < NIN; ++k) a[k] = v[k][j];
        const CT r = fn.template eval<NIN>(P.prog, a);
        if (tl.full || map_valid(P, tl, 0, t, j)) store_elem<CT, UNIFORM>(ob, o0 + P.g_joff[0][j], P.dtype[0], P.conj[0], r);
    }
}

} // namespace sb
