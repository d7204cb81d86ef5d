// tma_tile.hpp -- consumer side of the TMA-staged map kernel (host/device neutral).
//
// The producer (one elected thread, kernels.cuh) keeps `nstage` tiles of every input in flight with
// cp.async.bulk.tensor; consumers read a landed stage in OUTPUT order, evaluate f and store coalesced.
// Out-of-bounds box elements are zero-filled by the TMA unit, so only the stores of edge tiles are masked.
#pragma once
#include "map_tile.hpp"

namespace sb {

template <int NIN> struct TmaThread {
    uint32_t s_t[NIN]; // (swizzled) shared-memory byte offset contributed by t, input k
};

// dense byte offset of tile coordinate `f` along output-order slot `slot` inside operand k's box set
SB_HD uint32_t tma_slot_offset(const TmaOperand &o, int slot, int f)
{
    if (slot == o.inner_slot) {
        const int lo = f & ((1 << o.split_bits) - 1);
        return (uint32_t)(lo * o.d_lo[slot] + (f >> o.split_bits) * o.d_hi);
    }
    return (uint32_t)(f * o.d_lo[slot]);
}

template <int NIN> SB_HD void tma_thread_init(const MapParams &P, const TmaParams &T, int t, TmaThread<NIN> &th)
{
    const OrderTab &oo = P.order[0];
#pragma unroll
    for (int k = 0; k < NIN; ++k) {
        uint32_t d = 0;
        if (k < T.nin) {
            for (int i = 0; i < oo.n; ++i) d += tma_slot_offset(T.op[k], i, field_of(oo, i, lin_t(t, P.vbits)));
            if (T.op[k].swizzle) d = swizzle128(d);
        }
        th.s_t[k] = d;
    }
}

// consume one landed stage: `stage` points at the stage base (1024-byte aligned)
template <class CT, int RC, int NIN, int EPT>
SB_HD void tma_consume(const MapParams &P, const TmaParams &T, const TmaThread<NIN> &th, const MapThread<1> &th0, const MapTile<1> &tl, int t,
                       const unsigned char *stage)
{
    constexpr int V = VecOf<CT>::V;
    constexpr bool VEC = V > 1 && (EPT % V == 0);
    CT v[NIN][EPT];
#pragma unroll
    for (int k = 0; k < NIN; ++k) {
        if (k < T.nin) {
            const unsigned char *s = stage + T.op[k].smem_off;
            if (VEC && P.vbits > 0 && !T.op[k].swizzle) { // un-swizzled (direct) operand: 128-bit shared-memory loads
#pragma unroll
                for (int j = 0; j < EPT; j += V) {
                    const typename VecOf<CT>::type x = *reinterpret_cast<const typename VecOf<CT>::type *>(s + (th.s_t[k] ^ (uint32_t)T.op[k].s_joff[j]));
#pragma unroll
                    for (int u = 0; u < V; ++u) v[k][j + u] = x.e[u];
                }
            } else {
#pragma unroll
                for (int j = 0; j < EPT; ++j) v[k][j] = *reinterpret_cast<const CT *>(s + (th.s_t[k] ^ (uint32_t)T.op[k].s_joff[j]));
            }
        } else {
#pragma unroll
            for (int j = 0; j < EPT; ++j) v[k][j] = make<CT>(0.0, 0.0);
        }
    }
    ElemFn<CT, RC> fn;
    unsigned char *ob = const_cast<unsigned char *>(tl.ptr[0]);
    if (P.uniform & 0x100) { // diagnostic (SB_DEBUG=nostore): keep the loads and the math, drop the stores
#pragma unroll
        for (int j = 0; j < EPT; ++j) {
            CT a[NIN];
#pragma unroll
            for (int k = 0; k < NIN; ++k) a[k] = v[k][j];
            const CT r = fn.template eval<NIN>(P.prog, a);
            if (re_of(r) == (typename traits<CT>::real)1.2345678e-30f) store_elem<CT, true>(ob + P.g_joff[0][j], P.dtype[0], 0, r);
        }
    } else if (tl.full && VEC && P.gvec[0]) {
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
            store_vec16<CT>(ob + P.g_joff[0][j], x);
        }
    } else if (tl.full) {
#pragma unroll
        for (int j = 0; j < EPT; ++j) {
            CT a[NIN];
#pragma unroll
            for (int k = 0; k < NIN; ++k) a[k] = v[k][j];
            store_elem<CT, true>(ob + P.g_joff[0][j], P.dtype[0], 0, fn.template eval<NIN>(P.prog, a));
        }
    } else {
        const uint32_t rg = map_tile_rem(P, tl.id);
#pragma unroll
        for (int j = 0; j < EPT; ++j) {
            CT a[NIN];
#pragma unroll
            for (int k = 0; k < NIN; ++k) a[k] = v[k][j];
            const CT r = fn.template eval<NIN>(P.prog, a);
            if (map_valid(P, th0, rg, 0, j)) store_elem<CT, true>(ob + P.g_joff[0][j], P.dtype[0], 0, r);
        }
    }
}

// producer-side coordinates of tile `id` for input k, box q: tensor-map coordinates (own dim order)
SB_HD void tma_box_coords(const MapParams &P, const TmaOperand &o, uint32_t id, int q, int32_t (&crd)[TMA_MAXRANK])
{
    int32_t origin[MAXD];
    for (int d = 0; d < MAXD; ++d) origin[d] = 0;
    for (int d = 0; d < P.ndim; ++d) {
        uint32_t qq, c;
        fast_divmod(P.tdiv[d], id, qq, c);
        id = qq;
        origin[d] = (int32_t)map_tile_origin(P, d, c);
    }
    for (int i = 0; i < TMA_MAXRANK; ++i) crd[i] = (i < o.rank) ? origin[o.cdim[i]] : 0;
    crd[0] += q * o.inner_step;
}

} // namespace sb
