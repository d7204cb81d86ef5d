// reduce_tile.hpp -- body of the tiled strided map-reduce kernel (host/device neutral).
//
// Reduce mode of the reference kernel: `A1[I1] = op(A1[I1], f(A2[I2], ...))` with `initop` applied once per
// distinct output first (src/mapreduce.jl:314, :351-382).  Canonical dims are split into KEPT dims (output
// stride != 0) and REDUCED dims (output stride == 0).  A CTA owns one output tile and one split of the
// reduced index space; it walks its split tile by tile, every thread keeping EPT private accumulators whose
// output coordinate never changes (steps move only along reduced dims), loads coalesced in the input's
// fastest-stride order.  At the end the THREADS*EPT accumulators are combined through shared memory
// (warp-shuffle fold when >= 32 partials per output), and either applied to the output
// (`op(initop(out), partial)`, single split) or written to a partials buffer that `reduce_finalize`
// folds in a fixed order -- deterministic, no floating-point atomics.  This is the GPU analog of the
// reference's per-task partial slots + serial fold (src/mapreduce.jl:153-170), extended to partial
// reductions, which the reference never parallelises over reduced dims (:172-174).
#pragma once
#include "functors.hpp"

namespace sb {

template <int NIN> struct RedCta {
    const unsigned char *ptr[NIN]; // input k+1: base + kept-tile origin + thread offset (bytes)
    uint32_t out_tile, split;
    bool kept_full;
};

template <int NIN> SB_HD void red_cta_init(const ReduceParams &P, uint32_t bid, int t, RedCta<NIN> &c)
{
    uint32_t q, r;
    fast_divmod(P.outdiv, bid, q, r);
    c.out_tile = r;
    c.split = q;
    int64_t off[NIN];
#pragma unroll
    for (int k = 0; k < NIN; ++k) {
        int64_t g = 0;
        if (k + 1 < P.nops)
            for (int i = 0; i < P.order.n; ++i) g += (int64_t)field_of(P.order, i, t) * P.g_tstr[k + 1][i];
        off[k] = g;
    }
    uint32_t id = c.out_tile;
    bool full = true;
#pragma unroll
    for (int d = 0; d < MAXD; ++d) {
        if (d < P.nkept) {
            uint32_t qq, cd;
            fast_divmod(P.tdiv[d], id, qq, cd);
            id = qq;
            full = full && ((int32_t)cd < P.nfull[d]);
#pragma unroll
            for (int k = 0; k < NIN; ++k)
                if (k + 1 < P.nops) off[k] += (int64_t)cd * P.tstep[k + 1][d];
        }
    }
#pragma unroll
    for (int k = 0; k < NIN; ++k) c.ptr[k] = P.base[k + 1 < P.nops ? k + 1 : 0] + off[k];
    c.kept_full = full;
}

// decode reduction step -> byte offset of the step origin per input; returns "tile is interior"
template <int NIN> SB_HD bool red_step_init(const ReduceParams &P, const RedCta<NIN> &c, uint32_t step, int64_t (&soff)[NIN])
{
#pragma unroll
    for (int k = 0; k < NIN; ++k) soff[k] = 0;
    bool full = c.kept_full;
    uint32_t id = step;
#pragma unroll
    for (int d = 0; d < MAXD; ++d) {
        if (d >= P.nkept && d < P.ndim) {
            uint32_t q, cd;
            fast_divmod(P.tdiv[d], id, q, cd);
            id = q;
            full = full && ((int32_t)cd < P.nfull[d]);
#pragma unroll
            for (int k = 0; k < NIN; ++k)
                if (k + 1 < P.nops) soff[k] += (int64_t)cd * P.tstep[k + 1][d];
        }
    }
    return full;
}

// packed remaining extents (R | guard) for (out_tile, step): edge tiles only (slow path)
SB_HD uint32_t red_rem(const ReduceParams &P, uint32_t out_tile, uint32_t step)
{
    int64_t origin[MAXD];
    uint32_t id = out_tile;
    for (int d = 0; d < P.ndim; ++d) {
        if (d == P.nkept) id = step;
        uint32_t q, cd;
        fast_divmod(P.tdiv[d], id, q, cd);
        id = q;
        origin[d] = (int64_t)cd * P.tile_b[d];
    }
    int32_t rem[MAXTD];
    for (int i = 0; i < P.ntd; ++i) {
        const int64_t r = P.dims[P.tdim[i]] - origin[P.tdim[i]];
        rem[i] = r > 0x7fffffff ? 0x7fffffff : (int32_t)r;
    }
    return pack_rem(rem, P.ntd, P.cpos, P.cbits, P.guard);
}

// Accumulation phase of one CTA for thread t.  Leaves the EPT accumulators in shared memory.
template <class AT, int RC, int NIN, int EPT, bool UNIFORM>
SB_HD void red_accumulate(const ReduceParams &P, uint32_t bid, int t, AT *smem)
{
    RedCta<NIN> c;
    red_cta_init<NIN>(P, bid, t, c);
    AT acc[EPT];
#pragma unroll
    for (int j = 0; j < EPT; ++j) acc[j] = red_neutral<AT>(P.op);
    uint32_t c_toff = 0;
    for (int i = 0; i < P.order.n; ++i) c_toff += (uint32_t)field_of(P.order, i, t) * P.c_tstr[i];
    ElemFn<AT, RC> fn;
    const int64_t s0 = (int64_t)c.split * P.steps_per_split;
    int64_t s1 = s0 + P.steps_per_split;
    if (s1 > P.nrsteps) s1 = P.nrsteps;
    for (int64_t step = s0; step < s1; ++step) {
        int64_t soff[NIN];
        const bool full = red_step_init<NIN>(P, c, (uint32_t)step, soff);
        AT v[NIN][EPT];
        if (full) {
#pragma unroll
            for (int k = 0; k < NIN; ++k) {
                const unsigned char *b = c.ptr[k] + soff[k];
#pragma unroll
                for (int j = 0; j < EPT; ++j)
                    v[k][j] = (k + 1 < P.nops) ? load_elem<AT, UNIFORM>(b + P.g_joff[k + 1][j], P.dtype[k + 1], P.conj[k + 1]) : make<AT>(0.0, 0.0);
            }
#pragma unroll
            for (int j = 0; j < EPT; ++j) {
                AT a[NIN];
#pragma unroll
                for (int k = 0; k < NIN; ++k) a[k] = v[k][j];
                acc[j] = red_apply<AT>(P.op, acc[j], fn.template eval<NIN>(P.prog, a));
            }
        } else {
            const uint32_t rg = red_rem(P, c.out_tile, (uint32_t)step);
#pragma unroll
            for (int j = 0; j < EPT; ++j) {
                if (!packed_valid(rg, c_toff + P.c_joff[j], P.guard)) continue;
                AT a[NIN];
#pragma unroll
                for (int k = 0; k < NIN; ++k)
                    a[k] = (k + 1 < P.nops) ? load_elem<AT, UNIFORM>(c.ptr[k] + soff[k] + P.g_joff[k + 1][j], P.dtype[k + 1], P.conj[k + 1])
                                            : make<AT>(0.0, 0.0);
                acc[j] = red_apply<AT>(P.op, acc[j], fn.template eval<NIN>(P.prog, a));
            }
        }
    }
    int32_t s_toff = 0;
    for (int i = 0; i < P.order.n; ++i) s_toff += field_of(P.order, i, t) * P.s_tstr[i];
#pragma unroll
    for (int j = 0; j < EPT; ++j) smem[s_toff + P.s_joff[j]] = acc[j];
}

// partial of output o seen by one lane (warp-per-output layout [o][r]); folded across the warp afterwards
template <class AT> SB_HD AT red_lane_partial(const ReduceParams &P, const AT *smem, int o, int lane)
{
    AT p = red_neutral<AT>(P.op);
    const AT *row = smem + (int64_t)o * P.nred_tile;
    for (int r = lane; r < P.nred_tile; r += 32) p = red_apply<AT>(P.op, p, row[r]);
    return p;
}
// thread-per-output layout [r][o]
template <class AT> SB_HD AT red_thread_partial(const ReduceParams &P, const AT *smem, int o)
{
    AT p = red_neutral<AT>(P.op);
    for (int r = 0; r < P.nred_tile; ++r) p = red_apply<AT>(P.op, p, smem[(int64_t)r * P.nout_tile + o]);
    return p;
}

// output number o of a tile -> (valid, BYTE offset in the output)
SB_HD bool red_out_locate(const ReduceParams &P, uint32_t out_tile, int o, int64_t &off)
{
    uint32_t id = out_tile;
    int64_t origin[MAXD];
    off = 0;
    for (int d = 0; d < P.nkept; ++d) {
        uint32_t q, cd;
        fast_divmod(P.tdiv[d], id, q, cd);
        id = q;
        origin[d] = (int64_t)cd * P.tile_b[d];
        off += (int64_t)cd * P.tstep[0][d];
    }
    bool ok = true;
    for (int i = 0; i < P.kept_order.n; ++i) {
        const int d = P.tdim[P.kept_order.td[i]];
        const int64_t cd = field_of(P.kept_order, i, o);
        ok = ok && (origin[d] + cd < P.dims[d]);
        off += cd * P.strides[0][d] * dtype_size(P.dtype[0]);
    }
    return ok;
}

// lane-0 / owning-thread epilogue for output o with in-CTA partial p
template <class AT, bool UNIFORM> SB_HD void red_finish(const ReduceParams &P, uint32_t bid, int o, AT p)
{
    uint32_t split, out_tile;
    fast_divmod(P.outdiv, bid, split, out_tile);
    if (P.nsplit > 1) {
        reinterpret_cast<AT *>(P.scratch)[((int64_t)split * P.nouttiles + out_tile) * P.nout_tile + o] = p;
        return;
    }
    int64_t off;
    if (!red_out_locate(P, out_tile, o, off)) return;
    AT x = load_elem<AT, UNIFORM>(P.base[0] + off, P.dtype[0], P.conj[0]);
    x = init_apply<AT>(P.initop, P.init_re, P.init_im, x);
    store_elem<AT, UNIFORM>(P.base[0] + off, P.dtype[0], P.conj[0], red_apply<AT>(P.op, x, p));
}

// partials are written by other SMs in the same launch: read them through L2 (ld.global.cg), never a stale L1 line
template <class AT> SB_HD AT load_partial(const AT *p)
{
#if defined(__CUDA_ARCH__)
    AT v;
    constexpr int W = sizeof(AT) / 4;
    const unsigned int *src = reinterpret_cast<const unsigned int *>(p);
    unsigned int *dst = reinterpret_cast<unsigned int *>(&v);
#pragma unroll
    for (int i = 0; i < W; ++i) dst[i] = __ldcg(src + i);
    return v;
#else
    return *p;
#endif
}

// finalize: one WARP per output (out_tile, o).  Lane l folds partials of splits l, l+32, ... in split
// order; the 32 lane results are then combined by a shuffle butterfly (fixed order -> deterministic).
template <class AT> SB_HD AT red_finalize_lane(const ReduceParams &P, int64_t out_idx, int lane)
{
    const AT *sc = reinterpret_cast<const AT *>(P.scratch);
    const int64_t stride = P.nouttiles * (int64_t)P.nout_tile;
    AT p = red_neutral<AT>(P.op);
    for (int s = lane; s < P.nsplit; s += 256) { // eight independent loads in flight, folded in split order
        AT v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = (s + 32 * u < P.nsplit) ? load_partial(sc + (int64_t)(s + 32 * u) * stride + out_idx) : red_neutral<AT>(P.op);
#pragma unroll
        for (int u = 0; u < 8; ++u) p = red_apply<AT>(P.op, p, v[u]);
    }
    return p;
}
template <class AT, bool UNIFORM> SB_HD void red_finalize_store(const ReduceParams &P, int64_t out_idx, AT p)
{
    const int64_t out_tile = out_idx / P.nout_tile;
    const int o = (int)(out_idx - out_tile * P.nout_tile);
    int64_t off;
    if (!red_out_locate(P, (uint32_t)out_tile, o, off)) return;
    AT x = load_elem<AT, UNIFORM>(P.base[0] + off, P.dtype[0], P.conj[0]);
    x = init_apply<AT>(P.initop, P.init_re, P.init_im, x);
    store_elem<AT, UNIFORM>(P.base[0] + off, P.dtype[0], P.conj[0], red_apply<AT>(P.op, x, p));
}

} // namespace sb
