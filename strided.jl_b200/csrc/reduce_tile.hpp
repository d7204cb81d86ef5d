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

struct RedCta {
    int64_t out_tile, split;
    int64_t kbase[MAXO]; // element offset of the output tile origin (kept dims), per operand
    int32_t rem[MAXTD];  // remaining extent per tile-dim slot; kept slots fixed, reduced slots per step
    bool kept_full;
};

SB_HD void red_cta_init(const ReduceParams &P, int64_t bid, RedCta &c)
{
    c.out_tile = bid % P.nouttiles;
    c.split = bid / P.nouttiles;
    for (int k = 0; k < MAXO; ++k) c.kbase[k] = 0;
    int64_t id = c.out_tile;
    bool full = true;
    for (int i = 0; i < MAXTD; ++i) c.rem[i] = 1;
    for (int d = 0; d < P.nkept; ++d) {
        const int64_t q = id / P.ntile[d];
        const int64_t cd = id - q * P.ntile[d];
        id = q;
        const int64_t origin = cd * P.tile_b[d];
        for (int k = 0; k < P.nops; ++k) c.kbase[k] += origin * P.strides[k][d];
        const int64_t r = P.dims[d] - origin;
        full = full && (r >= P.tile_b[d]);
        for (int i = 0; i < P.ntd; ++i)
            if (P.tdim[i] == d) c.rem[i] = r > 0x7fffffff ? 0x7fffffff : (int32_t)r;
    }
    c.kept_full = full;
}

// decode reduction step -> per-operand offset of the step origin; updates rem[] of reduced slots
SB_HD bool red_step_init(const ReduceParams &P, RedCta &c, int64_t step, int64_t (&sbase)[MAXO])
{
    for (int k = 0; k < MAXO; ++k) sbase[k] = c.kbase[k];
    bool full = c.kept_full;
    int64_t id = step;
    for (int d = P.nkept; d < P.ndim; ++d) {
        const int64_t q = id / P.ntile[d];
        const int64_t cd = id - q * P.ntile[d];
        id = q;
        const int64_t origin = cd * P.tile_b[d];
        for (int k = 1; k < P.nops; ++k) sbase[k] += origin * P.strides[k][d];
        const int64_t r = P.dims[d] - origin;
        full = full && (r >= P.tile_b[d]);
        for (int i = 0; i < P.ntd; ++i)
            if (P.tdim[i] == d) c.rem[i] = r > 0x7fffffff ? 0x7fffffff : (int32_t)r;
    }
    return full;
}

SB_HD bool red_valid(const ReduceParams &P, const RedCta &c, int t, int j)
{
    bool ok = true;
    for (int i = 0; i < P.order.n; ++i) {
        const int f = field_of(P.order, i, t) + (int)P.jfield[j][i];
        ok = ok && (f < c.rem[P.order.td[i]]);
    }
    return ok;
}

// Accumulation phase of one CTA for thread t.  Leaves the EPT accumulators in shared memory.
template <class AT, int RC, int NIN, int EPT, bool UNIFORM>
SB_HD void red_accumulate(const ReduceParams &P, int64_t bid, int t, AT *smem)
{
    RedCta c;
    red_cta_init(P, bid, c);
    int64_t g_toff[NIN];
#pragma unroll
    for (int k = 1; k <= NIN; ++k) {
        int64_t g = 0;
        if (k < P.nops)
            for (int i = 0; i < P.order.n; ++i) g += (int64_t)field_of(P.order, i, t) * P.g_tstr[k][i];
        g_toff[k - 1] = g;
    }
    AT acc[EPT];
#pragma unroll
    for (int j = 0; j < EPT; ++j) acc[j] = red_neutral<AT>(P.op);
    ElemFn<AT, RC> fn;
    const int64_t s0 = c.split * P.steps_per_split;
    int64_t s1 = s0 + P.steps_per_split;
    if (s1 > P.nrsteps) s1 = P.nrsteps;
    for (int64_t step = s0; step < s1; ++step) {
        int64_t sbase[MAXO];
        const bool full = red_step_init(P, c, step, sbase);
        AT v[NIN][EPT];
#pragma unroll
        for (int k = 1; k <= NIN; ++k) {
            if (k >= P.nops) {
#pragma unroll
                for (int j = 0; j < EPT; ++j) v[k - 1][j] = make<AT>(0.0, 0.0);
                continue;
            }
            const unsigned char *b = P.base[k];
            const int64_t o0 = sbase[k] + g_toff[k - 1];
            if (full) {
#pragma unroll
                for (int j = 0; j < EPT; ++j)
                    v[k - 1][j] = load_elem<AT, UNIFORM>(b, o0 + P.g_joff[k][j], P.dtype[k], P.conj[k]);
            } else {
#pragma unroll
                for (int j = 0; j < EPT; ++j) {
                    AT x = make<AT>(0.0, 0.0);
                    if (red_valid(P, c, t, j)) x = load_elem<AT, UNIFORM>(b, o0 + P.g_joff[k][j], P.dtype[k], P.conj[k]);
                    v[k - 1][j] = x;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < EPT; ++j) {
            AT a[NIN];
#pragma unroll
            for (int k = 0; k < NIN; ++k) a[k] = v[k][j];
            const AT r = fn.template eval<NIN>(P.prog, a);
            if (full || red_valid(P, c, t, j)) acc[j] = red_apply<AT>(P.op, acc[j], r);
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

// output number o of a tile -> (valid, element offset in the output)
SB_HD bool red_out_locate(const ReduceParams &P, int64_t out_tile, int o, int64_t &off)
{
    // tile origin over kept dims
    int64_t id = out_tile;
    int64_t origin[MAXD];
    off = 0;
    for (int d = 0; d < P.nkept; ++d) {
        const int64_t q = id / P.ntile[d];
        origin[d] = (id - q * P.ntile[d]) * P.tile_b[d];
        id = q;
        off += origin[d] * P.strides[0][d];
    }
    bool ok = true;
    for (int i = 0; i < P.kept_order.n; ++i) {
        const int d = P.tdim[P.kept_order.td[i]];
        const int64_t cd = field_of(P.kept_order, i, o);
        ok = ok && (origin[d] + cd < P.dims[d]);
        off += cd * P.strides[0][d];
    }
    return ok;
}

// lane-0 / owning-thread epilogue for output o with in-CTA partial p
template <class AT, bool UNIFORM> SB_HD void red_finish(const ReduceParams &P, int64_t bid, int o, AT p)
{
    const int64_t out_tile = bid % P.nouttiles, split = bid / P.nouttiles;
    if (P.nsplit > 1) {
        reinterpret_cast<AT *>(P.scratch)[(split * P.nouttiles + out_tile) * P.nout_tile + o] = p;
        return;
    }
    int64_t off;
    if (!red_out_locate(P, out_tile, o, off)) return;
    AT x = load_elem<AT, UNIFORM>(P.base[0], off, P.dtype[0], P.conj[0]);
    x = init_apply<AT>(P.initop, P.init_re, P.init_im, x);
    store_elem<AT, UNIFORM>(P.base[0], off, P.dtype[0], P.conj[0], red_apply<AT>(P.op, x, p));
}

// finalize kernel body: one thread per (out_tile, o); folds the nsplit partials in split order
template <class AT, bool UNIFORM> SB_HD void red_finalize(const ReduceParams &P, int64_t idx)
{
    const int64_t out_tile = idx / P.nout_tile;
    const int o = (int)(idx - out_tile * P.nout_tile);
    if (out_tile >= P.nouttiles) return;
    int64_t off;
    if (!red_out_locate(P, out_tile, o, off)) return;
    const AT *sc = reinterpret_cast<const AT *>(P.scratch);
    AT p = sc[(0 * P.nouttiles + out_tile) * P.nout_tile + o];
    for (int s = 1; s < P.nsplit; ++s) p = red_apply<AT>(P.op, p, sc[((int64_t)s * P.nouttiles + out_tile) * P.nout_tile + o]);
    AT x = load_elem<AT, UNIFORM>(P.base[0], off, P.dtype[0], P.conj[0]);
    x = init_apply<AT>(P.initop, P.init_re, P.init_im, x);
    store_elem<AT, UNIFORM>(P.base[0], off, P.dtype[0], P.conj[0], red_apply<AT>(P.op, x, p));
}

} // namespace sb
