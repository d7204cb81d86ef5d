// reduce_stream.hpp -- consumer side of the streamed complete reduction (host/device neutral; see common.hpp
// "StreamParams").  A landed stage holds, per input k, `chunk_bytes` consecutive bytes of that input at offset
// k * chunk_bytes.  Thread t folds the 16-byte vectors t, t + THREADS, ... of the chunk into four private accumulators
// (independent dependency chains), element by element through the plan's element function:
//     acc = op(acc, f(A2[i], A3[i], ...))                                        (reference src/mapreduce.jl:314)
// The combination order is fixed by (grid, chunk size) alone: results are reproducible run to run.
#pragma once
#include "functors.hpp"

namespace sb {

constexpr int STREAM_ACC = 4;

template <class AT> struct alignas(16) StreamVec {
    static constexpr int V = 16 / (int)sizeof(AT);
    AT v[V];
};

template <class AT, int RC, int NIN, int OP>
SB_HD void stream_fold_vec(const StreamArgs &P, const ElemFn<AT, RC> &fn, const StreamVec<AT> (&x)[NIN], AT (&acc)[StreamVec<AT>::V])
{
#pragma unroll
    for (int u = 0; u < StreamVec<AT>::V; ++u) { // one accumulator per vector lane: in the interleaved mode the lanes are different outputs
        AT a[NIN];
#pragma unroll
        for (int k = 0; k < NIN; ++k) a[k] = x[k].v[u];
        acc[u] = red_apply<AT>(OP, acc[u], fn.template eval<NIN>(P.prog, a)); // OP is a compile-time constant: the switch folds away
    }
}

// one landed chunk: `stage` = base of the stage, `nv` = whole vectors in it (same for every input)
template <class AT, int RC, int NIN, int OP>
SB_HD void stream_chunk_op(const StreamArgs &P, const StreamParams &S, const unsigned char *stage, int nv, int t, AT (&acc)[STREAM_ACC][StreamVec<AT>::V])
{
    ElemFn<AT, RC> fn;
    const StreamVec<AT> *in[NIN];
#pragma unroll
    for (int k = 0; k < NIN; ++k) in[k] = reinterpret_cast<const StreamVec<AT> *>(stage + (size_t)(k < S.nin ? k : 0) * S.chunk_bytes);
    int i = t;
    for (; i + (STREAM_ACC - 1) * THREADS < nv; i += STREAM_ACC * THREADS) {
        StreamVec<AT> x[STREAM_ACC][NIN];
#pragma unroll
        for (int q = 0; q < STREAM_ACC; ++q)
#pragma unroll
            for (int k = 0; k < NIN; ++k) x[q][k] = in[k][i + q * THREADS];
#pragma unroll
        for (int q = 0; q < STREAM_ACC; ++q) stream_fold_vec<AT, RC, NIN, OP>(P, fn, x[q], acc[q]);
    }
#pragma unroll
    for (int q = 0; q < STREAM_ACC - 1; ++q) { // at most STREAM_ACC - 1 vectors are left for this thread
        if (i < nv) {
            StreamVec<AT> x[NIN];
#pragma unroll
            for (int k = 0; k < NIN; ++k) x[k] = in[k][i];
            stream_fold_vec<AT, RC, NIN, OP>(P, fn, x, acc[q]);
            i += THREADS;
        }
    }
}

// (the reduction operator is dispatched once per chunk, not once per element)
template <class AT, int RC, int NIN>
SB_HD void stream_chunk(const StreamArgs &P, const StreamParams &S, const unsigned char *stage, int nv, int t, AT (&acc)[STREAM_ACC][StreamVec<AT>::V])
{
    switch (P.op) {
    case OP_ADD: stream_chunk_op<AT, RC, NIN, OP_ADD>(P, S, stage, nv, t, acc); break;
    case OP_MUL: stream_chunk_op<AT, RC, NIN, OP_MUL>(P, S, stage, nv, t, acc); break;
    case OP_MIN: stream_chunk_op<AT, RC, NIN, OP_MIN>(P, S, stage, nv, t, acc); break;
    default: stream_chunk_op<AT, RC, NIN, OP_MAX>(P, S, stage, nv, t, acc); break;
    }
}

// byte offset of output number o (kept dim 0 fastest): in input k's parent (k >= 0) or in the output (k < 0)
SB_HD int64_t stream_out_offset(const StreamParams &S, int o, int k)
{
    int64_t off = 0;
#pragma unroll
    for (int d = 0; d < STREAM_MAXKD; ++d) {
        if (d < S.nkd) {
            const int64_t c = (int64_t)o % S.kdims[d];
            o = (int)((int64_t)o / S.kdims[d]);
            off += c * (k < 0 ? S.kout_bytes[d] : S.kin_bytes[k][d]);
        }
    }
    return off;
}

// the < 16-byte rest of output o's runs (elements [vec_bytes / sizeof(AT), nelem)), read straight from the operands
template <class AT, int RC, int NIN> SB_HD AT stream_rest(const StreamArgs &P, const StreamParams &S, int o, AT acc)
{
    ElemFn<AT, RC> fn;
    for (int64_t e = S.vec_bytes / (int64_t)sizeof(AT); e < S.nelem; ++e) {
        AT a[NIN];
#pragma unroll
        for (int k = 0; k < NIN; ++k) {
            const int kk = k < S.nin ? k : 0;
            a[k] = reinterpret_cast<const AT *>(P.base[kk + 1] + stream_out_offset(S, o, kk))[e];
        }
        acc = red_apply<AT>(P.op, acc, fn.template eval<NIN>(P.prog, a));
    }
    return acc;
}

// out[o] = op(initop(out[o]), total)   (reference src/mapreduce.jl:314 with the initop of :351-382)
template <class AT> SB_HD void stream_store(const StreamArgs &P, const StreamParams &S, int o, AT total)
{
    unsigned char *dst = P.base[0] + stream_out_offset(S, o, -1);
    AT x = load_elem<AT, true>(dst, P.dtype[0], P.conj[0]);
    x = init_apply<AT>(P.initop, P.init_re, P.init_im, x);
    store_elem<AT, true>(dst, P.dtype[0], P.conj[0], red_apply<AT>(P.op, x, total));
}

// the accumulators of a thread: dense mode -> one value (slots in order, then lanes in order)
template <class AT> SB_HD AT stream_thread_total(const StreamArgs &P, const AT (&acc)[STREAM_ACC][StreamVec<AT>::V])
{
    AT p = red_neutral<AT>(P.op);
#pragma unroll
    for (int u = 0; u < StreamVec<AT>::V; ++u) {
        AT pu = acc[0][u];
#pragma unroll
        for (int q = 1; q < STREAM_ACC; ++q) pu = red_apply<AT>(P.op, pu, acc[q][u]);
        p = u == 0 ? pu : red_apply<AT>(P.op, p, pu);
    }
    return p;
}
// interleaved mode -> one value per vector lane u (output number (t mod G) * V + u)
template <class AT> SB_HD AT stream_thread_lane_total(const StreamArgs &P, const AT (&acc)[STREAM_ACC][StreamVec<AT>::V], int u)
{
    AT pu = acc[0][u];
#pragma unroll
    for (int q = 1; q < STREAM_ACC; ++q) pu = red_apply<AT>(P.op, pu, acc[q][u]);
    return pu;
}

} // namespace sb
