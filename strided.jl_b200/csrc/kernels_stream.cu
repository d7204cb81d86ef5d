// kernels_stream.cu -- instantiations of the streamed complete reduction (uniform dtypes, <= 3 inputs).
// Mirrored by planner.cpp: stream_instantiated().
#include "stream_kernel.cuh"
namespace sb {
const StreamEntry *stream_table(int *n)
{
    static const StreamEntry tab[] = {
        SB_STREAM_ENTRY(double, F64, RC_COPY, 1),      SB_STREAM_ENTRY(double, F64, RC_ABS2, 1),
        SB_STREAM_ENTRY(double, F64, RC_INTERP, 1),    SB_STREAM_ENTRY(double, F64, RC_INTERP, 2),
        SB_STREAM_ENTRY(double, F64, RC_INTERP, 3),
        SB_STREAM_ENTRY(double, F64, RC_S_ABS, 1),     SB_STREAM_ENTRY(double, F64, RC_S_MUL2, 2),
        SB_STREAM_ENTRY(float, F32, RC_S_ABS, 1),      SB_STREAM_ENTRY(float, F32, RC_S_MUL2, 2),
        SB_STREAM_ENTRY(float, F32, RC_COPY, 1),       SB_STREAM_ENTRY(float, F32, RC_ABS2, 1),
        SB_STREAM_ENTRY(float, F32, RC_INTERP, 1),     SB_STREAM_ENTRY(float, F32, RC_INTERP, 2),
        SB_STREAM_ENTRY(float, F32, RC_INTERP, 3),
        SB_STREAM_ENTRY(cx<float>, C32, RC_COPY, 1),   SB_STREAM_ENTRY(cx<float>, C32, RC_INTERP, 1),
        SB_STREAM_ENTRY(cx<float>, C32, RC_INTERP, 2), SB_STREAM_ENTRY(cx<float>, C32, RC_INTERP, 3),
        SB_STREAM_ENTRY(cx<double>, C64, RC_COPY, 1),  SB_STREAM_ENTRY(cx<double>, C64, RC_INTERP, 1),
        SB_STREAM_ENTRY(cx<double>, C64, RC_INTERP, 2), SB_STREAM_ENTRY(cx<double>, C64, RC_INTERP, 3),
    };
    *n = (int)(sizeof(tab) / sizeof(tab[0]));
    return tab;
}
const StreamEntry *find_stream_kernel(const KernelKey &k)
{
    int n = 0;
    const StreamEntry *t = stream_table(&n);
    for (int i = 0; i < n; ++i) {
        const KernelKey &e = t[i].key;
        if (e.ct == k.ct && e.recipe == k.recipe && e.nin == k.nin) return &t[i];
    }
    return nullptr;
}
} // namespace sb
