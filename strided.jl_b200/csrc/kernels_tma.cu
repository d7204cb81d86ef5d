// kernels_tma.cu -- instantiations of the TMA-pipelined map kernel (uniform dtypes, <= 4 inputs).
// Mirrored by planner.cpp: tma_instantiated().
#include "tma_kernel.cuh"
namespace sb {
const TmaEntry *tma_table(int *n)
{
    static const TmaEntry tab[] = {
        SB_TMA_ENTRY(double, F64, RC_COPY, 1, 8),     SB_TMA_ENTRY(double, F64, RC_SCALE, 1, 8),
        SB_TMA_ENTRY(double, F64, RC_ADD2, 2, 8),     SB_TMA_ENTRY(double, F64, RC_ADD2_MUL, 2, 8),
        SB_TMA_ENTRY(double, F64, RC_ADD2_DIV, 2, 8), SB_TMA_ENTRY(double, F64, RC_AXPY, 2, 8),
        SB_TMA_ENTRY(double, F64, RC_AXPBY, 2, 8),    SB_TMA_ENTRY(double, F64, RC_INTERP, 1, 8),
        SB_TMA_ENTRY(double, F64, RC_INTERP, 2, 8),   SB_TMA_ENTRY(double, F64, RC_INTERP, 4, 8),
        SB_TMA_ENTRY(float, F32, RC_COPY, 1, 8),      SB_TMA_ENTRY(float, F32, RC_SCALE, 1, 8),
        SB_TMA_ENTRY(float, F32, RC_ADD2, 2, 8),      SB_TMA_ENTRY(float, F32, RC_ADD2_MUL, 2, 8),
        SB_TMA_ENTRY(float, F32, RC_ADD2_DIV, 2, 8),  SB_TMA_ENTRY(float, F32, RC_AXPY, 2, 8),
        SB_TMA_ENTRY(float, F32, RC_AXPBY, 2, 8),     SB_TMA_ENTRY(float, F32, RC_INTERP, 1, 8),
        SB_TMA_ENTRY(float, F32, RC_INTERP, 2, 8),    SB_TMA_ENTRY(float, F32, RC_INTERP, 4, 8),
        SB_TMA_ENTRY(cx<float>, C32, RC_COPY, 1, 8),  SB_TMA_ENTRY(cx<float>, C32, RC_INTERP, 1, 8),
        SB_TMA_ENTRY(cx<float>, C32, RC_INTERP, 2, 8),
    };
    *n = (int)(sizeof(tab) / sizeof(tab[0]));
    return tab;
}
const TmaEntry *find_tma_kernel(const KernelKey &k)
{
    int n = 0;
    const TmaEntry *t = tma_table(&n);
    for (int i = 0; i < n; ++i) {
        const KernelKey &e = t[i].key;
        if (e.ct == k.ct && e.recipe == k.recipe && e.nin == k.nin && e.ept == k.ept) return &t[i];
    }
    return nullptr;
}
} // namespace sb
