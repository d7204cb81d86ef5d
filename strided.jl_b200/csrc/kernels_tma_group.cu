// kernels_tma_group.cu -- instantiations of the GROUPED TMA map kernel (several problems of one plan in one launch,
// sb_mapreduce_batch): the entries of kernels_tma.cu with at most TMA_GROUP_MAXIN inputs.
#include "tma_kernel.cuh"
namespace sb {
const TmaGroupEntry *find_tma_group_kernel(const KernelKey &k)
{
    static const TmaGroupEntry tab[] = {
        SB_TMA_GROUP_ENTRY(double, F64, RC_COPY, 1, 8),     SB_TMA_GROUP_ENTRY(double, F64, RC_SCALE, 1, 8),
        SB_TMA_GROUP_ENTRY(double, F64, RC_ADD2, 2, 8),     SB_TMA_GROUP_ENTRY(double, F64, RC_ADD2_MUL, 2, 8),
        SB_TMA_GROUP_ENTRY(double, F64, RC_ADD2_DIV, 2, 8), SB_TMA_GROUP_ENTRY(double, F64, RC_AXPY, 2, 8),
        SB_TMA_GROUP_ENTRY(double, F64, RC_AXPBY, 2, 8),    SB_TMA_GROUP_ENTRY(double, F64, RC_INTERP, 1, 8),
        SB_TMA_GROUP_ENTRY(double, F64, RC_INTERP, 2, 8),
        SB_TMA_GROUP_ENTRY(float, F32, RC_COPY, 1, 8),      SB_TMA_GROUP_ENTRY(float, F32, RC_SCALE, 1, 8),
        SB_TMA_GROUP_ENTRY(float, F32, RC_ADD2, 2, 8),      SB_TMA_GROUP_ENTRY(float, F32, RC_ADD2_MUL, 2, 8),
        SB_TMA_GROUP_ENTRY(float, F32, RC_ADD2_DIV, 2, 8),  SB_TMA_GROUP_ENTRY(float, F32, RC_AXPY, 2, 8),
        SB_TMA_GROUP_ENTRY(float, F32, RC_AXPBY, 2, 8),     SB_TMA_GROUP_ENTRY(float, F32, RC_INTERP, 1, 8),
        SB_TMA_GROUP_ENTRY(float, F32, RC_INTERP, 2, 8),
        SB_TMA_GROUP_ENTRY(cx<float>, C32, RC_COPY, 1, 8),  SB_TMA_GROUP_ENTRY(cx<float>, C32, RC_INTERP, 1, 8),
        SB_TMA_GROUP_ENTRY(cx<float>, C32, RC_INTERP, 2, 8),
    };
    for (const TmaGroupEntry &e : tab)
        if (e.key.ct == k.ct && e.key.recipe == k.recipe && e.key.nin == k.nin && e.key.ept == k.ept) return &e;
    return nullptr;
}
} // namespace sb
