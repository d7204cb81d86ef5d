// kernels_map_group.cu -- instantiations of the GROUPED LSU map kernel (several problems of one plan in one launch,
// sb_mapreduce_batch): uniform single-input copies / scalings (`permutedims!`, `B .= c .* A'`) and two-input sums.
#include "kernels.cuh"
namespace sb {
const MapGroupEntry *find_map_group_kernel(const KernelKey &k)
{
    static const MapGroupEntry tab[] = {
        SB_MAP_GROUP_ENTRY(double, F64, RC_COPY, 1, 4),  SB_MAP_GROUP_ENTRY(double, F64, RC_COPY, 1, 8),
        SB_MAP_GROUP_ENTRY(double, F64, RC_SCALE, 1, 4), SB_MAP_GROUP_ENTRY(double, F64, RC_SCALE, 1, 8),
        SB_MAP_GROUP_ENTRY(double, F64, RC_ADD2, 2, 4),  SB_MAP_GROUP_ENTRY(double, F64, RC_ADD2, 2, 8),
        SB_MAP_GROUP_ENTRY(float, F32, RC_COPY, 1, 4),   SB_MAP_GROUP_ENTRY(float, F32, RC_COPY, 1, 8),
        SB_MAP_GROUP_ENTRY(float, F32, RC_SCALE, 1, 4),  SB_MAP_GROUP_ENTRY(float, F32, RC_SCALE, 1, 8),
        SB_MAP_GROUP_ENTRY(cx<float>, C32, RC_COPY, 1, 4), SB_MAP_GROUP_ENTRY(cx<float>, C32, RC_COPY, 1, 8),
    };
    if (!k.uniform) return nullptr;
    for (const MapGroupEntry &e : tab)
        if (e.key.ct == k.ct && e.key.recipe == k.recipe && e.key.nin == k.nin && e.key.ept == k.ept) return &e;
    return nullptr;
}
} // namespace sb
