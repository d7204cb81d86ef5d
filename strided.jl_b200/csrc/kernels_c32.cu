// kernels_c32.cu -- instantiations for compute type cx<float> (one translation unit per type so that
// the build can run nvcc in parallel).  Which (recipe, NIN, EPT, UNIFORM) tuples exist is mirrored by
// planner.cpp: recipe_instantiated(), template_nin(), default_ept().
#include "kernels.cuh"
namespace sb {
const MapEntry *map_table_c32(int *n)
{
    static const MapEntry tab[] = {
        SB_MAP_ENTRY(cx<float>, C32, RC_COPY, 1, 8, 1),
        SB_MAP_ENTRY(cx<float>, C32, RC_SCALE, 1, 8, 1),
        SB_MAP_ENTRY(cx<float>, C32, RC_COPY, 1, 8, 0),
        SB_MAP_ENTRY(cx<float>, C32, RC_INTERP, 1, 8, 1),
        SB_MAP_ENTRY(cx<float>, C32, RC_INTERP, 1, 8, 0),
        SB_MAP_ENTRY(cx<float>, C32, RC_INTERP, 2, 8, 1),
        SB_MAP_ENTRY(cx<float>, C32, RC_INTERP, 2, 8, 0),
        SB_MAP_ENTRY(cx<float>, C32, RC_INTERP, 4, 8, 1),
        SB_MAP_ENTRY(cx<float>, C32, RC_INTERP, 4, 8, 0),
        SB_MAP_ENTRY(cx<float>, C32, RC_INTERP, 7, 8, 1),
        SB_MAP_ENTRY(cx<float>, C32, RC_INTERP, 7, 8, 0),
    };
    *n = (int)(sizeof(tab) / sizeof(tab[0]));
    return tab;
}
const ReduceEntry *reduce_table_c32(int *n)
{
    static const ReduceEntry tab[] = {
        SB_RED_ENTRY(cx<float>, C32, RC_COPY, 1, 8, 1),
        SB_RED_ENTRY(cx<float>, C32, RC_ABS2, 1, 8, 1),
        SB_RED_ENTRY(cx<float>, C32, RC_INTERP, 1, 8, 1),
        SB_RED_ENTRY(cx<float>, C32, RC_INTERP, 1, 8, 0),
        SB_RED_ENTRY(cx<float>, C32, RC_INTERP, 2, 8, 1),
        SB_RED_ENTRY(cx<float>, C32, RC_INTERP, 2, 8, 0),
        SB_RED_ENTRY(cx<float>, C32, RC_INTERP, 3, 8, 1),
        SB_RED_ENTRY(cx<float>, C32, RC_INTERP, 3, 8, 0),
    };
    *n = (int)(sizeof(tab) / sizeof(tab[0]));
    return tab;
}
} // namespace sb
