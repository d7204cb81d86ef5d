// kernels_c64.cu -- instantiations for compute type cx<double> (one translation unit per type so that
// the build can run nvcc in parallel).  Which (recipe, NIN, EPT, UNIFORM) tuples exist is mirrored by
// planner.cpp: recipe_instantiated(), template_nin(), default_ept().
#include "kernels.cuh"
namespace sb {
const MapEntry *map_table_c64(int *n)
{
    static const MapEntry tab[] = {
        SB_MAP_ENTRY(cx<double>, C64, RC_COPY, 1, 4, 1),
        SB_MAP_ENTRY(cx<double>, C64, RC_SCALE, 1, 4, 1),
        SB_MAP_ENTRY(cx<double>, C64, RC_COPY, 1, 4, 0),
        SB_MAP_ENTRY(cx<double>, C64, RC_INTERP, 1, 4, 1),
        SB_MAP_ENTRY(cx<double>, C64, RC_INTERP, 1, 4, 0),
        SB_MAP_ENTRY(cx<double>, C64, RC_INTERP, 2, 4, 1),
        SB_MAP_ENTRY(cx<double>, C64, RC_INTERP, 2, 4, 0),
        SB_MAP_ENTRY(cx<double>, C64, RC_INTERP, 4, 4, 1),
        SB_MAP_ENTRY(cx<double>, C64, RC_INTERP, 4, 4, 0),
        SB_MAP_ENTRY(cx<double>, C64, RC_INTERP, 7, 4, 1),
        SB_MAP_ENTRY(cx<double>, C64, RC_INTERP, 7, 4, 0),
    };
    *n = (int)(sizeof(tab) / sizeof(tab[0]));
    return tab;
}
const ReduceEntry *reduce_table_c64(int *n)
{
    static const ReduceEntry tab[] = {
        SB_RED_ENTRY(cx<double>, C64, RC_COPY, 1, 4, 1),
        SB_RED_ENTRY(cx<double>, C64, RC_ABS2, 1, 4, 1),
        SB_RED_ENTRY(cx<double>, C64, RC_INTERP, 1, 4, 1),
        SB_RED_ENTRY(cx<double>, C64, RC_INTERP, 1, 4, 0),
        SB_RED_ENTRY(cx<double>, C64, RC_INTERP, 2, 4, 1),
        SB_RED_ENTRY(cx<double>, C64, RC_INTERP, 2, 4, 0),
        SB_RED_ENTRY(cx<double>, C64, RC_INTERP, 3, 4, 1),
        SB_RED_ENTRY(cx<double>, C64, RC_INTERP, 3, 4, 0),
    };
    *n = (int)(sizeof(tab) / sizeof(tab[0]));
    return tab;
}
} // namespace sb
