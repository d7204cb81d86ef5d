// kernels_f32.cu -- instantiations for compute type float (one translation unit per type so that
// the build can run nvcc in parallel).  Which (recipe, NIN, EPT, UNIFORM) tuples exist is mirrored by
// planner.cpp: recipe_instantiated(), template_nin(), default_ept().
#include "kernels.cuh"
namespace sb {
const MapEntry *map_table_f32(int *n)
{
    static const MapEntry tab[] = {
        SB_MAP_ENTRY(float, F32, RC_COPY, 1, 8, 1),
        SB_MAP_ENTRY(float, F32, RC_SCALE, 1, 8, 1),
        SB_MAP_ENTRY(float, F32, RC_ABS2, 1, 8, 1),
        SB_MAP_ENTRY(float, F32, RC_ADD2, 2, 8, 1),
        SB_MAP_ENTRY(float, F32, RC_ADD2_DIV, 2, 8, 1),
        SB_MAP_ENTRY(float, F32, RC_ADD2_MUL, 2, 8, 1),
        SB_MAP_ENTRY(float, F32, RC_AXPY, 2, 8, 1),
        SB_MAP_ENTRY(float, F32, RC_AXPBY, 2, 8, 1),
        SB_MAP_ENTRY(float, F32, RC_SUM3, 3, 8, 1),
        SB_MAP_ENTRY(float, F32, RC_SUM4, 4, 8, 1),
        SB_MAP_ENTRY(float, F32, RC_COPY, 1, 8, 0),
        SB_MAP_ENTRY(float, F32, RC_INTERP, 1, 8, 1),
        SB_MAP_ENTRY(float, F32, RC_INTERP, 1, 8, 0),
        SB_MAP_ENTRY(float, F32, RC_INTERP, 2, 8, 1),
        SB_MAP_ENTRY(float, F32, RC_INTERP, 2, 8, 0),
        SB_MAP_ENTRY(float, F32, RC_INTERP, 4, 8, 1),
        SB_MAP_ENTRY(float, F32, RC_INTERP, 4, 8, 0),
        SB_MAP_ENTRY(float, F32, RC_INTERP, 7, 8, 1),
        SB_MAP_ENTRY(float, F32, RC_INTERP, 7, 8, 0),
        SB_MAP_ENTRY(float, F32, RC_COPY, 1, 16, 1),
        SB_MAP_ENTRY(float, F32, RC_SCALE, 1, 16, 1),
        SB_MAP_ENTRY(float, F32, RC_ABS2, 1, 16, 1),
        SB_MAP_ENTRY(float, F32, RC_ADD2, 2, 16, 1),
        SB_MAP_ENTRY(float, F32, RC_ADD2_DIV, 2, 16, 1),
        SB_MAP_ENTRY(float, F32, RC_ADD2_MUL, 2, 16, 1),
        SB_MAP_ENTRY(float, F32, RC_AXPY, 2, 16, 1),
        SB_MAP_ENTRY(float, F32, RC_AXPBY, 2, 16, 1),
        SB_MAP_ENTRY(float, F32, RC_SUM3, 3, 16, 1),
        SB_MAP_ENTRY(float, F32, RC_SUM4, 4, 16, 1),
        SB_MAP_ENTRY(float, F32, RC_COPY, 1, 16, 0),
        SB_MAP_ENTRY(float, F32, RC_INTERP, 1, 16, 1),
        SB_MAP_ENTRY(float, F32, RC_INTERP, 1, 16, 0),
        SB_MAP_ENTRY(float, F32, RC_INTERP, 2, 16, 1),
        SB_MAP_ENTRY(float, F32, RC_INTERP, 2, 16, 0),
        SB_MAP_ENTRY(float, F32, RC_INTERP, 4, 16, 1),
        SB_MAP_ENTRY(float, F32, RC_INTERP, 4, 16, 0),
        SB_MAP_ENTRY(float, F32, RC_INTERP, 7, 16, 1),
        SB_MAP_ENTRY(float, F32, RC_INTERP, 7, 16, 0),
    };
    *n = (int)(sizeof(tab) / sizeof(tab[0]));
    return tab;
}
const ReduceEntry *reduce_table_f32(int *n)
{
    static const ReduceEntry tab[] = {
        SB_RED_ENTRY(float, F32, RC_COPY, 1, 8, 1),
        SB_RED_ENTRY(float, F32, RC_ABS2, 1, 8, 1),
        SB_RED_ENTRY(float, F32, RC_INTERP, 1, 8, 1),
        SB_RED_ENTRY(float, F32, RC_INTERP, 1, 8, 0),
        SB_RED_ENTRY(float, F32, RC_INTERP, 2, 8, 1),
        SB_RED_ENTRY(float, F32, RC_INTERP, 2, 8, 0),
        SB_RED_ENTRY(float, F32, RC_INTERP, 3, 8, 1),
        SB_RED_ENTRY(float, F32, RC_INTERP, 3, 8, 0),
    };
    *n = (int)(sizeof(tab) / sizeof(tab[0]));
    return tab;
}
} // namespace sb
