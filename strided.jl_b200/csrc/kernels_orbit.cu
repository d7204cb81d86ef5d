// kernels_orbit.cu -- instantiations of the alias-fused ("orbit") map kernel.  Mirrored by planner.cpp: orbit_instantiated().
#include "orbit_kernel.cuh"
namespace sb {
#define SB_ORBIT_EPTS(CT, DT, RC, NIN)                                                                               \
    SB_ORBIT_ENTRY(CT, DT, RC, NIN, 4, 8), SB_ORBIT_ENTRY(CT, DT, RC, NIN, 8, 8), SB_ORBIT_ENTRY(CT, DT, RC, NIN, 16, 8),  \
        SB_ORBIT_ENTRY(CT, DT, RC, NIN, 2, 9), SB_ORBIT_ENTRY(CT, DT, RC, NIN, 4, 9), SB_ORBIT_ENTRY(CT, DT, RC, NIN, 8, 9)
#define SB_ORBIT_TYPE(CT, DT)                                                                                        \
    SB_ORBIT_EPTS(CT, DT, RC_ADD2, 2), SB_ORBIT_EPTS(CT, DT, RC_ADD2_MUL, 2), SB_ORBIT_EPTS(CT, DT, RC_ADD2_DIV, 2),   \
        SB_ORBIT_EPTS(CT, DT, RC_AXPY, 2), SB_ORBIT_EPTS(CT, DT, RC_AXPBY, 2), SB_ORBIT_EPTS(CT, DT, RC_SUM3, 3),      \
        SB_ORBIT_EPTS(CT, DT, RC_SUM4, 4), SB_ORBIT_EPTS(CT, DT, RC_INTERP, 2), SB_ORBIT_EPTS(CT, DT, RC_INTERP, 3),          \
        SB_ORBIT_EPTS(CT, DT, RC_INTERP, 4)
const OrbitEntry *orbit_table(int *n)
{
    static const OrbitEntry tab[] = {SB_ORBIT_TYPE(float, F32), SB_ORBIT_TYPE(double, F64)};
    *n = (int)(sizeof(tab) / sizeof(tab[0]));
    return tab;
}
const OrbitEntry *find_orbit_kernel(const KernelKey &k, int logt)
{
    int n = 0;
    const OrbitEntry *t = orbit_table(&n);
    for (int i = 0; i < n; ++i) {
        const KernelKey &e = t[i].key;
        if (e.ct == k.ct && e.recipe == k.recipe && e.nin == k.nin && e.ept == k.ept && t[i].logt == logt) return &t[i];
    }
    return nullptr;
}
} // namespace sb
