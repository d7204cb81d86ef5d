// orbit_tile.hpp -- consumer side of the alias-fused ("orbit") map kernel (host/device neutral).
//
// A landed stage holds the parent blocks of one orbit (dense TMA boxes, parent dim order).  For output tile m the
// element u = t + 256*j of thread t sits at tile coordinates x = M u (GF(2)-linear, chosen by the planner); input k
// reads it from block slot[m][k] at byte offset  T_k(t) XOR J_k[j]  and the result goes to the staging buffer (dense
// box in OUTPUT dim order, later written by one TMA store) at  T_0(t) XOR J_0[j].  Replaces the blocked inner loop
// nest of `_mapreduce_kernel!` (reference src/mapreduce.jl:311, :339-349) for aliased permuted views.
//
// All buffers are addressed as offsets from the ring base; block offsets are multiples of tile_bytes (a power of
// two) and T XOR J < tile_bytes, so  block + (T ^ J) == (block | T) ^ J : one logic op per access.
#pragma once
#include "functors.hpp"

namespace sb {

template <int NIN> struct OrbitThread {
    uint32_t T[NIN + 1];
};

template <int NIN> SB_HD void orbit_thread_init(const OrbitParams &O, int t, OrbitThread<NIN> &th)
{
#pragma unroll
    for (int k = 0; k <= NIN; ++k) {
        uint32_t a = 0;
#pragma unroll
        for (int i = 0; i < ORB_MAXLOGT; ++i)
            if (i < O.log_threads && ((t >> i) & 1)) a ^= O.tcol[k][i];
        th.T[k] = a;
    }
}

// One output tile.  `ring`: base of the input ring; `stage_off`: byte offset of the landed stage; `slots`: slot[m][0..3]
// packed little-endian; `sbuf_off`: byte offset of this tile's staging buffer (all offsets multiples of tile_bytes).
template <class CT, int RC, int NIN, int EPT>
SB_HD void orbit_compute(const OrbitParams &O, const OrbitThread<NIN> &th, unsigned char *ring, uint32_t stage_off, uint32_t slots,
                         uint32_t sbuf_off)
{
#ifndef SB_ORBIT_CH
#define SB_ORBIT_CH 8
#endif
    constexpr int CH = EPT < SB_ORBIT_CH ? EPT : SB_ORBIT_CH; // elements in flight per thread: NIN * CH shared-memory loads
    uint32_t bt[NIN + 1];
    bt[0] = sbuf_off | th.T[0];
#pragma unroll
    for (int k = 0; k < NIN; ++k) bt[k + 1] = (stage_off + ((slots >> (8 * k)) & 0xffu) * (uint32_t)O.tile_bytes) | th.T[k + 1];
    ElemFn<CT, RC> fn;
#pragma unroll
    for (int j0 = 0; j0 < EPT; j0 += CH) {
        CT v[NIN][CH];
#pragma unroll
        for (int k = 0; k < NIN; ++k)
#pragma unroll
            for (int u = 0; u < CH; ++u) v[k][u] = *reinterpret_cast<const CT *>(ring + (bt[k + 1] ^ O.jtab[k + 1][j0 + u]));
#pragma unroll
        for (int u = 0; u < CH; ++u) {
            CT a[NIN];
#pragma unroll
            for (int k = 0; k < NIN; ++k) a[k] = v[k][u];
            *reinterpret_cast<CT *>(ring + (bt[0] ^ O.jtab[0][j0 + u])) = fn.template eval<NIN>(O.prog, a);
        }
    }
}

// direct-store pass of one staged tile (after the tile barrier): thread t copies its 16-byte groups to the output
struct OrbitVec16 {
    alignas(16) uint32_t w[4];
};
SB_HD int64_t orbit_store_toff(const OrbitParams &O, int t)
{
    int64_t a = 0;
#pragma unroll
    for (int i = 0; i < ORB_MAXLOGT; ++i)
        if (i < O.log_threads && ((t >> i) & 1)) a += O.st_tcol[i];
    return a;
}
SB_HD void orbit_store_direct(const OrbitParams &O, int t, int64_t st_t, const unsigned char *ring, uint32_t sbuf_off, unsigned char *out_tile)
{
    const unsigned char *src = ring + sbuf_off + 16u * (uint32_t)t;
    unsigned char *dst = out_tile + st_t;
#pragma unroll 4
    for (int r = 0; r < O.st_groups; ++r) {
#if defined(__CUDA_ARCH__)
        const OrbitVec16 v = *reinterpret_cast<const OrbitVec16 *>(src + (size_t)r * (16u << O.log_threads));
        __stcs(reinterpret_cast<uint4 *>(dst + O.st_roff[r]), make_uint4(v.w[0], v.w[1], v.w[2], v.w[3])); // streaming: never re-read
#else
        for (int b = 0; b < 16; ++b) dst[O.st_roff[r] + b] = src[(size_t)r * (16u << O.log_threads) + b]; // (host emulation: no alignment assumed)
#endif
    }
}

} // namespace sb
