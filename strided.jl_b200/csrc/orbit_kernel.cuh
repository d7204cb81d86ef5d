// orbit_kernel.cuh -- sm_100a kernel: alias-fused strided map (see common.hpp "orbit" and orbit_tile.hpp).
//
//   producer warp : per work item (orbit): waits for a free stage, copies the item record (TMA coordinates, slot table;
//                   fetched two items ahead) into shared memory, arms the stage's full barrier with ntile * tile_bytes and
//                   lets lane s issue the cp.async.bulk.tensor of parent block s                (SASS: UTMALDG, SYNCS)
//   consumer warps: wait on the full barrier; for each output tile of the orbit compute T*EPT elements from shared memory
//                   into one of K (2..4) staging buffers (fence.proxy.async), meet on a named barrier, and thread 0 issues
//                   the TMA store of the tile (cp.async.bulk.tensor ... bulk_group; SASS: UTMASTG) -- up to K-1 stores are
//                   in flight while the next tiles are computed; every consumer thread releases the stage itself.
//                   (opt-in: 128-bit st.global from the staging buffer instead of the TMA store, two buffers)
// Every kernel executes griddepcontrol.launch_dependents / griddepcontrol.wait (PDL, common.hpp).
// Out-of-bounds parts of edge blocks are zero-filled on load and clipped on store by the TMA unit: no masks anywhere.
#pragma once
#include "tma_kernel.cuh"
#include "orbit_tile.hpp"

namespace sb {

template <int RANK> __device__ __forceinline__ void tma_store(const CUtensorMap *map, uint32_t src, const int32_t *c);
template <> __device__ __forceinline__ void tma_store<1>(const CUtensorMap *map, uint32_t src, const int32_t *c)
{
    asm volatile("cp.async.bulk.tensor.1d.global.shared::cta.bulk_group [%0, {%2}], [%1];" ::"l"(map), "r"(src), "r"(c[0]) : "memory");
}
template <> __device__ __forceinline__ void tma_store<2>(const CUtensorMap *map, uint32_t src, const int32_t *c)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c[0]), "r"(c[1])
                 : "memory");
}
template <> __device__ __forceinline__ void tma_store<3>(const CUtensorMap *map, uint32_t src, const int32_t *c)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(src), "r"(c[0]),
                 "r"(c[1]), "r"(c[2])
                 : "memory");
}
template <> __device__ __forceinline__ void tma_store<4>(const CUtensorMap *map, uint32_t src, const int32_t *c)
{
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map), "r"(src), "r"(c[0]),
                 "r"(c[1]), "r"(c[2]), "r"(c[3])
                 : "memory");
}
template <> __device__ __forceinline__ void tma_store<5>(const CUtensorMap *map, uint32_t src, const int32_t *c)
{
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(map), "r"(src),
                 "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4])
                 : "memory");
}
__device__ __forceinline__ void tma_load_rank(int rank, uint32_t dst, const CUtensorMap *map, uint32_t bar, const int32_t *c)
{
    switch (rank) {
    case 1: tma_load<1>(dst, map, bar, c); break;
    case 2: tma_load<2>(dst, map, bar, c); break;
    case 3: tma_load<3>(dst, map, bar, c); break;
    case 4: tma_load<4>(dst, map, bar, c); break;
    default: tma_load<5>(dst, map, bar, c); break;
    }
}
__device__ __forceinline__ void tma_store_rank(int rank, const CUtensorMap *map, uint32_t src, const int32_t *c)
{
    switch (rank) {
    case 1: tma_store<1>(map, src, c); break;
    case 2: tma_store<2>(map, src, c); break;
    case 3: tma_store<3>(map, src, c); break;
    case 4: tma_store<4>(map, src, c); break;
    default: tma_store<5>(map, src, c); break;
    }
}

constexpr int ORB_MAXSTAGE = 8;

template <class CT, int RC, int NIN, int EPT, int LOGT>
__global__ void __launch_bounds__((1 << LOGT) + 32, LOGT == 8 ? 2 : 1)
map_orbit_kernel(const __grid_constant__ OrbitParams O, const __grid_constant__ CUtensorMap min, const __grid_constant__ CUtensorMap mout)
{
    extern __shared__ unsigned char sb_orbit_smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[ORB_MAXSTAGE];
    __shared__ __align__(8) uint64_t empty_bar[ORB_MAXSTAGE];
    __shared__ __align__(16) uint32_t item_smem[ORB_MAXSTAGE][64];
    // ring (nstage stages of gmax blocks) followed by the staging buffers; TMA needs 128-byte aligned boxes
    unsigned char *ring = sb_orbit_smem_raw + ((0u - smem_u32(sb_orbit_smem_raw)) & 127u); // (offset form keeps the address space known: LDS/STS)
    const uint32_t ring_u32 = smem_u32(ring);
    constexpr int NT = 1 << LOGT; // consumer threads
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int S = O.nstage;
    pdl_launch_dependents();
    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 32); // every producer lane arrives: each releases its own item words
            mbar_init(smem_u32(&empty_bar[s]), NT);  // every consumer thread arrives after its last read of the stage
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t nitems = (uint32_t)O.nitems;
    const uint32_t grid = gridDim.x;
    constexpr int IW = (int)(sizeof(OrbitItem) / 4); // words per work item
    static_assert(sizeof(OrbitItem) % 4 == 0 && IW <= 64, "OrbitItem is copied as <= 2 words per producer lane");
    if (warp == NT / 32) {
        // ---------------- producer warp ----------------
        // The work item (coordinates, slots) travels with the stage: the producer copies it into shared memory before
        // arming the full barrier, so that no consumer ever waits on a dependent global load (measured: the item loads
        // on the consumers' critical path cost ~1 us per item, profiles/r01_v9_orbit_first.txt).
        int stage = 0;
        uint32_t parity = 1; // a fresh barrier passes a wait on parity 1: every stage starts out empty
        // item words are fetched TWO items ahead (registers a*/b*): one L2 round trip per item would otherwise bound the loop
        uint32_t a0 = 0, a1 = 0, b0 = 0, b1 = 0;
        auto fetch = [&](uint32_t p, uint32_t &x0, uint32_t &x1) {
            if (p < nitems) {
                const uint32_t *src = reinterpret_cast<const uint32_t *>(O.items + p);
                x0 = src[lane];
                if (lane + 32 < IW) x1 = src[lane + 32];
            }
        };
        fetch(blockIdx.x, a0, a1); // (the item table is written once, at plan creation: safe to read before pdl_wait)
        fetch(blockIdx.x + grid, b0, b1);
        pdl_wait(); // the parent may be the previous kernel's output
        for (uint32_t pos = blockIdx.x; pos < nitems; pos += grid) {
            mbar_wait(smem_u32(&empty_bar[stage]), parity);
            uint32_t *d = item_smem[stage];
            d[lane] = a0;
            if (lane + 32 < IW) d[lane + 32] = a1;
            __syncwarp();
            a0 = b0;
            a1 = b1;
            fetch(pos + 2 * grid, b0, b1);
            const OrbitItem *it = reinterpret_cast<const OrbitItem *>(d);
            const int ntile = it->nblock; // parent blocks of this item (>= its output tiles)
            const uint32_t fb = smem_u32(&full_bar[stage]);
            if (lane == 0) mbar_expect_tx(fb, (uint32_t)(ntile * O.tile_bytes)); // arrive (release) + transaction bytes
            else mbar_arrive(fb);                                                // arrive (release) of this lane's item words
            __syncwarp();
            if (O.debug & 1) {
                if (lane == 0) asm volatile("mbarrier.complete_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(fb), "r"((uint32_t)(ntile * O.tile_bytes)) : "memory");
            } else if (lane < ntile) {
                tma_load_rank(O.rank, ring_u32 + (uint32_t)(stage * O.stage_bytes + lane * O.tile_bytes), &min, fb, it->pcrd[lane]);
            }
            if (++stage == S) {
                stage = 0;
                parity ^= 1u;
            }
        }
    } else {
        // ---------------- consumer warps ----------------
        OrbitThread<NIN> th;
        orbit_thread_init<NIN>(O, tid, th);
        const uint32_t staging0 = (uint32_t)(S * O.stage_bytes);
        int stage = 0;
        uint32_t parity = 0, sbuf = 0, tcount = 0;
        const int K = O.nstaging;
        const int64_t st_t = orbit_store_toff(O, tid);
        pdl_wait(); // the output may still be read or written by the previous kernel
        for (uint32_t pos = blockIdx.x; pos < nitems; pos += grid) {
            mbar_wait(smem_u32(&full_bar[stage]), parity);
            const OrbitItem *it = reinterpret_cast<const OrbitItem *>(item_smem[stage]);
            const int ntile = it->ntile;
            for (int m = 0; m < ntile; ++m) {
                const uint32_t slots = *reinterpret_cast<const uint32_t *>(it->slot[m]);
                const uint32_t sbuf_off = staging0 + sbuf * (uint32_t)O.tile_bytes;
                if (!(O.debug & 4)) orbit_compute<CT, RC, NIN, EPT>(O, th, ring, (uint32_t)(stage * O.stage_bytes), slots, sbuf_off);
                const bool direct_now = O.direct_store == 1 || (O.direct_store == 2 && (tcount & 1u)); // 2: alternate TMA store / st.global
                ++tcount;
                if (direct_now) {
                    // two staging buffers: a thread can only reach the writes of tile m+2 (same buffer) through the barrier
                    // of tile m+1, which every thread passes after its own reads of tile m below
                    if (!(O.debug & 16)) asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
                    if (!(O.debug & 2)) orbit_store_direct(O, tid, st_t, ring, sbuf_off, O.out_base + it->ooff[m]);
                } else {
                    if (!(O.debug & 8)) asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // generic-proxy writes -> visible to the TMA store
                    // the NEXT tile is computed into the buffer the store of K-1 tiles ago reads from: that store must have
                    // finished reading before anyone passes the barrier (at most K-2 younger stores may still be pending)
                    if (tid == 0) {
                        // (alternating mode: every second tile is a TMA store, so the buffer is reused two stores later)
                        switch (O.direct_store == 2 ? (K >= 4 ? 3 : 2) : K) {
                        case 2: asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); break;
                        case 3: asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); break;
                        default: asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory"); break;
                        }
                    }
                    if (!(O.debug & 16)) asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
                    if (tid == 0 && !(O.debug & 2)) {
                        tma_store_rank(O.rank, &mout, ring_u32 + sbuf_off, it->ocrd[m]);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
                if (++sbuf == (uint32_t)K) sbuf = 0;
            }
            mbar_arrive(smem_u32(&empty_bar[stage])); // this thread is done with the stage (and its item copy)
            if (++stage == S) {
                stage = 0;
                parity ^= 1u;
            }
        }
        if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); // stores complete before the CTA retires
    }
}

struct OrbitEntry {
    KernelKey key;
    int logt;
    cudaError_t (*launch)(const OrbitParams &, const CUtensorMap *, int grid, size_t smem, cudaStream_t);
    cudaError_t (*occupancy)(int *nblocks, size_t smem);
    const void *func;
};

template <class CT, int RC, int NIN, int EPT, int LOGT> struct OrbitLaunch {
    static cudaError_t launch(const OrbitParams &O, const CUtensorMap *maps, int grid, size_t smem, cudaStream_t s)
    {
        auto k = map_orbit_kernel<CT, RC, NIN, EPT, LOGT>;
        cudaError_t e = ensure_dynamic_smem((const void *)k, smem);
        if (e != cudaSuccess) return e;
        return launch_pdl(k, grid, (1 << LOGT) + 32, smem, s, O, maps[0], maps[1]);
    }
    static cudaError_t occupancy(int *nb, size_t smem)
    {
        auto k = map_orbit_kernel<CT, RC, NIN, EPT, LOGT>;
        cudaError_t e = ensure_dynamic_smem((const void *)k, smem);
        if (e != cudaSuccess) return e;
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(nb, k, (1 << LOGT) + 32, smem);
    }
    static const void *func() { return (const void *)map_orbit_kernel<CT, RC, NIN, EPT, LOGT>; }
};

#define SB_ORBIT_ENTRY(CT, DT, RC, NIN, EPT, LOGT)                                                                   \
    OrbitEntry { KernelKey{DT, RC, NIN, EPT, 1}, LOGT, &OrbitLaunch<CT, RC, NIN, EPT, LOGT>::launch,                 \
                 &OrbitLaunch<CT, RC, NIN, EPT, LOGT>::occupancy, OrbitLaunch<CT, RC, NIN, EPT, LOGT>::func() }

const OrbitEntry *orbit_table(int *n);
const OrbitEntry *find_orbit_kernel(const KernelKey &k, int logt);

} // namespace sb
