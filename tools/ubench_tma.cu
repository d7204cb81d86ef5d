// ubench_tma.cu -- measurement aid (not product code): what the TMA unit and the LSU path of a B200 SM sustain for tiled
// 2-D copies of a Float64 N x N matrix (the traffic pattern of config 2).  Build + run (GPU box):
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/ubench_tma tools/ubench_tma.cu && /tmp/ubench_tma
// Modes: 0 TMA load + TMA store, 1 TMA load only, 2 TMA store only, 3 TMA load + st.global.v4 store, 4 st.global only,
//        5 ld.global.v4 + st.global.v4 through registers (plain tiled copy)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t b, uint32_t ph)
{
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b), "r"(ph) : "memory");
}
__device__ __forceinline__ void tma_ld(uint32_t dst, const CUtensorMap *m, uint32_t bar, int x, int y)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(m), "r"(bar), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void tma_st(const CUtensorMap *m, uint32_t src, int x, int y)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(m), "r"(src), "r"(x), "r"(y) : "memory");
}

struct P {
    int n, bx, by, ntx, nty, stages, mode, transposed_order;
    const double *src;
    double *dst;
};

__global__ void __launch_bounds__(256) k(const __grid_constant__ P p, const __grid_constant__ CUtensorMap ms, const __grid_constant__ CUtensorMap md)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t full[8];
    const int tid = threadIdx.x, S = p.stages;
    const uint32_t tile_bytes = (uint32_t)(p.bx * p.by * 8);
    if (tid == 0) {
        for (int s = 0; s < S; ++s) mbar_init(s32(&full[s]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int ntiles = p.ntx * p.nty, grid = gridDim.x;
    const int mine = (ntiles - (int)blockIdx.x + grid - 1) / grid;
    const bool tload = p.mode == 0 || p.mode == 1 || p.mode == 3;
    auto coords = [&](int i, int &x, int &y) {
        const int t = blockIdx.x + i * grid;
        int a = t % p.ntx, b = t / p.ntx;
        if (p.transposed_order) { a = t / p.nty; b = t % p.nty; }
        x = a * p.bx;
        y = b * p.by;
    };
    auto issue = [&](int i) {
        int x, y;
        coords(i, x, y);
        const uint32_t fb = s32(&full[i % S]);
        mbar_expect(fb, tile_bytes);
        tma_ld(s32(smem) + (uint32_t)(i % S) * tile_bytes, &ms, fb, x, y);
    };
    if (tload && tid == 0)
        for (int i = 0; i < S - 1 && i < mine; ++i) issue(i);
    for (int i = 0; i < mine; ++i) {
        int x, y;
        coords(i, x, y);
        const int st = i % S;
        if (tload) {
            if (tid == 0 && i + S - 1 < mine) {
                // stage (i+S-1)%S == (i-1)%S was consumed in iteration i-1 (TMA store: wait until it has been read)
                if (p.mode == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                issue(i + S - 1);
            }
            mbar_wait(s32(&full[st]), (uint32_t)((i / S) & 1));
        }
        if (p.mode == 0 || p.mode == 2) {
            if (tid == 0) {
                if (p.mode == 2) asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
                tma_st(&md, s32(smem) + (uint32_t)st * tile_bytes, x, y);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        } else if (p.mode == 3 || p.mode == 4) {
            // all threads: 16-byte groups of the dense box, row-major
            const int gpr = p.bx / 2; // groups per row
            const unsigned char *sb = smem + (size_t)st * tile_bytes;
            for (int g = tid; g < gpr * p.by; g += 256) {
                const int r = g / gpr, c = g % gpr;
                const float4 v = *reinterpret_cast<const float4 *>(sb + (size_t)g * 16);
                __stcs(reinterpret_cast<float4 *>(p.dst + (size_t)(y + r) * p.n + x + c * 2), v);
            }
        } else if (p.mode == 5) {
            const int gpr = p.bx / 2;
            for (int g = tid; g < gpr * p.by; g += 256) {
                const int r = g / gpr, c = g % gpr;
                const float4 v = __ldcs(reinterpret_cast<const float4 *>(p.src + (size_t)(y + r) * p.n + x + c * 2));
                __stcs(reinterpret_cast<float4 *>(p.dst + (size_t)(y + r) * p.n + x + c * 2), v);
            }
        }
        if (p.mode != 5 && p.mode != 2) __syncthreads();
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*EncFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                          const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main()
{
    const int n = 4096;
    double *a, *b;
    CK(cudaMalloc(&a, (size_t)n * n * 8));
    CK(cudaMalloc(&b, (size_t)n * n * 8));
    CK(cudaMemset(a, 1, (size_t)n * n * 8));
    void *fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
    EncFn enc = (EncFn)fp;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const char *mname[] = {"tma_ld+tma_st", "tma_ld only", "tma_st only", "tma_ld+stg", "stg only", "ldg+stg"};
    struct Cfg { int bx, by, stages, cta_per_sm; };
    const Cfg cfgs[] = {{32, 32, 4, 2}, {32, 32, 4, 4}, {64, 32, 4, 2}, {64, 64, 2, 2}, {64, 64, 3, 2}, {128, 32, 3, 2}, {256, 16, 3, 2}, {4, 256, 3, 2}};
    printf("%-16s %-10s %-6s %-5s %-6s %9s %9s\n", "mode", "box", "stages", "cta/sm", "order", "us", "TB/s");
    for (const Cfg &c : cfgs)
        for (int mode = 0; mode < 6; ++mode)
            for (int tr = 0; tr < 2; ++tr) {
                if (mode == 5 && (c.stages != 4 && c.stages != 2)) continue;
                CUtensorMap ms, md;
                cuuint64_t gd[2] = {(cuuint64_t)n, (cuuint64_t)n}, gs[1] = {(cuuint64_t)n * 8};
                cuuint32_t box[2] = {(cuuint32_t)c.bx, (cuuint32_t)c.by}, es[2] = {1, 1};
                if (enc(&ms, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, a, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS ||
                    enc(&md, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, b, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
                    printf("encode failed\n");
                    continue;
                }
                P p{n, c.bx, c.by, n / c.bx, n / c.by, c.stages, mode, tr, a, b};
                const size_t smem = (size_t)c.stages * c.bx * c.by * 8;
                const int grid = 148 * c.cta_per_sm;
                for (int w = 0; w < 3; ++w) k<<<grid, 256, smem>>>(p, ms, md);
                CK(cudaDeviceSynchronize());
                const int reps = 20;
                CK(cudaEventRecord(e0));
                for (int r = 0; r < reps; ++r) k<<<grid, 256, smem>>>(p, ms, md);
                CK(cudaEventRecord(e1));
                CK(cudaDeviceSynchronize());
                float ms_ = 0;
                CK(cudaEventElapsedTime(&ms_, e0, e1));
                const double us = ms_ * 1e3 / reps;
                const double bytes = (double)n * n * 8 * ((mode == 1 || mode == 2 || mode == 4) ? 1 : 2);
                char bs[32];
                snprintf(bs, sizeof bs, "%dx%d", c.bx, c.by);
                printf("%-16s %-10s %-6d %-5d %-6s %9.2f %9.3f\n", mname[mode], bs, c.stages, c.cta_per_sm, tr ? "col" : "row", us, bytes / us / 1e6);
            }
    return 0;
}
