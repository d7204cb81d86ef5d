// ubench_reduce.cu -- measurement aid (not product code): how fast can ONE launch reduce a dense Float64 vector to a
// scalar (sum of squares, the per-GPU share of BASELINE config 5: 4096 x 4096 -> 1), and where does the time go.
// Build + run (GPU box):
//   nvcc -std=c++17 -O3 --fmad=false -gencode arch=compute_100a,code=sm_100a -o /tmp/ubench_reduce tools/ubench_reduce.cu && /tmp/ubench_reduce
// Variants (all deterministic: partials are folded in a fixed order, no floating-point atomics):
//   ldg    persistent CTAs, U x 128-bit ld.global.nc per thread in flight, block fold, partial + arrival counter, last CTA folds
//          tail=0: __threadfence + atomicAdd by thread 0 + 2 barriers (what reduce_tile_kernel does)
//          tail=1: thread 0 stores the partial and arrives with ONE atom.add.acq_rel.gpu (no fence by 256 threads)
//   bulk   cp.async.bulk (1-D) ring: one elected thread keeps S chunks of C bytes in flight per CTA from cycle 0,
//          consumer warps fold from shared memory
//   empty  kernel that only does the PDL handshake (launch floor of a graph node)
// Every number is a CUDA-graph replay of R back-to-back launches (PDL edges), divided by R.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct RP {
    const double *x;
    double *out;
    double *partials;
    unsigned *counter;
    long long n; // elements, multiple of 2
    int tail;
    int chunk_bytes, stages;
};

__device__ __forceinline__ double2 ldg_stream(const double2 *p)
{
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

template <int THREADS> __device__ __forceinline__ double block_fold(double p, double *sm)
{
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) p += __shfl_xor_sync(0xffffffffu, p, m);
    if (lane == 0) sm[warp] = p;
    __syncthreads();
    double q = 0.0;
    if (warp == 0) {
        q = lane < THREADS / 32 ? sm[lane] : 0.0;
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) q += __shfl_xor_sync(0xffffffffu, q, m);
    }
    return q; // valid in warp 0
}

// fold of the grid's partials by the last-arriving CTA
template <int THREADS> __device__ __forceinline__ void finish(const RP &P, double q, double *sm)
{
    const int t = threadIdx.x;
    __shared__ unsigned last;
    if (P.tail == 0) {
        if (t == 0) P.partials[blockIdx.x] = q;
        __threadfence();
        __syncthreads();
        if (t == 0) last = atomicAdd(P.counter, 1u) == gridDim.x - 1u;
        __syncthreads();
        if (!last) return;
        __threadfence();
    } else {
        if (t == 0) {
            P.partials[blockIdx.x] = q;
            unsigned old;
            asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(P.counter) : "memory");
            last = old == gridDim.x - 1u;
        }
        __syncthreads();
        if (!last) return;
    }
    double p = 0.0;
    for (int i = t; i < (int)gridDim.x; i += THREADS) p += __ldcg(P.partials + i);
    __syncthreads();
    const double tot = block_fold<THREADS>(p, sm);
    if (t == 0) {
        P.out[0] = P.out[0] * 0.0 + tot; // op(initop(out), total) with a read of the output, like the product
        *P.counter = 0u;
    }
}

template <int THREADS, int U> __global__ void __launch_bounds__(THREADS) k_ldg(const __grid_constant__ RP P)
{
    __shared__ double sm[32];
    pdl_launch();
    const long long nvec = P.n >> 1;
    const double2 *x = reinterpret_cast<const double2 *>(P.x);
    const long long stride = (long long)gridDim.x * THREADS;
    long long i = (long long)blockIdx.x * THREADS + threadIdx.x;
    double acc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) acc[u] = 0.0;
    pdl_wait();
    for (; i + (U - 1) * stride < nvec; i += U * stride) {
        double2 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = ldg_stream(x + i + u * stride);
#pragma unroll
        for (int u = 0; u < U; ++u) acc[u] += v[u].x * v[u].x + v[u].y * v[u].y;
    }
    for (; i < nvec; i += stride) {
        const double2 v = ldg_stream(x + i);
        acc[0] += v.x * v.x + v.y * v.y;
    }
    double p = 0.0;
#pragma unroll
    for (int u = 0; u < U; ++u) p += acc[u];
    const double q = block_fold<THREADS>(p, sm);
    finish<THREADS>(P, q, sm);
}

// contiguous CTA slabs instead of grid-stride interleave (DRAM page locality per CTA)
template <int THREADS, int U> __global__ void __launch_bounds__(THREADS) k_ldg_slab(const __grid_constant__ RP P)
{
    __shared__ double sm[32];
    pdl_launch();
    const long long nvec = P.n >> 1;
    const double2 *x = reinterpret_cast<const double2 *>(P.x);
    const long long per = (nvec + gridDim.x - 1) / gridDim.x;
    const long long b0 = per * blockIdx.x, b1 = b0 + per < nvec ? b0 + per : nvec;
    double acc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) acc[u] = 0.0;
    long long i = b0 + threadIdx.x;
    pdl_wait();
    for (; i + (U - 1) * THREADS < b1; i += U * THREADS) {
        double2 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = ldg_stream(x + i + u * THREADS);
#pragma unroll
        for (int u = 0; u < U; ++u) acc[u] += v[u].x * v[u].x + v[u].y * v[u].y;
    }
    for (; i < b1; i += THREADS) {
        const double2 v = ldg_stream(x + i);
        acc[0] += v.x * v.x + v.y * v.y;
    }
    double p = 0.0;
#pragma unroll
    for (int u = 0; u < U; ++u) p += acc[u];
    const double q = block_fold<THREADS>(p, sm);
    finish<THREADS>(P, q, sm);
}

// cp.async.bulk ring
__device__ __forceinline__ void mbar_init(uint32_t b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(b), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint32_t b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t b, uint32_t ph)
{
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b), "r"(ph) : "memory");
}
__device__ __forceinline__ void bulk_ld(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <int THREADS> __global__ void __launch_bounds__(THREADS + 32) k_bulk(const __grid_constant__ RP P)
{
    extern __shared__ __align__(128) unsigned char ring[];
    __shared__ __align__(8) uint64_t full[16], empty[16];
    __shared__ double sm[32];
    const int tid = threadIdx.x, S = P.stages, C = P.chunk_bytes;
    pdl_launch();
    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(s32(&full[s]), 1);
            mbar_init(s32(&empty[s]), THREADS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long bytes = P.n * 8;
    const long long nchunks = (bytes + C - 1) / C;
    const unsigned char *src = reinterpret_cast<const unsigned char *>(P.x);
    if (tid >= THREADS) { // producer warp
        if (tid == THREADS) {
            int st = 0;
            uint32_t par = 1;
            pdl_wait();
            for (long long c = blockIdx.x; c < nchunks; c += gridDim.x) {
                mbar_wait(s32(&empty[st]), par);
                const long long off = c * C;
                const uint32_t nb = (uint32_t)(bytes - off < C ? bytes - off : C);
                mbar_expect(s32(&full[st]), nb);
                bulk_ld(s32(ring) + (uint32_t)st * C, src + off, nb, s32(&full[st]));
                if (++st == S) { st = 0; par ^= 1u; }
            }
        }
        return;
    }
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
    int st = 0;
    uint32_t par = 0;
    for (long long c = blockIdx.x; c < nchunks; c += gridDim.x) {
        mbar_wait(s32(&full[st]), par);
        const long long off = c * C;
        const int nb = (int)(bytes - off < C ? bytes - off : C);
        const double2 *p = reinterpret_cast<const double2 *>(ring + (size_t)st * C);
        const int nv = nb >> 4;
        int i = tid;
        for (; i + 3 * THREADS < nv; i += 4 * THREADS) {
            const double2 a = p[i], b = p[i + THREADS], c2 = p[i + 2 * THREADS], d = p[i + 3 * THREADS];
            acc0 += a.x * a.x + a.y * a.y;
            acc1 += b.x * b.x + b.y * b.y;
            acc2 += c2.x * c2.x + c2.y * c2.y;
            acc3 += d.x * d.x + d.y * d.y;
        }
        for (; i < nv; i += THREADS) {
            const double2 a = p[i];
            acc0 += a.x * a.x + a.y * a.y;
        }
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(s32(&empty[st]));
        if (++st == S) { st = 0; par ^= 1u; }
    }
    // named barrier over the consumer threads only inside block_fold would deadlock with the retired producer warp:
    // use bar.sync with an explicit count
    double p = (acc0 + acc1) + (acc2 + acc3);
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) p += __shfl_xor_sync(0xffffffffu, p, m);
    if (lane == 0) sm[warp] = p;
    asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory");
    double q = 0.0;
    if (warp == 0) {
        q = lane < THREADS / 32 ? sm[lane] : 0.0;
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) q += __shfl_xor_sync(0xffffffffu, q, m);
    }
    __shared__ unsigned last;
    if (tid == 0) {
        P.partials[blockIdx.x] = q;
        unsigned old;
        asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(P.counter) : "memory");
        last = old == gridDim.x - 1u;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory");
    if (!last) return;
    double pp = 0.0;
    for (int i = tid; i < (int)gridDim.x; i += THREADS) pp += __ldcg(P.partials + i);
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) pp += __shfl_xor_sync(0xffffffffu, pp, m);
    asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory");
    if (lane == 0) sm[warp] = pp;
    asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory");
    if (tid == 0) {
        double tot = 0.0;
        for (int w = 0; w < THREADS / 32; ++w) tot += sm[w];
        P.out[0] = P.out[0] * 0.0 + tot;
        *P.counter = 0u;
    }
}

__global__ void k_empty(const __grid_constant__ RP P)
{
    pdl_launch();
    pdl_wait();
    if (P.n < 0) P.out[0] = 0.0;
}

template <class K> static float time_graph(K launch, int reps, cudaStream_t s)
{
    for (int i = 0; i < 3; ++i) launch(i);
    CK(cudaStreamSynchronize(s));
    cudaGraph_t g;
    cudaGraphExec_t ge;
    CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < reps; ++i) launch(i);
    CK(cudaStreamEndCapture(s, &g));
    CK(cudaGraphInstantiate(&ge, g, 0));
    CK(cudaGraphLaunch(ge, s));
    CK(cudaStreamSynchronize(s));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(e0, s));
        CK(cudaGraphLaunch(ge, s));
        CK(cudaEventRecord(e1, s));
        CK(cudaStreamSynchronize(s));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaGraphExecDestroy(ge));
    CK(cudaGraphDestroy(g));
    return best * 1e3f / reps; // us per launch
}

template <class... A> static void launch_pdl(void (*k)(A...), int grid, int block, size_t smem, cudaStream_t s, const RP &p, bool pdl)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    CK(cudaLaunchKernelEx(&cfg, k, p));
}

#ifndef UBENCH_NO_MAIN
int main(int argc, char **argv)
{
    const long long n = argc > 1 ? atoll(argv[1]) : (1ll << 24);
    const int nbuf = 3; // rotate over 3 buffers (> L2 in total even for 2^23)
    cudaStream_t s;
    CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    std::vector<double *> bufs(nbuf);
    std::vector<double> h(n);
    for (long long i = 0; i < n; ++i) h[i] = (double)((i * 2654435761u) % 1000) / 1000.0 - 0.5;
    double want = 0.0;
    for (long long i = 0; i < n; ++i) want += h[i] * h[i];
    for (int b = 0; b < nbuf; ++b) {
        CK(cudaMalloc(&bufs[b], n * 8));
        CK(cudaMemcpy(bufs[b], h.data(), n * 8, cudaMemcpyHostToDevice));
    }
    double *out, *partials;
    unsigned *counter;
    CK(cudaMalloc(&out, 8));
    CK(cudaMalloc(&partials, 8 * 8192));
    CK(cudaMalloc(&counter, 4));
    CK(cudaMemset(counter, 0, 4));
    CK(cudaMemset(out, 0, 8));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    printf("# n=%lld (%.1f MiB) SMs=%d; us per launch (graph of 60, PDL), rot = rotating over %d buffers, same = one buffer\n", n, n * 8.0 / 1048576.0, sms, nbuf);
    auto report = [&](const char *name, int grid, int threads, int u, int tail, float us_rot, float us_same) {
        double got;
        CK(cudaMemcpy(&got, out, 8, cudaMemcpyDeviceToHost));
        const double rel = fabs(got - want) / want;
        printf("%-10s grid=%5d thr=%4d U=%2d tail=%d  rot %7.2f us %6.0f GB/s   same %7.2f us %6.0f GB/s   relerr %.1e\n", name, grid, threads, u, tail, us_rot,
               n * 8.0 / us_rot * 1e-3, us_same, n * 8.0 / us_same * 1e-3, rel);
        fflush(stdout);
    };
    RP P{};
    P.out = out;
    P.partials = partials;
    P.counter = counter;
    P.n = n;
    const int R = 60;
    {
        P.x = bufs[0];
        for (int pdl = 0; pdl < 2; ++pdl) {
            float us = time_graph([&](int) { launch_pdl(k_empty, 148, 256, 0, s, P, pdl != 0); }, 200, s);
            printf("empty kernel, grid 148, pdl=%d: %.2f us per graph node\n", pdl, us);
        }
    }
#define RUN_LDG(KERN, NAME, TH, U)                                                                                     \
    for (int mult = 1; mult <= 8; mult *= 2) {                                                                          \
        const int grid = sms * mult;                                                                                    \
        if (grid * TH > sms * 2048) continue;                                                                           \
        for (int tail = 0; tail < 2; ++tail) {                                                                          \
            P.tail = tail;                                                                                              \
            float a = time_graph([&](int i) { RP q = P; q.x = bufs[i % nbuf]; launch_pdl(KERN<TH, U>, grid, TH, 0, s, q, true); }, R, s); \
            float b = time_graph([&](int i) { RP q = P; q.x = bufs[0]; launch_pdl(KERN<TH, U>, grid, TH, 0, s, q, true); }, R, s);        \
            report(NAME, grid, TH, U, tail, a, b);                                                                      \
        }                                                                                                               \
    }
    const bool quick = argc > 2; // calibration run: only the best variant of each family
    if (quick) {
        P.tail = 1;
        float a = time_graph([&](int i) { RP q = P; q.x = bufs[i % nbuf]; launch_pdl(k_ldg<512, 4>, sms * 4, 512, 0, s, q, true); }, R, s);
        float b = time_graph([&](int i) { RP q = P; q.x = bufs[0]; launch_pdl(k_ldg<512, 4>, sms * 4, 512, 0, s, q, true); }, R, s);
        report("ldg", sms * 4, 512, 4, 1, a, b);
        const int chunk = 32768, stages = 4;
        CK(cudaFuncSetAttribute(k_bulk<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, chunk * stages));
        P.chunk_bytes = chunk;
        P.stages = stages;
        a = time_graph([&](int i) { RP q = P; q.x = bufs[i % nbuf]; launch_pdl(k_bulk<256>, sms, 288, (size_t)chunk * stages, s, q, true); }, R, s);
        b = time_graph([&](int i) { RP q = P; q.x = bufs[0]; launch_pdl(k_bulk<256>, sms, 288, (size_t)chunk * stages, s, q, true); }, R, s);
        report("bulk32k", sms, 256, stages, 1, a, b);
        return 0;
    }
    RUN_LDG(k_ldg, "ldg", 256, 4)
    RUN_LDG(k_ldg, "ldg", 256, 8)
    RUN_LDG(k_ldg, "ldg", 512, 4)
    RUN_LDG(k_ldg, "ldg", 512, 8)
    RUN_LDG(k_ldg, "ldg", 1024, 4)
    RUN_LDG(k_ldg_slab, "ldg_slab", 256, 8)
    RUN_LDG(k_ldg_slab, "ldg_slab", 512, 8)
    // bulk ring
    for (int chunk : {8192, 16384, 32768}) {
        for (int stages : {4, 8}) {
            for (int mult = 1; mult <= 2; ++mult) {
                const size_t smem = (size_t)chunk * stages;
                if (smem * mult > 220 * 1024) continue;
                CK(cudaFuncSetAttribute(k_bulk<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                CK(cudaFuncSetAttribute(k_bulk<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                P.chunk_bytes = chunk;
                P.stages = stages;
                const int grid = sms * mult;
                float a = time_graph([&](int i) { RP q = P; q.x = bufs[i % nbuf]; launch_pdl(k_bulk<256>, grid, 288, smem, s, q, true); }, R, s);
                float b = time_graph([&](int i) { RP q = P; q.x = bufs[0]; launch_pdl(k_bulk<256>, grid, 288, smem, s, q, true); }, R, s);
                char nm[64];
                snprintf(nm, sizeof nm, "bulk%dk", chunk / 1024);
                report(nm, grid, 256, stages, 1, a, b);
                a = time_graph([&](int i) { RP q = P; q.x = bufs[i % nbuf]; launch_pdl(k_bulk<512>, grid, 544, smem, s, q, true); }, R, s);
                b = time_graph([&](int i) { RP q = P; q.x = bufs[0]; launch_pdl(k_bulk<512>, grid, 544, smem, s, q, true); }, R, s);
                report(nm, grid, 512, stages, 1, a, b);
            }
        }
    }
    return 0;
}
#endif // UBENCH_NO_MAIN
