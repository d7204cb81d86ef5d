// ubench_stream_product.cu -- measurement aid: the PRODUCT's reduce_stream_kernel launched from the same bare C harness as
// the stand-alone ring of ubench_reduce.cu (k_bulk), side by side on one box, to separate "kernel code" from "harness /
// plan" when the two disagree (product 23.9 us vs stand-alone 22.4 us on the per-GPU share of config 5).
//   nvcc -std=c++17 -O3 --fmad=false --expt-relaxed-constexpr -gencode arch=compute_100a,code=sm_100a -I strided.jl_b200/csrc \
//        -o /tmp/ubench_stream_product tools/ubench_stream_product.cu && /tmp/ubench_stream_product
#define UBENCH_NO_MAIN
#include "ubench_reduce.cu"
#undef mbar_init
namespace ub {
using ::time_graph;
}
#include "../strided.jl_b200/csrc/stream_kernel.cuh"

namespace sb {
EnvCache &env_cache()
{
    static EnvCache c;
    return c;
}
void env_reload() {}
cudaError_t ensure_dynamic_smem(const void *func, size_t smem) { return cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); }
} // namespace sb

int main(int argc, char **argv)
{
    const long long n = argc > 1 ? atoll(argv[1]) : (1ll << 24);
    const int nbuf = 3;
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
    const int R = 60;
    auto check = [&](const char *name, float a, float b) {
        double got;
        CK(cudaMemcpy(&got, out, 8, cudaMemcpyDeviceToHost));
        printf("%-34s rot %7.2f us %6.0f GB/s   same %7.2f us   relerr %.1e\n", name, a, n * 8.0 / a * 1e-3, b, fabs(got - want) / want);
        fflush(stdout);
    };
    // stand-alone ring (ubench_reduce.cu k_bulk): 32 KB x 4 stages, 148 CTAs
    {
        RP P{};
        P.out = out;
        P.partials = partials;
        P.counter = counter;
        P.n = n;
        P.tail = 1;
        P.chunk_bytes = 32768;
        P.stages = 4;
        CK(cudaFuncSetAttribute(k_bulk<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 * 4));
        float a = time_graph([&](int i) { RP q = P; q.x = bufs[i % nbuf]; launch_pdl(k_bulk<256>, sms, 288, (size_t)32768 * 4, s, q, true); }, R, s);
        float b = time_graph([&](int i) { RP q = P; q.x = bufs[0]; launch_pdl(k_bulk<256>, sms, 288, (size_t)32768 * 4, s, q, true); }, R, s);
        check("stand-alone k_bulk 32K x 4", a, b);
    }
    // the product kernel, same harness
    for (int variant = 0; variant < 3; ++variant) {
        sb::StreamArgs A;
        memset(&A, 0, sizeof A);
        A.base[0] = (unsigned char *)out;
        A.dtype[0] = A.dtype[1] = sb::F64;
        A.op = sb::OP_ADD;
        A.initop = sb::INIT_ZERO;
        A.scratch = (unsigned char *)partials;
        A.counters = counter;
        A.peer.world = 1;
        A.prog.recipe = sb::RC_ABS2;
        sb::StreamParams &S = A.S;
        S.nelem = n;
        S.vec_bytes = n * 8;
        S.nin = 1;
        S.chunk_bytes = variant == 2 ? 16384 : 32768;
        S.nstage = variant == 1 ? 3 : (variant == 2 ? 8 : 4);
        S.stage_bytes = S.chunk_bytes;
        S.nchunks = (S.vec_bytes + S.chunk_bytes - 1) / S.chunk_bytes;
        S.nout = 1;
        S.nkd = 0;
        const size_t smem = (size_t)S.nstage * S.stage_bytes + 128;
        auto k = sb::reduce_stream_kernel<double, sb::RC_ABS2, 1>;
        CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        auto launch = [&](const double *x) {
            sb::StreamArgs q = A;
            q.base[1] = (unsigned char *)x;
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(sms);
            cfg.blockDim = dim3(sb::STREAM_THREADS);
            cfg.dynamicSmemBytes = smem;
            cfg.stream = s;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            CK(cudaLaunchKernelEx(&cfg, k, q));
        };
        float a = time_graph([&](int i) { launch(bufs[i % nbuf]); }, R, s);
        float b = time_graph([&](int) { launch(bufs[0]); }, R, s);
        char nm[96];
        snprintf(nm, sizeof nm, "product reduce_stream %dK x %d", S.chunk_bytes / 1024, S.nstage);
        check(nm, a, b);
    }
    return 0;
}
