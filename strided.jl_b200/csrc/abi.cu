// abi.cu -- the C ABI of include/strided_b200.h: context, plan cache, launches, host staging.
//
// There is NO CPU execution path here: without a CUDA device every compute entry point returns
// SB_E_NODEVICE.  (sb_plan_describe is pure host planning and works anywhere.)
#include "orbit_kernel.cuh"
#include "stream_kernel.cuh"
#include "jit.hpp"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>
#include <algorithm>
#include <array>

using namespace sb;

namespace sb {

static const MapEntry *lookup_map(const MapEntry *tab, int n, const KernelKey &k)
{
    for (int i = 0; i < n; ++i) {
        const KernelKey &e = tab[i].key;
        if (e.ct == k.ct && e.recipe == k.recipe && e.nin == k.nin && e.ept == k.ept && e.uniform == k.uniform) return &tab[i];
    }
    return nullptr;
}
static const ReduceEntry *lookup_red(const ReduceEntry *tab, int n, const KernelKey &k)
{
    for (int i = 0; i < n; ++i) {
        const KernelKey &e = tab[i].key;
        if (e.ct == k.ct && e.recipe == k.recipe && e.nin == k.nin && e.ept == k.ept && e.uniform == k.uniform) return &tab[i];
    }
    return nullptr;
}
const MapEntry *find_map_kernel(const KernelKey &k)
{
    int n = 0;
    const MapEntry *t = k.ct == F32 ? map_table_f32(&n) : k.ct == F64 ? map_table_f64(&n) : k.ct == C32 ? map_table_c32(&n) : map_table_c64(&n);
    return lookup_map(t, n, k);
}
const ReduceEntry *find_reduce_kernel(const KernelKey &k)
{
    int n = 0;
    const ReduceEntry *t = k.ct == F32 ? reduce_table_f32(&n) : k.ct == F64 ? reduce_table_f64(&n) : k.ct == C32 ? reduce_table_c32(&n) : reduce_table_c64(&n);
    return lookup_red(t, n, k);
}

} // namespace sb

static thread_local std::string g_tls_err;

namespace sb {
EnvCache &env_cache()
{
    static EnvCache c;
    return c;
}
void env_reload()
{
    EnvCache &c = env_cache();
    c.pdl = std::getenv("SB_NO_PDL") == nullptr;
    c.fused_peer = std::getenv("SB_NO_FUSED_PEER") == nullptr;
    const char *e = std::getenv("SB_JIT_MIN_ELEMENTS");
    c.jit_min_elements = e ? std::atoll(e) : (1ll << 18);
    c.jit_sync = std::getenv("SB_JIT_SYNC") != nullptr;
    c.no_group = std::getenv("SB_NO_GROUP") != nullptr;
    const char *tm = std::getenv("SB_PLAN_TABLE_MB");
    c.plan_table_mb = tm ? std::max(1ll, std::atoll(tm)) : 512;
}
cudaError_t ensure_dynamic_smem(const void *func, size_t smem)
{
    static std::mutex mu;
    static std::map<std::pair<int, const void *>, size_t> have;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(mu);
    size_t &h = have[std::make_pair(dev, func)];
    if (smem <= h) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) h = smem;
    return e;
}
} // namespace sb

// cuTensorMapEncodeTiled is fetched through the runtime (no link-time dependency on libcuda, so the library
// also loads on machines without a driver -- where it can only plan, not compute).
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn()
{
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        return (EncodeTiledFn)p;
    }();
    return fn;
}

static bool encode_tma_maps(const Plan &plan, CUtensorMap *maps)
{
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return false;
    static const CUtensorMapL2promotion l2promo = []() {
        const char *e = std::getenv("SB_TMA_L2PROMO"); // tuning knob: 0 none, 64, 128 (default), 256
        const int v = e ? std::atoi(e) : 128;
        return v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : v == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
               : v == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    }();
    const int nin = plan.tma.nin;
    for (int k = 0; k < nin; ++k) {
        const Plan::TmaGlobal &g = plan.tma_global[k];
        void *base = plan.map.base[k + 1];
        if (((uintptr_t)base & 15u) != 0) return false;
        cuuint64_t gdim[TMA_MAXRANK], gstr[TMA_MAXRANK];
        cuuint32_t box[TMA_MAXRANK], estr[TMA_MAXRANK];
        for (int i = 0; i < g.rank; ++i) {
            gdim[i] = g.gdim[i];
            box[i] = g.box[i];
            estr[i] = 1;
            if (i > 0) gstr[i - 1] = g.gstride_bytes[i];
        }
        const CUtensorMapDataType dt = g.elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                       : (plan.key.ct == F64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_UINT64);
        const CUresult r = enc(&maps[k], dt, (cuuint32_t)g.rank, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               g.swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, l2promo,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return false;
    }
    for (int k = nin; k < TMA_MAXIN; ++k) maps[k] = maps[0];
    return true;
}

// parent (load) and output (store) maps of the alias-fused orbit kernel: dense boxes, no swizzle
static bool encode_orbit_maps(const Plan &plan, CUtensorMap *maps)
{
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return false;
    for (int w = 0; w < 2; ++w) {
        const Plan::TmaGlobal &g = plan.orbit_global[w];
        void *base = plan.map.base[w == 0 ? 1 : 0];
        if (((uintptr_t)base & 15u) != 0) return false;
        cuuint64_t gdim[TMA_MAXRANK], gstr[TMA_MAXRANK];
        cuuint32_t box[TMA_MAXRANK], estr[TMA_MAXRANK];
        for (int i = 0; i < g.rank; ++i) {
            gdim[i] = g.gdim[i];
            box[i] = g.box[i];
            estr[i] = 1;
            if (i > 0) gstr[i - 1] = g.gstride_bytes[i];
        }
        const CUtensorMapDataType dt = g.elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
        const CUresult r = enc(&maps[w], dt, (cuuint32_t)g.rank, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return false;
    }
    return true;
}

struct CachedPlan {
    Plan plan;
    void *dev_order = nullptr;
    void *dev_desc = nullptr;
    void *dev_lsu = nullptr; // MapParams::lsu_desc
    void *dev_orbit = nullptr;
};

struct sb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    bool sync = true;
    DeviceInfo dev;
    std::mutex mu;
    std::string err;
    sb_stats stats{};
    // reduce partials
    void *scratch = nullptr;
    size_t scratch_bytes = 0;
    size_t table_bytes = 0; // device memory held by plan tables (tile orders, per-tile records, orbit items) of the cached plans
    bool clear_pending = false;
    // occupancy cache: kernel function -> (smem -> blocks/SM)
    std::map<std::pair<const void *, size_t>, int> occ;
    // host staging pool (sb_mapreduce_host)
    bool host_zero_copy = true; // SB_HOST_ZERO_COPY=0 turns the zero-copy mode of synchronous calls off
    void *stage = nullptr;
    size_t stage_bytes = 0;
    std::unordered_map<std::string, CachedPlan> plans;
    // peer group for reductions across GPUs (sb_peer_*): exchange buffers of all ranks, mapped through CUDA IPC
    int peer_rank = 0, peer_world = 1;
    uint32_t *peer_epoch_dev = nullptr;          // call counter of the collective calls, advanced BY THE KERNELS (graph-replay safe)
    volatile uint32_t *peer_err_host = nullptr;  // mapped host word: a kernel sets it when a peer did not show up in time
    uint32_t *peer_err_dev = nullptr;            // ... its device address
    int64_t peer_timeout_cycles = 0;
    void *peer_local = nullptr;                  // this rank's buffer (cudaMalloc)
    void *peer_buf[SB_PEER_MAX_WORLD] = {nullptr}; // [g] = rank g's buffer as seen from this process
    void *peer_tmp = nullptr;                    // local partial (SB_PEER_MAX_OUT elements of up to 16 bytes)
    // sb_mapreduce_batch: side streams between a fork and a join on `stream`
    static constexpr int NSIDE = 8;
    cudaStream_t side[NSIDE] = {nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[NSIDE] = {nullptr};
};

static bool stream_is_capturing(sb_ctx *ctx)
{
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(ctx->stream, &st) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return st != cudaStreamCaptureStatusNone;
}

static void clear_plans(sb_ctx *ctx)
{
    cudaStreamSynchronize(ctx->stream);
    for (auto &kv : ctx->plans)
        if (kv.second.dev_order) cudaFree(kv.second.dev_order);
    for (auto &kv : ctx->plans)
        if (kv.second.dev_desc) cudaFree(kv.second.dev_desc);
    for (auto &kv : ctx->plans)
        if (kv.second.dev_lsu) cudaFree(kv.second.dev_lsu);
    for (auto &kv : ctx->plans)
        if (kv.second.dev_orbit) cudaFree(kv.second.dev_orbit);
    ctx->plans.clear();
    ctx->table_bytes = 0;
    ctx->clear_pending = false;
}

// plan tables over budget (lookup_plan): drop the cache at an API entry, where nobody holds plan pointers
static void maybe_clear_plans(sb_ctx *ctx)
{
    if (ctx->clear_pending && !stream_is_capturing(ctx)) clear_plans(ctx);
}

static int set_err(sb_ctx *ctx, int code, const std::string &msg)
{
    g_tls_err = msg;
    if (ctx) ctx->err = msg;
    return code;
}
static int cuda_fail(sb_ctx *ctx, cudaError_t e, const char *what)
{
    return set_err(ctx, SB_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

// a collective kernel that gave up waiting for a peer raised the mapped flag: report it once (the context stays usable)
static int peer_check(sb_ctx *ctx)
{
    if (ctx->peer_err_host && *ctx->peer_err_host) {
        *ctx->peer_err_host = 0u;
        return set_err(ctx, SB_E_CUDA, "sb_mapreduce_allreduce: a peer rank did not arrive within SB_PEER_TIMEOUT_MS; the result of that call is invalid");
    }
    return SB_OK;
}

extern "C" {

int sb_abi_version(void) { return SB_ABI_VERSION; }

int sb_shutdown(void)
{
    sb::jit_join_workers();
    return SB_OK;
}

const char *sb_last_error(sb_ctx *ctx)
{
    if (ctx) return ctx->err.c_str();
    return g_tls_err.c_str();
}

int sb_ctx_create(int device, void *stream, sb_ctx **out)
{
    if (!out) return set_err(nullptr, SB_E_INVALID, "sb_ctx_create: null out");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        return set_err(nullptr, SB_E_NODEVICE, "no CUDA device available: strided_b200 has no CPU execution path");
    }
    if (device < 0 || device >= ndev) return set_err(nullptr, SB_E_INVALID, "sb_ctx_create: bad device index");
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaSetDevice");
    sb_ctx *c = new (std::nothrow) sb_ctx();
    if (!c) return set_err(nullptr, SB_E_NOMEM, "out of host memory");
    c->device = device;
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) {
        delete c;
        return cuda_fail(nullptr, e, "cudaGetDeviceProperties");
    }
    c->dev.sm_count = prop.multiProcessorCount;
    c->dev.ctas_per_sm = 4;
    if (const char *e = std::getenv("SB_HOST_ZERO_COPY")) c->host_zero_copy = std::atoi(e) != 0;
    env_reload();
    if (stream) {
        c->stream = (cudaStream_t)stream;
    } else {
        e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            delete c;
            return cuda_fail(nullptr, e, "cudaStreamCreate");
        }
        c->own_stream = true;
    }
    *out = c;
    return SB_OK;
}

int sb_ctx_destroy(sb_ctx *ctx)
{
    if (!ctx) return SB_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    clear_plans(ctx);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->stage) cudaFree(ctx->stage);
    sb_peer_detach(ctx);
    if (ctx->peer_local) cudaFree(ctx->peer_local);
    if (ctx->peer_tmp) cudaFree(ctx->peer_tmp);
    for (int q = 0; q < sb_ctx::NSIDE; ++q) {
        if (ctx->side[q]) cudaStreamDestroy(ctx->side[q]);
        if (ctx->ev_join[q]) cudaEventDestroy(ctx->ev_join[q]);
    }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->peer_epoch_dev) cudaFree(ctx->peer_epoch_dev);
    if (ctx->peer_err_host) cudaFreeHost((void *)ctx->peer_err_host);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return SB_OK;
}

int sb_ctx_set_stream(sb_ctx *ctx, void *stream)
{
    if (!ctx) return set_err(nullptr, SB_E_INVALID, "null ctx");
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (ctx->own_stream) {
        cudaStreamSynchronize(ctx->stream);
        cudaStreamDestroy(ctx->stream);
        ctx->own_stream = false;
    }
    ctx->stream = (cudaStream_t)stream;
    return SB_OK;
}

int sb_ctx_reload_env(sb_ctx *ctx)
{
    if (!ctx) return set_err(nullptr, SB_E_INVALID, "null ctx");
    std::lock_guard<std::mutex> lk(ctx->mu);
    env_reload();
    const char *e = std::getenv("SB_HOST_ZERO_COPY");
    ctx->host_zero_copy = e ? std::atoi(e) != 0 : true;
    clear_plans(ctx); // plan-time knobs (SB_NO_TMA, SB_ORBIT_*, ...) take effect for the plans built from now on
    return SB_OK;
}

int sb_ctx_set_sync(sb_ctx *ctx, int sync)
{
    if (!ctx) return set_err(nullptr, SB_E_INVALID, "null ctx");
    ctx->sync = sync != 0;
    return SB_OK;
}

int sb_sync(sb_ctx *ctx)
{
    if (!ctx) return set_err(nullptr, SB_E_INVALID, "null ctx");
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaStreamSynchronize");
    return peer_check(ctx);
}

int sb_malloc(sb_ctx *ctx, size_t bytes, void **out)
{
    if (!ctx || !out) return set_err(ctx, SB_E_INVALID, "sb_malloc: null argument");
    cudaSetDevice(ctx->device);
    cudaError_t e = cudaMalloc(out, bytes ? bytes : 1);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaMalloc");
    return SB_OK;
}
int sb_free(sb_ctx *ctx, void *ptr)
{
    if (!ctx) return set_err(nullptr, SB_E_INVALID, "null ctx");
    cudaError_t e = cudaFree(ptr);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaFree");
    return SB_OK;
}
int sb_memcpy_h2d(sb_ctx *ctx, void *dst, const void *src, size_t bytes)
{
    if (!ctx) return set_err(nullptr, SB_E_INVALID, "null ctx");
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "memcpy h2d");
    ctx->stats.h2d_bytes += bytes;
    return SB_OK;
}
int sb_memcpy_d2h(sb_ctx *ctx, void *dst, const void *src, size_t bytes)
{
    if (!ctx) return set_err(nullptr, SB_E_INVALID, "null ctx");
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "memcpy d2h");
    ctx->stats.d2h_bytes += bytes;
    return SB_OK;
}

int sb_get_stats(sb_ctx *ctx, sb_stats *out)
{
    if (!ctx || !out) return set_err(ctx, SB_E_INVALID, "null argument");
    *out = ctx->stats;
    return SB_OK;
}
int sb_reset_stats(sb_ctx *ctx)
{
    if (!ctx) return set_err(nullptr, SB_E_INVALID, "null ctx");
    ctx->stats = sb_stats{};
    return SB_OK;
}

int sb_plan_describe(sb_ctx *ctx, const sb_desc *desc, char *buf, size_t buflen)
{
    if (!desc || !buf || buflen == 0) return set_err(ctx, SB_E_INVALID, "sb_plan_describe: null argument");
    Plan plan;
    std::string err;
    DeviceInfo dev = ctx ? ctx->dev : DeviceInfo{};
    int rc = build_plan(*desc, dev, plan, err);
    if (rc != SB_OK) return set_err(ctx, rc, err);
    std::string s = describe_plan(plan);
    std::snprintf(buf, buflen, "%s", s.c_str());
    return SB_OK;
}

} // extern "C"

static void operand_range(const sb_desc &d, int k, int64_t &lo, int64_t &hi);

// ---- launch -----------------------------------------------------------------------------------------------
static int occupancy_of(sb_ctx *ctx, const void *func, cudaError_t (*occ)(int *, size_t), size_t smem, int &nb)
{
    auto key = std::make_pair(func, smem);
    auto it = ctx->occ.find(key);
    if (it != ctx->occ.end()) {
        nb = it->second;
        return SB_OK;
    }
    cudaError_t e = occ(&nb, smem);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "occupancy query");
    if (nb < 1) nb = 1;
    ctx->occ[key] = nb;
    return SB_OK;
}

// launch of a run-time compiled kernel (same bodies, same PDL contract as the static ones)
static cudaError_t launch_jit(const void *fn, unsigned grid, size_t smem, cudaStream_t s, void **args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelExC(&cfg, fn, args);
}

typedef std::unordered_map<std::string, CachedPlan>::iterator PlanIt;

// plan cache: everything but the base pointers (the reference re-plans on every call; config 3 is ~2 us
// of device work, so planning must not be on the critical path).  `hostlink`: the operands live in pinned HOST memory and
// are read / written by the kernel itself (zero-copy mode of sb_mapreduce_host): the planner then fuses aliased views
// from two views on, because every byte that crosses the host link twice costs twice.
static int lookup_plan(sb_ctx *ctx, const sb_desc &desc, bool hostlink, PlanIt &hit, bool grouped = false)
{
    sb_desc keyd;
    std::memset(&keyd, 0, sizeof keyd); // field-wise copy below: struct padding must not leak into the key
    keyd.ndim = desc.ndim;
    keyd.nops = desc.nops;
    std::memcpy(keyd.dims, desc.dims, sizeof keyd.dims);
    std::memcpy(keyd.strides, desc.strides, sizeof keyd.strides);
    std::memcpy(keyd.dtype, desc.dtype, sizeof keyd.dtype);
    std::memcpy(keyd.conj, desc.conj, sizeof keyd.conj);
    keyd.ntok = desc.ntok;
    for (int i = 0; i < desc.ntok && i < SB_MAX_TOKENS; ++i) {
        keyd.prog[i].kind = desc.prog[i].kind;
        keyd.prog[i].a = desc.prog[i].a;
        keyd.prog[i].re = desc.prog[i].re;
        keyd.prog[i].im = desc.prog[i].im;
    }
    keyd.op = desc.op;
    keyd.initop = desc.initop;
    keyd.init_re = desc.init_re;
    keyd.init_im = desc.init_im;
    if (keyd.ndim < 0 || keyd.ndim > SB_MAX_DIMS || keyd.nops < 1 || keyd.nops > SB_MAX_OPS || keyd.ntok < 0 || keyd.ntok > SB_MAX_TOKENS) {
        Plan bad;
        std::string err;
        return set_err(ctx, build_plan(desc, ctx->dev, bad, err), err); // let the planner produce the status + message
    }
    for (int k = 0; k < SB_MAX_OPS; ++k) { // keep only the ALIAS pattern of the bases (which operands share a parent pointer)
        uintptr_t first = 0;
        if (k < desc.nops)
            for (int q = k; q >= 0; --q)
                if (desc.base[q] == desc.base[k]) first = (uintptr_t)q + 1;
        if (k < desc.nops) first |= ((uintptr_t)desc.base[k] & 15u) << 8; // 16-byte alignment decides the vector / TMA paths
        keyd.base[k] = (void *)first;
    }
    for (int i = keyd.ndim; i < SB_MAX_DIMS; ++i) keyd.dims[i] = 0;
    for (int k = 0; k < SB_MAX_OPS; ++k)
        for (int i = 0; i < SB_MAX_DIMS; ++i)
            if (k >= keyd.nops || i >= keyd.ndim) keyd.strides[k][i] = 0;
    for (int k = keyd.nops; k < SB_MAX_OPS; ++k) keyd.dtype[k] = keyd.conj[k] = 0;
    for (int i = (keyd.ntok > 0 ? keyd.ntok : 0); i < SB_MAX_TOKENS; ++i) keyd.prog[i] = sb_tok{0, 0, 0.0, 0.0};
    std::string key((const char *)&keyd, sizeof keyd);
    key.push_back(hostlink ? 'H' : 'D');
    if (grouped) key.push_back('G'); // plan for a grouped launch (sb_mapreduce_batch): sized for many problems of this shape
    int rc;
    hit = ctx->plans.find(key);
    if (hit == ctx->plans.end()) {
        Plan fresh;
        std::string err;
        DeviceInfo dinfo = ctx->dev;
        dinfo.host_link = hostlink;
        dinfo.grouped = grouped;
        rc = build_plan(desc, dinfo, fresh, err);
        if (rc != SB_OK) return set_err(ctx, rc, err);
        ctx->stats.plans_built++;
        if (ctx->plans.size() > 4096) ctx->clear_pending = true; // (dropped at the next API entry, see maybe_clear_plans)
        CachedPlan cp;
        cp.plan = std::move(fresh);
        const bool tables = !cp.plan.tile_order.empty() || !cp.plan.tile_desc.empty() || !cp.plan.orbit_items.empty() || !cp.plan.lsu_desc.empty();
        if (tables) {
            // plan tables live on the device with the plan.  Uploading them allocates and synchronises, which is illegal
            // while the stream is being captured into a CUDA graph: say so instead of invalidating the capture.
            cudaSetDevice(ctx->device);
            if (stream_is_capturing(ctx))
                return set_err(ctx, SB_E_UNSUPPORTED, "first call of this plan inside a CUDA-graph capture: run the call once outside the capture (plan tables are uploaded on the first call)");
            // (a plain synchronous copy: nothing of this plan is in flight yet)
            // (plan tables are a cache: beyond SB_PLAN_TABLE_MB = 512 MB per context everything is dropped and rebuilt on demand)
            const size_t incoming = cp.plan.tile_order.size() * sizeof(int32_t) + cp.plan.tile_desc.size() * sizeof(TileDesc) +
                                    cp.plan.orbit_items.size() * sizeof(OrbitItem) + cp.plan.lsu_desc.size() * sizeof(int64_t);
            if (ctx->table_bytes + incoming > ((size_t)env_cache().plan_table_mb << 20)) ctx->clear_pending = true; // (dropped at the next API entry: callers may hold plan pointers)
            ctx->table_bytes += incoming;
            auto upload = [&](void **dst, const void *src, size_t bytes) -> cudaError_t {
                cudaError_t e = cudaMalloc(dst, bytes);
                if (e == cudaSuccess) e = cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
                return e;
            };
            cudaError_t e = cudaSuccess;
            if (!cp.plan.tile_order.empty()) e = upload(&cp.dev_order, cp.plan.tile_order.data(), cp.plan.tile_order.size() * sizeof(int32_t)); // alias-aware launch order
            if (e == cudaSuccess && !cp.plan.tile_desc.empty()) e = upload(&cp.dev_desc, cp.plan.tile_desc.data(), cp.plan.tile_desc.size() * sizeof(TileDesc)); // per-tile records of the TMA path
            if (e == cudaSuccess && !cp.plan.orbit_items.empty()) e = upload(&cp.dev_orbit, cp.plan.orbit_items.data(), cp.plan.orbit_items.size() * sizeof(OrbitItem)); // work items of the orbit kernel
            if (e == cudaSuccess && !cp.plan.lsu_desc.empty()) e = upload(&cp.dev_lsu, cp.plan.lsu_desc.data(), cp.plan.lsu_desc.size() * sizeof(int64_t)); // per-tile records of the LSU kernel
            if (e != cudaSuccess) { // nothing half-built stays behind
                if (cp.dev_order) cudaFree(cp.dev_order);
                if (cp.dev_desc) cudaFree(cp.dev_desc);
                if (cp.dev_lsu) cudaFree(cp.dev_lsu);
                if (cp.dev_orbit) cudaFree(cp.dev_orbit);
                return cuda_fail(ctx, e, "plan table upload");
            }
            cp.plan.tile_order.clear();
            cp.plan.tile_order.shrink_to_fit();
            cp.plan.tile_desc.clear();
            cp.plan.tile_desc.shrink_to_fit();
            cp.plan.lsu_desc.clear();
            cp.plan.lsu_desc.shrink_to_fit();
            cp.plan.orbit_items.clear();
            cp.plan.orbit_items.shrink_to_fit();
        }
        hit = ctx->plans.emplace(std::move(key), std::move(cp)).first;
    } else {
        ctx->stats.plans_cached++;
    }
    return SB_OK;
}

// `peer` != nullptr: collective call (sb_mapreduce_allreduce).  If the plan is a reduction with a single output tile whose
// accumulator type equals the output type, the exchange across GPUs is fused into the reduction kernel (*fused = true);
// otherwise NOTHING is launched (*fused = false) and the caller runs the two-kernel path.
static int run_desc(sb_ctx *ctx, const sb_desc &desc, const PeerLink *peer = nullptr, bool *fused = nullptr, bool hostlink = false)
{
    PlanIt hit;
    int rc = lookup_plan(ctx, desc, hostlink, hit);
    if (rc != SB_OK) return rc;
    Plan plan = hit->second.plan; // copy: bases are bound per call
    plan.map.tile_order = (const int32_t *)hit->second.dev_order;
    plan.map.tile_desc = (const TileDesc *)hit->second.dev_desc;
    plan.map.lsu_desc = (const int64_t *)hit->second.dev_lsu;
    plan.orbit.items = (const OrbitItem *)hit->second.dev_orbit;
    for (int k = 0; k < MAXO; ++k) {
        plan.map.base[k] = (unsigned char *)desc.base[plan.base_src[k] < desc.nops ? plan.base_src[k] : 0];
        plan.red.base[k] = plan.map.base[k];
    }
    if (plan.kind == PLAN_MAP && plan.map.shift_last && output_overlaps_inputs(desc)) {
        // in-place update: elements must be computed exactly once -> masked edge tiles, LSU kernel (the per-tile records
        // of the TMA variant were built for the shifted tiling)
        plan.map.shift_last = 0;
        plan.tma_ok = false;
        plan.map.lsu_desc = nullptr; // (the records were built for the shifted tiling)
    }
    if (peer) {
        const bool can = plan.kind == PLAN_REDUCE && plan.red.nouttiles == 1 && plan.red.nout_tile <= PEER_MAX_OUT && plan.key.ct != C64 &&
                         desc.dtype[0] == plan.key.ct && env_cache().fused_peer;
        *fused = can;
        if (!can) return SB_OK;
        plan.red.peer = *peer;
    }
    if (plan.kind == PLAN_NOOP) return SB_OK;
    cudaSetDevice(ctx->device);
    // Run-time specialised element function (NVRTC) instead of the interpreter, once the problem is large enough to
    // amortise a ~2 s compile (the GPU analog of MINTHREADLENGTH, reference src/mapreduce.jl:141).
    const JitKernel *jk = nullptr;
    // (plans with an alias-fused orbit variant keep the in-kernel interpreter: the orbit kernel is bound by its memory
    //  request rate, not by instruction issue, and beats the generic kernel + JIT by 2-3x on aliased views)
    // (... and reductions for which the streamed kernel has its own functor: faster than the NVRTC-specialised tile kernel)
    if (plan.key.recipe == RC_INTERP && jit_enabled() && !(plan.kind == PLAN_MAP && plan.orbit_ok) &&
        !(plan.kind == PLAN_REDUCE && plan.stream_ok && plan.stream_recipe != RC_INTERP)) {
        const int64_t thr = (int64_t)env_cache().jit_min_elements;
        if (plan.elements >= thr)
            jk = jit_get(plan.kind == PLAN_MAP ? JIT_MAP : JIT_REDUCE, plan.key, plan.kind == PLAN_MAP ? plan.map.prog : plan.red.prog,
                         env_cache().jit_sync || plan.needs_jit);
    }
    if (plan.needs_jit) { // expression tree deeper than the interpreter's register stack: straight-line NVRTC code only, at any size
        if (!jk && jit_enabled()) jk = jit_get(plan.kind == PLAN_MAP ? JIT_MAP : JIT_REDUCE, plan.key, plan.kind == PLAN_MAP ? plan.map.prog : plan.red.prog, true);
        if (!jk) return set_err(ctx, SB_E_UNSUPPORTED, "program stack deeper than 4 needs the NVRTC-specialised kernel, which is unavailable (SB_NO_JIT / no libnvrtc)");
    }
    if (jk && plan.kind == PLAN_MAP) {
        const void *fn = (const void *)jk->fn;
        if (plan.smem_bytes > 48 * 1024 && ensure_dynamic_smem(fn, (size_t)plan.smem_bytes) != cudaSuccess) {
            cudaGetLastError();
            jk = nullptr;
        }
        if (jk) {
            int nb = jk->min_blocks;
            auto okey = std::make_pair(fn, (size_t)plan.smem_bytes);
            auto oit = ctx->occ.find(okey);
            if (oit != ctx->occ.end()) nb = oit->second;
            else {
                int q = 0;
                if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, fn, THREADS, (size_t)plan.smem_bytes) == cudaSuccess && q > 0) nb = q;
                else cudaGetLastError();
                ctx->occ[okey] = nb;
            }
            int64_t grid = std::min<int64_t>(plan.map.ntiles, (int64_t)ctx->dev.sm_count * nb);
            if (grid < 1) grid = 1;
            void *args[] = {(void *)&plan.map};
            cudaError_t e = launch_jit(fn, (unsigned)grid, (size_t)plan.smem_bytes, ctx->stream, args);
            if (e != cudaSuccess) return cuda_fail(ctx, e, "jit map launch");
            ctx->stats.launches++;
            ctx->stats.jit_launches++;
            return SB_OK;
        }
    }
    if (plan.kind == PLAN_MAP && plan.orbit_ok && plan.orbit.items) { // alias-fused orbits: parent read once, output written once
        // bind-time conditions: the output must not overlap the parent (blocks are re-read after other tiles were stored)
        int64_t lo0, hi0, lo1, hi1;
        operand_range(desc, plan.base_src[0], lo0, hi0);
        operand_range(desc, plan.base_src[1], lo1, hi1);
        const uintptr_t o0 = (uintptr_t)plan.map.base[0], p0 = (uintptr_t)plan.map.base[1];
        const bool overlap = (o0 + (uintptr_t)lo0 < p0 + (uintptr_t)hi1) && (p0 + (uintptr_t)lo1 < o0 + (uintptr_t)hi0);
        const OrbitEntry *ok = overlap ? nullptr : find_orbit_kernel(KernelKey{plan.key.ct, plan.key.recipe, plan.orbit.nin, plan.orbit.ept, 1}, plan.orbit.log_threads);
        alignas(64) CUtensorMap maps[2];
        if (ok && encode_orbit_maps(plan, maps)) {
            plan.orbit.out_base = plan.map.base[0];
            int nb = 1;
            rc = occupancy_of(ctx, ok->func, ok->occupancy, (size_t)plan.orbit_smem_bytes, nb);
            if (rc != SB_OK) return rc;
            int64_t grid = std::min<int64_t>(plan.orbit.nitems, (int64_t)ctx->dev.sm_count * nb);
            if (grid < 1) grid = 1;
            cudaError_t e = ok->launch(plan.orbit, maps, (int)grid, (size_t)plan.orbit_smem_bytes, ctx->stream);
            if (e != cudaSuccess) return cuda_fail(ctx, e, "map_orbit launch");
            ctx->stats.launches++;
            return SB_OK;
        }
    }
    if (plan.kind == PLAN_MAP && plan.tma_ok) { // TMA-pipelined variant when the inputs qualify at bind time
        const TmaEntry *tk = find_tma_kernel(plan.key);
        alignas(64) CUtensorMap maps[TMA_MAXIN];
        if (tk && encode_tma_maps(plan, maps)) {
            int nb = 1;
            rc = occupancy_of(ctx, tk->func, tk->occupancy, (size_t)plan.tma_smem_bytes, nb);
            if (rc != SB_OK) return rc;
            int64_t grid = std::min<int64_t>(plan.map.ntiles, (int64_t)ctx->dev.sm_count * nb);
            if (grid < 1) grid = 1;
            cudaError_t e = tk->launch(plan.map, plan.tma, maps, (int)grid, (size_t)plan.tma_smem_bytes, ctx->stream);
            if (e != cudaSuccess) return cuda_fail(ctx, e, "map_tma launch");
            ctx->stats.launches++;
            return SB_OK;
        }
    }
    if (plan.kind == PLAN_MAP) {
        const MapEntry *k = find_map_kernel(plan.key);
        if (!k) return set_err(ctx, SB_E_UNSUPPORTED, "no map kernel instantiated for this plan");
        int nb = 1;
        rc = occupancy_of(ctx, k->func, k->occupancy, (size_t)plan.smem_bytes, nb);
        if (rc != SB_OK) return rc;
        int64_t grid = std::min<int64_t>(plan.map.ntiles, (int64_t)ctx->dev.sm_count * nb);
        if (grid < 1) grid = 1;
        cudaError_t e = k->launch(plan.map, (int)grid, (size_t)plan.smem_bytes, ctx->stream);
        if (e != cudaSuccess) return cuda_fail(ctx, e, "map_tile launch");
        ctx->stats.launches++;
    } else {
        const ReduceEntry *k = find_reduce_kernel(plan.key);
        if (!k) return set_err(ctx, SB_E_UNSUPPORTED, "no reduce kernel instantiated for this plan");
        // fixed-size counter region in front of the partials (a split plan has fewer output tiles than resident CTAs),
        // so that partials of one plan can never alias the counters of another
        const size_t counters_bytes = 64 * 1024;
        if (plan.red.nsplit > 1 && (size_t)plan.red.nouttiles * 4 > counters_bytes) return set_err(ctx, SB_E_UNSUPPORTED, "too many output tiles for a split reduction");
        size_t scratch_need = plan.red.nsplit > 1 ? counters_bytes + (size_t)plan.scratch_bytes : 0;
        if (plan.stream_ok) scratch_need = std::max(scratch_need, counters_bytes + (size_t)plan.stream_grid * (size_t)plan.stream.nout * 16);
        if (scratch_need > ctx->scratch_bytes) {
            if (stream_is_capturing(ctx))
                return set_err(ctx, SB_E_UNSUPPORTED, "the reduction scratch buffer has to grow inside a CUDA-graph capture: run the call once outside the capture");
            if (ctx->scratch) {
                cudaStreamSynchronize(ctx->stream);
                cudaFree(ctx->scratch);
                ctx->scratch = nullptr;
                ctx->scratch_bytes = 0;
            }
            size_t want = std::max<size_t>(scratch_need, 1 << 20);
            cudaError_t e = cudaMalloc(&ctx->scratch, want);
            if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaMalloc(scratch)");
            e = cudaMemsetAsync(ctx->scratch, 0, want, ctx->stream); // arrival counters start at zero; kernels re-arm them
            if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaMemset(scratch)");
            ctx->scratch_bytes = want;
        }
        plan.red.counters = (uint32_t *)ctx->scratch;
        plan.red.scratch = (unsigned char *)ctx->scratch + counters_bytes;
        cudaError_t e;
        if (!jk && plan.stream_ok) { // streamed complete reduction: dense inputs, 16-byte aligned at bind time
            bool aligned = true;
            for (int q = 1; q <= plan.stream.nin; ++q) aligned = aligned && (((uintptr_t)plan.red.base[q] & 15u) == 0);
            KernelKey skey = plan.key;
            if (plan.stream_recipe != RC_INTERP) { // a functor only the streamed kernel has (common.hpp RC_S_*)
                skey.recipe = plan.stream_recipe;
                skey.nin = plan.stream.nin;
            }
            const StreamEntry *sk = aligned ? find_stream_kernel(skey) : nullptr;
            if (sk) {
                StreamArgs sa;
                std::memset(&sa, 0, sizeof sa);
                for (int q = 0; q <= plan.stream.nin; ++q) {
                    sa.base[q] = plan.red.base[q];
                    sa.dtype[q] = plan.red.dtype[q];
                    sa.conj[q] = plan.red.conj[q];
                }
                sa.op = plan.red.op;
                sa.initop = plan.red.initop;
                sa.init_re = plan.red.init_re;
                sa.init_im = plan.red.init_im;
                sa.scratch = plan.red.scratch;
                sa.counters = plan.red.counters;
                sa.peer = plan.red.peer;
                sa.S = plan.stream;
                sa.prog = plan.red.prog;
                e = sk->launch(sa, (int)plan.stream_grid, (size_t)plan.stream_smem_bytes, ctx->stream);
                if (e != cudaSuccess) return cuda_fail(ctx, e, "reduce_stream launch");
                ctx->stats.launches++;
                return SB_OK;
            }
        }
        if (jk) {
            void *args[] = {(void *)&plan.red};
            e = launch_jit((const void *)jk->fn, (unsigned)plan.grid, (size_t)plan.smem_bytes, ctx->stream, args);
        } else {
            e = k->launch(plan.red, (int)plan.grid, (size_t)plan.smem_bytes, ctx->stream);
        }
        if (e != cudaSuccess) return cuda_fail(ctx, e, "reduce_tile launch");
        ctx->stats.launches++;
        if (jk) ctx->stats.jit_launches++;
        // (the fold of the split partials is fused into reduce_tile: last-arriving CTA per output tile)
    }
    return SB_OK;
}

extern "C" int sb_mapreduce(sb_ctx *ctx, const sb_desc *desc)
{
    if (!ctx) return set_err(nullptr, SB_E_INVALID, "sb_mapreduce: null ctx");
    if (!desc) return set_err(ctx, SB_E_INVALID, "sb_mapreduce: null desc");
    std::lock_guard<std::mutex> lk(ctx->mu);
    maybe_clear_plans(ctx);
    int rc = run_desc(ctx, *desc);
    if (rc != SB_OK) return rc;
    if (ctx->sync) {
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) return cuda_fail(ctx, e, "kernel execution");
    }
    return SB_OK;
}

// ---- batches of independent calls ---------------------------------------------------------------------------------
// Small problems (the README-sized shapes: 1000^2, 32^4) are bound by launch and DRAM latency, not by bandwidth: one after
// the other they cost ~4 us each whatever the kernel does.  A batch spreads its INDEPENDENT map calls over side streams
// between a fork and a join on the ctx's stream, so they overlap; under stream capture this becomes parallel branches of
// the CUDA graph.  A call is independent when it is a map whose output byte range overlaps no other call's operands; all
// other calls (reductions share the ctx's scratch; chains such as B = f(A); C = g(B)) run in order after the join.
static void desc_ranges(const sb_desc &d, uintptr_t (&lo)[SB_MAX_OPS], uintptr_t (&hi)[SB_MAX_OPS])
{
    for (int k = 0; k < d.nops && k < SB_MAX_OPS; ++k) {
        int64_t l, h;
        operand_range(d, k, l, h);
        lo[k] = (uintptr_t)d.base[k] + (uintptr_t)l;
        hi[k] = (uintptr_t)d.base[k] + (uintptr_t)h;
    }
}

// Several problems of ONE cached plan (same dims, strides, eltypes, program, alias pattern; different base pointers) that
// take the TMA ring kernel: one grouped launch (tma_kernel.cuh "GROUP") instead of one launch each.  Returns false when
// the plan or a problem does not qualify at bind time (the caller then issues the calls one by one).
static bool group_plan_ok(sb_ctx *ctx, const CachedPlan &cp)
{
    const Plan &pl = cp.plan;
    (void)ctx;
    if (pl.kind != PLAN_MAP || pl.needs_jit || env_cache().no_group) return false;
    if (pl.orbit_ok && cp.dev_orbit) return false;
    if (pl.key.recipe == RC_INTERP && jit_enabled() && pl.elements >= (int64_t)env_cache().jit_min_elements) return false; // NVRTC-specialised kernel
    if (pl.tma_ok) return pl.tma.nin <= TMA_GROUP_MAXIN && find_tma_group_kernel(pl.key) != nullptr;
    return find_map_group_kernel(pl.key) != nullptr; // LSU kernel
}

static int run_group(sb_ctx *ctx, const sb_desc *descs, const int *idx, int cnt, const CachedPlan &cp, bool *done)
{
    *done = false;
    Plan plan = cp.plan;
    plan.map.tile_order = (const int32_t *)cp.dev_order;
    plan.map.tile_desc = (const TileDesc *)cp.dev_desc;
    plan.map.lsu_desc = (const int64_t *)cp.dev_lsu;
    for (int p = 0; p < cnt; ++p)
        if (plan.map.shift_last && output_overlaps_inputs(descs[idx[p]])) return SB_OK; // in-place update: masked edge tiles, one by one
    auto bind = [&](const sb_desc &d) {
        for (int k = 0; k < MAXO; ++k) plan.map.base[k] = (unsigned char *)d.base[plan.base_src[k] < d.nops ? plan.base_src[k] : 0];
    };
    cudaSetDevice(ctx->device);
    if (plan.tma_ok) {
        const TmaGroupEntry *gk = find_tma_group_kernel(plan.key);
        if (!gk) return SB_OK;
        TmaGroup G;
        std::memset(&G, 0, sizeof G);
        G.nprob = cnt;
        for (int p = 0; p < cnt; ++p) {
            bind(descs[idx[p]]);
            alignas(64) CUtensorMap maps[TMA_MAXIN];
            if (!encode_tma_maps(plan, maps)) return SB_OK;
            for (int k = 0; k < TMA_GROUP_MAXIN; ++k) G.maps[p][k] = maps[k < plan.tma.nin ? k : 0];
            G.out[p] = plan.map.base[0];
        }
        int nb = 1;
        int rc = occupancy_of(ctx, gk->func, gk->occupancy, (size_t)plan.tma_smem_bytes, nb);
        if (rc != SB_OK) return rc;
        int64_t grid = std::min<int64_t>(plan.map.ntiles * cnt, (int64_t)ctx->dev.sm_count * nb);
        if (grid < 1) grid = 1;
        cudaError_t e = gk->launch(plan.map, plan.tma, G, (int)grid, (size_t)plan.tma_smem_bytes, ctx->stream);
        if (e != cudaSuccess) return cuda_fail(ctx, e, "map_tma_group launch");
    } else {
        const MapGroupEntry *gk = find_map_group_kernel(plan.key);
        if (!gk) return SB_OK;
        MapGroup G;
        std::memset(&G, 0, sizeof G);
        G.nprob = cnt;
        unsigned char *base0[MAXO];
        bind(descs[idx[0]]);
        for (int k = 0; k < MAXO; ++k) base0[k] = plan.map.base[k];
        for (int p = 0; p < cnt; ++p) {
            bind(descs[idx[p]]);
            for (int k = 0; k < MAXO; ++k) {
                // (128-bit accesses were planned from problem 0's alignment: the key carries base & 15 per operand, so every problem of the plan agrees)
                G.delta[p][k] = (int64_t)((intptr_t)plan.map.base[k] - (intptr_t)base0[k]);
            }
        }
        for (int k = 0; k < MAXO; ++k) plan.map.base[k] = base0[k];
        int nb = 1;
        int rc = occupancy_of(ctx, gk->func, gk->occupancy, (size_t)plan.smem_bytes, nb);
        if (rc != SB_OK) return rc;
        int64_t grid = std::min<int64_t>(plan.map.ntiles * cnt, (int64_t)ctx->dev.sm_count * nb);
        if (grid < 1) grid = 1;
        cudaError_t e = gk->launch(plan.map, G, (int)grid, (size_t)plan.smem_bytes, ctx->stream);
        if (e != cudaSuccess) return cuda_fail(ctx, e, "map_tile_group launch");
    }
    ctx->stats.launches++;
    ctx->stats.grouped_calls += (uint64_t)cnt;
    *done = true;
    return SB_OK;
}

extern "C" int sb_mapreduce_batch(sb_ctx *ctx, int n, const sb_desc *descs)
{
    if (!ctx) return set_err(nullptr, SB_E_INVALID, "sb_mapreduce_batch: null ctx");
    if (n < 0 || (n > 0 && !descs)) return set_err(ctx, SB_E_INVALID, "sb_mapreduce_batch: bad arguments");
    std::lock_guard<std::mutex> lk(ctx->mu);
    maybe_clear_plans(ctx);
    cudaSetDevice(ctx->device);
    std::vector<char> parallel((size_t)n, 0);
    {
        std::vector<std::array<uintptr_t, SB_MAX_OPS>> lo((size_t)n), hi((size_t)n);
        std::vector<char> sane((size_t)n, 0);
        for (int i = 0; i < n; ++i) {
            const sb_desc &d = descs[i];
            sane[i] = d.ndim >= 0 && d.ndim <= SB_MAX_DIMS && d.nops >= 1 && d.nops <= SB_MAX_OPS;
            if (!sane[i]) continue;
            bool empty = false;
            for (int q = 0; q < d.ndim; ++q) empty |= d.dims[q] <= 0;
            if (empty) { // (zero-size problems touch at most the output through initop: leave them on the main stream)
                sane[i] = 0;
                continue;
            }
            uintptr_t l[SB_MAX_OPS] = {0}, h[SB_MAX_OPS] = {0};
            desc_ranges(d, l, h);
            for (int k = 0; k < SB_MAX_OPS; ++k) {
                lo[i][k] = l[k];
                hi[i][k] = h[k];
            }
        }
        for (int i = 0; i < n; ++i) {
            if (!sane[i] || descs[i].op != SB_OP_NONE) continue;
            bool indep = true;
            for (int j = 0; j < n && indep; ++j) {
                if (j == i) continue;
                if (!sane[j]) { // unknown footprint: be conservative
                    indep = false;
                    break;
                }
                auto overlap = [&](int a, int ka, int b, int kb) { return lo[a][ka] < hi[b][kb] && lo[b][kb] < hi[a][ka]; };
                for (int k = 0; k < descs[j].nops && indep; ++k)
                    if (overlap(i, 0, j, k)) indep = false; // my output against every operand of j (write-write, j reads what I write)
                for (int q = 1; q < descs[i].nops && indep; ++q)
                    if (overlap(i, q, j, 0)) indep = false; // my inputs against j's output (I read what j writes)
            }
            parallel[i] = indep ? 1 : 0;
        }
    }
    // units of independent work: a single call, or a group of calls that share one plan (one launch, see run_group)
    struct Unit {
        std::vector<int> idx;
        const CachedPlan *cp = nullptr; // != nullptr: grouped launch
    };
    std::vector<Unit> units;
    {
        std::vector<const CachedPlan *> cps((size_t)n, nullptr);
        if (ctx->plans.size() < 4000) // (lookup_plan drops the whole cache beyond 4096 entries: no pointers across that)
            for (int i = 0; i < n; ++i) {
                if (!parallel[i]) continue;
                PlanIt hit;
                const int rc = lookup_plan(ctx, descs[i], false, hit);
                if (rc != SB_OK) return rc;
                if (hit->second.plan.kind == PLAN_MAP) cps[i] = &hit->second; // (same cached plan <=> same shapes, strides, eltypes, program)
            }
        std::vector<char> taken((size_t)n, 0);
        for (int i = 0; i < n; ++i) {
            if (!parallel[i] || taken[i]) continue;
            Unit u;
            u.idx.push_back(i);
            taken[i] = 1;
            if (cps[i]) {
                for (int j = i + 1; j < n && (int)u.idx.size() < (TMA_GROUP_MAX < MAP_GROUP_MAX ? TMA_GROUP_MAX : MAP_GROUP_MAX); ++j)
                    if (parallel[j] && !taken[j] && cps[j] == cps[i]) {
                        u.idx.push_back(j);
                        taken[j] = 1;
                    }
                if (u.idx.size() >= 2) { // the plan of a grouped launch is sized for many problems (DeviceInfo::grouped)
                    PlanIt gh;
                    const int rc = lookup_plan(ctx, descs[i], false, gh, true);
                    if (rc != SB_OK) return rc;
                    if (group_plan_ok(ctx, gh->second)) u.cp = &gh->second;
                }
            }
            if (!u.cp && u.idx.size() >= 2) { // not groupable: independent single calls, one unit each (they overlap on the side streams)
                for (int q : u.idx) {
                    Unit one;
                    one.idx.push_back(q);
                    units.push_back(std::move(one));
                }
                continue;
            }
            units.push_back(std::move(u));
        }
    }
    auto run_unit = [&](const Unit &u) -> int {
        if (u.cp) {
            bool done = false;
            const int rc = run_group(ctx, descs, u.idx.data(), (int)u.idx.size(), *u.cp, &done);
            if (rc != SB_OK || done) return rc;
        }
        for (int i : u.idx) {
            const int rc = run_desc(ctx, descs[i]);
            if (rc != SB_OK) return rc;
        }
        return SB_OK;
    };
    const int npar = (int)units.size();
    cudaStream_t main_stream = ctx->stream;
    int used = 0;
    if (npar >= 2) {
        used = npar < sb_ctx::NSIDE ? npar : sb_ctx::NSIDE;
        for (int q = 0; q < used; ++q) {
            if (!ctx->side[q]) {
                if (stream_is_capturing(ctx)) return set_err(ctx, SB_E_UNSUPPORTED, "first sb_mapreduce_batch inside a CUDA-graph capture: run a batch once outside the capture");
                cudaError_t e = cudaStreamCreateWithFlags(&ctx->side[q], cudaStreamNonBlocking);
                if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_join[q], cudaEventDisableTiming);
                if (e != cudaSuccess) return cuda_fail(ctx, e, "batch side stream");
            }
        }
        if (!ctx->ev_fork) {
            cudaError_t e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
            if (e != cudaSuccess) return cuda_fail(ctx, e, "batch fork event");
        }
        cudaError_t e = cudaEventRecord(ctx->ev_fork, main_stream);
        for (int q = 0; q < used && e == cudaSuccess; ++q) e = cudaStreamWaitEvent(ctx->side[q], ctx->ev_fork, 0);
        if (e != cudaSuccess) return cuda_fail(ctx, e, "batch fork");
        int slot = 0, rc = SB_OK;
        for (size_t q = 0; q < units.size() && rc == SB_OK; ++q) {
            ctx->stream = ctx->side[slot % used];
            ++slot;
            rc = run_unit(units[q]);
        }
        ctx->stream = main_stream;
        for (int q = 0; q < used; ++q) { // always join, also on error: the side streams must not stay forked (capture!)
            cudaError_t e2 = cudaEventRecord(ctx->ev_join[q], ctx->side[q]);
            if (e2 == cudaSuccess) e2 = cudaStreamWaitEvent(main_stream, ctx->ev_join[q], 0);
            if (e2 != cudaSuccess && rc == SB_OK) rc = cuda_fail(ctx, e2, "batch join");
        }
        if (rc != SB_OK) return rc;
    } else if (npar == 1 && units[0].cp) { // one group and nothing else to overlap with: on the ctx's stream, in order
        const int rc = run_unit(units[0]);
        if (rc != SB_OK) return rc;
        for (int i : units[0].idx) parallel[i] = 2; // done
    }
    for (int i = 0; i < n; ++i) {
        if ((npar >= 2 && parallel[i]) || parallel[i] == 2) continue;
        const int rc = run_desc(ctx, descs[i]);
        if (rc != SB_OK) return rc;
    }
    ctx->stats.batches++;
    if (ctx->sync) {
        cudaError_t e = cudaStreamSynchronize(main_stream);
        if (e != cudaSuccess) return cuda_fail(ctx, e, "sb_mapreduce_batch");
    }
    return SB_OK;
}

// ---- reductions across GPUs through peer memory (no NCCL on the path) -------------------------------------------
// Exchange buffer of a rank:  [flag of rank 0 | ... | flag of rank 7]  (128 bytes apart)  then
// data[parity][rank][SB_PEER_MAX_OUT] in 16-byte slots.  Rank r pushes its partial into slot [epoch & 1][r] of EVERY
// rank's buffer (plain stores to mapped peer memory: NVLink), fences, raises flag r in every buffer to `epoch`, waits
// until its own buffer shows all flags >= epoch, then folds the world partials in rank order.  Two parities because a
// fast rank may already push call n+1 while a slow one still folds call n (it cannot be two calls ahead: it needs the
// slow rank's flag of call n+1 first).
namespace {
static_assert(PEER_MAX_WORLD == SB_PEER_MAX_WORLD && PEER_MAX_OUT == SB_PEER_MAX_OUT, "header and kernels agree on the buffer layout");
constexpr size_t PEER_BUF_BYTES = PEER_DATA_OFF + 2 * (size_t)SB_PEER_MAX_WORLD * SB_PEER_MAX_OUT * PEER_SLOT;

struct PeerKernelParams {
    int32_t world, rank, nout, op, initop, local_empty, nkept, out_dtype, out_conj;
    int32_t pad_;
    uint32_t *epoch_ptr; // device-resident call counter (see PeerLink)
    uint32_t *err_flag;
    int64_t timeout_cycles;
    double init_re, init_im;
    int64_t kdims[MAXD], kstr_bytes[MAXD]; // kept dims of the output view and their byte strides
    unsigned char *buf[SB_PEER_MAX_WORLD];
    const unsigned char *tmp; // this rank's partial: dense, element type = output dtype
    unsigned char *out;
};

template <class AT> __global__ void __launch_bounds__(256) peer_allreduce_kernel(const __grid_constant__ PeerKernelParams K)
{
    constexpr int W = sizeof(AT) / 4;
    const int t = threadIdx.x;
    pdl_launch_dependents();
    pdl_wait(); // the local reduction wrote K.tmp
    const uint32_t epoch = __ldcg(K.epoch_ptr) + 1u; // this call's number; advanced on the device (graph replays get fresh epochs)
    __syncthreads();
    if (t == 0) __stcg(K.epoch_ptr, epoch);
    const size_t par = (size_t)(epoch & 1u) * SB_PEER_MAX_WORLD;
    if constexpr (W <= 2) {
        // Low-latency path for 4- and 8-byte elements: every 32-bit half travels in ONE 8-byte store together with the
        // epoch ({word, epoch}; 8-byte stores to peer memory are delivered atomically), so the receiver needs no flag,
        // no fence and no barrier: each thread pushes its element to all ranks and then polls its own copies.
        for (int o = t; o < K.nout; o += 256) {
            union {
                AT v;
                uint32_t w[2];
            } u;
            u.w[1] = 0u;
            u.v = K.local_empty ? red_neutral<AT>(K.op) : reinterpret_cast<const AT *>(K.tmp)[o];
            const size_t off = PEER_DATA_OFF + ((par + (size_t)K.rank) * SB_PEER_MAX_OUT + (size_t)o) * PEER_SLOT;
            for (int g = 0; g < K.world; ++g) {
                unsigned char *dst = K.buf[g] + off;
                asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(dst), "r"(u.w[0]), "r"(epoch) : "memory");
                asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(dst + 8), "r"(u.w[1]), "r"(epoch) : "memory");
            }
            AT tot = red_neutral<AT>(K.op);
            const long long t0 = clock64();
            for (int g = 0; g < K.world; ++g) {
                const unsigned char *src = K.buf[K.rank] + PEER_DATA_OFF + ((par + (size_t)g) * SB_PEER_MAX_OUT + (size_t)o) * PEER_SLOT;
                uint32_t a0, e0, a1, e1;
                for (;;) {
                    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(a0), "=r"(e0) : "l"(src) : "memory");
                    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(a1), "=r"(e1) : "l"(src + 8) : "memory");
                    if (e0 == epoch && e1 == epoch) break;
                    if (clock64() - t0 > K.timeout_cycles) { // a missing peer must not hang the GPU: flag it and go on
                        *reinterpret_cast<volatile uint32_t *>(K.err_flag) = 1u;
                        a0 = a1 = 0u;
                        break;
                    }
                }
                u.w[0] = a0;
                u.w[1] = a1;
                tot = g == 0 ? u.v : red_apply<AT>(K.op, tot, u.v);
            }
            int64_t ooff = 0;
            int rem = o;
            for (int d = 0; d < K.nkept; ++d) {
                const int c = rem % (int)K.kdims[d];
                rem /= (int)K.kdims[d];
                ooff += (int64_t)c * K.kstr_bytes[d];
            }
            AT x = load_elem<AT, false>(K.out + ooff, K.out_dtype, K.out_conj);
            x = init_apply<AT>(K.initop, K.init_re, K.init_im, x);
            store_elem<AT, false>(K.out + ooff, K.out_dtype, K.out_conj, red_apply<AT>(K.op, x, tot));
        }
    } else {
    for (int o = t; o < K.nout; o += 256) {
        union {
            AT v;
            uint32_t w[W];
        } u;
        u.v = K.local_empty ? red_neutral<AT>(K.op) : reinterpret_cast<const AT *>(K.tmp)[o];
        const size_t off = PEER_DATA_OFF + ((par + (size_t)K.rank) * SB_PEER_MAX_OUT + (size_t)o) * PEER_SLOT;
        for (int g = 0; g < K.world; ++g) {
            volatile uint32_t *dst = reinterpret_cast<volatile uint32_t *>(K.buf[g] + off);
#pragma unroll
            for (int i = 0; i < W; ++i) dst[i] = u.w[i];
        }
    }
    __threadfence_system();
    __syncthreads();
    if (t < K.world) { // thread g: tell rank g that this rank's partial has landed, then wait for rank g's
        *reinterpret_cast<volatile uint32_t *>(K.buf[t] + PEER_FLAG_STRIDE * (size_t)K.rank) = epoch;
        volatile uint32_t *f = reinterpret_cast<volatile uint32_t *>(K.buf[K.rank] + PEER_FLAG_STRIDE * (size_t)t);
        const long long t0 = clock64();
        while ((int32_t)(*f - epoch) < 0)
            if (clock64() - t0 > K.timeout_cycles) { // a missing peer must not hang the GPU: flag it and go on
                *reinterpret_cast<volatile uint32_t *>(K.err_flag) = 1u;
                break;
            }
    }
    __threadfence_system();
    __syncthreads();
    for (int o = t; o < K.nout; o += 256) {
        AT tot = red_neutral<AT>(K.op);
        for (int g = 0; g < K.world; ++g) {
            union {
                AT v;
                uint32_t w[W];
            } u;
            const volatile uint32_t *src = reinterpret_cast<const volatile uint32_t *>(
                K.buf[K.rank] + PEER_DATA_OFF + ((par + (size_t)g) * SB_PEER_MAX_OUT + (size_t)o) * PEER_SLOT);
#pragma unroll
            for (int i = 0; i < W; ++i) u.w[i] = src[i];
            tot = g == 0 ? u.v : red_apply<AT>(K.op, tot, u.v);
        }
        int64_t off = 0;
        int rem = o;
        for (int d = 0; d < K.nkept; ++d) {
            const int c = rem % (int)K.kdims[d];
            rem /= (int)K.kdims[d];
            off += (int64_t)c * K.kstr_bytes[d];
        }
        AT x = load_elem<AT, false>(K.out + off, K.out_dtype, K.out_conj);
        x = init_apply<AT>(K.initop, K.init_re, K.init_im, x);
        store_elem<AT, false>(K.out + off, K.out_dtype, K.out_conj, red_apply<AT>(K.op, x, tot));
    }
    } // (wide elements)
}
} // namespace

extern "C" int sb_peer_export(sb_ctx *ctx, unsigned char handle_out[SB_IPC_HANDLE_BYTES])
{
    if (!ctx || !handle_out) return set_err(ctx, SB_E_INVALID, "sb_peer_export: null argument");
    std::lock_guard<std::mutex> lk(ctx->mu);
    cudaSetDevice(ctx->device);
    if (!ctx->peer_local) {
        cudaError_t e = cudaMalloc(&ctx->peer_local, PEER_BUF_BYTES);
        if (e == cudaSuccess) e = cudaMalloc(&ctx->peer_tmp, (size_t)SB_PEER_MAX_OUT * PEER_SLOT);
        if (e == cudaSuccess) e = cudaMalloc((void **)&ctx->peer_epoch_dev, 256);
        void *hp = nullptr;
        if (e == cudaSuccess) e = cudaHostAlloc(&hp, 64, cudaHostAllocMapped);
        if (e == cudaSuccess) {
            ctx->peer_err_host = (volatile uint32_t *)hp;
            *ctx->peer_err_host = 0u;
            e = cudaHostGetDevicePointer((void **)&ctx->peer_err_dev, hp, 0);
        }
        if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaMalloc(peer buffer)");
        // how long a rank waits for its peers before it gives up (flag + SB_E_CUDA, the context stays usable): ranks may
        // legitimately be seconds apart (first-call plan builds, NVRTC compiles, host stalls) -- NCCL tolerates that too
        const char *tm = std::getenv("SB_PEER_TIMEOUT_MS");
        const double ms = tm ? std::atof(tm) : 30000.0;
        int khz = 0;
        if (cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, ctx->device) != cudaSuccess || khz <= 0) khz = 2000000;
        ctx->peer_timeout_cycles = (int64_t)(ms * (double)khz);
    }
    cudaError_t e = cudaMemset(ctx->peer_local, 0, PEER_BUF_BYTES);
    if (e == cudaSuccess) e = cudaMemset(ctx->peer_epoch_dev, 0, 256);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaMemset(peer buffer)");
    cudaIpcMemHandle_t h;
    static_assert(sizeof(cudaIpcMemHandle_t) == SB_IPC_HANDLE_BYTES, "CUDA IPC handle size");
    e = cudaIpcGetMemHandle(&h, ctx->peer_local);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaIpcGetMemHandle");
    std::memcpy(handle_out, &h, sizeof h);
    return SB_OK;
}

extern "C" int sb_peer_detach(sb_ctx *ctx)
{
    if (!ctx) return SB_OK;
    for (int g = 0; g < SB_PEER_MAX_WORLD; ++g) {
        if (ctx->peer_buf[g] && ctx->peer_buf[g] != ctx->peer_local) cudaIpcCloseMemHandle(ctx->peer_buf[g]);
        ctx->peer_buf[g] = nullptr;
    }
    ctx->peer_world = 1;
    ctx->peer_rank = 0;
    return SB_OK;
}

extern "C" int sb_peer_attach(sb_ctx *ctx, int rank, int world, const unsigned char *handles)
{
    if (!ctx || !handles) return set_err(ctx, SB_E_INVALID, "sb_peer_attach: null argument");
    if (world < 1 || world > SB_PEER_MAX_WORLD || rank < 0 || rank >= world) return set_err(ctx, SB_E_INVALID, "sb_peer_attach: bad rank/world");
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!ctx->peer_local) return set_err(ctx, SB_E_INVALID, "sb_peer_attach: call sb_peer_export first");
    cudaSetDevice(ctx->device);
    sb_peer_detach(ctx);
    for (int g = 0; g < world; ++g) {
        if (g == rank) {
            ctx->peer_buf[g] = ctx->peer_local;
            continue;
        }
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handles + (size_t)g * SB_IPC_HANDLE_BYTES, sizeof h);
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            sb_peer_detach(ctx);
            return cuda_fail(ctx, e, "cudaIpcOpenMemHandle (peer exchange buffer)");
        }
        ctx->peer_buf[g] = p;
    }
    ctx->peer_rank = rank;
    ctx->peer_world = world;
    // (the call counter and the slots are reset TOGETHER, by sb_peer_export only: a re-attach without a fresh export keeps
    //  counting where the previous group stopped, so stale slots can never match a new epoch)
    *ctx->peer_err_host = 0u;
    return SB_OK;
}

extern "C" int sb_mapreduce_allreduce(sb_ctx *ctx, const sb_desc *desc)
{
    if (!ctx) return set_err(nullptr, SB_E_INVALID, "sb_mapreduce_allreduce: null ctx");
    if (!desc) return set_err(ctx, SB_E_INVALID, "sb_mapreduce_allreduce: null desc");
    if (ctx->peer_world <= 1) return sb_mapreduce(ctx, desc);
    std::lock_guard<std::mutex> lk(ctx->mu);
    const sb_desc &d = *desc;
    if (d.ndim < 0 || d.ndim > SB_MAX_DIMS || d.nops < 2 || d.nops > SB_MAX_OPS) return set_err(ctx, SB_E_INVALID, "bad ndim/nops");
    if (d.op < SB_OP_ADD || d.op > SB_OP_MAX) return set_err(ctx, SB_E_INVALID, "sb_mapreduce_allreduce needs a reduction operator");
    if (d.dtype[0] < SB_F32 || d.dtype[0] > SB_C64) return set_err(ctx, SB_E_INVALID, "bad dtype");
    PeerKernelParams K;
    std::memset(&K, 0, sizeof K);
    int64_t nout = 1;
    bool empty = false;
    sb_desc loc = d; // the local reduction into the dense temporary, started from the neutral element
    // kept dims in ascending |output stride| (the planner's canonical order): the logical output index used on the wire
    // is the same whether a rank takes the fused or the two-kernel path
    int kd[SB_MAX_DIMS], nk = 0;
    for (int i = 0; i < d.ndim; ++i) {
        if (d.dims[i] < 0) return set_err(ctx, SB_E_SHAPE, "negative dim");
        if (d.dims[i] == 0) empty = true;
        loc.strides[0][i] = 0;
        if (d.strides[0][i] != 0 && d.dims[i] != 1) kd[nk++] = i;
    }
    std::stable_sort(kd, kd + nk, [&](int a, int b) {
        const int64_t sa = d.strides[0][a] < 0 ? -d.strides[0][a] : d.strides[0][a], sb_ = d.strides[0][b] < 0 ? -d.strides[0][b] : d.strides[0][b];
        return sa < sb_;
    });
    for (int q = 0; q < nk; ++q) {
        const int i = kd[q];
        loc.strides[0][i] = nout;
        K.kdims[K.nkept] = d.dims[i];
        K.kstr_bytes[K.nkept] = d.strides[0][i] * (int64_t)dtype_size(d.dtype[0]);
        K.nkept++;
        nout *= d.dims[i];
    }
    if (nout > SB_PEER_MAX_OUT) return set_err(ctx, SB_E_UNSUPPORTED, "sb_mapreduce_allreduce: more than SB_PEER_MAX_OUT outputs");
    cudaSetDevice(ctx->device);
    PeerLink link;
    std::memset(&link, 0, sizeof link);
    link.world = ctx->peer_world;
    link.rank = ctx->peer_rank;
    link.epoch_ptr = ctx->peer_epoch_dev;
    link.err_flag = ctx->peer_err_dev;
    link.timeout_cycles = ctx->peer_timeout_cycles;
    for (int g = 0; g < ctx->peer_world; ++g) link.buf[g] = (unsigned char *)ctx->peer_buf[g];
    if (!empty) { // preferred: the exchange fused into the reduction kernel (one launch)
        bool fused = false;
        const int rc = run_desc(ctx, d, &link, &fused);
        if (rc != SB_OK) return rc;
        if (fused) {
            if (ctx->sync) {
                cudaError_t e = cudaStreamSynchronize(ctx->stream);
                if (e != cudaSuccess) return cuda_fail(ctx, e, "sb_mapreduce_allreduce");
                return peer_check(ctx);
            }
            return SB_OK;
        }
    }
    const double neutral = d.op == SB_OP_ADD ? 0.0 : d.op == SB_OP_MUL ? 1.0 : d.op == SB_OP_MIN ? (double)INFINITY : -(double)INFINITY;
    if (!empty) {
        loc.base[0] = ctx->peer_tmp;
        loc.conj[0] = 0;
        loc.initop = SB_INIT_CONST;
        loc.init_re = neutral;
        loc.init_im = 0.0;
        const int rc = run_desc(ctx, loc);
        if (rc != SB_OK) return rc; // (every rank sees the same program and dtypes: a planning error is collective)
    }
    K.world = ctx->peer_world;
    K.rank = ctx->peer_rank;
    K.nout = (int32_t)nout;
    K.op = d.op;
    K.initop = d.initop;
    K.init_re = d.init_re;
    K.init_im = d.init_im;
    K.local_empty = empty ? 1 : 0;
    K.out_dtype = d.dtype[0];
    K.out_conj = d.conj[0] && (d.dtype[0] == SB_C32 || d.dtype[0] == SB_C64);
    K.epoch_ptr = link.epoch_ptr;
    K.err_flag = link.err_flag;
    K.timeout_cycles = link.timeout_cycles;
    for (int g = 0; g < ctx->peer_world; ++g) K.buf[g] = link.buf[g];
    K.tmp = (const unsigned char *)ctx->peer_tmp;
    K.out = (unsigned char *)d.base[0];
    cudaError_t e;
    switch (d.dtype[0]) {
    case SB_F32: e = launch_pdl(peer_allreduce_kernel<float>, 1, 256, 0, ctx->stream, K); break;
    case SB_F64: e = launch_pdl(peer_allreduce_kernel<double>, 1, 256, 0, ctx->stream, K); break;
    case SB_C32: e = launch_pdl(peer_allreduce_kernel<cx<float>>, 1, 256, 0, ctx->stream, K); break;
    default: e = launch_pdl(peer_allreduce_kernel<cx<double>>, 1, 256, 0, ctx->stream, K); break;
    }
    if (e != cudaSuccess) return cuda_fail(ctx, e, "peer_allreduce launch");
    ctx->stats.launches++;
    if (ctx->sync) {
        e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) return cuda_fail(ctx, e, "sb_mapreduce_allreduce");
        return peer_check(ctx);
    }
    return SB_OK;
}

// ---- host-resident operands ---------------------------------------------------------------------------------
// Byte range touched by operand k: [lo, hi) relative to base[k].
static void operand_range(const sb_desc &d, int k, int64_t &lo, int64_t &hi)
{
    const int es = dtype_size(d.dtype[k]);
    int64_t mn = 0, mx = 0;
    for (int i = 0; i < d.ndim; ++i) {
        const int64_t ext = (d.dims[i] - 1) * d.strides[k][i];
        if (ext < 0) mn += ext;
        else mx += ext;
    }
    lo = mn * es;
    hi = (mx + 1) * es;
}

extern "C" int sb_mapreduce_host(sb_ctx *ctx, const sb_desc *desc)
{
    if (!ctx) return set_err(nullptr, SB_E_INVALID, "sb_mapreduce_host: null ctx");
    if (!desc) return set_err(ctx, SB_E_INVALID, "sb_mapreduce_host: null desc");
    std::lock_guard<std::mutex> lk(ctx->mu);
    const sb_desc &d = *desc;
    if (d.ndim < 0 || d.ndim > SB_MAX_DIMS || d.nops < 1 || d.nops > SB_MAX_OPS) return set_err(ctx, SB_E_INVALID, "bad ndim/nops");
    for (int i = 0; i < d.ndim; ++i) {
        if (d.dims[i] < 0) return set_err(ctx, SB_E_SHAPE, "negative dim");
        if (d.dims[i] == 0) { // nothing is read; only the (rare) initop-on-empty case touches the output
            sb_desc E;
            if (!empty_initop_desc(d, E)) return SB_OK;
            break;
        }
    }
    for (int k = 0; k < d.nops; ++k)
        if (d.dtype[k] < SB_F32 || d.dtype[k] > SB_C64) return set_err(ctx, SB_E_INVALID, "bad dtype");
    // 1. host byte ranges per operand; merge overlapping ranges (aliased views of one parent) into segments
    struct Seg {
        uintptr_t lo, hi;
        bool has_in, has_out;
        size_t dev_off;
    };
    std::vector<Seg> segs;
    uintptr_t olo[SB_MAX_OPS], ohi[SB_MAX_OPS];
    bool any_zero = false;
    for (int i = 0; i < d.ndim; ++i) any_zero |= d.dims[i] == 0;
    for (int k = 0; k < d.nops; ++k) {
        int64_t lo, hi;
        if (any_zero) {
            lo = 0;
            hi = 0;
            if (k == 0) { // output range of the initop-on-empty case: use kept dims only
                sb_desc E;
                empty_initop_desc(d, E);
                operand_range(E, 0, lo, hi);
            }
        } else
            operand_range(d, k, lo, hi);
        olo[k] = (uintptr_t)d.base[k] + lo;
        ohi[k] = (uintptr_t)d.base[k] + hi;
        if (ohi[k] > olo[k]) segs.push_back(Seg{olo[k], ohi[k], k > 0, k == 0, 0});
    }
    std::sort(segs.begin(), segs.end(), [](const Seg &a, const Seg &b) { return a.lo < b.lo; });
    std::vector<Seg> merged;
    for (const Seg &s : segs) {
        if (!merged.empty() && s.lo < merged.back().hi) {
            merged.back().hi = std::max(merged.back().hi, s.hi);
            merged.back().has_in |= s.has_in;
            merged.back().has_out |= s.has_out;
        } else
            merged.push_back(s);
    }
    size_t total = 0;
    for (Seg &s : merged) {
        s.dev_off = total;
        total += ((s.hi - s.lo) + 255) & ~(size_t)255;
    }
    cudaSetDevice(ctx->device);
    // does the output need its previous contents on the device?  yes if it is read (op / initop) or if
    // the written elements do not cover its byte range densely (strided destination)
    bool out_dense = !any_zero;
    if (out_dense) {
        int64_t cnt = 1;
        for (int i = 0; i < d.ndim; ++i)
            if (d.strides[0][i] != 0) cnt *= d.dims[i];
        out_dense = (uintptr_t)(cnt * dtype_size(d.dtype[0])) == (ohi[0] - olo[0]);
    }
    // ---- zero-copy mode ------------------------------------------------------------------------------------------
    // A synchronous call on PINNED host operands: the kernel reads its inputs straight from host memory and writes the
    // result straight into host memory (unified addressing), so the two directions of the host link overlap inside ONE
    // launch instead of running H2D copy -> kernel -> D2H copy back to back.  Measured on config 2 (profiles/
    // r02_a_exp.txt): 3.56 ms against 4.62 ms staged.  Taken only when every input byte crosses the link once: maps
    // (no read of the output), inputs that do not alias each other -- or alias as permuted views of one parent, which
    // the orbit kernel fetches once -- and no broadcast (zero-stride) re-reads.  Stream-ordered (sync == 0) callers keep
    // the staged path: pipelining consecutive calls on the copy engines gets closer to the link rate than SM-issued
    // stores do (94 vs 72 GB/s on config 2).
    if (ctx->sync && ctx->host_zero_copy && d.op == SB_OP_NONE && out_dense && d.nops >= 2 && !merged.empty()) {
        bool ok = true, alias = false;
        for (int k = 1; k < d.nops && ok; ++k)
            for (int i = 0; i < d.ndim; ++i)
                if (d.strides[k][i] == 0 && d.dims[i] > 1) ok = false;
        for (const Seg &s0 : merged) {
            int nin_here = 0;
            for (int k = 1; k < d.nops; ++k)
                if (olo[k] >= s0.lo && olo[k] < s0.hi) ++nin_here;
            if (nin_here > 1) alias = true;
            if (s0.has_in && s0.has_out) ok = false; // in-place: keep the staged path
        }
        std::vector<uintptr_t> devlo(merged.size(), 0);
        for (size_t q = 0; q < merged.size() && ok; ++q) {
            cudaPointerAttributes a0, a1;
            if (cudaPointerGetAttributes(&a0, (const void *)merged[q].lo) != cudaSuccess ||
                cudaPointerGetAttributes(&a1, (const void *)(merged[q].hi - 1)) != cudaSuccess) {
                cudaGetLastError();
                ok = false;
                break;
            }
            if (a0.type != cudaMemoryTypeHost || a1.type != cudaMemoryTypeHost || !a0.devicePointer || !a1.devicePointer ||
                (uintptr_t)a1.devicePointer - (uintptr_t)a0.devicePointer != merged[q].hi - 1 - merged[q].lo)
                ok = false;
            else
                devlo[q] = (uintptr_t)a0.devicePointer;
        }
        if (ok) {
            sb_desc dz = d;
            for (int k = 0; k < d.nops; ++k)
                for (size_t q = 0; q < merged.size(); ++q)
                    if (olo[k] >= merged[q].lo && olo[k] < merged[q].hi) dz.base[k] = (void *)(devlo[q] + ((uintptr_t)d.base[k] - merged[q].lo));
            if (alias) { // only worth it if the aliased views are fetched once (orbit plan)
                PlanIt hit;
                const int rc0 = lookup_plan(ctx, dz, true, hit);
                if (rc0 != SB_OK || !hit->second.plan.orbit_ok) ok = false;
            }
            if (ok) {
                const int rc1 = run_desc(ctx, dz, nullptr, nullptr, true);
                if (rc1 != SB_OK) return rc1;
                for (const Seg &s0 : merged) {
                    if (s0.has_in) ctx->stats.h2d_bytes += s0.hi - s0.lo;
                    if (s0.has_out) ctx->stats.d2h_bytes += s0.hi - s0.lo;
                }
                ctx->stats.zero_copy_calls++;
                cudaError_t e = cudaStreamSynchronize(ctx->stream);
                if (e != cudaSuccess) return cuda_fail(ctx, e, "sb_mapreduce_host (zero-copy)");
                return SB_OK;
            }
        }
    }
    if (total > ctx->stage_bytes) {
        if (ctx->stage) {
            cudaStreamSynchronize(ctx->stream);
            cudaFree(ctx->stage);
            ctx->stage = nullptr;
            ctx->stage_bytes = 0;
        }
        cudaError_t e = cudaMalloc(&ctx->stage, total);
        if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaMalloc(stage)");
        ctx->stage_bytes = total;
    }
    // 2. stage the inputs (and the output's previous contents when they are read, see out_dense above)
    const bool out_read = d.op != SB_OP_NONE || !out_dense;
    for (const Seg &s : merged) {
        if (!(s.has_in || (s.has_out && out_read))) continue;
        cudaError_t e = cudaMemcpyAsync((char *)ctx->stage + s.dev_off, (const void *)s.lo, s.hi - s.lo, cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) return cuda_fail(ctx, e, "stage h2d");
        ctx->stats.h2d_bytes += s.hi - s.lo;
    }
    // 3. run on device pointers
    sb_desc dd = d;
    for (int k = 0; k < d.nops; ++k) {
        const Seg *home = nullptr;
        for (const Seg &s : merged)
            if (olo[k] >= s.lo && olo[k] < s.hi) home = &s;
        if (!home) { // empty range (zero-size problem)
            dd.base[k] = ctx->stage;
            continue;
        }
        dd.base[k] = (char *)ctx->stage + home->dev_off + ((uintptr_t)d.base[k] - home->lo);
    }
    int rc = run_desc(ctx, dd);
    if (rc != SB_OK) return rc;
    // 4. output back
    for (const Seg &s : merged) {
        if (!s.has_out) continue;
        cudaError_t e = cudaMemcpyAsync((void *)s.lo, (char *)ctx->stage + s.dev_off, s.hi - s.lo, cudaMemcpyDeviceToHost, ctx->stream);
        if (e != cudaSuccess) return cuda_fail(ctx, e, "stage d2h");
        ctx->stats.d2h_bytes += s.hi - s.lo;
    }
    if (ctx->sync) { // sync == 0: stream-ordered (H2D, kernel, D2H are all enqueued); the caller syncs with sb_sync
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) return cuda_fail(ctx, e, "sb_mapreduce_host");
    }
    return SB_OK;
}
