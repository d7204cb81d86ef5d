// jit.cu -- see jit.hpp.  libnvrtc is dlopen'ed (no link-time dependency).
#include "jit.hpp"

#include <dlfcn.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <mutex>
#include <sstream>
#include <chrono>
#include <thread>
#include <vector>

namespace sb {

namespace {

typedef struct _nvrtcProgram *nvrtcProgram;
typedef int nvrtcResult;
struct Nvrtc {
    void *h = nullptr;
    nvrtcResult (*CreateProgram)(nvrtcProgram *, const char *, const char *, int, const char *const *, const char *const *) = nullptr;
    nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char *const *) = nullptr;
    nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t *) = nullptr;
    nvrtcResult (*GetCUBIN)(nvrtcProgram, char *) = nullptr;
    nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t *) = nullptr;
    nvrtcResult (*GetProgramLog)(nvrtcProgram, char *) = nullptr;
    nvrtcResult (*DestroyProgram)(nvrtcProgram *) = nullptr;
    nvrtcResult (*Version)(int *, int *) = nullptr;
    bool ok = false;
};

std::mutex g_mu;
std::string g_log;

Nvrtc &nvrtc()
{
    static Nvrtc n = []() {
        Nvrtc r;
        const char *names[] = {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so"};
        for (const char *nm : names) {
            r.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (r.h) break;
        }
        if (!r.h) return r;
#define LOADSYM(field, sym) *(void **)(&r.field) = dlsym(r.h, sym)
        LOADSYM(CreateProgram, "nvrtcCreateProgram");
        LOADSYM(CompileProgram, "nvrtcCompileProgram");
        LOADSYM(GetCUBINSize, "nvrtcGetCUBINSize");
        LOADSYM(GetCUBIN, "nvrtcGetCUBIN");
        LOADSYM(GetProgramLogSize, "nvrtcGetProgramLogSize");
        LOADSYM(GetProgramLog, "nvrtcGetProgramLog");
        LOADSYM(DestroyProgram, "nvrtcDestroyProgram");
        LOADSYM(Version, "nvrtcVersion");
#undef LOADSYM
        r.ok = r.CreateProgram && r.CompileProgram && r.GetCUBINSize && r.GetCUBIN && r.GetProgramLogSize && r.GetProgramLog && r.DestroyProgram;
        return r;
    }();
    return n;
}

std::string csrc_dir()
{
    Dl_info info;
    if (dladdr((const void *)&jit_last_log, &info) && info.dli_fname) {
        std::string p(info.dli_fname);
        const size_t k = p.find_last_of('/');
        return (k == std::string::npos ? std::string(".") : p.substr(0, k)) + "/csrc";
    }
    return "csrc";
}

uint64_t fnv64(const std::string &s)
{
    uint64_t h = 1469598103934665603ull;
    for (unsigned char c : s) {
        h ^= c;
        h *= 1099511628211ull;
    }
    return h;
}

const char *ct_name(int ct) { return ct == F32 ? "float" : ct == F64 ? "double" : ct == C32 ? "sb::cx<float>" : "sb::cx<double>"; }

// straight-line C++ for the postfix program; constants stay run-time parameters (p.tok[i].re/im)
bool gen_functor(const Program &prog, std::ostringstream &os)
{
    os << "namespace sb {\ntemplate <class CT> struct ElemFn<CT, RC_JIT> {\n"
          "    template <int NIN> SB_HD CT eval(const Program &p, const CT *a) const\n    {\n";
    std::vector<int> st;
    int nv = 0;
    if (prog.ntok == 0) {
        os << "        return a[0];\n    }\n};\n}\n";
        return true;
    }
    for (int i = 0; i < prog.ntok; ++i) {
        const Tok &t = prog.tok[i];
        if (t.kind == TOK_ARG) {
            os << "        const CT v" << nv << " = a[" << t.a << "];\n";
            st.push_back(nv++);
        } else if (t.kind == TOK_CONST) {
            os << "        const CT v" << nv << " = make<CT>(p.tok[" << i << "].re, p.tok[" << i << "].im);\n";
            st.push_back(nv++);
        } else if (t.a < 32) {
            if (st.empty()) return false;
            const int x = st.back();
            st.pop_back();
            os << "        const CT v" << nv << " = call1(" << t.a << ", v" << x << ");\n";
            st.push_back(nv++);
        } else {
            if (st.size() < 2) return false;
            const int y = st.back();
            st.pop_back();
            const int x = st.back();
            st.pop_back();
            os << "        const CT v" << nv << " = call2<CT>(" << t.a << ", v" << x << ", v" << y << ");\n";
            st.push_back(nv++);
        }
    }
    if (st.size() != 1) return false;
    os << "        return v" << st.back() << ";\n    }\n};\n}\n";
    return true;
}

std::string structure_key(int kind, const KernelKey &k, const Program &prog)
{
    std::ostringstream os;
    os << kind << ":" << k.ct << ":" << k.nin << ":" << k.ept << ":" << k.uniform << ":";
    for (int i = 0; i < prog.ntok; ++i) os << prog.tok[i].kind << "," << (prog.tok[i].kind == TOK_CONST ? 0 : prog.tok[i].a) << ";";
    return os.str();
}

std::string cache_dir()
{
    if (const char *e = std::getenv("SB_JIT_CACHE")) return e;
    const char *home = std::getenv("HOME");
    if (home && *home) return std::string(home) + "/.cache/strided_b200";
    return "/tmp/strided_b200-" + std::to_string((unsigned long)geteuid()); // no HOME: a per-user directory, checked below
}

// The on-disk cache is only used inside a directory that belongs to this user and that nobody else can write to (it is
// created 0700, parents as needed); otherwise kernels are compiled per process and nothing is read from or written to disk.
bool private_dir(const std::string &dir)
{
    for (size_t k = 1; k <= dir.size(); ++k) // mkdir -p
        if (k == dir.size() || dir[k] == '/') {
            const std::string part = dir.substr(0, k);
            if (!part.empty()) mkdir(part.c_str(), k == dir.size() ? 0700 : 0755);
        }
    struct stat st;
    if (lstat(dir.c_str(), &st) != 0 || !S_ISDIR(st.st_mode)) return false;
    return st.st_uid == geteuid() && (st.st_mode & (S_IWGRP | S_IWOTH)) == 0;
}

// cache file = 8-byte magic, 8-byte FNV-1a of the payload, payload: a truncated or foreign file is rejected, not loaded
const char kMagic[8] = {'S', 'B', 'J', 'I', 'T', '0', '0', '2'};
bool read_cache_file(const std::string &path, std::vector<char> &cubin)
{
    const int fd = open(path.c_str(), O_RDONLY | O_NOFOLLOW | O_CLOEXEC);
    if (fd < 0) return false;
    struct stat st;
    bool ok = fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_uid == geteuid() && st.st_size > 16 && st.st_size < (64 << 20);
    std::vector<char> buf;
    if (ok) {
        buf.resize((size_t)st.st_size);
        size_t got = 0;
        while (got < buf.size()) {
            const ssize_t r = read(fd, buf.data() + got, buf.size() - got);
            if (r <= 0) break;
            got += (size_t)r;
        }
        ok = got == buf.size();
    }
    close(fd);
    if (!ok || std::memcmp(buf.data(), kMagic, 8) != 0) return false;
    uint64_t want;
    std::memcpy(&want, buf.data() + 8, 8);
    if (fnv64(std::string(buf.data() + 16, buf.size() - 16)) != want) return false;
    cubin.assign(buf.begin() + 16, buf.end());
    return true;
}
void write_cache_file(const std::string &path, const std::vector<char> &cubin)
{
    const std::string tmp = path + ".tmp" + std::to_string((long)getpid());
    const int fd = open(tmp.c_str(), O_WRONLY | O_CREAT | O_EXCL | O_NOFOLLOW | O_CLOEXEC, 0600);
    if (fd < 0) return;
    const uint64_t h = fnv64(std::string(cubin.data(), cubin.size()));
    bool ok = write(fd, kMagic, 8) == 8 && write(fd, &h, 8) == 8;
    size_t put = 0;
    while (ok && put < cubin.size()) {
        const ssize_t r = write(fd, cubin.data() + put, cubin.size() - put);
        if (r <= 0) ok = false;
        else put += (size_t)r;
    }
    close(fd);
    if (ok) std::rename(tmp.c_str(), path.c_str());
    else unlink(tmp.c_str());
}

bool load_cubin(const std::vector<char> &cubin, JitKernel &out)
{
    cudaLibrary_t lib = nullptr;
    if (cudaLibraryLoadData(&lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    cudaKernel_t fn = nullptr;
    if (cudaLibraryGetKernel(&fn, lib, "sb_jit_kernel") != cudaSuccess) {
        cudaGetLastError();
        cudaLibraryUnload(lib);
        return false;
    }
    out.lib = lib;
    out.fn = fn;
    return true;
}

} // namespace

const char *jit_last_log() { return g_log.c_str(); }

bool jit_enabled()
{
    static const bool on = !std::getenv("SB_NO_JIT");
    return on && nvrtc().ok;
}

// ---- compile jobs ------------------------------------------------------------------------------------------------------
// NVRTC takes ~2 s per kernel.  By default the compile runs on a worker thread while the caller keeps using the in-kernel
// interpreter (same arithmetic, bit-identical results: tests/test_gpu_jit.py); the specialised kernel takes over at the
// first call after the cubin is ready.  `wait` (plans that cannot run on the interpreter, SB_JIT_SYNC=1) blocks instead.
struct Job {
    int state = 0; // 0 compiling, 1 cubin ready (not loaded yet), 2 loaded, 3 failed
    std::vector<char> cubin;
    std::string log;
    int min_blocks = 1;
    JitKernel kernel;
};
std::map<std::string, Job> g_jobs;
std::vector<std::thread> g_workers;
std::mutex g_workers_mu;

// Waits for the compiles that are still running on worker threads.  A process that exits while NVRTC is compiling would
// tear libnvrtc's own static state down under the worker (libnvrtc is dlopen'ed AFTER this library, so its destructors run
// BEFORE ours): the handler is therefore registered with atexit at every spawn -- later than anything NVRTC has registered
// up to then -- and hosts call it explicitly through sb_shutdown (Python: atexit hook in abi.py; Julia: atexit in the glue).
void jit_join_workers()
{
    std::vector<std::thread> ws;
    {
        std::lock_guard<std::mutex> lk(g_workers_mu);
        ws.swap(g_workers);
    }
    for (auto &t : ws)
        if (t.joinable()) t.join();
}

struct WorkerJoin {
    ~WorkerJoin() { jit_join_workers(); }
} g_worker_join;

// everything that does not need a CUDA context: source generation, disk cache, NVRTC
static void compile_job(const std::string skey, int kind, const KernelKey key, const Program prog)
{
    std::vector<char> cubin;
    std::string log;
    bool ok = false;
    const int words = key.nin * key.ept * (dtype_size(key.ct) / 4);
    const int minb = words <= 32 ? 4 : (words <= 64 ? 2 : 1);
    do {
        std::ostringstream src;
        src << "#include \"functors.hpp\"\n";
        if (!gen_functor(prog, src)) {
            log = "jit: malformed program";
            break;
        }
        src << "#include \"kernel_bodies.cuh\"\n";
        const char *ct = ct_name(key.ct);
        if (kind == JIT_MAP)
            src << "extern \"C\" __global__ void __launch_bounds__(" << THREADS << ", " << minb << ") sb_jit_kernel(const __grid_constant__ sb::MapParams P)\n{\n"
                << "    sb::map_tile_body<" << ct << ", sb::RC_JIT, " << key.nin << ", " << key.ept << ", " << (key.uniform ? "true" : "false") << ">(P);\n}\n";
        else
            src << "extern \"C\" __global__ void __launch_bounds__(" << THREADS << ", " << minb << ") sb_jit_kernel(const __grid_constant__ sb::ReduceParams P)\n{\n"
                << "    sb::reduce_tile_body<" << ct << ", sb::RC_JIT, " << key.nin << ", " << key.ept << ", " << (key.uniform ? "true" : "false") << ">(P);\n}\n";
        const std::string source = src.str();
        const std::string inc1 = "-I" + csrc_dir();
        const char *cinc = std::getenv("SB_CUDA_INCLUDE");
        const std::string inc2 = std::string("-I") + (cinc ? cinc : "/usr/local/cuda/include");
        int vmaj = 0, vmin = 0;
        if (nvrtc().Version) nvrtc().Version(&vmaj, &vmin);
        // disk cache: the kernel bodies are part of the hash through their modification times
        std::ostringstream hk;
        hk << source << inc1 << vmaj << "." << vmin;
        for (const char *f : {"/common.hpp", "/elem.hpp", "/functors.hpp", "/map_tile.hpp", "/reduce_tile.hpp", "/kernel_bodies.cuh"}) {
            struct stat stt;
            if (stat((csrc_dir() + f).c_str(), &stt) == 0) hk << stt.st_mtime << ":" << stt.st_size << ";";
        }
        char fname[64];
        std::snprintf(fname, sizeof fname, "/%016llx.cubin", (unsigned long long)fnv64(hk.str()));
        const std::string dir = cache_dir(), path = dir + fname;
        const bool use_disk = private_dir(dir);
        if (use_disk && read_cache_file(path, cubin)) {
            ok = true;
            break;
        }
        nvrtcProgram pr = nullptr;
        if (nvrtc().CreateProgram(&pr, source.c_str(), "sb_jit.cu", 0, nullptr, nullptr) != 0) {
            log = "jit: nvrtcCreateProgram failed";
            break;
        }
        const char *opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "--fmad=false", "-lineinfo", inc1.c_str(), inc2.c_str()};
        const nvrtcResult rc = nvrtc().CompileProgram(pr, 6, opts);
        if (rc != 0) {
            size_t n = 0;
            nvrtc().GetProgramLogSize(pr, &n);
            std::string l(n, '\0');
            if (n) nvrtc().GetProgramLog(pr, &l[0]);
            log = "jit: compile failed: " + l.substr(0, 2000);
            nvrtc().DestroyProgram(&pr);
            if (std::getenv("SB_JIT_VERBOSE")) std::fprintf(stderr, "%s\n%s\n", log.c_str(), source.c_str());
            break;
        }
        size_t sz = 0;
        nvrtc().GetCUBINSize(pr, &sz);
        cubin.resize(sz);
        nvrtc().GetCUBIN(pr, cubin.data());
        nvrtc().DestroyProgram(&pr);
        if (use_disk) write_cache_file(path, cubin);
        ok = true;
    } while (false);
    std::lock_guard<std::mutex> lk(g_mu);
    Job &j = g_jobs[skey];
    j.min_blocks = minb;
    if (ok) {
        j.cubin = std::move(cubin);
        j.state = 1;
    } else {
        j.log = log;
        g_log = log;
        j.state = 3;
    }
}

const JitKernel *jit_get(int kind, const KernelKey &key, const Program &prog, bool wait)
{
    if (!jit_enabled()) return nullptr;
    const std::string skey = structure_key(kind, key, prog);
    std::unique_lock<std::mutex> lk(g_mu);
    auto it = g_jobs.find(skey);
    if (it == g_jobs.end()) {
        g_jobs[skey]; // state 0: compiling
        if (wait) {
            lk.unlock();
            compile_job(skey, kind, key, prog);
            lk.lock();
        } else {
            {
                std::lock_guard<std::mutex> wl(g_workers_mu);
                g_workers.emplace_back(compile_job, skey, kind, key, prog);
            }
            std::atexit(jit_join_workers);
            return nullptr; // the interpreter serves this call
        }
        it = g_jobs.find(skey);
    }
    if (it->second.state == 0 && wait) { // somebody else's compile is in flight: wait for it
        while (it->second.state == 0) {
            lk.unlock();
            std::this_thread::sleep_for(std::chrono::milliseconds(2));
            lk.lock();
            it = g_jobs.find(skey);
        }
    }
    Job &j = it->second;
    if (j.state == 1) { // cubin ready: load it into the CALLER's CUDA context (the worker thread never touches CUDA)
        if (load_cubin(j.cubin, j.kernel)) {
            j.kernel.min_blocks = j.min_blocks;
            j.state = 2;
        } else {
            g_log = "jit: cudaLibraryLoadData failed";
            j.state = 3;
        }
        j.cubin.clear();
        j.cubin.shrink_to_fit();
    }
    return j.state == 2 ? &j.kernel : nullptr;
}

} // namespace sb
