// jit.cu -- see jit.hpp.  libnvrtc is dlopen'ed (no link-time dependency).
#include "jit.hpp"

#include <dlfcn.h>
#include <sys/stat.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <mutex>
#include <sstream>
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
std::map<std::string, JitKernel> g_cache;   // key -> kernel (fn == nullptr: failed, do not retry)

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
    return std::string(home ? home : "/tmp") + "/.cache/strided_b200";
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

const JitKernel *jit_get(int kind, const KernelKey &key, const Program &prog)
{
    if (!jit_enabled()) return nullptr;
    std::lock_guard<std::mutex> lk(g_mu);
    const std::string skey = structure_key(kind, key, prog);
    auto it = g_cache.find(skey);
    if (it != g_cache.end()) return it->second.fn ? &it->second : nullptr;
    JitKernel &slot = g_cache[skey]; // fn == nullptr until success

    const int words = key.nin * key.ept * (dtype_size(key.ct) / 4);
    const int minb = words <= 32 ? 4 : (words <= 64 ? 2 : 1);
    std::ostringstream src;
    src << "#include \"functors.hpp\"\n";
    if (!gen_functor(prog, src)) {
        g_log = "jit: malformed program";
        return nullptr;
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
    {
        std::ifstream f(path, std::ios::binary);
        if (f) {
            std::vector<char> cubin((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
            if (!cubin.empty() && load_cubin(cubin, slot)) {
                slot.min_blocks = minb;
                return &slot;
            }
        }
    }
    nvrtcProgram pr = nullptr;
    if (nvrtc().CreateProgram(&pr, source.c_str(), "sb_jit.cu", 0, nullptr, nullptr) != 0) {
        g_log = "jit: nvrtcCreateProgram failed";
        return nullptr;
    }
    const char *opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "--fmad=false", "-lineinfo", inc1.c_str(), inc2.c_str()};
    const nvrtcResult rc = nvrtc().CompileProgram(pr, 6, opts);
    if (rc != 0) {
        size_t n = 0;
        nvrtc().GetProgramLogSize(pr, &n);
        std::string log(n, '\0');
        if (n) nvrtc().GetProgramLog(pr, &log[0]);
        g_log = "jit: compile failed: " + log.substr(0, 2000);
        nvrtc().DestroyProgram(&pr);
        if (std::getenv("SB_JIT_VERBOSE")) std::fprintf(stderr, "%s\n%s\n", g_log.c_str(), source.c_str());
        return nullptr;
    }
    size_t sz = 0;
    nvrtc().GetCUBINSize(pr, &sz);
    std::vector<char> cubin(sz);
    nvrtc().GetCUBIN(pr, cubin.data());
    nvrtc().DestroyProgram(&pr);
    if (!load_cubin(cubin, slot)) {
        g_log = "jit: cudaLibraryLoadData failed";
        return nullptr;
    }
    slot.min_blocks = minb;
    mkdir(dir.c_str(), 0755);
    {
        std::ofstream f(path + ".tmp", std::ios::binary);
        if (f) {
            f.write(cubin.data(), (std::streamsize)cubin.size());
            f.close();
            std::rename((path + ".tmp").c_str(), path.c_str());
        }
    }
    return &slot;
}

} // namespace sb
