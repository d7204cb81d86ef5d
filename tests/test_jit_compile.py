"""CPU suite: the source that csrc/jit.cu hands to NVRTC at run time (a generated element functor + the SAME kernel
bodies the static kernels use) must compile for sm_100a.  NVRTC needs no GPU, so a change to kernel_bodies.cuh that
would silently push every non-recipe expression back onto the interpreter is caught here."""
import ctypes as C
import glob
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "strided.jl_b200", "csrc")


def _nvrtc():
    for pat in ("/usr/local/cuda/lib64/libnvrtc.so*", "/usr/local/cuda/targets/*/lib/libnvrtc.so*"):
        for p in sorted(glob.glob(pat)):
            if "builtins" in p:
                continue
            try:
                return C.CDLL(p)
            except OSError:
                pass
    return None


FUNCTOR = """namespace sb {
template <class CT> struct ElemFn<CT, RC_JIT> {
    template <int NIN> SB_HD CT eval(const Program &p, const CT *a) const
    {
        const CT v0 = a[0];
        const CT v1 = call1(10, v0);
        const CT v2 = a[1];
        const CT v3 = make<CT>(p.tok[3].re, p.tok[3].im);
        const CT v4 = call2<CT>(34, v2, v3);
        const CT v5 = call2<CT>(32, v1, v4);
        return v5;
    }
};
}
"""


def _all_functions_functor():
    """one functor that calls EVERY function id of the ABI (sb_fn): whatever program the planner hands to the JIT compiles"""
    unary = list(range(0, 14))
    binary = [32, 33, 34, 35, 36, 37, 38]
    lines = ["        const CT v0 = a[0];", "        const CT v1 = a[1];"]
    n = 2
    for f in unary:
        lines.append(f"        const CT v{n} = call1({f}, v{n - 1});")
        n += 1
    for f in binary:
        lines.append(f"        const CT v{n} = call2<CT>({f}, v{n - 1}, v{n - 2});")
        n += 1
    body = "\n".join(lines)
    return ("namespace sb {\ntemplate <class CT> struct ElemFn<CT, RC_JIT> {\n"
            "    template <int NIN> SB_HD CT eval(const Program &p, const CT *a) const\n    {\n" + body +
            f"\n        return v{n - 1};\n    }}\n}};\n}}\n")


def _source(kind, ct):
    body = ('extern "C" __global__ void __launch_bounds__(256, 2) sb_jit_kernel(const __grid_constant__ sb::%sParams P)\n{\n'
            "    sb::%s_tile_body<%s, sb::RC_JIT, 2, %d, true>(P);\n}\n")
    if kind == "map":
        return '#include "functors.hpp"\n' + FUNCTOR + '#include "kernel_bodies.cuh"\n' + body % ("Map", "map", ct, 8)
    return '#include "functors.hpp"\n' + FUNCTOR + '#include "kernel_bodies.cuh"\n' + body % ("Reduce", "reduce", ct, 8 if ct != "sb::cx<double>" else 4)


@pytest.mark.parametrize("kind", ["map", "reduce"])
@pytest.mark.parametrize("ct", ["float", "double", "sb::cx<float>", "sb::cx<double>"])
def test_runtime_specialised_source_compiles_with_nvrtc(kind, ct):
    lib = _nvrtc()
    if lib is None:
        pytest.skip("libnvrtc not found in this image")
    prog = C.c_void_p()
    src = _source(kind, ct).encode()
    assert lib.nvrtcCreateProgram(C.byref(prog), src, b"sb_jit.cu", 0, None, None) == 0
    opts = [b"--gpu-architecture=sm_100a", b"-std=c++17", b"--fmad=false", b"-lineinfo", b"-I" + CSRC.encode(), b"-I/usr/local/cuda/include"]
    arr = (C.c_char_p * len(opts))(*opts)
    rc = lib.nvrtcCompileProgram(prog, len(opts), arr)
    n = C.c_size_t()
    lib.nvrtcGetProgramLogSize(prog, C.byref(n))
    log = C.create_string_buffer(n.value + 1)
    lib.nvrtcGetProgramLog(prog, log)
    lib.nvrtcDestroyProgram(C.byref(prog))
    assert rc == 0, log.value.decode()[:3000]


@pytest.mark.parametrize("ct", ["float", "double", "sb::cx<float>", "sb::cx<double>"])
def test_every_function_id_compiles_in_a_runtime_functor(ct):
    lib = _nvrtc()
    if lib is None:
        pytest.skip("libnvrtc not found in this image")
    src = ('#include "functors.hpp"\n' + _all_functions_functor() + '#include "kernel_bodies.cuh"\n'
           'extern "C" __global__ void __launch_bounds__(256, 2) sb_jit_kernel(const __grid_constant__ sb::MapParams P)\n{\n'
           "    sb::map_tile_body<%s, sb::RC_JIT, 2, 4, true>(P);\n}\n" % ct).encode()
    prog = C.c_void_p()
    assert lib.nvrtcCreateProgram(C.byref(prog), src, b"sb_jit.cu", 0, None, None) == 0
    opts = [b"--gpu-architecture=sm_100a", b"-std=c++17", b"--fmad=false", b"-I" + CSRC.encode(), b"-I/usr/local/cuda/include"]
    arr = (C.c_char_p * len(opts))(*opts)
    rc = lib.nvrtcCompileProgram(prog, len(opts), arr)
    n = C.c_size_t()
    lib.nvrtcGetProgramLogSize(prog, C.byref(n))
    log = C.create_string_buffer(n.value + 1)
    lib.nvrtcGetProgramLog(prog, log)
    lib.nvrtcDestroyProgram(C.byref(prog))
    assert rc == 0, log.value.decode()[:3000]
