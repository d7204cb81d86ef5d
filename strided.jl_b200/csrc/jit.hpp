// jit.hpp -- run-time specialisation of the element function with NVRTC.
//
// Julia compiles one `_mapreduce_kernel!` per callable `f` (it is @generated on the types of f, op, initop; reference
// src/mapreduce.jl:229-234).  The pre-instantiated recipes cover the common trees; for everything else the postfix
// program is turned into straight-line C++ (one statement per token, `call1/call2` with literal function ids, constants
// left as run-time parameters), compiled against the SAME kernel bodies (kernel_bodies.cuh) for sm_100a, cached in
// memory and on disk.  If NVRTC is missing or the compile fails, the caller keeps using the in-kernel interpreter.
#pragma once
#include "planner.hpp"
#include <cuda_runtime.h>
#include <string>

namespace sb {

struct JitKernel {
    cudaKernel_t fn = nullptr;
    cudaLibrary_t lib = nullptr;
    int min_blocks = 1;
};

enum JitKind : int { JIT_MAP = 0, JIT_REDUCE = 1 };

// nullptr: JIT unavailable / failed for this key (the reason is kept in jit_last_log()) -- or, with wait == false, still
// being compiled on a worker thread (the caller runs the interpreter meanwhile and asks again on its next call).
const JitKernel *jit_get(int kind, const KernelKey &key, const Program &prog, bool wait);
const char *jit_last_log();
bool jit_enabled();
void jit_join_workers(); // waits for background compiles (atexit / sb_shutdown)

} // namespace sb
