/*
 * strided_b200.h -- C ABI of the B200-native strided map / permute / reduce engine.
 *
 * This is the drop-in boundary for ONE hot path of Jutho/Strided.jl (v2.3.2): everything that
 * funnels into
 *
 *     _mapreduce_fuse!(f, op, initop, dims, arrays)                     src/mapreduce.jl:98-117
 *     _mapreduce_order!(f, op, initop, dims, strides, arrays)           src/mapreduce.jl:119-139
 *     _mapreduce_block!(f, op, initop, dims, strides, offsets, costs, arrays)   :142-180
 *     _mapreduce_threaded!(...)                                         :195-227
 *     _mapreduce_kernel!(f, op, initop, dims, blocks, arrays, strides, offsets) :229-425
 *
 * i.e. `map!`/`copy!`/`permutedims!`/`adjoint!`/`conj!` (mapreduce.jl:2-53), the broadcast
 * `copyto!(dest::StridedView, bc::Broadcasted{StridedArrayStyle})` (broadcast.jl:27-37) and
 * `mapreduce`/`mapreducedim!`/`_mapreducedim!` (mapreduce.jl:16-30, 55-96).
 *
 * The reference has no FFI; its "operator API" is Julia dispatch that reaches the signatures above.
 * A maintainer's Julia glue `ccall`s the entry points below in place of `_mapreduce_block!`
 * (see INTEGRATION.md for the binding).  Plain C types only: pointers, sizes, PODs.
 *
 * Conventions kept from the reference:
 *   - operand 0 is the OUTPUT, operands 1..nops-1 the inputs (arrays[1] / arrays[2:end]);
 *   - all operands have the same rank `ndim` and the same `dims`; broadcast / reduction is expressed
 *     with ZERO strides exactly as `promoteshape1` does (broadcast.jl:56-65);
 *   - strides are in ELEMENTS of the operand's own dtype, any sign, column-major habit but no
 *     ordering assumed (StridedViews.jl layout contract);
 *   - `base[k]` already includes the view offset: it points at element (1,1,...,1) of the view,
 *     i.e. `pointer(parent, offset+1)` (mapreduce.jl:268);
 *   - `op == SB_OP_NONE`  <=> map mode  `A1[I1] = f(A2[I2], ...)`            (mapreduce.jl:311)
 *     otherwise reduce mode `A1[I1] = op(A1[I1], f(A2[I2], ...))`           (mapreduce.jl:314)
 *     with `initop` applied exactly once per distinct output element first   (mapreduce.jl:351-382);
 *   - `conj[k] != 0` <=> the view's `op` is `conj`/`adjoint` on a complex eltype: the element is
 *     conjugated on load, and (operand 0) on store (ParentIndex get/setindex, mapreduce.jl:276-278).
 *
 * Element functions outside the pre-compiled recipes run either through an in-kernel interpreter or, for problems of
 * at least SB_JIT_MIN_ELEMENTS elements (default 2^18), through a kernel specialised at run time with NVRTC (compiled on a
 * worker thread while the interpreter keeps serving the calls; SB_JIT_SYNC=1 blocks instead) and cached on disk
 * (~/.cache/strided_b200, private to the user); SB_NO_JIT=1 disables the latter.
 *
 * Error behaviour: no entry point throws or aborts; every one returns an `sb_status`.  The glue maps
 * SB_E_SHAPE -> DimensionMismatch (mapreduce.jl:43-46, broadcast.jl:61), SB_E_UNSUPPORTED -> fall
 * back to the original CPU method, anything else -> ErrorException(sb_last_error(ctx)).
 */
#ifndef STRIDED_B200_H
#define STRIDED_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB_ABI_VERSION 1
#define SB_MAX_DIMS 8   /* rank after the caller's own fusing; Strided.jl tests use N <= 6        */
#define SB_MAX_OPS 8    /* output + up to 7 inputs                                                  */
#define SB_MAX_TOKENS 48

typedef enum sb_status {
    SB_OK = 0,
    SB_E_INVALID = -1,     /* null pointer, bad enum, ndim/nops out of range, malformed program     */
    SB_E_SHAPE = -2,       /* negative dim                                    -> DimensionMismatch  */
    SB_E_UNSUPPORTED = -3, /* legal in the reference but not on the device path -> CPU fallback      */
    SB_E_CUDA = -4,        /* a CUDA runtime/driver call failed; see sb_last_error                  */
    SB_E_NOMEM = -5,
    SB_E_NODEVICE = -6     /* no CUDA device / driver: the product path fails loudly, never on CPU  */
} sb_status;

/* element types of StridedView parents on this path (othertests.jl:2,18,47,69,110) */
typedef enum sb_dtype { SB_F32 = 0, SB_F64 = 1, SB_C32 = 2, SB_C64 = 3 } sb_dtype;

/* reduction operator `op`; the neutral elements are those of _init_reduction! (mapreduce.jl:182-187).
 * `&` and `|` (mapreduce.jl:186-187) are deliberately absent: they reduce Bool arrays, and Bool is not a device eltype
 * (sb_dtype); the glue keeps such calls on the reference's CPU method.  Predicate COUNTS (`count(x -> x < 0, A)`,
 * othertests.jl:116) are on the device path: SB_FN_LT yields 0/1 in the array's eltype and the count is an SB_OP_ADD. */
typedef enum sb_op { SB_OP_NONE = 0, SB_OP_ADD = 1, SB_OP_MUL = 2, SB_OP_MIN = 3, SB_OP_MAX = 4 } sb_op;

/* `initop` flavours seen on the path (linalg.jl:145-158, othertests.jl:76-102) */
typedef enum sb_initop {
    SB_INIT_NONE = 0,     /* initop === nothing : existing output contents participate             */
    SB_INIT_ZERO = 1,     /* zero / x->0                                                             */
    SB_INIT_IDENTITY = 2, /* identity (same result as NONE, kept for fidelity)                       */
    SB_INIT_SCALE = 3,    /* x -> beta*x                                                             */
    SB_INIT_CONST = 4,    /* x -> beta      (mapreduce(...; init=beta))                              */
    SB_INIT_CONJ = 5      /* conj                                                                    */
} sb_initop;

/* The element function `f` is a POSTFIX program mirroring the CaptureArgs tree that broadcast.jl:67-98
 * builds: arguments are consumed depth-first, left to right (capturestridedargs, broadcast.jl:41-46);
 * scalars / Refs are baked in as constants (make_capture, broadcast.jl:81-83). */
typedef enum sb_tok_kind {
    SB_TOK_ARG = 0,   /* push input number `a` (0-based: input a is operand a+1)                     */
    SB_TOK_CONST = 1, /* push the constant (re, im)                                                  */
    SB_TOK_CALL = 2   /* pop arity(fn) values, push fn(...)                                          */
} sb_tok_kind;

typedef enum sb_fn {
    /* unary */
    SB_FN_IDENTITY = 0, SB_FN_NEG = 1, SB_FN_CONJ = 2, SB_FN_ABS = 3, SB_FN_ABS2 = 4, SB_FN_REAL = 5,
    SB_FN_IMAG = 6, SB_FN_SQRT = 7, SB_FN_EXP = 8, SB_FN_LOG = 9, SB_FN_SIN = 10, SB_FN_COS = 11,
    SB_FN_TANH = 12, SB_FN_INV = 13,
    /* binary (n-ary `+`/`*` are emitted as left folds, as Julia evaluates them) */
    SB_FN_ADD = 32, SB_FN_SUB = 33, SB_FN_MUL = 34, SB_FN_DIV = 35, SB_FN_MAX = 36, SB_FN_MIN = 37,
    SB_FN_LT = 38 /* real(x) < real(y) ? 1 : 0  (predicate counts, othertests.jl:116) */
} sb_fn;

typedef struct sb_tok {
    int32_t kind; /* sb_tok_kind */
    int32_t a;    /* ARG: input index; CALL: sb_fn                                                  */
    double re;    /* CONST                                                                          */
    double im;
} sb_tok;

/* One `_mapreduce_fuse!` call. */
typedef struct sb_desc {
    int32_t ndim;                             /* 0..SB_MAX_DIMS                                     */
    int32_t nops;                             /* 1..SB_MAX_OPS ; operand 0 = output                 */
    int64_t dims[SB_MAX_DIMS];
    int64_t strides[SB_MAX_OPS][SB_MAX_DIMS]; /* elements; 0 and negative legal                     */
    void *base[SB_MAX_OPS];                   /* device (sb_mapreduce) or host (sb_mapreduce_host)  */
    int32_t dtype[SB_MAX_OPS];                /* sb_dtype                                           */
    int32_t conj[SB_MAX_OPS];
    int32_t ntok;
    sb_tok prog[SB_MAX_TOKENS];               /* f ; ntok == 0 means identity of input 0            */
    int32_t op;                               /* sb_op                                              */
    int32_t initop;                           /* sb_initop                                          */
    double init_re, init_im;                  /* beta for SCALE / CONST                             */
} sb_desc;

typedef struct sb_ctx sb_ctx;

/* ---- context ------------------------------------------------------------------------------------
 * GPU analog of the thread-count globals (Strided.jl:18-35): one ctx = one device + one stream.
 * `stream` is a cudaStream_t (NULL = the ctx's own non-blocking stream).  Fails with SB_E_NODEVICE
 * when there is no usable GPU -- there is no CPU execution path in this library. */
int sb_ctx_create(int device, void *stream, sb_ctx **out);
int sb_ctx_destroy(sb_ctx *ctx);
int sb_ctx_set_stream(sb_ctx *ctx, void *stream);
/* sync != 0 (default): sb_mapreduce returns after the result is complete, like the reference, which
 * joins its tasks before returning (mapreduce.jl:223).  sync == 0: stream-ordered (benchmarks). */
int sb_ctx_set_sync(sb_ctx *ctx, int sync);
/* The SB_* environment knobs are read when a ctx is created, not on the launch path.  A host program that changes them
 * afterwards (tests, tuning tools) calls this: re-reads them and drops the ctx's cached plans. */
int sb_ctx_reload_env(sb_ctx *ctx);
int sb_sync(sb_ctx *ctx);
const char *sb_last_error(sb_ctx *ctx); /* ctx may be NULL: last error of the calling thread        */
int sb_abi_version(void);
/* Waits for the background NVRTC compiles of this process (jit.cu).  Call it from the host's exit hook (Python `atexit`,
 * Julia `atexit`): a process must not run its exit handlers while a worker thread is still inside NVRTC.  Contexts stay
 * valid; idempotent; the library also registers it with atexit itself as a second line of defence. */
int sb_shutdown(void);

/* ---- device memory helpers for hosts without their own CUDA allocator (the Julia glue) ---------- */
int sb_malloc(sb_ctx *ctx, size_t bytes, void **out);
int sb_free(sb_ctx *ctx, void *ptr);
int sb_memcpy_h2d(sb_ctx *ctx, void *dst, const void *src, size_t bytes);
int sb_memcpy_d2h(sb_ctx *ctx, void *dst, const void *src, size_t bytes);

/* ---- the hot path -------------------------------------------------------------------------------
 * Replaces _mapreduce_block!/_mapreduce_threaded!/_mapreduce_kernel! (mapreduce.jl:142-425) for
 * device-resident operands.  Thread-safe per ctx. */
int sb_mapreduce(sb_ctx *ctx, const sb_desc *desc);

/* n calls as ONE batch (device pointers).  Small problems are bound by launch and DRAM latency (~4 us each on a B200,
 * whatever the kernel does); the batch runs its INDEPENDENT map calls -- output byte range disjoint from every other
 * call's operands -- concurrently on side streams between a fork and a join on the ctx's stream, everything else
 * (reductions, chains such as B = f(A); C = g(B)) in order after the join.  Independent calls that share ONE plan (same
 * dims, strides, eltypes and program; only the base pointers differ) are merged into a single grouped launch of up to 16
 * problems (`sb_stats.grouped_calls`; TMA ring kernel, or the LSU kernel for copies / scalings / two-input sums): a block
 * of equal-shape `permutedims!` then streams at HBM speed instead of paying one launch + one DRAM round trip per
 * statement.  The group has its own cached plan, sized for all its problems.  Results are the same as n sb_mapreduce calls
 * in order.  Capturable in a CUDA graph (parallel branches) once a batch has run outside the capture.  The reference
 * issues one `_mapreduce_fuse!` per statement (src/mapreduce.jl:98); this is the entry a glue uses for a block of
 * independent `@strided` statements. */
int sb_mapreduce_batch(sb_ctx *ctx, int n, const sb_desc *descs);

/* Same call for HOST-resident operands (plain Julia `Array` parents): stages every distinct parent
 * range to the device once (aliased views share one copy), runs sb_mapreduce, copies the output
 * range back.  This is the end-to-end entry `bench.py` times as "e2e".
 * Zero-copy mode: a SYNCHRONOUS call (sb_ctx_set_sync != 0) whose operands are all pinned / registered host memory
 * (cudaHostAlloc, cudaHostRegister) and whose inputs would cross the host link exactly once (a map without broadcast
 * re-reads; aliased permuted views of one parent are fetched once by the orbit kernel) is served by ONE kernel that
 * reads and writes host memory directly, so both directions of the link overlap.  SB_HOST_ZERO_COPY=0 disables it. */
int sb_mapreduce_host(sb_ctx *ctx, const sb_desc *desc);

/* ---- reductions across GPUs (one process per GPU) --------------------------------------------------
 * GPU analog of the per-task partial slots + serial fold of a threaded complete reduction
 * (`threadedout`, src/mapreduce.jl:153-170): every rank reduces ITS slab, the partials of all ranks are exchanged
 * through peer memory over NVLink by a kernel of this library (no NCCL call on the path) and folded in RANK ORDER on
 * every rank, so all ranks obtain bit-identical results:
 *
 *     out_r = op(initop(out_r), partial_0 (op) partial_1 (op) ... (op) partial_{world-1})
 *
 *   sb_peer_export  allocates this rank's exchange buffer and returns its 64-byte CUDA IPC handle;
 *   sb_peer_attach  takes the handles of ALL ranks (world * 64 bytes, rank order; exchanged by the host program,
 *                   e.g. torch.distributed.all_gather_object / MPI) and maps the peers' buffers;
 *   sb_mapreduce_allreduce  is COLLECTIVE: every attached rank must call it, in the same order, with descriptors whose
 *                   OUTPUT has the same number of elements (<= SB_PEER_MAX_OUT) and dtype.  `desc` describes the local
 *                   reduction exactly as for sb_mapreduce (device pointers; a zero-size local slab contributes the
 *                   neutral element).  world == 1 (or not attached): identical to sb_mapreduce.
 * The call number that tags the exchanged values lives in device memory and is advanced by the exchanging kernel, so a
 * collective call may be captured in a CUDA graph and replayed (every replay on every rank, in the same order).
 * sb_peer_export resets that counter together with the exchange buffer; sb_peer_attach does not touch either, so the
 * host program must put a barrier between the attach of all ranks and the first collective call.
 * A rank that waits longer than SB_PEER_TIMEOUT_MS (environment, default 30000) for a peer gives up WITHOUT trapping:
 * the call (sync mode) or the next sb_sync returns SB_E_CUDA, the result of that call is invalid, the context stays
 * usable. */
#define SB_PEER_MAX_OUT 1024
#define SB_PEER_MAX_WORLD 8
#define SB_IPC_HANDLE_BYTES 64
int sb_peer_export(sb_ctx *ctx, unsigned char handle_out[SB_IPC_HANDLE_BYTES]);
int sb_peer_attach(sb_ctx *ctx, int rank, int world, const unsigned char *handles /* world * 64 bytes */);
int sb_peer_detach(sb_ctx *ctx);
int sb_mapreduce_allreduce(sb_ctx *ctx, const sb_desc *desc);

/* ---- introspection (no GPU needed) --------------------------------------------------------------
 * CUDA graphs: a call may be captured (cudaStreamBeginCapture on the ctx's stream, sync == 0) once its plan exists, i.e.
 * after the same call has run once outside the capture; a first call inside a capture returns SB_E_UNSUPPORTED with a
 * message instead of invalidating the capture (plan tables and scratch are allocated on the first call).
 *
 * Writes a one-line JSON description of the plan the host planner picks for `desc` (kernel family,
 * canonical dims, tile extents, staged operands, grid) into buf.  ctx may be NULL. */
int sb_plan_describe(sb_ctx *ctx, const sb_desc *desc, char *buf, size_t buflen);

/* Counters: kernels launched / bytes staged since ctx creation (bench.py's gpu_launches, e2e bytes) */
typedef struct sb_stats {
    uint64_t launches;
    uint64_t h2d_bytes;
    uint64_t d2h_bytes;
    uint64_t plans_built;
    uint64_t plans_cached;
    uint64_t jit_launches; /* launches of NVRTC-specialised kernels (subset of `launches`) */
    uint64_t zero_copy_calls; /* sb_mapreduce_host calls served without staging (kernel reads/writes pinned host memory) */
    uint64_t batches;         /* sb_mapreduce_batch calls */
    uint64_t grouped_calls;   /* calls of a batch that shared ONE launch with other calls of the same plan */
} sb_stats;
int sb_get_stats(sb_ctx *ctx, sb_stats *out);
int sb_reset_stats(sb_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* STRIDED_B200_H */
