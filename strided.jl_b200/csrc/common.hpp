// common.hpp -- PODs shared by the host planner and the sm_100a kernels.
//
// Design: every `_mapreduce_fuse!` call (reference src/mapreduce.jl:98) becomes a
// *tile plan*.  The index space is cut into power-of-two boxes ("tiles"); a CTA of THREADS threads owns a
// tile and each thread owns EPT elements `lin = t + j*THREADS` of it.  Because all tile extents and
// THREADS are powers of two, the coordinates of `lin` under ANY traversal order split into bit fields that
// come either from `t` or from `j`, so every address functional (global offset, shared-memory slot) is
//     F(t, j) = toff(t) + joff[j]
// with toff computed once per thread and joff a small table in kernel-parameter space.  Each operand is
// loaded in ITS OWN fastest-stride order (coalesced), staged through shared memory when that order differs
// from the output's, and consumed in the output's order.  This replaces the reference's cache-blocked loop
// nest `_mapreduce_kernel!` (src/mapreduce.jl:229-425) and its task bisection (:195-227).
#pragma once
#if defined(__CUDACC_RTC__) // NVRTC (csrc/jit.cu): no libc headers
#include <cuda/std/cstdint>
using cuda::std::int8_t;
using cuda::std::int16_t;
using cuda::std::int32_t;
using cuda::std::int64_t;
using cuda::std::uint8_t;
using cuda::std::uint16_t;
using cuda::std::uint32_t;
using cuda::std::uint64_t;
using cuda::std::uintptr_t;
#else
#include <stdint.h>
#endif

#if defined(__CUDACC__)
#define SB_HD __host__ __device__ __forceinline__
#define SB_D __device__ __forceinline__
#else
#define SB_HD inline
#define SB_D inline
#endif

namespace sb {

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------------
// Every kernel of this library is launched with programmaticStreamSerializationAllowed: in a sequence of calls (what a
// Julia program issuing one `@strided` statement after another produces) the NEXT kernel's launch latency and prologue
// (parameter loads, mbarrier init, per-thread offset functionals) overlap the tail of the previous one.  Correctness:
// every kernel executes griddepcontrol.wait -- which returns only when the preceding grid has completed and its
// memory is visible -- before its first access to global memory that a previous kernel may have written or read.
// (Both instructions are no-ops for a kernel that was not launched programmatically.)
#if defined(__CUDA_ARCH__)
SB_D void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
SB_D void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#else
SB_D void pdl_launch_dependents() {}
SB_D void pdl_wait() {}
#endif

constexpr int MAXD = 8;    // SB_MAX_DIMS
constexpr int MAXO = 8;    // SB_MAX_OPS (operand 0 = output)
constexpr int MAXIN = MAXO - 1;
constexpr int MAXTD = 6;   // tile dims with extent > 1
constexpr int MAXEPT = 16; // elements per thread per tile
constexpr int THREADS = 256;
constexpr int LOG_THREADS = 8;
constexpr int MAXTOK = 48;

enum DType : int { F32 = 0, F64 = 1, C32 = 2, C64 = 3 };
SB_HD int dtype_size(int dt) { return dt == F32 ? 4 : (dt == C64 ? 16 : 8); }

// ---- element program ------------------------------------------------------------------------------
enum TokKind : int { TOK_ARG = 0, TOK_CONST = 1, TOK_CALL = 2 };
enum Fn : int {
    FN_IDENTITY = 0, FN_NEG = 1, FN_CONJ = 2, FN_ABS = 3, FN_ABS2 = 4, FN_REAL = 5, FN_IMAG = 6, FN_SQRT = 7,
    FN_EXP = 8, FN_LOG = 9, FN_SIN = 10, FN_COS = 11, FN_TANH = 12, FN_INV = 13,
    FN_ADD = 32, FN_SUB = 33, FN_MUL = 34, FN_DIV = 35, FN_MAX = 36, FN_MIN = 37, FN_LT = 38
};
enum RedOp : int { OP_NONE = 0, OP_ADD = 1, OP_MUL = 2, OP_MIN = 3, OP_MAX = 4 };
enum InitOp : int { INIT_NONE = 0, INIT_ZERO = 1, INIT_IDENTITY = 2, INIT_SCALE = 3, INIT_CONST = 4, INIT_CONJ = 5 };

struct Tok {
    int32_t kind;
    int32_t a;
    double re, im;
};

// Pre-instantiated element functions ("recipes"): Julia specialises `_mapreduce_kernel!` on `f`
// (it is @generated on the callable's type); a precompiled engine cannot, so the planner pattern-matches
// the program against these and falls back to the in-kernel interpreter otherwise.
enum Recipe : int {
    RC_INTERP = 0,   // generic postfix interpreter
    RC_COPY,         // x0                           copy!/permutedims!/adjoint!       mapreduce.jl:2-14
    RC_SCALE,        // c0 * x0                      C1  `3 .* A'`, rmul!/lmul!        linalg.jl:2-3
    RC_ADD2,         // x0 + x1
    RC_ADD2_DIV,     // (x0 + x1) / c0               C2  `(A .+ A') ./ 2`
    RC_ADD2_MUL,     // (x0 + x1) * c0
    RC_SUM3,         // (x0 + x1) + x2
    RC_SUM4,         // ((x0 + x1) + x2) + x3        C4  4-way permutedims sum
    RC_AXPY,         // c0 * x0 + x1                 axpy!                             linalg.jl:23-31
    RC_AXPBY,        // c0 * x0 + c1 * x1            axpby!                            linalg.jl:32-42
    RC_ABS2,         // abs2(x0)                     C5  mapreduce(abs2, +, A; dims)
    RC_COUNT_,
    RC_JIT = RC_COUNT_, // element function compiled at run time by NVRTC from the postfix program (csrc/jit.cu)
    // functors of the STREAMED reduction only (plan.stream_recipe; the tile kernels run these programs through the
    // interpreter / NVRTC): `maximum(abs, A)` and dot-like `sum(x .* y)` ran at 0.14 / 0.28 of peak on the interpreter
    RC_S_ABS = 32,  // abs(x0)
    RC_S_MUL2 = 33  // x0 * x1
};

struct Program {
    int32_t recipe;
    int32_t ntok;
    double c0re, c0im, c1re, c1im; // recipe constants
    Tok tok[MAXTOK];
};

// ---- traversal order of a tile ---------------------------------------------------------------------
// slot i of an order covers bits [shift[i], shift[i]+bits[i]) of the in-tile linear index and is the
// coordinate along tile-dim `td[i]` (an index into MapPlan::tdim).
struct OrderTab {
    uint8_t n;
    uint8_t td[MAXTD];
    uint8_t shift[MAXTD];
    uint8_t bits[MAXTD];
};

// ---- division by a launch-time constant (tile decode) ---------------------------------------------------
// q = x / n for x < 2^31 with one mul.hi + shift (round-up magic number, Granlund & Montgomery).  Replaces
// the 64-bit div/mod chain that dominated the first profile (profiles/r01_v0_*: ~79 instr/element).
struct FastDiv {
    uint32_t mul, shr, n;
};
SB_HD uint32_t sb_umulhi(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}
SB_HD void fast_divmod(const FastDiv &f, uint32_t x, uint32_t &q, uint32_t &r)
{
    q = (f.n == 1u) ? x : (sb_umulhi(x, f.mul) >> f.shr);
    r = x - q * f.n;
}
#if !defined(__CUDACC_RTC__)
inline FastDiv make_fastdiv(uint32_t n)
{
    FastDiv f{0u, 0u, n};
    if (n <= 1u) return f;
    uint32_t lg = 0;
    while ((1ull << lg) < (uint64_t)n) ++lg;
    const uint32_t p = 31u + lg;
    f.mul = (uint32_t)(((1ull << p) + n - 1ull) / n);
    f.shr = p - 32u;
    return f;
}
#endif

// ---- per-tile descriptor table (TMA path) ----------------------------------------------------------------
// The no-load/no-store diagnostic (profiles/r01_v8_*) showed the TMA kernel bound by its own per-tile instruction
// overhead (~270 warp-instructions per tile, of which ~80 touch data).  Everything that depends only on the tile is
// therefore precomputed on the host, in LAUNCH order: one 32-byte record per tile, read with a single uniform load.
struct TileDesc {
    int32_t origin[5];  // element origin per canonical dim (TMA coordinates are a permutation of these)
    uint32_t id_full;   // bit 31: interior tile (no masking); bits 0..30: tile id (edge masks)
    int64_t out_off;    // byte offset of the tile origin in the output
};

// ---- map plan (kernel parameter block) -------------------------------------------------------------
struct MapParams {
    int32_t ndim;  // canonical dims (size-1 dropped, fused, sorted by |output stride|)
    int32_t nops;  // output + inputs
    int32_t ntd;   // tile dims (extent > 1)
    int32_t nstaged;
    int64_t dims[MAXD];
    int32_t tile_b[MAXD]; // tile extent per canonical dim (1 for grid-only dims)
    int32_t ntile[MAXD];  // ceil(dims/tile_b)
    int32_t nfull[MAXD];  // dims/tile_b: tile coordinate c is an interior (unmasked) tile iff c < nfull
    // Shifted last tile: along a dim whose extent is not a multiple of the (power-of-two) tile extent, the LAST tile starts
    // at dims - tile_b instead of (ntile-1)*tile_b, i.e. it overlaps its neighbour by `excess` = ntile*tile_b - dims
    // elements, which are simply computed twice (same inputs, same bits).  Every tile is then a full, unmasked tile: for
    // odd extents (41^4: 70 % of the tiles touch an edge) the masked slow path disappears.  Only for maps whose output
    // overlaps no input (decided at bind time; in-place updates keep the masked edge tiles).
    int32_t excess[MAXD]; // 0: dim is tiled exactly or is shorter than one tile
    int32_t shift_last;
    // Balanced tiles (umask = 1): the thread map stays a power-of-two box of 2^cbits per tile dim, but only the first
    // tile_b[d] <= 2^cbits coordinates of it are used -- tile_b[d] = ceil(dims / ceil(dims / 2^cbits)) -- so that
    // ceil(dims/2^cbits) tiles cover the dim with (almost) no overlap: 70 = 3 x 24 instead of 3 x 32 with 26 elements
    // computed twice.  Every tile then has the SAME valid region: its packed mask `urg` is a plan constant, no per-tile
    // decode.  Idle lanes cost issue slots, not memory requests (reversal permute of 70^4: L2<->SM traffic 1.88x -> 1.06x).
    int32_t umask;
    uint32_t urg;    // packed (tile_b - 1 | guard) of a balanced tile (see pack_rem)
    int32_t pad_um_;
    FastDiv tdiv[MAXD];   // division by ntile[d]
    int64_t tstep[MAXO][MAXD]; // BYTES per tile step along d: tile_b[d] * strides[k][d] * sizeof(elem k)
    uint8_t tdim[MAXTD];  // tile-dim slot -> canonical dim
    int64_t ntiles;
    const int32_t *tile_order; // optional device table: launch position -> tile id (alias-aware order)
    const TileDesc *tile_desc; // optional device table (TMA path): launch position -> precomputed tile record
    // optional device table (LSU kernel): launch position -> {tile id | full << 31 | packed edge mask << 32, byte offset of the tile origin in operand
    // 0, 1, ..., nops-1} as nops + 1 int64 words.  ncu on the 91^4 reversal showed ~400 instructions per thread and tile,
    // most of them the per-tile decode (4 magic divisions, 64-bit multiply-adds per operand and dim, shifted-tile
    // corrections) that every thread repeats: with the record it is one uniform load and one add per operand.
    const int64_t *lsu_desc;
    int32_t lsu_prefetch; // 1: the kernel fetches the NEXT tile's record while it works on the current one (and the first one before griddepcontrol.wait)
    int32_t pad_lsu_;
    unsigned char *base[MAXO];
    int64_t strides[MAXO][MAXD]; // elements
    uint8_t dtype[MAXO];
    uint8_t conj[MAXO];
    uint8_t staged[MAXO];  // 1: loaded in own order, transposed through shared memory
    OrderTab order[MAXO];  // order[0] = output order; order[k] = load order of operand k
    // all address functionals are in BYTES
    int64_t g_tstr[MAXO][MAXTD]; // global byte stride per order slot (for toff)
    int64_t g_joff[MAXO][MAXEPT];
    int32_t w_tstr[MAXO][MAXTD]; // shared-memory byte stride per OWN-order slot (staged operands)
    int32_t w_joff[MAXO][MAXEPT];
    int32_t r_tstr[MAXO][MAXTD]; // shared-memory byte stride per OUTPUT-order slot
    int32_t r_joff[MAXO][MAXEPT];
    int32_t smem_off[MAXO];      // byte offset of the operand's staging buffer (elements stored as CT)
    // edge masks: tile coordinates packed into one word, field of tile-dim td at bit cpos[td] with one guard bit on
    // top; element (t, j) of operand k has packed coordinates c_toff(t) + c_joff[k][j]; it is inside the array iff
    // ((R | guard) - C) & guard == guard with R = packed (remaining extent - 1).  One subtract per element.
    uint32_t c_tstr[MAXO][MAXTD]; // 1 << cpos[td] per order slot
    uint32_t c_joff[MAXO][MAXEPT];
    uint32_t guard;
    uint8_t cpos[MAXTD];
    uint8_t cbits[MAXTD];
    int32_t ept;
    int32_t uniform; // all dtypes == compute type and no conj flags
    int32_t vbits;   // log2 of the per-thread vector length V (elements): 16 bytes / sizeof(compute type), or 0
    uint8_t gvec[MAXO]; // operand k: V consecutive elements of its load traversal are contiguous + 16-B aligned in HBM
    uint8_t svec[MAXO]; // staged operand k: ... and in its staging buffer (128-bit shared-memory stores)
    Program prog;
};

// Several problems of ONE map plan in one launch of the LSU kernel (sb_mapreduce_batch): problem p's operand k lives
// `delta[p][k]` bytes from problem 0's (whose bases are MapParams::base); launch position g = p * ntiles + tile.
constexpr int MAP_GROUP_MAX = 16;
struct MapGroup {
    int32_t nprob;
    int32_t pad_;
    int64_t delta[MAP_GROUP_MAX][MAXO];
};

// ---- TMA-staged map plan ------------------------------------------------------------------------------------
// Same tile decomposition as MapParams, but every INPUT tile is fetched by the Tensor Memory Accelerator
// (cp.async.bulk.tensor) into a multi-stage shared-memory ring, several tiles ahead of the consumers.
// Operands whose fastest dim is not the output's ("staged": the permutedims / transpose case) are written with
// the 128-byte swizzle so that the transposed read is (almost) bank-conflict free; the consumer address of
// element (t, j) is  S_t XOR s_joff[j]  because the dense box layout is a power-of-two bit-field layout and
// the swizzle  D ^ (((D >> 7) & 7) << 4)  is linear over GF(2).
constexpr int TMA_MAXIN = 4;
constexpr int TMA_MAXRANK = 5;
struct TmaOperand {
    int32_t rank;                // tensor-map rank = canonical ndim, dims in the operand's own stride order
    int32_t nbox;                // boxes per tile (inner dim split into 128-byte rows when swizzled)
    int32_t box_bytes;
    int32_t smem_off;            // byte offset inside a stage (multiple of 1024)
    int32_t inner_step;          // inner-dim elements per box
    int32_t swizzle;             // 0: none, 1: 128-byte
    uint8_t cdim[TMA_MAXRANK];   // tensor-map dim i -> canonical dim
    uint8_t pad_[3];
    int32_t box[TMA_MAXRANK];    // box extent per tensor-map dim
    // consumer functional: dense byte offset per OUTPUT-order slot
    int32_t d_lo[MAXTD];
    int32_t inner_slot;          // output-order slot of this operand's inner dim (-1: not a tile dim)
    int32_t split_bits;          // inner coordinate bits that stay inside one box row
    int32_t d_hi;                // byte stride of the box index
    int32_t s_joff[MAXEPT];      // (swizzled) j part
};
struct TmaParams {
    int32_t nin;
    int32_t nstage;
    int32_t stage_bytes;
    int32_t pad_;
    TmaOperand op[TMA_MAXIN];
};
SB_HD uint32_t swizzle128(uint32_t d) { return d ^ (((d >> 7) & 7u) << 4); }

// ---- alias-fused ("orbit") map plan --------------------------------------------------------------------------
// When every input is a dim-permuted view of ONE parent (A and A' in `(A .+ A') ./ 2`; the four rotations of the
// 4-way permutedims sum, README.md:91-104), the output tiles fall into orbits under the group generated by the
// permutations, and the tiles of one orbit need exactly the parent blocks of that orbit.  A work item is one
// orbit: its <= ORB_MAXG parent blocks are fetched ONCE by TMA (cp.async.bulk.tensor, dense box), every output
// tile of the orbit is computed from shared memory (each view reads "its" block through a permuted address
// functional), staged in output layout and written back with a TMA store.  SM<->L2 traffic drops from
// (nin + 1) x to 2 x the array; the reference gets the same effect from cache blocking inside one task
// (src/mapreduce.jl:463-500).
//
// Thread t of T = 2^log_threads consumers handles elements u = t + T*j of a tile THROUGH A GF(2)-LINEAR MAP x = M u onto the tile's bit-field
// coordinates, chosen by the planner so that all views' shared-memory reads and the staging write of a warp are
// bank-conflict free.  Every address functional is again  T(t) XOR J[j].
constexpr int ORB_MAXG = 4;      // tiles (= parent blocks) per work item
constexpr int ORB_MAXIN = 4;
constexpr int ORB_MAXLOGT = 9;    // consumer threads: 2^8 (two CTAs per SM possible) or 2^9 (one big CTA per SM), + 1 producer warp
struct OrbitItem {
    int32_t ntile;  // output tiles computed by this work item
    int32_t nblock; // parent blocks loaded for it (>= ntile: small problems split an orbit's tiles over several items,
                    // each of which loads all the orbit's blocks, so that every SM gets work)
    int32_t pcrd[ORB_MAXG][TMA_MAXRANK]; // TMA coordinates (parent dim order) of parent block s
    int32_t ocrd[ORB_MAXG][TMA_MAXRANK]; // TMA coordinates (output dim order) of output tile m
    uint8_t slot[ORB_MAXG][ORB_MAXIN];   // input k of output tile m reads parent block slot[m][k]
    int64_t ooff[ORB_MAXG];              // byte offset of output tile m's origin (direct-store mode)
};
struct OrbitParams {
    int32_t nin, rank;
    int32_t log_threads; // log2 of the consumer thread count (8 or 9)
    int32_t pad0_;
    int32_t nstage;      // depth of the input ring (stages of gmax blocks)
    int32_t tile_bytes;  // one parent block = one output tile
    int32_t stage_bytes; // gmax * tile_bytes
    int32_t gmax;
    int32_t ept;
    int32_t nitems;
    int32_t debug;       // diagnostics only (SB_DEBUG, tools/): 1 no TMA loads, 2 no TMA stores, 4 no compute, 8 no proxy fence, 16 no tile barrier -- results are WRONG
    int32_t nstaging;    // output staging buffers (TMA stores in flight + 1), 2..4
    const OrbitItem *items;
    // direct-store mode (no edge tiles): the staged tile is written by all consumer threads with 128-bit st.global,
    // thread t owning the 16-byte groups g = t + 256 r of the staging buffer; the TMA unit then only serves the loads
    // (it processes box rows at ~2 cycles each, which bounds 32-byte-row tiles when it has to do both directions)
    int32_t direct_store;
    int32_t st_groups;            // groups per thread per tile = tile_bytes / (16 * consumer threads)
    unsigned char *out_base;
    int64_t st_tcol[ORB_MAXLOGT]; // global byte offset contributed by bit i of t
    int64_t st_roff[8];           // ... by r
    uint32_t tcol[ORB_MAXIN + 1][ORB_MAXLOGT]; // byte-address image of thread bit i; view 0 = output staging, k = input k
    uint32_t jtab[ORB_MAXIN + 1][MAXEPT];      // byte-address image of j
    Program prog;
};

// ---- peer group (reductions across GPUs, include/strided_b200.h sb_peer_*) ------------------------------------
// Exchange buffer of a rank: 8 flag words 128 bytes apart (wide elements only), then data[parity][rank][PEER_MAX_OUT]
// in 16-byte slots.  Elements of <= 8 bytes travel as two 8-byte stores {32-bit half, epoch} (see abi.cu).
constexpr int PEER_MAX_WORLD = 8;
constexpr int PEER_MAX_OUT = 1024;
constexpr uint32_t PEER_FLAG_STRIDE = 128;
constexpr uint32_t PEER_DATA_OFF = PEER_MAX_WORLD * PEER_FLAG_STRIDE;
constexpr uint32_t PEER_SLOT = 16;
struct PeerLink {
    int32_t world, rank; // world <= 1: no exchange
    // The call number ("epoch") lives in DEVICE memory and is advanced by the exchanging kernel itself, so that a CUDA
    // graph that captured a collective call uses a fresh epoch on every replay (a host-side counter baked into the
    // kernel parameters would make every replay accept the previous replay's slots).
    uint32_t *epoch_ptr;
    uint32_t *err_flag;     // set to 1 (mapped host memory) when a peer did not show up within timeout_cycles
    int64_t timeout_cycles; // clock64() ticks; the kernel then gives up WITHOUT trapping (the context stays usable)
    unsigned char *buf[PEER_MAX_WORLD]; // rank g's exchange buffer as mapped into this process
};

// ---- reduce plan ------------------------------------------------------------------------------------
struct ReduceParams {
    int32_t ndim, nops, ntd;
    int32_t nkept;               // canonical dims [0, nkept) are kept, [nkept, ndim) reduced
    int64_t dims[MAXD];
    int32_t tile_b[MAXD];
    int32_t ntile[MAXD];
    int32_t nfull[MAXD];
    FastDiv tdiv[MAXD];
    int64_t tstep[MAXO][MAXD];   // BYTES per tile step along d
    uint8_t tdim[MAXTD];
    FastDiv outdiv;              // division by nouttiles (block id -> split, out tile)
    int64_t nouttiles;           // product of ntile over kept dims
    int64_t nrsteps;             // product of ntile over reduced dims
    int64_t steps_per_split;
    int32_t nsplit;
    unsigned char *base[MAXO];
    int64_t strides[MAXO][MAXD];
    uint8_t dtype[MAXO];
    uint8_t conj[MAXO];
    OrderTab order;              // load order (input 1's fastest-stride order) over ALL tile dims
    int64_t g_tstr[MAXO][MAXTD]; // BYTES
    int64_t g_joff[MAXO][MAXEPT];
    uint32_t c_tstr[MAXTD]; // packed edge-mask coordinates (see MapParams)
    uint32_t c_joff[MAXEPT];
    uint32_t guard;
    uint8_t cpos[MAXTD];
    uint8_t cbits[MAXTD];
    int32_t s_tstr[MAXTD];       // shared-memory slot (elements) of (t, j) for the final in-CTA combine
    int32_t s_joff[MAXEPT];
    int32_t nout_tile;           // outputs per tile  = prod of kept tile extents
    int32_t nred_tile;           // partials per output = E / nout_tile
    int32_t warp_per_output;     // 1: layout [o][r], warp folds one output; 0: layout [r][o], thread per output
    OrderTab kept_order;         // decode of output number o -> kept tile coords (output order)
    unsigned char *scratch;      // partials [nsplit][nouttiles][nout_tile] of the ACCUMULATOR type
    uint32_t *counters;          // one arrival counter per output tile: the LAST split CTA to arrive folds the partials
    int32_t ept;
    int32_t uniform;
    int32_t op;
    int32_t initop;
    double init_re, init_im;
    PeerLink peer;               // fused exchange of the final values across GPUs (single output tile plans)
    Program prog;
};

// ---- streamed complete reduction (reduce_stream_kernel) -----------------------------------------------------------
// Complete reductions of dense operands (`sum(abs2, A)`, `mapreduce(f, op, A, B)` without `dims`; the per-GPU share of
// BASELINE config 5) are one long contiguous stream per input: no tile decode is needed at all.  One persistent CTA per
// SM; a producer thread keeps `nstage` chunks of `chunk_bytes` per input in flight with cp.async.bulk (1-D bulk copy,
// mbarrier complete_tx) from the first cycle on, eight consumer warps fold the landed chunks out of shared memory with
// 128-bit loads; CTA partials meet in a fixed order in the last-arriving CTA (one acq_rel atomic per CTA, no fences).
// GPU analog of the per-task partial slots + serial fold of the reference (src/mapreduce.jl:153-170).
// The same kernel serves a FEW outputs whose reduction ranges are each one dense run (`mapreduce(f, op, A; dims=(1,2))`
// of a 3-D array; config 5 with several dense slices per GPU): the outputs are streamed one after the other through the
// same ring, every CTA leaves one partial per output.
constexpr int STREAM_MAXKD = 4;   // kept dims (after canonical fusing)
constexpr int STREAM_MAXOUT = 64; // outputs
struct StreamParams {
    int64_t nelem;       // elements per input PER OUTPUT (every run is dense, stride 1, the accumulator's type)
    int64_t vec_bytes;   // bytes of a run covered by whole 16-byte vectors (the <16-byte rest is folded by one thread)
    int64_t nchunks;     // chunks per output: ceil(vec_bytes / chunk_bytes)
    int32_t nin;
    int32_t chunk_bytes; // per input per stage (multiple of 16)
    int32_t nstage;
    int32_t stage_bytes; // nin * chunk_bytes
    int32_t nout;        // outputs = product of kdims (1: complete reduction)
    int32_t nkd;         // kept dims
    // Interleaved mode (inter_g > 0): the kept dim is the INNERMOST, contiguous dim of the inputs (column-major
    // `mapreduce(f, op, A; dims=(2,3))`, BASELINE config 5 on one GPU): the whole input is ONE dense run in which element e
    // belongs to output e mod K.  With K * sizeof(T) = inter_g * 16 bytes (a power of two <= 512) every 16-byte vector
    // lane of thread t always meets the same output ((t mod inter_g) * V + lane), so the same ring serves it with one
    // accumulator per lane and an epilogue that folds lanes of equal class.
    int32_t inter_g;
    int32_t pad_inter_;
    int64_t kdims[STREAM_MAXKD];
    int64_t kin_bytes[MAXIN][STREAM_MAXKD]; // byte stride of input k along kept dim d (multiple of 16)
    int64_t kout_bytes[STREAM_MAXKD];       // byte stride of the output along kept dim d
};

// Kernel parameter block of reduce_stream_kernel: only what the streamed reduction reads (~1.9 KB; the tile plan's
// ReduceParams is 4.6 KB, and the launch cost of a graph node grows with the size of its parameter block).  Member
// names follow ReduceParams so that the shared helpers (peer exchange, element functions) take either.
struct StreamArgs {
    unsigned char *base[4]; // 0: output, 1..3: inputs
    uint8_t dtype[4];
    uint8_t conj[4];
    int32_t op, initop;
    double init_re, init_im;
    unsigned char *scratch; // CTA partials [nout][grid] of the accumulator type
    uint32_t *counters;     // arrival counter
    PeerLink peer;
    StreamParams S;
    Program prog;
};

// element origin of tile coordinate c along canonical dim d (see MapParams::excess)
SB_HD int64_t map_tile_origin(const MapParams &P, int d, uint32_t c)
{
    const int64_t o = (int64_t)c * P.tile_b[d];
    return (P.shift_last && P.excess[d] != 0 && (int32_t)c == P.ntile[d] - 1) ? o - P.excess[d] : o;
}

// ---- in-tile linear index of element (t, j) -----------------------------------------------------------------
// V = 2^vbits consecutive elements of the traversal (16 bytes) belong to the same thread, so that global loads/
// stores and staging-buffer writes can be 128-bit wherever the operand is contiguous along the traversal's fastest
// dim:  lin = (j mod V) + V * (t + THREADS * (j div V)).  t-bits and j-bits stay disjoint bit fields of lin.
SB_HD int lin_t(int t, int vbits) { return t << vbits; }
SB_HD int lin_j(int j, int vbits) { return (j & ((1 << vbits) - 1)) | ((j >> vbits) << (vbits + LOG_THREADS)); }

// packed "remaining extent - 1" of a tile (R | guard), from the per-tile-dim remaining extents
SB_HD uint32_t pack_rem(const int32_t *rem, int ntd, const uint8_t *cpos, const uint8_t *cbits, uint32_t guard)
{
    uint32_t r = guard;
    for (int i = 0; i < ntd; ++i) {
        const int32_t mx = (1 << cbits[i]) - 1;
        int32_t v = rem[i] - 1;
        if (v > mx) v = mx;
        if (v < 0) return 0; // empty: nothing is valid (guard bits cleared)
        r |= (uint32_t)v << cpos[i];
    }
    return r;
}
SB_HD bool packed_valid(uint32_t rg, uint32_t c, uint32_t guard) { return ((rg - c) & guard) == guard; }

// ---- small HD helpers ---------------------------------------------------------------------------------
SB_HD int field_of(const OrderTab &o, int slot, int lin)
{
    return (lin >> o.shift[slot]) & ((1 << o.bits[slot]) - 1);
}

} // namespace sb
