// planner.hpp -- host planner: sb_desc -> canonical problem -> tile plan (MapParams / ReduceParams).
//
// GPU counterpart of the reference planner `_mapreduce_fuse!` / `_mapreduce_order!` / `_computeblocks`
// (src/mapreduce.jl:98-139, 463-520).  Only two ideas carry over: fusing dims that are contiguous in every
// operand (:103-115) and ranking dims by stride order (:119-139).  The cache-line block model (:503-520) and
// the task bisection (:195-227) are CPU-specific and are replaced by shared-memory tiles and a CTA grid.
#pragma once
#include "common.hpp"
#include "../../include/strided_b200.h"
#include <string>
#include <vector>

namespace sb {

enum PlanKind : int { PLAN_NOOP = 0, PLAN_MAP = 1, PLAN_REDUCE = 2 };

struct DeviceInfo {
    int sm_count = 148;        // B200
    int ctas_per_sm = 4;       // resident CTAs of THREADS threads assumed for grid sizing
    bool host_link = false;    // operands are pinned HOST memory accessed by the kernel (zero-copy): fuse aliased views from 2 views on
    bool grouped = false;      // the plan serves a grouped launch (several problems of this shape in one grid): never "small"
};

struct KernelKey {
    int ct;      // compute dtype (DType)
    int recipe;  // Recipe
    int nin;     // template NIN
    int ept;     // template EPT
    int uniform; // template UNIFORM
};

struct Plan {
    int kind = PLAN_NOOP;
    KernelKey key{};
    MapParams map{};
    ReduceParams red{};
    int64_t grid = 0;
    int64_t smem_bytes = 0;
    int64_t scratch_bytes = 0;      // reduce partials
    int64_t finalize_threads = 0;   // > 0: launch reduce_finalize
    int64_t elements = 0;           // size of the canonical index space
    std::string family;             // "map_tile" | "reduce_tile" | "noop"
    int base_src[MAXO] = {0, 1, 2, 3, 4, 5, 6, 7}; // canonical operand k reads sb_desc::base[base_src[k]]
    // optional alias-aware tile order (host copy; uploaded by the ctx)
    std::vector<int32_t> tile_order; // launch position -> tile id (empty: natural order)
    // TMA-pipelined variant of the same map plan (chosen at bind time when the input bases are 16-byte aligned)
    bool tma_ok = false;
    TmaParams tma{};
    struct TmaGlobal { // what cuTensorMapEncodeTiled needs, minus the base address
        int rank = 0, elem_bytes = 0, swizzle = 0;
        uint64_t gdim[TMA_MAXRANK] = {0};
        uint64_t gstride_bytes[TMA_MAXRANK] = {0}; // entry i is the stride of dim i (entry 0 unused)
        uint32_t box[TMA_MAXRANK] = {0};
    } tma_global[TMA_MAXIN];
    int64_t tma_smem_bytes = 0;
    std::vector<TileDesc> tile_desc; // per-tile records in launch order (uploaded with the plan)
    int stream_recipe = 0;           // RC_INTERP, or a functor only the streamed reduction has (RC_S_ABS, RC_S_MUL2)
    std::vector<int64_t> lsu_desc;   // per-tile records of the LSU map kernel (MapParams::lsu_desc), launch order
    // alias-fused ("orbit") variant: all inputs are permuted views of one parent (see common.hpp).  Preferred over the
    // TMA ring when available; needs 16-byte aligned bases and an output that does not overlap the parent (bind time).
    bool orbit_ok = false;
    OrbitParams orbit{};
    TmaGlobal orbit_global[2];           // [0] parent (load), [1] output (store)
    std::vector<OrbitItem> orbit_items;  // work items in launch order (uploaded with the plan)
    int64_t orbit_smem_bytes = 0;
    int orbit_tile_b[MAXD] = {0};
    // streamed variant of a complete reduction over dense inputs (cp.async.bulk ring, one CTA per SM); chosen at bind time
    // when the input bases are 16-byte aligned
    bool stream_ok = false;
    StreamParams stream{};
    int64_t stream_grid = 0;
    int64_t stream_smem_bytes = 0;
    // the program needs more than the interpreter's 4 stack slots: only the NVRTC-specialised kernel can run it
    bool needs_jit = false;
    std::string note;
};

// Returns sb_status.  `err` receives a message on failure.
int build_plan(const sb_desc &d, const DeviceInfo &dev, Plan &plan, std::string &err);

// `_mapreducedim!` with a zero-size dim applies initop to a non-empty output (reference mapreduce.jl:88-91):
// rewrites D into the equivalent `map!(initop, out, out)`; false when nothing has to be done.
bool empty_initop_desc(const sb_desc &D, sb_desc &E);

// true when the output's byte range overlaps an input's (or the descriptor is malformed): the shifted-last-tile plans
// recompute a few elements and therefore need an output that no input aliases (decided per call, at bind time)
bool output_overlaps_inputs(const sb_desc &d);

// one-line JSON (sb_plan_describe)
std::string describe_plan(const Plan &plan);

// shared-memory bank model used by the padding search (exposed for tests)
int smem_wavefronts(const int32_t *elem_addr, int nlanes, int elem_bytes);

} // namespace sb
