/*
 * strided_ref.h -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the Strided.jl hot path.
 *
 * PARITY UNPINNED: Julia is not installed here or on the GPU box, the reference ships no golden
 * vectors (its tests are differential against Base `Array` semantics on Julia-RNG data, SURVEY.md
 * section 4), so this oracle cannot be checked against reference-produced output.  It is pinned only
 * by (i) the hand-derived planner known-answers of SURVEY.md section 8(a) and (ii) agreement with the
 * independent NumPy semantic oracle (oracle/semantic.py), which restates the Base semantics the
 * reference's own tests compare against.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use
 * anything under oracle/.  The product library never links or calls it.
 */
#ifndef STRIDED_REF_H
#define STRIDED_REF_H
#include "../include/strided_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define REF_MAX_LEAVES 256

typedef struct ref_plan_out {
    int32_t ndim;                       /* rank (unchanged by fusing: fused-away dims become 1)      */
    int64_t fused_dims[SB_MAX_DIMS];    /* after _mapreduce_fuse!           mapreduce.jl:98-117      */
    int64_t importance[SB_MAX_DIMS];    /* _mapreduce_order!                mapreduce.jl:126-132     */
    int32_t perm[SB_MAX_DIMS];          /* 1-based sortperm, as in the reference / SURVEY KATs       */
    int64_t dims[SB_MAX_DIMS];          /* ordered dims                                              */
    int64_t strides[SB_MAX_OPS][SB_MAX_DIMS]; /* ordered strides                                     */
    int64_t costs[SB_MAX_DIMS];         /* mapreduce.jl:137                                          */
    int64_t blocks[SB_MAX_DIMS];        /* _computeblocks                   mapreduce.jl:463-500     */
    int64_t region_bytes;               /* totalmemoryregion(blocks)        mapreduce.jl:503-520     */
    int32_t complete_reduction;         /* branch taken at mapreduce.jl:153                          */
    int32_t nleaves;                    /* leaves of the bisection at `nthreads`  :195-227           */
    int64_t leaf_dims[REF_MAX_LEAVES][SB_MAX_DIMS];
} ref_plan_out;

/* Plan only (no data touched): what the reference's planner/scheduler would decide. */
int ref_plan(const sb_desc *desc, int nthreads, ref_plan_out *out);

/* Execute on HOST pointers with `nthreads` tasks (1 = Strided.disable_threads()). */
int ref_mapreduce(const sb_desc *desc, int nthreads);

const char *ref_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
