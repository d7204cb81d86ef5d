/*
 * strided_ref.c -- TEST INFRASTRUCTURE ONLY (see strided_ref.h: PARITY UNPINNED).
 *
 * CPU restatement, in C11 + pthreads, of the Strided.jl v2.3.2 hot path.  Each function names the
 * reference lines it follows (all in /root/reference/src/mapreduce.jl unless noted).  It is the parity
 * checker for the CUDA engine and the timed "reference CPU path" of bench.py; it is never part of
 * the product.  Nothing here is copied: Julia tuples/@generated code are re-expressed as plain loops.
 */
#include "strided_ref.h"

#include <complex.h>
#undef I /* we use I for index pointers, as the reference does */
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MINTHREADLENGTH ((int64_t)1 << 15) /* :141 */
#define BLOCKMEMORYSIZE ((int64_t)1 << 15) /* :462 */
#define CACHELINE 64                        /* :502 */

static _Thread_local char g_err[256];
const char *ref_last_error(void) { return g_err; }
static int fail(int code, const char *msg)
{
    snprintf(g_err, sizeof g_err, "%s", msg);
    return code;
}

enum { CT_F32 = 0, CT_F64 = 1, CT_C32 = 2, CT_C64 = 3 };

typedef struct ref_job ref_job;
typedef void (*inner_fn)(const ref_job *, char **, int64_t);
typedef void (*initel_fn)(const ref_job *, char *);

struct ref_job {
    int N, M;
    int64_t dims[SB_MAX_DIMS];
    int64_t strides[SB_MAX_OPS][SB_MAX_DIMS]; /* ordered, elements */
    int64_t bstr[SB_MAX_OPS][SB_MAX_DIMS];    /* ordered, bytes    */
    int64_t costs[SB_MAX_DIMS];
    int64_t blocks[SB_MAX_DIMS];
    char *base[SB_MAX_OPS];
    int dtype[SB_MAX_OPS], conj[SB_MAX_OPS], esize[SB_MAX_OPS];
    int ntok;
    sb_tok prog[SB_MAX_TOKENS];
    int op, initop;
    double init_re, init_im;
    int ct;
    double c0, c1; /* constants of the specialised recipes */
    inner_fn inner;
    initel_fn initel;
    /* plan recording */
    ref_plan_out *rec;
};

static const int ESIZE[4] = {4, 8, 8, 16};

/* ------------------------------------------------------------------------------------------------
 * typed element code: four instantiations of ref_eval.inc
 * ---------------------------------------------------------------------------------------------- */
#define CT float
#define RT float
#define SUF f32
#define CPLX 0
#define MK(re, im) ((float)(re))
#define RE(x) (x)
#define IM(x) (0.0f)
#define R_SQRT sqrtf
#define R_EXP expf
#define R_LOG logf
#define R_SIN sinf
#define R_COS cosf
#define R_TANH tanhf
#define R_FABS fabsf
#include "ref_eval.inc"
#undef CT
#undef RT
#undef SUF
#undef CPLX
#undef MK
#undef RE
#undef IM
#undef R_SQRT
#undef R_EXP
#undef R_LOG
#undef R_SIN
#undef R_COS
#undef R_TANH
#undef R_FABS

#define CT double
#define RT double
#define SUF f64
#define CPLX 0
#define MK(re, im) ((double)(re))
#define RE(x) (x)
#define IM(x) (0.0)
#define R_SQRT sqrt
#define R_EXP exp
#define R_LOG log
#define R_SIN sin
#define R_COS cos
#define R_TANH tanh
#define R_FABS fabs
#include "ref_eval.inc"
#undef CT
#undef RT
#undef SUF
#undef CPLX
#undef MK
#undef RE
#undef IM
#undef R_SQRT
#undef R_EXP
#undef R_LOG
#undef R_SIN
#undef R_COS
#undef R_TANH
#undef R_FABS

#define CT float complex
#define RT float
#define SUF c32
#define CPLX 1
#define MK(re, im) CMPLXF((re), (im))
#define RE(x) crealf(x)
#define IM(x) cimagf(x)
#define R_HYPOT hypotf
#define C_SQRT csqrtf
#define C_EXP cexpf
#define C_LOG clogf
#define C_SIN csinf
#define C_COS ccosf
#define C_TANH ctanhf
#include "ref_eval.inc"
#undef CT
#undef RT
#undef SUF
#undef CPLX
#undef MK
#undef RE
#undef IM
#undef R_HYPOT
#undef C_SQRT
#undef C_EXP
#undef C_LOG
#undef C_SIN
#undef C_COS
#undef C_TANH

#define CT double complex
#define RT double
#define SUF c64
#define CPLX 1
#define MK(re, im) CMPLX((re), (im))
#define RE(x) creal(x)
#define IM(x) cimag(x)
#define R_HYPOT hypot
#define C_SQRT csqrt
#define C_EXP cexp
#define C_LOG clog
#define C_SIN csin
#define C_COS ccos
#define C_TANH ctanh
#include "ref_eval.inc"
#undef CT
#undef RT
#undef SUF
#undef CPLX
#undef MK
#undef RE
#undef IM

/* ------------------------------------------------------------------------------------------------
 * specialised innermost loops: Julia compiles one kernel per (f, op, eltypes); these are the same
 * @simd loops (mapreduce.jl:320-336) for the recipes of the BASELINE configs, so the CPU baseline
 * is not handicapped by the generic postfix interpreter.  Same arithmetic, same rounding.
 * ---------------------------------------------------------------------------------------------- */
#define DEF_COPY(NAME_, TY)                                                                          \
    static void NAME_(const ref_job *jb, char **I, int64_t d1)                                       \
    {                                                                                                \
        const int64_t so = jb->strides[0][0], si = jb->strides[1][0];                                \
        TY *restrict o = (TY *)I[0];                                                                 \
        const TY *restrict a = (const TY *)I[1];                                                     \
        if (so == 1 && si == 1) for (int64_t j = 0; j < d1; ++j) o[j] = a[j];                        \
        else if (so == 1) for (int64_t j = 0; j < d1; ++j) o[j] = a[j * si];                         \
        else for (int64_t j = 0; j < d1; ++j) o[j * so] = a[j * si];                                 \
    }
DEF_COPY(inner_copy4, uint32_t)
DEF_COPY(inner_copy8, uint64_t)
typedef struct { uint64_t a, b; } u128_t;
DEF_COPY(inner_copy16, u128_t)

#define DEF_SCALE(NAME_, TY)                                                                         \
    static void NAME_(const ref_job *jb, char **I, int64_t d1) /* c0 * x  (C1: 3 .* A') */           \
    {                                                                                                \
        const int64_t so = jb->strides[0][0], si = jb->strides[1][0];                                \
        const TY c = (TY)jb->c0;                                                                     \
        TY *restrict o = (TY *)I[0];                                                                 \
        const TY *restrict a = (const TY *)I[1];                                                     \
        if (so == 1 && si == 1) for (int64_t j = 0; j < d1; ++j) o[j] = c * a[j];                    \
        else if (so == 1) for (int64_t j = 0; j < d1; ++j) o[j] = c * a[j * si];                     \
        else for (int64_t j = 0; j < d1; ++j) o[j * so] = c * a[j * si];                             \
    }
DEF_SCALE(inner_scale_f32, float)
DEF_SCALE(inner_scale_f64, double)

#define DEF_ADD2DIV(NAME_, TY)                                                                       \
    static void NAME_(const ref_job *jb, char **I, int64_t d1) /* (x + y) / c0  (C2) */              \
    {                                                                                                \
        const int64_t so = jb->strides[0][0], s1 = jb->strides[1][0], s2 = jb->strides[2][0];        \
        const TY c = (TY)jb->c0;                                                                     \
        TY *restrict o = (TY *)I[0];                                                                 \
        const TY *restrict a = (const TY *)I[1];                                                     \
        const TY *restrict b = (const TY *)I[2];                                                     \
        if (so == 1 && s1 == 1) for (int64_t j = 0; j < d1; ++j) o[j] = (a[j] + b[j * s2]) / c;      \
        else for (int64_t j = 0; j < d1; ++j) o[j * so] = (a[j * s1] + b[j * s2]) / c;               \
    }
DEF_ADD2DIV(inner_add2div_f32, float)
DEF_ADD2DIV(inner_add2div_f64, double)

#define DEF_SUM4(NAME_, TY)                                                                          \
    static void NAME_(const ref_job *jb, char **I, int64_t d1) /* ((x+y)+z)+w  (C4) */               \
    {                                                                                                \
        const int64_t so = jb->strides[0][0], s1 = jb->strides[1][0], s2 = jb->strides[2][0],        \
                      s3 = jb->strides[3][0], s4 = jb->strides[4][0];                                \
        TY *restrict o = (TY *)I[0];                                                                 \
        const TY *restrict a = (const TY *)I[1];                                                     \
        const TY *restrict b = (const TY *)I[2];                                                     \
        const TY *restrict c = (const TY *)I[3];                                                     \
        const TY *restrict d = (const TY *)I[4];                                                     \
        for (int64_t j = 0; j < d1; ++j)                                                             \
            o[j * so] = ((a[j * s1] + b[j * s2]) + c[j * s3]) + d[j * s4];                           \
    }
DEF_SUM4(inner_sum4_f32, float)
DEF_SUM4(inner_sum4_f64, double)

#define DEF_ABS2ADD(NAME_, TY)                                                                       \
    static void NAME_(const ref_job *jb, char **I, int64_t d1) /* op=+, f=abs2  (C5) */              \
    {                                                                                                \
        const int64_t so = jb->strides[0][0], si = jb->strides[1][0];                                \
        TY *restrict o = (TY *)I[0];                                                                 \
        const TY *restrict a = (const TY *)I[1];                                                     \
        if (so == 0) { /* register accumulator; @simd licenses the reassociation (:320-327) */       \
            TY acc = o[0];                                                                           \
            if (si == 1) {                                                                           \
                _Pragma("omp simd reduction(+ : acc)") for (int64_t j = 0; j < d1; ++j) acc += a[j] * a[j]; \
            } else {                                                                                 \
                _Pragma("omp simd reduction(+ : acc)") for (int64_t j = 0; j < d1; ++j) acc += a[j * si] * a[j * si]; \
            }                                                                                        \
            o[0] = acc;                                                                              \
        } else if (so == 1 && si == 1) {                                                             \
            for (int64_t j = 0; j < d1; ++j) o[j] += a[j] * a[j];                                    \
        } else {                                                                                     \
            for (int64_t j = 0; j < d1; ++j) o[j * so] += a[j * si] * a[j * si];                     \
        }                                                                                            \
    }
DEF_ABS2ADD(inner_abs2add_f32, float)
DEF_ABS2ADD(inner_abs2add_f64, double)

#define DEF_IDADD(NAME_, TY)                                                                         \
    static void NAME_(const ref_job *jb, char **I, int64_t d1) /* op=+, f=identity (sum) */          \
    {                                                                                                \
        const int64_t so = jb->strides[0][0], si = jb->strides[1][0];                                \
        TY *restrict o = (TY *)I[0];                                                                 \
        const TY *restrict a = (const TY *)I[1];                                                     \
        if (so == 0) {                                                                               \
            TY acc = o[0];                                                                           \
            _Pragma("omp simd reduction(+ : acc)") for (int64_t j = 0; j < d1; ++j) acc += a[j * si]; \
            o[0] = acc;                                                                              \
        } else {                                                                                     \
            for (int64_t j = 0; j < d1; ++j) o[j * so] += a[j * si];                                 \
        }                                                                                            \
    }
DEF_IDADD(inner_idadd_f32, float)
DEF_IDADD(inner_idadd_f64, double)

static int is_arg(const sb_tok *t, int a) { return t->kind == SB_TOK_ARG && t->a == a; }
static int is_call(const sb_tok *t, int fn) { return t->kind == SB_TOK_CALL && t->a == fn; }
static int is_const_real(const sb_tok *t) { return t->kind == SB_TOK_CONST && t->im == 0.0; }

/* choose the innermost loop; falls back to the generic interpreter loop of ref_eval.inc */
static void pick_inner(ref_job *jb)
{
    static const inner_fn generic[4] = {inner_f32, inner_f64, inner_c32, inner_c64};
    static const initel_fn initel[4] = {initel_f32, initel_f64, initel_c32, initel_c64};
    jb->inner = generic[jb->ct];
    jb->initel = initel[jb->ct];
    int uniform = 1;
    for (int k = 0; k < jb->M; ++k)
        if (jb->dtype[k] != jb->dtype[0] || jb->conj[k]) uniform = 0;
    if (!uniform || jb->ct != jb->dtype[0]) return;
    const sb_tok *p = jb->prog;
    const int n = jb->ntok, real = (jb->ct == CT_F32 || jb->ct == CT_F64), f64 = (jb->ct == CT_F64);
    if (jb->op == SB_OP_NONE && jb->strides[0][0] != 0) {
        if (jb->M == 2 && (n == 0 || (n == 1 && is_arg(p, 0)) || (n == 2 && is_arg(p, 0) && is_call(p + 1, SB_FN_IDENTITY)))) {
            jb->inner = jb->esize[0] == 4 ? inner_copy4 : jb->esize[0] == 8 ? inner_copy8 : inner_copy16;
        } else if (real && jb->M == 2 && n == 3 && is_const_real(p) && is_arg(p + 1, 0) && is_call(p + 2, SB_FN_MUL)) {
            jb->c0 = p[0].re;
            jb->inner = f64 ? inner_scale_f64 : inner_scale_f32;
        } else if (real && jb->M == 3 && n == 5 && is_arg(p, 0) && is_arg(p + 1, 1) && is_call(p + 2, SB_FN_ADD) &&
                   is_const_real(p + 3) && is_call(p + 4, SB_FN_DIV)) {
            jb->c0 = p[3].re;
            jb->inner = f64 ? inner_add2div_f64 : inner_add2div_f32;
        } else if (real && jb->M == 5 && n == 7 && is_arg(p, 0) && is_arg(p + 1, 1) && is_call(p + 2, SB_FN_ADD) &&
                   is_arg(p + 3, 2) && is_call(p + 4, SB_FN_ADD) && is_arg(p + 5, 3) && is_call(p + 6, SB_FN_ADD)) {
            jb->inner = f64 ? inner_sum4_f64 : inner_sum4_f32;
        }
    } else if (jb->op == SB_OP_ADD && real && jb->M == 2) {
        if (n == 2 && is_arg(p, 0) && is_call(p + 1, SB_FN_ABS2)) jb->inner = f64 ? inner_abs2add_f64 : inner_abs2add_f32;
        else if (n == 0 || (n == 1 && is_arg(p, 0))) jb->inner = f64 ? inner_idadd_f64 : inner_idadd_f32;
    }
}

/* ------------------------------------------------------------------------------------------------
 * planner helpers
 * ---------------------------------------------------------------------------------------------- */

/* indexorder  :427-441 */
static void indexorder(int N, const int64_t *s, int64_t *order)
{
    for (int i = 0; i < N; ++i) {
        int64_t si = llabs(s[i]);
        if (si == 0) { order[i] = 1; continue; }
        int64_t k = 1;
        for (int j = 0; j < N; ++j)
            if (s[j] != 0 && llabs(s[j]) < si) ++k;
        order[i] = k;
    }
}

/* _lastargmax  :452-460 */
static int lastargmax(int n, const int64_t *t)
{
    int i = 0;
    for (int j = 1; j < n; ++j)
        if (t[j] >= t[i]) i = j;
    return i;
}

/* totalmemoryregion  :503-520  (dims/strides given from index lo on) */
static int64_t totalmemoryregion(int N, int lo, int M, const int64_t *dims, int64_t bytestrides[][SB_MAX_DIMS])
{
    int64_t region = 0;
    for (int k = 0; k < M; ++k) {
        int64_t contig = 0, nblocks = 1;
        for (int i = lo; i < N; ++i) {
            int64_t s = bytestrides[k][i];
            if (s < CACHELINE) contig += (dims[i] - 1) * s;
            else nblocks *= dims[i];
        }
        contig = contig / CACHELINE + 1;
        region += CACHELINE * contig * nblocks;
    }
    return region;
}

/* _computeblocks  :463-500 ; the Julia recursion on tail(...) becomes the index `lo` */
static void computeblocks(int N, int lo, int M, const int64_t *dims, const int64_t *costs,
                          int64_t bytestrides[][SB_MAX_DIMS], int64_t strideorders[][SB_MAX_DIMS],
                          int64_t *blocks)
{
    if (lo >= N) return;
    if (totalmemoryregion(N, lo, M, dims, bytestrides) <= BLOCKMEMORYSIZE) { /* :474 */
        for (int i = lo; i < N; ++i) blocks[i] = dims[i];
        return;
    }
    int64_t minorder = strideorders[0][lo];
    for (int k = 0; k < M; ++k)
        for (int i = lo; i < N; ++i)
            if (strideorders[k][i] < minorder) minorder = strideorders[k][i];
    int allfirst = 1;
    for (int k = 0; k < M; ++k)
        if (strideorders[k][lo] != minorder) allfirst = 0;
    if (allfirst) { /* :478-483 */
        blocks[lo] = dims[lo];
        computeblocks(N, lo + 1, M, dims, costs, bytestrides, strideorders, blocks);
        return;
    }
    int64_t minbs = bytestrides[0][lo];
    for (int k = 0; k < M; ++k)
        for (int i = lo; i < N; ++i)
            if (bytestrides[k][i] < minbs) minbs = bytestrides[k][i];
    if (minbs > BLOCKMEMORYSIZE) { /* :485-487 */
        for (int i = lo; i < N; ++i) blocks[i] = 1;
        return;
    }
    int64_t b[SB_MAX_DIMS], w[SB_MAX_DIMS];
    for (int i = 0; i < N; ++i) b[i] = dims[i];
    const int n = N - lo;
    /* Deviation (termination only): with NEGATIVE strides the reference's costs are negative (signed `min` at :137,
     * SURVEY.md appendix E.2), `_lastargmax` can then keep choosing a dim whose block is already 1 and the reference's
     * `while` loops would not terminate.  Whenever the chosen dim cannot shrink, the largest remaining block is shrunk
     * instead; identical to the reference in every case where the reference terminates. */
    for (;;) { /* :491-494 */
        if (totalmemoryregion(N, lo, M, b, bytestrides) < 2 * BLOCKMEMORYSIZE) break;
        for (int i = 0; i < n; ++i) w[i] = (b[lo + i] - 1) * costs[lo + i];
        int i = lo + lastargmax(n, w);
        if (b[i] <= 1) {
            i = -1;
            for (int q = lo; q < N; ++q)
                if (b[q] > 1 && (i < 0 || b[q] >= b[i])) i = q;
            if (i < 0) break;
        }
        b[i] = (b[i] + 1) >> 1;
    }
    for (;;) { /* :495-498 */
        if (totalmemoryregion(N, lo, M, b, bytestrides) <= BLOCKMEMORYSIZE) break;
        for (int i = 0; i < n; ++i) w[i] = (b[lo + i] - 1) * costs[lo + i];
        int i = lo + lastargmax(n, w);
        if (b[i] <= 1) {
            i = -1;
            for (int q = lo; q < N; ++q)
                if (b[q] > 1 && (i < 0 || b[q] >= b[i])) i = q;
            if (i < 0) break;
        }
        b[i] -= 1;
    }
    for (int i = lo; i < N; ++i) blocks[i] = b[i];
}

static int64_t prod(int N, const int64_t *d)
{
    int64_t p = 1;
    for (int i = 0; i < N; ++i) p *= d[i];
    return p;
}

/* _length  :443-447 */
static int64_t outlength(int N, const int64_t *dims, const int64_t *s0)
{
    int64_t p = 1;
    for (int i = 0; i < N; ++i) p *= (s0[i] == 0 ? 1 : dims[i]);
    return p;
}

/* ------------------------------------------------------------------------------------------------
 * _mapreduce_kernel!  :229-425   (Appendix A of SURVEY.md shows the N=2, M=2 expansion)
 * ---------------------------------------------------------------------------------------------- */

/* initop pre-pass over the DISTINCT outputs of one block  (:351-375) */
static void init_nest(const ref_job *jb, int level, const int64_t *d, char *o)
{
    const int64_t dd = (jb->strides[0][level] == 0) ? 1 : d[level]; /* :356,:367 */
    if (level == 0) {
        for (int64_t j = 0; j < dd; ++j) jb->initel(jb, o + j * jb->bstr[0][0]);
        return;
    }
    for (int64_t j = 0; j < dd; ++j) init_nest(jb, level - 1, d, o + j * jb->bstr[0][level]);
}

/* inner loop nest over one block  (:339-349), dim N outermost ... dim 2, then the @simd line */
static void inner_nest(const ref_job *jb, int level, const int64_t *d, char **I)
{
    if (level == 0) {
        jb->inner(jb, I, d[0]);
        return;
    }
    char *P[SB_MAX_OPS];
    for (int k = 0; k < jb->M; ++k) P[k] = I[k];
    for (int64_t j = 0; j < d[level]; ++j) {
        inner_nest(jb, level - 1, d, P);
        for (int k = 0; k < jb->M; ++k) P[k] += jb->bstr[k][level]; /* :343-344 */
    }
}

/* block loops  (:385-394 / :403-414): level N-1 is the outermost */
static void block_loops(const ref_job *jb, int level, const int64_t *dims, int64_t *d, char **I, int init_above,
                        int use_init)
{
    if (level < 0) {
        if (use_init && init_above) init_nest(jb, jb->N - 1, d, I[0]); /* :377-379 */
        inner_nest(jb, jb->N - 1, d, I);
        return;
    }
    int init = init_above; /* :405 */
    char *P[SB_MAX_OPS];
    for (int k = 0; k < jb->M; ++k) P[k] = I[k];
    for (int64_t J = 0; J < dims[level]; J += jb->blocks[level]) {
        const int64_t rem = dims[level] - J;
        d[level] = jb->blocks[level] < rem ? jb->blocks[level] : rem; /* :407 */
        block_loops(jb, level - 1, dims, d, P, init, use_init);
        init = init && (jb->strides[0][level] > 0); /* :409 (quirk kept: '> 0') */
        for (int k = 0; k < jb->M; ++k) P[k] += d[level] * jb->bstr[k][level]; /* :303 */
    }
}

static void run_kernel(const ref_job *jb, const int64_t *dims, char **ptrs, int use_init)
{
    int64_t d[SB_MAX_DIMS];
    char *I[SB_MAX_OPS];
    for (int k = 0; k < jb->M; ++k) I[k] = ptrs[k];
    for (int i = 0; i < jb->N; ++i)
        if (dims[i] <= 0) return;
    block_loops(jb, jb->N - 1, dims, d, I, 1, use_init);
}

/* ------------------------------------------------------------------------------------------------
 * _mapreduce_threaded!  :195-227   (Threads.@spawn -> pthread_create, wait -> pthread_join)
 * ---------------------------------------------------------------------------------------------- */
typedef struct thr_arg {
    const ref_job *jb;
    int64_t dims[SB_MAX_DIMS];
    char *ptrs[SB_MAX_OPS];
    const int64_t *costs;
    int nthreads;
    int64_t spacing_bytes;
    int taskindex;
    int use_init;
} thr_arg;

static void threaded(const ref_job *jb, const int64_t *dims, char **ptrs, const int64_t *costs, int nthreads,
                     int64_t spacing_bytes, int taskindex, int use_init);

static void *thr_entry(void *p)
{
    thr_arg *a = (thr_arg *)p;
    threaded(a->jb, a->dims, a->ptrs, a->costs, a->nthreads, a->spacing_bytes, a->taskindex, a->use_init);
    return NULL;
}

static void leaf(const ref_job *jb, const int64_t *dims, char **ptrs, int64_t spacing_bytes, int taskindex,
                 int use_init)
{
    if (jb->rec) { /* plan-only: record the leaf */
        ref_plan_out *r = jb->rec;
        if (r->nleaves < REF_MAX_LEAVES) memcpy(r->leaf_dims[r->nleaves], dims, sizeof(int64_t) * SB_MAX_DIMS);
        r->nleaves++;
        return;
    }
    char *P[SB_MAX_OPS];
    for (int k = 0; k < jb->M; ++k) P[k] = ptrs[k];
    P[0] += spacing_bytes * (taskindex - 1); /* :199,:205 */
    run_kernel(jb, dims, P, use_init);
}

static void threaded(const ref_job *jb, const int64_t *dims, char **ptrs, const int64_t *costs, int nthreads,
                     int64_t spacing_bytes, int taskindex, int use_init)
{
    const int N = jb->N;
    if (nthreads == 1 || prod(N, dims) <= MINTHREADLENGTH) { /* :198 */
        leaf(jb, dims, ptrs, spacing_bytes, taskindex, use_init);
        return;
    }
    int64_t w[SB_MAX_DIMS];
    for (int i = 0; i < N; ++i) w[i] = (dims[i] - 1) * costs[i];
    const int i = lastargmax(N, w); /* :203 */
    const int64_t lim = jb->blocks[i] < 1024 ? jb->blocks[i] : 1024;
    if (costs[i] == 0 || dims[i] <= lim) { /* :204 */
        leaf(jb, dims, ptrs, spacing_bytes, taskindex, use_init);
        return;
    }
    const int64_t di = dims[i], ndi = di >> 1; /* :209-210 */
    const int nn = nthreads >> 1;
    thr_arg a;
    a.jb = jb;
    memcpy(a.dims, dims, sizeof a.dims);
    a.dims[i] = ndi;
    for (int k = 0; k < jb->M; ++k) a.ptrs[k] = ptrs[k];
    a.costs = costs;
    a.nthreads = nn;
    a.spacing_bytes = spacing_bytes;
    a.taskindex = taskindex;
    a.use_init = use_init;
    pthread_t th;
    int spawned = 0;
    if (!jb->rec) spawned = (pthread_create(&th, NULL, thr_entry, &a) == 0);
    if (!spawned) thr_entry(&a);
    int64_t d2[SB_MAX_DIMS];
    char *p2[SB_MAX_OPS];
    memcpy(d2, dims, sizeof d2);
    d2[i] = di - ndi; /* :219 */
    for (int k = 0; k < jb->M; ++k) p2[k] = ptrs[k] + ndi * jb->bstr[k][i]; /* :217-218 */
    threaded(jb, d2, p2, costs, nthreads - nn, spacing_bytes, taskindex + nn, use_init);
    if (spawned) pthread_join(th, NULL);
}

/* ------------------------------------------------------------------------------------------------
 * _mapreduce_fuse! -> _mapreduce_order! -> _mapreduce_block!
 * ---------------------------------------------------------------------------------------------- */
static int setup(const sb_desc *D, ref_job *jb, ref_plan_out *rec)
{
    memset(jb, 0, sizeof *jb);
    if (!D) return fail(SB_E_INVALID, "null desc");
    if (D->ndim < 0 || D->ndim > SB_MAX_DIMS) return fail(SB_E_INVALID, "ndim out of range");
    if (D->nops < 1 || D->nops > SB_MAX_OPS) return fail(SB_E_INVALID, "nops out of range");
    if (D->ntok < 0 || D->ntok > SB_MAX_TOKENS) return fail(SB_E_INVALID, "ntok out of range");
    const int M = D->nops;
    int N = D->ndim;
    int64_t dims[SB_MAX_DIMS], S[SB_MAX_OPS][SB_MAX_DIMS];
    for (int i = 0; i < N; ++i) {
        if (D->dims[i] < 0) return fail(SB_E_SHAPE, "negative dim");
        dims[i] = D->dims[i];
        for (int k = 0; k < M; ++k) S[k][i] = D->strides[k][i];
    }
    if (N == 0) { /* rank-0: one element */
        N = 1;
        dims[0] = 1;
        for (int k = 0; k < M; ++k) S[k][0] = 1;
    }
    /* program sanity + compute type (promote over inputs and typed constants) */
    int cplx = 0, dbl = 0, sp = 0;
    for (int k = (M > 1 ? 1 : 0); k < M; ++k) {
        if (D->dtype[k] < 0 || D->dtype[k] > 3) return fail(SB_E_INVALID, "bad dtype");
        if (D->dtype[k] == SB_C32 || D->dtype[k] == SB_C64) cplx = 1;
        if (D->dtype[k] == SB_F64 || D->dtype[k] == SB_C64) dbl = 1;
    }
    for (int i = 0; i < D->ntok; ++i) {
        const sb_tok *t = &D->prog[i];
        if (t->kind == SB_TOK_ARG) {
            if (t->a < 0 || t->a >= M - 1) return fail(SB_E_INVALID, "ARG index out of range");
            ++sp;
        } else if (t->kind == SB_TOK_CONST) {
            if (t->im != 0.0) cplx = 1;
            if (t->a == 2) dbl = 1;
            ++sp;
        } else if (t->kind == SB_TOK_CALL) {
            const int ar = t->a < 32 ? 1 : 2;
            if (sp < ar) return fail(SB_E_INVALID, "program stack underflow");
            sp -= ar - 1;
        } else
            return fail(SB_E_INVALID, "bad token");
    }
    if (D->ntok > 0 && sp != 1) return fail(SB_E_INVALID, "program leaves != 1 value");
    if (D->ntok == 0 && M < 2) return fail(SB_E_INVALID, "identity program needs an input");
    if (D->op != SB_OP_NONE) { /* the output participates in op(...) */
        if (D->dtype[0] == SB_C32 || D->dtype[0] == SB_C64) cplx = 1;
        if (D->dtype[0] == SB_F64 || D->dtype[0] == SB_C64) dbl = 1;
    }
    jb->ct = (cplx ? 2 : 0) + (dbl ? 1 : 0);

    /* _mapreduce_fuse!  :103-115 */
    for (int i = N - 1; i >= 1; --i) {
        int merge = 1;
        for (int k = 0; k < M; ++k)
            if (S[k][i] != dims[i - 1] * S[k][i - 1]) { merge = 0; break; }
        if (merge) {
            dims[i - 1] = dims[i - 1] * dims[i];
            dims[i] = 1;
        }
    }
    if (rec) {
        memset(rec, 0, sizeof *rec);
        rec->ndim = N;
        for (int i = 0; i < N; ++i) rec->fused_dims[i] = dims[i];
    }

    /* _mapreduce_order!  :121-137 */
    int g = 0;
    for (unsigned v = (unsigned)(M + 1); v; v >>= 1) ++g; /* 8*sizeof(Int) - leading_zeros(M+1) */
    int64_t imp[SB_MAX_DIMS], ord[SB_MAX_DIMS];
    indexorder(N, S[0], ord);
    for (int i = 0; i < N; ++i) imp[i] = 2 * ((int64_t)1 << (g * (N - ord[i])));
    for (int k = 1; k < M; ++k) {
        indexorder(N, S[k], ord);
        for (int i = 0; i < N; ++i) imp[i] += (int64_t)1 << (g * (N - ord[i]));
    }
    for (int i = 0; i < N; ++i) imp[i] *= (dims[i] > 1); /* :132 */
    int p[SB_MAX_DIMS];
    for (int i = 0; i < N; ++i) p[i] = i;
    for (int i = 1; i < N; ++i) { /* stable insertion sort, descending importance  (:133) */
        int v = p[i], j = i - 1;
        while (j >= 0 && imp[p[j]] < imp[v]) { p[j + 1] = p[j]; --j; }
        p[j + 1] = v;
    }
    jb->N = N;
    jb->M = M;
    for (int i = 0; i < N; ++i) {
        jb->dims[i] = dims[p[i]];
        for (int k = 0; k < M; ++k) jb->strides[k][i] = S[k][p[i]];
    }
    for (int k = 0; k < M; ++k) {
        jb->dtype[k] = D->dtype[k];
        jb->conj[k] = D->conj[k];
        jb->esize[k] = ESIZE[D->dtype[k]];
        jb->base[k] = (char *)D->base[k];
        for (int i = 0; i < N; ++i) jb->bstr[k][i] = jb->strides[k][i] * jb->esize[k]; /* :144 */
    }
    for (int i = 0; i < N; ++i) { /* :137 */
        int64_t m = jb->strides[0][i];
        for (int k = 1; k < M; ++k)
            if (jb->strides[k][i] < m) m = jb->strides[k][i];
        jb->costs[i] = (m == 0) ? 1 : (m << 1);
    }
    jb->ntok = D->ntok;
    memcpy(jb->prog, D->prog, sizeof(sb_tok) * (size_t)D->ntok);
    jb->op = D->op;
    jb->initop = D->initop;
    jb->init_re = D->init_re;
    jb->init_im = D->init_im;

    /* _mapreduce_block!  :144-146 */
    int64_t so[SB_MAX_OPS][SB_MAX_DIMS];
    for (int k = 0; k < M; ++k) indexorder(N, jb->strides[k], so[k]);
    computeblocks(N, 0, M, jb->dims, jb->costs, jb->bstr, so, jb->blocks);
    pick_inner(jb);

    if (rec) {
        for (int i = 0; i < N; ++i) {
            rec->importance[i] = imp[i];
            rec->perm[i] = p[i] + 1;
            rec->dims[i] = jb->dims[i];
            rec->costs[i] = jb->costs[i];
            rec->blocks[i] = jb->blocks[i];
            for (int k = 0; k < M; ++k) rec->strides[k][i] = jb->strides[k][i];
        }
        rec->region_bytes = totalmemoryregion(N, 0, M, jb->blocks, jb->bstr);
    }
    return SB_OK;
}

/* _mapreduce_block!  :151-178 */
static int block_dispatch(ref_job *jb, int nthreads, ref_plan_out *rec)
{
    const int N = jb->N;
    const int use_init = jb->initop != SB_INIT_NONE && jb->op != SB_OP_NONE;
    jb->rec = rec;
    for (int i = 0; i < N; ++i)
        if (jb->dims[i] == 0) return SB_OK;
    if (nthreads <= 1 || prod(N, jb->dims) <= MINTHREADLENGTH) { /* :151 */
        leaf(jb, jb->dims, jb->base, 0, 1, use_init);
        return SB_OK;
    }
    if (jb->op != SB_OP_NONE && outlength(N, jb->dims, jb->strides[0]) == 1) { /* :153 complete reduction */
        if (rec) rec->complete_reduction = 1;
        const int es = jb->esize[0];
        const int64_t spacing = 64 / es > 1 ? 64 / es : 1; /* :155 */
        const int64_t spacing_bytes = spacing * es;
        char *slots = NULL;
        char *ptrs[SB_MAX_OPS];
        for (int k = 0; k < jb->M; ++k) ptrs[k] = jb->base[k];
        if (!rec) {
            slots = (char *)aligned_alloc(64, (size_t)(spacing_bytes * nthreads + 64));
            if (!slots) return fail(SB_E_NOMEM, "alloc");
            static void (*const neutral[4])(const ref_job *, const char *, char *, int, int64_t) = {
                neutral_f32, neutral_f64, neutral_c32, neutral_c64};
            neutral[jb->ct](jb, jb->base[0], slots, nthreads, spacing_bytes); /* :157-161 */
            ptrs[0] = slots;
        }
        /* slots are plain (un-conjugated) storage */
        ref_job jb2 = *jb;
        jb2.conj[0] = 0;
        jb2.rec = rec;
        threaded(&jb2, jb->dims, ptrs, jb->costs, nthreads, spacing_bytes, 1, 0); /* :164 (initop=nothing) */
        if (!rec) {
            static void (*const fold[4])(const ref_job *, char *, const char *, int, int64_t) = {fold_f32, fold_f64,
                                                                                                  fold_c32, fold_c64};
            fold[jb->ct](jb, jb->base[0], slots, nthreads, spacing_bytes); /* :167-170 */
            free(slots);
        }
        return SB_OK;
    }
    int64_t costs[SB_MAX_DIMS];
    for (int i = 0; i < N; ++i) costs[i] = jb->costs[i] * (jb->strides[0][i] != 0); /* :172 */
    threaded(jb, jb->dims, jb->base, costs, nthreads, 0, 1, use_init);             /* :176 */
    return SB_OK;
}

/* map! with any zero dim returns early (:48); _mapreducedim! with a zero dim applies initop to a
 * non-empty output (:88-91). */
static int handle_empty(const sb_desc *D, int *handled)
{
    *handled = 0;
    int anyzero = 0;
    for (int i = 0; i < D->ndim; ++i)
        if (D->dims[i] == 0) anyzero = 1;
    if (!anyzero) return SB_OK;
    *handled = 1;
    if (D->op == SB_OP_NONE || D->initop == SB_INIT_NONE) return SB_OK;
    /* output elements: dims where the output stride is non-zero; empty if one of those is 0 */
    sb_desc E;
    memset(&E, 0, sizeof E);
    E.ndim = 0;
    for (int i = 0; i < D->ndim; ++i) {
        if (D->strides[0][i] == 0 && D->dims[i] != 1) continue; /* reduced (or zero-size reduced) dim */
        if (D->dims[i] == 0) return SB_OK;                       /* empty output */
        E.dims[E.ndim] = D->dims[i];
        E.strides[0][E.ndim] = D->strides[0][i];
        E.strides[1][E.ndim] = D->strides[0][i];
        E.ndim++;
    }
    E.nops = 2;
    E.base[0] = E.base[1] = D->base[0];
    E.dtype[0] = E.dtype[1] = D->dtype[0];
    E.conj[0] = E.conj[1] = D->conj[0];
    E.op = SB_OP_NONE;
    E.initop = SB_INIT_NONE;
    switch (D->initop) { /* map!(initop, out, out) */
    case SB_INIT_ZERO: E.ntok = 1; E.prog[0] = (sb_tok){SB_TOK_CONST, 0, 0.0, 0.0}; break;
    case SB_INIT_CONST: E.ntok = 1; E.prog[0] = (sb_tok){SB_TOK_CONST, 0, D->init_re, D->init_im}; break;
    case SB_INIT_SCALE:
        E.ntok = 3;
        E.prog[0] = (sb_tok){SB_TOK_CONST, 0, D->init_re, D->init_im};
        E.prog[1] = (sb_tok){SB_TOK_ARG, 0, 0, 0};
        E.prog[2] = (sb_tok){SB_TOK_CALL, SB_FN_MUL, 0, 0};
        break;
    case SB_INIT_CONJ:
        E.ntok = 2;
        E.prog[0] = (sb_tok){SB_TOK_ARG, 0, 0, 0};
        E.prog[1] = (sb_tok){SB_TOK_CALL, SB_FN_CONJ, 0, 0};
        break;
    default: return SB_OK;
    }
    return ref_mapreduce(&E, 1);
}

int ref_plan(const sb_desc *desc, int nthreads, ref_plan_out *out)
{
    ref_job jb;
    if (!out) return fail(SB_E_INVALID, "null out");
    int rc = setup(desc, &jb, out);
    if (rc != SB_OK) return rc;
    return block_dispatch(&jb, nthreads, out);
}

int ref_mapreduce(const sb_desc *desc, int nthreads)
{
    ref_job jb;
    int handled = 0;
    if (!desc) return fail(SB_E_INVALID, "null desc");
    for (int i = 0; i < desc->ndim && i < SB_MAX_DIMS; ++i)
        if (desc->dims[i] < 0) return fail(SB_E_SHAPE, "negative dim");
    int rc = handle_empty(desc, &handled);
    if (rc != SB_OK || handled) return rc;
    rc = setup(desc, &jb, NULL);
    if (rc != SB_OK) return rc;
    if (jb.op == SB_OP_NONE && jb.M < 2 && jb.ntok == 0) return fail(SB_E_INVALID, "map without inputs");
    return block_dispatch(&jb, nthreads, NULL);
}
