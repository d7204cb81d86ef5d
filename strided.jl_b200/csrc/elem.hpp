// elem.hpp -- element types, typed loads/stores and the element function set (host/device neutral).
//
// Arithmetic follows Julia's scalar semantics for the eltypes on the path (Float32/64, ComplexF32/64):
// one IEEE rounding per node, no mul-add contraction (the build passes --fmad=false; Julia never
// contracts `a*x + y`, reference src/linalg.jl:28), n-ary `+` is a left fold, `min`/`max` propagate NaN
// and order signed zeros.
#pragma once
#include "common.hpp"
#if !defined(__CUDACC_RTC__)
#include <math.h>
#endif

namespace sb {

template <class R> struct alignas(2 * sizeof(R)) cx {
    R re, im;
};

template <class T> struct traits;
template <> struct traits<float> {
    using real = float;
    static constexpr bool cplx = false;
    static constexpr int dt = F32;
};
template <> struct traits<double> {
    using real = double;
    static constexpr bool cplx = false;
    static constexpr int dt = F64;
};
template <> struct traits<cx<float>> {
    using real = float;
    static constexpr bool cplx = true;
    static constexpr int dt = C32;
};
template <> struct traits<cx<double>> {
    using real = double;
    static constexpr bool cplx = true;
    static constexpr int dt = C64;
};

// ---- construction / access ----------------------------------------------------------------------------
template <class T> SB_HD T make(double re, double im);
template <> SB_HD float make<float>(double re, double) { return (float)re; }
template <> SB_HD double make<double>(double re, double) { return re; }
template <> SB_HD cx<float> make<cx<float>>(double re, double im) { return cx<float>{(float)re, (float)im}; }
template <> SB_HD cx<double> make<cx<double>>(double re, double im) { return cx<double>{re, im}; }

SB_HD float re_of(float x) { return x; }
SB_HD double re_of(double x) { return x; }
template <class R> SB_HD R re_of(cx<R> x) { return x.re; }
SB_HD float im_of(float) { return 0.f; }
SB_HD double im_of(double) { return 0.0; }
template <class R> SB_HD R im_of(cx<R> x) { return x.im; }

// ---- complex arithmetic (Julia's formulas: base/complex.jl `*`, `/` use the textbook product and
// Smith-like scaling; here product is textbook, division is the robust scaled form) ---------------------
template <class R> SB_HD cx<R> operator+(cx<R> a, cx<R> b) { return cx<R>{a.re + b.re, a.im + b.im}; }
template <class R> SB_HD cx<R> operator-(cx<R> a, cx<R> b) { return cx<R>{a.re - b.re, a.im - b.im}; }
template <class R> SB_HD cx<R> operator-(cx<R> a) { return cx<R>{-a.re, -a.im}; }
template <class R> SB_HD cx<R> operator*(cx<R> a, cx<R> b)
{
    return cx<R>{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
template <class R> SB_HD cx<R> operator/(cx<R> a, cx<R> b)
{
    // Smith's algorithm
    if (fabs((double)b.re) >= fabs((double)b.im)) {
        R r = b.im / b.re, d = b.re + b.im * r;
        return cx<R>{(a.re + a.im * r) / d, (a.im - a.re * r) / d};
    } else {
        R r = b.re / b.im, d = b.re * r + b.im;
        return cx<R>{(a.re * r + a.im) / d, (a.im * r - a.re) / d};
    }
}

// ---- Julia min/max ------------------------------------------------------------------------------------
template <class R> SB_HD bool sb_signbit(R x);
template <> SB_HD bool sb_signbit<float>(float x)
{
#if defined(__CUDA_ARCH__)
    return (__float_as_uint(x) >> 31) != 0;
#else
    return signbit(x);
#endif
}
template <> SB_HD bool sb_signbit<double>(double x)
{
#if defined(__CUDA_ARCH__)
    return (__double2hiint(x) >> 31) != 0;
#else
    return signbit(x);
#endif
}
template <class R> SB_HD R jl_max(R a, R b)
{
    if (a != a) return a;
    if (b != b) return b;
    if (a == b) return sb_signbit(a) ? b : a;
    return a > b ? a : b;
}
template <class R> SB_HD R jl_min(R a, R b)
{
    if (a != a) return a;
    if (b != b) return b;
    if (a == b) return sb_signbit(a) ? a : b;
    return a < b ? a : b;
}

// ---- real math wrappers (float versions call the f-suffixed functions) ---------------------------------
SB_HD float m_sqrt(float x) { return sqrtf(x); }
SB_HD double m_sqrt(double x) { return sqrt(x); }
SB_HD float m_exp(float x) { return expf(x); }
SB_HD double m_exp(double x) { return exp(x); }
SB_HD float m_log(float x) { return logf(x); }
SB_HD double m_log(double x) { return log(x); }
SB_HD float m_sin(float x) { return sinf(x); }
SB_HD double m_sin(double x) { return sin(x); }
SB_HD float m_cos(float x) { return cosf(x); }
SB_HD double m_cos(double x) { return cos(x); }
SB_HD float m_tanh(float x) { return tanhf(x); }
SB_HD double m_tanh(double x) { return tanh(x); }
SB_HD float m_abs(float x) { return fabsf(x); }
SB_HD double m_abs(double x) { return fabs(x); }
SB_HD float m_hypot(float x, float y) { return hypotf(x, y); }
SB_HD double m_hypot(double x, double y) { return hypot(x, y); }
SB_HD float m_sinh(float x) { return sinhf(x); }
SB_HD double m_sinh(double x) { return sinh(x); }
SB_HD float m_cosh(float x) { return coshf(x); }
SB_HD double m_cosh(double x) { return cosh(x); }
SB_HD float m_atan2(float y, float x) { return atan2f(y, x); }
SB_HD double m_atan2(double y, double x) { return atan2(y, x); }

// ---- unary / binary element functions -----------------------------------------------------------------
template <class R> SB_HD R call1(int fn, R x)
{
    switch (fn) {
    case FN_NEG: return -x;
    case FN_ABS: return m_abs(x);
    case FN_ABS2: return x * x;
    case FN_IMAG: return (R)0;
    case FN_SQRT: return m_sqrt(x);
    case FN_EXP: return m_exp(x);
    case FN_LOG: return m_log(x);
    case FN_SIN: return m_sin(x);
    case FN_COS: return m_cos(x);
    case FN_TANH: return m_tanh(x);
    case FN_INV: return (R)1 / x;
    default: return x; // identity, conj, real
    }
}
template <class R> SB_HD cx<R> call1(int fn, cx<R> x)
{
    switch (fn) {
    case FN_NEG: return -x;
    case FN_CONJ: return cx<R>{x.re, -x.im};
    case FN_ABS: return cx<R>{m_hypot(x.re, x.im), (R)0};
    case FN_ABS2: return cx<R>{x.re * x.re + x.im * x.im, (R)0};
    case FN_REAL: return cx<R>{x.re, (R)0};
    case FN_IMAG: return cx<R>{x.im, (R)0};
    case FN_SQRT: {
        R m = m_hypot(x.re, x.im);
        if (m == (R)0) return cx<R>{(R)0, x.im};
        R s = m_sqrt((m + m_abs(x.re)) * (R)0.5);
        if (x.re >= (R)0) return cx<R>{s, x.im / (s + s)};
        R im = sb_signbit(x.im) ? -s : s;
        return cx<R>{m_abs(x.im) / (s + s), im};
    }
    case FN_EXP: {
        R e = m_exp(x.re);
        if (x.im == (R)0) return cx<R>{e, x.im}; // Julia's exp(::Complex): a zero imaginary part stays zero (exp(x) may be Inf)
        return cx<R>{e * m_cos(x.im), e * m_sin(x.im)};
    }
    case FN_LOG: return cx<R>{m_log(m_hypot(x.re, x.im)), m_atan2(x.im, x.re)};
    case FN_SIN: return cx<R>{m_sin(x.re) * m_cosh(x.im), m_cos(x.re) * m_sinh(x.im)};
    case FN_COS: return cx<R>{m_cos(x.re) * m_cosh(x.im), -(m_sin(x.re) * m_sinh(x.im))};
    case FN_TANH: {
        R a = m_tanh(x.re), t = (R)(sin((double)x.im) / cos((double)x.im));
        cx<R> num{a, t}, den{(R)1, a * t};
        return num / den;
    }
    case FN_INV: return cx<R>{(R)1, (R)0} / x;
    default: return x;
    }
}
template <class T> SB_HD T call2(int fn, T x, T y)
{
    using R = typename traits<T>::real;
    switch (fn) {
    case FN_ADD: return x + y;
    case FN_SUB: return x - y;
    case FN_MUL: return x * y;
    case FN_DIV: return x / y;
    case FN_MAX: return make<T>((double)jl_max<R>(re_of(x), re_of(y)), 0.0);
    case FN_MIN: return make<T>((double)jl_min<R>(re_of(x), re_of(y)), 0.0);
    case FN_LT: return make<T>(re_of(x) < re_of(y) ? 1.0 : 0.0, 0.0);
    default: return x;
    }
}

template <class T> SB_HD T conj_of(T x) { return x; }
template <class R> SB_HD cx<R> conj_of(cx<R> x) { return cx<R>{x.re, -x.im}; }

// ---- CUDA vector type of the same size (for cache-hinted accesses) ---------------------------------------
template <class T> struct vec_of;
#if defined(__CUDACC__)
template <> struct vec_of<float> { using type = float; };
template <> struct vec_of<double> { using type = double; };
template <> struct vec_of<cx<float>> { using type = float2; };
template <> struct vec_of<cx<double>> { using type = double2; };
SB_HD float to_vec(float x) { return x; }
SB_HD double to_vec(double x) { return x; }
SB_HD float2 to_vec(cx<float> x) { return make_float2(x.re, x.im); }
SB_HD double2 to_vec(cx<double> x) { return make_double2(x.re, x.im); }
#endif

// ---- typed global loads/stores -------------------------------------------------------------------------
// `p` is the BYTE address of the element.  `UNIFORM`: storage type == compute type and no conj flag: a plain
// typed access (8/16-byte types become one LDG.64/LDG.128).
template <class CT, bool UNIFORM> SB_HD CT load_elem(const unsigned char *p, int dtype, int cj)
{
    if (UNIFORM) return *reinterpret_cast<const CT *>(p);
    CT v;
    switch (dtype) {
    case F32: v = make<CT>((double)*reinterpret_cast<const float *>(p), 0.0); break;
    case F64: v = make<CT>(*reinterpret_cast<const double *>(p), 0.0); break;
    case C32: v = make<CT>((double)reinterpret_cast<const float *>(p)[0], (double)reinterpret_cast<const float *>(p)[1]); break;
    default: v = make<CT>(reinterpret_cast<const double *>(p)[0], reinterpret_cast<const double *>(p)[1]); break;
    }
    return cj ? conj_of(v) : v;
}

template <class CT, bool UNIFORM> SB_HD void store_elem(unsigned char *p, int dtype, int cj, CT v)
{
    if (UNIFORM) {
#if defined(__CUDA_ARCH__)
        __stcs(reinterpret_cast<typename vec_of<CT>::type *>(p), to_vec(v)); // streaming store: the output is never re-read
#else
        *reinterpret_cast<CT *>(p) = v;
#endif
        return;
    }
    if (cj) v = conj_of(v);
    switch (dtype) {
    case F32: *reinterpret_cast<float *>(p) = (float)re_of(v); break;
    case F64: *reinterpret_cast<double *>(p) = (double)re_of(v); break;
    case C32:
        reinterpret_cast<float *>(p)[0] = (float)re_of(v);
        reinterpret_cast<float *>(p)[1] = (float)im_of(v);
        break;
    default:
        reinterpret_cast<double *>(p)[0] = (double)re_of(v);
        reinterpret_cast<double *>(p)[1] = (double)im_of(v);
        break;
    }
}

SB_HD double sb_inf()
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double(0x7ff0000000000000LL);
#else
    return (double)INFINITY;
#endif
}

// ---- reduction operator / initop ----------------------------------------------------------------------
template <class T> SB_HD T red_apply(int op, T a, T b)
{
    using R = typename traits<T>::real;
    switch (op) {
    case OP_ADD: return a + b;
    case OP_MUL: return a * b;
    case OP_MIN: return make<T>((double)jl_min<R>(re_of(a), re_of(b)), 0.0);
    default: return make<T>((double)jl_max<R>(re_of(a), re_of(b)), 0.0);
    }
}
// neutral elements of _init_reduction! (reference src/mapreduce.jl:182-187); min/max use +-Inf here
// because every partial is later folded with the (initialised) output value.
template <class T> SB_HD T red_neutral(int op)
{
    switch (op) {
    case OP_ADD: return make<T>(0.0, 0.0);
    case OP_MUL: return make<T>(1.0, 0.0);
    case OP_MIN: return make<T>(sb_inf(), 0.0);
    default: return make<T>(-sb_inf(), 0.0);
    }
}
template <class T> SB_HD T init_apply(int initop, double bre, double bim, T x)
{
    switch (initop) {
    case INIT_ZERO: return make<T>(0.0, 0.0);
    case INIT_SCALE: return make<T>(bre, bim) * x;
    case INIT_CONST: return make<T>(bre, bim);
    case INIT_CONJ: return conj_of(x);
    default: return x;
    }
}

} // namespace sb
