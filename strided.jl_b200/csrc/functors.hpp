// functors.hpp -- the element function `f` of a plan: pre-instantiated recipes and the postfix interpreter.
//
// The reference evaluates the CaptureArgs tree built by make_capture (src/broadcast.jl:75-83) through
// `consume` (:86-98); Julia inlines it.  Here the planner flattens the tree to postfix tokens (sb_tok,
// include/strided_b200.h) and either matches a recipe (compiled functor below) or runs the interpreter.
#pragma once
#include "elem.hpp"

namespace sb {

template <int RC> struct RecipeTag {
    static constexpr int value = RC;
};

// every functor: CT eval<NIN>(const Program&, const CT* args) -- args[k] = value of input k, k < NIN
template <class CT, int RC> struct ElemFn;

template <class CT> struct ElemFn<CT, RC_COPY> {
    template <int NIN> SB_HD CT eval(const Program &, const CT *a) const { return a[0]; }
};
template <class CT> struct ElemFn<CT, RC_SCALE> {
    template <int NIN> SB_HD CT eval(const Program &p, const CT *a) const { return make<CT>(p.c0re, p.c0im) * a[0]; }
};
template <class CT> struct ElemFn<CT, RC_ADD2> {
    template <int NIN> SB_HD CT eval(const Program &, const CT *a) const { return a[0] + a[1]; }
};
template <class CT> struct ElemFn<CT, RC_ADD2_DIV> {
    template <int NIN> SB_HD CT eval(const Program &p, const CT *a) const { return (a[0] + a[1]) / make<CT>(p.c0re, p.c0im); }
};
template <class CT> struct ElemFn<CT, RC_ADD2_MUL> {
    template <int NIN> SB_HD CT eval(const Program &p, const CT *a) const { return (a[0] + a[1]) * make<CT>(p.c0re, p.c0im); }
};
template <class CT> struct ElemFn<CT, RC_SUM3> {
    template <int NIN> SB_HD CT eval(const Program &, const CT *a) const { return (a[0] + a[1]) + a[2]; }
};
template <class CT> struct ElemFn<CT, RC_SUM4> {
    template <int NIN> SB_HD CT eval(const Program &, const CT *a) const { return ((a[0] + a[1]) + a[2]) + a[3]; }
};
template <class CT> struct ElemFn<CT, RC_AXPY> {
    template <int NIN> SB_HD CT eval(const Program &p, const CT *a) const { return make<CT>(p.c0re, p.c0im) * a[0] + a[1]; }
};
template <class CT> struct ElemFn<CT, RC_AXPBY> {
    template <int NIN> SB_HD CT eval(const Program &p, const CT *a) const
    {
        return make<CT>(p.c0re, p.c0im) * a[0] + make<CT>(p.c1re, p.c1im) * a[1];
    }
};
template <class CT> struct ElemFn<CT, RC_ABS2> {
    template <int NIN> SB_HD CT eval(const Program &, const CT *a) const { return call1(FN_ABS2, a[0]); }
};

template <class CT> struct ElemFn<CT, RC_S_ABS> {
    template <int NIN> SB_HD CT eval(const Program &, const CT *a) const { return call1(FN_ABS, a[0]); }
};
template <class CT> struct ElemFn<CT, RC_S_MUL2> {
    template <int NIN> SB_HD CT eval(const Program &, const CT *a) const { return call2(FN_MUL, a[0], a[NIN > 1 ? 1 : 0]); }
};

// Generic interpreter.  The value stack lives in registers: depth is bounded by 4 (the planner rejects
// deeper programs with SB_E_UNSUPPORTED) and pushes/pops shift a fixed window, so no local memory is used.
template <class CT> struct ElemFn<CT, RC_INTERP> {
    template <int NIN> SB_HD CT eval(const Program &p, const CT *a) const
    {
        CT s0 = a[0], s1 = s0, s2 = s0, s3 = s0;
        const int n = p.ntok;
        for (int i = 0; i < n; ++i) {
            const int kind = p.tok[i].kind, x = p.tok[i].a;
            if (kind == TOK_CALL) {
                if (x < 32) {
                    s0 = call1(x, s0);
                } else {
                    s0 = call2(x, s1, s0);
                    s1 = s2;
                    s2 = s3;
                }
            } else {
                s3 = s2;
                s2 = s1;
                s1 = s0;
                if (kind == TOK_ARG) {
                    CT v = a[0];
#pragma unroll
                    for (int k = 1; k < NIN; ++k)
                        if (x == k) v = a[k];
                    s0 = v;
                } else {
                    s0 = make<CT>(p.tok[i].re, p.tok[i].im);
                }
            }
        }
        return s0;
    }
};

} // namespace sb
