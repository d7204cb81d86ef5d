// planner.cpp -- see planner.hpp.  Pure host C++ (no CUDA): unit-tested on CPU through sb_plan_describe and
// through the thread-grid emulator in tests/emul/.
#include "planner.hpp"
#include "tma_tile.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <vector>

namespace sb {

namespace {

struct Canon {
    int ndim = 0, nops = 0, nkept = 0;
    int64_t dims[MAXD] = {0};
    int64_t strides[MAXO][MAXD] = {{0}};
    unsigned char *base[MAXO] = {nullptr};
    int dtype[MAXO] = {0};
    int conj[MAXO] = {0};
    int op = OP_NONE, initop = INIT_NONE;
    double init_re = 0, init_im = 0;
    int ntok = 0;
    Tok tok[MAXTOK];
    int ct = F64;
    int src[MAXO] = {0, 1, 2, 3, 4, 5, 6, 7}; // canonical operand -> sb_desc operand (plan cache re-binds bases)
    int depth = 0;                            // operand-stack depth the program needs (the interpreter keeps 4 in registers)
};

int64_t iabs64(int64_t x) { return x < 0 ? -x : x; }
int ilog2_ceil(int64_t x)
{
    int b = 0;
    while (((int64_t)1 << b) < x) ++b;
    return b;
}

bool is_cplx(int dt) { return dt == C32 || dt == C64; }
bool is_dbl(int dt) { return dt == F64 || dt == C64; }

// ---- program checks ---------------------------------------------------------------------------------
int check_program(const Canon &c, std::string &err, int &maxdepth)
{
    int sp = 0;
    maxdepth = 0;
    for (int i = 0; i < c.ntok; ++i) {
        const Tok &t = c.tok[i];
        if (t.kind == TOK_ARG) {
            if (t.a < 0 || t.a >= c.nops - 1) { err = "program: ARG index out of range"; return SB_E_INVALID; }
            ++sp;
        } else if (t.kind == TOK_CONST) {
            ++sp;
        } else if (t.kind == TOK_CALL) {
            const bool unary = (t.a >= FN_IDENTITY && t.a <= FN_INV);
            const bool binary = (t.a >= FN_ADD && t.a <= FN_LT);
            if (!unary && !binary) { err = "program: unknown function id"; return SB_E_INVALID; }
            const int ar = unary ? 1 : 2;
            if (sp < ar) { err = "program: stack underflow"; return SB_E_INVALID; }
            sp -= ar - 1;
        } else {
            err = "program: bad token kind";
            return SB_E_INVALID;
        }
        maxdepth = std::max(maxdepth, sp);
    }
    if (c.ntok > 0 && sp != 1) { err = "program: must leave exactly one value"; return SB_E_INVALID; }
    if (c.ntok == 0 && c.nops < 2) { err = "identity program needs an input"; return SB_E_INVALID; }
    return SB_OK;
}

bool tok_arg(const Tok &t, int a) { return t.kind == TOK_ARG && t.a == a; }
bool tok_call(const Tok &t, int fn) { return t.kind == TOK_CALL && t.a == fn; }
bool tok_const(const Tok &t) { return t.kind == TOK_CONST; }
bool is_pow2_double(double c)
{
    if (c == 0.0 || !std::isfinite(c)) return false;
    int e;
    const double m = std::frexp(std::fabs(c), &e);
    return m == 0.5 && e > -1000 && e < 1000;
}

void match_recipe(const Canon &c, Program &p)
{
    p.recipe = RC_INTERP;
    p.ntok = c.ntok;
    p.c0re = p.c0im = p.c1re = p.c1im = 0;
    for (int i = 0; i < c.ntok; ++i) p.tok[i] = c.tok[i];
    const Tok *t = c.tok;
    const int n = c.ntok, nin = c.nops - 1;
    auto set0 = [&](const Tok &k) { p.c0re = k.re; p.c0im = k.im; };
    auto set1 = [&](const Tok &k) { p.c1re = k.re; p.c1im = k.im; };
    if (nin == 1 && (n == 0 || (n == 1 && tok_arg(t[0], 0)) || (n == 2 && tok_arg(t[0], 0) && tok_call(t[1], FN_IDENTITY)))) {
        p.recipe = RC_COPY;
    } else if (nin == 1 && n == 3 && tok_const(t[0]) && tok_arg(t[1], 0) && tok_call(t[2], FN_MUL)) {
        p.recipe = RC_SCALE; set0(t[0]);
    } else if (nin == 1 && n == 3 && tok_arg(t[0], 0) && tok_const(t[1]) && tok_call(t[2], FN_MUL) && t[1].im == 0.0) {
        p.recipe = RC_SCALE; set0(t[1]); // x*c == c*x bitwise for real c
    } else if (nin == 1 && n == 2 && tok_arg(t[0], 0) && tok_call(t[1], FN_ABS2)) {
        p.recipe = RC_ABS2;
    } else if (nin == 2 && n == 3 && tok_arg(t[0], 0) && tok_arg(t[1], 1) && tok_call(t[2], FN_ADD)) {
        p.recipe = RC_ADD2;
    } else if (nin == 2 && n == 5 && tok_arg(t[0], 0) && tok_arg(t[1], 1) && tok_call(t[2], FN_ADD) && tok_const(t[3]) &&
               (tok_call(t[4], FN_DIV) || tok_call(t[4], FN_MUL))) {
        set0(t[3]);
        if (tok_call(t[4], FN_MUL)) p.recipe = RC_ADD2_MUL;
        else if (t[3].im == 0.0 && is_pow2_double(t[3].re)) { // x / 2^k == x * 2^-k exactly (both correctly rounded)
            p.recipe = RC_ADD2_MUL;
            p.c0re = 1.0 / t[3].re;
        } else p.recipe = RC_ADD2_DIV;
    } else if (nin == 3 && n == 5 && tok_arg(t[0], 0) && tok_arg(t[1], 1) && tok_call(t[2], FN_ADD) && tok_arg(t[3], 2) &&
               tok_call(t[4], FN_ADD)) {
        p.recipe = RC_SUM3;
    } else if (nin == 4 && n == 7 && tok_arg(t[0], 0) && tok_arg(t[1], 1) && tok_call(t[2], FN_ADD) && tok_arg(t[3], 2) &&
               tok_call(t[4], FN_ADD) && tok_arg(t[5], 3) && tok_call(t[6], FN_ADD)) {
        p.recipe = RC_SUM4;
    } else if (nin == 2 && n == 5 && tok_const(t[0]) && tok_arg(t[1], 0) && tok_call(t[2], FN_MUL) && tok_arg(t[3], 1) &&
               tok_call(t[4], FN_ADD)) {
        p.recipe = RC_AXPY; set0(t[0]);
    } else if (nin == 2 && n == 7 && tok_const(t[0]) && tok_arg(t[1], 0) && tok_call(t[2], FN_MUL) && tok_const(t[3]) &&
               tok_arg(t[4], 1) && tok_call(t[5], FN_MUL) && tok_call(t[6], FN_ADD)) {
        p.recipe = RC_AXPBY; set0(t[0]); set1(t[3]);
    }
}

// ---- canonicalisation ---------------------------------------------------------------------------------
int canonicalise(const sb_desc &d, Canon &c, bool &noop, std::string &err)
{
    noop = false;
    if (d.ndim < 0 || d.ndim > SB_MAX_DIMS) { err = "ndim out of range"; return SB_E_INVALID; }
    if (d.nops < 1 || d.nops > SB_MAX_OPS) { err = "nops out of range"; return SB_E_INVALID; }
    if (d.ntok < 0 || d.ntok > SB_MAX_TOKENS) { err = "ntok out of range"; return SB_E_INVALID; }
    if (d.op < SB_OP_NONE || d.op > SB_OP_MAX) { err = "bad op"; return SB_E_INVALID; }
    if (d.initop < SB_INIT_NONE || d.initop > SB_INIT_CONJ) { err = "bad initop"; return SB_E_INVALID; }
    for (int k = 0; k < d.nops; ++k)
        if (d.dtype[k] < SB_F32 || d.dtype[k] > SB_C64) { err = "bad dtype"; return SB_E_INVALID; }
    for (int i = 0; i < d.ndim; ++i)
        if (d.dims[i] < 0) { err = "negative dim"; return SB_E_SHAPE; }

    int opsel[MAXO] = {0, 1, 2, 3, 4, 5, 6, 7}; // canonical operand -> descriptor operand (identical inputs are merged below)
    c.nops = d.nops;
    c.op = d.op;
    c.initop = (d.op == SB_OP_NONE) ? (int)INIT_NONE : d.initop;
    c.init_re = d.init_re;
    c.init_im = d.init_im;
    c.ntok = d.ntok;
    for (int i = 0; i < d.ntok; ++i) {
        c.tok[i].kind = d.prog[i].kind;
        c.tok[i].a = d.prog[i].a;
        c.tok[i].re = d.prog[i].re;
        c.tok[i].im = d.prog[i].im;
    }
    for (int k = 0; k < d.nops; ++k) {
        c.base[k] = (unsigned char *)d.base[k];
        c.dtype[k] = d.dtype[k];
        c.conj[k] = d.conj[k] && is_cplx(d.dtype[k]);
    }
    if (c.op != OP_NONE && c.nops < 2) { err = "reduction without an input"; return SB_E_INVALID; }
    int depth = 0;
    int rc = check_program(c, err, depth);
    if (rc != SB_OK) return rc;
    // Identical input views (same base, eltype, conj flag and strides) are ONE operand: the broadcast front end captures
    // every occurrence of an array separately (`A .* exp.(-2 .* A) .+ sin.(A .* A)` arrives with four copies of A,
    // capturestridedargs, src/broadcast.jl:41-46); loading it once is free and keeps such trees inside SB_MAX_OPS.
    {
        int remap[MAXO], keep[MAXO], nk = 1;
        keep[0] = 0;
        for (int k = 1; k < c.nops; ++k) {
            int same = -1;
            for (int q = 1; q < nk && same < 0; ++q) {
                const int j = keep[q];
                bool eq = c.base[j] == c.base[k] && c.dtype[j] == c.dtype[k] && c.conj[j] == c.conj[k];
                for (int i = 0; i < d.ndim && eq; ++i)
                    if (d.dims[i] != 1 && d.strides[j][i] != d.strides[k][i]) eq = false;
                if (eq) same = q;
            }
            if (same >= 0) remap[k] = same;
            else {
                remap[k] = nk;
                keep[nk++] = k;
            }
        }
        if (nk < c.nops) {
            for (int i = 0; i < c.ntok; ++i)
                if (c.tok[i].kind == TOK_ARG) c.tok[i].a = remap[c.tok[i].a + 1] - 1;
            for (int q = 1; q < nk; ++q) {
                c.base[q] = c.base[keep[q]];
                c.dtype[q] = c.dtype[keep[q]];
                c.conj[q] = c.conj[keep[q]];
                c.src[q] = keep[q];
            }
            c.nops = nk;
        }
        for (int q = 0; q < MAXO; ++q) opsel[q] = q < nk ? keep[q] : q;
    }

    // compute type: promote inputs and typed constants (Julia promotes per node; see DESIGN.md "compute type")
    bool cplx = false, dbl = false;
    for (int k = (c.nops > 1 ? 1 : 0); k < c.nops; ++k) {
        cplx |= is_cplx(c.dtype[k]);
        dbl |= is_dbl(c.dtype[k]);
    }
    for (int i = 0; i < c.ntok; ++i)
        if (c.tok[i].kind == TOK_CONST) {
            if (c.tok[i].im != 0.0) cplx = true;
            if (c.tok[i].a == 2) dbl = true;
        }
    if (c.op != OP_NONE) {
        cplx |= is_cplx(c.dtype[0]);
        dbl |= is_dbl(c.dtype[0]);
        if (c.initop == INIT_SCALE || c.initop == INIT_CONST)
            if (c.init_im != 0.0) cplx = true;
    }
    c.ct = (cplx ? 2 : 0) + (dbl ? 1 : 0);

    // zero-size dims: map! returns early (reference mapreduce.jl:48); _mapreducedim! applies initop to a
    // non-empty output (:88-91) -- the latter is rewritten by the caller (abi) into a map; here: noop.
    for (int i = 0; i < d.ndim; ++i)
        if (d.dims[i] == 0) { noop = true; return SB_OK; }

    // drop size-1 dims
    int n = 0;
    for (int i = 0; i < d.ndim; ++i) {
        if (d.dims[i] == 1) continue;
        c.dims[n] = d.dims[i];
        for (int k = 0; k < c.nops; ++k) c.strides[k][n] = d.strides[opsel[k]][i];
        ++n;
    }
    if (n == 0) { // a single element
        c.dims[0] = 1;
        for (int k = 0; k < c.nops; ++k) c.strides[k][0] = 1;
        n = 1;
    }
    c.ndim = n;

    bool any_reduced = false;
    for (int i = 0; i < n; ++i)
        if (c.strides[0][i] == 0 && c.dims[i] > 1) any_reduced = true;
    if (c.op == OP_NONE && any_reduced) {
        // map mode with a zero output stride is "last write wins" in the reference's serial loop order
        // (mapreduce.jl:312, :320-327): undefined for a parallel engine.
        err = "map with a zero-stride (broadcast) output dim";
        return SB_E_UNSUPPORTED;
    }

    if (c.op != OP_NONE && !any_reduced) {
        // no reduced dim: out = op(initop(out), f(in...)) elementwise -> a map with the output as extra input
        if (c.nops >= MAXO) { err = "too many operands for in-place op"; return SB_E_UNSUPPORTED; }
        const int k = c.nops;
        c.base[k] = c.base[0];
        c.src[k] = 0;
        c.dtype[k] = c.dtype[0];
        c.conj[k] = c.conj[0];
        for (int i = 0; i < n; ++i) c.strides[k][i] = c.strides[0][i];
        c.nops = k + 1;
        std::vector<Tok> t;
        const Tok argout{TOK_ARG, k - 1, 0, 0};
        switch (c.initop) {
        case INIT_ZERO: t.push_back(Tok{TOK_CONST, 0, 0, 0}); break;
        case INIT_CONST: t.push_back(Tok{TOK_CONST, 0, c.init_re, c.init_im}); break;
        case INIT_SCALE:
            t.push_back(Tok{TOK_CONST, 0, c.init_re, c.init_im});
            t.push_back(argout);
            t.push_back(Tok{TOK_CALL, FN_MUL, 0, 0});
            break;
        case INIT_CONJ:
            t.push_back(argout);
            t.push_back(Tok{TOK_CALL, FN_CONJ, 0, 0});
            break;
        default: t.push_back(argout); break;
        }
        if (c.ntok == 0) t.push_back(Tok{TOK_ARG, 0, 0, 0});
        for (int i = 0; i < c.ntok; ++i) t.push_back(c.tok[i]);
        const int fn = c.op == OP_ADD ? FN_ADD : c.op == OP_MUL ? FN_MUL : c.op == OP_MIN ? FN_MIN : FN_MAX;
        t.push_back(Tok{TOK_CALL, fn, 0, 0});
        if ((int)t.size() > MAXTOK) { err = "program too long"; return SB_E_UNSUPPORTED; }
        c.ntok = (int)t.size();
        for (int i = 0; i < c.ntok; ++i) c.tok[i] = t[i];
        c.op = OP_NONE;
        c.initop = INIT_NONE;
        rc = check_program(c, err, depth);
        if (rc != SB_OK) return rc;
    }
    // The in-kernel interpreter keeps a 4-deep operand stack in registers; deeper expression trees (the reference has no
    // limit, src/broadcast.jl:67-98) are legal here too but can only run through the NVRTC-specialised kernel, which is
    // straight-line code without a stack: the plan is marked `needs_jit` and the launch fails with SB_E_UNSUPPORTED
    // (-> CPU method) only where NVRTC is unavailable.
    c.depth = depth;
    if (c.nops < 2) { // `out .= const`: give the functor a (never read) input so that a[0] exists
        c.base[1] = c.base[0];
        c.src[1] = 0;
        c.dtype[1] = c.dtype[0];
        c.conj[1] = 0;
        for (int i = 0; i < n; ++i) c.strides[1][i] = c.strides[0][i];
        c.nops = 2;
    }

    // sort: kept dims by |output stride|, then reduced dims by |first input stride| (zero strides last)
    int idx[MAXD];
    for (int i = 0; i < n; ++i) idx[i] = i;
    auto key_in = [&](int i) {
        int64_t s = iabs64(c.strides[1][i]);
        return s == 0 ? INT64_MAX : s;
    };
    std::stable_sort(idx, idx + n, [&](int a, int b) {
        const bool ra = c.strides[0][a] == 0, rb = c.strides[0][b] == 0;
        if (ra != rb) return !ra; // kept first
        if (!ra) {
            const int64_t sa = iabs64(c.strides[0][a]), sb_ = iabs64(c.strides[0][b]);
            if (sa != sb_) return sa < sb_;
        }
        return key_in(a) < key_in(b);
    });
    Canon s = c;
    for (int i = 0; i < n; ++i) {
        s.dims[i] = c.dims[idx[i]];
        for (int k = 0; k < c.nops; ++k) s.strides[k][i] = c.strides[k][idx[i]];
    }
    // fuse neighbours that are contiguous in EVERY operand (the rule of mapreduce.jl:105-110, applied after
    // sorting so that it finds every fusable pair -- SURVEY.md appendix E.5)
    int m = 0;
    for (int i = 0; i < n; ++i) {
        bool merge = m > 0;
        if (merge)
            for (int k = 0; k < s.nops; ++k)
                if (s.strides[k][i] != s.dims[m - 1] * s.strides[k][m - 1]) { merge = false; break; }
        if (merge) {
            s.dims[m - 1] *= s.dims[i];
        } else {
            s.dims[m] = s.dims[i];
            for (int k = 0; k < s.nops; ++k) s.strides[k][m] = s.strides[k][i];
            ++m;
        }
    }
    s.ndim = m;
    s.nkept = 0;
    for (int i = 0; i < m; ++i)
        if (s.strides[0][i] != 0 || s.op == OP_NONE) s.nkept++;
    if (s.op != OP_NONE && m == 1 && s.dims[0] == 1) s.nkept = 1;
    c = s;
    return SB_OK;
}

// ---- order tables ---------------------------------------------------------------------------------------
// slots: tile-dim indices (into tdim) in traversal order, fastest first
void make_order(const int *slots, int n, const int *tbits, OrderTab &o)
{
    o.n = (uint8_t)n;
    int sh = 0;
    for (int i = 0; i < n; ++i) {
        o.td[i] = (uint8_t)slots[i];
        o.shift[i] = (uint8_t)sh;
        o.bits[i] = (uint8_t)tbits[slots[i]];
        sh += tbits[slots[i]];
    }
    for (int i = n; i < MAXTD; ++i) o.td[i] = o.shift[i] = o.bits[i] = 0;
}
bool same_order(const OrderTab &a, const OrderTab &b)
{
    if (a.n != b.n) return false;
    for (int i = 0; i < a.n; ++i)
        if (a.td[i] != b.td[i]) return false;
    return true;
}

// order of the tile dims for operand strides `s` (by |stride| ascending, zero strides last)
void operand_order(const Canon &c, int k, const int *tdim, int ntd, const int *tbits, OrderTab &o)
{
    int slots[MAXTD];
    for (int i = 0; i < ntd; ++i) slots[i] = i;
    std::stable_sort(slots, slots + ntd, [&](int a, int b) {
        int64_t sa = iabs64(c.strides[k][tdim[a]]), sb_ = iabs64(c.strides[k][tdim[b]]);
        if (sa == 0) sa = INT64_MAX;
        if (sb_ == 0) sb_ = INT64_MAX;
        return sa < sb_;
    });
    make_order(slots, ntd, tbits, o);
}

} // namespace

// ---- shared-memory bank model -----------------------------------------------------------------------------
// 32 banks x 4 B.  An access of `elem_bytes` per lane is served in groups of 128/elem_bytes lanes (whole
// warp for 4 B, half-warps for 8 B, quarter-warps for 16 B); within a group the number of wavefronts is
// the largest number of DISTINCT addresses that fall into one bank.
int smem_wavefronts(const int32_t *elem_addr, int nlanes, int elem_bytes)
{
    const int words = elem_bytes / 4;
    const int group = std::max(1, 32 / words);
    int total = 0;
    for (int g0 = 0; g0 < nlanes; g0 += group) {
        int worst = 1;
        for (int bank = 0; bank < 32; bank += words) {
            int cnt = 0;
            int32_t seen[32];
            for (int l = g0; l < std::min(nlanes, g0 + group); ++l) {
                const int b = (int)(((int64_t)elem_addr[l] * words) % 32);
                if (b != bank) continue;
                bool dup = false;
                for (int q = 0; q < cnt; ++q)
                    if (seen[q] == elem_addr[l]) dup = true;
                if (!dup) seen[cnt++] = elem_addr[l];
            }
            worst = std::max(worst, cnt);
        }
        total += worst;
    }
    return total;
}

namespace {

// choose padded shared-memory strides (in elements) for a staged operand.
// own = its load order, out = output order; returns sigma per tile-dim slot and the buffer length.
int32_t choose_smem_strides(const OrderTab &own, const OrderTab &out, int ntd, int elem_bytes, int vbits, bool vec, int32_t *sigma)
{
    const int n = own.n, V = 1 << vbits;
    int pads[MAXTD] = {0};
    int bestpads[MAXTD] = {0};
    long best_cost = -1;
    int32_t best_len = 0;
    const int maxpad = 8, unit = vec ? V : 1; // vector stores need every row start on a 16-byte boundary
    const int npad = std::max(0, n - 1);
    long combos = 1;
    for (int i = 0; i < npad; ++i) combos *= (maxpad + 1);
    for (long cidx = 0; cidx < combos; ++cidx) {
        long r = cidx;
        for (int i = 0; i < npad; ++i) {
            pads[i] = (int)(r % (maxpad + 1)) * unit;
            r /= (maxpad + 1);
        }
        int32_t sg[MAXTD] = {0};
        int32_t cur = 1;
        for (int i = 0; i < n; ++i) {
            sg[own.td[i]] = cur;
            cur = cur * (1 << own.bits[i]) + (i < npad ? pads[i] : 0);
        }
        int32_t len = 1;
        for (int i = 0; i < n; ++i) len += ((1 << own.bits[i]) - 1) * sg[own.td[i]];
        int32_t wa[32], ra[32];
        for (int l = 0; l < 32; ++l) {
            const int lin = lin_t(l, vbits);
            int32_t w = 0, rr = 0;
            for (int i = 0; i < own.n; ++i) w += field_of(own, i, lin) * sg[own.td[i]];
            for (int i = 0; i < out.n; ++i) rr += field_of(out, i, lin) * sg[out.td[i]];
            wa[l] = vec ? w / V : w; // 16-byte units for vector stores
            ra[l] = rr;
        }
        const long cost = smem_wavefronts(wa, 32, vec ? 16 : elem_bytes) + smem_wavefronts(ra, 32, elem_bytes);
        if (best_cost < 0 || cost < best_cost || (cost == best_cost && len < best_len)) {
            best_cost = cost;
            best_len = len;
            for (int i = 0; i < npad; ++i) bestpads[i] = pads[i];
        }
    }
    int32_t cur = 1;
    for (int i = 0; i < ntd; ++i) sigma[i] = 0;
    for (int i = 0; i < n; ++i) {
        sigma[own.td[i]] = cur;
        cur = cur * (1 << own.bits[i]) + (i < npad ? bestpads[i] : 0);
    }
    return best_len;
}

// which (compute type, EPT) tuples exist for the pre-instantiated recipes (tools/gen_kernel_units.py)
bool ept_instantiated(int ct, int recipe, int nin_t, int ept, bool uniform)
{
    const int base = (ct == C64) ? 4 : 8;
    if (recipe == RC_INTERP || !uniform) return ept == base;
    const int words = nin_t * ept * (dtype_size(ct) / 4);
    switch (ct) {
    case F32: return (ept == 4 || ept == 8 || ept == 16) && !(ept == 16 && words > 64);
    case F64: return (ept == 4 || ept == 8 || ept == 16) && !(ept == 16 && words > 64);
    case C32: return ept == 4 || ept == 8;
    default: return ept == 4;
    }
}

int default_ept(int ct, int recipe, int nin_t, bool uniform, int64_t needed, int64_t elements, const DeviceInfo &dev)
{
    int ept = (ct == C64) ? 4 : 8;
    if (ct == F32 && needed > THREADS * 8) ept = 16; // four hot dims x 32-byte runs (config 4)
    // small problems: more, smaller tiles so that every SM gets an even share (one-wave kernels)
    while (ept > 4 && elements / ((int64_t)THREADS * ept) < 4 * (int64_t)dev.sm_count && needed <= THREADS * (ept / 2)) ept /= 2;
    if (const char *e = std::getenv("SB_FORCE_EPT")) { // tuning knob (tools/, never set in production)
        const int v = std::atoi(e);
        if (v == 4 || v == 8 || v == 16) ept = v;
    }
    while (ept > 4 && !ept_instantiated(ct, recipe, nin_t, ept, uniform)) ept /= 2;
    if (!ept_instantiated(ct, recipe, nin_t, ept, uniform)) ept = (ct == C64) ? 4 : 8;
    return ept;
}

int template_nin(int recipe, int nin)
{
    switch (recipe) {
    case RC_COPY: case RC_SCALE: case RC_ABS2: return 1;
    case RC_ADD2: case RC_ADD2_DIV: case RC_ADD2_MUL: case RC_AXPY: case RC_AXPBY: return 2;
    case RC_SUM3: return 3;
    case RC_SUM4: return 4;
    default: return nin <= 1 ? 1 : nin == 2 ? 2 : nin <= 4 ? 4 : MAXIN;
    }
}

bool recipe_instantiated(int ct, int recipe, bool uniform, bool reduce)
{
    if (recipe == RC_INTERP) return true;
    if (reduce) return uniform && (recipe == RC_COPY || recipe == RC_ABS2);
    if (recipe == RC_COPY) return true; // uniform and converting copies
    if (!uniform) return false;
    if (ct == F32 || ct == F64) return true;
    return recipe == RC_SCALE;
}

// packed edge-mask coordinates: field of tile-dim td at bit cpos[td], one guard bit above each field
uint32_t fill_cpos(int ntd, const int *tbits, uint8_t *cpos, uint8_t *cbits)
{
    uint32_t guard = 0;
    int pos = 0;
    for (int i = 0; i < ntd; ++i) {
        cpos[i] = (uint8_t)pos;
        cbits[i] = (uint8_t)tbits[i];
        pos += tbits[i];
        guard |= 1u << pos;
        pos += 1;
    }
    return guard;
}
void fill_common_tables(const OrderTab &o, int ept, int vbits, const uint8_t *cpos, uint32_t *c_tstr, uint32_t *c_joff)
{
    for (int i = 0; i < o.n; ++i) c_tstr[i] = 1u << cpos[o.td[i]];
    for (int j = 0; j < ept; ++j) {
        uint32_t c = 0;
        for (int i = 0; i < o.n; ++i) c += (uint32_t)field_of(o, i, lin_j(j, vbits)) * c_tstr[i];
        c_joff[j] = c;
    }
}

// ---- hot-dims-first tile order -----------------------------------------------------------------------------
// Tile ids count the output's dims in order (dim 0 fastest), so the CTAs of one wave write neighbouring output tiles --
// but a transposing input's contiguous dim may be the SLOWEST-counted one (reversal `permutedims!(B, A, (4,3,2,1))` of
// 91^4: the three tiles that share a 728-byte row of A run 12558 tiles apart): every 256-byte run is then fetched alone
// from its DRAM page, and the sectors that straddle two tiles are fetched twice.  This order counts the hot dims (the
// output's fastest dim, then every input's fastest dim) first, so that the tiles of one (output-fastest x input-fastest)
// plane run in the same wave.
bool build_hot_order(const MapParams &P, const bool *hot, std::vector<int32_t> &order)
{
    const int n = P.ndim;
    if (P.ntiles > (1 << 22) || P.ntiles < 2) return false;
    // hot dims with the SHORTEST tile run first: neighbouring tiles then complete the short runs into long ones within a wave
    // (rotation `permutedims!(B, A, (2,3,4,1))` of 70^4: tile [128, 16], the input's 128-byte runs are completed to 560-byte
    // rows by the 5 tiles along dim 1, which ran 2680 tiles apart in the natural order); ties keep the output's dim first
    int perm[MAXD], np = 0;
    for (int d = 0; d < n; ++d)
        if (hot[d]) perm[np++] = d;
    std::stable_sort(perm, perm + np, [&](int a, int b) { return P.tile_b[a] < P.tile_b[b]; });
    // ... as long as few tiles cover that dim (the rows are then completed within one wave: 70^4 rotation 95.4 -> 65.9 us,
    // 100^4 418 -> 269 us = 0.91); a long dim counted first loses (5001^2 transpose 69.1 -> 71.9 us): output's dim first then
    if (np > 0 && P.ntile[perm[0]] > 32) {
        np = 0;
        for (int d = 0; d < n; ++d)
            if (hot[d]) perm[np++] = d;
    }
    for (int d = 0; d < n; ++d)
        if (!hot[d]) perm[np++] = d;
    // identical to the natural order unless a multi-tile dim is overtaken by a later multi-tile dim
    bool same = true;
    {
        int last = -1;
        for (int s2 = 0; s2 < n; ++s2) {
            if (P.ntile[perm[s2]] <= 1) continue;
            if (perm[s2] < last) same = false;
            last = perm[s2];
        }
    }
    if (same) return false;
    int64_t mul[MAXD], m = 1;
    for (int d = 0; d < n; ++d) {
        mul[d] = m;
        m *= P.ntile[d];
    }
    order.resize((size_t)P.ntiles);
    int32_t cc[MAXD] = {0};
    for (int64_t pos = 0; pos < P.ntiles; ++pos) {
        int64_t id = 0;
        for (int d = 0; d < n; ++d) id += cc[d] * mul[d];
        order[(size_t)pos] = (int32_t)id;
        for (int s2 = 0; s2 < n; ++s2) {
            const int d = perm[s2];
            if (++cc[d] < P.ntile[d]) break;
            cc[d] = 0;
        }
    }
    return true;
}

// ---- alias-aware tile order ------------------------------------------------------------------------------
// Inputs that are dim-permuted views of the SAME parent (A and A' in `(A .+ A') ./ 2`; the four rotations of
// config 4) read every parent byte once per view.  The reference's answer is cache blocking inside one task;
// on the GPU the cache that can serve the second touch is L2 -- if both touches happen close in time.  This
// builds a launch order in which the tiles whose DIRECT read regions are each other's permuted images run
// next to each other (same wave): square "superblocks" closed under the permutation are visited orbit by
// orbit.  Result: DRAM reads approach the compulsory bytes (SURVEY.md section 7 "Config 2 aliasing").
bool build_tile_order(const Canon &c, const MapParams &P, std::vector<int32_t> &order)
{
    order.clear();
    const int n = c.ndim;
    if (P.ntiles < 4 || P.ntiles > (1 << 22)) return false;
    // generators: dim permutations pi with  s_k'[d] == s_k[pi(d)]  for aliasing inputs k < k'
    std::vector<std::vector<int>> gens;
    for (int k2 = 2; k2 < c.nops; ++k2)
        for (int k1 = 1; k1 < k2; ++k1) {
            if (c.base[k1] != c.base[k2] || c.base[k1] == nullptr || c.dtype[k1] != c.dtype[k2]) continue;
            std::vector<int> pi(n, -1);
            std::vector<bool> used(n, false);
            bool ok = true, ident = true;
            for (int d = 0; d < n && ok; ++d) {
                int hit = -1;
                for (int e = 0; e < n; ++e)
                    if (!used[e] && c.strides[k1][e] == c.strides[k2][d] && c.strides[k1][e] != 0 && c.dims[e] == c.dims[d]) {
                        hit = e;
                        break;
                    }
                if (hit < 0) ok = false;
                else {
                    pi[d] = hit;
                    used[hit] = true;
                    if (hit != d) ident = false;
                }
            }
            if (ok && !ident) gens.push_back(pi);
        }
    if (gens.empty()) return false;
    // superblock extents: constant along the orbits of every generator
    int64_t S[MAXD];
    for (int d = 0; d < n; ++d) S[d] = P.tile_b[d];
    for (bool changed = true; changed;) {
        changed = false;
        for (const auto &pi : gens)
            for (int d = 0; d < n; ++d) {
                const int64_t m = std::max(S[d], S[pi[d]]);
                if (S[d] != m || S[pi[d]] != m) {
                    S[d] = S[pi[d]] = m;
                    changed = true;
                }
            }
    }
    int64_t nsb[MAXD], r[MAXD], nsb_total = 1;
    for (int d = 0; d < n; ++d) {
        nsb[d] = (c.dims[d] + S[d] - 1) / S[d];
        r[d] = S[d] / P.tile_b[d];
        nsb_total *= nsb[d];
    }
    if (nsb_total > (1 << 22)) return false;
    std::vector<uint8_t> seen((size_t)nsb_total, 0);
    order.reserve((size_t)P.ntiles);
    std::vector<int64_t> orbit, stack;
    auto decode = [&](int64_t id, int64_t *cc) {
        for (int d = 0; d < n; ++d) {
            cc[d] = id % nsb[d];
            id /= nsb[d];
        }
    };
    auto encode = [&](const int64_t *cc) {
        int64_t id = 0;
        for (int d = n - 1; d >= 0; --d) id = id * nsb[d] + cc[d];
        return id;
    };
    for (int64_t sb0 = 0; sb0 < nsb_total; ++sb0) {
        if (seen[(size_t)sb0]) continue;
        orbit.clear();
        stack.assign(1, sb0);
        seen[(size_t)sb0] = 1;
        while (!stack.empty()) {
            const int64_t cur = stack.back();
            stack.pop_back();
            orbit.push_back(cur);
            int64_t cc[MAXD], nc[MAXD];
            decode(cur, cc);
            for (const auto &pi : gens)
                for (int dir = 0; dir < 2; ++dir) {
                    for (int d = 0; d < n; ++d) {
                        if (dir == 0) nc[pi[d]] = cc[d]; // Phi
                        else nc[d] = cc[pi[d]];          // Phi^-1
                    }
                    const int64_t nid = encode(nc);
                    if (!seen[(size_t)nid]) {
                        seen[(size_t)nid] = 1;
                        stack.push_back(nid);
                    }
                }
        }
        for (const int64_t sb : orbit) { // all tiles of the superblock, natural order (dim 0 fastest)
            int64_t cc[MAXD], lo[MAXD], hi[MAXD], t[MAXD];
            decode(sb, cc);
            bool empty = false;
            for (int d = 0; d < n; ++d) {
                lo[d] = cc[d] * r[d];
                hi[d] = std::min<int64_t>(lo[d] + r[d], P.ntile[d]);
                t[d] = lo[d];
                if (lo[d] >= hi[d]) empty = true;
            }
            if (empty) continue;
            for (;;) {
                int64_t id = 0;
                for (int d = n - 1; d >= 0; --d) id = id * P.ntile[d] + t[d];
                order.push_back((int32_t)id);
                int d = 0;
                for (; d < n; ++d) {
                    if (++t[d] < hi[d]) break;
                    t[d] = lo[d];
                }
                if (d == n) break;
            }
        }
    }
    if ((int64_t)order.size() != P.ntiles) { // cannot happen; keep the natural order rather than a broken table
        order.clear();
        return false;
    }
    return true;
}

// ---- TMA-staged variant ------------------------------------------------------------------------------------
bool tma_instantiated(int ct, int recipe, int nin_t, int ept)
{
    if (ept != 8) return false;
    if (ct == C32) return (recipe == RC_COPY && nin_t == 1) || (recipe == RC_INTERP && (nin_t == 1 || nin_t == 2));
    if (ct != F32 && ct != F64) return false;
    switch (recipe) {
    case RC_COPY: case RC_SCALE: return nin_t == 1;
    case RC_ADD2: case RC_ADD2_MUL: case RC_ADD2_DIV: case RC_AXPY: case RC_AXPBY: return nin_t == 2;
    case RC_INTERP: return nin_t == 1 || nin_t == 2 || nin_t == 4;
    default: return false;
    }
}

// Cheap pre-check used when the tile size is chosen: could plan_tma accept this problem with 2048-element tiles?
bool tma_candidate(const Canon &c, int recipe, int nin_t, bool uniform)
{
    if (std::getenv("SB_NO_TMA")) return false;
    const int n = c.ndim, nin = c.nops - 1, esz = dtype_size(c.ct);
    if (!uniform || nin < 1 || nin > TMA_MAXIN || n > TMA_MAXRANK || n < 2) return false;
    if (!tma_instantiated(c.ct, recipe, nin_t, 8)) return false;
    bool transposing = false;
    for (int k = 1; k <= nin; ++k) {
        int fast = -1;
        for (int d = 0; d < n; ++d) {
            if (c.strides[k][d] <= 0) return false;
            if (c.strides[k][d] == 1) fast = d;
            else if ((c.strides[k][d] * esz) % 16 != 0) return false;
        }
        if (fast < 0) return false;
        if (fast != 0) transposing = true; // canonical dim 0 is the output's fastest
    }
    return transposing && c.strides[0][0] == 1;
}

// Fills plan.tma / plan.tma_global when every input tile of the map plan P can be fetched by the TMA unit.
bool plan_tma(const Canon &c, const MapParams &P, const int *tdim, Plan &plan, const DeviceInfo &dev)
{
    if (std::getenv("SB_NO_TMA")) return false;
    const int n = c.ndim, nin = c.nops - 1, esz = dtype_size(c.ct);
    if (P.nstaged < 1 || !P.uniform || nin < 1 || nin > TMA_MAXIN || n > TMA_MAXRANK) return false;
    if (!tma_instantiated(plan.key.ct, plan.key.recipe, plan.key.nin, plan.key.ept)) return false;
    for (int d = 0; d < n; ++d)
        if (P.tile_b[d] > 256 || c.dims[d] >= ((int64_t)1 << 31)) return false;
    TmaParams T;
    std::memset(&T, 0, sizeof T);
    T.nin = nin;
    int32_t off = 0;
    const OrderTab &oo = P.order[0];
    for (int k = 1; k <= nin; ++k) {
        TmaOperand &o = T.op[k - 1];
        Plan::TmaGlobal &g = plan.tma_global[k - 1];
        g = Plan::TmaGlobal{};
        int ord[MAXD];
        for (int d = 0; d < n; ++d) {
            if (c.strides[k][d] <= 0) return false;
            ord[d] = d;
        }
        std::stable_sort(ord, ord + n, [&](int a, int b) { return c.strides[k][a] < c.strides[k][b]; });
        if (c.strides[k][ord[0]] != 1) return false;
        o.rank = n;
        g.rank = n;
        g.elem_bytes = esz;
        int64_t rows = 1;
        for (int i = 0; i < n; ++i) {
            const int d = ord[i];
            o.cdim[i] = (uint8_t)d;
            o.box[i] = P.tile_b[d];
            g.gdim[i] = (uint64_t)c.dims[d];
            g.gstride_bytes[i] = (uint64_t)c.strides[k][d] * (uint64_t)esz;
            if (i > 0) {
                if (g.gstride_bytes[i] % 16 != 0 || g.gstride_bytes[i] >= ((uint64_t)1 << 40)) return false;
                rows *= P.tile_b[d];
            }
        }
        const int b = P.tile_b[ord[0]];
        if (P.staged[k]) {
            const int ipb = 128 / esz;
            if (b < ipb || b % ipb != 0 || rows < 8) return false;
            // a transposing operand with a single 128-byte row per box line (16 Float64 along its contiguous dim: what odd
            // extents such as 70^4 end up with) is faster through the LSU kernel: 105 vs 141 us on the 70^4 reversal
            // (profiles/r02_j_odd_extents_tma_vs_lsu_shift.txt); from 256-byte rows on the TMA ring wins (54^4: 33 vs 54 us)
            // -- for 8-byte elements.  4-byte elements are the other way round: their 2048-element tiles ([64, 32]) have
            // 128-byte transposing rows by construction, and the ring beats the LSU kernel by 20-45 % once the problem
            // is not launch-bound (Float32 `3 .* A'` 4096^2: 35.8 -> 24.5 us = 0.84, 8192^2 0.63 -> 0.90; `(A .+ A') ./ 2`
            // 4000^2: 39.0 -> 29.1 us; 1000^2 loses: 3.42 -> 3.82 us; profiles/r02_y_f32_*).
            int64_t total_elems = 1;
            for (int d = 0; d < n; ++d) total_elems *= c.dims[d];
            const bool narrow_ok = std::getenv("SB_TMA_NARROW") || (esz == 4 && !std::getenv("SB_NO_TMA_NARROW") && total_elems * esz * 2 >= ((int64_t)16 << 20));
            if (b * esz < 256 && !narrow_ok) return false;
            o.swizzle = 1;
            o.nbox = b / ipb;
            o.inner_step = ipb;
            o.box[0] = ipb;
            o.box_bytes = (int32_t)(128 * rows);
        } else {
            if ((b * esz) % 16 != 0 || b * esz < 16) return false;
            // a direct operand is read in output order: its inner dim must be the output's fastest tile dim
            if (ord[0] != tdim[oo.td[0]]) return false;
            o.swizzle = 0;
            o.nbox = 1;
            o.inner_step = b;
            o.box_bytes = (int32_t)(rows * b * esz);
        }
        g.swizzle = o.swizzle;
        for (int i = 0; i < n; ++i) g.box[i] = (uint32_t)o.box[i];
        o.smem_off = off;
        off += ((o.box_bytes * o.nbox) + 1023) & ~1023;
        // consumer functional per OUTPUT-order slot
        o.inner_slot = -1;
        o.split_bits = 0;
        o.d_hi = 0;
        for (int s = 0; s < oo.n; ++s) {
            const int cd = tdim[oo.td[s]];
            int i = 0;
            while (i < n && ord[i] != cd) ++i;
            if (i == n) return false;
            if (o.swizzle) {
                if (i == 0) {
                    o.d_lo[s] = esz;
                    o.inner_slot = s;
                    o.split_bits = ilog2_ceil(o.inner_step);
                    o.d_hi = o.box_bytes;
                } else {
                    int64_t st = 128;
                    for (int q = 1; q < i; ++q) st *= o.box[q];
                    o.d_lo[s] = (int32_t)st;
                }
            } else {
                int64_t st = esz;
                for (int q = 0; q < i; ++q) st *= o.box[q];
                o.d_lo[s] = (int32_t)st;
            }
        }
        for (int j = 0; j < P.ept; ++j) {
            uint32_t d = 0;
            for (int s = 0; s < oo.n; ++s) d += tma_slot_offset(o, s, field_of(oo, s, lin_j(j, P.vbits)));
            o.s_joff[j] = (int32_t)(o.swizzle ? swizzle128(d) : d);
        }
    }
    int nboxes = 0;
    for (int k = 0; k < nin; ++k) nboxes += T.op[k].nbox;
    if (nboxes > 32) return false; // one producer lane per box
    T.stage_bytes = off;
    // 2 stages of <= 32 KB leave room for 3 resident CTAs per SM (27 warps), measured best for the 2-input case
    // (profiles/r01_v8_tile_desc_table.txt); single-input transposes (16 KB stages) get 4 stages
    int ns = (int)(65536 / std::max(1, off));
    if (off > 110 * 1024) return false;
    if (ns < 2) ns = 2;
    if (ns > 4) ns = 4;
    // one-wave problems (at most two tiles per resident CTA): deeper rings only cost shared memory, i.e. resident CTAs
    if (P.ntiles <= (int64_t)dev.sm_count * 8) ns = 2;
    if (const char *e = std::getenv("SB_TMA_STAGES")) {
        const int v = std::atoi(e);
        if (v >= 2 && v <= 8) ns = v;
    }
    T.nstage = ns;
    plan.tma = T;
    plan.tma_smem_bytes = (int64_t)ns * off + 1024;
    if (plan.tma_smem_bytes > 220 * 1024) return false;
    plan.tma_ok = true;
    return true;
}

// ---- alias-fused ("orbit") variant ---------------------------------------------------------------------------
// mirrored by csrc/kernels_orbit.cu
bool orbit_instantiated(int ct, int recipe, int nin, int ept, int logt)
{
    if (ct != F32 && ct != F64) return false;
    if (logt == 8 ? (ept != 4 && ept != 8 && ept != 16) : (logt != 9 || (ept != 2 && ept != 4 && ept != 8))) return false;
    switch (recipe) {
    case RC_ADD2: case RC_ADD2_MUL: case RC_ADD2_DIV: case RC_AXPY: case RC_AXPBY: return nin == 2;
    case RC_SUM3: return nin == 3;
    case RC_SUM4: return nin == 4;
    case RC_INTERP: return nin >= 2 && nin <= 4; // any other element function over aliased views: in-kernel interpreter
    default: return false;
    }
}

int gf2_rank(const uint32_t *v, int n)
{
    uint32_t basis[32];
    int r = 0;
    for (int i = 0; i < n; ++i) {
        uint32_t x = v[i];
        for (int q = 0; q < r; ++q)
            if ((x ^ basis[q]) < x) x ^= basis[q];
        if (x) {
            basis[r++] = x;
            for (int q = r - 1; q > 0 && basis[q] > basis[q - 1]; --q) std::swap(basis[q], basis[q - 1]);
        }
    }
    return r;
}

// ---- geometry of an alias family --------------------------------------------------------------------------------
struct OrbitGeom {
    int n = 0, nin = 0, esz = 0;
    int pord[MAXD] = {0};             // parent position -> canonical dim (input 1's strides ascending)
    int q[ORB_MAXIN + 1][MAXD] = {{0}}; // q[k][d]: parent position of canonical dim d under input k
    bool moved[MAXD] = {false};       // dim d sits at different parent positions under different views
    int nmoved = 0;
    int tb[MAXD] = {0};               // log2 tile extent per canonical dim
    int ebits = 0;                    // log2 elements per tile
};

// Is every input a dim-permuted, TMA-describable view of ONE parent, the output TMA-storable, and the output's fastest
// dim moved by at least one of the permutations (the transposing case)?  Fills pord / q / moved.
bool orbit_family(const Canon &c, OrbitGeom &G)
{
    const int n = G.n = c.ndim, nin = G.nin = c.nops - 1, esz = G.esz = dtype_size(c.ct);
    if (c.op != OP_NONE || nin < 2 || nin > ORB_MAXIN || n < 2 || n > TMA_MAXRANK) return false;
    if (c.ct != F32 && c.ct != F64) return false;
    for (int k = 0; k <= nin; ++k)
        if (c.dtype[k] != c.ct || c.conj[k]) return false;
    for (int k = 2; k <= nin; ++k)
        if (c.base[k] != c.base[1]) return false;
    if (c.base[0] == c.base[1]) return false;
    for (int d = 0; d < n; ++d) {
        G.pord[d] = d;
        if (c.strides[1][d] <= 0 || c.strides[0][d] <= 0 || c.dims[d] >= ((int64_t)1 << 31)) return false;
    }
    std::stable_sort(G.pord, G.pord + n, [&](int a, int b) { return c.strides[1][a] < c.strides[1][b]; });
    if (c.strides[1][G.pord[0]] != 1 || c.strides[0][0] != 1) return false;
    for (int i = 1; i < n; ++i) {
        const int64_t sp = c.strides[1][G.pord[i]], so = c.strides[0][i];
        if (sp == c.strides[1][G.pord[i - 1]]) return false;
        if ((sp * esz) % 16 != 0 || (so * esz) % 16 != 0) return false;
        if (sp * esz >= ((int64_t)1 << 40) || so * esz >= ((int64_t)1 << 40)) return false;
        if (so <= c.strides[0][i - 1]) return false;
    }
    for (int k = 1; k <= nin; ++k) {
        bool used[MAXD] = {false};
        for (int d = 0; d < n; ++d) {
            int hit = -1;
            for (int i = 0; i < n; ++i)
                if (!used[i] && c.strides[1][G.pord[i]] == c.strides[k][d] && c.dims[G.pord[i]] == c.dims[d]) hit = i;
            if (hit < 0) return false;
            used[hit] = true;
            G.q[k][d] = hit;
        }
    }
    G.nmoved = 0;
    for (int d = 0; d < n; ++d) {
        G.moved[d] = false;
        for (int k = 2; k <= nin; ++k)
            if (G.q[k][d] != G.q[1][d]) G.moved[d] = true;
        G.nmoved += G.moved[d];
    }
    return G.moved[0] && G.nmoved >= 2;
}

// Tile extents: 2^bb along every moved dim (a cube: the tile set must be closed under the permutations), batch bits along
// unmoved dims; 1024..4096 elements, <= 16 KB (else <= 32 KB), at least one 32-byte sector per box row.
bool orbit_tile_bits(const Canon &c, OrbitGeom &G)
{
    const int n = G.n, esz = G.esz;
    int cap[MAXD], capmin = 30;
    for (int d = 0; d < n; ++d) {
        cap[d] = ilog2_ceil(c.dims[d]);
        if (G.moved[d]) capmin = std::min(capmin, cap[d]);
    }
    auto waste_of = [&](int d, int bits) {
        const int64_t t = (int64_t)1 << bits;
        return (double)(((c.dims[d] + t - 1) / t) * t) / (double)c.dims[d];
    };
    int bbmax = std::min(12 / G.nmoved, capmin);
    for (;; --bbmax) {
        if (bbmax < 1) return false;
        double w = 1.0;
        for (int d = 0; d < n; ++d)
            if (G.moved[d]) w = std::max(w, waste_of(d, bbmax));
        if (w <= 1.2) break;
    }
    int forced_bb = 0;
    if (const char *e = std::getenv("SB_ORBIT_BITS")) forced_bb = std::atoi(e); // tuning knob: log2 cube edge
    for (int pass = 0; pass < 2; ++pass) {
        const int64_t limit = pass == 0 ? 16384 : 32768;
        for (int bb = bbmax; bb >= 1; --bb) {
            if (forced_bb > 0 && bb != forced_bb) continue;
            if (((int64_t)esz << bb) < 32) break;
            int eb = G.nmoved * bb;
            if (eb > 12 || ((int64_t)esz << eb) > limit) continue;
            int t2[MAXD] = {0};
            for (int d = 0; d < n; ++d)
                if (G.moved[d]) t2[d] = bb;
            for (int d = 0; d < n && eb < 12; ++d) { // batch bits on unmoved dims (same box position in every view)
                if (G.moved[d]) continue;
                int add = 0;
                while (eb + add < 12 && add < cap[d] && ((int64_t)esz << (eb + add + 1)) <= limit && waste_of(d, add + 1) <= 1.2) ++add;
                t2[d] = add;
                eb += add;
            }
            if (eb < 10) continue;
            for (int d = 0; d < n; ++d) G.tb[d] = t2[d];
            G.ebits = eb;
            return true;
        }
    }
    return false;
}

// Work items: the orbits of the tile grid under  c -> pb_1^-1(pb_k(c))  (pb_k: tile -> parent block of input k), in
// launch order.  False when an orbit has more than ORB_MAXG tiles or the grid is too small / too large.
bool orbit_items(const Canon &c, const OrbitGeom &G, const DeviceInfo &dev, std::vector<OrbitItem> &items, int &gmax)
{
    const int n = G.n, nin = G.nin, esz = G.esz;
    const int *tb = G.tb;
    int64_t ntile[MAXD], ntiles = 1;
    for (int d = 0; d < n; ++d) {
        ntile[d] = (c.dims[d] + ((int64_t)1 << tb[d]) - 1) >> tb[d];
        ntiles *= ntile[d];
    }
    if (ntiles < 2 || ntiles > (1 << 18)) return false; // (one 216-byte record per orbit lives on the device with the plan)
    auto decode = [&](int64_t id, int64_t *cc) {
        for (int d = 0; d < n; ++d) {
            cc[d] = id % ntile[d];
            id /= ntile[d];
        }
    };
    auto encode = [&](const int64_t *cc) {
        int64_t id = 0;
        for (int d = n - 1; d >= 0; --d) id = id * ntile[d] + cc[d];
        return id;
    };
    int inv1[MAXD]; // parent position -> canonical dim under input 1
    for (int d = 0; d < n; ++d) inv1[G.q[1][d]] = d;
    auto image = [&](int k, const int64_t *cc, int64_t *nc) { // tile whose input-1 block is input k's block of tile cc
        for (int d = 0; d < n; ++d) nc[inv1[G.q[k][d]]] = cc[d];
    };
    // Launch order of the orbit representatives: by cubic SUPER-BLOCKS of 2^sup tiles along every moved dim.  The items
    // in flight at any time then cover, in EVERY view's block family, runs of 2^sup adjacent blocks along that family's
    // contiguous dim (DRAM page locality for the transposed families, whose neighbours would otherwise be a whole
    // row of tiles apart in launch order).
    int sup = 2;
    if (const char *e = std::getenv("SB_ORBIT_SUPER")) sup = std::max(0, std::min(6, std::atoi(e)));
    std::vector<int64_t> visit;
    visit.reserve((size_t)ntiles);
    {
        int64_t nsb[MAXD], nsb_total = 1, sbx[MAXD];
        for (int d = 0; d < n; ++d) {
            sbx[d] = G.moved[d] ? ((int64_t)1 << sup) : 1;
            nsb[d] = (ntile[d] + sbx[d] - 1) / sbx[d];
            nsb_total *= nsb[d];
        }
        for (int64_t sb0 = 0; sb0 < nsb_total; ++sb0) {
            int64_t rest = sb0, lo[MAXD], hi[MAXD], t[MAXD];
            for (int d = 0; d < n; ++d) {
                lo[d] = (rest % nsb[d]) * sbx[d];
                rest /= nsb[d];
                hi[d] = std::min(lo[d] + sbx[d], ntile[d]);
                t[d] = lo[d];
            }
            for (;;) {
                visit.push_back(encode(t));
                int d = 0;
                for (; d < n; ++d) {
                    if (++t[d] < hi[d]) break;
                    t[d] = lo[d];
                }
                if (d == n) break;
            }
        }
    }
    std::vector<uint8_t> seen((size_t)ntiles, 0);
    std::vector<int64_t> orb;
    items.clear();
    gmax = 1;
    for (const int64_t t0 : visit) {
        if (seen[(size_t)t0]) continue;
        orb.assign(1, t0);
        seen[(size_t)t0] = 1;
        for (size_t h = 0; h < orb.size(); ++h) {
            int64_t cc[MAXD], nc[MAXD];
            decode(orb[h], cc);
            for (int k = 2; k <= nin; ++k) {
                image(k, cc, nc);
                for (int d = 0; d < n; ++d)
                    if (nc[d] >= ntile[d]) return false; // cannot happen: permuted dims have equal extents
                const int64_t id = encode(nc);
                if (!seen[(size_t)id]) {
                    seen[(size_t)id] = 1;
                    orb.push_back(id);
                    if ((int)orb.size() > ORB_MAXG) return false;
                }
            }
        }
        OrbitItem it;
        std::memset(&it, 0, sizeof it);
        it.ntile = (int32_t)orb.size();
        it.nblock = it.ntile;
        gmax = std::max(gmax, it.ntile);
        for (int m = 0; m < it.ntile; ++m) {
            int64_t cc[MAXD], nc[MAXD];
            decode(orb[(size_t)m], cc);
            for (int d = 0; d < n; ++d) {
                it.ocrd[m][d] = (int32_t)(cc[d] << tb[d]);
                it.ooff[m] += (cc[d] << tb[d]) * c.strides[0][d] * esz;
            }
            for (int i = 0; i < n; ++i) it.pcrd[m][i] = (int32_t)(cc[G.pord[i]] << tb[G.pord[i]]);
            it.slot[m][0] = (uint8_t)m;
            for (int k = 2; k <= nin; ++k) {
                image(k, cc, nc);
                const int64_t id = encode(nc);
                int s = -1;
                for (int z = 0; z < it.ntile; ++z)
                    if (orb[(size_t)z] == id) s = z;
                if (s < 0) return false;
                it.slot[m][k - 1] = (uint8_t)s;
            }
        }
        items.push_back(it);
    }
    // One-wave problems (fewer orbits than SMs: the README's Float64 32^4 4-way sum has 70): the output tiles of an orbit
    // are dealt out to 2 or 4 work items, each of which loads ALL the orbit's parent blocks (a few extra L2 reads) and
    // computes its share of the tiles, so that (almost) every SM gets an item instead of half of them idling.
    {
        int f = 1;
        while (f < 4 && f * 2 <= gmax && (int64_t)items.size() * f * 2 <= (int64_t)dev.sm_count) f *= 2;
        if (const char *e = std::getenv("SB_ORBIT_SPLIT")) { // tuning knob (tools/): force the split factor
            const int v = std::atoi(e);
            if ((v == 1 || v == 2 || v == 4) && v <= gmax) f = v;
        }
        if (f > 1) {
            std::vector<OrbitItem> split;
            for (const OrbitItem &it : items) {
                const int parts = std::min(f, it.ntile);
                for (int q = 0; q < parts; ++q) {
                    const int m0 = it.ntile * q / parts, m1 = it.ntile * (q + 1) / parts;
                    OrbitItem s2 = it; // same parent blocks (pcrd, nblock)
                    s2.ntile = m1 - m0;
                    for (int m = 0; m < ORB_MAXG; ++m) {
                        const int src = m0 + m < it.ntile ? m0 + m : m0;
                        std::memcpy(s2.ocrd[m], it.ocrd[src], sizeof s2.ocrd[m]);
                        std::memcpy(s2.slot[m], it.slot[src], sizeof s2.slot[m]);
                        s2.ooff[m] = it.ooff[src];
                    }
                    split.push_back(s2);
                }
            }
            items.swap(split);
        }
    }
    // Longest items first: the short orbits (tiles on a diagonal of the tile grid: fewer distinct images) go to the END of
    // the launch order, so that the CTAs that get one item more than the others in the last round get a short one
    // (config 4: 1044 items on 148 CTAs = 7 rounds + 8 items; those 8 are now 1- and 2-tile orbits).
    const int gm = gmax;
    std::stable_partition(items.begin(), items.end(), [gm](const OrbitItem &it) { return it.ntile == gm; });
    return true;
}

// Thread map x = M u over GF(2): the columns col[0..B) of M (images of the lane, warp and iteration bits) such that every
// view's shared-memory access of a warp -- the element-address bits below the bank width, ebit[v][.] -- is conflict free:
// the lane columns restricted to each view's bank bits must be linearly independent.  Randomised search with a fixed seed.
bool orbit_thread_map(const OrbitGeom &G, int (*ebit)[16], uint32_t *col)
{
    const int n = G.n, nin = G.nin, B = G.ebits;
    const int L = G.esz == 4 ? 5 : 4; // lanes served together: a warp of 4-byte or a half-warp of 8-byte accesses
    int xshift[MAXD], pshift[MAXD];
    for (int d = 0, s = 0; d < n; ++d) {
        xshift[d] = s;
        s += G.tb[d];
    }
    for (int i = 0, s = 0; i < n; ++i) {
        pshift[i] = s;
        s += G.tb[G.pord[i]];
    }
    for (int d = 0; d < n; ++d) // element-address bit of tile-coordinate bit p under view v (0 = staging / output layout)
        for (int r = 0; r < G.tb[d]; ++r) {
            const int p = xshift[d] + r;
            ebit[0][p] = p;
            for (int k = 1; k <= nin; ++k) ebit[k][p] = pshift[G.q[k][d]] + r;
        }
    auto restricted = [&](int v, uint32_t m) {
        uint32_t r = 0;
        for (int p = 0; p < B; ++p)
            if (((m >> p) & 1u) && ebit[v][p] < L) r ^= 1u << ebit[v][p];
        return r;
    };
    bool ok = false;
    uint64_t rng = 0x9E3779B97F4A7C15ull;
    for (int attempt = 0; attempt < 256 && !ok; ++attempt) {
        int nc = 0;
        bool stuck = false;
        while (nc < L && !stuck) {
            bool got = false;
            for (int tries = 0; tries < 4096 && !got; ++tries) {
                rng = rng * 6364136223846793005ull + 1442695040888963407ull;
                uint32_t cand = (uint32_t)(rng >> 33) & ((1u << B) - 1u);
                if (attempt == 0 && tries < B) cand = 1u << tries; // unit vectors first: the identity map when it already works
                if (!cand) continue;
                col[nc] = cand;
                bool good = gf2_rank(col, nc + 1) == nc + 1;
                for (int v = 0; v <= nin && good; ++v) {
                    uint32_t rv[8];
                    for (int z = 0; z <= nc; ++z) rv[z] = restricted(v, col[z]);
                    good = gf2_rank(rv, nc + 1) == nc + 1;
                }
                got = good;
            }
            if (got) ++nc;
            else stuck = true;
        }
        ok = !stuck;
    }
    if (!ok) return false;
    int ncol = L;
    for (int p = 0; p < B && ncol < B; ++p) { // complete to a basis with unit vectors
        col[ncol] = 1u << p;
        if (gf2_rank(col, ncol + 1) == ncol + 1) ++ncol;
    }
    return ncol == B;
}

// Fills plan.orbit* when every input is a dim-permuted view of one parent and the output's fastest dim is moved
// by at least one of the permutations (the transposing case).  `prog` is the matched recipe of the map plan.
bool plan_orbit(const Canon &c, const Program &prog, Plan &plan, const DeviceInfo &dev)
{
    if (std::getenv("SB_NO_ORBIT")) return false;
    // Two aliased views (A and A'): the TMA ring kernel (two loads through L2, alias-aware tile order) measures
    // 43.7 us on config 2 against 44.8-46 us here (profiles/r01_v9_orbit_tma_vs_lsu_breakdown.txt) -- both are bound by
    // the ~20 B/clk/SM the TMA unit moves -- so the fused path is taken from three views on, where it wins 3x.
    // (across the host link -- zero-copy mode of sb_mapreduce_host -- a second fetch of the parent costs a full transfer)
    if (c.nops - 1 == 2 && !dev.host_link && !std::getenv("SB_ORBIT_NIN2")) return false;
    OrbitGeom G;
    if (!orbit_family(c, G) || !orbit_tile_bits(c, G)) return false;
    const int n = G.n, nin = G.nin, esz = G.esz, ebits = G.ebits;
    const int *tb = G.tb, *pord = G.pord;
    const int32_t tile_bytes = esz << ebits;
    // consumer threads: 256.  512 (16 warps) are instantiated too but measure the same (config 4: 30.0 vs 29.7 us,
    // profiles/r01_v9_orbit_threads_lpt.txt): the tile loop is bound by shared-memory bandwidth and by the SM's memory
    // request rate (32-byte box rows), not by latency.
    int logt = 8;
    if (const char *e = std::getenv("SB_ORBIT_LOGT")) {
        const int v = std::atoi(e);
        if ((v == 8 || v == 9) && ebits - v >= 1) logt = v;
    }
    const int ept = 1 << (ebits - logt);
    if (!orbit_instantiated(c.ct, prog.recipe, nin, ept, logt)) return false;
    const int nthreads = 1 << logt;

    std::vector<OrbitItem> items;
    int gmax = 1;
    if (!orbit_items(c, G, dev, items, gmax)) return false;

    const int B = ebits, lg = esz == 4 ? 2 : 3;
    int ebit[ORB_MAXIN + 1][16];
    uint32_t col[16];
    if (!orbit_thread_map(G, ebit, col)) return false;

    OrbitParams &O = plan.orbit;
    std::memset(&O, 0, sizeof O);
    O.nin = nin;
    O.rank = n;
    O.tile_bytes = tile_bytes;
    O.gmax = gmax;
    O.stage_bytes = gmax * tile_bytes;
    O.ept = ept;
    O.log_threads = logt;
    O.nitems = (int32_t)items.size();
    O.prog = prog;
#ifdef SB_DIAG // diagnostics build only (make DIAG=1, tools/): results are WRONG with these; not compiled into the product
    if (const char *dbg = std::getenv("SB_DEBUG")) {
        if (std::strstr(dbg, "noload")) O.debug |= 1;
        if (std::strstr(dbg, "nostore")) O.debug |= 2;
        if (std::strstr(dbg, "nocompute")) O.debug |= 4;
        if (std::strstr(dbg, "nofence")) O.debug |= 8;
        if (std::strstr(dbg, "nobar")) O.debug |= 16;
    }
#endif
    auto addr_image = [&](int v, uint32_t m) {
        uint32_t a = 0;
        for (int p = 0; p < B; ++p)
            if ((m >> p) & 1u) a ^= (uint32_t)1 << (ebit[v][p] + lg);
        return a;
    };
    for (int v = 0; v <= nin; ++v) {
        for (int i = 0; i < logt; ++i) O.tcol[v][i] = addr_image(v, col[i]);
        for (int j = 0; j < ept; ++j) {
            uint32_t a = 0;
            for (int i = 0; i + logt < B; ++i)
                if ((j >> i) & 1) a ^= addr_image(v, col[logt + i]);
            O.jtab[v][j] = a;
        }
    }
    // direct-store mode: no edge tiles, and a 16-byte group stays inside the output's contiguous dim
    {
        const int lgV = esz == 4 ? 2 : 1;
        // measured (profiles/r01_v9_orbit_directstore_graph.txt): 128-bit st.global from the staging buffer is slower than
        // the TMA store for 32-byte rows (C4: 36.4 vs 28.6 us) and equal for 256-byte rows (C2) -> opt-in only
        bool direct = tb[0] >= lgV && std::getenv("SB_ORBIT_DIRECT") != nullptr;
        for (int d = 0; d < n; ++d)
            if (c.dims[d] % ((int64_t)1 << tb[d]) != 0) direct = false;
        O.direct_store = direct ? (std::atoi(std::getenv("SB_ORBIT_DIRECT")) == 2 ? 2 : 1) : 0; // 2: alternate TMA store / st.global per tile (needs 4 staging buffers)
        O.st_groups = tile_bytes / (16 * nthreads);
        int xshift[MAXD];
        for (int d = 0, sft = 0; d < n; ++d) {
            xshift[d] = sft;
            sft += tb[d];
        }
        auto xbit_off = [&](int p) -> int64_t { // global byte offset of tile-coordinate bit p (output order)
            for (int d = 0; d < n; ++d)
                if (p >= xshift[d] && p < xshift[d] + tb[d]) return ((int64_t)1 << (p - xshift[d])) * c.strides[0][d] * esz;
            return 0;
        };
        for (int i = 0; i < logt; ++i) O.st_tcol[i] = xbit_off(i + lgV);
        for (int r = 0; r < 8; ++r) {
            int64_t a = 0;
            for (int i = 0; i < 3; ++i)
                if ((r >> i) & 1) a += xbit_off(logt + lgV + i);
            O.st_roff[r] = a;
        }
        if (O.st_groups < 1 || O.st_groups > 8) O.direct_store = 0;
    }
    // guard: the bank model must agree (every warp access of every view is conflict-free)
    for (int v = 0; v <= nin; ++v)
        for (int w = 0; w < nthreads / 32; ++w)
            for (int j = 0; j < ept; ++j) {
                int32_t ea[32];
                for (int l = 0; l < 32; ++l) {
                    uint32_t a = O.jtab[v][j];
                    const int t = w * 32 + l;
                    for (int i = 0; i < logt; ++i)
                        if ((t >> i) & 1) a ^= O.tcol[v][i];
                    ea[l] = (int32_t)(a >> lg);
                }
                if (smem_wavefronts(ea, 32, esz) != (esz == 4 ? 1 : 2)) return false;
            }
    // shared memory: `ns` input stages + `ks` output staging buffers.  The TMA unit serves loads and stores in order,
    // so a store queues behind the prefetch of the next stage: with only two staging buffers the consumers stalled on
    // it every tile (profiles/r01_v9_orbit_diag.txt: loads and stores each cost 2 us alone, 10 us together).
    int ks = O.direct_store == 1 ? 2 : 4, ns = (int)((208 * 1024 - ks * (int64_t)tile_bytes) / O.stage_bytes);
    if (ns < 2) {
        ks = 2;
        ns = (int)((208 * 1024 - 2 * (int64_t)tile_bytes) / O.stage_bytes);
    }
    ns = std::max(1, std::min(4, ns));
    if (const char *e = std::getenv("SB_ORBIT_STAGES")) {
        const int v = std::atoi(e);
        if (v >= 1 && v <= 8) ns = v;
    }
    if (const char *e = std::getenv("SB_ORBIT_STAGING")) {
        const int v = std::atoi(e);
        if (v >= 2 && v <= 4 && !O.direct_store) ks = v;
    }
    O.nstage = ns;
    O.nstaging = ks;
    plan.orbit_smem_bytes = (int64_t)ns * O.stage_bytes + (int64_t)ks * tile_bytes + 128;
    if (plan.orbit_smem_bytes > 224 * 1024) return false;
    Plan::TmaGlobal &gp = plan.orbit_global[0], &go = plan.orbit_global[1];
    gp = Plan::TmaGlobal{};
    go = Plan::TmaGlobal{};
    gp.rank = go.rank = n;
    gp.elem_bytes = go.elem_bytes = esz;
    for (int i = 0; i < n; ++i) {
        gp.gdim[i] = (uint64_t)c.dims[pord[i]];
        gp.gstride_bytes[i] = (uint64_t)c.strides[1][pord[i]] * (uint64_t)esz;
        gp.box[i] = 1u << tb[pord[i]];
        go.gdim[i] = (uint64_t)c.dims[i];
        go.gstride_bytes[i] = (uint64_t)c.strides[0][i] * (uint64_t)esz;
        go.box[i] = 1u << tb[i];
    }
    for (int d = 0; d < MAXD; ++d) plan.orbit_tile_b[d] = d < n ? (1 << tb[d]) : 0;
    plan.orbit_items = std::move(items);
    plan.orbit_ok = true;
    return true;
}

int plan_map(const Canon &c, const DeviceInfo &dev, Plan &plan, std::string &err, bool balanced = false)
{
    MapParams &P = plan.map;
    std::memset(&P, 0, sizeof P);
    const int n = c.ndim, nops = c.nops, nin = nops - 1;
    const int esz = dtype_size(c.ct);
    match_recipe(c, P.prog);
    bool uniform = true;
    for (int k = 0; k < nops; ++k)
        if (c.dtype[k] != c.ct || c.conj[k]) uniform = false;
    if (!recipe_instantiated(c.ct, P.prog.recipe, uniform, false)) P.prog.recipe = RC_INTERP;

    // hot dims: the output's fastest dim and every input's fastest dim
    int fastest[MAXO];
    for (int k = 0; k < nops; ++k) {
        int best = -1;
        for (int i = 0; i < n; ++i) {
            if (c.strides[k][i] == 0) continue;
            if (best < 0 || iabs64(c.strides[k][i]) < iabs64(c.strides[k][best])) best = i;
        }
        fastest[k] = best; // -1: scalar broadcast
    }
    bool hot[MAXD] = {false};
    hot[0] = true;
    for (int k = 1; k < nops; ++k)
        if (fastest[k] >= 0) hot[fastest[k]] = true;
    int nhot = 0;
    for (int i = 0; i < n; ++i) nhot += hot[i];

    int cap[MAXD], tb[MAXD];
    for (int i = 0; i < n; ++i) {
        cap[i] = ilog2_ceil(c.dims[i]);
        tb[i] = 0;
    }
    const int minrun_bits = ilog2_ceil(std::max(1, 32 / esz)); // 32-byte sector
    int64_t needed = 1;
    for (int i = 0; i < n; ++i)
        if (hot[i]) needed <<= std::min(cap[i], minrun_bits);
    int64_t elements = 1;
    for (int i = 0; i < n; ++i) elements *= c.dims[i];
    // Tile extents are powers of two; an extent that does not divide the dim wastes masked lanes (profiles/
    // r01_v5_sweep_before_tilefix.txt: 70^4 ran at 0.18 of peak with 64-wide tiles).  lim[i] = the largest extent whose
    // padding waste ceil(n/b)*b/n stays <= 1.2 (never below a 32-byte run for hot dims).
    // Balanced mode (MapParams::umask): a box of 2^bits holds ceil(dims / ntile) used coordinates; nothing is computed
    // twice, the cost of a big box is its idle lanes -- "waste" is box / used extent, bounded by 1 / (lane utilisation).
    auto bal_ext = [&](int i, int bits) {
        const int64_t t = (int64_t)1 << bits;
        const int64_t nt = (c.dims[i] + t - 1) / t;
        return (c.dims[i] + nt - 1) / nt;
    };
    auto waste_of = [&](int i, int bits) {
        const int64_t t = (int64_t)1 << bits;
        if (balanced) return (double)t / (double)bal_ext(i, bits);
        return (double)(((c.dims[i] + t - 1) / t) * t) / (double)c.dims[i];
    };
    int lim[MAXD];
    double max_waste = 1.2; // (balanced mode: idle-lane factor of a box; larger boxes with more idle lanes measured slower)
    if (const char *e = std::getenv(balanced ? "SB_BAL_WASTE" : "SB_WASTE")) max_waste = std::max(1.0, std::atof(e)); // tuning knob
    for (int i = 0; i < n; ++i) {
        const int lo = hot[i] ? std::min(cap[i], minrun_bits) : 0;
        int bbits = cap[i];
        while (bbits > lo && waste_of(i, bbits) > max_waste) --bbits;
        lim[i] = bbits;
    }
    auto ntile_dims = [&]() {
        int q = 0;
        for (int i = 0; i < n; ++i) q += tb[i] > 0;
        return q;
    };
    auto fill = [&](int ebits_try) {
        int used_ = 0;
        for (int i = 0; i < n; ++i) tb[i] = 0;
        // phase 1: a sector-sized run along every hot dim
        for (int i = 0; i < n && used_ < ebits_try; ++i)
            if (hot[i]) {
                const int bb = std::min(std::min(cap[i], minrun_bits), ebits_try - used_);
                tb[i] = bb;
                used_ += bb;
            }
        // phase 2: grow the hot dim with the shortest run (output dim first on ties)
        while (used_ < ebits_try) {
            int pick = -1;
            for (int i = 0; i < n; ++i)
                if (hot[i] && tb[i] < lim[i] && (pick < 0 || tb[i] < tb[pick])) pick = i;
            if (pick < 0) break;
            tb[pick]++;
            used_++;
        }
        // phase 3: other dims, in output order
        for (int i = 0; i < n && used_ < ebits_try; ++i) {
            if (hot[i] || lim[i] == 0) continue;
            if (ntile_dims() >= MAXTD) break;
            const int bb = std::min(lim[i], ebits_try - used_);
            tb[i] = bb;
            used_ += bb;
        }
        return used_;
    };
    int ept = default_ept(c.ct, P.prog.recipe, template_nin(P.prog.recipe, nin), uniform, needed, elements, dev);
    // Small transposing problems (BASELINE configs 1 and 3: 1000^2 `3 .* A'`, 32^4 permutedims) keep the 2048-element
    // tile of the TMA ring kernel instead of shrinking to 1024 elements for the LSU kernel: measured 4.5 -> 3.9 us
    // (config 3) and 5.9 -> 4.1 us (config 1), profiles/r02_a_exp.txt.
    // With per-tile records (MapParams::lsu_desc) the LSU kernel beats the TMA ring on SINGLE-input transposes in two cases
    // (profiles/r02_v_tma_vs_lsu_with_records.txt): small problems (one or two waves: 32^4 permutedims 4.03 -> 3.40 us,
    // 1000^2 `3 .* A'` 3.72 -> 3.42 us) and tilings without edge tiles (64^4 48.3 -> 45.3 us, 96^4 257.8 -> 222.6 us, 128^4
    // 835 -> 697 us = 0.94 of peak).  The TMA ring keeps two-input maps (`(A .+ A') ./ 2`: 42.4 vs 46.3 us) and extents that
    // leave edge tiles, which the TMA unit clips for free (54^4: 31.9 vs 50.9 us; 3000^2: 24.7 vs 26.2 us).
    const bool lsu_pref_on = !std::getenv("SB_NO_PREFER_LSU") && !std::getenv("SB_NO_LSU_DESC") && nin == 1 && esz == 8;
    const bool small_single = lsu_pref_on && !dev.grouped && elements * esz * 2 <= ((int64_t)32 << 20);
    if (ept < 8 && !std::getenv("SB_FORCE_EPT") && !small_single && tma_candidate(c, P.prog.recipe, template_nin(P.prog.recipe, nin), uniform)) ept = 8;
    if (!std::getenv("SB_FORCE_EPT")) { // prefer the largest tile that can be filled without padding waste
        int best = ept, best_left = 1 << 30;
        for (int e = ept; e >= 4; e /= 2) {
            if (!ept_instantiated(c.ct, P.prog.recipe, template_nin(P.prog.recipe, nin), e, uniform)) continue;
            const int eb = LOG_THREADS + ilog2_ceil(e);
            const int left = eb - fill(eb);
            if (left < best_left) {
                best_left = left;
                best = e;
            }
            if (left == 0) break;
        }
        ept = best;
    }
    const int ebits = LOG_THREADS + ilog2_ceil(ept);
    int used = fill(ebits);
    // phase 3b: still room -> raise extents towards the full dim, least additional waste first
    while (used < ebits) {
        int pick = -1;
        double pw = 0;
        for (int i = 0; i < n; ++i) {
            if (tb[i] >= cap[i] || (tb[i] == 0 && ntile_dims() >= MAXTD)) continue;
            const double w = waste_of(i, tb[i] + 1) / waste_of(i, tb[i]);
            if (pick < 0 || w < pw) {
                pick = i;
                pw = w;
            }
        }
        if (pick < 0) break;
        tb[pick]++;
        used++;
    }
    // phase 4: pad (elements beyond the array are masked)
    if (used < ebits) {
        tb[0] += ebits - used;
        used = ebits;
    }
    if (const char *e = std::getenv("SB_TILE_BITS")) { // tuning knob: "b0,b1,..." log2 tile extents per canonical dim
        int v[MAXD] = {0}, cnt = 0, sum = 0;
        for (const char *q = e; *q && cnt < MAXD;) {
            v[cnt++] = std::atoi(q);
            while (*q && *q != ',') ++q;
            if (*q == ',') ++q;
        }
        for (int i = 0; i < cnt; ++i) sum += v[i];
        if (cnt == n && sum == ebits)
            for (int i = 0; i < n; ++i) tb[i] = v[i];
    }
    if (ntile_dims() > MAXTD) { err = "too many tile dims"; return SB_E_UNSUPPORTED; }

    // tile-dim slots in canonical (output) order
    int tdim[MAXTD], tbits[MAXTD], ntd = 0;
    for (int i = 0; i < n; ++i)
        if (tb[i] > 0) {
            tdim[ntd] = i;
            tbits[ntd] = tb[i];
            ++ntd;
        }
    P.ndim = n;
    P.nops = nops;
    P.ntd = ntd;
    P.ept = ept;
    P.uniform = uniform;
    P.ntiles = 1;
    // Balanced tiles (common.hpp MapParams::umask): the same number of tiles per dim, but each uses only
    // ceil(dims / ntile) of the box's 2^tb coordinates (rounded up to whole 16-byte groups), so neighbouring tiles do not
    // overlap and nothing is computed twice.
    const int Vb = (uniform && esz < 16) ? 16 / esz : 1;
    bool bal_vec_ok = true; // whole 16-byte groups are valid or not: every balanced extent is a multiple of V
    for (int i = 0; i < n; ++i) {
        P.dims[i] = c.dims[i];
        P.tile_b[i] = 1 << tb[i];
        if (balanced && tb[i] > 0) {
            const int64_t box = (int64_t)1 << tb[i];
            int64_t ext = bal_ext(i, tb[i]);
            if (ext % Vb != 0) {
                const int64_t up = (ext + Vb - 1) / Vb * Vb;
                if (up <= box && up <= c.dims[i]) ext = up;
                else bal_vec_ok = false;
            }
            if (ext < box) {
                P.tile_b[i] = (int32_t)ext;
                P.umask = 1;
            }
        }
        const int64_t nt = (c.dims[i] + P.tile_b[i] - 1) / P.tile_b[i];
        if (nt > 0x7fffffff) { err = "dim too large"; return SB_E_UNSUPPORTED; }
        P.ntile[i] = (int32_t)nt;
        P.nfull[i] = (int32_t)(c.dims[i] / P.tile_b[i]);
        P.tdiv[i] = make_fastdiv((uint32_t)nt);
        P.ntiles *= nt;
        // shifted last tile (common.hpp MapParams::excess): only where the dim holds at least one whole tile
        P.excess[i] = (P.tile_b[i] > 1 && c.dims[i] >= P.tile_b[i]) ? (int32_t)(nt * P.tile_b[i] - c.dims[i]) : 0;
        if (P.excess[i] != 0 && !std::getenv("SB_NO_SHIFT")) P.shift_last = 1;
    }
    if (P.ntiles > 0x7fffffff) { err = "too many tiles"; return SB_E_UNSUPPORTED; }
    for (int i = 0; i < ntd; ++i) P.tdim[i] = (uint8_t)tdim[i];
    for (int k = 0; k < nops; ++k) {
        P.base[k] = c.base[k];
        P.dtype[k] = (uint8_t)c.dtype[k];
        P.conj[k] = (uint8_t)c.conj[k];
        for (int i = 0; i < n; ++i) {
            P.strides[k][i] = c.strides[k][i];
            P.tstep[k][i] = (int64_t)P.tile_b[i] * c.strides[k][i] * dtype_size(c.dtype[k]);
        }
    }
    // orders
    int ident[MAXTD] = {0};
    for (int i = 0; i < ntd; ++i) ident[i] = i;
    make_order(ident, ntd, tbits, P.order[0]);
    int32_t smem_elems = 0;
    P.nstaged = 0;
    for (int k = 1; k < nops; ++k) {
        OrderTab own;
        operand_order(c, k, tdim, ntd, tbits, own);
        const bool bcast0 = ntd > 0 && c.strides[k][tdim[0]] == 0; // broadcast along the output's fastest tile dim
        const bool direct = ntd == 0 || same_order(own, P.order[0]) || own.td[0] == P.order[0].td[0] || bcast0 ||
                            fastest[k] < 0;
        if (direct) {
            P.staged[k] = 0;
            P.order[k] = P.order[0];
        } else {
            P.staged[k] = 1;
            P.order[k] = own;
            P.nstaged++;
        }
    }
    P.guard = fill_cpos(ntd, tbits, P.cpos, P.cbits);
    if (P.umask) {
        int32_t ext[MAXTD];
        for (int i = 0; i < ntd; ++i) ext[i] = P.tile_b[tdim[i]];
        P.urg = pack_rem(ext, ntd, P.cpos, P.cbits, P.guard);
    }
    // per-thread vector length: 16 bytes of the compute type when storage == compute type
    // measured (profiles/r01_v4_vector_experiment.txt): 128-bit accesses pay off for all-direct plans and for 4-byte
    // eltypes; for 8-byte staged plans the extra shared-memory conflicts of the 128-bit path cost more than they save
    const bool want_vec = (P.nstaged == 0 || esz == 4) && !std::getenv("SB_NO_VEC") && (!P.umask || bal_vec_ok);
    const int V = (uniform && esz < 16 && want_vec) ? 16 / esz : 1;
    const int vbits = (V > 1 && ept % V == 0) ? ilog2_ceil(V) : 0;
    P.vbits = vbits;
#ifdef SB_DIAG // diagnostics build only (make DIAG=1, tools/): results are WRONG with SB_DEBUG; not compiled into the product
    if (const char *dbg = std::getenv("SB_DEBUG")) {
        if (std::strstr(dbg, "nostore")) P.uniform |= 0x100;
        if (std::strstr(dbg, "noload")) P.uniform |= 0x200;
    }
    if (std::getenv("SB_TMA_NOPREFETCH")) P.uniform |= 0x400; // A/B switch: tile records loaded where they are used
#endif
    auto aligned16 = [&](int k) {
        if (((uintptr_t)c.base[k] & 15u) != 0) return false;
        for (int i = 0; i < n; ++i) {
            if ((P.tstep[k][i] % 16) != 0) return false;
            // (the shifted last tile starts `excess` elements before the regular grid position)
            if (P.shift_last && P.excess[i] != 0 && ((int64_t)P.excess[i] * c.strides[k][i] * dtype_size(c.dtype[k])) % 16 != 0) return false;
        }
        return true;
    };
    // functionals
    for (int k = 0; k < nops; ++k) {
        const OrderTab &o = P.order[k];
        for (int i = 0; i < o.n; ++i) P.g_tstr[k][i] = c.strides[k][tdim[o.td[i]]] * dtype_size(c.dtype[k]); // bytes
        for (int j = 0; j < ept; ++j) {
            int64_t g = 0;
            for (int i = 0; i < o.n; ++i) g += (int64_t)field_of(o, i, lin_j(j, vbits)) * P.g_tstr[k][i];
            P.g_joff[k][j] = g;
        }
        fill_common_tables(o, ept, vbits, P.cpos, P.c_tstr[k], P.c_joff[k]);
        // 128-bit global access: the V elements of a group are contiguous and every group start is 16-byte aligned
        bool gv = vbits > 0 && o.n > 0 && o.bits[0] >= vbits && c.strides[k][tdim[o.td[0]]] == 1 && aligned16(k);
        for (int i = 1; i < o.n && gv; ++i)
            if ((P.g_tstr[k][i] % 16) != 0) gv = false;
        P.gvec[k] = gv ? 1 : 0;
        if (k > 0 && P.staged[k]) {
            int32_t sigma[MAXTD];
            const bool sv = vbits > 0 && o.bits[0] >= vbits;
            const int32_t len = choose_smem_strides(o, P.order[0], ntd, esz, vbits, sv, sigma);
            P.svec[k] = sv ? 1 : 0;
            P.smem_off[k] = smem_elems * esz; // bytes; staged values are stored as the compute type
            smem_elems += (len + 3) & ~3;
            const OrderTab &oo = P.order[0];
            for (int i = 0; i < o.n; ++i) P.w_tstr[k][i] = sigma[o.td[i]] * esz;
            for (int i = 0; i < oo.n; ++i) P.r_tstr[k][i] = sigma[oo.td[i]] * esz;
            for (int j = 0; j < ept; ++j) {
                int32_t w = 0, r = 0;
                for (int i = 0; i < o.n; ++i) w += field_of(o, i, lin_j(j, vbits)) * P.w_tstr[k][i];
                for (int i = 0; i < oo.n; ++i) r += field_of(oo, i, lin_j(j, vbits)) * P.r_tstr[k][i];
                P.w_joff[k][j] = w;
                P.r_joff[k][j] = r;
            }
        }
    }
    plan.kind = PLAN_MAP;
    for (int k = 0; k < MAXO; ++k) plan.base_src[k] = c.src[k];
    plan.family = "map_tile";
    plan.key = KernelKey{c.ct, P.prog.recipe, template_nin(P.prog.recipe, nin), ept, uniform ? 1 : 0};
    plan.smem_bytes = (int64_t)smem_elems * esz;
    if (plan.smem_bytes > 200 * 1024) { err = "staging buffers exceed shared memory"; return SB_E_UNSUPPORTED; }
    plan.grid = std::min<int64_t>(P.ntiles, (int64_t)dev.sm_count * dev.ctas_per_sm);
    if (!P.umask && build_tile_order(c, P, plan.tile_order)) plan.note = "alias-aware tile order";
    else {
        plan.tile_order.clear();
        if (P.nstaged > 0 && !std::getenv("SB_NO_HOT_ORDER") && build_hot_order(P, hot, plan.tile_order)) plan.note = "hot-dims-first tile order";
    }
    bool exact_tiling = true;
    for (int i = 0; i < n; ++i) exact_tiling = exact_tiling && (c.dims[i] % P.tile_b[i] == 0);
    const bool prefer_lsu = small_single || (lsu_pref_on && exact_tiling);
    const bool tma = !P.umask && !prefer_lsu && plan_tma(c, P, tdim, plan, dev); // (TMA boxes and the orbit cubes are whole power-of-two boxes)
    if (plan.note == "hot-dims-first tile order") {
        // measured (profiles/r02_q_hot_order.txt): reversal permutes on the LSU kernel gain 9-13 % (70^4 105.6 -> 92.8 us,
        // 91^4 260.8 -> 233.9 us, 100^4 415.9 -> 360.8 us); the TMA ring loses 1-4 % (54^4, 64^4) and an L2-resident
        // problem (41^4, 45 MB) loses the table lookup -- so only beyond L2 size and only for the LSU kernel
        int64_t bytes = 0;
        for (int k = 0; k < nops; ++k) bytes += elements * dtype_size(c.dtype[k]);
        if (tma || bytes < ((int64_t)64 << 20)) {
            plan.tile_order.clear();
            plan.note.clear();
        }
    }
    if (tma) {
        // The TMA unit clips edge boxes for free, so pulling the last tile back only pays while the recomputed part is
        // small: 54^4 reversal, 10 of 54 columns twice: 35.3 us shifted vs 33.1 us masked; 4002^2: 45.6 vs 46.8 us.
        bool any = false;
        for (int i = 0; i < n; ++i) {
            if (P.excess[i] != 0 && (int64_t)P.excess[i] * 100 > 15 * c.dims[i]) P.excess[i] = 0;
            any = any || P.excess[i] != 0;
        }
        if (!any) P.shift_last = 0;
    }
    if (tma && P.ntiles <= (1 << 20) && !std::getenv("SB_NO_TILE_DESC")) {
        plan.tile_desc.resize((size_t)P.ntiles);
        for (int64_t pos = 0; pos < P.ntiles; ++pos) {
            uint32_t id = plan.tile_order.empty() ? (uint32_t)pos : (uint32_t)plan.tile_order[(size_t)pos];
            TileDesc &td = plan.tile_desc[(size_t)pos];
            td.id_full = id;
            td.out_off = 0;
            bool full = true;
            for (int d = 0; d < 5; ++d) td.origin[d] = 0;
            for (int d = 0; d < n; ++d) {
                const uint32_t cd = id % (uint32_t)P.ntile[d];
                id /= (uint32_t)P.ntile[d];
                const bool shifted = P.shift_last && P.excess[d] != 0 && (int32_t)cd == P.ntile[d] - 1;
                td.origin[d] = (int32_t)map_tile_origin(P, d, cd);
                td.out_off += (int64_t)cd * P.tstep[0][d] - (shifted ? (int64_t)P.excess[d] * P.strides[0][d] * dtype_size(P.dtype[0]) : 0);
                full = full && (shifted || (int32_t)cd < P.nfull[d]);
            }
            if (full) td.id_full |= 0x80000000u;
        }
    }
    // per-tile records for the LSU kernel (MapParams::lsu_desc): multi-dim plans with enough tiles to pay for the table
    int64_t lsu_min_tiles = 256;
    if (const char *e = std::getenv("SB_LSU_DESC_MIN")) lsu_min_tiles = std::max<int64_t>(1, std::atoll(e)); // (tests: exercise the table on small cases)
    if (!tma && n >= 2 && P.ntiles >= lsu_min_tiles && P.ntiles * (int64_t)(nops + 1) * 8 <= ((int64_t)8 << 20) && !std::getenv("SB_NO_LSU_DESC")) {
        const int W = nops + 1;
        P.lsu_prefetch = std::getenv("SB_LSU_NO_PREFETCH") ? 0 : 1;
        plan.lsu_desc.resize((size_t)P.ntiles * (size_t)W);
        for (int64_t pos = 0; pos < P.ntiles; ++pos) {
            uint32_t id = plan.tile_order.empty() ? (uint32_t)pos : (uint32_t)plan.tile_order[(size_t)pos];
            int64_t *r = &plan.lsu_desc[(size_t)pos * (size_t)W];
            uint32_t w = id;
            bool full = true;
            for (int k = 0; k < nops; ++k) r[1 + k] = 0;
            for (int d = 0; d < n; ++d) {
                const uint32_t cd = id % (uint32_t)P.ntile[d];
                id /= (uint32_t)P.ntile[d];
                const bool shifted = P.shift_last && P.excess[d] != 0 && (int32_t)cd == P.ntile[d] - 1;
                full = full && (shifted || (int32_t)cd < P.nfull[d]);
                for (int k = 0; k < nops; ++k)
                    r[1 + k] += (int64_t)cd * P.tstep[k][d] - (shifted ? (int64_t)P.excess[d] * P.strides[k][d] * dtype_size(P.dtype[k]) : 0);
            }
            if (full) w |= 0x80000000u;
            // edge tiles carry their packed mask: no per-tile decode in the kernel at all (a dim shorter than its box -- 54 in a
            // 64-wide tile -- makes EVERY tile an edge tile: Float32 54^4 reversal 49.1 -> 17 us)
            const uint32_t rg = full ? 0u : map_tile_rem(P, w & 0x7fffffffu);
            r[0] = (int64_t)(((uint64_t)rg << 32) | (uint64_t)w);
        }
    }
    plan.elements = 1;
    for (int i = 0; i < n; ++i) plan.elements *= c.dims[i];
    if (uniform && !P.umask && plan_orbit(c, P.prog, plan, dev)) plan.note = "alias-fused orbits";
    if (P.umask) plan.note = "balanced tiles";
    return SB_OK;
}

// mirrored by csrc/kernels_stream.cu
bool stream_instantiated(int ct, int recipe, int nin_t)
{
    if (recipe == RC_INTERP) return nin_t >= 1 && nin_t <= 3;
    if (recipe == RC_COPY) return nin_t == 1;
    if (recipe == RC_ABS2 || recipe == RC_S_ABS) return nin_t == 1 && (ct == F32 || ct == F64);
    if (recipe == RC_S_MUL2) return nin_t == 2 && (ct == F32 || ct == F64);
    return false;
}

// Streamed variant (common.hpp "StreamParams"): every output's reduction range is ONE dense run of the accumulator's
// type in every input -- complete reductions (no kept dim) and reductions over the leading dims of a dense array
// (`mapreduce(f, op, A; dims=(1,2))`, config 5 with several dense slices per GPU) with at most STREAM_MAXOUT outputs.
void plan_stream(const Canon &c, const DeviceInfo &dev, bool uniform, Plan &plan)
{
    if (std::getenv("SB_NO_STREAM")) return;
    const int nin = c.nops - 1, esz = dtype_size(c.ct);
    if (c.ndim != c.nkept + 1 || c.nkept > STREAM_MAXKD || !uniform || nin < 1 || nin > 3) return;
    const int r = c.nkept; // the (fused) reduced dim
    plan.stream_recipe = RC_INTERP;
    if (plan.key.recipe == RC_INTERP && !std::getenv("SB_NO_STREAM_FUNCTORS")) { // functors of the streamed kernel only (common.hpp RC_S_*)
        const Tok *t = c.tok;
        if (nin == 1 && c.ntok == 2 && tok_arg(t[0], 0) && tok_call(t[1], FN_ABS) && stream_instantiated(plan.key.ct, RC_S_ABS, 1)) plan.stream_recipe = RC_S_ABS;
        if (nin == 2 && c.ntok == 3 && tok_arg(t[0], 0) && tok_arg(t[1], 1) && tok_call(t[2], FN_MUL) && stream_instantiated(plan.key.ct, RC_S_MUL2, 2)) plan.stream_recipe = RC_S_MUL2;
    }
    if (plan.stream_recipe == RC_INTERP && !stream_instantiated(plan.key.ct, plan.key.recipe, plan.key.nin)) return;
    StreamParams &S = plan.stream;
    std::memset(&S, 0, sizeof S);
    // interleaved mode: ONE kept dim that is the innermost, contiguous dim of every input, the reduced dim dense on top of
    // it (column-major `mapreduce(f, op, A; dims=(2,3))`: BASELINE config 5 on one GPU)
    if (c.nkept == 1 && !std::getenv("SB_NO_STREAM_INTER")) {
        const int64_t K = c.dims[0], kb = K * esz;
        bool ok = kb >= 16 && kb <= 512 && (kb & (kb - 1)) == 0 && K <= STREAM_MAXOUT;
        for (int k = 1; k <= nin && ok; ++k) ok = c.strides[k][0] == 1 && c.strides[k][1] == K;
        if (ok) {
            S.inter_g = (int32_t)(kb / 16);
            S.nkd = 1;
            S.kdims[0] = K;
            S.kout_bytes[0] = c.strides[0][0] * dtype_size(c.dtype[0]);
            S.nout = (int32_t)K;
            S.nelem = K * c.dims[1];
            if (c.dims[1] > ((int64_t)1 << 56) / kb) return;
            S.vec_bytes = S.nelem * esz; // a multiple of K * esz >= 16
        }
    }
    if (S.inter_g == 0) {
    for (int k = 1; k <= nin; ++k)
        if (c.strides[k][r] != 1) return;
    int64_t nout = 1;
    S.nkd = c.nkept;
    for (int d = 0; d < c.nkept; ++d) {
        S.kdims[d] = c.dims[d];
        nout *= c.dims[d];
        if (nout > STREAM_MAXOUT) return;
        S.kout_bytes[d] = c.strides[0][d] * dtype_size(c.dtype[0]);
        for (int k = 1; k <= nin; ++k) {
            S.kin_bytes[k - 1][d] = c.strides[k][d] * esz;
            if (S.kin_bytes[k - 1][d] % 16 != 0) return; // every run starts on a 16-byte boundary (cp.async.bulk)
        }
    }
    S.nout = (int32_t)nout;
    S.nelem = c.dims[r];
    if (S.nelem > ((int64_t)1 << 56) / esz) return;
    S.vec_bytes = (S.nelem * esz) & ~(int64_t)15;
    } // (dense runs)
    if (S.vec_bytes < 64 * 1024) return; // (one or two CTAs of the tiled kernel do as well below that)
    S.nin = nin;
    // chunk per input: 32 KB / nin at most (four stages of 32 KB keep ~19 MB in flight over 148 SMs), smaller when the
    // problem would otherwise leave SMs without a chunk: aim for two chunks per CTA and output
    int64_t chunk = 32 * 1024 / (nin == 3 ? 4 : nin);
    const int64_t want = S.vec_bytes / (2 * (int64_t)dev.sm_count);
    while (chunk > 4096 && chunk > want) chunk >>= 1;
    if (const char *e = std::getenv("SB_STREAM_CHUNK")) { // tuning knob (tools/)
        const int64_t v = std::atoll(e);
        if (v >= 1024 && v <= 65536 && (v & (v - 1)) == 0) chunk = v;
    }
    S.chunk_bytes = (int32_t)chunk;
    S.stage_bytes = (int32_t)(chunk * nin);
    S.nchunks = (S.vec_bytes + chunk - 1) / chunk;
    int ns = (int)(128 * 1024 / S.stage_bytes);
    ns = std::max(2, std::min(8, ns));
    if (const char *e = std::getenv("SB_STREAM_STAGES")) {
        const int v = std::atoi(e);
        if (v >= 2 && v <= 8) ns = v;
    }
    while (ns > 2 && (int64_t)ns * S.stage_bytes > 192 * 1024) --ns;
    S.nstage = ns;
    plan.stream_grid = std::min<int64_t>(dev.sm_count, S.nchunks);
    plan.stream_smem_bytes = (int64_t)ns * S.stage_bytes + 128;
    plan.stream_ok = true;
}

int plan_reduce(const Canon &c, const DeviceInfo &dev, Plan &plan, std::string &err)
{
    ReduceParams &P = plan.red;
    std::memset(&P, 0, sizeof P);
    const int n = c.ndim, nops = c.nops, nin = nops - 1;
    const int esz = dtype_size(c.ct);
    match_recipe(c, P.prog);
    bool uniform = true;
    for (int k = 0; k < nops; ++k)
        if (c.dtype[k] != c.ct || c.conj[k]) uniform = false;
    if (!recipe_instantiated(c.ct, P.prog.recipe, uniform, true)) P.prog.recipe = RC_INTERP;
    if (nin > 3) { err = "reduction over more than 3 inputs"; return SB_E_UNSUPPORTED; }

    const int ept = (c.ct == C64) ? 4 : 8;
    const int ebits = LOG_THREADS + ilog2_ceil(ept);
    // dims in the first input's stride order
    int ord[MAXD];
    for (int i = 0; i < n; ++i) ord[i] = i;
    std::stable_sort(ord, ord + n, [&](int a, int b) {
        int64_t sa = iabs64(c.strides[1][a]), sb_ = iabs64(c.strides[1][b]);
        if (sa == 0) sa = INT64_MAX;
        if (sb_ == 0) sb_ = INT64_MAX;
        return sa < sb_;
    });
    int tb[MAXD] = {0};
    int used = 0, ntdims = 0;
    // A KEPT dim takes at most a 2 KB run of the tile and at most 1/8 of all outputs: `sum(A; dims=2)` of a column-major
    // 4096^2 matrix used to give all 2048 tile elements to the contiguous kept dim -- 2 output tiles, so 293 splits of the
    // reduced dim to fill the GPU, and ONE last-arriving CTA per output tile folding 293 x 2048 partials with one dependent L2
    // round trip per eight outputs: 1011 us = 0.02 of peak.  With 256 outputs x 8 reduced rows per tile (16 output tiles, 37
    // splits) and a thread-per-output fold: 29 us = 0.70; 8192^2 771 -> 124 us; 256^3 dims=2 105 -> 35 us; 128 x 512 x 512
    // dims=(2,3) 136 -> 53 us; Float32 8192^2 571 -> 74 us (profiles/r02_z_reduce_dims_*.txt: no single cap wins everywhere,
    // the cost follows the number of splits).
    int kcap_bytes = 2048;
    if (const char *e = std::getenv("SB_KEPT_CAP_BYTES")) kcap_bytes = std::max(64, std::atoi(e)); // tuning knob
    int64_t total_kept = 1;
    for (int i = 0; i < c.nkept; ++i) total_kept *= c.dims[i];
    const int kcap = std::max(3, std::min(ilog2_ceil(kcap_bytes / esz), ilog2_ceil(total_kept) - 3));
    // The mirror case: the contiguous dim is REDUCED and there are many outputs (`sum(A; dims=1)`, column sums): giving it all
    // 2048 tile elements makes one CTA per output (4096 tiny CTAs for 4096^2).  A 2 KB run leaves bits for the kept dims.
    // Column sums: 4096^2 38.4 -> 31.8 us, 8192^2 112 -> 100 us (0.82), Float32 8192^2 84.8 -> 61.7 us
    // (profiles/r02_z_reduce_dims_rcap.txt).  SB_RED_CAP_BYTES=0 switches it off.
    int rcap = ilog2_ceil(2048 / esz);
    if (const char *e = std::getenv("SB_RED_CAP_BYTES")) rcap = std::atoi(e) > 0 ? ilog2_ceil(std::max(64, std::atoi(e)) / esz) : 0;
    for (int q = 0; q < n && used < ebits && ntdims < MAXTD; ++q) {
        const int i = ord[q];
        int b = std::min(ilog2_ceil(c.dims[i]), ebits - used);
        if (i < c.nkept && !std::getenv("SB_NO_KEPT_CAP")) b = std::min(b, kcap);
        if (i >= c.nkept && q == 0 && rcap > 0 && c.nkept > 0) b = std::min(b, rcap); // (see rcap above)
        if (b == 0) continue;
        tb[i] = b;
        used += b;
        ntdims++;
    }
    for (int q = 0; q < n && used < ebits; ++q) { // bits nobody else wanted go back to the capped kept dims
        const int i = ord[q];
        if (tb[i] == 0) continue;
        const int more = std::min(ilog2_ceil(c.dims[i]) - tb[i], ebits - used);
        if (more > 0) {
            tb[i] += more;
            used += more;
        }
    }
    if (used < ebits) { // pad along the fastest dim (masked)
        if (ntdims == 0) ntdims = 1;
        tb[ord[0]] += ebits - used;
        used = ebits;
    }
    int tdim[MAXTD], tbits[MAXTD], ntd = 0;
    for (int i = 0; i < n; ++i)
        if (tb[i] > 0) {
            tdim[ntd] = i;
            tbits[ntd] = tb[i];
            ++ntd;
        }
    P.ndim = n;
    P.nops = nops;
    P.ntd = ntd;
    P.nkept = c.nkept;
    P.ept = ept;
    P.uniform = uniform;
    P.op = c.op;
    P.initop = c.initop;
    P.init_re = c.init_re;
    P.init_im = c.init_im;
    P.nouttiles = 1;
    P.nrsteps = 1;
    for (int i = 0; i < n; ++i) {
        P.dims[i] = c.dims[i];
        P.tile_b[i] = 1 << tb[i];
        const int64_t nt = (c.dims[i] + P.tile_b[i] - 1) / P.tile_b[i];
        if (nt > 0x7fffffff) { err = "dim too large"; return SB_E_UNSUPPORTED; }
        P.ntile[i] = (int32_t)nt;
        P.nfull[i] = (int32_t)(c.dims[i] / P.tile_b[i]);
        P.tdiv[i] = make_fastdiv((uint32_t)nt);
        if (i < c.nkept) P.nouttiles *= nt;
        else P.nrsteps *= nt;
    }
    if (P.nouttiles > 0x7fffffff || P.nrsteps > 0x7fffffff) { err = "too many tiles"; return SB_E_UNSUPPORTED; }
    P.outdiv = make_fastdiv((uint32_t)P.nouttiles);
    for (int i = 0; i < ntd; ++i) P.tdim[i] = (uint8_t)tdim[i];
    for (int k = 0; k < nops; ++k) {
        P.base[k] = c.base[k];
        P.dtype[k] = (uint8_t)c.dtype[k];
        P.conj[k] = (uint8_t)c.conj[k];
        for (int i = 0; i < n; ++i) {
            P.strides[k][i] = c.strides[k][i];
            P.tstep[k][i] = (int64_t)P.tile_b[i] * c.strides[k][i] * dtype_size(c.dtype[k]);
        }
    }
    operand_order(c, 1, tdim, ntd, tbits, P.order);
    for (int k = 1; k < nops; ++k) {
        for (int i = 0; i < P.order.n; ++i) P.g_tstr[k][i] = c.strides[k][tdim[P.order.td[i]]] * dtype_size(c.dtype[k]); // bytes
        for (int j = 0; j < ept; ++j) {
            int64_t g = 0;
            for (int i = 0; i < P.order.n; ++i) g += (int64_t)field_of(P.order, i, j * THREADS) * P.g_tstr[k][i];
            P.g_joff[k][j] = g;
        }
    }
    P.guard = fill_cpos(ntd, tbits, P.cpos, P.cbits);
    fill_common_tables(P.order, ept, 0, P.cpos, P.c_tstr, P.c_joff);
    // in-CTA combine layout
    int kslots[MAXTD], nk = 0;
    int32_t kdense[MAXTD] = {0}, rdense[MAXTD] = {0};
    int32_t nout = 1, nred = 1;
    for (int i = 0; i < ntd; ++i)
        if (tdim[i] < c.nkept) {
            kslots[nk++] = i;
            kdense[i] = nout;
            nout <<= tbits[i];
        }
    for (int q = 0; q < P.order.n; ++q) {
        const int i = P.order.td[q];
        if (tdim[i] >= c.nkept) {
            rdense[i] = nred;
            nred <<= tbits[i];
        }
    }
    P.nout_tile = nout;
    P.nred_tile = nred;
    P.warp_per_output = nred >= 32 ? 1 : 0;
    make_order(kslots, nk, tbits, P.kept_order);
    for (int q = 0; q < P.order.n; ++q) {
        const int i = P.order.td[q];
        if (tdim[i] < c.nkept) P.s_tstr[q] = kdense[i] * (P.warp_per_output ? nred : 1);
        else P.s_tstr[q] = rdense[i] * (P.warp_per_output ? 1 : nout);
    }
    for (int j = 0; j < ept; ++j) {
        int32_t s = 0;
        for (int q = 0; q < P.order.n; ++q) s += field_of(P.order, q, j * THREADS) * P.s_tstr[q];
        P.s_joff[j] = s;
    }
    // splits of the reduced index space
    const int64_t target = (int64_t)dev.sm_count * dev.ctas_per_sm;
    int64_t nsplit = 1;
    // at least 4 tile steps (>= 32 KB of input) per split: tiny reductions are latency-bound, more partials only
    // lengthen the final fold
    if (P.nouttiles < target) nsplit = std::min<int64_t>(std::max<int64_t>(1, P.nrsteps / 4), (target + P.nouttiles - 1) / P.nouttiles);
    if (nsplit < 1) nsplit = 1;
    P.steps_per_split = std::max<int64_t>(1, (P.nrsteps + nsplit - 1) / nsplit);
    nsplit = (P.nrsteps + P.steps_per_split - 1) / P.steps_per_split;
    if (nsplit > 0x7fffffff) { err = "too many splits"; return SB_E_UNSUPPORTED; }
    P.nsplit = (int32_t)nsplit;
    plan.kind = PLAN_REDUCE;
    for (int k = 0; k < MAXO; ++k) plan.base_src[k] = c.src[k];
    plan.family = "reduce_tile";
    plan.key = KernelKey{c.ct, P.prog.recipe, P.prog.recipe == RC_INTERP ? (nin <= 1 ? 1 : nin == 2 ? 2 : 3) : 1, ept,
                         uniform ? 1 : 0};
    plan.smem_bytes = (int64_t)THREADS * ept * esz;
    plan.grid = P.nouttiles * nsplit;
    plan.scratch_bytes = nsplit > 1 ? nsplit * P.nouttiles * (int64_t)nout * esz : 0;
    plan.finalize_threads = nsplit > 1 ? P.nouttiles * (int64_t)nout : 0;
    plan.elements = 1;
    for (int i = 0; i < n; ++i) plan.elements *= c.dims[i];
    if (plan.grid > 0x7fffffff) { err = "grid too large"; return SB_E_UNSUPPORTED; }
    plan_stream(c, dev, uniform, plan);
    return SB_OK;
}

} // namespace

// Does the byte range of the output overlap the byte range of any input?  (in-place updates such as `Y .= a .* X .+ Y`)
bool output_overlaps_inputs(const sb_desc &d)
{
    if (d.ndim < 0 || d.ndim > SB_MAX_DIMS || d.nops < 1 || d.nops > SB_MAX_OPS) return true;
    uintptr_t lo[SB_MAX_OPS], hi[SB_MAX_OPS];
    for (int k = 0; k < d.nops; ++k) {
        if (d.dtype[k] < SB_F32 || d.dtype[k] > SB_C64) return true;
        const int es = dtype_size(d.dtype[k]);
        int64_t mn = 0, mx = 0;
        for (int i = 0; i < d.ndim; ++i) {
            if (d.dims[i] <= 0) return true;
            const int64_t ext = (d.dims[i] - 1) * d.strides[k][i];
            if (ext < 0) mn += ext;
            else mx += ext;
        }
        lo[k] = (uintptr_t)d.base[k] + (uintptr_t)(mn * es);
        hi[k] = (uintptr_t)d.base[k] + (uintptr_t)((mx + 1) * es);
    }
    for (int k = 1; k < d.nops; ++k)
        if (lo[0] < hi[k] && lo[k] < hi[0]) return true;
    return false;
}

// _mapreducedim! with a zero-size dim applies initop to a non-empty output (reference mapreduce.jl:88-91)
bool empty_initop_desc(const sb_desc &D, sb_desc &E)
{
    bool anyzero = false;
    for (int i = 0; i < D.ndim; ++i)
        if (D.dims[i] == 0) anyzero = true;
    if (!anyzero || D.op == SB_OP_NONE || D.initop == SB_INIT_NONE || D.initop == SB_INIT_IDENTITY) return false;
    std::memset(&E, 0, sizeof E);
    for (int i = 0; i < D.ndim; ++i) {
        if (D.strides[0][i] == 0 && D.dims[i] != 1) continue;
        if (D.dims[i] == 0) return false;
        E.dims[E.ndim] = D.dims[i];
        E.strides[0][E.ndim] = D.strides[0][i];
        E.strides[1][E.ndim] = D.strides[0][i];
        E.ndim++;
    }
    E.nops = 2;
    E.base[0] = E.base[1] = D.base[0];
    E.dtype[0] = E.dtype[1] = D.dtype[0];
    E.conj[0] = E.conj[1] = D.conj[0];
    E.op = SB_OP_NONE;
    switch (D.initop) {
    case SB_INIT_ZERO: E.ntok = 1; E.prog[0] = sb_tok{SB_TOK_CONST, 0, 0.0, 0.0}; break;
    case SB_INIT_CONST: E.ntok = 1; E.prog[0] = sb_tok{SB_TOK_CONST, 0, D.init_re, D.init_im}; break;
    case SB_INIT_SCALE:
        E.ntok = 3;
        E.prog[0] = sb_tok{SB_TOK_CONST, 0, D.init_re, D.init_im};
        E.prog[1] = sb_tok{SB_TOK_ARG, 0, 0, 0};
        E.prog[2] = sb_tok{SB_TOK_CALL, SB_FN_MUL, 0, 0};
        break;
    default:
        E.ntok = 2;
        E.prog[0] = sb_tok{SB_TOK_ARG, 0, 0, 0};
        E.prog[1] = sb_tok{SB_TOK_CALL, SB_FN_CONJ, 0, 0};
        break;
    }
    return true;
}

int build_plan(const sb_desc &d, const DeviceInfo &dev, Plan &plan, std::string &err)
{
    sb_desc E;
    if (empty_initop_desc(d, E)) { // map!(initop, out, out) over the kept dims
        const int rc = build_plan(E, dev, plan, err);
        for (int k = 0; k < MAXO; ++k) plan.base_src[k] = 0;
        return rc;
    }
    Canon c;
    bool noop = false;
    int rc = canonicalise(d, c, noop, err);
    if (rc != SB_OK) return rc;
    if (noop) {
        plan.kind = PLAN_NOOP;
        plan.family = "noop";
        return SB_OK;
    }
    // an index space that cannot exist in 180 GB of HBM is rejected here, before any product of extents can overflow
    // (the planners below multiply tile counts in int64)
    {
        long double total = 1.0L;
        for (int i = 0; i < c.ndim; ++i) total *= (long double)c.dims[i];
        if (total > 281474976710656.0L) { // 2^48 elements
            err = "index space larger than 2^48 elements";
            return SB_E_UNSUPPORTED;
        }
    }
    rc = c.op == OP_NONE ? plan_map(c, dev, plan, err) : plan_reduce(c, dev, plan, err);
    // Odd extents on the LSU kernel: power-of-two tiles with a shifted last tile recompute ntile*2^b / dims of every such
    // dim.  Balanced tiles (MapParams::umask) remove the recomputation -- and measure the SAME or slower (reversal permute of
    // 70^4: 105.9 us with 14-of-16 balanced tiles vs 105.6 us shifted; 130.5 us with 24-of-32 tiles;
    // profiles/r02_q_odd_extents_balanced.txt): the kernel is bound by per-tile instruction issue and the latency of short
    // unaligned runs, not by L2<->SM bytes.  Default only where a dim is shorter than its box; SB_BALANCED=1 forces it.
    if (rc == SB_OK && plan.kind == PLAN_MAP && !plan.tma_ok && !plan.orbit_ok && !std::getenv("SB_NO_BALANCED")) {
        const MapParams &P = plan.map;
        double w_pow2 = 1.0; // work of the power-of-two tiling relative to the array (recomputed or masked coordinates)
        bool short_dim = false; // a dim shorter than its box: EVERY tile is an edge tile of the power-of-two tiling
        for (int i = 0; i < P.ndim; ++i)
            if (P.tile_b[i] > 1) {
                w_pow2 *= (double)((int64_t)P.ntile[i] * P.tile_b[i]) / (double)c.dims[i];
                short_dim = short_dim || c.dims[i] < P.tile_b[i];
            }
        // measured with per-tile records (profiles/r02_z_f32_odd_balanced_*.txt, r02_z_mask_records.txt): with a short dim the
        // balanced plan wins clearly (Float32 54^4 reversal 31.5 -> 17.7 us: the plan-constant mask replaces the edge path of
        // every tile); without one it is within +-7 % of the shifted tiling (41^4 +7 %, 70^4 -1 %) and stays opt-in
        if ((short_dim && w_pow2 > 1.02) || (std::getenv("SB_BALANCED") && w_pow2 > 1.10)) {
            Plan alt;
            std::string err2;
            if (plan_map(c, dev, alt, err2, true) == SB_OK && alt.map.umask) plan = std::move(alt);
        }
    }
    if (rc == SB_OK && c.depth > 4) {
        if (plan.key.recipe != RC_INTERP) { err = "internal: deep program matched a recipe"; return SB_E_INVALID; }
        plan.needs_jit = true;
        plan.orbit_ok = plan.tma_ok = plan.stream_ok = false; // (those variants run the interpreter for RC_INTERP plans)
    }
    return rc;
}

std::string describe_plan(const Plan &p)
{
    std::ostringstream os;
    os << "{\"family\":\"" << p.family << "\"";
    if (p.kind == PLAN_NOOP) {
        os << "}";
        return os.str();
    }
    static const char *ctn[] = {"f32", "f64", "c32", "c64"};
    static const char *rcn[] = {"interp", "copy", "scale", "add2", "add2_div", "add2_mul", "sum3", "sum4", "axpy", "axpby", "abs2"};
    os << ",\"ct\":\"" << ctn[p.key.ct] << "\",\"recipe\":\"" << rcn[p.key.recipe] << "\",\"nin_t\":" << p.key.nin
       << ",\"ept\":" << p.key.ept << ",\"uniform\":" << p.key.uniform << ",\"grid\":" << p.grid
       << ",\"smem_bytes\":" << p.smem_bytes << ",\"elements\":" << p.elements;
    if (p.needs_jit) os << ",\"needs_jit\":1";
    if (p.kind == PLAN_MAP && p.map.shift_last) os << ",\"shift_last\":1";
    if (p.kind == PLAN_MAP && p.map.umask) os << ",\"balanced\":1";
    if (p.kind == PLAN_MAP && !p.lsu_desc.empty()) os << ",\"lsu_desc\":1";
    auto arr64 = [&](const char *name, const int64_t *v, int n) {
        os << ",\"" << name << "\":[";
        for (int i = 0; i < n; ++i) os << (i ? "," : "") << v[i];
        os << "]";
    };
    auto arr32 = [&](const char *name, const int32_t *v, int n) {
        os << ",\"" << name << "\":[";
        for (int i = 0; i < n; ++i) os << (i ? "," : "") << v[i];
        os << "]";
    };
    if (p.kind == PLAN_MAP) {
        const MapParams &P = p.map;
        arr64("dims", P.dims, P.ndim);
        arr32("tile", P.tile_b, P.ndim);
        if (p.orbit_ok) {
            os << ",\"orbit\":{\"items\":" << p.orbit.nitems << ",\"gmax\":" << p.orbit.gmax << ",\"ept\":" << p.orbit.ept << ",\"threads\":" << (1 << p.orbit.log_threads)
               << ",\"nstage\":" << p.orbit.nstage << ",\"nstaging\":" << p.orbit.nstaging << ",\"direct_store\":" << p.orbit.direct_store << ",\"tile_bytes\":" << p.orbit.tile_bytes << ",\"smem_bytes\":" << p.orbit_smem_bytes;
            arr32("tile", p.orbit_tile_b, P.ndim);
            os << "}";
        }
        os << ",\"ntiles\":" << P.ntiles << ",\"tile_order\":" << (p.tile_order.empty() ? 0 : 1) << ",\"tma\":" << (p.tma_ok ? p.tma.nstage : 0);
        // (`grid` / `smem_bytes` above describe the generic map_tile launch; the TMA ring kernel, when it binds, runs
        //  min(ntiles, SMs x resident CTAs) CTAs of 288 threads with this much dynamic shared memory)
        if (p.tma_ok) os << ",\"tma_smem_bytes\":" << p.tma_smem_bytes << ",\"tma_threads\":288";
        os << ",\"nstaged\":" << P.nstaged << ",\"staged\":[";
        for (int k = 0; k < P.nops; ++k) os << (k ? "," : "") << (int)P.staged[k];
        os << "],\"strides\":[";
        for (int k = 0; k < P.nops; ++k) {
            os << (k ? "," : "") << "[";
            for (int i = 0; i < P.ndim; ++i) os << (i ? "," : "") << P.strides[k][i];
            os << "]";
        }
        os << "]";
    } else {
        const ReduceParams &P = p.red;
        arr64("dims", P.dims, P.ndim);
        arr32("tile", P.tile_b, P.ndim);
        os << ",\"nkept\":" << P.nkept << ",\"nouttiles\":" << P.nouttiles << ",\"nrsteps\":" << P.nrsteps
           << ",\"nsplit\":" << P.nsplit << ",\"steps_per_split\":" << P.steps_per_split
           << ",\"nout_tile\":" << P.nout_tile << ",\"nred_tile\":" << P.nred_tile
           << ",\"warp_per_output\":" << P.warp_per_output << ",\"scratch_bytes\":" << p.scratch_bytes;
        if (p.stream_ok)
            os << ",\"stream\":{\"functor\":" << p.stream_recipe << ",\"grid\":" << p.stream_grid << ",\"nout\":" << p.stream.nout << ",\"interleaved\":" << p.stream.inter_g << ",\"chunk_bytes\":" << p.stream.chunk_bytes << ",\"nstage\":" << p.stream.nstage
               << ",\"nchunks\":" << p.stream.nchunks << ",\"smem_bytes\":" << p.stream_smem_bytes << "}";
    }
    os << "}";
    return os.str();
}

} // namespace sb
