// emul.cpp -- TEST INFRASTRUCTURE ONLY: CPU thread-grid emulation of the CUDA tile kernels.
//
// There is no GPU in the build container, so the index machinery (planner tables, per-thread offset
// functionals, staging-buffer slots, edge masks, split reductions) is verified here by running the SAME
// host/device-neutral kernel bodies (csrc/map_tile.hpp, csrc/reduce_tile.hpp) for every (block, thread) of
// the launch on the CPU.  This library is built and loaded only by tests/ (CPU suite); the product library
// contains no CPU execution path and never links this file.
#include "../../strided.jl_b200/csrc/map_tile.hpp"
#include "../../strided.jl_b200/csrc/reduce_tile.hpp"
#include "../../strided.jl_b200/csrc/planner.hpp"
#include "../../strided.jl_b200/csrc/tma_tile.hpp"
#include "../../strided.jl_b200/csrc/orbit_tile.hpp"
#include "../../strided.jl_b200/csrc/reduce_stream.hpp"
#include <cstdlib>

#include <cstring>
#include <string>
#include <vector>

using namespace sb;

static std::string g_err;

template <class CT, int RC, int NIN, int EPT, bool U> static void run_map(const Plan &plan, int grid)
{
    const MapParams &P = plan.map;
    std::vector<unsigned char> smem((size_t)plan.smem_bytes + 64);
    std::vector<MapThread<NIN + 1>> th(THREADS);
    for (int t = 0; t < THREADS; ++t) map_thread_init<NIN + 1>(P, t, th[t]);
    using Regs = CT[NIN][EPT];
    std::vector<char> regbuf(sizeof(Regs) * THREADS);
    Regs *v = reinterpret_cast<Regs *>(regbuf.data());
    for (int b = 0; b < grid; ++b)
        for (uint32_t pos = (uint32_t)b; pos < (uint32_t)P.ntiles; pos += (uint32_t)grid) {
            std::vector<MapTile<NIN + 1>> tl(THREADS); // the tile descriptor is per thread (it folds the thread offset in)
            for (int t = 0; t < THREADS; ++t) {
                map_tile_init<NIN + 1>(P, th[t], pos, tl[t]);
                map_phase1<CT, NIN, EPT, U>(P, th[t], tl[t], t, v[t], smem.data());
            }
            for (int t = 0; t < THREADS; ++t) map_phase2<CT, RC, NIN, EPT, U>(P, th[t], tl[t], t, v[t], smem.data());
        }
}

// TMA-staged variant: the box copies the Tensor Memory Accelerator would perform (dense box, innermost dim first,
// out-of-bounds elements zero-filled, 128-byte swizzle on the shared-memory address) are emulated, then the
// real consumer body runs for every thread.
template <class CT, int RC, int NIN, int EPT> static void run_tma(const Plan &plan, int grid)
{
    const MapParams &P = plan.map;
    const TmaParams &T = plan.tma;
    std::vector<unsigned char> raw((size_t)T.stage_bytes + 2048);
    unsigned char *stage = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(raw.data()) + 1023) & ~(uintptr_t)1023);
    std::vector<MapThread<1>> th0(THREADS);
    std::vector<TmaThread<NIN>> th(THREADS);
    for (int t = 0; t < THREADS; ++t) {
        map_thread_init<1>(P, t, th0[t]);
        tma_thread_init<NIN>(P, T, t, th[t]);
    }
    for (int b = 0; b < grid; ++b)
        for (uint32_t pos = (uint32_t)b; pos < (uint32_t)P.ntiles; pos += (uint32_t)grid) {
            const uint32_t id = P.tile_order ? (uint32_t)P.tile_order[pos] : pos;
            std::memset(stage, 0xCD, (size_t)T.stage_bytes);
            for (int k = 0; k < T.nin; ++k) {
                const TmaOperand &o = T.op[k];
                const Plan::TmaGlobal &g = plan.tma_global[k];
                const unsigned char *base = P.base[k + 1];
                for (int q = 0; q < o.nbox; ++q) {
                    int32_t crd[TMA_MAXRANK];
                    tma_box_coords(P, o, id, q, crd);
                    int64_t nelem = 1;
                    for (int i = 0; i < o.rank; ++i) nelem *= o.box[i];
                    for (int64_t e = 0; e < nelem; ++e) {
                        int64_t rest = e, src = 0;
                        bool oob = false;
                        for (int i = 0; i < o.rank; ++i) {
                            const int64_t ix = rest % o.box[i];
                            rest /= o.box[i];
                            const int64_t gi = (int64_t)crd[i] + ix;
                            if (gi < 0 || gi >= (int64_t)g.gdim[i]) oob = true;
                            src += gi * (i == 0 ? (int64_t)g.elem_bytes : (int64_t)g.gstride_bytes[i]);
                        }
                        uint32_t dense = (uint32_t)(e * g.elem_bytes);
                        uint32_t off = (uint32_t)o.smem_off + (uint32_t)(q * o.box_bytes) + dense;
                        if (o.swizzle) off = swizzle128(off);
                        if (oob) std::memset(stage + off, 0, (size_t)g.elem_bytes);
                        else std::memcpy(stage + off, base + src, (size_t)g.elem_bytes);
                    }
                }
            }
            for (int t = 0; t < THREADS; ++t) {
                MapTile<1> tl;
                map_tile_init<1>(P, th0[t], pos, tl);
                tma_consume<CT, RC, NIN, EPT>(P, T, th[t], th0[t], tl, t, stage);
            }
        }
}

// Alias-fused orbit variant: the TMA box loads of the parent blocks (dense box in parent dim order, zero-filled out of
// bounds) and the TMA store of the staged output tile (clipped) are emulated; the real consumer body runs per thread.
static void emul_box(const Plan::TmaGlobal &g, const int32_t *crd, unsigned char *gbase, unsigned char *sm, bool store)
{
    int64_t nelem = 1;
    for (int i = 0; i < g.rank; ++i) nelem *= g.box[i];
    for (int64_t e = 0; e < nelem; ++e) {
        int64_t rest = e, off = 0;
        bool oob = false;
        for (int i = 0; i < g.rank; ++i) {
            const int64_t ix = rest % g.box[i];
            rest /= g.box[i];
            const int64_t gi = (int64_t)crd[i] + ix;
            if (gi < 0 || gi >= (int64_t)g.gdim[i]) oob = true;
            off += gi * (i == 0 ? (int64_t)g.elem_bytes : (int64_t)g.gstride_bytes[i]);
        }
        unsigned char *s = sm + e * g.elem_bytes;
        if (store) {
            if (!oob) std::memcpy(gbase + off, s, (size_t)g.elem_bytes);
        } else {
            if (oob) std::memset(s, 0, (size_t)g.elem_bytes);
            else std::memcpy(s, gbase + off, (size_t)g.elem_bytes);
        }
    }
}

template <class CT, int RC, int NIN, int EPT> static void run_orbit(const Plan &plan, int grid)
{
    const OrbitParams &O = plan.orbit;
    std::vector<unsigned char> ring((size_t)O.nstage * O.stage_bytes + (size_t)O.nstaging * O.tile_bytes);
    const int NT = 1 << O.log_threads;
    std::vector<OrbitThread<NIN>> th(NT);
    for (int t = 0; t < NT; ++t) orbit_thread_init<NIN>(O, t, th[t]);
    const uint32_t staging0 = (uint32_t)(O.nstage * O.stage_bytes);
    for (int b = 0; b < grid; ++b) {
        int stage = 0;
        uint32_t nout = 0;
        for (uint32_t pos = (uint32_t)b; pos < (uint32_t)O.nitems; pos += (uint32_t)grid) {
            const OrbitItem &it = plan.orbit_items[pos];
            std::memset(ring.data() + (size_t)stage * O.stage_bytes, 0xCD, (size_t)O.stage_bytes);
            for (int s = 0; s < it.nblock; ++s)
                emul_box(plan.orbit_global[0], it.pcrd[s], plan.map.base[1], ring.data() + (size_t)stage * O.stage_bytes + (size_t)s * O.tile_bytes, false);
            for (int m = 0; m < it.ntile; ++m) {
                uint32_t slots;
                std::memcpy(&slots, it.slot[m], 4);
                const uint32_t sbuf_off = staging0 + (nout % (uint32_t)O.nstaging) * (uint32_t)O.tile_bytes;
                for (int t = 0; t < NT; ++t)
                    orbit_compute<CT, RC, NIN, EPT>(O, th[t], ring.data(), (uint32_t)(stage * O.stage_bytes), slots, sbuf_off);
                if (O.direct_store) {
                    for (int t = 0; t < NT; ++t)
                        orbit_store_direct(O, t, orbit_store_toff(O, t), ring.data(), sbuf_off, plan.map.base[0] + it.ooff[m]);
                } else {
                    emul_box(plan.orbit_global[1], it.ocrd[m], plan.map.base[0], ring.data() + sbuf_off, true);
                }
                ++nout;
            }
            if (++stage == O.nstage) stage = 0;
        }
    }
}

template <class CT> static bool orbit_dispatch(const Plan &plan, int grid)
{
    const int rc = plan.key.recipe, nin = plan.orbit.nin, ept = plan.orbit.ept;
#define TRYO(R, N)                                                                                                   \
    if (rc == R && nin == N) {                                                                                       \
        if (ept == 2) { run_orbit<CT, R, N, 2>(plan, grid); return true; }                                           \
        if (ept == 4) { run_orbit<CT, R, N, 4>(plan, grid); return true; }                                           \
        if (ept == 8) { run_orbit<CT, R, N, 8>(plan, grid); return true; }                                           \
        if (ept == 16) { run_orbit<CT, R, N, 16>(plan, grid); return true; }                                         \
    }
    TRYO(RC_ADD2, 2) TRYO(RC_ADD2_MUL, 2) TRYO(RC_ADD2_DIV, 2) TRYO(RC_AXPY, 2) TRYO(RC_AXPBY, 2) TRYO(RC_SUM3, 3) TRYO(RC_SUM4, 4) TRYO(RC_INTERP, 2) TRYO(RC_INTERP, 3) TRYO(RC_INTERP, 4)
#undef TRYO
    return false;
}

template <class CT> static bool tma_dispatch(const Plan &plan, int grid)
{
    const KernelKey &k = plan.key;
    if (k.ept != 8) return false;
#define TRYT(R, N)                                                                                                   \
    if (k.recipe == R && k.nin == N) {                                                                               \
        run_tma<CT, R, N, 8>(plan, grid);                                                                            \
        return true;                                                                                                 \
    }
    TRYT(RC_COPY, 1) TRYT(RC_SCALE, 1) TRYT(RC_ADD2, 2) TRYT(RC_ADD2_MUL, 2) TRYT(RC_ADD2_DIV, 2) TRYT(RC_AXPY, 2) TRYT(RC_AXPBY, 2)
    TRYT(RC_INTERP, 1) TRYT(RC_INTERP, 2) TRYT(RC_INTERP, 4)
#undef TRYT
    return false;
}

template <class AT, int RC, int NIN, int EPT, bool U> static void run_reduce(const Plan &plan)
{
    const ReduceParams &P = plan.red;
    std::vector<AT> smem((size_t)THREADS * EPT);
    for (uint32_t bid = 0; bid < (uint32_t)plan.grid; ++bid) {
        for (int t = 0; t < THREADS; ++t) red_accumulate<AT, RC, NIN, EPT, U>(P, bid, t, smem.data());
        if (P.warp_per_output) {
            for (int warp = 0; warp < THREADS / 32; ++warp)
                for (int o = warp; o < P.nout_tile; o += THREADS / 32) {
                    AT p[32];
                    for (int lane = 0; lane < 32; ++lane) p[lane] = red_lane_partial<AT>(P, smem.data(), o, lane);
                    for (int m = 16; m >= 1; m >>= 1) { // butterfly, as __shfl_xor_sync
                        AT q[32];
                        for (int lane = 0; lane < 32; ++lane) q[lane] = red_apply<AT>(P.op, p[lane], p[lane ^ m]);
                        std::memcpy(p, q, sizeof p);
                    }
                    red_finish<AT, U>(P, bid, o, p[0]);
                }
        } else {
            for (int t = 0; t < THREADS; ++t)
                for (int o = t; o < P.nout_tile; o += THREADS) red_finish<AT, U>(P, bid, o, red_thread_partial<AT>(P, smem.data(), o));
        }
    }
    // the fused finalize: emulated after all CTAs (on the GPU the last-arriving CTA of each output tile does it)
    if (plan.finalize_threads > 0) {
        for (int64_t out_idx = 0; out_idx < plan.finalize_threads; ++out_idx) {
            AT p[32];
            for (int lane = 0; lane < 32; ++lane) p[lane] = red_finalize_lane<AT>(P, out_idx, lane);
            for (int m = 16; m >= 1; m >>= 1) {
                AT q[32];
                for (int lane = 0; lane < 32; ++lane) q[lane] = red_apply<AT>(P.op, p[lane], p[lane ^ m]);
                std::memcpy(p, q, sizeof p);
            }
            red_finalize_store<AT, U>(P, out_idx, p[0]);
        }
    }
}

// Streamed complete reduction: the cp.async.bulk chunk copies are emulated (plain copies of the chunk of every input into
// the stage), the real consumer body runs per thread; warp butterfly, CTA fold, partials and the last-CTA fold follow
// the kernel's order (csrc/stream_kernel.cuh).
template <class AT, class PT> static AT butterfly(const PT &P, AT *p)
{
    for (int m = 16; m >= 1; m >>= 1) {
        AT q[32];
        for (int lane = 0; lane < 32; ++lane) q[lane] = red_apply<AT>(P.op, p[lane], p[lane ^ m]);
        std::memcpy(p, q, sizeof q);
    }
    return p[0];
}
template <class AT, int RC, int NIN> static void run_stream(const Plan &plan)
{
    StreamArgs P; // the lean parameter block the launch builds (csrc/abi.cu)
    std::memset(&P, 0, sizeof P);
    for (int q = 0; q <= plan.stream.nin; ++q) {
        P.base[q] = plan.red.base[q];
        P.dtype[q] = plan.red.dtype[q];
        P.conj[q] = plan.red.conj[q];
    }
    P.op = plan.red.op;
    P.initop = plan.red.initop;
    P.init_re = plan.red.init_re;
    P.init_im = plan.red.init_im;
    P.S = plan.stream;
    P.prog = plan.red.prog;
    const StreamParams &S = P.S;
    const int grid = (int)plan.stream_grid;
    std::vector<unsigned char> raw((size_t)S.stage_bytes + 64);
    unsigned char *stage = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(raw.data()) + 15) & ~(uintptr_t)15);
    std::vector<AT> partials((size_t)grid * (size_t)S.nout);
    constexpr int V = StreamVec<AT>::V;
    using Acc = AT[STREAM_ACC][V];
    const int nruns = S.inter_g > 0 ? 1 : S.nout;
    for (int b = 0; b < grid; ++b) {
        for (int o = 0; o < nruns; ++o) {
            std::vector<char> accbuf(sizeof(Acc) * THREADS);
            Acc *acc = reinterpret_cast<Acc *>(accbuf.data());
            for (int t = 0; t < THREADS; ++t)
                for (int q = 0; q < STREAM_ACC; ++q)
                    for (int u = 0; u < V; ++u) acc[t][q][u] = red_neutral<AT>(P.op);
            for (int64_t c = b; c < S.nchunks; c += grid) {
                const int64_t off = c * (int64_t)S.chunk_bytes;
                const int64_t left = S.vec_bytes - off;
                const int64_t nb = left < S.chunk_bytes ? left : S.chunk_bytes;
                std::memset(stage, 0xCD, (size_t)S.stage_bytes);
                for (int k = 0; k < S.nin; ++k)
                    std::memcpy(stage + (size_t)k * S.chunk_bytes, P.base[k + 1] + stream_out_offset(S, o, k) + off, (size_t)nb);
                for (int t = 0; t < THREADS; ++t) stream_chunk<AT, RC, NIN>(P, S, stage, (int)(nb >> 4), t, acc[t]);
            }
            if (S.inter_g > 0) { // interleaved outputs: butterfly over lanes of equal class, [warp][output] rows, fold over warps
                const int G = S.inter_g;
                std::vector<AT> fold((size_t)(THREADS / 32) * STREAM_MAXOUT);
                for (int w = 0; w < THREADS / 32; ++w)
                    for (int u = 0; u < V; ++u) {
                        AT p[32];
                        for (int lane = 0; lane < 32; ++lane) p[lane] = stream_thread_lane_total<AT>(P, acc[w * 32 + lane], u);
                        for (int m = 16; m >= G; m >>= 1) {
                            AT q2[32];
                            for (int lane = 0; lane < 32; ++lane) q2[lane] = red_apply<AT>(P.op, p[lane], p[lane ^ m]);
                            std::memcpy(p, q2, sizeof q2);
                        }
                        for (int lane = 0; lane < G; ++lane) fold[(size_t)w * STREAM_MAXOUT + lane * V + u] = p[lane];
                    }
                for (int oo = 0; oo < S.nout; ++oo) {
                    AT q = fold[(size_t)oo];
                    for (int w = 1; w < THREADS / 32; ++w) q = red_apply<AT>(P.op, q, fold[(size_t)w * STREAM_MAXOUT + oo]);
                    partials[(size_t)oo * grid + b] = q;
                }
                break;
            }
            AT wres[THREADS / 32];
            for (int w = 0; w < THREADS / 32; ++w) {
                AT p[32];
                for (int lane = 0; lane < 32; ++lane) {
                    const int t = w * 32 + lane;
                    p[lane] = stream_thread_total<AT>(P, acc[t]);
                    if (b == 0 && t == 0) p[lane] = stream_rest<AT, RC, NIN>(P, S, o, p[lane]);
                }
                wres[w] = butterfly<AT, StreamArgs>(P, p);
            }
            AT q = wres[0];
            for (int w = 1; w < THREADS / 32; ++w) q = red_apply<AT>(P.op, q, wres[w]);
            partials[(size_t)o * grid + b] = q;
        }
    }
    for (int o = 0; o < S.nout; ++o) { // the last-arriving CTA: one warp per output, lane l folds CTAs l, l + 32, ...
        AT p[32];
        for (int lane = 0; lane < 32; ++lane) {
            AT r = red_neutral<AT>(P.op);
            for (int i = lane; i < grid; i += 32) r = red_apply<AT>(P.op, r, partials[(size_t)o * grid + i]);
            p[lane] = r;
        }
        stream_store<AT>(P, S, o, butterfly<AT, StreamArgs>(P, p));
    }
}
template <class AT> static bool stream_dispatch(const Plan &plan)
{
    const KernelKey &k = plan.key;
    if (k.recipe == RC_COPY && k.nin == 1) { run_stream<AT, RC_COPY, 1>(plan); return true; }
    if constexpr (!traits<AT>::cplx) {
        if (k.recipe == RC_ABS2 && k.nin == 1) { run_stream<AT, RC_ABS2, 1>(plan); return true; }
    }
    if (k.recipe == RC_INTERP && k.nin == 1) { run_stream<AT, RC_INTERP, 1>(plan); return true; }
    if (k.recipe == RC_INTERP && k.nin == 2) { run_stream<AT, RC_INTERP, 2>(plan); return true; }
    if (k.recipe == RC_INTERP && k.nin == 3) { run_stream<AT, RC_INTERP, 3>(plan); return true; }
    return false;
}

// dispatch over the instantiated tuples (mirror of csrc/kernels_*.cu)
template <class CT, bool U> static bool map_dispatch_interp(const Plan &plan, int grid)
{
    const int nin = plan.key.nin, ept = plan.key.ept;
#define TRY(N, E)                                                                                                    \
    if (nin == N && ept == E) {                                                                                      \
        run_map<CT, RC_INTERP, N, E, U>(plan, grid);                                                                 \
        return true;                                                                                                 \
    }
    TRY(1, 4) TRY(2, 4) TRY(4, 4) TRY(7, 4) TRY(1, 8) TRY(2, 8) TRY(4, 8) TRY(7, 8) TRY(1, 16) TRY(2, 16) TRY(4, 16) TRY(7, 16)
#undef TRY
    return false;
}

template <class CT> static bool map_dispatch(const Plan &plan, int grid)
{
    const KernelKey &k = plan.key;
    if (k.recipe == RC_INTERP) return k.uniform ? map_dispatch_interp<CT, true>(plan, grid) : map_dispatch_interp<CT, false>(plan, grid);
#define TRYR(R, N)                                                                                                   \
    if (k.recipe == R && k.nin == N && k.uniform) {                                                                  \
        if (k.ept == 4) { run_map<CT, R, N, 4, true>(plan, grid); return true; }                                     \
        if (k.ept == 8) { run_map<CT, R, N, 8, true>(plan, grid); return true; }                                     \
        if (k.ept == 16) { run_map<CT, R, N, 16, true>(plan, grid); return true; }                                   \
    }
    TRYR(RC_COPY, 1) TRYR(RC_SCALE, 1) TRYR(RC_ABS2, 1) TRYR(RC_ADD2, 2) TRYR(RC_ADD2_DIV, 2) TRYR(RC_ADD2_MUL, 2)
    TRYR(RC_AXPY, 2) TRYR(RC_AXPBY, 2) TRYR(RC_SUM3, 3) TRYR(RC_SUM4, 4)
#undef TRYR
    if (k.recipe == RC_COPY && k.nin == 1 && !k.uniform) {
        if (k.ept == 4) { run_map<CT, RC_COPY, 1, 4, false>(plan, grid); return true; }
        if (k.ept == 8) { run_map<CT, RC_COPY, 1, 8, false>(plan, grid); return true; }
        if (k.ept == 16) { run_map<CT, RC_COPY, 1, 16, false>(plan, grid); return true; }
    }
    return false;
}

template <class AT, int EPT> static bool red_dispatch(const Plan &plan)
{
    const KernelKey &k = plan.key;
    if (k.ept != EPT) return false;
    if (k.recipe == RC_COPY && k.uniform) { run_reduce<AT, RC_COPY, 1, EPT, true>(plan); return true; }
    if (k.recipe == RC_ABS2 && k.uniform) { run_reduce<AT, RC_ABS2, 1, EPT, true>(plan); return true; }
    if (k.recipe == RC_INTERP) {
#define TRY(N)                                                                                                       \
    if (k.nin == N) {                                                                                                \
        if (k.uniform) run_reduce<AT, RC_INTERP, N, EPT, true>(plan);                                                \
        else run_reduce<AT, RC_INTERP, N, EPT, false>(plan);                                                         \
        return true;                                                                                                 \
    }
        TRY(1) TRY(2) TRY(3)
#undef TRY
    }
    return false;
}

extern "C" const char *emul_last_error(void) { return g_err.c_str(); }

// Runs `desc` (HOST pointers) exactly as the CUDA launch would be organised.  `grid_limit` > 0 overrides the
// persistent-grid size of the map kernel (to exercise the tile loop with few CTAs).
extern "C" int emul_mapreduce(const sb_desc *desc, int grid_limit)
{
    Plan plan;
    DeviceInfo dev;
    int rc = build_plan(*desc, dev, plan, g_err);
    if (rc != SB_OK) return rc;
    if (plan.kind == PLAN_NOOP) return SB_OK;
    if (plan.needs_jit) { // (the emulator runs the interpreter bodies; deep programs exist only as NVRTC kernels)
        g_err = "emul: program needs the NVRTC path";
        return SB_E_UNSUPPORTED;
    }
    if (!plan.tile_order.empty()) plan.map.tile_order = plan.tile_order.data();
    if (!plan.lsu_desc.empty()) plan.map.lsu_desc = plan.lsu_desc.data();
    if (plan.kind == PLAN_MAP && plan.map.shift_last && output_overlaps_inputs(*desc)) { // as the launch does (csrc/abi.cu)
        plan.map.shift_last = 0;
        plan.tma_ok = false;
        plan.map.lsu_desc = nullptr;
    }
    bool ok = false;
    if (plan.kind == PLAN_MAP) {
        int grid = (int)plan.grid;
        if (grid_limit > 0 && grid > grid_limit) grid = grid_limit;
        if (plan.orbit_ok && !std::getenv("SB_EMUL_NO_ORBIT")) {
            bool aligned = ((reinterpret_cast<uintptr_t>(plan.map.base[0]) | reinterpret_cast<uintptr_t>(plan.map.base[1])) & 15u) == 0;
            if (aligned) {
                int g2 = (int)std::min<int64_t>(plan.orbit.nitems, 148);
                if (grid_limit > 0 && g2 > grid_limit) g2 = grid_limit;
                ok = plan.key.ct == F32 ? orbit_dispatch<float>(plan, g2) : plan.key.ct == F64 ? orbit_dispatch<double>(plan, g2) : false;
                if (ok) return SB_OK;
            }
        }
        if (plan.tma_ok && !std::getenv("SB_EMUL_NO_TMA")) {
            bool aligned = true;
            for (int k = 1; k < plan.map.nops; ++k) aligned = aligned && ((reinterpret_cast<uintptr_t>(plan.map.base[k]) & 15u) == 0);
            if (aligned) {
                switch (plan.key.ct) {
                case F32: ok = tma_dispatch<float>(plan, grid); break;
                case F64: ok = tma_dispatch<double>(plan, grid); break;
                case C32: ok = tma_dispatch<cx<float>>(plan, grid); break;
                default: ok = false; break;
                }
                if (ok) return SB_OK;
            }
        }
        switch (plan.key.ct) {
        case F32: ok = map_dispatch<float>(plan, grid); break;
        case F64: ok = map_dispatch<double>(plan, grid); break;
        case C32: ok = map_dispatch<cx<float>>(plan, grid); break;
        default: ok = map_dispatch<cx<double>>(plan, grid); break;
        }
    } else {
        std::vector<unsigned char> scratch((size_t)plan.scratch_bytes + 64);
        plan.red.scratch = scratch.data();
        if (plan.stream_ok && !std::getenv("SB_EMUL_NO_STREAM")) {
            bool aligned = true;
            for (int k = 1; k <= plan.stream.nin; ++k) aligned = aligned && ((reinterpret_cast<uintptr_t>(plan.red.base[k]) & 15u) == 0);
            if (aligned) {
                switch (plan.key.ct) {
                case F32: ok = stream_dispatch<float>(plan); break;
                case F64: ok = stream_dispatch<double>(plan); break;
                case C32: ok = stream_dispatch<cx<float>>(plan); break;
                default: ok = stream_dispatch<cx<double>>(plan); break;
                }
                if (ok) return SB_OK;
            }
        }
        switch (plan.key.ct) {
        case F32: ok = red_dispatch<float, 8>(plan); break;
        case F64: ok = red_dispatch<double, 8>(plan); break;
        case C32: ok = red_dispatch<cx<float>, 8>(plan); break;
        default: ok = red_dispatch<cx<double>, 4>(plan); break;
        }
    }
    if (!ok) {
        g_err = "emul: no instantiation for this kernel key";
        return SB_E_UNSUPPORTED;
    }
    return SB_OK;
}
