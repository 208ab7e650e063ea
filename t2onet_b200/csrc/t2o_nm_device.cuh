// t2o_nm_device.cuh -- the device-resident Nelder-Mead step (one warp per fit), shared by the round-by-round kernels of
// t2o_nm.cu and the resident planner kernel of t2o_score.cu.  See t2o_nm.cu for the algorithm and its provenance
// (scipy.optimize._minimize_neldermead as utils/beam_search.py:88 calls it).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <math_constants.h>

#include "../../include/t2o.h"
#include "t2o_common.cuh"

namespace t2o {


constexpr int NM_MAXN = T2O_MAX_OP_PARAMS;      // 24
constexpr int NM_ROWS = NM_MAXN + 1;            // simplex vertices
enum { NM_INIT = 0, NM_REFLECT = 1, NM_EXPAND = 2, NM_CONTRACT_OUT = 3, NM_CONTRACT_IN = 4, NM_SHRINK = 5, NM_DONE = 6 };
enum { CTL_N = 0, CTL_PHASE = 1, CTL_K = 2, CTL_FCALLS = 3, CTL_ITERS = 4, CTL_STATUS = 5, CTL_OP = 6, CTL_RES = 7 };
enum { VEC_XBAR = 0, VEC_XR = 1, VEC_PEND = 2 };

struct NMArgs {
    t2o_nm_state st;
    int P;
    const int *n_dims, *prob_op;    // start only
    const double *x0;               // start only
    const float *l1_sum;            // advance only
    float numel;
    float *cand_param;
    int *cand_op;
    double nonz_scale, zdelt, xatol, fatol;
};

// development probe (-DT2O_RES_PROBE): clocks per stage of the advance of 24-parameter fits, summed over all of them
#ifdef T2O_RES_PROBE
static __device__ unsigned long long g_nmp[12];
#define T2O_NMP(i) if (w.lane == 0 && w.N == NM_MAXN && w.store_result) { const long long now_ = clock64(); atomicAdd(&g_nmp[i], (unsigned long long)(now_ - w.pt)); w.pt = now_; }
#else
#define T2O_NMP(i)
#endif

struct NMWarp {
#ifdef T2O_RES_PROBE
    long long pt;
#endif
    // per-problem views
    double *sim, *vec, *fxr, *xbest, *fbest, *fsim;
    int *ctl, *perm;
    int ld;                         // doubles between the rows of sim / vec (NM_MAXN in the caller's arrays, N when packed in shared memory)
    bool store_result;              // nm_finish writes xbest / fbest (false in the CTAs that only shadow a fit)
    float *cparam;
    int *cop;
    int N, lane;
    // lane i <= N: function value and physical row of sorted position i
    double f;
    int row;
    int fcalls, iters, maxfun;
    // the step's control state (uniform over the lanes): what the pending vertex is (phase, k), scipy's status once finished,
    // the value of the reflected vertex -- in registers between nm_load and nm_store
    int phase, k, status;
    double fxrv;
};

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
// Double-precision compares and maxima through INTEGER instructions.  One warp per fit runs this code on the critical path of
// every evaluation, so what counts is the latency of each dependent instruction, and FP64 compares / min-max / |x| are
// multi-instruction, low-rate sequences -- FP64 is kept for the roundings scipy's arithmetic needs (sums, differences, the
// centroid) and everything that only ORDERS values works on bit patterns: the pattern of |x| orders like |x|, a NaN's is
// above +inf's, and function values (L1 distances: >= +0, +inf or NaN) order like their own patterns.
constexpr long long NM_INF_BITS = 0x7ff0000000000000ll;
__device__ __forceinline__ long long nm_abs_bits(double x) {
    // (through the two words: written as one 64-bit mask the compiler turns it back into an FP64 |x| instruction)
    return ((long long)(__double2hiint(x) & 0x7fffffff) << 32) | (long long)(unsigned)__double2loint(x);
}
__device__ __forceinline__ bool nm_isnan(double x) { return nm_abs_bits(x) > NM_INF_BITS; }
// a < b / a <= b as IEEE answers them, for a, b that are not negative (-0.0 included) unless NaN
__device__ __forceinline__ bool nm_lt(double a, double b) {
    return !nm_isnan(a) && !nm_isnan(b) && __double_as_longlong(a) < __double_as_longlong(b);
}
__device__ __forceinline__ bool nm_le(double a, double b) {
    return !nm_isnan(a) && !nm_isnan(b) && __double_as_longlong(a) <= __double_as_longlong(b);
}
// the sort key of a function value: NaN last, like numpy (as a pattern)
__device__ __forceinline__ long long nm_key(double f) { return nm_isnan(f) ? NM_INF_BITS : __double_as_longlong(f); }
// the maximum over the warp of non-negative patterns: two 32-bit reductions (REDUX) -- the high words, then the low words of
// the lanes that hold the largest high word
__device__ __forceinline__ long long warp_max_bits(long long bits) {
    const unsigned hi = (unsigned)(bits >> 32), lo = (unsigned)bits;
    const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
    const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
    return ((long long)mhi << 32) | (long long)mlo;
}

// stable rank sort of positions 0..N by function value (NaN last, like numpy).  Not inlined: it runs at the end of the initial
// simplex, after a shrink and on an exhausted budget only, and one copy keeps the kernels' code small.  (Values in and out in
// registers: a reference to the NMWarp would put the whole struct on the stack.)
struct NMSorted { double f; int row; };
static __device__ __noinline__ NMSorted nm_sort_impl(double f, int row, int N, int i) {
    const double key = (i <= N) ? (isnan(f) ? CUDART_INF : f) : CUDART_INF;
    int rank = 0;
#pragma unroll 4
    for (int j = 0; j <= N; ++j) {
        const double kj = shfl_d(key, j);
        rank += (kj < key || (kj == key && j < i)) ? 1 : 0;
    }
    NMSorted out{f, row};
#pragma unroll 4
    for (int j = 0; j <= N; ++j) {
        const int rj = __shfl_sync(0xffffffffu, rank, j);
        const double fj = shfl_d(f, j);
        const int rowj = __shfl_sync(0xffffffffu, row, j);
        if (rj == i) { out.f = fj; out.row = rowj; }
    }
    return out;
}
__device__ __forceinline__ void nm_sort(NMWarp &w) {
    const NMSorted r = nm_sort_impl(w.f, w.row, w.N, w.lane);
    w.f = r.f; w.row = r.row;
}

// The same order when only position N changed and positions 0..N-1 are already in order (every accepted vertex replaces the
// worst one): rank_i = i + [key_N < key_i] and rank_N = #{j < N: key_j <= key_N}, i.e. an insertion -- one ballot instead of
// 2 (N + 1) shuffle rounds.
__device__ __forceinline__ void nm_sort_last(NMWarp &w) {
    const int i = w.lane, N = w.N;
    const long long key = nm_key(w.f);
    const long long keyN = __shfl_sync(0xffffffffu, key, N);
    const int pos = __popc(__ballot_sync(0xffffffffu, i < N && key <= keyN));
    const double fN = shfl_d(w.f, N), fup = __shfl_up_sync(0xffffffffu, w.f, 1);
    const int rowN = __shfl_sync(0xffffffffu, w.row, N), rup = __shfl_up_sync(0xffffffffu, w.row, 1);
    if (i == pos) { w.f = fN; w.row = rowN; }
    else if (i > pos && i <= N) { w.f = fup; w.row = rup; }
}

// evaluate `x` (lane d holds x[d]) next: scipy's func() wrapper.  Returns false if the evaluation budget is spent
// (_MaxFun raised before the evaluation).
__device__ __forceinline__ bool nm_propose(NMWarp &w, double x, int phase, int k) {
    if (w.fcalls >= w.maxfun) return false;
    w.fcalls += 1;
    if (w.lane < NM_MAXN) {
        w.cparam[w.lane] = w.lane < w.N ? (float)x : 0.0f;      // float64 -> float32, as torch.tensor([param], dtype=torch.float)
        if (w.lane < w.N) w.vec[VEC_PEND * w.ld + w.lane] = x;
    }
    w.phase = phase; w.k = k;
    return true;
}

__device__ __forceinline__ void nm_finish(NMWarp &w, const NMArgs &a) {
    const int row0 = __shfl_sync(0xffffffffu, w.row, 0);
    const double f0 = shfl_d(w.f, 0);
    if (w.store_result && w.lane < NM_MAXN) w.xbest[w.lane] = w.lane < w.N ? w.sim[row0 * w.ld + w.lane] : 0.0;
    w.phase = NM_DONE;
    w.status = w.fcalls >= w.maxfun ? 1 : (w.iters >= w.maxfun ? 2 : 0);               // maxiter == maxfun == 200 N
    if (w.lane == 0) {
        if (w.store_result) *w.fbest = f0;
        *w.cop = T2O_OP_SKIP;
    }
}

// top of scipy's while loop: budget, convergence test, centroid, reflection.  NC: the number of parameters when it is known
// at compile time (the loop over the vertices unrolls: its shuffles and loads issue together), 0: w.N
template <int NC>
__device__ __forceinline__ void nm_begin_iteration_n(NMWarp &w, const NMArgs &a) {
    const int N = NC ? NC : w.N, d = w.lane;
    bool finish = !(w.fcalls < w.maxfun && w.iters < w.maxfun);
    if (!finish) {
        const int row0 = __shfl_sync(0xffffffffu, w.row, 0);
        const int rowN = __shfl_sync(0xffffffffu, w.row, N);
        // np.max(np.abs(sim[1:] - sim[0])) <= xatol and np.max(np.abs(fsim[0] - fsim[1:])) <= fatol: the values first -- when
        // they are apart (or NaN) the fit goes on whatever the vertices are, and the pass below only sums the centroid
        const double f0 = shfl_d(w.f, 0);
        const long long dfb = (d >= 1 && d <= N) ? nm_abs_bits(__dsub_rn(f0, w.f)) : 0ll;
        const bool need_dx = warp_max_bits(dfb) <= __double_as_longlong(a.fatol);         // (a NaN's pattern is above any tolerance)
        // (lanes >= N read coordinate 0 and their results are dropped: no divergent branch around the loads, so the unrolled
        // loop's shuffles and loads issue together)
        const int dd = d < N ? d : 0;
        const double x0d = w.sim[row0 * w.ld + dd];
        const double worst = w.sim[rowN * w.ld + dd];
        T2O_NMP(3)
        // one pass over the vertices in order: centroid of all but the worst, np.add.reduce(sim[:-1], 0) / N (row by row), and
        // (only when the values are close) the largest distance to the best vertex
        double s = x0d;
        long long dxb = 0ll;
#pragma unroll(NC ? NC : 4)
        for (int k = 1; k < N; ++k) {
            const int rk = __shfl_sync(0xffffffffu, w.row, k);
            const double v = w.sim[rk * w.ld + dd];
            s = __dadd_rn(s, v);
            if (need_dx) dxb = max(dxb, nm_abs_bits(__dsub_rn(v, x0d)));
        }
        T2O_NMP(4)
        if (need_dx) {
            dxb = max(dxb, nm_abs_bits(__dsub_rn(worst, x0d)));
            finish = warp_max_bits(d < N ? dxb : 0ll) <= __double_as_longlong(a.xatol);    // (NaN: above)
        }
        T2O_NMP(5)
        if (!finish) {
            double xr = 0.0;
            if (d < N) {
                const double xbar = __ddiv_rn(s, (double)N);
                xr = __dsub_rn(__dmul_rn(2.0, xbar), worst);           // (1 + rho) * xbar - rho * sim[-1]
                w.vec[VEC_XBAR * w.ld + d] = xbar;
                w.vec[VEC_XR * w.ld + d] = xr;
            }
            if (!nm_propose(w, xr, NM_REFLECT, 0)) {                    // _MaxFun: the iteration is dropped, sort, leave the loop
                nm_sort(w);
                finish = true;
            }
        }
    }
    T2O_NMP(6)
    if (finish) nm_finish(w, a);
}
// (the color operator's 24 and the tone operator's 8 parameters are what the planner's long fits have)
__device__ __forceinline__ void nm_begin_iteration(NMWarp &w, const NMArgs &a) {
    if (w.N == NM_MAXN) nm_begin_iteration_n<NM_MAXN>(w, a);
    else if (w.N == 8) nm_begin_iteration_n<8>(w, a);
    else nm_begin_iteration_n<0>(w, a);
}

// only_last: the iteration replaced the worst vertex and nothing else (positions 0..N-1 are still in order)
__device__ __forceinline__ void nm_end_iteration(NMWarp &w, const NMArgs &a, bool aborted, bool only_last = false) {
    if (!aborted) w.iters += 1;
    if (only_last) nm_sort_last(w);
    else nm_sort(w);
    T2O_NMP(2)
    nm_begin_iteration(w, a);
}

// replace the worst vertex by the vector stored at vec[which]
__device__ __forceinline__ void nm_accept(NMWarp &w, int which, double fval) {
    const int N = w.N;
    const int rowN = __shfl_sync(0xffffffffu, w.row, N);
    if (w.lane < N) w.sim[rowN * w.ld + w.lane] = w.vec[which * w.ld + w.lane];
    if (w.lane == N) w.f = fval;
}

// shrink vertex j towards the best one and evaluate it; false when the budget is spent (sim[j] is already moved, as in scipy)
__device__ __forceinline__ bool nm_shrink_vertex(NMWarp &w, int j) {
    const int row0 = __shfl_sync(0xffffffffu, w.row, 0);
    const int rowj = __shfl_sync(0xffffffffu, w.row, j);
    double x = 0.0;
    if (w.lane < w.N) {
        const double b = w.sim[row0 * w.ld + w.lane];
        const double v = w.sim[rowj * w.ld + w.lane];
        x = __dadd_rn(b, __dmul_rn(0.5, __dsub_rn(v, b)));     // sim[0] + sigma * (sim[j] - sim[0])
        w.sim[rowj * w.ld + w.lane] = x;
    }
    return nm_propose(w, x, NM_SHRINK, j);
}


// views of fit p in the caller-allocated state arrays
__device__ __forceinline__ void nm_bind(NMWarp &w, const NMArgs &a, int p, int lane) {
    w.sim = a.st.sim + (size_t)p * NM_ROWS * NM_MAXN;
    w.vec = a.st.vec + (size_t)p * 3 * NM_MAXN;
    w.fxr = a.st.fxr + p;
    w.xbest = a.st.xbest + (size_t)p * NM_MAXN;
    w.fbest = a.st.fbest + p;
    w.ctl = a.st.ctl + (size_t)p * 8;
    w.cparam = a.cand_param + (size_t)p * NM_MAXN;
    w.cop = a.cand_op + p;
    w.fsim = a.st.fsim + (size_t)p * NM_ROWS;
    w.perm = a.st.perm + (size_t)p * NM_ROWS;
    w.ld = NM_MAXN;
    w.lane = lane;
    w.store_result = true;
}

// The fit's control state and sorted values: memory (the caller's arrays, or the resident kernel's shared memory) -> registers
__device__ __forceinline__ void nm_load(NMWarp &w) {
    const int lane = w.lane;
    const int N = w.ctl[CTL_N];
    w.N = N; w.maxfun = 200 * N;
    w.phase = w.ctl[CTL_PHASE]; w.k = w.ctl[CTL_K]; w.fcalls = w.ctl[CTL_FCALLS]; w.iters = w.ctl[CTL_ITERS];
    w.status = w.ctl[CTL_STATUS];
    w.f = lane <= N ? w.fsim[lane] : CUDART_INF;
    w.row = lane <= N ? w.perm[lane] : lane;
    w.fxrv = *w.fxr;
}
// ... and back
__device__ __forceinline__ void nm_store(NMWarp &w) {
    const int lane = w.lane;
    __syncwarp();
    if (lane <= w.N) { w.fsim[lane] = w.f; w.perm[lane] = w.row; }
    if (lane == 0) {
        w.ctl[CTL_PHASE] = w.phase; w.ctl[CTL_K] = w.k; w.ctl[CTL_FCALLS] = w.fcalls; w.ctl[CTL_ITERS] = w.iters;
        w.ctl[CTL_STATUS] = w.status;
        *w.fxr = w.fxrv;
    }
    __syncwarp();
}

// the same state as a value: what a warp that owns a fit for many steps keeps in registers between them
struct NMRegs { double f, fxrv; int row, phase, k, fcalls, iters, status, N; };
__device__ __forceinline__ NMRegs nm_regs(const NMWarp &w) { return NMRegs{w.f, w.fxrv, w.row, w.phase, w.k, w.fcalls, w.iters, w.status, w.N}; }
__device__ __forceinline__ void nm_set_regs(NMWarp &w, const NMRegs &r) {
    w.f = r.f; w.fxrv = r.fxrv; w.row = r.row; w.phase = r.phase; w.k = r.k; w.fcalls = r.fcalls; w.iters = r.iters;
    w.status = r.status; w.N = r.N; w.maxfun = 200 * r.N;
}

// One evaluation result for the fit `w` holds (nm_load; l1 = the L1 sum of its pending vertex): consume it, move the simplex,
// propose the next vertex (or finish).  Called by all 32 lanes of one warp; a finished fit returns at once.
__device__ __forceinline__ void nm_step(NMWarp &w, const NMArgs &a, float l1) {
    const int lane = w.lane;
    {
#ifdef T2O_RES_PROBE
        w.pt = clock64();
#endif
        const int phase = w.phase;
        if (phase == NM_DONE) return;
        const int N = w.N, k = w.k;
        // the value of the pending vertex, (x1 - x2).norm(1) / numel -> .item(): fp32, then widened.  torch's CUDA
        // division by a host scalar multiplies by the rounded reciprocal; so does this, to the bit
        const double fv = (double)__fmul_rn(l1, __frcp_rn(a.numel));
        const double fxr = w.fxrv;
        const double f0 = shfl_d(w.f, 0), fN = shfl_d(w.f, N), fN1 = shfl_d(w.f, N >= 1 ? N - 1 : 0);
        T2O_NMP(0)
        // how the iteration ends: 0 not yet, 1 completed (full sort), 2 completed, only the worst vertex replaced (insertion),
        // 3 dropped by the evaluation budget -- ONE copy of the sort / convergence test / centroid code behind the switch
        int fin = 0;
        switch (phase) {
            case NM_INIT:
                if (lane == k) w.f = fv;
                if (k < N) {
                    const double x = lane < N ? w.sim[(k + 1) * w.ld + lane] : 0.0;
                    nm_propose(w, x, NM_INIT, k + 1);
                } else {
                    w.iters = 0;                                 // (scipy's counter starts at 1 after the initial simplex)
                    fin = 1;
                }
                break;
            case NM_REFLECT:
                w.fxrv = fv;
                if (nm_lt(fv, f0)) {
                    double xe = 0.0;
                    const int rowN = __shfl_sync(0xffffffffu, w.row, N);
                    if (lane < N) {
                        const double xbar = w.vec[VEC_XBAR * w.ld + lane], worst = w.sim[rowN * w.ld + lane];
                        xe = __dsub_rn(__dmul_rn(3.0, xbar), __dmul_rn(2.0, worst));         // (1 + rho chi) xbar - rho chi sim[-1]
                    }
                    if (!nm_propose(w, xe, NM_EXPAND, 0)) fin = 3;
                } else if (nm_lt(fv, fN1)) {
                    nm_accept(w, VEC_XR, fv);
                    fin = 2;
                } else {
                    const int rowN = __shfl_sync(0xffffffffu, w.row, N);
                    double xc = 0.0;
                    const bool outside = nm_lt(fv, fN);
                    if (lane < N) {
                        const double xbar = w.vec[VEC_XBAR * w.ld + lane], worst = w.sim[rowN * w.ld + lane];
                        xc = outside ? __dsub_rn(__dmul_rn(1.5, xbar), __dmul_rn(0.5, worst))       // (1 + psi rho) xbar - psi rho sim[-1]
                                     : __dadd_rn(__dmul_rn(0.5, xbar), __dmul_rn(0.5, worst));      // (1 - psi) xbar + psi sim[-1]
                    }
                    if (!nm_propose(w, xc, outside ? NM_CONTRACT_OUT : NM_CONTRACT_IN, 0)) fin = 3;
                }
                break;
            case NM_EXPAND:
                if (nm_lt(fv, fxr)) nm_accept(w, VEC_PEND, fv);
                else nm_accept(w, VEC_XR, fxr);
                fin = 2;
                break;
            case NM_CONTRACT_OUT:
            case NM_CONTRACT_IN:
                if (phase == NM_CONTRACT_OUT ? nm_le(fv, fxr) : nm_lt(fv, fN)) {
                    nm_accept(w, VEC_PEND, fv);
                    fin = 2;
                } else if (!nm_shrink_vertex(w, 1)) {
                    fin = 3;
                }
                break;
            case NM_SHRINK:
                if (lane == k) w.f = fv;
                if (k < N) {
                    if (!nm_shrink_vertex(w, k + 1)) fin = 3;
                } else {
                    fin = 1;
                }
                break;
            default: break;
        }
        T2O_NMP(1)
        if (fin) nm_end_iteration(w, a, fin == 3, fin == 2);
    }
    T2O_NMP(7)
#ifdef T2O_RES_PROBE
    if (w.lane == 0 && w.N == NM_MAXN && w.store_result) atomicAdd(&g_nmp[8], 1ull);
#endif
}

// load, step, store: the round-by-round driver's advance (state in the caller's device arrays)
__device__ __forceinline__ void nm_advance_fit(const NMArgs &a, int p, int lane, float l1) {
    NMWarp w;
    nm_bind(w, a, p, lane);
    nm_load(w);
    if (w.phase == NM_DONE) return;
    nm_step(w, a, l1);
    nm_store(w);
}

}  // namespace t2o
