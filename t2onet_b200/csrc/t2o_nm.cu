// t2o_nm.cu -- device-resident Nelder-Mead for the operation planner (sm_100a).
//
// Replaces: scipy.optimize.minimize(func, param0, method='Nelder-Mead') as the reference calls it once per
// (state, operator) pair (utils/beam_search.py:88 inside get_param_naive, :65-91).  scipy is third-party and
// un-pinned by the reference; the algorithm restated here is scipy 1.x `_minimize_neldermead` with its defaults
// (rho = 1, chi = 2, psi = sigma = 0.5, nonzdelt = 0.05, zdelt = 0.00025, xatol = fatol = 1e-4,
// maxiter = maxfev = 200 N, the evaluation-count guard that aborts the running iteration) -- the same
// restatement as the host coroutine t2onet_b200/nelder_mead.py, which the tests compare it with.
//
// Why on the device: one planner step fits every (beam state, operator) pair of every image in flight; each fit is a
// strictly sequential chain of up to 200 N function evaluations, and an evaluation is one candidate of
// t2o_score_candidates.  With the simplex bookkeeping on the host every evaluation round costs a launch, a
// device->host read of the scores and a host->device copy of the next vertices; here a round is two launches
// (t2o_nm_advance, t2o_score_candidates) that exchange vertices and scores in device memory, so rounds can be
// enqueued back to back (or replayed as a CUDA graph) without any host synchronisation.
//
// One warp owns one fit: lane d owns coordinate d of every vertex (N <= 24), lane i owns sorted position i of the
// simplex (its function value and the row it lives in); the control flow is warp-uniform.  The simplex arithmetic
// is float64 in scipy's operation order with explicitly rounded operations (no FMA contraction), so that equal
// function values give the same vertex sequence.  Sorting is a stable rank sort (numpy's argsort is not stable:
// exact ties between function values may order differently).
#include <cuda_runtime.h>
#include <math.h>
#include <math_constants.h>

#include <cstring>

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

struct NMWarp {
    // per-problem views
    double *sim, *vec, *fxr, *xbest, *fbest;
    int *ctl;
    float *cparam;
    int *cop;
    int N, lane;
    // lane i <= N: function value and physical row of sorted position i
    double f;
    int row;
    int fcalls, iters, maxfun;
};

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ double warp_max_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// stable rank sort of positions 0..N by function value (NaN last, like numpy)
__device__ __forceinline__ void nm_sort(NMWarp &w) {
    const int i = w.lane, N = w.N;
    const double key = (i <= N) ? (isnan(w.f) ? CUDART_INF : w.f) : CUDART_INF;
    int rank = 0;
    for (int j = 0; j <= N; ++j) {
        const double kj = shfl_d(key, j);
        rank += (kj < key || (kj == key && j < i)) ? 1 : 0;
    }
    double nf = w.f;
    int nrow = w.row;
    for (int j = 0; j <= N; ++j) {
        const int rj = __shfl_sync(0xffffffffu, rank, j);
        const double fj = shfl_d(w.f, j);
        const int rowj = __shfl_sync(0xffffffffu, w.row, j);
        if (rj == i) { nf = fj; nrow = rowj; }
    }
    w.f = nf; w.row = nrow;
}

// evaluate `x` (lane d holds x[d]) next: scipy's func() wrapper.  Returns false if the evaluation budget is spent
// (_MaxFun raised before the evaluation).
__device__ __forceinline__ bool nm_propose(NMWarp &w, double x, int phase, int k) {
    if (w.fcalls >= w.maxfun) return false;
    w.fcalls += 1;
    if (w.lane < NM_MAXN) {
        w.cparam[w.lane] = w.lane < w.N ? (float)x : 0.0f;      // float64 -> float32, as torch.tensor([param], dtype=torch.float)
        if (w.lane < w.N) w.vec[VEC_PEND * NM_MAXN + w.lane] = x;
    }
    if (w.lane == 0) { w.ctl[CTL_PHASE] = phase; w.ctl[CTL_K] = k; }
    return true;
}

__device__ __forceinline__ void nm_finish(NMWarp &w, const NMArgs &a) {
    const int row0 = __shfl_sync(0xffffffffu, w.row, 0);
    const double f0 = shfl_d(w.f, 0);
    if (w.lane < NM_MAXN) w.xbest[w.lane] = w.lane < w.N ? w.sim[row0 * NM_MAXN + w.lane] : 0.0;
    if (w.lane == 0) {
        *w.fbest = f0;
        w.ctl[CTL_PHASE] = NM_DONE;
        w.ctl[CTL_STATUS] = w.fcalls >= w.maxfun ? 1 : (w.iters >= w.maxfun ? 2 : 0);   // maxiter == maxfun == 200 N
        *w.cop = T2O_OP_SKIP;
    }
}

// top of scipy's while loop: budget, convergence test, centroid, reflection
__device__ void nm_begin_iteration(NMWarp &w, const NMArgs &a) {
    const int N = w.N, d = w.lane;
    if (!(w.fcalls < w.maxfun && w.iters < w.maxfun)) { nm_finish(w, a); return; }
    const int row0 = __shfl_sync(0xffffffffu, w.row, 0);
    const int rowN = __shfl_sync(0xffffffffu, w.row, N);
    // np.max(np.abs(sim[1:] - sim[0])) <= xatol and np.max(np.abs(fsim[0] - fsim[1:])) <= fatol
    double dx = 0.0, s = 0.0;
    bool bad = false;
    const double x0d = d < N ? w.sim[row0 * NM_MAXN + d] : 0.0;
    for (int k = 1; k <= N; ++k) {
        const int rk = __shfl_sync(0xffffffffu, w.row, k);
        if (d < N) {
            const double v = w.sim[rk * NM_MAXN + d];
            const double e = fabs(__dsub_rn(v, x0d));
            bad |= isnan(e);
            dx = fmax(dx, e);
        }
    }
    // centroid of all but the worst vertex: np.add.reduce(sim[:-1], 0) / N  (row by row)
    for (int k = 0; k < N; ++k) {
        const int rk = __shfl_sync(0xffffffffu, w.row, k);
        if (d < N) { const double v = w.sim[rk * NM_MAXN + d]; s = k == 0 ? v : __dadd_rn(s, v); }
    }
    const double f0 = shfl_d(w.f, 0);
    double df = (d >= 1 && d <= N) ? fabs(__dsub_rn(f0, w.f)) : 0.0;
    bad |= isnan(df);
    dx = warp_max_d(dx);
    df = warp_max_d(df);
    const bool anybad = __any_sync(0xffffffffu, bad);
    if (!anybad && dx <= a.xatol && df <= a.fatol) { nm_finish(w, a); return; }
    double xr = 0.0;
    if (d < N) {
        const double xbar = __ddiv_rn(s, (double)N);
        const double worst = w.sim[rowN * NM_MAXN + d];
        xr = __dsub_rn(__dmul_rn(2.0, xbar), worst);           // (1 + rho) * xbar - rho * sim[-1]
        w.vec[VEC_XBAR * NM_MAXN + d] = xbar;
        w.vec[VEC_XR * NM_MAXN + d] = xr;
    }
    if (!nm_propose(w, xr, NM_REFLECT, 0)) {                    // _MaxFun: the iteration is dropped, sort, leave the loop
        nm_sort(w);
        nm_finish(w, a);
    }
}

__device__ __forceinline__ void nm_end_iteration(NMWarp &w, const NMArgs &a, bool aborted) {
    if (!aborted) w.iters += 1;
    nm_sort(w);
    nm_begin_iteration(w, a);
}

// replace the worst vertex by the vector stored at vec[which]
__device__ __forceinline__ void nm_accept(NMWarp &w, int which, double fval) {
    const int N = w.N;
    const int rowN = __shfl_sync(0xffffffffu, w.row, N);
    if (w.lane < N) w.sim[rowN * NM_MAXN + w.lane] = w.vec[which * NM_MAXN + w.lane];
    if (w.lane == N) w.f = fval;
}

// shrink vertex j towards the best one and evaluate it; false when the budget is spent (sim[j] is already moved, as in scipy)
__device__ __forceinline__ bool nm_shrink_vertex(NMWarp &w, int j) {
    const int row0 = __shfl_sync(0xffffffffu, w.row, 0);
    const int rowj = __shfl_sync(0xffffffffu, w.row, j);
    double x = 0.0;
    if (w.lane < w.N) {
        const double b = w.sim[row0 * NM_MAXN + w.lane];
        const double v = w.sim[rowj * NM_MAXN + w.lane];
        x = __dadd_rn(b, __dmul_rn(0.5, __dsub_rn(v, b)));     // sim[0] + sigma * (sim[j] - sim[0])
        w.sim[rowj * NM_MAXN + w.lane] = x;
    }
    return nm_propose(w, x, NM_SHRINK, j);
}

template <bool START>
__global__ void __launch_bounds__(128) nm_kernel(const __grid_constant__ NMArgs a) {
    const int lane = threadIdx.x & 31;
    const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= a.P) return;
    NMWarp w;
    w.sim = a.st.sim + (size_t)p * NM_ROWS * NM_MAXN;
    w.vec = a.st.vec + (size_t)p * 3 * NM_MAXN;
    w.fxr = a.st.fxr + p;
    w.xbest = a.st.xbest + (size_t)p * NM_MAXN;
    w.fbest = a.st.fbest + p;
    w.ctl = a.st.ctl + (size_t)p * 8;
    w.cparam = a.cand_param + (size_t)p * NM_MAXN;
    w.cop = a.cand_op + p;
    w.lane = lane;
    double *fsim = a.st.fsim + (size_t)p * NM_ROWS;
    int *perm = a.st.perm + (size_t)p * NM_ROWS;

    if constexpr (START) {
        const int N = a.n_dims[p];
        w.N = N; w.maxfun = 200 * N; w.fcalls = 0; w.iters = 0;
        // sim[0] = x0; sim[k+1] = x0 with coordinate k stepped: (1 + nonzdelt) * y[k] if y[k] != 0 else zdelt
        const double x0d = lane < N ? a.x0[(size_t)p * NM_MAXN + lane] : 0.0;
        for (int k = 0; k <= N; ++k) {
            double v = x0d;
            if (k >= 1 && lane == k - 1) v = (x0d != 0.0) ? __dmul_rn(a.nonz_scale, x0d) : a.zdelt;
            if (lane < N) w.sim[k * NM_MAXN + lane] = v;
        }
        w.f = CUDART_INF; w.row = lane;
        if (lane == 0) {
            w.ctl[CTL_N] = N; w.ctl[CTL_STATUS] = 0; w.ctl[CTL_OP] = a.prob_op[p]; w.ctl[CTL_RES] = 0;
            *w.cop = a.prob_op[p];
        }
        __syncwarp();
        nm_propose(w, x0d, NM_INIT, 0);
    } else {
        const int phase = w.ctl[CTL_PHASE];
        if (phase == NM_DONE) return;
        const int N = w.ctl[CTL_N], k = w.ctl[CTL_K];
        w.N = N; w.maxfun = 200 * N; w.fcalls = w.ctl[CTL_FCALLS]; w.iters = w.ctl[CTL_ITERS];
        w.f = lane <= N ? fsim[lane] : CUDART_INF;
        w.row = lane <= N ? perm[lane] : lane;
        // the value of the pending vertex, (x1 - x2).norm(1) / numel -> .item(): fp32, then widened.  torch's CUDA
        // division by a host scalar multiplies by the rounded reciprocal; so does this, to the bit
        const double fv = (double)__fmul_rn(a.l1_sum[p], __frcp_rn(a.numel));
        const double fxr = *w.fxr;
        __syncwarp();
        const double f0 = shfl_d(w.f, 0), fN = shfl_d(w.f, N), fN1 = shfl_d(w.f, N >= 1 ? N - 1 : 0);
        switch (phase) {
            case NM_INIT:
                if (lane == k) w.f = fv;
                if (k < N) {
                    const double x = lane < N ? w.sim[(k + 1) * NM_MAXN + lane] : 0.0;
                    nm_propose(w, x, NM_INIT, k + 1);
                } else {
                    nm_sort(w);
                    w.iters = 1;
                    nm_begin_iteration(w, a);
                }
                break;
            case NM_REFLECT:
                if (lane == 0) *w.fxr = fv;
                if (fv < f0) {
                    double xe = 0.0;
                    const int rowN = __shfl_sync(0xffffffffu, w.row, N);
                    if (lane < N) {
                        const double xbar = w.vec[VEC_XBAR * NM_MAXN + lane], worst = w.sim[rowN * NM_MAXN + lane];
                        xe = __dsub_rn(__dmul_rn(3.0, xbar), __dmul_rn(2.0, worst));         // (1 + rho chi) xbar - rho chi sim[-1]
                    }
                    if (!nm_propose(w, xe, NM_EXPAND, 0)) nm_end_iteration(w, a, true);
                } else if (fv < fN1) {
                    nm_accept(w, VEC_XR, fv);
                    nm_end_iteration(w, a, false);
                } else {
                    const int rowN = __shfl_sync(0xffffffffu, w.row, N);
                    double xc = 0.0;
                    const bool outside = fv < fN;
                    if (lane < N) {
                        const double xbar = w.vec[VEC_XBAR * NM_MAXN + lane], worst = w.sim[rowN * NM_MAXN + lane];
                        xc = outside ? __dsub_rn(__dmul_rn(1.5, xbar), __dmul_rn(0.5, worst))       // (1 + psi rho) xbar - psi rho sim[-1]
                                     : __dadd_rn(__dmul_rn(0.5, xbar), __dmul_rn(0.5, worst));      // (1 - psi) xbar + psi sim[-1]
                    }
                    if (!nm_propose(w, xc, outside ? NM_CONTRACT_OUT : NM_CONTRACT_IN, 0)) nm_end_iteration(w, a, true);
                }
                break;
            case NM_EXPAND:
                if (fv < fxr) nm_accept(w, VEC_PEND, fv);
                else nm_accept(w, VEC_XR, fxr);
                nm_end_iteration(w, a, false);
                break;
            case NM_CONTRACT_OUT:
            case NM_CONTRACT_IN:
                if (phase == NM_CONTRACT_OUT ? (fv <= fxr) : (fv < fN)) {
                    nm_accept(w, VEC_PEND, fv);
                    nm_end_iteration(w, a, false);
                } else if (!nm_shrink_vertex(w, 1)) {
                    nm_end_iteration(w, a, true);
                }
                break;
            case NM_SHRINK:
                if (lane == k) w.f = fv;
                if (k < N) {
                    if (!nm_shrink_vertex(w, k + 1)) nm_end_iteration(w, a, true);
                } else {
                    nm_end_iteration(w, a, false);
                }
                break;
            default: break;
        }
    }
    __syncwarp();
    if (lane <= w.N) { fsim[lane] = w.f; perm[lane] = w.row; }
    if (lane == 0) { w.ctl[CTL_FCALLS] = w.fcalls; w.ctl[CTL_ITERS] = w.iters; }
}

// ---------------------------------------------------------------- top-k of the candidates' scores
// The beam selection of a planner step (utils/beam_search.py:252-256: np.argsort of the candidates' distances, the first
// beam_size kept) for many searches at once: segment s of `values` holds the candidates of search s; the k smallest come
// out in ascending order, exact ties by the smaller index (numpy's kind='stable'; the reference's default quicksort leaves
// the order of ties unspecified), NaN last.  One warp per segment: pass j picks the smallest (value, index) key that is
// greater than the key picked in pass j - 1, so nothing is marked or moved.
__global__ void __launch_bounds__(128) topk_min_kernel(const float *__restrict__ values, const int *__restrict__ seg_begin, int n_seg, int k,
                                                       int *__restrict__ out_idx, float *__restrict__ out_val) {
    const int lane = threadIdx.x & 31;
    const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (s >= n_seg) return;
    const int b = seg_begin[s], e = seg_begin[s + 1];
    float last_v = -CUDART_INF_F;
    int last_i = -1;
    for (int j = 0; j < k; ++j) {
        float best_v = CUDART_INF_F;
        int best_i = 0x7fffffff;
        for (int i = b + lane; i < e; i += 32) {
            float v = values[i];
            if (isnan(v)) v = CUDART_INF_F;
            const bool after = v > last_v || (v == last_v && i > last_i);
            if (after && (v < best_v || (v == best_v && i < best_i))) { best_v = v; best_i = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best_v, o);
            const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
            if (ov < best_v || (ov == best_v && oi < best_i)) { best_v = ov; best_i = oi; }
        }
        const bool found = best_i != 0x7fffffff;
        if (lane == 0) {
            out_idx[(size_t)s * k + j] = found ? best_i : -1;
            out_val[(size_t)s * k + j] = found ? values[best_i] : CUDART_INF_F;
        }
        if (!found) { last_v = CUDART_INF_F; last_i = 0x7fffffff; } else { last_v = best_v; last_i = best_i; }
    }
}

int topk_min(const float *values, const int *seg_begin, int n_seg, int k, int *out_idx, float *out_val, cudaStream_t stream) {
    if (!values || !seg_begin || !out_idx || !out_val || n_seg < 1 || k < 1) return T2O_ERR_INVALID_ARG;
    topk_min_kernel<<<(n_seg + 3) / 4, 128, 0, stream>>>(values, seg_begin, n_seg, k, out_idx, out_val);
    T2O_CUDA_OK(cudaGetLastError());
    return T2O_OK;
}

static int nm_check(const t2o_nm_state *st, int P) {
    if (!st || P < 1) return T2O_ERR_INVALID_ARG;
    if (!st->sim || !st->fsim || !st->vec || !st->fxr || !st->perm || !st->ctl || !st->xbest || !st->fbest) return T2O_ERR_INVALID_ARG;
    return T2O_OK;
}

int nm_start(const t2o_nm_state *st, int P, const int *n_dims, const int *prob_op, const double *x0,
             float *cand_param, int *cand_op, cudaStream_t stream) {
    int s = nm_check(st, P);
    if (s != T2O_OK) return s;
    if (!n_dims || !prob_op || !x0 || !cand_param || !cand_op) return T2O_ERR_INVALID_ARG;
    NMArgs a;
    memset(&a, 0, sizeof(a));
    a.st = *st; a.P = P; a.n_dims = n_dims; a.prob_op = prob_op; a.x0 = x0; a.cand_param = cand_param; a.cand_op = cand_op;
    a.nonz_scale = 1 + 0.05; a.zdelt = 0.00025; a.xatol = 1e-4; a.fatol = 1e-4;
    nm_kernel<true><<<(P + 3) / 4, 128, 0, stream>>>(a);
    T2O_CUDA_OK(cudaGetLastError());
    return T2O_OK;
}

int nm_advance(const t2o_nm_state *st, int P, const float *l1_sum, float numel, float *cand_param, int *cand_op,
               cudaStream_t stream) {
    int s = nm_check(st, P);
    if (s != T2O_OK) return s;
    if (!l1_sum || !cand_param || !cand_op || !(numel > 0.0f)) return T2O_ERR_INVALID_ARG;
    NMArgs a;
    memset(&a, 0, sizeof(a));
    a.st = *st; a.P = P; a.l1_sum = l1_sum; a.numel = numel; a.cand_param = cand_param; a.cand_op = cand_op;
    a.nonz_scale = 1 + 0.05; a.zdelt = 0.00025; a.xatol = 1e-4; a.fatol = 1e-4;
    nm_kernel<false><<<(P + 3) / 4, 128, 0, stream>>>(a);
    T2O_CUDA_OK(cudaGetLastError());
    return T2O_OK;
}

}  // namespace t2o
