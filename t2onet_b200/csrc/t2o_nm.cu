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
#include <cstring>

#include "t2o_nm_device.cuh"

namespace t2o {

template <bool START>
__global__ void __launch_bounds__(128) nm_kernel(const __grid_constant__ NMArgs a) {
    const int lane = threadIdx.x & 31;
    const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (p >= a.P) return;
    if constexpr (START) {
        NMWarp w;
        nm_bind(w, a, p, lane);
        const int N = a.n_dims[p];
        w.N = N; w.maxfun = 200 * N; w.fcalls = 0; w.iters = 0; w.status = 0; w.fxrv = 0.0;
        // sim[0] = x0; sim[k+1] = x0 with coordinate k stepped: (1 + nonzdelt) * y[k] if y[k] != 0 else zdelt
        const double x0d = lane < N ? a.x0[(size_t)p * NM_MAXN + lane] : 0.0;
        for (int k = 0; k <= N; ++k) {
            double v = x0d;
            if (k >= 1 && lane == k - 1) v = (x0d != 0.0) ? __dmul_rn(a.nonz_scale, x0d) : a.zdelt;
            if (lane < N) w.sim[k * NM_MAXN + lane] = v;
        }
        w.f = CUDART_INF; w.row = lane;
        if (lane == 0) {
            w.ctl[CTL_N] = N; w.ctl[CTL_OP] = a.prob_op[p]; w.ctl[CTL_RES] = 0;
            *w.cop = a.prob_op[p];
        }
        __syncwarp();
        nm_propose(w, x0d, NM_INIT, 0);
        nm_store(w);
    } else {
        nm_advance_fit(a, p, lane, a.l1_sum[p]);
    }
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
