/*
 * t2o.h -- C-ABI of the B200-native T2ONet operator / planner hot path.
 *
 * The reference (jshi31/T2ONet) has no FFI: its operator API is a Python class surface
 * (models/operators.py, executors/executor.py, utils/beam_search.py).  The entry points
 * below are what a binding for that surface calls instead of the chain of eager PyTorch
 * kernels; t2onet_b200/ binds them with ctypes (see INTEGRATION.md for the stub a
 * maintainer of the reference would add).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its comment says "host"
 *   - images are float32, NCHW planar, contiguous: (B, 3, H, W), values nominally in [0, 1]
 *   - masks are float32 (B, 1|3, H, W) or NULL (= all ones, models/operators.py:123)
 *   - parameters are float32 rows: op k of image b reads
 *         params[b * param_stride + param_off[k] + i],  i < t2o_num_params(op_ids[k])
 *     with the reference's layouts (tone: (L,), color: (3, L) index c*L+i, models/operators.py:578,608)
 *   - all work is enqueued on `stream`; nothing synchronises the device
 *   - `workspace` is caller-owned device scratch of >= t2o_workspace_bytes(...) bytes whose first 256 KiB
 *     (the per-image / per-state arrival counters) must be ZERO before its first use; every call leaves
 *     them zeroed again, whatever the batch size, so one workspace can serve all entry points in turn
 *     (never two launches that overlap in time); the bytes behind the counters are plain scratch
 *   - every function returns a t2o_status (0 = ok) and never throws; inputs are never modified
 */
#ifndef T2O_H_
#define T2O_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define T2O_VERSION 104          /* major*100 + minor */
#define T2O_MAX_CHAIN 8          /* operators fused in one launch */
#define T2O_MAX_CURVE_STEPS 8    /* cfg.curve_steps (options/fiveK_base_options.py:50) */
#define T2O_MAX_OP_PARAMS 24     /* color: 3 * curve_steps */

/* flags of t2o_chain_forward */
#define T2O_FLAG_RAW_PROCESS 1   /* n_ops == 1 only: write Operator.process(img, param) itself
                                    (models/operators.py:128), without the mask blend and clamp */

/* Operator ids = reference Executor indices (executors/executor.py:30) + extension ids for the
 * operator classes that exist without an Executor slot (models/operators.py:186,527 and :298,373,414). */
enum t2o_op {
    T2O_OP_SKIP = -2,         /* t2o_score_candidates only: the candidate is not evaluated (a finished Nelder-Mead fit) */
    T2O_OP_IDENTITY = -1,     /* executors/executor.py:44-46 (op_ind < 0): passthrough, no clamp */
    T2O_OP_BRIGHTNESS = 0,    /* models/operators.py:277-283 */
    T2O_OP_CONTRAST = 1,      /* models/operators.py:240-245 */
    T2O_OP_SATURATION = 2,    /* models/operators.py:473-479 */
    T2O_OP_COLOR = 3,         /* models/operators.py:607-616 */
    T2O_OP_INPAINT = 4,       /* models/operators.py:680-682 -- NOT implemented (deep CNN, out of scope) */
    T2O_OP_TONE = 5,          /* models/operators.py:571-585 */
    T2O_OP_SHARPNESS = 6,     /* models/operators.py:351-358 */
    T2O_OP_WHITE = 7,         /* models/operators.py:510-512 */
    T2O_OP_EXPOSURE = 8,      /* models/operators.py:209-210 */
    T2O_OP_WHITEBALANCE = 9,  /* models/operators.py:548-549 */
    T2O_OP_BNW = 10,          /* models/operators.py:314-316   lerp(img, luminance, p) */
    T2O_OP_BLUR = 11,         /* models/operators.py:397-404   lerp(img, 3x3 Gaussian (sigma 2, zero padding) * img, p): a stencil
                                 operator like sharpness -- a launch holds at most one of the two */
    T2O_OP_HUE = 12           /* models/operators.py:432-438   hsv_to_rgb(p, s, v): every pixel's hue replaced by p (radians) */
};

enum t2o_status {
    T2O_OK = 0,
    T2O_ERR_INVALID_ARG = 1,      /* NULL where required, bad sizes, bad op id */
    T2O_ERR_UNSUPPORTED = 2,      /* inpaint, curve_steps > 8, an operator type twice in one backward launch, ... */
    T2O_ERR_WORKSPACE = 3,        /* workspace too small */
    T2O_ERR_CUDA = 4,             /* a CUDA runtime call failed (see t2o_last_cuda_error) */
    T2O_ERR_NO_DEVICE = 5         /* no sm_100 device / driver entry point missing */
};

typedef struct CUstream_st *t2o_stream_t;   /* == cudaStream_t */

int t2o_version(void);
const char *t2o_status_string(int status);
const char *t2o_last_cuda_error(void);
/* number of parameters of one operator (num_op_param of each class in models/operators.py) */
int t2o_num_params(int op_id, int curve_steps);

/* Scratch needed by the chain entry points for a (B, 3, H, W) batch with `param_stride` floats per row. */
size_t t2o_workspace_bytes(int B, int H, int W, int param_stride);
/* Scratch needed by t2o_score_candidates for S states and C candidates on (H, W) images. */
size_t t2o_score_workspace_bytes(int S, int C, int H, int W);

/*
 * K fused Operator.execute steps (models/operators.py:112-131: process -> out*mask + img*(1-mask)
 * -> clamp(0,1)), i.e. K successive Executor.execute calls with specified parameters
 * (executors/executor.py:33-55), in ONE pass over HBM.  K == 1 is a single Operator.execute.
 *   out      (B,3,H,W) or NULL      edited image
 *   target   (B,3,H,W) or NULL      with l1_sum: per-image sum |out - target| (the numerator of
 *   l1_sum   (B,)      or NULL      get_dist 'L1', utils/beam_search.py:170-173, and of the training L1)
 * At most one sharpness operator per launch (the binding splits longer chains).
 */
int t2o_chain_forward(int n_ops, const int *op_ids /*host*/, const int *param_off /*host*/,
                      const float *img, const float *mask, int mask_ch,
                      const float *params, int param_stride,
                      const float *target, float *out, float *l1_sum,
                      int B, int H, int W, int curve_steps, int flags,
                      void *workspace, size_t workspace_bytes, t2o_stream_t stream);

/*
 * Backward of t2o_chain_forward with the forward RECOMPUTED in the same pass (no stored
 * intermediates; replaces autograd through the ~50 saved planes per operator).
 * Upstream gradient, one of:
 *   grad_out (B,3,H,W)                         dLoss/d(out), or
 *   grad_out == NULL, target + grad_l1 (B,)    dLoss/d(out) = grad_l1[b] * sign(out - target)
 *                                              (the fused L1; grad_l1[b] = dLoss/d(l1_sum[b]))
 * Outputs:
 *   grad_params (B, param_stride)   every column of the row is written (0 outside the used slots)
 *   grad_img    (B,3,H,W) or NULL   dLoss/d(img)
 *   out, l1_sum           or NULL   the forward results, for a fused forward+backward step
 * Every operator type (sharpness included) may appear at most once per launch: the kernel keeps one register
 * accumulator slot per type (T2O_ERR_UNSUPPORTED otherwise; the binding splits longer chains into launches).
 * Gradient conventions follow torch autograd on the reference graph: clamp passes the gradient on
 * the closed interval, the contrast luminance clamp splits ties 0.5/0.5, HSV max/min route to the
 * first tied channel (exact for gray pixels; two-channel ties see DESIGN.md).
 */
int t2o_chain_backward(int n_ops, const int *op_ids /*host*/, const int *param_off /*host*/,
                       const float *img, const float *mask, int mask_ch,
                       const float *params, int param_stride,
                       const float *grad_out, const float *target, const float *grad_l1,
                       float *grad_params, float *grad_img, float *out, float *l1_sum,
                       int B, int H, int W, int curve_steps,
                       void *workspace, size_t workspace_bytes, t2o_stream_t stream);

/*
 * Per-row chains: the Actor call sites (models/actor.py:100-114 divide_op_group + :165,252,340): every batch row
 * applies its OWN operator (K = 1 per decoding step; K > 1 = K decoding steps with known parameters).  Replaces
 * the reference's group-by-operator loop (torch.unique + index_select gathers + one Executor.execute per group +
 * cat + index_select back) by one launch per tiling over the original tensors: no image copies, no host sync.
 *   row_ops       (B, K) int32, DEVICE: operator id of step k of row b (T2O_OP_IDENTITY = <END>, passes through)
 *   row_ops_host  the same ids on the HOST, or NULL.  With a host copy the rows are validated up front (an
 *                 operator type twice in one row of a backward call, inpaint, a second stencil -> error status)
 *                 and only the tilings some row needs are launched.  NULL requires K == 1; the kernels then
 *                 treat an invalid id as identity and set bit 0 of *status (DEVICE uint32, may be NULL).
 *   params        (B, param_stride): step k of a row reads its parameters at column k * param_slot
 *                 (param_slot >= 3 * curve_steps; 24 = the Actor's zero-padded parameter rows, models/actor.py:166)
 *   grad_params   same layout, every column written (0 outside the row's used slots)
 * Everything else as in t2o_chain_forward / t2o_chain_backward.
 */
int t2o_rows_forward(int K, const int32_t *row_ops, const int32_t *row_ops_host /*host*/, int param_slot,
                     const float *img, const float *mask, int mask_ch,
                     const float *params, int param_stride,
                     const float *target, float *out, float *l1_sum, uint32_t *status,
                     int B, int H, int W, int curve_steps,
                     void *workspace, size_t workspace_bytes, t2o_stream_t stream);
int t2o_rows_backward(int K, const int32_t *row_ops, const int32_t *row_ops_host /*host*/, int param_slot,
                      const float *img, const float *mask, int mask_ch,
                      const float *params, int param_stride,
                      const float *grad_out, const float *target, const float *grad_l1,
                      float *grad_params, float *grad_img, float *out, float *l1_sum, uint32_t *status,
                      int B, int H, int W, int curve_steps,
                      void *workspace, size_t workspace_bytes, t2o_stream_t stream);

/* Per-image sum |a - b| over n floats per image: get_dist(x1, x2, 'L1') * numel, utils/beam_search.py:170-173.
 * workspace: >= 256 KiB + B * (n_per_image / 4096 + 2) * 4 bytes. */
int t2o_l1_sum(const float *a, const float *b, float *l1_sum, int B, int64_t n_per_image,
               void *workspace, size_t workspace_bytes, t2o_stream_t stream);

/*
 * SSIM of image pairs (evaluation metric, utils/ssim/__init__.py:19-41 as utils/eval.py:57-60 uses it: 11x11 Gaussian
 * window, sigma 1.5, zero padding, C1 = 0.01^2, C2 = 0.03^2): ssim_sum[b] = sum over channels and pixels of the SSIM map
 * of (img1[b], img2[b]); the reference's value is ssim_sum / (C*H*W) (its mean over the batch: the sums' total / numel).
 * img1, img2 (B, C, H, W) float32 contiguous.  One pass over HBM, nothing stored but the sums.
 */
size_t t2o_ssim_workspace_bytes(int B, int C, int H, int W);
int t2o_ssim_sum(const float *img1, const float *img2, float *ssim_sum, int B, int C, int H, int W,
                 void *workspace, size_t workspace_bytes, t2o_stream_t stream);

/*
 * 8-bit image <-> float32 tensor conversions (utils/visual_utils.py of the reference), so that images cross PCIe as the
 * uint8 arrays they are and the x / 255 happens in HBM.  Bit-identical to the reference's host arithmetic:
 *   t2o_u8_to_f32      dst[i] = float(src[i]) / 255  (IEEE division, utils/visual_utils.py:46,67)       any layout, n values
 *   t2o_f32_to_u8      dst[i] = uint8(src[i] * 255)  (truncation, utils/visual_utils.py:55-57; clamped to [0, 255])
 *   t2o_img2tensor     img2tensor (utils/visual_utils.py:61-70): (N, H, W, 3) BGR uint8 -> (N, 3, H, W) RGB float32 / 255
 *   t2o_tensor2img     tensor2img (utils/visual_utils.py:50-58): (N, 3, H, W) RGB float32 -> (N, H, W, 3) BGR uint8
 * The flat pair needs 16-byte aligned pointers; the float side of the layout pair too.
 */
int t2o_u8_to_f32(const uint8_t *src, float *dst, int64_t n, t2o_stream_t stream);
int t2o_f32_to_u8(const float *src, uint8_t *dst, int64_t n, t2o_stream_t stream);
int t2o_img2tensor(const uint8_t *hwc_bgr, float *chw_rgb, int N, int H, int W, t2o_stream_t stream);
int t2o_tensor2img(const float *chw_rgb, uint8_t *hwc_bgr, int N, int H, int W, t2o_stream_t stream);

/*
 * Planner candidate scoring (the inner loop of get_param_naive / beam_search,
 * utils/beam_search.py:77-87,229-237): candidate c applies operator cand_op[c] with parameters
 * cand_param[c*24 ..] to state image cand_state[c] and is scored against that state's target,
 *     l1_sum[c] = sum |clamp(op(state; param)) - target|      (no image is written).
 * Candidates MUST be sorted by cand_state (ascending); cand_begin[s] .. cand_begin[s+1] are the
 * candidates of state s (S + 1 ints).  state_target[s] selects the target image of state s
 * (NULL: target s % T).  Each (state, target) tile is staged once in shared memory by TMA and
 * every candidate of that state is evaluated from there.
 */
int t2o_score_candidates(const float *states, int S, const float *targets, int T,
                         const int32_t *state_target, const int32_t *cand_begin,
                         const int32_t *cand_op, const float *cand_param, int C,
                         float *l1_sum, int H, int W, int curve_steps,
                         void *workspace, size_t workspace_bytes, t2o_stream_t stream);

/*
 * The same with masks, as the GIER planner driver hands them over (preprocess/gen_greedy_seqs_GIER.py:60-62: a list of
 * (1, 3, H, W) masks, the global all-ones one first, and the operator each local mask belongs to): candidate c edits
 * its state inside mask cand_mask[c] of `masks` (n_masks, mask_ch, H, W; mask_ch 1 or 3; fp32 in [0, 1]),
 *     l1_sum[c] = sum |clamp(op(state; param) * mask + state * (1 - mask)) - target|    (models/operators.py:129-130),
 * or everywhere if cand_mask[c] < 0.
 */
int t2o_score_candidates_masked(const float *states, int S, const float *targets, int T,
                                const int32_t *state_target, const int32_t *cand_begin,
                                const int32_t *cand_op, const float *cand_param,
                                const int32_t *cand_mask, const float *masks, int n_masks, int mask_ch, int C,
                                float *l1_sum, int H, int W, int curve_steps,
                                void *workspace, size_t workspace_bytes, t2o_stream_t stream);

/*
 * The beam selection of a planner step for many searches at once (utils/beam_search.py:252-256: np.argsort of the candidates'
 * distances, the first beam_size kept): values[seg_begin[s] .. seg_begin[s+1]) are the distances of search s
 * (n_seg + 1 ints); out_idx / out_val (n_seg, k) receive the k smallest in ascending order -- global indices into `values`,
 * exact ties by the smaller index, NaN last; -1 / +inf where a segment holds fewer than k values.
 */
int t2o_topk_min(const float *values, const int32_t *seg_begin, int n_seg, int k,
                 int32_t *out_idx, float *out_val, t2o_stream_t stream);

/*
 * Device-resident Nelder-Mead: P independent fits argmin_param L1(op(state; param), target), one per (state, operator)
 * pair of a planner step -- scipy.optimize.minimize(func, param0, method='Nelder-Mead') as utils/beam_search.py:88
 * calls it (scipy defaults: xatol = fatol = 1e-4, maxiter = maxfev = 200 N, initial simplex step 5 % / 2.5e-4),
 * with float64 simplex arithmetic in scipy's operation order.  A fit is a sequential chain of evaluations; an
 * evaluation is candidate p of t2o_score_candidates (fits and candidates share the index, sorted by state):
 *     t2o_nm_start(...)                                   builds the simplices, writes the first vertices
 *     repeat: t2o_score_candidates(...cand_op, cand_param... -> l1_sum);  t2o_nm_advance(... l1_sum ...)
 * Nothing synchronises with the host: vertices (cand_param, float32 rows of 24) and scores (l1_sum) stay in device
 * memory, so rounds can be enqueued back to back or captured in a CUDA graph.  A finished fit sets its
 * cand_op[p] = T2O_OP_SKIP (the scorer then ignores it), ctl[p][1] = 6 and leaves its result in xbest / fbest.
 * All state is caller-allocated device memory; no initialisation is required before t2o_nm_start.
 */
typedef struct {
    double *sim;      /* (P, 25, 24) simplex vertices */
    double *fsim;     /* (P, 25)     function values by sorted position */
    double *vec;      /* (P, 3, 24)  centroid, reflected point, pending point */
    double *fxr;      /* (P,)        value of the reflected point */
    double *xbest;    /* (P, 24)     result: best vertex (valid once the fit is done) */
    double *fbest;    /* (P,)        result: its function value */
    int32_t *perm;    /* (P, 25)     simplex row of each sorted position */
    int32_t *ctl;     /* (P, 8)      n, phase (6 = done), k, function calls, iterations, scipy status (0 ok, 1 maxfev, 2 maxiter), op, 0 */
} t2o_nm_state;

/* n_dims (P,) int32: parameters of each fit (1..24); prob_op (P,) int32: its operator id; x0 (P, 24) float64 start
 * points (utils/beam_search.py:150-153: zeros, ones for the curve operators). */
int t2o_nm_start(const t2o_nm_state *state /*host struct of device pointers*/, int P,
                 const int32_t *n_dims, const int32_t *prob_op, const double *x0,
                 float *cand_param, int32_t *cand_op, t2o_stream_t stream);
/* l1_sum (P,): the scorer's output for the pending vertices; numel = 3*H*W*batch of get_dist's mean
 * (utils/beam_search.py:172): the function value is float32(l1_sum * float32(1 / numel)) -- torch's CUDA division by a
 * host scalar -- widened to float64, as .item() gives. */
int t2o_nm_advance(const t2o_nm_state *state /*host struct of device pointers*/, int P,
                   const float *l1_sum, float numel, float *cand_param, int32_t *cand_op, t2o_stream_t stream);

/*
 * Every fit of a planner step in ONE launch (replaces the thousands of get_dist calls scipy makes per fit,
 * utils/beam_search.py:65-91): after t2o_nm_start, a cluster of CTAs (one per tile of the image) keeps a state and its target
 * in shared memory for the whole life of the state's fits, every CTA keeps (and steps) the fits' simplices there too, and
 * the cluster iterates score -> advance, exchanging only the tiles' partial sums -- the same arithmetic as rounds of t2o_score_candidates + t2o_nm_advance, bit for
 * bit, without re-staging the images or touching device memory between evaluations.
 * fits_begin[s] .. fits_begin[s+1] are the fits of state s (S + 1 ints, at most 8 fits per state; fits sorted by state as for
 * t2o_score_candidates); fit_mask / masks as in t2o_score_candidates_masked (NULL: none); host_fits_begin / host_fit_op are
 * HOST copies of fits_begin and of the fits' operators (they size the launch's shared memory); max_rounds bounds the
 * evaluations per fit (200 * 24 + 8 covers scipy's maxfev); fits it leaves unfinished can go on, here or in rounds.
 * Returns T2O_ERR_UNSUPPORTED where the shape is not eligible (more than 16 tiles of 32 x 128 pixels per image, more than 8
 * fits per state, W % 4 != 0, no TMA, fits too large for the shared memory): run the rounds instead.  workspace: >= 256 KiB.
 */
int t2o_nm_run_resident(const float *states, int S, const float *targets, int T, const int32_t *state_target,
                        const int32_t *fits_begin, const int32_t *fit_mask, const float *masks, int n_masks, int mask_ch,
                        const t2o_nm_state *state /*host struct of device pointers*/, int P, float numel,
                        float *cand_param, int32_t *cand_op, const int32_t *host_fits_begin, const int32_t *host_fit_op,
                        int H, int W, int curve_steps, int max_rounds, void *workspace, size_t workspace_bytes, t2o_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* T2O_H_ */
