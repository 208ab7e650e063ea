// t2o_cabi.cu -- the extern "C" boundary declared in include/t2o.h.  Plain pointers and sizes
// only; every entry point validates its arguments and returns a t2o_status.
#include <cuda_runtime.h>

#include "../../include/t2o.h"
#include "t2o_math.cuh"

namespace t2o {
size_t chain_workspace_bytes(int B, int H, int W, int pstride);
int chain_forward(int n_ops, const int *op_ids, const int *param_off, const float *img, const float *mask, int mask_ch,
                  const float *params, int pstride, const float *target, float *out, float *l1_sum,
                  int B, int H, int W, int L, int flags, void *ws, size_t ws_bytes, cudaStream_t stream);
int chain_backward(int n_ops, const int *op_ids, const int *param_off, const float *img, const float *mask, int mask_ch,
                   const float *params, int pstride, const float *grad_out, const float *target, const float *grad_l1,
                   float *grad_params, float *grad_img, float *out, float *l1_sum,
                   int B, int H, int W, int L, void *ws, size_t ws_bytes, cudaStream_t stream);
int rows_forward(int K, const int *row_ops, const int *row_ops_host, int slot, const float *img, const float *mask, int mask_ch,
                 const float *params, int pstride, const float *target, float *out, float *l1_sum, unsigned int *status,
                 int B, int H, int W, int L, void *ws, size_t ws_bytes, cudaStream_t stream);
int rows_backward(int K, const int *row_ops, const int *row_ops_host, int slot, const float *img, const float *mask, int mask_ch,
                  const float *params, int pstride, const float *grad_out, const float *target, const float *grad_l1,
                  float *grad_params, float *grad_img, float *out, float *l1_sum, unsigned int *status,
                  int B, int H, int W, int L, void *ws, size_t ws_bytes, cudaStream_t stream);
int l1_sum_launch(const float *pa, const float *pb, float *l1_sum, int B, long long n, void *ws, size_t ws_bytes,
                  cudaStream_t stream);
size_t score_workspace_bytes(int S, int C, int H, int W);
int score_candidates(const float *states, int S, const float *targets, int T, const int *state_target,
                     const int *cand_begin, const int *cand_op, const float *cand_param, const int *cand_mask,
                     const float *masks, int n_masks, int mask_ch, int C, float *l1_sum,
                     int H, int W, int L, void *ws, size_t ws_bytes, cudaStream_t stream);
int nm_run_resident(const float *states, int S, const float *targets, int T, const int *state_target, const int *fits_begin,
                    const int *cand_mask, const float *masks, int n_masks, int mask_ch,
                    const t2o_nm_state *st, int P, float numel, float *cand_param, int *cand_op,
                    const int *h_fits_begin, const int *h_fit_op,
                    int H, int W, int L, int max_rounds, void *ws, size_t ws_bytes, cudaStream_t stream);
int topk_min(const float *values, const int *seg_begin, int n_seg, int k, int *out_idx, float *out_val, cudaStream_t stream);
int nm_start(const t2o_nm_state *st, int P, const int *n_dims, const int *prob_op, const double *x0,
             float *cand_param, int *cand_op, cudaStream_t stream);
int nm_advance(const t2o_nm_state *st, int P, const float *l1_sum, float numel, float *cand_param, int *cand_op,
               cudaStream_t stream);
size_t ssim_workspace_bytes(int B, int C, int H, int W);
int ssim_sum(const float *img1, const float *img2, float *out, int B, int C, int H, int W, void *ws, size_t ws_bytes,
             cudaStream_t stream);
int convert_u8_to_f32(const uint8_t *src, float *dst, long long n, cudaStream_t stream);
int convert_f32_to_u8(const float *src, uint8_t *dst, long long n, cudaStream_t stream);
int convert_img2tensor(const uint8_t *hwc_bgr, float *chw_rgb, int N, int H, int W, cudaStream_t stream);
int convert_tensor2img(const float *chw_rgb, uint8_t *hwc_bgr, int N, int H, int W, cudaStream_t stream);
const char *last_cuda_error();
}  // namespace t2o

extern "C" {

int t2o_version(void) { return T2O_VERSION; }

const char *t2o_status_string(int status) {
    switch (status) {
        case T2O_OK: return "ok";
        case T2O_ERR_INVALID_ARG: return "invalid argument";
        case T2O_ERR_UNSUPPORTED: return "unsupported configuration";
        case T2O_ERR_WORKSPACE: return "workspace missing or too small";
        case T2O_ERR_CUDA: return "CUDA runtime error";
        case T2O_ERR_NO_DEVICE: return "no usable device / driver entry point";
        default: return "unknown status";
    }
}

const char *t2o_last_cuda_error(void) { return t2o::last_cuda_error(); }

int t2o_num_params(int op_id, int curve_steps) {
    if (op_id < t2o::OP_IDENTITY || op_id >= t2o::OP_COUNT) return -1;
    return t2o::op_num_params(op_id, curve_steps);
}

size_t t2o_workspace_bytes(int B, int H, int W, int param_stride) {
    if (B < 1 || H < 1 || W < 1) return 0;
    return t2o::chain_workspace_bytes(B, H, W, param_stride);
}

size_t t2o_score_workspace_bytes(int S, int C, int H, int W) {
    if (H < 1 || W < 1) return 0;
    return t2o::score_workspace_bytes(S, C, H, W);
}

int t2o_chain_forward(int n_ops, const int *op_ids, const int *param_off, const float *img, const float *mask, int mask_ch,
                      const float *params, int param_stride, const float *target, float *out, float *l1_sum,
                      int B, int H, int W, int curve_steps, int flags, void *workspace, size_t workspace_bytes,
                      t2o_stream_t stream) {
    return t2o::chain_forward(n_ops, op_ids, param_off, img, mask, mask_ch, params, param_stride, target, out, l1_sum,
                              B, H, W, curve_steps, flags, workspace, workspace_bytes, (cudaStream_t)stream);
}

int t2o_chain_backward(int n_ops, const int *op_ids, const int *param_off, const float *img, const float *mask, int mask_ch,
                       const float *params, int param_stride, const float *grad_out, const float *target,
                       const float *grad_l1, float *grad_params, float *grad_img, float *out, float *l1_sum,
                       int B, int H, int W, int curve_steps, void *workspace, size_t workspace_bytes, t2o_stream_t stream) {
    return t2o::chain_backward(n_ops, op_ids, param_off, img, mask, mask_ch, params, param_stride, grad_out, target, grad_l1,
                               grad_params, grad_img, out, l1_sum, B, H, W, curve_steps, workspace, workspace_bytes,
                               (cudaStream_t)stream);
}

int t2o_rows_forward(int K, const int32_t *row_ops, const int32_t *row_ops_host, int param_slot, const float *img,
                     const float *mask, int mask_ch, const float *params, int param_stride, const float *target, float *out,
                     float *l1_sum, uint32_t *status, int B, int H, int W, int curve_steps, void *workspace,
                     size_t workspace_bytes, t2o_stream_t stream) {
    return t2o::rows_forward(K, row_ops, row_ops_host, param_slot, img, mask, mask_ch, params, param_stride, target, out,
                             l1_sum, status, B, H, W, curve_steps, workspace, workspace_bytes, (cudaStream_t)stream);
}

int t2o_rows_backward(int K, const int32_t *row_ops, const int32_t *row_ops_host, int param_slot, const float *img,
                      const float *mask, int mask_ch, const float *params, int param_stride, const float *grad_out,
                      const float *target, const float *grad_l1, float *grad_params, float *grad_img, float *out,
                      float *l1_sum, uint32_t *status, int B, int H, int W, int curve_steps, void *workspace,
                      size_t workspace_bytes, t2o_stream_t stream) {
    return t2o::rows_backward(K, row_ops, row_ops_host, param_slot, img, mask, mask_ch, params, param_stride, grad_out,
                              target, grad_l1, grad_params, grad_img, out, l1_sum, status, B, H, W, curve_steps, workspace,
                              workspace_bytes, (cudaStream_t)stream);
}

int t2o_l1_sum(const float *a, const float *b, float *l1_sum, int B, int64_t n_per_image, void *workspace,
               size_t workspace_bytes, t2o_stream_t stream) {
    return t2o::l1_sum_launch(a, b, l1_sum, B, (long long)n_per_image, workspace, workspace_bytes, (cudaStream_t)stream);
}

int t2o_score_candidates(const float *states, int S, const float *targets, int T, const int32_t *state_target,
                         const int32_t *cand_begin, const int32_t *cand_op, const float *cand_param, int C, float *l1_sum,
                         int H, int W, int curve_steps, void *workspace, size_t workspace_bytes, t2o_stream_t stream) {
    return t2o::score_candidates(states, S, targets, T, state_target, cand_begin, cand_op, cand_param, nullptr, nullptr, 0, 0, C,
                                 l1_sum, H, W, curve_steps, workspace, workspace_bytes, (cudaStream_t)stream);
}

int t2o_score_candidates_masked(const float *states, int S, const float *targets, int T, const int32_t *state_target,
                                const int32_t *cand_begin, const int32_t *cand_op, const float *cand_param,
                                const int32_t *cand_mask, const float *masks, int n_masks, int mask_ch, int C, float *l1_sum,
                                int H, int W, int curve_steps, void *workspace, size_t workspace_bytes, t2o_stream_t stream) {
    return t2o::score_candidates(states, S, targets, T, state_target, cand_begin, cand_op, cand_param, cand_mask, masks, n_masks,
                                 mask_ch, C, l1_sum, H, W, curve_steps, workspace, workspace_bytes, (cudaStream_t)stream);
}

int t2o_nm_run_resident(const float *states, int S, const float *targets, int T, const int32_t *state_target,
                        const int32_t *fits_begin, const int32_t *fit_mask, const float *masks, int n_masks, int mask_ch,
                        const t2o_nm_state *state, int P, float numel, float *cand_param, int32_t *cand_op,
                        const int32_t *host_fits_begin, const int32_t *host_fit_op,
                        int H, int W, int curve_steps, int max_rounds, void *workspace, size_t workspace_bytes, t2o_stream_t stream) {
    return t2o::nm_run_resident(states, S, targets, T, state_target, fits_begin, fit_mask, masks, n_masks, mask_ch, state, P, numel,
                                cand_param, cand_op, host_fits_begin, host_fit_op, H, W, curve_steps, max_rounds, workspace,
                                workspace_bytes, (cudaStream_t)stream);
}

int t2o_topk_min(const float *values, const int32_t *seg_begin, int n_seg, int k, int32_t *out_idx, float *out_val,
                 t2o_stream_t stream) {
    return t2o::topk_min(values, seg_begin, n_seg, k, out_idx, out_val, (cudaStream_t)stream);
}

int t2o_nm_start(const t2o_nm_state *state, int P, const int32_t *n_dims, const int32_t *prob_op, const double *x0,
                 float *cand_param, int32_t *cand_op, t2o_stream_t stream) {
    return t2o::nm_start(state, P, n_dims, prob_op, x0, cand_param, cand_op, (cudaStream_t)stream);
}

int t2o_nm_advance(const t2o_nm_state *state, int P, const float *l1_sum, float numel, float *cand_param,
                   int32_t *cand_op, t2o_stream_t stream) {
    return t2o::nm_advance(state, P, l1_sum, numel, cand_param, cand_op, (cudaStream_t)stream);
}

size_t t2o_ssim_workspace_bytes(int B, int C, int H, int W) {
    if (B < 1 || C < 1 || H < 1 || W < 1) return 0;
    return t2o::ssim_workspace_bytes(B, C, H, W);
}

int t2o_ssim_sum(const float *img1, const float *img2, float *ssim_sum, int B, int C, int H, int W, void *workspace,
                 size_t workspace_bytes, t2o_stream_t stream) {
    return t2o::ssim_sum(img1, img2, ssim_sum, B, C, H, W, workspace, workspace_bytes, (cudaStream_t)stream);
}

int t2o_u8_to_f32(const uint8_t *src, float *dst, int64_t n, t2o_stream_t stream) {
    return t2o::convert_u8_to_f32(src, dst, (long long)n, (cudaStream_t)stream);
}

int t2o_f32_to_u8(const float *src, uint8_t *dst, int64_t n, t2o_stream_t stream) {
    return t2o::convert_f32_to_u8(src, dst, (long long)n, (cudaStream_t)stream);
}

int t2o_img2tensor(const uint8_t *hwc_bgr, float *chw_rgb, int N, int H, int W, t2o_stream_t stream) {
    return t2o::convert_img2tensor(hwc_bgr, chw_rgb, N, H, W, (cudaStream_t)stream);
}

int t2o_tensor2img(const float *chw_rgb, uint8_t *hwc_bgr, int N, int H, int W, t2o_stream_t stream) {
    return t2o::convert_tensor2img(chw_rgb, hwc_bgr, N, H, W, (cudaStream_t)stream);
}

}  // extern "C"
