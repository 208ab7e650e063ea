// t2o_step.cu -- host side of t2o_chain_backward: geometry, workspace carving, launch selection for the fused
// forward + L1 + backward kernels of t2o_step_kernels.cuh.
#include <cstdlib>

#include "t2o_step_kernels.cuh"

namespace t2o {

size_t chain_workspace_bytes(int B, int H, int W, int pstride);
struct Workspace {
    unsigned int *counters;
    float *part_l1, *part_gp;
};
Workspace carve_workspace(void *ws, int B, int H, int W);
size_t max_tiles(int H, int W);

static int make_step_desc(int n_ops, const int *op_ids, const int *param_off, int L, int pstride, StepDesc &d) {
    if (!param_off) return T2O_ERR_INVALID_ARG;
    return build_step_desc(n_ops, op_ids, param_off, 0, L, pstride, d);
}

static int pick_vec(const void *const *ptrs, int nptr, size_t plane) {
    int vec = (plane % 4 == 0) ? 4 : (plane % 2 == 0 ? 2 : 1);
    for (int i = 0; i < nptr; ++i) {
        if (!ptrs[i]) continue;
        const uintptr_t a = (uintptr_t)ptrs[i];
        while (vec > 1 && (a % (vec * 4)) != 0) vec >>= 1;
    }
    return vec;
}

// resident CTAs per SM of a kernel at a given dynamic shared memory size (cached by the caller's static)
template <typename K>
static int resident_ctas(K kernel, int nth, size_t smem) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, nth, smem) != cudaSuccess || nb < 1) nb = 1;
    return nb;
}

static void geom_step_flat(StepGeom &g, int B, int H, int W, int vec, int slots, int SNT) {
    memset(&g, 0, sizeof(g));
    g.B = B; g.H = H; g.W = W;
    const long long plane = (long long)H * W;
    g.ngroups = plane / vec;
    const long long total = g.ngroups * B;
    // one wave of CTAs unless a thread would then own more than ~32 groups
    long long waves = total / ((long long)slots * SNT * 32);
    if (waves < 1) waves = 1;
    if (waves > 16) waves = 16;
    long long cg = (total + slots * waves - 1) / (slots * waves);
    const long long quantum = SNT;
    cg = (cg + quantum - 1) / quantum * quantum;
    const long long min_groups = (1024 + vec - 1) / vec;       // a chunk holds >= 1024 pixels (bounds the workspace)
    if (cg < min_groups) cg = (min_groups + quantum - 1) / quantum * quantum;
    g.chunk_groups = (int)cg;
    g.nchunks = (int)((g.ngroups + cg - 1) / cg);
    if (g.nchunks < 1) g.nchunks = 1;
}

// rows_extra: halo rows a band computes beyond its own (4 for the forward + backward pipeline, 2 for the forward one)
void geom_step_rows(StepGeom &g, int B, int H, int W, int vec, int slots, int SNT, int rows_extra) {
    const int SNW = SNT / 32;
    memset(&g, 0, sizeof(g));
    g.B = B; g.H = H; g.W = W;
    g.Wg = W / vec;
    if (g.Wg <= 32) {
        g.HL = 0; g.IW = g.Wg; g.strips = 1;
    } else {
        g.HL = vec >= 2 ? 1 : 2; g.IW = 32 - 2 * g.HL;
        g.strips = (g.Wg + g.IW - 1) / g.IW;
    }
    // band height: minimise waves x (steps + 1) over the number of bands; a band holds >= 16 rows
    long long best = -1;
    int best_hb = H;
    const int max_nb = H / 16 > 1 ? H / 16 : 1;
    for (int nb = 1; nb <= max_nb; ++nb) {
        const int hb = (H + nb - 1) / nb;
        const int bands = (H + hb - 1) / hb;
        const int steps = (hb + rows_extra + SNW - 1) / SNW;
        const long long ctas = (long long)B * g.strips * bands;
        const long long waves = (ctas + slots - 1) / slots;
        const long long cost = waves * (steps + 1);
        if (best < 0 || cost < best) { best = cost; best_hb = hb; }
    }
    g.HB = best_hb;
    g.bands = (H + g.HB - 1) / g.HB;
    g.steps = (g.HB + rows_extra + SNW - 1) / SNW;
    g.nchunks = g.strips * g.bands;
}

static size_t rows_smem_bytes(int n, int sharp, int vec, bool has_mask, int SNT) {
    const int RING = SNT / 32 + 2;
    const size_t ringf = (size_t)RING * 3 * 34 * vec;
    const size_t ntp = sharp > 1 ? sharp - 1 : 0;
    const size_t ntq = n - sharp - 1;
    const size_t stage = (size_t)3 * SNT * vec;            // per-thread staging slots of the asynchronous row copies
    return (stage + (has_mask ? 3 : 2) * ringf + ntp * RING * 3 * 32 * vec + ntq * 3 * SNT * vec) * sizeof(float);
}

template <int VEC, bool HM, int NTH, unsigned int SP, int MINB = 2>
static int launch_step_sp(StepArgs &a, cudaStream_t stream) {
    const bool rows = a.ch.sharp >= 0;
    const int B = a.g.B, H = a.g.H, W = a.g.W;
    size_t smem;
    if (!rows) {
        if constexpr (SP == 0u || sp_sharp(SP) < 0) {
            smem = (size_t)(a.ch.n + 2) * 3 * NTH * VEC * sizeof(float);        // tape + two staging slot sets
            int st = step_set_smem(step_flat_kernel<VEC, HM, NTH, false, SP>, smem);
            if (st) return st;
            geom_step_flat(a.g, B, H, W, VEC, NUM_SMS * resident_ctas(step_flat_kernel<VEC, HM, NTH, false, SP>, NTH, smem), NTH);
            dim3 grid(a.g.nchunks, B);
            step_flat_kernel<VEC, HM, NTH, false, SP><<<grid, NTH, smem, stream>>>(a);
        }
    } else {
        if constexpr (SP == 0u || sp_sharp(SP) >= 0) {
            smem = rows_smem_bytes(a.ch.n, a.ch.sharp, VEC, HM, NTH);
            int st = step_set_smem(step_sharp_kernel<VEC, HM, NTH, false, SP, MINB>, smem);
            if (st) return st;
            geom_step_rows(a.g, B, H, W, VEC, NUM_SMS * resident_ctas(step_sharp_kernel<VEC, HM, NTH, false, SP, MINB>, NTH, smem), NTH, 4);
            dim3 grid(a.g.nchunks, B);
            step_sharp_kernel<VEC, HM, NTH, false, SP, MINB><<<grid, NTH, smem, stream>>>(a);
        }
    }
    T2O_CUDA_OK(cudaGetLastError());
    return T2O_OK;
}

// T2O_NO_SPECIALIZED=1 in the environment forces the run-time dispatched kernels; read at every launch so that
// the parity tests can compare both paths in one process.
static bool use_specialized() {
    const char *e = getenv("T2O_NO_SPECIALIZED");
    return !(e && e[0] == '1');
}

template <int VEC, bool HM, int NTH>
static int launch_step(StepArgs &a, cudaStream_t stream) {
    if constexpr (VEC == 4 && !HM) {       // chain-specialised instantiations (unmasked, 128-bit groups)
        if (use_specialized()) {
            if (a.ch.ops_packed == SP_C6) return launch_step_sp<VEC, HM, NTH, SP_C6>(a, stream);
            if (a.ch.ops_packed == SP_P5) return launch_step_sp<VEC, HM, NTH, SP_P5>(a, stream);
            if (a.ch.ops_packed == SP_S1) return launch_step_sp<VEC, HM, NTH, SP_S1, 4>(a, stream);      // light kernels: 4 CTAs per SM
            if (a.ch.ops_packed == SP_B1) return launch_step_sp<VEC, HM, NTH, SP_B1, 3>(a, stream);
        }
    }
    return launch_step_sp<VEC, HM, NTH, 0u>(a, stream);
}

// Per-row chains: one launch of each tiling; a CTA whose row belongs to the other tiling exits at once.
// `paths` bit 0: some row has no stencil, bit 1: some row has one (3 when the host does not know the rows).
template <int VEC, bool HM, int NTH>
static int launch_step_rows(StepArgs &a, int paths, cudaStream_t stream) {
    const int B = a.g.B, H = a.g.H, W = a.g.W, K = a.rows_K;
    if (paths & 1) {
        const size_t smem = (size_t)(K + 2) * 3 * NTH * VEC * sizeof(float);        // tape + two staging slot sets
        int st = step_set_smem(step_flat_kernel<VEC, HM, NTH, true>, smem);
        if (st) return st;
        geom_step_flat(a.g, B, H, W, VEC, NUM_SMS * resident_ctas(step_flat_kernel<VEC, HM, NTH, true>, NTH, smem), NTH);
        dim3 grid(a.g.nchunks, B);
        step_flat_kernel<VEC, HM, NTH, true><<<grid, NTH, smem, stream>>>(a);
        T2O_CUDA_OK(cudaGetLastError());
    }
    if (paths & 2) {
        // shared memory for the worst row: the stencil last (K - 1 operators on the ring tape) or first (K - 1 per-thread tapes)
        const size_t s_last = rows_smem_bytes(K, K - 1, VEC, HM, NTH), s_first = rows_smem_bytes(K, 0, VEC, HM, NTH);
        const size_t smem = s_last > s_first ? s_last : s_first;
        int st = step_set_smem(step_sharp_kernel<VEC, HM, NTH, true>, smem);
        if (st) return st;
        geom_step_rows(a.g, B, H, W, VEC, NUM_SMS * resident_ctas(step_sharp_kernel<VEC, HM, NTH, true>, NTH, smem), NTH, 4);
        dim3 grid(a.g.nchunks, B);
        step_sharp_kernel<VEC, HM, NTH, true><<<grid, NTH, smem, stream>>>(a);
        T2O_CUDA_OK(cudaGetLastError());
    }
    return T2O_OK;
}

template <bool HM>
static int launch_step_vec(int vec, StepArgs &a, int rows_paths, cudaStream_t stream) {
    if (rows_paths) {
        if (vec == 4) return launch_step_rows<4, HM, 256>(a, rows_paths, stream);
        if (vec == 2) return launch_step_rows<2, HM, 256>(a, rows_paths, stream);
        return launch_step_rows<1, HM, 256>(a, rows_paths, stream);
    }
    if (vec == 4) return launch_step<4, HM, 256>(a, stream);
    if (vec == 2) return launch_step<2, HM, 256>(a, stream);
    return launch_step<1, HM, 256>(a, stream);
}

int chain_backward(int n_ops, const int *op_ids, const int *param_off, const float *img, const float *mask, int mask_ch,
                   const float *params, int pstride, const float *grad_out, const float *target, const float *grad_l1,
                   float *grad_params, float *grad_img, float *out, float *l1_sum,
                   int B, int H, int W, int L, void *ws, size_t ws_bytes, cudaStream_t stream) {
    StepArgs a;
    memset(&a, 0, sizeof(a));
    int st = make_step_desc(n_ops, op_ids, param_off, L, pstride, a.ch);
    if (st != T2O_OK) return st;
    if (!img || B < 1 || H < 1 || W < 1 || (pstride > 0 && !params)) return T2O_ERR_INVALID_ARG;
    if (mask && mask_ch != 1 && mask_ch != 3) return T2O_ERR_INVALID_ARG;
    if (!grad_out && (!target || !grad_l1)) return T2O_ERR_INVALID_ARG;
    if (l1_sum && !target) return T2O_ERR_INVALID_ARG;
    if (!grad_params && !grad_img) return T2O_ERR_INVALID_ARG;
    if (!ws || ws_bytes < chain_workspace_bytes(B, H, W, pstride)) return T2O_ERR_WORKSPACE;
    if (B > 65535) return T2O_ERR_UNSUPPORTED;
    Workspace w = carve_workspace(ws, B, H, W);
    a.img = img; a.mask = mask; a.params = params; a.grad_out = grad_out; a.target = target; a.grad_l1 = grad_l1;
    a.grad_params = grad_params; a.grad_img = grad_img; a.out = out; a.l1_sum = l1_sum;
    a.part_l1 = w.part_l1; a.part_gp = w.part_gp; a.counters = w.counters; a.mask_ch = mask_ch; a.pstride = pstride;
    a.g.B = B; a.g.H = H; a.g.W = W;
    const size_t plane = (size_t)H * W;
    const void *ptrs[] = {img, mask, target, out, grad_out, grad_img};
    int vec = pick_vec(ptrs, 6, plane);
    if (a.ch.sharp >= 0)
        while (vec > 1 && W % vec != 0) vec >>= 1;
    return mask ? launch_step_vec<true>(vec, a, 0, stream) : launch_step_vec<false>(vec, a, 0, stream);
}

int rows_paths(int K, const int *row_ops_host, int B, int slot, int L, int pstride, bool backward, int &paths);

// Per-row chains, backward (+ optional fused forward outputs): autograd through the Actor's per-row operator step.
int rows_backward(int K, const int *row_ops, const int *row_ops_host, int slot, const float *img, const float *mask, int mask_ch,
                  const float *params, int pstride, const float *grad_out, const float *target, const float *grad_l1,
                  float *grad_params, float *grad_img, float *out, float *l1_sum, unsigned int *status,
                  int B, int H, int W, int L, void *ws, size_t ws_bytes, cudaStream_t stream) {
    if (!row_ops || !img || !params || B < 1 || H < 1 || W < 1) return T2O_ERR_INVALID_ARG;
    if (mask && mask_ch != 1 && mask_ch != 3) return T2O_ERR_INVALID_ARG;
    if (!grad_out && (!target || !grad_l1)) return T2O_ERR_INVALID_ARG;
    if (l1_sum && !target) return T2O_ERR_INVALID_ARG;
    if (!grad_params && !grad_img) return T2O_ERR_INVALID_ARG;
    if (!ws || ws_bytes < chain_workspace_bytes(B, H, W, pstride)) return T2O_ERR_WORKSPACE;
    if (B > 65535) return T2O_ERR_UNSUPPORTED;
    int paths = 0;
    int st = rows_paths(K, row_ops_host, B, slot, L, pstride, true, paths);
    if (st != T2O_OK) return st;
    StepArgs a;
    memset(&a, 0, sizeof(a));
    a.ch.n = K; a.ch.L = L; a.ch.sharp = -1;
    Workspace w = carve_workspace(ws, B, H, W);
    a.img = img; a.mask = mask; a.params = params; a.grad_out = grad_out; a.target = target; a.grad_l1 = grad_l1;
    a.grad_params = grad_params; a.grad_img = grad_img; a.out = out; a.l1_sum = l1_sum;
    a.part_l1 = w.part_l1; a.part_gp = w.part_gp; a.counters = w.counters; a.mask_ch = mask_ch; a.pstride = pstride;
    a.row_ops = row_ops; a.rows_K = K; a.rows_slot = slot; a.status = status;
    a.g.B = B; a.g.H = H; a.g.W = W;
    const size_t plane = (size_t)H * W;
    const void *ptrs[] = {img, mask, target, out, grad_out, grad_img};
    int vec = pick_vec(ptrs, 6, plane);
    if (paths & 2)
        while (vec > 1 && W % vec != 0) vec >>= 1;
    return mask ? launch_step_vec<true>(vec, a, paths, stream) : launch_step_vec<false>(vec, a, paths, stream);
}

}  // namespace t2o
