// t2o_chain_kernels.cuh -- fused operator-chain kernels for sm_100a (templates; instantiated by t2o_chain*.cu).
//
//   chain_fwd_kernel   K x Operator.execute (+ per-image L1 to a target) in one pass over HBM
//   chain_bwd_kernel   the same chain recomputed and differentiated in one pass: parameter
//                      gradients (+ optional image gradient, + optional forward outputs)
//
// Replaces: K successive Executor.execute calls (executors/executor.py:33-55 ->
// models/operators.py:112-131), get_dist 'L1' (utils/beam_search.py:170-173) and autograd through
// them.  Elementwise and HBM-bound by bytes, FP32-issue-bound in practice: no tensor cores by design.
//
// Data layout: images NCHW planar fp32; a thread owns VEC consecutive pixels of the three planes
// (128-bit coalesced accesses for VEC = 4).  Per-(image, op) tables (curve slopes / offsets,
// scalars) live in shared memory.  A sharpness operator turns the launch into a 2-D tiling with a
// halo: the operators before it are evaluated on tile + halo into shared memory, the 3x3 stencil
// and the operators after it run from there, so the chain still makes one pass over HBM.
// Reductions (L1, parameter gradients) are warp-shuffle -> shared -> one partial per CTA; the last
// CTA of each image sums the partials in a fixed order (deterministic, no float atomics).
#pragma once
#include <cstdio>
#include <cstring>

#include "t2o_common.cuh"
#include "../../include/t2o.h"

namespace t2o {

// ---------------------------------------------------------------- operator dispatch over pixel groups
// One switch per G x VEC pixels: the dispatch (uniform branch on the op id) is amortised over them.
template <int VEC, int G, bool HM>
__device__ __forceinline__ void apply_op_grp(int op, const float *tab, int L, float (&x)[G][3][VEC],
                                             const float (&m)[G][3][VEC], bool raw = false) {
#define T2O_CASE(OPC)                                                                                          \
    case OPC:                                                                                                  \
        _Pragma("unroll") for (int q = 0; q < G; ++q)                                                          \
            _Pragma("unroll") for (int v = 0; v < VEC; ++v)                                                    \
                op_apply<HM>(OPC, tab, L, x[q][0][v], x[q][1][v], x[q][2][v], m[q][0][v], m[q][1][v], m[q][2][v], raw); \
        break;
    switch (op) {
        T2O_CASE(OP_BRIGHTNESS) T2O_CASE(OP_CONTRAST) T2O_CASE(OP_SATURATION) T2O_CASE(OP_COLOR)
        T2O_CASE(OP_TONE) T2O_CASE(OP_WHITE) T2O_CASE(OP_EXPOSURE) T2O_CASE(OP_WHITEBALANCE)
        default: break;
    }
#undef T2O_CASE
}
template <int VEC, bool HM>
__device__ __forceinline__ void apply_op_vec(int op, const float *tab, int L, float (&x)[3][VEC],
                                             const float (&m)[3][VEC], bool raw = false) {
    apply_op_grp<VEC, 1, HM>(op, tab, L, reinterpret_cast<float(&)[1][3][VEC]>(x),
                             reinterpret_cast<const float(&)[1][3][VEC]>(m), raw);
}

// out = op(in): the backward kernels chain sv[k] -> sv[k+1] so every operator input stays in registers
template <int VEC, bool HM>
__device__ __forceinline__ void apply_io_vec(int op, const float *tab, int L, const float (&in)[3][VEC],
                                             float (&out)[3][VEC], const float (&m)[3][VEC]) {
#define T2O_CASE(OPC)                                                                                          \
    case OPC:                                                                                                  \
        _Pragma("unroll") for (int v = 0; v < VEC; ++v)                                                        \
            op_apply_io<HM>(OPC, tab, L, in[0][v], in[1][v], in[2][v], out[0][v], out[1][v], out[2][v],        \
                            m[0][v], m[1][v], m[2][v]);                                                        \
        break;
    switch (op) {
        T2O_CASE(OP_BRIGHTNESS) T2O_CASE(OP_CONTRAST) T2O_CASE(OP_SATURATION) T2O_CASE(OP_COLOR)
        T2O_CASE(OP_TONE) T2O_CASE(OP_WHITE) T2O_CASE(OP_EXPOSURE) T2O_CASE(OP_WHITEBALANCE)
        default:
            _Pragma("unroll") for (int v = 0; v < VEC; ++v) { out[0][v] = in[0][v]; out[1][v] = in[1][v]; out[2][v] = in[2][v]; }
            break;
    }
#undef T2O_CASE
}

template <int VEC, bool HM>
__device__ __forceinline__ void bwd_op_vec(int op, const float *tab, int L, const float (&x)[3][VEC],
                                           const float (&m)[3][VEC], float (&g)[3][VEC],
                                           float *acc, const Hist &hist, bool own) {
#define T2O_CASE(OPC)                                                                                   \
    case OPC:                                                                                           \
        _Pragma("unroll") for (int v = 0; v < VEC; ++v)                                                 \
            pointwise_bwd<HM>(OPC, tab, L, x[0][v], x[1][v], x[2][v], m[0][v], m[1][v], m[2][v],        \
                              g[0][v], g[1][v], g[2][v], acc, hist, own);                               \
        break;
    switch (op) {
        T2O_CASE(OP_BRIGHTNESS) T2O_CASE(OP_CONTRAST) T2O_CASE(OP_SATURATION) T2O_CASE(OP_COLOR)
        T2O_CASE(OP_TONE) T2O_CASE(OP_WHITE) T2O_CASE(OP_EXPOSURE) T2O_CASE(OP_WHITEBALANCE)
        default: break;
    }
#undef T2O_CASE
}

template <int VEC, bool HM>
__device__ __forceinline__ void ld_mask_t(const float *mask_b, int mask_ch, size_t plane, size_t off, float (&m)[3][VEC]) {
    if constexpr (HM) ld_mask<VEC>(mask_b, mask_ch, plane, off, m);
}

// 5-point stencil of one plane around a VEC-pixel group held in a shared-memory region.
//   row: pointer to the group's first float in its row;  rstride: floats per region row
//   lo / hi: first / one-past-last valid float offset relative to `row` within that row
template <int VEC>
__device__ __forceinline__ void stencil_group(const float *row, int rstride, int lo, int hi,
                                              float (&ctr)[VEC], float (&lap)[VEC]) {
    float up[VEC], dn[VEC];
    lds_vec<VEC>(row, ctr);
    lds_vec<VEC>(row - rstride, up);
    lds_vec<VEC>(row + rstride, dn);
    const float lf = (-1 >= lo) ? row[-1] : 0.0f;
    const float rt = (VEC < hi) ? row[VEC] : 0.0f;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        const float l = v > 0 ? ctr[v - 1] : lf;
        const float r = v < VEC - 1 ? ctr[v + 1] : rt;
        lap[v] = laplace(ctr[v], up[v], dn[v], l, r);
    }
}

// n / d for the small region indices of the 2-D kernels: mul = ceil(2^32 / d), exact for n * d < 2^32
__device__ __forceinline__ int fast_div(int n, int d, unsigned int mul) {
    return d == 1 ? n : (int)__umulhi((unsigned int)n, mul);
}

template <int VEC>
__device__ __forceinline__ void zero_px(float (&x)[3][VEC]) {
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int v = 0; v < VEC; ++v) x[c][v] = 0.0f;
}
template <int VEC>
__device__ __forceinline__ float l1_px(const float (&x)[3][VEC], const float (&t)[3][VEC]) {
    float s = 0.0f;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int v = 0; v < VEC; ++v) s += fabsf(x[c][v] - t[c][v]);
    return s;
}

// =========================================================================================== forward
struct FwdArgs {
    ChainDesc ch;
    Geom g;
    const float *img, *mask, *params, *target;
    float *out, *l1_sum;
    float *part_l1;
    unsigned int *counters;
    int mask_ch, pstride;
    int raw;                // T2O_FLAG_RAW_PROCESS
};

template <int VEC, bool SHARP, bool HM>
__global__ void __launch_bounds__(NT) chain_fwd_kernel(const __grid_constant__ FwdArgs a) {
    extern __shared__ __align__(16) float dyn_smem[];
    __shared__ __align__(16) float tabs[MAX_CHAIN][TAB];
    __shared__ float red[32];
    __shared__ int last_flag;

    const int tid = threadIdx.x;
    const int b = blockIdx.y, tile = blockIdx.x;
    const int n = a.ch.n, L = a.ch.L;
    const size_t plane = (size_t)a.g.H * a.g.W;
    const float *img_b = a.img + (size_t)b * 3 * plane;
    const float *tgt_b = a.target ? a.target + (size_t)b * 3 * plane : nullptr;
    float *out_b = a.out ? a.out + (size_t)b * 3 * plane : nullptr;
    const float *mask_b = HM ? a.mask + (size_t)b * a.mask_ch * plane : nullptr;
    const bool raw = a.raw != 0;

    if (tid < n) build_table(a.ch.op[tid], a.params + (size_t)b * a.pstride + a.ch.poff[tid], L, tabs[tid]);
    __syncthreads();

    float l1 = 0.0f;
    if constexpr (!SHARP) {
        const long long g0 = (long long)tile * a.g.tile_groups;
        long long g1 = g0 + a.g.tile_groups;
        if (g1 > a.g.ngroups) g1 = a.g.ngroups;
        long long gi = g0 + tid;
        // two groups (2 x VEC pixels) per operator dispatch while both exist
        for (; gi + NT < g1; gi += 2 * NT) {
            const size_t off0 = (size_t)gi * VEC, off1 = (size_t)(gi + NT) * VEC;
            float x[2][3][VEC], m[2][3][VEC];
            ld_px<VEC>(img_b, plane, off0, x[0]);
            ld_px<VEC>(img_b, plane, off1, x[1]);
            ld_mask_t<VEC, HM>(mask_b, a.mask_ch, plane, off0, m[0]);
            ld_mask_t<VEC, HM>(mask_b, a.mask_ch, plane, off1, m[1]);
            for (int k = 0; k < n; ++k) apply_op_grp<VEC, 2, HM>(a.ch.op[k], tabs[k], L, x, m, raw);
            if (tgt_b) {
                float t[3][VEC];
                ld_px<VEC>(tgt_b, plane, off0, t);
                l1 += l1_px<VEC>(x[0], t);
                ld_px<VEC>(tgt_b, plane, off1, t);
                l1 += l1_px<VEC>(x[1], t);
            }
            if (out_b) { st_px<VEC>(out_b, plane, off0, x[0]); st_px<VEC>(out_b, plane, off1, x[1]); }
        }
        for (; gi < g1; gi += NT) {
            const size_t off = (size_t)gi * VEC;
            float x[3][VEC], m[3][VEC];
            ld_px<VEC>(img_b, plane, off, x);
            ld_mask_t<VEC, HM>(mask_b, a.mask_ch, plane, off, m);
            for (int k = 0; k < n; ++k) apply_op_vec<VEC, HM>(a.ch.op[k], tabs[k], L, x, m, raw);
            if (tgt_b) {
                float t[3][VEC];
                ld_px<VEC>(tgt_b, plane, off, t);
                l1 += l1_px<VEC>(x, t);
            }
            if (out_b) st_px<VEC>(out_b, plane, off, x);
        }
    } else {
        const int H = a.g.H, Wg = a.g.Wg, TH = a.g.TH, TWg = a.g.TWg;
        const int ty = fast_div(tile, a.g.tiles_x, a.g.mul_tiles_x), tx = tile - ty * a.g.tiles_x;
        const int y0 = ty * TH, xg0 = tx * TWg;
        const int RH = TH + 2, RWg = TWg + 2;
        const int rstride = RWg * VEC;                 // floats per region row
        const int cstride = RH * rstride;              // floats per region plane
        const int sp = a.ch.sharp;                     // 0 <= sp < n
        // ---- phase A: operators before the stencil on tile + halo -> shared memory
        for (int idx = tid; idx < RH * RWg; idx += NT) {
            const int ry = fast_div(idx, RWg, a.g.mul_rw), rxg = idx - ry * RWg;
            const int y = y0 - 1 + ry, xg = xg0 - 1 + rxg;
            float x[3][VEC];
            if (y >= 0 && y < H && xg >= 0 && xg < Wg) {
                const size_t off = (size_t)y * a.g.W + (size_t)xg * VEC;
                float m[3][VEC];
                ld_px<VEC>(img_b, plane, off, x);
                if (sp > 0) {
                    ld_mask_t<VEC, HM>(mask_b, a.mask_ch, plane, off, m);
                    for (int k = 0; k < sp; ++k) apply_op_vec<VEC, HM>(a.ch.op[k], tabs[k], L, x, m);
                }
            } else {                                    // zero padding of the stencil input
                zero_px<VEC>(x);
            }
            float *dst = dyn_smem + ry * rstride + rxg * VEC;
#pragma unroll
            for (int c = 0; c < 3; ++c) st_vec<VEC>(dst + c * cstride, x[c]);
        }
        __syncthreads();
        // ---- phase B: stencil + remaining operators on the tile interior
        const float p = tabs[sp][0];
        for (int idx = tid; idx < TH * TWg; idx += NT) {
            const int ly = fast_div(idx, TWg, a.g.mul_tw), lxg = idx - ly * TWg;
            const int y = y0 + ly, xg = xg0 + lxg;
            if (y >= H || xg >= Wg) continue;
            const size_t off = (size_t)y * a.g.W + (size_t)xg * VEC;
            float x[3][VEC], m[3][VEC];
            ld_mask_t<VEC, HM>(mask_b, a.mask_ch, plane, off, m);
            const float *src = dyn_smem + (ly + 1) * rstride + (lxg + 1) * VEC;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float ctr[VEC], lap[VEC];
                stencil_group<VEC>(src + c * cstride, rstride, -(lxg + 1) * VEC, (RWg - lxg - 1) * VEC, ctr, lap);
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    const float yv = fmaf(p, lap[v], ctr[v]);
                    x[c][v] = raw ? yv : sat01(blend<HM>(yv, ctr[v], m[c][v]));
                }
            }
            for (int k = sp + 1; k < n; ++k) apply_op_vec<VEC, HM>(a.ch.op[k], tabs[k], L, x, m);
            if (tgt_b) {
                float t[3][VEC];
                ld_px<VEC>(tgt_b, plane, off, t);
                l1 += l1_px<VEC>(x, t);
            }
            if (out_b) st_px<VEC>(out_b, plane, off, x);
        }
    }

    if (a.l1_sum) {
        const float s = block_sum(l1, red);
        const int ntiles = a.g.ntiles;
        if (tid == 0) a.part_l1[(size_t)b * ntiles + tile] = s;
        if (arrive_is_last(a.counters + b, (unsigned)ntiles, &last_flag)) {
            float v = 0.0f;
            for (int t = tid; t < ntiles; t += NT) v += __ldcg(a.part_l1 + (size_t)b * ntiles + t);
            v = block_sum(v, red);
            if (tid == 0) a.l1_sum[b] = v;
        }
    }
}

// =========================================================================================== backward
struct BwdArgs {
    ChainDesc ch;
    Geom g;
    const float *img, *mask, *params, *grad_out, *target, *grad_l1;
    float *grad_params, *grad_img, *out, *l1_sum;
    float *part_l1, *part_gp;
    unsigned int *counters;
    int mask_ch, pstride;
};

// upstream gradient of a pixel group: explicit grad_out, or the fused L1: gl1 * sign(out - target)
template <int VEC>
__device__ __forceinline__ void upstream(const float *go_b, const float *tgt_b, size_t plane,
                                         size_t off, float gl1, const float (&x)[3][VEC], float (&g)[3][VEC],
                                         float &l1, bool own) {
    if (go_b) {
        ld_px<VEC>(go_b, plane, off, g);
        if (tgt_b && own) {
            float t[3][VEC];
            ld_px<VEC>(tgt_b, plane, off, t);
            l1 += l1_px<VEC>(x, t);
        }
    } else {
        float t[3][VEC];
        ld_px<VEC>(tgt_b, plane, off, t);
        float s = 0.0f;
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const float d = x[c][v] - t[c][v];
                g[c][v] = d != 0.0f ? copysignf(gl1, gl1 < 0.0f ? -d : d) : 0.0f;
                s += fabsf(d);
            }
        if (own) l1 += s;
    }
}

template <int VEC, int KMAX, bool SHARP, bool HM>
__global__ void __launch_bounds__(NT, 2) chain_bwd_kernel(const __grid_constant__ BwdArgs a) {
    extern __shared__ __align__(16) float dyn_smem[];
    __shared__ __align__(16) float tabs[MAX_CHAIN][TAB];
    __shared__ float rowbuf[MAX_PSTRIDE];
    __shared__ float red[32];
    __shared__ int last_flag;

    const int tid = threadIdx.x;
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int n = a.ch.n, L = a.ch.L;
    const size_t plane = (size_t)a.g.H * a.g.W;
    const float *img_b = a.img + (size_t)b * 3 * plane;
    const float *tgt_b = a.target ? a.target + (size_t)b * 3 * plane : nullptr;
    const float *go_b = a.grad_out ? a.grad_out + (size_t)b * 3 * plane : nullptr;
    float *out_b = a.out ? a.out + (size_t)b * 3 * plane : nullptr;
    float *gi_b = a.grad_img ? a.grad_img + (size_t)b * 3 * plane : nullptr;
    const float *mask_b = HM ? a.mask + (size_t)b * a.mask_ch * plane : nullptr;
    const float gl1 = a.grad_l1 ? a.grad_l1[b] : 0.0f;

    // dynamic shared memory: [curve moments: hist_total x NT float2] [2-D regions (SHARP)]
    F2 *hist_mem = reinterpret_cast<F2 *>(dyn_smem);
    float *region = dyn_smem + (size_t)a.ch.hist_total * NT * 2;
    for (int i = tid; i < a.ch.hist_total * NT * 2; i += NT) dyn_smem[i] = 0.0f;
    if (tid < n) build_table(a.ch.op[tid], a.params + (size_t)b * a.pstride + a.ch.poff[tid], L, tabs[tid]);
    for (int i = tid; i < MAX_PSTRIDE; i += NT) rowbuf[i] = 0.0f;
    __syncthreads();

    float acc[KMAX][3];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) { acc[k][0] = 0.0f; acc[k][1] = 0.0f; acc[k][2] = 0.0f; }
    float l1 = 0.0f;
    float acc_sharp = 0.0f;                              // parameter gradient of the stencil operator
    F2 *hist_t = hist_mem + tid;

    if constexpr (!SHARP) {
        const long long g0 = (long long)chunk * a.g.tile_groups;
        long long g1 = g0 + a.g.tile_groups;
        if (g1 > a.g.ngroups) g1 = a.g.ngroups;
        for (long long gi = g0 + tid; gi < g1; gi += NT) {
            const size_t off = (size_t)gi * VEC;
            float x[3][VEC], m[3][VEC], g[3][VEC];
            float sv[KMAX + 1][3][VEC];                 // sv[k] = input of operator k, sv[n] = output
            ld_px<VEC>(img_b, plane, off, sv[0]);
            ld_mask_t<VEC, HM>(mask_b, a.mask_ch, plane, off, m);
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
                if (k < n) {
                    apply_io_vec<VEC, HM>(a.ch.op[k], tabs[k], L, sv[k], sv[k + 1], m);
                    if (k == n - 1) {
#pragma unroll
                        for (int c = 0; c < 3; ++c)
#pragma unroll
                            for (int v = 0; v < VEC; ++v) x[c][v] = sv[k + 1][c][v];
                    }
                }
            }
            upstream<VEC>(go_b, tgt_b, plane, off, gl1, x, g, l1, true);
            if (out_b) st_px<VEC>(out_b, plane, off, x);
#pragma unroll
            for (int k = KMAX - 1; k >= 0; --k) {
                if (k < n)
                    bwd_op_vec<VEC, HM>(a.ch.op[k], tabs[k], L, sv[k], m, g, acc[k],
                                        Hist{hist_t + a.ch.hoff[k] * NT, NT}, true);
            }
            if (gi_b) st_px<VEC>(gi_b, plane, off, g);
        }
    } else {
        const int H = a.g.H, Wg = a.g.Wg, TH = a.g.TH, TWg = a.g.TWg;
        constexpr int HGX = VEC == 1 ? 2 : 1;          // halo groups of the X region
        const int XH = TH + 4, XWg = TWg + 2 * HGX;    // X : operators-before-stencil output, tile + 2
        const int GH = TH + 2, GWg = TWg + 2;          // GY: gradient at the stencil output, tile + 1
        const int xrs = XWg * VEC, xcs = XH * xrs;
        const int grs = GWg * VEC, gcs = GH * grs;
        float *Xs = region;
        float *GYs = Xs + 3 * xcs;
        float *GDs = GYs + 3 * gcs;                    // only with a mask
        const int sp = a.ch.sharp;
        const float p = tabs[sp][0];
        const int t_begin = chunk * a.g.tiles_per_cta;
        int t_end = t_begin + a.g.tiles_per_cta;
        if (t_end > a.g.ntiles) t_end = a.g.ntiles;
        for (int tile = t_begin; tile < t_end; ++tile) {
            const int ty = fast_div(tile, a.g.tiles_x, a.g.mul_tiles_x), tx = tile - ty * a.g.tiles_x;
            const int y0 = ty * TH, xg0 = tx * TWg;
            // ---- phase A: X on tile + 2
            for (int idx = tid; idx < XH * XWg; idx += NT) {
                const int ry = fast_div(idx, XWg, a.g.mul_xw), rxg = idx - ry * XWg;
                const int y = y0 - 2 + ry, xg = xg0 - HGX + rxg;
                float x[3][VEC];
                if (y >= 0 && y < H && xg >= 0 && xg < Wg) {
                    const size_t off = (size_t)y * a.g.W + (size_t)xg * VEC;
                    float m[3][VEC];
                    ld_px<VEC>(img_b, plane, off, x);
                    if (sp > 0) {
                        ld_mask_t<VEC, HM>(mask_b, a.mask_ch, plane, off, m);
                        for (int k = 0; k < sp; ++k) apply_op_vec<VEC, HM>(a.ch.op[k], tabs[k], L, x, m);
                    }
                } else {
                    zero_px<VEC>(x);
                }
                float *dst = Xs + ry * xrs + rxg * VEC;
#pragma unroll
                for (int c = 0; c < 3; ++c) st_vec<VEC>(dst + c * xcs, x[c]);
            }
            __syncthreads();
            // ---- phase B: stencil + operators after it, forward and backward, on tile + 1 -> GY
            for (int idx = tid; idx < GH * GWg; idx += NT) {
                const int ry = fast_div(idx, GWg, a.g.mul_gw), rxg = idx - ry * GWg;
                const int y = y0 - 1 + ry, xg = xg0 - 1 + rxg;
                float gy[3][VEC], gd[3][VEC];
                const bool inb = y >= 0 && y < H && xg >= 0 && xg < Wg;
                if (inb) {
                    const bool own = ry >= 1 && ry <= TH && rxg >= 1 && rxg <= TWg;
                    const size_t off = (size_t)y * a.g.W + (size_t)xg * VEC;
                    float x[3][VEC], m[3][VEC], g[3][VEC], ctr[3][VEC], lap[3][VEC];
                    float sv[KMAX + 1][3][VEC];
                    ld_mask_t<VEC, HM>(mask_b, a.mask_ch, plane, off, m);
                    const int xgx = rxg - 1 + HGX;          // group index inside the X region
                    const float *src = Xs + (ry + 1) * xrs + xgx * VEC;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        stencil_group<VEC>(src + c * xcs, xrs, -xgx * VEC, (XWg - xgx) * VEC, ctr[c], lap[c]);
#pragma unroll
                        for (int v = 0; v < VEC; ++v)
                            x[c][v] = sat01(blend<HM>(fmaf(p, lap[c][v], ctr[c][v]), ctr[c][v], m[c][v]));
                    }
#pragma unroll
                    for (int k = 1; k < KMAX; ++k) {
                        if (k > sp && k < n) {
                            if (k == sp + 1) {
#pragma unroll
                                for (int c = 0; c < 3; ++c)
#pragma unroll
                                    for (int v = 0; v < VEC; ++v) sv[k][c][v] = x[c][v];
                            }
                            apply_io_vec<VEC, HM>(a.ch.op[k], tabs[k], L, sv[k], sv[k + 1], m);
                            if (k == n - 1) {
#pragma unroll
                                for (int c = 0; c < 3; ++c)
#pragma unroll
                                    for (int v = 0; v < VEC; ++v) x[c][v] = sv[k + 1][c][v];
                            }
                        }
                    }
                    upstream<VEC>(go_b, tgt_b, plane, off, gl1, x, g, l1, own);
                    if (out_b && own) st_px<VEC>(out_b, plane, off, x);
#pragma unroll
                    for (int k = KMAX - 1; k >= 1; --k) {
                        if (k > sp && k < n)
                            bwd_op_vec<VEC, HM>(a.ch.op[k], tabs[k], L, sv[k], m, g, acc[k],
                                                Hist{hist_t + a.ch.hoff[k] * NT, NT}, own);
                    }
                    float accp = 0.0f;
#pragma unroll
                    for (int c = 0; c < 3; ++c)
#pragma unroll
                        for (int v = 0; v < VEC; ++v) {
                            blend_bwd<HM>(fmaf(p, lap[c][v], ctr[c][v]), ctr[c][v], m[c][v], g[c][v], gy[c][v], gd[c][v]);
                            accp = fmaf(gy[c][v], lap[c][v], accp);
                        }
                    if (own) acc_sharp += accp;
                } else {                                    // outside the image: no stencil output there
                    zero_px<VEC>(gy);
                    zero_px<VEC>(gd);
                }
                float *dst = GYs + ry * grs + rxg * VEC;
#pragma unroll
                for (int c = 0; c < 3; ++c) st_vec<VEC>(dst + c * gcs, gy[c]);
                if constexpr (HM) {
                    float *dd = GDs + ry * grs + rxg * VEC;
#pragma unroll
                    for (int c = 0; c < 3; ++c) st_vec<VEC>(dd + c * gcs, gd[c]);
                }
            }
            __syncthreads();
            // ---- phase C: transposed stencil, then the operators before it, on the tile interior
            if (gi_b != nullptr || sp > 0) {
                for (int idx = tid; idx < TH * TWg; idx += NT) {
                    const int ly = fast_div(idx, TWg, a.g.mul_tw), lxg = idx - ly * TWg;
                    const int y = y0 + ly, xg = xg0 + lxg;
                    if (y >= H || xg >= Wg) continue;
                    const size_t off = (size_t)y * a.g.W + (size_t)xg * VEC;
                    float g[3][VEC];
                    const float *src = GYs + (ly + 1) * grs + (lxg + 1) * VEC;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        float ctr[VEC], lap[VEC];
                        stencil_group<VEC>(src + c * gcs, grs, -(lxg + 1) * VEC, (GWg - lxg - 1) * VEC, ctr, lap);
                        float gdv[VEC];
                        if constexpr (HM) lds_vec<VEC>(GDs + c * gcs + (ly + 1) * grs + (lxg + 1) * VEC, gdv);
#pragma unroll
                        for (int v = 0; v < VEC; ++v) g[c][v] = fmaf(p, lap[v], ctr[v]) + (HM ? gdv[v] : 0.0f);
                    }
                    if (sp > 0) {
                        float m[3][VEC];
                        float sv[KMAX][3][VEC];
                        ld_px<VEC>(img_b, plane, off, sv[0]);
                        ld_mask_t<VEC, HM>(mask_b, a.mask_ch, plane, off, m);
#pragma unroll
                        for (int k = 0; k < KMAX - 2; ++k) {
                            if (k + 1 < sp) apply_io_vec<VEC, HM>(a.ch.op[k], tabs[k], L, sv[k], sv[k + 1], m);
                        }
#pragma unroll
                        for (int k = KMAX - 2; k >= 0; --k) {
                            if (k < sp)
                                bwd_op_vec<VEC, HM>(a.ch.op[k], tabs[k], L, sv[k], m, g, acc[k],
                                                    Hist{hist_t + a.ch.hoff[k] * NT, NT}, true);
                        }
                    }
                    if (gi_b) st_px<VEC>(gi_b, plane, off, g);
                }
            }
            __syncthreads();      // the regions are rewritten by the next tile
        }
    }

    // ---------------------------------------------------------------- CTA partials
    const int nchunks = a.g.nchunks;
    if (a.grad_params) {
        __syncthreads();                                 // all curve moments are in shared memory
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            if (k < n) {
                const int op = a.ch.op[k];
                if (op_is_curve(op)) {
                    // block-reduce the (A, Bx) moments of this operator: warp w reduces slots w, w+8, ...
                    const int nslots = op_hist_slots(op);
                    const int lane = tid & 31, warp = tid >> 5;
                    F2 *hk = hist_mem + (size_t)a.ch.hoff[k] * NT;
                    for (int s = warp; s < nslots; s += NT / 32) {
                        float va = 0.0f, vb = 0.0f;
#pragma unroll
                        for (int i = 0; i < NT / 32; ++i) { const F2 e = hk[s * NT + lane + 32 * i]; va += e.a; vb += e.b; }
                        va = warp_sum(va); vb = warp_sum(vb);
                        if (lane == 0) { hk[s * NT].a = va; hk[s * NT].b = vb; }   // slot total parked in element 0
                    }
                    const int ncur = op == OP_TONE ? 1 : 3;
                    float cs[3] = {0.0f, 0.0f, 0.0f};
                    for (int c = 0; c < ncur; ++c) cs[c] = block_sum(acc[k][c], red);   // C = sum g*y (also a barrier)
                    __syncthreads();
                    if (tid == 0) {
                        for (int c = 0; c < ncur; ++c) {
                            float A[NBIN], Bx[NBIN], gk[MAX_L];
                            for (int s = 0; s < NBIN; ++s) { const F2 e = hk[(c * NBIN + s) * NT]; A[s] = e.a; Bx[s] = e.b; }
                            curve_param_grad(tabs[k] + c * CT, L, A, Bx, cs[c], gk);
                            for (int i = 0; i < L; ++i) rowbuf[a.ch.poff[k] + c * L + i] = gk[i];
                        }
                    }
                } else if (op == OP_SHARPNESS) {
                    const float s = block_sum(acc_sharp, red);
                    if (tid == 0) rowbuf[a.ch.poff[k]] = s;
                } else if (op == OP_WHITEBALANCE) {
                    for (int i = 0; i < 3; ++i) {
                        const float s = block_sum(acc[k][i], red);
                        if (tid == 0) rowbuf[a.ch.poff[k] + i] = s;
                    }
                } else if (op >= 0 && op != OP_WHITE) {
                    const float s = block_sum(acc[k][0], red);
                    if (tid == 0) rowbuf[a.ch.poff[k]] = s;
                }
            }
        }
        __syncthreads();
        float *prow = a.part_gp + ((size_t)b * nchunks + chunk) * a.pstride;
        for (int i = tid; i < a.pstride; i += NT) prow[i] = rowbuf[i];
    }
    if (a.l1_sum) {
        const float s = block_sum(l1, red);
        if (tid == 0) a.part_l1[(size_t)b * nchunks + chunk] = s;
    }
    if (a.grad_params || a.l1_sum) {
        if (arrive_is_last(a.counters + b, (unsigned)nchunks, &last_flag)) {
            if (a.grad_params)
                reduce_columns(a.part_gp + (size_t)b * nchunks * a.pstride, nchunks, a.pstride,
                               a.grad_params + (size_t)b * a.pstride);
            if (a.l1_sum) {
                float v = 0.0f;
                for (int t = tid; t < nchunks; t += NT) v += __ldcg(a.part_l1 + (size_t)b * nchunks + t);
                v = block_sum(v, red);
                if (tid == 0) a.l1_sum[b] = v;
            }
        }
    }
}

// =========================================================================================== L1 only
struct L1Args {
    const float *a, *b;
    float *l1_sum, *part;
    unsigned int *counters;
    long long n;          // floats per image
    int ntiles;
    long long tile_elems;
};

template <int VEC>
__global__ void __launch_bounds__(NT) l1_sum_kernel(const __grid_constant__ L1Args a) {
    __shared__ float red[32];
    __shared__ int last_flag;
    const int tid = threadIdx.x, b = blockIdx.y, tile = blockIdx.x;
    const float *pa = a.a + (size_t)b * a.n, *pb = a.b + (size_t)b * a.n;
    const long long e0 = (long long)tile * a.tile_elems;
    long long e1 = e0 + a.tile_elems;
    if (e1 > a.n) e1 = a.n;
    float s = 0.0f;
    for (long long e = e0 + (long long)tid * VEC; e < e1; e += (long long)NT * VEC) {
        float x[VEC], y[VEC];
        ld_vec<VEC>(pa + e, x);
        ld_vec<VEC>(pb + e, y);
#pragma unroll
        for (int v = 0; v < VEC; ++v) s += fabsf(x[v] - y[v]);
    }
    s = block_sum(s, red);
    if (tid == 0) a.part[(size_t)b * a.ntiles + tile] = s;
    if (arrive_is_last(a.counters + b, (unsigned)a.ntiles, &last_flag)) {
        float v = 0.0f;
        for (int t = tid; t < a.ntiles; t += NT) v += __ldcg(a.part + (size_t)b * a.ntiles + t);
        v = block_sum(v, red);
        if (tid == 0) a.l1_sum[b] = v;
    }
}

// Opt in to the dynamic shared memory a launch needs.  The 48 KB default limit counts static +
// dynamic bytes, so anything above 32 KB is configured explicitly.
template <typename K>
static int set_smem(K kernel, size_t bytes) {
    if (bytes > 32 * 1024) {
        if (bytes > 227 * 1024) return T2O_ERR_UNSUPPORTED;
        T2O_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    }
    return T2O_OK;
}

template <int VEC, int KMAX, bool SHARP, bool HM>
static int launch_bwd(const BwdArgs &a, size_t smem, cudaStream_t stream) {
    int st = set_smem(chain_bwd_kernel<VEC, KMAX, SHARP, HM>, smem);
    if (st) return st;
    dim3 grid(a.g.nchunks, a.g.B);
    chain_bwd_kernel<VEC, KMAX, SHARP, HM><<<grid, NT, smem, stream>>>(a);
    T2O_CUDA_OK(cudaGetLastError());
    return T2O_OK;
}
template <bool SHARP, bool HM>
static int launch_bwd_sel(int vec, bool small_chain, const BwdArgs &a, size_t smem, cudaStream_t stream) {
    if (small_chain) {
        if (vec == 4) return launch_bwd<4, 2, SHARP, HM>(a, smem, stream);
        if (vec == 2) return launch_bwd<2, 2, SHARP, HM>(a, smem, stream);
        return launch_bwd<1, 2, SHARP, HM>(a, smem, stream);
    }
    if (vec == 2) return launch_bwd<2, MAX_CHAIN, SHARP, HM>(a, smem, stream);
    return launch_bwd<1, MAX_CHAIN, SHARP, HM>(a, smem, stream);
}

// one translation unit per (SHARP, HM) pair, so the 20 backward instantiations compile in parallel
int launch_bwd_flat_nomask(int vec, bool small_chain, const BwdArgs &a, size_t smem, cudaStream_t stream);
int launch_bwd_flat_mask(int vec, bool small_chain, const BwdArgs &a, size_t smem, cudaStream_t stream);
int launch_bwd_sharp_nomask(int vec, bool small_chain, const BwdArgs &a, size_t smem, cudaStream_t stream);
int launch_bwd_sharp_mask(int vec, bool small_chain, const BwdArgs &a, size_t smem, cudaStream_t stream);

}  // namespace t2o
