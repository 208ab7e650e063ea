// t2o_chain.cu -- fused operator-chain kernels for sm_100a.
//
//   chain_fwd_kernel   K x Operator.execute (+ per-image L1 to a target) in one pass over HBM
//   chain_bwd_kernel   the same chain recomputed and differentiated in one pass: parameter
//                      gradients (+ optional image gradient, + optional forward outputs)
//
// Replaces: K successive Executor.execute calls (executors/executor.py:33-55 ->
// models/operators.py:112-131), get_dist 'L1' (utils/beam_search.py:170-173) and autograd through
// them.  Elementwise and HBM-bound: no tensor cores by design.
//
// Data layout: images NCHW planar fp32; a thread owns VEC consecutive pixels of the three planes
// (128-bit coalesced accesses for VEC = 4).  Per-(image, op) tables (curve slopes / offsets,
// scalars) live in shared memory.  A sharpness operator turns the launch into a 2-D tiling with a
// halo: the operators before it are evaluated on tile + halo into shared memory, the 3x3 stencil
// and the operators after it run from there, so the chain still makes one pass over HBM.
// Reductions (L1, parameter gradients) are warp-shuffle -> shared -> one partial per CTA; the last
// CTA of each image sums the partials in a fixed order (deterministic, no float atomics).
#include <cstdio>
#include <cstring>

#include "t2o_common.cuh"
#include "../../include/t2o.h"

namespace t2o {

// ---------------------------------------------------------------- operator dispatch over a pixel group
template <int VEC>
__device__ __forceinline__ void apply_op_vec(int op, const float *tab, int L, float (&x)[3][VEC],
                                             const float (&m)[3][VEC], bool has_mask, bool raw = false) {
#define T2O_CASE(OPC)                                                                                   \
    case OPC:                                                                                           \
        _Pragma("unroll") for (int v = 0; v < VEC; ++v)                                                 \
            op_apply(OPC, tab, L, x[0][v], x[1][v], x[2][v], m[0][v], m[1][v], m[2][v], has_mask, raw); \
        break;
    switch (op) {
        T2O_CASE(OP_BRIGHTNESS) T2O_CASE(OP_CONTRAST) T2O_CASE(OP_SATURATION) T2O_CASE(OP_COLOR)
        T2O_CASE(OP_TONE) T2O_CASE(OP_WHITE) T2O_CASE(OP_EXPOSURE) T2O_CASE(OP_WHITEBALANCE)
        default: break;
    }
#undef T2O_CASE
}

template <int VEC>
__device__ __forceinline__ void bwd_op_vec(int op, const float *tab, int L, const float (&x)[3][VEC],
                                           const float (&m)[3][VEC], bool has_mask, float (&g)[3][VEC],
                                           float *acc, const Hist &hist, bool own) {
#define T2O_CASE(OPC)                                                                                   \
    case OPC:                                                                                           \
        _Pragma("unroll") for (int v = 0; v < VEC; ++v)                                                 \
            pointwise_bwd(OPC, tab, L, x[0][v], x[1][v], x[2][v], m[0][v], m[1][v], m[2][v], has_mask,  \
                          g[0][v], g[1][v], g[2][v], acc, hist, own);                                   \
        break;
    switch (op) {
        T2O_CASE(OP_BRIGHTNESS) T2O_CASE(OP_CONTRAST) T2O_CASE(OP_SATURATION) T2O_CASE(OP_COLOR)
        T2O_CASE(OP_TONE) T2O_CASE(OP_WHITE) T2O_CASE(OP_EXPOSURE) T2O_CASE(OP_WHITEBALANCE)
        default: break;
    }
#undef T2O_CASE
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

template <int VEC, bool SHARP>
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
    const float *mask_b = a.mask ? a.mask + (size_t)b * a.mask_ch * plane : nullptr;
    const bool has_mask = mask_b != nullptr;

    if (tid < n) build_table(a.ch.op[tid], a.params + (size_t)b * a.pstride + a.ch.poff[tid], L, tabs[tid]);
    __syncthreads();

    float l1 = 0.0f;
    if constexpr (!SHARP) {
        const long long g0 = (long long)tile * a.g.tile_groups;
        long long g1 = g0 + a.g.tile_groups;
        if (g1 > a.g.ngroups) g1 = a.g.ngroups;
        for (long long gi = g0 + tid; gi < g1; gi += NT) {
            const size_t off = (size_t)gi * VEC;
            float x[3][VEC], m[3][VEC], t[3][VEC];
            ld_px<VEC>(img_b, plane, off, x);
            ld_mask<VEC>(mask_b, a.mask_ch, plane, off, m);
            if (tgt_b) ld_px<VEC>(tgt_b, plane, off, t);
            for (int k = 0; k < n; ++k) apply_op_vec<VEC>(a.ch.op[k], tabs[k], L, x, m, has_mask, a.raw != 0);
            if (tgt_b) {
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int v = 0; v < VEC; ++v) l1 += fabsf(x[c][v] - t[c][v]);
            }
            if (out_b) st_px<VEC>(out_b, plane, off, x);
        }
    } else {
        const int H = a.g.H, Wg = a.g.Wg, TH = a.g.TH, TWg = a.g.TWg;
        const int ty = tile / a.g.tiles_x, tx = tile - ty * a.g.tiles_x;
        const int y0 = ty * TH, xg0 = tx * TWg;
        const int RH = TH + 2, RWg = TWg + 2;
        const int rstride = RWg * VEC;                 // floats per region row
        const int cstride = RH * rstride;              // floats per region plane
        const int sp = a.ch.sharp;                     // 0 <= sp < n
        // ---- phase A: operators before the stencil on tile + halo -> shared memory
        for (int idx = tid; idx < RH * RWg; idx += NT) {
            const int ry = idx / RWg, rxg = idx - ry * RWg;
            const int y = y0 - 1 + ry, xg = xg0 - 1 + rxg;
            float x[3][VEC];
            if (y >= 0 && y < H && xg >= 0 && xg < Wg) {
                const size_t off = (size_t)y * a.g.W + (size_t)xg * VEC;
                float m[3][VEC];
                ld_px<VEC>(img_b, plane, off, x);
                if (sp > 0) {
                    ld_mask<VEC>(mask_b, a.mask_ch, plane, off, m);
                    for (int k = 0; k < sp; ++k) apply_op_vec<VEC>(a.ch.op[k], tabs[k], L, x, m, has_mask);
                }
            } else {                                    // zero padding of the stencil input
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int v = 0; v < VEC; ++v) x[c][v] = 0.0f;
            }
            float *dst = dyn_smem + ry * rstride + rxg * VEC;
#pragma unroll
            for (int c = 0; c < 3; ++c) st_vec<VEC>(dst + c * cstride, x[c]);
        }
        __syncthreads();
        // ---- phase B: stencil + remaining operators on the tile interior
        const float p = tabs[sp][0];
        for (int idx = tid; idx < TH * TWg; idx += NT) {
            const int ly = idx / TWg, lxg = idx - ly * TWg;
            const int y = y0 + ly, xg = xg0 + lxg;
            if (y >= H || xg >= Wg) continue;
            const size_t off = (size_t)y * a.g.W + (size_t)xg * VEC;
            float x[3][VEC], m[3][VEC], t[3][VEC];
            ld_mask<VEC>(mask_b, a.mask_ch, plane, off, m);
            if (tgt_b) ld_px<VEC>(tgt_b, plane, off, t);
            const float *src = dyn_smem + (ly + 1) * rstride + (lxg + 1) * VEC;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float ctr[VEC], lap[VEC];
                stencil_group<VEC>(src + c * cstride, rstride, -(lxg + 1) * VEC, (RWg - lxg - 1) * VEC, ctr, lap);
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    const float yv = fmaf(p, lap[v], ctr[v]);
                    x[c][v] = a.raw ? yv : sat01(blend(yv, ctr[v], m[c][v], has_mask));
                }
            }
            for (int k = sp + 1; k < n; ++k) apply_op_vec<VEC>(a.ch.op[k], tabs[k], L, x, m, has_mask);
            if (tgt_b) {
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int v = 0; v < VEC; ++v) l1 += fabsf(x[c][v] - t[c][v]);
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
__device__ __forceinline__ void upstream(const BwdArgs &a, const float *go_b, const float *tgt_b, size_t plane,
                                         size_t off, float gl1, const float (&x)[3][VEC], float (&g)[3][VEC],
                                         float &l1, bool own) {
    if (go_b) {
        ld_px<VEC>(go_b, plane, off, g);
        if (tgt_b && own) {
            float t[3][VEC];
            ld_px<VEC>(tgt_b, plane, off, t);
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int v = 0; v < VEC; ++v) l1 += fabsf(x[c][v] - t[c][v]);
        }
    } else {
        float t[3][VEC];
        ld_px<VEC>(tgt_b, plane, off, t);
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const float d = x[c][v] - t[c][v];
                g[c][v] = d > 0.0f ? gl1 : (d < 0.0f ? -gl1 : 0.0f);
                if (own) l1 += fabsf(d);
            }
    }
}

template <int VEC, int KMAX, bool SHARP>
__global__ void __launch_bounds__(NT) chain_bwd_kernel(const __grid_constant__ BwdArgs a) {
    extern __shared__ __align__(16) float dyn_smem[];
    __shared__ __align__(16) float tabs[MAX_CHAIN][TAB];
    __shared__ float rowbuf[MAX_PSTRIDE];
    __shared__ float red[32];
    __shared__ int last_flag;

    const int tid = threadIdx.x;
    const int b = blockIdx.y, tile = blockIdx.x;
    const int n = a.ch.n, L = a.ch.L;
    const size_t plane = (size_t)a.g.H * a.g.W;
    const float *img_b = a.img + (size_t)b * 3 * plane;
    const float *tgt_b = a.target ? a.target + (size_t)b * 3 * plane : nullptr;
    const float *go_b = a.grad_out ? a.grad_out + (size_t)b * 3 * plane : nullptr;
    float *out_b = a.out ? a.out + (size_t)b * 3 * plane : nullptr;
    float *gi_b = a.grad_img ? a.grad_img + (size_t)b * 3 * plane : nullptr;
    const float *mask_b = a.mask ? a.mask + (size_t)b * a.mask_ch * plane : nullptr;
    const bool has_mask = mask_b != nullptr;
    const float gl1 = a.grad_l1 ? a.grad_l1[b] : 0.0f;

    // dynamic shared memory: [curve moments: hist_total x NT] [2-D regions (SHARP)]
    float *hist_mem = dyn_smem;
    float *region = dyn_smem + (size_t)a.ch.hist_total * NT;
    for (int i = tid; i < a.ch.hist_total * NT; i += NT) hist_mem[i] = 0.0f;
    if (tid < n) build_table(a.ch.op[tid], a.params + (size_t)b * a.pstride + a.ch.poff[tid], L, tabs[tid]);
    for (int i = tid; i < MAX_PSTRIDE; i += NT) rowbuf[i] = 0.0f;
    __syncthreads();

    float acc[KMAX][3];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) { acc[k][0] = 0.0f; acc[k][1] = 0.0f; acc[k][2] = 0.0f; }
    float l1 = 0.0f;
    float acc_sharp = 0.0f;                              // parameter gradient of the stencil operator
    float *hist_t = hist_mem + tid;

    if constexpr (!SHARP) {
        const long long g0 = (long long)tile * a.g.tile_groups;
        long long g1 = g0 + a.g.tile_groups;
        if (g1 > a.g.ngroups) g1 = a.g.ngroups;
        for (long long gi = g0 + tid; gi < g1; gi += NT) {
            const size_t off = (size_t)gi * VEC;
            float x[3][VEC], m[3][VEC], g[3][VEC];
            float sv[KMAX][3][VEC];
            ld_px<VEC>(img_b, plane, off, x);
            ld_mask<VEC>(mask_b, a.mask_ch, plane, off, m);
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
                if (k < n) {
#pragma unroll
                    for (int c = 0; c < 3; ++c)
#pragma unroll
                        for (int v = 0; v < VEC; ++v) sv[k][c][v] = x[c][v];
                    apply_op_vec<VEC>(a.ch.op[k], tabs[k], L, x, m, has_mask);
                }
            }
            upstream<VEC>(a, go_b, tgt_b, plane, off, gl1, x, g, l1, true);
            if (out_b) st_px<VEC>(out_b, plane, off, x);
#pragma unroll
            for (int k = KMAX - 1; k >= 0; --k) {
                if (k < n)
                    bwd_op_vec<VEC>(a.ch.op[k], tabs[k], L, sv[k], m, has_mask, g, acc[k],
                                    Hist{hist_t + a.ch.hoff[k] * NT, NT}, true);
            }
            if (gi_b) st_px<VEC>(gi_b, plane, off, g);
        }
    } else {
        const int H = a.g.H, Wg = a.g.Wg, TH = a.g.TH, TWg = a.g.TWg;
        const int ty = tile / a.g.tiles_x, tx = tile - ty * a.g.tiles_x;
        const int y0 = ty * TH, xg0 = tx * TWg;
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
        // ---- phase A: X on tile + 2
        for (int idx = tid; idx < XH * XWg; idx += NT) {
            const int ry = idx / XWg, rxg = idx - ry * XWg;
            const int y = y0 - 2 + ry, xg = xg0 - HGX + rxg;
            float x[3][VEC];
            if (y >= 0 && y < H && xg >= 0 && xg < Wg) {
                const size_t off = (size_t)y * a.g.W + (size_t)xg * VEC;
                float m[3][VEC];
                ld_px<VEC>(img_b, plane, off, x);
                if (sp > 0) {
                    ld_mask<VEC>(mask_b, a.mask_ch, plane, off, m);
                    for (int k = 0; k < sp; ++k) apply_op_vec<VEC>(a.ch.op[k], tabs[k], L, x, m, has_mask);
                }
            } else {
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int v = 0; v < VEC; ++v) x[c][v] = 0.0f;
            }
            float *dst = Xs + ry * xrs + rxg * VEC;
#pragma unroll
            for (int c = 0; c < 3; ++c) st_vec<VEC>(dst + c * xcs, x[c]);
        }
        __syncthreads();
        // ---- phase B: stencil + operators after it, forward and backward, on tile + 1 -> GY
        for (int idx = tid; idx < GH * GWg; idx += NT) {
            const int ry = idx / GWg, rxg = idx - ry * GWg;
            const int y = y0 - 1 + ry, xg = xg0 - 1 + rxg;
            float gy[3][VEC], gd[3][VEC];
            const bool inb = y >= 0 && y < H && xg >= 0 && xg < Wg;
            if (inb) {
                const bool own = ry >= 1 && ry <= TH && rxg >= 1 && rxg <= TWg;
                const size_t off = (size_t)y * a.g.W + (size_t)xg * VEC;
                float x[3][VEC], m[3][VEC], g[3][VEC], ctr[3][VEC], lap[3][VEC];
                float sv[KMAX][3][VEC];
                ld_mask<VEC>(mask_b, a.mask_ch, plane, off, m);
                const int xgx = rxg - 1 + HGX;          // group index inside the X region
                const float *src = Xs + (ry + 1) * xrs + xgx * VEC;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    stencil_group<VEC>(src + c * xcs, xrs, -xgx * VEC, (XWg - xgx) * VEC, ctr[c], lap[c]);
#pragma unroll
                    for (int v = 0; v < VEC; ++v)
                        x[c][v] = sat01(blend(fmaf(p, lap[c][v], ctr[c][v]), ctr[c][v], m[c][v], has_mask));
                }
#pragma unroll
                for (int k = 1; k < KMAX; ++k) {
                    if (k > sp && k < n) {
#pragma unroll
                        for (int c = 0; c < 3; ++c)
#pragma unroll
                            for (int v = 0; v < VEC; ++v) sv[k][c][v] = x[c][v];
                        apply_op_vec<VEC>(a.ch.op[k], tabs[k], L, x, m, has_mask);
                    }
                }
                upstream<VEC>(a, go_b, tgt_b, plane, off, gl1, x, g, l1, own);
                if (out_b && own) st_px<VEC>(out_b, plane, off, x);
#pragma unroll
                for (int k = KMAX - 1; k >= 1; --k) {
                    if (k > sp && k < n)
                        bwd_op_vec<VEC>(a.ch.op[k], tabs[k], L, sv[k], m, has_mask, g, acc[k],
                                        Hist{hist_t + a.ch.hoff[k] * NT, NT}, own);
                }
                float accp = 0.0f;
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        blend_bwd(fmaf(p, lap[c][v], ctr[c][v]), ctr[c][v], m[c][v], has_mask, g[c][v], gy[c][v], gd[c][v]);
                        accp = fmaf(gy[c][v], lap[c][v], accp);
                    }
                if (own) acc_sharp += accp;
            } else {                                    // outside the image: no stencil output there
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int v = 0; v < VEC; ++v) { gy[c][v] = 0.0f; gd[c][v] = 0.0f; }
            }
            float *dst = GYs + ry * grs + rxg * VEC;
#pragma unroll
            for (int c = 0; c < 3; ++c) st_vec<VEC>(dst + c * gcs, gy[c]);
            if (has_mask) {
                float *dd = GDs + ry * grs + rxg * VEC;
#pragma unroll
                for (int c = 0; c < 3; ++c) st_vec<VEC>(dd + c * gcs, gd[c]);
            }
        }
        __syncthreads();
        // ---- phase C: transposed stencil, then the operators before it, on the tile interior
        if (gi_b != nullptr || sp > 0) {
            for (int idx = tid; idx < TH * TWg; idx += NT) {
                const int ly = idx / TWg, lxg = idx - ly * TWg;
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
                    if (has_mask) lds_vec<VEC>(GDs + c * gcs + (ly + 1) * grs + (lxg + 1) * VEC, gdv);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) g[c][v] = fmaf(p, lap[v], ctr[v]) + (has_mask ? gdv[v] : 0.0f);
                }
                if (sp > 0) {
                    float x[3][VEC], m[3][VEC];
                    float sv[KMAX][3][VEC];
                    ld_px<VEC>(img_b, plane, off, x);
                    ld_mask<VEC>(mask_b, a.mask_ch, plane, off, m);
#pragma unroll
                    for (int k = 0; k < KMAX - 1; ++k) {
                        if (k < sp) {
#pragma unroll
                            for (int c = 0; c < 3; ++c)
#pragma unroll
                                for (int v = 0; v < VEC; ++v) sv[k][c][v] = x[c][v];
                            apply_op_vec<VEC>(a.ch.op[k], tabs[k], L, x, m, has_mask);
                        }
                    }
#pragma unroll
                    for (int k = KMAX - 2; k >= 0; --k) {
                        if (k < sp)
                            bwd_op_vec<VEC>(a.ch.op[k], tabs[k], L, sv[k], m, has_mask, g, acc[k],
                                            Hist{hist_t + a.ch.hoff[k] * NT, NT}, true);
                    }
                }
                if (gi_b) st_px<VEC>(gi_b, plane, off, g);
            }
        }
    }

    // ---------------------------------------------------------------- CTA partials
    const int ntiles = a.g.ntiles;
    if (a.grad_params) {
        __syncthreads();                                 // all curve moments are in shared memory
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            if (k < n) {
                const int op = a.ch.op[k];
                if (op_is_curve(op)) {
                    // block-reduce the moments of this operator: warp w reduces slots w, w+8, ...
                    const int nslots = op_hist_floats(op);
                    const int lane = tid & 31, warp = tid >> 5;
                    float *hk = hist_mem + (size_t)a.ch.hoff[k] * NT;
                    for (int s = warp; s < nslots; s += NT / 32) {
                        float v = 0.0f;
#pragma unroll
                        for (int i = 0; i < NT / 32; ++i) v += hk[s * NT + lane + 32 * i];
                        v = warp_sum(v);
                        if (lane == 0) hk[s * NT] = v;    // slot total parked in element 0
                    }
                    __syncthreads();
                    const int ncur = op == OP_TONE ? 1 : 3;
                    if (tid < ncur) {
                        float mom[HIST], gk[MAX_L];
                        for (int s = 0; s < HIST; ++s) mom[s] = hk[(tid * HIST + s) * NT];
                        curve_param_grad(tabs[k] + tid * CT, L, mom, gk);
                        for (int i = 0; i < L; ++i) rowbuf[a.ch.poff[k] + tid * L + i] = gk[i];
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
        float *prow = a.part_gp + ((size_t)b * ntiles + tile) * a.pstride;
        for (int i = tid; i < a.pstride; i += NT) prow[i] = rowbuf[i];
    }
    if (a.l1_sum) {
        const float s = block_sum(l1, red);
        if (tid == 0) a.part_l1[(size_t)b * ntiles + tile] = s;
    }
    if (a.grad_params || a.l1_sum) {
        if (arrive_is_last(a.counters + b, (unsigned)ntiles, &last_flag)) {
            if (a.grad_params)
                reduce_columns(a.part_gp + (size_t)b * ntiles * a.pstride, ntiles, a.pstride,
                               a.grad_params + (size_t)b * a.pstride);
            if (a.l1_sum) {
                float v = 0.0f;
                for (int t = tid; t < ntiles; t += NT) v += __ldcg(a.part_l1 + (size_t)b * ntiles + t);
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

// =========================================================================================== host side
static thread_local char g_cuda_err[256] = "";
void set_cuda_error(cudaError_t e) { snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", cudaGetErrorName(e), cudaGetErrorString(e)); }
const char *last_cuda_error() { return g_cuda_err; }

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// workspace: [counters: B x u32][L1 partials: B x maxtiles][grad partials: B x maxtiles x pstride]
// upper bound of tiles per image for every tiling a launcher can pick (tiles are >= 32 x 32 px or >= 1024 px flat)
static inline size_t max_tiles(int H, int W) { return ((size_t)W / 32 + 2) * ((size_t)H / 32 + 2); }
size_t chain_workspace_bytes(int B, int H, int W, int pstride) {
    const size_t mt = max_tiles(H, W);
    return align_up((size_t)B * 4, 256) + align_up((size_t)B * mt * 4, 256) + align_up((size_t)B * mt * (size_t)(pstride > 0 ? pstride : 1) * 4, 256);
}
struct Workspace {
    unsigned int *counters;
    float *part_l1, *part_gp;
};
static Workspace carve(void *ws, int B, int H, int W) {
    const size_t mt = max_tiles(H, W);
    char *p = (char *)ws;
    Workspace w;
    w.counters = (unsigned int *)p; p += align_up((size_t)B * 4, 256);
    w.part_l1 = (float *)p; p += align_up((size_t)B * mt * 4, 256);
    w.part_gp = (float *)p;
    return w;
}

static int make_desc(int n_ops, const int *op_ids, const int *param_off, int L, int pstride, ChainDesc &d) {
    if (n_ops < 1 || n_ops > MAX_CHAIN || !op_ids || !param_off) return T2O_ERR_INVALID_ARG;
    if (L < 1 || L > MAX_L) return T2O_ERR_UNSUPPORTED;
    if (pstride < 0 || pstride > MAX_PSTRIDE) return T2O_ERR_UNSUPPORTED;
    d.n = n_ops; d.L = L; d.sharp = -1; d.hist_total = 0;
    for (int k = 0; k < MAX_CHAIN; ++k) { d.op[k] = OP_IDENTITY; d.poff[k] = 0; d.hoff[k] = 0; }
    for (int k = 0; k < n_ops; ++k) {
        const int op = op_ids[k];
        if (op == OP_INPAINT) return T2O_ERR_UNSUPPORTED;
        if (op < OP_IDENTITY || op >= OP_COUNT) return T2O_ERR_INVALID_ARG;
        if (param_off[k] < 0 || param_off[k] + op_num_params(op, L) > pstride) return T2O_ERR_INVALID_ARG;
        if (op == OP_SHARPNESS) {
            if (d.sharp >= 0) return T2O_ERR_UNSUPPORTED;
            d.sharp = k;
        }
        d.op[k] = op; d.poff[k] = param_off[k]; d.hoff[k] = d.hist_total;
        d.hist_total += op_hist_floats(op);
    }
    return T2O_OK;
}

static int pick_vec_flat(const void *const *ptrs, int nptr, size_t plane) {
    int vec = (plane % 4 == 0) ? 4 : (plane % 2 == 0 ? 2 : 1);
    for (int i = 0; i < nptr; ++i) {
        if (!ptrs[i]) continue;
        const uintptr_t a = (uintptr_t)ptrs[i];
        while (vec > 1 && (a % (vec * 4)) != 0) vec >>= 1;
    }
    return vec;
}

// flat tiling (no stencil): contiguous ranges of groups
static void geom_flat(Geom &g, int B, int H, int W, int vec) {
    g.B = B; g.H = H; g.W = W; g.Wg = 0; g.tiles_x = g.tiles_y = 0; g.TH = g.TWg = 0;
    const long long plane = (long long)H * W;
    g.ngroups = plane / vec;
    const long long total_px = plane * B;
    long long tile_px = total_px / (NUM_SMS * 8);
    const long long quantum = (long long)NT * vec;
    tile_px = (tile_px + quantum - 1) / quantum * quantum;
    if (tile_px < MIN_TILE_PX) tile_px = MIN_TILE_PX;
    if (tile_px > 8192) tile_px = 8192;
    if (tile_px < quantum) tile_px = quantum;
    g.tile_groups = (int)(tile_px / vec);
    g.ntiles = (int)((g.ngroups + g.tile_groups - 1) / g.tile_groups);
    if (g.ntiles < 1) g.ntiles = 1;
}

// 2-D tiling (one stencil in the chain); tile area >= MIN_TILE_PX unless the image is smaller
static void geom_2d(Geom &g, int B, int H, int W, int vec, int tw_px, int th) {
    g.B = B; g.H = H; g.W = W; g.Wg = W / vec; g.ngroups = 0; g.tile_groups = 0;
    int TWg = tw_px / vec;
    if (TWg > g.Wg) TWg = g.Wg;
    if (TWg < 1) TWg = 1;
    int TH = th;
    if (TH > H) TH = H;
    while ((long long)TH * TWg * vec < MIN_TILE_PX && TH < H) TH = TH * 2 < H ? TH * 2 : H;
    g.TH = TH; g.TWg = TWg;
    g.tiles_x = (g.Wg + TWg - 1) / TWg;
    g.tiles_y = (H + TH - 1) / TH;
    g.ntiles = g.tiles_x * g.tiles_y;
}

// Opt in to the dynamic shared memory a launch needs.  The 48 KB default limit counts static +
// dynamic bytes, so anything above 32 KB is configured explicitly (once per kernel and size).
template <typename K>
static int set_smem(K kernel, size_t bytes) {
    static size_t configured = 0;          // one instance per kernel type K ... but K is a pointer type
    (void)configured;
    if (bytes > 32 * 1024) {
        if (bytes > 227 * 1024) return T2O_ERR_UNSUPPORTED;
        T2O_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    }
    return T2O_OK;
}

int chain_forward(int n_ops, const int *op_ids, const int *param_off, const float *img, const float *mask, int mask_ch,
                  const float *params, int pstride, const float *target, float *out, float *l1_sum,
                  int B, int H, int W, int L, int flags, void *ws, size_t ws_bytes, cudaStream_t stream) {
    FwdArgs a;
    if ((flags & ~T2O_FLAG_RAW_PROCESS) != 0 || ((flags & T2O_FLAG_RAW_PROCESS) && n_ops != 1)) return T2O_ERR_INVALID_ARG;
    a.raw = (flags & T2O_FLAG_RAW_PROCESS) ? 1 : 0;
    int st = make_desc(n_ops, op_ids, param_off, L, pstride, a.ch);
    if (st != T2O_OK) return st;
    if (!img || B < 1 || H < 1 || W < 1 || (pstride > 0 && !params)) return T2O_ERR_INVALID_ARG;
    if (mask && mask_ch != 1 && mask_ch != 3) return T2O_ERR_INVALID_ARG;
    if (l1_sum && !target) return T2O_ERR_INVALID_ARG;
    if (!out && !l1_sum) return T2O_ERR_INVALID_ARG;
    if (l1_sum && (!ws || ws_bytes < chain_workspace_bytes(B, H, W, pstride))) return T2O_ERR_WORKSPACE;
    if (B > 65535) return T2O_ERR_UNSUPPORTED;
    Workspace w = carve(ws, B, H, W);
    a.img = img; a.mask = mask; a.params = params; a.target = target; a.out = out; a.l1_sum = l1_sum;
    a.part_l1 = w.part_l1; a.counters = w.counters; a.mask_ch = mask_ch; a.pstride = pstride;
    const size_t plane = (size_t)H * W;
    const void *ptrs[] = {img, mask, target, out};
    if (a.ch.sharp < 0) {
        const int vec = pick_vec_flat(ptrs, 4, plane);
        geom_flat(a.g, B, H, W, vec);
        dim3 grid(a.g.ntiles, B);
        if (vec == 4) chain_fwd_kernel<4, false><<<grid, NT, 0, stream>>>(a);
        else if (vec == 2) chain_fwd_kernel<2, false><<<grid, NT, 0, stream>>>(a);
        else chain_fwd_kernel<1, false><<<grid, NT, 0, stream>>>(a);
    } else {
        int vec = pick_vec_flat(ptrs, 4, plane);
        while (vec > 1 && W % vec != 0) vec >>= 1;
        geom_2d(a.g, B, H, W, vec, 256, 32);
        const size_t smem = (size_t)3 * (a.g.TH + 2) * (a.g.TWg + 2) * vec * sizeof(float);
        dim3 grid(a.g.ntiles, B);
        if (vec == 4) { st = set_smem(chain_fwd_kernel<4, true>, smem); if (st) return st; chain_fwd_kernel<4, true><<<grid, NT, smem, stream>>>(a); }
        else if (vec == 2) { st = set_smem(chain_fwd_kernel<2, true>, smem); if (st) return st; chain_fwd_kernel<2, true><<<grid, NT, smem, stream>>>(a); }
        else { st = set_smem(chain_fwd_kernel<1, true>, smem); if (st) return st; chain_fwd_kernel<1, true><<<grid, NT, smem, stream>>>(a); }
    }
    T2O_CUDA_OK(cudaGetLastError());
    return T2O_OK;
}

template <int VEC, int KMAX, bool SHARP>
static int launch_bwd(const BwdArgs &a, size_t smem, cudaStream_t stream) {
    int st = set_smem(chain_bwd_kernel<VEC, KMAX, SHARP>, smem);
    if (st) return st;
    dim3 grid(a.g.ntiles, a.g.B);
    chain_bwd_kernel<VEC, KMAX, SHARP><<<grid, NT, smem, stream>>>(a);
    T2O_CUDA_OK(cudaGetLastError());
    return T2O_OK;
}

int chain_backward(int n_ops, const int *op_ids, const int *param_off, const float *img, const float *mask, int mask_ch,
                   const float *params, int pstride, const float *grad_out, const float *target, const float *grad_l1,
                   float *grad_params, float *grad_img, float *out, float *l1_sum,
                   int B, int H, int W, int L, void *ws, size_t ws_bytes, cudaStream_t stream) {
    BwdArgs a;
    int st = make_desc(n_ops, op_ids, param_off, L, pstride, a.ch);
    if (st != T2O_OK) return st;
    if (!img || B < 1 || H < 1 || W < 1 || (pstride > 0 && !params)) return T2O_ERR_INVALID_ARG;
    if (mask && mask_ch != 1 && mask_ch != 3) return T2O_ERR_INVALID_ARG;
    if (!grad_out && (!target || !grad_l1)) return T2O_ERR_INVALID_ARG;
    if (l1_sum && !target) return T2O_ERR_INVALID_ARG;
    if (!grad_params && !grad_img) return T2O_ERR_INVALID_ARG;
    if (!ws || ws_bytes < chain_workspace_bytes(B, H, W, pstride)) return T2O_ERR_WORKSPACE;
    if (B > 65535) return T2O_ERR_UNSUPPORTED;
    Workspace w = carve(ws, B, H, W);
    a.img = img; a.mask = mask; a.params = params; a.grad_out = grad_out; a.target = target; a.grad_l1 = grad_l1;
    a.grad_params = grad_params; a.grad_img = grad_img; a.out = out; a.l1_sum = l1_sum;
    a.part_l1 = w.part_l1; a.part_gp = w.part_gp; a.counters = w.counters; a.mask_ch = mask_ch; a.pstride = pstride;
    const size_t plane = (size_t)H * W;
    const void *ptrs[] = {img, mask, target, out, grad_out, grad_img};
    const size_t hist_bytes = (size_t)a.ch.hist_total * NT * sizeof(float);
    int vec = pick_vec_flat(ptrs, 6, plane);
    const bool small_chain = n_ops <= 2;
    if (!small_chain && vec > 2) vec = 2;              // register budget: KMAX x 3 x VEC saved inputs
    if (a.ch.sharp < 0) {
        geom_flat(a.g, B, H, W, vec);
        if (small_chain) {
            if (vec == 4) return launch_bwd<4, 2, false>(a, hist_bytes, stream);
            if (vec == 2) return launch_bwd<2, 2, false>(a, hist_bytes, stream);
            return launch_bwd<1, 2, false>(a, hist_bytes, stream);
        }
        if (vec == 2) return launch_bwd<2, MAX_CHAIN, false>(a, hist_bytes, stream);
        return launch_bwd<1, MAX_CHAIN, false>(a, hist_bytes, stream);
    }
    while (vec > 1 && W % vec != 0) vec >>= 1;
    // largest tile whose X (tile+2), GY (tile+1) [, GD] regions and curve moments leave room for 2 CTAs / SM
    static const int cand[4][2] = {{32, 128}, {32, 64}, {16, 64}, {32, 32}};
    size_t smem = 0;
    for (int i = 0; i < 4; ++i) {
        geom_2d(a.g, B, H, W, vec, cand[i][1], cand[i][0]);
        const int hgx = vec == 1 ? 2 : 1;
        const size_t xs = (size_t)3 * (a.g.TH + 4) * (a.g.TWg + 2 * hgx) * vec;
        const size_t gs = (size_t)3 * (a.g.TH + 2) * (a.g.TWg + 2) * vec;
        smem = hist_bytes + (xs + gs * (mask ? 2 : 1)) * sizeof(float);
        if (smem <= 110 * 1024) break;
    }
    if (small_chain) {
        if (vec == 4) return launch_bwd<4, 2, true>(a, smem, stream);
        if (vec == 2) return launch_bwd<2, 2, true>(a, smem, stream);
        return launch_bwd<1, 2, true>(a, smem, stream);
    }
    if (vec == 2) return launch_bwd<2, MAX_CHAIN, true>(a, smem, stream);
    return launch_bwd<1, MAX_CHAIN, true>(a, smem, stream);
}

int l1_sum_launch(const float *pa, const float *pb, float *l1_sum, int B, long long n, void *ws, size_t ws_bytes,
                  cudaStream_t stream) {
    if (!pa || !pb || !l1_sum || B < 1 || n < 1) return T2O_ERR_INVALID_ARG;
    if (B > 65535) return T2O_ERR_UNSUPPORTED;
    L1Args a;
    const void *ptrs[] = {pa, pb};
    const int vec = pick_vec_flat(ptrs, 2, (size_t)n);
    long long tile = (n * B) / (NUM_SMS * 8);
    const long long quantum = (long long)NT * 4;
    tile = (tile + quantum - 1) / quantum * quantum;
    if (tile < 4096) tile = 4096;
    if (tile > 65536) tile = 65536;
    a.ntiles = (int)((n + tile - 1) / tile);
    const size_t need = align_up((size_t)B * 4, 256) + (size_t)B * a.ntiles * 4;
    if (!ws || ws_bytes < need) return T2O_ERR_WORKSPACE;
    a.a = pa; a.b = pb; a.l1_sum = l1_sum; a.n = n; a.tile_elems = tile;
    a.counters = (unsigned int *)ws;
    a.part = (float *)((char *)ws + align_up((size_t)B * 4, 256));
    dim3 grid(a.ntiles, B);
    if (vec == 4) l1_sum_kernel<4><<<grid, NT, 0, stream>>>(a);
    else if (vec == 2) l1_sum_kernel<2><<<grid, NT, 0, stream>>>(a);
    else l1_sum_kernel<1><<<grid, NT, 0, stream>>>(a);
    T2O_CUDA_OK(cudaGetLastError());
    return T2O_OK;
}

}  // namespace t2o
