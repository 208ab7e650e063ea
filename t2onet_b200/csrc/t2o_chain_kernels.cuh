// t2o_chain_kernels.cuh -- fused operator-chain kernels for sm_100a (templates; instantiated by t2o_chain*.cu).
//
//   chain_fwd_kernel   K x Operator.execute (+ per-image L1 to a target) in one pass over HBM
//   l1_sum_kernel      get_dist 'L1'
// (the fused forward + backward kernels live in t2o_step_kernels.cuh)
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

}  // namespace t2o
