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
// scalars) live in shared memory.  A sharpness operator turns the launch into a row pipeline down a
// column strip (chain_fwd_rows_kernel), so the chain still makes one pass over HBM.
// Reductions (L1, parameter gradients) are warp-shuffle -> shared -> one partial per CTA; the last
// CTA of each image sums the partials in a fixed order (deterministic, no float atomics).
#pragma once
#include <cstdio>
#include <cstring>

#include "t2o_step_kernels.cuh"

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
        T2O_CASE(OP_BNW) T2O_CASE(OP_HUE)
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

template <int VEC>
__device__ __forceinline__ float l1_px(const float (&x)[3][VEC], const float (&t)[3][VEC]) {
    float s = 0.0f;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int v = 0; v < VEC; ++v) s += fabsf(x[c][v] - t[c][v]);
    return s;
}

// Forward launch descriptor of one chain; host for uniform chains (t2o_chain.cu), device (once per CTA) for
// per-row chains.  param_off == nullptr: operator k reads its parameters at column k * slot.
T2O_HD int build_chain_desc(int n_ops, const int *op_ids, const int *param_off, int slot, int L, int pstride, ChainDesc &d) {
    if (n_ops < 1 || n_ops > MAX_CHAIN || !op_ids) return T2O_ERR_INVALID_ARG;
    if (L < 1 || L > MAX_L) return T2O_ERR_UNSUPPORTED;
    if (pstride < 0 || pstride > MAX_PSTRIDE) return T2O_ERR_UNSUPPORTED;
    d.n = n_ops; d.L = L; d.sharp = -1;
    for (int k = 0; k < MAX_CHAIN; ++k) { d.op[k] = OP_IDENTITY; d.poff[k] = 0; }
    for (int k = 0; k < n_ops; ++k) {
        const int op = op_ids[k];
        if (op == OP_INPAINT) return T2O_ERR_UNSUPPORTED;
        if (op < OP_IDENTITY || op >= OP_COUNT) return T2O_ERR_INVALID_ARG;
        const int po = param_off ? param_off[k] : k * slot;
        if (po < 0 || po + op_num_params(op, L) > pstride) return T2O_ERR_INVALID_ARG;
        if (op_is_stencil(op)) {                               // one stencil operator (sharpness or blur) per launch
            if (d.sharp >= 0) return T2O_ERR_UNSUPPORTED;
            d.sharp = k;
        }
        d.op[k] = op; d.poff[k] = po;
    }
    return T2O_OK;
}
// bit k: the input of operator k lies in [0, 1] (some operator before it clamped its output)
T2O_HD int chain_clamped_bits(const int *op, int n) {
    int bits = 0;
    for (int k = 1; k < n; ++k)
        if (op[k - 1] >= 0 || ((bits >> (k - 1)) & 1)) bits |= 1 << k;
    return bits;
}
// per-row chains: the row's descriptor, built by thread 0 (an invalid row becomes an identity chain and flags *status)
__device__ __forceinline__ void rows_build_chain_desc(const int *row_ops, int K, int slot, int L, int pstride,
                                                      unsigned int *status, int b, ChainDesc &d) {
    if (threadIdx.x == 0) {
        int ops[MAX_CHAIN];
        for (int k = 0; k < MAX_CHAIN; ++k) ops[k] = k < K ? row_ops[(size_t)b * K + k] : OP_IDENTITY;
        if (build_chain_desc(K, ops, nullptr, slot, L, pstride, d) != T2O_OK) {
            for (int k = 0; k < MAX_CHAIN; ++k) ops[k] = OP_IDENTITY;
            build_chain_desc(K, ops, nullptr, 0, L, pstride, d);
            if (status) atomicOr(status, 1u);
        }
    }
    __syncthreads();
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
    const int *row_ops;     // per-row chains (ROWS kernels): (B, rows_K) operator ids in device memory
    int rows_K, rows_slot;
    unsigned int *status;
};

template <int VEC, bool HM, bool ROWS>
__global__ void __launch_bounds__(NT) chain_fwd_kernel(const __grid_constant__ FwdArgs a) {
    __shared__ __align__(16) float tabs[MAX_CHAIN][TAB];
    __shared__ float red[32];
    __shared__ int last_flag;
    __shared__ ChainDesc rdesc;

    const int tid = threadIdx.x;
    const int b = blockIdx.y, tile = blockIdx.x;
    if constexpr (ROWS) {
        rows_build_chain_desc(a.row_ops, a.rows_K, a.rows_slot, a.ch.L, a.pstride, a.status, b, rdesc);
        if (rdesc.sharp >= 0) return;                       // rows with a stencil belong to chain_fwd_rows_kernel
    }
    const ChainDesc &ch = ROWS ? rdesc : a.ch;
    const int n = ch.n, L = ch.L;
    const size_t plane = (size_t)a.g.H * a.g.W;
    const float *img_b = a.img + (size_t)b * 3 * plane;
    const float *tgt_b = a.target ? a.target + (size_t)b * 3 * plane : nullptr;
    float *out_b = a.out ? a.out + (size_t)b * 3 * plane : nullptr;
    const float *mask_b = HM ? a.mask + (size_t)b * a.mask_ch * plane : nullptr;
    const bool raw = a.raw != 0;

    for (int k = tid >> 5; k < n; k += NT / 32) build_table_lanes<false>(ch.op[k], tid & 31, a.params + (size_t)b * a.pstride + ch.poff[k], L, tabs[k]);
    __syncthreads();

    float l1 = 0.0f;
    {
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
            for (int k = 0; k < n; ++k) apply_op_grp<VEC, 2, HM>(ch.op[k], tabs[k], L, x, m, raw);
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
            for (int k = 0; k < n; ++k) apply_op_vec<VEC, HM>(ch.op[k], tabs[k], L, x, m, raw);
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

// =========================================================================================== forward, one stencil
// Row pipeline down a column strip (same layout as step_sharp_kernel): every warp owns one image row per step;
// phase A writes X = (operators before the stencil)(img) of row rA into a ring of NW + 2 rows, phase B applies
// the stencil and the remaining operators to row rA - 1 and stores / scores it.  One pass over HBM; only the two
// halo rows per band and the two halo lanes per strip are recomputed.
struct FwdRowsArgs {
    ChainDesc ch;
    StepGeom g;
    const float *img, *mask, *params, *target;
    float *out, *l1_sum;
    float *part_l1;
    unsigned int *counters;
    int mask_ch, pstride;
    int raw;                // T2O_FLAG_RAW_PROCESS
    int clamped;            // bit k: the input of operator k lies in [0, 1]
    const int *row_ops;     // per-row chains (ROWS kernels)
    int rows_K, rows_slot;
    unsigned int *status;
};

// SP != 0: the chain is known at compile time (its ops_packed word, see t2o_step_kernels.cuh): the operator loops
// unroll and every operator switch folds away.
template <int VEC, bool HM, bool ROWS, unsigned int SP = 0u>
__global__ void __launch_bounds__(NT, 4) chain_fwd_rows_kernel(const __grid_constant__ FwdRowsArgs a) {
    static_assert(!(ROWS && SP), "per-row chains are dispatched at run time");
    constexpr int UNR = SP ? MAX_CHAIN : 1;
    constexpr int SPN = sp_count(SP), SPS = sp_sharp(SP), SPC = sp_clamped(SP);
    // ring of 2 NW + 2 rows: phase A of the next step never overwrites a row phase B of this step still reads,
    // so one barrier per step (after phase A) is enough
    constexpr int NW = NT / 32, RING = 2 * NW + 2;
    constexpr int ROWF = 34 * VEC, SLOTF = 3 * ROWF, RINGF = RING * SLOTF;
    extern __shared__ __align__(16) float dyn_smem[];           // the X ring
    __shared__ __align__(16) float tabs[MAX_CHAIN][TAB];
    __shared__ float red[32];
    __shared__ int last_flag;
    __shared__ ChainDesc rdesc;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int strip = chunk % a.g.strips, band = chunk / a.g.strips;
    if constexpr (ROWS) {
        rows_build_chain_desc(a.row_ops, a.rows_K, a.rows_slot, a.ch.L, a.pstride, a.status, b, rdesc);
        if (rdesc.sharp < 0) return;                        // rows without a stencil belong to chain_fwd_kernel
    }
    const ChainDesc &ch = ROWS ? rdesc : a.ch;
    const int n = SP ? SPN : ch.n, L = ch.L, sp = SP ? SPS : ch.sharp;
    const int H = a.g.H, W = a.g.W, Wg = a.g.Wg;
    const size_t plane = (size_t)H * W;
    const float *img_b = a.img + (size_t)b * 3 * plane;
    const float *tgt_b = a.target ? a.target + (size_t)b * 3 * plane : nullptr;
    float *out_b = a.out ? a.out + (size_t)b * 3 * plane : nullptr;
    const float *mask_b = HM ? a.mask + (size_t)b * a.mask_ch * plane : nullptr;
    const bool raw = a.raw != 0;

    const int HL = a.g.HL;
    const int gx = strip * a.g.IW - HL + lane;
    const bool lane_on = HL > 0 || lane < Wg;
    const bool col_ok = lane_on && gx >= 0 && gx < Wg;
    const bool interior = col_ok && lane >= HL && lane < 32 - HL;
    const int ya = band * a.g.HB;
    const int yb = ya + a.g.HB < H ? ya + a.g.HB : H;
    const size_t coff = (size_t)gx * VEC;

    float *Xc = dyn_smem + (1 + lane) * VEC;
    for (int i = tid; i < RINGF; i += NT) dyn_smem[i] = 0.0f;
    for (int k = tid >> 5; k < n; k += NT / 32) build_table_lanes<false>(ch.op[k], tid & 31, a.params + (size_t)b * a.pstride + ch.poff[k], L, tabs[k]);

    const int clamped = SP ? SPC : (ROWS ? chain_clamped_bits(ch.op, n) : a.clamped);
    float l1 = 0.0f;
    int rA = ya - 1 + warp, sA = warp;
    // L1 prefetches one phase ahead (this kernel keeps 4 CTAs per SM and a small ring, so the L1 has room; the
    // cp.async staging of the step kernels costs a CTA of occupancy here and measured slower)
    if (col_ok && rA >= 0 && rA < H && rA <= yb) prefetch_px(img_b, plane, (size_t)rA * W + coff);
    __syncthreads();
    const float p = tabs[sp][0];
    const bool blur = SP ? sp_blur(SP) : ch.op[sp] == OP_BLUR;  // which stencil
#pragma unroll 1
    for (int s = 0; s < a.g.steps; ++s) {
        const int rB = rA - 1;
        const bool do_b = interior && rB >= ya && rB < yb;
        // ---------------- phase A: X on row rA
        if (do_b && tgt_b) prefetch_px(tgt_b, plane, (size_t)rB * W + coff);                     // phase B's target row
        if (lane_on) {
            float x[3][VEC];
            if (col_ok && rA >= 0 && rA < H && rA <= yb) {
                ld_px<VEC>(img_b, plane, (size_t)rA * W + coff, x);
                if (sp > 0) {
                    float m[3][VEC];
                    ldm<VEC, HM>(mask_b, a.mask_ch, plane, (size_t)rA * W + coff, m);
#pragma unroll UNR
                    for (int k = 0; k < sp; ++k)
                        fwd_op_grp<VEC, HM>(SP ? packed_op(SP, k) : ch.op[k], tabs[k], L, x, m, (clamped >> k) & 1);
                }
            } else {
                zero3<VEC>(x);                                              // outside the image: the stencil's zero padding
            }
            float *dst = Xc + sA * SLOTF;
#pragma unroll
            for (int c = 0; c < 3; ++c) st_vec<VEC>(dst + c * ROWF, x[c]);
        }
        __syncthreads();
        // ---------------- phase B: stencil + remaining operators on row rA - 1
        {
            const int rN = rA + NW;                                         // phase A's row of the next step
            if (col_ok && rN >= 0 && rN < H && rN <= yb) prefetch_px(img_b, plane, (size_t)rN * W + coff);
            if (do_b) {
                const int sB = sA >= 1 ? sA - 1 : RING - 1;
                const int sU = sB >= 1 ? sB - 1 : RING - 1, sD = sB + 1 < RING ? sB + 1 : 0;
                const size_t off = (size_t)rB * W + coff;
                float x[3][VEC], m[3][VEC];
                ldm<VEC, HM>(mask_b, a.mask_ch, plane, off, m);
                const float *xb = Xc + sB * SLOTF, *xu = Xc + sU * SLOTF, *xd = Xc + sD * SLOTF;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    float ctr[VEC], lap[VEC];
                    stencil_any<VEC>(blur, xb + c * ROWF, xu + c * ROWF, xd + c * ROWF, ctr, lap);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        const float yv = fmaf(p, lap[v], ctr[v]);
                        x[c][v] = raw ? yv : sat01(blend<HM>(yv, ctr[v], m[c][v]));
                    }
                }
#pragma unroll UNR
                for (int k = sp + 1; k < n; ++k)
                    fwd_op_grp<VEC, HM>(SP ? packed_op(SP, k) : ch.op[k], tabs[k], L, x, m, (clamped >> k) & 1);
                if (tgt_b) {
                    float t[3][VEC];
                    ld_px<VEC>(tgt_b, plane, off, t);
                    l1 += l1_px<VEC>(x, t);
                }
                if (out_b) st_px<VEC>(out_b, plane, off, x);
            }
        }
        rA += NW;
        sA += NW;
        if (sA >= RING) sA -= RING;
    }

    if (a.l1_sum) {
        const float s = block_sum(l1, red);
        const int nchunks = a.g.nchunks;
        if (tid == 0) a.part_l1[(size_t)b * nchunks + chunk] = s;
        if (arrive_is_last(a.counters + b, (unsigned)nchunks, &last_flag)) {
            float v = 0.0f;
            for (int t = tid; t < nchunks; t += NT) v += __ldcg(a.part_l1 + (size_t)b * nchunks + t);
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

}  // namespace t2o
