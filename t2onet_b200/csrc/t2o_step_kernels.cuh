// t2o_step_kernels.cuh -- the fused "step" kernels for sm_100a: operator chain forward + L1 + backward in ONE
// launch (t2o_chain_backward).  Instantiated by t2o_step.cu.
//
// Replaces: K x Executor.execute (executors/executor.py:33-55 -> models/operators.py:112-131), the L1 of
// get_dist / the training loss (utils/beam_search.py:170-173, experiments/t2onet/train_seq2seqL1.py:85) and
// torch autograd through all of it.
//
// Design:
//   * a thread owns VEC consecutive pixels of the three planes (128-bit coalesced LDG/STG for VEC = 4);
//   * the chain is a runtime loop over operators with ONE uniform switch per loop iteration; the operator
//     inputs the backward sweep needs (the "tape") live in shared memory, not in registers, so the loop is
//     not unrolled and the kernel keeps 4-pixel groups with <= 128 registers;
//   * parameter gradients accumulate in registers, one slot per operator type (GradAcc), and are reduced
//     once per CTA: 32-value warp transpose-reductions -> shared -> one partial row per CTA -> the last CTA
//     of the image sums the partial rows in a fixed order (deterministic, no float atomics);
//   * a sharpness operator turns the kernel into a row pipeline down a column strip: every warp owns one
//     image row per step; X = (operators before the stencil)(img) is produced one step ahead into a ring of
//     R + 2 rows, the stencil + the operators after it + the loss + their backward run one row behind and
//     leave the gradient at the stencil output in a second ring, the transposed stencil and the backward of
//     the operators before the stencil run two rows behind, reading their tape from a third ring.  Nothing is
//     recomputed except the 4 halo rows per band and the 2 halo lanes per strip.
#pragma once
#include <cstdio>
#include <cstring>

#include "t2o_common.cuh"
#include "../../include/t2o.h"

namespace t2o {

// Threads per CTA are a template parameter NTH (256 or 192): NTH / 32 warps = image rows per pipeline step (a warp
// owns one row, a lane one group); the pipeline rings hold NTH / 32 + 2 rows.
constexpr int SNW_MAX = 8;

struct StepDesc {
    int n, L, sharp;              // operators, curve steps, index of the sharpness operator or -1
    int k_tone, k_color;          // chain position of the tone / color operator or -1 (their 1/S lives in the tables)
    int clamped;                  // bit k: the input of operator k is the clamped output of another operator (in [0, 1])
    unsigned int ops_packed;      // 4 bits per operator: op[k] + 1 (register-resident copy of op[] for the dispatch loops)
    int op[MAX_CHAIN];
    int poff[MAX_CHAIN];
    int slot_col[ACC_SLOTS];      // parameter column fed by each accumulator slot, or -1
};

struct StepGeom {
    int B, H, W;
    // flat (no stencil): an image is `ngroups` VEC-pixel groups, a CTA takes a contiguous chunk of them
    long long ngroups;
    int chunk_groups;
    int nchunks;                  // CTAs per image (both tilings)
    // row pipeline (one stencil)
    int Wg;                       // groups per image row
    int HL;                       // halo lanes on each side of a strip (0: one strip spans the image width, <= 32 groups)
    int IW;                       // interior lanes (groups) per strip
    int strips, bands, HB, steps;
};

struct StepArgs {
    StepDesc ch;
    StepGeom g;
    const float *img, *mask, *params, *grad_out, *target, *grad_l1;
    float *grad_params, *grad_img, *out, *l1_sum;
    float *part_l1, *part_gp;
    unsigned int *counters;
    int mask_ch, pstride;
    // per-row chains (ROWS kernels): operator ids (B, rows_K) in device memory, operator k of a row reads its
    // parameters at column k * rows_slot; *status gets bit 0 set if a row holds an invalid chain
    const int *row_ops;
    int rows_K, rows_slot;
    unsigned int *status;
};

// Launch descriptor of one chain; runs on the host for uniform chains (t2o_step.cu) and on the device, once per
// CTA, for per-row chains.  param_off == nullptr: operator k reads its parameters at column k * slot.
T2O_HD int build_step_desc(int n_ops, const int *op_ids, const int *param_off, int slot, int L, int pstride, StepDesc &d) {
    if (n_ops < 1 || n_ops > MAX_CHAIN || !op_ids) return T2O_ERR_INVALID_ARG;
    if (L < 1 || L > MAX_L) return T2O_ERR_UNSUPPORTED;
    if (pstride < 0 || pstride > MAX_PSTRIDE) return T2O_ERR_UNSUPPORTED;
    d.n = n_ops; d.L = L; d.sharp = -1; d.k_tone = -1; d.k_color = -1; d.clamped = 0; d.ops_packed = 0u;
    for (int k = 0; k < MAX_CHAIN; ++k) { d.op[k] = OP_IDENTITY; d.poff[k] = 0; }
    for (int i = 0; i < ACC_SLOTS; ++i) d.slot_col[i] = -1;
    unsigned seen = 0u;
    for (int k = 0; k < n_ops; ++k) {
        const int op = op_ids[k];
        if (op == OP_INPAINT) return T2O_ERR_UNSUPPORTED;
        if (op < OP_IDENTITY || op >= OP_COUNT) return T2O_ERR_INVALID_ARG;
        const int po = param_off ? param_off[k] : k * slot;
        if (po < 0 || po + op_num_params(op, L) > pstride) return T2O_ERR_INVALID_ARG;
        d.op[k] = op; d.poff[k] = po;
        d.ops_packed |= (unsigned)(op + 1) << (4 * k);
        // the input of operator k lies in [0, 1] if the last non-identity operator before it exists (its output is clamped)
        if (k > 0 && (d.op[k - 1] >= 0 || ((d.clamped >> (k - 1)) & 1))) d.clamped |= 1 << k;
        if (op < 0) continue;
        // one accumulator slot per operator type: a launch holds each type at most once (the binding splits)
        if ((seen >> op) & 1u) return T2O_ERR_UNSUPPORTED;
        seen |= 1u << op;
        switch (op) {
            case OP_SHARPNESS: case OP_BLUR:                    // one stencil operator per launch; both feed the `sharp` slot
                if (d.sharp >= 0) return T2O_ERR_UNSUPPORTED;
                d.sharp = k; d.slot_col[ACC_SHARP] = po;
                break;
            case OP_BNW: d.slot_col[ACC_BNW] = po; break;
            case OP_HUE: d.slot_col[ACC_HUE] = po; break;
            case OP_BRIGHTNESS: d.slot_col[ACC_BRIGHT] = po; break;
            case OP_CONTRAST: d.slot_col[ACC_CONTRAST] = po; break;
            case OP_SATURATION: d.slot_col[ACC_SATUR] = po; break;
            case OP_EXPOSURE: d.slot_col[ACC_EXPO] = po; break;
            case OP_WHITEBALANCE: for (int c = 0; c < 3; ++c) d.slot_col[ACC_WB + c] = po + c; break;
            case OP_TONE:
                d.k_tone = k;
                for (int i = 0; i < L; ++i) d.slot_col[ACC_TONE + i] = po + i;
                break;
            case OP_COLOR:
                d.k_color = k;
                for (int c = 0; c < 3; ++c)
                    for (int i = 0; i < L; ++i) d.slot_col[ACC_COLOR + c * MAX_L + i] = po + c * L + i;
                break;
            default: break;     // white: no parameter gradient (models/operators.py:510-512 ignores the parameter)
        }
    }
    return T2O_OK;
}

// operator id of chain position k from the packed register copy (two ALU instructions; a c[][] / shared-memory
// load indexed by k would put its latency in front of every dispatch branch)
__host__ __device__ __forceinline__ constexpr int packed_op(unsigned int ops_packed, int k) { return (int)((ops_packed >> (4 * k)) & 15u) - 1; }

// Chain-specialised instantiations: a kernel's SP template argument is the ops_packed word of ONE chain known at
// compile time (0 = any chain, dispatched at run time).  With SP != 0 the operator loops unroll completely, every
// operator switch folds away and the compiler schedules across operator boundaries.  The canonical FiveK order
// brightness, contrast, saturation, color, tone, sharpness (preprocess/gen_greedy_seqs_FiveK.py:39) is instantiated.
__host__ __device__ constexpr unsigned int pack_ops(int o0 = -1, int o1 = -1, int o2 = -1, int o3 = -1, int o4 = -1, int o5 = -1, int o6 = -1, int o7 = -1) {
    return (unsigned)(o0 + 1) | (unsigned)(o1 + 1) << 4 | (unsigned)(o2 + 1) << 8 | (unsigned)(o3 + 1) << 12 |
           (unsigned)(o4 + 1) << 16 | (unsigned)(o5 + 1) << 20 | (unsigned)(o6 + 1) << 24 | (unsigned)(o7 + 1) << 28;
}
__host__ __device__ constexpr int sp_count(unsigned int sp) { int n = 0; for (int k = 0; k < MAX_CHAIN; ++k) if ((sp >> (4 * k)) & 15u) n = k + 1; return n; }
__host__ __device__ constexpr int sp_sharp(unsigned int sp) { for (int k = 0; k < MAX_CHAIN; ++k) if (packed_op(sp, k) == OP_SHARPNESS || packed_op(sp, k) == OP_BLUR) return k; return -1; }
__host__ __device__ constexpr bool sp_blur(unsigned int sp) { return sp_sharp(sp) >= 0 && packed_op(sp, sp_sharp(sp)) == OP_BLUR; }
__host__ __device__ constexpr int sp_clamped(unsigned int sp) {
    int c = 0;
    for (int k = 1; k < MAX_CHAIN; ++k) if (packed_op(sp, k - 1) >= 0 || ((c >> (k - 1)) & 1)) c |= 1 << k;
    return c;
}
constexpr unsigned int SP_C6 = pack_ops(OP_BRIGHTNESS, OP_CONTRAST, OP_SATURATION, OP_COLOR, OP_TONE, OP_SHARPNESS);
constexpr unsigned int SP_P5 = pack_ops(OP_BRIGHTNESS, OP_CONTRAST, OP_SATURATION, OP_COLOR, OP_TONE);
constexpr unsigned int SP_S1 = pack_ops(OP_SHARPNESS);      // a single sharpness step: the launch behind Executor.execute(img, 6, ...)'s backward
constexpr unsigned int SP_B1 = pack_ops(OP_BLUR);

// ---------------------------------------------------------------- operator dispatch over one pixel group
// `cl`: the operator's input is known to lie in [0, 1] (lets the curve operators skip their input clamp)
template <int VEC, bool HM>
__device__ __forceinline__ void fwd_op_grp(int op, const float *tab, int L, float (&x)[3][VEC], const float (&m)[3][VEC], bool cl) {
#define T2O_CASE(OPC, CL)                                                                                      \
        _Pragma("unroll") for (int v = 0; v < VEC; ++v)                                                        \
            op_apply<HM, CL>(OPC, tab, L, x[0][v], x[1][v], x[2][v], m[0][v], m[1][v], m[2][v]);
    switch (op) {
        case OP_BRIGHTNESS: T2O_CASE(OP_BRIGHTNESS, false) break;
        case OP_CONTRAST: T2O_CASE(OP_CONTRAST, false) break;
        case OP_SATURATION: T2O_CASE(OP_SATURATION, false) break;
        case OP_COLOR: if (cl) { T2O_CASE(OP_COLOR, true) } else { T2O_CASE(OP_COLOR, false) } break;
        case OP_TONE: if (cl) { T2O_CASE(OP_TONE, true) } else { T2O_CASE(OP_TONE, false) } break;
        case OP_WHITE: T2O_CASE(OP_WHITE, false) break;
        case OP_EXPOSURE: T2O_CASE(OP_EXPOSURE, false) break;
        case OP_WHITEBALANCE: T2O_CASE(OP_WHITEBALANCE, false) break;
        case OP_BNW: T2O_CASE(OP_BNW, false) break;
        case OP_HUE: T2O_CASE(OP_HUE, false) break;
        default: break;
    }
#undef T2O_CASE
}

// every value of the group lies in [0, 1]: one unsigned compare on the bits of each pixel's max and min
template <int VEC>
__device__ __forceinline__ bool grp_in01(const float (&x)[3][VEC]) {
    bool ok = true;
#pragma unroll
    for (int v = 0; v < VEC; ++v) ok = ok && in01(max3(x[0][v], x[1][v], x[2][v])) && in01(min3(x[0][v], x[1][v], x[2][v]));
    return ok;
}

template <int VEC, bool HM>
__device__ __forceinline__ void bwd_op_grp(int op, const float *tab, int L, const float (&x)[3][VEC],
                                           const float (&m)[3][VEC], float (&g)[3][VEC], GradAcc &A, bool own, bool cl) {
#define T2O_CASE_NG(OPC, CL, NG)                                                                            \
        _Pragma("unroll") for (int v = 0; v < VEC; ++v)                                                     \
            pointwise_bwd<HM, CL, NG>(OPC, tab, L, x[0][v], x[1][v], x[2][v], m[0][v], m[1][v], m[2][v],    \
                                      g[0][v], g[1][v], g[2][v], A, own);
#define T2O_CASE(OPC, CL) T2O_CASE_NG(OPC, CL, false)
    // curves that map [0, 1] into [0, 1] (every curve with non-negative parameters) take the variant without clamp gates:
    // a branch that is uniform over the CTA (one image, one table)
    // brightness / saturation on an input that is not known to be clamped (the first operator of a chain): the gates can
    // only fire for pixels outside [0, 1]; a thread whose pixels are all inside takes the gate-free variant
#define T2O_CASE_HSV(OPC)                                                                                    \
        if (cl) { T2O_CASE(OPC, true) }                                                                      \
        else if (!HM && grp_in01<VEC>(x)) { T2O_CASE(OPC, true) }                                            \
        else { T2O_CASE(OPC, false) }
#define T2O_CASE_CURVE(OPC, CL)                                                                              \
        if (!HM && curve_in_range(OPC, tab)) { T2O_CASE_NG(OPC, CL, true) } else { T2O_CASE_NG(OPC, CL, false) }
    switch (op) {
        case OP_BRIGHTNESS: T2O_CASE_HSV(OP_BRIGHTNESS) break;
        case OP_CONTRAST: T2O_CASE(OP_CONTRAST, false) break;
        case OP_SATURATION: T2O_CASE_HSV(OP_SATURATION) break;
        case OP_COLOR: if (cl) { T2O_CASE_CURVE(OP_COLOR, true) } else { T2O_CASE_CURVE(OP_COLOR, false) } break;
        case OP_TONE: if (cl) { T2O_CASE_CURVE(OP_TONE, true) } else { T2O_CASE_CURVE(OP_TONE, false) } break;
        case OP_WHITE: T2O_CASE(OP_WHITE, false) break;
        case OP_EXPOSURE: T2O_CASE(OP_EXPOSURE, false) break;
        case OP_WHITEBALANCE: T2O_CASE(OP_WHITEBALANCE, false) break;
        case OP_BNW: T2O_CASE(OP_BNW, false) break;
        case OP_HUE: T2O_CASE(OP_HUE, false) break;
        default: break;
    }
#undef T2O_CASE
#undef T2O_CASE_NG
#undef T2O_CASE_CURVE
#undef T2O_CASE_HSV
}

template <int VEC, bool HM>
__device__ __forceinline__ void ldm(const float *mask_b, int mask_ch, size_t plane, size_t off, float (&m)[3][VEC]) {
    if constexpr (HM) ld_mask<VEC>(mask_b, mask_ch, plane, off, m);
}

// tape entries: three VEC-wide vectors `cstride` vectors apart
template <int VEC>
__device__ __forceinline__ void tape_st(typename VecT<VEC>::type *p, int cstride, const float (&x)[3][VEC]) {
#pragma unroll
    for (int c = 0; c < 3; ++c) st_vec<VEC>(reinterpret_cast<float *>(p + c * cstride), x[c]);
}
template <int VEC>
__device__ __forceinline__ void tape_ld(const typename VecT<VEC>::type *p, int cstride, float (&x)[3][VEC]) {
#pragma unroll
    for (int c = 0; c < 3; ++c) lds_vec<VEC>(reinterpret_cast<const float *>(p + c * cstride), x[c]);
}

// L1 prefetch of the three planes of a pixel group: the row pipeline issues it one phase ahead of the load, so
// the phases (which all warps of a CTA enter together, right after a barrier) do not start on a DRAM round trip
__device__ __forceinline__ void prefetch_px(const float *base, size_t plane, size_t off) {
    asm volatile("prefetch.global.L1 [%0];" ::"l"(base + off));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(base + plane + off));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(base + 2 * plane + off));
}

template <int VEC>
__device__ __forceinline__ void zero3(float (&x)[3][VEC]) {
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int v = 0; v < VEC; ++v) x[c][v] = 0.0f;
}

// upstream gradient of a pixel group: explicit grad_out, or the fused L1: gl1 * sign(out - target).
// `u` holds the group's values of the upstream operand (grad_out if there is one, else the target), already loaded.
template <int VEC>
__device__ __forceinline__ void upstream_grad_ld(bool has_go, const float (&u)[3][VEC], const float *tgt_b, size_t plane, size_t off,
                                                 float gl1, const float (&x)[3][VEC], float (&g)[3][VEC], float &l1, bool own) {
    if (has_go) {
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int v = 0; v < VEC; ++v) g[c][v] = u[c][v];
        if (tgt_b && own) {
            float t[3][VEC];
            ld_px<VEC>(tgt_b, plane, off, t);
            float s = 0.0f;
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int v = 0; v < VEC; ++v) s += fabsf(x[c][v] - t[c][v]);
            l1 += s;
        }
    } else {
        float s = 0.0f;
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const float d = x[c][v] - u[c][v];
                const float sg = __uint_as_float(__float_as_uint(gl1) ^ (__float_as_uint(d) & 0x80000000u));   // gl1 * sign(d)
                g[c][v] = d != 0.0f ? sg : 0.0f;
                s += fabsf(d);
            }
        if (own) l1 += s;
    }
}
template <int VEC>
__device__ __forceinline__ void upstream_grad(const float *go_b, const float *tgt_b, size_t plane, size_t off, float gl1,
                                              const float (&x)[3][VEC], float (&g)[3][VEC], float &l1, bool own) {
    float u[3][VEC];
    ld_px<VEC>(go_b ? go_b : tgt_b, plane, off, u);
    upstream_grad_ld<VEC>(go_b != nullptr, u, tgt_b, plane, off, gl1, x, g, l1, own);
}

// 5-point stencil of one plane around a VEC-pixel group of a ring row.  `row` points at the group's first
// float; the rows above / below are `up` / `dn`; row[-1] and row[VEC] are always addressable (zero pads).
template <int VEC>
__device__ __forceinline__ void stencil_ring(const float *row, const float *up, const float *dn,
                                             float (&ctr)[VEC], float (&lap)[VEC]) {
    float u[VEC], d[VEC];
    lds_vec<VEC>(row, ctr);
    lds_vec<VEC>(up, u);
    lds_vec<VEC>(dn, d);
    const float lf = row[-1], rt = row[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        const float l = v > 0 ? ctr[v - 1] : lf;
        const float r = v < VEC - 1 ? ctr[v + 1] : rt;
        lap[v] = laplace(ctr[v], u[v], d[v], l, r);
    }
}

// The blur's 9-point variant: the same plus the four diagonal neighbours (up[-1], up[VEC], dn[-1], dn[VEC] are addressable too).
template <int VEC>
__device__ __forceinline__ void stencil_ring9(const float *row, const float *up, const float *dn,
                                              float (&ctr)[VEC], float (&lap)[VEC]) {
    float u[VEC], d[VEC];
    lds_vec<VEC>(row, ctr);
    lds_vec<VEC>(up, u);
    lds_vec<VEC>(dn, d);
    const float lf = row[-1], rt = row[VEC], ul = up[-1], ur = up[VEC], dl = dn[-1], dr = dn[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        const float l = v > 0 ? ctr[v - 1] : lf, r = v < VEC - 1 ? ctr[v + 1] : rt;
        const float a = v > 0 ? u[v - 1] : ul, b = v < VEC - 1 ? u[v + 1] : ur;
        const float e = v > 0 ? d[v - 1] : dl, f = v < VEC - 1 ? d[v + 1] : dr;
        lap[v] = blur_delta(ctr[v], (u[v] + d[v]) + (l + r), (a + b) + (e + f));
    }
}
// S(x) of the launch's stencil operator (warp-uniform choice): y = x + p * S(x)
template <int VEC>
__device__ __forceinline__ void stencil_any(bool blur, const float *row, const float *up, const float *dn,
                                            float (&ctr)[VEC], float (&lap)[VEC]) {
    if (blur) stencil_ring9<VEC>(row, up, dn, ctr, lap);
    else stencil_ring<VEC>(row, up, dn, ctr, lap);
}

// ---------------------------------------------------------------- warp transpose-reductions
// On return lane l holds sum over the warp of v[l] (N = 32), or of v[l & 15] (N = 16): N - 1 shuffles.
template <int N>
__device__ __forceinline__ float warp_reduce_multi(float *v, int lane) {
#pragma unroll
    for (int off = N / 2; off >= 1; off >>= 1) {
        const bool hi = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            const float send = hi ? v[i] : v[i + off];
            const float keep = hi ? v[i + off] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    float r = v[0];
    if (N == 16) r += __shfl_xor_sync(0xffffffffu, r, 16);
    return r;
}

struct StepShared {
    float tabs[MAX_CHAIN][TAB];
    float wred[SNW_MAX][ACC_SLOTS];
    float tot[ACC_SLOTS];
    float rowbuf[MAX_PSTRIDE];
    float red[32];
    int last_flag;
    StepDesc rdesc;               // per-row chains: this row's descriptor (ROWS kernels)
};

// Per-row chains: thread 0 builds the row's descriptor; an invalid row becomes an identity chain and flags *status.
// Ends with a barrier; every thread then reads sh.rdesc.
__device__ __forceinline__ void rows_build_desc(const StepArgs &a, StepShared &sh, int b) {
    if (threadIdx.x == 0) {
        int ops[MAX_CHAIN];
        for (int k = 0; k < MAX_CHAIN; ++k) ops[k] = k < a.rows_K ? a.row_ops[(size_t)b * a.rows_K + k] : OP_IDENTITY;
        if (build_step_desc(a.rows_K, ops, nullptr, a.rows_slot, a.ch.L, a.pstride, sh.rdesc) != T2O_OK) {
            for (int k = 0; k < MAX_CHAIN; ++k) ops[k] = OP_IDENTITY;
            build_step_desc(a.rows_K, ops, nullptr, 0, a.ch.L, a.pstride, sh.rdesc);
            if (a.status) atomicOr(a.status, 1u);
        }
    }
    __syncthreads();
}

// CTA partials of the parameter gradients and the L1, then "the last CTA of the image finishes"
template <int NTH>
__device__ __forceinline__ void step_epilogue(const StepArgs &a, const StepDesc &ch, StepShared &sh, GradAcc &A, float l1, int b, int chunk) {
    constexpr int SNT = NTH, SNW = NTH / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nchunks = a.g.nchunks;
    if (a.grad_params) {
        float v[ACC_SLOTS];
        acc_to_slots(A, v);
        const float r0 = warp_reduce_multi<32>(v, lane);
        const float r1 = warp_reduce_multi<16>(v + 32, lane);
        sh.wred[warp][lane] = r0;
        if (lane < 16) sh.wred[warp][32 + lane] = r1;
        for (int i = tid; i < MAX_PSTRIDE; i += SNT) sh.rowbuf[i] = 0.0f;
        __syncthreads();
        if (tid < ACC_SLOTS) {
            float s = 0.0f;
#pragma unroll
            for (int w = 0; w < SNW; ++w) s += sh.wred[w][tid];
            sh.tot[tid] = s;
        }
        __syncthreads();
        if (tid < ACC_SLOTS) {
            const int col = ch.slot_col[tid];
            if (col >= 0) {
                float val = sh.tot[tid];
                if (tid < ACC_COLOR + 3 * MAX_L) {
                    const int c = tid / MAX_L;
                    val = curve_param_grad(sh.tabs[ch.k_color] + c * CT, ch.L, sh.tot + ACC_COLOR + c * MAX_L, tid - c * MAX_L);
                } else if (tid >= ACC_TONE && tid < ACC_TONE + MAX_L) {
                    val = curve_param_grad(sh.tabs[ch.k_tone], ch.L, sh.tot + ACC_TONE, tid - ACC_TONE);
                }
                sh.rowbuf[col] = val;
            }
        }
        __syncthreads();
        float *prow = a.part_gp + ((size_t)b * nchunks + chunk) * a.pstride;
        for (int i = tid; i < a.pstride; i += SNT) prow[i] = sh.rowbuf[i];
    }
    if (a.l1_sum) {
        const float s = block_sum(l1, sh.red);
        if (tid == 0) a.part_l1[(size_t)b * nchunks + chunk] = s;
    }
    if (a.grad_params || a.l1_sum) {
        if (arrive_is_last(a.counters + b, (unsigned)nchunks, &sh.last_flag)) {
            if (a.grad_params)
                reduce_columns(a.part_gp + (size_t)b * nchunks * a.pstride, nchunks, a.pstride,
                               a.grad_params + (size_t)b * a.pstride);
            if (a.l1_sum) {
                float v = 0.0f;
                for (int t = tid; t < nchunks; t += SNT) v += __ldcg(a.part_l1 + (size_t)b * nchunks + t);
                v = block_sum(v, sh.red);
                if (tid == 0) a.l1_sum[b] = v;
            }
        }
    }
}

// =========================================================================================== flat (no stencil)
template <int VEC, bool HM, int NTH, bool ROWS, unsigned int SP = 0u>
__global__ void __launch_bounds__(NTH, 2) step_flat_kernel(const __grid_constant__ StepArgs a) {
    static_assert(!(ROWS && SP), "per-row chains are dispatched at run time");
    constexpr int UNR = SP ? MAX_CHAIN : 1;
    constexpr int SPN = sp_count(SP), SPS = sp_sharp(SP), SPC = sp_clamped(SP);   // compile-time chain facts (SP != 0)
    using V = typename VecT<VEC>::type;
    constexpr int SNT = NTH;
    extern __shared__ __align__(16) float dyn_smem[];
    __shared__ __align__(16) StepShared sh;

    const int tid = threadIdx.x;
    const int b = blockIdx.y, chunk = blockIdx.x;
    if constexpr (ROWS) {
        rows_build_desc(a, sh, b);
        if (sh.rdesc.sharp >= 0) return;                    // rows with a stencil belong to step_sharp_kernel
    }
    const StepDesc &ch = ROWS ? sh.rdesc : a.ch;
    const int n = SP ? SPN : ch.n, L = ch.L;
    const size_t plane = (size_t)a.g.H * a.g.W;
    const float *img_b = a.img + (size_t)b * 3 * plane;
    const float *tgt_b = a.target ? a.target + (size_t)b * 3 * plane : nullptr;
    const float *go_b = a.grad_out ? a.grad_out + (size_t)b * 3 * plane : nullptr;
    float *out_b = a.out ? a.out + (size_t)b * 3 * plane : nullptr;
    float *gi_b = a.grad_img ? a.grad_img + (size_t)b * 3 * plane : nullptr;
    const float *mask_b = HM ? a.mask + (size_t)b * a.mask_ch * plane : nullptr;
    const float gl1 = a.grad_l1 ? a.grad_l1[b] : 0.0f;

    // (a warp per operator, a lane per curve record)
    for (int k = tid >> 5; k < n; k += SNT / 32) build_table_lanes<true>(ch.op[k], tid & 31, a.params + (size_t)b * a.pstride + ch.poff[k], L, sh.tabs[k]);
    __syncthreads();

    // dynamic shared memory: [staging slots: image 3 x SNT vectors, upstream 3 x SNT vectors][tape]
    constexpr int STGF = SNT * VEC;
    float *stg_i = dyn_smem + tid * VEC, *stg_u = stg_i + 3 * STGF;
    V *tape = reinterpret_cast<V *>(dyn_smem + 6 * STGF) + tid;   // input of operator k, plane c: tape[(k*3 + c) * SNT]
    const unsigned int opsp = SP ? SP : ch.ops_packed;
    const int clamped = SP ? SPC : ch.clamped;
    const float *up_b = go_b ? go_b : tgt_b;
    GradAcc A;
    acc_zero(A);
    float l1 = 0.0f;

    const long long g0 = (long long)chunk * a.g.chunk_groups;
    long long g1 = g0 + a.g.chunk_groups;
    if (g1 > a.g.ngroups) g1 = a.g.ngroups;
    // the image group of the next iteration and the upstream group of this one are in flight (cp.async into the
    // thread's own staging slots) while the forward sweep runs
    if (g0 + tid < g1) cp_async_px<VEC>(stg_i, STGF, img_b, plane, (size_t)(g0 + tid) * VEC);
    for (long long gi = g0 + tid; gi < g1; gi += SNT) {
        const size_t off = (size_t)gi * VEC;
        float x[3][VEC], m[3][VEC], g[3][VEC];
        cp_async_wait_all();
        lds_px<VEC>(stg_i, STGF, x);
        cp_async_px<VEC>(stg_u, STGF, up_b, plane, off);
        if (gi + SNT < g1) cp_async_px<VEC>(stg_i, STGF, img_b, plane, off + (size_t)SNT * VEC);
        ldm<VEC, HM>(mask_b, a.mask_ch, plane, off, m);
#pragma unroll UNR
        for (int k = 0; k < n; ++k) {
            tape_st<VEC>(tape + k * 3 * SNT, SNT, x);
            fwd_op_grp<VEC, HM>(packed_op(opsp, k), sh.tabs[k], L, x, m, (clamped >> k) & 1);
        }
        {
            float u[3][VEC];
            cp_async_wait_all();
            lds_px<VEC>(stg_u, STGF, u);
            upstream_grad_ld<VEC>(go_b != nullptr, u, tgt_b, plane, off, gl1, x, g, l1, true);
        }
        if (out_b) st_px<VEC>(out_b, plane, off, x);
#pragma unroll UNR
        for (int k = n - 1; k >= 0; --k) {
            tape_ld<VEC>(tape + k * 3 * SNT, SNT, x);
            bwd_op_grp<VEC, HM>(packed_op(opsp, k), sh.tabs[k], L, x, m, g, A, true, (clamped >> k) & 1);
        }
        if (gi_b) st_px<VEC>(gi_b, plane, off, g);
    }
    step_epilogue<NTH>(a, ch, sh, A, l1, b, chunk);
}

// =========================================================================================== row pipeline (one stencil)
// Shared-memory layout (compile-time strides): a ring row holds 3 planes of 34 groups (one zero pad group each
// side of the 32 lanes); the tape of the operators before the stencil holds, per operator 1 .. sp-1, RING rows of
// 3 x 32 vectors; the operators after the stencil keep a per-thread tape.
template <int VEC, bool HM, int NTH, bool ROWS, unsigned int SP = 0u, int MINB = 2>
__global__ void __launch_bounds__(NTH, MINB) step_sharp_kernel(const __grid_constant__ StepArgs a) {
    static_assert(!(ROWS && SP), "per-row chains are dispatched at run time");
    constexpr int UNR = SP ? MAX_CHAIN : 1;
    constexpr int SPN = sp_count(SP), SPS = sp_sharp(SP), SPC = sp_clamped(SP);   // compile-time chain facts (SP != 0)
    using V = typename VecT<VEC>::type;
    constexpr int SNT = NTH, SNW = NTH / 32, RING = SNW + 2;
    constexpr int ROWF = 34 * VEC;                 // floats of one ring row of one plane
    constexpr int SLOTF = 3 * ROWF;                // floats of one ring row
    constexpr int RINGF = RING * SLOTF;            // floats of one ring
    constexpr int TSLOT = 3 * 32;                  // vectors of one tape row
    extern __shared__ __align__(16) float dyn_smem[];
    __shared__ __align__(16) StepShared sh;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int strip = chunk % a.g.strips, band = chunk / a.g.strips;
    if constexpr (ROWS) {
        rows_build_desc(a, sh, b);
        if (sh.rdesc.sharp < 0) return;                     // rows without a stencil belong to step_flat_kernel
    }
    const StepDesc &ch = ROWS ? sh.rdesc : a.ch;
    const int n = SP ? SPN : ch.n, L = ch.L, sp = SP ? SPS : ch.sharp;
    const int H = a.g.H, W = a.g.W, Wg = a.g.Wg;
    const size_t plane = (size_t)H * W;
    const float *img_b = a.img + (size_t)b * 3 * plane;
    const float *tgt_b = a.target ? a.target + (size_t)b * 3 * plane : nullptr;
    const float *go_b = a.grad_out ? a.grad_out + (size_t)b * 3 * plane : nullptr;
    float *out_b = a.out ? a.out + (size_t)b * 3 * plane : nullptr;
    float *gi_b = a.grad_img ? a.grad_img + (size_t)b * 3 * plane : nullptr;
    const float *mask_b = HM ? a.mask + (size_t)b * a.mask_ch * plane : nullptr;
    const float gl1 = a.grad_l1 ? a.grad_l1[b] : 0.0f;

    // lane -> group of the strip; HL halo lanes on each side unless one strip spans the image
    const int HL = a.g.HL;
    const int gx = strip * a.g.IW - HL + lane;                     // group index in the image row
    const bool lane_on = HL > 0 || lane < Wg;                      // lanes that own a ring column
    const bool col_ok = lane_on && gx >= 0 && gx < Wg;             // ... whose column is inside the image
    const bool interior = col_ok && lane >= HL && lane < 32 - HL;  // ... and inside the strip proper
    const int ya = band * a.g.HB;
    const int yb = ya + a.g.HB < H ? ya + a.g.HB : H;

    // dynamic shared memory: [staging slots 3 x SNT vectors][rings][tapes]
    float *stg = dyn_smem + tid * VEC;                             // this thread's staging slots (planes STGF floats apart)
    constexpr int STGF = SNT * VEC;
    float *rings = dyn_smem + 3 * STGF;
    float *Xc = rings + (1 + lane) * VEC;                          // this lane's column in the X / GY / GD rings
    float *GYc = Xc + RINGF;
    float *GDc = GYc + RINGF;
    constexpr int NRING = HM ? 3 : 2;
    V *tapeP = reinterpret_cast<V *>(rings + NRING * RINGF) + lane;      // [(k-1) * RING + slot][c][32], k = 1 .. sp-1
    const int ntp = sp > 1 ? sp - 1 : 0;
    V *tapeQ = reinterpret_cast<V *>(rings + NRING * RINGF) + ntp * RING * TSLOT + tid;      // [(k-sp-1)][c][SNT]
    for (int i = tid; i < NRING * RINGF; i += SNT) rings[i] = 0.0f;
    // (a warp per operator, a lane per curve record)
    for (int k = tid >> 5; k < n; k += SNT / 32) build_table_lanes<true>(ch.op[k], tid & 31, a.params + (size_t)b * a.pstride + ch.poff[k], L, sh.tabs[k]);
    __syncthreads();

    const float p = sh.tabs[sp][0];
    const bool blur = SP ? sp_blur(SP) : ch.op[sp] == OP_BLUR;     // which stencil
    const bool need_c = gi_b != nullptr || sp > 0;
    const int clamped = SP ? SPC : ch.clamped;
    const unsigned int opsp = SP ? SP : ch.ops_packed;
    const float *up_b = go_b ? go_b : tgt_b;                       // phase B's upstream operand: grad_out, else the target
    const size_t coff = (size_t)gx * VEC;
    GradAcc A;
    acc_zero(A);
    float l1 = 0.0f;

    // Global rows travel through the staging slots as asynchronous copies: the image row of phase A is requested a
    // whole phase C ahead, the upstream row of phase B a phase A ahead.  A thread reads back only its own slots
    // (cp.async.wait_all, no barrier), and requests the next row only after it has read the previous one.
    int rA = ya - 2 + warp, sA = warp;                             // row produced in phase A and its ring slot
    if (col_ok && rA >= 0 && rA < H && rA <= yb + 1) cp_async_px<VEC>(stg, STGF, img_b, plane, (size_t)rA * W + coff);
#pragma unroll 1
    for (int s = 0; s < a.g.steps; ++s) {
        const int rB = rA - 1;
        const bool in_b = lane_on && rB >= ya - 1 && rB <= yb;                       // phase B works on this lane's row
        const bool img_b_ok = in_b && col_ok && rB >= 0 && rB < H;                   // ... and the row is inside the image
        // ---------------- phase A: X = (operators before the stencil)(img) on row rA
        const bool in_a = col_ok && rA >= 0 && rA < H && rA <= yb + 1;               // phase A has an image row to work on
        float xa[3][VEC];
        if (in_a) {
            cp_async_wait_all();
            lds_px<VEC>(stg, STGF, xa);
        }
        if (img_b_ok) cp_async_px<VEC>(stg, STGF, up_b, plane, (size_t)rB * W + coff);   // phase B's upstream row
        if (lane_on) {
            float (&x)[3][VEC] = xa;
            if (in_a) {
                if (sp > 0) {
                    float m[3][VEC];
                    ldm<VEC, HM>(mask_b, a.mask_ch, plane, (size_t)rA * W + coff, m);
#pragma unroll UNR
                    for (int k = 0; k < sp; ++k) {
                        if (k > 0 && interior) tape_st<VEC>(tapeP + ((k - 1) * RING + sA) * TSLOT, 32, x);
                        fwd_op_grp<VEC, HM>(packed_op(opsp, k), sh.tabs[k], L, x, m, (clamped >> k) & 1);
                    }
                }
            } else {                                               // outside the image: the stencil's zero padding
                zero3<VEC>(x);
            }
            float *dst = Xc + sA * SLOTF;
#pragma unroll
            for (int c = 0; c < 3; ++c) st_vec<VEC>(dst + c * ROWF, x[c]);
        }
        __syncthreads();
        // ---------------- phase B: stencil, operators after it, loss, their backward on row rA - 1 -> GY ring
        if (in_b) {
            const int sB = sA >= 1 ? sA - 1 : RING - 1;
            float gy[3][VEC], gd[3][VEC];
            if (img_b_ok) {
                const bool own = interior && rB >= ya && rB < yb;
                const size_t off = (size_t)rB * W + coff;
                const int sU = sB >= 1 ? sB - 1 : RING - 1, sD = sB + 1 < RING ? sB + 1 : 0;
                float x[3][VEC], m[3][VEC], g[3][VEC], ctr[3][VEC], lap[3][VEC];
                ldm<VEC, HM>(mask_b, a.mask_ch, plane, off, m);
                const float *xb = Xc + sB * SLOTF, *xu = Xc + sU * SLOTF, *xd = Xc + sD * SLOTF;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    stencil_any<VEC>(blur, xb + c * ROWF, xu + c * ROWF, xd + c * ROWF, ctr[c], lap[c]);
#pragma unroll
                    for (int v = 0; v < VEC; ++v)
                        x[c][v] = sat01(blend<HM>(fmaf(p, lap[c][v], ctr[c][v]), ctr[c][v], m[c][v]));
                }
#pragma unroll UNR
                for (int k = sp + 1; k < n; ++k) {
                    tape_st<VEC>(tapeQ + (k - sp - 1) * 3 * SNT, SNT, x);
                    fwd_op_grp<VEC, HM>(packed_op(opsp, k), sh.tabs[k], L, x, m, (clamped >> k) & 1);
                }
                {
                    float u[3][VEC];
                    cp_async_wait_all();
                    lds_px<VEC>(stg, STGF, u);
                    upstream_grad_ld<VEC>(go_b != nullptr, u, tgt_b, plane, off, gl1, x, g, l1, own);
                }
                if (out_b && own) st_px<VEC>(out_b, plane, off, x);
#pragma unroll UNR
                for (int k = n - 1; k > sp; --k) {
                    tape_ld<VEC>(tapeQ + (k - sp - 1) * 3 * SNT, SNT, x);
                    bwd_op_grp<VEC, HM>(packed_op(opsp, k), sh.tabs[k], L, x, m, g, A, own, (clamped >> k) & 1);
                }
                float accp = 0.0f;
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        blend_bwd<HM>(fmaf(p, lap[c][v], ctr[c][v]), ctr[c][v], m[c][v], g[c][v], gy[c][v], gd[c][v]);
                        accp = fmaf(gy[c][v], lap[c][v], accp);
                    }
                if (own) A.sharp += accp;
            } else {                                               // outside the image: no stencil output there
                zero3<VEC>(gy);
                zero3<VEC>(gd);
            }
            if (need_c) {
                float *dst = GYc + sB * SLOTF;
#pragma unroll
                for (int c = 0; c < 3; ++c) st_vec<VEC>(dst + c * ROWF, gy[c]);
                if constexpr (HM) {
                    float *dd = GDc + sB * SLOTF;
#pragma unroll
                    for (int c = 0; c < 3; ++c) st_vec<VEC>(dd + c * ROWF, gd[c]);
                }
            }
        }
        __syncthreads();
        // ---------------- phase C: transposed stencil, backward of the operators before it, on row rA - 2.
        // No barrier closes it: phase A of the next step writes ring / tape slots that phase C either does not read
        // (X ring) or reads from the same thread earlier in program order (tape slot sC == next sA); the GY ring is
        // rewritten only after the next step's first barrier.
        {
            const int rN = rA + SNW;                               // phase A's image row of the next step
            if (col_ok && rN >= 0 && rN < H && rN <= yb + 1) cp_async_px<VEC>(stg, STGF, img_b, plane, (size_t)rN * W + coff);
        }
        if (need_c) {
            const int rC = rA - 2;
            if (interior && rC >= ya && rC < yb) {
                const int sC = sA >= 2 ? sA - 2 : sA - 2 + RING;
                const int sU = sC >= 1 ? sC - 1 : RING - 1, sD = sC + 1 < RING ? sC + 1 : 0;
                const size_t off = (size_t)rC * W + coff;
                float g[3][VEC];
                const float *yc = GYc + sC * SLOTF, *yu = GYc + sU * SLOTF, *yd = GYc + sD * SLOTF;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    float ctr[VEC], lap[VEC];
                    stencil_any<VEC>(blur, yc + c * ROWF, yu + c * ROWF, yd + c * ROWF, ctr, lap);
                    float gdv[VEC];
                    if constexpr (HM) lds_vec<VEC>(GDc + sC * SLOTF + c * ROWF, gdv);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) g[c][v] = fmaf(p, lap[v], ctr[v]) + (HM ? gdv[v] : 0.0f);
                }
                if (sp > 0) {
                    float x[3][VEC], m[3][VEC];
                    ldm<VEC, HM>(mask_b, a.mask_ch, plane, off, m);
#pragma unroll UNR
                    for (int k = sp - 1; k >= 1; --k) {
                        tape_ld<VEC>(tapeP + ((k - 1) * RING + sC) * TSLOT, 32, x);
                        bwd_op_grp<VEC, HM>(packed_op(opsp, k), sh.tabs[k], L, x, m, g, A, true, (clamped >> k) & 1);
                    }
                    ld_px<VEC>(img_b, plane, off, x);
                    bwd_op_grp<VEC, HM>(packed_op(opsp, 0), sh.tabs[0], L, x, m, g, A, true, false);
                }
                if (gi_b) st_px<VEC>(gi_b, plane, off, g);
            }
        }
        rA += SNW;
        sA += SNW;
        if (sA >= RING) sA -= RING;
    }
    cp_async_wait_all();
    step_epilogue<NTH>(a, ch, sh, A, l1, b, chunk);
}

// Opt in to the dynamic shared memory a launch needs.
template <typename K>
static int step_set_smem(K kernel, size_t bytes) {
    if (bytes > 200 * 1024) return T2O_ERR_UNSUPPORTED;
    if (bytes > 40 * 1024)
        T2O_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return T2O_OK;
}

}  // namespace t2o
