// t2o_score.cu -- planner candidate scoring for sm_100a.
//
// Replaces the inner loop of the operation planner: for one state image, C x
//   executor.execute(img, op, None, specified_param=param)  ->  get_dist(pred, target, 'L1').item()
// (utils/beam_search.py:77-87 inside scipy's Nelder-Mead, and :229-237 for the beam candidates).
//
// One CTA owns a (state, target) tile.  The tile (+ a 1-pixel halo of the state for the
// sharpness stencil, zero-filled outside the image by the TMA unit = conv2d's zero padding) is
// staged ONCE in shared memory with cp.async.bulk.tensor (TMA) completing on an mbarrier; then
// every warp takes candidates of that state round-robin and evaluates operator + |out - target|
// over the whole tile from shared memory -- no block-level barrier inside the candidate loop, no
// HBM traffic per candidate.  Per-(candidate, tile) partials are summed by the last CTA of the
// state in a fixed order.  FP32-ALU bound once C is large; HBM traffic is 24 B/px per STATE.
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "t2o_common.cuh"
#include "../../include/t2o.h"

namespace t2o {

constexpr int SCORE_NT = 256;
constexpr int SCORE_NW = SCORE_NT / 32;

struct ScoreArgs {
    const float *states, *targets, *cand_param;
    const int *state_target, *cand_begin, *cand_op;
    const float *masks;         // (n_masks, mask_ch, H, W) or null
    const int *cand_mask;       // mask of candidate c, or -1 (masked launches only)
    int mask_ch;
    float *l1_sum, *part;
    unsigned int *counters;
    int S, T, C, H, W, L;
    int TH, TW;                 // tile rows / tile width in pixels (multiple of VEC)
    int tiles_x, ntiles;        // tiles per image
    int nsplit;                 // CTAs sharing one (state, tile): candidates are dealt round-robin
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tma_load_4d(void *dst, const CUtensorMap *tm, uint64_t *bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// One pixel group (VEC pixels x 3 channels) of a candidate: x = op(state group), sum += |x - target| element by element in
// (channel, pixel) order.  `src` points at the group's first pixel in the staged state tile (rows `spitch` floats apart,
// channels `cs` floats apart, one halo row / HX halo floats around it), `tsrc` at the same pixel of the staged target
// (channels `ct` floats apart).  Shared by both scorer kernels so that their sums agree to the bit.
// HM: the launch carries masks (Operator.execute's out * mask + img * (1 - mask), models/operators.py:129): `mptr` points at
// the group's first pixel in the candidate's mask (global memory, channels `mcs` floats apart; 0 for a 1-channel mask),
// or is null for a candidate without a mask (mask = 1: blend(y, x, 1) == y to the bit).
template <int VEC, bool HM>
__device__ __forceinline__ void score_group(float &sum, int op, const float *tab, int L, float p, const float *src, int spitch, int cs,
                                            const float *tsrc, int ct, const float *mptr, size_t mcs) {
    float x[3][VEC], t[3][VEC], m[3][VEC];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        if (HM && mptr) ld_vec<VEC>(mptr + c * mcs, m[c]);
        else {
#pragma unroll
            for (int v = 0; v < VEC; ++v) m[c][v] = 1.0f;
        }
    }
    if (op == OP_SHARPNESS) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float *row = src + c * cs;
            float ctr[VEC], up[VEC], dn[VEC];
            lds_vec<VEC>(row, ctr);
            lds_vec<VEC>(row - spitch, up);
            lds_vec<VEC>(row + spitch, dn);
            const float lf = row[-1], rt = row[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const float l = v > 0 ? ctr[v - 1] : lf;
                const float r = v < VEC - 1 ? ctr[v + 1] : rt;
                x[c][v] = sat01(blend<HM>(fmaf(p, laplace(ctr[v], up[v], dn[v], l, r), ctr[v]), ctr[v], m[c][v]));
            }
        }
    } else if (op == OP_BLUR) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float *row = src + c * cs, *ru = row - spitch, *rd = row + spitch;
            float ctr[VEC], up[VEC], dn[VEC];
            lds_vec<VEC>(row, ctr);
            lds_vec<VEC>(ru, up);
            lds_vec<VEC>(rd, dn);
            const float lf = row[-1], rt = row[VEC], ul = ru[-1], ur = ru[VEC], dl = rd[-1], dr = rd[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const float l = v > 0 ? ctr[v - 1] : lf, r = v < VEC - 1 ? ctr[v + 1] : rt;
                const float e = v > 0 ? up[v - 1] : ul, f = v < VEC - 1 ? up[v + 1] : ur;
                const float g = v > 0 ? dn[v - 1] : dl, h = v < VEC - 1 ? dn[v + 1] : dr;
                x[c][v] = sat01(blend<HM>(fmaf(p, blur_delta(ctr[v], (up[v] + dn[v]) + (l + r), (e + f) + (g + h)), ctr[v]), ctr[v], m[c][v]));
            }
        }
    } else {
#pragma unroll
        for (int c = 0; c < 3; ++c) lds_vec<VEC>(src + c * cs, x[c]);
        switch (op) {
#define T2O_CASE(OPC)                                                                                         \
    case OPC:                                                                                                 \
        _Pragma("unroll") for (int v = 0; v < VEC; ++v)                                                       \
            op_apply<HM>(OPC, tab, L, x[0][v], x[1][v], x[2][v], m[0][v], m[1][v], m[2][v]);                       \
        break;
            T2O_CASE(OP_BRIGHTNESS) T2O_CASE(OP_CONTRAST) T2O_CASE(OP_SATURATION) T2O_CASE(OP_COLOR)
            T2O_CASE(OP_TONE) T2O_CASE(OP_WHITE) T2O_CASE(OP_EXPOSURE) T2O_CASE(OP_WHITEBALANCE)
            T2O_CASE(OP_BNW) T2O_CASE(OP_HUE)
#undef T2O_CASE
            default: break;
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        lds_vec<VEC>(tsrc + c * ct, t[c]);
#pragma unroll
        for (int v = 0; v < VEC; ++v) sum += fabsf(x[c][v] - t[c][v]);
    }
}

// VEC = 4: W % 4 == 0, state rows padded by 4 floats each side (keeps 128-bit LDS aligned)
// VEC = 1: any W, 1 float each side
template <int VEC, bool USE_TMA, bool HM = false>
__global__ void __launch_bounds__(SCORE_NT) score_kernel(const __grid_constant__ CUtensorMap tm_state,
                                                         const __grid_constant__ CUtensorMap tm_target,
                                                         const __grid_constant__ ScoreArgs a) {
    extern __shared__ __align__(16) unsigned char dyn_raw[];
    __shared__ __align__(16) float wtab[SCORE_NW][TAB];
    __shared__ float wsum[SCORE_NW][SCORE_NW];
    __shared__ __align__(8) uint64_t bar;
    __shared__ int last_flag;

    constexpr int HX = VEC == 4 ? 4 : 1;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per_state = a.ntiles * a.nsplit;
    const int s = blockIdx.x / per_state;
    const int rem = blockIdx.x - s * per_state;
    const int tile = rem / a.nsplit, split = rem - tile * a.nsplit;
    const int ty = tile / a.tiles_x, tx = tile - ty * a.tiles_x;
    const int y0 = ty * a.TH, x0 = tx * a.TW;
    const int H = a.H, W = a.W, TH = a.TH, TW = a.TW;
    const int spitch = TW + 2 * HX, srows = TH + 2;
    const int t_idx = a.state_target ? a.state_target[s] : s % a.T;

    // a state whose candidates are all T2O_OP_SKIP (finished fits) is not even staged; every CTA of the state takes
    // the same decision, so the state's arrival counter stays untouched
    const int cbeg = a.cand_begin[s], cend = a.cand_begin[s + 1];
    {
        bool any = false;
        for (int ci = cbeg; ci < cend; ++ci) any |= a.cand_op[ci] != OP_SKIP;
        if (!any) return;
    }

    const bool coop = a.nsplit == 1 && cend - cbeg <= SCORE_NW;      // few candidates: all warps share each candidate's tile

    float *sS = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(dyn_raw) + 127) & ~(uintptr_t)127);
    float *sT = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(sS) + ((3 * srows * spitch * 4 + 127) & ~127));
    const size_t plane = (size_t)H * W;

    if constexpr (USE_TMA) {
        if (tid == 0) {
            mbar_init(&bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0) {
            const uint32_t bytes = (uint32_t)(3 * srows * spitch + 3 * TH * TW) * 4u;
            mbar_expect_tx(&bar, bytes);
            tma_load_4d(sS, &tm_state, &bar, x0 - HX, y0 - 1, 0, s);
            tma_load_4d(sT, &tm_target, &bar, x0, y0, 0, t_idx);
        }
        // few candidates: their tables are built while the tiles are in flight
        if (coop && warp < cend - cbeg && lane == 0) {
            const int op = a.cand_op[cbeg + warp];
            if (op != OP_SKIP) build_table<false>(op, a.cand_param + (size_t)(cbeg + warp) * T2O_MAX_OP_PARAMS, a.L, wtab[warp]);
        }
        // bounded wait: a broken descriptor must fail loudly, not hang the GPU
        bool done = false;
        for (int it = 0; it < (1 << 20) && !done; ++it) done = mbar_try_wait(&bar, 0);
        if (!done) __trap();
    } else {
        const float *sb = a.states + (size_t)s * 3 * plane;
        const float *tb = a.targets + (size_t)t_idx * 3 * plane;
        for (int i = tid; i < 3 * srows * spitch; i += SCORE_NT) {
            const int c = i / (srows * spitch), r = (i / spitch) % srows, col = i % spitch;
            const int y = y0 - 1 + r, x = x0 - HX + col;
            sS[i] = (y >= 0 && y < H && x >= 0 && x < W) ? __ldg(sb + c * plane + (size_t)y * W + x) : 0.0f;
        }
        for (int i = tid; i < 3 * TH * TW; i += SCORE_NT) {
            const int c = i / (TH * TW), r = (i / TW) % TH, col = i % TW;
            const int y = y0 + r, x = x0 + col;
            sT[i] = (y < H && x < W) ? __ldg(tb + c * plane + (size_t)y * W + x) : 0.0f;
        }
        if (coop && warp < cend - cbeg && lane == 0) {
            const int op = a.cand_op[cbeg + warp];
            if (op != OP_SKIP) build_table<false>(op, a.cand_param + (size_t)(cbeg + warp) * T2O_MAX_OP_PARAMS, a.L, wtab[warp]);
        }
        __syncthreads();
    }

    const int TWg = TW / VEC;
    const int ngroups = TH * TWg;
    // one lane's share of a candidate's |op(state) - target| over the tile: groups first, first + stride, ...
    const size_t mcs = HM && a.mask_ch == 3 ? plane : 0;
    // the mask image of candidate ci, or null
    auto cand_mask_img = [&](int ci) -> const float * {
        if (!HM || !a.cand_mask) return nullptr;
        const int mi = a.cand_mask[ci];
        return mi >= 0 ? a.masks + (size_t)mi * a.mask_ch * plane : nullptr;
    };
    auto tile_sum = [&](int op, const float *tab, int first, int stride, const float *mimg) -> float {
        float sum = 0.0f;
        const float p = tab[0];
        for (int gi = first; gi < ngroups; gi += stride) {
            const int ly = gi / TWg, lx = (gi - ly * TWg) * VEC;
            if (y0 + ly >= H || x0 + lx >= W) continue;       // ragged edge (W % VEC == 0)
            score_group<VEC, HM>(sum, op, tab, a.L, p, sS + (ly + 1) * spitch + HX + lx, spitch, srows * spitch,
                                 sT + ly * TW + lx, TH * TW, mimg ? mimg + (size_t)(y0 + ly) * W + x0 + lx : nullptr, mcs);
        }
        return sum;
    };

    if (coop) {
        // Few candidates (the Nelder-Mead rounds: at most one per operator and state): warp w builds the table of
        // candidate w, then ALL warps share every candidate's tile (warp w takes groups w*32 + lane, + 256, ...) and
        // the per-warp sums are added in warp order -- instead of one warp per candidate and the others idle.
        const int ncand = cend - cbeg;
        __syncthreads();                 // the tables (built above, while the tiles were in flight)
        for (int c = 0; c < ncand; ++c) {
            const int op = a.cand_op[cbeg + c];
            if (op == OP_SKIP) continue;
            const float sum = warp_sum(tile_sum(op, wtab[c], warp * 32 + lane, SCORE_NT, cand_mask_img(cbeg + c)));
            if (lane == 0) wsum[c][warp] = sum;
        }
        __syncthreads();
        if (tid < ncand && a.cand_op[cbeg + tid] != OP_SKIP) {
            float v = 0.0f;
#pragma unroll
            for (int w = 0; w < SCORE_NW; ++w) v += wsum[tid][w];
            a.part[(size_t)(cbeg + tid) * a.ntiles + tile] = v;
        }
    } else {
        float *tab = wtab[warp];
        for (int ci = cbeg + split * SCORE_NW + warp; ci < cend; ci += a.nsplit * SCORE_NW) {
            const int op = a.cand_op[ci];
            if (op == OP_SKIP) continue;
            __syncwarp();
            if (lane == 0) build_table<false>(op, a.cand_param + (size_t)ci * T2O_MAX_OP_PARAMS, a.L, tab);
            __syncwarp();
            const float sum = warp_sum(tile_sum(op, tab, lane, 32, cand_mask_img(ci)));
            if (lane == 0) a.part[(size_t)ci * a.ntiles + tile] = sum;
        }
    }

    // last CTA of this state sums the per-tile partials of all its candidates (fixed order)
    if (arrive_is_last(a.counters + s, (unsigned)per_state, &last_flag)) {
        for (int ci = cbeg + warp; ci < cend; ci += SCORE_NW) {
            float v = 0.0f;
            for (int t = lane; t < a.ntiles; t += 32) v += __ldcg(a.part + (size_t)ci * a.ntiles + t);
            v = warp_sum(v);
            if (lane == 0) a.l1_sum[ci] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------- host
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        else
            (void)cudaGetLastError();
    }
    return fn;
}

// (N, 3, H, W) fp32 tensor seen as a 4-D tiled map with box (bw, bh, 3, 1); out-of-bounds -> 0
static bool make_map(CUtensorMap *tm, const float *base, int N, int H, int W, int bw, int bh) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return false;
    cuuint64_t dims[4] = {(cuuint64_t)W, (cuuint64_t)H, 3, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)3 * H * W * 4};
    cuuint32_t box[4] = {(cuuint32_t)bw, (cuuint32_t)bh, 3, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    return fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void *)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline size_t score_max_tiles(int H, int W) { return ((size_t)W / 32 + 2) * ((size_t)H / 8 + 2); }

size_t score_workspace_bytes(int S, int C, int H, int W) {
    return COUNTER_REGION + (size_t)(C > 0 ? C : 1) * score_max_tiles(H, W) * 4;
}

int score_candidates(const float *states, int S, const float *targets, int T, const int *state_target,
                     const int *cand_begin, const int *cand_op, const float *cand_param, const int *cand_mask,
                     const float *masks, int n_masks, int mask_ch, int C, float *l1_sum,
                     int H, int W, int L, void *ws, size_t ws_bytes, cudaStream_t stream) {
    if (!states || !targets || !cand_begin || !cand_op || !cand_param || !l1_sum) return T2O_ERR_INVALID_ARG;
    const bool hm = masks != nullptr;
    if (hm && (!cand_mask || n_masks < 1 || (mask_ch != 1 && mask_ch != 3))) return T2O_ERR_INVALID_ARG;
    if (S < 1 || T < 1 || C < 0 || H < 1 || W < 1) return T2O_ERR_INVALID_ARG;
    if ((size_t)S * sizeof(unsigned int) > COUNTER_REGION) return T2O_ERR_UNSUPPORTED;
    if (L < 1 || L > MAX_L) return T2O_ERR_UNSUPPORTED;
    if (C == 0) return T2O_OK;
    if (!ws || ws_bytes < score_workspace_bytes(S, C, H, W)) return T2O_ERR_WORKSPACE;
    ScoreArgs a;
    a.states = states; a.targets = targets; a.cand_param = cand_param; a.state_target = state_target;
    a.cand_begin = cand_begin; a.cand_op = cand_op; a.l1_sum = l1_sum;
    a.masks = masks; a.cand_mask = hm ? cand_mask : nullptr; a.mask_ch = hm ? mask_ch : 0;
    a.counters = (unsigned int *)ws;
    a.part = (float *)((char *)ws + COUNTER_REGION);
    a.S = S; a.T = T; a.C = C; a.H = H; a.W = W; a.L = L;
    const bool aligned = ((uintptr_t)states % 16 == 0) && ((uintptr_t)targets % 16 == 0);
    const int vec = (W % 4 == 0 && (!hm || (uintptr_t)masks % 16 == 0)) ? 4 : 1;
    // tile: up to 128 px wide, 32 rows (state 3x34x136 + target 3x32x128 floats = 105 KB -> 2 CTAs / SM)
    int TW = W < 128 ? (W + vec - 1) / vec * vec : 128;
    int TH = H < 32 ? H : 32;          // (16- and 8-row tiles measured slower in every regime: the cost is per CTA)
    a.TH = TH; a.TW = TW;
    a.tiles_x = (W + TW - 1) / TW;
    a.ntiles = a.tiles_x * ((H + TH - 1) / TH);
    const int hx = vec == 4 ? 4 : 1;
    const size_t s_floats = (size_t)3 * (TH + 2) * (TW + 2 * hx), t_floats = (size_t)3 * TH * TW;
    const bool use_tma = vec == 4 && aligned && get_encode_fn() != nullptr && TW + 8 <= 256 && TH + 2 <= 256;
    const long long ctas = (long long)a.ntiles * S;
    int nsplit = 1;
    if (ctas < 3 * NUM_SMS) {
        const int want = (int)((3 * NUM_SMS + ctas - 1) / ctas);
        const int avg = (C + S - 1) / S;
        const int maxsplit = (avg + SCORE_NW - 1) / SCORE_NW;
        nsplit = want < maxsplit ? want : maxsplit;
        if (nsplit < 1) nsplit = 1;
    }
    a.nsplit = nsplit;
    const size_t smem = align_up(s_floats * 4, 128) + t_floats * 4 + 128;
    if (smem > 227 * 1024) return T2O_ERR_UNSUPPORTED;
    const long long grid = ctas * nsplit;
    if (grid > 0x7fffffffLL) return T2O_ERR_UNSUPPORTED;
    CUtensorMap tms, tmt;
    memset(&tms, 0, sizeof(tms)); memset(&tmt, 0, sizeof(tmt));
#define T2O_LAUNCH_SCORE(V, TMA, M)                                                                                              \
    do {                                                                                                                         \
        T2O_CUDA_OK(cudaFuncSetAttribute(score_kernel<V, TMA, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
        score_kernel<V, TMA, M><<<(unsigned)grid, SCORE_NT, smem, stream>>>(tms, tmt, a);                                         \
    } while (0)
    if (use_tma) {
        if (!make_map(&tms, states, S, H, W, TW + 8, TH + 2) || !make_map(&tmt, targets, T, H, W, TW, TH)) return T2O_ERR_NO_DEVICE;
        if (hm) T2O_LAUNCH_SCORE(4, true, true); else T2O_LAUNCH_SCORE(4, true, false);
    } else if (vec == 4) {
        if (hm) T2O_LAUNCH_SCORE(4, false, true); else T2O_LAUNCH_SCORE(4, false, false);
    } else {
        if (hm) T2O_LAUNCH_SCORE(1, false, true); else T2O_LAUNCH_SCORE(1, false, false);
    }
#undef T2O_LAUNCH_SCORE
    T2O_CUDA_OK(cudaGetLastError());
    return T2O_OK;
}

}  // namespace t2o
