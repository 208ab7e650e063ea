// t2o_ssim.cu -- SSIM of image pairs in one pass over HBM (sm_100a).
//
// Replaces: utils/ssim/__init__.py:19-41 (`_ssim`, the pytorch_ssim recipe the reference's evaluation uses,
// utils/eval.py:57-60): five depthwise 11x11 Gaussian (sigma 1.5, zero padding 5) convolutions -- of img1, img2,
// img1^2, img2^2, img1*img2 -- followed by ~15 elementwise launches and a mean.  Here persistent CTAs (three per SM)
// walk over (plane, tile) items of one image: per item a (28+10) x (64+16) patch of both images is staged in shared
// memory with zero-filling cp.async copies, the Gaussian runs separably (the reference's 2-D window IS the outer
// product of its 1-D window, :13-14) -- horizontally on the five products with a register sliding window, then
// vertically -- and the SSIM map is evaluated and summed in registers: 8 B per pixel and plane of HBM traffic, nothing
// written but one partial per CTA.  Per-image sums are finished by the image's last CTA in a fixed order.
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>

#include "../../include/t2o.h"
#include "t2o_common.cuh"

namespace t2o {

constexpr int SS_R = 5, SS_K = 2 * SS_R + 1;        // window radius / taps
constexpr int SS_TH = 28, SS_TW = 64;               // output tile: 73 KB of shared memory, three CTAs per SM
constexpr int SS_RG = SS_TH / 4;                    // output rows per thread in the vertical pass (256 threads = 64 columns x 4)
constexpr int SS_LM = 8;                            // patch column 0 is image column x0 - 8 (16-byte aligned when W % 4 == 0)
constexpr int SS_PH = SS_TH + 2 * SS_R, SS_PITCH = SS_TW + 2 * SS_LM;   // patch rows / floats per patch row (80)
constexpr int SS_NT = 256;

struct SsimArgs {
    const float *a, *b;
    float *ssim_sum, *part;
    unsigned int *counters;
    int C, H, W, tiles_x, ntiles, nchunks;
    float w[SS_K];
};

// V4: W % 4 == 0 and 16-byte aligned images -> the patch is staged with 16-byte cp.async copies (a 4-pixel vector is
// either wholly inside the image or wholly outside); otherwise 4-byte copies.  Zero fill (src-size 0) outside the image
// = conv2d's zero padding.  A CTA walks over items (plane, tile) of ONE image and keeps its sum in registers.
template <bool V4>
__global__ void __launch_bounds__(SS_NT, 3) ssim_kernel(const __grid_constant__ SsimArgs a) {
    extern __shared__ __align__(16) float sm[];
    float *sA = sm, *sB = sA + SS_PH * SS_PITCH;
    float *sH = sB + SS_PH * SS_PITCH;              // [5][SS_PH][SS_TW]
    __shared__ float red[32];
    __shared__ int last_flag;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int chunk = blockIdx.x, b = blockIdx.y;
    const int H = a.H, W = a.W;
    const size_t plane = (size_t)H * W;
    const int items = a.ntiles * a.C;
    float sum = 0.0f;

    for (int item = chunk; item < items; item += a.nchunks) {
        const int c = item / a.ntiles, tile = item - c * a.ntiles;
        const int ty = tile / a.tiles_x, tx = tile - ty * a.tiles_x;
        const int y0 = ty * SS_TH, x0 = tx * SS_TW;
        const float *pa = a.a + ((size_t)b * a.C + c) * plane, *pb = a.b + ((size_t)b * a.C + c) * plane;
        // ---- stage the patch of both images, a warp per patch row, everything in flight at once
        for (int r = warp; r < SS_PH; r += SS_NT / 32) {
            const int y = y0 - SS_R + r;
            const bool row_in = y >= 0 && y < H;
            const size_t roff = (size_t)(row_in ? y : 0) * W;
            if constexpr (V4) {
                if (lane < SS_PITCH / 4) {
                    const int x = x0 - SS_LM + 4 * lane;
                    const bool in = row_in && x >= 0 && x < W;
                    const size_t off = in ? roff + x : 0;
                    const unsigned int da = (unsigned int)__cvta_generic_to_shared(sA + r * SS_PITCH + 4 * lane);
                    const unsigned int db = (unsigned int)__cvta_generic_to_shared(sB + r * SS_PITCH + 4 * lane);
                    const unsigned int n = in ? 16u : 0u;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(da), "l"(pa + off), "r"(n) : "memory");
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(db), "l"(pb + off), "r"(n) : "memory");
                }
            } else {
#pragma unroll
                for (int pc = SS_LM - SS_R + lane; pc < SS_LM + SS_TW + SS_R; pc += 32) {
                    const int x = x0 - SS_LM + pc;
                    const bool in = row_in && x >= 0 && x < W;
                    const size_t off = in ? roff + x : 0;
                    const unsigned int da = (unsigned int)__cvta_generic_to_shared(sA + r * SS_PITCH + pc);
                    const unsigned int db = (unsigned int)__cvta_generic_to_shared(sB + r * SS_PITCH + pc);
                    const unsigned int n = in ? 4u : 0u;
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(da), "l"(pa + off), "r"(n) : "memory");
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(db), "l"(pb + off), "r"(n) : "memory");
                }
            }
        }
        cp_async_wait_all();
        __syncthreads();            // the patch is complete -- and every thread has left the previous item's vertical pass
        // ---- horizontal pass: item = (patch row, 4 output columns c4 .. c4+3): inputs are patch columns c4+3 .. c4+16,
        // read as five 128-bit loads per image (consecutive lanes, consecutive 16-byte words: conflict-free)
        for (int it = tid; it < SS_PH * (SS_TW / 4); it += SS_NT) {
            const int r = it / (SS_TW / 4), c4 = (it - r * (SS_TW / 4)) * 4;
            float va[20], vb[20];
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const float4 ta = *reinterpret_cast<const float4 *>(sA + r * SS_PITCH + c4 + 4 * k);
                const float4 tb = *reinterpret_cast<const float4 *>(sB + r * SS_PITCH + c4 + 4 * k);
                va[4 * k] = ta.x; va[4 * k + 1] = ta.y; va[4 * k + 2] = ta.z; va[4 * k + 3] = ta.w;
                vb[4 * k] = tb.x; vb[4 * k + 1] = tb.y; vb[4 * k + 2] = tb.z; vb[4 * k + 3] = tb.w;
            }
            float o[5][4];
#pragma unroll
            for (int q = 0; q < 5; ++q)
#pragma unroll
                for (int j = 0; j < 4; ++j) o[q][j] = 0.0f;
#pragma unroll
            for (int k = 0; k < 4 + 2 * SS_R; ++k) {
                const float x = va[k + SS_LM - SS_R], y = vb[k + SS_LM - SS_R], xx = x * x, yy = y * y, xy = x * y;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int t = k - j;                 // tap index of input k for output j
                    if (t >= 0 && t < SS_K) {
                        const float w = a.w[t];
                        o[0][j] = fmaf(w, x, o[0][j]); o[1][j] = fmaf(w, y, o[1][j]);
                        o[2][j] = fmaf(w, xx, o[2][j]); o[3][j] = fmaf(w, yy, o[3][j]); o[4][j] = fmaf(w, xy, o[4][j]);
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < 5; ++q)
                *reinterpret_cast<float4 *>(sH + (q * SS_PH + r) * SS_TW + c4) = make_float4(o[q][0], o[q][1], o[q][2], o[q][3]);
        }
        __syncthreads();
        // ---- vertical pass + SSIM map: thread = (output column, SS_RG output rows)
        {
            const int col = tid % SS_TW, r8 = (tid / SS_TW) * SS_RG;
            float o[5][SS_RG];
#pragma unroll
            for (int q = 0; q < 5; ++q)
#pragma unroll
                for (int j = 0; j < SS_RG; ++j) o[q][j] = 0.0f;
#pragma unroll
            for (int k = 0; k < SS_RG + 2 * SS_R; ++k) {
                float v[5];
#pragma unroll
                for (int q = 0; q < 5; ++q) v[q] = sH[(q * SS_PH + r8 + k) * SS_TW + col];
#pragma unroll
                for (int j = 0; j < SS_RG; ++j) {
                    const int t = k - j;
                    if (t >= 0 && t < SS_K) {
                        const float w = a.w[t];
#pragma unroll
                        for (int q = 0; q < 5; ++q) o[q][j] = fmaf(w, v[q], o[q][j]);
                    }
                }
            }
            const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
#pragma unroll
            for (int j = 0; j < SS_RG; ++j) {
                const float mu1 = o[0][j], mu2 = o[1][j];
                const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
                const float s1 = o[2][j] - mu1_sq, s2 = o[3][j] - mu2_sq, s12 = o[4][j] - mu12;
                const float num = (2.0f * mu12 + C1) * (2.0f * s12 + C2);
                const float den = (mu1_sq + mu2_sq + C1) * (s1 + s2 + C2);
                const float v = fdiv(num, den);          // MUFU.RCP, <= 1 ulp (the tolerance on the mean is 1e-5)
                if (y0 + r8 + j < H && x0 + col < W) sum += v;
            }
        }
    }
    const float s = block_sum(sum, red);
    if (tid == 0) a.part[(size_t)b * a.nchunks + chunk] = s;
    if (arrive_is_last(a.counters + b, (unsigned)a.nchunks, &last_flag)) {
        float v = 0.0f;
        for (int t = tid; t < a.nchunks; t += SS_NT) v += __ldcg(a.part + (size_t)b * a.nchunks + t);
        v = block_sum(v, red);
        if (tid == 0) a.ssim_sum[b] = v;
    }
}

static inline size_t ssim_tiles(int H, int W) { return (size_t)((H + SS_TH - 1) / SS_TH) * ((W + SS_TW - 1) / SS_TW); }

size_t ssim_workspace_bytes(int B, int C, int H, int W) {
    return COUNTER_REGION + (size_t)(B > 0 ? B : 1) * (C > 0 ? C : 1) * ssim_tiles(H, W) * sizeof(float);
}

int ssim_sum(const float *img1, const float *img2, float *out, int B, int C, int H, int W, void *ws, size_t ws_bytes,
             cudaStream_t stream) {
    if (!img1 || !img2 || !out || B < 1 || C < 1 || H < 1 || W < 1) return T2O_ERR_INVALID_ARG;
    if (B > 65535) return T2O_ERR_UNSUPPORTED;
    if (!ws || ws_bytes < ssim_workspace_bytes(B, C, H, W)) return T2O_ERR_WORKSPACE;
    SsimArgs a;
    memset(&a, 0, sizeof(a));
    a.a = img1; a.b = img2; a.ssim_sum = out;
    a.counters = (unsigned int *)ws;
    a.part = (float *)((char *)ws + COUNTER_REGION);
    a.C = C; a.H = H; a.W = W;
    a.tiles_x = (W + SS_TW - 1) / SS_TW;
    a.ntiles = (int)ssim_tiles(H, W);
    // utils/ssim/__init__.py:8-10: gauss = Tensor([exp(-(x - 5)^2 / (2 sigma^2))]) / sum, sigma = 1.5 (double exp, float32 tensor)
    float g[SS_K], gs = 0.0f;
    for (int x = 0; x < SS_K; ++x) g[x] = (float)exp(-(double)((x - SS_R) * (x - SS_R)) / (2.0 * 1.5 * 1.5));
    for (int x = 0; x < SS_K; ++x) gs += g[x];
    for (int x = 0; x < SS_K; ++x) a.w[x] = g[x] / gs;
    const size_t smem = (size_t)(2 * SS_PH * SS_PITCH + 5 * SS_PH * SS_TW) * sizeof(float);
    // CTAs per image: the resident CTAs (3 per SM) shared out over the batch, at most one per item
    const int items = a.ntiles * C;
    int nchunks = (3 * NUM_SMS) / B;
    if (nchunks < 1) nchunks = 1;
    if (nchunks > items) nchunks = items;
    a.nchunks = nchunks;
    const bool v4 = (W % 4 == 0) && ((uintptr_t)img1 % 16 == 0) && ((uintptr_t)img2 % 16 == 0);
    dim3 grid(nchunks, B);
    if (v4) {
        T2O_CUDA_OK(cudaFuncSetAttribute(ssim_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // per device
        ssim_kernel<true><<<grid, SS_NT, smem, stream>>>(a);
    } else {
        T2O_CUDA_OK(cudaFuncSetAttribute(ssim_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ssim_kernel<false><<<grid, SS_NT, smem, stream>>>(a);
    }
    T2O_CUDA_OK(cudaGetLastError());
    return T2O_OK;
}

}  // namespace t2o
