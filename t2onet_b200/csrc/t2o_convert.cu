// t2o_convert.cu -- 8-bit image <-> float32 tensor conversions on the device.
//
// Replaces, bit for bit (file:line in /root/reference):
//   img2tensor / load_train_img / load_infer_img   utils/visual_utils.py:46,61-70   BGR HWC uint8 -> RGB CHW float32, x / 255
//   tensor2img                                     utils/visual_utils.py:50-58      RGB CHW float32 -> BGR HWC uint8, (x * 255) truncated
// so that images cross PCIe as the 8-bit arrays they are (a quarter of the float32 bytes) and the x / 255 happens in HBM.
// Both layouts are served: planar (N, 3, H, W) uint8 <-> float32 (the layout the kernels use) and cv2's interleaved
// (N, H, W, 3) BGR.  Memory-bound: 15 B per pixel (3 read + 12 written, or the reverse).
//
//   x / 255 is an IEEE fp32 division in the reference (torch true-divide of a uint8 tensor by a Python int), so the
//   device uses __fdiv_rn, not a multiplication by 1/255 (which differs in the last bit for 126 of the 256 values).
//   (x * 255).astype(np.uint8) truncates toward zero; values outside [0, 255] are undefined in numpy -- clamped here.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/t2o.h"
#include "t2o_common.cuh"

namespace t2o {

__device__ __forceinline__ float u8_to_unit(unsigned int v) { return __fdiv_rn((float)v, 255.0f); }
__device__ __forceinline__ unsigned int unit_to_u8(float x) {
    const float y = __fmul_rn(x, 255.0f);
    return (unsigned int)__float2int_rz(fminf(fmaxf(y, 0.0f), 255.0f));
}

// planar: n bytes <-> n floats, 16 values per thread and iteration (one 128-bit load, four 128-bit stores)
__global__ void __launch_bounds__(256) u8_to_f32_kernel(const uint8_t *__restrict__ src, float *__restrict__ dst, long long n) {
    const long long n16 = n >> 4;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
        const uint4 w = __ldg(reinterpret_cast<const uint4 *>(src) + i);
        const unsigned int ws[4] = {w.x, w.y, w.z, w.w};
        float4 *o = reinterpret_cast<float4 *>(dst) + 4 * i;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            o[k] = make_float4(u8_to_unit(ws[k] & 255u), u8_to_unit((ws[k] >> 8) & 255u), u8_to_unit((ws[k] >> 16) & 255u),
                               u8_to_unit(ws[k] >> 24));
    }
    for (long long i = (n16 << 4) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        dst[i] = u8_to_unit(src[i]);
}

__global__ void __launch_bounds__(256) f32_to_u8_kernel(const float *__restrict__ src, uint8_t *__restrict__ dst, long long n) {
    const long long n16 = n >> 4;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
        const float4 *s = reinterpret_cast<const float4 *>(src) + 4 * i;
        unsigned int ws[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float4 v = __ldg(s + k);
            ws[k] = unit_to_u8(v.x) | unit_to_u8(v.y) << 8 | unit_to_u8(v.z) << 16 | unit_to_u8(v.w) << 24;
        }
        reinterpret_cast<uint4 *>(dst)[i] = make_uint4(ws[0], ws[1], ws[2], ws[3]);
    }
    for (long long i = (n16 << 4) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        dst[i] = (uint8_t)unit_to_u8(src[i]);
}

// interleaved BGR (N, H, W, 3) uint8 <-> planar RGB (N, 3, H, W) float32.  A thread owns 4 consecutive pixels of one
// image: 12 interleaved bytes = three aligned 32-bit words (plane sizes that are not a multiple of 4 take the scalar
// path for their last pixels), three float4 on the planar side.
__global__ void __launch_bounds__(256) hwc_bgr_to_chw_kernel(const uint8_t *__restrict__ src, float *__restrict__ dst, int N, long long plane) {
    const long long groups = plane >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (int b = blockIdx.y; b < N; b += gridDim.y) {
        const uint8_t *s = src + (size_t)b * 3 * plane;
        float *d = dst + (size_t)b * 3 * plane;
        const bool aligned = ((size_t)s & 3) == 0 && (plane & 3) == 0;
        for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += stride) {
            unsigned char px[12];
            if (aligned) {
                const unsigned int *w = reinterpret_cast<const unsigned int *>(s) + 3 * g;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const unsigned int v = __ldg(w + k);
                    px[4 * k] = v & 255u; px[4 * k + 1] = (v >> 8) & 255u; px[4 * k + 2] = (v >> 16) & 255u; px[4 * k + 3] = v >> 24;
                }
            } else {
#pragma unroll
                for (int k = 0; k < 12; ++k) px[k] = s[12 * g + k];
            }
            // px = b0 g0 r0 b1 g1 r1 ...; output plane 0 = R, 1 = G, 2 = B (img[:, :, ::-1], utils/visual_utils.py:66)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float4 v = make_float4(u8_to_unit(px[2 - c]), u8_to_unit(px[5 - c]), u8_to_unit(px[8 - c]), u8_to_unit(px[11 - c]));
                if (aligned) *reinterpret_cast<float4 *>(d + c * plane + 4 * g) = v;
                else { float *o = d + c * plane + 4 * g; o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
            }
        }
        for (long long i = (groups << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += stride)
#pragma unroll
            for (int c = 0; c < 3; ++c) d[c * plane + i] = u8_to_unit(s[3 * i + 2 - c]);
    }
}

__global__ void __launch_bounds__(256) chw_to_hwc_bgr_kernel(const float *__restrict__ src, uint8_t *__restrict__ dst, int N, long long plane) {
    const long long groups = plane >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (int b = blockIdx.y; b < N; b += gridDim.y) {
        const float *s = src + (size_t)b * 3 * plane;
        uint8_t *d = dst + (size_t)b * 3 * plane;
        const bool aligned = ((size_t)d & 3) == 0 && (plane & 3) == 0;
        for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += stride) {
            unsigned int px[12];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float v[4];
                if (aligned) { const float4 t = __ldg(reinterpret_cast<const float4 *>(s + c * plane) + g); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
                else { const float *p = s + c * plane + 4 * g; v[0] = p[0]; v[1] = p[1]; v[2] = p[2]; v[3] = p[3]; }
#pragma unroll
                for (int k = 0; k < 4; ++k) px[3 * k + 2 - c] = unit_to_u8(v[k]);
            }
            if (aligned) {
                unsigned int *w = reinterpret_cast<unsigned int *>(d) + 3 * g;
#pragma unroll
                for (int k = 0; k < 3; ++k) w[k] = px[4 * k] | px[4 * k + 1] << 8 | px[4 * k + 2] << 16 | px[4 * k + 3] << 24;
            } else {
#pragma unroll
                for (int k = 0; k < 12; ++k) d[12 * g + k] = (uint8_t)px[k];
            }
        }
        for (long long i = (groups << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += stride)
#pragma unroll
            for (int c = 0; c < 3; ++c) d[3 * i + 2 - c] = (uint8_t)unit_to_u8(s[c * plane + i]);
    }
}

static int grid_for(long long items) {
    long long g = (items + 255) / 256;
    const long long cap = (long long)NUM_SMS * 8;          // 8 resident CTAs of 256 threads per SM, grid-stride beyond
    if (g > cap) g = cap;
    return g < 1 ? 1 : (int)g;
}

int convert_u8_to_f32(const uint8_t *src, float *dst, long long n, cudaStream_t stream) {
    if (!src || !dst || n < 0) return T2O_ERR_INVALID_ARG;
    if (n == 0) return T2O_OK;
    if (((size_t)src & 15) || ((size_t)dst & 15)) return T2O_ERR_INVALID_ARG;
    u8_to_f32_kernel<<<grid_for(n / 16 + 1), 256, 0, stream>>>(src, dst, n);
    T2O_CUDA_OK(cudaGetLastError());
    return T2O_OK;
}

int convert_f32_to_u8(const float *src, uint8_t *dst, long long n, cudaStream_t stream) {
    if (!src || !dst || n < 0) return T2O_ERR_INVALID_ARG;
    if (n == 0) return T2O_OK;
    if (((size_t)src & 15) || ((size_t)dst & 15)) return T2O_ERR_INVALID_ARG;
    f32_to_u8_kernel<<<grid_for(n / 16 + 1), 256, 0, stream>>>(src, dst, n);
    T2O_CUDA_OK(cudaGetLastError());
    return T2O_OK;
}

int convert_img2tensor(const uint8_t *hwc_bgr, float *chw_rgb, int N, int H, int W, cudaStream_t stream) {
    if (!hwc_bgr || !chw_rgb || N < 1 || H < 1 || W < 1) return T2O_ERR_INVALID_ARG;
    if ((size_t)chw_rgb & 15) return T2O_ERR_INVALID_ARG;
    const long long plane = (long long)H * W;
    dim3 grid(grid_for(plane / 4 + 1), N < 65535 ? N : 65535);
    hwc_bgr_to_chw_kernel<<<grid, 256, 0, stream>>>(hwc_bgr, chw_rgb, N, plane);
    T2O_CUDA_OK(cudaGetLastError());
    return T2O_OK;
}

int convert_tensor2img(const float *chw_rgb, uint8_t *hwc_bgr, int N, int H, int W, cudaStream_t stream) {
    if (!hwc_bgr || !chw_rgb || N < 1 || H < 1 || W < 1) return T2O_ERR_INVALID_ARG;
    if ((size_t)chw_rgb & 15) return T2O_ERR_INVALID_ARG;
    const long long plane = (long long)H * W;
    dim3 grid(grid_for(plane / 4 + 1), N < 65535 ? N : 65535);
    chw_to_hwc_bgr_kernel<<<grid, 256, 0, stream>>>(chw_rgb, hwc_bgr, N, plane);
    T2O_CUDA_OK(cudaGetLastError());
    return T2O_OK;
}

}  // namespace t2o
