// t2o_common.cuh -- launch descriptors and device helpers shared by the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "t2o_math.cuh"

namespace t2o {

constexpr int NT = 256;           // threads per CTA of the chain kernels
constexpr int NUM_SMS = 148;      // B200
constexpr int MIN_TILE_PX = 1024; // smallest tile any launcher picks (bounds the workspace)
constexpr int MAX_PSTRIDE = 256;  // floats per parameter row
// Every entry point keeps its arrival counters (one per image / state, left at 0 again by the last CTA) in the first
// COUNTER_REGION bytes of the workspace and its partial sums behind them: whatever batch sizes share one workspace,
// one call's partials never land on another call's counters.
constexpr size_t COUNTER_REGION = 65536 * sizeof(unsigned int);

// One fused chain, uniform over the batch.
struct ChainDesc {
    int n;                  // number of operators
    int L;                  // curve steps
    int sharp;              // index of the (single) sharpness operator, or -1
    int op[MAX_CHAIN];
    int poff[MAX_CHAIN];    // column of the operator's first parameter
};

// Launch geometry.  Without sharpness an image is a flat array of `ngroups` VEC-pixel groups and
// a tile is a contiguous range of `tile_groups`; with sharpness tiles are TH x TWg groups with a halo.
struct Geom {
    int B, H, W;
    int Wg;                 // W / VEC (2-D tiling only)
    int tiles_x, tiles_y;   // 2-D tiling only
    int ntiles;             // tiles per image
    int nchunks;            // CTAs per image (gridDim.x): nchunks * tiles_per_cta >= ntiles
    int tiles_per_cta;
    int TH, TWg;            // 2-D tile, in rows / groups
    int tile_groups;        // 1-D tiling
    long long ngroups;      // 1-D tiling: H*W / VEC
    unsigned int mul_tiles_x, mul_tw, mul_rw, mul_gw, mul_xw;   // ceil(2^32 / d) magic numbers (fast_div)
};

// ---------------------------------------------------------------- vector global access
template <int VEC> struct VecT;
template <> struct VecT<1> { using type = float; };
template <> struct VecT<2> { using type = float2; };
template <> struct VecT<4> { using type = float4; };

template <int VEC>
__device__ __forceinline__ void ld_vec(const float *p, float (&v)[VEC]) {
    if constexpr (VEC == 4) { const float4 t = __ldg(reinterpret_cast<const float4 *>(p)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    else if constexpr (VEC == 2) { const float2 t = __ldg(reinterpret_cast<const float2 *>(p)); v[0] = t.x; v[1] = t.y; }
    else { v[0] = __ldg(p); }
}
template <int VEC>
__device__ __forceinline__ void st_vec(float *p, const float (&v)[VEC]) {
    if constexpr (VEC == 4) { *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]); }
    else if constexpr (VEC == 2) { *reinterpret_cast<float2 *>(p) = make_float2(v[0], v[1]); }
    else { *p = v[0]; }
}
// shared-memory flavours (plain loads, no read-only path)
template <int VEC>
__device__ __forceinline__ void lds_vec(const float *p, float (&v)[VEC]) {
    if constexpr (VEC == 4) { const float4 t = *reinterpret_cast<const float4 *>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    else if constexpr (VEC == 2) { const float2 t = *reinterpret_cast<const float2 *>(p); v[0] = t.x; v[1] = t.y; }
    else { v[0] = *p; }
}

// asynchronous global -> shared copy of one VEC-float vector (LDGSTS: the data never passes through registers, so a
// thread can have its next rows in flight while it computes); a thread only ever reads back what it copied itself,
// hence cp_async_wait_all() alone -- no barrier -- makes the data visible to it
template <int VEC>
__device__ __forceinline__ void cp_async_vec(float *smem_dst, const float *gsrc) {
    const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
    if constexpr (VEC == 4) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
    else if constexpr (VEC == 2) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// explicit groups, for a thread that keeps two copies in flight and needs only the older one: commit after every
// (possibly empty) request so that the group count is the same on every path, then wait until <= N groups are pending
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// three planes of a pixel group -> a thread's staging slots (`sstride` floats between the planes)
template <int VEC>
__device__ __forceinline__ void cp_async_px(float *stg, int sstride, const float *base, size_t plane, size_t off) {
    cp_async_vec<VEC>(stg, base + off);
    cp_async_vec<VEC>(stg + sstride, base + plane + off);
    cp_async_vec<VEC>(stg + 2 * sstride, base + 2 * plane + off);
}
template <int VEC>
__device__ __forceinline__ void lds_px(const float *stg, int sstride, float (&x)[3][VEC]) {
    lds_vec<VEC>(stg, x[0]);
    lds_vec<VEC>(stg + sstride, x[1]);
    lds_vec<VEC>(stg + 2 * sstride, x[2]);
}

// pixel-group loads: three planes (+ optional mask planes)
template <int VEC>
__device__ __forceinline__ void ld_px(const float *base, size_t plane, size_t off, float (&x)[3][VEC]) {
    ld_vec<VEC>(base + off, x[0]);
    ld_vec<VEC>(base + plane + off, x[1]);
    ld_vec<VEC>(base + 2 * plane + off, x[2]);
}
template <int VEC>
__device__ __forceinline__ void st_px(float *base, size_t plane, size_t off, const float (&x)[3][VEC]) {
    st_vec<VEC>(base + off, x[0]);
    st_vec<VEC>(base + plane + off, x[1]);
    st_vec<VEC>(base + 2 * plane + off, x[2]);
}
template <int VEC>
__device__ __forceinline__ void ld_mask(const float *mask_b, int mask_ch, size_t plane, size_t off, float (&m)[3][VEC]) {
    if (mask_b == nullptr) {
#pragma unroll
        for (int v = 0; v < VEC; ++v) { m[0][v] = 1.0f; m[1][v] = 1.0f; m[2][v] = 1.0f; }
    } else if (mask_ch == 3) {
        ld_px<VEC>(mask_b, plane, off, m);
    } else {
        ld_vec<VEC>(mask_b + off, m[0]);
#pragma unroll
        for (int v = 0; v < VEC; ++v) { m[1][v] = m[0][v]; m[2][v] = m[0][v]; }
    }
}

// ---------------------------------------------------------------- reductions (fixed order => deterministic)
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// Sum over the CTA; result valid in every thread of warp 0.  `red` holds >= 32 floats.
__device__ __forceinline__ float block_sum(float v, float *red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();                 // protect `red` from the previous use
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float r = 0.0f;
    if (warp == 0) {
        r = lane < nw ? red[lane] : 0.0f;
        r = warp_sum(r);
    }
    return r;
}

// "last CTA of the image finishes the reduction": returns true in every thread of exactly one
// CTA per counter, after all CTAs' partials are visible.  The counter is left at 0 again.
__device__ __forceinline__ bool arrive_is_last(unsigned int *counter, unsigned int total, int *flag_smem) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int ticket = atomicAdd(counter, 1u);
        const int last = (ticket == total - 1);
        if (last) *counter = 0u;
        *flag_smem = last;
    }
    __syncthreads();
    const bool last = *flag_smem != 0;
    if (last) __threadfence();
    return last;
}

// Column sums of a (ntiles, ncols) partial matrix by one CTA: out[col] = sum_t part[t*ncols + col].
// One warp per column, lanes stride the tiles, then a shuffle tree: fixed order.
__device__ __forceinline__ void reduce_columns(const float *part, int ntiles, int ncols, float *out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int col = warp; col < ncols; col += nw) {
        float s = 0.0f;
#pragma unroll 4
        for (int t = lane; t < ntiles; t += 32) s += __ldcg(part + (size_t)t * ncols + col);
        s = warp_sum(s);
        if (lane == 0) out[col] = s;
    }
}

#define T2O_CUDA_OK(expr)                                   \
    do {                                                    \
        cudaError_t e__ = (expr);                           \
        if (e__ != cudaSuccess) { t2o::set_cuda_error(e__); return T2O_ERR_CUDA; } \
    } while (0)

void set_cuda_error(cudaError_t e);

}  // namespace t2o
