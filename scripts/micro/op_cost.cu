// op_cost.cu -- static SASS instruction counts of each operator's forward / backward body on a 4-pixel group
// (development aid): nvcc -cubin, then count instructions per kernel and subtract the identity baseline.
#include "../../t2onet_b200/csrc/t2o_step_kernels.cuh"
using namespace t2o;
template <int OP, bool CL>
__global__ void fwd(const float4 *in, float4 *out, const float *tab) {
    __shared__ float st[TAB];
    if (threadIdx.x < TAB) st[threadIdx.x] = tab[threadIdx.x];
    __syncthreads();
    float x[3][4], m[3][4];
    for (int c = 0; c < 3; ++c) { float4 v = in[c * 1024 + threadIdx.x]; x[c][0] = v.x; x[c][1] = v.y; x[c][2] = v.z; x[c][3] = v.w; }
#pragma unroll
    for (int v = 0; v < 4; ++v) op_apply<false, CL>(OP, st, 8, x[0][v], x[1][v], x[2][v], 1.f, 1.f, 1.f);
    for (int c = 0; c < 3; ++c) out[c * 1024 + threadIdx.x] = make_float4(x[c][0], x[c][1], x[c][2], x[c][3]);
}
template <int OP, bool CL>
__global__ void bwd(const float4 *in, float4 *out, const float *tab, float *accout) {
    __shared__ float st[TAB];
    if (threadIdx.x < TAB) st[threadIdx.x] = tab[threadIdx.x];
    __syncthreads();
    float x[3][4], g[3][4];
    for (int c = 0; c < 3; ++c) { float4 v = in[c * 1024 + threadIdx.x]; x[c][0] = v.x; x[c][1] = v.y; x[c][2] = v.z; x[c][3] = v.w; }
    for (int c = 0; c < 3; ++c) { float4 v = in[(c + 3) * 1024 + threadIdx.x]; g[c][0] = v.x; g[c][1] = v.y; g[c][2] = v.z; g[c][3] = v.w; }
    GradAcc A; acc_zero(A);
#pragma unroll
    for (int v = 0; v < 4; ++v) pointwise_bwd<false, CL>(OP, st, 8, x[0][v], x[1][v], x[2][v], 1.f, 1.f, 1.f, g[0][v], g[1][v], g[2][v], A, true);
    for (int c = 0; c < 3; ++c) out[c * 1024 + threadIdx.x] = make_float4(g[c][0], g[c][1], g[c][2], g[c][3]);
    float s[ACC_SLOTS]; acc_to_slots(A, s); float t = 0; for (int i = 0; i < ACC_SLOTS; ++i) t += s[i] * (i + 1);
    accout[threadIdx.x] = t;
}
#define INST(OP) template __global__ void fwd<OP, true>(const float4 *, float4 *, const float *); \
                 template __global__ void bwd<OP, true>(const float4 *, float4 *, const float *, float *);
INST(-1) INST(0) INST(1) INST(2) INST(3) INST(5) INST(8) INST(9)
