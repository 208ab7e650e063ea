// Static instruction cost of each operator's forward / backward pixel code (4 pixels per thread, no mask):
//   nvcc -arch=sm_100a -cubin -o /tmp/op_cost.cubin scripts/micro/op_cost.cu && cuobjdump -sass /tmp/op_cost.cubin
#include "../../t2onet_b200/csrc/t2o_math.cuh"
using namespace t2o;
struct Io { const float4 *x; const float4 *g; float4 *o; float *acc; const float *tab; };
#define LOAD4(P, A) { float4 q = P; A[0] = q.x; A[1] = q.y; A[2] = q.z; A[3] = q.w; }
template <int OP, bool CL>
__global__ void bwd_k(Io io) {
    __shared__ __align__(16) float tab[TAB];
    if (threadIdx.x < TAB) tab[threadIdx.x] = io.tab[threadIdx.x];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float x[3][4], g[3][4];
    for (int c = 0; c < 3; ++c) { LOAD4(io.x[i * 3 + c], x[c]); LOAD4(io.g[i * 3 + c], g[c]); }
    GradAcc A; acc_zero(A);
    {
#pragma unroll
    for (int v = 0; v < 4; ++v)
        pointwise_bwd<false, CL>(OP, tab, 8, x[0][v], x[1][v], x[2][v], 1.f, 1.f, 1.f, g[0][v], g[1][v], g[2][v], A, true);
    }
    for (int c = 0; c < 3; ++c) io.o[i * 3 + c] = make_float4(g[c][0], g[c][1], g[c][2], g[c][3]);
    float v[ACC_SLOTS]; acc_to_slots(A, v);
    float s = 0; for (int k = 0; k < ACC_SLOTS; ++k) s += v[k] * (k + 1);
    io.acc[i] = s;
}
template <int OP, bool CL>
__global__ void fwd_k(Io io) {
    __shared__ __align__(16) float tab[TAB];
    if (threadIdx.x < TAB) tab[threadIdx.x] = io.tab[threadIdx.x];
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float x[3][4];
    for (int c = 0; c < 3; ++c) { LOAD4(io.x[i * 3 + c], x[c]); }
    {
#pragma unroll
    for (int v = 0; v < 4; ++v) op_apply<false, CL>(OP, tab, 8, x[0][v], x[1][v], x[2][v], 1.f, 1.f, 1.f);
    }
    for (int c = 0; c < 3; ++c) io.o[i * 3 + c] = make_float4(x[c][0], x[c][1], x[c][2], x[c][3]);
}
#define INST(OP, CL) template __global__ void bwd_k<OP, CL>(Io); template __global__ void fwd_k<OP, CL>(Io);
INST(OP_BRIGHTNESS, false) INST(OP_CONTRAST, false) INST(OP_SATURATION, false) INST(OP_COLOR, true) INST(OP_TONE, true)
INST(OP_COLOR, false) INST(OP_IDENTITY, false)
