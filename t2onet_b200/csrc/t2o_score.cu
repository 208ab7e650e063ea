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

#include <cooperative_groups.h>

#include "t2o_common.cuh"
#include "t2o_nm_device.cuh"
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
// distributed shared memory: the address of `p` (this CTA's shared memory) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(const void *p, int rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
    return r;
}
// one 32-bit word into another CTA's shared memory, completing 4 bytes of the transaction count of an mbarrier there: the
// receiver that sees the barrier's phase complete sees the data (no fence, no cluster barrier)
__device__ __forceinline__ void st_async_u32(uint32_t remote_addr, uint32_t value, uint32_t remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
                 ::"r"(remote_addr), "r"(value), "r"(remote_bar) : "memory");
}
// wait for a phase with a suspend-time hint: the warp sleeps in hardware until the phase completes (or ~1 ms pass) instead of
// re-issuing try_wait -- without the hint the waiting warps' polling was a quarter of all instructions the resident kernel
// issued (ncu), taken from the arithmetic of the CTA next to them
__device__ __forceinline__ bool mbar_try_wait_sleep(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(1000000u) : "memory");
    return ok != 0;
}
// the same hand-over inside ONE CTA (an image of a single tile: a cluster of one block has no distributed shared memory to
// address): a plain store, then a CTA fence, then the transaction bytes completed on the local mbarrier -- a release pattern
// that the waiters' try_wait (acquire) pairs with.  (compute-sanitizer's racecheck does not model mbarrier ordering and
// reports these stores against the waiters' reads; memcheck is clean.)
__device__ __forceinline__ void st_local_u32(void *dst, uint32_t value, uint64_t *bar) {
    *reinterpret_cast<volatile uint32_t *>(dst) = value;
    __threadfence_block();
    asm volatile("mbarrier.complete_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(4u) : "memory");
}
__device__ __forceinline__ void mbar_wait_or_trap(uint64_t *bar, uint32_t parity) {
    if (mbar_try_wait_sleep(bar, parity)) return;
    // a broken protocol must fail loudly, not hang the GPU: trap after 20 s of wall clock (a round takes microseconds; the
    // bound is in time, not in polls, because how long one try_wait sleeps is up to the hardware)
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        for (int it = 0; it < 64; ++it)
            if (mbar_try_wait_sleep(bar, parity)) return;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 20000000000ull) __trap();
    }
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
// (split in two so that the resident kernel loads a group once for all the fits of its state: xc / t are the group's centre
// pixels of the state and its target pixels)
// clamped_in: every value of the state tile lies in [0, 1] (true of every state a planner meets: images and clamped edits), so
// the curve operators' input clamp is the identity and is skipped -- the same bits, one instruction per channel less
template <int VEC, bool HM>
__device__ __forceinline__ void score_group_loaded(float &sum, int op, const float *tab, int L, float p, const float (&xc)[3][VEC],
                                                   const float (&t)[3][VEC], const float *src, int spitch, int cs,
                                                   const float *mptr, size_t mcs, bool clamped_in = false) {
    float x[3][VEC], m[3][VEC];
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
            float up[VEC], dn[VEC];
            lds_vec<VEC>(row - spitch, up);
            lds_vec<VEC>(row + spitch, dn);
            const float lf = row[-1], rt = row[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const float l = v > 0 ? xc[c][v - 1] : lf;
                const float r = v < VEC - 1 ? xc[c][v + 1] : rt;
                x[c][v] = sat01(blend<HM>(fmaf(p, laplace(xc[c][v], up[v], dn[v], l, r), xc[c][v]), xc[c][v], m[c][v]));
            }
        }
    } else if (op == OP_BLUR) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float *row = src + c * cs, *ru = row - spitch, *rd = row + spitch;
            float up[VEC], dn[VEC];
            lds_vec<VEC>(ru, up);
            lds_vec<VEC>(rd, dn);
            const float lf = row[-1], rt = row[VEC], ul = ru[-1], ur = ru[VEC], dl = rd[-1], dr = rd[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const float l = v > 0 ? xc[c][v - 1] : lf, r = v < VEC - 1 ? xc[c][v + 1] : rt;
                const float e = v > 0 ? up[v - 1] : ul, f = v < VEC - 1 ? up[v + 1] : ur;
                const float g = v > 0 ? dn[v - 1] : dl, h = v < VEC - 1 ? dn[v + 1] : dr;
                x[c][v] = sat01(blend<HM>(fmaf(p, blur_delta(xc[c][v], (up[v] + dn[v]) + (l + r), (e + f) + (g + h)), xc[c][v]), xc[c][v], m[c][v]));
            }
        }
    } else {
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int v = 0; v < VEC; ++v) x[c][v] = xc[c][v];
        switch (op) {
#define T2O_CASE(OPC)                                                                                         \
    case OPC:                                                                                                 \
        _Pragma("unroll") for (int v = 0; v < VEC; ++v)                                                       \
            op_apply<HM>(OPC, tab, L, x[0][v], x[1][v], x[2][v], m[0][v], m[1][v], m[2][v]);                       \
        break;
#define T2O_CASE_CURVE(OPC)                                                                                   \
    case OPC:                                                                                                 \
        if (clamped_in) {                                                                                     \
            _Pragma("unroll") for (int v = 0; v < VEC; ++v)                                                   \
                op_apply<HM, true>(OPC, tab, L, x[0][v], x[1][v], x[2][v], m[0][v], m[1][v], m[2][v]);             \
        } else {                                                                                              \
            _Pragma("unroll") for (int v = 0; v < VEC; ++v)                                                   \
                op_apply<HM>(OPC, tab, L, x[0][v], x[1][v], x[2][v], m[0][v], m[1][v], m[2][v]);                   \
        }                                                                                                     \
        break;
            T2O_CASE(OP_BRIGHTNESS) T2O_CASE(OP_CONTRAST) T2O_CASE(OP_SATURATION) T2O_CASE_CURVE(OP_COLOR)
            T2O_CASE_CURVE(OP_TONE) T2O_CASE(OP_WHITE) T2O_CASE(OP_EXPOSURE) T2O_CASE(OP_WHITEBALANCE)
            T2O_CASE(OP_BNW) T2O_CASE(OP_HUE)
#undef T2O_CASE
#undef T2O_CASE_CURVE
            default: break;
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int v = 0; v < VEC; ++v) sum += fabsf(x[c][v] - t[c][v]);
}

template <int VEC, bool HM>
__device__ __forceinline__ void score_group(float &sum, int op, const float *tab, int L, float p, const float *src, int spitch, int cs,
                                            const float *tsrc, int ct, const float *mptr, size_t mcs, bool clamped_in = false) {
    float xc[3][VEC], t[3][VEC];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        lds_vec<VEC>(src + c * cs, xc[c]);
        lds_vec<VEC>(tsrc + c * ct, t[c]);
    }
    score_group_loaded<VEC, HM>(sum, op, tab, L, p, xc, t, src, spitch, cs, mptr, mcs, clamped_in);
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

    // (aligned by an offset into the array, not by integer arithmetic on the address: the pointers stay shared-memory pointers,
    // so the tiles are read with LDS rather than generic loads)
    float *sS = reinterpret_cast<float *>(dyn_raw + ((128u - (smem_u32(dyn_raw) & 127u)) & 127u));
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
            // one warp, the whole tile -- summed in the ORDER of the cooperative path (the share of each of its eight warps,
            // reduced, then added in warp order), so that a candidate's score does not depend on how many candidates share
            // its state: a fit is the same fit whether six or sixteen others run beside it
            const float *mimg = cand_mask_img(ci);
            float sum = 0.0f;
            for (int vw = 0; vw < SCORE_NW; ++vw) sum += warp_sum(tile_sum(op, tab, vw * 32 + lane, SCORE_NT, mimg));
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

// ------------------------------------------------------------------------------------------- resident Nelder-Mead
// The fits of a planner step as ONE launch.  Round by round (t2o_score_candidates + t2o_nm_advance, thousands of times) every
// round re-stages every state: a CTA per (state, tile) pays a candidate scan, a TMA round trip and an arrival-counter
// epilogue for ~0.6 us of arithmetic, and the host pays a launch pair per round.  Here a CLUSTER of ntiles CTAs owns a state
// for the whole life of its fits: each CTA stages its tile of the state and of the target ONCE; the leader (CTA 0) copies the
// Nelder-Mead state of the state's fits (simplex, values, counters) from the caller's arrays into shared memory, packed by the
// fits' dimensions -- EVERY CTA of the cluster does, and every CTA steps every fit itself (warp c: fit c): the step is
// deterministic, so all copies agree on the next vertex without a word exchanged, and only the tile partials travel:
//     every CTA: tables of the live fits' vertices (one lane per curve record), |op(tile) - target tile| per fit
//                (score_kernel's cooperative path), its tile partials -> every CTA's shared memory (st.async over distributed
//                shared memory, completing the transaction count of the receiver's mbarrier: no fence, no cluster barrier)
//     every CTA: the fit warps wait on that mbarrier, sum the partials in tile order, consume the score and propose the next
//                vertex (nm_step, on shared memory and registers)
// until every fit of the state is finished; the leader (CTA 0) writes results and the fits' state back, and the cluster takes
// the next state (atomic work counter).  One hand-over per evaluation; two partial buffers / mbarriers alternate by round
// parity, so a CTA that runs one round ahead can never complete a phase its neighbour still waits on.
// Vertices and scores never leave the SMs.  The arithmetic is the round-by-round path's, bit for bit: the same tiles, the
// same groups per thread in the same order, the same warp / tile reduction order, the same Nelder-Mead code --
// tests/test_gpu_nm.py compares the two exactly.
namespace cg = cooperative_groups;
constexpr int RES_MAXT = 16;     // tiles per state = cluster size (8 is the portable limit; 9 .. 16 -- a 256 x 256 image -- are opted into)

struct ResidentShared {
    float wsum[SCORE_NW][SCORE_NW];
    float part[2][SCORE_NW][RES_MAXT];   // tile partials of every fit, sent by every CTA of the cluster to every CTA (completing sbar);
                                         // two buffers / barriers, by round parity: a CTA one round ahead cannot touch an open phase
    float pend[SCORE_NW][NM_MAXN];       // the vertex / operator each fit evaluates next (written by this CTA's own Nelder-Mead step)
    int pop[SCORE_NW];
    int cmask[SCORE_NW];                 // the fits' masks
    int nmoff[SCORE_NW];                 // byte offset of fit c in the Nelder-Mead region, -1: not loaded
    int nmN[SCORE_NW];                   // its number of parameters
    int toff[SCORE_NW];                  // offset of fit c's table (floats)
    int next_state;                      // the state the cluster works on (written by the leader into every CTA)
    uint64_t bar;                        // the tiles' TMA loads
    uint64_t sbar[2];                    // a round's tile partials have arrived (4 bytes per live fit and CTA)
};

// floats of an operator's table / bytes of a fit's Nelder-Mead state in shared memory (host and device agree on these)
__host__ __device__ inline int res_tab_floats(int op) { return op == OP_COLOR ? 3 * CT : (op == OP_TONE ? CT : 8); }
__host__ __device__ inline int res_nm_bytes(int N) { return (8 * (N * N + 5 * N + 2) + 4 * (N + 9) + 7) & ~7; }

// fit c of the leader: views into the packed region [sim (N+1) x N | vec 3 x N | fsim N+1 | fxr | perm N+1 | ctl 8]
__device__ __forceinline__ void res_bind(NMWarp &w, unsigned char *base, int N, const NMArgs &nm, int p, int lane, float *cprm, int *cop,
                                         bool store_result) {
    double *d = reinterpret_cast<double *>(base);
    w.sim = d; d += (N + 1) * N;
    w.vec = d; d += 3 * N;
    w.fsim = d; d += N + 1;
    w.fxr = d; d += 1;
    int *i = reinterpret_cast<int *>(d);
    w.perm = i; i += N + 1;
    w.ctl = i;
    w.xbest = nm.st.xbest + (size_t)p * NM_MAXN;
    w.fbest = nm.st.fbest + p;
    w.cparam = cprm;
    w.cop = cop;
    w.ld = N;
    w.lane = lane;
    w.store_result = store_result;
    // (these views are shared memory: lets the Nelder-Mead code, written for generic pointers, use LDS / STS)
    __builtin_assume(__isShared(w.sim)); __builtin_assume(__isShared(w.vec)); __builtin_assume(__isShared(w.fsim));
    __builtin_assume(__isShared(w.fxr)); __builtin_assume(__isShared(w.perm)); __builtin_assume(__isShared(w.ctl));
    __builtin_assume(__isShared(w.cparam)); __builtin_assume(__isShared(w.cop));
}

template <bool HM>
__global__ void __launch_bounds__(SCORE_NT, 2) nm_resident_kernel(const __grid_constant__ CUtensorMap tm_state,
                                                                 const __grid_constant__ CUtensorMap tm_target,
                                                                 const __grid_constant__ ScoreArgs a, const __grid_constant__ NMArgs nm,
                                                                 unsigned int *work_counter, int max_rounds, int tab_cap, int nm_cap,
                                                                 unsigned long long *probe) {
    constexpr int VEC = 4, HX = 4;
    extern __shared__ __align__(16) unsigned char dyn_raw[];
    __shared__ __align__(16) ResidentShared sh;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();                  // = tile index
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = a.H, W = a.W, TH = a.TH, TW = a.TW;
    const int spitch = TW + 2 * HX, srows = TH + 2;
    const int ty = rank / a.tiles_x, tx = rank - ty * a.tiles_x;
    const int y0 = ty * TH, x0 = tx * TW;
    const int TWg = TW / VEC, ngroups = TH * TWg;
    const int twg_magic = 65536 / TWg + 1;                       // gi / TWg == (gi * twg_magic) >> 16 for gi * TWg < 65536
    const size_t plane = (size_t)H * W;
    const size_t mcs = HM && a.mask_ch == 3 ? plane : 0;
    // (aligned by an offset into the array, not by integer arithmetic on the address: the pointers stay shared-memory pointers,
    // so the tiles are read with LDS rather than generic loads)
    float *sS = reinterpret_cast<float *>(dyn_raw + ((128u - (smem_u32(dyn_raw) & 127u)) & 127u));
    float *sT = reinterpret_cast<float *>(reinterpret_cast<unsigned char *>(sS) + ((3 * srows * spitch * 4 + 127) & ~127));
    float *tabs = sT + 3 * TH * TW;                                              // tab_cap floats (16-byte aligned)
    unsigned char *nmreg = reinterpret_cast<unsigned char *>(tabs + tab_cap);      // nm_cap bytes (leader only)
    if (tid == 0) {
        mbar_init(&sh.bar, 1);
        mbar_init(&sh.sbar[0], 1);
        mbar_init(&sh.sbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t bar_phase = 0, sphase0 = 0, sphase1 = 0;             // (parities of the tile barrier and of the two partial barriers)
    unsigned int gr = 0;                                         // rounds this CTA has played (the same in every CTA of the cluster)
    const bool single = a.ntiles == 1;
    // every CTA of the cluster is running (and its mbarriers are initialised) before anyone writes into its shared memory
    cluster.sync();

    for (;;) {
        // ---- the cluster's next state
        if (rank == 0 && tid == 0) {
            const int nxt = (int)atomicAdd(work_counter, 1u);
            for (int r = 0; r < a.ntiles; ++r) (single ? &sh : cluster.map_shared_rank(&sh, r))->next_state = nxt;
        }
        cluster.sync();
        const int s = sh.next_state;
        if (s >= a.S) break;
        const int cbeg = a.cand_begin[s], cend = a.cand_begin[s + 1];
        const int m = cend - cbeg;                               // <= SCORE_NW (checked by the host)
        {
            bool any = false;
            for (int ci = cbeg; ci < cend; ++ci) any |= __ldcg(a.cand_op + ci) != OP_SKIP;
            if (!any) continue;                                  // (every CTA of the cluster takes the same decision)
        }
        const int t_idx = a.state_target ? a.state_target[s] : s % a.T;
        if (tid == 0) {
            const uint32_t bytes = (uint32_t)(3 * srows * spitch + 3 * TH * TW) * 4u;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&sh.bar, bytes);
            tma_load_4d(sS, &tm_state, &sh.bar, x0 - HX, y0 - 1, 0, s);
            tma_load_4d(sT, &tm_target, &sh.bar, x0, y0, 0, t_idx);
        }
        // ---- EVERY CTA: the fits' Nelder-Mead state -> its own shared memory (warp c: fit c).  The CTAs of a cluster run the
        // same deterministic step on the same numbers, so each knows every next vertex without being told: only the tile
        // partials travel.  The leader's copy is the one that writes results and goes back to the caller's arrays.
        NMRegs nr = NMRegs{0.0, 0.0, 0, NM_DONE, 0, 0, 0, 0, 1};
        int my_to = 0;                                           // warp c < m: where fit c's table lives
        bool mylive = false;                                     // warp c < m: fit c has not finished
        if (warp < m) {
            const int p = cbeg + warp;
            // a state whose fits do not fit the shared memory the host sized is left alone (its fits stay unfinished)
            int off = 0, N = 0, nm_total = 0, tab_total = 0;
            for (int c = 0; c < m; ++c) {
                const int Nc = __ldcg(nm.st.ctl + (size_t)(cbeg + c) * 8 + CTL_N);
                const int tf = res_tab_floats(__ldcg(a.cand_op + cbeg + c));
                if (c < warp) { off += res_nm_bytes(Nc); my_to += tf; }
                if (c == warp) N = Nc;
                nm_total += res_nm_bytes(Nc);
                tab_total += tf;
            }
            const bool fits_ok = nm_total <= nm_cap && tab_total <= tab_cap;
            const int op = fits_ok ? __ldcg(a.cand_op + p) : OP_SKIP;
            const float prm = lane < NM_MAXN ? __ldcg(a.cand_param + (size_t)p * NM_MAXN + lane) : 0.0f;
            if (op != OP_SKIP) {
                NMWarp w;
                res_bind(w, nmreg + off, N, nm, p, lane, sh.pend[warp], &sh.pop[warp], rank == 0);
                const double *gsim = nm.st.sim + (size_t)p * NM_ROWS * NM_MAXN, *gvec = nm.st.vec + (size_t)p * 3 * NM_MAXN;
                if (lane < N) {
                    for (int r = 0; r <= N; ++r) w.sim[r * N + lane] = __ldcg(gsim + r * NM_MAXN + lane);
                    for (int r = 0; r < 3; ++r) w.vec[r * N + lane] = __ldcg(gvec + r * NM_MAXN + lane);
                }
                if (lane <= N) {
                    w.fsim[lane] = __ldcg(nm.st.fsim + (size_t)p * NM_ROWS + lane);
                    w.perm[lane] = __ldcg(nm.st.perm + (size_t)p * NM_ROWS + lane);
                }
                if (lane < 8) w.ctl[lane] = __ldcg(nm.st.ctl + (size_t)p * 8 + lane);
                if (lane == 0) *w.fxr = __ldcg(nm.st.fxr + p);
                __syncwarp();
                nm_load(w);
                nr = nm_regs(w);                                 // (the warp keeps the fit's control state in registers from here on)
                mylive = true;
            }
            if (lane == 0) {
                sh.nmoff[warp] = op != OP_SKIP ? off : -1; sh.nmN[warp] = N; sh.pop[warp] = op;
                sh.toff[warp] = my_to; sh.cmask[warp] = (HM && a.cand_mask) ? a.cand_mask[p] : -1;
            }
            if (lane < NM_MAXN) sh.pend[warp][lane] = prm;
            __syncwarp();
        }
        mbar_wait_or_trap(&sh.bar, bar_phase);                   // the tiles
        bar_phase ^= 1u;
        // is every value of the state tile (halo included; out-of-image zeros count) in [0, 1]?  One look per state.
        bool clamped_in = true;
        for (int i = tid; i < 3 * srows * spitch; i += SCORE_NT) clamped_in &= in01(sS[i]);
        clamped_in = __syncthreads_and(clamped_in) != 0;         // (also: pop / toff / cmask of every warp are in place)
        // ---- rounds
#ifdef T2O_RES_PROBE
        long long pc[6] = {0, 0, 0, 0, 0, 0}, pt = 0;
#define T2O_PROBE(i) { const long long now = clock64(); pc[i] += now - pt; pt = now; }
#else
#define T2O_PROBE(i)
#endif
        for (int round = 0; round < max_rounds; ++round) {
#ifdef T2O_RES_PROBE
            pt = clock64();
#endif
            // the table of this warp's fit, from the vertex its own step proposed
            if (warp < m && mylive) {
                const int op = sh.pop[warp];
                mylive = op != OP_SKIP;
                if (mylive) build_table_lanes<false>(op, lane, sh.pend[warp], a.L, tabs + my_to);
            }
            __syncthreads();
            int nlive = 0;
            for (int c = 0; c < m; ++c) nlive += sh.pop[c] != OP_SKIP ? 1 : 0;
            if (nlive == 0) break;
            const int pb = (int)(gr & 1u);
            gr += 1u;
            if (tid == 0) mbar_expect_tx(&sh.sbar[pb], (uint32_t)(a.ntiles * nlive) * 4u);
            T2O_PROBE(1)
            // fit by fit, all warps share the tile (score_kernel's cooperative path: thread tid takes groups tid, tid + 256, ...)
#pragma unroll 1
            for (int c = 0; c < m; ++c) {
                const int op = sh.pop[c];
                if (op == OP_SKIP) continue;
                const float *tab = tabs + sh.toff[c];
                const float p = tab[0];
                const float *mimg = nullptr;
                if (HM) { const int mi = sh.cmask[c]; if (mi >= 0) mimg = a.masks + (size_t)mi * a.mask_ch * plane; }
                float sum = 0.0f;
#pragma unroll 1
                for (int gi = tid; gi < ngroups; gi += SCORE_NT) {
                    const int ly = (gi * twg_magic) >> 16, lx = (gi - ly * TWg) * VEC;     // gi / TWg (exact: gi < 1024, TWg <= 32)
                    if (y0 + ly >= H || x0 + lx >= W) continue;
                    score_group<VEC, HM>(sum, op, tab, a.L, p, sS + (ly + 1) * spitch + HX + lx, spitch, srows * spitch,
                                         sT + ly * TW + lx, TH * TW, mimg ? mimg + (size_t)(y0 + ly) * W + x0 + lx : nullptr, mcs,
                                         clamped_in);
                }
                sum = warp_sum(sum);
                if (lane == 0) sh.wsum[c][warp] = sum;
            }
            __syncthreads();
            T2O_PROBE(2)
            // this tile's partial of every live fit -> every CTA of the cluster (st.async completing the receiver's mbarrier)
            if (tid < m && sh.pop[tid] != OP_SKIP) {
                float v = 0.0f;
#pragma unroll
                for (int w = 0; w < SCORE_NW; ++w) v += sh.wsum[tid][w];
                const uint32_t pv = __float_as_uint(v);
                if (single) st_local_u32(&sh.part[pb][tid][0], pv, &sh.sbar[pb]);
                else
                    for (int r = 0; r < a.ntiles; ++r) st_async_u32(mapa_u32(&sh.part[pb][tid][rank], r), pv, mapa_u32(&sh.sbar[pb], r));
            }
            // the warps that own a live fit wait for the cluster's partials, sum them in tile order and step their fit -- in
            // every CTA alike; the others go on to the barrier behind the next tables (which costs no issue slots)
            if (warp < m && mylive) {
                mbar_wait_or_trap(&sh.sbar[pb], pb ? sphase1 : sphase0);
                T2O_PROBE(3)
                float v = 0.0f;
                for (int t = lane; t < a.ntiles; t += 32) v += sh.part[pb][warp][t];
                v = warp_sum(v);
                NMWarp w;
                res_bind(w, nmreg + sh.nmoff[warp], sh.nmN[warp], nm, cbeg + warp, lane, sh.pend[warp], &sh.pop[warp], rank == 0);
                nm_set_regs(w, nr);
                nm_step(w, nm, v);
                nr = nm_regs(w);
                __syncwarp();
            }
            if (pb) sphase1 ^= 1u; else sphase0 ^= 1u;
            T2O_PROBE(4)
#ifdef T2O_RES_PROBE
            pc[5] += 1;
#endif
        }
#ifdef T2O_RES_PROBE
        // per-phase clocks of state 0's cluster: lane 0 of every warp of CTAs 0 and 1
        if (probe && s == 0 && rank < 2 && lane == 0)
            for (int i = 0; i < 6; ++i) probe[(rank * SCORE_NW + warp) * 6 + i] = (unsigned long long)pc[i];
        if (probe && s == 0 && rank == 0 && tid < 12) probe[128 + tid] = g_nmp[tid];      // (running totals of the launch so far)
#endif
        // ---- leader: the fits' state back into the caller's arrays (finished, or stopped by max_rounds: the rounds can go on)
        if (rank == 0 && warp < m && sh.nmoff[warp] >= 0) {
            __syncwarp();                                        // (the lanes read what other lanes of the warp wrote in the last step)
            const int p = cbeg + warp, N = sh.nmN[warp];
            NMWarp w;
            res_bind(w, nmreg + sh.nmoff[warp], N, nm, p, lane, sh.pend[warp], &sh.pop[warp], true);
            nm_set_regs(w, nr);
            nm_store(w);                                         // registers -> the shared-memory copy that goes back below
            double *gsim = nm.st.sim + (size_t)p * NM_ROWS * NM_MAXN, *gvec = nm.st.vec + (size_t)p * 3 * NM_MAXN;
            if (lane < N) {
                for (int r = 0; r <= N; ++r) gsim[r * NM_MAXN + lane] = w.sim[r * N + lane];
                for (int r = 0; r < 3; ++r) gvec[r * NM_MAXN + lane] = w.vec[r * N + lane];
            }
            if (lane <= N) {
                nm.st.fsim[(size_t)p * NM_ROWS + lane] = w.fsim[lane];
                nm.st.perm[(size_t)p * NM_ROWS + lane] = w.perm[lane];
            }
            if (lane < 8) nm.st.ctl[(size_t)p * 8 + lane] = w.ctl[lane];
            if (lane == 0) { nm.st.fxr[p] = *w.fxr; nm.cand_op[p] = sh.pop[warp]; }
            if (lane < NM_MAXN) nm.cand_param[(size_t)p * NM_MAXN + lane] = sh.pend[warp][lane];
        }
    }
}

// ------------------------------------------------------------------------------------------- host
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn lookup_encode_fn() {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPointByVersion("cuTensorMapEncodeTiled", &p, 12000, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
        return (EncodeTiledFn)p;
    (void)cudaGetLastError();
    return nullptr;
}
// (a function-local static: initialised once, also when two planner threads make their first call together)
static EncodeTiledFn get_encode_fn() {
    static const EncodeTiledFn fn = lookup_encode_fn();
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

// Every fit of a planner step in one launch (see nm_resident_kernel).  T2O_ERR_UNSUPPORTED: the shape is not eligible (the
// caller then runs the rounds itself).  fits_begin[s] .. fits_begin[s+1] are the fits of state s (at most 8); h_fits_begin and
// h_fit_op are HOST copies of fits_begin and of the fits' operators: they size the shared memory of the launch.
int nm_run_resident(const float *states, int S, const float *targets, int T, const int *state_target, const int *fits_begin,
                    const int *cand_mask, const float *masks, int n_masks, int mask_ch,
                    const t2o_nm_state *st, int P, float numel, float *cand_param, int *cand_op,
                    const int *h_fits_begin, const int *h_fit_op,
                    int H, int W, int L, int max_rounds, void *ws, size_t ws_bytes, cudaStream_t stream) {
    if (!states || !targets || !fits_begin || !st || !cand_param || !cand_op || !h_fits_begin || !h_fit_op) return T2O_ERR_INVALID_ARG;
    if (S < 1 || T < 1 || P < 1 || H < 1 || W < 1 || !(numel > 0.0f) || max_rounds < 1) return T2O_ERR_INVALID_ARG;
    if (L < 1 || L > MAX_L) return T2O_ERR_UNSUPPORTED;
    if (!ws || ws_bytes < COUNTER_REGION) return T2O_ERR_WORKSPACE;
    const bool hm = masks != nullptr;
    if (hm && (!cand_mask || n_masks < 1 || (mask_ch != 1 && mask_ch != 3))) return T2O_ERR_INVALID_ARG;
    const bool aligned = ((uintptr_t)states % 16 == 0) && ((uintptr_t)targets % 16 == 0) && (!hm || (uintptr_t)masks % 16 == 0);
    if (W % 4 != 0 || !aligned || get_encode_fn() == nullptr) return T2O_ERR_UNSUPPORTED;
    // shared memory of a state's fits: operator tables and Nelder-Mead state, the largest need over the states
    int tab_cap = 0, nm_cap = 0;
    if (h_fits_begin[0] != 0 || h_fits_begin[S] != P) return T2O_ERR_INVALID_ARG;
    for (int s = 0; s < S; ++s) {
        const int b = h_fits_begin[s], e = h_fits_begin[s + 1];
        if (e < b || e - b > SCORE_NW) return e < b ? T2O_ERR_INVALID_ARG : T2O_ERR_UNSUPPORTED;
        int tf = 0, nb = 0;
        for (int c = b; c < e; ++c) {
            const int n = t2o_num_params(h_fit_op[c], L);
            if (n < 1 || n > NM_MAXN) return T2O_ERR_INVALID_ARG;
            tf += res_tab_floats(h_fit_op[c]);
            nb += res_nm_bytes(n);
        }
        tab_cap = tf > tab_cap ? tf : tab_cap;
        nm_cap = nb > nm_cap ? nb : nm_cap;
    }
    ScoreArgs a;
    memset(&a, 0, sizeof(a));
    a.states = states; a.targets = targets; a.cand_param = cand_param; a.state_target = state_target;
    a.cand_begin = fits_begin; a.cand_op = cand_op;
    a.masks = masks; a.cand_mask = hm ? cand_mask : nullptr; a.mask_ch = hm ? mask_ch : 0;
    a.S = S; a.T = T; a.C = P; a.H = H; a.W = W; a.L = L;
    a.TW = W < 128 ? W : 128;
    a.TH = H < 32 ? H : 32;
    a.tiles_x = (W + a.TW - 1) / a.TW;
    a.ntiles = a.tiles_x * ((H + a.TH - 1) / a.TH);
    a.nsplit = 1;
    if (a.ntiles > RES_MAXT || a.TW + 8 > 256 || a.TH + 2 > 256) return T2O_ERR_UNSUPPORTED;
    NMArgs nm;
    memset(&nm, 0, sizeof(nm));
    nm.st = *st; nm.P = P; nm.numel = numel; nm.cand_param = cand_param; nm.cand_op = cand_op;
    nm.nonz_scale = 1 + 0.05; nm.zdelt = 0.00025; nm.xatol = 1e-4; nm.fatol = 1e-4;
    const size_t s_floats = (size_t)3 * (a.TH + 2) * (a.TW + 8), t_floats = (size_t)3 * a.TH * a.TW;
    const size_t smem = align_up(s_floats * 4, 128) + t_floats * 4 + (size_t)tab_cap * 4 + (size_t)nm_cap + 128;
    auto kernel = hm ? nm_resident_kernel<true> : nm_resident_kernel<false>;
    cudaFuncAttributes fa;
    T2O_CUDA_OK(cudaFuncGetAttributes(&fa, kernel));
    if (smem + fa.sharedSizeBytes > (size_t)227 * 1024) return T2O_ERR_UNSUPPORTED;
    CUtensorMap tms, tmt;
    memset(&tms, 0, sizeof(tms)); memset(&tmt, 0, sizeof(tmt));
    if (!make_map(&tms, states, S, H, W, a.TW + 8, a.TH + 2) || !make_map(&tmt, targets, T, H, W, a.TW, a.TH)) return T2O_ERR_NO_DEVICE;
    // the work counter: the last word of the workspace's counter region (no entry point has 65 536 images or states); it is
    // zero on entry and left at zero, like every other counter there
    unsigned int *counter = (unsigned int *)ws + (COUNTER_REGION / sizeof(unsigned int) - 1);
    T2O_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (a.ntiles > 8 && cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
        (void)cudaGetLastError();
        return T2O_ERR_UNSUPPORTED;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)a.ntiles; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(SCORE_NT); cfg.dynamicSmemBytes = smem; cfg.stream = stream; cfg.attrs = attr; cfg.numAttrs = 1;
    cfg.gridDim = dim3((unsigned)a.ntiles);
    int nclusters = 0;
    if (cudaOccupancyMaxActiveClusters(&nclusters, kernel, &cfg) != cudaSuccess || nclusters < 1) { (void)cudaGetLastError(); return T2O_ERR_UNSUPPORTED; }
    if (nclusters > S) nclusters = S;
    cfg.gridDim = dim3((unsigned)(nclusters * a.ntiles));
    // (development builds with -DT2O_RES_PROBE leave per-phase clocks behind the counter region)
    unsigned long long *probe = ws_bytes >= COUNTER_REGION + 4096 ? (unsigned long long *)((unsigned char *)ws + COUNTER_REGION) : nullptr;
    T2O_CUDA_OK(cudaLaunchKernelEx(&cfg, kernel, tms, tmt, a, nm, counter, max_rounds, tab_cap, nm_cap, probe));
    T2O_CUDA_OK(cudaGetLastError());
    T2O_CUDA_OK(cudaMemsetAsync(counter, 0, sizeof(unsigned int), stream));
    return T2O_OK;
}

}  // namespace t2o
