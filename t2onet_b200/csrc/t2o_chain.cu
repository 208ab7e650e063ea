// t2o_chain.cu -- host side of the fused operator-chain kernels: geometry, workspace, launch selection.
// The kernels live in t2o_chain_kernels.cuh; the backward instantiations are spread over t2o_chain_bwd_*.cu.
#include <cstdlib>

#include "t2o_chain_kernels.cuh"

namespace t2o {

// =========================================================================================== host side
static thread_local char g_cuda_err[256] = "";
void set_cuda_error(cudaError_t e) { snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", cudaGetErrorName(e), cudaGetErrorString(e)); }
const char *last_cuda_error() { return g_cuda_err; }

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// workspace: [counters: B x u32][L1 partials: B x maxtiles][grad partials: B x maxtiles x pstride]
// upper bound of partials per image for every tiling a launcher can pick: forward tiles are >= 32 x 32 px or
// >= 1024 px flat; step-kernel chunks are >= 1024 px flat or (strip of >= 28 groups) x (band of >= 16 rows)
size_t max_tiles(int H, int W) { return ((size_t)W / 28 + 2) * ((size_t)H / 16 + 2); }
size_t chain_workspace_bytes(int B, int H, int W, int pstride) {
    const size_t mt = max_tiles(H, W);
    return COUNTER_REGION + align_up((size_t)B * mt * 4, 256) + align_up((size_t)B * mt * (size_t)(pstride > 0 ? pstride : 1) * 4, 256);
}
struct Workspace {
    unsigned int *counters;
    float *part_l1, *part_gp;
};
Workspace carve_workspace(void *ws, int B, int H, int W) {
    const size_t mt = max_tiles(H, W);
    char *p = (char *)ws;
    Workspace w;
    w.counters = (unsigned int *)p; p += COUNTER_REGION;             // B <= 65535 counters
    w.part_l1 = (float *)p; p += align_up((size_t)B * mt * 4, 256);
    w.part_gp = (float *)p;
    return w;
}

static int make_desc(int n_ops, const int *op_ids, const int *param_off, int L, int pstride, ChainDesc &d) {
    if (!param_off) return T2O_ERR_INVALID_ARG;
    return build_chain_desc(n_ops, op_ids, param_off, 0, L, pstride, d);
}

static int pick_vec_flat(const void *const *ptrs, int nptr, size_t plane) {
    int vec = (plane % 4 == 0) ? 4 : (plane % 2 == 0 ? 2 : 1);
    for (int i = 0; i < nptr; ++i) {
        if (!ptrs[i]) continue;
        const uintptr_t a = (uintptr_t)ptrs[i];
        while (vec > 1 && (a % (vec * 4)) != 0) vec >>= 1;
    }
    return vec;
}


// flat tiling (no stencil): contiguous ranges of groups
static void geom_flat(Geom &g, int B, int H, int W, int vec) {
    memset(&g, 0, sizeof(g));
    g.B = B; g.H = H; g.W = W;
    const long long plane = (long long)H * W;
    g.ngroups = plane / vec;
    const long long total_px = plane * B;
    long long tile_px = total_px / (NUM_SMS * 8);
    const long long quantum = (long long)NT * vec * 2;
    tile_px = (tile_px + quantum - 1) / quantum * quantum;
    if (tile_px < MIN_TILE_PX) tile_px = MIN_TILE_PX;
    if (tile_px > 16384) tile_px = 16384;
    if (tile_px < quantum) tile_px = quantum;
    g.tile_groups = (int)(tile_px / vec);
    g.ntiles = (int)((g.ngroups + g.tile_groups - 1) / g.tile_groups);
    if (g.ntiles < 1) g.ntiles = 1;
    g.nchunks = g.ntiles;
    g.tiles_per_cta = 1;
}

template <int VEC, bool HM, bool ROWS>
static int launch_fwd(const FwdArgs &a, cudaStream_t stream) {
    dim3 grid(a.g.ntiles, a.g.B);
    chain_fwd_kernel<VEC, HM, ROWS><<<grid, NT, 0, stream>>>(a);
    T2O_CUDA_OK(cudaGetLastError());
    return T2O_OK;
}
template <bool HM, bool ROWS>
static int launch_fwd_vec(int vec, const FwdArgs &a, cudaStream_t stream) {
    if (vec == 4) return launch_fwd<4, HM, ROWS>(a, stream);
    if (vec == 2) return launch_fwd<2, HM, ROWS>(a, stream);
    return launch_fwd<1, HM, ROWS>(a, stream);
}

void geom_step_rows(StepGeom &g, int B, int H, int W, int vec, int slots, int SNT, int rows_extra);

template <int VEC, bool HM, bool ROWS, unsigned int SP>
static int launch_fwd_rows_sp(FwdRowsArgs &a, int B, int H, int W, cudaStream_t stream) {
    const size_t smem = (size_t)(2 * (NT / 32) + 2) * 3 * 34 * VEC * sizeof(float);        // the X ring
    static int resident = 0;
    if (resident == 0) {
        if (smem > 40 * 1024)
            T2O_CUDA_OK(cudaFuncSetAttribute(chain_fwd_rows_kernel<VEC, HM, ROWS, SP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, chain_fwd_rows_kernel<VEC, HM, ROWS, SP>, NT, smem) != cudaSuccess || nb < 1) nb = 1;
        resident = nb;
    }
    geom_step_rows(a.g, B, H, W, VEC, NUM_SMS * resident, NT, 2);
    dim3 grid(a.g.nchunks, B);
    chain_fwd_rows_kernel<VEC, HM, ROWS, SP><<<grid, NT, smem, stream>>>(a);
    T2O_CUDA_OK(cudaGetLastError());
    return T2O_OK;
}
template <int VEC, bool HM, bool ROWS>
static int launch_fwd_rows(FwdRowsArgs &a, int B, int H, int W, cudaStream_t stream) {
    if constexpr (VEC == 4 && !HM && !ROWS) {      // chain-specialised instantiation (unmasked, 128-bit groups)
        const char *e = getenv("T2O_NO_SPECIALIZED");
        if (!(e && e[0] == '1') && a.ch.n <= MAX_CHAIN) {
            unsigned int packed = 0u;
            for (int k = 0; k < a.ch.n; ++k) packed |= (unsigned)(a.ch.op[k] + 1) << (4 * k);
            if (packed == SP_C6) return launch_fwd_rows_sp<VEC, HM, ROWS, SP_C6>(a, B, H, W, stream);
            if (packed == SP_S1) return launch_fwd_rows_sp<VEC, HM, ROWS, SP_S1>(a, B, H, W, stream);
            if (packed == SP_B1) return launch_fwd_rows_sp<VEC, HM, ROWS, SP_B1>(a, B, H, W, stream);
        }
    }
    return launch_fwd_rows_sp<VEC, HM, ROWS, 0u>(a, B, H, W, stream);
}
template <bool HM, bool ROWS>
static int launch_fwd_rows_vec(int vec, FwdRowsArgs &a, int B, int H, int W, cudaStream_t stream) {
    if (vec == 4) return launch_fwd_rows<4, HM, ROWS>(a, B, H, W, stream);
    if (vec == 2) return launch_fwd_rows<2, HM, ROWS>(a, B, H, W, stream);
    return launch_fwd_rows<1, HM, ROWS>(a, B, H, W, stream);
}

int chain_forward(int n_ops, const int *op_ids, const int *param_off, const float *img, const float *mask, int mask_ch,
                  const float *params, int pstride, const float *target, float *out, float *l1_sum,
                  int B, int H, int W, int L, int flags, void *ws, size_t ws_bytes, cudaStream_t stream) {
    FwdArgs a;
    memset(&a, 0, sizeof(a));
    if ((flags & ~T2O_FLAG_RAW_PROCESS) != 0 || ((flags & T2O_FLAG_RAW_PROCESS) && n_ops != 1)) return T2O_ERR_INVALID_ARG;
    a.raw = (flags & T2O_FLAG_RAW_PROCESS) ? 1 : 0;
    int st = make_desc(n_ops, op_ids, param_off, L, pstride, a.ch);
    if (st != T2O_OK) return st;
    if (!img || B < 1 || H < 1 || W < 1 || (pstride > 0 && !params)) return T2O_ERR_INVALID_ARG;
    if (mask && mask_ch != 1 && mask_ch != 3) return T2O_ERR_INVALID_ARG;
    if (l1_sum && !target) return T2O_ERR_INVALID_ARG;
    if (!out && !l1_sum) return T2O_ERR_INVALID_ARG;
    if (l1_sum && (!ws || ws_bytes < chain_workspace_bytes(B, H, W, pstride))) return T2O_ERR_WORKSPACE;
    if (B > 65535) return T2O_ERR_UNSUPPORTED;
    Workspace w = carve_workspace(ws, B, H, W);
    a.img = img; a.mask = mask; a.params = params; a.target = target; a.out = out; a.l1_sum = l1_sum;
    a.part_l1 = w.part_l1; a.counters = w.counters; a.mask_ch = mask_ch; a.pstride = pstride;
    const size_t plane = (size_t)H * W;
    const void *ptrs[] = {img, mask, target, out};
    int vec = pick_vec_flat(ptrs, 4, plane);
    if (a.ch.sharp < 0) {
        geom_flat(a.g, B, H, W, vec);
        return mask ? launch_fwd_vec<true, false>(vec, a, stream) : launch_fwd_vec<false, false>(vec, a, stream);
    }
    while (vec > 1 && W % vec != 0) vec >>= 1;
    FwdRowsArgs r;
    memset(&r, 0, sizeof(r));
    r.ch = a.ch;
    r.img = img; r.mask = mask; r.params = params; r.target = target; r.out = out; r.l1_sum = l1_sum;
    r.part_l1 = w.part_l1; r.counters = w.counters; r.mask_ch = mask_ch; r.pstride = pstride; r.raw = a.raw;
    r.clamped = chain_clamped_bits(a.ch.op, n_ops);
    return mask ? launch_fwd_rows_vec<true, false>(vec, r, B, H, W, stream) : launch_fwd_rows_vec<false, false>(vec, r, B, H, W, stream);
}

// Host-side check of per-row chains when the caller also has the operator ids on the host: validates every row and
// reports which tilings are needed (bit 0: rows without a stencil, bit 1: rows with one).  Without a host copy only
// single-operator rows are accepted (nothing but the id range can be wrong; the kernels flag that in *status).
int rows_paths(int K, const int *row_ops_host, int B, int slot, int L, int pstride, bool backward, int &paths) {
    if (K < 1 || K > MAX_CHAIN || slot < 1 || (long long)K * slot > pstride) return T2O_ERR_INVALID_ARG;
    if (pstride > MAX_PSTRIDE || L < 1 || L > MAX_L || slot < 3 * L) return T2O_ERR_UNSUPPORTED;
    if (!row_ops_host) {
        if (K != 1) return T2O_ERR_INVALID_ARG;
        paths = 3;
        return T2O_OK;
    }
    paths = 0;
    for (int b = 0; b < B; ++b) {
        int st;
        int sharp;
        if (backward) {
            StepDesc d;
            st = build_step_desc(K, row_ops_host + (size_t)b * K, nullptr, slot, L, pstride, d);
            sharp = d.sharp;
        } else {
            ChainDesc d;
            st = build_chain_desc(K, row_ops_host + (size_t)b * K, nullptr, slot, L, pstride, d);
            sharp = d.sharp;
        }
        if (st != T2O_OK) return st;
        paths |= sharp >= 0 ? 2 : 1;
    }
    return T2O_OK;
}

// Per-row chains, forward (Actor call sites models/actor.py:165,252,340: one operator per batch row and decoding step).
int rows_forward(int K, const int *row_ops, const int *row_ops_host, int slot, const float *img, const float *mask, int mask_ch,
                 const float *params, int pstride, const float *target, float *out, float *l1_sum, unsigned int *status,
                 int B, int H, int W, int L, void *ws, size_t ws_bytes, cudaStream_t stream) {
    if (!row_ops || !img || !params || B < 1 || H < 1 || W < 1) return T2O_ERR_INVALID_ARG;
    if (mask && mask_ch != 1 && mask_ch != 3) return T2O_ERR_INVALID_ARG;
    if (l1_sum && !target) return T2O_ERR_INVALID_ARG;
    if (!out && !l1_sum) return T2O_ERR_INVALID_ARG;
    if (l1_sum && (!ws || ws_bytes < chain_workspace_bytes(B, H, W, pstride))) return T2O_ERR_WORKSPACE;
    if (B > 65535) return T2O_ERR_UNSUPPORTED;
    int paths = 0;
    int st = rows_paths(K, row_ops_host, B, slot, L, pstride, false, paths);
    if (st != T2O_OK) return st;
    Workspace w = carve_workspace(ws, B, H, W);
    const size_t plane = (size_t)H * W;
    const void *ptrs[] = {img, mask, target, out};
    int vec = pick_vec_flat(ptrs, 4, plane);
    if (paths & 2)
        while (vec > 1 && W % vec != 0) vec >>= 1;
    if (paths & 1) {
        FwdArgs a;
        memset(&a, 0, sizeof(a));
        a.ch.n = K; a.ch.L = L; a.ch.sharp = -1;
        a.img = img; a.mask = mask; a.params = params; a.target = target; a.out = out; a.l1_sum = l1_sum;
        a.part_l1 = w.part_l1; a.counters = w.counters; a.mask_ch = mask_ch; a.pstride = pstride;
        a.row_ops = row_ops; a.rows_K = K; a.rows_slot = slot; a.status = status;
        geom_flat(a.g, B, H, W, vec);
        st = mask ? launch_fwd_vec<true, true>(vec, a, stream) : launch_fwd_vec<false, true>(vec, a, stream);
        if (st != T2O_OK) return st;
    }
    if (paths & 2) {
        FwdRowsArgs r;
        memset(&r, 0, sizeof(r));
        r.ch.n = K; r.ch.L = L; r.ch.sharp = -1;
        r.img = img; r.mask = mask; r.params = params; r.target = target; r.out = out; r.l1_sum = l1_sum;
        r.part_l1 = w.part_l1; r.counters = w.counters; r.mask_ch = mask_ch; r.pstride = pstride;
        r.row_ops = row_ops; r.rows_K = K; r.rows_slot = slot; r.status = status;
        st = mask ? launch_fwd_rows_vec<true, true>(vec, r, B, H, W, stream) : launch_fwd_rows_vec<false, true>(vec, r, B, H, W, stream);
        if (st != T2O_OK) return st;
    }
    return T2O_OK;
}

int l1_sum_launch(const float *pa, const float *pb, float *l1_sum, int B, long long n, void *ws, size_t ws_bytes,
                  cudaStream_t stream) {
    if (!pa || !pb || !l1_sum || B < 1 || n < 1) return T2O_ERR_INVALID_ARG;
    if (B > 65535) return T2O_ERR_UNSUPPORTED;
    L1Args a;
    const void *ptrs[] = {pa, pb};
    const int vec = pick_vec_flat(ptrs, 2, (size_t)n);
    long long tile = (n * B) / (NUM_SMS * 8);
    const long long quantum = (long long)NT * 4;
    tile = (tile + quantum - 1) / quantum * quantum;
    if (tile < 4096) tile = 4096;
    if (tile > 65536) tile = 65536;
    a.ntiles = (int)((n + tile - 1) / tile);
    const size_t need = COUNTER_REGION + (size_t)B * a.ntiles * 4;
    if (!ws || ws_bytes < need) return T2O_ERR_WORKSPACE;
    a.a = pa; a.b = pb; a.l1_sum = l1_sum; a.n = n; a.tile_elems = tile;
    a.counters = (unsigned int *)ws;
    a.part = (float *)((char *)ws + COUNTER_REGION);
    dim3 grid(a.ntiles, B);
    if (vec == 4) l1_sum_kernel<4><<<grid, NT, 0, stream>>>(a);
    else if (vec == 2) l1_sum_kernel<2><<<grid, NT, 0, stream>>>(a);
    else l1_sum_kernel<1><<<grid, NT, 0, stream>>>(a);
    T2O_CUDA_OK(cudaGetLastError());
    return T2O_OK;
}

}  // namespace t2o
