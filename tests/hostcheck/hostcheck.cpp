// hostcheck.cpp -- TEST INFRASTRUCTURE ONLY.
// Runs the product's per-pixel arithmetic (t2onet_b200/csrc/t2o_math.cuh, the code the CUDA
// kernels inline) on the CPU over whole images, so that the closed forms and hand-derived
// gradients can be checked against the oracle in the GPU-less authoring container.  It stores
// every intermediate image (no tiling, no recompute) and is never loaded by the product.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../t2onet_b200/csrc/t2o_math.cuh"

using namespace t2o;

// S(x) of the stencil operators at one pixel of a plane (zero padding): laplace for sharpness, (G - delta) for blur
static inline float stencil_at(int op, const float *pc, int H, int W, int yy, int xx) {
    auto at = [&](int y, int x) { return (y >= 0 && y < H && x >= 0 && x < W) ? pc[(size_t)y * W + x] : 0.f; };
    const float ctr = at(yy, xx), up = at(yy - 1, xx), dn = at(yy + 1, xx), lf = at(yy, xx - 1), rt = at(yy, xx + 1);
    if (op == t2o::OP_BLUR)
        return t2o::blur_delta(ctr, (up + dn) + (lf + rt), (at(yy - 1, xx - 1) + at(yy - 1, xx + 1)) + (at(yy + 1, xx - 1) + at(yy + 1, xx + 1)));
    return t2o::laplace(ctr, up, dn, lf, rt);
}

extern "C" int hc_chain(int n_ops, const int *ops, const int *poff, const float *img, const float *mask, int mask_ch,
                        const float *params, int pstride, const float *grad_out, const float *target,
                        const float *grad_l1, float *out, float *l1_sum, float *grad_params, float *grad_img,
                        int B, int H, int W, int L) {
    const size_t plane = (size_t)H * W;
    const bool has_mask = mask != nullptr;
    std::vector<float> tabs(MAX_CHAIN * TAB);
    std::vector<std::vector<float>> xs(n_ops + 1, std::vector<float>(3 * plane));
    std::vector<float> g(3 * plane), gn(3 * plane), gyv(3 * plane);
    for (int b = 0; b < B; ++b) {
        const float *ib = img + (size_t)b * 3 * plane;
        const float *mb = has_mask ? mask + (size_t)b * mask_ch * plane : nullptr;
        auto M = [&](int c, size_t i) { return has_mask ? mb[(mask_ch == 3 ? c : 0) * plane + i] : 1.0f; };
        for (int k = 0; k < n_ops; ++k) build_table(ops[k], params + (size_t)b * pstride + poff[k], L, &tabs[k * TAB]);
        std::memcpy(xs[0].data(), ib, 3 * plane * sizeof(float));
        // forward
        for (int k = 0; k < n_ops; ++k) {
            const float *x = xs[k].data();
            float *y = xs[k + 1].data();
            const float *tab = &tabs[k * TAB];
            if (op_is_stencil(ops[k])) {
                for (int c = 0; c < 3; ++c)
                    for (int yy = 0; yy < H; ++yy)
                        for (int xx = 0; xx < W; ++xx) {
                            const size_t i = (size_t)yy * W + xx;
                            const float *pc = x + c * plane;
                            const float ctr = pc[i];
                            const float v = fmaf(tab[0], stencil_at(ops[k], pc, H, W, yy, xx), ctr);
                            y[c * plane + i] = sat01(has_mask ? blend<true>(v, ctr, M(c, i)) : v);
                        }
            } else {
                for (size_t i = 0; i < plane; ++i) {
                    float r = x[i], gg = x[plane + i], bb = x[2 * plane + i];
                    const bool cl = k > 0 && ops[k - 1] >= 0;      // input = clamped output of the previous operator
                    if (has_mask) { if (cl) op_apply<true, true>(ops[k], tab, L, r, gg, bb, M(0, i), M(1, i), M(2, i)); else op_apply<true, false>(ops[k], tab, L, r, gg, bb, M(0, i), M(1, i), M(2, i)); }
                    else { if (cl) op_apply<false, true>(ops[k], tab, L, r, gg, bb, 1.f, 1.f, 1.f); else op_apply<false, false>(ops[k], tab, L, r, gg, bb, 1.f, 1.f, 1.f); }
                    y[i] = r; y[plane + i] = gg; y[2 * plane + i] = bb;
                }
            }
        }
        const float *fin = xs[n_ops].data();
        if (out) std::memcpy(out + (size_t)b * 3 * plane, fin, 3 * plane * sizeof(float));
        if (l1_sum && target) {
            double s = 0;
            for (size_t i = 0; i < 3 * plane; ++i) s += fabsf(fin[i] - target[(size_t)b * 3 * plane + i]);
            l1_sum[b] = (float)s;
        }
        if (!grad_params && !grad_img) continue;
        // upstream gradient
        for (size_t i = 0; i < 3 * plane; ++i) {
            if (grad_out) g[i] = grad_out[(size_t)b * 3 * plane + i];
            else {
                const float d = fin[i] - target[(size_t)b * 3 * plane + i];
                g[i] = grad_l1[b] * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
            }
        }
        if (grad_params) std::memset(grad_params + (size_t)b * pstride, 0, pstride * sizeof(float));
        for (int k = n_ops - 1; k >= 0; --k) {
            const float *x = xs[k].data();
            const float *tab = &tabs[k * TAB];
            std::vector<double> accd(ACC_SLOTS, 0.0);
            if (op_is_stencil(ops[k])) {
                const float p = tab[0];
                double accp = 0;
                for (int c = 0; c < 3; ++c)
                    for (int yy = 0; yy < H; ++yy)
                        for (int xx = 0; xx < W; ++xx) {
                            const size_t i = (size_t)yy * W + xx;
                            const float *pc = x + c * plane;
                            const float ctr = pc[i];
                            const float lap = stencil_at(ops[k], pc, H, W, yy, xx);
                            float gy, gd;
                            if (has_mask) blend_bwd<true>(fmaf(p, lap, ctr), ctr, M(c, i), g[c * plane + i], gy, gd);
                            else blend_bwd<false>(fmaf(p, lap, ctr), ctr, 1.f, g[c * plane + i], gy, gd);
                            gyv[c * plane + i] = gy;
                            gn[c * plane + i] = gd;
                            accp += (double)gy * lap;
                        }
                for (int c = 0; c < 3; ++c)
                    for (int yy = 0; yy < H; ++yy)
                        for (int xx = 0; xx < W; ++xx) {
                            const size_t i = (size_t)yy * W + xx;
                            const float *pg = gyv.data() + c * plane;
                            gn[c * plane + i] += pg[i] + p * stencil_at(ops[k], pg, H, W, yy, xx);   // symmetric stencil: S^T = S
                        }
                if (grad_params) grad_params[(size_t)b * pstride + poff[k]] = (float)accp;
                g.swap(gn);
                continue;
            }
            for (size_t i = 0; i < plane; ++i) {
                float gr = g[i], gg = g[plane + i], gb = g[2 * plane + i];
                GradAcc A;
                acc_zero(A);
                const bool cl = k > 0 && ops[k - 1] >= 0;
                const float xr = x[i], xg = x[plane + i], xb = x[2 * plane + i];
                if (has_mask) {
                    if (cl) pointwise_bwd<true, true>(ops[k], tab, L, xr, xg, xb, M(0, i), M(1, i), M(2, i), gr, gg, gb, A, true);
                    else pointwise_bwd<true, false>(ops[k], tab, L, xr, xg, xb, M(0, i), M(1, i), M(2, i), gr, gg, gb, A, true);
                } else if (op_is_curve(ops[k]) && curve_in_range(ops[k], tab)) {       // as bwd_op_grp dispatches
                    if (cl) pointwise_bwd<false, true, true>(ops[k], tab, L, xr, xg, xb, 1.f, 1.f, 1.f, gr, gg, gb, A, true);
                    else pointwise_bwd<false, false, true>(ops[k], tab, L, xr, xg, xb, 1.f, 1.f, 1.f, gr, gg, gb, A, true);
                } else {
                    if (cl) pointwise_bwd<false, true>(ops[k], tab, L, xr, xg, xb, 1.f, 1.f, 1.f, gr, gg, gb, A, true);
                    else pointwise_bwd<false, false>(ops[k], tab, L, xr, xg, xb, 1.f, 1.f, 1.f, gr, gg, gb, A, true);
                }
                g[i] = gr; g[plane + i] = gg; g[2 * plane + i] = gb;
                float v[ACC_SLOTS];
                acc_to_slots(A, v);
                for (int t = 0; t < ACC_SLOTS; ++t) accd[t] += v[t];
            }
            if (grad_params) {
                float *gp = grad_params + (size_t)b * pstride + poff[k];
                switch (ops[k]) {
                    case OP_TONE: {
                        float G[MAX_L];
                        for (int i = 0; i < MAX_L; ++i) G[i] = (float)accd[ACC_TONE + i];
                        for (int i = 0; i < L; ++i) gp[i] = curve_param_grad(tab, L, G, i);
                        break;
                    }
                    case OP_COLOR:
                        for (int c = 0; c < 3; ++c) {
                            float G[MAX_L];
                            for (int i = 0; i < MAX_L; ++i) G[i] = (float)accd[ACC_COLOR + c * MAX_L + i];
                            for (int i = 0; i < L; ++i) gp[c * L + i] = curve_param_grad(tab + c * CT, L, G, i);
                        }
                        break;
                    case OP_BRIGHTNESS: gp[0] = (float)accd[ACC_BRIGHT]; break;
                    case OP_CONTRAST: gp[0] = (float)accd[ACC_CONTRAST]; break;
                    case OP_SATURATION: gp[0] = (float)accd[ACC_SATUR]; break;
                    case OP_EXPOSURE: gp[0] = (float)accd[ACC_EXPO]; break;
                    case OP_BNW: gp[0] = (float)accd[ACC_BNW]; break;
                    case OP_HUE: gp[0] = (float)accd[ACC_HUE]; break;
                    case OP_WHITEBALANCE: for (int t = 0; t < 3; ++t) gp[t] = (float)accd[ACC_WB + t]; break;
                    default: break;
                }
            }
        }
        if (grad_img) std::memcpy(grad_img + (size_t)b * 3 * plane, g.data(), 3 * plane * sizeof(float));
    }
    return 0;
}
