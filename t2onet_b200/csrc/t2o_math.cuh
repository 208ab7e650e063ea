// t2o_math.cuh -- per-pixel arithmetic of the T2ONet global editing operators (forward and
// backward), shared by every kernel.  Host+device so the same code can be checked on a CPU
// against the oracle in this GPU-less authoring container (tests/hostcheck); the product only
// ever runs it inside the sm_100a kernels.
//
// Reference semantics (file:line in /root/reference):
//   Operator.execute        models/operators.py:112-131   y = process(x,p); z = y*m + x*(1-m); out = clamp(z,0,1)
//   Brightness / Saturation models/operators.py:277-283 / 473-479 over kornia rgb_to_hsv / hsv_to_rgb
//   Contrast                models/operators.py:240-245   (rgb2lum utils/operator_utils.py:9, lerp :5)
//   Tone / Color curves     models/operators.py:571-585 / 607-616
//   Sharpness               models/operators.py:351-358   (3x3 Laplacian, zero padding)
//   Exposure / WB / White   models/operators.py:209-210 / 548-549 / 510-512
//
// Closed forms used instead of the reference's op-by-op evaluation:
//   HSV round trip:   with v = max, mn = min, d = v - mn, u_c = v - c, s = d/(v+eps):
//       brightness    y_c = v' * (1 - u_c/(v+eps)),            v' = clamp(v(1+p), 0, 1)
//       saturation    y_c = v  * (1 - u_c * s'/d),             s' = clamp(s(1+p), 0, 1)   (d == 0: y_c = v)
//     (f*s of hsv_to_rgb is u_c/(v+eps) exactly; the sector select is continuous, so no branch)
//   Curves:           j = min(floor(L x), L-1);  y = k'_j x + Q_j,  k' = k L/S, Q_j = L/S sum_{i<j} k_i/L - k'_j j/L
//   Curve param grads: 2L+1 moments per curve (A_j = sum g, Bx_j = sum g x over bin j, C = sum g y)
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define T2O_HD __host__ __device__ __forceinline__
#else
#define T2O_HD inline
#endif

namespace t2o {

enum : int {
    OP_IDENTITY = -1, OP_BRIGHTNESS = 0, OP_CONTRAST = 1, OP_SATURATION = 2, OP_COLOR = 3, OP_INPAINT = 4,
    OP_TONE = 5, OP_SHARPNESS = 6, OP_WHITE = 7, OP_EXPOSURE = 8, OP_WHITEBALANCE = 9, OP_COUNT = 10
};

constexpr int MAX_CHAIN = 8;
constexpr int MAX_L = 8;
constexpr int TAB = 64;    // floats in one (image, op) table
constexpr int CT = 20;     // floats in one curve table: k'[8], Q[8], 1/S, L/S, pad
constexpr int HIST = 17;   // moments of one curve: A[8], Bx[8], C
constexpr float HSV_EPS = 1e-6f;     // kornia.rgb_to_hsv eps
constexpr float LUM_EPS = 1e-6f;     // models/operators.py:244
constexpr float CURVE_EPS = 1e-10f;  // models/operators.py:579,610
constexpr float LN2_F = 0.6931471805599453f;   // np.log(2) cast to fp32 (models/operators.py:210)
constexpr float PI_F = 3.14159265358979323846f;

T2O_HD int op_num_params(int op, int L) {
    switch (op) {
        case OP_COLOR: return 3 * L;
        case OP_TONE: return L;
        case OP_WHITEBALANCE: return 3;
        case OP_IDENTITY: return 0;
        default: return 1;
    }
}
T2O_HD bool op_is_curve(int op) { return op == OP_TONE || op == OP_COLOR; }
T2O_HD int op_hist_floats(int op) { return op == OP_TONE ? HIST : (op == OP_COLOR ? 3 * HIST : 0); }

// ---------------------------------------------------------------- scalar helpers
T2O_HD float sat01(float z) {
#if defined(__CUDA_ARCH__)
    return __saturatef(z);
#else
    return fminf(fmaxf(z, 0.0f), 1.0f);
#endif
}
T2O_HD bool in01(float z) { return z >= 0.0f && z <= 1.0f; }   // torch.clamp backward: closed interval
T2O_HD float rcp(float x) { return 1.0f / x; }
T2O_HD float fdiv(float a, float b) { return a / b; }
T2O_HD float mul_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    volatile float r = a * b; return r;
#endif
}
T2O_HD float add_rn(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    volatile float r = a + b; return r;
#endif
}
T2O_HD float cospi_f(float x) {
#if defined(__CUDA_ARCH__)
    return cospif(x);
#else
    return cosf(PI_F * x);
#endif
}
T2O_HD float sinpi_f(float x) {
#if defined(__CUDA_ARCH__)
    return sinpif(x);
#else
    return sinf(PI_F * x);
#endif
}
T2O_HD float max3(float a, float b, float c) { return fmaxf(a, fmaxf(b, c)); }
T2O_HD float min3(float a, float b, float c) { return fminf(a, fminf(b, c)); }
// first index attaining the max / min (torch max(dim)/min(dim) tie rule on the reference graph)
T2O_HD int argmax3(float r, float g, float b) { return (r >= g && r >= b) ? 0 : (g >= b ? 1 : 2); }
T2O_HD int argmin3(float r, float g, float b) { return (r <= g && r <= b) ? 0 : (g <= b ? 1 : 2); }
// rgb2lum with the reference's rounding sequence (utils/operator_utils.py:10), so that the exact
// ties lum == 0 / lum == 1 of the contrast clamp fall where the reference's do
T2O_HD float lum_rn(float r, float g, float b) {
    return add_rn(add_rn(mul_rn(0.27f, r), mul_rn(0.67f, g)), mul_rn(0.06f, b));
}

// ---------------------------------------------------------------- per-(image, op) tables
// scalar ops : tab[0] = p, tab[1] = derived (1+p | 1-p | 2^p)
// whitebal.  : tab[0..2]
// tone       : one curve table at tab[0]
// color      : three curve tables at tab[0], tab[CT], tab[2*CT]
T2O_HD void build_curve(const float *k, int L, float *ct) {
    float S = 0.0f;
    for (int i = 0; i < L; ++i) S += k[i];
    S += CURVE_EPS;
    const float scale = (float)L / S;
    const float invL = 1.0f / (float)L;
    float prefix = 0.0f;
    for (int j = 0; j < MAX_L; ++j) {
        if (j < L) {
            const float kp = k[j] * scale;
            ct[j] = kp;
            ct[MAX_L + j] = prefix * scale - kp * ((float)j * invL);
            prefix += k[j] * invL;
        } else {
            ct[j] = 0.0f;
            ct[MAX_L + j] = 0.0f;
        }
    }
    ct[16] = 1.0f / S;
    ct[17] = scale;
    // The curve maps 1 -> sum(k)/(sum(k)+1e-10) < 1, but the rounded table can land one ulp above 1,
    // which would make the output clamp swallow the gradient of every saturated (x == 1.0) pixel.
    // Pull the last segment back so that y(1) <= 1 (a < 1e-6 shift; reference rounding there is
    // platform dependent anyway).
    float q_end = ct[MAX_L + L - 1];
    if (fmaf(ct[L - 1], 1.0f, q_end) < 1.0f + 1e-5f) {
        for (int it = 0; it < 8 && fmaf(ct[L - 1], 1.0f, q_end) > 1.0f; ++it) q_end = nextafterf(q_end, -4.0f);
        ct[MAX_L + L - 1] = q_end;
    }
}

T2O_HD void build_table(int op, const float *p, int L, float *tab) {
    switch (op) {
        case OP_BRIGHTNESS: case OP_SATURATION: tab[0] = p[0]; tab[1] = 1.0f + p[0]; break;
        case OP_CONTRAST: tab[0] = p[0]; tab[1] = 1.0f - p[0]; break;
        case OP_SHARPNESS: case OP_WHITE: tab[0] = p[0]; break;
        case OP_EXPOSURE: tab[0] = p[0]; tab[1] = expf(p[0] * LN2_F); break;
        case OP_WHITEBALANCE: tab[0] = p[0]; tab[1] = p[1]; tab[2] = p[2]; break;
        case OP_TONE: build_curve(p, L, tab); break;
        case OP_COLOR: for (int c = 0; c < 3; ++c) build_curve(p + c * L, L, tab + c * CT); break;
        default: break;
    }
}

// ---------------------------------------------------------------- blend + clamp (models/operators.py:129-130)
T2O_HD float blend(float y, float x, float m, bool has_mask) { return has_mask ? fmaf(y, m, x * (1.0f - m)) : y; }

// ---------------------------------------------------------------- forward: y = process(x; p)   (pre-blend)
T2O_HD void brightness_y(const float *tab, float r, float g, float b, float &yr, float &yg, float &yb) {
    const float v = max3(r, g, b);
    const float inv = rcp(v + HSV_EPS);
    const float v2 = sat01(v * tab[1]);
    yr = v2 * (1.0f - (v - r) * inv);
    yg = v2 * (1.0f - (v - g) * inv);
    yb = v2 * (1.0f - (v - b) * inv);
}

T2O_HD void saturation_y(const float *tab, float r, float g, float b, float &yr, float &yg, float &yb) {
    const float v = max3(r, g, b), mn = min3(r, g, b);
    const float d = v - mn;
    const float s = fdiv(d, v + HSV_EPS);
    const float s2 = sat01(s * tab[1]);
    const float rho = d > 0.0f ? fdiv(s2, d) : 0.0f;
    const float vr = v * rho;
    yr = v - (v - r) * vr;
    yg = v - (v - g) * vr;
    yb = v - (v - b) * vr;
}

T2O_HD void contrast_y(const float *tab, float r, float g, float b, float &yr, float &yg, float &yb) {
    const float L = sat01(lum_rn(r, g, b));
    const float cl = 0.5f - 0.5f * cospi_f(L);
    const float R = fdiv(cl, L + LUM_EPS);
    const float F = fmaf(tab[0], R, tab[1]);      // (1-p) + p*R
    yr = r * F; yg = g * F; yb = b * F;
}

T2O_HD int curve_bin(float xs, int L) {
    const int j = (int)(xs * (float)L);
    return j < L - 1 ? j : L - 1;
}
T2O_HD float curve_y(const float *ct, int L, float x) {
    const float xs = sat01(x);
    const int j = curve_bin(xs, L);
    return fmaf(ct[j], xs, ct[MAX_L + j]);
}

// y = x + p * laplace(x), zero padding handled by the caller (neighbours outside the image are 0)
T2O_HD float laplace(float c, float up, float dn, float lf, float rt) { return 4.0f * c - up - dn - lf - rt; }

// pointwise operators only (sharpness needs neighbours and is applied by the kernels)
T2O_HD void op_y(int op, const float *tab, int L, float r, float g, float b, float &yr, float &yg, float &yb) {
    switch (op) {
        case OP_BRIGHTNESS: brightness_y(tab, r, g, b, yr, yg, yb); break;
        case OP_CONTRAST: contrast_y(tab, r, g, b, yr, yg, yb); break;
        case OP_SATURATION: saturation_y(tab, r, g, b, yr, yg, yb); break;
        case OP_COLOR: yr = curve_y(tab, L, r); yg = curve_y(tab + CT, L, g); yb = curve_y(tab + 2 * CT, L, b); break;
        case OP_TONE: yr = curve_y(tab, L, r); yg = curve_y(tab, L, g); yb = curve_y(tab, L, b); break;
        case OP_WHITE: yr = 1.0f; yg = 1.0f; yb = 1.0f; break;
        case OP_EXPOSURE: yr = r * tab[1]; yg = g * tab[1]; yb = b * tab[1]; break;
        case OP_WHITEBALANCE: yr = r * tab[0]; yg = g * tab[1]; yb = b * tab[2]; break;
        default: yr = r; yg = g; yb = b; break;
    }
}

// one full Operator.execute on a pixel (pointwise operators): x <- clamp(blend(process(x)))
T2O_HD void op_apply(int op, const float *tab, int L, float &r, float &g, float &b,
                     float mr, float mg, float mb, bool has_mask, bool raw = false) {
    if (op < 0) return;                       // identity: no clamp (executors/executor.py:44-46)
    float yr, yg, yb;
    op_y(op, tab, L, r, g, b, yr, yg, yb);
    if (raw) { r = yr; g = yg; b = yb; return; }   // Operator.process only
    r = sat01(blend(yr, r, mr, has_mask));
    g = sat01(blend(yg, g, mg, has_mask));
    b = sat01(blend(yb, b, mb, has_mask));
}

// ---------------------------------------------------------------- backward
// Gradient through blend + clamp: g (dLoss/d out) -> gy (dLoss/d y) and gd (direct path to x).
T2O_HD void blend_bwd(float y, float x, float m, bool has_mask, float g, float &gy, float &gd) {
    const float z = blend(y, x, m, has_mask);
    const float gz = in01(z) ? g : 0.0f;
    gy = has_mask ? gz * m : gz;
    gd = has_mask ? gz * (1.0f - m) : 0.0f;
}

// Histogram sink for the curve moments.  `h[slot * stride]` is private to the calling thread.
struct Hist {
    float *h;
    int stride;
    T2O_HD void add(int slot, float v) const { h[slot * stride] += v; }
};

// Each *_bwd takes the operator input x = (r,g,b), the mask, the upstream gradient
// (gr,gg,gb) = dLoss/d(out) and returns dLoss/d(x) in place.  `acc` receives the parameter
// gradient contributions when `own` is true (halo pixels recompute but must not accumulate).
T2O_HD void brightness_bwd(const float *tab, float r, float g, float b, float mr, float mg, float mb, bool has_mask,
                           float &gr, float &gg, float &gb, float *acc, bool own) {
    const float q = tab[1];
    const float v = max3(r, g, b), mn = min3(r, g, b);
    const float inv = rcp(v + HSV_EPS);
    const float t = v * q;
    const float v2 = sat01(t);
    const float ipq = in01(t) ? q : 0.0f;
    const float wr = 1.0f - (v - r) * inv, wg = 1.0f - (v - g) * inv, wb = 1.0f - (v - b) * inv;
    float gyr, gyg, gyb, gdr, gdg, gdb;
    blend_bwd(v2 * wr, r, mr, has_mask, gr, gyr, gdr);
    blend_bwd(v2 * wg, g, mg, has_mask, gg, gyg, gdg);
    blend_bwd(v2 * wb, b, mb, has_mask, gb, gyb, gdb);
    const float G = gyr * wr + gyg * wg + gyb * wb;                   // dLoss/d v'
    if (own) acc[0] += in01(t) ? v * G : 0.0f;
    if (v == mn) {          // gray pixel: the reference routes everything through max -> channel 0
        gr = gdr + G * ipq; gg = gdg; gb = gdb;
        return;
    }
    const float k = v2 * inv;
    const float Gv = G * (ipq - k);
    const int im = argmax3(r, g, b);
    gr = gdr + gyr * k + (im == 0 ? Gv : 0.0f);
    gg = gdg + gyg * k + (im == 1 ? Gv : 0.0f);
    gb = gdb + gyb * k + (im == 2 ? Gv : 0.0f);
}

T2O_HD void saturation_bwd(const float *tab, float r, float g, float b, float mr, float mg, float mb, bool has_mask,
                           float &gr, float &gg, float &gb, float *acc, bool own) {
    const float q = tab[1];
    const float v = max3(r, g, b), mn = min3(r, g, b);
    const float d = v - mn;
    const float inv = rcp(v + HSV_EPS);
    const float s = fdiv(d, v + HSV_EPS);
    const float t = s * q;
    const float s2 = sat01(t);
    const float rho = d > 0.0f ? fdiv(s2, d) : 0.0f;
    const float ur = v - r, ug = v - g, ub = v - b;
    const float vr = v * rho;
    float gyr, gyg, gyb, gdr, gdg, gdb;
    blend_bwd(v - ur * vr, r, mr, has_mask, gr, gyr, gdr);
    blend_bwd(v - ug * vr, g, mg, has_mask, gg, gyg, gdg);
    blend_bwd(v - ub * vr, b, mb, has_mask, gb, gyb, gdb);
    const float Sg = gyr + gyg + gyb;
    if (!(d > 0.0f)) {      // gray pixel: y = v for every channel, max -> channel 0
        gr = gdr + Sg; gg = gdg; gb = gdb;
        return;
    }
    const float Su = gyr * ur + gyg * ug + gyb * ub;
    float rho_v, rho_mn;
    if (in01(t)) {          // s' = s q  ->  rho = q / (v + eps)
        if (own) acc[0] -= v * inv * Su;
        rho_v = -q * inv * inv; rho_mn = 0.0f;
    } else if (t > 1.0f) {  // s' = 1    ->  rho = 1 / d
        const float id = rcp(d);
        rho_v = -id * id; rho_mn = id * id;
    } else {                // s' = 0
        rho_v = 0.0f; rho_mn = 0.0f;
    }
    const float Gv = Sg - rho * Su - vr * Sg - v * rho_v * Su;
    const float Gmn = -v * rho_mn * Su;
    const int im = argmax3(r, g, b), in = argmin3(r, g, b);
    gr = gdr + gyr * vr + (im == 0 ? Gv : 0.0f) + (in == 0 ? Gmn : 0.0f);
    gg = gdg + gyg * vr + (im == 1 ? Gv : 0.0f) + (in == 1 ? Gmn : 0.0f);
    gb = gdb + gyb * vr + (im == 2 ? Gv : 0.0f) + (in == 2 ? Gmn : 0.0f);
}

T2O_HD void contrast_bwd(const float *tab, float r, float g, float b, float mr, float mg, float mb, bool has_mask,
                         float &gr, float &gg, float &gb, float *acc, bool own) {
    const float p = tab[0];
    const float lum = lum_rn(r, g, b);
    const float L = sat01(lum);
    // torch.min(torch.max(lum, 0), 1): binary max/min split the gradient 0.5/0.5 on ties
    const float f0 = lum > 0.0f ? 1.0f : (lum == 0.0f ? 0.5f : 0.0f);
    const float f1 = L < 1.0f ? 1.0f : (fmaxf(lum, 0.0f) == 1.0f ? 0.5f : 0.0f);
    const float cl = 0.5f - 0.5f * cospi_f(L);
    const float dcl = 0.5f * PI_F * sinpi_f(L);
    const float iden = rcp(L + LUM_EPS);
    const float R = cl * iden;
    const float dR = (dcl - R) * iden;
    const float F = fmaf(p, R, tab[1]);
    float gyr, gyg, gyb, gdr, gdg, gdb;
    blend_bwd(r * F, r, mr, has_mask, gr, gyr, gdr);
    blend_bwd(g * F, g, mg, has_mask, gg, gyg, gdg);
    blend_bwd(b * F, b, mb, has_mask, gb, gyb, gdb);
    const float Sgc = gyr * r + gyg * g + gyb * b;
    if (own) acc[0] += (R - 1.0f) * Sgc;
    const float k = p * dR * f0 * f1 * Sgc;
    gr = gdr + gyr * F + 0.27f * k;
    gg = gdg + gyg * F + 0.67f * k;
    gb = gdb + gyb * F + 0.06f * k;
}

// one channel of a curve operator
T2O_HD float curve_bwd(const float *ct, int L, float x, float m, bool has_mask, float g, const Hist &hist, bool own) {
    const float xs = sat01(x);
    const int j = curve_bin(xs, L);
    const float kp = ct[j];
    const float y = fmaf(kp, xs, ct[MAX_L + j]);
    float gy, gd;
    blend_bwd(y, x, m, has_mask, g, gy, gd);
    if (own) {
        hist.add(j, gy);
        hist.add(MAX_L + j, gy * xs);
        hist.add(2 * MAX_L, gy * y);
    }
    float slope = kp;
    if (j > 0 && xs * (float)L == (float)j) slope += ct[j - 1];   // exact knot: both clamp terms pass
    return gd + (in01(x) ? gy * slope : 0.0f);
}

// dLoss/dk_i of one curve from its block-reduced moments (A, Bx, C): k has L entries
T2O_HD void curve_param_grad(const float *ct, int L, const float *mom, float *gk) {
    const float invS = ct[16], scale = ct[17];
    float tail = 0.0f;                       // sum_{j > i} A_j
    for (int i = L - 1; i >= 0; --i) {
        const float x0 = (float)i / (float)L;
        const float sgc = (mom[MAX_L + i] - x0 * mom[i]) + tail / (float)L;
        gk[i] = scale * sgc - invS * mom[2 * MAX_L];
        tail += mom[i];
    }
}

T2O_HD void pointwise_bwd(int op, const float *tab, int L, float r, float g, float b,
                          float mr, float mg, float mb, bool has_mask,
                          float &gr, float &gg, float &gb, float *acc, const Hist &hist, bool own) {
    switch (op) {
        case OP_BRIGHTNESS: brightness_bwd(tab, r, g, b, mr, mg, mb, has_mask, gr, gg, gb, acc, own); break;
        case OP_CONTRAST: contrast_bwd(tab, r, g, b, mr, mg, mb, has_mask, gr, gg, gb, acc, own); break;
        case OP_SATURATION: saturation_bwd(tab, r, g, b, mr, mg, mb, has_mask, gr, gg, gb, acc, own); break;
        case OP_TONE:
            gr = curve_bwd(tab, L, r, mr, has_mask, gr, hist, own);
            gg = curve_bwd(tab, L, g, mg, has_mask, gg, hist, own);
            gb = curve_bwd(tab, L, b, mb, has_mask, gb, hist, own);
            break;
        case OP_COLOR: {
            Hist h1{hist.h + HIST * hist.stride, hist.stride}, h2{hist.h + 2 * HIST * hist.stride, hist.stride};
            gr = curve_bwd(tab, L, r, mr, has_mask, gr, hist, own);
            gg = curve_bwd(tab + CT, L, g, mg, has_mask, gg, h1, own);
            gb = curve_bwd(tab + 2 * CT, L, b, mb, has_mask, gb, h2, own);
            break;
        }
        case OP_WHITE: {
            float gy, gd;
            blend_bwd(1.0f, r, mr, has_mask, gr, gy, gd); gr = gd;
            blend_bwd(1.0f, g, mg, has_mask, gg, gy, gd); gg = gd;
            blend_bwd(1.0f, b, mb, has_mask, gb, gy, gd); gb = gd;
            break;
        }
        case OP_EXPOSURE: {
            const float e = tab[1];
            float gyr, gyg, gyb, gdr, gdg, gdb;
            blend_bwd(r * e, r, mr, has_mask, gr, gyr, gdr);
            blend_bwd(g * e, g, mg, has_mask, gg, gyg, gdg);
            blend_bwd(b * e, b, mb, has_mask, gb, gyb, gdb);
            if (own) acc[0] += LN2_F * e * (gyr * r + gyg * g + gyb * b);
            gr = gdr + gyr * e; gg = gdg + gyg * e; gb = gdb + gyb * e;
            break;
        }
        case OP_WHITEBALANCE: {
            float gyr, gyg, gyb, gdr, gdg, gdb;
            blend_bwd(r * tab[0], r, mr, has_mask, gr, gyr, gdr);
            blend_bwd(g * tab[1], g, mg, has_mask, gg, gyg, gdg);
            blend_bwd(b * tab[2], b, mb, has_mask, gb, gyb, gdb);
            if (own) { acc[0] += gyr * r; acc[1] += gyg * g; acc[2] += gyb * b; }
            gr = gdr + gyr * tab[0]; gg = gdg + gyg * tab[1]; gb = gdb + gyb * tab[2];
            break;
        }
        default: break;     // identity: gradient passes unchanged
    }
}

}  // namespace t2o
