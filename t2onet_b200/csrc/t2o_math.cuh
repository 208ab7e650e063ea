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
//   Contrast:         0.5 - 0.5 cos(pi L) = sin^2(pi L / 2), odd/even polynomials in L on [0, 1]
//   Curves:           j = floor(L x);  y = k'_j x + Q_j,  k' = k L/S, Q_j = L/S sum_{i<j} k_i/L - k'_j j/L
//                     (one 16-byte segment record per j: k'_j, Q_j and the extra slope an exact knot passes)
//   Curve param grads: G_i = sum g clamp(L x - i, 0, 1), C = sum g y  ->  dLoss/dk_i = (G_i - C) / S
//
// The instruction budget matters: at the HBM roofline a B200 SM has ~2 cycles per pixel, so the
// hot paths avoid IEEE division (MUFU.RCP, 1 ulp), libm trigonometry and float->int conversions.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define T2O_HD __host__ __device__ __forceinline__
#else
#define T2O_HD inline
#endif

namespace t2o {

enum : int {
    OP_SKIP = -2, OP_IDENTITY = -1, OP_BRIGHTNESS = 0, OP_CONTRAST = 1, OP_SATURATION = 2, OP_COLOR = 3, OP_INPAINT = 4,
    OP_TONE = 5, OP_SHARPNESS = 6, OP_WHITE = 7, OP_EXPOSURE = 8, OP_WHITEBALANCE = 9,
    OP_BNW = 10, OP_BLUR = 11, OP_HUE = 12, OP_COUNT = 13
};

constexpr int MAX_CHAIN = 8;
constexpr int MAX_L = 8;    // even (the gradient accumulators are paired)
constexpr int TAB = 128;   // floats in one (image, op) table
constexpr int CT = 40;     // floats in one curve table: 9 segments x (k', Q, knot slope, 0), then 1/S, L/S
constexpr int CT_INVS = 36, CT_SCALE = 37, CT_INRANGE = 38;   // CT_INRANGE: 1.0f if the curve maps [0, 1] into [0, 1] (see build_curve)
constexpr int NBIN = MAX_L + 1;   // segments of one curve table (entry L repeats L-1: it serves x == 1.0)
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
// operators that read the pixel's 3x3 neighbourhood: y = x + p * S(x) with a symmetric stencil S
T2O_HD bool op_is_stencil(int op) { return op == OP_SHARPNESS || op == OP_BLUR; }
// BlurOperator (models/operators.py:373-411): lerp(img, G * img, p) = img + p * (G - delta) * img with G the 3x3 Gaussian of
// get_gaussian_kernel(3, 2) (:685-717), zero padding 1; the fp32 weights below are the reference's own (centre, edge, corner)
constexpr float BLUR_WC = 0.13080118596553802f, BLUR_WE = 0.11543164402246475f, BLUR_WK = 0.10186807066202164f;
constexpr float TWO_PI_F = 6.283185307179586f;

struct F2 { float a, b; };   // == float2 without needing vector_types.h on the host
struct F4 { float a, b, c, d; };   // == float4

// ---------------------------------------------------------------- scalar helpers
T2O_HD float sat01(float z) {
#if defined(__CUDA_ARCH__)
    return __saturatef(z);
#else
    return fminf(fmaxf(z, 0.0f), 1.0f);
#endif
}
// torch.clamp backward passes the gradient on the closed interval [0, 1].  On the device this is ONE
// unsigned compare of the bit pattern (non-negative floats order like unsigned ints; negatives and NaN
// have larger patterns).  The only deviation is z == -0.0f, which counts as outside.
T2O_HD bool in01(float z) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(z) <= 0x3F800000u;
#else
    return z >= 0.0f && z <= 1.0f;
#endif
}
// 1/x as a bare MUFU.RCP (<= 1 ulp) instead of the ~10-instruction IEEE sequence
T2O_HD float rcp(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#else
    return 1.0f / x;
#endif
}
T2O_HD float fdiv(float a, float b) { return a * rcp(b); }
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
// sin(pi/2 L) and cos(pi/2 L) for L in [0, 1]: Chebyshev-fitted polynomials in L^2, |err| < 2e-7 in fp32
T2O_HD float sin_halfpi(float L) {
    const float t = L * L;
    float p = 0.0001516751217423007f;
    p = fmaf(p, t, -0.004674150608479977f);
    p = fmaf(p, t, 0.07968991994857788f);
    p = fmaf(p, t, -0.6459637880325317f);
    p = fmaf(p, t, 1.5707963705062866f);
    return p * L;
}
T2O_HD float cos_halfpi(float L) {
    const float t = L * L;
    float p = -2.3824535674066283e-05f;
    p = fmaf(p, t, 0.0009177238680422306f);
    p = fmaf(p, t, -0.02086268737912178f);
    p = fmaf(p, t, 0.2536693215370178f);
    p = fmaf(p, t, -1.2337005138397217f);
    p = fmaf(p, t, 1.0f);
    return p;
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
// curve table: segment j = 0..L at ct[4j]: (k'_j, Q_j, K_j, k'_j) with K_j = k'_j + k'_{j-1} for 0 < j < L, else k'_j
//              (the forward loads the first half of a record, the gate-free backward the second, the gated backward all of it):
//              at an exact knot x = j/L both neighbouring clamp terms of the reference pass the gradient (closed
//              intervals), so dy/dx = K_j there.  Segment L repeats L-1 and serves x == 1.0.
//              ct[CT_INVS] = 1/S, ct[CT_SCALE] = L/S
// BWD: also set CT_INRANGE (only the backward kernels read it; the forward kernels and the scorer, which builds a table per
// candidate and round, skip the 2 L evaluations)
// One record of a curve table and the scalars every record needs.  Every product and sum is an explicitly rounded operation
// (mul_rn / fmaf), so that the serial builder below and the lane-parallel one of the resident planner kernel (one lane per
// record, t2o_score.cu) produce the same bits whatever the compiler would contract.
struct CurveScalars { float S, scale, invL; };
// (fixed trip counts with a predicate: the loops unroll, the loads of k[] issue together)
T2O_HD CurveScalars curve_scalars(const float *k, int L) {
    float S = 0.0f;
#pragma unroll
    for (int i = 0; i < MAX_L; ++i)
        if (i < L) S += k[i];
    S += CURVE_EPS;
    return CurveScalars{S, (float)L / S, 1.0f / (float)L};
}
// prefix_j = sum_{i<j} k_i / L, accumulated in index order
T2O_HD float curve_prefix(const float *k, int j, float invL) {
    float prefix = 0.0f;
#pragma unroll
    for (int i = 0; i < MAX_L; ++i)
        if (i < j) prefix = fmaf(k[i], invL, prefix);
    return prefix;
}
// record j < L from prefix_j: (k'_j, Q_j, K_j, k'_j)
T2O_HD F4 curve_record(const float *k, int L, int j, const CurveScalars &cs, float prefix) {
    const float kp = mul_rn(k[j], cs.scale);
    const float q = fmaf(-kp, mul_rn((float)j, cs.invL), mul_rn(prefix, cs.scale));
    const float K = (j > 0) ? fmaf(k[j - 1], cs.scale, kp) : kp;
    return F4{kp, q, K, kp};
}
// The curve maps 1 -> sum(k)/(sum(k)+1e-10) < 1, but the rounded table can land one ulp above 1, which would make the
// output clamp swallow the gradient of every saturated (x == 1.0) pixel.  Pull the last segment back so that y(1) <= 1 (a
// < 1e-6 shift; reference rounding there is platform dependent anyway).
T2O_HD float curve_pull_back(float kp_last, float q_last) {
    float q_end = q_last;
    if (fmaf(kp_last, 1.0f, q_end) < 1.0f + 1e-5f)
        for (int it = 0; it < 8; ++it) {
            // one ulp of y(1) can be many ulps of a small Q: step by the excess, and by at least one ulp of Q
            const float y1 = fmaf(kp_last, 1.0f, q_end);
            if (!(y1 > 1.0f)) break;
            q_end = fminf(nextafterf(q_end, -4.0f), q_end - (y1 - 1.0f));
        }
    return q_end;
}

// BWD: also set CT_INRANGE (only the backward kernels read it; the forward kernels and the scorer, which builds a table per
// candidate and round, skip the 2 L evaluations)
template <bool BWD = true>
T2O_HD void build_curve(const float *k, int L, float *ct) {
    const CurveScalars cs = curve_scalars(k, L);
    const float invL = cs.invL;
    float prefix = 0.0f;
    for (int j = 0; j < NBIN; ++j) {
        F4 r = F4{0.0f, 0.0f, 0.0f, 0.0f};
        if (j < L) {
            r = curve_record(k, L, j, cs, prefix);
            prefix = fmaf(k[j], invL, prefix);
        }
        ct[4 * j] = r.a; ct[4 * j + 1] = r.b; ct[4 * j + 2] = r.c; ct[4 * j + 3] = r.d;
    }
    ct[4 * (L - 1) + 1] = curve_pull_back(ct[4 * (L - 1)], ct[4 * (L - 1) + 1]);
    ct[4 * L] = ct[4 * (L - 1)];
    ct[4 * L + 1] = ct[4 * (L - 1) + 1];
    ct[4 * L + 2] = ct[4 * (L - 1)];
    ct[4 * L + 3] = ct[4 * (L - 1)];
    ct[CT_INVS] = 1.0f / cs.S;
    ct[CT_SCALE] = cs.scale;
    // Does the output clamp ever cut this curve?  y is linear on a segment and fmaf is monotone in x, so it is enough to
    // evaluate both ends of every segment the way the kernels do.  All k_i >= 0 (every curve the Actor's regressors
    // produce) gives 1: the backward then skips the clamp gate of this curve.
    if (!BWD) return;
    bool inr = true;
    for (int j = 0; j < L; ++j) {
        const float y0 = fmaf(ct[4 * j], (float)j * invL, ct[4 * j + 1]), y1 = fmaf(ct[4 * j], (float)(j + 1) * invL, ct[4 * j + 1]);
        inr = inr && y0 >= 0.0f && y0 <= 1.0f && y1 >= 0.0f && y1 <= 1.0f;
    }
    ct[CT_INRANGE] = inr ? 1.0f : 0.0f;
}

template <bool BWD = true>
T2O_HD void build_table(int op, const float *p, int L, float *tab) {
    switch (op) {
        case OP_BRIGHTNESS: case OP_SATURATION: tab[0] = p[0]; tab[1] = 1.0f + p[0]; break;
        case OP_CONTRAST: tab[0] = p[0]; tab[1] = 1.0f - p[0]; break;
        case OP_SHARPNESS: case OP_WHITE: tab[0] = p[0]; break;
        case OP_BLUR: tab[0] = p[0]; break;
        case OP_BNW: tab[0] = p[0]; tab[1] = 1.0f - p[0]; break;
        case OP_HUE: {
            // hsv_to_rgb with a constant hue (models/operators.py:432-438): the sector hi and the fraction f are
            // per-image constants; channel c is v * (1 - a_c * s) with a_c in {0, 1, f, 1 - f} by sector
            const float h6 = (p[0] / TWO_PI_F) * 6.0f;
            float m6 = fmodf(h6, 6.0f);
            if (m6 < 0.0f) m6 += 6.0f;                        // torch's % takes the sign of the divisor
            float hi = fmodf(floorf(h6), 6.0f);
            if (hi < 0.0f) hi += 6.0f;
            const float f = m6 - hi, df = 6.0f / TWO_PI_F;
            const int sec = (int)hi;
            // rows of the gather table (v, q, p, p, t, v | t, v, v, q, p, p | p, p, t, v, v, q): v -> 0, p -> 1, q -> f, t -> 1 - f
            // (two bits per sector, packed: no local array, so the kernels keep no stack frame for it)
            const unsigned int code[3] = {856u, 1411u, 2101u};     // {0,2,1,1,3,0}, {3,0,0,2,1,1}, {1,1,3,0,0,2}
            const int sc = sec < 0 ? 0 : (sec > 5 ? 5 : sec);
            tab[0] = p[0];
            for (int c = 0; c < 3; ++c) {
                const int k = (int)((code[c] >> (2 * sc)) & 3u);
                tab[1 + c] = k == 0 ? 0.0f : (k == 1 ? 1.0f : (k == 2 ? f : 1.0f - f));
                tab[4 + c] = k == 2 ? df : (k == 3 ? -df : 0.0f);
            }
            break;
        }
        case OP_EXPOSURE: tab[0] = p[0]; tab[1] = expf(p[0] * LN2_F); break;
        case OP_WHITEBALANCE: tab[0] = p[0]; tab[1] = p[1]; tab[2] = p[2]; break;
        case OP_TONE: build_curve<BWD>(p, L, tab); break;
        case OP_COLOR: for (int c = 0; c < 3; ++c) build_curve<BWD>(p + c * L, L, tab + c * CT); break;
        default: break;
    }
}

// The same split in three: part c builds the c-th curve of a color operator, part 0 everything else -- so that three
// threads per operator share the one table whose construction is long (3 x 9 segments, each with a division).
template <bool BWD = true>
T2O_HD void build_table_part(int op, int part, const float *p, int L, float *tab) {
    if (op == OP_COLOR) build_curve<BWD>(p + part * L, L, tab + part * CT);
    else if (part == 0) build_table<BWD>(op, p, L, tab);
}

#if defined(__CUDACC__)
// One warp builds one operator's table: one lane per curve record (3 L lanes of a color operator, L of a tone operator),
// lane 0 everything else -- the same functions in the same order as build_curve, so the same bits; the in-range flag is the
// AND of the lanes' segment checks (a ballot).  Called by all 32 lanes of the warp.  The kernels give every operator of a
// chain its own warp: the serial builders (three threads of one warp, one curve each, different operators one after the
// other) kept a 128 x 128 step's other 250 threads at the first barrier for ~2 us.
template <bool BWD = true>
__device__ __forceinline__ void build_table_lanes_l(int op, int lane, const float *p, const int L, float *tab);
// (curve_steps is 8 everywhere in the reference: with the constant the lane split is a shift, 1 / L and the loop predicates
// fold -- the same operations on the same values)
template <bool BWD = true>
__device__ __forceinline__ void build_table_lanes(int op, int lane, const float *p, int L, float *tab) {
    if (L == MAX_L) build_table_lanes_l<BWD>(op, lane, p, MAX_L, tab);
    else build_table_lanes_l<BWD>(op, lane, p, L, tab);
}
template <bool BWD>
__device__ __forceinline__ void build_table_lanes_l(int op, int lane, const float *p, const int L, float *tab) {
    if (op == OP_COLOR || op == OP_TONE) {
        const int part = lane / L, j = lane - part * L;
        const bool active = part < (op == OP_COLOR ? 3 : 1);
        bool ok = true;
        if (active) {
            const float *k = p + part * L;
            float *ct = tab + part * CT;
            const CurveScalars cs = curve_scalars(k, L);
            F4 r = curve_record(k, L, j, cs, curve_prefix(k, j, cs.invL));
            if (j == L - 1) {
                r.b = curve_pull_back(r.a, r.b);
                ct[4 * L] = r.a; ct[4 * L + 1] = r.b; ct[4 * L + 2] = r.a; ct[4 * L + 3] = r.a;
                ct[CT_INVS] = 1.0f / cs.S;
                ct[CT_SCALE] = cs.scale;
            }
            ct[4 * j] = r.a; ct[4 * j + 1] = r.b; ct[4 * j + 2] = r.c; ct[4 * j + 3] = r.d;
            if (BWD) {
                const float y0 = fmaf(r.a, (float)j * cs.invL, r.b), y1 = fmaf(r.a, (float)(j + 1) * cs.invL, r.b);
                ok = y0 >= 0.0f && y0 <= 1.0f && y1 >= 0.0f && y1 <= 1.0f;
            }
        }
        if (BWD) {
            const unsigned int bad = __ballot_sync(0xffffffffu, !ok);
            const unsigned int mine = ((1u << L) - 1u) << (part * L);           // the lanes of this lane's curve
            if (active && j == 0) tab[part * CT + CT_INRANGE] = (bad & mine) ? 0.0f : 1.0f;
        }
    } else if (lane == 0) {
        build_table<BWD>(op, p, L, tab);
    }
}
#endif

// ---------------------------------------------------------------- blend + clamp (models/operators.py:129-130)
template <bool HM>
T2O_HD float blend(float y, float x, float m) { return HM ? fmaf(y, m, x * (1.0f - m)) : y; }

// ---------------------------------------------------------------- forward: y = process(x; p)   (pre-blend)
T2O_HD void brightness_y(const float *tab, float r, float g, float b, float &yr, float &yg, float &yb) {
    const float v = max3(r, g, b);
    const float inv = rcp(v + HSV_EPS);
    const float v2 = sat01(v * tab[1]);
    yr = v2 * fmaf(r - v, inv, 1.0f);
    yg = v2 * fmaf(g - v, inv, 1.0f);
    yb = v2 * fmaf(b - v, inv, 1.0f);
}

T2O_HD void saturation_y(const float *tab, float r, float g, float b, float &yr, float &yg, float &yb) {
    const float v = max3(r, g, b), mn = min3(r, g, b);
    const float d = v - mn;
    const float s2 = sat01(d * rcp(v + HSV_EPS) * tab[1]);
    const float vr = v * (s2 * rcp(fmaxf(d, 1e-30f)));      // d == 0  =>  s2 == 0  =>  vr == 0
    yr = fmaf(r - v, vr, v);
    yg = fmaf(g - v, vr, v);
    yb = fmaf(b - v, vr, v);
}

T2O_HD void contrast_y(const float *tab, float r, float g, float b, float &yr, float &yg, float &yb) {
    const float L = sat01(lum_rn(r, g, b));
    const float sh = sin_halfpi(L);
    const float R = sh * sh * rcp(L + LUM_EPS);     // (0.5 - 0.5 cos(pi L)) / (L + eps)
    const float F = fmaf(tab[0], R, tab[1]);        // (1-p) + p*R
    yr = r * F; yg = g * F; yb = b * F;
}

// BNWOperator.process (models/operators.py:314-316): lerp(img, rgb2lum(img), p)
T2O_HD void bnw_y(const float *tab, float r, float g, float b, float &yr, float &yg, float &yb) {
    const float pl = tab[0] * lum_rn(r, g, b);
    yr = fmaf(tab[1], r, pl); yg = fmaf(tab[1], g, pl); yb = fmaf(tab[1], b, pl);
}

// HueOperator.process (models/operators.py:432-438): hsv_to_rgb(param, s, v) -- see build_table
T2O_HD void hue_y(const float *tab, float r, float g, float b, float &yr, float &yg, float &yb) {
    const float v = max3(r, g, b), mn = min3(r, g, b);
    const float vs = v * ((v - mn) * rcp(v + HSV_EPS));
    yr = fmaf(-tab[1], vs, v); yg = fmaf(-tab[2], vs, v); yb = fmaf(-tab[3], vs, v);
}

// Segment lookup of a clamped input xs in [0, 1]: bin j = floor(L xs) in 0..L, record j of the table.
// On the device the bin comes from a round-down add against 2.0f: with tt = xs * (L * 2^-22) the sum 2 + tt rounds down to
// 2 + j * 2^-22 (the ulp of [2, 4) is 2^-22), whose bit pattern is 0x40000000 + j.  Shifted left by 4 the exponent drops
// out and 16 j is left: the record's SHARED-MEMORY address is ONE shift-add on top of the table's base, no mask (xs is
// clamped, so j <= L always), no float->int conversion.  Every kernel keeps its operator tables in shared memory
// (StepShared::tabs, ChainShared::tabs, the scorer's wtab).  The scaling by 2^-22 is exact, so tt == u - 2 exactly when
// L xs is an integer (the knot test of the backward), and fma(tt, 2^22, -i) is the correctly rounded L xs - i.
constexpr float CURVE_TT = 1.0f / 4194304.0f, CURVE_TT_INV = 4194304.0f;   // 2^-22, 2^22
// -> tt = L xs 2^-22, tfs = j 2^-22; seg = (k'_j, Q_j[, K_j, 0])
T2O_HD F2 curve_seg2(const float *ct, float xs, int L, float &tt, float &tfs) {
    tt = xs * ((float)L * CURVE_TT);
    F2 seg;
#if defined(__CUDA_ARCH__)
    const float u = __fadd_rd(tt, 2.0f);
    tfs = u - 2.0f;
    asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(seg.a), "=f"(seg.b)
        : "r"((__float_as_uint(u) << 4) + (unsigned)__cvta_generic_to_shared(ct)));
#else
    const float tf = floorf(tt * CURVE_TT_INV);
    tfs = tf * CURVE_TT;
    seg.a = ct[4 * (int)tf]; seg.b = ct[4 * (int)tf + 1];
#endif
    return seg;
}
// the second half of the record: (K_j, k'_j)
T2O_HD F2 curve_seg2hi(const float *ct, float xs, int L, float &tt, float &tfs) {
    tt = xs * ((float)L * CURVE_TT);
    F2 seg;
#if defined(__CUDA_ARCH__)
    const float u = __fadd_rd(tt, 2.0f);
    tfs = u - 2.0f;
    asm("ld.shared.v2.f32 {%0, %1}, [%2+8];" : "=f"(seg.a), "=f"(seg.b)
        : "r"((__float_as_uint(u) << 4) + (unsigned)__cvta_generic_to_shared(ct)));
#else
    const float tf = floorf(tt * CURVE_TT_INV);
    tfs = tf * CURVE_TT;
    seg.a = ct[4 * (int)tf + 2]; seg.b = ct[4 * (int)tf + 3];
#endif
    return seg;
}
T2O_HD F4 curve_seg4(const float *ct, float xs, int L, float &tt, float &tfs) {
    tt = xs * ((float)L * CURVE_TT);
    F4 seg;
#if defined(__CUDA_ARCH__)
    const float u = __fadd_rd(tt, 2.0f);
    tfs = u - 2.0f;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(seg.a), "=f"(seg.b), "=f"(seg.c), "=f"(seg.d)
        : "r"((__float_as_uint(u) << 4) + (unsigned)__cvta_generic_to_shared(ct)));
#else
    const float tf = floorf(tt * CURVE_TT_INV);
    tfs = tf * CURVE_TT;
    const float *q = ct + 4 * (int)tf;
    seg.a = q[0]; seg.b = q[1]; seg.c = q[2]; seg.d = q[3];
#endif
    return seg;
}
// CL: the input is known to lie in [0, 1] (it is the clamped output of the previous operator)
template <bool CL>
T2O_HD float curve_y(const float *ct, int L, float x) {
    const float xs = CL ? x : sat01(x);
    float tt, tfs;
    const F2 seg = curve_seg2(ct, xs, L, tt, tfs);
    return fmaf(seg.a, xs, seg.b);
}

// y = x + p * laplace(x), zero padding handled by the caller (neighbours outside the image are 0)
T2O_HD float laplace(float c, float up, float dn, float lf, float rt) { return fmaf(4.0f, c, -((up + dn) + (lf + rt))); }
// y = x + p * blur_delta(x): (G - delta) * x of the 3x3 Gaussian; `edge` = up + dn + lf + rt, `corner` = the four diagonals
T2O_HD float blur_delta(float c, float edge, float corner) { return fmaf(BLUR_WC - 1.0f, c, fmaf(BLUR_WE, edge, BLUR_WK * corner)); }

// pointwise operators only (sharpness needs neighbours and is applied by the kernels)
template <bool CL = false>
T2O_HD void op_y(int op, const float *tab, int L, float r, float g, float b, float &yr, float &yg, float &yb) {
    switch (op) {
        case OP_BRIGHTNESS: brightness_y(tab, r, g, b, yr, yg, yb); break;
        case OP_CONTRAST: contrast_y(tab, r, g, b, yr, yg, yb); break;
        case OP_SATURATION: saturation_y(tab, r, g, b, yr, yg, yb); break;
        case OP_COLOR: yr = curve_y<CL>(tab, L, r); yg = curve_y<CL>(tab + CT, L, g); yb = curve_y<CL>(tab + 2 * CT, L, b); break;
        case OP_TONE: yr = curve_y<CL>(tab, L, r); yg = curve_y<CL>(tab, L, g); yb = curve_y<CL>(tab, L, b); break;
        case OP_WHITE: yr = 1.0f; yg = 1.0f; yb = 1.0f; break;
        case OP_EXPOSURE: yr = r * tab[1]; yg = g * tab[1]; yb = b * tab[1]; break;
        case OP_WHITEBALANCE: yr = r * tab[0]; yg = g * tab[1]; yb = b * tab[2]; break;
        case OP_BNW: bnw_y(tab, r, g, b, yr, yg, yb); break;
        case OP_HUE: hue_y(tab, r, g, b, yr, yg, yb); break;
        default: yr = r; yg = g; yb = b; break;
    }
}

// one full Operator.execute on a pixel (pointwise operators): x <- clamp(blend(process(x)))
template <bool HM, bool CL = false>
T2O_HD void op_apply(int op, const float *tab, int L, float &r, float &g, float &b,
                     float mr, float mg, float mb, bool raw = false) {
    if (op < 0) return;                       // identity: no clamp (executors/executor.py:44-46)
    float yr, yg, yb;
    op_y<CL>(op, tab, L, r, g, b, yr, yg, yb);
    if (raw) { r = yr; g = yg; b = yb; return; }   // Operator.process only
    r = sat01(blend<HM>(yr, r, mr));
    g = sat01(blend<HM>(yg, g, mg));
    b = sat01(blend<HM>(yb, b, mb));
}

// ---------------------------------------------------------------- backward
// Gradient through blend + clamp: g (dLoss/d out) -> gy (dLoss/d y) and gd (direct path to x).
template <bool HM>
T2O_HD void blend_bwd(float y, float x, float m, float g, float &gy, float &gd) {
    const float z = blend<HM>(y, x, m);
    const float gz = in01(z) ? g : 0.0f;
    gy = HM ? gz * m : gz;
    gd = HM ? gz * (1.0f - m) : 0.0f;
}

// Per-thread parameter-gradient accumulators: one slot per operator TYPE (a backward launch holds each
// operator type at most once, the binding splits longer chains), statically indexed so they live in registers.
// Curves: G[i] = sum g * clamp(L x - i, 0, 1); with y = sum_i k_i clamp(L x - i, 0, 1) / S and S = sum k + eps
// (models/operators.py:579-585, 610-616):  C = sum g * y = sum_i k_i G[i] / S  and  dLoss/dk_i = (G[i] - C) / S.
// The curve slots are kept as pairs so that the device accumulates two bins per packed FFMA2 (sm_100 fp32x2).
struct GradAcc {
    F2 color[3][MAX_L / 2];
    float bright, contrast, satur, expo, sharp;     // `sharp` serves the launch's one stencil operator (sharpness or blur)
    F2 tone[MAX_L / 2];
    float wb[3];
    float bnw, hue;
};
// c + a * b on both halves: one FFMA2 issue slot on the device
T2O_HD F2 fma2(F2 a, F2 b, F2 c) {
#if defined(__CUDA_ARCH__)
    const float2 r = __ffma2_rn(make_float2(a.a, a.b), make_float2(b.a, b.b), make_float2(c.a, c.b));
    return F2{r.x, r.y};
#else
    return F2{fmaf(a.a, b.a, c.a), fmaf(a.b, b.b, c.b)};
#endif
}
constexpr int ACC_SLOTS = 48;     // slot layout used by the kernels' reduction (see acc_slot_* below)
constexpr int ACC_COLOR = 0, ACC_BRIGHT = 27, ACC_CONTRAST = 28, ACC_SATUR = 29, ACC_EXPO = 30,
              ACC_SHARP = 31, ACC_TONE = 32, ACC_WB = 41, ACC_BNW = 24, ACC_HUE = 25;
T2O_HD void acc_zero(GradAcc &a) {
    for (int c = 0; c < 3; ++c) { for (int i = 0; i < MAX_L / 2; ++i) a.color[c][i] = F2{0.0f, 0.0f}; a.wb[c] = 0.0f; }
    for (int i = 0; i < MAX_L / 2; ++i) a.tone[i] = F2{0.0f, 0.0f};
    a.bright = 0.0f; a.contrast = 0.0f; a.satur = 0.0f; a.expo = 0.0f; a.sharp = 0.0f; a.bnw = 0.0f; a.hue = 0.0f;
}
T2O_HD void acc_to_slots(const GradAcc &a, float *v) {     // v[ACC_SLOTS]
    for (int i = 0; i < ACC_SLOTS; ++i) v[i] = 0.0f;
    for (int c = 0; c < 3; ++c) {
        for (int i = 0; i < MAX_L / 2; ++i) { v[ACC_COLOR + c * MAX_L + 2 * i] = a.color[c][i].a; v[ACC_COLOR + c * MAX_L + 2 * i + 1] = a.color[c][i].b; }
        v[ACC_WB + c] = a.wb[c];
    }
    for (int i = 0; i < MAX_L / 2; ++i) { v[ACC_TONE + 2 * i] = a.tone[i].a; v[ACC_TONE + 2 * i + 1] = a.tone[i].b; }
    v[ACC_BRIGHT] = a.bright; v[ACC_CONTRAST] = a.contrast; v[ACC_SATUR] = a.satur;
    v[ACC_EXPO] = a.expo; v[ACC_SHARP] = a.sharp; v[ACC_BNW] = a.bnw; v[ACC_HUE] = a.hue;
}

// Each *_bwd takes the operator input x = (r,g,b), the mask, the upstream gradient
// (gr,gg,gb) = dLoss/d(out) and returns dLoss/d(x) in place.  The parameter-gradient contribution goes
// to `acc` when `own` is true (halo pixels recompute but must not accumulate).
// Routing of a gradient that arrives at max(r, g, b) / min(r, g, b): torch's max(dim) / min(dim) send it to the FIRST
// channel attaining the value (v is one of r, g, b, so equality tests find it).
T2O_HD void route3(float r, float g, float b, float v, float G, float &gr, float &gg, float &gb) {
    const bool e0 = r == v, e1 = !e0 && g == v, e2 = !e0 && !e1;     // three independent predicated adds, no branch
    if (e0) gr += G;
    if (e1) gg += G;
    if (e2) gb += G;
}

// CL: the input is the clamped output of another operator (in [0, 1]).  The outputs of brightness and saturation then lie
// in [0, 1] by construction (y_c = v' w_c with v', w_c in [0, 1]; y_c = v - u_c v s'/d in [0, v]), so the output clamp
// always passes the gradient and its gate (recompute y_c, compare, select) is dropped.
template <bool HM, bool CL = false>
T2O_HD void brightness_bwd(const float *tab, float r, float g, float b, float mr, float mg, float mb,
                           float &gr, float &gg, float &gb, float &acc, bool own) {
    const float q = tab[1];
    const float v = max3(r, g, b), mn = min3(r, g, b);
    const float inv = rcp(v + HSV_EPS);
    const float t = v * q;
    const float v2 = sat01(t);
    const bool ip = in01(t);
    const float ipq = ip ? q : 0.0f;
    const float wr = fmaf(r - v, inv, 1.0f), wg = fmaf(g - v, inv, 1.0f), wb = fmaf(b - v, inv, 1.0f);
    float gyr, gyg, gyb, gdr, gdg, gdb;
    if (HM || !CL) {
        blend_bwd<HM>(v2 * wr, r, mr, gr, gyr, gdr);
        blend_bwd<HM>(v2 * wg, g, mg, gg, gyg, gdg);
        blend_bwd<HM>(v2 * wb, b, mb, gb, gyb, gdb);
    } else {
        gyr = gr; gyg = gg; gyb = gb; gdr = 0.0f; gdg = 0.0f; gdb = 0.0f;
    }
    const float G = gyr * wr + gyg * wg + gyb * wb;                   // dLoss/d v'
    if (own) acc = fmaf(ip ? v : 0.0f, G, acc);
    // y_c = v' (1 + (c - v) inv): the direct path has slope k = v' inv, the path through v = max the slope
    // G (dv'/dv - k).  A gray pixel (v == mn) has no direct path in the reference (hsv_to_rgb with s = 0 returns v for
    // every channel): everything goes through max -> channel 0, which is k = 0 here.
    const float k = v == mn ? 0.0f : v2 * inv;
    gr = fmaf(gyr, k, gdr); gg = fmaf(gyg, k, gdg); gb = fmaf(gyb, k, gdb);
    route3(r, g, b, v, G * (ipq - k), gr, gg, gb);
}

template <bool HM, bool CL = false>
T2O_HD void saturation_bwd(const float *tab, float r, float g, float b, float mr, float mg, float mb,
                           float &gr, float &gg, float &gb, float &acc, bool own) {
    const float q = tab[1];
    const float v = max3(r, g, b), mn = min3(r, g, b);
    const float d = v - mn;
    const float inv = rcp(v + HSV_EPS);
    const float t = d * inv * q;
    const float s2 = sat01(t);
    const float id = rcp(fmaxf(d, 1e-30f));
    const float rho = s2 * id;                                         // s' / d  (d == 0  =>  s' == 0  =>  rho == 0)
    const float ur = v - r, ug = v - g, ub = v - b;
    const float vr = v * rho;
    float gyr, gyg, gyb, gdr, gdg, gdb;
    if (HM || !CL) {
        blend_bwd<HM>(fmaf(-ur, vr, v), r, mr, gr, gyr, gdr);
        blend_bwd<HM>(fmaf(-ug, vr, v), g, mg, gg, gyg, gdg);
        blend_bwd<HM>(fmaf(-ub, vr, v), b, mb, gb, gyb, gdb);
    } else {
        gyr = gr; gyg = gg; gyb = gb; gdr = 0.0f; gdg = 0.0f; gdb = 0.0f;
    }
    const float Sg = gyr + gyg + gyb;
    const float Su = gyr * ur + gyg * ug + gyb * ub;
    // y_c = v - u_c v rho with rho = s'/d in three regimes of s' = clamp(s q): linear (rho = q/(v+eps)), saturated at 1
    // (rho = 1/d), cut at 0 (rho = 0).  In all of them d rho/dv = -rho e and d rho/d mn = rho eB with
    // e = 1/(v+eps) | 1/d | 0 and eB = 0 | 1/d | 0.  A gray pixel (d == 0) is the linear regime with rho = 0: all that is
    // left is Sg through max -> channel 0, as the reference routes it.
    const bool lin = in01(t);
    const float eB = t > 1.0f ? id : 0.0f;
    const float e = lin ? inv : eB;
    if (own) acc = fmaf(lin ? -(v * inv) : 0.0f, Su, acc);
    const float A1 = fmaf(-(v * e), rho, rho);                          // rho + v d rho/dv
    const float Gv = fmaf(-Su, A1, fmaf(-vr, Sg, Sg));
    const float Gmn = -(vr * eB) * Su;
    gr = fmaf(gyr, vr, gdr); gg = fmaf(gyg, vr, gdg); gb = fmaf(gyb, vr, gdb);
    route3(r, g, b, v, Gv, gr, gg, gb);
    route3(r, g, b, mn, Gmn, gr, gg, gb);
}

template <bool HM>
T2O_HD void contrast_bwd(const float *tab, float r, float g, float b, float mr, float mg, float mb,
                         float &gr, float &gg, float &gb, float &acc, bool own) {
    const float p = tab[0];
    const float lum = lum_rn(r, g, b);
    const float L = sat01(lum);
    // torch.min(torch.max(lum, 0), 1) splits the gradient 0.5 / 0.5 on ties.  Only the upper one can matter: for
    // lum <= 0 the factor it multiplies, dR below, is exactly 0 (L = 0 gives sin = 0).
    const float pf = lum < 1.0f ? p : (lum == 1.0f ? 0.5f * p : 0.0f);
    const float sh = sin_halfpi(L), ch = cos_halfpi(L);
    const float cl = sh * sh;                       // 0.5 - 0.5 cos(pi L)
    const float dcl = PI_F * sh * ch;               // 0.5 pi sin(pi L)
    const float iden = rcp(L + LUM_EPS);
    const float R = cl * iden;
    const float dR = (dcl - R) * iden;
    const float F = fmaf(p, R, tab[1]);
    float gyr, gyg, gyb, gdr, gdg, gdb;
    blend_bwd<HM>(r * F, r, mr, gr, gyr, gdr);
    blend_bwd<HM>(g * F, g, mg, gg, gyg, gdg);
    blend_bwd<HM>(b * F, b, mb, gb, gyb, gdb);
    const float Sgc = gyr * r + gyg * g + gyb * b;
    if (own) acc = fmaf(R - 1.0f, Sgc, acc);
    const float k = pf * dR * Sgc;
    gr = fmaf(0.27f, k, fmaf(gyr, F, gdr));
    gg = fmaf(0.67f, k, fmaf(gyg, F, gdg));
    gb = fmaf(0.06f, k, fmaf(gyb, F, gdb));
}

template <bool HM>
T2O_HD void bnw_bwd(const float *tab, float r, float g, float b, float mr, float mg, float mb,
                    float &gr, float &gg, float &gb, float &acc, bool own) {
    const float p = tab[0], q = tab[1];
    const float lum = lum_rn(r, g, b);
    const float pl = p * lum;
    float gyr, gyg, gyb, gdr, gdg, gdb;
    blend_bwd<HM>(fmaf(q, r, pl), r, mr, gr, gyr, gdr);
    blend_bwd<HM>(fmaf(q, g, pl), g, mg, gg, gyg, gdg);
    blend_bwd<HM>(fmaf(q, b, pl), b, mb, gb, gyb, gdb);
    const float Sg = gyr + gyg + gyb;
    if (own) acc += gyr * (lum - r) + gyg * (lum - g) + gyb * (lum - b);
    const float k = p * Sg;
    gr = fmaf(gyr, q, gdr) + 0.27f * k;
    gg = fmaf(gyg, q, gdg) + 0.67f * k;
    gb = fmaf(gyb, q, gdb) + 0.06f * k;
}

template <bool HM>
T2O_HD void hue_bwd(const float *tab, float r, float g, float b, float mr, float mg, float mb,
                    float &gr, float &gg, float &gb, float &acc, bool own) {
    const float v = max3(r, g, b), mn = min3(r, g, b);
    const float inv = rcp(v + HSV_EPS);
    const float s = (v - mn) * inv, kap = v * inv;
    const float vs = v * s;
    float gyr, gyg, gyb, gdr, gdg, gdb;
    blend_bwd<HM>(fmaf(-tab[1], vs, v), r, mr, gr, gyr, gdr);
    blend_bwd<HM>(fmaf(-tab[2], vs, v), g, mg, gg, gyg, gdg);
    blend_bwd<HM>(fmaf(-tab[3], vs, v), b, mb, gb, gyb, gdb);
    const float Sg = gyr + gyg + gyb;
    const float Sa = tab[1] * gyr + tab[2] * gyg + tab[3] * gyb;
    if (own) acc -= vs * (tab[4] * gyr + tab[5] * gyg + tab[6] * gyb);
    // y_c = v - a_c v (v - mn) / (v + eps): d/dv = 1 - a_c (s + kap (1 - s)), d/dmn = a_c kap; max / min route to the
    // first tied channel (gray pixels: both to channel 0, which then receives exactly Sg)
    const float Gv = Sg - Sa * fmaf(kap, 1.0f - s, s);
    const float Gmn = Sa * kap;
    const int im = argmax3(r, g, b), in = argmin3(r, g, b);
    gr = gdr + (im == 0 ? Gv : 0.0f) + (in == 0 ? Gmn : 0.0f);
    gg = gdg + (im == 1 ? Gv : 0.0f) + (in == 1 ? Gmn : 0.0f);
    gb = gdb + (im == 2 ? Gv : 0.0f) + (in == 2 ? Gmn : 0.0f);
}

// one channel of a curve operator: G[i] += g * clamp(L x - i, 0, 1)
// NG: the curve maps [0, 1] into [0, 1] (ct[CT_INRANGE]) and there is no mask: the output clamp passes every gradient
template <bool HM, bool CL, bool NG = false>
T2O_HD float curve_bwd(const float *ct, int L, float x, float m, float g, F2 *G, bool own) {
    const float xs = CL ? x : sat01(x);
    float tt, tfs, gy = g, gd = 0.0f, k_knot, k_seg;
    if (HM || !NG) {
        const F4 seg = curve_seg4(ct, xs, L, tt, tfs);
        blend_bwd<HM>(fmaf(seg.a, xs, seg.b), x, m, g, gy, gd);
        k_knot = seg.c; k_seg = seg.a;
    } else {
        const F2 seg = curve_seg2hi(ct, xs, L, tt, tfs);
        k_knot = seg.a; k_seg = seg.b;
    }
    const float ga = own ? gy : 0.0f;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < MAX_L / 2; ++i)
        G[i] = fma2(F2{ga, ga}, F2{sat01(fmaf(tt, CURVE_TT_INV, -(float)(2 * i))), sat01(fmaf(tt, CURVE_TT_INV, -(float)(2 * i + 1)))}, G[i]);
    const float slope = tt == tfs ? k_knot : k_seg;                 // exact knot: both clamp terms pass
    const float gx = gy * slope;
    return gd + ((CL || in01(x)) ? gx : 0.0f);
}
// may the backward of this curve operator skip its clamp gates?  (uniform per image and operator)
T2O_HD bool curve_in_range(int op, const float *tab) {
    return op == OP_TONE ? tab[CT_INRANGE] != 0.0f
                         : (tab[CT_INRANGE] != 0.0f && tab[CT + CT_INRANGE] != 0.0f && tab[2 * CT + CT_INRANGE] != 0.0f);
}

// dLoss/dk_i of one curve from its reduced accumulators G[0..L): (G[i] - C) / S with C = sum_j (k_j / S) G[j]
// (k_j / S = k'_j / L from the segment table)
T2O_HD float curve_param_grad(const float *ct, int L, const float *G, int i) {
    float C = 0.0f;
    for (int j = 0; j < L; ++j) C = fmaf(ct[4 * j], G[j], C);
    return ct[CT_INVS] * (G[i] - C / (float)L);
}

template <bool HM, bool CL = false, bool NG = false>
T2O_HD void pointwise_bwd(int op, const float *tab, int L, float r, float g, float b,
                          float mr, float mg, float mb,
                          float &gr, float &gg, float &gb, GradAcc &A, bool own) {
    switch (op) {
        case OP_BRIGHTNESS: brightness_bwd<HM, CL>(tab, r, g, b, mr, mg, mb, gr, gg, gb, A.bright, own); break;
        case OP_CONTRAST: contrast_bwd<HM>(tab, r, g, b, mr, mg, mb, gr, gg, gb, A.contrast, own); break;
        case OP_SATURATION: saturation_bwd<HM, CL>(tab, r, g, b, mr, mg, mb, gr, gg, gb, A.satur, own); break;
        case OP_TONE:
            gr = curve_bwd<HM, CL, NG>(tab, L, r, mr, gr, A.tone, own);
            gg = curve_bwd<HM, CL, NG>(tab, L, g, mg, gg, A.tone, own);
            gb = curve_bwd<HM, CL, NG>(tab, L, b, mb, gb, A.tone, own);
            break;
        case OP_COLOR:
            gr = curve_bwd<HM, CL, NG>(tab, L, r, mr, gr, A.color[0], own);
            gg = curve_bwd<HM, CL, NG>(tab + CT, L, g, mg, gg, A.color[1], own);
            gb = curve_bwd<HM, CL, NG>(tab + 2 * CT, L, b, mb, gb, A.color[2], own);
            break;
        case OP_BNW: bnw_bwd<HM>(tab, r, g, b, mr, mg, mb, gr, gg, gb, A.bnw, own); break;
        case OP_HUE: hue_bwd<HM>(tab, r, g, b, mr, mg, mb, gr, gg, gb, A.hue, own); break;
        case OP_WHITE: {
            float gy, gd;
            blend_bwd<HM>(1.0f, r, mr, gr, gy, gd); gr = gd;
            blend_bwd<HM>(1.0f, g, mg, gg, gy, gd); gg = gd;
            blend_bwd<HM>(1.0f, b, mb, gb, gy, gd); gb = gd;
            break;
        }
        case OP_EXPOSURE: {
            const float e = tab[1];
            float gyr, gyg, gyb, gdr, gdg, gdb;
            blend_bwd<HM>(r * e, r, mr, gr, gyr, gdr);
            blend_bwd<HM>(g * e, g, mg, gg, gyg, gdg);
            blend_bwd<HM>(b * e, b, mb, gb, gyb, gdb);
            if (own) A.expo += LN2_F * e * (gyr * r + gyg * g + gyb * b);
            gr = fmaf(gyr, e, gdr); gg = fmaf(gyg, e, gdg); gb = fmaf(gyb, e, gdb);
            break;
        }
        case OP_WHITEBALANCE: {
            float gyr, gyg, gyb, gdr, gdg, gdb;
            blend_bwd<HM>(r * tab[0], r, mr, gr, gyr, gdr);
            blend_bwd<HM>(g * tab[1], g, mg, gg, gyg, gdg);
            blend_bwd<HM>(b * tab[2], b, mb, gb, gyb, gdb);
            if (own) { A.wb[0] = fmaf(gyr, r, A.wb[0]); A.wb[1] = fmaf(gyg, g, A.wb[1]); A.wb[2] = fmaf(gyb, b, A.wb[2]); }
            gr = fmaf(gyr, tab[0], gdr); gg = fmaf(gyg, tab[1], gdg); gb = fmaf(gyb, tab[2], gdb);
            break;
        }
        default: break;     // identity: gradient passes unchanged
    }
}

}  // namespace t2o
