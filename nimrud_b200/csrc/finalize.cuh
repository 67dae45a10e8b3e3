// finalize.cuh -- from integer neighborhood moments to the feature columns.
//
// The search set is a lattice, so a neighbor is an integer cell offset j = k - c from the query's
// anchor cell c.  Every kernel accumulates n, sum(j) and sum(j j^T) EXACTLY in integers; this file
// turns them into the reference's columns (nimrud/minimal/features.py:21-57):
//     population                         n
//     centroid distance                  | q - mean(centres) | = e * | f - 0.5 - mean(j) |
//     l_max / sum(l), l_mid / sum(l)      eigenvalues of cov = e^2 (n S2 - S1 S1^T) / (n (n-1))
// (the ratios are scale free, so the integer matrix n*S2 - S1*S1^T is solved directly).
// undefined features are zeros (minimal/multiscale.py:4-5): ratios when n < 2 or the trace is 0,
// centroid when n == 0.
#pragma once
#include "common.cuh"

namespace nbr {

struct Moments {
    long long n;
    long long s1[3];
    long long s2[6];   // xx xy xz yy yz zz
};

#ifdef __CUDACC__
// single-instruction reciprocal / square root (MUFU, ~1 ulp, denormals flushed).  1.0f / x, sqrtf and
// __frcp_rn all carry a correctly-rounded slow path (~15 instructions and a call each); every use below
// either feeds a float32 guess that is refined in float64 or a column with a 1e-4 tolerance.
__device__ __forceinline__ float rcp_fast(float x)
{
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sqrt_fast(float x)
{
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// eigenvalues (descending) of a symmetric positive semi-definite 3x3 with unit trace, and the unit
// eigenvector of the smallest one.
//
// the trigonometric closed form alone loses half the digits on a (near-)double eigenvalue
// (acos near +-1), which is the normal case here: lines give l2 = l3 = 0, flat patches l1 = l2.
// so it is only used to pick the ISOLATED eigenvalue (largest if the half-determinant r >= 0,
// smallest otherwise) and its eigenvector, in float32; that eigenpair is then deflated in float64
// (Rayleigh quotient + the 2x2 block in the orthogonal complement through the cancellation-free
// m +- hypot(.,.) form).  the float32 error of the eigenvector enters at second order only, so every
// eigenvalue comes out with absolute error ~1e-13, clusters included.
__device__ __forceinline__ void sym_mul(const double a[6], const double x[3], double y[3])
{
    y[0] = a[0] * x[0] + a[1] * x[1] + a[2] * x[2];
    y[1] = a[1] * x[0] + a[3] * x[1] + a[4] * x[2];
    y[2] = a[2] * x[0] + a[4] * x[1] + a[5] * x[2];
}

// float64 closed form (trigonometric); only used when the matrix is (nearly) a multiple of the
// identity, where the three eigenvalues are within ~1e-3 of each other and half-precision loss on a
// double root is harmless.
// arguments and result by value: a pointer to the caller's arrays would pin them in local memory on the hot path too.
struct Eig3 { double l0, l1, l2; };
static __device__ __noinline__ Eig3 eig3_near_isotropic(double a0, double a1, double a2, double a3, double a4, double a5)
{
    const double q = 1.0 / 3.0;
    Eig3 out = {q, q, q};
    const double p1 = a1 * a1 + a2 * a2 + a4 * a4;
    const double d0 = a0 - q, d1 = a3 - q, d2 = a5 - q;
    const double p2 = d0 * d0 + d1 * d1 + d2 * d2 + 2.0 * p1;
    if (!(p2 > 1e-28)) return out;
    const double p = sqrt(p2 / 6.0);
    const double ip = 1.0 / p;
    const double b0 = d0 * ip, b1 = a1 * ip, b2 = a2 * ip, b3 = d1 * ip, b4 = a4 * ip, b5 = d2 * ip;
    double r = 0.5 * (b0 * (b3 * b5 - b4 * b4) - b1 * (b1 * b5 - b4 * b2) + b2 * (b1 * b4 - b3 * b2));
    r = fmin(1.0, fmax(-1.0, r));
    const double phi = acos(r) * (1.0 / 3.0);
    const double e0 = q + 2.0 * p * cos(phi), e2 = q + 2.0 * p * cos(phi + 2.0943951023931954923);
    out.l0 = e0; out.l2 = e2; out.l1 = 1.0 - e0 - e2;
    return out;
}

// WANT_NORMAL: also the unit eigenvectors of the smallest (normal) and of the largest (lead) eigenvalue
template <bool WANT_NORMAL>
__device__ __forceinline__ void eig3_unit_trace(const double a[6], double l[3], double normal[3], double lead[3])
{
    normal[0] = 0.0; normal[1] = 0.0; normal[2] = 1.0;
    lead[0] = 1.0; lead[1] = 0.0; lead[2] = 0.0;
    // ---- float32: which eigenvalue is isolated, and its eigenvector to ~1e-6
    const float qf = 1.0f / 3.0f;
    const float f0 = (float)a[0], f1 = (float)a[1], f2 = (float)a[2], f3 = (float)a[3], f4 = (float)a[4],
                f5 = (float)a[5];
    const float g0 = f0 - qf, g1 = f3 - qf, g2 = f5 - qf;
    const float p2 = g0 * g0 + g1 * g1 + g2 * g2 + 2.0f * (f1 * f1 + f2 * f2 + f4 * f4);
    if (!(p2 > 1e-7f)) { const Eig3 e = eig3_near_isotropic(a[0], a[1], a[2], a[3], a[4], a[5]); l[0] = e.l0; l[1] = e.l1; l[2] = e.l2; return; }
    const float p26 = p2 * (1.0f / 6.0f);
    const float ip = rsqrtf(p26);
    const float p = p26 * ip;
    const float b0 = g0 * ip, b1 = f1 * ip, b2 = f2 * ip, b3 = g1 * ip, b4 = f4 * ip, b5 = g2 * ip;
    float r = 0.5f * (b0 * (b3 * b5 - b4 * b4) - b1 * (b1 * b5 - b4 * b2) + b2 * (b1 * b4 - b3 * b2));
    r = fminf(1.0f, fmaxf(-1.0f, r));
    const float phi = acosf(r) * (1.0f / 3.0f);
    const bool top = r >= 0.0f;          // true: the largest eigenvalue is the isolated one
    const float lamf = top ? qf + 2.0f * p * __cosf(phi) : qf + 2.0f * p * __cosf(phi + 2.0943951f);
    // best-conditioned cross product of rows of (A - lam I)
    const float r00 = f0 - lamf, r11 = f3 - lamf, r22 = f5 - lamf;
    const float c0x = f1 * f4 - f2 * r11, c0y = f2 * f1 - r00 * f4, c0z = r00 * r11 - f1 * f1;      // row0 x row1
    const float c1x = f1 * r22 - f2 * f4, c1y = f2 * f2 - r00 * r22, c1z = r00 * f4 - f1 * f2;      // row0 x row2
    const float c2x = r11 * r22 - f4 * f4, c2y = f4 * f2 - f1 * r22, c2z = f1 * f4 - r11 * f2;      // row1 x row2
    const float n0 = c0x * c0x + c0y * c0y + c0z * c0z;
    const float n1 = c1x * c1x + c1y * c1y + c1z * c1z;
    const float n2 = c2x * c2x + c2y * c2y + c2z * c2z;
    float vx = c0x, vy = c0y, vz = c0z, nn = n0;
    if (n1 > nn) { nn = n1; vx = c1x; vy = c1y; vz = c1z; }
    if (n2 > nn) { nn = n2; vx = c2x; vy = c2y; vz = c2z; }
    if (!(nn > 1e-30f)) { const Eig3 e = eig3_near_isotropic(a[0], a[1], a[2], a[3], a[4], a[5]); l[0] = e.l0; l[1] = e.l1; l[2] = e.l2; return; }
    const float invf = rsqrtf(nn);

    // ---- float64: exact deflation around that (approximate) eigenvector.  an error d in v only
    // enters the results at second order (spread * d^2 ~ 1e-13).  v is NOT renormalised in float64: with
    // s = v.v, u1 = e_k x v (exactly orthogonal to v, entries are copies of v's) and u2 = v x u1,
    //     lam = v.Av / s,   m00 = u1.Au1 / nu,   m11 = u2.Au2 / (s nu),   m01^2 = (u1.Au2)^2 / (s nu^2),  nu = |u1|^2,
    // so the only divisions are two reciprocals (float32 seed + Newton) and the only root is the
    // discriminant's (float32 rsqrt seed + one Newton step): no float64 sqrt/rsqrt library calls.
    const double v[3] = {(double)(vx * invf), (double)(vy * invf), (double)(vz * invf)};
    const double s = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];      // 1 +- 1e-6
    double rs = 2.0 - s;
    rs = rs * (2.0 - s * rs);                                        // 1/s to 1e-24
    double av[3];
    sym_mul(a, v, av);
    const double lam = (v[0] * av[0] + v[1] * av[1] + v[2] * av[2]) * rs;           // Rayleigh quotient

    double u1[3], u2[3];
    const float ax = fabsf(vx), ay = fabsf(vy), az = fabsf(vz);
    double nu;
    if (ax <= ay && ax <= az) { u1[0] = 0.0; u1[1] = -v[2]; u1[2] = v[1]; nu = s - v[0] * v[0]; }
    else if (ay <= az)        { u1[0] = v[2]; u1[1] = 0.0; u1[2] = -v[0]; nu = s - v[1] * v[1]; }
    else                      { u1[0] = -v[1]; u1[1] = v[0]; u1[2] = 0.0; nu = s - v[2] * v[2]; }
    u2[0] = v[1] * u1[2] - v[2] * u1[1];
    u2[1] = v[2] * u1[0] - v[0] * u1[2];
    u2[2] = v[0] * u1[1] - v[1] * u1[0];
    double in1 = (double)rcp_fast((float)nu);                        // nu = |u1|^2 >= 2/3 s
    in1 = in1 * (2.0 - nu * in1);
    double au1[3], au2[3];
    sym_mul(a, u1, au1);
    sym_mul(a, u2, au2);
    const double m00 = (u1[0] * au1[0] + u1[1] * au1[1] + u1[2] * au1[2]) * in1;
    const double m11 = (u2[0] * au2[0] + u2[1] * au2[1] + u2[2] * au2[2]) * in1 * rs;
    const double c01 = (u1[0] * au2[0] + u1[1] * au2[1] + u1[2] * au2[2]) * in1;   // m01 * sqrt(s)
    const double mid = 0.5 * (m00 + m11);
    const double hd = 0.5 * (m00 - m11);
    const double disc = hd * hd + c01 * c01 * rs;
    double rad = 0.0;
    if (disc > 1.0e-30) {                                            // below: a double root to 1e-15
        const double y0 = (double)rsqrtf((float)disc);
        const double y1 = y0 * (1.5 - 0.5 * disc * y0 * y0);
        rad = disc * y1;
    }
    const double hi = mid + rad, lo = mid - rad;
    if (top) {
        l[0] = lam; l[1] = hi; l[2] = lo;
        if (WANT_NORMAL) {
            // eigenvector of the 2x2 block for `lo`, mapped back through the normalised u1, u2
            const double m01 = c01 * rsqrt(s);
            double w0 = m01, w1 = lo - m00;
            if (fabs(lo - m11) > fabs(w1)) { w0 = lo - m11; w1 = m01; }
            const double wn = w0 * w0 + w1 * w1;
            const double i1 = rsqrt(nu), i2 = rsqrt(nu * s);
            if (wn > 0.0) {
                const double iw = rsqrt(wn);
                w0 *= iw * i1; w1 *= iw * i2;
                normal[0] = w0 * u1[0] + w1 * u2[0];
                normal[1] = w0 * u1[1] + w1 * u2[1];
                normal[2] = w0 * u1[2] + w1 * u2[2];
            } else {
                normal[0] = u2[0] * i2; normal[1] = u2[1] * i2; normal[2] = u2[2] * i2;
            }
            const double iv = rsqrt(s);
            lead[0] = v[0] * iv; lead[1] = v[1] * iv; lead[2] = v[2] * iv;
        }
    } else {
        l[0] = hi; l[1] = lo; l[2] = lam;
        if (WANT_NORMAL) {
            const double iv = rsqrt(s);
            normal[0] = v[0] * iv; normal[1] = v[1] * iv; normal[2] = v[2] * iv;
            // eigenvector of the 2x2 block for `hi`, mapped back through the normalised u1, u2
            const double m01 = c01 * iv;
            double w0 = m01, w1 = hi - m00;
            if (fabs(hi - m11) > fabs(w1)) { w0 = hi - m11; w1 = m01; }
            const double wn = w0 * w0 + w1 * w1;
            const double i1 = rsqrt(nu), i2 = rsqrt(nu * s);
            if (wn > 0.0) {
                const double iw = rsqrt(wn);
                w0 *= iw * i1; w1 *= iw * i2;
                lead[0] = w0 * u1[0] + w1 * u2[0];
                lead[1] = w0 * u1[1] + w1 * u2[1];
                lead[2] = w0 * u1[2] + w1 * u2[2];
            } else {
                lead[0] = u1[0] * i1; lead[1] = u1[1] * i1; lead[2] = u1[2] * i1;
            }
        }
    }
}

// ---- eigenvalue ratios from SMALL INTEGER matrices (the window kernels' fast path, reference columns only)
//
// the window kernels hand over n*S2 - S1*S1^T as exact integers below 2^28.  for those the ratios do not need an
// eigenvector: with c1 = sum of the principal 2x2 minors / tr^2 (EXACT numerator: int64) and c0 = det / tr^3 the
// normalised spectrum solves  l^3 - l^2 + c1 l - c0 = 0.
//   * the isolated root (largest if the depressed cubic's cos(theta) >= 0, smallest otherwise) is well conditioned:
//     float32 trigonometric estimate, two Newton steps in float64;
//   * the other two follow from the deflated quadratic  m^2 - s m + P,  s = 1 - l,  P = c1 - l s.  a (near-)double
//     root of the pair costs sqrt(1e-16) = 1e-8 absolute, i.e. <= 1e-5 relative for any pair a window of integer
//     offsets can produce (the smallest non-zero normalised eigenvalue of such a matrix is > 1e-4) -- and the one case
//     where the pair is a double root at ZERO (collinear voxels, rank <= 1) is detected exactly: c1's numerator is 0.
// about a third of the instructions of the eigenvector-deflation route (eig3_unit_trace), which stays in use for the
// extended columns, the interval / exact / kNN kernels (larger integers) and near-isotropic matrices.
#ifndef NBR_FAST_EIG
#define NBR_FAST_EIG 1
#endif
__device__ __forceinline__ bool eig_ratios_small_int(int a0, int a1, int a2, int a3, int a4, int a5, double tr, double it,
                                                     double &l0, double &l1)
{
    const long long m01 = (long long)a0 * a3 - (long long)a1 * a1;
    const long long m02 = (long long)a0 * a5 - (long long)a2 * a2;
    const long long m12 = (long long)a3 * a5 - (long long)a4 * a4;
    const long long C1 = m01 + m02 + m12;                   // exact; > 0 unless the voxels are collinear
    if (C1 <= 0) { l0 = 1.0; l1 = 0.0; return true; }
    const double d01 = (double)((long long)a1 * a5 - (long long)a4 * a2), d02 = (double)((long long)a1 * a4 - (long long)a3 * a2);
    const double D = (double)a0 * (double)m12 - (double)a1 * d01 + (double)a2 * d02;
    it = it * (2.0 - tr * it);                               // 1 / tr to the last bit: c1 and c0 feed a square root below
    const double it2 = it * it;
    const double c1 = (double)C1 * it2;
    const double c0 = fmax(D * it2 * it, 0.0);
    // depressed cubic m^3 + a m + b, l = 1/3 + m
    const float p2 = (float)(1.0 / 3.0 - c1) * (1.0f / 3.0f);          // -a / 3
    if (!(p2 > 1.0e-7f)) return false;                                  // near-isotropic: the careful route
    const float bf = (float)(c1 * (1.0 / 3.0) - 2.0 / 27.0 - c0);
    const float ip = rsqrtf(p2);
    const float pf = p2 * ip;
    float r = -0.5f * bf * ip * ip * ip;                                // cos(theta)
    r = fminf(1.0f, fmaxf(-1.0f, r));
    const float phi = acosf(r) * (1.0f / 3.0f);
    const bool top = r >= 0.0f;                                         // the largest root is the isolated one
    double lam = 1.0 / 3.0 + (double)(2.0f * pf * __cosf(top ? phi : phi + 2.0943951f));
#pragma unroll
    for (int itn = 0; itn < 2; ++itn) {
        const double pl = ((lam - 1.0) * lam + c1) * lam - c0;
        const double dp = (3.0 * lam - 2.0) * lam + c1;                 // (l - la)(l - lb): away from 0 at the isolated root
        lam -= pl * (double)rcp_fast((float)dp);
    }
    const double s = 1.0 - lam, mid = 0.5 * s;
    const double P = c1 - lam * s;
    const double disc = mid * mid - P;
    double rad = 0.0;
    if (disc > 1.0e-30) {
        const double y0 = (double)rsqrtf((float)disc);
        rad = disc * (y0 * (1.5 - 0.5 * disc * y0 * y0));
    }
    if (top) { l0 = lam; l1 = mid + rad; }
    else     { l0 = mid + rad; l1 = mid - rad; }
    return true;
}

// x > 0, else x == 0 and y > 0, else z > 0: the sign convention of the emitted leading eigenvectors
__device__ __forceinline__ void canonical_sign_xy(double w[3])
{
    if (w[0] < 0 || (w[0] == 0 && (w[1] < 0 || (w[1] == 0 && w[2] < 0)))) { w[0] = -w[0]; w[1] = -w[1]; w[2] = -w[2]; }
}

// the 22 extension columns (not on the reference path; kept out of line so the hot kernels stay small)
static __device__ __noinline__ void extended_descriptors(const double l[3], double v[3], double lead[3], double trace,
                                                        const double a[6], double ext[NBR_COLS_EXTENDED - 4])
{
    const double e1 = fmax(l[0], 0.0), e2 = fmax(l[1], 0.0), e3 = fmax(l[2], 0.0);
    const double i1 = 1.0 / e1;
    ext[0] = (e1 - e2) * i1;               // linearity
    ext[1] = (e2 - e3) * i1;               // planarity
    ext[2] = e3 * i1;                      // sphericity
    ext[3] = cbrt(e1 * e2 * e3);           // omnivariance
    ext[4] = (e1 - e3) * i1;               // anisotropy
    double ent = 0.0;
    if (e1 > 0) ent -= e1 * log(e1);
    if (e2 > 0) ent -= e2 * log(e2);
    if (e3 > 0) ent -= e3 * log(e3);
    ext[5] = ent;                          // eigenentropy
    ext[6] = e3;                           // change of curvature
    if (v[2] < 0 || (v[2] == 0 && (v[1] < 0 || (v[1] == 0 && v[0] < 0)))) {
        v[0] = -v[0]; v[1] = -v[1]; v[2] = -v[2];
    }
    ext[7] = 1.0 - fabs(v[2]);             // verticality
    ext[8] = v[0]; ext[9] = v[1]; ext[10] = v[2];
    ext[11] = trace;                       // trace of the ddof=1 covariance
#pragma unroll
    for (int i = 0; i < 6; ++i) ext[12 + i] = a[i] * trace;     // covariance (ddof = 1): xx xy xz yy yz zz; a has unit trace
    // x, y of the unit eigenvectors of the largest and the middle eigenvalue (the legacy OG_MSO keeps the first two
    // components of two eigenvectors, nimrud/prototypes/mso.py:1498-1539); middle = normal x lead
    double second[3] = {v[1] * lead[2] - v[2] * lead[1], v[2] * lead[0] - v[0] * lead[2], v[0] * lead[1] - v[1] * lead[0]};
    canonical_sign_xy(lead);
    canonical_sign_xy(second);
    ext[18] = lead[0]; ext[19] = lead[1]; ext[20] = second[0]; ext[21] = second[1];
}

// the columns from: n, the centroid distance, and the UNNORMALISED matrix a = n*S2 - S1*S1^T
// (exact integers converted to double).  writes 4 (reference) or 26 (extended) columns at out[0..]
template <typename OutT>
__device__ __forceinline__ void emit_core(long long n_int, double centroid, double a[6], double edge, OutT *out,
                                          int descriptor_mask, const int *small_ints = nullptr)
{
    const int ncol = (descriptor_mask & NBR_DESC_EXTENDED) ? NBR_COLS_EXTENDED : NBR_COLS_REFERENCE;
    const double n = (double)n_int;
    double l0 = 0.0, l1 = 0.0;
    constexpr int NEXT = NBR_COLS_EXTENDED - NBR_COLS_REFERENCE;
    double ext[NEXT];
#pragma unroll
    for (int i = 0; i < NEXT; ++i) ext[i] = 0.0;
    if (n_int >= 2) {
        const double tr = a[0] + a[3] + a[5];
        if (tr > 0.0) {
            // 1/tr: float32 seed + one Newton step (relative error ~1e-14)
            const double r0 = (double)rcp_fast((float)tr);
            const double it = r0 * (2.0 - tr * r0);
            bool done = false;
#if NBR_FAST_EIG
            if (small_ints && !(descriptor_mask & NBR_DESC_EXTENDED))
                done = eig_ratios_small_int(small_ints[0], small_ints[1], small_ints[2], small_ints[3], small_ints[4], small_ints[5],
                                            tr, it, l0, l1);
#endif
            if (!done) {
#pragma unroll
                for (int i = 0; i < 6; ++i) a[i] *= it;
                double l[3], v[3], lead[3];
                if (descriptor_mask & NBR_DESC_EXTENDED) eig3_unit_trace<true>(a, l, v, lead);
                else eig3_unit_trace<false>(a, l, v, lead);
                l0 = l[0];
                l1 = l[1];
                if ((descriptor_mask & NBR_DESC_EXTENDED) && n_int >= 3)
                    extended_descriptors(l, v, lead, tr * edge * edge / (n * (n - 1.0)), a, ext);
            }
        }
    }
    out[0] = (OutT)n;
    out[1] = (OutT)centroid;
    out[2] = (OutT)l0;
    out[3] = (OutT)l1;
    if (ncol > 4) {
#pragma unroll
        for (int i = 0; i < NEXT; ++i) out[4 + i] = (OutT)ext[i];
    }
}

// int64 moments relative to the anchor cell (exact kernel, kNN)
template <typename OutT>
__device__ __forceinline__ void emit_features(const Moments &m, const double f[3], double edge, OutT *out,
                                              int descriptor_mask)
{
    double centroid = 0.0;
    if (m.n > 0) {
        const double inv = 1.0 / (double)m.n;
        const double dx = (f[0] - 0.5 - (double)m.s1[0] * inv) * edge;
        const double dy = (f[1] - 0.5 - (double)m.s1[1] * inv) * edge;
        const double dz = (f[2] - 0.5 - (double)m.s1[2] * inv) * edge;
        centroid = sqrt(dx * dx + dy * dy + dz * dz);
    }
    double a[6];
    a[0] = (double)(m.n * m.s2[0] - m.s1[0] * m.s1[0]);
    a[1] = (double)(m.n * m.s2[1] - m.s1[0] * m.s1[1]);
    a[2] = (double)(m.n * m.s2[2] - m.s1[0] * m.s1[2]);
    a[3] = (double)(m.n * m.s2[3] - m.s1[1] * m.s1[1]);
    a[4] = (double)(m.n * m.s2[4] - m.s1[1] * m.s1[2]);
    a[5] = (double)(m.n * m.s2[5] - m.s1[2] * m.s1[2]);
    emit_core<OutT>(m.n, centroid, a, edge, out, descriptor_mask);
}

// int32 moments in WINDOW coordinates j' = j + W (row kernel).  the matrix n*S2 - S1*S1^T does not
// depend on the origin, only the centroid needs the shift.  fm[a] = f[a] - 0.5 + W (float32): the query
// in window units.  SMALL: every product fits int32 (W <= 6).
template <typename OutT>
__device__ __forceinline__ void emit_features_window(int n, int sx, int sy, int sz, int sxx, int sxy, int sxz, int syy,
                                                     int syz, int szz, float fxm, float fym, float fzm, bool SMALL,
                                                     double edge, OutT *out, int descriptor_mask)
{
    double centroid = 0.0;
    if (n > 0) {
        if (sizeof(OutT) == 8) {
            // float64 rows: the sums are exact integers, only the query's window position (fxm ..., float32) limits
            // the column: |error| <= 3e-7 * edge, relative accuracy is kept when the centroid is close to the query
            const double inv = 1.0 / (double)n;
            const double dx = (double)fxm - (double)sx * inv, dy = (double)fym - (double)sy * inv, dz = (double)fzm - (double)sz * inv;
            centroid = sqrt(dx * dx + dy * dy + dz * dz) * edge;
        } else {
            const float inv = rcp_fast((float)n);
            const float dx = fxm - (float)sx * inv, dy = fym - (float)sy * inv, dz = fzm - (float)sz * inv;
            centroid = (double)sqrt_fast(dx * dx + dy * dy + dz * dz) * edge;
        }
    }
    double a[6];
    int ai[6] = {0, 0, 0, 0, 0, 0};
    if (SMALL) {
        ai[0] = n * sxx - sx * sx; ai[1] = n * sxy - sx * sy; ai[2] = n * sxz - sx * sz;
        ai[3] = n * syy - sy * sy; ai[4] = n * syz - sy * sz; ai[5] = n * szz - sz * sz;
#pragma unroll
        for (int i = 0; i < 6; ++i) a[i] = (double)ai[i];
    } else {
        const long long N = n, X = sx, Y = sy, Z = sz;
        a[0] = (double)(N * sxx - X * X); a[1] = (double)(N * sxy - X * Y); a[2] = (double)(N * sxz - X * Z);
        a[3] = (double)(N * syy - Y * Y); a[4] = (double)(N * syz - Y * Z); a[5] = (double)(N * szz - Z * Z);
    }
    emit_core<OutT>(n, centroid, a, edge, out, descriptor_mask, SMALL ? ai : nullptr);
}

// anchor cell and fractional position of a query on one axis.  c is clamped so that far-away
// queries cannot overflow; f in [0,1) when unclamped.
__device__ __forceinline__ void query_anchor(double q, const GridDev &g, int a, int &c, double &f)
{
    // the anchor only positions the window, so a reciprocal multiply is as good as the division
    const double u = (q - g.minc[a]) * g.inv_edge;
    const double cf = floor(u);
    f = u - cf;
    // local cell coordinate: saturating conversion + integer clamp
    c = max(min(__double2int_rn(cf - (double)g.cell_lo[a]), 1000000000), -1000000000);
}
#endif

}  // namespace nbr
