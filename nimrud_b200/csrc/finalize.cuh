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
// eigenvalues (descending) of a symmetric positive semi-definite 3x3 with unit trace, and the unit
// eigenvector of the smallest one.
//
// the trigonometric closed form alone loses half the digits on a (near-)double eigenvalue
// (acos near +-1), which is the normal case here: lines give l2 = l3 = 0, flat patches l1 = l2.
// so it is only used to pick the ISOLATED eigenvalue (largest if the half-determinant r >= 0,
// smallest otherwise), which it gets to full precision; that eigenpair is deflated and the other two
// come from the 2x2 block in the orthogonal complement through the cancellation-free
// m +- hypot(.,.) form.  absolute error ~1e-16 on every eigenvalue, clusters included.
__device__ __forceinline__ void sym_mul(const double a[6], const double x[3], double y[3])
{
    y[0] = a[0] * x[0] + a[1] * x[1] + a[2] * x[2];
    y[1] = a[1] * x[0] + a[3] * x[1] + a[4] * x[2];
    y[2] = a[2] * x[0] + a[4] * x[1] + a[5] * x[2];
}

__device__ __forceinline__ void eig3_unit_trace(const double a[6], double l[3], double normal[3])
{
    const double q = 1.0 / 3.0;
    normal[0] = 0.0; normal[1] = 0.0; normal[2] = 1.0;
    const double p1 = a[1] * a[1] + a[2] * a[2] + a[4] * a[4];
    const double d0 = a[0] - q, d1 = a[3] - q, d2 = a[5] - q;
    const double p2 = d0 * d0 + d1 * d1 + d2 * d2 + 2.0 * p1;
    if (!(p2 > 1e-28)) { l[0] = l[1] = l[2] = q; return; }
    const double p = sqrt(p2 / 6.0);
    const double ip = 1.0 / p;
    const double b0 = d0 * ip, b1 = a[1] * ip, b2 = a[2] * ip, b3 = d1 * ip, b4 = a[4] * ip, b5 = d2 * ip;
    double r = 0.5 * (b0 * (b3 * b5 - b4 * b4) - b1 * (b1 * b5 - b4 * b2) + b2 * (b1 * b4 - b3 * b2));
    r = fmin(1.0, fmax(-1.0, r));
    const double phi = acos(r) * (1.0 / 3.0);
    const bool top = r >= 0.0;          // true: the largest eigenvalue is the isolated one
    double lam = top ? q + 2.0 * p * cos(phi) : q + 2.0 * p * cos(phi + 2.0943951023931954923);

    // eigenvector of the isolated eigenvalue: best-conditioned cross product of rows of (A - lam I)
    const double r0[3] = {a[0] - lam, a[1], a[2]};
    const double r1[3] = {a[1], a[3] - lam, a[4]};
    const double r2[3] = {a[2], a[4], a[5] - lam};
    double c0[3] = {r0[1] * r1[2] - r0[2] * r1[1], r0[2] * r1[0] - r0[0] * r1[2], r0[0] * r1[1] - r0[1] * r1[0]};
    double c1[3] = {r0[1] * r2[2] - r0[2] * r2[1], r0[2] * r2[0] - r0[0] * r2[2], r0[0] * r2[1] - r0[1] * r2[0]};
    double c2[3] = {r1[1] * r2[2] - r1[2] * r2[1], r1[2] * r2[0] - r1[0] * r2[2], r1[0] * r2[1] - r1[1] * r2[0]};
    const double n0 = c0[0] * c0[0] + c0[1] * c0[1] + c0[2] * c0[2];
    const double n1 = c1[0] * c1[0] + c1[1] * c1[1] + c1[2] * c1[2];
    const double n2 = c2[0] * c2[0] + c2[1] * c2[1] + c2[2] * c2[2];
    double v[3], nn = n0;
    v[0] = c0[0]; v[1] = c0[1]; v[2] = c0[2];
    if (n1 > nn) { nn = n1; v[0] = c1[0]; v[1] = c1[1]; v[2] = c1[2]; }
    if (n2 > nn) { nn = n2; v[0] = c2[0]; v[1] = c2[1]; v[2] = c2[2]; }
    if (!(nn > 1e-60)) {
        // (numerically) three equal eigenvalues: the closed form is as good as anything
        const double e0 = q + 2.0 * p * cos(phi), e2 = q + 2.0 * p * cos(phi + 2.0943951023931954923);
        l[0] = e0; l[2] = e2; l[1] = 1.0 - e0 - e2;
        return;
    }
    const double inv = rsqrt(nn);
    v[0] *= inv; v[1] *= inv; v[2] *= inv;
    double av[3];
    sym_mul(a, v, av);
    lam = v[0] * av[0] + v[1] * av[1] + v[2] * av[2];           // Rayleigh quotient

    // orthonormal basis of the complement
    double u1[3], u2[3];
    const double ax = fabs(v[0]), ay = fabs(v[1]), az = fabs(v[2]);
    if (ax <= ay && ax <= az) { u1[0] = 0.0; u1[1] = -v[2]; u1[2] = v[1]; }
    else if (ay <= az)        { u1[0] = v[2]; u1[1] = 0.0; u1[2] = -v[0]; }
    else                      { u1[0] = -v[1]; u1[1] = v[0]; u1[2] = 0.0; }
    const double iu = rsqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
    u1[0] *= iu; u1[1] *= iu; u1[2] *= iu;
    u2[0] = v[1] * u1[2] - v[2] * u1[1];
    u2[1] = v[2] * u1[0] - v[0] * u1[2];
    u2[2] = v[0] * u1[1] - v[1] * u1[0];
    double au1[3], au2[3];
    sym_mul(a, u1, au1);
    sym_mul(a, u2, au2);
    const double m00 = u1[0] * au1[0] + u1[1] * au1[1] + u1[2] * au1[2];
    const double m01 = u1[0] * au2[0] + u1[1] * au2[1] + u1[2] * au2[2];
    const double m11 = u2[0] * au2[0] + u2[1] * au2[1] + u2[2] * au2[2];
    const double mid = 0.5 * (m00 + m11);
    const double hd = 0.5 * (m00 - m11);
    const double rad = sqrt(hd * hd + m01 * m01);
    const double hi = mid + rad, lo = mid - rad;
    if (top) {
        l[0] = lam; l[1] = hi; l[2] = lo;
        // eigenvector of the 2x2 block for `lo`, mapped back
        double w0 = m01, w1 = lo - m00;
        if (fabs(lo - m11) > fabs(w1)) { w0 = lo - m11; w1 = m01; }
        const double wn = w0 * w0 + w1 * w1;
        if (wn > 0.0) {
            const double iw = rsqrt(wn);
            w0 *= iw; w1 *= iw;
            normal[0] = w0 * u1[0] + w1 * u2[0];
            normal[1] = w0 * u1[1] + w1 * u2[1];
            normal[2] = w0 * u1[2] + w1 * u2[2];
        } else {
            normal[0] = u2[0]; normal[1] = u2[1]; normal[2] = u2[2];
        }
    } else {
        l[0] = hi; l[1] = lo; l[2] = lam;
        normal[0] = v[0]; normal[1] = v[1]; normal[2] = v[2];
    }
}

// writes 4 (reference) or 16 (extended) columns at out[0..]
template <typename OutT>
__device__ __forceinline__ void emit_features(const Moments &m, const double f[3], double edge, OutT *out,
                                              int descriptor_mask)
{
    const int ncol = (descriptor_mask & NBR_DESC_EXTENDED) ? NBR_COLS_EXTENDED : NBR_COLS_REFERENCE;
    double col[NBR_COLS_EXTENDED];
#pragma unroll
    for (int i = 0; i < NBR_COLS_EXTENDED; ++i) col[i] = 0.0;
    const double n = (double)m.n;
    col[0] = n;
    if (m.n > 0) {
        const double inv = 1.0 / n;
        const double dx = (f[0] - 0.5 - (double)m.s1[0] * inv) * edge;
        const double dy = (f[1] - 0.5 - (double)m.s1[1] * inv) * edge;
        const double dz = (f[2] - 0.5 - (double)m.s1[2] * inv) * edge;
        col[1] = sqrt(dx * dx + dy * dy + dz * dz);
    }
    if (m.n >= 2) {
        // exact integers: n*S2 - S1*S1^T  (= n^2 * biased covariance in cell units)
        double a[6];
        a[0] = (double)(m.n * m.s2[0] - m.s1[0] * m.s1[0]);
        a[1] = (double)(m.n * m.s2[1] - m.s1[0] * m.s1[1]);
        a[2] = (double)(m.n * m.s2[2] - m.s1[0] * m.s1[2]);
        a[3] = (double)(m.n * m.s2[3] - m.s1[1] * m.s1[1]);
        a[4] = (double)(m.n * m.s2[4] - m.s1[1] * m.s1[2]);
        a[5] = (double)(m.n * m.s2[5] - m.s1[2] * m.s1[2]);
        const double tr = a[0] + a[3] + a[5];
        if (tr > 0.0) {
            const double it = 1.0 / tr;
#pragma unroll
            for (int i = 0; i < 6; ++i) a[i] *= it;
            double l[3], v[3];
            eig3_unit_trace(a, l, v);
            col[2] = l[0];
            col[3] = l[1];
            if ((descriptor_mask & NBR_DESC_EXTENDED) && m.n >= 3) {
                const double e1 = fmax(l[0], 0.0), e2 = fmax(l[1], 0.0), e3 = fmax(l[2], 0.0);
                const double i1 = 1.0 / e1;
                col[4] = (e1 - e2) * i1;               // linearity
                col[5] = (e2 - e3) * i1;               // planarity
                col[6] = e3 * i1;                      // sphericity
                col[7] = cbrt(e1 * e2 * e3);           // omnivariance
                col[8] = (e1 - e3) * i1;               // anisotropy
                double ent = 0.0;
                if (e1 > 0) ent -= e1 * log(e1);
                if (e2 > 0) ent -= e2 * log(e2);
                if (e3 > 0) ent -= e3 * log(e3);
                col[9] = ent;                          // eigenentropy
                col[10] = e3;                          // change of curvature
                if (v[2] < 0 || (v[2] == 0 && (v[1] < 0 || (v[1] == 0 && v[0] < 0)))) {
                    v[0] = -v[0]; v[1] = -v[1]; v[2] = -v[2];
                }
                col[11] = 1.0 - fabs(v[2]);            // verticality
                col[12] = v[0]; col[13] = v[1]; col[14] = v[2];
                col[15] = tr * edge * edge / (n * (n - 1.0));   // trace of the ddof=1 covariance
            }
        }
    }
    for (int i = 0; i < ncol; ++i) out[i] = (OutT)col[i];
}

// anchor cell and fractional position of a query on one axis.  c is clamped so that far-away
// queries cannot overflow; f in [0,1) when unclamped.
__device__ __forceinline__ void query_anchor(double q, double minc, double edge, int &c, double &f)
{
    const double u = __ddiv_rn(__dsub_rn(q, minc), edge);
    double cf = floor(u);
    cf = fmin(fmax(cf, -1.0e9), 1.0e9);
    c = (int)cf;
    f = u - cf;
}
#endif

}  // namespace nbr
