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
// eigenvalues of a symmetric 3x3 with unit trace, descending.  trigonometric closed form in float64.
__device__ __forceinline__ void eig3_unit_trace(const double a[6], double l[3])
{
    const double q = 1.0 / 3.0;
    const double p1 = a[1] * a[1] + a[2] * a[2] + a[4] * a[4];
    const double d0 = a[0] - q, d1 = a[3] - q, d2 = a[5] - q;
    const double p2 = d0 * d0 + d1 * d1 + d2 * d2 + 2.0 * p1;
    if (!(p2 > 1e-30)) { l[0] = l[1] = l[2] = q; return; }
    const double p = sqrt(p2 / 6.0);
    const double ip = 1.0 / p;
    const double b0 = d0 * ip, b1 = a[1] * ip, b2 = a[2] * ip, b3 = d1 * ip, b4 = a[4] * ip, b5 = d2 * ip;
    double r = 0.5 * (b0 * (b3 * b5 - b4 * b4) - b1 * (b1 * b5 - b4 * b2) + b2 * (b1 * b4 - b3 * b2));
    r = fmin(1.0, fmax(-1.0, r));
    const double phi = acos(r) * (1.0 / 3.0);
    const double e0 = q + 2.0 * p * cos(phi);
    const double e2 = q + 2.0 * p * cos(phi + 2.0943951023931954923);
    l[0] = e0;
    l[2] = e2;
    l[1] = 1.0 - e0 - e2;
}

// unit eigenvector of the symmetric matrix `a` for eigenvalue lam (best conditioned cross product)
__device__ __forceinline__ void eigvec3(const double a[6], double lam, double v[3])
{
    const double r0[3] = {a[0] - lam, a[1], a[2]};
    const double r1[3] = {a[1], a[3] - lam, a[4]};
    const double r2[3] = {a[2], a[4], a[5] - lam};
    double c[3][3] = {
        {r0[1] * r1[2] - r0[2] * r1[1], r0[2] * r1[0] - r0[0] * r1[2], r0[0] * r1[1] - r0[1] * r1[0]},
        {r0[1] * r2[2] - r0[2] * r2[1], r0[2] * r2[0] - r0[0] * r2[2], r0[0] * r2[1] - r0[1] * r2[0]},
        {r1[1] * r2[2] - r1[2] * r2[1], r1[2] * r2[0] - r1[0] * r2[2], r1[0] * r2[1] - r1[1] * r2[0]}};
    int best = 0;
    double bn = -1.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        double n2 = c[i][0] * c[i][0] + c[i][1] * c[i][1] + c[i][2] * c[i][2];
        if (n2 > bn) { bn = n2; best = i; }
    }
    if (!(bn > 0.0)) { v[0] = 0; v[1] = 0; v[2] = 1; return; }
    const double s = rsqrt(bn);
    v[0] = c[best][0] * s; v[1] = c[best][1] * s; v[2] = c[best][2] * s;
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
            double l[3];
            eig3_unit_trace(a, l);
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
                double v[3];
                eigvec3(a, l[2], v);
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
