/*
 * mofem_oracle.c -- CPU restatement of MOFEM's stiffness-assembly hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for the CUDA path
 * in mofem_b200/.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference leg may build, load or call it.  The product
 * (mofem_b200/) never links or imports anything under oracle/.
 *
 * Parity is PINNED: tests/golden/ holds outputs of the unmodified reference
 * (junbinhuang/MOFEM, imported from /root/reference by tests/golden/make_golden.py)
 * for simpleOFE3, AMOREBracket and small synthetic overlays; tests/test_oracle.py
 * checks every function here against them.
 *
 * Every function restates one reference function and cites it (paths relative
 * to the reference root).  The reference's third-party arithmetic is restated
 * from its published algorithms:
 *   numpy.linalg.solve -> LAPACK dgesv (partial-pivot LU; OpenBLAS 0.3.30,
 *       numpy 2.3.5 in the survey container -- the reference pins no version).
 *       For 2x2 systems the exact operation order of that build was recovered
 *       by bit-comparison (see lu2_*): l=a10*(1/a00); u11=a11-l*a01 (unfused);
 *       y1=fma(-l,b0,b1); x0=fma(-a01,x1,b0); back-substitution divides for one
 *       right-hand side and multiplies by the reciprocal for several.
 *   numpy.linalg.det   -> dgetrf, then sign*exp(sum(log|u_ii|)) (numpy's
 *       umath_linalg slogdet path).
 *   scipy.sparse.coo_matrix(...).tocsr() -> coo_tocsr, csr_sort_indices,
 *       csr_sum_duplicates (scipy/sparse/sparsetools): rows bucketed in COO
 *       order, columns sorted per row, duplicates summed, explicit zeros kept.
 *
 * Plain C99, no dependencies beyond libm and pthreads (threads are used only
 * to time the CPU baseline on all host cores).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "mofem_oracle.h"

/* ------------------------------------------------------------------ */
/* numpy.linalg restatements                                           */
/* ------------------------------------------------------------------ */

typedef struct { double a00, a01, l, u11; int swapped; } lu2_t;

/* dgetrf on a 2x2 row-major matrix (partial pivoting on column 0). */
static lu2_t lu2_factor(const double *A)
{
    lu2_t f;
    double a00 = A[0], a01 = A[1], a10 = A[2], a11 = A[3];
    f.swapped = fabs(a10) > fabs(a00);
    if (f.swapped) { double t; t = a00; a00 = a10; a10 = t; t = a01; a01 = a11; a11 = t; }
    f.a00 = a00; f.a01 = a01;
    f.l = a10 * (1.0 / a00);
    f.u11 = a11 - f.l * a01;
    return f;
}

/* numpy.linalg.solve(A, b) for one right-hand side (b is a vector). */
static void np_solve2_vec(const double *A, const double *b, double *x)
{
    lu2_t f = lu2_factor(A);
    double b0 = b[0], b1 = b[1];
    if (f.swapped) { double t = b0; b0 = b1; b1 = t; }
    double y1 = fma(-f.l, b0, b1);
    double x1 = y1 / f.u11;
    double x0 = fma(-f.a01, x1, b0) / f.a00;
    x[0] = x0; x[1] = x1;
}

/* numpy.linalg.solve(A, B) with B 2xk row-major, k >= 2. */
static void np_solve2_mat(const double *A, const double *B, int k, double *X)
{
    lu2_t f = lu2_factor(A);
    double ru = 1.0 / f.u11, ra = 1.0 / f.a00;
    for (int j = 0; j < k; j++) {
        double b0 = B[j], b1 = B[k + j];
        if (f.swapped) { double t = b0; b0 = b1; b1 = t; }
        double y1 = fma(-f.l, b0, b1);
        double x1 = y1 * ru;
        double x0 = fma(-f.a01, x1, b0) * ra;
        X[j] = x0; X[k + j] = x1;
    }
}

/* numpy.linalg.det of a 2x2 matrix. */
static double np_det2(const double *A)
{
    lu2_t f = lu2_factor(A);
    double sign = f.swapped ? -1.0 : 1.0;
    double logdet = 0.0;
    double d[2] = { f.a00, f.u11 };
    for (int i = 0; i < 2; i++) {
        double a = d[i];
        if (a < 0.0) { sign = -sign; a = -a; }
        logdet += log(a);
    }
    return sign * exp(logdet);
}

/* numpy.linalg.solve(A, B): A nxn, B nxk row-major, general small n (ICM 4x4). */
static void np_solve_n(const double *Ain, int n, const double *Bin, int k, double *X)
{
    double A[16 * 16];
    for (int i = 0; i < n * n; i++) A[i] = Ain[i];
    for (int i = 0; i < n * k; i++) X[i] = Bin[i];
    for (int c = 0; c < n; c++) {
        int p = c; double best = fabs(A[c * n + c]);
        for (int r = c + 1; r < n; r++) if (fabs(A[r * n + c]) > best) { best = fabs(A[r * n + c]); p = r; }
        if (p != c) {
            for (int j = 0; j < n; j++) { double t = A[c * n + j]; A[c * n + j] = A[p * n + j]; A[p * n + j] = t; }
            for (int j = 0; j < k; j++) { double t = X[c * k + j]; X[c * k + j] = X[p * k + j]; X[p * k + j] = t; }
        }
        double rp = 1.0 / A[c * n + c];
        for (int r = c + 1; r < n; r++) {
            double l = A[r * n + c] * rp;
            A[r * n + c] = l;
            for (int j = c + 1; j < n; j++) A[r * n + j] -= l * A[c * n + j];
        }
    }
    for (int j = 0; j < k; j++) {
        for (int r = 1; r < n; r++) {
            double s = X[r * k + j];
            for (int c = 0; c < r; c++) s -= A[r * n + c] * X[c * k + j];
            X[r * k + j] = s;
        }
        for (int r = n - 1; r >= 0; r--) {
            double s = X[r * k + j];
            for (int c = r + 1; c < n; c++) s -= A[r * n + c] * X[c * k + j];
            X[r * k + j] = s / A[r * n + r];
        }
    }
}

/* ------------------------------------------------------------------ */
/* numpy.matmul rounding orders (OpenBLAS 0.3.30 kernels, recovered by  */
/* bit-comparison against numpy 2.3.5; see DESIGN.md "arithmetic pin") */
/* ------------------------------------------------------------------ */

/* dgemm inner product, k-loop forward: (m,n)@(n,p) with m>=2, p>=2. */
static double dot_gemm(const double *a, int sa, const double *b, int sb, int n)
{
    double acc = a[0] * b[0];
    for (int k = 1; k < n; k++) acc = fma(a[k * sa], b[k * sb], acc);
    return acc;
}

/* (1,n)@(n,p): dgemv_n on the transposed operand; 4-column blocks as
 * (y + fma(a0,b0,a1*b1)) + fma(a2,b2,a3*b3), leftover columns as fma(a,b,y). */
static double dot_vecmat(const double *a, const double *b, int sb, int n)
{
    double y = 0.0;
    int k = 0;
    for (; k + 4 <= n; k += 4) {
        double f01 = fma(a[k], b[k * sb], a[k + 1] * b[(k + 1) * sb]);
        double f23 = fma(a[k + 2], b[(k + 2) * sb], a[k + 3] * b[(k + 3) * sb]);
        y = (k == 0) ? f01 + f23 : (y + f01) + f23;
    }
    for (; k < n; k++) y = (k == 0) ? a[0] * b[0] : fma(a[k], b[k * sb], y);
    return y;
}

/* (2,3)@(3,) and (2,6)@(6,): dgemv_t (dot-product kernel with SIMD lanes). */
static double dot_gemv3(const double *a, const double *b)
{
    return fma(a[2], b[2], fma(a[0], b[0], a[1] * b[1]));
}
static double dot_gemv6(const double *a, const double *b)
{
    return ((a[0] * b[0] + a[2] * b[2]) + (a[1] * b[1] + a[3] * b[3])) + fma(a[4], b[4], a[5] * b[5]);
}

/* ------------------------------------------------------------------ */
/* otherFunctions/numericalIntegration.py                              */
/* ------------------------------------------------------------------ */

#define T13 (1.0 / 3.0)
/* gaussTri: otherFunctions/numericalIntegration.py:5-114 (15-digit literals kept). */
int orc_gauss_tri(int n, double *pos, double *wei)
{
    static const double p1[] = { T13, T13, T13 };
    static const double w1[] = { 1.0 };
    static const double p3[] = { 0.5, 0.5, 0.0, 0.5, 0.0, 0.5, 0.0, 0.5, 0.5 };
    static const double w3[] = { T13, T13, T13 };
    static const double p4[] = { T13, T13, T13, 3.0 / 5, 1.0 / 5, 1.0 / 5, 1.0 / 5, 3.0 / 5, 1.0 / 5, 1.0 / 5, 1.0 / 5, 3.0 / 5 };
    static const double w4[] = { -0.562500000000000, 0.520833333333333, 0.520833333333333, 0.520833333333333 };
    static const double p6[] = {
        0.816847572980459, 0.091576213509771, 0.091576213509771,
        0.091576213509771, 0.816847572980459, 0.091576213509771,
        0.091576213509771, 0.091576213509771, 0.816847572980459,
        0.108103018168070, 0.445948490915965, 0.445948490915965,
        0.445948490915965, 0.108103018168070, 0.445948490915965,
        0.445948490915965, 0.445948490915965, 0.108103018168070 };
    static const double w6[] = { 0.109951743655322, 0.109951743655322, 0.109951743655322,
        0.223381589678011, 0.223381589678011, 0.223381589678011 };
    static const double p7[] = { T13, T13, T13,
        0.797426985353087, 0.101286507323457, 0.101286507323457,
        0.101286507323457, 0.797426985353087, 0.101286507323457,
        0.101286507323457, 0.101286507323457, 0.797426985353087,
        0.059715871789770, 0.470142064105115, 0.470142064105115,
        0.470142064105115, 0.059715871789770, 0.470142064105115,
        0.470142064105115, 0.470142064105115, 0.059715871789770 };
    static const double w7[] = { 0.225000000000000,
        0.125939180544827, 0.125939180544827, 0.125939180544827,
        0.132394152788506, 0.132394152788506, 0.132394152788506 };
    static const double p9[] = {
        0.124949503233232, 0.437525248383384, 0.437525248383384,
        0.437525248383384, 0.124949503233232, 0.437525248383384,
        0.437525248383384, 0.437525248383384, 0.124949503233232,
        0.797112651860071, 0.165409927389841, 0.037477420750088,
        0.165409927389841, 0.797112651860071, 0.037477420750088,
        0.037477420750088, 0.797112651860071, 0.165409927389841,
        0.037477420750088, 0.165409927389841, 0.797112651860071,
        0.797112651860071, 0.037477420750088, 0.165409927389841,
        0.165409927389841, 0.037477420750088, 0.797112651860071 };
    static const double w9[] = { 0.205950504760887, 0.205950504760887, 0.205950504760887,
        0.063691414286223, 0.063691414286223, 0.063691414286223,
        0.063691414286223, 0.063691414286223, 0.063691414286223 };
    static const double p12[] = {
        0.873821971016996, 0.063089014491502, 0.063089014491502,
        0.063089014491502, 0.873821971016996, 0.063089014491502,
        0.063089014491502, 0.063089014491502, 0.873821971016996,
        0.501426509658179, 0.249286745170910, 0.249286745170910,
        0.249286745170910, 0.501426509658179, 0.249286745170910,
        0.249286745170910, 0.249286745170910, 0.501426509658179,
        0.636502499121399, 0.310352451033785, 0.053145049844816,
        0.310352451033785, 0.636502499121399, 0.053145049844816,
        0.053145049844816, 0.636502499121399, 0.310352451033785,
        0.053145049844816, 0.310352451033785, 0.636502499121399,
        0.636502499121399, 0.053145049844816, 0.310352451033785,
        0.310352451033785, 0.053145049844816, 0.636502499121399 };
    static const double w12[] = { 0.050844906370207, 0.050844906370207, 0.050844906370207,
        0.116786275726379, 0.116786275726379, 0.116786275726379,
        0.082851075618374, 0.082851075618374, 0.082851075618374,
        0.082851075618374, 0.082851075618374, 0.082851075618374 };
    static const double p13[] = { T13, T13, T13,
        0.479308067841920, 0.260345966079040, 0.260345966079040,
        0.260345966079040, 0.479308067841920, 0.260345966079040,
        0.260345966079040, 0.260345966079040, 0.479308067841920,
        0.869739794195568, 0.065130102902216, 0.065130102902216,
        0.065130102902216, 0.869739794195568, 0.065130102902216,
        0.065130102902216, 0.065130102902216, 0.869739794195568,
        0.638444188569810, 0.312865496004874, 0.048690315425316,
        0.638444188569810, 0.048690315425316, 0.312865496004874,
        0.312865496004874, 0.638444188569810, 0.048690315425316,
        0.312865496004874, 0.048690315425316, 0.638444188569810,
        0.048690315425316, 0.638444188569810, 0.312865496004874,
        0.048690315425316, 0.312865496004874, 0.638444188569810 };
    static const double w13[] = { -0.149570044467682,
        0.175615257433208, 0.175615257433208, 0.175615257433208,
        0.053347235608838, 0.053347235608838, 0.053347235608838,
        0.077113760890257, 0.077113760890257, 0.077113760890257,
        0.077113760890257, 0.077113760890257, 0.077113760890257 };
    static const double p16[] = { T13, T13, T13,
        0.081414823414554, 0.459292588292723, 0.459292588292723,
        0.459292588292723, 0.081414823414554, 0.459292588292723,
        0.459292588292723, 0.459292588292723, 0.081414823414554,
        0.658861384496480, 0.170569307751760, 0.170569307751760,
        0.170569307751760, 0.658861384496480, 0.170569307751760,
        0.170569307751760, 0.170569307751760, 0.658861384496480,
        0.898905543365938, 0.050547228317031, 0.050547228317031,
        0.050547228317031, 0.898905543365938, 0.050547228317031,
        0.050547228317031, 0.050547228317031, 0.898905543365938,
        0.008394777409958, 0.263112829634638, 0.728492392955404,
        0.008394777409958, 0.728492392955404, 0.263112829634638,
        0.263112829634638, 0.008394777409958, 0.728492392955404,
        0.263112829634638, 0.728492392955404, 0.008394777409958,
        0.728492392955404, 0.008394777409958, 0.263112829634638,
        0.728492392955404, 0.263112829634638, 0.008394777409958 };
    static const double w16[] = { 0.144315607677787,
        0.095091634267285, 0.095091634267285, 0.095091634267285,
        0.103217370534718, 0.103217370534718, 0.103217370534718,
        0.032458497623198, 0.032458497623198, 0.032458497623198,
        0.027230314174435, 0.027230314174435, 0.027230314174435,
        0.027230314174435, 0.027230314174435, 0.027230314174435 };
    const double *p, *w;
    switch (n) {
    case 1: p = p1; w = w1; break;
    case 3: p = p3; w = w3; break;
    case 4: p = p4; w = w4; break;
    case 6: p = p6; w = w6; break;
    case 7: p = p7; w = w7; break;
    case 9: p = p9; w = w9; break;
    case 12: p = p12; w = w12; break;
    case 13: p = p13; w = w13; break;
    case 16: p = p16; w = w16; break;
    default: return -1; /* ValueError("Wrong number of quadrature points!") */
    }
    memcpy(pos, p, sizeof(double) * 3 * n);
    memcpy(wei, w, sizeof(double) * n);
    return 0;
}

/* gaussQuad: otherFunctions/numericalIntegration.py:116-134. */
int orc_gauss_quad(int n, double *pos, double *wei)
{
    switch (n) {
    case 1: pos[0] = 0.0; wei[0] = 2.0; return 0;
    case 2: pos[0] = -0.577350269189626; pos[1] = 0.577350269189626; wei[0] = wei[1] = 1.0; return 0;
    case 3: pos[0] = -0.774596669241483; pos[1] = 0.0; pos[2] = 0.774596669241483;
        wei[0] = 5.0 / 9; wei[1] = 8.0 / 9; wei[2] = 5.0 / 9; return 0;
    case 4: pos[0] = -0.861136311594053; pos[1] = -0.339981043584856; pos[2] = 0.339981043584856; pos[3] = 0.861136311594053;
        wei[0] = 0.347854845137454; wei[1] = 0.652145154862546; wei[2] = 0.652145154862546; wei[3] = 0.347854845137454; return 0;
    case 5: pos[0] = -0.906179845938664; pos[1] = -0.538469310105683; pos[2] = 0.0; pos[3] = 0.538469310105683; pos[4] = 0.906179845938664;
        wei[0] = 0.236926885056189; wei[1] = 0.478628670499366; wei[2] = 128.0 / 225; wei[3] = 0.478628670499366; wei[4] = 0.236926885056189; return 0;
    default: return -1;
    }
}

/* ------------------------------------------------------------------ */
/* elementLibrary/shapeFunction.py + invShapeFunction.py               */
/* ------------------------------------------------------------------ */

/* N (n values) and dN (2 x n) of the 4-node quad: shapeFunction.py:34-40,
 * invShapeFunction.py:143-158. */
static void quad_N_dN(double r, double s, double *N, double *dN)
{
    N[0] = 0.25 * (1.0 - r) * (1.0 - s);
    N[1] = 0.25 * (1.0 + r) * (1.0 - s);
    N[2] = 0.25 * (1.0 + r) * (1.0 + s);
    N[3] = 0.25 * (1.0 - r) * (1.0 + s);
    dN[0] = -0.25 * (1.0 - s); dN[1] = 0.25 * (1.0 - s); dN[2] = 0.25 * (1.0 + s); dN[3] = -0.25 * (1.0 + s);
    dN[4] = -0.25 * (1.0 - r); dN[5] = -0.25 * (1.0 + r); dN[6] = 0.25 * (1.0 + r); dN[7] = 0.25 * (1.0 - r);
}

/* 9-node quad: shapeFunction.py:95-131, invShapeFunction.py:160-196.
 * r**2 in the reference is r*r. */
static void quadquad_N_dN(double r, double s, double *N, double *dN)
{
    double r2 = r * r, s2 = s * s;
    N[0] = 0.25 * (1.0 - r) * (1.0 - s) + 0.25 * (1.0 - r2) * (1.0 - s2) - 0.25 * (1.0 - s) * (1.0 - r2) - 0.25 * (1.0 - r) * (1.0 - s2);
    N[1] = 0.25 * (1.0 + r) * (1.0 - s) + 0.25 * (1.0 - r2) * (1.0 - s2) - 0.25 * (1.0 - s) * (1.0 - r2) - 0.25 * (1.0 + r) * (1.0 - s2);
    N[2] = 0.25 * (1.0 + r) * (1.0 + s) + 0.25 * (1.0 - r2) * (1.0 - s2) - 0.25 * (1.0 + r) * (1.0 - s2) - 0.25 * (1.0 + s) * (1.0 - r2);
    N[3] = 0.25 * (1.0 - r) * (1.0 + s) + 0.25 * (1.0 - r2) * (1.0 - s2) - 0.25 * (1.0 + s) * (1.0 - r2) - 0.25 * (1.0 - r) * (1.0 - s2);
    N[4] = 0.5 * (1.0 - s) * (1.0 - r2) - 0.5 * (1.0 - r2) * (1.0 - s2);
    N[5] = 0.5 * (1.0 + r) * (1.0 - s2) - 0.5 * (1.0 - r2) * (1.0 - s2);
    N[6] = 0.5 * (1.0 + s) * (1.0 - r2) - 0.5 * (1.0 - r2) * (1.0 - s2);
    N[7] = 0.5 * (1.0 - r) * (1.0 - s2) - 0.5 * (1.0 - r2) * (1.0 - s2);
    N[8] = (1.0 - r2) * (1.0 - s2);
    dN[0] = -0.25 * (1.0 - s) - 0.5 * r * (1.0 - s2) + 0.5 * r * (1.0 - s) + 0.25 * (1.0 - s2);
    dN[1] = 0.25 * (1.0 - s) - 0.5 * r * (1.0 - s2) + 0.5 * r * (1.0 - s) - 0.25 * (1.0 - s2);
    dN[2] = 0.25 * (1.0 + s) - 0.5 * r * (1.0 - s2) - 0.25 * (1.0 - s2) + 0.5 * r * (1.0 + s);
    dN[3] = -0.25 * (1.0 + s) - 0.5 * r * (1.0 - s2) + 0.5 * r * (1.0 + s) + 0.25 * (1.0 - s2);
    dN[4] = -r * (1.0 - s) + r * (1.0 - s2);
    dN[5] = 0.5 * (1.0 - s2) + r * (1.0 - s2);
    dN[6] = -r * (1.0 + s) + r * (1.0 - s2);
    dN[7] = -0.5 * (1.0 - s2) + r * (1.0 - s2);
    dN[8] = -2.0 * r * (1.0 - s2);
    dN[9] = -0.25 * (1.0 - r) - 0.5 * s * (1.0 - r2) + 0.25 * (1.0 - r2) + 0.5 * s * (1.0 - r);
    dN[10] = -0.25 * (1.0 + r) - 0.5 * s * (1.0 - r2) + 0.25 * (1.0 - r2) + 0.5 * s * (1.0 + r);
    dN[11] = 0.25 * (1.0 + r) - 0.5 * s * (1.0 - r2) + 0.5 * s * (1.0 + r) - 0.25 * (1.0 - r2);
    dN[12] = 0.25 * (1.0 - r) - 0.5 * s * (1.0 - r2) - 0.25 * (1.0 - r2) + 0.5 * s * (1.0 - r);
    dN[13] = -0.5 * (1.0 - r2) + s * (1.0 - r2);
    dN[14] = -s * (1.0 + r) + s * (1.0 - r2);
    dN[15] = 0.5 * (1.0 - r2) + s * (1.0 - r2);
    dN[16] = -s * (1.0 - r) + s * (1.0 - r2);
    dN[17] = -2.0 * s * (1.0 - r2);
}

/* 3-node triangle: shapeFunction.py:189-193. */
static void tri_N_dN(double r, double s, double *N, double *dN)
{
    N[0] = r; N[1] = s; N[2] = 1.0 - r - s;
    dN[0] = 1.0; dN[1] = 0.0; dN[2] = -1.0;
    dN[3] = 0.0; dN[4] = 1.0; dN[5] = -1.0;
}

/* 6-node triangle: shapeFunction.py:223-241, invShapeFunction.py:210-231. */
static void triquad_N_dN(double r, double s, double *N, double *dN)
{
    double t = 1.0 - r - s;
    N[0] = r - 2.0 * r * s - 2.0 * r * t;
    N[1] = s - 2.0 * s * t - 2.0 * r * s;
    N[2] = t - 2.0 * s * t - 2.0 * r * t;
    N[3] = 4.0 * r * s; N[4] = 4.0 * s * t; N[5] = 4.0 * r * t;
    dN[0] = 1.0 - 2.0 * s - 2.0 * t + 2.0 * r;
    dN[1] = 0.0;
    dN[2] = -1.0 + 2.0 * s - 2.0 * t + 2.0 * r;
    dN[3] = 4.0 * s; dN[4] = -4.0 * s; dN[5] = 4.0 * t - 4.0 * r;
    dN[6] = 0.0;
    dN[7] = 1.0 - 2.0 * t + 2.0 * s - 2.0 * r;
    dN[8] = -1.0 - 2.0 * t + 2.0 * s + 2.0 * r;
    dN[9] = 4.0 * r; dN[10] = 4.0 * t - 4.0 * s; dN[11] = -4.0 * r;
}

static void N_dN(int n, double r, double s, double *N, double *dN)
{
    if (n == 4) quad_N_dN(r, s, N, dN);
    else if (n == 9) quadquad_N_dN(r, s, N, dN);
    else if (n == 3) tri_N_dN(r, s, N, dN);
    else triquad_N_dN(r, s, N, dN);
}

/* J = dN @ coord (2x2), as np.matmul(dNmat,coord). */
static void jac(const double *dN, const double *coord, int n, double *J)
{
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++) {
            J[2 * i + j] = dot_gemm(dN + i * n, 1, coord + j, 2, n);
        }
}

/* Common body of shapeQuadFE:31, shapeQuadQuadFE:92, shapeTriFE:186,
 * shapeTriQuadFE:221: N (n), dudx (2 x n) = solve(J,dN), Jacobian = det(J). */
static void shape_generic(const double *coord, int n, double r, double s,
                          double *N, double *dudx, double *Jacobian)
{
    double dN[18], J[4];
    N_dN(n, r, s, N, dN);
    jac(dN, coord, n, J);
    *Jacobian = np_det2(J);
    np_solve2_mat(J, dN, n, dudx);
}

/* Nmat (2 x 2n) and Bmat (3 x 2n) exactly as the reference lays them out. */
void orc_shape(const double *coord, int n, double r, double s,
               double *Nmat, double *Bmat, double *Jacobian)
{
    double N[9], dudx[18];
    shape_generic(coord, n, r, s, N, dudx, Jacobian);
    int w = 2 * n;
    memset(Nmat, 0, sizeof(double) * 2 * w);
    memset(Bmat, 0, sizeof(double) * 3 * w);
    for (int a = 0; a < n; a++) {
        Nmat[2 * a] = N[a];
        Nmat[w + 2 * a + 1] = N[a];
        Bmat[2 * a] = dudx[a];
        Bmat[w + 2 * a + 1] = dudx[n + a];
        Bmat[2 * w + 2 * a] = dudx[n + a];
        Bmat[2 * w + 2 * a + 1] = dudx[a];
    }
}

/* shapeTriFE2:206 / shapeTriQuadFE2:255 : N (1 x nt), dudx (2 x nt), det. */
static void shape_tri2(const double *coordTri, int nt, double r, double s,
                       double *N, double *dudx, double *Jacobian)
{
    shape_generic(coordTri, nt, r, s, N, dudx, Jacobian);
}

/* invTri: invShapeFunction.py:21-39 (returns 3 area coordinates). */
void orc_inv_tri(const double *c, const double *xy, double *iso)
{
    double area = (c[2] * c[5] - c[4] * c[3] - c[0] * (c[5] - c[3]) + c[1] * (c[4] - c[2])) / 2;
    double a[3] = { c[2] * c[5] - c[4] * c[3], c[4] * c[1] - c[0] * c[5], c[0] * c[3] - c[2] * c[1] };
    double b[3] = { c[3] - c[5], c[5] - c[1], c[1] - c[3] };
    double cc[3] = { c[4] - c[2], c[0] - c[4], c[2] - c[0] };
    for (int i = 0; i < 3; i++) iso[i] = (a[i] + b[i] * xy[0] + cc[i] * xy[1]) / 2.0 / area;
}

/* invQuad:41, invQuadQuad:75, invTriQuad:109 -- Newton inverse map.
 * Returns the number of Newton updates, or -1 for "Too many iterations!". */
int orc_inv_newton(const double *coord, int n, const double *xy, double *iso)
{
    double elementSize = fmax(fabs(coord[4] - coord[0]), fabs(coord[5] - coord[1]));
    double tol = elementSize * 1e-15; /* 10**(-15) is the float 1e-15 */
    double cn[18], N[9], dN[18], res[2], J[4], Jt[4], d[2], xn[2];
    if (n == 6) { iso[0] = 1.0 / 3; iso[1] = 1.0 / 3; } else { iso[0] = 0.0; iso[1] = 0.0; }
    xn[0] = xy[0] - coord[0]; xn[1] = xy[1] - coord[1];
    for (int i = 0; i < n; i++) { cn[2 * i] = coord[2 * i] - coord[0]; cn[2 * i + 1] = coord[2 * i + 1] - coord[1]; }
    int count = 0;
    for (;;) {
        N_dN(n, iso[0], iso[1], N, dN);
        for (int j = 0; j < 2; j++) {
            res[j] = xn[j] - dot_vecmat(N, cn + j, 2, n);
        }
        if (!(fmax(fabs(res[0]), fabs(res[1])) > tol)) break;
        jac(dN, cn, n, J);
        Jt[0] = J[0]; Jt[1] = J[2]; Jt[2] = J[1]; Jt[3] = J[3];
        np_solve2_vec(Jt, res, d);
        iso[0] += d[0]; iso[1] += d[1];
        count++;
        if (count >= 20) return -1;
    }
    return count;
}

/* ------------------------------------------------------------------ */
/* elementLibrary/stiffnessMatrix.py                                   */
/* ------------------------------------------------------------------ */

/* out (m x n) = A^T (k x m)^T ... helper: C = A^T @ (D @ B), A: 3 x m, B: 3 x n. */
static void btdb(const double *A, int m, const double *D, const double *B, int n, double *C)
{
    double DB[3 * 18];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < n; j++) {
            DB[i * n + j] = fma(D[3 * i + 2], B[2 * n + j], fma(D[3 * i + 1], B[n + j], D[3 * i] * B[j]));
        }
    for (int i = 0; i < m; i++)
        for (int j = 0; j < n; j++) {
            C[i * n + j] = fma(A[2 * m + i], DB[2 * n + j], fma(A[m + i], DB[n + j], A[i] * DB[j]));
        }
}

/* triFE:9 (nInt=1) and triQuadFE:22 (nInt=3). */
static void tri_element(const double *coord, int n, const double *D, double *K)
{
    int nInt = (n == 3) ? 1 : 3, w = 2 * n;
    double pos[48], wei[16], Nmat[2 * 18], Bmat[3 * 18], C[18 * 18], Jac;
    orc_gauss_tri(nInt, pos, wei);
    memset(K, 0, sizeof(double) * w * w);
    for (int i = 0; i < nInt; i++) {
        orc_shape(coord, n, pos[3 * i], pos[3 * i + 1], Nmat, Bmat, &Jac);
        btdb(Bmat, w, D, Bmat, w, C);
        for (int e = 0; e < w * w; e++) K[e] += 0.5 * C[e] * Jac * wei[i];
    }
}

/* quadFE:35 (2x2 Gauss) and quadQuadFE:69 (3x3 Gauss). */
static void quad_element(const double *coord, int n, const double *D, double *K)
{
    int nInt = (n == 4) ? 2 : 3, w = 2 * n;
    double pos[5], wei[5], Nmat[2 * 18], Bmat[3 * 18], C[18 * 18], Jac;
    orc_gauss_quad(nInt, pos, wei);
    memset(K, 0, sizeof(double) * w * w);
    for (int i = 0; i < nInt; i++)
        for (int j = 0; j < nInt; j++) {
            orc_shape(coord, n, pos[i], pos[j], Nmat, Bmat, &Jac);
            btdb(Bmat, w, D, Bmat, w, C);
            for (int e = 0; e < w * w; e++) K[e] += C[e] * Jac * wei[i] * wei[j];
        }
}

/* ICMFE: stiffnessMatrix.py:49-67 with shapeICMFE: shapeFunction.py:53-73. */
static void icm_element(const double *coord, const double *D, double *K)
{
    double pos[5], wei[5];
    double H[16], E[8 * 4], N[4], dN[8], J[4], dudx[8], dGdx[4];
    double Bmat[3 * 8], Gmat[3 * 4], C[64], CH[16], CE[32];
    static const double dGcoef[2] = { -2.0, -2.0 };
    (void)dGcoef;
    orc_gauss_quad(2, pos, wei);
    memset(K, 0, sizeof(double) * 64); memset(H, 0, sizeof H); memset(E, 0, sizeof E);
    for (int i = 0; i < 2; i++)
        for (int j = 0; j < 2; j++) {
            double r = pos[i], s = pos[j];
            quad_N_dN(r, s, N, dN);
            double dG[4] = { -2.0 * r, 0.0, 0.0, -2.0 * s };
            jac(dN, coord, 4, J);
            double Jac = np_det2(J);
            np_solve2_mat(J, dN, 4, dudx);
            np_solve2_mat(J, dG, 2, dGdx);
            memset(Bmat, 0, sizeof Bmat); memset(Gmat, 0, sizeof Gmat);
            for (int a = 0; a < 4; a++) {
                Bmat[2 * a] = dudx[a]; Bmat[8 + 2 * a + 1] = dudx[4 + a];
                Bmat[16 + 2 * a] = dudx[4 + a]; Bmat[16 + 2 * a + 1] = dudx[a];
            }
            for (int a = 0; a < 2; a++) {
                Gmat[2 * a] = dGdx[a]; Gmat[4 + 2 * a + 1] = dGdx[2 + a];
                Gmat[8 + 2 * a] = dGdx[2 + a]; Gmat[8 + 2 * a + 1] = dGdx[a];
            }
            btdb(Bmat, 8, D, Bmat, 8, C);
            btdb(Gmat, 4, D, Gmat, 4, CH);
            btdb(Bmat, 8, D, Gmat, 4, CE);
            for (int e = 0; e < 64; e++) K[e] += C[e] * Jac * wei[i] * wei[j];
            for (int e = 0; e < 16; e++) H[e] += CH[e] * Jac * wei[i] * wei[j];
            for (int e = 0; e < 32; e++) E[e] += CE[e] * Jac * wei[i] * wei[j];
        }
    /* Kloc -= Eloc @ solve(Hloc, Eloc.T) */
    double Et[4 * 8], X[4 * 8];
    for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) Et[j * 8 + i] = E[i * 4 + j];
    np_solve_n(H, 4, Et, 8, X);
    for (int i = 0; i < 8; i++)
        for (int j = 0; j < 8; j++) {
            K[i * 8 + j] -= dot_gemm(E + i * 4, 1, X + j, 8, 4);
        }
}

/* Regular-element dispatch used by the solvers (AMORE.py:42-56, 310-327,
 * 733-748; traditionalElement.py:36-48, 188-198, 339-351).
 * quad4_rule: 0 = quadFE, 1 = ICMFE. */
int orc_element_stiffness(const double *coord, int n, const double *D, int quad4_rule, double *K)
{
    if (n == 3 || n == 6) tri_element(coord, n, D, K);
    else if (n == 4 && quad4_rule) icm_element(coord, D, K);
    else if (n == 4 || n == 9) quad_element(coord, n, D, K);
    else return -1;
    return 0;
}

/* getLC: stiffnessMatrix.py:209-222 -> LC (3 x 2). */
static void get_lc(const double *coordTri, int nt, const double *rho, double r, double s, double *LC)
{
    double N[6], dudx[12], Jac, drho[2];
    shape_tri2(coordTri, nt, r, s, N, dudx, &Jac);
    if (nt == 3) {
        /* dudx@rho is a cancellation-prone 3-term sum on sliver triangles (cond(J) up
         * to 1e7 in AMOREBracket), so its rounding order is restated exactly:
         * OpenBLAS dgemv (numpy matmul (2,3)@(3,)) = fma(a2,b2, fma(a0,b0, a1*b1)). */
        for (int i = 0; i < 2; i++) drho[i] = dot_gemv3(dudx + 3 * i, rho);
    } else {
        double re[6] = { rho[0], rho[1], rho[2], 0.5 * (rho[0] + rho[1]), 0.5 * (rho[1] + rho[2]), 0.5 * (rho[2] + rho[0]) };
        for (int i = 0; i < 2; i++) drho[i] = dot_gemv6(dudx + 6 * i, re);
    }
    LC[0] = drho[0]; LC[1] = 0.0; LC[2] = 0.0; LC[3] = drho[1]; LC[4] = drho[1]; LC[5] = drho[0];
}

/* getStrainMatrix: stiffnessMatrix.py:192-207 -> (LC@Nmat)+(rho*Bmat), 3 x 2n.
 * Returns -1 when the Newton inverse map asserts. */
static int get_strain_matrix(const double *LC, const double *coord, int n, const double *xy,
                             double rho, double *Bt, int *newton)
{
    double iso[3], Nmat[2 * 18], Bmat[3 * 18], Jac;
    if (n == 3) orc_inv_tri(coord, xy, iso);
    else {
        int it = orc_inv_newton(coord, n, xy, iso);
        if (it < 0) return -1;
        if (newton) *newton += it;
    }
    orc_shape(coord, n, iso[0], iso[1], Nmat, Bmat, &Jac);
    int w = 2 * n;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < w; j++)
            Bt[i * w + j] = (LC[2 * i] * Nmat[j] + LC[2 * i + 1] * Nmat[w + j]) + rho * Bmat[i * w + j];
    return 0;
}

/* nInt selection: lowerOrderAMORE stiffnessMatrix.py:104-106,
 * quadraticAMORE stiffnessMatrix.py:137-144. */
int orc_amore_nint(int quadratic, int n1, int n2)
{
    int s = n1 + n2;
    if (!quadratic) return s == 6 ? 3 : (s == 7 ? 4 : 6);
    switch (s) {
    case 6: return 3; case 7: return 4; case 8: return 6; case 9: return 4;
    case 10: return 6; case 12: return 9; case 13: return 12; default: return 16;
    }
}

/* lowerOrderAMORE: stiffnessMatrix.py:99-129 and quadraticAMORE:131-190.
 * coordTri: nt x 2 (nt = 3 straight, 6 curved; curved only when quadratic).
 * coord2 == NULL means the diagonal call (Bmat2 is Bmat1).
 * Kloc (2 n1 x 2 n2) is OVERWRITTEN with this triangle's contribution. */
int orc_amore_block(int quadratic, const double *coordTri, int nt, const double *D,
                    const double *coord1, int n1, const double *rho1,
                    const double *coord2, int n2, const double *rho2, double *Kloc, int *newton)
{
    int diag = (coord2 == NULL);
    if (diag) { coord2 = coord1; n2 = n1; rho2 = rho1; }
    int nInt = orc_amore_nint(quadratic, n1, n2);
    double pos[48], wei[16], LC1[6], LC2[6], N[6], dudx[12], Jac, xy[2];
    double B1[3 * 18], B2[3 * 18], C[18 * 18];
    int w1 = 2 * n1, w2 = 2 * n2;
    orc_gauss_tri(nInt, pos, wei);
    memset(Kloc, 0, sizeof(double) * w1 * w2);
    if (nt == 3) {
        get_lc(coordTri, 3, rho1, 0.0, 0.0, LC1);
        if (!diag) get_lc(coordTri, 3, rho2, 0.0, 0.0, LC2);
    }
    for (int i = 0; i < nInt; i++) {
        double r1v, r2v = 0.0;
        if (nt == 6) {
            get_lc(coordTri, 6, rho1, pos[3 * i], pos[3 * i + 1], LC1);
            if (!diag) get_lc(coordTri, 6, rho2, pos[3 * i], pos[3 * i + 1], LC2);
        }
        shape_tri2(coordTri, nt, pos[3 * i], pos[3 * i + 1], N, dudx, &Jac);
        if (nt == 3) {
            r1v = dot_gemm(N, 1, rho1, 1, 3);   /* (1,3)@(3,): forward fma chain */
            if (!diag) r2v = dot_gemm(N, 1, rho2, 1, 3);
        } else {
            r1v = pos[3 * i] * rho1[0] + pos[3 * i + 1] * rho1[1] + pos[3 * i + 2] * rho1[2];
            if (!diag) r2v = pos[3 * i] * rho2[0] + pos[3 * i + 1] * rho2[1] + pos[3 * i + 2] * rho2[2];
        }
        for (int j = 0; j < 2; j++) {
            xy[j] = dot_vecmat(N, coordTri + j, 2, nt);
        }
        if (get_strain_matrix(LC1, coord1, n1, xy, r1v, B1, newton)) return -1;
        const double *B2p = B1;
        if (!diag) {
            if (get_strain_matrix(LC2, coord2, n2, xy, r2v, B2, newton)) return -1;
            B2p = B2;
        }
        btdb(B1, w1, D, B2p, w2, C);
        for (int e = 0; e < w1 * w2; e++) Kloc[e] += 0.5 * C[e] * Jac * wei[i];
    }
    return 0;
}

/* twoDMaterialMatrix: AMORE.py:1128-1146 / traditionalElement.py:496-514. */
int orc_material_matrix(double E, double nu, int problemType, double *D)
{
    memset(D, 0, sizeof(double) * 9);
    if (problemType == 1) {
        D[0] = E / (1.0 - nu * nu); D[1] = E * nu / (1.0 - nu * nu);
        D[3] = E * nu / (1.0 - nu * nu); D[4] = E / (1.0 - nu * nu);
        D[8] = E / 2.0 / (1.0 + nu);
    } else if (problemType == 2) {
        double cc = E * (1.0 - nu) / (1.0 + nu) / (1.0 - 2.0 * nu);
        D[0] = cc; D[1] = cc * nu / (1.0 - nu);
        D[3] = cc * nu / (1.0 - nu); D[4] = cc;
        D[8] = cc * (1.0 - 2.0 * nu) / (2.0 * (1.0 - nu));
    } else return -1;
    return 0;
}

/* ------------------------------------------------------------------ */
/* The cell loop: AMORE.py:59-123 / 328-398 / 751-821                  */
/* traditionalElement.py:36-48 / 188-198 / 339-351 (all cells regular) */
/* ------------------------------------------------------------------ */

static void elem_coord(const orc_problem *p, int e, double *coord)
{
    int n = p->elem_nn[e];
    for (int a = 0; a < n; a++) {
        int nd = p->elem_nodes[9 * (size_t)e + a];
        coord[2 * a] = p->node_xy[2 * (size_t)nd];
        coord[2 * a + 1] = p->node_xy[2 * (size_t)nd + 1];
    }
}

/* number of COO triplets cell c emits */
static int64_t cell_coo_count(const orc_problem *p, int64_t c)
{
    const int32_t *ce = p->cell_elem + (size_t)c * p->n_mesh;
    int64_t cnt = 0;
    if (p->cell_kind[c] == 0) {
        for (int j = 0; j < p->n_mesh; j++) if (ce[j] >= 0) { int n = p->elem_nn[ce[j]]; return 4 * (int64_t)n * n; }
        return 0;
    }
    for (int j = 0; j < p->n_mesh; j++) {
        if (ce[j] < 0) continue;
        int n1 = p->elem_nn[ce[j]];
        cnt += 4 * (int64_t)n1 * n1;
        for (int l = j + 1; l < p->n_mesh; l++) if (ce[l] >= 0) cnt += 8 * (int64_t)n1 * p->elem_nn[ce[l]];
    }
    return cnt;
}

int64_t orc_coo_count(const orc_problem *p)
{
    int64_t t = 0;
    for (int64_t c = 0; c < p->n_cell; c++) t += cell_coo_count(p, c);
    return t;
}

/* sparsifyElementMatrix: stiffnessMatrix.py:83-97. */
static int64_t sparsify(const double *K, const int32_t *num1, int n1, const int32_t *num2, int n2,
                        int transpose, int64_t *I, int64_t *J, double *V)
{
    int64_t k = 0;
    if (!transpose) {
        for (int i = 0; i < 2 * n1; i++)
            for (int j = 0; j < 2 * n2; j++) {
                I[k] = 2 * (int64_t)num1[i / 2] + i % 2; J[k] = 2 * (int64_t)num2[j / 2] + j % 2;
                V[k] = K[i * 2 * n2 + j]; k++;
            }
    } else { /* sparsifyElementMatrix(Kloc.transpose(), numbering2, numbering) */
        for (int i = 0; i < 2 * n2; i++)
            for (int j = 0; j < 2 * n1; j++) {
                I[k] = 2 * (int64_t)num2[i / 2] + i % 2; J[k] = 2 * (int64_t)num1[j / 2] + j % 2;
                V[k] = K[j * 2 * n2 + i]; k++;
            }
    }
    return k;
}

static int tri_coord(const orc_problem *p, int64_t t, double *ct)
{
    for (int k = 0; k < 6; k++) ct[k] = p->tri_xy[6 * (size_t)t + k];
    if (p->tri_curved && p->tri_curved[t]) {
        for (int k = 0; k < 6; k++) ct[6 + k] = p->tri_mid[6 * (size_t)t + k];
        return 6;
    }
    return 3;
}

static int cell_emit(const orc_problem *p, int64_t c, int64_t *I, int64_t *J, double *V, int64_t *newton)
{
    const int M = p->n_mesh;
    const int32_t *ce = p->cell_elem + (size_t)c * M;
    const int quadratic = p->solver != 1;
    const int quad4_rule = p->solver == 3;
    double coord[18], coord2[18], K[18 * 18], Kt[18 * 18], ct[12], rho[3], rho2[3];
    int64_t k = 0;
    int nw = 0;
    if (p->cell_kind[c] == 0) {
        for (int j = 0; j < M; j++) if (ce[j] >= 0) {
            int e = ce[j], n = p->elem_nn[e];
            if (p->solver == 1 && n != 3 && n != 4) return -2; /* "Wrong element numbering!" */
            elem_coord(p, e, coord);
            if (orc_element_stiffness(coord, n, p->mat_D + 9 * (size_t)p->elem_mat[e], quad4_rule, K)) return -2;
            sparsify(K, p->elem_nodes + 9 * (size_t)e, n, p->elem_nodes + 9 * (size_t)e, n, 0, I, J, V);
            return 0;
        }
        return 0;
    }
    for (int j = 0; j < M; j++) {
        if (ce[j] < 0) continue;
        int e1 = ce[j], n1 = p->elem_nn[e1];
        const double *D = p->mat_D + 9 * (size_t)p->elem_mat[e1];
        const int32_t *num1 = p->elem_nodes + 9 * (size_t)e1;
        elem_coord(p, e1, coord);
        memset(K, 0, sizeof(double) * 4 * n1 * n1);
        for (int64_t t = p->cell_tri_ptr[c]; t < p->cell_tri_ptr[c + 1]; t++) {
            int nt = tri_coord(p, t, ct);
            for (int v = 0; v < 3; v++) rho[v] = p->tri_rho[(3 * (size_t)t + v) * M + j];
            if (orc_amore_block(quadratic, ct, nt, D, coord, n1, rho, NULL, 0, NULL, Kt, &nw)) return -1;
            for (int e = 0; e < 4 * n1 * n1; e++) K[e] += Kt[e];
        }
        k += sparsify(K, num1, n1, num1, n1, 0, I + k, J + k, V + k);
        for (int l = j + 1; l < M; l++) {
            if (ce[l] < 0) continue;
            int e2 = ce[l], n2 = p->elem_nn[e2];
            const int32_t *num2 = p->elem_nodes + 9 * (size_t)e2;
            elem_coord(p, e2, coord2);
            memset(K, 0, sizeof(double) * 4 * n1 * n2);
            for (int64_t t = p->cell_tri_ptr[c]; t < p->cell_tri_ptr[c + 1]; t++) {
                int nt = tri_coord(p, t, ct);
                for (int v = 0; v < 3; v++) {
                    rho[v] = p->tri_rho[(3 * (size_t)t + v) * M + j];
                    rho2[v] = p->tri_rho[(3 * (size_t)t + v) * M + l];
                }
                if (orc_amore_block(quadratic, ct, nt, D, coord, n1, rho, coord2, n2, rho2, Kt, &nw)) return -1;
                for (int e = 0; e < 4 * n1 * n2; e++) K[e] += Kt[e];
            }
            k += sparsify(K, num1, n1, num2, n2, 0, I + k, J + k, V + k);
            k += sparsify(K, num1, n1, num2, n2, 1, I + k, J + k, V + k);
        }
    }
    if (newton) *newton = nw;
    return 0;
}

/* Emits the reference's (Iglo,Jglo,Vglo) in its own order.  cell_off (n_cell+1)
 * receives the triplet offset of each cell.  Returns 0, -1 (Newton assert),
 * -2 (wrong element numbering). newton_total (optional) = total Newton updates.
 * n_threads > 1 splits the cell loop over pthreads (used only to time the CPU
 * baseline on all host cores; the reference itself is single-threaded). */
typedef struct {
    const orc_problem *p; int64_t *I, *J; double *V; const int64_t *off;
    int64_t next, newton; int status; pthread_mutex_t mu;
} asm_job;

static void *asm_worker(void *arg)
{
    asm_job *jb = (asm_job *)arg;
    const int64_t chunk = 256;
    int64_t nwt = 0; int status = 0;
    for (;;) {
        pthread_mutex_lock(&jb->mu);
        int64_t b = jb->next; jb->next += chunk;
        pthread_mutex_unlock(&jb->mu);
        if (b >= jb->p->n_cell) break;
        int64_t e = b + chunk < jb->p->n_cell ? b + chunk : jb->p->n_cell;
        for (int64_t c = b; c < e; c++) {
            int64_t nw = 0;
            int rc = cell_emit(jb->p, c, jb->I + jb->off[c], jb->J + jb->off[c], jb->V + jb->off[c], &nw);
            nwt += nw;
            if (rc) status = rc;
        }
    }
    pthread_mutex_lock(&jb->mu);
    jb->newton += nwt;
    if (status) jb->status = status;
    pthread_mutex_unlock(&jb->mu);
    return NULL;
}

int orc_assemble_coo(const orc_problem *p, int64_t *I, int64_t *J, double *V, int64_t *cell_off,
                     int64_t *newton_total, int n_threads)
{
    int64_t *off = cell_off ? cell_off : (int64_t *)malloc(sizeof(int64_t) * (p->n_cell + 1));
    off[0] = 0;
    for (int64_t c = 0; c < p->n_cell; c++) off[c + 1] = off[c] + cell_coo_count(p, c);
    asm_job jb;
    jb.p = p; jb.I = I; jb.J = J; jb.V = V; jb.off = off; jb.next = 0; jb.newton = 0; jb.status = 0;
    pthread_mutex_init(&jb.mu, NULL);
    if (n_threads <= 1) asm_worker(&jb);
    else {
        if (n_threads > 256) n_threads = 256;
        pthread_t th[256];
        for (int t = 0; t < n_threads; t++) pthread_create(&th[t], NULL, asm_worker, &jb);
        for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
    }
    pthread_mutex_destroy(&jb.mu);
    if (newton_total) *newton_total = jb.newton;
    if (!cell_off) free(off);
    return jb.status;
}

/* ------------------------------------------------------------------ */
/* scipy.sparse coo_matrix(...).tocsr()                                 */
/* ------------------------------------------------------------------ */

typedef struct { int64_t col; double val; } kv_t;
static int kv_cmp(const void *a, const void *b)
{
    int64_t ca = ((const kv_t *)a)->col, cb = ((const kv_t *)b)->col;
    return (ca > cb) - (ca < cb);
}

/* coo_tocsr + csr_sort_indices + csr_sum_duplicates (scipy/sparse/sparsetools/
 * coo.h, csr.h).  Duplicates are summed in COO order within each (row, col)
 * (a stable per-row sort; scipy's std::sort is not stable, so the last bits of
 * sums with >2 duplicates may differ from scipy -- pattern is identical).
 * indptr has n_row+1 entries; indices/data need nnz_coo capacity.
 * Returns the CSR nnz. */
int64_t orc_coo_tocsr(int64_t n_row, int64_t nnz, const int64_t *I, const int64_t *J, const double *V,
                      int64_t *indptr, int64_t *indices, double *data)
{
    int64_t *cnt = (int64_t *)calloc((size_t)n_row + 1, sizeof(int64_t));
    for (int64_t k = 0; k < nnz; k++) cnt[I[k] + 1]++;
    for (int64_t r = 0; r < n_row; r++) cnt[r + 1] += cnt[r];
    kv_t *tmp = (kv_t *)malloc(sizeof(kv_t) * (size_t)(nnz > 0 ? nnz : 1));
    int64_t *cur = (int64_t *)malloc(sizeof(int64_t) * ((size_t)n_row + 1));
    memcpy(cur, cnt, sizeof(int64_t) * ((size_t)n_row + 1));
    for (int64_t k = 0; k < nnz; k++) { int64_t d = cur[I[k]]++; tmp[d].col = J[k]; tmp[d].val = V[k]; }
    int64_t out = 0;
    indptr[0] = 0;
    for (int64_t r = 0; r < n_row; r++) {
        int64_t b = cnt[r], e = cnt[r + 1];
        /* stable insertion/merge: qsort is not stable, so tag with position */
        if (e - b > 1) {
            /* small rows: insertion sort (stable) */
            for (int64_t i = b + 1; i < e; i++) {
                kv_t x = tmp[i]; int64_t j = i - 1;
                while (j >= b && tmp[j].col > x.col) { tmp[j + 1] = tmp[j]; j--; }
                tmp[j + 1] = x;
            }
        }
        int64_t i = b;
        while (i < e) {
            int64_t col = tmp[i].col; double s = tmp[i].val; i++;
            while (i < e && tmp[i].col == col) { s += tmp[i].val; i++; }
            indices[out] = col; data[out] = s; out++;
        }
        indptr[r + 1] = out;
    }
    (void)kv_cmp;
    free(cnt); free(tmp); free(cur);
    return out;
}

/* ------------------------------------------------------------------ */
/* weight functions                                   */
/* ------------------------------------------------------------------ */

/* np.inner of two 1-D float64 arrays = cblas_ddot (OpenBLAS): n < 32 runs the scalar tail loop
 * dot += y[i]*x[i], compiled with contraction -> fma chain starting from 0. */
static double np_inner(const double *a, const double *b, int n)
{
    double dot = 0.0;
    for (int i = 0; i < n; i++) dot = fma(b[i], a[i], dot);
    return dot;
}

/* planeSweepMeshes.rhoCalculation:991-1030, per overlay vertex.  vert_elem[v][m] = incident element of the mesh
 * slot m (-1 = none) of the face that first visits the vertex; elements carry their 3 or 4 corner vertices
 * (element.vertex) and corner weights.  rho[v][m]: weight of mesh m (slots >= 1 multiplied by 9, :1024-1025),
 * normalised by Python's left-to-right sum (:1027).  Returns 0, -1 "Too many iterations!", -2 ValueError (:1021). */
int orc_rho_eval(int64_t n_vert, int32_t n_mesh, const double *vert_xy, const int32_t *vert_elem,
                 const int32_t *elem_nv, const double *elem_xy, const double *elem_w, double *rho,
                 int64_t *newton_total)
{
    int64_t nw = 0;
    for (int64_t v = 0; v < n_vert; v++) {
        double *r = rho + v * n_mesh;
        const double *xy = vert_xy + 2 * v;
        for (int m = 0; m < n_mesh; m++) {
            r[m] = 0.0;
            int32_t e = vert_elem[v * n_mesh + m];
            if (e < 0) continue;
            const double *c = elem_xy + 8 * (int64_t)e, *w = elem_w + 4 * (int64_t)e;
            if (elem_nv[e] == 4) {
                double iso[2], N[4], dN[8];
                int it = orc_inv_newton(c, 4, xy, iso);
                if (it < 0) return -1;
                nw += it;
                quad_N_dN(iso[0], iso[1], N, dN);
                r[m] = np_inner(N, w, 4);
            } else if (elem_nv[e] == 3) {
                double iso[3];
                orc_inv_tri(c, xy, iso);
                r[m] = np_inner(iso, w, 3);
            } else
                return -2;
        }
        for (int m = 1; m < n_mesh; m++) r[m] *= 9;
        double s = 0.0;
        for (int m = 0; m < n_mesh; m++) s += r[m];
        for (int m = 0; m < n_mesh; m++) r[m] /= s;
    }
    if (newton_total) *newton_total = nw;
    return 0;
}
