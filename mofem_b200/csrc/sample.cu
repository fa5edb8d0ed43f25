// sample.cu -- displacement / stress sampler on the solution.
//
// Replaces plotTools/plotResults.getGraphDataFE:137 and getGraphDataOFE:15 as driven by
// outputFunctions/writeSolution.writeSampleData2D:41: every regular element and every integration triangle of an
// overlay cell ("unit") is sampled on an nS x nS grid of parent coordinates; per point X, Y, U, V, Sxx, Syy, Sxy.
//   regular quad4 / quad9   xy = N(r,s) X ; u = N u_e ; stress = D B u_e
//                           quad4 with solver 3: ICM recovery  D (B - G H^-1 E^T) u_e   (shapeICMFE:53, ICMFE:49)
//   regular tri3 / tri6     zeros (the reference's branches are `pass`)
//   overlay triangle        xy from the collapsed quad map of the (straight 3-node | curved 6-node) triangle; for every
//                           incident element k: xi_k = invmap_k(xy), u += rho_k N_k u_k,
//                           stress += rho_k D B_k u_k + D LC_k (N_k u_k); D = material of the FIRST incident element
// One thread per sample point.  Same device functions as the integration kernels (fem_device.cuh, -fmad=false).
#include "common.cuh"
#include "fem_device.cuh"

namespace mofem {

__constant__ double c_s2p[2] = {-0.577350269189626, 0.577350269189626};   // gaussQuad(2), numericalIntegration.py:121

template <int N>
__device__ __forceinline__ void load_elem_s(const DevProblem &p, int e, double *X, double *Y, const int32_t *&en)
{
    en = p.elem_nodes + (int64_t)e * 9;
#pragma unroll
    for (int a = 0; a < N; a++) {
        const double2 xy = *reinterpret_cast<const double2 *>(p.node_xy + 2 * (int64_t)en[a]);
        X[a] = xy.x; Y[a] = xy.y;
    }
}

// D applied to a strain vector (exx, eyy, gxy)
__device__ __forceinline__ void apply_D(const double *D, double exx, double eyy, double gxy, double &sxx, double &syy, double &sxy)
{
    sxx = D[0] * exx + D[1] * eyy + D[2] * gxy;
    syy = D[3] * exx + D[4] * eyy + D[5] * gxy;
    sxy = D[6] * exx + D[7] * eyy + D[8] * gxy;
}

// strain of the element field at parent point (r,s): B u_e, and the displacement N u_e
template <int N>
__device__ __forceinline__ void field_at(const double *X, const double *Y, const int32_t *en, const double *__restrict__ u,
                                         double r, double s, double &ux, double &uy, double &exx, double &eyy, double &gxy)
{
    double Nv[N], gx[N], gy[N];
    shape_grad<N>(X, Y, r, s, Nv, gx, gy);
    ux = uy = exx = eyy = gxy = 0.0;
#pragma unroll
    for (int a = 0; a < N; a++) {
        const double2 ua = *reinterpret_cast<const double2 *>(u + 2 * (int64_t)en[a]);
        ux += Nv[a] * ua.x; uy += Nv[a] * ua.y;
        exx += gx[a] * ua.x; eyy += gy[a] * ua.y; gxy += gy[a] * ua.x + gx[a] * ua.y;
    }
}

// ICM strain correction G H^-1 E^T u_e at (r,s): H, E integrated with the 2x2 rule as stiffnessMatrix.ICMFE:49-67
__device__ void icm_correction(const double *X, const double *Y, const double *D, const int32_t *en, const double *__restrict__ u,
                               double r, double s, double &cxx, double &cyy, double &cxy)
{
    double H[16], v[4];
    for (int i = 0; i < 16; i++) H[i] = 0.0;
    for (int i = 0; i < 4; i++) v[i] = 0.0;
    double ue[8];
    for (int a = 0; a < 4; a++) { ue[2 * a] = u[2 * (int64_t)en[a]]; ue[2 * a + 1] = u[2 * (int64_t)en[a] + 1]; }
    for (int q = 0; q < 4; q++) {
        const double rq = c_s2p[q >> 1], sq = c_s2p[q & 1];
        double Nv[4], dr[4], ds[4];
        Elem<4>::shape(rq, sq, Nv, dr, ds);
        double j00, j01, j10, j11;
        jacobian<4>(dr, ds, X, Y, j00, j01, j10, j11);
        LU2 f = lu2_factor(j00, j01, j10, j11);
        const double ru = 1.0 / f.u11, ra = 1.0 / f.a00;
        double gx[6], gy[6];
        for (int a = 0; a < 4; a++) lu2_solve_col(f, ru, ra, dr[a], ds[a], gx[a], gy[a]);
        lu2_solve_col(f, ru, ra, -2.0 * rq, 0.0, gx[4], gy[4]);
        lu2_solve_col(f, ru, ra, 0.0, -2.0 * sq, gx[5], gy[5]);
        const double w = lu2_det(f);
        // G (3 x 4): columns (gx4,0,gy4) (0,gy4,gx4) (gx5,0,gy5) (0,gy5,gx5);  strain of u_e at this Gauss point
        double G[3][4] = {{gx[4], 0.0, gx[5], 0.0}, {0.0, gy[4], 0.0, gy[5]}, {gy[4], gx[4], gy[5], gx[5]}};
        double exx = 0.0, eyy = 0.0, gxy = 0.0;
        for (int a = 0; a < 4; a++) { exx += gx[a] * ue[2 * a]; eyy += gy[a] * ue[2 * a + 1]; gxy += gy[a] * ue[2 * a] + gx[a] * ue[2 * a + 1]; }
        double sxx, syy, sxy;
        apply_D(D, exx, eyy, gxy, sxx, syy, sxy);             // D B u_e  (D symmetric: E^T u_e = sum G^T D B u_e w)
        for (int i = 0; i < 4; i++) {
            v[i] += (G[0][i] * sxx + G[1][i] * syy + G[2][i] * sxy) * w;
            double dg0, dg1, dg2;
            apply_D(D, G[0][i], G[1][i], G[2][i], dg0, dg1, dg2);
            for (int j = 0; j < 4; j++) H[4 * j + i] += (G[0][j] * dg0 + G[1][j] * dg1 + G[2][j] * dg2) * w;
        }
    }
    // w = H^-1 v (partial-pivot LU, as numpy.linalg.solve)
    for (int c = 0; c < 4; c++) {
        int pv = c; double best = fabs(H[4 * c + c]);
        for (int rr = c + 1; rr < 4; rr++) if (fabs(H[4 * rr + c]) > best) { best = fabs(H[4 * rr + c]); pv = rr; }
        if (pv != c) {
            for (int j = 0; j < 4; j++) { double t = H[4 * c + j]; H[4 * c + j] = H[4 * pv + j]; H[4 * pv + j] = t; }
            double t = v[c]; v[c] = v[pv]; v[pv] = t;
        }
        const double rp = 1.0 / H[4 * c + c];
        for (int rr = c + 1; rr < 4; rr++) {
            const double l = H[4 * rr + c] * rp;
            for (int j = c + 1; j < 4; j++) H[4 * rr + j] -= l * H[4 * c + j];
            v[rr] -= l * v[c];
        }
    }
    for (int rr = 3; rr >= 0; rr--) {
        double acc = v[rr];
        for (int c = rr + 1; c < 4; c++) acc -= H[4 * rr + c] * v[c];
        v[rr] = acc / H[4 * rr + rr];
    }
    // G at the sample point
    double Nv[4], dr[4], ds[4];
    Elem<4>::shape(r, s, Nv, dr, ds);
    double j00, j01, j10, j11;
    jacobian<4>(dr, ds, X, Y, j00, j01, j10, j11);
    LU2 f = lu2_factor(j00, j01, j10, j11);
    const double ru = 1.0 / f.u11, ra = 1.0 / f.a00;
    double g4x, g4y, g5x, g5y;
    lu2_solve_col(f, ru, ra, -2.0 * r, 0.0, g4x, g4y);
    lu2_solve_col(f, ru, ra, 0.0, -2.0 * s, g5x, g5y);
    cxx = g4x * v[0] + g5x * v[2];
    cyy = g4y * v[1] + g5y * v[3];
    cxy = g4y * v[0] + g4x * v[1] + g5y * v[2] + g5x * v[3];
}

// unit table: regular cell -> one unit (tri = -1); overlay cell -> one unit per triangle
__global__ void k_unit_counts(DevProblem p, int32_t *__restrict__ cnt)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= p.n_cell) return;
    cnt[c] = p.cell_kind[c] == CELL_REGULAR ? 1 : (int32_t)(p.cell_tri_ptr[c + 1] - p.cell_tri_ptr[c]);
}

__global__ void k_fill_units(DevProblem p, const int64_t *__restrict__ unit_off, int32_t *__restrict__ unit_cell,
                             int64_t *__restrict__ unit_tri)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= p.n_cell) return;
    int64_t k = unit_off[c];
    if (p.cell_kind[c] == CELL_REGULAR) { unit_cell[k] = (int32_t)c; unit_tri[k] = -1; return; }
    for (int64_t t = p.cell_tri_ptr[c]; t < p.cell_tri_ptr[c + 1]; t++, k++) { unit_cell[k] = (int32_t)c; unit_tri[k] = t; }
}

template <int N>
__device__ __forceinline__ bool overlay_element(const DevProblem &p, int e, const double *__restrict__ u, const double *D,
                                                double x, double y, double rho, double drx, double dry,
                                                double &U, double &V, double &sxx, double &syy, double &sxy)
{
    double X[N], Y[N];
    const int32_t *en;
    load_elem_s<N>(p, e, X, Y, en);
    double r, s;
    if (inverse_map<N>(X, Y, x, y, r, s) < 0) return false;
    double ux, uy, exx, eyy, gxy;
    field_at<N>(X, Y, en, u, r, s, ux, uy, exx, eyy, gxy);
    U += rho * ux; V += rho * uy;
    // rho D B u  +  D LC (N u),  LC = [[drx,0],[0,dry],[dry,drx]]
    double a, b, c;
    apply_D(D, rho * exx + drx * ux, rho * eyy + dry * uy, rho * gxy + dry * ux + drx * uy, a, b, c);
    sxx += a; syy += b; sxy += c;
    return true;
}

__global__ void __launch_bounds__(128)
k_sample(DevProblem p, const int32_t *__restrict__ unit_cell, const int64_t *__restrict__ unit_tri, int64_t n_unit, int nS,
         const double *__restrict__ grid, const double *__restrict__ u, double *__restrict__ out, int *err)
{
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int per = nS * nS;
    if (gid >= n_unit * per) return;
    const int64_t k = gid / per;
    const int ij = (int)(gid % per), i = ij / nS, j = ij % nS;
    const double r = grid[i], s = grid[j];
    const int64_t c = unit_cell[k], tr = unit_tri[k];
    const int M = p.n_mesh;
    const int32_t *ce = p.cell_elem + c * M;
    double X = 0, Y = 0, U = 0, V = 0, sxx = 0, syy = 0, sxy = 0;
    int e0 = -1;
    for (int m = 0; m < M; m++) if (ce[m] >= 0) { e0 = ce[m]; break; }
    double D[9];
    if (e0 >= 0) { const double *Dp = p.mat_D + 9 * (int64_t)p.elem_mat[e0]; for (int q = 0; q < 9; q++) D[q] = Dp[q]; }
    if (tr < 0) {
        // ---- regular element: getGraphDataFE -------------------------------------------------------------
        const int nn = e0 >= 0 ? p.elem_nn[e0] : 0;
        if (nn == 4) {
            double Xe[4], Ye[4], Nv[4], dr[4], ds[4];
            const int32_t *en;
            load_elem_s<4>(p, e0, Xe, Ye, en);
            Elem<4>::shape(r, s, Nv, dr, ds);
            for (int a = 0; a < 4; a++) { X += Nv[a] * Xe[a]; Y += Nv[a] * Ye[a]; }
            double exx, eyy, gxy;
            field_at<4>(Xe, Ye, en, u, r, s, U, V, exx, eyy, gxy);
            if (p.solver == 3) {
                double cxx, cyy, cxy;
                icm_correction(Xe, Ye, D, en, u, r, s, cxx, cyy, cxy);
                exx -= cxx; eyy -= cyy; gxy -= cxy;
            } else if (p.solver != 2) *err = MOFEM_E_ARG;      // the reference raises ValueError for other solvers
            apply_D(D, exx, eyy, gxy, sxx, syy, sxy);
        } else if (nn == 9) {
            double Xe[9], Ye[9], Nv[9], dr[9], ds[9];
            const int32_t *en;
            load_elem_s<9>(p, e0, Xe, Ye, en);
            Elem<9>::shape(r, s, Nv, dr, ds);
            for (int a = 0; a < 9; a++) { X += Nv[a] * Xe[a]; Y += Nv[a] * Ye[a]; }
            double exx, eyy, gxy;
            field_at<9>(Xe, Ye, en, u, r, s, U, V, exx, eyy, gxy);
            apply_D(D, exx, eyy, gxy, sxx, syy, sxy);
        }   // tri3 / tri6: zeros, like the reference
    } else {
        // ---- integration triangle of an overlay cell: getGraphDataOFE -------------------------------------------
        double tx[6], ty[6];
        for (int v = 0; v < 3; v++) { tx[v] = p.tri_xy[6 * tr + 2 * v]; ty[v] = p.tri_xy[6 * tr + 2 * v + 1]; }
        const bool curved = p.tri_curved && p.tri_curved[tr];
        TriPoint tp;
        double L1, L2;                      // area coordinates of the sample point in the triangle
        double Nt6[6], gxt[6], gyt[6];
        if (!curved) {
            // collapsed bilinear quad [c0,c1,c2,c2]
            double Nv[4], dr[4], ds[4];
            Elem<4>::shape(r, s, Nv, dr, ds);
            X = Nv[0] * tx[0] + Nv[1] * tx[1] + Nv[2] * tx[2] + Nv[3] * tx[2];
            Y = Nv[0] * ty[0] + Nv[1] * ty[1] + Nv[2] * ty[2] + Nv[3] * ty[2];
            inverse_map<3>(tx, ty, X, Y, L1, L2);
            tri_straight(tx, ty, tp);
        } else {
            for (int v = 0; v < 3; v++) { tx[3 + v] = p.tri_mid[6 * tr + 2 * v]; ty[3 + v] = p.tri_mid[6 * tr + 2 * v + 1]; }
            // collapsed 9-node quad: corners c0,c1,c2,c2; mid-sides c3,c4,c2,c5; centre c5  (plotResults.py:74-79)
            const double qx[9] = {tx[0], tx[1], tx[2], tx[2], tx[3], tx[4], tx[2], tx[5], tx[5]};
            const double qy[9] = {ty[0], ty[1], ty[2], ty[2], ty[3], ty[4], ty[2], ty[5], ty[5]};
            double Nv[9], dr[9], ds[9];
            Elem<9>::shape(r, s, Nv, dr, ds);
            for (int a = 0; a < 9; a++) { X += Nv[a] * qx[a]; Y += Nv[a] * qy[a]; }
            if (inverse_map<6>(tx, ty, X, Y, L1, L2) < 0) *err = MOFEM_E_NEWTON;
            shape_grad<6>(tx, ty, L1, L2, Nt6, gxt, gyt);
        }
        for (int m = 0; m < M; m++) {
            const int e = ce[m];
            if (e < 0) continue;
            double rho[3];
            for (int v = 0; v < 3; v++) rho[v] = p.tri_rho[(3 * tr + v) * M + m];
            double rhoV, drx, dry;
            if (!curved) {
                tri_grad_rho(tp, rho, drx, dry);
                rhoV = L1 * rho[0] + L2 * rho[1] + (1.0 - L1 - L2) * rho[2];
            } else {
                const double re[6] = {rho[0], rho[1], rho[2], 0.5 * (rho[0] + rho[1]), 0.5 * (rho[1] + rho[2]), 0.5 * (rho[2] + rho[0])};
                drx = dry = rhoV = 0.0;
                for (int a = 0; a < 6; a++) { drx += gxt[a] * re[a]; dry += gyt[a] * re[a]; rhoV += Nt6[a] * re[a]; }
            }
            const int nn = p.elem_nn[e];
            bool ok = true;
            if (nn == 4) ok = overlay_element<4>(p, e, u, D, X, Y, rhoV, drx, dry, U, V, sxx, syy, sxy);
            else if (nn == 9) ok = overlay_element<9>(p, e, u, D, X, Y, rhoV, drx, dry, U, V, sxx, syy, sxy);
            else *err = MOFEM_E_ELEMENT;          // the reference raises ValueError for triangular incident elements
            if (!ok) *err = MOFEM_E_NEWTON;
        }
    }
    double *o = out + (k * 7) * per + ij;
    o[0] = X; o[per] = Y; o[2 * per] = U; o[3 * per] = V; o[4 * per] = sxx; o[5 * per] = syy; o[6 * per] = sxy;
}

cudaError_t sample_count_units(const DevProblem &p, int32_t *cnt, cudaStream_t s)
{
    if (p.n_cell == 0) return cudaSuccess;
    k_unit_counts<<<(unsigned)((p.n_cell + 255) / 256), 256, 0, s>>>(p, cnt);
    return cudaGetLastError();
}

cudaError_t sample_fill_units(const DevProblem &p, const int64_t *unit_off, int32_t *unit_cell, int64_t *unit_tri, cudaStream_t s)
{
    if (p.n_cell == 0) return cudaSuccess;
    k_fill_units<<<(unsigned)((p.n_cell + 255) / 256), 256, 0, s>>>(p, unit_off, unit_cell, unit_tri);
    return cudaGetLastError();
}

cudaError_t sample_run(const DevProblem &p, const int32_t *unit_cell, const int64_t *unit_tri, int64_t n_unit, int nS,
                       const double *grid, const double *u, double *out, int *err, cudaStream_t s)
{
    const int64_t n = n_unit * nS * nS;
    if (n == 0) return cudaSuccess;
    k_sample<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(p, unit_cell, unit_tri, n_unit, nS, grid, u, out, err);
    return cudaGetLastError();
}

}  // namespace mofem
