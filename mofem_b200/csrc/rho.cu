// rho.cu -- weight functions of the overlay vertices.
//
// Replaces the per-vertex body of computGeometry/planeSweepMeshes.rhoCalculation:991-1030: for every overlay vertex and
// every mesh slot m with an incident element, the vertex is mapped back into that element (4 corners: Newton inverse of
// the bilinear map, invShapeFunction.invQuad:41; 3 corners: area coordinates, invTri:21), the corner weights are
// interpolated there (np.inner), slots >= 1 are multiplied by 9 (:1024-1025) and the row is normalised by its
// left-to-right sum (:1027).
//
// One thread per vertex: 16 B of coordinates and M element ids in (coalesced), M weights out; the element corner
// records (64 B of coordinates + 32 B of weights) are shared by the 4..9 vertices of the cells inside the element and
// hit L2.  Same device functions as the integration kernels (fem_device.cuh, compiled with -fmad=false).
#include "common.cuh"
#include "fem_device.cuh"

namespace mofem {

// cblas_ddot with n < 32 (OpenBLAS scalar tail): dot = fma(y_i, x_i, dot) from 0
template <int N>
__device__ __forceinline__ double np_inner(const double *a, const double *w)
{
    double dot = 0.0;
#pragma unroll
    for (int i = 0; i < N; i++) dot = fma(w[i], a[i], dot);
    return dot;
}

__global__ void __launch_bounds__(256)
k_rho(int64_t n_vert, int n_mesh, const double2 *__restrict__ vert_xy, const int32_t *__restrict__ vert_elem,
      int64_t n_elem, const int32_t *__restrict__ elem_nv, const double2 *__restrict__ elem_xy,
      const double2 *__restrict__ elem_w, double *__restrict__ rho, int *err, unsigned long long *newton_total)
{
    unsigned long long nw = 0;
    for (int64_t v = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; v < n_vert; v += (int64_t)gridDim.x * blockDim.x) {
        const double2 xy = vert_xy[v];
        double *out = rho + v * n_mesh;
        double sum = 0.0;
        bool fail = false;   // errors raise the flag and skip the vertex (no early return: the warp reduction below is full-mask)
        for (int m = 0; m < n_mesh && !fail; m++) {
            const int32_t e = vert_elem[v * n_mesh + m];
            double val = 0.0;
            if (e >= 0) {
                if (e >= n_elem) { atomicExch(err, MOFEM_E_ARG); fail = true; continue; }
                const int nv = elem_nv[e];
                double X[4], Y[4], W[4];
#pragma unroll
                for (int a = 0; a < 4; a++) {
                    const double2 c = __ldg(elem_xy + 4 * (int64_t)e + a);
                    X[a] = c.x; Y[a] = c.y;
                }
                const double2 w01 = __ldg(elem_w + 2 * (int64_t)e), w23 = __ldg(elem_w + 2 * (int64_t)e + 1);
                W[0] = w01.x; W[1] = w01.y; W[2] = w23.x; W[3] = w23.y;
                if (nv == 4) {
                    double r, s;
                    const int it = inverse_map<4>(X, Y, xy.x, xy.y, r, s);
                    if (it < 0) { atomicExch(err, MOFEM_E_NEWTON); fail = true; continue; }
                    nw += it;
                    double Nv[4], dr[4], ds[4];
                    Elem<4>::shape(r, s, Nv, dr, ds);
                    val = np_inner<4>(Nv, W);
                } else if (nv == 3) {
                    // invTri returns all three area coordinates (invShapeFunction.py:21-39)
                    const double area = (X[1] * Y[2] - X[2] * Y[1] - X[0] * (Y[2] - Y[1]) + Y[0] * (X[2] - X[1])) / 2;
                    const double a3[3] = {X[1] * Y[2] - X[2] * Y[1], X[2] * Y[0] - X[0] * Y[2], X[0] * Y[1] - X[1] * Y[0]};
                    const double b3[3] = {Y[1] - Y[2], Y[2] - Y[0], Y[0] - Y[1]};
                    const double c3[3] = {X[2] - X[1], X[0] - X[2], X[1] - X[0]};
                    double iso[3];
#pragma unroll
                    for (int i = 0; i < 3; i++) iso[i] = (a3[i] + b3[i] * xy.x + c3[i] * xy.y) / 2.0 / area;
                    val = np_inner<3>(iso, W);
                } else {
                    atomicExch(err, MOFEM_E_ELEMENT);
                    fail = true;
                    continue;
                }
            }
            if (m >= 1) val *= 9;
            out[m] = val;
            sum += val;
        }
        if (fail) continue;
        for (int m = 0; m < n_mesh; m++) out[m] = out[m] / sum;
    }
    if (newton_total) {
        for (int o = 16; o; o >>= 1) nw += __shfl_xor_sync(0xffffffffu, nw, o);
        if ((threadIdx.x & 31) == 0 && nw) atomicAdd(newton_total, nw);
    }
}

cudaError_t rho_run(int64_t n_vert, int n_mesh, const double *vert_xy, const int32_t *vert_elem, int64_t n_elem,
                    const int32_t *elem_nv, const double *elem_xy, const double *elem_w, double *rho, int *err,
                    unsigned long long *newton_total, cudaStream_t s)
{
    if (n_vert == 0) return cudaSuccess;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t want = (n_vert + 255) / 256;
    const int grid = (int)(want < (int64_t)sms * 8 ? want : (int64_t)sms * 8);
    k_rho<<<grid, 256, 0, s>>>(n_vert, n_mesh, reinterpret_cast<const double2 *>(vert_xy), vert_elem, n_elem, elem_nv,
                               reinterpret_cast<const double2 *>(elem_xy), reinterpret_cast<const double2 *>(elem_w), rho,
                               err, newton_total);
    return cudaGetLastError();
}

}  // namespace mofem
