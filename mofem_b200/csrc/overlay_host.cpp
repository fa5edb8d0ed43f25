// Host build of overlay_core.h -- the SAME source the CUDA kernels of overlay.cu compile -- as liboverlay_host.so:
//   * untimed preprocessing where no device is at hand yet (the multi-GPU driver builds each rank's strip before CUDA / NCCL
//     are initialised): mofem_b200/overlay_gen.py: patch_family(engine="native"), 10x faster than the NumPy generator and
//     without its forked worker processes;
//   * the CPU check that the kernel source reproduces the NumPy generator bit for bit (tests/test_overlay_core.py).
// Built by mofem_b200/build.py:  g++ -O2 -std=c++17 -ffp-contract=off -shared -fPIC overlay_host.cpp
#include "overlay_core.h"

using namespace mofem_ovl;

extern "C" void ovl_host_sizes(int nx, int ny, int b_lo, int b_hi, int n_mesh, int64_t *out)
{
    const Layout L = make_layout(nx, ny, b_lo, b_hi, n_mesh);
    out[0] = L.n_reg; out[1] = L.n_ring; out[2] = L.n_piece; out[3] = L.n_tri;
}

extern "C" void ovl_host_generate(int nx, int ny, int b_lo, int b_hi, int n_mesh, const double *X0, const double *X1,
                                  int32_t *cell_elem, int32_t *cell_kind, int64_t *cell_tri_ptr, double *tri_xy, double *tri_rho)
{
    const Layout L = make_layout(nx, ny, b_lo, b_hi, n_mesh);
    const Meshes m{nx, ny, L.Px, L.Py, X0, X1};
    const Outputs o{cell_elem, cell_kind, cell_tri_ptr, tri_xy, tri_rho};
    for (int b = b_lo; b < b_hi; b++)
        for (int a = 0; a < L.Px; a++) {
            for (int q = 0; q < 5; q++)
                for (int p = 0; p < 5; p++) gen_patch_element(m, L, o, b, a, q, p);
            for (int r = 0; r < RING_CELLS; r++) gen_ring_cell(m, L, o, b, a, r);
        }
    const int j0 = PERIOD * b_lo - 1 > 0 ? PERIOD * b_lo - 1 : 0, j1 = PERIOD * b_hi;
    for (int j = j0; j < j1; j++)
        for (int i = 0; i < nx; i++) gen_regular(L, o, j, i, untouched_before_row(j, nx) - untouched_before_row(j0, nx));
    cell_tri_ptr[L.n_cell()] = L.n_tri;
}

// test hook: the ear rule alone on one polygon of n <= 6 vertices (pts [n][2]); returns the triangle count
extern "C" int ovl_host_ear(const double *pts, int n, int32_t *tri_out)
{
    if (n < 3 || n > EAR_MAX) return -1;
    P2 p[EAR_MAX];
    for (int k = 0; k < n; k++) p[k] = P2{pts[2 * k], pts[2 * k + 1]};
    int tri[EAR_MAX - 2][3];
    const int nt = ear_triangulate(p, n, tri);
    for (int t = 0; t < nt; t++)
        for (int v = 0; v < 3; v++) tri_out[3 * t + v] = tri[t][v];
    return nt;
}
