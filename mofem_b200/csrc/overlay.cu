// overlay.cu -- the flattened overlay of the structured patch family, generated on the GPU.
//
// Replaces, for this family, the reference's host pipeline planeSweepMeshes.meshOverlay:354 + rhoCalculation:991 +
// triangulation.polygonTriangulation + the flattening of the object graph: pairwise convex clipping of every patch
// element against the four mesh-0 edges through it, the reference's ear rule on the pentagons / hexagons left outside the
// patch, weight functions by the inverse bilinear map.  The arithmetic lives in overlay_core.h, shared with the host build
// that tests/test_overlay_core.py compares bit for bit with mofem_b200/overlay_gen.py; this file only maps work items to
// threads.  All counts per patch are fixed (100 two-element cells with 2 triangles, 16 pentagons with 3, 4 hexagons
// with 4), so every output offset is closed-form: no scan, no atomics, every thread writes its own records.
// Compiled with -fmad=false (NumPy's element-wise arithmetic is unfused).
#include "common.cuh"
#include "overlay_core.h"

namespace mofem {
using namespace mofem_ovl;

__global__ void __launch_bounds__(128)
k_ovl_elements(Meshes m, Layout L, Outputs o)
{
    const int64_t n = (int64_t)(L.b_hi - L.b_lo) * L.Px * 25;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const int p = (int)(t % 5), q = (int)((t / 5) % 5);
        const int64_t pp = t / 25;
        gen_patch_element(m, L, o, L.b_lo + (int)(pp / L.Px), (int)(pp % L.Px), q, p);
    }
}

__global__ void __launch_bounds__(128)
k_ovl_ring(Meshes m, Layout L, Outputs o)
{
    const int64_t n = (int64_t)(L.b_hi - L.b_lo) * L.Px * RING_CELLS;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t pp = t / RING_CELLS;
        gen_ring_cell(m, L, o, L.b_lo + (int)(pp / L.Px), (int)(pp % L.Px), (int)(t % RING_CELLS));
    }
}

__global__ void __launch_bounds__(256)
k_ovl_regular(Layout L, Outputs o, int j0, int j1)
{
    const int64_t n = (int64_t)(j1 - j0) * L.nx;
    const int64_t before0 = untouched_before_row(j0, L.nx);
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        const int j = j0 + (int)(t / L.nx), i = (int)(t % L.nx);
        gen_regular(L, o, j, i, untouched_before_row(j, L.nx) - before0);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) o.cell_tri_ptr[L.n_cell()] = L.n_tri;
}

void overlay_sizes(int nx, int ny, int b_lo, int b_hi, int n_mesh, int64_t out[5])
{
    const Layout L = make_layout(nx, ny, b_lo, b_hi, n_mesh);
    out[0] = L.n_reg; out[1] = L.n_ring; out[2] = L.n_piece; out[3] = L.n_tri; out[4] = L.n_cell();
}

cudaError_t overlay_run(int nx, int ny, int b_lo, int b_hi, int n_mesh, const double *X0, const double *X1, int32_t *cell_elem,
                        int32_t *cell_kind, int64_t *cell_tri_ptr, double *tri_xy, double *tri_rho, cudaStream_t s)
{
    const Layout L = make_layout(nx, ny, b_lo, b_hi, n_mesh);
    const Meshes m{nx, ny, L.Px, L.Py, X0, X1};
    const Outputs o{cell_elem, cell_kind, cell_tri_ptr, tri_xy, tri_rho};
    const int64_t patches = (int64_t)(b_hi - b_lo) * L.Px;
    auto grid = [](int64_t n, int block) { const int64_t g = (n + block - 1) / block; return (unsigned)(g < 148 * 64 ? (g > 0 ? g : 1) : 148 * 64); };
    if (patches > 0) {
        k_ovl_elements<<<grid(patches * 25, 128), 128, 0, s>>>(m, L, o);
        k_ovl_ring<<<grid(patches * RING_CELLS, 128), 128, 0, s>>>(m, L, o);
    }
    const int j0 = PERIOD * b_lo - 1 > 0 ? PERIOD * b_lo - 1 : 0, j1 = PERIOD * b_hi;
    k_ovl_regular<<<grid((int64_t)(j1 - j0) * nx, 256), 256, 0, s>>>(L, o, j0, j1);
    return cudaGetLastError();
}

}  // namespace mofem
