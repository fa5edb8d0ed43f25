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

// ---- general families (overlay_general.h): tri3 patches, staggered third mesh -------------------------------------------
#include <string.h>

#include <string>
#include <thread>

#include "overlay_general.h"

namespace {
struct GeneralResult {
    mofem_ovg::Output out;
    int n_mesh = 0;
    std::string err;
};
}  // namespace

// Builds the cells of the mesh-0 element rows [j_lo, j_hi) (all columns).  patch_* arrays have n_patch_mesh entries:
// patches per direction, tri flag, first global element id, node coordinates [Py][Px][6][6][2].
// Returns an opaque handle (never null); ovl_general_error() tells whether the build failed.
extern "C" void *ovl_general_build(int nx, int ny, int j_lo, int j_hi, const double *X0, int n_patch_mesh, const int32_t *patch_px,
                                   const int32_t *patch_py, const int32_t *patch_tri, const int64_t *patch_elem0,
                                   const double *const *patch_X, int n_threads)
{
    using namespace mofem_ovg;
    GeneralResult *res = new GeneralResult;
    if (n_patch_mesh < 1 || n_patch_mesh > MAX_MESH - 1 || j_lo < 0 || j_hi > ny || j_lo >= j_hi) { res->err = "bad arguments"; return res; }
    Input in{};
    in.nx = nx; in.ny = ny; in.X0 = X0; in.n_patch_mesh = n_patch_mesh;
    for (int k = 0; k < n_patch_mesh; k++) in.pm[k] = PatchMesh{patch_px[k], patch_py[k], patch_tri[k], patch_elem0[k], patch_X[k]};
    res->n_mesh = in.n_mesh();
    const int rows = j_hi - j_lo;
    int T = n_threads < 1 ? 1 : n_threads;
    if (T > rows) T = rows;
    std::vector<Output> parts(T);
    std::vector<std::string> errs(T);
    auto work = [&](int t) {
        const int r0 = j_lo + (int)((int64_t)rows * t / T), r1 = j_lo + (int)((int64_t)rows * (t + 1) / T);
        try {
            for (int j = r0; j < r1; j++)
                for (int i = 0; i < nx; i++) gen_element(in, j, i, parts[t]);
        } catch (const std::exception &e) { errs[t] = e.what(); }
    };
    if (T == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < T; t++) th.emplace_back(work, t);
        for (auto &x : th) x.join();
    }
    for (int t = 0; t < T; t++) if (!errs[t].empty()) { res->err = errs[t]; return res; }
    Output &o = res->out;
    for (int t = 0; t < T; t++) {       // rows in order: the result does not depend on the thread count
        o.cell_elem.insert(o.cell_elem.end(), parts[t].cell_elem.begin(), parts[t].cell_elem.end());
        o.cell_kind.insert(o.cell_kind.end(), parts[t].cell_kind.begin(), parts[t].cell_kind.end());
        o.cell_ntri.insert(o.cell_ntri.end(), parts[t].cell_ntri.begin(), parts[t].cell_ntri.end());
        o.tri_xy.insert(o.tri_xy.end(), parts[t].tri_xy.begin(), parts[t].tri_xy.end());
        o.tri_rho.insert(o.tri_rho.end(), parts[t].tri_rho.begin(), parts[t].tri_rho.end());
        parts[t] = Output{};
    }
    return res;
}
extern "C" const char *ovl_general_error(void *h) { return static_cast<GeneralResult *>(h)->err.c_str(); }
extern "C" void ovl_general_sizes(void *h, int64_t out[2])
{
    const GeneralResult *r = static_cast<GeneralResult *>(h);
    out[0] = (int64_t)r->out.cell_kind.size(); out[1] = (int64_t)r->out.tri_xy.size() / 6;
}
extern "C" void ovl_general_copy(void *h, int32_t *cell_elem, int32_t *cell_kind, int64_t *cell_tri_ptr, double *tri_xy, double *tri_rho)
{
    const GeneralResult *r = static_cast<GeneralResult *>(h);
    const mofem_ovg::Output &o = r->out;
    if (!o.cell_elem.empty()) memcpy(cell_elem, o.cell_elem.data(), sizeof(int32_t) * o.cell_elem.size());
    if (!o.cell_kind.empty()) memcpy(cell_kind, o.cell_kind.data(), sizeof(int32_t) * o.cell_kind.size());
    int64_t acc = 0;
    for (size_t c = 0; c < o.cell_ntri.size(); c++) { cell_tri_ptr[c] = acc; acc += o.cell_ntri[c]; }
    cell_tri_ptr[o.cell_ntri.size()] = acc;
    if (!o.tri_xy.empty()) memcpy(tri_xy, o.tri_xy.data(), sizeof(double) * o.tri_xy.size());
    if (!o.tri_rho.empty()) memcpy(tri_rho, o.tri_rho.data(), sizeof(double) * o.tri_rho.size());
}
extern "C" void ovl_general_free(void *h) { delete static_cast<GeneralResult *>(h); }
