// overlay_core.h -- overlay of the structured patch family by pairwise convex clipping.
//
// For the synthetic patch family (mesh 0 = nx x ny quads, meshes >= 1 = disjoint 5 x 5 patches of the same
// spacing, period 8, straight patch borders) the reference's plane sweep + rhoCalculation + polygonTriangulation + flatten
// reduce to closed-form work per patch:
//   * every patch element contains exactly one mesh-0 node and is cut by the four mesh-0 edges leaving it into 4 convex
//     quadrilaterals ("pieces", two incident elements each, 2 triangles by the reference's ear rule);
//   * the 20 mesh-0 elements a patch border runs through keep a pentagon (edges) or hexagon (corners) outside the patch
//     ("ring cells", one incident element, triangulated by the reference's recursive rule triangulation.py:49-139);
//   * weight functions at the piece vertices: inverse bilinear map into the patch element, corner-weight interpolation,
//     x 9, normalisation (planeSweepMeshes.py:991-1030).
// The functions below restate, operation for operation, mofem_b200/overlay_gen.py (which is pinned against the reference's
// own sweep at n0 = 16); they are __host__ __device__ so that the SAME source is checked bit for bit against the NumPy
// generator on the CPU (tests/test_overlay_core.py builds it with the host compiler, -ffp-contract=off) and runs in the
// CUDA kernels of overlay.cu (-fmad=false).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define OVL_HD __host__ __device__ __forceinline__
#else
#define OVL_HD inline
#endif

namespace mofem_ovl {

constexpr int PERIOD = 8;
constexpr int RING_CELLS = 20;         // per patch: 16 pentagons + 4 hexagons
constexpr int RING_TRIS = 16 * 3 + 4 * 4;
constexpr int PIECES = 100;            // per patch: 25 elements x 4
struct P2 { double x, y; };

struct Meshes {
    int nx, ny, Px, Py;
    const double *X0;   // [ny+1][nx+1][2]
    const double *X1;   // [Py][Px][6][6][2]
    OVL_HD P2 n0(int j, int i) const { const double *p = X0 + 2 * ((int64_t)j * (nx + 1) + i); return P2{p[0], p[1]}; }
    OVL_HD P2 n1(int b, int a, int t, int s) const
    {
        const double *p = X1 + 2 * ((((int64_t)b * Px + a) * 6 + t) * 6 + s);
        return P2{p[0], p[1]};
    }
};

// overlay_gen._intersect: intersection of segment p-q with the line through c-d
OVL_HD P2 intersect(P2 p, P2 q, P2 c, P2 d)
{
    const double r0 = q.x - p.x, r1 = q.y - p.y, e0 = d.x - c.x, e1 = d.y - c.y;
    const double den = r0 * e1 - r1 * e0;
    const double t = ((c.x - p.x) * e1 - (c.y - p.y) * e0) / den;
    return P2{p.x + t * r0, p.y + t * r1};
}

// the four cut points of patch element (q, p) of patch (a, b), with the straight patch borders pinned exactly
struct Cuts { P2 C00, C10, C11, C01, Pn, IW, IE, IS, IN; };
OVL_HD Cuts element_cuts(const Meshes &m, int b, int a, int q, int p)
{
    Cuts c;
    c.C00 = m.n1(b, a, q, p); c.C10 = m.n1(b, a, q, p + 1); c.C11 = m.n1(b, a, q + 1, p + 1); c.C01 = m.n1(b, a, q + 1, p);
    const int ip = PERIOD * a + 2 + p, jp = PERIOD * b + 2 + q;
    c.Pn = m.n0(jp, ip);
    c.IW = intersect(c.Pn, m.n0(jp, ip - 1), c.C00, c.C01);
    c.IE = intersect(c.Pn, m.n0(jp, ip + 1), c.C10, c.C11);
    c.IS = intersect(c.Pn, m.n0(jp - 1, ip), c.C00, c.C10);
    c.IN = intersect(c.Pn, m.n0(jp + 1, ip), c.C01, c.C11);
    if (p == 0) c.IW.x = m.n1(b, a, 0, 0).x;
    if (p == 4) c.IE.x = m.n1(b, a, 0, 5).x;
    if (q == 0) c.IS.y = m.n1(b, a, 0, 0).y;
    if (q == 4) c.IN.y = m.n1(b, a, 5, 0).y;
    return c;
}

// overlay_gen._inv_quad_N: shape functions of a quad4 at the inverse image of xy (6 Newton updates from (0, 0))
OVL_HD void inv_quad_N(const P2 *coord, P2 xy, double *N)
{
    const double X0 = coord[0].x - coord[0].x, X1 = coord[1].x - coord[0].x, X2 = coord[2].x - coord[0].x, X3 = coord[3].x - coord[0].x;
    const double Y0 = coord[0].y - coord[0].y, Y1 = coord[1].y - coord[0].y, Y2 = coord[2].y - coord[0].y, Y3 = coord[3].y - coord[0].y;
    const double x = xy.x - coord[0].x, y = xy.y - coord[0].y;
    double r = 0.0, s = 0.0;
    for (int it = 0; it < 6; it++) {
        const double rm = 1.0 - r, rp = 1.0 + r, sm = 1.0 - s, sp = 1.0 + s;
        const double N0 = 0.25 * rm * sm, N1 = 0.25 * rp * sm, N2 = 0.25 * rp * sp, N3 = 0.25 * rm * sp;
        const double fx = x - (((N0 * X0 + N1 * X1) + N2 * X2) + N3 * X3);
        const double fy = y - (((N0 * Y0 + N1 * Y1) + N2 * Y2) + N3 * Y3);
        const double xr = 0.25 * (sm * (X1 - X0) + sp * (X2 - X3)), xs = 0.25 * (rm * (X3 - X0) + rp * (X2 - X1));
        const double yr = 0.25 * (sm * (Y1 - Y0) + sp * (Y2 - Y3)), ys = 0.25 * (rm * (Y3 - Y0) + rp * (Y2 - Y1));
        const double det = xr * ys - xs * yr;
        r = r + (fx * ys - fy * xs) / det;
        s = s + (fy * xr - fx * yr) / det;
    }
    const double rm = 1.0 - r, rp = 1.0 + r, sm = 1.0 - s, sp = 1.0 + s;
    N[0] = 0.25 * rm * sm; N[1] = 0.25 * rp * sm; N[2] = 0.25 * rp * sp; N[3] = 0.25 * rm * sp;
}

// node weight of patch node (t, s): 0 on the patch border (a mesh boundary inside the domain), 1 inside
OVL_HD double patch_weight(int t, int s) { return (t >= 1 && t <= 4 && s >= 1 && s <= 4) ? 1.0 : 0.0; }

// rho of the patch mesh at a piece vertex (rho of mesh 0 is 1 - that, computed as r0 / (r0 + r1) like the generator)
OVL_HD void piece_rho(const Cuts &c, int q, int p, P2 v, double &rho0, double &rho1)
{
    const P2 coord[4] = {c.C00, c.C10, c.C11, c.C01};
    const double w[4] = {patch_weight(q, p), patch_weight(q, p + 1), patch_weight(q + 1, p + 1), patch_weight(q + 1, p)};
    double N[4];
    inv_quad_N(coord, v, N);
    double sum = N[0] * w[0];                  // numpy sums the 4 products left to right
    for (int k = 1; k < 4; k++) sum += N[k] * w[k];
    const double r0 = 1.0, r1 = 9.0 * sum;
    rho0 = r0 / (r0 + r1);
    rho1 = r1 / (r0 + r1);
}

// the four pieces of a patch element: piece k has vertices piece_vertex(c, k, 0..3) and mesh-0 element (jp-1+dy, ip-1+dx)
OVL_HD P2 piece_vertex(const Cuts &c, int k, int v)
{
    switch (4 * k + v) {
    case 0: return c.C00; case 1: return c.IS; case 2: return c.Pn; case 3: return c.IW;
    case 4: return c.IS; case 5: return c.C10; case 6: return c.IE; case 7: return c.Pn;
    case 8: return c.Pn; case 9: return c.IE; case 10: return c.C11; case 11: return c.IN;
    case 12: return c.IW; case 13: return c.Pn; case 14: return c.IN; default: return c.C01;
    }
}

// overlay_gen._quad_triangles: vertex tables of the two triangles of a convex 4-gon (U = max y, then min x)
OVL_HD void quad_triangles(const P2 *poly, int tri[2][3])
{
    double ymax = poly[0].y;
    for (int k = 1; k < 4; k++) if (poly[k].y > ymax) ymax = poly[k].y;
    int u = -1;
    double best = INFINITY;
    for (int k = 0; k < 4; k++) {
        const double cand = (poly[k].y == ymax) ? poly[k].x : INFINITY;
        if (u < 0 || cand < best) { u = k; best = cand; }
    }
    tri[0][0] = (u + 1) & 3; tri[0][1] = (u + 2) & 3; tri[0][2] = (u + 3) & 3;
    tri[1][0] = (u + 3) & 3; tri[1][1] = u; tri[1][2] = (u + 1) & 3;
}

// ---- the reference's triangulation rule on small polygons ---------------------------------------------------------
// planeSweepMeshes.point_2D.__gt__: higher y first, then smaller x
OVL_HD bool point_gt(P2 a, P2 b) { return a.y != b.y ? a.y > b.y : a.x < b.x; }

OVL_HD double is_left(P2 p1, P2 p2, P2 p3)
{
    return ((p2.x - p1.x) * (p3.y - p1.y) - (p2.y - p1.y) * (p3.x - p1.x)) /
           sqrt((p2.y - p1.y) * (p2.y - p1.y) + (p2.x - p1.x) * (p2.x - p1.x));
}

// triangulation.isInPolygon:151-171 on the closed triangle (t0, t1, t2, t0); non-zero = inside or on the boundary
OVL_HD int in_triangle(P2 pt, P2 t0, P2 t1, P2 t2)
{
    const P2 poly[4] = {t0, t1, t2, t0};
    int w = 0;
    for (int i = 0; i < 3; i++) {
        const P2 a = poly[i], b = poly[i + 1];
        if ((a.x == pt.x && a.y == pt.y) || (b.x == pt.x && b.y == pt.y)) return 1;
        if (a.y == pt.y && pt.y == b.y && (pt.x - a.x) * (pt.x - b.x) <= 0) return 1;
        if (a.y <= pt.y) {
            if (b.y > pt.y) {
                const double v = is_left(a, b, pt);
                if (v > 0) w += 1;
                if (v == 0) return 1;
            }
        } else if (b.y <= pt.y) {
            const double v = is_left(a, b, pt);
            if (v < 0) w -= 1;
            if (v == 0) return 1;
        }
    }
    return w;
}

// overlay_gen.ear_triangulate (= triangulation.simpleTriangulation:49-139) for polygons of <= 6 vertices, without
// recursion: a work stack of sub-polygons; a pushed triangle is emitted when popped.  Triangles come out in the
// reference's order and vertex rotation.  Returns the number of triangles (n - 2).
constexpr int EAR_MAX = 6;
OVL_HD int ear_triangulate(const P2 *pts, int n, int tri[][3])
{
    struct Item { int n; int8_t id[EAR_MAX + 1]; };
    Item stack[8];
    int top = 0, nt = 0;
    stack[0].n = n;
    for (int k = 0; k < n; k++) stack[0].id[k] = (int8_t)k;
    top = 1;
    while (top > 0) {
        const Item it = stack[--top];
        if (it.n == 3) { tri[nt][0] = it.id[0]; tri[nt][1] = it.id[1]; tri[nt][2] = it.id[2]; nt++; continue; }
        int u = 0;
        for (int k = 1; k < it.n; k++) if (point_gt(pts[it.id[k]], pts[it.id[u]])) u = k;
        int8_t ring[EAR_MAX + 1];
        for (int k = 0; k < it.n; k++) ring[k] = it.id[(u + k) % it.n];
        const int U = ring[0], nxt = ring[1], prv = ring[it.n - 1];
        int best_k = -1;
        double best_d = 0.0;
        for (int k = 2; k < it.n - 1; k++) {
            if (in_triangle(pts[ring[k]], pts[U], pts[nxt], pts[prv])) {
                const double d = is_left(pts[nxt], pts[prv], pts[ring[k]]);
                if (best_k < 0 || d > best_d) { best_k = k; best_d = d; }
            }
        }
        // result = t2 + t1: push t1 first so that t2 is worked off (and emitted) first
        if (best_k < 0) {
            Item t1; t1.n = 3; t1.id[0] = (int8_t)prv; t1.id[1] = (int8_t)U; t1.id[2] = (int8_t)nxt;
            Item t2; t2.n = it.n - 1;
            for (int k = 1; k < it.n; k++) t2.id[k - 1] = ring[k];
            stack[top++] = t1; stack[top++] = t2;
        } else {
            Item t1; t1.n = best_k + 1;
            for (int k = 0; k <= best_k; k++) t1.id[k] = ring[k];
            Item t2; t2.n = it.n - best_k + 1;
            for (int k = best_k; k < it.n; k++) t2.id[k - best_k] = ring[k];
            t2.id[it.n - best_k] = (int8_t)U;
            stack[top++] = t1; stack[top++] = t2;
        }
    }
    return nt;
}

// ring cell r (0..19) of patch (aa, bb): its mesh-0 element, polygon (5 or 6 vertices) -- order of overlay_gen._ring_row
OVL_HD int ring_polygon(const Meshes &m, int bb, int aa, int r, P2 *poly, int64_t &elem)
{
    const int j1 = PERIOD * bb + 1, i1 = PERIOD * aa + 1;
    int u, v, kind;      // kind 0 left, 1 right, 2 bottom, 3 top, 4..7 corners
    if (r < 8) { v = 1 + r / 2; kind = r & 1; u = kind ? 5 : 0; }
    else if (r < 16) { u = 1 + (r - 8) / 2; kind = 2 + ((r - 8) & 1); v = kind == 3 ? 5 : 0; }
    else { kind = 4 + (r - 16); u = (kind == 5 || kind == 6) ? 5 : 0; v = (kind == 6 || kind == 7) ? 5 : 0; }
    const P2 n00 = m.n0(j1 + v, i1 + u), n10 = m.n0(j1 + v, i1 + u + 1), n11 = m.n0(j1 + v + 1, i1 + u + 1), n01 = m.n0(j1 + v + 1, i1 + u);
    elem = (int64_t)(j1 + v) * m.nx + i1 + u;
    auto iw = [&](int q) { return element_cuts(m, bb, aa, q, 0).IW; };
    auto ie = [&](int q) { return element_cuts(m, bb, aa, q, 4).IE; };
    auto is_ = [&](int p) { return element_cuts(m, bb, aa, 0, p).IS; };
    auto in_ = [&](int p) { return element_cuts(m, bb, aa, 4, p).IN; };
    switch (kind) {
    case 0: poly[0] = n00; poly[1] = iw(v - 1); poly[2] = m.n1(bb, aa, v, 0); poly[3] = iw(v); poly[4] = n01; return 5;
    case 1: poly[0] = ie(v - 1); poly[1] = n10; poly[2] = n11; poly[3] = ie(v); poly[4] = m.n1(bb, aa, v, 5); return 5;
    case 2: poly[0] = n00; poly[1] = n10; poly[2] = is_(u); poly[3] = m.n1(bb, aa, 0, u); poly[4] = is_(u - 1); return 5;
    case 3: poly[0] = in_(u - 1); poly[1] = m.n1(bb, aa, 5, u); poly[2] = in_(u); poly[3] = n11; poly[4] = n01; return 5;
    case 4: poly[0] = n00; poly[1] = n10; poly[2] = is_(0); poly[3] = m.n1(bb, aa, 0, 0); poly[4] = iw(0); poly[5] = n01; return 6;
    case 5: poly[0] = n00; poly[1] = n10; poly[2] = n11; poly[3] = ie(0); poly[4] = m.n1(bb, aa, 0, 5); poly[5] = is_(4); return 6;
    case 6: poly[0] = ie(4); poly[1] = n10; poly[2] = n11; poly[3] = n01; poly[4] = in_(4); poly[5] = m.n1(bb, aa, 5, 5); return 6;
    default: poly[0] = n00; poly[1] = iw(4); poly[2] = m.n1(bb, aa, 5, 0); poly[3] = in_(0); poly[4] = n11; poly[5] = n01; return 6;
    }
}
// first triangle of ring cell r inside its patch's block of RING_TRIS triangles
OVL_HD int ring_tri_offset(int r) { return r < 16 ? 3 * r : 48 + 4 * (r - 16); }

// ---- untouched mesh-0 elements (regular cells), in ascending element order ------------------------------------------
// element (j, i) is touched when 1 <= j mod 8 <= 6 and 1 <= i mod 8 <= 6
OVL_HD bool elem_touched(int j, int i) { const int jj = j % PERIOD, ii = i % PERIOD; return jj >= 1 && jj <= 6 && ii >= 1 && ii <= 6; }
// untouched elements in one element row: all nx of them, or 2 per period
OVL_HD int untouched_in_row(int j, int nx) { const int jj = j % PERIOD; return (jj >= 1 && jj <= 6) ? 2 * (nx / PERIOD) : nx; }
// untouched elements in the element rows 0 .. j-1
OVL_HD int64_t untouched_before_row(int j, int nx)
{
    const int64_t per_period = 2 * (int64_t)nx + 6 * 2 * (int64_t)(nx / PERIOD);
    int64_t c = (int64_t)(j / PERIOD) * per_period;
    for (int r = 0; r < j % PERIOD; r++) c += untouched_in_row(r, nx);
    return c;
}
// rank of untouched element (j, i) among the untouched elements of its row
OVL_HD int untouched_rank_in_row(int j, int i)
{
    const int jj = j % PERIOD;
    if (!(jj >= 1 && jj <= 6)) return i;
    return 2 * (i / PERIOD) + ((i % PERIOD) == 7 ? 1 : 0);
}

// ---- the flattened overlay, one work item at a time (identical code on the host and in the kernels) -----------------
struct Layout {
    int nx, ny, Px, Py, b_lo, b_hi, n_mesh;
    int64_t n_elem0, n_reg, n_ring, n_piece, n_tri;
    OVL_HD int64_t n_cell() const { return n_reg + n_ring + n_piece; }
};
struct Outputs {
    int32_t *cell_elem;     // [n_cell][n_mesh]
    int32_t *cell_kind;     // [n_cell]
    int64_t *cell_tri_ptr;  // [n_cell + 1]
    double *tri_xy;         // [n_tri][3][2]
    double *tri_rho;        // [n_tri][3][n_mesh]
};

inline Layout make_layout(int nx, int ny, int b_lo, int b_hi, int n_mesh)
{
    Layout L;
    L.nx = nx; L.ny = ny; L.Px = nx / PERIOD; L.Py = ny / PERIOD; L.b_lo = b_lo; L.b_hi = b_hi; L.n_mesh = n_mesh;
    L.n_elem0 = (int64_t)nx * ny;
    const int64_t patches = (int64_t)(b_hi - b_lo) * L.Px;
    L.n_ring = RING_CELLS * patches;
    L.n_piece = PIECES * patches;
    const int j0 = PERIOD * b_lo - 1 > 0 ? PERIOD * b_lo - 1 : 0, j1 = PERIOD * b_hi;
    L.n_reg = 0;
    for (int j = j0; j < j1; j++) L.n_reg += untouched_in_row(j, nx);
    L.n_tri = RING_TRIS * patches + 2 * L.n_piece;
    return L;
}

// mesh slot of patch (a, b): 1, or 1 + (a + b) % 2 for the 3-mesh variant
OVL_HD int patch_slot(const Layout &L, int b, int a) { return L.n_mesh == 3 ? 1 + ((a + b) & 1) : 1; }

OVL_HD void set_cell(const Layout &L, const Outputs &o, int64_t c, int kind, int64_t e0, int slot1, int64_t e1, int64_t tri0)
{
    for (int m = 0; m < L.n_mesh; m++) o.cell_elem[c * L.n_mesh + m] = -1;
    o.cell_elem[c * L.n_mesh] = (int32_t)e0;
    if (slot1 > 0) o.cell_elem[c * L.n_mesh + slot1] = (int32_t)e1;
    o.cell_kind[c] = kind;
    o.cell_tri_ptr[c] = tri0;
}

// work item 1: patch element (q, p) of patch (a, b) -> 4 two-element cells, 8 triangles
OVL_HD void gen_patch_element(const Meshes &m, const Layout &L, const Outputs &o, int b, int a, int q, int p)
{
    const Cuts c = element_cuts(m, b, a, q, p);
    const int slot = patch_slot(L, b, a);
    const int ip = PERIOD * a + 2 + p, jp = PERIOD * b + 2 + q;
    const int64_t e1 = L.n_elem0 + 25 * ((int64_t)b * L.Px + a) + 5 * q + p;
    const int64_t first = ((((int64_t)(b - L.b_lo) * L.Px + a) * 5 + q) * 5 + p) * 4;
    const int64_t ring_tris = (int64_t)RING_TRIS * (L.b_hi - L.b_lo) * L.Px;
    const int dx[4] = {0, 1, 1, 0}, dy[4] = {0, 0, 1, 1};
    for (int k = 0; k < 4; k++) {
        P2 poly[4];
        double rho0[4], rho1[4];
        for (int v = 0; v < 4; v++) { poly[v] = piece_vertex(c, k, v); piece_rho(c, q, p, poly[v], rho0[v], rho1[v]); }
        int tri[2][3];
        quad_triangles(poly, tri);
        const int64_t piece = first + k, t0 = ring_tris + 2 * piece;
        set_cell(L, o, L.n_reg + L.n_ring + piece, 1, (int64_t)(jp - 1 + dy[k]) * L.nx + (ip - 1 + dx[k]), slot, e1, t0);
        for (int t = 0; t < 2; t++)
            for (int v = 0; v < 3; v++) {
                const int64_t at = (t0 + t) * 3 + v;
                o.tri_xy[2 * at] = poly[tri[t][v]].x; o.tri_xy[2 * at + 1] = poly[tri[t][v]].y;
                for (int s = 0; s < L.n_mesh; s++) o.tri_rho[at * L.n_mesh + s] = 0.0;
                o.tri_rho[at * L.n_mesh] = rho0[tri[t][v]];
                o.tri_rho[at * L.n_mesh + slot] = rho1[tri[t][v]];
            }
    }
}

// work item 2: ring cell r of patch (a, b) -> one single-element cell, 3 or 4 triangles
OVL_HD void gen_ring_cell(const Meshes &m, const Layout &L, const Outputs &o, int b, int a, int r)
{
    P2 poly[EAR_MAX];
    int64_t elem;
    const int n = ring_polygon(m, b, a, r, poly, elem);
    int tri[EAR_MAX - 2][3];
    const int nt = ear_triangulate(poly, n, tri);
    const int64_t pp = (int64_t)(b - L.b_lo) * L.Px + a;
    const int64_t t0 = RING_TRIS * pp + ring_tri_offset(r);
    set_cell(L, o, L.n_reg + RING_CELLS * pp + r, 1, elem, 0, 0, t0);
    for (int t = 0; t < nt; t++)
        for (int v = 0; v < 3; v++) {
            const int64_t at = (t0 + t) * 3 + v;
            o.tri_xy[2 * at] = poly[tri[t][v]].x; o.tri_xy[2 * at + 1] = poly[tri[t][v]].y;
            for (int s = 0; s < L.n_mesh; s++) o.tri_rho[at * L.n_mesh + s] = 0.0;
            o.tri_rho[at * L.n_mesh] = 1.0;
        }
}

// work item 3: mesh-0 element (j, i) of the strip -> a regular cell if no patch touches it
OVL_HD void gen_regular(const Layout &L, const Outputs &o, int j, int i, int64_t row_base)
{
    if (elem_touched(j, i)) return;
    const int64_t c = row_base + untouched_rank_in_row(j, i);
    set_cell(L, o, c, 0, (int64_t)j * L.nx + i, 0, 0, 0);
}

}  // namespace mofem_ovl
