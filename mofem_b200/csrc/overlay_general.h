// overlay_general.h -- overlay of the general patch families by polygon clipping (host code).
//
// overlay_core.h covers the 2-mesh quad4 family in closed form.  The families of BASELINE configs 3b and 4 -- patch quads
// split into 3-node triangles, and a staggered third mesh that overlaps both other meshes, so that cells have 1, 2 and 3
// incident elements and up to 8 vertices -- need real clipping.  What the reference's plane sweep (planeSweepMeshes.py) +
// polygonTriangulation + rhoCalculation + flatten() hand to the solver is rebuilt per mesh-0 element:
//   * the element's quadrilateral is cut, patch mesh by patch mesh, into the pieces inside each patch element (clipping of
//     the piece against the convex element, Sutherland-Hodgman) and the piece outside the patch (the piece minus the patch's
//     axis-aligned rectangle; the patch nodes on the rectangle side that bounds the piece become vertices of it, exactly as
//     the faces of the reference's arrangement carry every vertex on their boundary);
//   * every piece is one overlay cell with the tuple of elements it lies in; a mesh-0 element no patch touches is a regular
//     cell;
//   * triangles: the reference's recursive "cut at the up-left-most vertex" rule (triangulation.py:49-139);
//   * weights: inverse map of the vertex into every incident element, corner-weight interpolation (0 on patch-border nodes),
//     x 9 for meshes >= 1, normalisation (planeSweepMeshes.py:991-1030).
// Pinned against the reference's own sweep at n0 = 16 and 32 (tests/golden/patch16_tri, patch16_m3s, patch32_m3s: same cell
// and triangle counts, same triangle set, same CSR pattern, values to 1e-12).  Intersection points are computed per piece in
// double precision (the sweep computes each arrangement vertex once, in long double): they agree to ~1e-16 relative, which
// is what the value tolerance absorbs; coordinates on the straight patch borders are pinned exactly so that the
// triangulation rule sees the same ties.
#pragma once
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <stdexcept>
#include <vector>

namespace mofem_ovg {

constexpr int PERIOD = 8;
constexpr int MAX_MESH = 4;
struct P2 { double x, y; };
typedef std::vector<P2> Poly;

struct PatchMesh {
    int Px, Py;          // patches per direction
    int tri;             // 1: every patch quad is two 3-node triangles
    int64_t elem0;       // global id of the first element of this mesh
    const double *X;     // [Py][Px][6][6][2]
    P2 node(int b, int a, int t, int s) const
    {
        const double *p = X + 2 * ((((int64_t)b * Px + a) * 6 + t) * 6 + s);
        return P2{p[0], p[1]};
    }
    int elems_per_patch() const { return tri ? 50 : 25; }
};

struct Input {
    int nx, ny;
    const double *X0;    // [ny+1][nx+1][2]
    int n_patch_mesh;
    PatchMesh pm[MAX_MESH - 1];
    int n_mesh() const { return 1 + n_patch_mesh; }
    P2 n0(int j, int i) const { const double *p = X0 + 2 * ((int64_t)j * (nx + 1) + i); return P2{p[0], p[1]}; }
};

struct Output {          // appended to, cell by cell
    std::vector<int32_t> cell_elem;   // [n_cell][n_mesh]
    std::vector<int32_t> cell_kind;
    std::vector<int64_t> cell_ntri;
    std::vector<double> tri_xy;       // [n_tri][3][2]
    std::vector<double> tri_rho;      // [n_tri][3][n_mesh]
};

// ---- small geometry ------------------------------------------------------------------------------------------------------
inline double cross(P2 a, P2 b, P2 c) { return (b.x - a.x) * (c.y - a.y) - (b.y - a.y) * (c.x - a.x); }
inline double area2(const Poly &p)
{
    double s = 0.0;
    for (size_t k = 0, n = p.size(); k < n; k++) { const P2 a = p[k], b = p[(k + 1) % n]; s += a.x * b.y - a.y * b.x; }
    return s;
}
struct Box { double x0, y0, x1, y1; };
inline Box bbox(const Poly &p)
{
    Box b{p[0].x, p[0].y, p[0].x, p[0].y};
    for (const P2 &v : p) { b.x0 = std::min(b.x0, v.x); b.x1 = std::max(b.x1, v.x); b.y0 = std::min(b.y0, v.y); b.y1 = std::max(b.y1, v.y); }
    return b;
}
inline bool overlap(const Box &a, const Box &b) { return a.x0 < b.x1 && b.x0 < a.x1 && a.y0 < b.y1 && b.y0 < a.y1; }

inline void dedupe(Poly &p, double tol)
{
    Poly q;
    for (const P2 &v : p)
        if (q.empty() || fabs(v.x - q.back().x) > tol || fabs(v.y - q.back().y) > tol) q.push_back(v);
    while (q.size() > 1 && fabs(q.front().x - q.back().x) <= tol && fabs(q.front().y - q.back().y) <= tol) q.pop_back();
    p.swap(q);
}

// subject (simple polygon, CCW, may be non-convex) clipped by a convex CCW polygon.  Intersections with an axis-aligned
// clip edge get that edge's coordinate exactly; vertices that land on a corner of the clipper take the corner's bits.
inline Poly clip_convex(const Poly &subject, const P2 *clipper, int nc, double tol)
{
    Poly cur = subject;
    for (int k = 0; k < nc && !cur.empty(); k++) {
        const P2 a = clipper[k], b = clipper[(k + 1) % nc];
        Poly out;
        const size_t n = cur.size();
        for (size_t i = 0; i < n; i++) {
            const P2 s = cur[i], e = cur[(i + 1) % n];
            const double ds = cross(a, b, s), de = cross(a, b, e);
            const bool in_s = ds >= 0.0, in_e = de >= 0.0;
            if (in_s) out.push_back(s);
            if (in_s != in_e) {
                const double t = ds / (ds - de);
                P2 x{s.x + t * (e.x - s.x), s.y + t * (e.y - s.y)};
                if (a.x == b.x) x.x = a.x;
                if (a.y == b.y) x.y = a.y;
                out.push_back(x);
            }
        }
        cur.swap(out);
    }
    for (P2 &v : cur)
        for (int k = 0; k < nc; k++)
            if (fabs(v.x - clipper[k].x) <= tol && fabs(v.y - clipper[k].y) <= tol) v = clipper[k];
    dedupe(cur, 0.0);
    return cur;
}

// ---- polygon minus the axis-aligned rectangle of a patch ---------------------------------------------------------------
struct Rect {
    double x0, y0, x1, y1;
    P2 border[20];       // the patch nodes on the rectangle, any order
    bool strictly_inside(P2 p) const { return p.x > x0 && p.x < x1 && p.y > y0 && p.y < y1; }
    // clockwise perimeter parameter in [0, 4): left side upwards, top rightwards, right downwards, bottom leftwards
    double param(P2 p, double tol) const
    {
        if (fabs(p.x - x0) <= tol && p.y < y1 - tol) return 0.0 + (p.y - y0) / (y1 - y0);
        if (fabs(p.y - y1) <= tol && p.x < x1 - tol) return 1.0 + (p.x - x0) / (x1 - x0);
        if (fabs(p.x - x1) <= tol && p.y > y0 + tol) return 2.0 + (y1 - p.y) / (y1 - y0);
        if (fabs(p.y - y0) <= tol) return 3.0 + (x1 - p.x) / (x1 - x0);
        throw std::runtime_error("overlay: point not on the patch rectangle");
    }
};

// the part of segment s-e inside the closed rectangle as a parameter interval (Liang-Barsky); false if empty
inline bool seg_in_rect(P2 s, P2 e, const Rect &r, double &t0, double &t1)
{
    t0 = 0.0; t1 = 1.0;
    const double dx = e.x - s.x, dy = e.y - s.y;
    const double p[4] = {-dx, dx, -dy, dy}, q[4] = {s.x - r.x0, r.x1 - s.x, s.y - r.y0, r.y1 - s.y};
    for (int k = 0; k < 4; k++) {
        if (p[k] == 0.0) { if (q[k] < 0.0) return false; }
        else {
            const double t = q[k] / p[k];
            if (p[k] < 0.0) { if (t > t1) return false; if (t > t0) t0 = t; }
            else { if (t < t0) return false; if (t < t1) t1 = t; }
        }
    }
    return t1 > t0;
}
inline P2 on_rect(P2 s, P2 e, double t, const Rect &r, double tol)
{
    P2 x{s.x + t * (e.x - s.x), s.y + t * (e.y - s.y)};
    if (fabs(x.x - r.x0) <= tol) x.x = r.x0;
    if (fabs(x.x - r.x1) <= tol) x.x = r.x1;
    if (fabs(x.y - r.y0) <= tol) x.y = r.y0;
    if (fabs(x.y - r.y1) <= tol) x.y = r.y1;
    return x;
}

// P \ rect for a simple CCW polygon P that the rectangle neither splits nor lies inside of.  Returns false when nothing of P
// is left; `changed` tells whether the rectangle took anything away.
inline bool minus_rect(const Poly &P, const Rect &r, double tol, Poly &out, bool &changed)
{
    const size_t n = P.size();
    changed = false;
    size_t n_in = 0;
    for (const P2 &v : P) n_in += r.strictly_inside(v);
    // walk the boundary; record the outside chain that starts at the exit point and ends at the entry point
    Poly chain;
    P2 enter{0, 0}, leave{0, 0};
    int n_enter = 0, n_leave = 0;
    std::vector<Poly> chains;
    Poly cur;
    bool outside = false, started = false;
    // start the walk at a vertex that is not strictly inside
    size_t k0 = n;
    for (size_t k = 0; k < n; k++) if (!r.strictly_inside(P[k])) { k0 = k; break; }
    if (k0 == n) { out.clear(); changed = true; return false; }   // all of P inside the rectangle
    (void)chain; (void)started;
    outside = true;
    cur.push_back(P[k0]);
    for (size_t step = 0; step < n; step++) {
        const P2 s = P[(k0 + step) % n], e = P[(k0 + step + 1) % n];
        double t0, t1;
        bool through = seg_in_rect(s, e, r, t0, t1);
        if (through) {     // a portion that merely runs along a side does not count
            const double tm = 0.5 * (t0 + t1);
            through = r.strictly_inside(P2{s.x + tm * (e.x - s.x), s.y + tm * (e.y - s.y)});
        }
        if (!through) {
            if (!outside) throw std::runtime_error("overlay: lost track of the rectangle");
            cur.push_back(e);
            continue;
        }
        changed = true;
        if (outside) {     // entering at t0
            const P2 en = on_rect(s, e, t0, r, tol);
            cur.push_back(en);
            enter = en; n_enter++;
            chains.push_back(cur);
            cur.clear();
            outside = false;
        }
        if (!r.strictly_inside(e)) {   // leaving at t1
            const P2 lv = on_rect(s, e, t1, r, tol);
            leave = lv; n_leave++;
            cur.clear();
            cur.push_back(lv);
            cur.push_back(e);
            outside = true;
        }
    }
    if (!changed) { out = P; return true; }
    if (n_enter != 1 || n_leave != 1 || !outside) throw std::runtime_error("overlay: a patch rectangle cuts a piece more than once");
    // the walk ended back at P[k0] on the chain that began at the exit point; the first chain began at P[k0]
    Poly res = cur;                                  // leave ... P[k0]
    res.pop_back();                                  // P[k0] is the head of chains[0]
    res.insert(res.end(), chains[0].begin(), chains[0].end());   // ... enter
    // back from `enter` to `leave` clockwise along the rectangle, picking up the patch nodes in between
    const double ue = r.param(enter, tol), ul = r.param(leave, tol);
    double span = ul - ue;
    if (span < 0.0) span += 4.0;
    if (span > 2.0) throw std::runtime_error("overlay: rectangle chain runs the wrong way round");
    std::vector<std::pair<double, int>> mid;
    for (int k = 0; k < 20; k++) {
        double d = r.param(r.border[k], tol) - ue;
        if (d < 0.0) d += 4.0;
        if (d > 1e-9 && d < span - 1e-9) mid.push_back({d, k});
    }
    std::sort(mid.begin(), mid.end());
    for (const auto &m : mid) res.push_back(r.border[m.second]);
    dedupe(res, 0.0);
    out.swap(res);
    return out.size() >= 3 && area2(out) > 0.0;
}

// ---- the reference's triangulation rule (overlay_gen.ear_triangulate = triangulation.simpleTriangulation:49-139) ------------
inline bool point_gt(P2 a, P2 b) { return a.y != b.y ? a.y > b.y : a.x < b.x; }
inline double is_left(P2 p1, P2 p2, P2 p3)
{
    return ((p2.x - p1.x) * (p3.y - p1.y) - (p2.y - p1.y) * (p3.x - p1.x)) /
           sqrt((p2.y - p1.y) * (p2.y - p1.y) + (p2.x - p1.x) * (p2.x - p1.x));
}
inline int in_triangle(P2 pt, P2 t0, P2 t1, P2 t2)     // triangulation.isInPolygon:151-171 on (t0, t1, t2, t0)
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
inline void ear(const Poly &pts, const std::vector<int> &ids, std::vector<int> &tris)
{
    const int n = (int)ids.size();
    if (n == 3) { tris.insert(tris.end(), ids.begin(), ids.end()); return; }
    int u = 0;
    for (int k = 1; k < n; k++) if (point_gt(pts[ids[k]], pts[ids[u]])) u = k;
    std::vector<int> ring(n);
    for (int k = 0; k < n; k++) ring[k] = ids[(u + k) % n];
    const int U = ring[0], nxt = ring[1], prv = ring[n - 1];
    int best_k = -1;
    double best_d = 0.0;
    for (int k = 2; k < n - 1; k++)
        if (in_triangle(pts[ring[k]], pts[U], pts[nxt], pts[prv])) {
            const double d = is_left(pts[nxt], pts[prv], pts[ring[k]]);
            if (best_k < 0 || d > best_d) { best_k = k; best_d = d; }     // the reference keeps the FIRST maximum
        }
    std::vector<int> t1, t2;
    if (best_k < 0) {
        t1 = {prv, U, nxt};
        t2.assign(ring.begin() + 1, ring.end());
    } else {
        t1.assign(ring.begin(), ring.begin() + best_k + 1);
        t2.assign(ring.begin() + best_k, ring.end());
        t2.push_back(U);
    }
    ear(pts, t2, tris);       // result = t2 + t1
    ear(pts, t1, tris);
}

// ---- weights --------------------------------------------------------------------------------------------------------------
// shape functions of a quad4 at the inverse image of xy (overlay_gen._inv_quad_N: 6 Newton updates from (0, 0))
inline void inv_quad_N(const P2 *c, P2 xy, double *N)
{
    const double X1 = c[1].x - c[0].x, X2 = c[2].x - c[0].x, X3 = c[3].x - c[0].x;
    const double Y1 = c[1].y - c[0].y, Y2 = c[2].y - c[0].y, Y3 = c[3].y - c[0].y;
    const double x = xy.x - c[0].x, y = xy.y - c[0].y;
    double r = 0.0, s = 0.0;
    for (int it = 0; it < 6; it++) {
        const double rm = 1.0 - r, rp = 1.0 + r, sm = 1.0 - s, sp = 1.0 + s;
        const double N1 = 0.25 * rp * sm, N2 = 0.25 * rp * sp, N3 = 0.25 * rm * sp;
        const double fx = x - ((N1 * X1 + N2 * X2) + N3 * X3), fy = y - ((N1 * Y1 + N2 * Y2) + N3 * Y3);
        const double xr = 0.25 * (sm * X1 + sp * (X2 - X3)), xs = 0.25 * (rm * X3 + rp * (X2 - X1));
        const double yr = 0.25 * (sm * Y1 + sp * (Y2 - Y3)), ys = 0.25 * (rm * Y3 + rp * (Y2 - Y1));
        const double det = xr * ys - xs * yr;
        r = r + (fx * ys - fy * xs) / det;
        s = s + (fy * xr - fx * yr) / det;
    }
    const double rm = 1.0 - r, rp = 1.0 + r, sm = 1.0 - s, sp = 1.0 + s;
    N[0] = 0.25 * rm * sm; N[1] = 0.25 * rp * sm; N[2] = 0.25 * rp * sp; N[3] = 0.25 * rm * sp;
}
// area coordinates of a 3-node triangle (invShapeFunction.invTri:21-39)
inline void inv_tri_N(const P2 *c, P2 xy, double *N)
{
    const double area = (c[1].x * c[2].y - c[2].x * c[1].y - c[0].x * (c[2].y - c[1].y) + c[0].y * (c[2].x - c[1].x)) / 2;
    const double a3[3] = {c[1].x * c[2].y - c[2].x * c[1].y, c[2].x * c[0].y - c[0].x * c[2].y, c[0].x * c[1].y - c[1].x * c[0].y};
    const double b3[3] = {c[1].y - c[2].y, c[2].y - c[0].y, c[0].y - c[1].y};
    const double c3[3] = {c[2].x - c[1].x, c[0].x - c[2].x, c[1].x - c[0].x};
    for (int i = 0; i < 3; i++) N[i] = (a3[i] + b3[i] * xy.x + c3[i] * xy.y) / 2.0 / area;
}

// a patch element: its corners (CCW), their weights (0 on the patch border) and its global id
struct PatchElem { P2 c[4]; double w[4]; int nc; int64_t id; };
inline double patch_weight(int t, int s) { return (t >= 1 && t <= 4 && s >= 1 && s <= 4) ? 1.0 : 0.0; }
inline PatchElem patch_elem(const PatchMesh &m, int b, int a, int q, int p, int half)
{
    // corners of quad (q, p): q0 = (q, p), q1 = (q, p+1), q2 = (q+1, p+1), q3 = (q+1, p)  as (row t, column s)
    const int T[4] = {q, q, q + 1, q + 1}, S[4] = {p, p + 1, p + 1, p};
    PatchElem e;
    const int64_t patch = (int64_t)b * m.Px + a;
    if (!m.tri) {
        e.nc = 4;
        for (int k = 0; k < 4; k++) { e.c[k] = m.node(b, a, T[k], S[k]); e.w[k] = patch_weight(T[k], S[k]); }
        e.id = m.elem0 + 25 * patch + 5 * q + p;
    } else {
        // (q0, q1, q2), (q0, q2, q3); on the anti-diagonal p + q == 4: (q0, q1, q3), (q1, q2, q3)   (overlay_gen.general_meshes)
        const bool anti = (p + q) == 4;
        const int idx[2][2][3] = {{{0, 1, 2}, {0, 2, 3}}, {{0, 1, 3}, {1, 2, 3}}};
        e.nc = 3;
        for (int k = 0; k < 3; k++) { const int v = idx[anti][half][k]; e.c[k] = m.node(b, a, T[v], S[v]); e.w[k] = patch_weight(T[v], S[v]); }
        e.id = m.elem0 + 50 * patch + 2 * (5 * q + p) + half;
    }
    return e;
}

struct Piece {
    Poly poly;
    int64_t elem[MAX_MESH];
    PatchElem pe[MAX_MESH];     // pe[m] valid when elem[m] >= 0, m >= 1
};

// every cell of mesh-0 element (j, i), appended to `o`
inline void gen_element(const Input &in, int j, int i, Output &o)
{
    const int M = in.n_mesh();
    const double h = 1.0 / in.nx, tol = 1e-9 * h;
    std::vector<Piece> pieces(1);
    pieces[0].poly = {in.n0(j, i), in.n0(j, i + 1), in.n0(j + 1, i + 1), in.n0(j + 1, i)};
    for (int m = 0; m < MAX_MESH; m++) pieces[0].elem[m] = -1;
    pieces[0].elem[0] = (int64_t)j * in.nx + i;
    const Box eb = bbox(pieces[0].poly);
    bool touched = false;
    for (int pmi = 0; pmi < in.n_patch_mesh; pmi++) {
        const PatchMesh &pm = in.pm[pmi];
        const int m = pmi + 1;
        // the (at most one) patch of this mesh whose rectangle reaches the element: patches are 5h wide, 3h apart
        int pa = -1, pb = -1;
        Rect r{};
        for (int b = 0; b < pm.Py && pa < 0; b++) {
            const double y0 = pm.node(b, 0, 0, 0).y, y1 = pm.node(b, 0, 5, 0).y;
            if (!(eb.y0 < y1 && y0 < eb.y1)) continue;
            for (int a = 0; a < pm.Px; a++) {
                const double x0 = pm.node(b, a, 0, 0).x, x1 = pm.node(b, a, 0, 5).x;
                if (eb.x0 < x1 && x0 < eb.x1) { pa = a; pb = b; r.x0 = x0; r.x1 = x1; r.y0 = y0; r.y1 = y1; break; }
            }
        }
        if (pa < 0) continue;
        {
            int nb = 0;
            for (int s = 0; s < 6; s++) { r.border[nb++] = pm.node(pb, pa, 0, s); r.border[nb++] = pm.node(pb, pa, 5, s); }
            for (int t = 1; t < 5; t++) { r.border[nb++] = pm.node(pb, pa, t, 0); r.border[nb++] = pm.node(pb, pa, t, 5); }
        }
        const Box rb{r.x0, r.y0, r.x1, r.y1};
        std::vector<Piece> next;
        for (const Piece &pc : pieces) {
            const Box pbx = bbox(pc.poly);
            if (!overlap(pbx, rb)) { next.push_back(pc); continue; }
            // inside the patch: one piece per patch element it reaches
            bool any_inside = false;
            for (int q = 0; q < 5; q++)
                for (int p = 0; p < 5; p++) {
                    const Box qb{std::min(pm.node(pb, pa, q, p).x, pm.node(pb, pa, q + 1, p).x), std::min(pm.node(pb, pa, q, p).y, pm.node(pb, pa, q, p + 1).y),
                                 std::max(pm.node(pb, pa, q, p + 1).x, pm.node(pb, pa, q + 1, p + 1).x), std::max(pm.node(pb, pa, q + 1, p).y, pm.node(pb, pa, q + 1, p + 1).y)};
                    if (!overlap(pbx, qb)) continue;
                    for (int half = 0; half < (pm.tri ? 2 : 1); half++) {
                        const PatchElem pe = patch_elem(pm, pb, pa, q, p, half);
                        Poly c = clip_convex(pc.poly, pe.c, pe.nc, tol);
                        if (c.size() < 3 || area2(c) <= 1e-12 * h * h) continue;
                        Piece np = pc;
                        np.poly.swap(c);
                        np.elem[m] = pe.id;
                        np.pe[m] = pe;
                        next.push_back(np);
                        any_inside = true;
                    }
                }
            // outside the patch
            Poly rest;
            bool changed;
            if (minus_rect(pc.poly, r, tol, rest, changed) && area2(rest) > 1e-12 * h * h) {
                Piece np = pc;
                np.poly.swap(rest);
                next.push_back(np);
            }
            if (any_inside || changed) touched = true;
        }
        pieces.swap(next);
    }
    if (!touched) {       // regular cell (quadFE on the mesh-0 element)
        for (int m = 0; m < M; m++) o.cell_elem.push_back(m == 0 ? (int32_t)pieces[0].elem[0] : -1);
        o.cell_kind.push_back(0);
        o.cell_ntri.push_back(0);
        return;
    }
    for (const Piece &pc : pieces) {
        const int nv = (int)pc.poly.size();
        // weights at the vertices
        std::vector<double> rho((size_t)nv * M, 0.0);
        for (int v = 0; v < nv; v++) {
            double rr[MAX_MESH] = {0, 0, 0, 0};
            rr[0] = 1.0;      // mesh 0: every node weight is 1 and the shape functions sum to 1
            double sum = rr[0];
            for (int m = 1; m < M; m++) {
                if (pc.elem[m] < 0) continue;
                double N[4];
                double s = 0.0;
                if (pc.pe[m].nc == 4) inv_quad_N(pc.pe[m].c, pc.poly[v], N); else inv_tri_N(pc.pe[m].c, pc.poly[v], N);
                s = N[0] * pc.pe[m].w[0];
                for (int k = 1; k < pc.pe[m].nc; k++) s += N[k] * pc.pe[m].w[k];
                rr[m] = 9.0 * s;
                sum += rr[m];
            }
            for (int m = 0; m < M; m++) rho[(size_t)v * M + m] = (m == 0 || pc.elem[m] >= 0) ? rr[m] / sum : 0.0;
        }
        std::vector<int> ids(nv), tris;
        for (int v = 0; v < nv; v++) ids[v] = v;
        ear(pc.poly, ids, tris);
        for (int m = 0; m < M; m++) o.cell_elem.push_back((int32_t)pc.elem[m]);
        o.cell_kind.push_back(1);
        o.cell_ntri.push_back((int64_t)tris.size() / 3);
        for (int v : tris) {
            o.tri_xy.push_back(pc.poly[v].x); o.tri_xy.push_back(pc.poly[v].y);
            for (int m = 0; m < M; m++) o.tri_rho.push_back(rho[(size_t)v * M + m]);
        }
    }
}

}  // namespace mofem_ovg
