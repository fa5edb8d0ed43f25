// integrate.cu -- block integration kernels (fp64).
//
// One task = one block the reference emits (AMORE.py:59-123 / 328-398 / 751-821):
//   * regular element stiffness (triFE / quadFE / ICMFE / triQuadFE / quadQuadFE,
//     stiffnessMatrix.py:9-81), or
//   * AMORE diagonal block (element j with itself) or off-diagonal block (j,l) summed over the
//     cell's triangles (stiffnessMatrix.py:99-190).
// Tasks are bucketed by class (element types, curved) so that a warp is homogeneous.
//
// Arithmetic: Kloc = sum_t sum_q (0.5 detJ_t w_q) B1~^T D B2~.  B~ has two distinct numbers per node
// (gx, gy), so each 2x2 node-pair sub-block is D applied to the four sums
//   S = sum w * [gx_a gx_b, gx_a gy_b, gy_a gx_b, gy_a gy_b];
// 4 FMAs per (node pair, point) instead of the reference's dense 3-row products.
//
// Output layout (blk_val): per task, n1*n2 node pairs row-major, 4 doubles each
//   [K(2a,2b), K(2a,2b+1), K(2a+1,2b), K(2a+1,2b+1)]   (one 32-byte sector per node pair).
//
// Compiled with -fmad=false (see fem_device.cuh); accumulations use explicit fma().
#include "common.cuh"
#include "fem_device.cuh"

namespace mofem {

// quadrature tables: otherFunctions/numericalIntegration.py:5-134 (literals verbatim)
#define T13 (1.0 / 3.0)
__constant__ double c_tp[71 * 3] = {
    // n=1
    T13, T13, T13,
    // n=3
    0.5, 0.5, 0.0, 0.5, 0.0, 0.5, 0.0, 0.5, 0.5,
    // n=4
    T13, T13, T13, 3.0 / 5, 1.0 / 5, 1.0 / 5, 1.0 / 5, 3.0 / 5, 1.0 / 5, 1.0 / 5, 1.0 / 5, 3.0 / 5,
    // n=6
    0.816847572980459, 0.091576213509771, 0.091576213509771,
    0.091576213509771, 0.816847572980459, 0.091576213509771,
    0.091576213509771, 0.091576213509771, 0.816847572980459,
    0.108103018168070, 0.445948490915965, 0.445948490915965,
    0.445948490915965, 0.108103018168070, 0.445948490915965,
    0.445948490915965, 0.445948490915965, 0.108103018168070,
    // n=7
    T13, T13, T13,
    0.797426985353087, 0.101286507323457, 0.101286507323457,
    0.101286507323457, 0.797426985353087, 0.101286507323457,
    0.101286507323457, 0.101286507323457, 0.797426985353087,
    0.059715871789770, 0.470142064105115, 0.470142064105115,
    0.470142064105115, 0.059715871789770, 0.470142064105115,
    0.470142064105115, 0.470142064105115, 0.059715871789770,
    // n=9
    0.124949503233232, 0.437525248383384, 0.437525248383384,
    0.437525248383384, 0.124949503233232, 0.437525248383384,
    0.437525248383384, 0.437525248383384, 0.124949503233232,
    0.797112651860071, 0.165409927389841, 0.037477420750088,
    0.165409927389841, 0.797112651860071, 0.037477420750088,
    0.037477420750088, 0.797112651860071, 0.165409927389841,
    0.037477420750088, 0.165409927389841, 0.797112651860071,
    0.797112651860071, 0.037477420750088, 0.165409927389841,
    0.165409927389841, 0.037477420750088, 0.797112651860071,
    // n=12
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
    0.310352451033785, 0.053145049844816, 0.636502499121399,
    // n=13
    T13, T13, T13,
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
    0.048690315425316, 0.312865496004874, 0.638444188569810,
    // n=16
    T13, T13, T13,
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
    0.728492392955404, 0.263112829634638, 0.008394777409958};
__constant__ double c_tw[71] = {
    1.0,
    T13, T13, T13,
    -0.562500000000000, 0.520833333333333, 0.520833333333333, 0.520833333333333,
    0.109951743655322, 0.109951743655322, 0.109951743655322, 0.223381589678011, 0.223381589678011, 0.223381589678011,
    0.225000000000000, 0.125939180544827, 0.125939180544827, 0.125939180544827,
    0.132394152788506, 0.132394152788506, 0.132394152788506,
    0.205950504760887, 0.205950504760887, 0.205950504760887,
    0.063691414286223, 0.063691414286223, 0.063691414286223, 0.063691414286223, 0.063691414286223, 0.063691414286223,
    0.050844906370207, 0.050844906370207, 0.050844906370207, 0.116786275726379, 0.116786275726379, 0.116786275726379,
    0.082851075618374, 0.082851075618374, 0.082851075618374, 0.082851075618374, 0.082851075618374, 0.082851075618374,
    -0.149570044467682, 0.175615257433208, 0.175615257433208, 0.175615257433208,
    0.053347235608838, 0.053347235608838, 0.053347235608838,
    0.077113760890257, 0.077113760890257, 0.077113760890257, 0.077113760890257, 0.077113760890257, 0.077113760890257,
    0.144315607677787, 0.095091634267285, 0.095091634267285, 0.095091634267285,
    0.103217370534718, 0.103217370534718, 0.103217370534718,
    0.032458497623198, 0.032458497623198, 0.032458497623198,
    0.027230314174435, 0.027230314174435, 0.027230314174435, 0.027230314174435, 0.027230314174435, 0.027230314174435};
// gaussQuad n=2 and n=3
__constant__ double c_q2p[2] = {-0.577350269189626, 0.577350269189626};
__constant__ double c_q3p[3] = {-0.774596669241483, 0.0, 0.774596669241483};
__constant__ double c_q3w[3] = {5.0 / 9, 8.0 / 9, 5.0 / 9};

// ---------------------------------------------------------------------------------------
// task construction
// ---------------------------------------------------------------------------------------

__global__ void k_cell_counts(DevProblem p, int32_t *__restrict__ ntask, int32_t *__restrict__ npair,
                              int32_t *__restrict__ ncoo, int *err_flag)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= p.n_cell) return;
    const int M = p.n_mesh;
    const int32_t *ce = p.cell_elem + c * M;
    int nt = 0, np = 0, nc = 0;
    if (p.cell_kind[c] == CELL_REGULAR) {
        for (int j = 0; j < M; j++)
            if (ce[j] >= 0) {
                int n = p.elem_nn[ce[j]];
                if (p.solver == 1 && n != 3 && n != 4) *err_flag = MOFEM_E_ELEMENT;
                nt = 1; np = n * n; nc = 4 * n * n;
                break;
            }
    } else {
        for (int j = 0; j < M; j++) {
            if (ce[j] < 0) continue;
            int n1 = p.elem_nn[ce[j]];
            if (p.solver == 1 && n1 != 3 && n1 != 4) *err_flag = MOFEM_E_ELEMENT;
            nt++; np += n1 * n1; nc += 4 * n1 * n1;
            for (int l = j + 1; l < M; l++)
                if (ce[l] >= 0) { int n2 = p.elem_nn[ce[l]]; nt++; np += n1 * n2; nc += 8 * n1 * n2; }
        }
    }
    ntask[c] = nt; npair[c] = np; ncoo[c] = nc;
}

__device__ __forceinline__ int task_class(const DevProblem &p, int64_t c, int n1, int n2, bool regular)
{
    if (regular) return n1 == 3 ? 0 : n1 == 4 ? (p.solver == 3 ? 2 : 1) : n1 == 6 ? 3 : 4;
    int curved = 0;
    if (p.tri_curved)
        for (int64_t t = p.cell_tri_ptr[c]; t < p.cell_tri_ptr[c + 1]; t++) curved |= p.tri_curved[t];
    return 8 + 2 * (4 * type_index(n1) + type_index(n2)) + curved;
}

// Cells whose incident elements are one or two 4-node quads, integrated over straight triangles, take the warp-cooperative
// cell kernel (k_cell_q4): all their tasks go to class FAST_CLASS, sub-bucket 0 = the first task of the cell.
__device__ __forceinline__ bool cell_is_fast(const DevProblem &p, int64_t c)
{
    if (p.cell_kind[c] != CELL_OVERLAY) return false;
    const int32_t *ce = p.cell_elem + c * p.n_mesh;
    int ne = 0;
    for (int j = 0; j < p.n_mesh; j++)
        if (ce[j] >= 0) { if (p.elem_nn[ce[j]] != 4) return false; ne++; }
    if (ne < 1 || ne > 2) return false;
    if (p.tri_curved)
        for (int64_t t = p.cell_tri_ptr[c]; t < p.cell_tri_ptr[c + 1]; t++) if (p.tri_curved[t]) return false;
    return p.cell_tri_ptr[c + 1] > p.cell_tri_ptr[c];
}

__device__ __forceinline__ int task_bucket(const DevProblem &p, int64_t c, int n1, int n2, bool regular, bool diag, int fast = 0)
{
    if (fast) return FAST_CLASS * N_SUB + (fast == 1 ? 0 : 1);
    const int cls = task_class(p, c, n1, n2, regular);
    int sub = 0;
    if (!regular) {
        const int64_t nt = p.cell_tri_ptr[c + 1] - p.cell_tri_ptr[c];
        sub = (diag ? 0 : 8) + (int)(nt < 7 ? nt : 7);
    }
    return cls * N_SUB + sub;
}

__global__ void k_fill_tasks(DevProblem p, const int64_t *__restrict__ task_off, const int64_t *__restrict__ pair_off,
                             DevTasks t, int use_fast)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= p.n_cell) return;
    const int M = p.n_mesh;
    const int32_t *ce = p.cell_elem + c * M;
    int64_t k = task_off[c], po = pair_off[c];
    if (p.cell_kind[c] == CELL_REGULAR) {
        for (int j = 0; j < M; j++)
            if (ce[j] >= 0) {
                int n = p.elem_nn[ce[j]];
                t.cell[k] = (int32_t)c; t.e1[k] = ce[j]; t.e2[k] = ce[j]; t.slots[k] = j | (j << 8);
                t.cls[k] = task_bucket(p, c, n, n, true, true); t.pair_off[k] = po;
                break;
            }
        return;
    }
    const bool fast = use_fast && cell_is_fast(p, c);
    const int64_t k0 = k;
    for (int j = 0; j < M; j++) {
        if (ce[j] < 0) continue;
        int n1 = p.elem_nn[ce[j]];
        t.cell[k] = (int32_t)c; t.e1[k] = ce[j]; t.e2[k] = ce[j]; t.slots[k] = j | (j << 8);
        t.cls[k] = task_bucket(p, c, n1, n1, false, true, fast ? (k == k0 ? 1 : 2) : 0); t.pair_off[k] = po;
        k++; po += n1 * n1;
        for (int l = j + 1; l < M; l++) {
            if (ce[l] < 0) continue;
            int n2 = p.elem_nn[ce[l]];
            t.cell[k] = (int32_t)c; t.e1[k] = ce[j]; t.e2[k] = ce[l]; t.slots[k] = j | (l << 8);
            t.cls[k] = task_bucket(p, c, n1, n2, false, false, fast ? 2 : 0); t.pair_off[k] = po;
            k++; po += n1 * n2;
        }
    }
}

// phase 0: histogram of classes; phase 1: scatter task ids into per-class lists (warp-aggregated cursors)
__global__ void k_bucket(DevTasks t, int64_t *__restrict__ cls_count, const int64_t *__restrict__ cls_start,
                         int32_t *__restrict__ cls_tasks, int phase)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = k < t.n_task;
    const int cls = active ? t.cls[k] : -1;
    const unsigned mask = __match_any_sync(0xffffffffu, cls);
    if (!active) return;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    const int rank = __popc(mask & ((1u << lane) - 1));
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd((unsigned long long *)&cls_count[cls], (unsigned long long)__popc(mask));
    base = __shfl_sync(mask, base, leader);
    if (phase == 1) cls_tasks[cls_start[cls] + (int64_t)base + rank] = (int32_t)k;
}

cudaError_t launch_cell_counts(const DevProblem &p, int32_t *ntask, int32_t *npair, int32_t *ncoo, int *err_flag,
                               cudaStream_t s)
{
    if (p.n_cell == 0) return cudaSuccess;
    k_cell_counts<<<(unsigned)((p.n_cell + 255) / 256), 256, 0, s>>>(p, ntask, npair, ncoo, err_flag);
    return cudaGetLastError();
}

cudaError_t launch_fill_tasks(const DevProblem &p, const int64_t *task_off, const int64_t *pair_off,
                              const DevTasks &t, int use_fast, cudaStream_t s)
{
    if (p.n_cell == 0) return cudaSuccess;
    k_fill_tasks<<<(unsigned)((p.n_cell + 255) / 256), 256, 0, s>>>(p, task_off, pair_off, t, use_fast);
    return cudaGetLastError();
}

cudaError_t launch_bucket_tasks(const DevTasks &t, int64_t *cls_count, const int64_t *cls_start, int32_t *cls_tasks,
                                int phase, cudaStream_t s)
{
    if (t.n_task == 0) return cudaSuccess;
    k_bucket<<<(unsigned)((t.n_task + 255) / 256), 256, 0, s>>>(t, cls_count, cls_start, cls_tasks, phase);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// integration
// ---------------------------------------------------------------------------------------

template <int N>
__device__ __forceinline__ void load_elem(const DevProblem &p, int e, double *X, double *Y)
{
    const int32_t *en = p.elem_nodes + (int64_t)e * 9;
#pragma unroll
    for (int a = 0; a < N; a++) {
        const double2 xy = *reinterpret_cast<const double2 *>(p.node_xy + 2 * (int64_t)en[a]);
        X[a] = xy.x; Y[a] = xy.y;
    }
}

// D applied to the four sums of one node pair -> the 2x2 sub-block, stored as one 32-byte sector
__device__ __forceinline__ void store_pair(double *out, const double *D, double sxx, double sxy, double syx, double syy)
{
    double kxx = D[0] * sxx + D[2] * sxy + D[6] * syx + D[8] * syy;
    double kxy = D[1] * sxy + D[2] * sxx + D[7] * syy + D[8] * syx;
    double kyx = D[3] * syx + D[5] * syy + D[6] * sxx + D[8] * sxy;
    double kyy = D[4] * syy + D[5] * syx + D[7] * sxy + D[8] * sxx;
    reinterpret_cast<double2 *>(out)[0] = make_double2(kxx, kxy);
    reinterpret_cast<double2 *>(out)[1] = make_double2(kyx, kyy);
}

// AMORE block (stiffnessMatrix.py:99-190): thread per task.
template <int N1, int N2, bool CURVED>
__global__ void __launch_bounds__(128)
k_amore(DevProblem p, DevTasks t, const int32_t *__restrict__ task_list, int64_t n_list, double *__restrict__ blk_val,
        unsigned long long *__restrict__ stats)
{
    const int64_t li = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long newton = 0, evals = 0;
    bool fail = false;
    if (li < n_list) {
        const int64_t k = task_list[li];
        const int e1 = t.e1[k], e2 = t.e2[k];
        const bool diag = (e1 == e2);
        const int64_t c = t.cell[k];
        const int sj = t.slots[k] & 0xff, sl = t.slots[k] >> 8;
        const int M = p.n_mesh;
        double X1[N1], Y1[N1], X2[N2], Y2[N2];
        load_elem<N1>(p, e1, X1, Y1);
        load_elem<N2>(p, e2, X2, Y2);
        double S[N1 * N2 * 4];
#pragma unroll
        for (int i = 0; i < N1 * N2 * 4; i++) S[i] = 0.0;
        const int nInt = amore_nint(p.solver != 1, N1, N2);
        const int qoff = tri_rule_offset(nInt);
        for (int64_t tr = p.cell_tri_ptr[c]; tr < p.cell_tri_ptr[c + 1]; tr++) {
            double tx[6], ty[6], rho1[3], rho2[3];
#pragma unroll
            for (int v = 0; v < 3; v++) {
                tx[v] = p.tri_xy[6 * tr + 2 * v]; ty[v] = p.tri_xy[6 * tr + 2 * v + 1];
                rho1[v] = p.tri_rho[(3 * tr + v) * M + sj];
                rho2[v] = p.tri_rho[(3 * tr + v) * M + sl];
            }
            bool curved = false;
            if (CURVED) {
                curved = p.tri_curved[tr] != 0;
                if (curved) {
#pragma unroll
                    for (int v = 0; v < 3; v++) { tx[3 + v] = p.tri_mid[6 * tr + 2 * v]; ty[3 + v] = p.tri_mid[6 * tr + 2 * v + 1]; }
                }
            }
            TriPoint tp;
            double dr1x = 0, dr1y = 0, dr2x = 0, dr2y = 0;
            if (!curved) {
                tri_straight(tx, ty, tp);
                tri_grad_rho(tp, rho1, dr1x, dr1y);
                if (!diag) tri_grad_rho(tp, rho2, dr2x, dr2y);
            }
            for (int q = 0; q < nInt; q++) {
                const double L1 = c_tp[3 * (qoff + q)], L2 = c_tp[3 * (qoff + q) + 1], L3 = c_tp[3 * (qoff + q) + 2];
                const double wq = c_tw[qoff + q];
                double x, y, r1v, r2v = 0.0, detJ;
                if (!CURVED || !curved) {
                    double Nt[3] = {L1, L2, 1.0 - L1 - L2};
                    x = dot_vecmat<3>(Nt, tx); y = dot_vecmat<3>(Nt, ty);
                    r1v = dot_gemm<3>(Nt, rho1);
                    if (!diag) r2v = dot_gemm<3>(Nt, rho2);
                    detJ = tp.detJ;
                } else {
                    // curved 6-node triangle: getLC per point (stiffnessMatrix.py:169-186, 213-218)
                    double Nt[6], gxt[6], gyt[6];
                    detJ = shape_grad<6>(tx, ty, L1, L2, Nt, gxt, gyt);
                    double re1[6] = {rho1[0], rho1[1], rho1[2], 0.5 * (rho1[0] + rho1[1]), 0.5 * (rho1[1] + rho1[2]), 0.5 * (rho1[2] + rho1[0])};
                    dr1x = ((gxt[0] * re1[0] + gxt[2] * re1[2]) + (gxt[1] * re1[1] + gxt[3] * re1[3])) + fma(gxt[4], re1[4], gxt[5] * re1[5]);
                    dr1y = ((gyt[0] * re1[0] + gyt[2] * re1[2]) + (gyt[1] * re1[1] + gyt[3] * re1[3])) + fma(gyt[4], re1[4], gyt[5] * re1[5]);
                    if (!diag) {
                        double re2[6] = {rho2[0], rho2[1], rho2[2], 0.5 * (rho2[0] + rho2[1]), 0.5 * (rho2[1] + rho2[2]), 0.5 * (rho2[2] + rho2[0])};
                        dr2x = ((gxt[0] * re2[0] + gxt[2] * re2[2]) + (gxt[1] * re2[1] + gxt[3] * re2[3])) + fma(gxt[4], re2[4], gxt[5] * re2[5]);
                        dr2y = ((gyt[0] * re2[0] + gyt[2] * re2[2]) + (gyt[1] * re2[1] + gyt[3] * re2[3])) + fma(gyt[4], re2[4], gyt[5] * re2[5]);
                    }
                    x = dot_vecmat<6>(Nt, tx); y = dot_vecmat<6>(Nt, ty);
                    r1v = L1 * rho1[0] + L2 * rho1[1] + L3 * rho1[2];
                    if (!diag) r2v = L1 * rho2[0] + L2 * rho2[1] + L3 * rho2[2];
                }
                double g1x[N1], g1y[N1], g2x[N2], g2y[N2];
                int it = strain_pair<N1>(X1, Y1, x, y, r1v, dr1x, dr1y, g1x, g1y);
                if (it < 0) fail = true; else newton += it;
                evals++;
                if (!diag) {
                    it = strain_pair<N2>(X2, Y2, x, y, r2v, dr2x, dr2y, g2x, g2y);
                    if (it < 0) fail = true; else newton += it;
                    evals++;
                } else {
#pragma unroll
                    for (int b = 0; b < N2; b++) { g2x[b] = g1x[b < N1 ? b : 0]; g2y[b] = g1y[b < N1 ? b : 0]; }
                }
                const double w = 0.5 * detJ * wq;
#pragma unroll
                for (int a = 0; a < N1; a++) {
                    const double wx = w * g1x[a], wy = w * g1y[a];
#pragma unroll
                    for (int b = 0; b < N2; b++) {
                        double *s4 = S + 4 * (a * N2 + b);
                        s4[0] = fma(wx, g2x[b], s4[0]);
                        s4[1] = fma(wx, g2y[b], s4[1]);
                        s4[2] = fma(wy, g2x[b], s4[2]);
                        s4[3] = fma(wy, g2y[b], s4[3]);
                    }
                }
            }
        }
        double D[9];
        const double *Dp = p.mat_D + 9 * (int64_t)p.elem_mat[e1];  // off-diagonal blocks use element j's material (AMORE.py:113)
#pragma unroll
        for (int i = 0; i < 9; i++) D[i] = Dp[i];
        double *out = blk_val + 4 * t.pair_off[k];
#pragma unroll
        for (int ab = 0; ab < N1 * N2; ab++) store_pair(out + 4 * ab, D, S[4 * ab], S[4 * ab + 1], S[4 * ab + 2], S[4 * ab + 3]);
    }
    // statistics: warp-reduce then one atomic per warp (integers: order-independent)
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        newton += __shfl_xor_sync(0xffffffffu, newton, d);
        evals += __shfl_xor_sync(0xffffffffu, evals, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (newton) atomicAdd(&stats[0], newton);
        if (evals) atomicAdd(&stats[1], evals);
    }
    if (fail) stats[2] = 1;
}

// ---------------------------------------------------------------------------------------------------------------
// Warp-cooperative cell kernel for the dominant case: overlay cells whose incident elements are one or two 4-node quads,
// straight integration triangles (stiffnessMatrix.lowerOrderAMORE:99 / quadraticAMORE:131 with n1 + n2 = 8 -> 6 points).
//
// Half a warp (16 lanes) owns one cell.
//   phase 1  lane q < 16 takes integration point q of the round (points are numbered triangle-major, 6 per triangle):
//            position, weights rho, ONE evaluation per incident element (Newton inverse map + shape gradients) -- the
//            reference evaluates element j for its diagonal block and again for the off-diagonal block; the values are
//            identical, so each (point, element) pair is evaluated once here -- and writes w*g and g of both elements
//            to shared memory (32 doubles per point).
//   phase 2  lane t < 4 * (blocks of the cell) owns row-node a of one block (diag j | off (j,l) | diag l) and accumulates
//            its 4 x 4 sums over the points of the round IN POINT ORDER with the same fma sequence as k_amore, so the
//            results are bit-identical to the thread-per-block kernel.
// Geometry is staged through shared memory, the per-cell reduction is a shared-memory transpose (no atomics).
// ---------------------------------------------------------------------------------------------------------------
constexpr int CELLQ_THREADS = 128;            // 8 cells per block
constexpr int CELLQ_PT_DOUBLES = 32;          // per point: wg1x[4] wg1y[4] g1x[4] g1y[4] wg2x[4] wg2y[4] g2x[4] g2y[4]

__global__ void __launch_bounds__(CELLQ_THREADS, 4)
k_cell_q4(DevProblem p, DevTasks t, const int32_t *__restrict__ first_task, int64_t n_cells, double *__restrict__ blk_val,
          unsigned long long *__restrict__ stats)
{
    __shared__ double pts[CELLQ_THREADS / 16][16][CELLQ_PT_DOUBLES];
    const int lane = threadIdx.x & 15, slot = threadIdx.x >> 4;
    const int64_t ci = (int64_t)blockIdx.x * (CELLQ_THREADS / 16) + slot;
    const bool live = ci < n_cells;
    unsigned long long newton = 0, evals = 0;
    bool fail = false;
    int64_t k0 = 0, c = 0;
    int e1 = 0, e2 = 0, sj = 0, sl = 0, ne = 1;
    if (live) {
        k0 = first_task[ci];
        c = t.cell[k0];
        e1 = t.e1[k0]; sj = t.slots[k0] & 0xff;
        // a second incident element shows up as the off-diagonal task right after the diagonal one
        if (k0 + 1 < t.n_task && t.cell[k0 + 1] == c) { ne = 2; e2 = t.e2[k0 + 1]; sl = t.slots[k0 + 1] >> 8; }
        else { e2 = e1; sl = sj; }
    }
    double X1[4], Y1[4], X2[4], Y2[4];
    if (live) { load_elem<4>(p, e1, X1, Y1); load_elem<4>(p, e2, X2, Y2); }
    const int64_t tr0 = live ? p.cell_tri_ptr[c] : 0;
    const int npts = live ? (int)(p.cell_tri_ptr[c + 1] - tr0) * 6 : 0;
    const int M = p.n_mesh;
    const int qoff = tri_rule_offset(6);
    // phase-2 role of this lane
    const int n_out = ne == 2 ? 12 : 4;
    const int blk = lane >> 2, a = lane & 3;          // blk 0: diag j, 1: off (j,l), 2: diag l
    const int wsrc = (blk == 2) ? 16 : 0;             // w*g of element l for the last block, of element j otherwise
    const int gsrc = (blk == 0) ? 8 : 24;             // g of element j for the first block, of element l otherwise
    double S[16];
#pragma unroll
    for (int i = 0; i < 16; i++) S[i] = 0.0;
    // all 32 lanes of the warp run the same number of rounds (the two cells of a warp may differ)
    const int npts_other = __shfl_xor_sync(0xffffffffu, npts, 16);
    const int rounds = ((npts > npts_other ? npts : npts_other) + 15) >> 4;
    for (int r = 0; r < rounds; r++) {
        const int q = r * 16 + lane;
        if (q < npts) {
            const int64_t tr = tr0 + q / 6;
            const int kq = q % 6;
            double tx[3], ty[3], rho1[3], rho2[3];
#pragma unroll
            for (int v = 0; v < 3; v++) {
                tx[v] = p.tri_xy[6 * tr + 2 * v]; ty[v] = p.tri_xy[6 * tr + 2 * v + 1];
                rho1[v] = p.tri_rho[(3 * tr + v) * M + sj];
                rho2[v] = p.tri_rho[(3 * tr + v) * M + sl];
            }
            TriPoint tp;
            tri_straight(tx, ty, tp);
            double dr1x, dr1y, dr2x = 0.0, dr2y = 0.0;
            tri_grad_rho(tp, rho1, dr1x, dr1y);
            if (ne == 2) tri_grad_rho(tp, rho2, dr2x, dr2y);
            const double L1 = c_tp[3 * (qoff + kq)], L2 = c_tp[3 * (qoff + kq) + 1];
            const double wq = c_tw[qoff + kq];
            double Nt[3] = {L1, L2, 1.0 - L1 - L2};
            const double x = dot_vecmat<3>(Nt, tx), y = dot_vecmat<3>(Nt, ty);
            const double r1v = dot_gemm<3>(Nt, rho1);
            const double w = 0.5 * tp.detJ * wq;
            double gx[4], gy[4];
            int it = strain_pair<4>(X1, Y1, x, y, r1v, dr1x, dr1y, gx, gy);
            if (it < 0) fail = true; else newton += (unsigned long long)it * (ne == 2 ? 2 : 1);   // the reference evaluates twice
            evals += (ne == 2 ? 2 : 1);
            double *row = pts[slot][lane];
#pragma unroll
            for (int b = 0; b < 4; b++) { row[b] = w * gx[b]; row[4 + b] = w * gy[b]; row[8 + b] = gx[b]; row[12 + b] = gy[b]; }
            if (ne == 2) {
                const double r2v = dot_gemm<3>(Nt, rho2);
                it = strain_pair<4>(X2, Y2, x, y, r2v, dr2x, dr2y, gx, gy);
                if (it < 0) fail = true; else newton += 2ull * it;
                evals += 2;
#pragma unroll
                for (int b = 0; b < 4; b++) { row[16 + b] = w * gx[b]; row[20 + b] = w * gy[b]; row[24 + b] = gx[b]; row[28 + b] = gy[b]; }
            }
        }
        __syncwarp();
        if (live && lane < n_out) {
            const int nq = npts - r * 16 < 16 ? npts - r * 16 : 16;
            for (int qq = 0; qq < nq; qq++) {
                const double *row = pts[slot][qq];
                const double wx = row[wsrc + a], wy = row[wsrc + 4 + a];
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const double g2x = row[gsrc + b], g2y = row[gsrc + 4 + b];
                    S[4 * b + 0] = fma(wx, g2x, S[4 * b + 0]);
                    S[4 * b + 1] = fma(wx, g2y, S[4 * b + 1]);
                    S[4 * b + 2] = fma(wy, g2x, S[4 * b + 2]);
                    S[4 * b + 3] = fma(wy, g2y, S[4 * b + 3]);
                }
            }
        }
        __syncwarp();
    }
    if (live && lane < n_out) {
        // diag j and off (j,l) use element j's material (AMORE.py:113), diag l its own
        const double *Dp = p.mat_D + 9 * (int64_t)p.elem_mat[blk == 2 ? e2 : e1];
        double D[9];
#pragma unroll
        for (int i = 0; i < 9; i++) D[i] = Dp[i];
        double *out = blk_val + 4 * (t.pair_off[k0 + blk] + 4 * a);
#pragma unroll
        for (int b = 0; b < 4; b++) store_pair(out + 4 * b, D, S[4 * b], S[4 * b + 1], S[4 * b + 2], S[4 * b + 3]);
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        newton += __shfl_xor_sync(0xffffffffu, newton, d);
        evals += __shfl_xor_sync(0xffffffffu, evals, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (newton) atomicAdd(&stats[0], newton);
        if (evals) atomicAdd(&stats[1], evals);
    }
    if (fail) stats[2] = 1;
}

// Regular elements: triFE (1 pt) / triQuadFE (3 pts) / quadFE (2x2) / quadQuadFE (3x3), stiffnessMatrix.py:9-81.
template <int N>
__global__ void __launch_bounds__(128)
k_regular(DevProblem p, DevTasks t, const int32_t *__restrict__ task_list, int64_t n_list, double *__restrict__ blk_val)
{
    const int64_t li = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= n_list) return;
    const int64_t k = task_list[li];
    const int e = t.e1[k];
    double X[N], Y[N];
    load_elem<N>(p, e, X, Y);
    double S[N * N * 4];
#pragma unroll
    for (int i = 0; i < N * N * 4; i++) S[i] = 0.0;
    constexpr bool TRI = (N == 3 || N == 6);
    constexpr int NQ = TRI ? (N == 3 ? 1 : 3) : (N == 4 ? 4 : 9);
    for (int q = 0; q < NQ; q++) {
        double r, s, w1, w2 = 1.0;
        if (TRI) {
            const int off = (N == 3) ? 0 : 1;
            r = c_tp[3 * (off + q)]; s = c_tp[3 * (off + q) + 1]; w1 = c_tw[off + q];
        } else if (N == 4) {
            r = c_q2p[q >> 1]; s = c_q2p[q & 1]; w1 = 1.0; w2 = 1.0;
        } else {
            r = c_q3p[q / 3]; s = c_q3p[q % 3]; w1 = c_q3w[q / 3]; w2 = c_q3w[q % 3];
        }
        double Nv[N], gx[N], gy[N];
        const double detJ = shape_grad<N>(X, Y, r, s, Nv, gx, gy);
        const double w = TRI ? 0.5 * detJ * w1 : detJ * w1 * w2;
#pragma unroll
        for (int a = 0; a < N; a++) {
            const double wx = w * gx[a], wy = w * gy[a];
#pragma unroll
            for (int b = 0; b < N; b++) {
                double *s4 = S + 4 * (a * N + b);
                s4[0] = fma(wx, gx[b], s4[0]);
                s4[1] = fma(wx, gy[b], s4[1]);
                s4[2] = fma(wy, gx[b], s4[2]);
                s4[3] = fma(wy, gy[b], s4[3]);
            }
        }
    }
    double D[9];
    const double *Dp = p.mat_D + 9 * (int64_t)p.elem_mat[e];
#pragma unroll
    for (int i = 0; i < 9; i++) D[i] = Dp[i];
    double *out = blk_val + 4 * t.pair_off[k];
#pragma unroll
    for (int ab = 0; ab < N * N; ab++) store_pair(out + 4 * ab, D, S[4 * ab], S[4 * ab + 1], S[4 * ab + 2], S[4 * ab + 3]);
}

// ICMFE (stiffnessMatrix.py:49-67, shapeFunction.py:53-73): quad4 + two incompatible modes, static condensation.
__global__ void __launch_bounds__(128)
k_icm(DevProblem p, DevTasks t, const int32_t *__restrict__ task_list, int64_t n_list, double *__restrict__ blk_val)
{
    const int64_t li = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= n_list) return;
    const int64_t k = task_list[li];
    const int e = t.e1[k];
    double X[4], Y[4];
    load_elem<4>(p, e, X, Y);
    double D[9];
    const double *Dp = p.mat_D + 9 * (int64_t)p.elem_mat[e];
    for (int i = 0; i < 9; i++) D[i] = Dp[i];
    // dense 12 x 12: [K (8x8) E (8x4); E^T H (4x4)] built from the 3 x 12 matrix [B G]
    double A[12 * 12];
    for (int i = 0; i < 144; i++) A[i] = 0.0;
    for (int q = 0; q < 4; q++) {
        const double r = c_q2p[q >> 1], s = c_q2p[q & 1];
        double Nv[4], dr[4], ds[4];
        Elem<4>::shape(r, s, Nv, dr, ds);
        double j00, j01, j10, j11;
        jacobian<4>(dr, ds, X, Y, j00, j01, j10, j11);
        LU2 f = lu2_factor(j00, j01, j10, j11);
        const double ru = 1.0 / f.u11, ra = 1.0 / f.a00;
        double gx[6], gy[6];
        for (int a = 0; a < 4; a++) lu2_solve_col(f, ru, ra, dr[a], ds[a], gx[a], gy[a]);
        lu2_solve_col(f, ru, ra, -2.0 * r, 0.0, gx[4], gy[4]);
        lu2_solve_col(f, ru, ra, 0.0, -2.0 * s, gx[5], gy[5]);
        const double w = lu2_det(f);  // weights are 1
        // BG (3 x 12): column 2a = (gx,0,gy), column 2a+1 = (0,gy,gx)
        double BG[3][12], DBG[3][12];
        for (int a = 0; a < 6; a++) {
            BG[0][2 * a] = gx[a]; BG[1][2 * a] = 0.0; BG[2][2 * a] = gy[a];
            BG[0][2 * a + 1] = 0.0; BG[1][2 * a + 1] = gy[a]; BG[2][2 * a + 1] = gx[a];
        }
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 12; j++)
                DBG[i][j] = fma(D[3 * i + 2], BG[2][j], fma(D[3 * i + 1], BG[1][j], D[3 * i] * BG[0][j]));
        for (int i = 0; i < 12; i++)
            for (int j = 0; j < 12; j++) {
                double v = fma(BG[2][i], DBG[2][j], fma(BG[1][i], DBG[1][j], BG[0][i] * DBG[0][j]));
                A[12 * i + j] += v * w * 1.0 * 1.0;
            }
    }
    // X = solve(H, E^T)  (4 x 8), partial-pivot LU as numpy.linalg.solve
    double H[16], Xs[4 * 8];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) H[4 * i + j] = A[12 * (8 + i) + 8 + j];
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 8; j++) Xs[8 * i + j] = A[12 * j + 8 + i];  // Eloc^T
    for (int c = 0; c < 4; c++) {
        int pv = c; double best = fabs(H[4 * c + c]);
        for (int r = c + 1; r < 4; r++) if (fabs(H[4 * r + c]) > best) { best = fabs(H[4 * r + c]); pv = r; }
        if (pv != c) {
            for (int j = 0; j < 4; j++) { double tt = H[4 * c + j]; H[4 * c + j] = H[4 * pv + j]; H[4 * pv + j] = tt; }
            for (int j = 0; j < 8; j++) { double tt = Xs[8 * c + j]; Xs[8 * c + j] = Xs[8 * pv + j]; Xs[8 * pv + j] = tt; }
        }
        const double rp = 1.0 / H[4 * c + c];
        for (int r = c + 1; r < 4; r++) {
            const double l = H[4 * r + c] * rp;
            H[4 * r + c] = l;
            for (int j = c + 1; j < 4; j++) H[4 * r + j] -= l * H[4 * c + j];
        }
    }
    for (int j = 0; j < 8; j++) {
        for (int r = 1; r < 4; r++) {
            double sacc = Xs[8 * r + j];
            for (int c = 0; c < r; c++) sacc -= H[4 * r + c] * Xs[8 * c + j];
            Xs[8 * r + j] = sacc;
        }
        for (int r = 3; r >= 0; r--) {
            double sacc = Xs[8 * r + j];
            for (int c = r + 1; c < 4; c++) sacc -= H[4 * r + c] * Xs[8 * c + j];
            Xs[8 * r + j] = sacc / H[4 * r + r];
        }
    }
    double *out = blk_val + 4 * t.pair_off[k];
    for (int i = 0; i < 8; i++)
        for (int j = 0; j < 8; j++) {
            double acc = A[12 * i + 8] * Xs[j];
            for (int c = 1; c < 4; c++) acc = fma(A[12 * i + 8 + c], Xs[8 * c + j], acc);
            const double v = A[12 * i + j] - acc;
            out[4 * ((i >> 1) * 4 + (j >> 1)) + 2 * (i & 1) + (j & 1)] = v;
        }
}

template <int N1, int N2>
static cudaError_t launch_amore(bool curved, const DevProblem &p, const DevTasks &t, const int32_t *list, int64_t n,
                                double *blk_val, unsigned long long *stats, cudaStream_t s)
{
    const unsigned grid = (unsigned)((n + 127) / 128);
    if (curved) k_amore<N1, N2, true><<<grid, 128, 0, s>>>(p, t, list, n, blk_val, stats);
    else k_amore<N1, N2, false><<<grid, 128, 0, s>>>(p, t, list, n, blk_val, stats);
    return cudaGetLastError();
}

template <int N1>
static cudaError_t launch_amore_n2(int i2, bool curved, const DevProblem &p, const DevTasks &t, const int32_t *list,
                                   int64_t n, double *blk_val, unsigned long long *stats, cudaStream_t s)
{
    switch (i2) {
    case 0: return launch_amore<N1, 3>(curved, p, t, list, n, blk_val, stats, s);
    case 1: return launch_amore<N1, 4>(curved, p, t, list, n, blk_val, stats, s);
    case 2: return launch_amore<N1, 6>(curved, p, t, list, n, blk_val, stats, s);
    default: return launch_amore<N1, 9>(curved, p, t, list, n, blk_val, stats, s);
    }
}

cudaError_t launch_integrate_class(int cls, const DevProblem &p, const DevTasks &t, const int32_t *list, int64_t n,
                                   double *blk_val, unsigned long long *stats, cudaStream_t s)
{
    if (n == 0) return cudaSuccess;
    const unsigned grid = (unsigned)((n + 127) / 128);
    if (cls == FAST_CLASS) {
        k_cell_q4<<<(unsigned)((n + CELLQ_THREADS / 16 - 1) / (CELLQ_THREADS / 16)), CELLQ_THREADS, 0, s>>>(p, t, list, n, blk_val, stats);
        return cudaGetLastError();
    }
    if (cls < 8) {
        switch (cls) {
        case 0: k_regular<3><<<grid, 128, 0, s>>>(p, t, list, n, blk_val); break;
        case 1: k_regular<4><<<grid, 128, 0, s>>>(p, t, list, n, blk_val); break;
        case 2: k_icm<<<grid, 128, 0, s>>>(p, t, list, n, blk_val); break;
        case 3: k_regular<6><<<grid, 128, 0, s>>>(p, t, list, n, blk_val); break;
        case 4: k_regular<9><<<grid, 128, 0, s>>>(p, t, list, n, blk_val); break;
        default: return cudaErrorInvalidValue;
        }
        return cudaGetLastError();
    }
    const int code = cls - 8;
    const bool curved = code & 1;
    const int i1 = (code >> 1) / 4, i2 = (code >> 1) % 4;
    switch (i1) {
    case 0: return launch_amore_n2<3>(i2, curved, p, t, list, n, blk_val, stats, s);
    case 1: return launch_amore_n2<4>(i2, curved, p, t, list, n, blk_val, stats, s);
    case 2: return launch_amore_n2<6>(i2, curved, p, t, list, n, blk_val, stats, s);
    default: return launch_amore_n2<9>(i2, curved, p, t, list, n, blk_val, stats, s);
    }
}

}  // namespace mofem
