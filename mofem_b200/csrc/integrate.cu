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
// Output (blk_val): one 32-byte record per CONTRIBUTION, [K(2a,2b), K(2a,2b+1), K(2a+1,2b), K(2a+1,2b+1)], stored straight at
// its position in the reduction order of the CSR build (csr.cu: pair_dst); node pair s of the task layout (per task n1*n2 pairs,
// row-major) goes to record pair_dst[s], the transposed copy of an off-diagonal pair to record pair_dst[off_t + s].
//
// Compiled with -fmad=false (see fem_device.cuh); accumulations use explicit fma().
#include <cstdlib>
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

// Overlay cells whose incident elements (1..3) are all 4-node quads, integrated over 1..5 straight triangles, take the cell
// kernel (k_cell_quads): all their tasks go to class FAST_CLASS; sub-bucket (ne - 1) * 5 + (T - 1) holds the FIRST task of each
// such cell (one launch per (ne, T)), sub-bucket 15 the remaining tasks (never iterated).  Returns the sub-bucket or -1.
__device__ __forceinline__ int cell_fast_sub(const DevProblem &p, int64_t c)
{
    if (p.cell_kind[c] != CELL_OVERLAY) return -1;
    const int32_t *ce = p.cell_elem + c * p.n_mesh;
    int ne = 0;
    for (int j = 0; j < p.n_mesh; j++)
        if (ce[j] >= 0) { if (p.elem_nn[ce[j]] != 4) return -1; ne++; }
    if (ne < 1 || ne > 3) return -1;
    const int64_t nt = p.cell_tri_ptr[c + 1] - p.cell_tri_ptr[c];
    if (nt < 1 || nt > 5) return -1;
    if (p.tri_curved)
        for (int64_t t = p.cell_tri_ptr[c]; t < p.cell_tri_ptr[c + 1]; t++) if (p.tri_curved[t]) return -1;
    return (ne - 1) * 5 + (int)nt - 1;
}

__device__ __forceinline__ int task_bucket(const DevProblem &p, int64_t c, int n1, int n2, bool regular, bool diag, int fast = -1)
{
    if (fast >= 0) return FAST_CLASS * N_SUB + fast;
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
    const int fsub = use_fast ? cell_fast_sub(p, c) : -1;
    const int64_t k0 = k;
    for (int j = 0; j < M; j++) {
        if (ce[j] < 0) continue;
        int n1 = p.elem_nn[ce[j]];
        t.cell[k] = (int32_t)c; t.e1[k] = ce[j]; t.e2[k] = ce[j]; t.slots[k] = j | (j << 8);
        t.cls[k] = task_bucket(p, c, n1, n1, false, true, fsub < 0 ? -1 : (k == k0 ? fsub : N_SUB - 1)); t.pair_off[k] = po;
        k++; po += n1 * n1;
        for (int l = j + 1; l < M; l++) {
            if (ce[l] < 0) continue;
            int n2 = p.elem_nn[ce[l]];
            t.cell[k] = (int32_t)c; t.e1[k] = ce[j]; t.e2[k] = ce[l]; t.slots[k] = j | (l << 8);
            t.cls[k] = task_bucket(p, c, n1, n2, false, false, fsub < 0 ? -1 : N_SUB - 1); t.pair_off[k] = po;
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

__device__ __forceinline__ void store_record(double *rec, double a, double b, double c, double d)
{   // one 256-bit store (STG.E.256): the record is written as a whole sector
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(rec), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

// D applied to the four sums of node pair s (task layout) -> the 2x2 sub-block, stored as one 32-byte record at the pair's
// position in the reduction order; off-diagonal blocks (both) also store the transposed sub-block
// D applied to the four sums of a node pair.  Not an ill-conditioned step (see fem_device.cuh): contracted into fused chains,
// 4 instructions per entry instead of 7.
__device__ __forceinline__ void apply_D(const double *D, double sxx, double sxy, double syx, double syy, double *k)
{
    k[0] = fma(D[8], syy, fma(D[6], syx, fma(D[2], sxy, D[0] * sxx)));
    k[1] = fma(D[8], syx, fma(D[7], syy, fma(D[2], sxx, D[1] * sxy)));
    k[2] = fma(D[8], sxy, fma(D[6], sxx, fma(D[5], syy, D[3] * syx)));
    k[3] = fma(D[8], sxx, fma(D[7], sxy, fma(D[5], syx, D[4] * syy)));
}
__device__ __forceinline__ void store_pair(const BlkOut &o, int64_t s, bool both, const double *D, double sxx, double sxy, double syx,
                                           double syy)
{
    double k[4];
    apply_D(D, sxx, sxy, syx, syy, k);
    store_record(o.val + 4 * (int64_t)o.dst[s], k[0], k[1], k[2], k[3]);
    if (both) store_record(o.val + 4 * (int64_t)o.dst[o.off_t + s], k[0], k[2], k[1], k[3]);
}

// AMORE block (stiffnessMatrix.py:99-190): thread per task.
template <int N1, int N2, bool CURVED>
__global__ void __launch_bounds__(128)
k_amore(DevProblem p, DevTasks t, const int32_t *__restrict__ task_list, int64_t n_list, BlkOut blk_val,
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
        const int64_t po = t.pair_off[k];
#pragma unroll
        for (int ab = 0; ab < N1 * N2; ab++) store_pair(blk_val, po + ab, !diag, D, S[4 * ab], S[4 * ab + 1], S[4 * ab + 2], S[4 * ab + 3]);
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
// Cell kernel for the dominant case: overlay cells whose NE <= 3 incident elements are all 4-node quads, integrated over
// T <= 5 straight triangles (stiffnessMatrix.lowerOrderAMORE:99 / quadraticAMORE:131 with n1 + n2 = 8 -> the 6-point rule for
// EVERY block of the cell, so all blocks share their integration points).
//
// The reference evaluates the strain matrix of element j at every point once for the diagonal block (j,j) and again for every
// off-diagonal block it takes part in (AMORE.py:59-123): NE evaluations per (point, element).  The values are identical, so here
// each (point, element) pair is evaluated ONCE.  A CTA takes a batch of CB cells and runs three phases over shared memory:
//   phase 0  thread per (cell, triangle, element): triangle corners, det J, rho_e at the corners and grad(rho_e) (constant on a
//            straight triangle) -> smem; the thread of triangle 0 also fetches the element's node coordinates and material id.
//            Everything is three dependent loads away from the cell record the symbolic phase wrote (CqRec).
//   phase 1  thread per EVALUATION (cell, point, element): position, rho, Newton inverse map, shape gradients, the compressed
//            strain pair (gx[4], gy[4]) -> one record of 8 doubles in smem.  This is where the fp64 pipe is kept busy.
//   phase 2  four threads per BLOCK (diag j | off (j,l)), each owning one row node = 16 of the block's 64 sums: they walk the
//            points IN ORDER and accumulate with the same fma sequence as k_amore, reading the records of elements j and l from
//            smem, then apply D and store the 2x2 sub-blocks as whole 32-byte records at their positions in the reduction
//            order (record numbers fetched before the accumulation; a second, transposed record for off-diagonal blocks).
//            Results are bit-identical to the thread-per-block kernel.
// A CTA is ONE warp (barriers cost nothing) with a batch of ~4 cells; about 20 CTAs are resident per SM and sit in different
// phases, so the fp64-dense Newton phase of some overlaps the load latency of phase 0 and the stores of phase 2 of others.
// Strain-pair records are 10 doubles apart (8 used; 128-bit stores of a quarter warp cover all 32 banks) and the per-cell stride
// is 4 mod 16 doubles, so the records of neighbouring cells start in different banks.
// ---------------------------------------------------------------------------------------------------------------
#ifndef CQ_THREADS
#define CQ_THREADS 32                         // one warp per CTA: __syncthreads costs nothing and the ~20 resident CTAs of an SM are in different phases
#endif
#ifndef CQ_MINB
#define CQ_MINB 20                            // resident CTAs per SM the kernel is compiled for (register cap 65536 / CQ_THREADS / CQ_MINB)
#endif
constexpr int CQ_MAX_T = 5;                   // triangles per cell handled here (sub-bucket = (NE - 1) * 5 + T - 1)
constexpr int CQ_SMEM_BUDGET = (227 * 1024) / CQ_MINB - 1024 - 256;   // dynamic bytes per CTA

constexpr int CQ_ROW = 10;                    // doubles between two strain-pair records: 8 quarter-warp lanes x 16 bytes at this stride cover all 32 banks
struct CqLayout {       // offsets in doubles inside one cell record
    int rows, exy, tri, trie, meta, stride;
};
__host__ __device__ inline CqLayout cq_layout(int ne, int T)
{
    CqLayout L;
    L.rows = 0;                                  // [T*6][ne][10] gx[4] gy[4] . .   (16-byte aligned records, 128-bit accesses)
    L.exy = L.rows + ne * T * 6 * CQ_ROW;        // [ne][10]      X[4] Y[4] . .
    L.tri = L.exy + ne * CQ_ROW;                 // [T][7]        tx[3] ty[3] detJ
    L.trie = L.tri + T * 7;                      // [T][ne][5]    rho[3] drho_x drho_y
    L.meta = L.trie + T * ne * 5;                // int64: offset of the cell's first block in blk_val (node pairs); int32[ne]: material ids
    int n = L.meta + 1 + (ne + 1) / 2;
    L.stride = n + ((4 - n % 16) + 16) % 16;     // stride = 4 mod 16
    return L;
}

// one (incident elements, triangles) sub-bucket inside the launch: its CTAs are [cta_start, next cta_start)
struct CqBucket { int64_t list_off, n_cells; int32_t cta_start, ne, T, CB; };
struct CqPlan { int32_t n, n_cta; CqBucket b[N_SUB - 1]; };

// Symbolic phase: one 32-byte record per cell of the cell kernel, in the order of the sub-bucket lists, so that the kernel reaches
// its inputs in three dependent loads (record -> element nodes / material id -> coordinates / D) instead of five.
__global__ void k_cq_records(DevProblem p, DevTasks t, const int32_t *__restrict__ first_task, int64_t n, CqRec *__restrict__ rec)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t k0 = first_task[i];
    const int64_t c = t.cell[k0];
    const int32_t *ce = p.cell_elem + c * p.n_mesh;
    CqRec r;
    r.pair0 = t.pair_off[k0];
    r.tri0 = p.cell_tri_ptr[c];
    r.el[0] = r.el[1] = r.el[2] = -1;
    r.slots = 0;
    int e = 0;
    for (int j = 0; j < p.n_mesh && e < 3; j++)
        if (ce[j] >= 0) { r.el[e] = ce[j]; r.slots |= j << (8 * e); e++; }
    rec[i] = r;
}

cudaError_t launch_cq_records(const DevProblem &p, const DevTasks &t, const int32_t *first_task, int64_t n, CqRec *rec, cudaStream_t s)
{
    if (n == 0) return cudaSuccess;
    k_cq_records<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p, t, first_task, n, rec);
    return cudaGetLastError();
}

template <int NE>
__device__ __forceinline__ void
cq_body(const DevProblem &p, const CqRec *__restrict__ recs, const int64_t n_cells, const int T, const int CB, const int cta,
        BlkOut blk_val, unsigned long long *__restrict__ stats)
{
    extern __shared__ double sm[];
    constexpr int TPC = NE * (NE + 1) / 2;                 // blocks (tasks) per cell
    const int EPC = NE * T * 6;                            // evaluations per cell
    const CqLayout L = cq_layout(NE, T);
    double *rule = sm;                                     // the 6-point rule: L1[6] L2[6] (thread-dependent index: not from __constant__)
    double *cells = sm + 12;                               // 16-byte aligned (dynamic shared memory starts 16-byte aligned, 96 bytes in)
    const int tid = threadIdx.x;
    const int64_t c0 = (int64_t)cta * CB;
    const int nb = (int)(n_cells - c0 < CB ? n_cells - c0 : CB);
    const int M = p.n_mesh;
    const int qoff = tri_rule_offset(6);
    unsigned long long newton = 0, evals = 0;
    bool fail = false;

    // ---- phase 0: thread per (cell, triangle, element) -> corners, det J, rho_e, grad(rho_e); the thread of triangle 0 also
    //      fetches the element: node coordinates and material id.  All loads hang off the cell record.
    if (tid < 6) { rule[tid] = c_tp[3 * (qoff + tid)]; rule[6 + tid] = c_tp[3 * (qoff + tid) + 1]; }
    for (int i = tid; i < nb * T * NE; i += CQ_THREADS) {
        const int ci = i / (T * NE), r = i - ci * (T * NE), tr = r / NE, e = r - tr * NE;
        const int4 *rp = reinterpret_cast<const int4 *>(recs + c0 + ci);
        const int4 r0 = __ldg(rp), r1 = __ldg(rp + 1);       // pair0 | tri0 ; el[3] | slots
        const int64_t trg = (((int64_t)r0.w << 32) | (uint32_t)r0.z) + tr;
        const int el = e == 0 ? r1.x : e == 1 ? r1.y : r1.z;
        const int slot = (r1.w >> (8 * e)) & 0xff;
        double *cell = cells + (size_t)ci * L.stride;
        const bool lead = tr == 0;
        int en[4] = {0, 0, 0, 0};
        int mat = 0;
        if (lead) {
            const int32_t *enp = p.elem_nodes + (int64_t)el * 9;
#pragma unroll
            for (int a = 0; a < 4; a++) en[a] = enp[a];
            mat = p.elem_mat[el];
        }
        double tx[3], ty[3], rho[3];
#pragma unroll
        for (int v = 0; v < 3; v++) {
            const double2 c2 = *reinterpret_cast<const double2 *>(p.tri_xy + 6 * trg + 2 * v);
            tx[v] = c2.x; ty[v] = c2.y;
            rho[v] = p.tri_rho[(3 * trg + v) * M + slot];
        }
        if (lead) {
            double *xy = cell + L.exy + e * CQ_ROW;
#pragma unroll
            for (int a = 0; a < 4; a++) {
                const double2 v = *reinterpret_cast<const double2 *>(p.node_xy + 2 * (int64_t)en[a]);
                xy[a] = v.x; xy[4 + a] = v.y;
            }
            reinterpret_cast<int32_t *>(cell + L.meta + 1)[e] = mat;
            if (e == 0) *reinterpret_cast<int64_t *>(cell + L.meta) = ((int64_t)r0.y << 32) | (uint32_t)r0.x;
        }
        TriPoint tp;
        tri_straight(tx, ty, tp);
        double gx, gy;
        tri_grad_rho(tp, rho, gx, gy);
        if (e == 0) {
            double *tg = cell + L.tri + tr * 7;
#pragma unroll
            for (int v = 0; v < 3; v++) { tg[v] = tx[v]; tg[3 + v] = ty[v]; }
            tg[6] = tp.detJ;
        }
        double *te = cell + L.trie + (tr * NE + e) * 5;
        te[0] = rho[0]; te[1] = rho[1]; te[2] = rho[2]; te[3] = gx; te[4] = gy;
    }
    __syncthreads();
    // ---- phase 1: one evaluation per (cell, point, element)
    int ci = tid / EPC, rem = tid - ci * EPC;              // (cell, evaluation in the cell) advanced without a division per round
    for (int ev = tid; ev < nb * EPC; ev += CQ_THREADS, rem += CQ_THREADS) {
        while (rem >= EPC) { rem -= EPC; ci++; }
        const int q = rem / NE, e = rem - q * NE, tr = q / 6, kq = q - tr * 6;
        double *cell = cells + (size_t)ci * L.stride;
        const double *tg = cell + L.tri + tr * 7;
        const double *te = cell + L.trie + (tr * NE + e) * 5;
        const double2 *xy = reinterpret_cast<const double2 *>(cell + L.exy + e * CQ_ROW);
        const double tx[3] = {tg[0], tg[1], tg[2]}, ty[3] = {tg[3], tg[4], tg[5]}, rho[3] = {te[0], te[1], te[2]};
        const double L1 = rule[kq], L2 = rule[6 + kq];
        const double Nt[3] = {L1, L2, 1.0 - L1 - L2};
        const double x = dot_vecmat<3>(Nt, tx), y = dot_vecmat<3>(Nt, ty);
        const double rv = dot_gemm<3>(Nt, rho);
        double gx[4], gy[4];
        const double2 x01 = xy[0], x23 = xy[1], y01 = xy[2], y23 = xy[3];
        const double X[4] = {x01.x, x01.y, x23.x, x23.y}, Y[4] = {y01.x, y01.y, y23.x, y23.y};
        const int it = strain_pair<4>(X, Y, x, y, rv, te[3], te[4], gx, gy);
        if (it < 0) fail = true; else newton += (unsigned long long)(it * NE);   // the reference evaluates NE times (1 diag + NE-1 off)
        evals += NE;
        double2 *row = reinterpret_cast<double2 *>(cell + L.rows + (q * NE + e) * CQ_ROW);
        row[0] = make_double2(gx[0], gx[1]); row[1] = make_double2(gx[2], gx[3]);
        row[2] = make_double2(gy[0], gy[1]); row[3] = make_double2(gy[2], gy[3]);
    }
    __syncthreads();
    // ---- phase 2: four threads per block, one row node x 4 column nodes x 4 sums each
    for (int item = tid; item < nb * TPC * 4; item += CQ_THREADS) {
        const int ci = item / (TPC * 4), r = item - ci * (TPC * 4), tk = r >> 2, a = r & 3;
        int ej, el;
        if (NE == 1) { ej = 0; el = 0; }
        else if (NE == 2) { ej = tk == 2 ? 1 : 0; el = tk == 0 ? 0 : 1; }
        else { ej = tk < 3 ? 0 : (tk < 5 ? 1 : 2); el = tk < 3 ? tk : (tk < 5 ? tk - 2 : 2); }
        const double *cell = cells + (size_t)ci * L.stride;
        const double *r1 = cell + L.rows + ej * CQ_ROW + a;
        const double2 *r2 = reinterpret_cast<const double2 *>(cell + L.rows + el * CQ_ROW);
        const double *tg = cell + L.tri + 6;
        // the blocks of a cell are consecutive in the task layout, 16 node pairs each; the records of this thread's four node pairs
        // (and of their transposed copies) are fetched now and used after the accumulation
        const int64_t s0 = *reinterpret_cast<const int64_t *>(cell + L.meta) + 16 * tk + 4 * a;
        const bool both = ej != el;
        uint32_t rec[4], rec_t[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int b = 0; b < 4; b++) rec[b] = __ldg(blk_val.dst + s0 + b);
        if (both) {
#pragma unroll
            for (int b = 0; b < 4; b++) rec_t[b] = __ldg(blk_val.dst + blk_val.off_t + s0 + b);
        }
        double S[16];
#pragma unroll
        for (int i = 0; i < 16; i++) S[i] = 0.0;
        for (int tr = 0; tr < T; tr++) {
            const double hd = 0.5 * tg[tr * 7];
#pragma unroll
            for (int kq = 0; kq < 6; kq++) {
                const double w = hd * c_tw[qoff + kq];
                const double wx = w * r1[0], wy = w * r1[4];
                const double2 b0 = r2[0], b1 = r2[1], b2 = r2[2], b3 = r2[3];
                const double gbx[4] = {b0.x, b0.y, b1.x, b1.y}, gby[4] = {b2.x, b2.y, b3.x, b3.y};
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const double g2x = gbx[b], g2y = gby[b];
                    S[4 * b + 0] = fma(wx, g2x, S[4 * b + 0]);
                    S[4 * b + 1] = fma(wx, g2y, S[4 * b + 1]);
                    S[4 * b + 2] = fma(wy, g2x, S[4 * b + 2]);
                    S[4 * b + 3] = fma(wy, g2y, S[4 * b + 3]);
                }
                r1 += NE * CQ_ROW; r2 += NE * (CQ_ROW / 2);
            }
        }
        // diag j and off (j,l) use element j's material (AMORE.py:113)
        const double *Dp = p.mat_D + 9 * (int64_t)reinterpret_cast<const int32_t *>(cell + L.meta + 1)[ej];
        double D[9];
#pragma unroll
        for (int i = 0; i < 9; i++) D[i] = Dp[i];
        // keep the fetched record numbers opaque until here: otherwise their address arithmetic is scheduled right behind the loads
        // and the warp waits for them BEFORE the accumulation instead of during it
#pragma unroll
        for (int b = 0; b < 4; b++) { asm volatile("" : "+r"(rec[b])); asm volatile("" : "+r"(rec_t[b])); }
#pragma unroll
        for (int b = 0; b < 4; b++) {
            double k[4];
            apply_D(D, S[4 * b], S[4 * b + 1], S[4 * b + 2], S[4 * b + 3], k);
            store_record(blk_val.val + 4 * (int64_t)rec[b], k[0], k[1], k[2], k[3]);
            if (both) store_record(blk_val.val + 4 * (int64_t)rec_t[b], k[0], k[2], k[1], k[3]);
        }
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

// All sub-buckets in ONE launch (the tails of the small sub-buckets overlap the big ones): a CTA looks its sub-bucket up in the plan.
__global__ void __launch_bounds__(CQ_THREADS, CQ_MINB)
k_cell_quads(DevProblem p, const CqRec *__restrict__ recs, CqPlan plan, BlkOut blk_val,
             unsigned long long *__restrict__ stats)
{
    int i = 0;
    while (i + 1 < plan.n && (int)blockIdx.x >= plan.b[i + 1].cta_start) i++;
    const CqBucket &b = plan.b[i];
    const int cta = (int)blockIdx.x - b.cta_start;
    switch (b.ne) {
    case 1: cq_body<1>(p, recs + b.list_off, b.n_cells, b.T, b.CB, cta, blk_val, stats); break;
    case 2: cq_body<2>(p, recs + b.list_off, b.n_cells, b.T, b.CB, cta, blk_val, stats); break;
    default: cq_body<3>(p, recs + b.list_off, b.n_cells, b.T, b.CB, cta, blk_val, stats); break;
    }
}

// cells per CTA: the largest batch inside the shared-memory budget whose evaluations (phase 1, ~70 % of the work) and block rows
// (phase 2) fill whole rounds of CQ_THREADS threads best
static int cq_batch(int ne, int T)
{
    const int tpc = ne * (ne + 1) / 2, epc = ne * T * 6;
    const int cbmax = (CQ_SMEM_BUDGET - 12 * 8) / (cq_layout(ne, T).stride * 8);
    int best = 0;
    double best_eff = -1.0;
    for (int cb = 1; cb <= cbmax; cb++) {
        const int e1 = cb * epc, e2 = cb * tpc * 4;
        const double f1 = (double)e1 / (((e1 + CQ_THREADS - 1) / CQ_THREADS) * CQ_THREADS);
        const double f2 = (double)e2 / (((e2 + CQ_THREADS - 1) / CQ_THREADS) * CQ_THREADS);
        const double eff = 0.7 * f1 + 0.3 * f2;
        if (eff >= best_eff - 1e-9 && (eff > best_eff + 0.01 || cb > best)) { if (eff > best_eff) best_eff = eff; best = cb; }
    }
    return best;
}

// sub_count: cells in the N_SUB sub-buckets of class FAST_CLASS; sub-bucket (ne - 1) * 5 + T - 1 = the cells with ne incident quads
// and T triangles, whose records follow each other in `recs` in sub-bucket order (the last sub-bucket holds no cells)
cudaError_t launch_cell_quads(const DevProblem &p, const CqRec *recs, const int64_t *sub_count, BlkOut blk_val,
                              unsigned long long *stats, cudaStream_t s)
{
    CqPlan plan{};
    size_t smem = 0;
    int64_t n_cta = 0, sub_start[N_SUB];
    sub_start[0] = 0;
    for (int sb = 1; sb < N_SUB; sb++) sub_start[sb] = sub_start[sb - 1] + sub_count[sb - 1];
    for (int sb = N_SUB - 2; sb >= 0; sb--) {         // heaviest CTAs first: 3 elements before 2 before 1, many triangles before few
        if (!sub_count[sb]) continue;
        const int ne = sb / CQ_MAX_T + 1, T = sb % CQ_MAX_T + 1;
        int CB = cq_batch(ne, T);
        if (const char *ov = getenv("MOFEM_CQ_CB")) {      // tuning override: cells per CTA (clamped to the hardware limit)
            const int lim = (int)((227 * 1024 - 1024 - 96) / (cq_layout(ne, T).stride * 8));
            CB = atoi(ov) < lim ? atoi(ov) : lim;
        }
        if (CB < 1) return cudaErrorInvalidValue;
        CqBucket &b = plan.b[plan.n++];
        b.list_off = sub_start[sb]; b.n_cells = sub_count[sb]; b.cta_start = (int32_t)n_cta; b.ne = ne; b.T = T; b.CB = CB;
        n_cta += (sub_count[sb] + CB - 1) / CB;
        const size_t need = (size_t)CB * cq_layout(ne, T).stride * 8 + 12 * 8;
        if (need > smem) smem = need;
    }
    if (!plan.n) return cudaSuccess;
    if (n_cta > 0x7fffffff) return cudaErrorInvalidValue;
    plan.n_cta = (int32_t)n_cta;
    cudaError_t e = cudaFuncSetAttribute(k_cell_quads, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem > (size_t)CQ_SMEM_BUDGET ? smem : (size_t)CQ_SMEM_BUDGET));
    if (e != cudaSuccess) return e;
    k_cell_quads<<<(unsigned)n_cta, CQ_THREADS, smem, s>>>(p, recs, plan, blk_val, stats);
    return cudaGetLastError();
}

// Regular elements: triFE (1 pt) / triQuadFE (3 pts) / quadFE (2x2) / quadQuadFE (3x3), stiffnessMatrix.py:9-81.
template <int N>
__global__ void __launch_bounds__(128)
k_regular(DevProblem p, DevTasks t, const int32_t *__restrict__ task_list, int64_t n_list, BlkOut blk_val)
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
    const int64_t po = t.pair_off[k];
#pragma unroll
    for (int ab = 0; ab < N * N; ab++) store_pair(blk_val, po + ab, false, D, S[4 * ab], S[4 * ab + 1], S[4 * ab + 2], S[4 * ab + 3]);
}

// ICMFE (stiffnessMatrix.py:49-67, shapeFunction.py:53-73): quad4 + two incompatible modes, static condensation.
__global__ void __launch_bounds__(128)
k_icm(DevProblem p, DevTasks t, const int32_t *__restrict__ task_list, int64_t n_list, BlkOut blk_val)
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
    const int64_t po = t.pair_off[k];
    for (int i = 0; i < 8; i += 2)
        for (int j = 0; j < 8; j += 2) {
            double v[4];
            for (int pq = 0; pq < 4; pq++) {
                const int ii = i + (pq >> 1), jj = j + (pq & 1);
                double acc = A[12 * ii + 8] * Xs[jj];
                for (int c = 1; c < 4; c++) acc = fma(A[12 * ii + 8 + c], Xs[8 * c + jj], acc);
                v[pq] = A[12 * ii + jj] - acc;
            }
            store_record(blk_val.val + 4 * (int64_t)blk_val.dst[po + (i >> 1) * 4 + (j >> 1)], v[0], v[1], v[2], v[3]);
        }
}

template <int N1, int N2>
static cudaError_t launch_amore(bool curved, const DevProblem &p, const DevTasks &t, const int32_t *list, int64_t n,
                                BlkOut blk_val, unsigned long long *stats, cudaStream_t s)
{
    const unsigned grid = (unsigned)((n + 127) / 128);
    if (curved) k_amore<N1, N2, true><<<grid, 128, 0, s>>>(p, t, list, n, blk_val, stats);
    else k_amore<N1, N2, false><<<grid, 128, 0, s>>>(p, t, list, n, blk_val, stats);
    return cudaGetLastError();
}

template <int N1>
static cudaError_t launch_amore_n2(int i2, bool curved, const DevProblem &p, const DevTasks &t, const int32_t *list,
                                   int64_t n, BlkOut blk_val, unsigned long long *stats, cudaStream_t s)
{
    switch (i2) {
    case 0: return launch_amore<N1, 3>(curved, p, t, list, n, blk_val, stats, s);
    case 1: return launch_amore<N1, 4>(curved, p, t, list, n, blk_val, stats, s);
    case 2: return launch_amore<N1, 6>(curved, p, t, list, n, blk_val, stats, s);
    default: return launch_amore<N1, 9>(curved, p, t, list, n, blk_val, stats, s);
    }
}

cudaError_t launch_integrate_class(int cls, const DevProblem &p, const DevTasks &t, const int32_t *list, int64_t n,
                                   BlkOut blk_val, unsigned long long *stats, cudaStream_t s)
{
    if (n == 0) return cudaSuccess;
    const unsigned grid = (unsigned)((n + 127) / 128);
    if (cls == FAST_CLASS) return cudaErrorInvalidValue;      // launch_cell_quads, one launch per (ne, T) sub-bucket
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
