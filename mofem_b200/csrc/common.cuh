// common.cuh -- shared declarations of the mofem_b200 CUDA library (internal).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/mofem_b200.h"
#include "nccl_shim.h"

namespace mofem {

// cell kinds / task classes -------------------------------------------------------------
enum { CELL_REGULAR = 0, CELL_OVERLAY = 1 };

// class id of a task (one emitted block):
//   0..4   regular element rules: tri3, quad4 (quadFE), quad4 (ICMFE), tri6, quad9
//   8 + 2*(4*i1+i2) + curved   AMORE blocks, i = index of n in {3,4,6,9}
constexpr int N_CLASS = 8 + 32;
// tasks of one class are further grouped by (diagonal / off-diagonal, triangle count) so that the threads of a warp
// run the same trip counts; a bucket id is class * N_SUB + sub
constexpr int N_SUB = 16;
constexpr int FAST_CLASS = 5;   // all-quad4 overlay cells on the cell kernel: sub-bucket (ne-1)*5 + T-1 = first tasks of the cells with ne elements / T triangles
constexpr int N_BUCKET = N_CLASS * N_SUB;
__host__ __device__ inline int type_index(int nn) { return nn == 3 ? 0 : nn == 4 ? 1 : nn == 6 ? 2 : 3; }

// device view of the flattened problem
struct DevProblem {
    int64_t n_node, n_elem, n_cell, n_tri;
    int32_t n_mesh, n_mat, solver;
    const double *node_xy;
    const int32_t *elem_nodes, *elem_nn, *elem_mat;
    const double *mat_D;
    const int32_t *cell_elem, *cell_kind;
    const int64_t *cell_tri_ptr;
    const double *tri_xy, *tri_rho, *tri_mid;
    const uint8_t *tri_curved;
};

// one task = one emitted block (regular element matrix, AMORE diagonal or off-diagonal block)
// where the integration kernels store: blk_val in reduction order (csr.cu) through the inverse map of the sort
struct BlkOut {
    double *val;           // [4 Q] one 32-byte record per contribution
    const uint32_t *dst;   // [off_t + n_pair] record of node pair s (task layout); [off_t + s]: record of its transposed copy
    int64_t off_t;
};

struct DevTasks {
    int64_t n_task;
    int32_t *cell;      // [n_task]
    int32_t *e1, *e2;   // [n_task] global element ids (e2 == e1 for diagonal / regular)
    int32_t *slots;     // [n_task] mesh slot j | (l << 8)
    int32_t *cls;       // [n_task] bucket id = class id * N_SUB + sub-bucket
    int64_t *pair_off;  // [n_task] offset of the block in blk_val, in node pairs (x4 doubles)
};

#define MOFEM_CUDA_OK(call)                                                                      \
    do {                                                                                         \
        cudaError_t e__ = (call);                                                                \
        if (e__ != cudaSuccess) {                                                                \
            return ctx_fail(ctx, MOFEM_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
        }                                                                                        \
    } while (0)

// scan.cu
cudaError_t exclusive_scan_i64(const int64_t *in, int64_t *out, int64_t n, int64_t *total_dev, cudaStream_t s,
                               int64_t *launches);
cudaError_t exclusive_scan_i32_to_i64(const int32_t *in, int64_t *out, int64_t n, int64_t *total_dev, cudaStream_t s,
                                      int64_t *launches);

// integrate.cu
cudaError_t launch_cell_counts(const DevProblem &p, int32_t *ntask, int32_t *npair, int32_t *ncoo, int *err_flag,
                               cudaStream_t s);
cudaError_t launch_fill_tasks(const DevProblem &p, const int64_t *task_off, const int64_t *pair_off,
                              const DevTasks &t, int use_fast, cudaStream_t s);
cudaError_t launch_bucket_tasks(const DevTasks &t, int64_t *cls_count /*[N_CLASS] dev, zeroed*/,
                                const int64_t *cls_start /*[N_CLASS] dev or null*/, int32_t *cls_tasks,
                                int phase, cudaStream_t s);
// integrates all tasks of one class; stats[0] += newton updates, stats[1] += evaluations, stats[2] = error flag
cudaError_t launch_integrate_class(int cls, const DevProblem &p, const DevTasks &t, const int32_t *task_list,
                                   int64_t n_list, BlkOut blk_val, unsigned long long *stats, cudaStream_t s);

// all-quad4 overlay cells with 1..3 incident elements and 1..5 straight triangles (class FAST_CLASS): one evaluation per
// (point, element), blocks contracted out of shared memory, one launch.  Symbolic phase: one record per cell in sub-bucket order.
struct alignas(32) CqRec {
    int64_t pair0;     // offset of the cell's first block in blk_val (node pairs); its blocks follow each other, 16 pairs each
    int64_t tri0;      // first integration triangle
    int32_t el[3];     // incident elements in mesh order
    int32_t slots;     // their mesh slots, 8 bits each
};
cudaError_t launch_cq_records(const DevProblem &p, const DevTasks &t, const int32_t *first_task, int64_t n, CqRec *rec, cudaStream_t s);
cudaError_t launch_cell_quads(const DevProblem &p, const CqRec *recs, const int64_t *sub_count /*[N_SUB]*/, BlkOut blk_val,
                              unsigned long long *stats, cudaStream_t s);

// csr.cu
struct DevCsr {
    int64_t n_node;    // node rows
    int64_t n_entry;   // Q: contributions (node pairs incl. transposed copies)
    int64_t n_upair;   // U: unique node pairs
    int64_t *row_start;   // [n_node+1] start of each node row in the entry array
    int64_t *node_ptr;    // [n_node+1] unique pairs per node row (prefix)
    int32_t *node_idx;    // [U] column node of each unique pair
    int32_t *pair_row;    // [U] row node of each unique pair
    int64_t *contrib_ptr; // [U+1] range of contributions of each unique pair
    uint32_t *pair_dst;   // [2 n_pair] record of blk_val (reduction order) that holds node pair s of the task layout; [n_pair + s]: its transposed copy (off-diagonal blocks only)
};
cudaError_t csr_count_rows(const DevProblem &p, const DevTasks &t, int32_t *row_cnt, cudaStream_t s);
cudaError_t csr_fill_rows(const DevProblem &p, const DevTasks &t, const int64_t *row_start, int32_t *row_cursor,
                          uint64_t *entries, cudaStream_t s);
cudaError_t csr_sort_rows(int64_t n_node, const int64_t *row_start, uint64_t *entries, int32_t *row_uniq,
                          uint64_t *scratch, int *err_flag, int phase, cudaStream_t s);
cudaError_t csr_emit(int64_t n_node, const int64_t *row_start, const uint64_t *entries, const int64_t *node_ptr,
                     int32_t *node_idx, int32_t *pair_row, int64_t *contrib_ptr, uint32_t *contrib_src, cudaStream_t s);
cudaError_t csr_pair_dst(const DevCsr &c, const uint32_t *contrib_src /*[Q] sorted sources*/, int64_t n_pair, cudaStream_t s);
cudaError_t csr_numeric(const DevCsr &c, const double *blk_val, double *data, cudaStream_t s);
cudaError_t csr_export_pattern(const DevCsr &c, int32_t *indptr, int32_t *indices, cudaStream_t s);

// sample.cu
cudaError_t sample_count_units(const DevProblem &p, int32_t *cnt, cudaStream_t s);
cudaError_t sample_fill_units(const DevProblem &p, const int64_t *unit_off, int32_t *unit_cell, int64_t *unit_tri, cudaStream_t s);
cudaError_t sample_run(const DevProblem &p, const int32_t *unit_cell, const int64_t *unit_tri, int64_t n_unit, int nS,
                       const double *grid, const double *u, double *out, int *err, cudaStream_t s);

// rho.cu
cudaError_t rho_run(int64_t n_vert, int n_mesh, const double *vert_xy, const int32_t *vert_elem, int64_t n_elem,
                    const int32_t *elem_nv, const double *elem_xy, const double *elem_w, double *rho, int *err,
                    unsigned long long *newton_total, cudaStream_t s);

// overlay.cu
void overlay_sizes(int nx, int ny, int b_lo, int b_hi, int n_mesh, int64_t out[5]);
cudaError_t overlay_run(int nx, int ny, int b_lo, int b_hi, int n_mesh, const double *X0, const double *X1, int32_t *cell_elem,
                        int32_t *cell_kind, int64_t *cell_tri_ptr, double *tri_xy, double *tri_rho, cudaStream_t s);

// pcg.cu
struct DevSystem {
    int64_t n;            // local scalar unknowns (2 * local nodes, ghosts included)
    int64_t n_own;        // scalar rows this rank owns (== n on a single GPU); owned nodes come first
    const int64_t *node_ptr;
    const int32_t *node_idx;
    const double *data;   // scalar CSR values
    int64_t n_blk = 0;    // 2x2 sub-blocks stored in the rows of this rank (node_ptr[n_own / 2])
};

// Row-partitioned operation: local node order is [owned | ghosts grouped by owner rank].
struct DistPlan {
    int world = 1, rank = 0;
    ncclComm_t comm = nullptr;
    int64_t n_owned_node = 0;
    std::vector<int> nb_rank;                  // neighbour ranks
    std::vector<int64_t> send_ptr, recv_ptr;   // [n_nb + 1], in nodes; recv offsets are relative to the ghost segment
    int32_t *send_idx = nullptr;               // device: owned local node ids to send, neighbour by neighbour
    double *sendbuf = nullptr;                 // device: 2 doubles per sent node
    // ---- NVLink peer-memory path (optional; set up by mofem_p2p_export / mofem_p2p_import) ----------------
    bool p2p = false;
    double *xbuf = nullptr;                    // this rank's exchange buffer (cudaMalloc, IPC-exported)
    std::vector<double *> peer_host;           // every rank's exchange buffer mapped into this process
    double **peer = nullptr;                   // device copy of peer_host
    int64_t n_ghost = 0;                       // ghost nodes of this rank (capacity of the halo area)
    std::vector<int64_t> send_remote_off;      // per neighbour: where this rank's entries start in ITS ghost segment
    int32_t *send_nb = nullptr;                // device [n_send]: neighbour slot of every sent node
    int64_t *send_dst = nullptr;               // device [n_send]: index inside this rank's segment of the neighbour's halo area
    int64_t *nb_ghost_off = nullptr;           // device [n_nb]: offset (doubles) of that segment in the neighbour's buffer
    int *nb_rank_dev = nullptr;                // device [n_nb]
    int *wait_rank = nullptr;                  // device [n_wait]: ranks this rank receives ghosts from
    int n_wait = 0;
    int64_t epoch = 0;                         // solves done on the peer path (stamps are unique per solve)
    int32_t *halo_rows = nullptr;              // device: owned rows that reference a ghost column, ascending (built by pcg_setup)
    int64_t n_halo_rows = 0, halo_rows_cap = 0;
};

constexpr int P2P_MAX_WORLD = 64;              // one thread per rank in the mailbox code, shared staging of that size
// Exchange-buffer layout (doubles), identical on every rank (world = W <= P2P_MAX_WORLD):
//   mailbox  [site 0..1][parity 0..1][src rank][4 x u64] = {stamp32 | 32 payload bits} x 4 (two doubles)   16 W words
//   hflag    [parity 0..1][src rank]                                          2 W doubles
//   slab     p, r, dinv, q, x, b, ufix -- the PCG vectors live HERE on the peer path (p first: [owned | ghost nodes][2]),
//            so that neighbours write their boundary entries straight into the ghost segment of p and the SpMV gathers
//            from one contiguous vector
struct P2PDev {
    double *const *peer;   // device array [world]
    double *local;         // == peer[rank]
    int world, rank;
    double stamp;          // stamp of the CURRENT iteration (epoch * 2^32 + it + 1)
    int parity;            // it & 1
    int *err;              // set to 1 when a wait times out
    unsigned int stamp32;  // 32-bit stamp carried inside every mailbox word
};
__host__ __device__ inline int64_t p2p_mbox_off(int W, int site, int parity, int src) { return ((int64_t)(site * 2 + parity) * W + src) * 4; }
__host__ __device__ inline int64_t p2p_hflag_off(int W, int parity, int src) { return 16 * (int64_t)W + (int64_t)parity * W + src; }
__host__ __device__ inline int64_t p2p_p_off(int W) { return (18 * (int64_t)W + 15) & ~(int64_t)15; }   // the slab starts on a 128-byte line
inline int64_t p2p_slab_pad(int64_t n_local_dof) { return (n_local_dof + 31) & ~(int64_t)31; }
// mailboxes + flags + the whole PCG vector slab (p first): one allocation, one L2 persistence window
inline int64_t p2p_buffer_doubles(int W, int64_t n_local_node) { return p2p_p_off(W) + 7 * p2p_slab_pad(2 * n_local_node) + 8; }

cudaError_t spmv(const DevSystem &A, const double *x, double *y, const uint8_t *fixed /*or null*/, cudaStream_t s);

constexpr int PCG_MAX_PART = 4096;
struct PcgBuffers {
    double *x, *r, *p, *q, *dinv, *b, *ufix;  // n each
    double *part_a, *part_b, *part_c;          // PCG_MAX_PART each
    double *sc;                                // 16 scalars (pcg.cu: SC_*)
    uint8_t *row_halo;                         // [local nodes] 1 = row references a ghost column (peer path)
    unsigned int *counters;                    // 8 words: arrival counters of the grid-wide reductions, [6] = time-out flag (peer wait / grid barrier),
                                               // [4] = arrivals of the vector kernel, [7] = largest node-row degree (written by pcg_setup)
    // TMA-ring SpMV (spmv_ring.cuh): tile table built by pcg_setup; capacity in tiles
    void *ring_tiles = nullptr;
    int64_t ring_cap = 0;
    int n_tile = 0;
    unsigned int max_deg = 0;                  // largest node-row degree of the owned rows (pcg_setup)
    int use_ring = 1, use_pdl = 1;             // tuning switches (mofem_set_option)
    double floor_rtol = 1e-8;                  // largest true relative residual accepted as "fp64 floor reached" (converged = 2)
};
int64_t ring_tile_capacity(int64_t n_node, int64_t n_blk);
// builds u_fix, b, D^-1, the ring tile table and (row-partitioned, peer transport) the list of rows that gather ghosts;
// synchronises the stream once (sizes of the compacted lists)
cudaError_t pcg_setup(const DevSystem &A, const double *f, const uint8_t *fixed, const double *fixed_val,
                      PcgBuffers &w, DistPlan *dist, cudaStream_t s, int64_t *launches);
cudaError_t pcg_run(const DevSystem &A, const uint8_t *fixed, PcgBuffers &w, DistPlan *dist, double rtol, int maxit,
                    int chunk, mofem_solve_info_t *info, cudaStream_t s, int64_t *launches, cudaEvent_t *prof, double *prof_ms,
                    int prof_every);
cudaError_t pcg_finish(const DevSystem &A, PcgBuffers &w, const DistPlan *dist, double *u_dev, cudaStream_t s);
cudaError_t energy_of(const DevSystem &A, const double *u, PcgBuffers &w, const DistPlan *dist, double *out_dev, cudaStream_t s);

}  // namespace mofem
