// csr.cu -- deterministic COO -> CSR: integer symbolic phase + segmented reduce (no float atomics).
//
// Replaces scipy.sparse.coo_matrix((V,(I,J))).tocsr() (AMORE.py:125-129 etc.; scipy sparsetools
// coo_tocsr + csr_sort_indices + csr_sum_duplicates).  Every block the reference emits is dense over
// (nodes of element 1) x (nodes of element 2) x 2x2 DOFs, so the whole pattern is determined at NODE level:
// a node pair (a,b) owns the scalar entries (2a+p, 2b+q).  The symbolic phase therefore works on node pairs
// (4x fewer keys than COO triplets):
//   1. count contributions per row node (integer atomics), prefix-sum -> row buckets
//   2. fill buckets with keys (col_node << 32) | src,  src = (node pair in task order << 1) | transposed
//      -- src increases in emission order, so sorting the 64-bit key orders a row by (column, emission order)
//      whatever order the atomics filled the bucket in
//   3. one warp per row: bitonic sort of the bucket (in registers for rows <= 512, through shared memory above), count distinct columns
//   4. prefix-sum -> node_ptr; emit node_idx / contrib_ptr / the sorted sources
//   5. inverse map pair_dst: the record of blk_val every contribution is stored in
// The integration kernels store every 2x2 block value straight into its record (blk_val is in REDUCTION order: the transposed
// copy of an off-diagonal pair is a second 32-byte record), so the numeric phase is a pure, coalesced stream.  Record order:
// the records of 32 consecutive unique node pairs (one warp of the numeric kernel) occupy the range their sorted runs occupy,
// but ROUND-major: first the first contribution of every pair of the group, then the second of those that have one, ...
// (a sliced-ELL layout without padding).  In round j the lanes that still have a contribution read consecutive records; every
// pair still sums its contributions in emission order -- a fixed order: bit-reproducible run to run -- and writes the 2x2
// sub-block into the scalar CSR rows 2a and 2a+1.
// Scalar CSR layout: row 2a = [4*node_ptr[a], +2*deg), row 2a+1 = [4*node_ptr[a]+2*deg, 4*node_ptr[a+1]);
// columns sorted, explicit zeros kept -- identical to scipy's canonical form.
#include "common.cuh"

namespace mofem {

// ---- 1/2: row buckets -------------------------------------------------------------------
// thread per task: the block (e1 x e2) contributes n2 entries to each row node of e1 and, if off-diagonal,
// n1 transposed entries to each row node of e2.
__global__ void k_count_rows(DevProblem p, DevTasks t, int32_t *__restrict__ row_cnt)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= t.n_task) return;
    const int e1 = t.e1[k], e2 = t.e2[k];
    const int n1 = p.elem_nn[e1], n2 = p.elem_nn[e2];
    const int32_t *en1 = p.elem_nodes + (int64_t)e1 * 9, *en2 = p.elem_nodes + (int64_t)e2 * 9;
    for (int a = 0; a < n1; a++) atomicAdd(&row_cnt[en1[a]], n2);
    if (e1 != e2)
        for (int b = 0; b < n2; b++) atomicAdd(&row_cnt[en2[b]], n1);
}

__global__ void k_fill_rows(DevProblem p, DevTasks t, const int64_t *__restrict__ row_start,
                            int32_t *__restrict__ row_cursor, uint64_t *__restrict__ entries)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= t.n_task) return;
    const int e1 = t.e1[k], e2 = t.e2[k];
    const int n1 = p.elem_nn[e1], n2 = p.elem_nn[e2];
    const int32_t *en1 = p.elem_nodes + (int64_t)e1 * 9, *en2 = p.elem_nodes + (int64_t)e2 * 9;
    const uint64_t po = (uint64_t)t.pair_off[k];
    for (int a = 0; a < n1; a++) {
        const int row = en1[a];
        const int64_t pos = row_start[row] + atomicAdd(&row_cursor[row], n2);
        for (int b = 0; b < n2; b++)
            entries[pos + b] = ((uint64_t)(uint32_t)en2[b] << 32) | ((po + (uint64_t)(a * n2 + b)) << 1);
    }
    if (e1 != e2) {
        for (int b = 0; b < n2; b++) {
            const int row = en2[b];
            const int64_t pos = row_start[row] + atomicAdd(&row_cursor[row], n1);
            for (int a = 0; a < n1; a++)
                entries[pos + a] = ((uint64_t)(uint32_t)en1[a] << 32) | ((po + (uint64_t)(a * n2 + b)) << 1) | 1u;
        }
    }
}

// ---- 3: per-row sort -----------------------------------------------------------------------
constexpr int SORT_WARPS = 4;
constexpr int SORT_SMEM_KEYS = 1024;  // per warp; longer rows take the global-memory path

__device__ __forceinline__ void warp_bitonic(uint64_t *keys, int P, int lane)
{
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncwarp();
            for (int i = lane; i < (P >> 1); i += 32) {
                const int lo = (i / stride) * (stride << 1) + (i % stride);
                const int hi = lo + stride;
                const bool up = ((lo & size) == 0);
                const uint64_t a = keys[lo], b = keys[hi];
                if ((a > b) == up) { keys[lo] = b; keys[hi] = a; }
            }
        }
    }
    __syncwarp();
}

__global__ void __launch_bounds__(SORT_WARPS * 32)
k_sort_rows(int64_t n_node, const int64_t *__restrict__ row_start, uint64_t *__restrict__ entries,
            int32_t *__restrict__ row_uniq, uint64_t *__restrict__ scratch, int *err_flag, int64_t min_len)
{
    __shared__ uint64_t sm[SORT_WARPS][SORT_SMEM_KEYS];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t row = (int64_t)blockIdx.x * SORT_WARPS + w;
    if (row >= n_node) return;
    const int64_t b = row_start[row];
    const int64_t L = row_start[row + 1] - b;
    if (L <= min_len) return;        // short rows were sorted by k_sort_rows_rank
    uint64_t *keys;
    int P = 1;
    while (P < L) P <<= 1;
    if (P <= SORT_SMEM_KEYS) keys = sm[w];
    else {
        // long row: sort in a global scratch segment of the same padded size (rare: > 1024 contributions to one node row)
        if (scratch == nullptr) { if (lane == 0) *err_flag = MOFEM_E_LIMIT; return; }
        keys = scratch + 2 * b;  // P < 2L, segments are disjoint
    }
    for (int i = lane; i < P; i += 32) keys[i] = (i < L) ? entries[b + i] : ~0ull;
    warp_bitonic(keys, P, lane);
    int uniq = 0;
    for (int64_t i = lane; i < L; i += 32) {
        const uint64_t kx = keys[i];
        entries[b + i] = kx;
        if (i == 0 || (uint32_t)(keys[i - 1] >> 32) != (uint32_t)(kx >> 32)) uniq++;
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) uniq += __shfl_xor_sync(0xffffffffu, uniq, d);
    if (lane == 0) row_uniq[row] = uniq;
}

// Rows of <= RANK_MAX contributions (the bench matrix averages 71 per node row, the 3-mesh family 130): bitonic network IN REGISTERS.  A warp sorts
// one row; lane l holds the M = 2 | 4 | 8 | 16 consecutive keys  l*M .. l*M + M-1  of the row padded to 32*M with ~0.  Compare-exchange
// steps whose partner is M or more keys away go through shuffles (the partner sits in lane l ^ (stride / M), same slot), closer
// ones stay inside the lane; no shared-memory traffic and no barriers between the steps -- only the load is staged through
// shared memory (coalesced read of the row, blocked layout for the lanes).  Measured against the O(L^2) rank sort it replaces
// (every lane counts, over one broadcast pass through the row, how many keys are smaller than its own): see DESIGN.md.  Longer rows
// are left to k_sort_rows (flagged through *long_rows).
constexpr int RANK_WARPS = 8;
constexpr int RANK_MAX = 512;

__device__ __forceinline__ uint64_t shfl_xor_u64(uint64_t v, int lane_mask)
{
    const uint32_t lo = __shfl_xor_sync(0xffffffffu, (uint32_t)v, lane_mask);
    const uint32_t hi = __shfl_xor_sync(0xffffffffu, (uint32_t)(v >> 32), lane_mask);
    return ((uint64_t)hi << 32) | lo;
}

template <int M>
__device__ __forceinline__ void warp_bitonic_regs(uint64_t (&k)[M], int lane)
{
    const int i0 = lane * M;       // index of this lane's first key
#pragma unroll
    for (int size = 2; size <= 32 * M; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride >= M) {     // partner in another lane, same slot
                const bool up = (i0 & size) == 0, lower = (i0 & stride) == 0;
                const bool take_min = (up == lower);
#pragma unroll
                for (int r = 0; r < M; r++) {
                    const uint64_t o = shfl_xor_u64(k[r], stride / M);
                    const bool mine_smaller = k[r] < o;
                    k[r] = (mine_smaller == take_min) ? k[r] : o;
                }
            } else {               // partner inside the lane: slots r and r ^ stride
#pragma unroll
                for (int r = 0; r < M; r++) {
                    if ((r & stride) == 0) {
                        const int r2 = r | stride;
                        const bool up = (size >= M) ? ((i0 & size) == 0) : ((r & size) == 0);
                        const uint64_t a = k[r], c = k[r2];
                        const bool sw = (a > c) == up;
                        k[r] = sw ? c : a; k[r2] = sw ? a : c;
                    }
                }
            }
        }
    }
}

// sorts the row staged in `keys` (L valid, padded here), writes it back to entries[b ..] and returns the lane's count of
// distinct columns
template <int M>
__device__ __forceinline__ int sort_row_regs(uint64_t *__restrict__ keys, int L, int lane, uint64_t *__restrict__ out)
{
    uint64_t k[M];
#pragma unroll
    for (int r = 0; r < M; r++) { const int i = lane * M + r; k[r] = i < L ? keys[i] : ~0ull; }
    warp_bitonic_regs<M>(k, lane);
    // the key before this lane's first one (distinct-column test across the lane boundary)
    uint64_t prev;
    {
        const uint32_t lo = __shfl_up_sync(0xffffffffu, (uint32_t)k[M - 1], 1);
        const uint32_t hi = __shfl_up_sync(0xffffffffu, (uint32_t)(k[M - 1] >> 32), 1);
        prev = ((uint64_t)hi << 32) | lo;
    }
    int uniq = 0;
#pragma unroll
    for (int r = 0; r < M; r++) {
        const int i = lane * M + r;
        if (i < L) {
            out[i] = k[r];
            if (i == 0 || (uint32_t)(prev >> 32) != (uint32_t)(k[r] >> 32)) uniq++;
        }
        prev = k[r];
    }
    return uniq;
}

__global__ void __launch_bounds__(RANK_WARPS * 32)
k_sort_rows_rank(int64_t n_node, const int64_t *__restrict__ row_start, uint64_t *__restrict__ entries,
                 int32_t *__restrict__ row_uniq, int *long_rows)
{
    __shared__ uint64_t sm_in[RANK_WARPS][RANK_MAX];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t row = (int64_t)blockIdx.x * RANK_WARPS + w;
    if (row >= n_node) return;
    const int64_t b = row_start[row];
    const int64_t L64 = row_start[row + 1] - b;
    if (L64 == 0) { if (lane == 0) row_uniq[row] = 0; return; }
    if (L64 > RANK_MAX) { if (lane == 0) *long_rows = 1; return; }
    const int L = (int)L64;
    uint64_t *keys = sm_in[w];
    for (int i = lane; i < L; i += 32) keys[i] = entries[b + i];
    __syncwarp();
    int uniq;
    if (L <= 64) uniq = sort_row_regs<2>(keys, L, lane, entries + b);
    else if (L <= 128) uniq = sort_row_regs<4>(keys, L, lane, entries + b);
    else if (L <= 256) uniq = sort_row_regs<8>(keys, L, lane, entries + b);
    else uniq = sort_row_regs<16>(keys, L, lane, entries + b);
#pragma unroll
    for (int d = 16; d; d >>= 1) uniq += __shfl_xor_sync(0xffffffffu, uniq, d);
    if (lane == 0) row_uniq[row] = uniq;
}

// ---- 4: emit node-level CSR + contribution lists ------------------------------------------
__global__ void __launch_bounds__(128)
k_emit(int64_t n_node, const int64_t *__restrict__ row_start, const uint64_t *__restrict__ entries,
       const int64_t *__restrict__ node_ptr, int32_t *__restrict__ node_idx, int32_t *__restrict__ pair_row,
       int64_t *__restrict__ contrib_ptr, uint32_t *__restrict__ contrib_src)
{
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n_node) return;
    const int64_t b = row_start[row], L = row_start[row + 1] - b;
    int64_t u = node_ptr[row];
    for (int64_t base = 0; base < L; base += 32) {
        const int64_t i = base + lane;
        bool head = false;
        uint64_t kx = 0;
        if (i < L) {
            kx = entries[b + i];
            head = (i == 0) || ((uint32_t)(entries[b + i - 1] >> 32) != (uint32_t)(kx >> 32));
            contrib_src[b + i] = (uint32_t)kx;
        }
        const unsigned m = __ballot_sync(0xffffffffu, head);
        if (head) {
            const int64_t pos = u + __popc(m & ((1u << lane) - 1));
            node_idx[pos] = (int32_t)(kx >> 32);
            pair_row[pos] = (int32_t)row;
            contrib_ptr[pos] = b + i;
        }
        u += __popc(m);
    }
}

// ---- 5: record of every contribution (inverse map of the round-major order) -----------------------------------
// thread per unique node pair, warp = group of 32 consecutive pairs; the numeric kernel walks the same rounds
__global__ void __launch_bounds__(256)
k_pair_dst(int64_t n_upair, const int64_t *__restrict__ contrib_ptr, const uint32_t *__restrict__ contrib_src,
           uint32_t *__restrict__ pair_dst, int64_t n_pair)
{
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool valid = u < n_upair;
    const int64_t cb = valid ? contrib_ptr[u] : 0;
    const int len = valid ? (int)(contrib_ptr[u + 1] - cb) : 0;
    int64_t base = __shfl_sync(0xffffffffu, cb, 0);
    int maxlen = len;
#pragma unroll
    for (int d = 16; d; d >>= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, d));
    for (int j = 0; j < maxlen; j++) {
        const unsigned m = __ballot_sync(0xffffffffu, len > j);
        if (len > j) {
            const uint32_t src = contrib_src[cb + j];
            pair_dst[((src & 1u) ? n_pair : 0) + (int64_t)(src >> 1)] = (uint32_t)(base + __popc(m & ((1u << lane) - 1)));
        }
        base += __popc(m);
    }
}

// ---- numeric: segmented reduce in emission order ---------------------------------------------
// thread per unique node pair, warp = group of 32 consecutive pairs whose records are stored round-major (see above): in round j
// the lanes with more than j contributions read consecutive 32-byte records (one 256-bit load each)
__global__ void __launch_bounds__(256)
k_numeric(int64_t n_upair, const int64_t *__restrict__ node_ptr, const int32_t *__restrict__ pair_row,
          const int64_t *__restrict__ contrib_ptr, const double *__restrict__ blk_val, double *__restrict__ data)
{
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool valid = u < n_upair;
    const int64_t cb = valid ? contrib_ptr[u] : 0;
    const int len = valid ? (int)(contrib_ptr[u + 1] - cb) : 0;
    int64_t base = __shfl_sync(0xffffffffu, cb, 0);
    int maxlen = len;
#pragma unroll
    for (int d = 16; d; d >>= 1) maxlen = max(maxlen, __shfl_xor_sync(0xffffffffu, maxlen, d));
    double kxx = 0.0, kxy = 0.0, kyx = 0.0, kyy = 0.0;
    for (int j = 0; j < maxlen; j++) {
        const unsigned m = __ballot_sync(0xffffffffu, len > j);
        if (len > j) {
            double a0, a1, a2, a3;
            asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a0), "=d"(a1), "=d"(a2), "=d"(a3)
                         : "l"(blk_val + 4 * (base + __popc(m & ((1u << lane) - 1)))));
            kxx += a0; kxy += a1; kyx += a2; kyy += a3;
        }
        base += __popc(m);
    }
    if (!valid) return;
    const int64_t a = pair_row[u];
    const int64_t r0 = node_ptr[a], deg = node_ptr[a + 1] - r0;
    const int64_t kk = u - r0;
    *reinterpret_cast<double2 *>(data + 4 * r0 + 2 * kk) = make_double2(kxx, kxy);
    *reinterpret_cast<double2 *>(data + 4 * r0 + 2 * deg + 2 * kk) = make_double2(kyx, kyy);
}

// scalar CSR index arrays (int32, scipy layout) from the node-level pattern
__global__ void k_export_indptr(int64_t n_node, const int64_t *__restrict__ node_ptr, int32_t *__restrict__ indptr)
{
    const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a > n_node) return;
    if (a == n_node) { indptr[2 * a] = (int32_t)(4 * node_ptr[a]); return; }
    const int64_t r0 = node_ptr[a], deg = node_ptr[a + 1] - r0;
    indptr[2 * a] = (int32_t)(4 * r0);
    indptr[2 * a + 1] = (int32_t)(4 * r0 + 2 * deg);
}

__global__ void k_export_indices(int64_t n_upair, const int64_t *__restrict__ node_ptr, const int32_t *__restrict__ pair_row,
                                 const int32_t *__restrict__ node_idx, int32_t *__restrict__ indices)
{
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_upair) return;
    const int64_t a = pair_row[u];
    const int64_t r0 = node_ptr[a], deg = node_ptr[a + 1] - r0, kk = u - r0;
    const int32_t c = 2 * node_idx[u];
    indices[4 * r0 + 2 * kk] = c; indices[4 * r0 + 2 * kk + 1] = c + 1;
    indices[4 * r0 + 2 * deg + 2 * kk] = c; indices[4 * r0 + 2 * deg + 2 * kk + 1] = c + 1;
}

// ---- launchers -------------------------------------------------------------------------------
cudaError_t csr_count_rows(const DevProblem &p, const DevTasks &t, int32_t *row_cnt, cudaStream_t s)
{
    if (t.n_task == 0) return cudaSuccess;
    k_count_rows<<<(unsigned)((t.n_task + 255) / 256), 256, 0, s>>>(p, t, row_cnt);
    return cudaGetLastError();
}

cudaError_t csr_fill_rows(const DevProblem &p, const DevTasks &t, const int64_t *row_start, int32_t *row_cursor,
                          uint64_t *entries, cudaStream_t s)
{
    if (t.n_task == 0) return cudaSuccess;
    k_fill_rows<<<(unsigned)((t.n_task + 255) / 256), 256, 0, s>>>(p, t, row_start, row_cursor, entries);
    return cudaGetLastError();
}

// phase 0: rank sort of the short rows, *err_flag = 1 if longer rows exist; phase 1: bitonic sort of the rows longer than
// RANK_MAX (shared memory up to 1024 keys, else `scratch`; *err_flag = MOFEM_E_LIMIT if scratch is needed but null)
cudaError_t csr_sort_rows(int64_t n_node, const int64_t *row_start, uint64_t *entries, int32_t *row_uniq,
                          uint64_t *scratch, int *err_flag, int phase, cudaStream_t s)
{
    if (n_node == 0) return cudaSuccess;
    if (phase == 0)
        k_sort_rows_rank<<<(unsigned)((n_node + RANK_WARPS - 1) / RANK_WARPS), RANK_WARPS * 32, 0, s>>>(n_node, row_start, entries,
                                                                                                       row_uniq, err_flag);
    else
        k_sort_rows<<<(unsigned)((n_node + SORT_WARPS - 1) / SORT_WARPS), SORT_WARPS * 32, 0, s>>>(
            n_node, row_start, entries, row_uniq, scratch, err_flag, (int64_t)RANK_MAX);
    return cudaGetLastError();
}

cudaError_t csr_emit(int64_t n_node, const int64_t *row_start, const uint64_t *entries, const int64_t *node_ptr,
                     int32_t *node_idx, int32_t *pair_row, int64_t *contrib_ptr, uint32_t *contrib_src, cudaStream_t s)
{
    if (n_node == 0) return cudaSuccess;
    k_emit<<<(unsigned)((n_node * 32 + 127) / 128), 128, 0, s>>>(n_node, row_start, entries, node_ptr, node_idx,
                                                                  pair_row, contrib_ptr, contrib_src);
    return cudaGetLastError();
}

cudaError_t csr_pair_dst(const DevCsr &c, const uint32_t *contrib_src, int64_t n_pair, cudaStream_t s)
{
    if (c.n_upair == 0) return cudaSuccess;
    k_pair_dst<<<(unsigned)((c.n_upair + 255) / 256), 256, 0, s>>>(c.n_upair, c.contrib_ptr, contrib_src, c.pair_dst, n_pair);
    return cudaGetLastError();
}

cudaError_t csr_numeric(const DevCsr &c, const double *blk_val, double *data, cudaStream_t s)
{
    if (c.n_upair == 0) return cudaSuccess;
    k_numeric<<<(unsigned)((c.n_upair + 255) / 256), 256, 0, s>>>(c.n_upair, c.node_ptr, c.pair_row, c.contrib_ptr,
                                                                   blk_val, data);
    return cudaGetLastError();
}

cudaError_t csr_export_pattern(const DevCsr &c, int32_t *indptr, int32_t *indices, cudaStream_t s)
{
    if (indptr) k_export_indptr<<<(unsigned)((c.n_node + 1 + 255) / 256), 256, 0, s>>>(c.n_node, c.node_ptr, indptr);
    if (indices && c.n_upair)
        k_export_indices<<<(unsigned)((c.n_upair + 255) / 256), 256, 0, s>>>(c.n_upair, c.node_ptr, c.pair_row,
                                                                             c.node_idx, indices);
    return cudaGetLastError();
}

}  // namespace mofem
