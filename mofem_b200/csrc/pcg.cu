// pcg.cu -- Jacobi-preconditioned CG on the assembled matrix (replaces scipy spsolve,
// AMORE.py:264,683,1106; traditionalElement.py:146,296,461).
//
// The Dirichlet elimination of the reference (delete fixed rows/columns, AMORE.py:213-255) is applied as a mask:
// search directions are zero on fixed DOFs and the SpMV zeroes fixed rows, so the iteration is exactly CG on the
// reduced free-free system without building a second matrix.
//
// The matrix is the scalar CSR viewed at node level (one column index per 2x2 sub-block: 9 bytes per stored scalar
// instead of 12).  Two SpMV kernels read it:
//   k_spmv_ring (spmv_ring.cuh)  the SpMV of the PCG loop: a producer warp streams row tiles into a shared-memory ring with
//                                cp.async.bulk (TMA), consumer warps multiply out of shared memory; _p2p variant with the
//                                NVLink halo exchange inside the kernel
//   k_spmv                       row kernel (8 lanes per node row chase pointer -> {index, values} -> x through global
//                                loads, 4 independent load groups per lane, evict-first streaming): mofem_spmv, residual
//                                checks, energy, the NCCL transport, rows too long for a ring stage
// The vector part of an iteration is ONE kernel, k_pcg_vec:
//     r -= alpha q, sums of r.z and r.r  |  arrive  |  x += alpha p  |  wait  |  p = z + beta p
// The second reduction of the iteration (beta needs r.z of the NEW residual) is the arrive/wait pair -- a grid barrier on one
// GPU, the peers' mailboxes on the peer transport -- and the x update, which needs neither sum, sits between the two so that
// the barrier / NVLink latency is spent working.  (A single-reduction PCG, which gets beta from q.z and q.D^-1.q summed
// inside the SpMV, was built and measured: the extra sums cost the SpMV more than the hidden second reduction costs here;
// DESIGN.md has the numbers.)  The kernels are chained with programmatic dependent launch on one GPU and on the peer
// transport: the ring's producer fills its stages while the vector kernel drains.
//
// Dot products have a fixed reduction order (bit-reproducible): per-thread sums over a static assignment, per-CTA
// partials summed in index order, per-rank sums added in rank order.  alpha/beta live on the device; the vector kernel
// that sees r.r <= tol^2 raises a stop flag which turns the rest of the enqueued chunk into no-ops, so the host polls only
// every PCG_CHUNK iterations.
#include <algorithm>

#include "common.cuh"
#include "spmv_ring.cuh"

namespace mofem {

constexpr int VEC_BLOCK = 256;
constexpr int SPMV_BLOCK = 128;
constexpr int SPMV_LANES = 8;    // lanes cooperating on one node row
constexpr int SPMV_UNROLL = 4;   // independent load groups per lane
constexpr int SPMV_MINB = 12;    // resident blocks per SM (42 registers per thread)
constexpr int SPMV_GRID_CAP = 148 * 24;
static_assert(SPMV_GRID_CAP <= PCG_MAX_PART, "partials buffer too small");

// scalar slots in PcgBuffers::sc: (rz, rr) pairs double-buffered by iteration parity at [0,1] and [2,3] -- each pair
// is contiguous so that the row-partitioned solve all-reduces it in one call.
enum { SC_BNORM2 = 4, SC_RR_TRUE = 5, SC_ENERGY = 6, SC_PQ = 7, SC_DONE = 8, SC_IT_DONE = 9, SC_PART_RZ = 10, SC_PART_PQ = 12, SC_COUNT = 16 };
// SC_PART_*: per-rank partials on the NCCL path (all-reduced INTO the global slots, so that re-reducing after the device
// has stopped is idempotent).
// SC_DONE: set by the direction kernel of the iteration whose recurrence residual met the tolerance; every later kernel
// of the chunk returns at once, so the host can poll rarely (large chunks) without paying for overshoot iterations.
__host__ __device__ inline int sc_rz(int it) { return 2 * (it & 1); }
__host__ __device__ inline int sc_rr(int it) { return 2 * (it & 1) + 1; }

// block-wide sum, result valid in every thread; fixed order
__device__ __forceinline__ double block_sum(double v, double *sm)
{
#pragma unroll
    for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sm[w] = v;
    __syncthreads();
    double s = 0.0;
    const int nw = blockDim.x >> 5;
    for (int i = 0; i < nw; i++) s += sm[i];
    return s;
}

// Publishes this block's partial; the last block to arrive returns true after which `total` (valid in every
// thread of that block) is the sum of all partials in index order -- independent of which block came last.
__device__ __forceinline__ bool grid_sum(double partial, double *__restrict__ part, unsigned int *counter, double *sm,
                                         double &total)
{
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
        part[blockIdx.x] = partial;
        __threadfence();
        const unsigned int ticket = atomicAdd(counter, 1u);
        is_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
    double v = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) v += __ldcg(part + i);
    total = block_sum(v, sm);
    if (threadIdx.x == 0) *counter = 0u;
    return true;
}

// ---- NVLink peer-memory primitives (row-partitioned solve, one process per GPU) ---------------------------------
// Every rank owns an exchange buffer that all peers map (CUDA IPC).  A value travels as {payload, stamp}: the writer
// stores the payload into the consumer's buffer, fences at system scope, then stores the stamp; the consumer polls the
// stamp in its OWN memory.  Stamps grow monotonically (epoch, iteration), two parity slots make reuse safe: a rank can
// only be two rounds ahead of a peer after that peer has consumed the older round.
constexpr long long P2P_TIMEOUT_CLOCKS = 40000000000ll;   // ~20 s at 2 GHz: a peer died (or a CTA never became resident)
__device__ __forceinline__ bool p2p_wait(const volatile double *flag, double stamp, int *err)
{
    long long t0 = 0;
    for (unsigned spins = 0; *flag != stamp; spins++) {
        if ((spins & 0xfff) == 0xfff) {          // look at the clock every 4096 polls
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > P2P_TIMEOUT_CLOCKS) { *err = 1; return false; }
        }
    }
    return true;
}

// All-reduce of one or two doubles in "LL" form (the trick NCCL's low-latency protocol uses): every 8-byte word carries
// 32 bits of payload and a 32-bit stamp, and an aligned 8-byte store is atomic over NVLink -- so no fence and no
// separate flag: one-way latency only.
__device__ __forceinline__ unsigned int p2p_stamp32(const P2PDev &pp) { return pp.stamp32; }

// producer side: called by ONE block after it holds this rank's partial(s)
__device__ __forceinline__ void p2p_publish(const P2PDev &pp, int site, double v0, double v1)
{
    if ((int)threadIdx.x < pp.world) {
        volatile unsigned long long *slot =
            reinterpret_cast<volatile unsigned long long *>(pp.peer[threadIdx.x] + p2p_mbox_off(pp.world, site, pp.parity, pp.rank));
        const unsigned long long f = (unsigned long long)pp.stamp32 << 32;
        const unsigned long long a = (unsigned long long)__double_as_longlong(v0), c = (unsigned long long)__double_as_longlong(v1);
        slot[0] = f | (a & 0xffffffffull); slot[1] = f | (a >> 32);
        slot[2] = f | (c & 0xffffffffull); slot[3] = f | (c >> 32);
    }
}

// consumer side: every block sums the W partials in rank order (bit-identical on every rank and block)
__device__ __forceinline__ void p2p_collect(const P2PDev &pp, int site, double &v0, double &v1)
{
    __shared__ double sv[2][P2P_MAX_WORLD];
    if ((int)threadIdx.x < pp.world) {
        const volatile unsigned long long *slot =
            reinterpret_cast<const volatile unsigned long long *>(pp.local + p2p_mbox_off(pp.world, site, pp.parity, threadIdx.x));
        unsigned long long wv[4];
        long long t0 = 0;
        for (unsigned spins = 0;; spins++) {
            bool ok = true;
#pragma unroll
            for (int j = 0; j < 4; j++) { wv[j] = slot[j]; ok = ok && (unsigned int)(wv[j] >> 32) == pp.stamp32; }
            if (ok) break;
            if ((spins & 0xfff) == 0xfff) {
                const long long now = clock64();
                if (t0 == 0) t0 = now;
                else if (now - t0 > P2P_TIMEOUT_CLOCKS) { *pp.err = 1; break; }
            }
        }
        sv[0][threadIdx.x] = __longlong_as_double((long long)((wv[0] & 0xffffffffull) | (wv[1] << 32)));
        sv[1][threadIdx.x] = __longlong_as_double((long long)((wv[2] & 0xffffffffull) | (wv[3] << 32)));
    }
    __syncthreads();
    double a = 0.0, c = 0.0;
    for (int r = 0; r < pp.world; r++) { a += sv[0][r]; c += sv[1][r]; }
    v0 = a; v1 = c;
    __syncthreads();
}

struct HaloPush {          // peer path: what a kernel needs to deliver this rank's boundary entries of p to the neighbours
    int64_t n_send;
    const int32_t *send_idx, *send_nb;
    const int64_t *send_dst, *nb_ghost_off;
    const int *nb_rank;
    int n_nb;
};

// Ghost entries of x are written by the neighbours WHILE a peer-transport SpMV runs, so the rows that gather them must not
// read x through the non-coherent path (ld.global.nc is only defined for data that stays constant for the kernel's
// lifetime, and an interior row may already have pulled the sector shared by the last owned and the first ghost node into
// L1): COHERENT loads x with ld.global.cg (L2, the point the NVLink stores land in).
// one node row (two scalar rows) by a group of SPMV_LANES lanes; the sums are valid in every lane of the group
template <bool COHERENT = false>
__device__ __forceinline__ void spmv_row(int64_t a, int sub, const int64_t *__restrict__ node_ptr,
                                         const int32_t *__restrict__ node_idx, const double *__restrict__ data,
                                         const double *x, double &y0, double &y1)
{
    const int64_t r0 = node_ptr[a];
    const int deg = (int)(node_ptr[a + 1] - r0);
    const double2 *d0 = reinterpret_cast<const double2 *>(data + 4 * r0);
    const double2 *d1 = reinterpret_cast<const double2 *>(data + 4 * r0 + 2 * deg);
    for (int kb = sub; kb < deg; kb += SPMV_LANES * SPMV_UNROLL) {
        double2 v0[SPMV_UNROLL], v1[SPMV_UNROLL], xv[SPMV_UNROLL];
#pragma unroll
        for (int u = 0; u < SPMV_UNROLL; u++) {
            const int kk = kb + u * SPMV_LANES;
            if (kk < deg) {
                const int c = __ldg(node_idx + r0 + kk);
                v0[u] = __ldcs(d0 + kk);
                v1[u] = __ldcs(d1 + kk);
                xv[u] = COHERENT ? __ldcg(reinterpret_cast<const double2 *>(x + 2 * (int64_t)c))
                                 : __ldg(reinterpret_cast<const double2 *>(x + 2 * (int64_t)c));
            } else {
                v0[u] = v1[u] = xv[u] = make_double2(0.0, 0.0);
            }
        }
#pragma unroll
        for (int u = 0; u < SPMV_UNROLL; u++) {
            y0 = fma(v0[u].x, xv[u].x, y0); y0 = fma(v0[u].y, xv[u].y, y0);
            y1 = fma(v1[u].x, xv[u].x, y1); y1 = fma(v1[u].y, xv[u].y, y1);
        }
    }
}

// y = K x on node rows; fixed rows -> 0.  Optionally the dot product x.y goes to *dot_out.  Ghost columns are the tail
// of x.
// P2P (row-partitioned, peer transport) -- one kernel does the halo exchange, the product and the all-reduce hand-off:
//   * block 0 first writes this rank's boundary entries of x into the ghost segment of the neighbours' x over NVLink and
//     raises its stamp there;
//   * every block first multiplies its rows that touch no ghost column; the few that do (halo_rows) are dealt out afterwards,
//     behind a wait for the neighbours' stamps -- the NVLink latency hides behind the interior rows;
//   * the last block publishes the dot-product partial into every peer's mailbox.
template <bool DOT, bool P2P>
__global__ void __launch_bounds__(SPMV_BLOCK, SPMV_MINB)
k_spmv(int64_t n_node, const int64_t *__restrict__ node_ptr, const int32_t *__restrict__ node_idx,
       const double *__restrict__ data, const double *x, double *__restrict__ y,
       const uint8_t *__restrict__ fixed, double *__restrict__ part_dot, unsigned int *counter, double *dot_out,
       P2PDev pp, int n_wait, const int *__restrict__ wait_rank, const uint8_t *__restrict__ row_halo,
       const int32_t *__restrict__ halo_rows, int64_t n_halo, HaloPush hp, const double *done_flag)
{
    __shared__ double sm[SPMV_BLOCK / 32];
    ring::grid_dep_wait();   // PDL: no-op unless launched with the programmatic-serialization attribute
    if (DOT && done_flag && __ldcg(done_flag) != 0.0) return;
    if (P2P && blockIdx.x == 0 && hp.n_nb) {
        const double2 *x2 = reinterpret_cast<const double2 *>(x);
        for (int64_t i = threadIdx.x; i < hp.n_send; i += blockDim.x) {
            const int k = hp.send_nb[i];
            double2 *dst = reinterpret_cast<double2 *>(pp.peer[hp.nb_rank[k]] + hp.nb_ghost_off[k]) + hp.send_dst[i];
            *dst = x2[hp.send_idx[i]];
        }
        __threadfence_system();
        __syncthreads();
        if ((int)threadIdx.x < hp.n_nb)
            *reinterpret_cast<volatile double *>(pp.peer[hp.nb_rank[threadIdx.x]] + p2p_hflag_off(pp.world, pp.parity, pp.rank)) = pp.stamp;
    }
    const int sub = threadIdx.x % SPMV_LANES;
    constexpr int64_t rows_per_block = SPMV_BLOCK / SPMV_LANES;
    double dot = 0.0;
    auto finish_row = [&](int64_t a, double y0, double y1) {
#pragma unroll
        for (int d = SPMV_LANES / 2; d; d >>= 1) {
            y0 += __shfl_xor_sync(0xffffffffu, y0, d);
            y1 += __shfl_xor_sync(0xffffffffu, y1, d);
        }
        if (a >= 0 && sub == 0) {
            if (fixed) {
                const uchar2 fx = *reinterpret_cast<const uchar2 *>(fixed + 2 * a);
                if (fx.x) y0 = 0.0;
                if (fx.y) y1 = 0.0;
            }
            *reinterpret_cast<double2 *>(y + 2 * a) = make_double2(y0, y1);
            if (DOT) {
                const double2 xa = *reinterpret_cast<const double2 *>(x + 2 * a);
                dot = fma(xa.x, y0, dot); dot = fma(xa.y, y1, dot);
            }
        }
    };
    for (int64_t row0 = (int64_t)blockIdx.x * rows_per_block; row0 < n_node; row0 += (int64_t)gridDim.x * rows_per_block) {
        int64_t a = row0 + threadIdx.x / SPMV_LANES;
        if (a >= n_node || (P2P && row_halo[a])) a = -1;        // halo rows come last (below)
        double y0 = 0.0, y1 = 0.0;
        if (a >= 0) spmv_row(a, sub, node_ptr, node_idx, data, x, y0, y1);
        finish_row(a, y0, y1);
    }
    if (P2P) {
        // rows that gather ghost entries: a compact list, dealt block by block AFTER the interior rows, behind a wait for
        // the neighbours' stamps -- by then the halo has long arrived and the NVLink latency is hidden
        const int64_t h0 = (int64_t)blockIdx.x * rows_per_block;
        if (h0 < n_halo) {
            if ((int)threadIdx.x < n_wait)
                p2p_wait(reinterpret_cast<const volatile double *>(pp.local + p2p_hflag_off(pp.world, pp.parity, wait_rank[threadIdx.x])),
                         pp.stamp, pp.err);
            __syncthreads();
            __threadfence();
            for (int64_t i0 = h0; i0 < n_halo; i0 += (int64_t)gridDim.x * rows_per_block) {
                const int64_t i = i0 + threadIdx.x / SPMV_LANES;
                const int64_t a = i < n_halo ? halo_rows[i] : -1;
                double y0 = 0.0, y1 = 0.0;
                if (a >= 0) spmv_row<true>(a, sub, node_ptr, node_idx, data, x, y0, y1);
                finish_row(a, y0, y1);
            }
        }
    }
    ring::grid_dep_launch();   // the dependent (update kernel) may become resident from here on
    if (DOT) {
        const double s = block_sum(dot, sm);
        double total;
        if (grid_sum(s, part_dot, counter, sm, total)) {
            if (P2P) p2p_publish(pp, 0, total, 0.0);
            else if (threadIdx.x == 0) *dot_out = total;
        }
    }
}

// ---- TMA-ring SpMV of the single-GPU PCG loop (spmv_ring.cuh) -------------------------------------------------------
// 2 CTAs per SM, 12 consumer warps + 1 producer warp each, 2 stages of <= 960 rows + sub-blocks (tuned with
// tools/spmv_lab2.cu inside a PCG-shaped loop: 85 us against 97 us for the row kernel on the 1 M-cell bench matrix).
using PcgRing = RingCfg<960, 128, 2, 12>;
constexpr int RING_GRID = 148 * 2;
static_assert(RING_GRID <= PCG_MAX_PART, "partials buffer too small");

int64_t ring_tile_capacity(int64_t n_node, int64_t n_blk) { return (n_node + n_blk) / PcgRing::SPAN + 2; }

// tile t = node rows a with  t*SPAN <= a + node_ptr[a] < (t+1)*SPAN  (the key grows by 1 + degree per row, so a tile
// holds <= SPAN rows and < SPAN + maxdeg sub-blocks); one thread per tile, two binary searches
__global__ void k_ring_tiles(int64_t n_node, const int64_t *__restrict__ node_ptr, int n_tile, RingTile *__restrict__ tiles,
                             const uint8_t *__restrict__ row_mask)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tile) return;
    auto first_row_with_key_ge = [&](int64_t key) {
        int64_t lo = 0, hi = n_node;          // answer in [0, n_node]
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (mid + node_ptr[mid] >= key) hi = mid; else lo = mid + 1;
        }
        return lo;
    };
    const int64_t a0 = first_row_with_key_ge((int64_t)t * PcgRing::SPAN), a1 = first_row_with_key_ge((int64_t)(t + 1) * PcgRing::SPAN);
    RingTile rt;
    rt.a0 = (int32_t)a0; rt.a1 = (int32_t)a1; rt.b0 = node_ptr[a0]; rt.b1 = node_ptr[a1]; rt.flags = 0;
    if (row_mask)
        for (int64_t a = a0; a < a1; a++) if (row_mask[a]) { rt.flags = 1; break; }
    tiles[t] = rt;
}

// q = K p and the CTA's partial of p.q (part_dot[blockIdx.x]: the vector kernel adds the partials up, so this kernel has no
// "last CTA" tail).  Launched with the PDL attribute the producer warp fills the ring while the vector kernel of the previous
// iteration is still draining (the matrix does not depend on it).
__global__ void __launch_bounds__(PcgRing::THREADS, 2)
k_spmv_ring(int n_tile, const RingTile *__restrict__ tiles, const int64_t *__restrict__ node_ptr,
            const int32_t *__restrict__ node_idx, const double *__restrict__ data, const double *__restrict__ x,
            double *__restrict__ y, const uint8_t *__restrict__ fixed, double *__restrict__ part_dot, const double *done_flag)
{
    extern __shared__ __align__(128) unsigned char ring_smem[];
    __shared__ double sm[32];
    double dot = 0.0;
    if (!ring_spmv_body<PcgRing, true>(ring_smem, n_tile, tiles, node_ptr, node_idx, data, x, y, fixed, dot, nullptr, done_flag)) return;
    const double s = block_sum(dot, sm);
    if (threadIdx.x == 0) part_dot[blockIdx.x] = s;
}

// Row-partitioned, peer transport: the same ring for the rows that touch no ghost column, with the halo exchange of k_spmv
// around it -- CTA 0 first stores this rank's boundary entries of x into the neighbours' ghost segments over NVLink and raises
// its stamp there; the rows that gather ghosts (a compact list, their tiles carry the mask flag) are multiplied after the ring,
// behind a wait for the neighbours' stamps; the last CTA publishes the p.q partial into every peer's mailbox.
__device__ __forceinline__ void consumer_barrier() { asm volatile("bar.sync 1, %0;" ::"r"(PcgRing::CW * 32) : "memory"); }

__global__ void __launch_bounds__(PcgRing::THREADS, 2)
k_spmv_ring_p2p(int n_tile, const RingTile *__restrict__ tiles, int64_t n_node, const int64_t *__restrict__ node_ptr,
                const int32_t *__restrict__ node_idx, const double *__restrict__ data, const double *x,
                double *__restrict__ y, const uint8_t *__restrict__ fixed, double *__restrict__ part_dot, unsigned int *counter,
                P2PDev pp, int n_wait, const int *__restrict__ wait_rank, const uint8_t *__restrict__ row_halo,
                const int32_t *__restrict__ halo_rows, int64_t n_halo, HaloPush hp, const double *done_flag)
{
    extern __shared__ __align__(128) unsigned char ring_smem[];
    __shared__ double sm[32];
    const bool consumer = (int)(threadIdx.x >> 5) < PcgRing::CW;
    constexpr int NC = PcgRing::CW * 32;
    if (blockIdx.x == 0 && hp.n_nb && consumer) {
        ring::grid_dep_wait();                     // x is the previous kernel's output (the producer warp does not wait here)
        if (!(done_flag && __ldcg(done_flag) != 0.0)) {
            const double2 *x2 = reinterpret_cast<const double2 *>(x);
            for (int64_t i = threadIdx.x; i < hp.n_send; i += NC) {
                const int k = hp.send_nb[i];
                double2 *dst = reinterpret_cast<double2 *>(pp.peer[hp.nb_rank[k]] + hp.nb_ghost_off[k]) + hp.send_dst[i];
                *dst = x2[hp.send_idx[i]];
            }
            __threadfence_system();
            consumer_barrier();
            if ((int)threadIdx.x < hp.n_nb)
                *reinterpret_cast<volatile double *>(pp.peer[hp.nb_rank[threadIdx.x]] + p2p_hflag_off(pp.world, pp.parity, pp.rank)) = pp.stamp;
        }
    }
    double dot = 0.0;
    if (!ring_spmv_body<PcgRing, false, true>(ring_smem, n_tile, tiles, node_ptr, node_idx, data, x, y, fixed, dot, row_halo, done_flag)) return;
    constexpr int64_t rows_per_block = NC / SPMV_LANES;
    const int64_t h0 = (int64_t)blockIdx.x * rows_per_block;
    if (consumer && h0 < n_halo) {
        if ((int)threadIdx.x < n_wait)
            p2p_wait(reinterpret_cast<const volatile double *>(pp.local + p2p_hflag_off(pp.world, pp.parity, wait_rank[threadIdx.x])),
                     pp.stamp, pp.err);
        consumer_barrier();
        __threadfence();
        const int sub = threadIdx.x % SPMV_LANES;
        for (int64_t i0 = h0; i0 < n_halo; i0 += (int64_t)gridDim.x * rows_per_block) {
            const int64_t i = i0 + threadIdx.x / SPMV_LANES;
            const int64_t a = i < n_halo ? halo_rows[i] : -1;
            double y0 = 0.0, y1 = 0.0;
            if (a >= 0) spmv_row<true>(a, sub, node_ptr, node_idx, data, x, y0, y1);
#pragma unroll
            for (int d = SPMV_LANES / 2; d; d >>= 1) {
                y0 += __shfl_xor_sync(0xffffffffu, y0, d);
                y1 += __shfl_xor_sync(0xffffffffu, y1, d);
            }
            if (a >= 0 && sub == 0) {
                if (fixed) {
                    const uchar2 fx = *reinterpret_cast<const uchar2 *>(fixed + 2 * a);
                    if (fx.x) y0 = 0.0;
                    if (fx.y) y1 = 0.0;
                }
                *reinterpret_cast<double2 *>(y + 2 * a) = make_double2(y0, y1);
                const double2 xa = *reinterpret_cast<const double2 *>(x + 2 * a);
                dot = fma(xa.x, y0, dot); dot = fma(xa.y, y1, dot);
            }
        }
    }
    (void)n_node;
    ring::grid_dep_launch();       // the vector kernel may become resident from here on
    const double s = block_sum(dot, sm);
    double total;
    if (grid_sum(s, part_dot, counter, sm, total)) p2p_publish(pp, 0, total, 0.0);
}

// kernel launch with (optionally) the programmatic-stream-serialization attribute: the kernel may start while its
// predecessor in the stream drains; it orders itself with griddepcontrol.wait
template <typename... KArgs, typename... Args>
static cudaError_t launch_ex(bool pdl, void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t s, Args... args)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// rows of the local matrix that reference a ghost column (peer path: they are multiplied after the halo has arrived)
__global__ void k_mark_halo_rows(int64_t n_node, const int64_t *__restrict__ node_ptr, const int32_t *__restrict__ node_idx,
                                 uint8_t *__restrict__ row_halo)
{
    const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n_node) return;
    uint8_t h = 0;
    for (int64_t k = node_ptr[a]; k < node_ptr[a + 1]; k++) h |= (node_idx[k] >= n_node);
    row_halo[a] = h;
}

// diagonal -> inverse (0 on fixed or empty rows)
__global__ void k_diag_inv(int64_t n_node, const int64_t *__restrict__ node_ptr, const int32_t *__restrict__ node_idx,
                           const double *__restrict__ data, const uint8_t *__restrict__ fixed, double *__restrict__ dinv,
                           unsigned int *max_deg)
{
    const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int dg = a < n_node ? (unsigned int)(node_ptr[a + 1] - node_ptr[a]) : 0u;
#pragma unroll
    for (int d = 16; d; d >>= 1) dg = max(dg, __shfl_xor_sync(0xffffffffu, dg, d));
    if ((threadIdx.x & 31) == 0) atomicMax(max_deg, dg);
    if (a >= n_node) return;
    const int64_t r0 = node_ptr[a], deg = node_ptr[a + 1] - r0;
    double dx = 0.0, dy = 0.0;
    for (int64_t kk = 0; kk < deg; kk++)
        if (node_idx[r0 + kk] == a) { dx = data[4 * r0 + 2 * kk]; dy = data[4 * r0 + 2 * deg + 2 * kk + 1]; break; }
    dinv[2 * a] = (fixed[2 * a] || dx == 0.0) ? 0.0 : 1.0 / dx;
    dinv[2 * a + 1] = (fixed[2 * a + 1] || dy == 0.0) ? 0.0 : 1.0 / dy;
}

__global__ void k_build_ufix(int64_t n, const uint8_t *__restrict__ fixed, const double *__restrict__ val, double *__restrict__ ufix)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ufix[i] = fixed[i] ? val[i] : 0.0;
}
// b = free ? f - K u_fix : 0   (AMORE.py:227 followed by the row deletions)
__global__ void k_build_rhs(int64_t n, const uint8_t *__restrict__ fixed, const double *__restrict__ f,
                            const double *__restrict__ Kufix, double *__restrict__ b)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) b[i] = fixed[i] ? 0.0 : f[i] - Kufix[i];
}

// Publishes two partials per block under one counter; the last block sums both lists in index order.
__device__ __forceinline__ bool grid_sum2(double s1, double s2, double *__restrict__ part1, double *__restrict__ part2,
                                          unsigned int *counter, double *sm, double &t1, double &t2)
{
    __shared__ bool is_last2;
    if (threadIdx.x == 0) {
        part1[blockIdx.x] = s1; part2[blockIdx.x] = s2;
        __threadfence();
        is_last2 = (atomicAdd(counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last2) return false;
    __threadfence();
    double a = 0.0, c = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) { a += __ldcg(part1 + i); c += __ldcg(part2 + i); }
    t1 = block_sum(a, sm); t2 = block_sum(c, sm);
    if (threadIdx.x == 0) *counter = 0u;
    return true;
}

// The vector kernels work on double2 (n = 2 * n_node is even).
// init: x = 0, r = b (or keep x, r), p = dinv*r ; rz = r.dinv.r, rr = r.r
__global__ void __launch_bounds__(VEC_BLOCK)
k_pcg_init(int64_t n2, const double2 *__restrict__ b, const double2 *__restrict__ dinv, double2 *__restrict__ x,
           double2 *__restrict__ r, double2 *__restrict__ p, double *__restrict__ part_rz, double *__restrict__ part_rr,
           unsigned int *counter, double *__restrict__ sc, int keep_x)
{
    __shared__ double sm[VEC_BLOCK / 32];
    double rz = 0.0, rr = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
        const double2 ri = keep_x ? r[i] : b[i];
        const double2 di = dinv[i];
        const double2 zi = make_double2(di.x * ri.x, di.y * ri.y);
        if (!keep_x) { x[i] = make_double2(0.0, 0.0); r[i] = ri; }
        p[i] = zi;
        rz = fma(ri.x, zi.x, rz); rz = fma(ri.y, zi.y, rz);
        rr = fma(ri.x, ri.x, rr); rr = fma(ri.y, ri.y, rr);
    }
    const double s1 = block_sum(rz, sm);
    const double s2 = block_sum(rr, sm);
    double a, c;
    if (grid_sum2(s1, s2, part_rz, part_rr, counter, sm, a, c) && threadIdx.x == 0) { sc[0] = a; sc[1] = c; }
}

// after the (optional) all-reduce of sc[0..1]: replicate into the other parity, remember ||b||^2
__global__ void k_pcg_init_finish(double *__restrict__ sc, int keep_x)
{
    sc[2] = sc[0]; sc[3] = sc[1];
    sc[SC_DONE] = 0.0; sc[SC_IT_DONE] = 0.0;
    if (!keep_x) sc[SC_BNORM2] = sc[1];
}

// x += alpha p ; r -= alpha q ; rz_new = r.dinv.r, rr = r.r      (alpha = rz / p.q)
template <bool P2P>
__global__ void __launch_bounds__(VEC_BLOCK)
k_pcg_update(int64_t n2, int it, double *__restrict__ sc, const double2 *__restrict__ p, const double2 *__restrict__ q,
             const double2 *__restrict__ dinv, double2 *__restrict__ x, double2 *__restrict__ r,
             double *__restrict__ part_rz, double *__restrict__ part_rr, unsigned int *counter, P2PDev pp, double *out2)
{
    __shared__ double sm[VEC_BLOCK / 32];
    ring::grid_dep_wait();        // PDL: no-ops unless launched with the programmatic-serialization attribute
    ring::grid_dep_launch();
    if (out2 == nullptr || out2 != sc + sc_rz(it + 1)) {   // row-partitioned runs poll rarely and rely on the stop flag
        if (sc[SC_DONE] != 0.0) return;
    }
    double pq, unused;
    if (P2P) p2p_collect(pp, 0, pq, unused);
    else pq = sc[SC_PQ];
    const double rz_old = sc[sc_rz(it)];
    const double alpha = (pq != 0.0) ? rz_old / pq : 0.0;
    double rz = 0.0, rr = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
        const double2 pi = p[i], qi = q[i], di = dinv[i];
        double2 xi = x[i], ri = r[i];
        xi.x = fma(alpha, pi.x, xi.x); xi.y = fma(alpha, pi.y, xi.y);
        ri.x = fma(-alpha, qi.x, ri.x); ri.y = fma(-alpha, qi.y, ri.y);
        x[i] = xi; r[i] = ri;
        const double zx = di.x * ri.x, zy = di.y * ri.y;
        rz = fma(ri.x, zx, rz); rz = fma(ri.y, zy, rz);
        rr = fma(ri.x, ri.x, rr); rr = fma(ri.y, ri.y, rr);
    }
    const double s1 = block_sum(rz, sm);
    const double s2 = block_sum(rr, sm);
    double a, c;
    if (grid_sum2(s1, s2, part_rz, part_rr, counter, sm, a, c)) {
        if (P2P) p2p_publish(pp, 1, a, c);
        else if (threadIdx.x == 0) { out2[0] = a; out2[1] = c; }
    }
}

// p = dinv*r + beta p      (beta = rz_new / rz_old)
template <bool P2P>
__global__ void __launch_bounds__(VEC_BLOCK)
k_pcg_direction(int64_t n2, int it, double *__restrict__ sc, const double2 *__restrict__ r,
                const double2 *__restrict__ dinv, double2 *__restrict__ p, P2PDev pp, double tol2, unsigned int *done_counter)
{
    ring::grid_dep_wait();
    ring::grid_dep_launch();
    if (done_counter && sc[SC_DONE] != 0.0) return;
    double rz_new, rr;
    if (P2P) {
        p2p_collect(pp, 1, rz_new, rr);
        if (blockIdx.x == 0 && threadIdx.x == 0) { sc[sc_rz(it + 1)] = rz_new; sc[sc_rr(it + 1)] = rr; }   // for the host poll
    } else {
        rz_new = sc[sc_rz(it + 1)];
        rr = sc[sc_rr(it + 1)];
    }
    if (done_counter && rr <= tol2) {
        // converged by the recurrence residual: the LAST block to get here raises the flag (the others may still be reading
        // it at their entry), so the remaining kernels of the chunk become no-ops; p is left as it is
        __shared__ bool last;
        __syncthreads();
        if (threadIdx.x == 0) {
            last = (atomicAdd(done_counter, 1u) == gridDim.x - 1);
            if (last) { sc[SC_IT_DONE] = (double)(it + 1); sc[SC_DONE] = 1.0; *done_counter = 0u; }
        }
        return;
    }
    const double rz_old = sc[sc_rz(it)];
    const double beta = (rz_old != 0.0) ? rz_new / rz_old : 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
        const double2 ri = r[i], di = dinv[i];
        double2 pi = p[i];
        pi.x = fma(beta, pi.x, di.x * ri.x); pi.y = fma(beta, pi.y, di.y * ri.y);
        p[i] = pi;
    }
}

// ---- the vector part of an iteration in ONE kernel -----------------------------------------------------------------------
//   phase 1   r -= alpha q ; z = dinv r ; partials of r.z and r.r                 (alpha = rz / p.q)
//   arrive    the CTA's partials are published: arrival counter; peer transport: the last CTA of this rank also writes the
//             rank's sums into every peer's mailbox
//   phase 2   x += alpha p                                                          (needs neither sum: hides the wait)
//   wait      single GPU: until every CTA has arrived (all CTAs are co-resident: the grid is sized by the occupancy query),
//             after which every CTA sums the partials in index order -- the same bits everywhere; peer transport: until the
//             W ranks' sums are in this rank's mailbox, summed in rank order
//   phase 3   p = z + beta p                                                        (beta = rz_new / rz)
// A thread owns the elements tid, tid + T, tid + 2T, ... and walks them in batches with the loads of a batch issued
// together; z of the FIRST batch stays in registers across the wait (all of z at the 1 M-cell bench size: 3 CTAs of 256
// threads per SM own 6.07 elements per thread), the later batches re-read r and dinv.  Every vector is then read once per
// iteration (q r dinv | p x in, r | x | p out, p a second time out of L2: 88 MB at the bench size).
// The iteration that meets the tolerance still updates x; p is left and the stop flag goes up.
constexpr int FUSE_BLOCK = 256, FUSE_EPT = 6, FUSE_BPS = 3;

// first batch of the residual update: D^-1 and r of the thread's first FUSE_EPT elements.  Neither is written by the SpMV, so the
// loads are issued BEFORE the grid dependency / the mailbox wait and overlap the tail of the SpMV.
__device__ __forceinline__ void vec_preload(int64_t n2, int64_t tid, int64_t nth, const double2 *__restrict__ dinv,
                                            const double2 *__restrict__ r, double2 (&dpre)[FUSE_EPT], double2 (&rpre)[FUSE_EPT])
{
#pragma unroll
    for (int e = 0; e < FUSE_EPT; e++) {
        const int64_t i = tid + e * nth;
        if (i < n2) { dpre[e] = dinv[i]; rpre[e] = r[i]; }
    }
}

__device__ __forceinline__ void vec_phase1(int64_t n2, int64_t tid, int64_t nth, double alpha, const double2 *__restrict__ q,
                                           const double2 *__restrict__ dinv, double2 *__restrict__ r, double2 (&zc)[FUSE_EPT],
                                           const double2 (&dpre)[FUSE_EPT], const double2 (&rpre)[FUSE_EPT], double &rz, double &rr)
{
    auto residual = [&](int64_t i, double2 qi, double2 di, double2 ri) {
        ri.x = fma(-alpha, qi.x, ri.x); ri.y = fma(-alpha, qi.y, ri.y);
        r[i] = ri;
        const double2 z = make_double2(di.x * ri.x, di.y * ri.y);
        rz = fma(ri.x, z.x, rz); rz = fma(ri.y, z.y, rz);
        rr = fma(ri.x, ri.x, rr); rr = fma(ri.y, ri.y, rr);
        return z;
    };
#pragma unroll
    for (int e = 0; e < FUSE_EPT; e++) {       // first batch: z stays in registers (the compiler interleaves the loads)
        const int64_t i = tid + e * nth;
        if (i < n2) zc[e] = residual(i, q[i], dpre[e], rpre[e]);
    }
    constexpr int B1 = 2;                      // later batches (large problems, DRAM-bound): loads of a batch issued together
    for (int64_t base = tid + FUSE_EPT * nth; base < n2; base += B1 * nth) {
        double2 qv[B1], dv[B1], rv[B1];
#pragma unroll
        for (int e = 0; e < B1; e++) {
            const int64_t i = base + e * nth;
            if (i < n2) { qv[e] = q[i]; dv[e] = dinv[i]; rv[e] = r[i]; }
        }
#pragma unroll
        for (int e = 0; e < B1; e++) {
            const int64_t i = base + e * nth;
            if (i < n2) residual(i, qv[e], dv[e], rv[e]);
        }
    }
}

// x += alpha p over the thread's elements
__device__ __forceinline__ void vec_phase2(int64_t n2, int64_t tid, int64_t nth, double alpha, const double2 *__restrict__ p,
                                           double2 *__restrict__ x)
{
    constexpr int B2 = 3;
    for (int64_t base = tid; base < n2; base += B2 * nth) {
        double2 pv[B2], xv[B2];
#pragma unroll
        for (int e = 0; e < B2; e++) {
            const int64_t i = base + e * nth;
            if (i < n2) { pv[e] = p[i]; xv[e] = x[i]; }
        }
#pragma unroll
        for (int e = 0; e < B2; e++) {
            const int64_t i = base + e * nth;
            if (i < n2) x[i] = make_double2(fma(alpha, pv[e].x, xv[e].x), fma(alpha, pv[e].y, xv[e].y));
        }
    }
}

// p = z + beta p over the thread's elements (z of the first batch from registers)
__device__ __forceinline__ void vec_phase3(int64_t n2, int64_t tid, int64_t nth, double beta, const double2 *__restrict__ dinv,
                                           const double2 *__restrict__ r, double2 *__restrict__ p, const double2 (&zc)[FUSE_EPT])
{
    {
        double2 pv[FUSE_EPT];
#pragma unroll
        for (int e = 0; e < FUSE_EPT; e++) {
            const int64_t i = tid + e * nth;
            if (i < n2) pv[e] = p[i];
        }
#pragma unroll
        for (int e = 0; e < FUSE_EPT; e++) {
            const int64_t i = tid + e * nth;
            if (i < n2) p[i] = make_double2(fma(beta, pv[e].x, zc[e].x), fma(beta, pv[e].y, zc[e].y));
        }
    }
    constexpr int B3 = 2;   // later batches (large problems, DRAM-bound): z recomputed from r and dinv
    for (int64_t base = tid + FUSE_EPT * nth; base < n2; base += B3 * nth) {
        double2 pv[B3], rv[B3], dv[B3];
#pragma unroll
        for (int e = 0; e < B3; e++) {
            const int64_t i = base + e * nth;
            if (i < n2) { pv[e] = p[i]; rv[e] = r[i]; dv[e] = dinv[i]; }
        }
#pragma unroll
        for (int e = 0; e < B3; e++) {
            const int64_t i = base + e * nth;
            if (i < n2) p[i] = make_double2(fma(beta, pv[e].x, dv[e].x * rv[e].x), fma(beta, pv[e].y, dv[e].y * rv[e].y));
        }
    }
}

// n_pq_part > 0 (single GPU, ring SpMV): p.q = sum of the SpMV's per-CTA partials part_pq[0 .. n_pq_part), added up here by
// every CTA in the same order; otherwise sc[SC_PQ] (row kernel) or the peers' mailboxes (P2P).
// `barrier` counts arrivals monotonically (zeroed by the host); `target` = arrivals once every CTA of THIS launch is in.
template <bool P2P>
__global__ void __launch_bounds__(FUSE_BLOCK, FUSE_BPS)
k_pcg_vec(int64_t n2, int it, double *__restrict__ sc, double2 *__restrict__ p, const double2 *__restrict__ q,
          const double2 *__restrict__ dinv, double2 *__restrict__ x, double2 *__restrict__ r, const double *__restrict__ part_pq,
          int n_pq_part, double *__restrict__ part_rz, double *__restrict__ part_rr, unsigned int *barrier, unsigned int target,
          P2PDev pp, double tol2, int *err)
{
    __shared__ double sm[FUSE_BLOCK / 32];
    __shared__ bool is_last;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    double2 dpre[FUSE_EPT], rpre[FUSE_EPT];
    vec_preload(n2, tid, nth, dinv, r, dpre, rpre);     // r and D^-1 are not the SpMV's to write: fetched ahead of the dependency
    ring::grid_dep_wait();        // PDL: no-op unless launched with the programmatic-serialization attribute
    if (__ldcg(sc + SC_DONE) != 0.0 || *reinterpret_cast<volatile int *>(err)) return;   // stopped, or an earlier wait timed out
    double pq, unused;
    if (P2P) p2p_collect(pp, 0, pq, unused);
    else if (n_pq_part > 0) {
        double a = 0.0;
        for (int j = threadIdx.x; j < n_pq_part; j += blockDim.x) a += __ldcg(part_pq + j);
        pq = block_sum(a, sm);
    } else pq = __ldcg(sc + SC_PQ);
    const double rz_old = __ldcg(sc + sc_rz(it));
    const double alpha = (pq != 0.0) ? rz_old / pq : 0.0;
    double2 zc[FUSE_EPT];
    double rz = 0.0, rr = 0.0;
    vec_phase1(n2, tid, nth, alpha, q, dinv, r, zc, dpre, rpre, rz, rr);
    const double s1 = block_sum(rz, sm);
    const double s2 = block_sum(rr, sm);
    // arrive
    if (threadIdx.x == 0) {
        part_rz[blockIdx.x] = s1; part_rr[blockIdx.x] = s2;
        __threadfence();
        is_last = (atomicAdd(barrier, 1u) + 1u == target);
    }
    if (P2P) {
        __syncthreads();
        if (is_last) {            // this rank's sums -> every peer's mailbox
            __threadfence();
            double a = 0.0, c = 0.0;
            for (int j = threadIdx.x; j < (int)gridDim.x; j += blockDim.x) { a += __ldcg(part_rz + j); c += __ldcg(part_rr + j); }
            const double t1 = block_sum(a, sm), t2 = block_sum(c, sm);
            p2p_publish(pp, 1, t1, t2);
        }
    }
    vec_phase2(n2, tid, nth, alpha, p, x);
    // wait
    double rz_new, rr_new;
    if (P2P) p2p_collect(pp, 1, rz_new, rr_new);
    else {
        if (threadIdx.x == 0) {
            long long t0 = 0;
            for (unsigned spins = 0; *reinterpret_cast<volatile unsigned int *>(barrier) < target; spins++) {
                if ((spins & 0xffff) == 0xffff) {      // a CTA that never became resident: give up instead of hanging
                    const long long now = clock64();
                    if (t0 == 0) t0 = now;
                    else if (now - t0 > P2P_TIMEOUT_CLOCKS) { *err = 1; break; }
                }
            }
            __threadfence();
        }
        __syncthreads();
        double a = 0.0, c = 0.0;
        for (int j = threadIdx.x; j < (int)gridDim.x; j += blockDim.x) { a += __ldcg(part_rz + j); c += __ldcg(part_rr + j); }
        rz_new = block_sum(a, sm); rr_new = block_sum(c, sm);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { sc[sc_rz(it + 1)] = rz_new; sc[sc_rr(it + 1)] = rr_new; }
    if (rr_new <= tol2) {
        // converged by the recurrence residual: every CTA sees the same two numbers and leaves p alone; the flag turns the rest
        // of the enqueued chunk into no-ops
        if (blockIdx.x == 0 && threadIdx.x == 0) { sc[SC_IT_DONE] = (double)(it + 1); __threadfence(); sc[SC_DONE] = 1.0; }
        return;
    }
    ring::grid_dep_launch();      // the next SpMV may become resident from here on (its producer warp pre-fills the ring)
    const double beta = (rz_old != 0.0) ? rz_new / rz_old : 0.0;
    vec_phase3(n2, tid, nth, beta, dinv, r, p, zc);
}

// r = b - q (true residual), rr_true = r.r
__global__ void __launch_bounds__(VEC_BLOCK)
k_true_residual(int64_t n, const double *__restrict__ b, const double *__restrict__ q, double *__restrict__ r,
                double *__restrict__ part_rr, unsigned int *counter, double *out)
{
    __shared__ double sm[VEC_BLOCK / 32];
    double rr = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double ri = b[i] - q[i];
        r[i] = ri;
        rr = fma(ri, ri, rr);
    }
    const double s = block_sum(rr, sm);
    double total;
    if (grid_sum(s, part_rr, counter, sm, total) && threadIdx.x == 0) *out = total;
}

__global__ void k_add_vec(int64_t n, const double *__restrict__ a, const double *__restrict__ b, double *__restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] + b[i];
}

// ---------------------------------------------------------------------------------------------
// row-partitioned operation (one process per GPU): halo exchange of ghost entries + all-reduce of the dot products
// ---------------------------------------------------------------------------------------------
__global__ void k_pack_halo(int64_t n_send, const int32_t *__restrict__ send_idx, const double2 *__restrict__ v,
                            double2 *__restrict__ buf)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_send) buf[i] = v[send_idx[i]];
}

#define NCCL_OK(call) do { ncclResult_t r__ = (call); if (r__ != ncclSuccess) return cudaErrorUnknown; } while (0)

// ghost entries (2 doubles per ghost node, grouped by owner) <- their owners' values of `v`, received into `ghost`
static cudaError_t halo_exchange(const DistPlan *d, const double *v, double *ghost, cudaStream_t s, int64_t *launches)
{
    if (!d || d->nb_rank.empty()) return cudaSuccess;
    const int64_t n_send = d->send_ptr.back();
    if (n_send) {
        k_pack_halo<<<(unsigned)((n_send + 255) / 256), 256, 0, s>>>(n_send, d->send_idx, reinterpret_cast<const double2 *>(v),
                                                                      reinterpret_cast<double2 *>(d->sendbuf));
        if (launches) ++*launches;
    }
    const NcclApi *nc = nccl_api();
    if (!nc) return cudaErrorUnknown;
    NCCL_OK(nc->GroupStart());
    for (size_t k = 0; k < d->nb_rank.size(); k++) {
        const int64_t ns = d->send_ptr[k + 1] - d->send_ptr[k], nr = d->recv_ptr[k + 1] - d->recv_ptr[k];
        if (ns) NCCL_OK(nc->Send(d->sendbuf + 2 * d->send_ptr[k], (size_t)(2 * ns), ncclDouble, d->nb_rank[k], d->comm, s));
        if (nr) NCCL_OK(nc->Recv(ghost + 2 * d->recv_ptr[k], (size_t)(2 * nr), ncclDouble, d->nb_rank[k], d->comm, s));
    }
    NCCL_OK(nc->GroupEnd());
    return cudaGetLastError();
}

static cudaError_t all_reduce(const DistPlan *d, double *v, int count, cudaStream_t s, const double *src = nullptr)
{
    if (!d || d->world <= 1) return cudaSuccess;
    const NcclApi *nc = nccl_api();
    if (!nc) return cudaErrorUnknown;
    NCCL_OK(nc->AllReduce(src ? src : v, v, (size_t)count, ncclDouble, ncclSum, d->comm, s));
    return cudaSuccess;
}

// ---------------------------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------------------------
static int vec_grid(int64_t n2, int64_t cap = 148 * 8)
{
    int64_t g = (n2 + VEC_BLOCK - 1) / VEC_BLOCK;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

static int spmv_grid(int64_t n_node)
{
    constexpr int64_t rows_per_block = SPMV_BLOCK / SPMV_LANES;
    int64_t g = (n_node + rows_per_block - 1) / rows_per_block;
    if (g > SPMV_GRID_CAP) g = SPMV_GRID_CAP;
    if (g < 1) g = 1;
    return (int)g;
}

static const P2PDev NO_P2P{nullptr, nullptr, 1, 0, 0.0, 0, nullptr, 0u};

// plain launch: ghosts (if any) are the tail of x
template <bool DOT>
static void launch_spmv(const DevSystem &A, const double *x, double *y, const uint8_t *fixed, double *part, unsigned int *counter,
                        double *dot_out, cudaStream_t s, const double *done = nullptr)
{
    const int64_t n_node = A.n_own / 2;
    k_spmv<DOT, false><<<spmv_grid(n_node), SPMV_BLOCK, 0, s>>>(n_node, A.node_ptr, A.node_idx, A.data, x, y, fixed, part, counter,
                                                                 dot_out, NO_P2P, 0, nullptr, nullptr, nullptr, 0, HaloPush{}, done);
}

// y = K x on the rows this rank owns (all rows on a single GPU); x must hold current ghost values
cudaError_t spmv(const DevSystem &A, const double *x, double *y, const uint8_t *fixed, cudaStream_t s)
{
    if (A.n_own == 0) return cudaSuccess;
    launch_spmv<false>(A, x, y, fixed, nullptr, nullptr, nullptr, s);
    return cudaGetLastError();
}

cudaError_t pcg_setup(const DevSystem &A, const double *f, const uint8_t *fixed, const double *fixed_val,
                      PcgBuffers &w, DistPlan *dist, cudaStream_t s, int64_t *launches)
{
    const int64_t n = A.n, n_own = A.n_own, n_node = n_own / 2;
    // fixed / fixed_val cover the ghost nodes too, so u_fix needs no exchange
    cudaError_t e = cudaMemsetAsync(w.counters, 0, sizeof(unsigned int) * 8, s);
    if (e != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(w.x, 0, sizeof(double) * n, s)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(w.p, 0, sizeof(double) * n, s)) != cudaSuccess) return e;
    k_build_ufix<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, fixed, fixed_val, w.ufix);
    e = spmv(A, w.ufix, w.q, nullptr, s);
    if (e != cudaSuccess) return e;
    if (n_own) {
        k_build_rhs<<<(unsigned)((n_own + 255) / 256), 256, 0, s>>>(n_own, fixed, f, w.q, w.b);
        k_diag_inv<<<(unsigned)((n_node + 255) / 256), 256, 0, s>>>(n_node, A.node_ptr, A.node_idx, A.data, fixed, w.dinv, w.counters + 7);
    }
    if (launches) *launches += 4;
    const bool multi = dist && dist->world > 1, peer = multi && dist->p2p;
    // sizes the host needs: largest row degree, sub-blocks in the owned rows, and (peer transport) which rows gather ghosts
    w.n_tile = 0; w.max_deg = 0;
    if (dist) dist->n_halo_rows = 0;
    if (!n_node) return cudaGetLastError();
    int64_t blk_own = 0;
    std::vector<uint8_t> mark;
    if ((e = cudaMemcpyAsync(&w.max_deg, w.counters + 7, sizeof(unsigned int), cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
    if ((e = cudaMemcpyAsync(&blk_own, A.node_ptr + n_node, sizeof(int64_t), cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
    if (peer) {
        k_mark_halo_rows<<<(unsigned)((n_node + 255) / 256), 256, 0, s>>>(n_node, A.node_ptr, A.node_idx, w.row_halo);
        if (launches) ++*launches;
        mark.resize((size_t)n_node);
        if ((e = cudaMemcpyAsync(mark.data(), w.row_halo, (size_t)n_node, cudaMemcpyDeviceToHost, s)) != cudaSuccess) return e;
    }
    if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;
    if (peer) {   // ascending list of the rows that reference a ghost column (a few thousand)
        std::vector<int32_t> rows;
        for (int64_t a = 0; a < n_node; a++) if (mark[a]) rows.push_back((int32_t)a);
        if ((int64_t)rows.size() > dist->halo_rows_cap) {
            if (dist->halo_rows) cudaFree(dist->halo_rows);
            dist->halo_rows = nullptr;
            dist->halo_rows_cap = (int64_t)rows.size() + 1024;
            if ((e = cudaMalloc(&dist->halo_rows, sizeof(int32_t) * dist->halo_rows_cap)) != cudaSuccess) { dist->halo_rows_cap = 0; return e; }
        }
        dist->n_halo_rows = (int64_t)rows.size();
        if (!rows.empty()) {
            if ((e = cudaMemcpyAsync(dist->halo_rows, rows.data(), sizeof(int32_t) * rows.size(), cudaMemcpyHostToDevice, s)) != cudaSuccess) return e;
            if ((e = cudaStreamSynchronize(s)) != cudaSuccess) return e;   // `rows` goes out of scope
        }
    }
    // tile table of the TMA-ring SpMV (the NCCL transport keeps the row kernel); peer transport: tiles with halo rows flagged
    if ((!multi || peer) && w.ring_tiles && w.max_deg <= (unsigned int)PcgRing::MAXDEG) {
        const int64_t nt = (n_node + blk_own + PcgRing::SPAN - 1) / PcgRing::SPAN;
        if (nt <= w.ring_cap && nt < (int64_t)1 << 30) {
            w.n_tile = (int)nt;
            k_ring_tiles<<<(unsigned)((nt + 127) / 128), 128, 0, s>>>(n_node, A.node_ptr, w.n_tile, reinterpret_cast<RingTile *>(w.ring_tiles),
                                                                     peer ? w.row_halo : nullptr);
            if (launches) ++*launches;
        }
    }
    return cudaGetLastError();
}

static cudaError_t true_residual(const DevSystem &A, const uint8_t *fixed, PcgBuffers &w, const DistPlan *dist,
                                 cudaStream_t s, double *rr_true, int64_t *launches)
{
    const int64_t n_own = A.n_own;
    cudaError_t e = halo_exchange(dist, w.x, w.x + n_own, s, launches);
    if (e != cudaSuccess) return e;
    launch_spmv<false>(A, w.x, w.q, fixed, nullptr, nullptr, nullptr, s);
    k_true_residual<<<vec_grid(n_own), VEC_BLOCK, 0, s>>>(n_own, w.b, w.q, w.r, w.part_b, w.counters + 2, w.sc + SC_RR_TRUE);
    if (launches) *launches += 2;
    if ((e = all_reduce(dist, w.sc + SC_RR_TRUE, 1, s)) != cudaSuccess) return e;
    e = cudaMemcpyAsync(rr_true, w.sc + SC_RR_TRUE, sizeof(double), cudaMemcpyDeviceToHost, s);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(s);
}

// Runs PCG; returns via info.  `chunk` iterations are enqueued between host polls (the device stops on its own).
// prof (optional): 2 pools of 4*chunk events; prof_ms[0..2] accumulate the device time of the SpMV (incl. halo exchange and the
// p.q all-reduce when row-partitioned) / the vector kernel (NCCL transport: update / direction kernels) of every
// prof_every-th iteration, prof_ms[3] the iterations profiled.  The events of a chunk are read back while the next chunk
// runs, so profiling does not stall the GPU.
//
// Row-partitioned, two transports:
//   NCCL  -- per iteration: pack + ncclSend/Recv of the ghost entries of p, ncclAllReduce(p.q), ncclAllReduce(rz, rr)
//   peer  -- (dist->p2p) no NCCL call inside the loop: CTA 0 of the SpMV writes this rank's boundary entries of p into the
//            neighbours' ghost segments over NVLink, the rows that gather ghosts wait for the neighbours' stamps; the last
//            CTA of the SpMV / of the vector kernel's first phase writes its sums into every peer's mailbox and the consumer
//            adds the W contributions in rank order.  The first iteration after (re)initialisation still uses NCCL for the halo.
cudaError_t pcg_run(const DevSystem &A, const uint8_t *fixed, PcgBuffers &w, DistPlan *dist, double rtol, int maxit,
                    int chunk, mofem_solve_info_t *info, cudaStream_t s, int64_t *launches, cudaEvent_t *prof, double *prof_ms, int prof_every)
{
    if (prof_every < 1) prof_every = 1;
    const int64_t n_own = A.n_own, n_node = n_own / 2, n2 = n_own / 2;
    const int gv = vec_grid(n2), gs = spmv_grid(n_node);
    cudaError_t e;
    double sc_host[SC_COUNT];
    const double2 *b2 = reinterpret_cast<const double2 *>(w.b), *dinv2 = reinterpret_cast<const double2 *>(w.dinv);
    double2 *x2 = reinterpret_cast<double2 *>(w.x), *r2 = reinterpret_cast<double2 *>(w.r);
    double2 *p2 = reinterpret_cast<double2 *>(w.p), *q2 = reinterpret_cast<double2 *>(w.q);
    const bool multi = dist && dist->world > 1, peer = multi && dist->p2p, single = !multi;
    if (peer && w.p != dist->xbuf + p2p_p_off(dist->world)) return cudaErrorInvalidValue;   // the slab lives in the exchange buffer
    const int W = dist ? dist->world : 1;
    const double stamp0 = peer ? (double)(++dist->epoch) * 4294967296.0 : 0.0;
    int *err_dev = reinterpret_cast<int *>(w.counters + 6);
    auto ppdev = [&](int it_for) {
        P2PDev pp = NO_P2P;
        if (peer) {
            pp.peer = dist->peer; pp.local = dist->xbuf; pp.world = W; pp.rank = dist->rank;
            pp.stamp = stamp0 + (double)it_for + 1.0; pp.parity = it_for & 1; pp.err = err_dev;
            pp.stamp32 = (unsigned int)(((dist->epoch & 0xfff) << 20) | (((int64_t)it_for + 1) & 0xfffff));
        }
        return pp;
    };
    auto init = [&](int keep_x) -> cudaError_t {
        k_pcg_init<<<gv, VEC_BLOCK, 0, s>>>(n2, b2, dinv2, x2, r2, p2, w.part_a, w.part_b, w.counters + 1, w.sc, keep_x);
        cudaError_t ee = all_reduce(dist, w.sc, 2, s);
        if (ee != cudaSuccess) return ee;
        k_pcg_init_finish<<<1, 1, 0, s>>>(w.sc, keep_x);
        if (launches) *launches += 2;
        return cudaGetLastError();
    };
    if ((e = init(0)) != cudaSuccess) return e;
    e = cudaMemcpyAsync(sc_host, w.sc, sizeof(double) * SC_COUNT, cudaMemcpyDeviceToHost, s);
    if (e != cudaSuccess) return e;
    e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) return e;
    const double bnorm2 = sc_host[SC_BNORM2];
    info->iterations = 0; info->converged = 0; info->restarts = 0; info->relres_rec = 1.0; info->relres_true = 1.0;
    if (bnorm2 == 0.0) {  // zero right-hand side: x = 0
        info->converged = 1; info->relres_rec = 0.0; info->relres_true = 0.0;
        return cudaSuccess;
    }
    const double tol2 = rtol * rtol * bnorm2;
    int it = 0;       // global iteration counter (parity selects the (rz, rr) pair)
    double prev_rr_true = 0.0;
    bool fresh = true;   // p was just (re)initialised: its halo has not been pushed by a previous iteration
    int pend_pool = -1, pend_n = 0, pool = 0;
    auto harvest = [&]() {
        if (!prof || pend_pool < 0) return;
        cudaEvent_t *ev = prof + (size_t)pend_pool * 4 * chunk;
        for (int k = 0; k < pend_n; k++)
            for (int j = 0; j < 3; j++) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, ev[4 * k + j], ev[4 * k + j + 1]);
                prof_ms[j] += ms;
            }
        prof_ms[3] += pend_n;
        pend_pool = -1;
    };
    auto done = [&](cudaError_t ee) { harvest(); return ee; };
    const int n_nb = dist ? (int)dist->nb_rank.size() : 0;
    const int64_t n_send = dist && !dist->send_ptr.empty() ? dist->send_ptr.back() : 0;
    // TMA-ring SpMV if every node row fits a ring stage (tile table built by pcg_setup)
    const bool ring_on = (single || peer) && w.use_ring && w.n_tile > 0;
    if (ring_on) {
        // per device (a process may hold contexts on several GPUs), so set on every solve: microseconds
        if ((e = cudaFuncSetAttribute(k_spmv_ring, cudaFuncAttributeMaxDynamicSharedMemorySize, PcgRing::SMEM_BYTES)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(k_spmv_ring_p2p, cudaFuncAttributeMaxDynamicSharedMemorySize, PcgRing::SMEM_BYTES)) != cudaSuccess) return e;
    }
    // the vector kernel waits inside the kernel for all its CTAs (grid barrier) or for the peers: every CTA must be resident
    int g_vec = 0;
    if (single || peer) {
        int bps = 0;
        const cudaError_t oe = single ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_pcg_vec<false>, FUSE_BLOCK, 0)
                                      : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_pcg_vec<true>, FUSE_BLOCK, 0);
        if (oe != cudaSuccess || bps < 1) return oe != cudaSuccess ? oe : cudaErrorLaunchOutOfResources;
        int dev = 0, sms = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int64_t want = (n2 + FUSE_BLOCK - 1) / FUSE_BLOCK, cap = (int64_t)sms * std::min(bps, FUSE_BPS);
        g_vec = (int)std::max<int64_t>(1, std::min<int64_t>(std::min(want, cap), PCG_MAX_PART));
        if ((e = cudaMemsetAsync(w.counters + 4, 0, sizeof(unsigned int), s)) != cudaSuccess) return e;
    }
    unsigned int vec_k = 0;    // vector-kernel launches since the arrival counter was last zeroed
    // programmatic dependent launch between the kernels of the loop; measured: costs time with the row kernel
    // (181 vs 133 us per iteration), and a run that times every iteration wants the kernels back to back
    const bool pdl = (single || peer) && w.use_pdl && ring_on && !(prof && prof_every == 1);
    const int g_ring = w.n_tile < RING_GRID ? w.n_tile : RING_GRID;
    const RingTile *tiles = reinterpret_cast<const RingTile *>(w.ring_tiles);
    const double *done_flag = w.sc + SC_DONE;
    while (it < maxit) {
        int todo = chunk;
        if (it + todo > maxit) todo = maxit - it;
        cudaEvent_t *ev_pool = prof ? prof + (size_t)pool * 4 * chunk : nullptr;
        int n_sampled = 0;
        for (int kk = 0; kk < todo; kk++, it++) {
            // profile: events around the kernels of every prof_every-th iteration (the others keep their PDL overlap)
            cudaEvent_t *ev = (ev_pool && it % prof_every == 0) ? ev_pool : nullptr;
            const int k = n_sampled;
            if (ev) { n_sampled++; cudaEventRecord(ev[4 * k], s); }
            const P2PDev pp = ppdev(it);
            if (single || peer) {
                if (single) {
                    if (ring_on)
                        e = launch_ex(pdl, k_spmv_ring, g_ring, PcgRing::THREADS, PcgRing::SMEM_BYTES, s, w.n_tile, tiles, A.node_ptr, A.node_idx,
                                      A.data, (const double *)w.p, w.q, fixed, w.part_c, done_flag);
                    else
                        e = launch_ex(false, k_spmv<true, false>, gs, SPMV_BLOCK, 0, s, n_node, A.node_ptr, A.node_idx, A.data, (const double *)w.p,
                                      w.q, fixed, w.part_c, w.counters, w.sc + SC_PQ, NO_P2P, 0, (const int *)nullptr, (const uint8_t *)nullptr,
                                      (const int32_t *)nullptr, (int64_t)0, HaloPush{}, done_flag);
                } else {
                    HaloPush hp{};
                    if (fresh) {   // halo of a freshly initialised p: NCCL, nothing to push or to wait for
                        if ((e = halo_exchange(dist, w.p, w.p + n_own, s, launches)) != cudaSuccess) return done(e);
                    } else {
                        hp.n_send = n_send; hp.send_idx = dist->send_idx; hp.send_nb = dist->send_nb; hp.send_dst = dist->send_dst;
                        hp.nb_ghost_off = dist->nb_ghost_off; hp.nb_rank = dist->nb_rank_dev; hp.n_nb = n_nb;
                    }
                    const int n_wait = fresh ? 0 : dist->n_wait;
                    if (ring_on)
                        e = launch_ex(pdl && !fresh, k_spmv_ring_p2p, g_ring, PcgRing::THREADS, PcgRing::SMEM_BYTES, s, w.n_tile, tiles, n_node,
                                      A.node_ptr, A.node_idx, A.data, (const double *)w.p, w.q, fixed, w.part_c, w.counters, pp, n_wait,
                                      (const int *)dist->wait_rank, (const uint8_t *)w.row_halo, (const int32_t *)dist->halo_rows,
                                      dist->n_halo_rows, hp, done_flag);
                    else
                        e = launch_ex(false, k_spmv<true, true>, gs, SPMV_BLOCK, 0, s, n_node, A.node_ptr, A.node_idx, A.data, (const double *)w.p,
                                      w.q, fixed, w.part_c, w.counters, (double *)nullptr, pp, n_wait, (const int *)dist->wait_rank,
                                      (const uint8_t *)w.row_halo, (const int32_t *)dist->halo_rows, dist->n_halo_rows, hp, done_flag);
                    fresh = false;
                }
                if (e != cudaSuccess) return done(e);
                if (ev) cudaEventRecord(ev[4 * k + 1], s);
                if ((uint64_t)(vec_k + 1) * (uint64_t)g_vec > 0x7fffffffull) {   // keep the arrival counter inside 32 bits
                    if ((e = cudaMemsetAsync(w.counters + 4, 0, sizeof(unsigned int), s)) != cudaSuccess) return done(e);
                    vec_k = 0;
                }
                ++vec_k;
                const unsigned int target = vec_k * (unsigned int)g_vec;
                if (single)
                    e = launch_ex(pdl, k_pcg_vec<false>, g_vec, FUSE_BLOCK, 0, s, n2, it, w.sc, p2, (const double2 *)q2, dinv2, x2, r2,
                                  (const double *)w.part_c, ring_on ? g_ring : 0, w.part_a, w.part_b, w.counters + 4, target, NO_P2P, tol2, err_dev);
                else
                    e = launch_ex(pdl, k_pcg_vec<true>, g_vec, FUSE_BLOCK, 0, s, n2, it, w.sc, p2, (const double2 *)q2, dinv2, x2, r2,
                                  (const double *)w.part_c, 0, w.part_a, w.part_b, w.counters + 4, target, pp, tol2, err_dev);
                if (e != cudaSuccess) return done(e);
                if (ev) { cudaEventRecord(ev[4 * k + 2], s); cudaEventRecord(ev[4 * k + 3], s); }
            } else {
                // NCCL transport: partials are all-reduced INTO the global slots, so re-reducing after the device has stopped is idempotent
                if ((e = halo_exchange(dist, w.p, w.p + n_own, s, launches)) != cudaSuccess) return done(e);
                launch_spmv<true>(A, w.p, w.q, fixed, w.part_c, w.counters, w.sc + SC_PART_PQ, s, done_flag);
                if ((e = all_reduce(dist, w.sc + SC_PQ, 1, s, w.sc + SC_PART_PQ)) != cudaSuccess) return done(e);
                if (ev) cudaEventRecord(ev[4 * k + 1], s);
                k_pcg_update<false><<<gv, VEC_BLOCK, 0, s>>>(n2, it, w.sc, p2, q2, dinv2, x2, r2, w.part_a, w.part_b,
                                                            w.counters + 1, NO_P2P, w.sc + SC_PART_RZ);
                if ((e = all_reduce(dist, w.sc + sc_rz(it + 1), 2, s, w.sc + SC_PART_RZ)) != cudaSuccess) return done(e);
                if (ev) cudaEventRecord(ev[4 * k + 2], s);
                k_pcg_direction<false><<<gv, VEC_BLOCK, 0, s>>>(n2, it, w.sc, r2, dinv2, p2, NO_P2P, tol2, w.counters + 5);
                if (ev) cudaEventRecord(ev[4 * k + 3], s);
            }
        }
        if (launches) *launches += ((single || peer) ? 2 : 3) * (int64_t)todo;
        harvest();  // events of the previous chunk, while this one runs
        e = cudaMemcpyAsync(sc_host, w.sc, sizeof(double) * SC_COUNT, cudaMemcpyDeviceToHost, s);
        if (e != cudaSuccess) return done(e);
        int err_host = 0;
        if (single || peer) e = cudaMemcpyAsync(&err_host, err_dev, sizeof(int), cudaMemcpyDeviceToHost, s);
        if (e != cudaSuccess) return done(e);
        e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) return done(e);
        if (err_host) return done(cudaErrorLaunchTimeout);   // a peer (or a CTA of the grid barrier) did not answer within P2P_TIMEOUT_CLOCKS (~20 s)
        if (prof) { pend_pool = pool; pend_n = n_sampled; pool ^= 1; }
        if (sc_host[SC_DONE] != 0.0) it = (int)sc_host[SC_IT_DONE];   // the device stopped early inside the chunk
        info->iterations = it;
        const double rr = sc_host[sc_rr(it)];
        info->relres_rec = sqrt(rr / bnorm2);
        if (!(rr == rr)) break;  // NaN: breakdown
        if (rr <= tol2) {
            // recurrence residual says converged: verify with the true residual r = b - A x
            double rr_true;
            e = true_residual(A, fixed, w, dist, s, &rr_true, launches);
            if (e != cudaSuccess) return done(e);
            info->relres_true = sqrt(rr_true / bnorm2);
            if (rr_true <= tol2) { info->converged = 1; return done(cudaSuccess); }
            // The recurrence residual drifted from b - A x.  Restart from the true residual (x kept, direction
            // reset).  If a restart cycle does not bring the true residual down any more, fp64 has reached its
            // attainable accuracy for this matrix (||A|| ||x|| eps / ||b||; AMOREBracket: ~1e-9, the reference's
            // own spsolve leaves 6e-11) -- stop; the result counts as converged (= 2) only if the true residual is
            // below max(rtol, floor_rtol), otherwise the caller gets MOFEM_E_NOCONV with the residual in info.
            if (info->restarts > 0 && rr_true > 0.25 * prev_rr_true) {
                const double lim = std::max(rtol, w.floor_rtol);
                info->converged = (rr_true <= lim * lim * bnorm2) ? 2 : 0;
                return done(cudaSuccess);
            }
            prev_rr_true = rr_true;
            info->restarts++;
            if (info->restarts > 20) break;
            if ((e = init(1)) != cudaSuccess) return done(e);  // publishes rz in both parity slots, lowers the stop flag
            fresh = true;
            // the launches behind the one that raised the flag left without arriving: realign the arrival counter
            if (g_vec) {
                if ((e = cudaMemsetAsync(w.counters + 4, 0, sizeof(unsigned int), s)) != cudaSuccess) return done(e);
                vec_k = 0;
            }
        }
    }
    // not converged by the recurrence test: report the true residual anyway
    double rr_true;
    e = true_residual(A, fixed, w, dist, s, &rr_true, launches);
    if (e != cudaSuccess) return done(e);
    info->relres_true = sqrt(rr_true / bnorm2);
    info->converged = rr_true <= tol2 ? 1 : 0;
    return done(cudaGetLastError());
}

// u = x + u_fix over all local DOFs (ghosts included)
cudaError_t pcg_finish(const DevSystem &A, PcgBuffers &w, const DistPlan *dist, double *u_dev, cudaStream_t s)
{
    cudaError_t e = halo_exchange(dist, w.x, w.x + A.n_own, s, nullptr);
    if (e != cudaSuccess) return e;
    k_add_vec<<<(unsigned)((A.n + 255) / 256), 256, 0, s>>>(A.n, w.x, w.ufix, u_dev);
    return cudaGetLastError();
}

// u^T K u  (the caller halves it); u must hold current ghost values
cudaError_t energy_of(const DevSystem &A, const double *u, PcgBuffers &w, const DistPlan *dist, double *out_dev, cudaStream_t s)
{
    launch_spmv<true>(A, u, w.q, nullptr, w.part_c, w.counters + 3, out_dev, s);
    cudaError_t e = all_reduce(dist, out_dev, 1, s);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

}  // namespace mofem
