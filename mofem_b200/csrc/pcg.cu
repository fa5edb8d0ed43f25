// pcg.cu -- Jacobi-preconditioned CG on the assembled matrix (replaces scipy spsolve,
// AMORE.py:264,683,1106; traditionalElement.py:146,296,461).
//
// The Dirichlet elimination of the reference (delete fixed rows/columns, AMORE.py:213-255) is applied as a mask:
// search directions are zero on fixed DOFs and the SpMV zeroes fixed rows, so the iteration is exactly CG on the
// reduced free-free system without building a second matrix.
//
// One iteration = TWO kernels and ONE reduction point (on any number of GPUs):
//   SpMV     q = K p, and for the rows it produces the three sums  p.q,  q.z,  q.D^-1.q   (z = D^-1 r).  Its last CTA adds
//            the two sums  r.z, r.r  of the CURRENT residual (left behind by the previous vector kernel) and hands the five
//            numbers on: into sc[] on one GPU, into every peer's mailbox over NVLink (peer transport), or to one
//            ncclAllReduce (NCCL transport).
//   vector   alpha = r.z / p.q;  the r.z of the NEXT residual follows from the sums without touching the vectors,
//            r'.z' = r.z - 2 alpha q.z + alpha^2 q.D^-1.q, so beta = r'.z' / r.z is known up front and one pass does
//            x += alpha p,  r -= alpha q,  p = D^-1 r + beta p  and the direct sums r.z, r.r of the new residual.
// The predicted r'.z' is used for beta only; alpha and the convergence test always use directly summed values, so the
// recurrence cannot drift.  The classical form needs a second reduction between the residual update and the direction
// update (two grid-wide / machine-wide sync points per iteration); this form has one.
//
// The matrix is the scalar CSR viewed at node level (one column index per 2x2 sub-block: 9 bytes per stored scalar
// instead of 12).  Two SpMV kernels read it:
//   k_spmv_ring (spmv_ring.cuh)  the SpMV of the PCG loop: a producer warp streams row tiles into a shared-memory ring with
//                                cp.async.bulk (TMA), consumer warps multiply out of shared memory; <true> variant with the
//                                NVLink halo exchange inside the kernel
//   k_spmv                       row kernel (8 lanes per node row chase pointer -> {index, values} -> x through global
//                                loads, 4 independent load groups per lane, evict-first streaming): mofem_spmv, residual
//                                checks, energy, the NCCL transport, rows too long for a ring stage
// The kernels of the loop are chained with programmatic dependent launch (single GPU and peer transport): the ring's
// producer fills its stages while the vector kernel drains.
//
// Dot products have a fixed reduction order (bit-reproducible): per-thread sums over a static assignment, per-CTA
// partials summed in index order by the last CTA to finish, per-rank sums added in rank order.  alpha/beta live on the
// device; the vector kernel that sees r.r <= tol^2 raises a stop flag which turns the rest of the enqueued chunk into
// no-ops, so the host polls only every PCG_CHUNK iterations.
#include <algorithm>

#include "common.cuh"
#include "spmv_ring.cuh"

namespace mofem {

constexpr int VEC_BLOCK = 256;
constexpr int SPMV_BLOCK = 128;
constexpr int SPMV_LANES = 8;    // lanes cooperating on one node row
constexpr int SPMV_UNROLL = 4;   // independent load groups per lane
constexpr int SPMV_MINB = 12;    // resident blocks per SM
constexpr int SPMV_GRID_CAP = 148 * 24;
static_assert(SPMV_GRID_CAP <= PCG_MAX_PART, "partials buffer too small");

struct Parts { double *p[5]; };   // per-CTA partial sums: [0..2] SpMV (p.q, q.z, q.D^-1.q), [3..4] vector kernel (r.z, r.r)

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

// Publishes this block's N partials under one counter; the last block to arrive returns true after which total[] (valid
// in every thread of that block) holds the sums of all partials in index order -- independent of which block came last.
template <int N>
__device__ __forceinline__ bool grid_sum_n(const double (&v)[N], double *const *part, unsigned int *counter, double *sm,
                                           double (&total)[N])
{
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int j = 0; j < N; j++) part[j][blockIdx.x] = v[j];
        __threadfence();
        is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
    double a[N];
#pragma unroll
    for (int j = 0; j < N; j++) a[j] = 0.0;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
#pragma unroll
        for (int j = 0; j < N; j++) a[j] += __ldcg(part[j] + i);
    }
#pragma unroll
    for (int j = 0; j < N; j++) total[j] = block_sum(a[j], sm);
    if (threadIdx.x == 0) *counter = 0u;
    return true;
}

__device__ __forceinline__ bool grid_sum(double partial, double *part, unsigned int *counter, double *sm, double &total)
{
    const double v[1] = {partial};
    double *const pp[1] = {part};
    double t[1];
    const bool last = grid_sum_n<1>(v, pp, counter, sm, t);
    total = t[0];
    return last;
}

// ---- NVLink peer-memory primitives (row-partitioned solve, one process per GPU) ---------------------------------
// Every rank owns an exchange buffer that all peers map (CUDA IPC).  Halo values travel as {payload, stamp}: the writer
// stores the payload into the consumer's buffer, fences at system scope, then stores the stamp; the consumer polls the
// stamp in its OWN memory.  Stamps grow monotonically (epoch, iteration), two parity slots make reuse safe: a rank can
// only be two rounds ahead of a peer after that peer has consumed the older round.
constexpr long long P2P_TIMEOUT_CLOCKS = 40000000000ll;   // ~20 s at 2 GHz: a peer died
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

// All-reduce of the five sums of an iteration in "LL" form (the trick NCCL's low-latency protocol uses): every 8-byte word
// carries 32 bits of payload and a 32-bit stamp, and an aligned 8-byte store is atomic over NVLink -- so no fence and no
// separate flag: one-way latency only.
// producer side: called by ONE block (>= world threads) after it holds this rank's five sums
__device__ __forceinline__ void p2p_publish5(const P2PDev &pp, const double (&v)[5])
{
    if ((int)threadIdx.x < pp.world) {
        volatile unsigned long long *slot =
            reinterpret_cast<volatile unsigned long long *>(pp.peer[threadIdx.x] + p2p_mbox_off(pp.world, pp.parity, pp.rank));
        const unsigned long long f = (unsigned long long)pp.stamp32 << 32;
#pragma unroll
        for (int j = 0; j < 5; j++) {
            const unsigned long long a = (unsigned long long)__double_as_longlong(v[j]);
            slot[2 * j] = f | (a & 0xffffffffull);
            slot[2 * j + 1] = f | (a >> 32);
        }
    }
}

// consumer side: every block (>= world threads) sums the W contributions in rank order (bit-identical on every rank and block)
__device__ __forceinline__ void p2p_collect5(const P2PDev &pp, double (&v)[5])
{
    __shared__ double sv[5][P2P_MAX_WORLD];
    if ((int)threadIdx.x < pp.world) {
        const volatile unsigned long long *slot =
            reinterpret_cast<const volatile unsigned long long *>(pp.local + p2p_mbox_off(pp.world, pp.parity, threadIdx.x));
        unsigned long long wv[10];
        long long t0 = 0;
        for (unsigned spins = 0;; spins++) {
            bool ok = true;
#pragma unroll
            for (int j = 0; j < 10; j++) { wv[j] = slot[j]; ok = ok && (unsigned int)(wv[j] >> 32) == pp.stamp32; }
            if (ok) break;
            if ((spins & 0xfff) == 0xfff) {
                const long long now = clock64();
                if (t0 == 0) t0 = now;
                else if (now - t0 > P2P_TIMEOUT_CLOCKS) { *pp.err = 1; break; }
            }
        }
#pragma unroll
        for (int j = 0; j < 5; j++)
            sv[j][threadIdx.x] = __longlong_as_double((long long)((wv[2 * j] & 0xffffffffull) | (wv[2 * j + 1] << 32)));
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 5; j++) {
        double a = 0.0;
        for (int r = 0; r < pp.world; r++) a += sv[j][r];
        v[j] = a;
    }
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
template <bool COHERENT>
__device__ __forceinline__ double2 load_x(const double2 *p)
{
    return COHERENT ? __ldcg(p) : __ldg(p);
}

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
                xv[u] = load_x<COHERENT>(reinterpret_cast<const double2 *>(x + 2 * (int64_t)c));
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

// reduces the lanes of a row group, masks fixed rows, stores y and adds the row to the sums (lane sub == 0)
template <int MODE>
__device__ __forceinline__ void spmv_finish_row(int64_t a, int sub, double y0, double y1, double *__restrict__ y,
                                                const uint8_t *__restrict__ fixed, const double *x, const double *r,
                                                const double *dinv, RingDots &dots)
{
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
        if (MODE == 1) {
            const double2 xa = *reinterpret_cast<const double2 *>(x + 2 * a);
            dots.pq = fma(xa.x, y0, dots.pq); dots.pq = fma(xa.y, y1, dots.pq);
        }
        if (MODE == 2)
            ring_row_dots(dots, *reinterpret_cast<const double2 *>(x + 2 * a), *reinterpret_cast<const double2 *>(r + 2 * a),
                          *reinterpret_cast<const double2 *>(dinv + 2 * a), y0, y1);
    }
}

// End of a PCG SpMV: the CTA's three sums -> grid totals; the last CTA appends r.z and r.r of the current residual and hands
// the five numbers on (out5: sc + SC_PQ on one GPU, sc + SC_PART for the NCCL all-reduce; peer transport: the mailboxes).
template <bool P2P>
__device__ __forceinline__ void pcg_spmv_epilogue(const RingDots &dots, Parts part, unsigned int *counter, double *sc,
                                                  double *out5, const P2PDev &pp, double *sm)
{
    const double v[3] = {block_sum(dots.pq, sm), block_sum(dots.qz, sm), block_sum(dots.qdq, sm)};
    double t[3];
    if (grid_sum_n<3>(v, part.p, counter, sm, t)) {
        const double five[5] = {t[0], t[1], t[2], __ldcg(sc + SC_LOC_RZ), __ldcg(sc + SC_LOC_RR)};
        if (P2P) p2p_publish5(pp, five);
        else if (threadIdx.x == 0) {
#pragma unroll
            for (int j = 0; j < 5; j++) out5[j] = five[j];
        }
    }
}

// this rank's boundary entries of x -> the ghost segments of the neighbours' x over NVLink, then the stamp; called by the
// `nthr` threads of ONE CTA that share `sync`
template <class Sync>
__device__ __forceinline__ void push_halo(const HaloPush &hp, const P2PDev &pp, const double *x, int nthr, Sync sync)
{
    const double2 *x2 = reinterpret_cast<const double2 *>(x);
    for (int64_t i = threadIdx.x; i < hp.n_send; i += nthr) {
        const int k = hp.send_nb[i];
        double2 *dst = reinterpret_cast<double2 *>(pp.peer[hp.nb_rank[k]] + hp.nb_ghost_off[k]) + hp.send_dst[i];
        *dst = x2[hp.send_idx[i]];
    }
    __threadfence_system();
    sync();
    if ((int)threadIdx.x < hp.n_nb)
        *reinterpret_cast<volatile double *>(pp.peer[hp.nb_rank[threadIdx.x]] + p2p_hflag_off(pp.world, pp.parity, pp.rank)) = pp.stamp;
}

// y = K x on node rows; fixed rows -> 0.  Ghost columns are the tail of x.
// MODE 0: product only.  MODE 1: also x.y -> *out (energy).  MODE 2: SpMV of the PCG loop (sums and hand-off as above).
// P2P (row-partitioned, peer transport) -- one kernel does the halo exchange, the product and the all-reduce hand-off:
//   * block 0 first writes this rank's boundary entries of x into the ghost segment of the neighbours' x over NVLink and
//     raises its stamp there;
//   * every block first multiplies its rows that touch no ghost column; the few that do (halo_rows) are dealt out afterwards,
//     behind a wait for the neighbours' stamps -- the NVLink latency hides behind the interior rows;
//   * the last block publishes the five sums into every peer's mailbox.
template <int MODE, bool P2P>
__global__ void __launch_bounds__(SPMV_BLOCK, SPMV_MINB)
k_spmv(int64_t n_node, const int64_t *__restrict__ node_ptr, const int32_t *__restrict__ node_idx,
       const double *__restrict__ data, const double *x, double *__restrict__ y,
       const uint8_t *__restrict__ fixed, const double *__restrict__ r, const double *__restrict__ dinv, Parts part,
       unsigned int *counter, double *sc, double *out, P2PDev pp, int n_wait, const int *__restrict__ wait_rank,
       const uint8_t *__restrict__ row_halo, const int32_t *__restrict__ halo_rows, int64_t n_halo, HaloPush hp)
{
    __shared__ double sm[SPMV_BLOCK / 32];
    ring::grid_dep_wait();   // PDL: no-op unless launched with the programmatic-serialization attribute
    if (MODE == 2 && __ldcg(sc + SC_DONE) != 0.0) return;
    if (P2P && blockIdx.x == 0 && hp.n_nb) push_halo(hp, pp, x, SPMV_BLOCK, [] { __syncthreads(); });
    const int sub = threadIdx.x % SPMV_LANES;
    constexpr int64_t rows_per_block = SPMV_BLOCK / SPMV_LANES;
    RingDots dots{0.0, 0.0, 0.0};
    for (int64_t row0 = (int64_t)blockIdx.x * rows_per_block; row0 < n_node; row0 += (int64_t)gridDim.x * rows_per_block) {
        int64_t a = row0 + threadIdx.x / SPMV_LANES;
        if (a >= n_node || (P2P && row_halo[a])) a = -1;        // halo rows come last (below)
        double y0 = 0.0, y1 = 0.0;
        if (a >= 0) spmv_row(a, sub, node_ptr, node_idx, data, x, y0, y1);
        spmv_finish_row<MODE>(a, sub, y0, y1, y, fixed, x, r, dinv, dots);
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
                spmv_finish_row<MODE>(a, sub, y0, y1, y, fixed, x, r, dinv, dots);
            }
        }
    }
    ring::grid_dep_launch();   // the dependent (vector kernel) may become resident from here on
    if (MODE == 1) {
        const double s = block_sum(dots.pq, sm);
        double total;
        if (grid_sum(s, part.p[0], counter, sm, total) && threadIdx.x == 0) *out = total;
    }
    if (MODE == 2) pcg_spmv_epilogue<P2P>(dots, part, counter, sc, out, pp, sm);
}

// ---- TMA-ring SpMV of the PCG loop (spmv_ring.cuh) -----------------------------------------------------------------
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
// q = K p and the sums of the iteration.  Launched with the PDL attribute the producer warp fills the ring while the vector
// kernel of the previous iteration is still draining (the matrix does not depend on it).
// P2P (row-partitioned, peer transport): the same ring for the rows that touch no ghost column, with the halo exchange of
// k_spmv around it -- CTA 0 first stores this rank's boundary entries of p into the neighbours' ghost segments over NVLink
// and raises its stamp there; the rows that gather ghosts (a compact list, their tiles carry the mask flag) are multiplied
// after the ring, behind a wait for the neighbours' stamps; the last CTA publishes the five sums into every peer's mailbox.
__device__ __forceinline__ void consumer_barrier() { asm volatile("bar.sync 1, %0;" ::"r"(PcgRing::CW * 32) : "memory"); }

template <bool P2P>
__global__ void __launch_bounds__(PcgRing::THREADS, 2)
k_spmv_ring(int n_tile, const RingTile *__restrict__ tiles, const int64_t *__restrict__ node_ptr,
            const int32_t *__restrict__ node_idx, const double *__restrict__ data, const double *x, double *__restrict__ y,
            const uint8_t *__restrict__ fixed, const double *__restrict__ r, const double *__restrict__ dinv, Parts part,
            unsigned int *counter, double *sc, P2PDev pp, int n_wait, const int *__restrict__ wait_rank,
            const uint8_t *__restrict__ row_halo, const int32_t *__restrict__ halo_rows, int64_t n_halo, HaloPush hp)
{
    extern __shared__ __align__(128) unsigned char ring_smem[];
    __shared__ double sm[32];
    const RingState rs = ring_prologue<PcgRing>(ring_smem, n_tile, tiles, node_ptr, node_idx, data, sc + SC_DONE);
    if (!rs.active) return;
    const bool consumer = (int)(threadIdx.x >> 5) < PcgRing::CW;
    constexpr int NC = PcgRing::CW * 32;
    if (P2P && blockIdx.x == 0 && hp.n_nb && consumer) push_halo(hp, pp, x, NC, [] { consumer_barrier(); });
    RingDots dots{0.0, 0.0, 0.0};
    ring_main<PcgRing, P2P>(ring_smem, rs, n_tile, tiles, node_ptr, node_idx, data, x, y, fixed, r, dinv, dots, row_halo);
    if (P2P) {
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
                spmv_finish_row<2>(a, sub, y0, y1, y, fixed, x, r, dinv, dots);
            }
        }
    }
    ring::grid_dep_launch();   // the vector kernel may become resident from here on
    pcg_spmv_epilogue<P2P>(dots, part, counter, sc, sc + SC_PQ, pp, sm);
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
// The vector kernels work on double2 (n = 2 * n_node is even).
// init: x = 0, r = b (or keep x, r), p = dinv*r ; this rank's r.z and r.r -> sc[SC_LOC_*] (the first SpMV hands them on);
// a fresh solve also leaves r.r = |b|^2 (this rank's part) in sc[SC_BNORM2]
__global__ void __launch_bounds__(VEC_BLOCK)
k_pcg_init(int64_t n2, const double2 *__restrict__ b, const double2 *__restrict__ dinv, double2 *__restrict__ x,
           double2 *__restrict__ r, double2 *__restrict__ p, Parts part, unsigned int *counter, double *__restrict__ sc, int keep_x)
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
    const double v[2] = {block_sum(rz, sm), block_sum(rr, sm)};
    double t[2];
    if (grid_sum_n<2>(v, part.p + 3, counter, sm, t) && threadIdx.x == 0) {
        sc[SC_LOC_RZ] = t[0]; sc[SC_LOC_RR] = t[1];
        sc[SC_DONE] = 0.0; sc[SC_IT_DONE] = 0.0;
        if (!keep_x) sc[SC_BNORM2] = t[1];
    }
}

// The vector part of iteration `it` (x has received `it` updates when it starts).  With the five sums of the iteration
//   alpha = r.z / p.q        r'.z' = r.z - 2 alpha q.z + alpha^2 q.D^-1.q        beta = r'.z' / r.z
// one pass does  x += alpha p ;  r -= alpha q ;  p = D^-1 r + beta p  and sums r.z, r.r of the new residual directly.
// If the current residual already meets the tolerance nothing is updated and the stop flag goes up.
constexpr int VEC_BATCH = 2, VEC_BPS = 3;
template <bool P2P>
__global__ void __launch_bounds__(VEC_BLOCK, VEC_BPS)
k_pcg_vec(int64_t n2, int it, double *__restrict__ sc, double2 *__restrict__ p, const double2 *__restrict__ q,
          const double2 *__restrict__ dinv, double2 *__restrict__ x, double2 *__restrict__ r, Parts part,
          unsigned int *counter, P2PDev pp, double tol2)
{
    __shared__ double sm[VEC_BLOCK / 32];
    ring::grid_dep_wait();        // PDL: no-op unless launched with the programmatic-serialization attribute
    if (__ldcg(sc + SC_DONE) != 0.0) return;
    double v[5];
    if (P2P) {
        p2p_collect5(pp, v);
        if (blockIdx.x == 0 && threadIdx.x == 0) {   // for the host poll
#pragma unroll
            for (int j = 0; j < 5; j++) sc[SC_PQ + j] = v[j];
        }
    } else {
#pragma unroll
        for (int j = 0; j < 5; j++) v[j] = __ldcg(sc + SC_PQ + j);
    }
    if (v[SC_RR] <= tol2) {
        // every CTA takes this branch (same five numbers everywhere) or returns above once the flag is up: x, r, p stay
        if (blockIdx.x == 0 && threadIdx.x == 0) { sc[SC_IT_DONE] = (double)it; __threadfence(); sc[SC_DONE] = 1.0; }
        return;
    }
    ring::grid_dep_launch();      // the next SpMV may become resident from here on (its producer warp pre-fills the ring)
    const double pq = v[SC_PQ], rz = v[SC_RZ];
    const double alpha = (pq != 0.0) ? rz / pq : 0.0;
    const double rz_next = fma(alpha, fma(alpha, v[SC_QDQ], -2.0 * v[SC_QZ]), rz);
    const double beta = (rz != 0.0 && rz_next > 0.0) ? rz_next / rz : 0.0;
    double s_rz = 0.0, s_rr = 0.0;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    for (int64_t base = tid; base < n2; base += VEC_BATCH * nth) {
        double2 pv[VEC_BATCH], qv[VEC_BATCH], dv[VEC_BATCH], xv[VEC_BATCH], rv[VEC_BATCH];
#pragma unroll
        for (int e = 0; e < VEC_BATCH; e++) {
            const int64_t i = base + e * nth;
            if (i < n2) { qv[e] = q[i]; rv[e] = r[i]; dv[e] = dinv[i]; pv[e] = p[i]; xv[e] = x[i]; }
        }
#pragma unroll
        for (int e = 0; e < VEC_BATCH; e++) {
            const int64_t i = base + e * nth;
            if (i < n2) {
                const double2 rn = make_double2(fma(-alpha, qv[e].x, rv[e].x), fma(-alpha, qv[e].y, rv[e].y));
                const double2 z = make_double2(dv[e].x * rn.x, dv[e].y * rn.y);
                x[i] = make_double2(fma(alpha, pv[e].x, xv[e].x), fma(alpha, pv[e].y, xv[e].y));
                r[i] = rn;
                p[i] = make_double2(fma(beta, pv[e].x, z.x), fma(beta, pv[e].y, z.y));
                s_rz = fma(rn.x, z.x, s_rz); s_rz = fma(rn.y, z.y, s_rz);
                s_rr = fma(rn.x, rn.x, s_rr); s_rr = fma(rn.y, rn.y, s_rr);
            }
        }
    }
    const double bs[2] = {block_sum(s_rz, sm), block_sum(s_rr, sm)};
    double t[2];
    if (grid_sum_n<2>(bs, part.p + 3, counter, sm, t) && threadIdx.x == 0) { sc[SC_LOC_RZ] = t[0]; sc[SC_LOC_RR] = t[1]; }
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
static Parts parts_of(const PcgBuffers &w)
{
    Parts pt;
    for (int j = 0; j < 5; j++) pt.p[j] = w.part[j];
    return pt;
}

// plain launch: ghosts (if any) are the tail of x.  MODE 0 product, MODE 1 product and x.y -> *out
template <int MODE>
static void launch_spmv(const DevSystem &A, const double *x, double *y, const uint8_t *fixed, const PcgBuffers *w, unsigned int *counter,
                        double *out, cudaStream_t s)
{
    const int64_t n_node = A.n_own / 2;
    k_spmv<MODE, false><<<spmv_grid(n_node), SPMV_BLOCK, 0, s>>>(n_node, A.node_ptr, A.node_idx, A.data, x, y, fixed, nullptr, nullptr,
                                                                  w ? parts_of(*w) : Parts{}, counter, nullptr, out, NO_P2P, 0, nullptr,
                                                                  nullptr, nullptr, 0, HaloPush{});
}

// y = K x on the rows this rank owns (all rows on a single GPU); x must hold current ghost values
cudaError_t spmv(const DevSystem &A, const double *x, double *y, const uint8_t *fixed, cudaStream_t s)
{
    if (A.n_own == 0) return cudaSuccess;
    launch_spmv<0>(A, x, y, fixed, nullptr, nullptr, nullptr, s);
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
    launch_spmv<0>(A, w.x, w.q, fixed, nullptr, nullptr, nullptr, s);
    k_true_residual<<<vec_grid(n_own), VEC_BLOCK, 0, s>>>(n_own, w.b, w.q, w.r, w.part[4], w.counters + 2, w.sc + SC_RR_TRUE);
    if (launches) *launches += 2;
    if ((e = all_reduce(dist, w.sc + SC_RR_TRUE, 1, s)) != cudaSuccess) return e;
    e = cudaMemcpyAsync(rr_true, w.sc + SC_RR_TRUE, sizeof(double), cudaMemcpyDeviceToHost, s);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(s);
}

// Runs PCG; returns via info.  `chunk` iterations are enqueued between host polls (the device stops on its own).
// prof (optional): 2 pools of 4*chunk events; prof_ms[0..1] accumulate the device time of the SpMV (incl. halo exchange
// and the all-reduce when row-partitioned) / the vector kernel of every prof_every-th iteration, prof_ms[3] the iterations
// profiled.  The events of a chunk are read back while the next chunk runs, so profiling does not stall the GPU.
//
// Row-partitioned, two transports:
//   NCCL  -- per iteration: pack + ncclSend/Recv of the ghost entries of p, ONE ncclAllReduce of the five sums
//   peer  -- (dist->p2p) no NCCL call inside the loop: CTA 0 of the SpMV writes this rank's boundary entries of p into the
//            neighbours' ghost segments over NVLink, the rows that gather ghosts wait for the neighbours' stamps; the last
//            CTA of the SpMV writes the five sums into every peer's mailbox and the vector kernel adds the W contributions
//            in rank order.  The first iteration after (re)initialisation still uses NCCL for the halo.
cudaError_t pcg_run(const DevSystem &A, const uint8_t *fixed, PcgBuffers &w, DistPlan *dist, double rtol, int maxit,
                    int chunk, mofem_solve_info_t *info, cudaStream_t s, int64_t *launches, cudaEvent_t *prof, double *prof_ms, int prof_every)
{
    if (prof_every < 1) prof_every = 1;
    const int64_t n_own = A.n_own, n_node = n_own / 2, n2 = n_own / 2;
    const int gv = vec_grid(n2, 148 * VEC_BPS), gs = spmv_grid(n_node);
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
    const Parts parts = parts_of(w);
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
        k_pcg_init<<<vec_grid(n2), VEC_BLOCK, 0, s>>>(n2, b2, dinv2, x2, r2, p2, parts, w.counters + 1, w.sc, keep_x);
        if (launches) ++*launches;
        if (!keep_x) {   // |b|^2 over all ranks (once per solve)
            const cudaError_t ee = all_reduce(dist, w.sc + SC_BNORM2, 1, s);
            if (ee != cudaSuccess) return ee;
        }
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
    int it = 0;       // updates x has received
    double prev_rr_true = 0.0;
    bool fresh = true;   // p was just (re)initialised: its halo has not been pushed by a previous iteration
    int pend_pool = -1, pend_n = 0, pool = 0;
    auto harvest = [&]() {
        if (!prof || pend_pool < 0) return;
        cudaEvent_t *ev = prof + (size_t)pend_pool * 4 * chunk;
        for (int k = 0; k < pend_n; k++)
            for (int j = 0; j < 2; j++) {
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
        if ((e = cudaFuncSetAttribute(k_spmv_ring<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PcgRing::SMEM_BYTES)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(k_spmv_ring<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PcgRing::SMEM_BYTES)) != cudaSuccess) return e;
    }
    // programmatic dependent launch between the kernels of the loop; measured: costs time with the row kernel
    // (181 vs 133 us per iteration), and a run that times every iteration wants the kernels back to back
    const bool pdl = (single || peer) && w.use_pdl && ring_on && !(prof && prof_every == 1);
    const int g_ring = w.n_tile < RING_GRID ? w.n_tile : RING_GRID;
    const RingTile *tiles = reinterpret_cast<const RingTile *>(w.ring_tiles);
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
            if (!peer) {
                if (multi && (e = halo_exchange(dist, w.p, w.p + n_own, s, launches)) != cudaSuccess) return done(e);
                if (ring_on)
                    e = launch_ex(pdl, k_spmv_ring<false>, g_ring, PcgRing::THREADS, PcgRing::SMEM_BYTES, s, w.n_tile, tiles, A.node_ptr,
                                  A.node_idx, A.data, (const double *)w.p, w.q, fixed, (const double *)w.r, (const double *)w.dinv, parts,
                                  w.counters, w.sc, NO_P2P, 0, (const int *)nullptr, (const uint8_t *)nullptr, (const int32_t *)nullptr,
                                  (int64_t)0, HaloPush{});
                else
                    e = launch_ex(false, k_spmv<2, false>, gs, SPMV_BLOCK, 0, s, n_node, A.node_ptr, A.node_idx, A.data, (const double *)w.p,
                                  w.q, fixed, (const double *)w.r, (const double *)w.dinv, parts, w.counters, w.sc,
                                  w.sc + (multi ? SC_PART : SC_PQ), NO_P2P, 0, (const int *)nullptr, (const uint8_t *)nullptr,
                                  (const int32_t *)nullptr, (int64_t)0, HaloPush{});
                if (e != cudaSuccess) return done(e);
                // NCCL: the partial sums are all-reduced INTO the global slots, so re-reducing after the device has stopped is idempotent
                if (multi && (e = all_reduce(dist, w.sc + SC_PQ, 5, s, w.sc + SC_PART)) != cudaSuccess) return done(e);
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
                    e = launch_ex(pdl && !fresh, k_spmv_ring<true>, g_ring, PcgRing::THREADS, PcgRing::SMEM_BYTES, s, w.n_tile, tiles, A.node_ptr,
                                  A.node_idx, A.data, (const double *)w.p, w.q, fixed, (const double *)w.r, (const double *)w.dinv, parts,
                                  w.counters, w.sc, pp, n_wait, (const int *)dist->wait_rank, (const uint8_t *)w.row_halo,
                                  (const int32_t *)dist->halo_rows, dist->n_halo_rows, hp);
                else
                    e = launch_ex(false, k_spmv<2, true>, gs, SPMV_BLOCK, 0, s, n_node, A.node_ptr, A.node_idx, A.data, (const double *)w.p, w.q,
                                  fixed, (const double *)w.r, (const double *)w.dinv, parts, w.counters, w.sc, (double *)nullptr, pp, n_wait,
                                  (const int *)dist->wait_rank, (const uint8_t *)w.row_halo, (const int32_t *)dist->halo_rows,
                                  dist->n_halo_rows, hp);
                if (e != cudaSuccess) return done(e);
                fresh = false;
            }
            if (ev) cudaEventRecord(ev[4 * k + 1], s);
            if (peer)
                e = launch_ex(pdl, k_pcg_vec<true>, gv, VEC_BLOCK, 0, s, n2, it, w.sc, p2, (const double2 *)q2, dinv2, x2, r2, parts,
                              w.counters + 1, pp, tol2);
            else
                e = launch_ex(pdl, k_pcg_vec<false>, gv, VEC_BLOCK, 0, s, n2, it, w.sc, p2, (const double2 *)q2, dinv2, x2, r2, parts,
                              w.counters + 1, NO_P2P, tol2);
            if (e != cudaSuccess) return done(e);
            if (ev) cudaEventRecord(ev[4 * k + 2], s);
        }
        if (launches) *launches += 2 * (int64_t)todo;
        harvest();  // events of the previous chunk, while this one runs
        e = cudaMemcpyAsync(sc_host, w.sc, sizeof(double) * SC_COUNT, cudaMemcpyDeviceToHost, s);
        if (e != cudaSuccess) return done(e);
        int err_host = 0;
        if (peer) e = cudaMemcpyAsync(&err_host, err_dev, sizeof(int), cudaMemcpyDeviceToHost, s);
        if (e != cudaSuccess) return done(e);
        e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) return done(e);
        if (err_host) return done(cudaErrorLaunchTimeout);   // a peer did not answer within P2P_TIMEOUT_CLOCKS (~20 s)
        if (prof) { pend_pool = pool; pend_n = n_sampled; pool ^= 1; }
        const bool stopped = sc_host[SC_DONE] != 0.0;
        if (stopped) it = (int)sc_host[SC_IT_DONE];   // the device stopped inside the chunk: x has `it` updates
        info->iterations = it;
        // r.r over all ranks as the last vector kernel that ran saw it (the residual BEFORE its update, or the one that stopped it)
        const double rr = sc_host[SC_RR];
        info->relres_rec = sqrt(rr / bnorm2);
        if (!(rr == rr)) break;  // NaN: breakdown
        if (stopped) {
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
            if ((e = init(1)) != cudaSuccess) return done(e);
            fresh = true;
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
    launch_spmv<1>(A, u, w.q, nullptr, &w, w.counters + 3, out_dev, s);
    cudaError_t e = all_reduce(dist, out_dev, 1, s);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
}

}  // namespace mofem
