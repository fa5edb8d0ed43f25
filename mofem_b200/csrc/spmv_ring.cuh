// spmv_ring.cuh -- block-CSR SpMV with the matrix streamed through shared memory by the TMA (1-D bulk async copies).
//
// The row kernel (pcg.cu: k_spmv) has every 8-lane group chase node_ptr -> {indices, values} -> x through three dependent
// global loads, so the bytes in flight per SM -- and with them the HBM rate -- are set by how many warps happen to sit in
// the "values" phase.  Here the stream is decoupled from the arithmetic:
//   * the matrix is cut into TILES of consecutive node rows (<= SPAN rows + sub-blocks each, built once per matrix);
//   * one PRODUCER warp per CTA walks the CTA's tiles and, per tile, issues three cp.async.bulk copies (values, column
//     indices, row pointers) into the next free stage of a shared-memory ring, completion on an mbarrier ("full");
//   * CW CONSUMER warps wait on "full", multiply their rows out of shared memory (8 lanes per node row, the only global
//     access left is the gather of x, which lives in L2/L1), and arrive on "empty" to hand the stage back.
// Warps move through the ring independently, so up to STAGES tiles per CTA are in flight regardless of what the
// consumers are doing.  With PDL the producer fills the ring's STAGES stages before the previous kernel of the PCG iteration
// has finished (the matrix does not depend on it); then both sides order themselves behind that kernel and look at the
// solver's stop flag -- if it is up, the producer lets its copies land and the whole CTA leaves (a no-op SpMV costs
// microseconds, not a pass over the matrix).
// Measured on the 1 M-cell bench matrix (ncu, kernels in sequence, warm L2): 82.9 us = 5.6 TB/s of DRAM traffic.  Things that
// were tried and made it slower are listed in DESIGN.md (staging the rows' vector entries through the ring, explicit
// ld.shared addressing, three fused dot products instead of one).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mofem {

struct alignas(16) RingTile {
    int64_t b0, b1;   // sub-blocks [b0, b1) of the node-level CSR
    int32_t a0, a1;   // node rows [a0, a1)
    int64_t flags;    // bit 0: some row of the tile is masked (row_skip), consumers then test every row
};
static_assert(sizeof(RingTile) == 32, "RingTile is read as 16-byte words");

template <int SPAN_, int MAXDEG_, int STAGES_, int CW_, bool EVICT_FIRST_ = true>
struct RingCfg {
    static constexpr bool EVICT_FIRST = EVICT_FIRST_;   // matrix copies carry an L2 evict-first hint (the vectors stay resident)
    static constexpr int SPAN = SPAN_;          // rows + sub-blocks per tile (key a + node_ptr[a])
    static constexpr int MAXDEG = MAXDEG_;      // largest node-row degree the ring can hold
    static constexpr int STAGES = STAGES_;
    static constexpr int CW = CW_;              // consumer warps
    static constexpr int BCAP = SPAN + MAXDEG;  // sub-blocks per stage
    static constexpr int VAL_OFF = 0;
    static constexpr int IDX_OFF = 32 * BCAP;
    static constexpr int PTR_OFF = IDX_OFF + ((4 * (BCAP + 4) + 15) & ~15);
    static constexpr int HDR_OFF = PTR_OFF + ((8 * (SPAN + 3) + 15) & ~15);
    static constexpr int STAGE_BYTES = (HDR_OFF + 32 + 127) & ~127;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 16 * STAGES + 128;
    static constexpr int THREADS = (CW + 1) * 32;
};

namespace ring {
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase)
{
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t pol)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void grid_dep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
}  // namespace ring

// y = K x on node rows (fixed rows -> 0); returns this thread's share of x.y in `dot` (consumer lanes with sub == 0).
// MASKED: rows with row_skip[row] != 0 are left out (their tiles carry flag bit 0).
// done_flag (may be null): the solver's stop flag, read behind the grid dependency; returns false (nothing computed, every
// thread of the CTA) when it is up.
// Called by every thread of a CTA of Cfg::THREADS threads with Cfg::SMEM_BYTES of dynamic shared memory.
// node_idx must be readable up to 12 bytes and node_ptr up to 8 bytes past their ends (the copies are 16-byte granular).
template <class Cfg, bool TRIG_LATE = false, bool MASKED = false>
__device__ __forceinline__ bool ring_spmv_body(unsigned char *smem, int n_tile, const RingTile *__restrict__ tiles,
                                               const int64_t *__restrict__ node_ptr, const int32_t *__restrict__ node_idx,
                                               const double *__restrict__ data, const double *__restrict__ x,
                                               double *__restrict__ y, const uint8_t *__restrict__ fixed, double &dot,
                                               const uint8_t *__restrict__ row_skip = nullptr, const double *done_flag = nullptr)
{
    constexpr int S = Cfg::STAGES, CW = Cfg::CW;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + S * Cfg::STAGE_BYTES);
    uint64_t *empty = full + S;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; s++) { ring::mbar_init(&full[s], 1); ring::mbar_init(&empty[s], CW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == CW) {
        // ---- producer: the matrix does not depend on the previous kernel, so the first STAGES tiles stream before the grid
        // dependency; the header of the next tile is fetched BEFORE the wait for its stage, so the copies go out the moment the
        // consumers hand the stage back
        if (lane == 0) {
            const uint64_t pol = Cfg::EVICT_FIRST ? ring::policy_evict_first() : 0;
            int k = 0;
            for (int t = blockIdx.x; t < n_tile; t += gridDim.x, k++) {
                const int s = k % S;
                unsigned char *st = smem + s * Cfg::STAGE_BYTES;
                const longlong2 h0 = __ldg(reinterpret_cast<const longlong2 *>(tiles + t));
                const longlong2 h1 = __ldg(reinterpret_cast<const longlong2 *>(tiles + t) + 1);   // {a0 | a1 << 32, flags}
                if (k == S) {
                    ring::grid_dep_wait();
                    if (done_flag && __ldcg(done_flag) != 0.0) break;
                }
                if (k >= S) ring::mbar_wait(&empty[s], ((k / S) - 1) & 1);
                *reinterpret_cast<longlong2 *>(st + Cfg::HDR_OFF) = h0;
                *reinterpret_cast<longlong2 *>(st + Cfg::HDR_OFF + 16) = h1;
                const int64_t b0 = h0.x, b1 = h0.y;
                const int a0 = (int)(h1.x & 0xffffffffll), a1 = (int)(h1.x >> 32);
                const int64_t i0 = b0 & ~(int64_t)3;
                const int p0 = a0 & ~1;
                const uint32_t bytes_v = (uint32_t)(b1 - b0) * 32u;
                const uint32_t bytes_i = bytes_v ? (((uint32_t)(b1 - i0) * 4u + 15u) & ~15u) : 0u;
                const uint32_t bytes_p = ((uint32_t)(a1 + 1 - p0) * 8u + 15u) & ~15u;
                ring::mbar_arrive_tx(&full[s], bytes_v + bytes_i + bytes_p);
                if (bytes_v) {
                    if (Cfg::EVICT_FIRST) {
                        ring::bulk_g2s_hint(st + Cfg::VAL_OFF, data + 4 * b0, bytes_v, &full[s], pol);
                        ring::bulk_g2s_hint(st + Cfg::IDX_OFF, node_idx + i0, bytes_i, &full[s], pol);
                    } else {
                        ring::bulk_g2s(st + Cfg::VAL_OFF, data + 4 * b0, bytes_v, &full[s]);
                        ring::bulk_g2s(st + Cfg::IDX_OFF, node_idx + i0, bytes_i, &full[s]);
                    }
                }
                ring::bulk_g2s(st + Cfg::PTR_OFF, node_ptr + p0, bytes_p, &full[s]);
            }
            // a CTA must not exit with bulk copies in flight: when the consumers leave early (stop flag) nobody waits on the
            // stages this lane filled, so it does
            ring::grid_dep_wait();
            if (done_flag && __ldcg(done_flag) != 0.0) {
                const int filled = k < S ? k : S;
                for (int s = 0; s < filled; s++) ring::mbar_wait(&full[s], 0);
            }
        }
        ring::grid_dep_wait();       // the other lanes of the warp: the flag may only be read behind the dependency
        if (!TRIG_LATE) ring::grid_dep_launch();
        return !(done_flag && __ldcg(done_flag) != 0.0);
    }
    // ---- consumers
    ring::grid_dep_wait();       // no-ops unless the kernel was launched with the programmatic-serialization attribute
    if (!TRIG_LATE) ring::grid_dep_launch();
    if (done_flag && __ldcg(done_flag) != 0.0) return false;
    const int sub = lane & 7, grp4 = lane >> 3;
    const double2 *x2 = reinterpret_cast<const double2 *>(x);
    int k = 0;
    for (int t = blockIdx.x; t < n_tile; t += gridDim.x, k++) {
        const int s = k % S;
        ring::mbar_wait(&full[s], (k / S) & 1);
        const unsigned char *st = smem + s * Cfg::STAGE_BYTES;
        const longlong2 h0 = *reinterpret_cast<const longlong2 *>(st + Cfg::HDR_OFF);
        const longlong2 h1 = *reinterpret_cast<const longlong2 *>(st + Cfg::HDR_OFF + 16);
        const int64_t b0 = h0.x;
        const int a0 = (int)(h1.x & 0xffffffffll), a1 = (int)(h1.x >> 32);
        const bool masked_tile = MASKED && (h1.y & 1);
        const double2 *vs = reinterpret_cast<const double2 *>(st + Cfg::VAL_OFF);
        const int32_t *is = reinterpret_cast<const int32_t *>(st + Cfg::IDX_OFF) + (int)(b0 & 3);
        const int64_t *ps = reinterpret_cast<const int64_t *>(st + Cfg::PTR_OFF) + (a0 & 1);
        for (int rbase = a0 + warp * 4; rbase < a1; rbase += CW * 4) {
            const int row = rbase + grp4;
            bool valid = row < a1;
            if (MASKED && masked_tile && valid && row_skip[row]) valid = false;   // left to the caller (rows that gather ghosts)
            int off = 0, deg = 0;
            if (valid) {
                const int64_t r0 = ps[row - a0];
                off = (int)(r0 - b0);
                deg = (int)(ps[row - a0 + 1] - r0);
            }
            double y0 = 0.0, y1 = 0.0;
            for (int kb = sub; kb < deg; kb += 32) {
                double2 v0[4], v1[4], xv[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int kk = kb + 8 * u;
                    if (kk < deg) {
                        const int c = is[off + kk];
                        v0[u] = vs[2 * off + kk];
                        v1[u] = vs[2 * off + deg + kk];
                        xv[u] = __ldg(x2 + c);
                    } else {
                        v0[u] = v1[u] = xv[u] = make_double2(0.0, 0.0);
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    y0 = fma(v0[u].x, xv[u].x, y0); y0 = fma(v0[u].y, xv[u].y, y0);
                    y1 = fma(v1[u].x, xv[u].x, y1); y1 = fma(v1[u].y, xv[u].y, y1);
                }
            }
#pragma unroll
            for (int d = 4; d; d >>= 1) {
                y0 += __shfl_xor_sync(0xffffffffu, y0, d);
                y1 += __shfl_xor_sync(0xffffffffu, y1, d);
            }
            if (valid && sub == 0) {
                if (fixed) {
                    const uchar2 fx = *reinterpret_cast<const uchar2 *>(fixed + 2 * (int64_t)row);
                    if (fx.x) y0 = 0.0;
                    if (fx.y) y1 = 0.0;
                }
                *reinterpret_cast<double2 *>(y + 2 * (int64_t)row) = make_double2(y0, y1);
                const double2 xa = __ldg(x2 + row);
                dot = fma(xa.x, y0, dot); dot = fma(xa.y, y1, dot);
            }
        }
        __syncwarp();
        if (lane == 0) ring::mbar_arrive(&empty[s]);
    }
    if (TRIG_LATE) ring::grid_dep_launch();
    return true;
}

}  // namespace mofem
