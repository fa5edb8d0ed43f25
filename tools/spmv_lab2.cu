// spmv_lab2.cu -- SpMV variants timed INSIDE a PCG-shaped loop (development tool, not product).
//   spmv_lab2 <dir>   reads node_ptr.bin (int64[n_node+1]), node_idx.bin (int32[U]), data.bin (f64[4U]), x.bin (tools/dump_matrix.py)
// Every iteration launches  q = K p  (variant under test), then an update-shaped and a direction-shaped vector kernel, as
// the solver does; the SpMV is timed by events around it (like mofem_set_profile) and the loop as a whole by wall events.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <cmath>
#include <string>
#include <vector>

#include "../mofem_b200/csrc/spmv_ring.cuh"

using namespace mofem;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <typename T> static std::vector<T> rd(const std::string &p)
{
    FILE *f = fopen(p.c_str(), "rb");
    if (!f) { printf("cannot open %s\n", p.c_str()); exit(1); }
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<T> v(n / sizeof(T));
    if (fread(v.data(), 1, n, f) != (size_t)n) exit(1);
    fclose(f);
    return v;
}

__device__ __forceinline__ void l2_prefetch(const void *p, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// ---- production row kernel (pcg.cu k_spmv<DOT=false>) + optional L2 bulk prefetch of the rows of the NEXT stride ------
template <int PF, bool PDL>
__global__ void __launch_bounds__(128, PF ? 10 : 12)
k_rows(int64_t n_node, const int64_t *__restrict__ node_ptr, const int32_t *__restrict__ node_idx,
       const double *__restrict__ data, const double *__restrict__ x, double *__restrict__ y)
{
    constexpr int LANES = 8, UNROLL = 4, BLOCK = 128;
    const int sub = threadIdx.x % LANES;
    const int64_t rows_per_block = BLOCK / LANES, stride = (int64_t)gridDim.x * rows_per_block;
    int64_t row0 = (int64_t)blockIdx.x * rows_per_block;
    // prefetch pipeline: (pb0, pb1) = sub-block range of the rows PF strides ahead
    if (PF && threadIdx.x == 0) {
        for (int d = 0; d < PF; d++) {
            const int64_t r = row0 + d * stride;
            if (r < n_node) {
                const int64_t e = r + rows_per_block < n_node ? r + rows_per_block : n_node;
                const int64_t b0 = node_ptr[r], b1 = node_ptr[e];
                if (b1 > b0) {
                    l2_prefetch(data + 4 * b0, (uint32_t)(b1 - b0) * 32u);
                    const int64_t i0 = b0 & ~(int64_t)3;
                    l2_prefetch(node_idx + i0, ((uint32_t)(b1 - i0) * 4u + 15u) & ~15u);
                }
            }
        }
    }
    if (PDL) { ring::grid_dep_wait(); ring::grid_dep_launch(); }
    for (; row0 < n_node; row0 += stride) {
        if (PF && threadIdx.x == 0) {
            const int64_t r = row0 + PF * stride;
            if (r < n_node) {
                const int64_t e = r + rows_per_block < n_node ? r + rows_per_block : n_node;
                const int64_t b0 = node_ptr[r], b1 = node_ptr[e];
                if (b1 > b0) {
                    l2_prefetch(data + 4 * b0, (uint32_t)(b1 - b0) * 32u);
                    const int64_t i0 = b0 & ~(int64_t)3;
                    l2_prefetch(node_idx + i0, ((uint32_t)(b1 - i0) * 4u + 15u) & ~15u);
                }
            }
        }
        const int64_t a = row0 + threadIdx.x / LANES;
        double y0 = 0.0, y1 = 0.0;
        if (a < n_node) {
            const int64_t r0 = node_ptr[a];
            const int deg = (int)(node_ptr[a + 1] - r0);
            const double2 *d0 = reinterpret_cast<const double2 *>(data + 4 * r0);
            const double2 *d1 = reinterpret_cast<const double2 *>(data + 4 * r0 + 2 * deg);
            for (int kb = sub; kb < deg; kb += LANES * UNROLL) {
                double2 v0[UNROLL], v1[UNROLL], xv[UNROLL];
#pragma unroll
                for (int u = 0; u < UNROLL; u++) {
                    const int kk = kb + u * LANES;
                    if (kk < deg) {
                        const int c = __ldg(node_idx + r0 + kk);
                        v0[u] = __ldcs(d0 + kk); v1[u] = __ldcs(d1 + kk);
                        xv[u] = __ldg(reinterpret_cast<const double2 *>(x + 2 * (int64_t)c));
                    } else { v0[u] = v1[u] = xv[u] = make_double2(0.0, 0.0); }
                }
#pragma unroll
                for (int u = 0; u < UNROLL; u++) {
                    y0 = fma(v0[u].x, xv[u].x, y0); y0 = fma(v0[u].y, xv[u].y, y0);
                    y1 = fma(v1[u].x, xv[u].x, y1); y1 = fma(v1[u].y, xv[u].y, y1);
                }
            }
        }
#pragma unroll
        for (int d = LANES / 2; d; d >>= 1) {
            y0 += __shfl_xor_sync(0xffffffffu, y0, d);
            y1 += __shfl_xor_sync(0xffffffffu, y1, d);
        }
        if (a < n_node && sub == 0) *reinterpret_cast<double2 *>(y + 2 * a) = make_double2(y0, y1);
    }
}

template <class Cfg, bool PDL, int MINB, bool LATE = false>
__global__ void __launch_bounds__(Cfg::THREADS, MINB)
k_ring(int n_tile, const RingTile *__restrict__ tiles, const int64_t *__restrict__ node_ptr,
       const int32_t *__restrict__ node_idx, const double *__restrict__ data, const double *__restrict__ x,
       double *__restrict__ y, double *__restrict__ part)
{
    extern __shared__ __align__(128) unsigned char smem[];
    double dot = 0.0;
    ring_spmv_body<Cfg, PDL, LATE>(smem, n_tile, tiles, node_ptr, node_idx, data, x, y, nullptr, dot);
    if (part && dot != 0.0 && (threadIdx.x & 255) == 0) part[blockIdx.x] = dot;   // keep the dot alive
}

// update-shaped: x += a p ; r -= a q (reads p q dinv x r, writes x r)    direction-shaped: p = dinv r + b p
template <bool PDL>
__global__ void __launch_bounds__(256)
k_update(int64_t n2, const double2 *__restrict__ p, const double2 *__restrict__ q, const double2 *__restrict__ dinv,
         double2 *__restrict__ x, double2 *__restrict__ r, double alpha)
{
    if (PDL) { ring::grid_dep_wait(); ring::grid_dep_launch(); }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
        const double2 pi = p[i], qi = q[i];
        double2 xi = x[i], ri = r[i];
        xi.x = fma(alpha, pi.x, xi.x); xi.y = fma(alpha, pi.y, xi.y);
        ri.x = fma(-alpha, qi.x, ri.x); ri.y = fma(-alpha, qi.y, ri.y);
        x[i] = xi; r[i] = ri;
    }
}
template <bool PDL>
__global__ void __launch_bounds__(256)
k_direction(int64_t n2, const double2 *__restrict__ r, const double2 *__restrict__ dinv, double2 *__restrict__ p, double beta)
{
    if (PDL) { ring::grid_dep_wait(); ring::grid_dep_launch(); }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
        const double2 ri = r[i], di = dinv[i];
        double2 pi = p[i];
        pi.x = fma(beta, pi.x, di.x * ri.x); pi.y = fma(beta, pi.y, di.y * ri.y);
        p[i] = pi;
    }
}

template <typename... KArgs, typename... Args>
static void launch(bool pdl, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    CK(cudaLaunchKernelEx(&cfg, kern, KArgs(args)...));
}

int main(int argc, char **argv)
{
    std::string dir = argc > 1 ? argv[1] : "/tmp/spmv_lab";
    auto node_ptr = rd<int64_t>(dir + "/node_ptr.bin");
    auto node_idx = rd<int32_t>(dir + "/node_idx.bin");
    auto data = rd<double>(dir + "/data.bin");
    auto xh = rd<double>(dir + "/x.bin");
    const int64_t n_node = (int64_t)node_ptr.size() - 1, U = node_idx.size(), n2 = n_node;
    int maxdeg = 0;
    for (int64_t a = 0; a < n_node; a++) maxdeg = std::max<int>(maxdeg, (int)(node_ptr[a + 1] - node_ptr[a]));
    printf("n_node %ld blocks %ld nnz %ld maxdeg %d\n", (long)n_node, (long)U, (long)(4 * U), maxdeg);
    int64_t *d_ptr; int32_t *d_idx; double *d_data, *d_p, *d_q, *d_x, *d_r, *d_dinv, *d_part;
    CK(cudaMalloc(&d_ptr, node_ptr.size() * 8 + 64)); CK(cudaMalloc(&d_idx, U * 4 + 64)); CK(cudaMalloc(&d_data, U * 32));
    for (double **v : {&d_p, &d_q, &d_x, &d_r, &d_dinv}) { CK(cudaMalloc(v, n_node * 16)); CK(cudaMemset(*v, 0, n_node * 16)); }
    CK(cudaMalloc(&d_part, 8 * 4096));
    CK(cudaMemcpy(d_ptr, node_ptr.data(), node_ptr.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_idx, node_idx.data(), U * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_data, data.data(), U * 32, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_p, xh.data(), n_node * 16, cudaMemcpyHostToDevice));
    const double bytes = 36.0 * U + 8.0 * (n_node + 1) + 32.0 * n_node;
    cudaStream_t s; CK(cudaStreamCreate(&s));
    const int ITER = 200;
    std::vector<cudaEvent_t> ev(2 * ITER + 2);
    for (size_t i = 0; i < ev.size(); i++) CK(cudaEventCreate(&ev[i]));
    std::vector<double> yref, yv(2 * n_node);
    const int gv = 148 * 8;

    auto tiles_for = [&](int span) {
        std::vector<RingTile> t;
        const int64_t key_end = n_node + U;
        int64_t a = 0;
        for (int64_t k0 = 0; k0 < key_end; k0 += span) {
            RingTile rt{};
            rt.a0 = (int32_t)a; rt.b0 = node_ptr[a];
            while (a < n_node && a + node_ptr[a] < k0 + span) a++;
            rt.a1 = (int32_t)a; rt.b1 = node_ptr[a];
            if (rt.a1 > rt.a0) t.push_back(rt);
        }
        return t;
    };

    // one configuration: `spmv(pdl)` launches the variant; returns nothing, prints the timings
    auto bench = [&](const char *name, bool pdl, auto spmv) {
        // correctness of q = K p first (p = x.bin, untouched: alpha = beta = 0 keeps the vectors)
        CK(cudaMemcpyAsync(d_p, xh.data(), n_node * 16, cudaMemcpyHostToDevice, s));
        CK(cudaMemsetAsync(d_q, 0, n_node * 16, s));
        spmv(false);
        CK(cudaStreamSynchronize(s));
        CK(cudaMemcpy(yv.data(), d_q, n_node * 16, cudaMemcpyDeviceToHost));
        if (yref.empty()) yref = yv;
        double err = 0, nrm = 0;
        for (int64_t i = 0; i < 2 * n_node; i++) { err = fmax(err, fabs(yv[i] - yref[i])); nrm = fmax(nrm, fabs(yref[i])); }
        // (a) per-SpMV events inside the loop (events between kernels: PDL cannot overlap here)
        float spmv_ms = 0.f, loop_ms = 0.f;
        for (int pass = 0; pass < 2; pass++) {
            for (int it = 0; it < ITER; it++) {
                CK(cudaEventRecord(ev[2 * it], s));
                spmv(false);
                CK(cudaEventRecord(ev[2 * it + 1], s));
                launch(false, k_update<false>, dim3(gv), dim3(256), 0, s, n2, (const double2 *)d_p, (const double2 *)d_q, (const double2 *)d_dinv, (double2 *)d_x, (double2 *)d_r, 0.0);
                launch(false, k_direction<false>, dim3(gv), dim3(256), 0, s, n2, (const double2 *)d_r, (const double2 *)d_dinv, (double2 *)d_p, 1.0);
            }
            CK(cudaStreamSynchronize(s));
        }
        for (int it = 0; it < ITER; it++) { float ms; cudaEventElapsedTime(&ms, ev[2 * it], ev[2 * it + 1]); spmv_ms += ms; }
        // (b) the loop as a whole, no events inside, with / without PDL
        for (int pass = 0; pass < 2; pass++) {
            CK(cudaEventRecord(ev[2 * ITER], s));
            for (int it = 0; it < ITER; it++) {
                spmv(pdl);
                if (pdl) {
                    launch(true, k_update<true>, dim3(gv), dim3(256), 0, s, n2, (const double2 *)d_p, (const double2 *)d_q, (const double2 *)d_dinv, (double2 *)d_x, (double2 *)d_r, 0.0);
                    launch(true, k_direction<true>, dim3(gv), dim3(256), 0, s, n2, (const double2 *)d_r, (const double2 *)d_dinv, (double2 *)d_p, 1.0);
                } else {
                    launch(false, k_update<false>, dim3(gv), dim3(256), 0, s, n2, (const double2 *)d_p, (const double2 *)d_q, (const double2 *)d_dinv, (double2 *)d_x, (double2 *)d_r, 0.0);
                    launch(false, k_direction<false>, dim3(gv), dim3(256), 0, s, n2, (const double2 *)d_r, (const double2 *)d_dinv, (double2 *)d_p, 1.0);
                }
            }
            CK(cudaEventRecord(ev[2 * ITER + 1], s));
            CK(cudaStreamSynchronize(s));
            cudaEventElapsedTime(&loop_ms, ev[2 * ITER], ev[2 * ITER + 1]);
        }
        printf("%-46s spmv %7.2f us %7.1f GB/s | iteration %7.2f us | relerr %.1e\n", name, spmv_ms / ITER * 1e3,
               bytes / (spmv_ms / ITER * 1e-3) / 1e9, loop_ms / ITER * 1e3, err / nrm);
        fflush(stdout);
    };

    const int64_t g_rows = std::min<int64_t>((n_node + 15) / 16, 148 * 24);
#define ROWS(PF)                                                                                                          \
    bench("rows pf=" #PF, false, [&](bool) { launch(false, k_rows<PF, false>, dim3((unsigned)g_rows), dim3(128), 0, s, n_node, (const int64_t *)d_ptr, (const int32_t *)d_idx, (const double *)d_data, (const double *)d_p, d_q); }); \
    bench("rows pf=" #PF " +pdl", true, [&](bool pdl) {                                                                \
        if (pdl) launch(true, k_rows<PF, true>, dim3((unsigned)g_rows), dim3(128), 0, s, n_node, (const int64_t *)d_ptr, (const int32_t *)d_idx, (const double *)d_data, (const double *)d_p, d_q); \
        else launch(false, k_rows<PF, false>, dim3((unsigned)g_rows), dim3(128), 0, s, n_node, (const int64_t *)d_ptr, (const int32_t *)d_idx, (const double *)d_data, (const double *)d_p, d_q); })
    ROWS(0);
    ROWS(1);

#define RING(SPAN, STAGES, CW, MINB, GRIDMUL, LATE)                                                                             \
    do {                                                                                                                  \
        using Cfg = RingCfg<SPAN, 64, STAGES, CW>;                                                                         \
        if (maxdeg > 64) { printf("ring: maxdeg %d > 64, skipped\n", maxdeg); break; }                                      \
        auto th = tiles_for(SPAN);                                                                                        \
        RingTile *d_tiles; CK(cudaMalloc(&d_tiles, th.size() * sizeof(RingTile)));                                         \
        CK(cudaMemcpy(d_tiles, th.data(), th.size() * sizeof(RingTile), cudaMemcpyHostToDevice));                          \
        CK(cudaFuncSetAttribute(k_ring<Cfg, false, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));  \
        CK(cudaFuncSetAttribute(k_ring<Cfg, true, MINB, LATE>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));   \
        const int nt = (int)th.size();                                                                                    \
        const int grid = std::min(nt, 148 * GRIDMUL);                                                                     \
        char nm[128];                                                                                                     \
        snprintf(nm, sizeof nm, "ring span=%d st=%d cw=%d minb=%d grid=148x%d smem=%dK late=%d", SPAN, STAGES, CW, MINB, GRIDMUL, Cfg::SMEM_BYTES / 1024, (int)LATE); \
        bench(nm, false, [&](bool) { launch(false, k_ring<Cfg, false, MINB>, dim3(grid), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, s, nt, (const RingTile *)d_tiles, (const int64_t *)d_ptr, (const int32_t *)d_idx, (const double *)d_data, (const double *)d_p, d_q, d_part); }); \
        std::string nm2 = std::string(nm) + " +pdl";                                                                      \
        bench(nm2.c_str(), true, [&](bool pdl) {                                                                          \
            if (pdl) launch(true, k_ring<Cfg, true, MINB, LATE>, dim3(grid), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, s, nt, (const RingTile *)d_tiles, (const int64_t *)d_ptr, (const int32_t *)d_idx, (const double *)d_data, (const double *)d_p, d_q, d_part); \
            else launch(false, k_ring<Cfg, false, MINB>, dim3(grid), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, s, nt, (const RingTile *)d_tiles, (const int64_t *)d_ptr, (const int32_t *)d_idx, (const double *)d_data, (const double *)d_p, d_q, d_part); }); \
        CK(cudaFree(d_tiles));                                                                                            \
    } while (0)
    RING(960, 2, 16, 2, 2, false);
    RING(960, 2, 16, 2, 2, true);
    RING(960, 2, 12, 2, 2, true);
    RING(960, 2, 14, 2, 2, true);
    RING(960, 2, 20, 2, 2, true);
    RING(704, 2, 10, 3, 3, true);
    RING(704, 2, 12, 3, 3, true);
    RING(704, 3, 12, 2, 2, true);
    RING(576, 2, 8, 3, 3, true);
    RING(576, 2, 8, 4, 4, true);
    RING(576, 2, 10, 3, 3, true);
    RING(576, 3, 10, 2, 2, true);
    RING(448, 2, 8, 4, 4, true);
    RING(448, 3, 8, 3, 3, true);
    RING(1216, 2, 20, 1, 1, true);
    RING(1216, 2, 16, 1, 1, true);
    RING(1440, 2, 24, 1, 1, true);
    return 0;
}
