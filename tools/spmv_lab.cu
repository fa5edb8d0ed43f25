// spmv_lab.cu -- micro-benchmark of block-CSR SpMV variants on the bench matrix (development tool, not product).
//   spmv_lab <dir>   reads node_ptr.bin (int64[n_node+1]), node_idx.bin (int32[U]), data.bin (f64[4U]), x.bin
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string>
#include <vector>
#include <cmath>

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

// ---- variant: LANES lanes per node row, optional streaming loads, unroll --------------------------------
template <int LANES, int BLOCK, int MINB, bool STREAM, int UNROLL>
__global__ void __launch_bounds__(BLOCK, MINB)
k_rows(int64_t n_node, const int64_t *__restrict__ node_ptr, const int32_t *__restrict__ node_idx,
       const double *__restrict__ data, const double *__restrict__ x, double *__restrict__ y)
{
    const int sub = threadIdx.x % LANES;
    const int64_t rows_per_block = BLOCK / LANES;
    for (int64_t row0 = (int64_t)blockIdx.x * rows_per_block; row0 < n_node; row0 += (int64_t)gridDim.x * rows_per_block) {
        const int64_t a = row0 + threadIdx.x / LANES;
        double y0 = 0.0, y1 = 0.0;
        if (a < n_node) {
            const int64_t r0 = node_ptr[a], deg = node_ptr[a + 1] - r0;
            const double2 *d0 = reinterpret_cast<const double2 *>(data + 4 * r0);
            const double2 *d1 = reinterpret_cast<const double2 *>(data + 4 * r0 + 2 * deg);
            for (int64_t kb = sub; kb < deg; kb += LANES * UNROLL) {
                double2 v0[UNROLL], v1[UNROLL], xv[UNROLL];
#pragma unroll
                for (int u = 0; u < UNROLL; u++) {
                    const int64_t kk = kb + u * LANES;
                    if (kk < deg) {
                        const int c = __ldg(node_idx + r0 + kk);
                        if (STREAM) { v0[u] = __ldcs(d0 + kk); v1[u] = __ldcs(d1 + kk); }
                        else { v0[u] = d0[kk]; v1[u] = d1[kk]; }
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

// ---- variant: as k_rows, node_ptr of the NEXT row prefetched while the current row is processed (software pipelining) --
template <int LANES, int BLOCK, int MINB, int UNROLL, bool IDXPF>
__global__ void __launch_bounds__(BLOCK, MINB)
k_rows_pipe(int64_t n_node, const int64_t *__restrict__ node_ptr, const int32_t *__restrict__ node_idx,
            const double *__restrict__ data, const double *__restrict__ x, double *__restrict__ y)
{
    const int sub = threadIdx.x % LANES;
    const int64_t rows_per_block = BLOCK / LANES;
    const int64_t stride = (int64_t)gridDim.x * rows_per_block;
    int64_t a = (int64_t)blockIdx.x * rows_per_block + threadIdx.x / LANES;
    int64_t r0 = 0, r1 = 0;
    if (a < n_node) { r0 = node_ptr[a]; r1 = node_ptr[a + 1]; }
    int cpf = 0;
    if (IDXPF && a < n_node && sub < r1 - r0) cpf = __ldg(node_idx + r0 + sub);
    for (int64_t row0 = (int64_t)blockIdx.x * rows_per_block; row0 < n_node; row0 += stride) {
        const int64_t an = a + stride;
        int64_t r0n = 0, r1n = 0;
        if (an < n_node) { r0n = node_ptr[an]; r1n = node_ptr[an + 1]; }
        double y0 = 0.0, y1 = 0.0;
        if (a < n_node) {
            const int deg = (int)(r1 - r0);
            const double2 *d0 = reinterpret_cast<const double2 *>(data + 4 * r0);
            const double2 *d1 = reinterpret_cast<const double2 *>(data + 4 * r0 + 2 * deg);
            for (int kb = sub; kb < deg; kb += LANES * UNROLL) {
                double2 v0[UNROLL], v1[UNROLL], xv[UNROLL];
#pragma unroll
                for (int u = 0; u < UNROLL; u++) {
                    const int kk = kb + u * LANES;
                    if (kk < deg) {
                        const int c = (IDXPF && u == 0 && kb == sub) ? cpf : __ldg(node_idx + r0 + kk);
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
        if (IDXPF && an < n_node && sub < r1n - r0n) cpf = __ldg(node_idx + r0n + sub);
#pragma unroll
        for (int d = LANES / 2; d; d >>= 1) {
            y0 += __shfl_xor_sync(0xffffffffu, y0, d);
            y1 += __shfl_xor_sync(0xffffffffu, y1, d);
        }
        if (a < n_node && sub == 0) *reinterpret_cast<double2 *>(y + 2 * a) = make_double2(y0, y1);
        a = an; r0 = r0n; r1 = r1n;
    }
}

// ---- variant: values staged through shared memory with 1-D bulk async copies (TMA), thread per block multiply,
//      thread per row fixed-order sum ------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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

// tile_row[t] .. tile_row[t+1]: node rows of tile t; every tile has <= TB blocks
template <int TB, int BLOCK, int STAGES>
__global__ void __launch_bounds__(BLOCK)
k_tma(int n_tile, const int32_t *__restrict__ tile_row, const int64_t *__restrict__ node_ptr,
      const int32_t *__restrict__ node_idx, const double *__restrict__ data, const double *__restrict__ x,
      double *__restrict__ y)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *vals = reinterpret_cast<double *>(smem_raw);                       // [STAGES][TB*4]
    double2 *part = reinterpret_cast<double2 *>(vals + (size_t)STAGES * TB * 4);  // [TB]
    int *rowptr = reinterpret_cast<int *>(part + TB);                          // [TB+1] local row starts
    __shared__ uint64_t bar[STAGES];
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) mbar_init(&bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // prologue: issue the first STAGES-1 tiles of this CTA
    int issued = 0;
    auto issue = [&](int t, int s) {
        const int64_t b0 = node_ptr[tile_row[t]], b1 = node_ptr[tile_row[t + 1]];
        const uint32_t bytes = (uint32_t)(b1 - b0) * 32u;
        mbar_expect_tx(&bar[s], bytes);
        if (bytes) bulk_g2s(vals + (size_t)s * TB * 4, data + 4 * b0, bytes, &bar[s]);
    };
    const int first = blockIdx.x, stride = gridDim.x;
    if (threadIdx.x == 0)
        for (int k = 0; k < STAGES - 1; k++) { int t = first + k * stride; if (t < n_tile) issue(t, k); }
    issued = STAGES - 1;
    int it = 0;
    for (int t = first; t < n_tile; t += stride, it++) {
        const int s = it % STAGES;
        if (threadIdx.x == 0) { int tn = first + (it + STAGES - 1) * stride; if (tn < n_tile) issue(tn, (it + STAGES - 1) % STAGES); }
        const int ra = tile_row[t], rb = tile_row[t + 1];
        const int64_t b0 = node_ptr[ra];
        const int nb = (int)(node_ptr[rb] - b0);
        for (int r = threadIdx.x; r <= rb - ra; r += BLOCK) rowptr[r] = (int)(node_ptr[ra + r] - b0);
        __syncthreads();
        mbar_wait(&bar[s], (it / STAGES) & 1);
        const double *v = vals + (size_t)s * TB * 4;
        // multiply: thread per block.  Need the row of each block: binary search in rowptr.
        for (int k = threadIdx.x; k < nb; k += BLOCK) {
            int lo = 0, hi = rb - ra;
            while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (rowptr[mid] <= k) lo = mid; else hi = mid; }
            const int rs = rowptr[lo], deg = rowptr[lo + 1] - rs, kk = k - rs;
            const double2 v0 = *reinterpret_cast<const double2 *>(v + 4 * rs + 2 * kk);
            const double2 v1 = *reinterpret_cast<const double2 *>(v + 4 * rs + 2 * deg + 2 * kk);
            const int c = __ldg(node_idx + b0 + k);
            const double2 xv = __ldg(reinterpret_cast<const double2 *>(x + 2 * (int64_t)c));
            part[k] = make_double2(fma(v0.y, xv.y, v0.x * xv.x), fma(v1.y, xv.y, v1.x * xv.x));
        }
        __syncthreads();
        for (int r = threadIdx.x; r < rb - ra; r += BLOCK) {
            double y0 = 0.0, y1 = 0.0;
            for (int k = rowptr[r]; k < rowptr[r + 1]; k++) { y0 += part[k].x; y1 += part[k].y; }
            *reinterpret_cast<double2 *>(y + 2 * (int64_t)(ra + r)) = make_double2(y0, y1);
        }
        __syncthreads();
    }
    (void)issued;
}

int main(int argc, char **argv)
{
    std::string dir = argc > 1 ? argv[1] : "/tmp/spmv_lab";
    auto node_ptr = rd<int64_t>(dir + "/node_ptr.bin");
    auto node_idx = rd<int32_t>(dir + "/node_idx.bin");
    auto data = rd<double>(dir + "/data.bin");
    auto x = rd<double>(dir + "/x.bin");
    const int64_t n_node = (int64_t)node_ptr.size() - 1, U = node_idx.size();
    printf("n_node %ld blocks %ld nnz %ld\n", (long)n_node, (long)U, (long)(4 * U));
    int64_t *d_ptr; int32_t *d_idx; double *d_data, *d_x, *d_y, *d_ref;
    CK(cudaMalloc(&d_ptr, node_ptr.size() * 8)); CK(cudaMalloc(&d_idx, U * 4)); CK(cudaMalloc(&d_data, U * 32));
    CK(cudaMalloc(&d_x, n_node * 16)); CK(cudaMalloc(&d_y, n_node * 16)); CK(cudaMalloc(&d_ref, n_node * 16));
    CK(cudaMemcpy(d_ptr, node_ptr.data(), node_ptr.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_idx, node_idx.data(), U * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_data, data.data(), U * 32, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_x, x.data(), n_node * 16, cudaMemcpyHostToDevice));
    const double bytes = 36.0 * U + 8.0 * (n_node + 1) + 32.0 * n_node;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    std::vector<double> yref(2 * n_node), yv(2 * n_node);
    bool have_ref = false;
    // cold-L2 timing: a 512 MB memset between launches evicts the index / pointer arrays, as the vector kernels of the PCG
    // loop do between two SpMVs
    char *flush; CK(cudaMalloc(&flush, 512u << 20));
    auto run_cold = [&](const char *name, auto launch) {
        float tot = 0;
        const int reps = 20;
        for (int i = 0; i < reps + 2; i++) {
            CK(cudaMemsetAsync(flush, i, 512u << 20));
            cudaEventRecord(e0);
            launch();
            cudaEventRecord(e1);
            CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (i >= 2) tot += ms;
        }
        printf("COLD %-40s %8.2f us  %7.1f GB/s\n", name, tot / reps * 1e3, bytes / (tot / reps * 1e-3) / 1e9);
    };
    auto run = [&](const char *name, auto launch) {
        CK(cudaMemset(d_y, 0, n_node * 16));
        for (int i = 0; i < 5; i++) launch();
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        const int reps = 50;
        for (int i = 0; i < reps; i++) launch();
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        CK(cudaMemcpy(yv.data(), d_y, n_node * 16, cudaMemcpyDeviceToHost));
        double err = 0, nrm = 0;
        if (!have_ref) { yref = yv; have_ref = true; }
        for (int64_t i = 0; i < 2 * n_node; i++) { err = fmax(err, fabs(yv[i] - yref[i])); nrm = fmax(nrm, fabs(yref[i])); }
        printf("%-44s %8.2f us  %7.1f GB/s  relerr %.1e\n", name, ms / reps * 1e3, bytes / (ms / reps * 1e-3) / 1e9, err / nrm);
    };
#define ROWS(L, B, M, S, UN, G)                                                                                  \
    run("rows lanes=" #L " block=" #B " minb=" #M " stream=" #S " unroll=" #UN " grid=" #G, [&]() {              \
        int64_t g = (n_node + (B / L) - 1) / (B / L); if (G > 0 && g > 148 * G) g = 148 * G;                       \
        k_rows<L, B, M, S, UN><<<(unsigned)g, B>>>(n_node, d_ptr, d_idx, d_data, d_x, d_y); })
#define ROWS(L, B, M, S, UN, G)                                                                                  \
    run("rows lanes=" #L " block=" #B " minb=" #M " stream=" #S " unroll=" #UN " grid=" #G, [&]() {              \
        int64_t g = (n_node + (B / L) - 1) / (B / L); if (G > 0 && g > 148 * G) g = 148 * G;                       \
        k_rows<L, B, M, S, UN><<<(unsigned)g, B>>>(n_node, d_ptr, d_idx, d_data, d_x, d_y); })
#define PIPE(L, B, M, UN, PF, G)                                                                                   \
    run("pipe lanes=" #L " block=" #B " minb=" #M " unroll=" #UN " idxpf=" #PF " grid=" #G, [&]() {               \
        int64_t g = (n_node + (B / L) - 1) / (B / L); if (G > 0 && g > 148 * G) g = 148 * G;                       \
        k_rows_pipe<L, B, M, UN, PF><<<(unsigned)g, B>>>(n_node, d_ptr, d_idx, d_data, d_x, d_y); })
#define COLD_ROWS(L, B, M, S, UN, G)                                                                              \
    run_cold("rows lanes=" #L " block=" #B " minb=" #M " unroll=" #UN " grid=" #G, [&]() {                          \
        int64_t g = (n_node + (B / L) - 1) / (B / L); if (G > 0 && g > 148 * G) g = 148 * G;                       \
        k_rows<L, B, M, S, UN><<<(unsigned)g, B>>>(n_node, d_ptr, d_idx, d_data, d_x, d_y); })
#define COLD_PIPE(L, B, M, UN, PF, G)                                                                             \
    run_cold("pipe lanes=" #L " block=" #B " minb=" #M " unroll=" #UN " idxpf=" #PF " grid=" #G, [&]() {           \
        int64_t g = (n_node + (B / L) - 1) / (B / L); if (G > 0 && g > 148 * G) g = 148 * G;                       \
        k_rows_pipe<L, B, M, UN, PF><<<(unsigned)g, B>>>(n_node, d_ptr, d_idx, d_data, d_x, d_y); })
    COLD_ROWS(8, 128, 12, true, 4, 24);
    COLD_PIPE(8, 128, 12, 4, false, 24);
    COLD_PIPE(8, 128, 12, 4, true, 24);
    COLD_PIPE(8, 128, 10, 4, false, 24);
    COLD_PIPE(8, 128, 10, 4, true, 24);
    COLD_PIPE(8, 256, 6, 4, false, 12);
    COLD_ROWS(8, 128, 12, true, 4, 0);
    COLD_ROWS(16, 128, 12, true, 2, 0);
    COLD_ROWS(8, 128, 12, true, 4, 24);
    PIPE(8, 128, 12, 4, false, 24);
    PIPE(8, 128, 12, 4, true, 24);
    PIPE(8, 128, 10, 4, false, 24);
    PIPE(8, 128, 10, 4, true, 24);
    PIPE(8, 128, 12, 4, false, 48);
    PIPE(8, 128, 12, 3, false, 24);
    PIPE(8, 256, 6, 4, false, 12);
    ROWS(8, 128, 12, true, 4, 24);
    ROWS(8, 128, 12, true, 4, 0);
    // TMA-staged variant: tiles of <= TB blocks
    auto make_tiles = [&](int TB) {
        std::vector<int32_t> tr; tr.push_back(0);
        int64_t a = 0;
        while (a < n_node) {
            int64_t b = a;
            while (b < n_node && node_ptr[b + 1] - node_ptr[a] <= TB) b++;
            if (b == a) { printf("row longer than tile\n"); exit(1); }
            tr.push_back((int32_t)b); a = b;
        }
        return tr;
    };
#define TMA(TBv, Bv, STv, GM)                                                                                    \
    {                                                                                                            \
        auto tr = make_tiles(TBv); int nt = (int)tr.size() - 1; int32_t *d_tr;                                   \
        CK(cudaMalloc(&d_tr, tr.size() * 4)); CK(cudaMemcpy(d_tr, tr.data(), tr.size() * 4, cudaMemcpyHostToDevice)); \
        size_t sm = (size_t)STv * TBv * 32 + (size_t)TBv * 16 + (TBv + 1) * 4 + 64;                             \
        CK(cudaFuncSetAttribute(k_tma<TBv, Bv, STv>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));     \
        run("tma TB=" #TBv " block=" #Bv " stages=" #STv " grid=148*" #GM, [&]() {                             \
            k_tma<TBv, Bv, STv><<<148 * GM, Bv, sm>>>(nt, d_tr, d_ptr, d_idx, d_data, d_x, d_y); });             \
        cudaFree(d_tr);                                                                                          \
    }
    TMA(2048, 512, 2, 1);
    // plain copy of the same number of bytes for reference
    {
        double *a, *b; size_t n = (size_t)(bytes / 16);
        CK(cudaMalloc(&a, n * 8)); CK(cudaMalloc(&b, n * 8));
        for (int i = 0; i < 3; i++) cudaMemcpyAsync(b, a, n * 8, cudaMemcpyDeviceToDevice);
        cudaEventRecord(e0);
        for (int i = 0; i < 20; i++) cudaMemcpyAsync(b, a, n * 8, cudaMemcpyDeviceToDevice);
        cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-44s %8.2f us  %7.1f GB/s (read+write)\n", "cudaMemcpy D2D same bytes", ms / 20 * 1e3, bytes / (ms / 20 * 1e-3) / 1e9);
    }
    return 0;
}
