// vec_lab.cu -- access-pattern study for the PCG vector kernels (development tool, not product).
//   vec_lab            times update-shaped kernels (reads p q dinv x r, writes x r; 56 B per scalar pair... 112 B per double2 element)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ void upd(int64_t i, double alpha, const double2 *p, const double2 *q, const double2 *dinv, double2 *x, double2 *r, double &acc)
{
    const double2 pi = p[i], qi = q[i], di = dinv[i];
    double2 xi = x[i], ri = r[i];
    xi.x = fma(alpha, pi.x, xi.x); xi.y = fma(alpha, pi.y, xi.y);
    ri.x = fma(-alpha, qi.x, ri.x); ri.y = fma(-alpha, qi.y, ri.y);
    x[i] = xi; r[i] = ri;
    acc = fma(ri.x, di.x * ri.x, acc); acc = fma(ri.y, di.y * ri.y, acc);
}

// grid-stride
template <int MINB>
__global__ void __launch_bounds__(256, MINB)
k_stride(int64_t n2, double alpha, const double2 *__restrict__ p, const double2 *__restrict__ q, const double2 *__restrict__ dinv,
         double2 *__restrict__ x, double2 *__restrict__ r, double *out)
{
    double acc = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) upd(i, alpha, p, q, dinv, x, r, acc);
    if (acc == 12345.678) out[0] = acc;
}
// every CTA owns one contiguous chunk
template <int MINB, int UN>
__global__ void __launch_bounds__(256, MINB)
k_chunk(int64_t n2, double alpha, const double2 *__restrict__ p, const double2 *__restrict__ q, const double2 *__restrict__ dinv,
        double2 *__restrict__ x, double2 *__restrict__ r, double *out)
{
    const int64_t per = (n2 + gridDim.x - 1) / gridDim.x;
    const int64_t b0 = (int64_t)blockIdx.x * per, b1 = b0 + per < n2 ? b0 + per : n2;
    double acc = 0;
    for (int64_t base = b0 + threadIdx.x; base < b1; base += (int64_t)UN * blockDim.x) {
#pragma unroll
        for (int u = 0; u < UN; u++) { const int64_t i = base + (int64_t)u * blockDim.x; if (i < b1) upd(i, alpha, p, q, dinv, x, r, acc); }
    }
    if (acc == 12345.678) out[0] = acc;
}
// plain copy and read-only sum for reference
__global__ void __launch_bounds__(256) k_copy(int64_t n2, const double2 *__restrict__ a, double2 *__restrict__ b)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) b[i] = a[i];
}
__global__ void __launch_bounds__(256) k_read5(int64_t n2, const double2 *__restrict__ a, const double2 *__restrict__ b, const double2 *__restrict__ c,
                                               const double2 *__restrict__ d, const double2 *__restrict__ e, double *out)
{
    double acc = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x)
        acc += a[i].x + b[i].y + c[i].x + d[i].y + e[i].x;
    if (acc == 12345.678) out[0] = acc;
}

int main()
{
    for (int64_t n2 : {(int64_t)690229, (int64_t)6764161}) {
        const int64_t npad = (2 * n2 + 31) & ~31;
        double *slab; CK(cudaMalloc(&slab, sizeof(double) * npad * 6));
        CK(cudaMemset(slab, 0, sizeof(double) * npad * 6));
        double2 *p = (double2 *)slab, *q = (double2 *)(slab + npad), *x = (double2 *)(slab + 2 * npad), *r = (double2 *)(slab + 3 * npad),
                *dinv = (double2 *)(slab + 4 * npad);
        double *out = slab + 5 * npad;
        char *flush; CK(cudaMalloc(&flush, 512u << 20));
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        auto run = [&](const char *name, double bytes, auto launch) {
            float tot = 0; const int reps = 10;
            for (int i = 0; i < reps + 2; i++) {
                CK(cudaMemsetAsync(flush, i, 512u << 20));
                cudaEventRecord(e0); launch(); cudaEventRecord(e1);
                CK(cudaDeviceSynchronize());
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (i >= 2) tot += ms;
            }
            printf("n2=%8ld %-34s %8.2f us %7.1f GB/s\n", (long)n2, name, tot / reps * 1e3, bytes / (tot / reps * 1e-3) / 1e9);
        };
        const double bu = 112.0 * n2;
        run("copy (1r 1w)", 32.0 * n2, [&] { k_copy<<<148 * 8, 256>>>(n2, p, q); });
        run("read5", 80.0 * n2, [&] { k_read5<<<148 * 8, 256>>>(n2, p, q, x, r, dinv, out); });
        run("update stride 8x256/SM", bu, [&] { k_stride<8><<<148 * 8, 256>>>(n2, 0.5, p, q, dinv, x, r, out); });
        run("update stride 4x256/SM", bu, [&] { k_stride<4><<<148 * 4, 256>>>(n2, 0.5, p, q, dinv, x, r, out); });
        run("update stride 16x256/SM(2 waves)", bu, [&] { k_stride<8><<<148 * 16, 256>>>(n2, 0.5, p, q, dinv, x, r, out); });
        run("update chunk 8/SM un1", bu, [&] { k_chunk<8, 1><<<148 * 8, 256>>>(n2, 0.5, p, q, dinv, x, r, out); });
        run("update chunk 8/SM un2", bu, [&] { k_chunk<8, 2><<<148 * 8, 256>>>(n2, 0.5, p, q, dinv, x, r, out); });
        run("update chunk 4/SM un4", bu, [&] { k_chunk<4, 4><<<148 * 4, 256>>>(n2, 0.5, p, q, dinv, x, r, out); });
        run("update chunk 4/SM un2", bu, [&] { k_chunk<4, 2><<<148 * 4, 256>>>(n2, 0.5, p, q, dinv, x, r, out); });
        run("update chunk 32/SM(4 waves) un1", bu, [&] { k_chunk<8, 1><<<148 * 32, 256>>>(n2, 0.5, p, q, dinv, x, r, out); });
        CK(cudaFree(slab)); CK(cudaFree(flush));
    }
    return 0;
}
