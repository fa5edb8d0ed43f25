// scan.cu -- exclusive prefix sums (int64) used by the task builder and the CSR symbolic phase.
// Three-phase scan: per-tile scan + tile sums, recursive scan of the tile sums, add back.
#include "common.cuh"

namespace mofem {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;  // per thread
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <typename TIn>
__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_tiles(const TIn *__restrict__ in, int64_t *__restrict__ out, int64_t n, int64_t *__restrict__ tile_sum)
{
    __shared__ int64_t warp_sum[SCAN_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int64_t v[SCAN_ITEMS];
    int64_t run = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        int64_t x = (base + i < n) ? (int64_t)in[base + i] : 0;
        v[i] = run;
        run += x;
    }
    // exclusive scan of `run` across the block
    int64_t incl = run;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int64_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) warp_sum[w] = incl;
    __syncthreads();
    if (w == 0) {
        int64_t s = (lane < SCAN_THREADS / 32) ? warp_sum[lane] : 0;
#pragma unroll
        for (int d = 1; d < SCAN_THREADS / 32; d <<= 1) {
            int64_t t = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= d) s += t;
        }
        if (lane < SCAN_THREADS / 32) warp_sum[lane] = s;  // inclusive over warps
    }
    __syncthreads();
    int64_t offs = (incl - run) + (w ? warp_sum[w - 1] : 0);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++)
        if (base + i < n) out[base + i] = v[i] + offs;
    if (threadIdx.x == SCAN_THREADS - 1 && tile_sum) tile_sum[blockIdx.x] = offs + run;
}

__global__ void k_scan_add(int64_t *__restrict__ out, int64_t n, const int64_t *__restrict__ tile_off)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] += tile_off[i / SCAN_TILE];
}

__global__ void k_scan_total(const int64_t *tile_off_last, const int64_t *tile_sum_last, int64_t *total)
{
    *total = *tile_off_last + *tile_sum_last;
}

template <typename TIn>
static cudaError_t scan_impl(const TIn *in, int64_t *out, int64_t n, int64_t *total_dev, cudaStream_t s,
                             int64_t *launches)
{
    if (n <= 0) {
        if (total_dev) return cudaMemsetAsync(total_dev, 0, sizeof(int64_t), s);
        return cudaSuccess;
    }
    const int64_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    int64_t *tile_sum = nullptr;
    cudaError_t e = cudaMallocAsync(&tile_sum, sizeof(int64_t) * (2 * tiles + 1), s);
    if (e != cudaSuccess) return e;
    int64_t *tile_off = tile_sum + tiles;
    k_scan_tiles<TIn><<<(unsigned)tiles, SCAN_THREADS, 0, s>>>(in, out, n, tile_sum);
    if (launches) ++*launches;
    if (tiles > 1) {
        e = scan_impl<int64_t>(tile_sum, tile_off, tiles, nullptr, s, launches);
        if (e != cudaSuccess) return e;
        k_scan_add<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(out, n, tile_off);
        if (launches) ++*launches;
    } else {
        cudaMemsetAsync(tile_off, 0, sizeof(int64_t), s);
    }
    if (total_dev) {
        k_scan_total<<<1, 1, 0, s>>>(tile_off + tiles - 1, tile_sum + tiles - 1, total_dev);
        if (launches) ++*launches;
    }
    e = cudaGetLastError();
    cudaFreeAsync(tile_sum, s);
    return e;
}

cudaError_t exclusive_scan_i64(const int64_t *in, int64_t *out, int64_t n, int64_t *total_dev, cudaStream_t s,
                               int64_t *launches)
{
    return scan_impl<int64_t>(in, out, n, total_dev, s, launches);
}

cudaError_t exclusive_scan_i32_to_i64(const int32_t *in, int64_t *out, int64_t n, int64_t *total_dev, cudaStream_t s,
                                      int64_t *launches)
{
    return scan_impl<int32_t>(in, out, n, total_dev, s, launches);
}

}  // namespace mofem
