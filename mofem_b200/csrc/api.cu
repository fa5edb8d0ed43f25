// api.cu -- the C-ABI (include/mofem_b200.h): context, device memory, phase orchestration.
#include <algorithm>
#include <cstring>
#include <map>

#include "common.cuh"

using namespace mofem;

struct mofem_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    // device copies of the problem
    DevProblem P{};
    std::vector<void *> problem_allocs;
    bool have_problem = false, have_symbolic = false, have_matrix = false, have_system = false, have_solution = false;
    // tasks
    DevTasks T{};
    int32_t *cls_tasks = nullptr;
    CqRec *cq_rec = nullptr;   // records of the cell kernel's cells (class FAST_CLASS), sub-bucket order
    int64_t cls_count[N_BUCKET]{}, cls_start[N_BUCKET]{};
    int64_t n_pair = 0, n_coo = 0;
    std::vector<void *> sym_allocs;
    // csr
    DevCsr C{};
    double *data = nullptr, *blk_val = nullptr;
    unsigned long long *stats = nullptr;  // [4]
    int *err_flag = nullptr;
    unsigned long long stats_host[4]{};
    // solve
    PcgBuffers W{};
    uint8_t *fixed = nullptr;
    double *rhs = nullptr, *fixed_val = nullptr, *u = nullptr;
    std::vector<void *> sys_allocs;
    int64_t sys_n = 0;
    double *slab = nullptr;
    size_t slab_hot_bytes = 0, slab_vec_bytes = 0;
    int l2_window = 5;    // leading vectors of the slab (p, r, dinv, q, x) inside the L2 persistence window
    bool slab_in_xbuf = false;
    int l2_persist = 0;   // L2 persistence window over the PCG vectors: neutral at 1 M cells since the matrix stream carries evict-first hints, costs 20 % at 10 M (measured) -> off
    int spmv_ring = 1;    // PCG SpMV streams the matrix through shared memory with the TMA (spmv_ring.cuh); 0: row kernel
    double floor_rtol = 1e-8;   // largest true relative residual mofem_solve accepts as "fp64 floor reached" (converged = 2)
    int pdl = 1;          // programmatic dependent launch between the two kernels of a PCG iteration (single GPU, peer transport)
    int cell_kernel = 1;  // 1: all-quad4 overlay cells on the cell kernel (one evaluation per (point, element); bit-identical to the thread-per-block kernels, 0)
    // timing
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double ms[6]{};
    int64_t launches[4]{};
    // optional per-kernel profile of the PCG loop
    // row-partitioned solve (one process per GPU)
    DistPlan dist;
    bool have_halo = false;
    std::vector<void *> dist_allocs;
    int profile = 0;      // 0 off, 1 events around the kernels of every PCG iteration, N > 1 of every N-th iteration
    std::vector<cudaEvent_t> prof_events;
    double prof_ms[4]{};
};
constexpr int PCG_CHUNK = 250;   // iterations enqueued between host polls (the device stops early on its own)

static int ctx_fail(mofem_ctx *ctx, int code, const std::string &msg)
{
    if (ctx) ctx->err = msg;
    return code;
}

#define CK(call) MOFEM_CUDA_OK(call)

static bool is_device_ptr(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

template <typename T>
static int dev_alloc(mofem_ctx *ctx, T **out, int64_t count, std::vector<void *> &pool)
{
    *out = nullptr;
    if (count <= 0) count = 1;
    void *p = nullptr;
    // stream-ordered allocation from the device's default pool (release threshold raised in mofem_create): after
    // the first pass over a problem size every allocation of the symbolic / solve phases is a pool hit
    cudaError_t e = cudaMallocAsync(&p, sizeof(T) * (size_t)count, ctx->stream);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return ctx_fail(ctx, MOFEM_E_NOMEM, std::string("cudaMallocAsync of ") + std::to_string(sizeof(T) * (size_t)count) + " bytes: " + cudaGetErrorString(e));
    }
    pool.push_back(p);
    *out = (T *)p;
    return MOFEM_OK;
}

static void free_pool(mofem_ctx *ctx, std::vector<void *> &pool)
{
    for (void *p : pool) cudaFreeAsync(p, ctx->stream);
    pool.clear();
}

template <typename T>
static int upload(mofem_ctx *ctx, const T *src, int64_t count, const T **out, std::vector<void *> &pool)
{
    *out = nullptr;
    if (!src) return MOFEM_OK;
    if (is_device_ptr(src)) { *out = src; return MOFEM_OK; }  // device-resident input: borrowed, zero-copy
    T *d = nullptr;
    int rc = dev_alloc(ctx, &d, count, pool);
    if (rc) return rc;
    if (count > 0) CK(cudaMemcpyAsync(d, src, sizeof(T) * (size_t)count, cudaMemcpyDefault, ctx->stream));
    *out = d;
    return MOFEM_OK;
}

static void tic(mofem_ctx *ctx) { cudaEventRecord(ctx->ev0, ctx->stream); }
static double toc(mofem_ctx *ctx)
{
    float ms = 0;
    cudaEventRecord(ctx->ev1, ctx->stream);
    cudaEventSynchronize(ctx->ev1);
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    return ms;
}

extern "C" {

static void p2p_close(mofem_ctx *ctx);

const char *mofem_version(void) { return "mofem_b200 0.1.0 (sm_100a)"; }

int mofem_create(mofem_ctx **out, int device)
{
    if (!out) return MOFEM_E_ARG;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) { cudaGetLastError(); return MOFEM_E_CUDA; }  // no GPU: fail loudly, no CPU fallback
    if (device < 0 || device >= ndev) return MOFEM_E_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return MOFEM_E_CUDA;
    mofem_ctx *ctx = new mofem_ctx();
    ctx->device = device;
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess) {
        delete ctx;
        return MOFEM_E_CUDA;
    }
    cudaMalloc(&ctx->stats, sizeof(unsigned long long) * 4);
    cudaMalloc(&ctx->err_flag, sizeof(int));
    cudaMemPool_t mp;
    if (cudaDeviceGetDefaultMemPool(&mp, device) == cudaSuccess) {
        uint64_t keep = UINT64_MAX;  // keep freed blocks cached in the pool instead of returning them to the driver
        cudaMemPoolSetAttribute(mp, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    *out = ctx;
    return MOFEM_OK;
}

void mofem_destroy(mofem_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    free_pool(ctx, ctx->problem_allocs); free_pool(ctx, ctx->sym_allocs); free_pool(ctx, ctx->sys_allocs);
    p2p_close(ctx);
    free_pool(ctx, ctx->dist_allocs);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->dist.comm && nccl_api()) nccl_api()->CommDestroy(ctx->dist.comm);
    cudaFree(ctx->stats); cudaFree(ctx->err_flag);
    cudaEventDestroy(ctx->ev0); cudaEventDestroy(ctx->ev1);
    for (cudaEvent_t ev : ctx->prof_events) cudaEventDestroy(ev);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char *mofem_last_error(const mofem_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }
void *mofem_stream(mofem_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int mofem_set_problem(mofem_ctx *ctx, const mofem_problem_t *p)
{
    if (!ctx || !p) return MOFEM_E_ARG;
    CK(cudaSetDevice(ctx->device));
    if (p->n_node < 0 || p->n_elem < 0 || p->n_cell < 0 || p->n_tri < 0 || p->n_mesh < 1 || p->n_mesh > 8 ||
        p->solver < 1 || p->solver > 3)
        return ctx_fail(ctx, MOFEM_E_ARG, "mofem_set_problem: bad sizes / solver");
    if (p->n_node >= (1ll << 30)) return ctx_fail(ctx, MOFEM_E_LIMIT, "n_node exceeds 2^30 (int32 DOF indices)");
    if (!p->node_xy || !p->elem_nodes || !p->elem_nn || !p->elem_mat || !p->mat_D || !p->cell_elem || !p->cell_kind ||
        !p->cell_tri_ptr || (p->n_tri && (!p->tri_xy || !p->tri_rho)))
        return ctx_fail(ctx, MOFEM_E_ARG, "mofem_set_problem: null array");
    if ((p->tri_curved != nullptr) != (p->tri_mid != nullptr))
        return ctx_fail(ctx, MOFEM_E_ARG, "tri_curved and tri_mid must be given together");
    free_pool(ctx, ctx->problem_allocs); free_pool(ctx, ctx->sym_allocs); free_pool(ctx, ctx->sys_allocs);
    ctx->have_problem = ctx->have_symbolic = ctx->have_matrix = ctx->have_system = ctx->have_solution = false;
    tic(ctx);
    DevProblem &P = ctx->P;
    P.n_node = p->n_node; P.n_elem = p->n_elem; P.n_cell = p->n_cell; P.n_tri = p->n_tri;
    P.n_mesh = p->n_mesh; P.n_mat = p->n_mat; P.solver = p->solver;
    auto &pool = ctx->problem_allocs;
    int rc;
    if ((rc = upload(ctx, p->node_xy, 2 * p->n_node, &P.node_xy, pool))) return rc;
    if ((rc = upload(ctx, p->elem_nodes, 9 * p->n_elem, &P.elem_nodes, pool))) return rc;
    if ((rc = upload(ctx, p->elem_nn, p->n_elem, &P.elem_nn, pool))) return rc;
    if ((rc = upload(ctx, p->elem_mat, p->n_elem, &P.elem_mat, pool))) return rc;
    if ((rc = upload(ctx, p->mat_D, 9 * (int64_t)p->n_mat, &P.mat_D, pool))) return rc;
    if ((rc = upload(ctx, p->cell_elem, p->n_cell * p->n_mesh, &P.cell_elem, pool))) return rc;
    if ((rc = upload(ctx, p->cell_kind, p->n_cell, &P.cell_kind, pool))) return rc;
    if ((rc = upload(ctx, p->cell_tri_ptr, p->n_cell + 1, &P.cell_tri_ptr, pool))) return rc;
    if ((rc = upload(ctx, p->tri_xy, 6 * p->n_tri, &P.tri_xy, pool))) return rc;
    if ((rc = upload(ctx, p->tri_rho, 3 * p->n_tri * p->n_mesh, &P.tri_rho, pool))) return rc;
    if ((rc = upload(ctx, p->tri_mid, 6 * p->n_tri, &P.tri_mid, pool))) return rc;
    if ((rc = upload(ctx, p->tri_curved, p->n_tri, &P.tri_curved, pool))) return rc;
    ctx->ms[4] = toc(ctx);
    CK(cudaGetLastError());
    ctx->have_problem = true;
    return MOFEM_OK;
}

int mofem_symbolic(mofem_ctx *ctx)
{
    if (!ctx || !ctx->have_problem) return ctx_fail(ctx, MOFEM_E_ARG, "mofem_symbolic: call mofem_set_problem first");
    CK(cudaSetDevice(ctx->device));
    free_pool(ctx, ctx->sym_allocs); free_pool(ctx, ctx->sys_allocs);
    ctx->have_symbolic = ctx->have_matrix = ctx->have_system = ctx->have_solution = false;
    cudaStream_t s = ctx->stream;
    const DevProblem &P = ctx->P;
    auto &pool = ctx->sym_allocs;
    int64_t &L = ctx->launches[0];
    L = 0;
    tic(ctx);
    int rc;
    // --- tasks -------------------------------------------------------------------------
    int32_t *ntask, *npair, *ncoo;
    int64_t *task_off, *pair_off, *coo_off;
    std::vector<void *> tmp;
    auto cleanup = [&]() { free_pool(ctx, tmp); };
    if ((rc = dev_alloc(ctx, &ntask, P.n_cell, tmp)) || (rc = dev_alloc(ctx, &npair, P.n_cell, tmp)) ||
        (rc = dev_alloc(ctx, &ncoo, P.n_cell, tmp)) || (rc = dev_alloc(ctx, &task_off, P.n_cell + 1, tmp)) ||
        (rc = dev_alloc(ctx, &pair_off, P.n_cell + 1, tmp)) || (rc = dev_alloc(ctx, &coo_off, P.n_cell + 1, tmp))) {
        cleanup(); return rc;
    }
#define CKT(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { cleanup(); return ctx_fail(ctx, MOFEM_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); } } while (0)
    CKT(cudaMemsetAsync(ctx->err_flag, 0, sizeof(int), s));
    CKT(launch_cell_counts(P, ntask, npair, ncoo, ctx->err_flag, s)); L++;
    CKT(exclusive_scan_i32_to_i64(ntask, task_off, P.n_cell, task_off + P.n_cell, s, &L));
    CKT(exclusive_scan_i32_to_i64(npair, pair_off, P.n_cell, pair_off + P.n_cell, s, &L));
    CKT(exclusive_scan_i32_to_i64(ncoo, coo_off, P.n_cell, coo_off + P.n_cell, s, &L));
    int64_t tot[3]; int eflag = 0;
    CKT(cudaMemcpyAsync(&tot[0], task_off + P.n_cell, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    CKT(cudaMemcpyAsync(&tot[1], pair_off + P.n_cell, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    CKT(cudaMemcpyAsync(&tot[2], coo_off + P.n_cell, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    CKT(cudaMemcpyAsync(&eflag, ctx->err_flag, sizeof(int), cudaMemcpyDeviceToHost, s));
    CKT(cudaStreamSynchronize(s));
    if (eflag == MOFEM_E_ELEMENT) { cleanup(); return ctx_fail(ctx, MOFEM_E_ELEMENT, "Wrong element numbering!"); }
    DevTasks &T = ctx->T;
    T.n_task = tot[0]; ctx->n_pair = tot[1]; ctx->n_coo = tot[2];
    if (ctx->n_pair >= (1ll << 31) || T.n_task >= (1ll << 31)) {
        cleanup();
        return ctx_fail(ctx, MOFEM_E_LIMIT, "more than 2^31 node pairs on one device: shard the cells over more GPUs");
    }
    if ((rc = dev_alloc(ctx, &T.cell, T.n_task, pool)) || (rc = dev_alloc(ctx, &T.e1, T.n_task, pool)) ||
        (rc = dev_alloc(ctx, &T.e2, T.n_task, pool)) || (rc = dev_alloc(ctx, &T.slots, T.n_task, pool)) ||
        (rc = dev_alloc(ctx, &T.cls, T.n_task, pool)) || (rc = dev_alloc(ctx, &T.pair_off, T.n_task, pool)) ||
        (rc = dev_alloc(ctx, &ctx->cls_tasks, T.n_task, pool))) { cleanup(); return rc; }
    CKT(launch_fill_tasks(P, task_off, pair_off, T, ctx->cell_kernel, s)); L++;
    // class buckets
    int64_t *cls_count_d, *cls_start_d;
    if ((rc = dev_alloc(ctx, &cls_count_d, N_BUCKET, tmp)) || (rc = dev_alloc(ctx, &cls_start_d, N_BUCKET, tmp))) { cleanup(); return rc; }
    CKT(cudaMemsetAsync(cls_count_d, 0, sizeof(int64_t) * N_BUCKET, s));
    CKT(launch_bucket_tasks(T, cls_count_d, nullptr, nullptr, 0, s)); L++;
    CKT(cudaMemcpyAsync(ctx->cls_count, cls_count_d, sizeof(int64_t) * N_BUCKET, cudaMemcpyDeviceToHost, s));
    CKT(cudaStreamSynchronize(s));
    int64_t acc = 0;
    for (int k = 0; k < N_BUCKET; k++) { ctx->cls_start[k] = acc; acc += ctx->cls_count[k]; }
    CKT(cudaMemcpyAsync(cls_start_d, ctx->cls_start, sizeof(int64_t) * N_BUCKET, cudaMemcpyHostToDevice, s));
    CKT(cudaMemsetAsync(cls_count_d, 0, sizeof(int64_t) * N_BUCKET, s));
    CKT(launch_bucket_tasks(T, cls_count_d, cls_start_d, ctx->cls_tasks, 1, s)); L++;
    {   // records of the cell kernel: the first tasks of its cells are sub-buckets 0 .. N_SUB-2 of class FAST_CLASS, contiguous in cls_tasks
        int64_t n_fast = 0;
        for (int b = 0; b < N_SUB - 1; b++) n_fast += ctx->cls_count[FAST_CLASS * N_SUB + b];
        ctx->cq_rec = nullptr;
        if (n_fast) {
            if ((rc = dev_alloc(ctx, &ctx->cq_rec, n_fast, pool))) { cleanup(); return rc; }
            CKT(launch_cq_records(P, T, ctx->cls_tasks + ctx->cls_start[FAST_CLASS * N_SUB], n_fast, ctx->cq_rec, s)); L++;
        }
    }
    // --- CSR symbolic ---------------------------------------------------------------------
    DevCsr &C = ctx->C;
    C.n_node = P.n_node;
    int32_t *row_cnt, *row_uniq;
    if ((rc = dev_alloc(ctx, &row_cnt, P.n_node, tmp)) || (rc = dev_alloc(ctx, &row_uniq, P.n_node, tmp)) ||
        (rc = dev_alloc(ctx, &C.row_start, P.n_node + 1, pool)) || (rc = dev_alloc(ctx, &C.node_ptr, P.n_node + 4, pool))) {
        cleanup(); return rc;
    }
    CKT(cudaMemsetAsync(row_cnt, 0, sizeof(int32_t) * (size_t)std::max<int64_t>(P.n_node, 1), s));
    CKT(csr_count_rows(P, T, row_cnt, s)); L++;
    CKT(exclusive_scan_i32_to_i64(row_cnt, C.row_start, P.n_node, C.row_start + P.n_node, s, &L));
    C.n_entry = ctx->n_coo / 4;        // every contribution (a node pair of a block, or its transposed copy) is four COO triplets: no read-back
    if (C.n_entry >= (1ll << 32)) {
        cleanup();
        return ctx_fail(ctx, MOFEM_E_LIMIT, "more than 2^32 block contributions on one device: shard the cells over more GPUs");
    }
    uint64_t *entries;
    if ((rc = dev_alloc(ctx, &entries, C.n_entry, tmp))) { cleanup(); return rc; }
    CKT(cudaMemsetAsync(row_cnt, 0, sizeof(int32_t) * (size_t)std::max<int64_t>(P.n_node, 1), s));
    CKT(csr_fill_rows(P, T, C.row_start, row_cnt, entries, s)); L++;
    CKT(cudaMemsetAsync(ctx->err_flag, 0, sizeof(int), s));
    CKT(csr_sort_rows(P.n_node, C.row_start, entries, row_uniq, nullptr, ctx->err_flag, 0, s)); L++;   // rank sort of the short rows
    CKT(cudaMemcpyAsync(&eflag, ctx->err_flag, sizeof(int), cudaMemcpyDeviceToHost, s));
    CKT(cudaStreamSynchronize(s));
    if (eflag) {                   // some node rows have > 512 contributions: shared-memory bitonic sort for those
        CKT(cudaMemsetAsync(ctx->err_flag, 0, sizeof(int), s));
        CKT(csr_sort_rows(P.n_node, C.row_start, entries, row_uniq, nullptr, ctx->err_flag, 1, s)); L++;
        CKT(cudaMemcpyAsync(&eflag, ctx->err_flag, sizeof(int), cudaMemcpyDeviceToHost, s));
        CKT(cudaStreamSynchronize(s));
        if (eflag == MOFEM_E_LIMIT) {  // ... and some > 1024: redo those with a global scratch area
            uint64_t *scratch;
            if ((rc = dev_alloc(ctx, &scratch, 2 * C.n_entry, tmp))) { cleanup(); return rc; }
            CKT(csr_sort_rows(P.n_node, C.row_start, entries, row_uniq, scratch, ctx->err_flag, 1, s)); L++;
        }
    }
    CKT(exclusive_scan_i32_to_i64(row_uniq, C.node_ptr, P.n_node, C.node_ptr + P.n_node, s, &L));
    CKT(cudaMemcpyAsync(&C.n_upair, C.node_ptr + P.n_node, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    CKT(cudaStreamSynchronize(s));
    if ((rc = dev_alloc(ctx, &C.node_idx, C.n_upair + 8, pool)) || (rc = dev_alloc(ctx, &C.pair_row, C.n_upair, pool)) ||
        (rc = dev_alloc(ctx, &C.contrib_ptr, C.n_upair + 1, pool)) || (rc = dev_alloc(ctx, &C.pair_dst, 2 * ctx->n_pair, pool)) ||
        (rc = dev_alloc(ctx, &ctx->data, 4 * C.n_upair, pool)) || (rc = dev_alloc(ctx, &ctx->blk_val, 4 * C.n_entry, pool))) {
        cleanup(); return rc;
    }
    uint32_t *contrib_src;
    if ((rc = dev_alloc(ctx, &contrib_src, C.n_entry, tmp))) { cleanup(); return rc; }
    CKT(csr_emit(P.n_node, C.row_start, entries, C.node_ptr, C.node_idx, C.pair_row, C.contrib_ptr, contrib_src, s)); L++;
    CKT(cudaMemcpyAsync(C.contrib_ptr + C.n_upair, &C.n_entry, sizeof(int64_t), cudaMemcpyHostToDevice, s));
    CKT(csr_pair_dst(C, contrib_src, ctx->n_pair, s)); L++;
    ctx->ms[0] = toc(ctx);
    CKT(cudaGetLastError());
    cleanup();
#undef CKT
    ctx->have_symbolic = true;
    return MOFEM_OK;
}

int mofem_assemble(mofem_ctx *ctx)
{
    if (!ctx || !ctx->have_symbolic) return ctx_fail(ctx, MOFEM_E_ARG, "mofem_assemble: call mofem_symbolic first");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    ctx->launches[1] = ctx->launches[2] = 0;
    CK(cudaMemsetAsync(ctx->stats, 0, sizeof(unsigned long long) * 4, s));
    tic(ctx);
    const BlkOut out{ctx->blk_val, ctx->C.pair_dst, ctx->n_pair};   // records in reduction order, addressed through the inverse map
    for (int k = 0; k < N_CLASS; k++) {
        if (k == FAST_CLASS) {      // all (incident elements, triangles) sub-buckets in one launch; the last sub-bucket holds no first tasks
            int64_t any = 0;
            for (int b = 0; b < N_SUB - 1; b++) any += ctx->cls_count[k * N_SUB + b];
            if (any) {
                CK(launch_cell_quads(ctx->P, ctx->cq_rec, ctx->cls_count + k * N_SUB, out, ctx->stats, s));
                ctx->launches[1]++;
            }
            continue;
        }
        int64_t n_cls = 0;
        for (int b = 0; b < N_SUB; b++) n_cls += ctx->cls_count[k * N_SUB + b];
        if (!n_cls) continue;
        CK(launch_integrate_class(k, ctx->P, ctx->T, ctx->cls_tasks + ctx->cls_start[k * N_SUB], n_cls, out,
                                  ctx->stats, s));
        ctx->launches[1]++;
    }
    ctx->ms[1] = toc(ctx);
    tic(ctx);
    CK(csr_numeric(ctx->C, ctx->blk_val, ctx->data, s));
    ctx->launches[2]++;
    ctx->ms[2] = toc(ctx);
    CK(cudaMemcpyAsync(ctx->stats_host, ctx->stats, sizeof(unsigned long long) * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaGetLastError());
    if (ctx->stats_host[2]) return ctx_fail(ctx, MOFEM_E_NEWTON, "Too many iterations!");
    ctx->have_matrix = true;
    ctx->have_system = ctx->have_solution = false;
    return MOFEM_OK;
}

int mofem_get_info(mofem_ctx *ctx, mofem_info_t *info)
{
    if (!ctx || !info || !ctx->have_symbolic) return ctx_fail(ctx, MOFEM_E_ARG, "mofem_get_info: no symbolic phase yet");
    info->n_dof = 2 * ctx->P.n_node;
    info->nnz = 4 * ctx->C.n_upair;
    info->n_block = ctx->T.n_task;
    info->n_coo = ctx->n_coo;
    info->n_pair = ctx->n_pair;
    info->newton_updates = (int64_t)ctx->stats_host[0];
    info->n_eval = (int64_t)ctx->stats_host[1];
    return MOFEM_OK;
}

int mofem_get_csr(mofem_ctx *ctx, int32_t *indptr, int32_t *indices, double *data)
{
    if (!ctx || !ctx->have_symbolic) return ctx_fail(ctx, MOFEM_E_ARG, "mofem_get_csr: call mofem_symbolic first");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const int64_t nnz = 4 * ctx->C.n_upair, nrow = 2 * ctx->P.n_node;
    if (nnz >= (1ll << 31)) return ctx_fail(ctx, MOFEM_E_LIMIT, "nnz >= 2^31: int32 CSR export not possible");
    std::vector<void *> tmp;
    int rc = MOFEM_OK;
    int32_t *ip = nullptr, *ix = nullptr;
    if (indptr) { if (is_device_ptr(indptr)) ip = indptr; else if ((rc = dev_alloc(ctx, &ip, nrow + 1, tmp))) return rc; }
    if (indices) { if (is_device_ptr(indices)) ix = indices; else if ((rc = dev_alloc(ctx, &ix, nnz, tmp))) { free_pool(ctx, tmp); return rc; } }
    cudaError_t e = csr_export_pattern(ctx->C, ip, ix, s);
    if (e == cudaSuccess && indptr && ip != indptr) e = cudaMemcpyAsync(indptr, ip, sizeof(int32_t) * (nrow + 1), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess && indices && ix != indices) e = cudaMemcpyAsync(indices, ix, sizeof(int32_t) * nnz, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess && data) {
        if (!ctx->have_matrix) { free_pool(ctx, tmp); return ctx_fail(ctx, MOFEM_E_ARG, "mofem_get_csr: values requested before mofem_assemble"); }
        e = cudaMemcpyAsync(data, ctx->data, sizeof(double) * nnz, cudaMemcpyDefault, s);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    free_pool(ctx, tmp);
    if (e != cudaSuccess) return ctx_fail(ctx, MOFEM_E_CUDA, std::string("mofem_get_csr: ") + cudaGetErrorString(e));
    return MOFEM_OK;
}

int mofem_get_coo_values(mofem_ctx *ctx, double *V)
{
    if (!ctx || !ctx->have_matrix || !V) return ctx_fail(ctx, MOFEM_E_ARG, "mofem_get_coo_values: call mofem_assemble first");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const int64_t nt = ctx->T.n_task;
    // pure re-layout (no arithmetic) of blk_val (reduction order) into the reference's emission order, done on the host
    std::vector<double> blk((size_t)(4 * ctx->C.n_entry));
    std::vector<uint32_t> dst((size_t)ctx->n_pair);
    CK(cudaMemcpyAsync(dst.data(), ctx->C.pair_dst, sizeof(uint32_t) * dst.size(), cudaMemcpyDeviceToHost, s));
    std::vector<int32_t> e1(nt), e2(nt), nn((size_t)ctx->P.n_elem);
    std::vector<int64_t> po(nt);
    CK(cudaMemcpyAsync(blk.data(), ctx->blk_val, sizeof(double) * blk.size(), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(e1.data(), ctx->T.e1, sizeof(int32_t) * nt, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(e2.data(), ctx->T.e2, sizeof(int32_t) * nt, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(po.data(), ctx->T.pair_off, sizeof(int64_t) * nt, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(nn.data(), ctx->P.elem_nn, sizeof(int32_t) * nn.size(), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    std::vector<double> host;
    double *out = V;
    const bool dev_out = is_device_ptr(V);
    if (dev_out) { host.resize((size_t)ctx->n_coo); out = host.data(); }
    int64_t off = 0;
    for (int64_t k = 0; k < nt; k++) {
        const int n1 = nn[e1[k]], n2 = nn[e2[k]];
        const uint32_t *d = dst.data() + po[k];      // record of node pair (a, c) of this block
        for (int a = 0; a < n1; a++)
            for (int c = 0; c < n2; c++)
                for (int p = 0; p < 2; p++)
                    for (int q = 0; q < 2; q++)
                        out[off + (int64_t)(2 * a + p) * (2 * n2) + 2 * c + q] = blk[4 * (size_t)d[a * n2 + c] + 2 * p + q];
        off += 4ll * n1 * n2;
        if (e1[k] != e2[k]) {
            for (int a = 0; a < n1; a++)
                for (int c = 0; c < n2; c++)
                    for (int p = 0; p < 2; p++)
                        for (int q = 0; q < 2; q++)
                            out[off + (int64_t)(2 * c + q) * (2 * n1) + 2 * a + p] = blk[4 * (size_t)d[a * n2 + c] + 2 * p + q];
            off += 4ll * n1 * n2;
        }
    }
    if (dev_out) { CK(cudaMemcpyAsync(V, host.data(), sizeof(double) * host.size(), cudaMemcpyHostToDevice, s)); CK(cudaStreamSynchronize(s)); }
    return MOFEM_OK;
}

static DevSystem system_of(mofem_ctx *ctx)
{
    DevSystem A;
    A.n = 2 * ctx->P.n_node; A.node_ptr = ctx->C.node_ptr; A.node_idx = ctx->C.node_idx; A.data = ctx->data;
    A.n_own = ctx->have_halo ? 2 * ctx->dist.n_owned_node : A.n;
    A.n_blk = ctx->C.n_upair;   // upper bound of the sub-blocks in the owned rows (exact on a single GPU)
    return A;
}
static DistPlan *dist_of(mofem_ctx *ctx) { return ctx->have_halo ? &ctx->dist : nullptr; }

int mofem_set_system(mofem_ctx *ctx, const double *rhs, const uint8_t *fixed, const double *fixed_value)
{
    if (!ctx || !ctx->have_matrix) return ctx_fail(ctx, MOFEM_E_ARG, "mofem_set_system: call mofem_assemble first");
    if (!rhs || !fixed || !fixed_value) return ctx_fail(ctx, MOFEM_E_ARG, "mofem_set_system: null array");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const int64_t n = 2 * ctx->P.n_node;
    int rc;
    const bool want_xbuf = ctx->have_halo && ctx->dist.p2p;
    if (ctx->sys_n != n || ctx->sys_allocs.empty() || want_xbuf != ctx->slab_in_xbuf) {
        free_pool(ctx, ctx->sys_allocs);
        auto &pool = ctx->sys_allocs;
        PcgBuffers &W = ctx->W;
        // the five vectors the PCG loop touches every iteration live in one slab (the L2 persistence window of mofem_solve);
        // on the peer path the slab is part of the IPC-exported exchange buffer
        const int64_t npad = p2p_slab_pad(n);
        double *slab;
        if (want_xbuf) {
            if (n != 2 * (ctx->dist.n_owned_node + ctx->dist.n_ghost))
                return ctx_fail(ctx, MOFEM_E_ARG, "mofem_set_system: problem size does not match the halo plan of the peer transport");
            slab = ctx->dist.xbuf + p2p_p_off(ctx->dist.world);
        } else if ((rc = dev_alloc(ctx, &slab, 7 * npad, pool))) return rc;
        ctx->slab_in_xbuf = want_xbuf;
        // order = how much a vector gains from staying in L2: p (gathered by the SpMV, read by both vector kernels), r and dinv
        // (both vector kernels), then q and x (one kernel each); "l2_window" picks how many of them the window covers
        W.p = slab; W.r = slab + npad; W.dinv = slab + 2 * npad; W.q = slab + 3 * npad; W.x = slab + 4 * npad;
        W.b = slab + 5 * npad; W.ufix = slab + 6 * npad;
        ctx->slab = slab; ctx->slab_vec_bytes = sizeof(double) * (size_t)npad; ctx->slab_hot_bytes = 5 * ctx->slab_vec_bytes;
        if ((rc = dev_alloc(ctx, &W.part_a, PCG_MAX_PART, pool)) ||
            (rc = dev_alloc(ctx, &W.part_b, PCG_MAX_PART, pool)) || (rc = dev_alloc(ctx, &W.part_c, PCG_MAX_PART, pool)) ||
            (rc = dev_alloc(ctx, &W.sc, 16, pool)) || (rc = dev_alloc(ctx, &W.counters, 8, pool)) || (rc = dev_alloc(ctx, &W.row_halo, n / 2 + 1, pool)) || (rc = dev_alloc(ctx, &ctx->fixed, n, pool)) ||
            (rc = dev_alloc(ctx, &ctx->rhs, n, pool)) || (rc = dev_alloc(ctx, &ctx->fixed_val, n, pool)) ||
            (rc = dev_alloc(ctx, &ctx->u, n, pool)))
            return rc;
        W.ring_cap = ring_tile_capacity(n / 2, ctx->C.n_upair);
        char *tiles;
        if ((rc = dev_alloc(ctx, &tiles, 32 * W.ring_cap, pool))) return rc;
        W.ring_tiles = tiles;
        ctx->sys_n = n;
    }
    tic(ctx);
    CK(cudaMemcpyAsync(ctx->rhs, rhs, sizeof(double) * n, cudaMemcpyDefault, s));
    CK(cudaMemcpyAsync(ctx->fixed, fixed, sizeof(uint8_t) * n, cudaMemcpyDefault, s));
    CK(cudaMemcpyAsync(ctx->fixed_val, fixed_value, sizeof(double) * n, cudaMemcpyDefault, s));
    ctx->ms[4] += toc(ctx);
    ctx->launches[3] = 0;
    if (ctx->have_halo && ctx->dist.n_owned_node > ctx->P.n_node) return ctx_fail(ctx, MOFEM_E_ARG, "halo plan does not match the problem");
    CK(pcg_setup(system_of(ctx), ctx->rhs, ctx->fixed, ctx->fixed_val, ctx->W, dist_of(ctx), s, &ctx->launches[3]));
    CK(cudaStreamSynchronize(s));
    ctx->have_system = true;
    ctx->have_solution = false;
    return MOFEM_OK;
}

int mofem_solve(mofem_ctx *ctx, double rtol, int32_t maxit, double *u, mofem_solve_info_t *info)
{
    if (!ctx || !ctx->have_system) return ctx_fail(ctx, MOFEM_E_ARG, "mofem_solve: call mofem_set_system first");
    if (!(rtol > 0) || maxit < 1) return ctx_fail(ctx, MOFEM_E_ARG, "mofem_solve: bad rtol / maxit");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const int64_t n = 2 * ctx->P.n_node;
    mofem_solve_info_t local;
    if (!info) info = &local;
    tic(ctx);
    cudaEvent_t *prof = nullptr;
    if (ctx->profile) {
        while ((int)ctx->prof_events.size() < 2 * 4 * PCG_CHUNK) {
            cudaEvent_t ev;
            CK(cudaEventCreate(&ev));
            ctx->prof_events.push_back(ev);
        }
        prof = ctx->prof_events.data();
        for (double &v : ctx->prof_ms) v = 0.0;
    }
    // L2 persistence: p, q, x, r, dinv (5 vectors) are re-read every iteration while 9 bytes per non-zero of matrix
    // stream through the 126 MB L2 once; pin the vectors, let the matrix loads (evict-first) bypass them.
    bool window = false;
    if (ctx->l2_persist && ctx->slab_hot_bytes) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, ctx->device) == cudaSuccess && prop.persistingL2CacheMaxSize > 0) {
            const size_t hot = std::min<size_t>(ctx->slab_hot_bytes, ctx->slab_vec_bytes * (size_t)std::max(1, std::min(5, ctx->l2_window)));
            const size_t want = std::min<size_t>(hot, (size_t)prop.persistingL2CacheMaxSize);
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
            cudaStreamAttrValue attr{};
            attr.accessPolicyWindow.base_ptr = ctx->slab;
            attr.accessPolicyWindow.num_bytes = std::min<size_t>(hot, (size_t)prop.accessPolicyMaxWindowSize);
            attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)want / (double)attr.accessPolicyWindow.num_bytes);
            attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            window = cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &attr) == cudaSuccess;
            cudaGetLastError();
        }
    }
    ctx->W.use_ring = ctx->spmv_ring; ctx->W.use_pdl = ctx->pdl; ctx->W.floor_rtol = ctx->floor_rtol;
    cudaError_t run_err = pcg_run(system_of(ctx), ctx->fixed, ctx->W, dist_of(ctx), rtol, maxit, PCG_CHUNK, info, s,
                                  &ctx->launches[3], prof, ctx->prof_ms, ctx->profile);
    if (window) {
        cudaStreamAttrValue attr{};
        attr.accessPolicyWindow.num_bytes = 0;
        cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &attr);
        cudaCtxResetPersistingL2Cache();
        cudaGetLastError();
    }
    CK(run_err);
    CK(pcg_finish(system_of(ctx), ctx->W, dist_of(ctx), ctx->u, s));
    ctx->launches[3]++;
    ctx->ms[3] = toc(ctx);
    ctx->have_solution = true;
    if (u) {
        tic(ctx);
        CK(cudaMemcpyAsync(u, ctx->u, sizeof(double) * n, cudaMemcpyDefault, s));
        ctx->ms[5] = toc(ctx);
    }
    CK(cudaStreamSynchronize(s));
    if (!info->converged) return ctx_fail(ctx, MOFEM_E_NOCONV, "PCG did not reach rtol within maxit");
    return MOFEM_OK;
}

int mofem_energy(mofem_ctx *ctx, double *energy)
{
    if (!ctx || !ctx->have_solution || !energy) return ctx_fail(ctx, MOFEM_E_ARG, "mofem_energy: call mofem_solve first");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    CK(energy_of(system_of(ctx), ctx->u, ctx->W, dist_of(ctx), ctx->W.sc + 6, s));
    double e = 0;
    CK(cudaMemcpyAsync(&e, ctx->W.sc + 6, sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    *energy = 0.5 * e;
    return MOFEM_OK;
}

int mofem_spmv(mofem_ctx *ctx, const double *x, double *y)
{
    if (!ctx || !ctx->have_matrix || !x || !y) return ctx_fail(ctx, MOFEM_E_ARG, "mofem_spmv: call mofem_assemble first");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const int64_t n = 2 * ctx->P.n_node;
    std::vector<void *> tmp;
    const double *xd = x; double *yd = y;
    int rc;
    if (!is_device_ptr(x)) { double *t; if ((rc = dev_alloc(ctx, &t, n, tmp))) return rc; CK(cudaMemcpyAsync(t, x, sizeof(double) * n, cudaMemcpyHostToDevice, s)); xd = t; }
    if (!is_device_ptr(y)) { double *t; if ((rc = dev_alloc(ctx, &t, n, tmp))) { free_pool(ctx, tmp); return rc; } yd = t; }
    // row-partitioned mode: only the owned rows are computed; the ghost rows of y are returned as zeros
    const DevSystem A = system_of(ctx);
    cudaError_t e = A.n_own < n ? cudaMemsetAsync(yd + A.n_own, 0, sizeof(double) * (size_t)(n - A.n_own), s) : cudaSuccess;
    if (e == cudaSuccess) e = spmv(A, xd, yd, nullptr, s);
    if (e == cudaSuccess && yd != y) e = cudaMemcpyAsync(y, yd, sizeof(double) * n, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    free_pool(ctx, tmp);
    if (e != cudaSuccess) return ctx_fail(ctx, MOFEM_E_CUDA, std::string("mofem_spmv: ") + cudaGetErrorString(e));
    return MOFEM_OK;
}

int mofem_assemble_solve(mofem_ctx *ctx, const mofem_problem_t *p, const double *rhs, const uint8_t *fixed,
                         const double *fixed_value, double rtol, int32_t maxit, double *u, double *energy,
                         mofem_solve_info_t *info)
{
    int rc;
    if ((rc = mofem_set_problem(ctx, p))) return rc;
    if ((rc = mofem_symbolic(ctx))) return rc;
    if ((rc = mofem_assemble(ctx))) return rc;
    if ((rc = mofem_set_system(ctx, rhs, fixed, fixed_value))) return rc;
    if ((rc = mofem_solve(ctx, rtol, maxit, u, info))) return rc;
    if (energy && (rc = mofem_energy(ctx, energy))) return rc;
    return MOFEM_OK;
}

// ---- row-partitioned solve -------------------------------------------------------------------------------------
int mofem_dist_unique_id(uint8_t id[128])
{
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    const NcclApi *nc = nccl_api();
    if (!nc) return MOFEM_E_CUDA;
    ncclUniqueId u;
    if (nc->GetUniqueId(&u) != ncclSuccess) return MOFEM_E_CUDA;
    memcpy(id, &u, 128);
    return MOFEM_OK;
}

int mofem_dist_init(mofem_ctx *ctx, const uint8_t id[128], int32_t rank, int32_t world)
{
    if (!ctx || !id || world < 1 || rank < 0 || rank >= world) return ctx_fail(ctx, MOFEM_E_ARG, "mofem_dist_init: bad arguments");
    CK(cudaSetDevice(ctx->device));
    const NcclApi *nc = nccl_api();
    if (!nc) return ctx_fail(ctx, MOFEM_E_CUDA, "libnccl.so.2 could not be loaded");
    if (ctx->dist.comm) { nc->CommDestroy(ctx->dist.comm); ctx->dist.comm = nullptr; }
    ncclUniqueId u;
    memcpy(&u, id, 128);
    ncclResult_t r = nc->CommInitRank(&ctx->dist.comm, world, u, rank);
    if (r != ncclSuccess) return ctx_fail(ctx, MOFEM_E_CUDA, std::string("ncclCommInitRank: ") + nc->GetErrorString(r));
    ctx->dist.world = world; ctx->dist.rank = rank;
    return MOFEM_OK;
}

int mofem_set_halo(mofem_ctx *ctx, int64_t n_owned_node, int32_t n_neighbor, const int32_t *neighbor_rank,
                   const int64_t *send_ptr, const int32_t *send_idx, const int64_t *recv_ptr)
{
    if (!ctx) return MOFEM_E_ARG;
    CK(cudaSetDevice(ctx->device));
    p2p_close(ctx);
    free_pool(ctx, ctx->dist_allocs);
    ctx->have_halo = false;
    ctx->have_system = ctx->have_solution = false;
    if (n_neighbor < 0 || n_owned_node < 0) return ctx_fail(ctx, MOFEM_E_ARG, "mofem_set_halo: negative size");
    if (n_neighbor > 0 && (!neighbor_rank || !send_ptr || !recv_ptr)) return ctx_fail(ctx, MOFEM_E_ARG, "mofem_set_halo: null array");
    if (n_neighbor > 0 && !ctx->dist.comm) return ctx_fail(ctx, MOFEM_E_ARG, "mofem_set_halo: call mofem_dist_init first");
    DistPlan &D = ctx->dist;
    D.n_owned_node = n_owned_node;
    D.nb_rank.assign(neighbor_rank, neighbor_rank + n_neighbor);
    D.send_ptr.assign(1, 0); D.recv_ptr.assign(1, 0);
    if (n_neighbor) { D.send_ptr.assign(send_ptr, send_ptr + n_neighbor + 1); D.recv_ptr.assign(recv_ptr, recv_ptr + n_neighbor + 1); }
    const int64_t n_send = D.send_ptr.back();
    if (n_send && !send_idx) return ctx_fail(ctx, MOFEM_E_ARG, "mofem_set_halo: null send_idx");
    int rc;
    if ((rc = dev_alloc(ctx, &D.send_idx, n_send, ctx->dist_allocs)) || (rc = dev_alloc(ctx, &D.sendbuf, 2 * n_send, ctx->dist_allocs))) return rc;
    if (n_send) CK(cudaMemcpyAsync(D.send_idx, send_idx, sizeof(int32_t) * n_send, cudaMemcpyDefault, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->have_halo = true;
    return MOFEM_OK;
}

static void p2p_close(mofem_ctx *ctx)
{
    DistPlan &D = ctx->dist;
    cudaStreamSynchronize(ctx->stream);
    if (ctx->slab_in_xbuf) {   // the PCG vectors lived in the exchange buffer: the system has to be set again
        free_pool(ctx, ctx->sys_allocs);
        ctx->sys_n = 0; ctx->slab_in_xbuf = false; ctx->slab = nullptr; ctx->slab_hot_bytes = 0; ctx->slab_vec_bytes = 0;
        ctx->have_system = ctx->have_solution = false;
    }
    for (int r = 0; r < (int)D.peer_host.size(); r++)
        if (r != D.rank && D.peer_host[r]) cudaIpcCloseMemHandle(D.peer_host[r]);
    D.peer_host.clear();
    if (D.xbuf) cudaFree(D.xbuf);
    if (D.peer) cudaFree(D.peer);
    if (D.send_nb) cudaFree(D.send_nb);
    if (D.send_dst) cudaFree(D.send_dst);
    if (D.nb_ghost_off) cudaFree(D.nb_ghost_off);
    if (D.nb_rank_dev) cudaFree(D.nb_rank_dev);
    if (D.wait_rank) cudaFree(D.wait_rank);
    if (D.halo_rows) cudaFree(D.halo_rows);
    D.halo_rows = nullptr; D.n_halo_rows = 0; D.halo_rows_cap = 0;
    D.xbuf = nullptr; D.peer = nullptr; D.send_nb = nullptr; D.send_dst = nullptr; D.nb_ghost_off = nullptr;
    D.nb_rank_dev = nullptr; D.wait_rank = nullptr; D.n_wait = 0;
    D.p2p = false;
    cudaGetLastError();
}

int mofem_p2p_export(mofem_ctx *ctx, uint8_t handle[64], int64_t *n_ghost)
{
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
    if (!ctx || !handle || !n_ghost) return MOFEM_E_ARG;
    if (!ctx->have_halo || !ctx->dist.comm) return ctx_fail(ctx, MOFEM_E_ARG, "mofem_p2p_export: call mofem_dist_init and mofem_set_halo first");
    CK(cudaSetDevice(ctx->device));
    p2p_close(ctx);
    DistPlan &D = ctx->dist;
    if (D.world > P2P_MAX_WORLD)   // one thread and one shared staging slot per rank in the mailbox code
        return ctx_fail(ctx, MOFEM_E_LIMIT, "mofem_p2p_export: peer transport supports at most " + std::to_string(P2P_MAX_WORLD) + " ranks (use the NCCL transport)");
    D.n_ghost = D.recv_ptr.back();
    const size_t bytes = sizeof(double) * (size_t)p2p_buffer_doubles(D.world, D.n_owned_node + D.n_ghost);
    CK(cudaMalloc(&D.xbuf, bytes));        // plain cudaMalloc: pool memory cannot be exported through legacy IPC handles
    CK(cudaMemset(D.xbuf, 0, bytes));
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, D.xbuf));
    memcpy(handle, &h, 64);
    *n_ghost = D.n_ghost;
    return MOFEM_OK;
}

int mofem_p2p_import(mofem_ctx *ctx, const uint8_t *handles, const int64_t *peer_n_owned, const int64_t *send_remote_off)
{
    if (!ctx || !handles || !peer_n_owned) return MOFEM_E_ARG;
    DistPlan &D = ctx->dist;
    if (!D.xbuf) return ctx_fail(ctx, MOFEM_E_ARG, "mofem_p2p_import: call mofem_p2p_export first");
    CK(cudaSetDevice(ctx->device));
    const int W = D.world, n_nb = (int)D.nb_rank.size();
    if (n_nb && !send_remote_off) return ctx_fail(ctx, MOFEM_E_ARG, "mofem_p2p_import: null send_remote_off");
    D.peer_host.assign(W, nullptr);
    for (int r = 0; r < W; r++) {
        if (r == D.rank) { D.peer_host[r] = D.xbuf; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + 64 * (size_t)r, 64);
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            p2p_close(ctx);
            return ctx_fail(ctx, MOFEM_E_CUDA, std::string("cudaIpcOpenMemHandle(rank ") + std::to_string(r) + "): " + cudaGetErrorString(e));
        }
        D.peer_host[r] = (double *)p;
    }
    const int64_t n_send = D.send_ptr.back();
    std::vector<int32_t> nb_of((size_t)n_send);
    std::vector<int64_t> dst((size_t)n_send), goff((size_t)std::max(n_nb, 1));
    std::vector<int> wait;
    for (int k = 0; k < n_nb; k++) {
        for (int64_t i = D.send_ptr[k]; i < D.send_ptr[k + 1]; i++) { nb_of[i] = k; dst[i] = i - D.send_ptr[k]; }
        // ghost segment of the neighbour's p = after its owned nodes; our entries start send_remote_off[k] nodes into it
        goff[k] = p2p_p_off(W) + 2 * (peer_n_owned[D.nb_rank[k]] + send_remote_off[k]);
        if (D.recv_ptr[k + 1] > D.recv_ptr[k]) wait.push_back(D.nb_rank[k]);
    }
    D.send_remote_off.assign(send_remote_off, send_remote_off + n_nb);
    D.n_wait = (int)wait.size();
    CK(cudaMalloc(&D.peer, sizeof(double *) * W));
    CK(cudaMalloc(&D.send_nb, sizeof(int32_t) * std::max<int64_t>(n_send, 1)));
    CK(cudaMalloc(&D.send_dst, sizeof(int64_t) * std::max<int64_t>(n_send, 1)));
    CK(cudaMalloc(&D.nb_ghost_off, sizeof(int64_t) * goff.size()));
    CK(cudaMalloc(&D.nb_rank_dev, sizeof(int) * std::max(n_nb, 1)));
    CK(cudaMalloc(&D.wait_rank, sizeof(int) * std::max(D.n_wait, 1)));
    CK(cudaMemcpy(D.peer, D.peer_host.data(), sizeof(double *) * W, cudaMemcpyHostToDevice));
    if (n_send) {
        CK(cudaMemcpy(D.send_nb, nb_of.data(), sizeof(int32_t) * n_send, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(D.send_dst, dst.data(), sizeof(int64_t) * n_send, cudaMemcpyHostToDevice));
    }
    CK(cudaMemcpy(D.nb_ghost_off, goff.data(), sizeof(int64_t) * goff.size(), cudaMemcpyHostToDevice));
    if (n_nb) CK(cudaMemcpy(D.nb_rank_dev, D.nb_rank.data(), sizeof(int) * n_nb, cudaMemcpyHostToDevice));
    if (D.n_wait) CK(cudaMemcpy(D.wait_rank, wait.data(), sizeof(int) * D.n_wait, cudaMemcpyHostToDevice));
    D.epoch = 0;
    D.p2p = true;
    return MOFEM_OK;
}

int mofem_p2p_close(mofem_ctx *ctx)
{
    if (!ctx) return MOFEM_E_ARG;
    p2p_close(ctx);
    return MOFEM_OK;
}

int mofem_clear_halo(mofem_ctx *ctx)
{
    if (!ctx) return MOFEM_E_ARG;
    p2p_close(ctx);
    free_pool(ctx, ctx->dist_allocs);
    ctx->have_halo = false;
    ctx->have_system = ctx->have_solution = false;
    return MOFEM_OK;
}

// ---- sampler --------------------------------------------------------------------
int mofem_sample(mofem_ctx *ctx, const double *u, int32_t n_sampling, int32_t solver, const double *grid, double *out,
                 int64_t out_capacity_units, int64_t *n_unit)
{
    if (!ctx || !ctx->have_problem || !n_unit) return ctx_fail(ctx, MOFEM_E_ARG, "mofem_sample: call mofem_set_problem first");
    if (n_sampling < 1 || n_sampling > 64 || (solver != 2 && solver != 3)) return ctx_fail(ctx, MOFEM_E_ARG, "mofem_sample: bad n_sampling / solver");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    DevProblem P = ctx->P;
    P.solver = solver;
    std::vector<void *> tmp;
    int rc;
    int32_t *cnt; int64_t *off;
    if ((rc = dev_alloc(ctx, &cnt, P.n_cell, tmp)) || (rc = dev_alloc(ctx, &off, P.n_cell + 1, tmp))) { free_pool(ctx, tmp); return rc; }
#define CKS(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { free_pool(ctx, tmp); return ctx_fail(ctx, MOFEM_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); } } while (0)
    CKS(sample_count_units(P, cnt, s));
    CKS(exclusive_scan_i32_to_i64(cnt, off, P.n_cell, off + P.n_cell, s, nullptr));
    int64_t nu = 0;
    CKS(cudaMemcpyAsync(&nu, off + P.n_cell, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    CKS(cudaStreamSynchronize(s));
    *n_unit = nu;
    if (!out) { free_pool(ctx, tmp); return MOFEM_OK; }          // size query
    if (!grid) { free_pool(ctx, tmp); return ctx_fail(ctx, MOFEM_E_ARG, "mofem_sample: null grid"); }
    if (out_capacity_units < nu) { free_pool(ctx, tmp); return ctx_fail(ctx, MOFEM_E_ARG, "mofem_sample: output buffer too small"); }
    const int64_t n_dof = 2 * P.n_node, per = (int64_t)n_sampling * n_sampling;
    int32_t *unit_cell; int64_t *unit_tri; double *grid_d, *out_d = out, *u_d = nullptr;
    if ((rc = dev_alloc(ctx, &unit_cell, nu, tmp)) || (rc = dev_alloc(ctx, &unit_tri, nu, tmp)) || (rc = dev_alloc(ctx, &grid_d, n_sampling, tmp))) {
        free_pool(ctx, tmp); return rc;
    }
    const double *uu = u;
    if (!u) {
        if (!ctx->have_solution) { free_pool(ctx, tmp); return ctx_fail(ctx, MOFEM_E_ARG, "mofem_sample: no displacement given and no solve done"); }
        uu = ctx->u;
    } else if (!is_device_ptr(u)) {
        if ((rc = dev_alloc(ctx, &u_d, n_dof, tmp))) { free_pool(ctx, tmp); return rc; }
        CKS(cudaMemcpyAsync(u_d, u, sizeof(double) * n_dof, cudaMemcpyHostToDevice, s));
        uu = u_d;
    }
    const bool out_host = !is_device_ptr(out);
    if (out_host && (rc = dev_alloc(ctx, &out_d, nu * 7 * per, tmp))) { free_pool(ctx, tmp); return rc; }
    CKS(cudaMemcpyAsync(grid_d, grid, sizeof(double) * n_sampling, cudaMemcpyDefault, s));
    CKS(cudaMemsetAsync(ctx->err_flag, 0, sizeof(int), s));
    CKS(sample_fill_units(P, off, unit_cell, unit_tri, s));
    CKS(sample_run(P, unit_cell, unit_tri, nu, n_sampling, grid_d, uu, out_d, ctx->err_flag, s));
    int eflag = 0;
    CKS(cudaMemcpyAsync(&eflag, ctx->err_flag, sizeof(int), cudaMemcpyDeviceToHost, s));
    if (out_host) CKS(cudaMemcpyAsync(out, out_d, sizeof(double) * nu * 7 * per, cudaMemcpyDeviceToHost, s));
    CKS(cudaStreamSynchronize(s));
#undef CKS
    free_pool(ctx, tmp);
    if (eflag == MOFEM_E_NEWTON) return ctx_fail(ctx, MOFEM_E_NEWTON, "Too many iterations!");
    if (eflag == MOFEM_E_ELEMENT) return ctx_fail(ctx, MOFEM_E_ELEMENT, "sampler: triangular incident elements are not supported (the reference raises ValueError)");
    if (eflag) return ctx_fail(ctx, MOFEM_E_ARG, "sampler: unsupported solver for a regular 4-node element");
    return MOFEM_OK;
}

// ---- weight functions of the overlay vertices -------------------------------------
int mofem_rho_eval(mofem_ctx *ctx, int64_t n_vert, int32_t n_mesh, const double *vert_xy, const int32_t *vert_elem,
                   int64_t n_elem, const int32_t *elem_nv, const double *elem_xy, const double *elem_w, double *rho,
                   int64_t *newton_updates, double *kernel_ms)
{
    if (!ctx) return MOFEM_E_ARG;
    if (n_vert < 0 || n_elem < 0 || n_mesh < 1 || n_mesh > 255) return ctx_fail(ctx, MOFEM_E_ARG, "mofem_rho_eval: bad sizes");
    if (n_vert > 0 && (!vert_xy || !vert_elem || !rho)) return ctx_fail(ctx, MOFEM_E_ARG, "mofem_rho_eval: null vertex array");
    if (n_elem > 0 && (!elem_nv || !elem_xy || !elem_w)) return ctx_fail(ctx, MOFEM_E_ARG, "mofem_rho_eval: null element array");
    if (newton_updates) *newton_updates = 0;
    if (kernel_ms) *kernel_ms = 0.0;
    if (n_vert == 0) return MOFEM_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    std::vector<void *> tmp;
    int rc;
    const double *vxy, *exy, *ew; const int32_t *vel, *env;
    if ((rc = upload(ctx, vert_xy, 2 * n_vert, &vxy, tmp)) || (rc = upload(ctx, vert_elem, n_vert * n_mesh, &vel, tmp)) ||
        (rc = upload(ctx, elem_nv, n_elem, &env, tmp)) || (rc = upload(ctx, elem_xy, 8 * n_elem, &exy, tmp)) ||
        (rc = upload(ctx, elem_w, 4 * n_elem, &ew, tmp))) { free_pool(ctx, tmp); return rc; }
    const bool out_host = !is_device_ptr(rho);
    double *rho_d = rho;
    unsigned long long *nw_d;
    int *err_d;
    if ((out_host && (rc = dev_alloc(ctx, &rho_d, n_vert * n_mesh, tmp))) || (rc = dev_alloc(ctx, &nw_d, 1, tmp)) ||
        (rc = dev_alloc(ctx, &err_d, 1, tmp))) { free_pool(ctx, tmp); return rc; }
#define CKS(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { free_pool(ctx, tmp); return ctx_fail(ctx, MOFEM_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); } } while (0)
    CKS(cudaMemsetAsync(nw_d, 0, sizeof(unsigned long long), s));
    CKS(cudaMemsetAsync(err_d, 0, sizeof(int), s));
    tic(ctx);
    CKS(rho_run(n_vert, n_mesh, vxy, vel, n_elem, env, exy, ew, rho_d, err_d, nw_d, s));
    CKS(cudaEventRecord(ctx->ev1, s));
    int eflag = 0;
    unsigned long long nw = 0;
    CKS(cudaMemcpyAsync(&eflag, err_d, sizeof(int), cudaMemcpyDeviceToHost, s));
    CKS(cudaMemcpyAsync(&nw, nw_d, sizeof(nw), cudaMemcpyDeviceToHost, s));
    if (out_host) CKS(cudaMemcpyAsync(rho, rho_d, sizeof(double) * n_vert * n_mesh, cudaMemcpyDeviceToHost, s));
    CKS(cudaStreamSynchronize(s));
    float ms = 0.f;
    CKS(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
#undef CKS
    free_pool(ctx, tmp);
    if (newton_updates) *newton_updates = (int64_t)nw;
    if (kernel_ms) *kernel_ms = ms;
    if (eflag == MOFEM_E_NEWTON) return ctx_fail(ctx, MOFEM_E_NEWTON, "Too many iterations!");
    if (eflag == MOFEM_E_ELEMENT) return ctx_fail(ctx, MOFEM_E_ELEMENT, "rho: incident element with neither 3 nor 4 corner vertices");
    if (eflag) return ctx_fail(ctx, MOFEM_E_ARG, "mofem_rho_eval: element index out of range");
    return MOFEM_OK;
}

// ---- overlay of the structured patch family ----------------------------------------
static bool overlay_args_ok(int32_t nx, int32_t ny, int32_t b_lo, int32_t b_hi, int32_t n_mesh)
{
    return nx >= 8 && ny >= 8 && nx % 8 == 0 && ny % 8 == 0 && b_lo >= 0 && b_lo < b_hi && b_hi <= ny / 8 && (n_mesh == 2 || n_mesh == 3);
}

int mofem_overlay_patch_sizes(int32_t nx, int32_t ny, int32_t b_lo, int32_t b_hi, int32_t n_mesh, int64_t sizes[5])
{
    if (!sizes || !overlay_args_ok(nx, ny, b_lo, b_hi, n_mesh)) return MOFEM_E_ARG;
    overlay_sizes(nx, ny, b_lo, b_hi, n_mesh, sizes);
    return MOFEM_OK;
}

int mofem_overlay_patch_family(mofem_ctx *ctx, int32_t nx, int32_t ny, int32_t b_lo, int32_t b_hi, int32_t n_mesh, const double *X0,
                               const double *X1, int32_t *cell_elem, int32_t *cell_kind, int64_t *cell_tri_ptr, double *tri_xy,
                               double *tri_rho, double *kernel_ms)
{
    if (!ctx) return MOFEM_E_ARG;
    if (!overlay_args_ok(nx, ny, b_lo, b_hi, n_mesh)) return ctx_fail(ctx, MOFEM_E_ARG, "mofem_overlay_patch_family: nx, ny must be multiples of 8, 0 <= b_lo < b_hi <= ny / 8, n_mesh 2 or 3");
    if (!X0 || !X1 || !cell_elem || !cell_kind || !cell_tri_ptr || !tri_xy || !tri_rho) return ctx_fail(ctx, MOFEM_E_ARG, "mofem_overlay_patch_family: null array");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    int64_t sz[5];
    overlay_sizes(nx, ny, b_lo, b_hi, n_mesh, sz);
    const int64_t n_cell = sz[4], n_tri = sz[3];
    if (n_cell >= ((int64_t)1 << 31)) return ctx_fail(ctx, MOFEM_E_LIMIT, "mofem_overlay_patch_family: more than 2^31 cells");
    std::vector<void *> tmp;
    int rc;
    const double *x0d, *x1d;
    if ((rc = upload(ctx, X0, 2 * (int64_t)(nx + 1) * (ny + 1), &x0d, tmp)) ||
        (rc = upload(ctx, X1, 72 * (int64_t)(nx / 8) * (ny / 8), &x1d, tmp))) { free_pool(ctx, tmp); return rc; }
    int32_t *ce = cell_elem, *ck = cell_kind; int64_t *tp = cell_tri_ptr; double *txy = tri_xy, *trho = tri_rho;
    const bool h_ce = !is_device_ptr(cell_elem), h_ck = !is_device_ptr(cell_kind), h_tp = !is_device_ptr(cell_tri_ptr),
               h_xy = !is_device_ptr(tri_xy), h_rho = !is_device_ptr(tri_rho);
    if ((h_ce && (rc = dev_alloc(ctx, &ce, n_cell * n_mesh, tmp))) || (h_ck && (rc = dev_alloc(ctx, &ck, n_cell, tmp))) ||
        (h_tp && (rc = dev_alloc(ctx, &tp, n_cell + 1, tmp))) || (h_xy && (rc = dev_alloc(ctx, &txy, 6 * n_tri, tmp))) ||
        (h_rho && (rc = dev_alloc(ctx, &trho, 3 * n_tri * n_mesh, tmp)))) { free_pool(ctx, tmp); return rc; }
#define CKS(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { free_pool(ctx, tmp); return ctx_fail(ctx, MOFEM_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); } } while (0)
    tic(ctx);
    CKS(overlay_run(nx, ny, b_lo, b_hi, n_mesh, x0d, x1d, ce, ck, tp, txy, trho, s));
    CKS(cudaEventRecord(ctx->ev1, s));
    if (h_ce) CKS(cudaMemcpyAsync(cell_elem, ce, sizeof(int32_t) * n_cell * n_mesh, cudaMemcpyDeviceToHost, s));
    if (h_ck) CKS(cudaMemcpyAsync(cell_kind, ck, sizeof(int32_t) * n_cell, cudaMemcpyDeviceToHost, s));
    if (h_tp) CKS(cudaMemcpyAsync(cell_tri_ptr, tp, sizeof(int64_t) * (n_cell + 1), cudaMemcpyDeviceToHost, s));
    if (h_xy) CKS(cudaMemcpyAsync(tri_xy, txy, sizeof(double) * 6 * n_tri, cudaMemcpyDeviceToHost, s));
    if (h_rho) CKS(cudaMemcpyAsync(tri_rho, trho, sizeof(double) * 3 * n_tri * n_mesh, cudaMemcpyDeviceToHost, s));
    CKS(cudaStreamSynchronize(s));
    float ms = 0.f;
    CKS(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
#undef CKS
    free_pool(ctx, tmp);
    if (kernel_ms) *kernel_ms = ms;
    return MOFEM_OK;
}

int mofem_set_profile(mofem_ctx *ctx, int on)
{
    if (!ctx) return MOFEM_E_ARG;
    ctx->profile = on < 0 ? 0 : on;
    return MOFEM_OK;
}

int mofem_set_option(mofem_ctx *ctx, const char *name, int64_t value)
{
    if (!ctx || !name) return MOFEM_E_ARG;
    if (!strcmp(name, "l2_persist")) { ctx->l2_persist = (int)value; return MOFEM_OK; }
    if (!strcmp(name, "l2_window")) { ctx->l2_window = (int)value; return MOFEM_OK; }
    if (!strcmp(name, "spmv_ring")) { ctx->spmv_ring = (int)value; return MOFEM_OK; }
    if (!strcmp(name, "floor_rtol_1e12")) { ctx->floor_rtol = (double)value * 1e-12; return MOFEM_OK; }
    if (!strcmp(name, "pdl")) { ctx->pdl = (int)value; return MOFEM_OK; }
    if (!strcmp(name, "cell_kernel")) { ctx->cell_kernel = (int)value; return MOFEM_OK; }   // takes effect at the next mofem_symbolic
    return ctx_fail(ctx, MOFEM_E_ARG, std::string("mofem_set_option: unknown option ") + name);
}

int mofem_get_pcg_profile(mofem_ctx *ctx, double ms[3], int64_t *iterations)
{
    if (!ctx) return MOFEM_E_ARG;
    if (ms) for (int i = 0; i < 3; i++) ms[i] = ctx->prof_ms[i];
    if (iterations) *iterations = (int64_t)ctx->prof_ms[3];
    return MOFEM_OK;
}

int mofem_get_timing(mofem_ctx *ctx, double ms[6], int64_t launches[4])
{
    if (!ctx) return MOFEM_E_ARG;
    if (ms) for (int i = 0; i < 6; i++) ms[i] = ctx->ms[i];
    if (launches) for (int i = 0; i < 4; i++) launches[i] = ctx->launches[i];
    return MOFEM_OK;
}

}  // extern "C"
