/*
 * mofem_b200.h -- C-ABI of the B200-native stiffness-assembly + PCG path.
 *
 * Drop-in boundary for the hot path of junbinhuang/MOFEM (pure Python): these
 * are the entry points a ctypes binding inside the reference's
 * linearSolvers/AMORE.py / traditionalElement.py would call (see
 * INTEGRATION.md for that binding).  Plain pointers and sizes only.
 *
 * Reference interfaces replaced (paths relative to the reference root):
 *   mofem_set_problem      <- the solvers' inputs: inputData[3..6]
 *                             (inputFunctions/inputF.py:165,204), polygons /
 *                             overlappingMesh (computGeometry/planeSweepMeshes.py:354,
 *                             meshTools/toDC.py:48-72), flattened by mofem_b200/flatten.py
 *   mofem_symbolic         <- scipy.sparse.coo_matrix(...).tocsr() pattern
 *                             (AMORE.py:129,404,827; traditionalElement.py:54,204,357)
 *   mofem_assemble         <- the element / cell loops AMORE.py:42-123, 307-398, 729-821,
 *                             traditionalElement.py:36-48,188-198,339-351 and the
 *                             arithmetic of elementLibrary/stiffnessMatrix.py,
 *                             shapeFunction.py, invShapeFunction.py,
 *                             otherFunctions/numericalIntegration.py
 *   mofem_set_system       <- constraint elimination AMORE.py:213-255 (clones :632-674,
 *                             1055-1097; traditionalElement.py:94-136 ...)
 *   mofem_solve            <- spsolve(Kglo,fglo) AMORE.py:264,683,1106;
 *                             traditionalElement.py:146,296,461 (Jacobi-PCG instead of SuperLU)
 *   mofem_energy           <- 0.5*u^T K u, AMORE.py:278,701,1124
 *
 * Error convention: every call returns MOFEM_OK (0) or a negative code;
 * mofem_last_error() gives the text.  CUDA errors are never swallowed.
 * The reference's own failures map to codes: MOFEM_E_NEWTON = AssertionError
 * "Too many iterations!" (invShapeFunction.py:68,102,136), MOFEM_E_ELEMENT =
 * ValueError "Wrong element numbering!" (AMORE.py:51,74).
 *
 * Threading: a context is single-threaded and synchronous (like the reference);
 * each context owns one CUDA stream on one device.
 */
#ifndef MOFEM_B200_H
#define MOFEM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOFEM_OK 0
#define MOFEM_E_CUDA -1      /* CUDA runtime error (text in mofem_last_error) */
#define MOFEM_E_ARG -2       /* bad argument / call order */
#define MOFEM_E_NEWTON -3    /* inverse isoparametric map did not converge in 20 updates */
#define MOFEM_E_ELEMENT -4   /* element node count not supported by the selected solver */
#define MOFEM_E_NOMEM -5
#define MOFEM_E_LIMIT -6     /* problem exceeds an index-width limit of this build */
#define MOFEM_E_NOCONV -7    /* PCG hit maxit before reaching rtol */

typedef struct mofem_ctx mofem_ctx;

/* Flattened problem (SoA).  Pointers may be HOST or DEVICE pointers (detected
 * with cudaPointerGetAttributes).  Host arrays are copied to the device inside
 * mofem_set_problem; device arrays are BORROWED (zero-copy) and must stay alive
 * and unchanged until the next mofem_set_problem / mofem_destroy.
 * Cell order is the reference's emission order: regular elements first, then
 * overlay cells in `polygons` order. */
typedef struct {
    int64_t n_node, n_elem, n_cell, n_tri;
    int32_t n_mesh, n_mat;
    int32_t solver;              /* 1 lowerOrderAMORE/FE, 2 quadraticAMORE/FE, 3 quadraticICMAMORE / ICMFE */
    int32_t _pad;
    const double *node_xy;       /* [n_node][2], mid-nodes completed by the getCoord rule */
    const int32_t *elem_nodes;   /* [n_elem][9], -1 padded */
    const int32_t *elem_nn;      /* [n_elem] in {3,4,6,9} */
    const int32_t *elem_mat;     /* [n_elem] */
    const double *mat_D;         /* [n_mat][3][3] */
    const int32_t *cell_elem;    /* [n_cell][n_mesh] global element id per mesh slot, -1 = none */
    const int32_t *cell_kind;    /* [n_cell] 0 regular element rule, 1 triangulated overlay cell */
    const int64_t *cell_tri_ptr; /* [n_cell+1] */
    const double *tri_xy;        /* [n_tri][3][2] */
    const double *tri_rho;       /* [n_tri][3][n_mesh] */
    const double *tri_mid;       /* [n_tri][3][2] mid-edge points of curved triangles, or NULL */
    const uint8_t *tri_curved;   /* [n_tri] or NULL */
} mofem_problem_t;

typedef struct {
    int64_t n_dof;        /* 2*n_node */
    int64_t nnz;          /* scalar CSR entries (explicit zeros kept) */
    int64_t n_block;      /* emitted element/cell blocks ("tasks") */
    int64_t n_coo;        /* COO triplets the reference would emit */
    int64_t n_pair;       /* 2x2 node-pair sub-blocks integrated */
    int64_t newton_updates; /* total Newton updates of the last mofem_assemble */
    int64_t n_eval;       /* (quadrature point, element) evaluations of the last mofem_assemble */
} mofem_info_t;

typedef struct {
    int32_t iterations;
    int32_t converged;    /* 0 no; 1 true residual <= rtol; 2 recurrence residual <= rtol and the true residual
                             stagnated at the fp64 attainable accuracy (reported in relres_true), below
                             max(rtol, "floor_rtol_1e12" option = 1e-8) -- a stagnation above that bound is 0 */
    double relres_true;   /* ||b-Ax|| / ||b|| recomputed at exit */
    double relres_rec;    /* recurrence residual at exit */
    int32_t restarts;     /* times the true residual replaced the recurrence residual */
    int32_t _pad;
} mofem_solve_info_t;

/* lifecycle ---------------------------------------------------------- */
int mofem_create(mofem_ctx **ctx, int device);
void mofem_destroy(mofem_ctx *ctx);
const char *mofem_last_error(const mofem_ctx *ctx);
const char *mofem_version(void);
/* CUDA stream the context launches on (cudaStream_t as void*), for event timing. */
void *mofem_stream(mofem_ctx *ctx);

/* assembly ----------------------------------------------------------- */
int mofem_set_problem(mofem_ctx *ctx, const mofem_problem_t *p);
int mofem_symbolic(mofem_ctx *ctx);   /* integer phase: CSR pattern + reduction plan */
int mofem_assemble(mofem_ctx *ctx);   /* integrate all blocks + deterministic reduce into CSR values */
int mofem_get_info(mofem_ctx *ctx, mofem_info_t *info);
/* Scalar CSR, identical in pattern to scipy coo->csr.  Output pointers may be host or device; any may be NULL. */
int mofem_get_csr(mofem_ctx *ctx, int32_t *indptr, int32_t *indices, double *data);
/* Parity/debug: the COO values (Vglo) in the reference's emission order, host or device pointer. */
int mofem_get_coo_values(mofem_ctx *ctx, double *V);

/* solve -------------------------------------------------------------- */
/* rhs: global force vector f (n_dof); fixed: 1 where the DOF is prescribed; fixed_value: prescribed values
 * (ignored where fixed==0).  Builds b = (f - K*u_fixed) on free DOFs (AMORE.py:227). */
int mofem_set_system(mofem_ctx *ctx, const double *rhs, const uint8_t *fixed, const double *fixed_value);
/* Jacobi-preconditioned CG on the free-free block; u (n_dof, host or device) gets fixed values + solution. */
int mofem_solve(mofem_ctx *ctx, double rtol, int32_t maxit, double *u, mofem_solve_info_t *info);
/* 0.5 * u^T K u with the full (unconstrained) K and the u of the last mofem_solve. */
int mofem_energy(mofem_ctx *ctx, double *energy);
/* y = K x with the assembled matrix (host or device pointers) -- used by tests and the multi-GPU driver. */
int mofem_spmv(mofem_ctx *ctx, const double *x, double *y);

/* one-call convenience used for end-to-end timing: host buffers in, host displacement out. */
int mofem_assemble_solve(mofem_ctx *ctx, const mofem_problem_t *p, const double *rhs, const uint8_t *fixed,
                         const double *fixed_value, double rtol, int32_t maxit, double *u,
                         double *energy, mofem_solve_info_t *info);

/* timing of the last calls, milliseconds measured with CUDA events on the context's stream:
 * [0] symbolic, [1] integrate kernels, [2] numeric reduce, [3] solve, [4] h2d, [5] d2h.
 * counts: launches[0..3] = kernel launches in symbolic / integrate / numeric / solve. */
int mofem_get_timing(mofem_ctx *ctx, double ms[6], int64_t launches[4]);

/* row-partitioned solve, one process per GPU ------------------------------------------------------------
 * Each rank holds a LOCAL problem (mofem_b200/partition.py): every cell that touches a node it owns, nodes
 * renumbered [owned | ghosts grouped by owner rank].  Assembly needs no exchange (rows of owned nodes are
 * complete); the PCG exchanges the ghost entries of the search direction once per iteration (ncclSend/Recv)
 * and all-reduces the dot products (NCCL over NVLink).  mofem_solve / mofem_energy / mofem_spmv then act on the
 * owned rows; vectors passed in and out are LOCAL (owned + ghost entries). */
int mofem_dist_unique_id(uint8_t id[128]);                 /* rank 0: ncclGetUniqueId, to be broadcast by the caller */
int mofem_dist_init(mofem_ctx *ctx, const uint8_t id[128], int32_t rank, int32_t world);
/* send_ptr/recv_ptr: [n_neighbor+1] in nodes; send_idx: local ids of owned nodes to send, neighbour by neighbour;
 * the ghosts received from neighbour k are local nodes n_owned_node + recv_ptr[k] .. + recv_ptr[k+1]. */
int mofem_set_halo(mofem_ctx *ctx, int64_t n_owned_node, int32_t n_neighbor, const int32_t *neighbor_rank,
                   const int64_t *send_ptr, const int32_t *send_idx, const int64_t *recv_ptr);
int mofem_clear_halo(mofem_ctx *ctx);
/* Optional NVLink peer-memory transport for the PCG loop (replaces the per-iteration NCCL calls): every rank exports an
 * exchange buffer (CUDA IPC), imports everybody else's, and from then on mofem_solve pushes halo entries and dot-product
 * partials straight into the peers' buffers from inside its kernels.  Call after mofem_set_halo, on every rank:
 *   mofem_p2p_export -> all-gather {handle, n_owned_node, recv offsets} by the caller -> mofem_p2p_import.
 * send_remote_off[k] = offset (nodes) of this rank's entries inside neighbour k's ghost segment (its recv_ptr for us).
 * A new mofem_set_halo / mofem_clear_halo drops the transport; mofem_solve then falls back to NCCL. */
int mofem_p2p_export(mofem_ctx *ctx, uint8_t handle[64], int64_t *n_ghost);
int mofem_p2p_import(mofem_ctx *ctx, const uint8_t *handles /* world x 64 */, const int64_t *peer_n_owned /* world */,
                     const int64_t *send_remote_off /* n_neighbor */);
int mofem_p2p_close(mofem_ctx *ctx);

/* Sampler (post-processing of the solution): replaces plotTools/plotResults.py:getGraphDataFE:137 / getGraphDataOFE:15 as
 * driven by outputFunctions/writeSolution.py:writeSampleData2D:41.  A "unit" is a regular element / untriangulated cell, or
 * one integration triangle of an overlay cell, in the problem's cell order.  out[unit][7][nS][nS] = X, Y, U, V, Sxx, Syy, Sxy
 * at the parent-grid points grid[i], grid[j] (np.linspace(-1, 1, nS) in the reference).  u: displacement (host or device,
 * NULL = the last mofem_solve).  out == NULL only returns *n_unit.  solver: 2 or 3 (stress recovery of regular 4-node
 * elements: plain D B u or the ICM form). */
int mofem_sample(mofem_ctx *ctx, const double *u, int32_t n_sampling, int32_t solver, const double *grid, double *out,
                 int64_t out_capacity_units, int64_t *n_unit);

/* Weight functions of the overlay vertices: replaces the per-vertex body of
 * computGeometry/planeSweepMeshes.py:rhoCalculation:991-1030.  vert_xy [n_vert][2]; vert_elem [n_vert][n_mesh] = incident element
 * (index into the elem_* tables) of mesh slot m of the face that first visits the vertex, -1 = none; elem_nv [n_elem] = 3 or 4
 * corner vertices (element.vertex), elem_xy [n_elem][4][2] their coordinates, elem_w [n_elem][4] their weights (1, or 0 on a
 * mesh boundary inside the domain, planeSweepMeshes.py:434-446).  rho [n_vert][n_mesh] (host or device, like every input):
 * interpolated weight per slot, slots >= 1 times 9, row normalised by its sum.  newton_updates / kernel_ms may be NULL.
 * Errors: MOFEM_E_NEWTON ("Too many iterations!", invShapeFunction.py:68), MOFEM_E_ELEMENT (the ValueError of :1021). */
int mofem_rho_eval(mofem_ctx *ctx, int64_t n_vert, int32_t n_mesh, const double *vert_xy, const int32_t *vert_elem,
                   int64_t n_elem, const int32_t *elem_nv, const double *elem_xy, const double *elem_w, double *rho,
                   int64_t *newton_updates, double *kernel_ms);

/* Overlay of the structured patch family on the GPU (mesh 0 = nx x ny quadrilaterals, nx and ny multiples of 8;
 * meshes >= 1 = disjoint 5 x 5 patches, one per 8 x 8 period, straight patch borders): replaces, for this family, the host
 * pipeline computGeometry/planeSweepMeshes.py:meshOverlay:354 + rhoCalculation:991 + triangulation.py:polygonTriangulation +
 * the flattening of the object graph.  Inputs: X0 [ny+1][nx+1][2] mesh-0 node coordinates, X1 [ny/8][nx/8][6][6][2] patch node
 * coordinates; [b_lo, b_hi) = patch rows to emit (plus the mesh-0 element row below them), n_mesh = 2, or 3 (patch (a, b) in mesh
 * 1 + (a + b) % 2).  Outputs = the cell / triangle arrays of mofem_problem_t, sized by mofem_overlay_patch_sizes
 * (sizes = {regular cells, ring cells, two-element cells, triangles, cells}); all pointers host or device.  The result is
 * bit-identical to mofem_b200/overlay_gen.py, which is pinned against the reference's own sweep. */
int mofem_overlay_patch_sizes(int32_t nx, int32_t ny, int32_t b_lo, int32_t b_hi, int32_t n_mesh, int64_t sizes[5]);
int mofem_overlay_patch_family(mofem_ctx *ctx, int32_t nx, int32_t ny, int32_t b_lo, int32_t b_hi, int32_t n_mesh, const double *X0,
                               const double *X1, int32_t *cell_elem, int32_t *cell_kind, int64_t *cell_tri_ptr, double *tri_xy,
                               double *tri_rho, double *kernel_ms);

/* Per-kernel profile of the PCG loop: on = 1 records CUDA events around the kernels of every iteration of mofem_solve on the
 * context's stream, on = N > 1 around those of every N-th iteration (the others keep their launch overlap), 0 = off.  ms[0..1] = summed device time of the SpMV (incl. halo exchange and the
 * all-reduce when row-partitioned) / the vector kernel over the sampled iterations of the last mofem_solve, ms[2] = 0,
 * *iterations = iterations profiled. */
int mofem_set_profile(mofem_ctx *ctx, int on);
/* Tuning switches (defaults in brackets): "spmv_ring" [1] PCG SpMV streams the matrix through shared memory with the TMA
 * (0: row kernel); "pdl" [1] programmatic dependent launch between the two kernels of a PCG iteration (single GPU and peer
 * transport, with the ring); "floor_rtol_1e12" [10000] = 1e12 x the largest TRUE relative residual mofem_solve still accepts
 * (converged = 2) when restarts no longer reduce it -- above max(rtol, that) the solve returns MOFEM_E_NOCONV; "l2_persist" [0] / "l2_window" [5] L2 persistence window over the first
 * N PCG vectors during mofem_solve (measured neutral at 1 M cells, harmful at 10 M); "cell_kernel" [1] integrates overlay
 * cells whose incident elements are all 4-node quads (<= 3 of them, <= 5 straight triangles) per CELL -- one evaluation per
 * (point, element) -- instead of per block (0: thread-per-block kernels for everything; bit-identical results, 2.2x slower on
 * the synthetic families -- see DESIGN.md; set before mofem_symbolic). */
int mofem_set_option(mofem_ctx *ctx, const char *name, int64_t value);
int mofem_get_pcg_profile(mofem_ctx *ctx, double ms[3], int64_t *iterations);

#ifdef __cplusplus
}
#endif
#endif
