/* mofem_oracle.h -- TEST INFRASTRUCTURE (parity oracle); see mofem_oracle.c. */
#ifndef MOFEM_ORACLE_H
#define MOFEM_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Flattened problem; same field meaning as mofem_problem_t in
 * include/mofem_b200.h (kept separate so the oracle shares no product code). */
typedef struct {
    int64_t n_node, n_elem, n_cell, n_tri;
    int32_t n_mesh, n_mat;
    int32_t solver;            /* 1 lowerOrder, 2 quadratic, 3 quadratic + ICM */
    int32_t _pad;
    const double *node_xy;     /* [n_node][2] */
    const int32_t *elem_nodes; /* [n_elem][9], -1 padded */
    const int32_t *elem_nn;    /* [n_elem] 3|4|6|9 */
    const int32_t *elem_mat;   /* [n_elem] */
    const double *mat_D;       /* [n_mat][3][3] */
    const int32_t *cell_elem;  /* [n_cell][n_mesh], -1 = none */
    const int32_t *cell_kind;  /* [n_cell] 0 regular, 1 triangulated */
    const int64_t *cell_tri_ptr; /* [n_cell+1] */
    const double *tri_xy;      /* [n_tri][3][2] */
    const double *tri_rho;     /* [n_tri][3][n_mesh] */
    const double *tri_mid;     /* [n_tri][3][2] or NULL */
    const uint8_t *tri_curved; /* [n_tri] or NULL */
} orc_problem;

int orc_gauss_tri(int n, double *pos, double *wei);
int orc_gauss_quad(int n, double *pos, double *wei);
void orc_shape(const double *coord, int n, double r, double s, double *Nmat, double *Bmat, double *Jacobian);
void orc_inv_tri(const double *c, const double *xy, double *iso);
int orc_inv_newton(const double *coord, int n, const double *xy, double *iso);
int orc_element_stiffness(const double *coord, int n, const double *D, int quad4_rule, double *K);
int orc_amore_nint(int quadratic, int n1, int n2);
int orc_amore_block(int quadratic, const double *coordTri, int nt, const double *D,
                    const double *coord1, int n1, const double *rho1,
                    const double *coord2, int n2, const double *rho2, double *Kloc, int *newton);
int orc_material_matrix(double E, double nu, int problemType, double *D);
int64_t orc_coo_count(const orc_problem *p);
int orc_assemble_coo(const orc_problem *p, int64_t *I, int64_t *J, double *V, int64_t *cell_off,
                     int64_t *newton_total, int n_threads);
int64_t orc_coo_tocsr(int64_t n_row, int64_t nnz, const int64_t *I, const int64_t *J, const double *V,
                      int64_t *indptr, int64_t *indices, double *data);

/* planeSweepMeshes.rhoCalculation:991-1030 on flat arrays (see mofem_oracle.c). */
int orc_rho_eval(int64_t n_vert, int32_t n_mesh, const double *vert_xy, const int32_t *vert_elem,
                 const int32_t *elem_nv, const double *elem_xy, const double *elem_w, double *rho,
                 int64_t *newton_total);

#ifdef __cplusplus
}
#endif
#endif
