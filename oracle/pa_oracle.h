/* TEST INFRASTRUCTURE ONLY — CPU restatement (plain C, scalar, no FMA contraction) of the
 * reference's algorithm for the hot path of SURVEY.md §8.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg may load this; the product (libb200pa.so) never does.
 *
 * Parity status: PINNED.  Every function below is checked bit-for-bit (or to 1 ulp-level
 * 1e-15 where noted) against outputs of the unmodified reference run in the build container
 * (oracle/_ref/ref_driver, fixtures under tests/golden/, generator tests/golden/make_golden.py).
 *
 * All paths are /root/reference-relative.  Layouts are the reference's:
 *   L-vector f64[ndofs]; E-vector f64[D,D,D,NE] (x fastest); q-data [Q,Q,Q,ncomp,NE];
 *   B,G column-major [Q,D] (b(q,d) = B[q + Q*d]); J [Q,Q,Q,3,3,NE] with J(q,row,col,e).
 */
#ifndef B200PA_ORACLE_H
#define B200PA_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

/* fem/restriction.cpp:66-106 — CSR (offsets, indices) from the E→L gather map */
void orc_restriction_tables(int ne, int nd, int ndofs, const int *gather_map, int *offsets, int *indices);
/* fem/restriction.cpp:109-129 */
void orc_restrict_mult(int ne, int nd, const int *gather_map, const double *x, double *y);
/* fem/restriction.cpp:152-186 (ADD=false) and :196-221 (abs != 0) */
void orc_restrict_mult_transpose(int ndofs, const int *offsets, const int *indices, const double *xE,
                                 double *yL, int abs);

/* fem/integ/bilininteg_diffusion_kernels.cpp:243-367, scalar-coefficient branch :349-362.
 * nc = number of entries of C (1 → constant). D is [Q,Q,Q,6,NE]. */
void orc_diffusion_setup(int Q1D, int NE, const double *W, const double *J, const double *C, long nc, double *D);
/* fem/integ/bilininteg_mass_pa.cpp:62-78 (map_type VALUE) */
void orc_mass_setup(int NQ, int NE, const double *W, const double *detJ, const double *C, long nc, double *v);

/* fem/integ/bilininteg_diffusion_kernels.hpp:989-1214 (symmetric) : yE += G^T D G xE */
void orc_diffusion_apply(int NE, int D1D, int Q1D, const double *B, const double *G, const double *D,
                         const double *xE, double *yE);
/* fem/integ/bilininteg_mass_kernels.hpp:807-1033 : yE += B^T v B xE */
void orc_mass_apply(int NE, int D1D, int Q1D, const double *B, const double *v, const double *xE, double *yE);
/* fem/integ/bilininteg_diffusion_kernels.hpp:369-484 */
void orc_diffusion_diag(int NE, int D1D, int Q1D, const double *B, const double *G, const double *D, double *dE);
/* fem/integ/bilininteg_mass_kernels.hpp:324-408 */
void orc_mass_diag(int NE, int D1D, int Q1D, const double *B, const double *v, double *dE);

/* The L→L operator: fem/bilinearform_ext.cpp:487-564 (gather, localY=0, Σ AddMultPA, scatter).
 * pa_diff / pa_mass may be NULL to skip that integrator. workE: 2*nd*NE doubles. */
typedef struct
{
   int NE, D1D, Q1D, ndofs;
   const int *gather_map, *offsets, *indices;
   const double *B, *G, *pa_diff, *pa_mass;
   int n_ess; const int *ess;          /* essential true dofs (linalg/operator.cpp:511-526) */
} orc_operator;
void orc_op_mult(const orc_operator *op, const double *x, double *y, double *workE);
/* fem/bilinearform_ext.cpp:370-454 : diag = AbsMultTranspose(Σ AssembleDiagonalPA) */
void orc_op_diag(const orc_operator *op, double *diag, double *workE);
/* linalg/operator.cpp:586-646 (DIAG_ONE). work: ndofs doubles (+ workE as above) */
void orc_constrained_mult(const orc_operator *op, const double *x, double *y, double *work, double *workE);
/* linalg/operator.cpp:559-584 : b -= A w (w = x on ess, 0 elsewhere); b[ess] = x[ess] */
void orc_eliminate_rhs(const orc_operator *op, const double *x, double *b, double *work2n, double *workE);

/* linalg/solvers.cpp:401-425 */
int orc_jacobi_setup(int n, const double *diag, int n_ess, const int *ess, double damping, double *dinv);
/* linalg/solvers.cpp:427-453 (iterative_mode == false) */
void orc_jacobi_mult(int n, const double *dinv, const double *r, double *z);
/* linalg/vector.cpp:1079-1152 + general/reducers.hpp:587-590 : strict left-to-right sum */
double orc_dot(long n, const double *a, const double *b);

/* linalg/solvers.cpp:869-1050 with OperatorJacobiSmoother as preconditioner and the
 * ConstrainedOperator of `op` as operator; iterative_mode = true.
 * Returns final_iter; *converged, *final_norm as the reference sets them; norms[i] = (B r, r)
 * after iteration i (i = 0..final_iter) when norms != NULL (size max_iter+1). */
/* OperatorChebyshevSmoother (linalg/solvers.cpp:455-657), PowerMethod (linalg/operator.cpp:871-928) */
int orc_chebyshev_coeffs(int order, double max_eig, double *coeffs);
void orc_chebyshev_mult(const orc_operator *op, const double *dinv, int order, const double *coeffs, const double *x,
                        double *y, double *work, double *workE);
double orc_power_method(const orc_operator *op, const double *dinv, double *v0, int num_steps, double tolerance);
int orc_pcg_prec(const orc_operator *op, const double *dinv, int cheb_order, double max_eig, const double *b, double *x,
                 double rel_tol, double abs_tol, int max_iter, int *converged, double *final_norm, double *norms);
int orc_pcg(const orc_operator *op, const double *dinv, const double *b, double *x,
            double rel_tol, double abs_tol, int max_iter, int *converged, double *final_norm,
            double *norms);

/* fem/qinterp/eval.hpp:131-193 (vdim 1): yq[Q,Q,Q,NE] = (B⊗B⊗B) xE */
void orc_qvalues(int NE, int D1D, int Q1D, const double *B, const double *xE, double *yq);
/* fem/qinterp/grad.hpp:233-374 (vdim 1, byVDIM, GRAD_PHYS): gq[3,Q,Q,Q,NE] */
void orc_qphysgrad(int NE, int D1D, int Q1D, const double *B, const double *G, const double *J,
                   const double *xE, double *gq);
/* fem/integ/lininteg_domain_kernels.hpp:164-298 (vdim 1, map VALUE): bE += B^T (W f detJ) */
void orc_domain_lf(int NE, int D1D, int Q1D, const double *B, const double *detJ, const double *W,
                   const double *f, long nf, double *bE);

#ifdef __cplusplus
}
#endif
#endif
