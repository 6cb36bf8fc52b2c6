/*
 * b200pa.h — C ABI of libb200pa.so: the B200 (sm_100a, FP64) partial-assembly hot path behind
 * the Pennes-bioheat / electrostatic solves of an MFEM-based cardiac-ablation solver.
 *
 * This is the drop-in boundary.  Every entry point names the reference
 * interface it replaces (paths relative to the reference tree, stock MFEM 4.9.1-dev).
 * Plain pointers and sizes only; no C++ or torch types.  The reference-side bindings a
 * maintainer adds (C++ subclasses of mfem::DiffusionIntegrator / MassIntegrator / Operator /
 * Solver forwarding to these functions) are shown in INTEGRATION.md and implemented in
 * cardiac-ablation-ecm2_b200/host/mfem_b200pa.hpp (compiled against the unmodified reference by
 * oracle/Makefile, checked on the GPU by tests/test_gpu_mfem_shim.py).
 *
 * Conventions
 *   - every function returns 0 on success, nonzero on error; b200pa_last_error() returns the
 *     message for the calling thread (≙ MFEM_VERIFY/MFEM_ABORT, general/error.hpp:26-64).
 *     There is NO CPU fallback: without a CUDA device every compute entry point fails.
 *   - layouts are the reference's: L-vector f64[ndofs]; E-vector f64[D,D,D,NE],
 *     x fastest (fem/restriction.cpp:116); q-data f64[Q,Q,Q,ncomp,NE]
 *     (fem/integ/bilininteg_diffusion_kernels.hpp:1009); B,G column-major [Q,D]
 *     (fem/fe/fe_base.cpp:2654-2655); J f64[Q,Q,Q,3,3,NE] with J(q,row,col,e)
 *     (fem/integ/bilininteg_diffusion_kernels.cpp:254); indices int32.
 *   - "dev" pointers are device pointers on the context's GPU; "any" pointers may be host or
 *     device (detected with cudaPointerGetAttributes) and are copied if they are host memory.
 *   - kernels are enqueued on the context's stream and are asynchronous unless stated.
 *   - supported: 3-D hexahedra, H1 scalar (vdim 1) spaces, orders 1..6 with the reference's
 *     default rules (D1D = p+1, Q1D = p+2) — (D1D,Q1D) in {(2,3),(3,4),(4,5),(5,6),(6,7),(7,8)};
 *     anything else is rejected (the reference would fall back to an unspecialised kernel,
 *     fem/kernel_dispatch.hpp:138-153).
 */
#ifndef B200PA_H
#define B200PA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200PA_VERSION 100

typedef struct b200pa_ctx_s *b200pa_ctx;      /* one per GPU / host thread (≙ mfem::Device)      */
typedef struct b200pa_space_s *b200pa_space;  /* ElementRestriction + DofToQuad + GeometricFactors */
typedef struct b200pa_form_s *b200pa_form;    /* PABilinearFormExtension (+ ConstrainedOperator)  */
typedef struct b200pa_comm_s *b200pa_comm;    /* shared-dof halo exchange + allreduce (NCCL)      */

int b200pa_version(void);
const char *b200pa_last_error(void);
/* number of kernels this library has launched in the calling process (bench.py's gpu_launches) */
long long b200pa_launch_count(void);

/* ------------------------------------------------------------------ context */
/* stream: a cudaStream_t to enqueue on, or NULL to create an owned stream. */
int b200pa_ctx_create(int device, void *stream, b200pa_ctx *out);
int b200pa_ctx_destroy(b200pa_ctx ctx);
int b200pa_ctx_sync(b200pa_ctx ctx);                       /* ≙ MFEM_STREAM_SYNC */
void *b200pa_ctx_stream(b200pa_ctx ctx);
/* device memory for callers that have no CUDA runtime of their own (a CPU build of the host
 * library): ≙ Memory<T>::New with MemoryType::DEVICE (general/mem_manager.hpp) */
int b200pa_malloc(b200pa_ctx ctx, size_t bytes, void **out_dev);
int b200pa_free(b200pa_ctx ctx, void *dev);
int b200pa_memset(b200pa_ctx ctx, void *dev, int value, size_t bytes);
/* page-locked host memory placed on the NUMA node the context's GPU is attached to (≙ MemoryType::HOST_PINNED,
 * general/mem_manager.cpp; the reference's cudaMallocHost leaves placement to the first-touch policy of whichever core
 * the rank runs on).  For the *_host entry points below at one rank per GPU: the vectors cross PCIe on every call and
 * should not cross the socket interconnect as well.  b200pa_host_node: node the block sits on, -1 = kernel default. */
int b200pa_host_alloc(b200pa_ctx ctx, size_t bytes, void **out_host);
int b200pa_host_free(b200pa_ctx ctx, void *host);   /* ctx may be NULL (the block remembers its device); waits for the device */
int b200pa_host_node(const void *host);
/* device-to-device copy of n doubles on the context's stream (Vector::operator=, linalg/vector.cpp:203-240) */
int b200pa_copy(b200pa_ctx ctx, long long n, const double *src_dev, double *dst_dev);
/* synchronous copies on the context's stream (≙ Vector::HostRead / Vector::Write of a device vector) */
int b200pa_ctx_upload(b200pa_ctx ctx, void *dst_dev, const void *src_host, size_t bytes);
int b200pa_ctx_download(b200pa_ctx ctx, void *dst_host, const void *src_dev, size_t bytes);

/* ------------------------------------------------- level 1: kernel-level API */
/* ElementRestriction::Mult, fem/restriction.cpp:109-129.  y[i] = ±x[gather_map[i]] */
int b200pa_restrict_mult(b200pa_ctx ctx, int ne, int nd, const int *gather_map_dev,
                         const double *x_dev, double *y_dev);
/* ElementRestriction::MultTranspose / AbsMultTranspose, fem/restriction.cpp:152-186, 196-221.
 * Atomic-free: one thread per L-dof, CSR (offsets, indices), ascending element order. */
int b200pa_restrict_mult_transpose(b200pa_ctx ctx, int ndofs, const int *offsets_dev,
                                   const int *indices_dev, const double *xE_dev, double *yL_dev,
                                   int abs);
/* internal::PADiffusionSetup3D, fem/integ/bilininteg_diffusion_kernels.cpp:243-367 (scalar
 * coefficient branch :349-362).  nc = 1 (constant) or Q^3*NE.  D is [Q,Q,Q,6,NE]. */
int b200pa_diffusion_setup(b200pa_ctx ctx, int q1d, int ne, const double *W_dev, const double *J_dev,
                           const double *C_dev, long long nc, double *D_dev);
/* MassIntegrator::AssemblePA inner kernel, fem/integ/bilininteg_mass_pa.cpp:62-78 */
int b200pa_mass_setup(b200pa_ctx ctx, int nq, int ne, const double *W_dev, const double *detJ_dev,
                      const double *C_dev, long long nc, double *v_dev);
/* DiffusionIntegrator::ApplyKernelType (fem/bilininteg.hpp:2181-2185) →
 * SmemPADiffusionApply3D, fem/integ/bilininteg_diffusion_kernels.hpp:989-1214:  yE += G^T D G xE */
int b200pa_diffusion_apply(b200pa_ctx ctx, int ne, int d1d, int q1d, const double *B_any,
                           const double *G_any, const double *D_dev, const double *xE_dev, double *yE_dev);
/* MassIntegrator::ApplyKernelType (fem/bilininteg.hpp:2387-2389) → SmemPAMassApply3D,
 * fem/integ/bilininteg_mass_kernels.hpp:1119-1144:  yE += B^T v B xE */
int b200pa_mass_apply(b200pa_ctx ctx, int ne, int d1d, int q1d, const double *B_any,
                      const double *v_dev, const double *xE_dev, double *yE_dev);
/* DiagonalKernelType: SmemPADiffusionDiagonal3D (…diffusion_kernels.hpp:369-484) and
 * SmemPAMassAssembleDiagonal3D (…mass_kernels.hpp:324-408):  dE += diag(element matrix) */
int b200pa_diffusion_diag(b200pa_ctx ctx, int ne, int d1d, int q1d, const double *B_any,
                          const double *G_any, const double *D_dev, double *dE_dev);
int b200pa_mass_diag(b200pa_ctx ctx, int ne, int d1d, int q1d, const double *B_any,
                     const double *v_dev, double *dE_dev);
/* quadrature_interpolator::Values3D, fem/qinterp/eval.hpp:131-193 (vdim 1): yq[Q,Q,Q,NE] */
int b200pa_qvalues(b200pa_ctx ctx, int ne, int d1d, int q1d, const double *B_any,
                   const double *xE_dev, double *yq_dev);
/* quadrature_interpolator::Derivatives3D<byVDIM,GRAD_PHYS>, fem/qinterp/grad.hpp:233-374:
 * gq[3,Q,Q,Q,NE] = J^{-T} grad_ref */
int b200pa_qphysgrad(b200pa_ctx ctx, int ne, int d1d, int q1d, const double *B_any,
                     const double *G_any, const double *J_dev, const double *xE_dev, double *gq_dev);
/* DLFEvalAssemble3D, fem/integ/lininteg_domain_kernels.hpp:164-298: bE += B^T (W f detJ);
 * nf = 1 (constant) or Q^3*NE */
int b200pa_domain_lf(b200pa_ctx ctx, int ne, int d1d, int q1d, const double *B_any,
                     const double *detJ_dev, const double *W_dev, const double *f_dev, long long nf,
                     double *bE_dev);
/* Vector::operator*, linalg/vector.cpp:1079-1152 (+ general/reducers.hpp:532-592).
 * Deterministic block-tree reduction; synchronous (returns the value on the host). */
int b200pa_dot(b200pa_ctx ctx, long long n, const double *a_dev, const double *b_dev, double *result_host);
/* add(v1, alpha, v2, v), linalg/vector.cpp:436-473:  v = v1 + alpha v2 */
int b200pa_add(b200pa_ctx ctx, long long n, const double *v1_dev, double alpha, const double *v2_dev,
               double *v_dev);
/* OperatorJacobiSmoother::Setup / ::Mult, linalg/solvers.cpp:401-425, 427-453 */
int b200pa_jacobi_setup(b200pa_ctx ctx, int n, const double *diag_dev, int n_ess, const int *ess_dev,
                        double damping, double *dinv_dev);
int b200pa_jacobi_mult(b200pa_ctx ctx, int n, const double *dinv_dev, const double *r_dev, double *z_dev);
/* q-point coefficient evaluation (the "user forall over Q-points" of SURVEY §3.2/§3.3):
 *   kind 0: out = a*(1 + b*(T - T0))                   k(T), sigma(T)
 *   kind 1: out = a                                      constant fill (rho c/dt + perfusion)
 *   kind 2: out = s*|g|^2 + a   (g = grad phi [3,..])   Joule source + perfusion source
 *           (semantics of miniapps/electromagnetics/joule_solver.cpp:898-906) */
int b200pa_coeff_eval(b200pa_ctx ctx, int kind, long long n, double a, double b, double T0,
                      const double *T_dev, const double *s_dev, const double *g_dev, double *out_dev);

/* -------------------------------------------- level 2: space / form / solver */
/* ElementRestriction ctor (fem/restriction.cpp:26-107) + DofToQuad (fem/fe/fe_base.cpp:2619-2662).
 * Copies gather_map (sign-encoded entries are rejected: H1 has none), builds the CSR
 * (offsets, indices) in ascending element order and its inverse (slot of every E-entry). */
int b200pa_space_create(b200pa_ctx ctx, int d1d, int q1d, int ne, int ndofs, const int *gather_map_any,
                        const double *B_any, const double *G_any, b200pa_space *out);
int b200pa_space_destroy(b200pa_space sp);
/* GeometricFactors (mesh/mesh.hpp:3086-3130): J[Q^3,3,3,NE], detJ[Q^3,NE], rule weights W[Q^3].
 * J/detJ are referenced if they are device pointers (caller keeps them alive), copied otherwise. */
int b200pa_space_set_geometry(b200pa_space sp, const double *W_any, const double *J_any,
                              const double *detJ_any);
/* GeometricFactors::Compute (mesh/mesh.cpp:15220-15273) for trilinear hexes, on the device:
 * vertices f64[3*nv], elem_vertices int32[8*NE] in the reference's hex vertex order.  The space
 * keeps the vertices and detJ; J[Q^3,3,3,NE] is NOT stored (set-up and q-point kernels rebuild it
 * per q-point) unless it is asked for through b200pa_space_J(). */
int b200pa_space_geometry_from_vertices(b200pa_space sp, const double *W_any, int nv,
                                        const double *vertices_any, const int *elem_vertices_any);
/* 1 when every element of the space is affine (vertices: a parallelepiped to 1e-13 of its edge lengths; stored
 * Jacobians: J constant over the element's q-points to 1e-13; decided on the device when the geometry is set),
 * 0 otherwise, -1 for a NULL space.  Affine meshes admit the factorised
 * diffusion q-data below. */
int b200pa_space_is_affine(b200pa_space sp);
/* read-only accessors to the device arrays (for tests and for the host mirror) */
const int *b200pa_space_offsets(b200pa_space sp);
const int *b200pa_space_indices(b200pa_space sp);
const int *b200pa_space_gather_map(b200pa_space sp);
const double *b200pa_space_J(b200pa_space sp);
const double *b200pa_space_detJ(b200pa_space sp);
const double *b200pa_space_W(b200pa_space sp);

/* q-point operators straight from an L-vector (the L->E gather is fused in, nothing E-sized or
 * q-sized is staged in HBM beyond the output):
 *   qvalues    ≙ QuadratureFunction::ProjectGridFunction (fem/qfunction.cpp:73-105)
 *   qphysgrad  ≙ ElementRestriction::Mult + QuadratureInterpolator::PhysDerivatives
 *                (fem/quadinterpolator.cpp:682-687), layout [3,Q,Q,Q,NE]
 *   coeff_linear: out_q = a (1 + b (T_q - T0))              k(T), sigma(T) of SURVEY §3.2/§3.3
 *   joule:        out_q = sigma_q |grad phi|_q^2 + add      (miniapps/electromagnetics/joule_solver.cpp:898-906)
 *   domain_lf  ≙ LinearForm(DomainLFIntegrator(f_q)) with UseFastAssembly (fem/linearform.cpp:162-184,
 *                fem/integ/lininteg_domain.cpp:22-59): b_L = R^T B^T (W f detJ); nf = 1 or Q^3*NE */
int b200pa_space_qvalues(b200pa_space sp, const double *xL_dev, double *yq_dev);
int b200pa_space_qphysgrad(b200pa_space sp, const double *xL_dev, double *gq_dev);
int b200pa_space_coeff_linear(b200pa_space sp, double a, double b, double T0, const double *TL_dev, double *out_q_dev);
int b200pa_space_joule(b200pa_space sp, const double *phiL_dev, const double *sigma_q_dev, double add, double *out_q_dev);
int b200pa_space_domain_lf(b200pa_space sp, const double *f_dev, long long nf, double *bL_dev);

/* PABilinearFormExtension (fem/bilinearform_ext.cpp:246-847): diffusion and/or mass on `sp`. */
int b200pa_form_create(b200pa_space sp, b200pa_form *out);
int b200pa_form_destroy(b200pa_form f);
/* DiffusionIntegrator::AssemblePA (fem/integ/bilininteg_diffusion_pa.cpp:89-142) with a constant
 * (nc = 1) or q-data (nc = Q^3*NE; ≙ QuadratureFunctionCoefficient, fem/coefficient.cpp:2059-2062)
 * scalar coefficient.  C may be NULL to drop the integrator. */
int b200pa_form_assemble_diffusion(b200pa_form f, const double *C_any, long long nc);
/* MassIntegrator::AssemblePA (fem/integ/bilininteg_mass_pa.cpp:24-79) */
int b200pa_form_assemble_mass(b200pa_form f, const double *C_any, long long nc);
/* use caller-provided pa_data (device) instead of assembling; NULL drops the integrator */
int b200pa_form_set_pa_data(b200pa_form f, const double *pa_diff_dev, const double *pa_mass_dev);
const double *b200pa_form_pa_diff(b200pa_form f);
const double *b200pa_form_pa_mass(b200pa_form f);
/* Factorised diffusion q-data.  On an affine element J is constant, so the reference's
 * D(q) = (w_q / det J) c_q adj(J) adj(J)^T (PADiffusionSetup3D, bilininteg_diffusion_kernels.cpp:243-367) is the
 * product of a per-ELEMENT tensor (6 doubles, kept by the space) and the scalar w_q c_q: 8 instead of 48 bytes per
 * q-point for AssemblePA to write and AddMultPA / AssembleDiagonalPA to read.  on = 1: the following
 * b200pa_form_assemble_diffusion calls store that form (b200pa_form_pa_diff then returns w_q c_q [Q^3,NE]); they
 * fail, loudly, when b200pa_space_is_affine() is not 1.  Results agree with the stored form to rounding. */
int b200pa_form_set_factorised(b200pa_form f, int on);
int b200pa_form_is_factorised(b200pa_form f);
/* Element-attribute markers: BilinearForm::AddDomainIntegrator(bfi, elem_marker) (fem/bilinearform.hpp; multi-material
 * domains: tissue / blood / electrode).  b200pa_space_set_attributes: Mesh::GetAttribute(e), one int >= 1 per element
 * (host or device).  b200pa_form_set_markers(which = 0 diffusion | 1 mass): marker[a-1] != 0 <=> the integrator acts
 * on the elements of attribute a (n_attr >= the largest attribute); marker = NULL removes it (re-assemble afterwards).
 *   apply     PABilinearFormExtension::AddMultWithMarkers (fem/bilinearform_ext.cpp:807-847): the element
 *             contributions of excluded elements are left out of the sum - here the integrator's q-data is zeroed on
 *             them at assembly, so the hot kernel is unchanged and those contributions are exactly 0;
 *   diagonal  PABilinearFormExtension::AssembleDiagonal (:370-454) zeroes, after EVERY integrator, the accumulated
 *             element diagonal of the elements that integrator's marker excludes - including what EARLIER integrators
 *             added there.  The form's integrator order is diffusion, then mass (as b200::PAOperator adds them), and
 *             this order dependence is reproduced: an element the MASS marker excludes has a zero element diagonal.
 * Markers need q-data assembled by the library (not b200pa_form_set_pa_data). */
int b200pa_space_set_attributes(b200pa_space sp, const int *attr_any);
int b200pa_form_set_markers(b200pa_form f, int which, int n_attr, const int *marker_host);
/* ConstrainedOperator ctor (linalg/operator.cpp:511-526): essential true-dof list, DIAG_ONE */
int b200pa_form_set_essential(b200pa_form f, int n_ess, const int *ess_any);
/* PABilinearFormExtension::Mult (fem/bilinearform_ext.cpp:487-564): y = A x, L→L, unconstrained.
 * One fused gather + sum-factorised contraction launch and one segmented E→L reduction. */
int b200pa_form_mult(b200pa_form f, const double *x_dev, double *y_dev);
/* profiling hook for bench.py's per-kernel roofline: phases bit 0 = the fused gather + element
 * kernel only (writes the E-sized scratch), bit 1 = the segmented E->L reduction only; 3 = mult */
int b200pa_form_mult_phases(b200pa_form f, const double *x_dev, double *y_dev, int phases);
/* ConstrainedOperator::Mult (linalg/operator.cpp:586-646, 710-714) */
int b200pa_form_constrained_mult(b200pa_form f, const double *x_dev, double *y_dev);
/* the same two with HOST vectors: copies x up, applies, copies y back, synchronises.  One GPU, >= 2^20 dofs: pipelined -
 * x tiles up, element chunks, E->L reduction of the tiles a chunk completes and y tiles down overlap on three streams
 * (page-locked vectors, e.g. b200pa_host_alloc, let the copies run asynchronously).  Environment: B200PA_NO_PIPELINE=1
 * selects the serial route, B200PA_PIPE_CHUNKS / B200PA_PIPE_TILE the plan (default 8 chunks, 32768 dofs per tile),
 * B200PA_PIPE_TRACE=1 prints the per-chunk timeline of every call on stderr. */
int b200pa_form_mult_host(b200pa_form f, int constrained, const double *x_host, double *y_host);
/* PABilinearFormExtension::AssembleDiagonal (fem/bilinearform_ext.cpp:370-454) */
int b200pa_form_assemble_diagonal(b200pa_form f, double *diag_dev);
/* DiffusionIntegrator::AssemblePA (fem/integ/bilininteg_diffusion_pa.cpp:89-142) + the form's AssembleDiagonal
 * (fem/bilinearform_ext.cpp:370-454) in ONE pass over the q-points - what every implicit step with k(T) needs.
 * On a mesh of affine elements (b200pa_space_is_affine) the kernel forms
 * W C per q-point, writes the integrator's q-data from it and takes the diagonal from the same values (the q-data
 * is never read back); the mass integrator is used as currently assembled.  Elsewhere it is
 * b200pa_form_assemble_diffusion followed by b200pa_form_assemble_diagonal.  Same results either way. */
int b200pa_form_assemble_diffusion_with_diagonal(b200pa_form f, const double *C_any, long long nc, double *diag_dev);
/* ConstrainedOperator::EliminateRHS (linalg/operator.cpp:559-584): b -= A w; b[ess] = x[ess] */
int b200pa_form_eliminate_rhs(b200pa_form f, const double *x_dev, double *b_dev);

/* CGSolver::Mult with OperatorJacobiSmoother (linalg/solvers.cpp:869-1050, 331-453), operator =
 * ConstrainedOperator(f), iterative_mode = true.  Stopping test, iteration numbering,
 * `converged`, `final_norm = sqrt((Br,r))` as the reference.  norms_host (may be NULL) receives
 * (B r, r) for iterations 0..final_iter (size max_iter+1).  Synchronous.
 * b, x: device pointers (b200pa_pcg_solve) or host pointers (b200pa_pcg_solve_host: copies
 * b and x0 up, x back). */
typedef struct
{
   int final_iter;
   int converged;
   double final_norm;
   double initial_norm;
} b200pa_pcg_result;
int b200pa_pcg_solve(b200pa_form f, const double *dinv_dev, const double *b_dev, double *x_dev,
                     double rel_tol, double abs_tol, int max_iter, b200pa_pcg_result *res,
                     double *norms_host);
int b200pa_pcg_solve_host(b200pa_form f, const double *dinv_dev, const double *b_host, double *x_host,
                          double rel_tol, double abs_tol, int max_iter, b200pa_pcg_result *res,
                          double *norms_host);

/* OperatorChebyshevSmoother (linalg/solvers.cpp:455-657) on the constrained PA operator, orders 1..5.
 *   b200pa_chebyshev_coeffs : ::Setup's polynomial coefficients (:571-621); host arithmetic, needs no device
 *   b200pa_power_method     : PowerMethod::EstimateLargestEigenvalue (linalg/operator.cpp:871-928) of Dinv*A, the
 *                             estimate the smoother's second constructor computes (:497-511: 10 steps, 1e-8, start
 *                             vector Vector::Randomize(12345) - b200pa_randomize); v0_dev is overwritten; with a
 *                             communicator on the form v0 must be a consistent L-vector
 *   b200pa_chebyshev_mult   : ::Mult (:623-657): y = p(Dinv A) Dinv x; dinv = b200pa_jacobi_setup(damping 1)
 *   b200pa_pcg_solve_chebyshev : CGSolver::Mult with that smoother as the preconditioner (same result struct,
 *                             stopping rule and residual history as b200pa_pcg_solve) */
int b200pa_chebyshev_coeffs(int order, double max_eig, double *coeffs_host);
int b200pa_power_method(b200pa_form f, const double *dinv_dev, double *v0_dev, int num_steps, double tolerance,
                        double *max_eig);
int b200pa_chebyshev_mult(b200pa_form f, const double *dinv_dev, int order, double max_eig, const double *x_dev,
                          double *y_dev);
int b200pa_pcg_solve_chebyshev(b200pa_form f, const double *dinv_dev, int order, double max_eig, const double *b_dev,
                               double *x_dev, double rel_tol, double abs_tol, int max_iter, b200pa_pcg_result *res,
                               double *norms_host);

/* ------------------------------------------------------ p-multigrid */
/* Order-refinement transfer between two forms on the SAME mesh (coarse order <= fine order):
 * TensorProductPRefinementTransferOperator::{Mult, MultTranspose} (fem/transfer.cpp:2223-2296, 2542-2592) wrapped
 * in the RectangularConstrainedOperator a GeometricMultigrid gives it (fem/multigrid.cpp:281-296): essential dofs of
 * either level (b200pa_form_set_essential) count as zero on input and are zeroed on output.
 * B_any: [DF x DC] column-major, the coarse 1-D basis at the fine nodes in lexicographic order - the DofToQuad::B
 * the reference operator builds (b200pa_basis_transfer for the synthetic builder's GLL-nodal H1 bases). */
typedef struct b200pa_transfer_s *b200pa_transfer;
int b200pa_transfer_create(b200pa_form coarse, b200pa_form fine, const double *B_any, b200pa_transfer *out);
int b200pa_transfer_destroy(b200pa_transfer t);
int b200pa_transfer_mult(b200pa_transfer t, const double *xc_dev, double *yf_dev);            /* prolongation */
int b200pa_transfer_mult_transpose(b200pa_transfer t, const double *xf_dev, double *yc_dev);  /* restriction  */
/* Multigrid (fem/multigrid.hpp, fem/multigrid.cpp:107-220; the hierarchy of examples/ex26.cpp): forms[0] is the coarsest
 * level, transfers[l] connects forms[l] and forms[l+1].  Levels >= 1: OperatorChebyshevSmoother (b200pa_mg_setup: order and
 * largest-eigenvalue estimate per level, <= 0 = the reference's power method); level 0: CGSolver to (rel_tol, abs_tol,
 * max_iter), unpreconditioned as in ex26 or with OperatorJacobiSmoother (jacobi = 1).  b200pa_mg_setup must be called again
 * after the forms are re-assembled.  b200pa_mg_mult = MultigridBase::Mult: one V- (or W-) cycle from a zero guess.
 * b200pa_pcg_solve_mg = CGSolver::Mult on the finest constrained operator preconditioned by the cycle.
 * Partitioned meshes: give EVERY level's form a communicator of its own (b200pa_form_set_comm; the shared-dof tables
 * differ per order) before b200pa_mg_create.  Vectors are consistent L-vectors on every level: the prolongation
 * broadcasts the owners' values of shared fine dofs, the restriction counts every fine dof once (through its owner) and
 * sums the coarse interface dofs across ranks, smoothers / coarse CG / outer CG use the all-reduced dots; the power
 * method starts from the owners' random values, so every rank gets the same eigenvalue estimate. */
typedef struct b200pa_mg_s *b200pa_mg;
int b200pa_mg_create(int nlevels, const b200pa_form *forms, const b200pa_transfer *transfers, b200pa_mg *out);
int b200pa_mg_destroy(b200pa_mg m);
int b200pa_mg_set_cycle(b200pa_mg m, int wcycle, int pre_smoothing_steps, int post_smoothing_steps);
int b200pa_mg_set_coarse_solver(b200pa_mg m, double rel_tol, double abs_tol, int max_iter, int jacobi);
int b200pa_mg_setup(b200pa_mg m, const int *order, const double *max_eig);
double b200pa_mg_max_eig(b200pa_mg m, int level);
int b200pa_mg_coarse_iterations(b200pa_mg m);
int b200pa_mg_mult(b200pa_mg m, const double *x_dev, double *y_dev);
int b200pa_pcg_solve_mg(b200pa_mg m, const double *b_dev, double *x_dev, double rel_tol, double abs_tol, int max_iter,
                        b200pa_pcg_result *res, double *norms_host);

/* ------------------------------------------------------ multi-GPU (one rank per GPU) */
/* Shared-dof exchange ≙ DeviceConformingProlongationOperator::{Mult,MultTranspose}
 * (fem/pfespace.cpp:5259-5532; GroupCommunicator, general/communication.cpp:723-1120) and
 * InnerProduct(comm,…) (linalg/vector.hpp:773-779), over NCCL instead of MPI.
 * Vectors handed to a form with a comm attached are *consistent L-vectors*: every copy of a
 * shared dof holds the same value; the owner (lowest sharing rank, owner_mask = 1) is the copy
 * that dots count.  P^T followed by P collapses into one symmetric neighbour exchange whose
 * per-dof summation order is ascending rank on every rank (bit-identical copies).
 * nccl_id: the 128-byte ncclUniqueId created on rank 0 (b200pa_comm_unique_id) and broadcast
 * by the host application (torch.distributed / MPI).  nccl_id == NULL creates a communicator
 * WITHOUT NCCL: the peer-memory path below is then its only transport (every exchange before
 * b200pa_comm_px_connect fails); this is what ranks sharing one GPU use - NCCL refuses them. */
int b200pa_comm_unique_id(unsigned char id_out[128]);
int b200pa_comm_create(b200pa_ctx ctx, const unsigned char nccl_id[128], int rank, int nranks, b200pa_comm *out);
int b200pa_comm_destroy(b200pa_comm c);
/* Neighbour tables (≙ GroupCommunicator::GetNeighborLDofTable, general/communication.hpp:294-298):
 * nbr_rank strictly ascending; shared_ldofs[shared_offsets[k] .. shared_offsets[k+1]) = local
 * L-dofs shared with neighbour k, in an order both sides agree on (ascending global dof id). */
int b200pa_comm_set_tables(b200pa_comm c, int ndofs, int n_nbr, const int *nbr_rank,
                           const int *shared_offsets, const int *shared_ldofs);
/* the host-side table construction behind set_tables (no GPU needed; used by the CPU tests).
 * sh_ldof[n_shared], sh_off[n_shared+1], sh_src[n_send+n_shared] (-1 = own value, else index
 * into the concatenated receive buffer; ascending rank order), owner_mask[ndofs].  Pass NULL
 * tables to query n_shared only. */
int b200pa_comm_build_tables(int rank, int ndofs, int n_nbr, const int *nbr_rank, const int *shared_offsets,
                             const int *shared_ldofs, int *n_shared_out, int *sh_ldof, int *sh_off, int *sh_src,
                             unsigned char *owner_mask);
const unsigned char *b200pa_comm_owner_mask(b200pa_comm c);   /* device pointer */
/* Peer-memory path over NVLink / NVSwitch (optional, after set_tables): the exchange and the scalar
 * all-reduce are then done by kernels that store straight into the peers' mailboxes (CUDA IPC mapped)
 * and synchronise with release/acquire flags - no NCCL call inside the PCG loop.
 *   1. every rank: b200pa_comm_px_prepare -> 64-byte cudaIpcMemHandle of its mailbox
 *   2. host application all-gathers (handle, nbr_rank[], shared_offsets[]) of every rank
 *   3. every rank: b200pa_comm_px_connect(handles[nranks*64], remote_off[n_nbr], remote_nsend[n_nbr]) where,
 *      for neighbour k = rank q, remote_off[k] = q's shared_offsets[index of this rank in q's nbr_rank] and
 *      remote_nsend[k] = q's shared_offsets[q's n_nbr]
 * Waits are bounded: a time-out raises a device error word, the kernels that follow skip their waits, and the
 * host-synchronous entry points (b200pa_pcg_solve*, b200pa_form_mult_host) fail with a message instead of
 * returning results computed from stale mailbox data; after the asynchronous ones (b200pa_form_mult, ...)
 * b200pa_comm_px_error() returns nonzero (and clears the word) if a wait timed out since the last query.
 * b200pa_comm_set_tables on a connected communicator tears the peer path down (mailbox sizes and peer
 * offsets came from the old tables): prepare / connect again, collectively. */
int b200pa_comm_px_prepare(b200pa_comm c, unsigned char handle_out[64]);
int b200pa_comm_px_connect(b200pa_comm c, const unsigned char *handles, const long long *remote_off,
                           const long long *remote_nsend);
int b200pa_comm_px_error(b200pa_comm c);
int b200pa_comm_px_enabled(b200pa_comm c);
int b200pa_comm_px_disable(b200pa_comm c);   /* back to NCCL; must be called on every rank */
/* fails unless c's tables are set (nranks > 1) and were built for the form's number of L-dofs */
int b200pa_form_set_comm(b200pa_form f, b200pa_comm c);
/* (P P^T) y: every copy of a shared dof <- sum of all copies.  (P R) x: <- the owner's value. */
int b200pa_comm_exchange_sum(b200pa_comm c, double *yL_dev);
int b200pa_comm_bcast(b200pa_comm c, double *xL_dev);
int b200pa_comm_allreduce_sum(b200pa_comm c, double *vals_dev, int n);

/* -------------------------------------------------- host-side problem builder (no GPU) */
/* Mesh::MakeCartesian3D (mesh/mesh.cpp:3683-3790, 4627-4635; space-filling-curve element order,
 * mesh/ncmesh.cpp:5435-5620) + H1 dof numbering (fem/fespace.cpp:3426-3533) + lexicographic
 * E-ordering (fem/restriction.cpp:44-62) for an nx*ny*nz hex mesh, order p — reproduces the
 * reference's gather_map exactly (tests/test_hexmesh.py).  All outputs are caller-allocated host
 * arrays; pass NULL to skip one.
 *   gather_map   int32[ne*(p+1)^3]        elem_vertices int32[8*ne]
 *   vertices     f64[3*nv]  ((x,y,z) per vertex, sx,sy,sz box; `skew` != 0 applies
 *                y += 0.2x, z += 0.3x as tests/unit/fem/test_pa_coeff.cpp:33-39)
 *   elem_ijk     int32[3*ne] lattice position of every element (SFC order)
 *   bdr_attr_of_dof  uint8[ndofs]: bit a-1 set if the dof lies on boundary attribute a (1..6)
 */
/* results hand-off in the reference's own text formats: Mesh::Print "MFEM mesh v1.0" (mesh/mesh.cpp:12239-12360) for the
 * mesh b200pa_hex_build numbers, and GridFunction::Save (fem/gridfunc.cpp:4142-4165) for a scalar H1 field of order p in
 * its L-dof numbering (host values).  GLVis / the reference load both; the loaded space has the builder's numbering. */
int b200pa_hex_write_mesh(const char *path, int nx, int ny, int nz, double sx, double sy, double sz, int skew);
int b200pa_write_gridfunction(const char *path, int p, long long n, const double *values_host);
/* ParaViewDataCollection::Save (fem/datacollection.hpp:584, fem/datacollection.cpp:887-1083; Mesh::PrintVTU,
 * mesh/mesh.cpp:12683-12890) for a hexahedral mesh without nodal GridFunction and scalar H1 fields of order p given in
 * the L-dof numbering of gather_map (host arrays): writes <prefix_path><collection>/Cycle%06d/proc%06d.vtu for this rank
 * and, on rank 0, Cycle%06d/data.pvtu and <collection>.pvd.  Every element carries (levels_of_detail+1)^3 uniformly
 * spaced points; high_order != 0: one VTK Lagrange hexahedron of order levels_of_detail per element (the reference's
 * SetHighOrderOutput(true)), else levels_of_detail^3 linear hexahedra.  format: 0 ascii, 1 binary (base64 Float64),
 * 2 binary32 (base64 Float32) - VTKFormat; zlib compression is not offered (compression level 0).
 * attributes: int32[ne] or NULL (all 1).  append = 0 starts a new .pvd (the first Save of a collection), 1 adds this
 * cycle to the existing one.  With several ranks every rank calls this with its own piece; only rank 0 needs nranks. */
int b200pa_paraview_save(const char *prefix_path, const char *collection, int cycle, double time, int rank, int nranks,
                         int p, long long ne, long long ndofs, const int *gather_map, const double *vertices,
                         const int *elem_vertices, const int *attributes, int nfields, const char *const *names,
                         const double *const *values_host, int levels_of_detail, int high_order, int format, int append);
int b200pa_hex_sizes(int nx, int ny, int nz, int p, long long *ne, long long *nv, long long *ndofs);
int b200pa_hex_build(int nx, int ny, int nz, int p, double sx, double sy, double sz, int skew,
                     int *gather_map, int *elem_vertices, double *vertices, int *elem_ijk,
                     unsigned char *bdr_attr_of_dof);
/* the same for the [ox,ox+nx) x [oy,oy+ny) x [oz,oz+nz) sub-box of a GNX x GNY x GNZ global mesh
 * (≙ Mesh::CartesianPartitioning + ParMesh, mesh/mesh.cpp:8966-9003, mesh/pmesh.cpp:106): local
 * numbering as if the sub-box were a mesh of its own, vertex coordinates / boundary attributes /
 * lattice coordinates (int32[3*ndofs], units of h/p) those of the global mesh */
int b200pa_hex_build_part(int GNX, int GNY, int GNZ, int ox, int oy, int oz, int nx, int ny, int nz, int p,
                          double sx, double sy, double sz, int skew, int *gather_map, int *elem_vertices,
                          double *vertices, int *elem_ijk, unsigned char *bdr_attr_of_dof, int *lattice);
/* lattice coordinates (ix,iy,iz in [0,p*n]) of every L-dof: int32[3*ndofs] */
int b200pa_hex_dof_lattice(int nx, int ny, int nz, int p, int *lattice);
/* Vector::Randomize(seed) (linalg/vector.cpp:955-967, rand_real linalg/vector.hpp:61-80):
 * srand(seed); out[i] = rand() / (RAND_MAX + 1.0) — the reference tests' and benchmarks' input */
int b200pa_randomize(int seed, long long n, double *out_host);
/* DofToQuad in TENSOR mode for H1 (GLL nodes) at the Gauss-Legendre rule with q1d points
 * (fem/fe/fe_base.cpp:2619-2662, fem/intrules.cpp): B,G f64[q1d*d1d] column-major, w1d f64[q1d],
 * W f64[q1d^3] */
int b200pa_basis(int p, int q1d, double *B, double *G, double *w1d, double *W, double *gll_nodes);
/* the 1-D matrix of the order-refinement transfer pc -> pf for these bases (b200pa_transfer_create): [pf+1, pc+1] column-major */
int b200pa_basis_transfer(int pc, int pf, double *B);

#ifdef __cplusplus
}
#endif
#endif /* B200PA_H */
