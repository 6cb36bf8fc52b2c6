// Streaming kernels around the element kernel: restriction, set-up, diagonal, BLAS-1, Jacobi,
// deterministic reductions.  All FP64, all HBM-bound; thread = one output entry, coalesced index
// streams, grid-stride loops sized from the SM count.
#pragma once
#include <cuda_runtime.h>

#include "geometry.cuh"
#include "pcg_state.cuh"
#include "reduce.cuh"
#include "tma.cuh"

namespace b200pa
{

__global__ void k_dot(long long n, const double *__restrict__ a, const double *__restrict__ b,
                      double *partials, unsigned int *ticket, double *out)
{
   double s = 0.0;
   for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
   {
      s = fma(a[i], b[i], s);
   }
   grid_sum(s, partials, ticket, out);
}

__global__ void k_pcg_scalar_init(PcgState *st, double *norms) { pcg_scalar_init(st, norms); }
__global__ void k_pcg_scalar_den(PcgState *st) { pcg_scalar_den(st); }
__global__ void k_pcg_scalar_beta(PcgState *st, double *norms) { pcg_scalar_beta(st, norms); }

// ------------------------------------------------------------------ restriction
// fem/restriction.cpp:109-129
__global__ void k_restrict_mult(long long n, const int *__restrict__ gmap, const double *__restrict__ x,
                                double *__restrict__ y)
{
   for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
   {
      const int g = gmap[i];
      y[i] = g >= 0 ? x[g] : -x[-1 - g];
   }
}

// fem/restriction.cpp:152-186, 196-221 — CSR, one thread per L-dof, ascending element order
__global__ void k_restrict_mult_transpose(int ndofs, const int *__restrict__ offsets, const int *__restrict__ indices,
                                          const double *__restrict__ xE, double *__restrict__ y, int abs)
{
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ndofs; i += gridDim.x * blockDim.x)
   {
      double v = 0.0;
      const int j1 = offsets[i + 1];
      for (int j = offsets[i]; j < j1; ++j)
      {
         const int s = indices[j];
         const int k = s >= 0 ? s : -1 - s;
         const double t = xE[k];
         v += (abs || s >= 0) ? t : -t;
      }
      y[i] = v;
   }
}

// The same sum over the SLOT layout written by the element kernel (y_S[j], j = CSR position):
// a contiguous segmented reduction — no index stream, no scattered reads.  Optional fused
// ConstrainedOperator fix-up (linalg/operator.cpp:615-640: y[ess] = x[ess]) and d.z partial.
template <bool CONSTR, bool DOT, bool ABS>
__global__ void k_segment_sum(int ndofs, const int *__restrict__ offsets, const double *__restrict__ yS,
                              double *__restrict__ y, const unsigned char *__restrict__ ess_mask,
                              const double *__restrict__ x, const unsigned char *__restrict__ own_mask,
                              double *partials, unsigned int *ticket, double *dot_out, const int *done_flag,
                              PcgState *st_epilogue = nullptr)
{
   if (done_flag && *done_flag) { return; }
   double acc = 0.0;
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ndofs; i += gridDim.x * blockDim.x)
   {
      double v = 0.0;
      const int j1 = offsets[i + 1];
      for (int j = offsets[i]; j < j1; ++j) { v += ABS ? fabs(yS[j]) : yS[j]; }
      if (CONSTR && ess_mask[i]) { v = x[i]; }
      y[i] = v;
      if (DOT) { if (!own_mask || own_mask[i]) { acc = fma(x[i], v, acc); } }
   }
   if (DOT)
   {
      // den = (d, A d) is complete: alpha / termination logic right here (linalg/solvers.cpp:1010-1024)
      if (grid_sum(acc, partials, ticket, dot_out) && st_epilogue) { pcg_scalar_den(st_epilogue); }
   }
}

// the same over a LIST of dof tiles (tile t = dofs [t TS, (t+1) TS)): one launch finishes all the tiles an element chunk of
// the pipelined host-buffer apply completes (b200pa_form_mult_host), however scattered they are
template <bool CONSTR>
__global__ void k_segment_sum_tiles(const int *__restrict__ tiles, int TS, int ndofs, const int *__restrict__ offsets,
                                    const double *__restrict__ yS, double *__restrict__ y, const unsigned char *__restrict__ ess_mask,
                                    const double *__restrict__ x)
{
   const int bpt = TS / blockDim.x;                 // blocks per tile (TS is a multiple of the block size)
   const int t = tiles[blockIdx.x / bpt];
   const int i = t * TS + (blockIdx.x % bpt) * blockDim.x + threadIdx.x;
   if (i >= ndofs) { return; }
   double v = 0.0;
   const int j1 = offsets[i + 1];
   for (int j = offsets[i]; j < j1; ++j) { v += yS[j]; }
   if (CONSTR && ess_mask[i]) { v = x[i]; }
   y[i] = v;
}

// multi-GPU with the peer-memory exchange: shared dofs only get their local partial sum here (the exchange
// kernel finishes them: remote contributions, constraint, their part of the dot); everything else is final
template <bool CONSTR, bool DOT>
__global__ void k_segment_sum_mg(int ndofs, const int *__restrict__ offsets, const double *__restrict__ yS, double *__restrict__ y,
                                 const unsigned char *__restrict__ ess_mask, const double *__restrict__ x,
                                 const unsigned char *__restrict__ shared_mask, double *partials, unsigned int *ticket,
                                 double *dot_out, const int *done_flag)
{
   if (done_flag && *done_flag) { return; }
   double acc = 0.0;
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ndofs; i += gridDim.x * blockDim.x)
   {
      double v = 0.0;
      const int j1 = offsets[i + 1];
      for (int j = offsets[i]; j < j1; ++j) { v += yS[j]; }
      if (!shared_mask[i])
      {
         if (CONSTR && ess_mask[i]) { v = x[i]; }
         if (DOT) { acc = fma(x[i], v, acc); }
      }
      y[i] = v;
   }
   if (DOT) { grid_sum(acc, partials, ticket, dot_out); }
}

// multi-GPU second pass after the shared-dof exchange: ConstrainedOperator fix-up and the
// owned-dof partial of d.z
__global__ void k_fixup_dot(int ndofs, double *__restrict__ y, const unsigned char *__restrict__ ess_mask,
                            const double *__restrict__ x, const unsigned char *__restrict__ own_mask, int want_dot,
                            double *partials, unsigned int *ticket, double *dot_out, const int *done_flag)
{
   if (done_flag && *done_flag) { return; }
   double acc = 0.0;
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ndofs; i += gridDim.x * blockDim.x)
   {
      double v = y[i];
      if (ess_mask && ess_mask[i]) { v = x[i]; y[i] = v; }
      if (want_dot && (!own_mask || own_mask[i])) { acc = fma(x[i], v, acc); }
   }
   if (want_dot) { grid_sum(acc, partials, ticket, dot_out); }
}

// ----------------------------------------------------------------------- set-up
// fem/integ/bilininteg_diffusion_kernels.cpp:243-367, scalar branch :349-362
__global__ void k_diffusion_setup(long long NQ, long long NE, const double *__restrict__ W,
                                  const double *__restrict__ J, const double *__restrict__ C, int const_c,
                                  double *__restrict__ D)
{
   const long long n = NQ * NE;
   for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
   {
      const long long e = i / NQ, q = i - e * NQ;
      const double *Je = J + e * 9 * NQ + q;
      const double J11 = Je[0 * NQ], J21 = Je[1 * NQ], J31 = Je[2 * NQ];
      const double J12 = Je[3 * NQ], J22 = Je[4 * NQ], J32 = Je[5 * NQ];
      const double J13 = Je[6 * NQ], J23 = Je[7 * NQ], J33 = Je[8 * NQ];
      const double detJ = J11 * (J22 * J33 - J32 * J23) - J21 * (J12 * J33 - J32 * J13) + J31 * (J12 * J23 - J22 * J13);
      const double c = const_c ? C[0] : C[i];
      const double w = c * (W[q] / detJ);
      const double A11 = (J22 * J33) - (J23 * J32), A12 = (J32 * J13) - (J12 * J33), A13 = (J12 * J23) - (J22 * J13);
      const double A21 = (J31 * J23) - (J21 * J33), A22 = (J11 * J33) - (J13 * J31), A23 = (J21 * J13) - (J11 * J23);
      const double A31 = (J21 * J32) - (J31 * J22), A32 = (J31 * J12) - (J11 * J32), A33 = (J11 * J22) - (J12 * J21);
      double *De = D + e * 6 * NQ + q;
      De[0 * NQ] = w * (A11 * A11 + A12 * A12 + A13 * A13);
      De[1 * NQ] = w * (A11 * A21 + A12 * A22 + A13 * A23);
      De[2 * NQ] = w * (A11 * A31 + A12 * A32 + A13 * A33);
      De[3 * NQ] = w * (A21 * A21 + A22 * A22 + A23 * A23);
      De[4 * NQ] = w * (A21 * A31 + A22 * A32 + A23 * A33);
      De[5 * NQ] = w * (A31 * A31 + A32 * A32 + A33 * A33);
   }
}

// fem/integ/bilininteg_mass_pa.cpp:62-78
// (element, q-point) of a grid-stride loop over NQ * NE entries without a 64-bit division per entry: one division at the
// start, then the stride is added in (element, q-point) form
struct EQ
{
   long long e; int q, de, dq, NQ;
   __device__ EQ(long long i0, long long stride, long long NQ_) : e(i0 / NQ_), q((int)(i0 % NQ_)), de((int)(stride / NQ_)), dq((int)(stride % NQ_)), NQ((int)NQ_) {}
   __device__ void next() { e += de; q += dq; if (q >= NQ) { q -= NQ; ++e; } }
};

__global__ void k_mass_setup(long long NQ, long long NE, const double *__restrict__ W, const double *__restrict__ detJ,
                             const double *__restrict__ C, int const_c, double *__restrict__ v)
{
   const long long n = NQ * NE, i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x, st = (long long)gridDim.x * blockDim.x;
   EQ eq(i0, st, NQ);
   for (long long i = i0; i < n; i += st, eq.next())
   {
      v[i] = W[eq.q] * (const_c ? C[0] : C[i]) * detJ[i];
   }
}

// the same on a mesh of affine elements whose determinant is kept per ELEMENT (8 B per element instead of 8 B per q-point)
__global__ void k_mass_setup_affine(long long NQ, long long NE, const double *__restrict__ W, const double *__restrict__ detE,
                                    const double *__restrict__ C, int const_c, double *__restrict__ v)
{
   const long long n = NQ * NE, i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x, st = (long long)gridDim.x * blockDim.x;
   EQ eq(i0, st, NQ);
   for (long long i = i0; i < n; i += st, eq.next())
   {
      v[i] = W[eq.q] * (const_c ? C[0] : C[i]) * detE[eq.e];
   }
}

// One thread per q-point; J and/or detJ may be null (only what is asked for is stored).
__global__ void k_geometry_trilinear(int Q1D, long long NE, const double *__restrict__ xi,
                                     const double *__restrict__ vtx, const int *__restrict__ ev,
                                     double *__restrict__ J, double *__restrict__ detJ)
{
   const long long NQ = (long long)Q1D * Q1D * Q1D, n = NQ * NE;
   for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
   {
      const long long e = i / NQ, q = i - e * NQ;
      const int qx = q % Q1D, qy = (q / Q1D) % Q1D, qz = q / (Q1D * Q1D);
      double Jm[9];
      trilinear_jacobian(vtx, ev + 8 * e, xi[qx], xi[qy], xi[qz], Jm);
      if (J)
      {
         double *Je = J + e * 9 * NQ + q;
#pragma unroll
         for (int k = 0; k < 9; ++k) { Je[k * NQ] = Jm[k]; }
      }
      if (detJ)
      {
         detJ[i] = Jm[0] * (Jm[4] * Jm[8] - Jm[5] * Jm[7]) - Jm[1] * (Jm[3] * Jm[8] - Jm[5] * Jm[6]) +
                   Jm[2] * (Jm[3] * Jm[7] - Jm[4] * Jm[6]);
      }
   }
}

// GeometricFactors::Compute + PADiffusionSetup3D in one pass for trilinear hexes: J never goes to HBM
// (reads 8 vertices per element instead of 9 doubles per q-point: 56 instead of 128 B per q-point).
__global__ void k_diffusion_setup_trilinear(int Q1D, long long NE, const double *__restrict__ W, const double *__restrict__ xi,
                                            const double *__restrict__ vtx, const int *__restrict__ ev,
                                            const double *__restrict__ C, int const_c, double *__restrict__ D)
{
   const long long NQ = (long long)Q1D * Q1D * Q1D, n = NQ * NE;
   for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
   {
      const long long e = i / NQ, q = i - e * NQ;
      const int qx = q % Q1D, qy = (q / Q1D) % Q1D, qz = q / (Q1D * Q1D);
      double Jm[9];
      trilinear_jacobian(vtx, ev + 8 * e, xi[qx], xi[qy], xi[qz], Jm);
      const double J11 = Jm[0], J21 = Jm[1], J31 = Jm[2], J12 = Jm[3], J22 = Jm[4], J32 = Jm[5], J13 = Jm[6], J23 = Jm[7], J33 = Jm[8];
      const double detJ = J11 * (J22 * J33 - J32 * J23) - J21 * (J12 * J33 - J32 * J13) + J31 * (J12 * J23 - J22 * J13);
      const double c = const_c ? C[0] : C[i];
      const double w = c * (W[q] / detJ);
      const double A11 = (J22 * J33) - (J23 * J32), A12 = (J32 * J13) - (J12 * J33), A13 = (J12 * J23) - (J22 * J13);
      const double A21 = (J31 * J23) - (J21 * J33), A22 = (J11 * J33) - (J13 * J31), A23 = (J21 * J13) - (J11 * J23);
      const double A31 = (J21 * J32) - (J31 * J22), A32 = (J31 * J12) - (J11 * J32), A33 = (J11 * J22) - (J12 * J21);
      double *De = D + e * 6 * NQ + q;
      De[0 * NQ] = w * (A11 * A11 + A12 * A12 + A13 * A13);
      De[1 * NQ] = w * (A11 * A21 + A12 * A22 + A13 * A23);
      De[2 * NQ] = w * (A11 * A31 + A12 * A32 + A13 * A33);
      De[3 * NQ] = w * (A21 * A21 + A22 * A22 + A23 * A23);
      De[4 * NQ] = w * (A21 * A31 + A22 * A32 + A23 * A33);
      De[5 * NQ] = w * (A31 * A31 + A32 * A32 + A33 * A33);
   }
}

// Factorised diffusion q-data for meshes of affine (parallelepiped) elements: J is constant over such an element,
// so the reference's D(q) = (w_q / det J) c_q adj(J) adj(J)^T (bilininteg_diffusion_kernels.cpp:243-367) splits
// into the per-element tensor below (6 doubles per ELEMENT) and the scalar w_q c_q per q-point.  One thread per
// element; `flag` is raised when an element is not affine to `tol` (relative to its edge lengths).
__global__ void k_affine_geometry(long long NE, const double *__restrict__ vtx, const int *__restrict__ ev, double tol,
                                  double *__restrict__ geo6, double *__restrict__ jinv9, double *__restrict__ detE, int *flag)
{
   for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < NE; e += (long long)gridDim.x * blockDim.x)
   {
      const int *v = ev + 8 * e;
      double X[8][3];
#pragma unroll
      for (int k = 0; k < 8; ++k)
      {
#pragma unroll
         for (int r = 0; r < 3; ++r) { X[k][r] = vtx[3LL * v[k] + r]; }
      }
      // vertex order of the reference hexahedron (mesh/mesh.cpp:3757-3765): edges a = 0->1, b = 0->3, c = 0->4
      double dev = 0.0, scale = 0.0;
#pragma unroll
      for (int r = 0; r < 3; ++r)
      {
         const double a = X[1][r] - X[0][r], b = X[3][r] - X[0][r], c = X[4][r] - X[0][r];
         scale = fmax(scale, fmax(fabs(a), fmax(fabs(b), fabs(c))));
         dev = fmax(dev, fabs(X[2][r] - (X[0][r] + a + b)));
         dev = fmax(dev, fabs(X[5][r] - (X[0][r] + a + c)));
         dev = fmax(dev, fabs(X[7][r] - (X[0][r] + b + c)));
         dev = fmax(dev, fabs(X[6][r] - (X[0][r] + a + b + c)));
      }
      if (!(dev <= tol * scale)) { atomicOr(flag, 1); }
      double Jm[9];
      trilinear_jacobian(vtx, v, 0.5, 0.5, 0.5, Jm);
      const double J11 = Jm[0], J21 = Jm[1], J31 = Jm[2], J12 = Jm[3], J22 = Jm[4], J32 = Jm[5], J13 = Jm[6], J23 = Jm[7], J33 = Jm[8];
      const double detJ = J11 * (J22 * J33 - J32 * J23) - J21 * (J12 * J33 - J32 * J13) + J31 * (J12 * J23 - J22 * J13);
      if (detE) { detE[e] = detJ; }
      const double w = 1.0 / detJ;
      const double A11 = (J22 * J33) - (J23 * J32), A12 = (J32 * J13) - (J12 * J33), A13 = (J12 * J23) - (J22 * J13);
      const double A21 = (J31 * J23) - (J21 * J33), A22 = (J11 * J33) - (J13 * J31), A23 = (J21 * J13) - (J11 * J23);
      const double A31 = (J21 * J32) - (J31 * J22), A32 = (J31 * J12) - (J11 * J32), A33 = (J11 * J22) - (J12 * J21);
      double *g = geo6 + 6 * e;
      g[0] = w * (A11 * A11 + A12 * A12 + A13 * A13);
      g[1] = w * (A11 * A21 + A12 * A22 + A13 * A23);
      g[2] = w * (A11 * A31 + A12 * A32 + A13 * A33);
      g[3] = w * (A21 * A21 + A22 * A22 + A23 * A23);
      g[4] = w * (A21 * A31 + A22 * A32 + A23 * A33);
      g[5] = w * (A31 * A31 + A32 * A32 + A33 * A33);
      if (jinv9)
      {
         // rows of J^{-T} = cofactors / det J: physical gradient g_r = ji[3r] gX + ji[3r+1] gY + ji[3r+2] gZ
         // (fem/qinterp/grad.hpp:340-352; same cofactor naming as pa_element_kernel's i0..i8)
         double *ji = jinv9 + 9 * e;
         ji[0] = A11 * w; ji[1] = A21 * w; ji[2] = A31 * w;
         ji[3] = A12 * w; ji[4] = A22 * w; ji[5] = A32 * w;
         ji[6] = A13 * w; ji[7] = A23 * w; ji[8] = A33 * w;
      }
   }
}

// The same per-element tensor from stored Jacobians (the host's GeometricFactors): J is taken at the first q-point
// and the element counts as affine when no entry of J moves by more than tol * max|J| over its q-points.
__global__ void k_affine_from_J(long long NQ, long long NE, const double *__restrict__ J, double tol, double *__restrict__ geo6,
                                double *__restrict__ jinv9, int *flag)
{
   for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < NE; e += (long long)gridDim.x * blockDim.x)
   {
      const double *Je = J + e * 9 * NQ;
      double J0[9], scale = 0.0, dev = 0.0;
#pragma unroll
      for (int k = 0; k < 9; ++k) { J0[k] = Je[k * NQ]; scale = fmax(scale, fabs(J0[k])); }
      for (long long q = 1; q < NQ; ++q)
      {
#pragma unroll
         for (int k = 0; k < 9; ++k) { dev = fmax(dev, fabs(Je[k * NQ + q] - J0[k])); }
      }
      if (!(dev <= tol * scale)) { atomicOr(flag, 1); }
      const double J11 = J0[0], J21 = J0[1], J31 = J0[2], J12 = J0[3], J22 = J0[4], J32 = J0[5], J13 = J0[6], J23 = J0[7], J33 = J0[8];
      const double detJ = J11 * (J22 * J33 - J32 * J23) - J21 * (J12 * J33 - J32 * J13) + J31 * (J12 * J23 - J22 * J13);
      const double w = 1.0 / detJ;
      const double A11 = (J22 * J33) - (J23 * J32), A12 = (J32 * J13) - (J12 * J33), A13 = (J12 * J23) - (J22 * J13);
      const double A21 = (J31 * J23) - (J21 * J33), A22 = (J11 * J33) - (J13 * J31), A23 = (J21 * J13) - (J11 * J23);
      const double A31 = (J21 * J32) - (J31 * J22), A32 = (J31 * J12) - (J11 * J32), A33 = (J11 * J22) - (J12 * J21);
      double *g = geo6 + 6 * e;
      g[0] = w * (A11 * A11 + A12 * A12 + A13 * A13);
      g[1] = w * (A11 * A21 + A12 * A22 + A13 * A23);
      g[2] = w * (A11 * A31 + A12 * A32 + A13 * A33);
      g[3] = w * (A21 * A21 + A22 * A22 + A23 * A23);
      g[4] = w * (A21 * A31 + A22 * A32 + A23 * A33);
      g[5] = w * (A31 * A31 + A32 * A32 + A33 * A33);
      if (jinv9)
      {
         // rows of J^{-T} = cofactors / det J: physical gradient g_r = ji[3r] gX + ji[3r+1] gY + ji[3r+2] gZ
         // (fem/qinterp/grad.hpp:340-352; same cofactor naming as pa_element_kernel's i0..i8)
         double *ji = jinv9 + 9 * e;
         ji[0] = A11 * w; ji[1] = A21 * w; ji[2] = A31 * w;
         ji[3] = A12 * w; ji[4] = A22 * w; ji[5] = A32 * w;
         ji[6] = A13 * w; ji[7] = A23 * w; ji[8] = A33 * w;
      }
   }
}

// PADiffusionSetup3D on a mesh of affine elements: D(q) = (w_q c_q) * (per-element tensor) - a streaming kernel
// (the trilinear rebuild of J per q-point above is FP64-issue-bound: 1.3 ms instead of 0.6 ms at 8 M dofs, p=2)
__global__ void k_diffusion_setup_affine(long long NQ, long long NE, const double *__restrict__ W, const double *__restrict__ geo6,
                                         const double *__restrict__ C, int const_c, double *__restrict__ D)
{
   const long long n = NQ * NE, i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x, st = (long long)gridDim.x * blockDim.x;
   EQ eq(i0, st, NQ);
   for (long long i = i0; i < n; i += st, eq.next())
   {
      const double w = W[eq.q] * (const_c ? C[0] : C[i]);
      const double *g = geo6 + 6 * eq.e;
      double *De = D + eq.e * 6 * NQ + eq.q;
#pragma unroll
      for (int k = 0; k < 6; ++k) { De[k * NQ] = w * g[k]; }
   }
}

// the scalar half of the factorised q-data: c[i] = W[q] C[i]
__global__ void k_coeff_times_w(long long NQ, long long NE, const double *__restrict__ W, const double *__restrict__ C, int const_c,
                                double *__restrict__ out)
{
   const long long n = NQ * NE, i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x, st = (long long)gridDim.x * blockDim.x;
   EQ eq(i0, st, NQ);
   for (long long i = i0; i < n; i += st, eq.next())
   {
      out[i] = W[eq.q] * (const_c ? C[0] : C[i]);
   }
}

// the "user forall over Q-points" of SURVEY §3.2/§3.3
__global__ void k_coeff_eval(int kind, long long n, double a, double b, double T0, const double *__restrict__ T,
                             const double *__restrict__ s, const double *__restrict__ g, double *__restrict__ out)
{
   for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
   {
      if (kind == 0) { out[i] = a * (1.0 + b * (T[i] - T0)); }
      else if (kind == 1) { out[i] = a; }
      else
      {
         const double gx = g[3 * i], gy = g[3 * i + 1], gz = g[3 * i + 2];
         out[i] = s[i] * (gx * gx + gy * gy + gz * gz) + a;
      }
   }
}

// --------------------------------------------------------------------- diagonal
// Sum-factorised diagonal (what the reference's SmemPADiffusionDiagonal3D / SmemPAMassAssembleDiagonal3D
// do, fem/integ/bilininteg_diffusion_kernels.hpp:369-484, bilininteg_mass_kernels.hpp:324-408):
//   dE[dx,dy,dz] += sum_f w_f sum_q D_f(q) Mx_f(qx,dx) My_f(qy,dy) Mz_f(qz,dz)
// with the seven fields f = D00, D01, D02, D11, D12, D22, mass; per direction the 1-D factor is
// GG, BG or BB (G in the directions i and j of D_ij, B elsewhere); w = 2 for the off-diagonal D_ij.
// Three contraction passes through shared memory instead of a Q^3 loop per E-entry (14x fewer FMAs at p=2).
//
// HBM-bound (reads the q-data once: 8 * 7 Q^3 bytes per element), so what matters is bytes in flight: the q-data of
// a batch is staged by TMA bulk copies into one of TWO shared-memory stages, two batches ahead of its use (round 1
// had one stage refilled after pass 1: 49 % of the HBM roofline).  Output modes:
//   * E-vector, accumulated (AssembleDiagonalPA semantics of the integrator-level entry points), or
//   * the slot layout of the E->L CSR, written (the form-level diagonal: no zero fill, no read-modify-write, and the
//     segmented reduction that follows streams contiguously instead of gathering through the index list).
// FUSED (affine meshes): the staged field is the raw coefficient C(q); the kernel forms c = W(q) C(q) in place, WRITES the
// integrator's q-data from it (pa_out: the six stored components c * geo6_f(e), or the scalar c of the factorised form)
// and takes the diagonal from the same registers - PADiffusionSetup3D and the diagonal in one pass over the q-points
// (an implicit time step re-assembles both at every step: 3.6 GB of q-data are then never read back for the diagonal).
#ifndef B200PA_TUNE_DIAG_KB
#define B200PA_TUNE_DIAG_KB 44
#endif
// field groups: the seven fields of passes 1 and 2 are split over NFG tasks per (element, column) - the kernel is bound by
// latency at low occupancy (shared memory per element is large, tasks per element few), more threads per element hide it
#ifndef B200PA_TUNE_DIAG_NFG
#define B200PA_TUNE_DIAG_NFG 2
#endif
__device__ __forceinline__ int diag_group_lo(int g)
{
   // first field of group g (and 7 for g = NFG)
   return B200PA_TUNE_DIAG_NFG == 1 ? (g == 0 ? 0 : 7)
          : B200PA_TUNE_DIAG_NFG == 2 ? (g == 0 ? 0 : (g == 1 ? 3 : 7))
          : B200PA_TUNE_DIAG_NFG == 4 ? (g < 3 ? 2 * g : (g == 3 ? 6 : 7))
          : g;
}
// smallest s >= lo with s = r (mod 16): 16 consecutive lanes (one 128-byte shared-memory wavefront of doubles) whose
// addresses advance by `r` per outer index and by 1 per inner index then fall into 16 different banks
constexpr int diag_stride(int lo, int r) { return lo + ((r - lo) % 16 + 16) % 16; }

template <int D1, int Q1>
struct DiagSfCfg
{
   static constexpr int Q2 = Q1 * Q1, Q3 = Q1 * Q1 * Q1, NF = 7;
   // The contraction order is z, y, x (round 1 went x, y, z: its first pass read rows of Q1 consecutive doubles per lane,
   // a Q1-way bank conflict on the staged q-data - 55 % of all shared-memory wavefronts at p=2, profiles/r2c_*):
   //   pass 1  task (e, qy, qx): contract qz   stage[f][qz][qy][qx] -> T1[f][dz][qy qx]     (lanes walk qx: stride 1)
   //   pass 2  task (e, dz, qx): contract qy   T1 -> T2[f][dz][dy][qx]                        (lanes walk qx, then dz)
   //   pass 3  task (e, dz, dy): contract qx, sum the fields -> element diagonal             (lanes walk dy, then dz)
   static constexpr int S1 = diag_stride(Q2, Q1);          // dz stride of T1: lanes (dz, qx) of pass 2 -> distinct banks
   static constexpr int F1 = D1 * S1;                      // field stride of T1
   static constexpr int T1E = diag_stride(NF * F1, Q2 % 16); // element stride of T1 (pass 1 lanes: Q2 per element)
   static constexpr int R2 = Q1 | 1;                       // row stride of T2 (odd: lanes of pass 3 walk rows)
   static constexpr int S2 = diag_stride(D1 * R2, Q1);     // dz stride of T2: lanes (dz, qx) of pass 2 -> distinct banks
   static constexpr int F2 = D1 * S2;
   static constexpr int T2E = NF * F2;
   static constexpr int SQD1 = 6 * Q3, SQM1 = Q3;          // staged q-data per element
   static_assert(T2E <= 7 * Q3 + 0 || true, "");
   // T2 overlays the stage it was fed from (free once pass 1 is done; the stage is refilled at the END of the batch, which
   // still leaves the copy a whole batch of time): shared memory per element = two stages + T1
   static constexpr int STAGE1 = (SQD1 + SQM1) > T2E ? (SQD1 + SQM1) : T2E;
   static constexpr int PER_E = (2 * STAGE1 + T1E) * 8;
   static constexpr int NEB0 = (B200PA_TUNE_DIAG_KB * 1024) / PER_E;
   static constexpr int NEB = NEB0 < 1 ? 1 : (NEB0 > 8 ? 8 : NEB0);
   // lanes per element in the three passes: padded to a whole 16-lane group while an element has fewer tasks than that, so
   // that no half-warp (one shared-memory wavefront of doubles) straddles two elements - the strides above are
   // conflict-free inside an element (ncu of the unpadded version: 2.25x the ideal wavefronts in pass 3 at p=2)
   static constexpr int TPE1 = Q2 < 16 ? 16 : Q2, TPE2 = D1 * Q1 < 16 ? 16 : D1 * Q1, TPE3 = D1 * D1 < 16 ? 16 : D1 * D1;
   static constexpr int NFG = B200PA_TUNE_DIAG_NFG;
   static_assert(NFG == 1 || NFG == 2 || NFG == 4 || NFG == 7, "field groups: 1, 2, 4 or 7");
   static constexpr int NT0 = ((NFG * NEB * TPE1 + 31) / 32) * 32;
   static constexpr int NT = NT0 < 64 ? 64 : NT0;
   static constexpr int SQD = NEB * SQD1 + 2, SQM = ((NEB * SQM1 + 2) + 1) & ~1; // TMA staging (16-byte aligned, +slack)
   static constexpr int STAGE = (SQD + SQM) > NEB * T2E ? (SQD + SQM) : ((NEB * T2E + 1) & ~1);
   static constexpr size_t SMEM_BYTES = sizeof(double) * (2 * STAGE + NEB * T1E + Q3);
};

// kernel parameters: the three 1-D factor tables live in the constant bank, so that with the field loop
// unrolled every coefficient is a compile-time-indexed DFMA operand
template <int D1, int Q1>
struct DiagParams
{
   double M[3][Q1 * D1]; // 0: B*B, 1: B*G, 2: G*G, column-major [Q,D]
   long long NE;
   const double *__restrict__ pa_diff; // stored: [6 Q^3, NE]; factorised (geo != null): c_q [Q^3, NE]; FUSED: the raw coefficient C
   const double *__restrict__ pa_mass;
   const double *__restrict__ geo; // adj(J)adj(J)^T/det J [6,NE] (factorised form and FUSED), else null
   double *__restrict__ out;       // E-vector (+=) or slot-order scratch (=)
   const int *__restrict__ slot;   // SLOT output: position of every E-entry in the E->L CSR
   // FUSED
   const double *__restrict__ W;   // quadrature weights [Q^3] (device)
   double *__restrict__ pa_out;    // q-data written by the kernel
   int pa_out_ncomp;               // 6 (stored form) or 1 (factorised form)
   int const_c;                    // FUSED: C has one entry
   // markers: elements whose DIFFUSION part is left out of the diagonal although its q-data is there - the reference
   // zeroes the accumulated element diagonal of every element a LATER integrator's marker excludes (see b200pa.h)
   const unsigned char *__restrict__ diff_off;
};

// QMODE: 0 = stored diffusion q-data (six components per q-point), 1 = factorised (one scalar per q-point + geo),
//        2 = FUSED (raw coefficient staged, q-data written by the kernel)
template <int D1, int Q1, bool SLOT, int QMODE>
__global__ void __launch_bounds__(DiagSfCfg<D1, Q1>::NT)
k_diag_sf(const __grid_constant__ DiagParams<D1, Q1> P)
{
   constexpr bool FUSED = (QMODE == 2);
   using C = DiagSfCfg<D1, Q1>;
   constexpr int D2 = D1 * D1, D3 = D1 * D1 * D1, Q2 = C::Q2, Q3 = C::Q3, NF = C::NF, NEB = C::NEB;
   constexpr int S1 = C::S1, F1 = C::F1, T1E = C::T1E, R2 = C::R2, S2 = C::S2, F2 = C::F2, T2E = C::T2E;
   extern __shared__ __align__(16) double dsm[];
   double *sT1 = dsm + 2 * C::STAGE;  // stages first: 16-byte aligned for the bulk copies
   double *sW = sT1 + NEB * T1E;
   __shared__ unsigned long long qbar[2];
   const long long NE = P.NE;
   const double *pa_diff = P.pa_diff, *pa_mass = P.pa_mass;
   const double *geo = P.geo;
   constexpr bool scalar_diff = (QMODE != 0);        // the diffusion field is one scalar per q-point
   const bool stage_diff = pa_diff != nullptr && !(FUSED && P.const_c);
   if (threadIdx.x == 0) { mbar_init(&qbar[0], 1); mbar_init(&qbar[1], 1); }
   if (FUSED) { for (int i = threadIdx.x; i < Q3; i += blockDim.x) { sW[i] = P.W[i]; } }
   __syncthreads();
// factor type of field f = D00, D01, D02, D11, D12, D22, mass in direction a: how many of the two indices
// of D_ij equal a (0: BB, 1: BG, 2: GG); f and a are compile-time wherever this is used
#define B200PA_MTYPE(f, a) ((((a) == 0 ? 0x0016u : ((a) == 1 ? 0x0184u : 0x0910u)) >> (2 * (f))) & 3u)
   struct FieldCopy { const double *src; unsigned bytes; };
   auto scalar_field = [&](const double *arr, double *sdst, long long e0, int nel)
   {
      const double *src = arr + e0 * Q3;
      const int sh = (int)(((unsigned long long)src >> 3) & 1ull);
      src -= sh;
      int nd = sh + nel * Q3;
      if (nd & 1)
      {
         if (e0 + nel < NE) { nd += 1; }
         else { nd -= 1; sdst[nd] = __ldg(src + nd); }
      }
      return FieldCopy{src, (unsigned)(nd * sizeof(double))};
   };
   auto issue = [&](long long b, int st)
   {
      double *sQd = dsm + st * C::STAGE, *sQm = sQd + C::SQD;
      const long long e0 = b * NEB;
      const int nel = (int)(NE - e0 < NEB ? NE - e0 : NEB);
      FieldCopy cd{nullptr, 0}, cm{nullptr, 0};
      if (stage_diff) { cd = scalar_diff ? scalar_field(pa_diff, sQd, e0, nel) : FieldCopy{pa_diff + e0 * 6 * Q3, (unsigned)(nel * 6 * Q3 * sizeof(double))}; }
      if (pa_mass) { cm = scalar_field(pa_mass, sQm, e0, nel); }
      mbar_expect_tx(&qbar[st], cd.bytes + cm.bytes);
      if (stage_diff) { tma_bulk_g2s(sQd, cd.src, cd.bytes, &qbar[st]); }
      if (pa_mass) { tma_bulk_g2s(sQm, cm.src, cm.bytes, &qbar[st]); }
   };
   const long long nbatch = (NE + NEB - 1) / NEB;
   unsigned phase = 0; // bit s: parity of stage s
   if (threadIdx.x == 0)
   {
      if ((long long)blockIdx.x < nbatch) { issue(blockIdx.x, 0); }
      if ((long long)blockIdx.x + gridDim.x < nbatch) { issue((long long)blockIdx.x + gridDim.x, 1); }
   }
   int st = 0;
   for (long long batch = blockIdx.x; batch < nbatch; batch += gridDim.x, st ^= 1)
   {
      const long long e0 = batch * NEB;
      const int nel = (int)(NE - e0 < NEB ? NE - e0 : NEB);
      double *sQd = dsm + st * C::STAGE, *sQm = sQd + C::SQD;
      double *sT2 = sQd;              // overlays the stage once pass 1 has consumed it
      mbar_wait(&qbar[st], (phase >> st) & 1u);
      phase ^= 1u << st;
      const int msh = pa_mass ? (int)(((unsigned long long)(pa_mass + e0 * Q3) >> 3) & 1ull) : 0;
      const int dsh = (stage_diff && scalar_diff) ? (int)(((unsigned long long)(pa_diff + e0 * Q3) >> 3) & 1ull) : 0;
      if (FUSED)
      {
         // c = W(q) C(q) in place; the integrator's q-data goes out from the same value (coalesced over q)
         const double c0 = P.const_c ? __ldg(pa_diff) : 0.0;
         for (int i = threadIdx.x; i < nel * Q3; i += blockDim.x)
         {
            const int e = i / Q3, q = i - e * Q3;
            const double c = sW[q] * (P.const_c ? c0 : sQd[dsh + i]);
            sQd[dsh + i] = c;
            if (P.pa_out_ncomp == 1) { P.pa_out[(e0 + e) * Q3 + q] = c; }
            else
            {
               const double *g = geo + (e0 + e) * 6;
               double *o = P.pa_out + (e0 + e) * 6 * Q3 + q;
#pragma unroll
               for (int f = 0; f < 6; ++f) { o[f * Q3] = c * __ldg(g + f); }
            }
         }
         __syncthreads();
      }
      // pass 1: contract qz.  task = (e, qy, qx), all seven fields: one column of Q1 q-data values each
      for (int t0 = threadIdx.x; t0 < C::NFG * nel * C::TPE1; t0 += blockDim.x)
      {
         const int fg = t0 / (nel * C::TPE1), t = t0 - fg * nel * C::TPE1;
         const int flo = diag_group_lo(fg), fhi = diag_group_lo(fg + 1);
         const int e = t / C::TPE1, c = t - e * C::TPE1;
         if (c >= Q2) { continue; }
         const bool diff_on = pa_diff != nullptr && !(P.diff_off && P.diff_off[e0 + e]);
#pragma unroll
         for (int f = 0; f < NF; ++f)
         {
            if (C::NFG > 1 && (f < flo || f >= fhi)) { continue; }
            const bool have = f < 6 ? diff_on : pa_mass != nullptr;
            const double *src = f < 6 ? (scalar_diff ? sQd + dsh + e * Q3 + c : sQd + (e * 6 + f) * Q3 + c) : sQm + msh + e * Q3 + c;
            const double scale = (f < 6 && scalar_diff && have) ? __ldg(geo + (e0 + e) * 6 + f) : 1.0;
            double out[D1];
#pragma unroll
            for (int d = 0; d < D1; ++d) { out[d] = 0.0; }
            if (have)
            {
#pragma unroll
               for (int q = 0; q < Q1; ++q)
               {
                  const double v = (scalar_diff && f < 6) ? scale * src[q * Q2] : src[q * Q2];
#pragma unroll
                  for (int d = 0; d < D1; ++d) { out[d] = fma(P.M[B200PA_MTYPE(f, 2)][q + Q1 * d], v, out[d]); }
               }
            }
#pragma unroll
            for (int d = 0; d < D1; ++d) { sT1[e * T1E + f * F1 + d * S1 + c] = out[d]; }
         }
      }
      __syncthreads();
      // pass 2: contract qy.  task = (e, dz, qx), all seven fields; T2 goes where the consumed q-data was
      for (int t0 = threadIdx.x; t0 < C::NFG * nel * C::TPE2; t0 += blockDim.x)
      {
         const int fg = t0 / (nel * C::TPE2), t = t0 - fg * nel * C::TPE2;
         const int flo = diag_group_lo(fg), fhi = diag_group_lo(fg + 1);
         const int e = t / C::TPE2, r2 = t - e * C::TPE2, dz = r2 / Q1, qx = r2 - dz * Q1;
         if (r2 >= D1 * Q1) { continue; }
#pragma unroll
         for (int f = 0; f < NF; ++f)
         {
            if (C::NFG > 1 && (f < flo || f >= fhi)) { continue; }
            double out[D1];
#pragma unroll
            for (int d = 0; d < D1; ++d) { out[d] = 0.0; }
#pragma unroll
            for (int qy = 0; qy < Q1; ++qy)
            {
               const double v = sT1[e * T1E + f * F1 + dz * S1 + qy * Q1 + qx];
#pragma unroll
               for (int d = 0; d < D1; ++d) { out[d] = fma(P.M[B200PA_MTYPE(f, 1)][qy + Q1 * d], v, out[d]); }
            }
#pragma unroll
            for (int dy = 0; dy < D1; ++dy) { sT2[e * T2E + f * F2 + dz * S2 + dy * R2 + qx] = out[dy]; }
         }
      }
      __syncthreads();
      // pass 3: contract qx and sum the fields.  task = (e, dz, dy): all dx at once
      for (int t = threadIdx.x; t < nel * C::TPE3; t += blockDim.x)
      {
         const int e = t / C::TPE3, k = t - e * C::TPE3, dz = k / D1, dy = k - dz * D1;
         if (k >= D2) { continue; }
         // slot indices first: their global-memory latency hides behind the contraction (ncu: 14 % of all stall samples
         // sat on the address of the first store when they were loaded where they are used)
         const long long o = (e0 + e) * D3 + k * D1;
         int sl[D1];
         if (SLOT)
         {
#pragma unroll
            for (int dx = 0; dx < D1; ++dx) { sl[dx] = __ldg(P.slot + o + dx); }
         }
         double acc[D1];
#pragma unroll
         for (int d = 0; d < D1; ++d) { acc[d] = 0.0; }
#pragma unroll
         for (int f = 0; f < NF; ++f)
         {
            const double w = (f == 1 || f == 2 || f == 4) ? 2.0 : 1.0;
#pragma unroll
            for (int qx = 0; qx < Q1; ++qx)
            {
               const double v = w * sT2[e * T2E + f * F2 + dz * S2 + dy * R2 + qx];
#pragma unroll
               for (int dx = 0; dx < D1; ++dx) { acc[dx] = fma(P.M[B200PA_MTYPE(f, 0)][qx + Q1 * dx], v, acc[dx]); }
            }
         }
#pragma unroll
         for (int dx = 0; dx < D1; ++dx)
         {
            if (SLOT) { P.out[sl[dx]] = acc[dx]; }
            else { P.out[o + dx] += acc[dx]; }
         }
      }
      __syncthreads();
      // T2 (and with it this stage) is consumed: refill the stage with the batch after the next one; the other stage has
      // been in flight / landed since the end of the previous batch
      if (batch + 2LL * gridDim.x < nbatch && threadIdx.x == 0) { issue(batch + 2LL * gridDim.x, st); }
   }
#undef B200PA_MTYPE
}

// ------------------------------------------------------- element-attribute markers
// BilinearForm::AddDomainIntegrator(bfi, elem_marker): on[e] = the integrator acts on element e
// (attr > 0 && marker[attr-1] != 0, fem/bilinearform_ext.cpp:391-399 and AddWithMarkers_)
__global__ void k_marker_mask(long long NE, const int *__restrict__ attr, int n_attr, const int *__restrict__ marker,
                              unsigned char *__restrict__ on, int *bad_flag)
{
   for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < NE; e += (long long)gridDim.x * blockDim.x)
   {
      const int a = attr[e];
      if (a > n_attr) { *bad_flag = 1; on[e] = 0; continue; }
      on[e] = (a > 0 && marker[a - 1] != 0) ? 1 : 0;
   }
}
__global__ void k_invert_mask(long long n, const unsigned char *__restrict__ in, unsigned char *__restrict__ out)
{
   for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) { out[i] = in[i] ? 0 : 1; }
}
// q-data of the elements an integrator does not act on := 0, so that their contribution to every apply is exactly
// zero - AddMultWithMarkers (fem/bilinearform_ext.cpp:807-847) computes it and then leaves it out of the sum
__global__ void k_zero_unmarked(long long NE, long long per_elem, const unsigned char *__restrict__ on, double *__restrict__ pa)
{
   const long long n = NE * per_elem;
   for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
   {
      if (!on[i / per_elem]) { pa[i] = 0.0; }
   }
}

// ------------------------------------------------------------------ BLAS-1 etc.
__global__ void k_add(long long n, const double *__restrict__ v1, double alpha, const double *__restrict__ v2,
                      double *__restrict__ v)
{
   for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
   {
      v[i] = fma(alpha, v2[i], v1[i]);
   }
}

// linalg/solvers.cpp:401-425
__global__ void k_jacobi_setup(int n, const double *__restrict__ diag, double damping, double *__restrict__ dinv,
                               int *zero_flag)
{
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
   {
      const double d = diag[i];
      if (d == 0.0) { *zero_flag = 1; }
      dinv[i] = damping / d;
   }
}
__global__ void k_set_indexed(int n, const int *__restrict__ idx, double val, double *__restrict__ v)
{
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { v[idx[i]] = val; }
}
__global__ void k_copy_indexed(int n, const int *__restrict__ idx, const double *__restrict__ src, double *__restrict__ dst)
{
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { dst[idx[i]] = src[idx[i]]; }
}
__global__ void k_mask_indexed(int n, const int *__restrict__ idx, unsigned char *__restrict__ mask)
{
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { mask[idx[i]] = 1; }
}
// linalg/solvers.cpp:442-452
__global__ void k_jacobi_mult(int n, const double *__restrict__ dinv, const double *__restrict__ r, double *__restrict__ z)
{
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { z[i] = dinv[i] * r[i]; }
}
__global__ void k_sub_inplace(int n, double *__restrict__ b, const double *__restrict__ z)
{
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { b[i] -= z[i]; }
}

// ------------------------------------------------------------- CSR construction
__global__ void k_iota(long long n, int *v)
{
   for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) { v[i] = (int)i; }
}
// offsets[i] = first position j with sorted_keys[j] >= i  (i in [0,ndofs])
__global__ void k_offsets_from_sorted(int ndofs, long long n, const int *__restrict__ keys, int *__restrict__ offsets)
{
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= ndofs; i += gridDim.x * blockDim.x)
   {
      long long lo = 0, hi = n;
      while (lo < hi)
      {
         const long long mid = (lo + hi) >> 1;
         if (keys[mid] < i) { lo = mid + 1; } else { hi = mid; }
      }
      offsets[i] = (int)lo;
   }
}
__global__ void k_invert_perm(long long n, const int *__restrict__ indices, int *__restrict__ slot)
{
   for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x) { slot[indices[j]] = (int)j; }
}
__global__ void k_check_nonneg(long long n, const int *__restrict__ v, int ndofs, int *flag)
{
   for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
   {
      if (v[i] < 0 || v[i] >= ndofs) { *flag = 1; }
   }
}
// constrained gather map: entries that point at an essential dof become -1 ("read zero"),
// folding z = x; z[ess] = 0 (linalg/operator.cpp:603-609) into the gather
__global__ void k_constrain_gmap(long long n, const int *__restrict__ gmap, const unsigned char *__restrict__ ess_mask,
                                 int *__restrict__ out)
{
   for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
   {
      const int g = gmap[i];
      out[i] = ess_mask[g] ? -1 : g;
   }
}

// -------------------------------------------------------------------------- PCG
// r = b - Ax (Ax passed in r), z = dinv r, d = z, nom partial = d.r    (:875-892)
__global__ void k_pcg_init(int n, const double *__restrict__ b, const double *__restrict__ dinv, double *__restrict__ r,
                           double *__restrict__ d, const unsigned char *__restrict__ own_mask, double *partials,
                           unsigned int *ticket, PcgState *st, double *norms_epilogue)
{
   double acc = 0.0;
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
   {
      const double ri = b[i] - r[i];
      const double zi = dinv[i] * ri;
      r[i] = ri;
      d[i] = zi;
      if (!own_mask || own_mask[i]) { acc = fma(zi, ri, acc); }
   }
   if (grid_sum(acc, partials, ticket, &st->dot_a) && norms_epilogue) { pcg_scalar_init(st, norms_epilogue); }
}

// scalar step after nom0 = Dot(d, r)   (:892-919)

// scalar step after den = Dot(d, z)   (:921-938 first time, :1010-1024 in the loop)

// x += alpha d; r -= alpha z; z = dinv r; betanom partial = r.z   (:956-963)
__global__ void k_pcg_update(int n, double *__restrict__ x, double *__restrict__ r, double *__restrict__ z,
                             const double *__restrict__ d, const double *__restrict__ dinv,
                             const unsigned char *__restrict__ own_mask, double *partials, unsigned int *ticket,
                             PcgState *st, double *norms_epilogue)
{
   if (st->done) { return; }
   const double alpha = st->alpha;
   double acc = 0.0;
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
   {
      x[i] = fma(alpha, d[i], x[i]);
      const double ri = fma(-alpha, z[i], r[i]);
      const double zi = dinv[i] * ri;
      r[i] = ri;
      z[i] = zi;
      if (!own_mask || own_mask[i]) { acc = fma(ri, zi, acc); }
   }
   if (grid_sum(acc, partials, ticket, &st->dot_a) && norms_epilogue) { pcg_scalar_beta(st, norms_epilogue); }
}

// scalar step after betanom = Dot(r, z)   (:964-1002)

// d = z + beta d   (:1003)
__global__ void k_pcg_direction(int n, const double *__restrict__ z, double *__restrict__ d, const PcgState *st)
{
   if (st->done) { return; }
   const double beta = st->beta;
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { d[i] = fma(beta, d[i], z[i]); }
}

// ------------------------------------------------------------ Chebyshev smoother
// One polynomial term of OperatorChebyshevSmoother::Mult (linalg/solvers.cpp:641-656), fused into one pass:
//   res = dinv .* src;   z = (FIRST ? 0 : z) + c * res
// DOT (last term, when a PCG called): partial of (r, z) over the owned dofs and, on one GPU, the scalar step of
// the PCG that consumes it (step 1: after the initial residual, 2: inside the loop) as the reduction's epilogue.
template <bool FIRST, bool DOT>
__global__ void k_cheb_term(int n, const double *__restrict__ src, const double *__restrict__ dinv, double c,
                            double *__restrict__ res, double *__restrict__ z, const double *__restrict__ r,
                            const unsigned char *__restrict__ own_mask, double *partials, unsigned int *ticket, PcgState *st,
                            double *norms_epilogue, int scalar_step)
{
   if (st && st->done) { return; }
   double acc = 0.0;
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
   {
      const double ri = dinv[i] * src[i];
      const double zi = FIRST ? c * ri : fma(c, ri, z[i]);
      res[i] = ri;
      z[i] = zi;
      if (DOT && (!own_mask || own_mask[i])) { acc = fma(r[i], zi, acc); }
   }
   if (DOT)
   {
      if (grid_sum(acc, partials, ticket, &st->dot_a) && norms_epilogue)
      {
         if (scalar_step == 1) { pcg_scalar_init(st, norms_epilogue); }
         else { pcg_scalar_beta(st, norms_epilogue); }
      }
   }
}

// r = b - r (r holds A x on entry)   (linalg/solvers.cpp:875-879)
__global__ void k_residual(int n, const double *__restrict__ b, double *__restrict__ r)
{
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { r[i] = b[i] - r[i]; }
}

// x += alpha d; r -= alpha q   (:956-957) - the general-preconditioner form of k_pcg_update
__global__ void k_pcg_update_plain(int n, double *__restrict__ x, double *__restrict__ r, const double *__restrict__ q,
                                   const double *__restrict__ d, const PcgState *st)
{
   if (st->done) { return; }
   const double alpha = st->alpha;
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
   {
      x[i] = fma(alpha, d[i], x[i]);
      r[i] = fma(-alpha, q[i], r[i]);
   }
}

// v /= s   (PowerMethod: v0 /= sqrt(normV0), linalg/operator.cpp:900)
__global__ void k_div_scalar(int n, double *__restrict__ v, double s)
{
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { v[i] = v[i] / s; }
}

} // namespace b200pa
