// p-multigrid building blocks: the order-refinement transfer between two H1 spaces on the same hex mesh and the small
// kernels of the V-cycle.  Reference: TensorProductPRefinementTransferOperator::{Mult, MultTranspose} with
// TransferKernels::{Prolongation3D, Restriction3D} (fem/transfer.cpp:2223-2296, 2542-2592) and MultigridBase::Cycle
// (fem/multigrid.cpp:164-220).  Not on the apply's hot path: runtime orders, one CTA per element, sum-factorised.
#pragma once
#include <cuda_runtime.h>

#include "pcg_state.cuh"
#include "reduce.cuh"

namespace b200pa
{

struct TransferParams
{
   long long NE;
   int DC, DF;                           // 1-D dofs of the coarse / fine space
   const double *__restrict__ B;         // [DF, DC] column-major: coarse basis at the fine nodes (DofToQuad::B)
   const int *__restrict__ gmap_c;       // coarse E -> L
   const int *__restrict__ gmap_f;       // fine E -> L
   const int *__restrict__ slot_f;       // fine: position of an E-entry in the E -> L CSR
   const int *__restrict__ off_f;        // fine CSR offsets: an E-entry is the FIRST of its dof <=> slot == offsets[dof]
                                         // (ElementRestriction::BooleanMask, fem/restriction.cpp)
   const unsigned char *__restrict__ ess_c, *__restrict__ ess_f; // essential-dof masks (RectangularConstrainedOperator) or null
   const int *__restrict__ slot_c;       // restriction: coarse slot layout
};

// shared memory: B [DF*DC] | a [DF^3] | b [DF^3]
__device__ __forceinline__ void mg_load_B(const TransferParams &P, double *sB)
{
   for (int i = threadIdx.x; i < P.DF * P.DC; i += blockDim.x) { sB[i] = P.B[i]; }
}

// y_fine = P x_coarse (L -> L).  RectangularConstrainedOperator::Mult (linalg/operator.cpp): coarse essential entries of x
// count as zero, fine essential entries of y are set to zero.
__global__ void k_mg_prolong(const TransferParams P, const double *__restrict__ xc, double *__restrict__ yf)
{
   extern __shared__ double sm[];
   const int DC = P.DC, DF = P.DF, DC3 = DC * DC * DC, DF3 = DF * DF * DF;
   double *sB = sm, *a = sB + DF * DC, *b = a + DF3;
   mg_load_B(P, sB);
   for (long long e = blockIdx.x; e < P.NE; e += gridDim.x)
   {
      __syncthreads();
      for (int i = threadIdx.x; i < DC3; i += blockDim.x)
      {
         const int g = P.gmap_c[e * DC3 + i];
         a[i] = (P.ess_c && P.ess_c[g]) ? 0.0 : xc[g];
      }
      __syncthreads();
      // x: b[dz][dy][qx] = sum_dx B(qx,dx) a[dz][dy][dx]
      for (int t = threadIdx.x; t < DC * DC * DF; t += blockDim.x)
      {
         const int qx = t % DF, r = t / DF;
         double s = 0.0;
         for (int dx = 0; dx < DC; ++dx) { s = fma(sB[qx + DF * dx], a[r * DC + dx], s); }
         b[r * DF + qx] = s;
      }
      __syncthreads();
      // y: a[dz][qy][qx] = sum_dy B(qy,dy) b[dz][dy][qx]
      for (int t = threadIdx.x; t < DC * DF * DF; t += blockDim.x)
      {
         const int qx = t % DF, qy = (t / DF) % DF, dz = t / (DF * DF);
         double s = 0.0;
         for (int dy = 0; dy < DC; ++dy) { s = fma(sB[qy + DF * dy], b[(dz * DC + dy) * DF + qx], s); }
         a[(dz * DF + qy) * DF + qx] = s;
      }
      __syncthreads();
      // z, and the one E-entry per fine dof that carries the value to the L-vector
      for (int t = threadIdx.x; t < DF3; t += blockDim.x)
      {
         const int qz = t / (DF * DF), c = t - qz * DF * DF;
         double s = 0.0;
         for (int dz = 0; dz < DC; ++dz) { s = fma(sB[qz + DF * dz], a[dz * DF * DF + c], s); }
         const long long k = e * DF3 + t;
         const int g = P.gmap_f[k];
         if (P.slot_f[k] == P.off_f[g]) { yf[g] = (P.ess_f && P.ess_f[g]) ? 0.0 : s; }
      }
   }
}

// y_coarse (slot layout) = P^T x_fine.  RectangularConstrainedOperator::MultTranspose: fine essential entries of x count as
// zero; every fine dof enters once (through its first E-entry); the coarse essential entries are zeroed by k_mg_sum_zero.
// own_f (multi-GPU): 0 on fine dofs another rank owns - they enter on the owner only.
__global__ void k_mg_restrict(const TransferParams P, const double *__restrict__ xf, const unsigned char *__restrict__ own_f,
                              double *__restrict__ yS)
{
   extern __shared__ double sm[];
   const int DC = P.DC, DF = P.DF, DC3 = DC * DC * DC, DF3 = DF * DF * DF;
   double *sB = sm, *a = sB + DF * DC, *b = a + DF3;
   mg_load_B(P, sB);
   for (long long e = blockIdx.x; e < P.NE; e += gridDim.x)
   {
      __syncthreads();
      for (int t = threadIdx.x; t < DF3; t += blockDim.x)
      {
         const long long k = e * DF3 + t;
         const int g = P.gmap_f[k];
         const bool take = P.slot_f[k] == P.off_f[g] && !(P.ess_f && P.ess_f[g]) && !(own_f && !own_f[g]);
         a[t] = take ? xf[g] : 0.0;
      }
      __syncthreads();
      // x^T: b[qz][qy][dx] = sum_qx B(qx,dx) a[qz][qy][qx]
      for (int t = threadIdx.x; t < DF * DF * DC; t += blockDim.x)
      {
         const int dx = t % DC, r = t / DC;
         double s = 0.0;
         for (int qx = 0; qx < DF; ++qx) { s = fma(sB[qx + DF * dx], a[r * DF + qx], s); }
         b[r * DC + dx] = s;
      }
      __syncthreads();
      // y^T: a[qz][dy][dx] = sum_qy B(qy,dy) b[qz][qy][dx]
      for (int t = threadIdx.x; t < DF * DC * DC; t += blockDim.x)
      {
         const int dx = t % DC, dy = (t / DC) % DC, qz = t / (DC * DC);
         double s = 0.0;
         for (int qy = 0; qy < DF; ++qy) { s = fma(sB[qy + DF * dy], b[(qz * DF + qy) * DC + dx], s); }
         a[(qz * DC + dy) * DC + dx] = s;
      }
      __syncthreads();
      // z^T -> the coarse slot layout (the segmented reduction that follows sums in ascending element order)
      for (int t = threadIdx.x; t < DC3; t += blockDim.x)
      {
         const int dz = t / (DC * DC), c = t - dz * DC * DC;
         double s = 0.0;
         for (int qz = 0; qz < DF; ++qz) { s = fma(sB[qz + DF * dz], a[qz * DC * DC + c], s); }
         yS[P.slot_c[e * DC3 + t]] = s;
      }
   }
}

// segmented E -> L sum of the coarse slot layout; essential coarse entries := 0
__global__ void k_mg_sum_zero(int ndofs, const int *__restrict__ offsets, const double *__restrict__ yS, double *__restrict__ y,
                              const unsigned char *__restrict__ ess)
{
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ndofs; i += gridDim.x * blockDim.x)
   {
      double v = 0.0;
      const int j1 = offsets[i + 1];
      for (int j = offsets[i]; j < j1; ++j) { v += yS[j]; }
      y[i] = (ess && ess[i]) ? 0.0 : v;
   }
}

// (a, b) over the dofs this rank owns (mask == NULL: all)
__global__ void k_dot_masked(int n, const double *__restrict__ a, const double *__restrict__ b, const unsigned char *__restrict__ own_mask,
                             double *partials, unsigned int *ticket, double *out)
{
   double acc = 0.0;
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
   {
      if (!own_mask || own_mask[i]) { acc = fma(a[i], b[i], acc); }
   }
   grid_sum(acc, partials, ticket, out);
}

// (r, z) over the owned dofs + the PCG scalar step that consumes it (1: after the initial residual, 2: in the loop), for
// preconditioners that are not fused into a vector pass of the loop (the multigrid cycle)
__global__ void k_dot_step(int n, const double *__restrict__ r, const double *__restrict__ z, const unsigned char *__restrict__ own_mask,
                           double *partials, unsigned int *ticket, PcgState *st, double *norms_epilogue, int scalar_step)
{
   if (st->done) { return; }
   double acc = 0.0;
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
   {
      if (!own_mask || own_mask[i]) { acc = fma(r[i], z[i], acc); }
   }
   if (grid_sum(acc, partials, ticket, &st->dot_a) && norms_epilogue)
   {
      if (scalar_step == 1) { pcg_scalar_init(st, norms_epilogue); }
      else { pcg_scalar_beta(st, norms_epilogue); }
   }
}

} // namespace b200pa
