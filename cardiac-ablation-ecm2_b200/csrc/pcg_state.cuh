// Device-resident scalar state of the PCG loop and its scalar steps (shared by the vector kernels, whose
// reductions run them as epilogues, and by the peer-memory all-reduce kernel of comm.cu).
#pragma once
#include <cuda_runtime.h>

namespace b200pa
{

// Device-resident scalar state of CGSolver::Mult (linalg/solvers.cpp:869-1050).
struct PcgState
{
   double nom, nom0, den, betanom, r0, alpha, beta;
   double dot_a, dot_b, dot_b2; // raw reduction results (before the scalar step / all-reduce); dot_b2 directly follows
                                // dot_b: the shared-dof part of d.Ad on the multi-GPU peer-memory path
   double rel_tol, abs_tol;
   int iter;                  // the reference's loop variable i
   int max_iter;
   int done, converged, final_iter, nonfinite;
};


// the scalar steps of the loop; run either as the epilogue of the reduction that produced their input
// (single GPU) or as 1-thread kernels after the all-reduce (multi-GPU)
__device__ __forceinline__ void pcg_scalar_init(PcgState *st, double *norms)
{
   const double nom = st->dot_a;
   st->nom = st->nom0 = nom;
   norms[0] = nom;
   st->iter = 1;
   if (!isfinite(nom)) { st->nonfinite = 1; st->done = 1; st->converged = 0; st->final_iter = 0; st->betanom = nom; return; }
   if (nom < 0.0) { st->done = 1; st->converged = 0; st->final_iter = 0; st->betanom = nom; return; }
   st->r0 = fmax(nom * st->rel_tol * st->rel_tol, st->abs_tol * st->abs_tol);
   st->betanom = nom;
   if (nom <= st->r0) { st->done = 1; st->converged = 1; st->final_iter = 0; }
}

__device__ __forceinline__ void pcg_scalar_den(PcgState *st)
{
   if (st->done) { return; }
   const double den = st->dot_b;
   st->den = den;
   if (!isfinite(den)) { st->nonfinite = 1; st->done = 1; st->converged = 0; st->final_iter = st->iter - 1; return; }
   if (den == 0.0)
   {
      // before the loop: final_iter = 0; inside: final_iter = i (already incremented)
      st->done = 1; st->converged = 0; st->final_iter = (st->iter == 1) ? 0 : st->iter;
      return;
   }
   st->alpha = st->nom / den;
}

__device__ __forceinline__ void pcg_scalar_beta(PcgState *st, double *norms)
{
   if (st->done) { return; }
   const double betanom = st->dot_a;
   const int i = st->iter;
   st->betanom = betanom;
   norms[i] = betanom;
   if (!isfinite(betanom)) { st->nonfinite = 1; st->done = 1; st->converged = 0; st->final_iter = i; return; }
   if (betanom < 0.0) { st->done = 1; st->converged = 0; st->final_iter = i; return; }
   if (betanom <= st->r0) { st->done = 1; st->converged = 1; st->final_iter = i; return; }
   if (i + 1 > st->max_iter) { st->done = 1; st->converged = 0; st->final_iter = st->max_iter; return; }
   st->iter = i + 1;
   st->beta = betanom / st->nom;
   st->nom = betanom; // (:1026; alpha of the next pass uses it)
}

} // namespace b200pa
