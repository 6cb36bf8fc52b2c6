// Trilinear hexahedron geometry shared by the set-up and q-point kernels.
#pragma once
#include <cuda_runtime.h>

namespace b200pa
{

// mesh/mesh.cpp:15220-15273 for trilinear hexes: J(q) = sum_v X_v (x) grad N_v(xi_q); vertex
// order of the reference hexahedron (mesh/mesh.cpp:3757-3765).  Jm[row + 3*col].
__device__ __forceinline__ void trilinear_jacobian(const double *__restrict__ vtx, const int *__restrict__ ev8, double x, double y,
                                                   double z, double Jm[9])
{
   const double bx[2] = {1.0 - x, x}, by[2] = {1.0 - y, y}, bz[2] = {1.0 - z, z};
   const double gm[2] = {-1.0, 1.0};
   // local vertex v -> (i,j,k) corner bits
   const int ci[8] = {0, 1, 1, 0, 0, 1, 1, 0}, cj[8] = {0, 0, 1, 1, 0, 0, 1, 1}, ck[8] = {0, 0, 0, 0, 1, 1, 1, 1};
#pragma unroll
   for (int k = 0; k < 9; ++k) { Jm[k] = 0.0; }
#pragma unroll
   for (int v = 0; v < 8; ++v)
   {
      const double *X = vtx + 3LL * ev8[v];
      const double d0 = gm[ci[v]] * by[cj[v]] * bz[ck[v]];
      const double d1 = bx[ci[v]] * gm[cj[v]] * bz[ck[v]];
      const double d2 = bx[ci[v]] * by[cj[v]] * gm[ck[v]];
#pragma unroll
      for (int r = 0; r < 3; ++r)
      {
         Jm[r + 0] = fma(X[r], d0, Jm[r + 0]);
         Jm[r + 3] = fma(X[r], d1, Jm[r + 3]);
         Jm[r + 6] = fma(X[r], d2, Jm[r + 6]);
      }
   }
}

} // namespace b200pa
