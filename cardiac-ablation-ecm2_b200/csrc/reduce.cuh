// Deterministic grid reduction shared by every kernel that produces a scalar.
#pragma once
#include <cuda_runtime.h>

namespace b200pa
{

// ------------------------------------------------------------------ reductions
// Deterministic: fixed shuffle tree per block, block partials summed in a fixed order by the
// last block to finish (ticket counter).  ≙ general/reducers.hpp:451-592 without the host join.
__device__ __forceinline__ double block_sum(double v)
{
   __shared__ double ws[32];
   const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
   for (int o = 16; o > 0; o >>= 1) { v += __shfl_down_sync(0xffffffffu, v, o); }
   if (lane == 0) { ws[w] = v; }
   __syncthreads();
   if (w == 0)
   {
      v = lane < nw ? ws[lane] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { v += __shfl_down_sync(0xffffffffu, v, o); }
   }
   return v; // valid in thread 0
}

// returns true in thread 0 of the last block to finish, after *out has been written: the place to run a
// scalar epilogue without another launch
__device__ __forceinline__ bool grid_sum(double v, double *partials, unsigned int *ticket, double *out)
{
   const double bs = block_sum(v);
   __shared__ bool last;
   if (threadIdx.x == 0)
   {
      partials[blockIdx.x] = bs;
      __threadfence();
      const unsigned int t = atomicInc(ticket, gridDim.x - 1); // wraps back to 0: self-resetting
      last = (t == gridDim.x - 1);
   }
   __syncthreads();
   if (last)
   {
      __threadfence();
      double s = 0.0;
      for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) { s += ((volatile double *)partials)[i]; }
      s = block_sum(s);
      if (threadIdx.x == 0) { *out = s; return true; }
   }
   return false;
}

} // namespace b200pa
