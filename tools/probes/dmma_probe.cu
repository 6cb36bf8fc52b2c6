// FP64 pipe probe for the DMMA experiment (north star: "DMMA tried for the small contractions, kept only if ncu shows a
// win"): sustained DFMA and DMMA.884 (mma.sync.m8n8k4.f64) rates on one B200, and the rate of the small contraction the
// apply kernel's row phases do (8 rows x Q inputs -> D outputs, basis in registers) written both ways with operands
// coming from shared memory as in the kernel.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_probe dmma_probe.cu && ./dmma_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
   asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int ILP>
__global__ void k_dfma(double *out, int iters, double a, double b)
{
   double acc[ILP];
#pragma unroll
   for (int i = 0; i < ILP; ++i) { acc[i] = threadIdx.x + i; }
   for (int it = 0; it < iters; ++it)
   {
#pragma unroll
      for (int i = 0; i < ILP; ++i) { acc[i] = fma(acc[i], a, b); }
   }
   double s = 0;
#pragma unroll
   for (int i = 0; i < ILP; ++i) { s += acc[i]; }
   out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void k_dmma(double *out, int iters, double a, double b)
{
   double c0[ILP], c1[ILP];
#pragma unroll
   for (int i = 0; i < ILP; ++i) { c0[i] = threadIdx.x + i; c1[i] = i; }
   for (int it = 0; it < iters; ++it)
   {
#pragma unroll
      for (int i = 0; i < ILP; ++i) { dmma884(c0[i], c1[i], a, b); }
   }
   double s = 0;
#pragma unroll
   for (int i = 0; i < ILP; ++i) { s += c0[i] + c1[i]; }
   out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// the row contraction of phase C1 at p = 6: rows (8 inputs, 3 fields) -> 7 outputs x 2, data in shared memory (row stride RS)
constexpr int Q = 8, D = 7, RS = 12, ROWS = 64;       // 64 rows per CTA of 64 threads, as the kernel at p = 6
struct BG { double B[Q * D], G[Q * D]; };
__global__ void k_rows_dfma(const __grid_constant__ BG P, double *out, int iters)
{
   __shared__ double s[3][ROWS * RS];
   for (int i = threadIdx.x; i < 3 * ROWS * RS; i += blockDim.x) { (&s[0][0])[i] = 1e-3 * (i % 17); }
   __syncthreads();
   const int row = threadIdx.x;
   double chk = 0;
   for (int it = 0; it < iters; ++it)
   {
      double r0[Q], r1[Q], r2[Q];
#pragma unroll
      for (int q = 0; q < Q; ++q) { r0[q] = s[0][row * RS + q]; r1[q] = s[1][row * RS + q]; r2[q] = s[2][row * RS + q]; }
#pragma unroll
      for (int d = 0; d < D; ++d)
      {
         double a = 0, b = 0;
#pragma unroll
         for (int q = 0; q < Q; ++q)
         {
            a = fma(P.G[q + Q * d], r0[q], a);
            b = fma(P.B[q + Q * d], r1[q], b);
            a = fma(P.B[q + Q * d], r2[q], a);
         }
         s[0][row * RS + d] = a; s[1][row * RS + d] = b;
      }
      __syncthreads();
      chk += s[0][row * RS];
   }
   out[blockIdx.x * blockDim.x + threadIdx.x] = chk;
}
__global__ void k_rows_dmma(const __grid_constant__ BG P, double *out, int iters)
{
   __shared__ double s[3][ROWS * RS];
   for (int i = threadIdx.x; i < 3 * ROWS * RS; i += blockDim.x) { (&s[0][0])[i] = 1e-3 * (i % 17); }
   __syncthreads();
   const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
   // B fragments (k = lane % 4 + 4 step, n = lane / 4 = output index d): the basis, in registers for the whole kernel
   double bB[2], bG[2];
#pragma unroll
   for (int st = 0; st < 2; ++st)
   {
      const int q = (lane & 3) + 4 * st, d = lane >> 2;
      bB[st] = d < D ? P.B[q + Q * d] : 0.0;
      bG[st] = d < D ? P.G[q + Q * d] : 0.0;
   }
   double chk = 0;
   for (int it = 0; it < iters; ++it)
   {
      for (int tile = warp; tile < ROWS / 8; tile += blockDim.x / 32)
      {
         const int row = tile * 8 + (lane >> 2);
         double a0[2], a1[2], a2[2];
#pragma unroll
         for (int st = 0; st < 2; ++st)
         {
            const int q = (lane & 3) + 4 * st;
            a0[st] = s[0][row * RS + q]; a1[st] = s[1][row * RS + q]; a2[st] = s[2][row * RS + q];
         }
         double c0 = 0, c1 = 0, e0 = 0, e1 = 0;
#pragma unroll
         for (int st = 0; st < 2; ++st)
         {
            dmma884(c0, c1, a0[st], bG[st]);
            dmma884(e0, e1, a1[st], bB[st]);
            dmma884(c0, c1, a2[st], bB[st]);
         }
         const int d = 2 * (lane & 3);
         s[0][row * RS + d] = c0; s[1][row * RS + d] = e0;
         if (d + 1 < D) { s[0][row * RS + d + 1] = c1; s[1][row * RS + d + 1] = e1; }
      }
      __syncthreads();
      chk += s[0][(threadIdx.x % ROWS) * RS];
   }
   out[blockIdx.x * blockDim.x + threadIdx.x] = chk;
}

template <typename F>
float timeit(F f)
{
   cudaEvent_t e0, e1;
   cudaEventCreate(&e0); cudaEventCreate(&e1);
   f();
   cudaDeviceSynchronize();
   cudaEventRecord(e0);
   f();
   cudaEventRecord(e1);
   cudaEventSynchronize(e1);
   float ms = 0;
   cudaEventElapsedTime(&ms, e0, e1);
   return ms;
}

int main()
{
   cudaDeviceProp p;
   cudaGetDeviceProperties(&p, 0);
   const int sms = p.multiProcessorCount;
   double *out;
   cudaMalloc(&out, sizeof(double) * sms * 16 * 1024);
   const int iters = 20000;
   printf("{\"device\": \"%s\", \"sms\": %d", p.name, sms);
   {
      const int blocks = sms * 4, threads = 256;
      float ms = timeit([&] { k_dfma<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
      printf(", \"dfma_tflops\": %.2f", 2.0 * blocks * threads * 8.0 * iters / (ms * 1e-3) / 1e12);
      ms = timeit([&] { k_dmma<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
      printf(", \"dmma884_tflops\": %.2f", 512.0 * blocks * (threads / 32) * 8.0 * iters / (ms * 1e-3) / 1e12);
      ms = timeit([&] { k_dmma<2><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
      printf(", \"dmma884_tflops_ilp2\": %.2f", 512.0 * blocks * (threads / 32) * 2.0 * iters / (ms * 1e-3) / 1e12);
   }
   {
      BG h;
      for (int i = 0; i < Q * D; ++i) { h.B[i] = 0.1 + 0.01 * i; h.G[i] = 0.2 - 0.01 * i; }
      const int blocks = sms * 4, it2 = 4000;
      float m1 = timeit([&] { k_rows_dfma<<<blocks, ROWS>>>(h, out, it2); });
      float m2 = timeit([&] { k_rows_dmma<<<blocks, ROWS>>>(h, out, it2); });
      const double rows = (double)blocks * ROWS * it2;
      printf(", \"row_phase_p6\": {\"dfma_Grows_per_s\": %.2f, \"dmma_Grows_per_s\": %.2f, \"dmma_over_dfma\": %.3f}", rows / (m1 * 1e-3) / 1e9,
             rows / (m2 * 1e-3) / 1e9, m1 / m2);
   }
   printf("}\n");
   return 0;
}
