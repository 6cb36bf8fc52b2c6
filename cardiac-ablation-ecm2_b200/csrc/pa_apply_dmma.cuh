// Order-6 variant of the hot kernel with the row phases on the FP64 tensor cores (DMMA, mma.sync.m8n8k4.f64).
//
// North star: "DMMA tried for the small contractions but kept only if ncu shows a win".  The probe
// (tools/probes/dmma_probe.cu, profiles/r2i_dmma_probe.json) measured on B200: DMMA.884 peak 37.1 TFLOP/s = the DFMA peak
// (33.9 measured), so there is no flop advantage - but the row contraction of this kernel at p = 6 runs 1.53x faster as
// DMMA with operands from shared memory: one DMMA replaces 8 warp-wide DFMAs AND their constant-bank operand fetches (sm_100
// ptxas never folds a constant into DFMA: every B/G entry costs an LDCU), the basis lives in 4 registers per lane, and a
// 254-register kernel becomes a ~130-register one.  At p = 6 the shapes fit: Q = 8 is exactly the M / N of the tile, D = 7
// pads to 8 (DMMA efficiency (7/8)^2); at p <= 5 the padding costs more than the issue slots saved.
//
// Per element (one element per batch, two warps, one slab = fixed dz per warp at a time):
//   phase A   X_dz [dy][dx] --(B | G along y)--> TB, TG [qy][dx] --(B | G along x)--> F0, F1, F2 [qy][qx]    10 DMMA / slab
//             the accumulator layout of the first product IS the A-operand layout of the second once the k-slots are
//             read as dx = 2 (lane % 4) + s: no shuffle, no shared-memory round trip between the two contractions
//   phase B   unchanged: one (qx,qy) column per thread, z-contraction, q-point operator, z^T (scalar DFMA)
//   phase C   R_f [qy][qx] --(x^T)--> S02^T, S1^T [dx][qy] --(y^T)--> OUT^T [dx][dy] -> y_S[slot]                 10 DMMA / slab
//             C1 and C2 fused in registers (round 1: through shared memory with a barrier in between)
// Shared-memory strides are chosen for the fragment accesses: rows of 8 doubles for sE (LDS.128 / STS.128 of the pairs
// (2 t, 2 t + 1), conflict-free per quarter warp), row stride 10 for the gathered x.
#pragma once
#include <cuda_runtime.h>

#include "pa_apply_kernel.cuh"

namespace b200pa
{

struct DmmaCfg
{
   static constexpr int D = 7, Q = 8, D2 = 49, D3 = 343, Q2 = 64, Q3 = 512;
   static constexpr int NT = 64;
   static constexpr int DP = 10;               // row stride of a gathered slab [dy][dx]
   static constexpr int SXS = D * DP;          // slab stride of sX
   static constexpr int SQ = Q2;               // (field, slab) stride of sE: rows of 8
   static constexpr int SX_DOUBLES = D * SXS;  // one x buffer
   static constexpr int SE_DOUBLES = 3 * D * SQ;
   static constexpr int IDX_OFF = (2 * SX_DOUBLES + SE_DOUBLES) * 8;
   static constexpr int QD_OFF = (IDX_OFF + 4 * D3 * 4 + 15) & ~15;
   static constexpr int SQD_DOUBLES = 6 * Q3 + 2, SQM_DOUBLES = Q3 + 2;
   static constexpr size_t SMEM_BYTES = QD_OFF + sizeof(double) * (SQD_DOUBLES + SQM_DOUBLES);
};

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
   asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <bool DIFF, bool MASS>
__global__ void __launch_bounds__(DmmaCfg::NT, 4)
pa_apply_dmma_kernel(const __grid_constant__ ElemParams<7, 8> P)
{
   using C = DmmaCfg;
   constexpr int D = C::D, Q = C::Q, D2 = C::D2, D3 = C::D3, Q2 = C::Q2, Q3 = C::Q3, NT = C::NT, DP = C::DP, SXS = C::SXS, SQ = C::SQ;
   constexpr int NIO = (D3 + NT - 1) / NT;
#define Bm(q, d) P.bg.B[(q) + Q * (d)]
#define Gm(q, d) P.bg.G[(q) + Q * (d)]
   extern __shared__ __align__(16) unsigned char smem_raw[];
   double *sX = reinterpret_cast<double *>(smem_raw);            // sX[2][D][SXS]
   double *sE = sX + 2 * C::SX_DOUBLES;                          // sE[3][D][SQ]
   int *sGi = reinterpret_cast<int *>(smem_raw + C::IDX_OFF);    // sGi[2][D3]: gather indices, two batches deep
   int *sSl = sGi + 2 * D3;                                      // sSl[2][D3]: slots, one batch ahead
   double *sQd = reinterpret_cast<double *>(smem_raw + C::QD_OFF);
   double *sQm = sQd + C::SQD_DOUBLES;
   __shared__ unsigned long long qbar;
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   if (P.done && *P.done) { return; }
   const int nbatch = P.NE;
   const unsigned long long pol = 0ull;
   const long long lim = (long long)P.NE * D3;

   auto copy_idx = [&](int *dst, const int *src, int b)
   {
      const long long base = (long long)b * D3;
      B200PA_UNROLL
      for (int r = 0; r < NIO; ++r)
      {
         const int t = tid + r * NT;
         if (t < D3)
         {
            if (base + t < lim) { cp_async4(dst + t, src + base + t); }
            else { dst[t] = -1; }
         }
      }
   };
   auto gather_x = [&](double *dst, const int *idx)
   {
      B200PA_UNROLL
      for (int r = 0; r < NIO; ++r)
      {
         const int t = tid + r * NT;
         if (t < D3)
         {
            const int slab = t / D2, k = t - slab * D2, dy = k / D, dx = k - dy * D;
            const int g = idx[t];
            cp_async8_zfill(dst + slab * SXS + dy * DP + dx, P.x + (g >= 0 ? g : 0), g >= 0);
         }
      }
   };
   auto tma_issue = [&](int b)
   {
      unsigned bytes = 0;
      if (DIFF) { bytes += (unsigned)(6 * Q3 * sizeof(double)); }
      if (MASS) { bytes += (unsigned)(Q3 * sizeof(double)); }
      mbar_expect_tx(&qbar, bytes);
      if (DIFF) { tma_bulk_g2s(sQd, P.pa_diff + (long long)b * 6 * Q3, (unsigned)(6 * Q3 * sizeof(double)), &qbar, pol); }
      if (MASS) { tma_bulk_g2s(sQm, P.pa_mass + (long long)b * Q3, (unsigned)(Q3 * sizeof(double)), &qbar, pol); }
   };

   // the basis as DMMA fragments, four registers per lane for the whole kernel:
   //   fa*[s] = M(q = lane / 4, d = 2 (lane % 4) + s)   A-operand of "M times data" (phase A, y) and B-operand of "data times M^T" (phase A, x)
   //   fc*[s] = M(q = 2 (lane % 4) + s, d = lane / 4)   A-operand of "M^T times data" (phase C, x^T) and B-operand of "data times M" (phase C, y^T)
   const int l4 = lane >> 2, t2 = 2 * (lane & 3);
   double faB[2], faG[2], fcB[2], fcG[2];
   B200PA_UNROLL
   for (int s = 0; s < 2; ++s)
   {
      const int d = t2 + s;
      faB[s] = d < D ? P.bg.B[l4 + Q * d] : 0.0;
      faG[s] = d < D ? P.bg.G[l4 + Q * d] : 0.0;
      fcB[s] = l4 < D ? P.bg.B[(t2 + s) + Q * l4] : 0.0;
      fcG[s] = l4 < D ? P.bg.G[(t2 + s) + Q * l4] : 0.0;
   }
   // the pad column / rows of the x buffers (dx = 7 is read by the fragments of lanes 28..31) must hold zeros, once
   for (int i = tid; i < 2 * C::SX_DOUBLES; i += NT) { sX[i] = 0.0; }
   unsigned qphase = 0;
   if (tid == 0) { mbar_init(&qbar, 1); }
   __syncthreads();

   int batch = blockIdx.x;
   if (batch < nbatch)
   {
      if (tid == 0) { tma_issue(batch); }
      copy_idx(sSl, P.slot, batch);
      copy_idx(sGi, P.gmap, batch);
      if (batch + (int)gridDim.x < nbatch) { copy_idx(sGi + D3, P.gmap, batch + gridDim.x); }
      cp_async_commit();
      cp_async_wait_all();
      gather_x(sX, sGi);
      cp_async_commit();
   }
   int cur = 0;
   for (; batch < nbatch; batch += gridDim.x, cur ^= 1)
   {
      const int next = batch + gridDim.x;
      const double *sXin = sX + cur * C::SX_DOUBLES;
      cp_async_wait_all();
      __syncthreads();                     // x of this batch has landed; everybody is done with the previous batch
      if (next < nbatch) { gather_x(sX + (cur ^ 1) * C::SX_DOUBLES, sGi + (cur ^ 1) * D3); }
      if (next + (int)gridDim.x < nbatch) { copy_idx(sGi + cur * D3, P.gmap, next + gridDim.x); }
      if (next < nbatch) { copy_idx(sSl + (cur ^ 1) * D3, P.slot, next); }
      cp_async_commit();

      // ------------------------------------------------------------- phase A: one slab per warp at a time
      for (int dz = warp; dz < D; dz += NT / 32)
      {
         // B-operand of the y-contraction: X[dy = t2 + s][dx = l4]  (the pad column dx = 7 and row dy = 7 read zeros / are skipped)
         const double *xs = sXin + dz * SXS + l4;
         const double x0 = xs[t2 * DP];
         const double x1 = (t2 + 1 < D) ? xs[(t2 + 1) * DP] : 0.0;
         double tb0 = 0.0, tb1 = 0.0, tg0 = 0.0, tg1 = 0.0;
         dmma884(tb0, tb1, faB[0], x0); dmma884(tb0, tb1, faB[1], x1);     // TB[qy = l4][dx = t2 + j]
         if (DIFF) { dmma884(tg0, tg1, faG[0], x0); dmma884(tg0, tg1, faG[1], x1); }
         // x-contraction: the accumulators are the A-operand (k-slot (s, t) <-> dx = 2 t + s), the basis fragment the B-operand
         double f00 = 0.0, f01 = 0.0, f10 = 0.0, f11 = 0.0, f20 = 0.0, f21 = 0.0;
         dmma884(f20, f21, tb0, faB[0]); dmma884(f20, f21, tb1, faB[1]);   // Bx By
         if (DIFF)
         {
            dmma884(f00, f01, tb0, faG[0]); dmma884(f00, f01, tb1, faG[1]); // Gx By
            dmma884(f10, f11, tg0, faB[0]); dmma884(f10, f11, tg1, faB[1]); // Bx Gy
         }
         double *o = sE + dz * SQ + Q * l4 + t2;                            // F[qy = l4][qx = t2 + j]
         if (DIFF)
         {
            *reinterpret_cast<double2 *>(o + 0 * D * SQ) = make_double2(f00, f01);
            *reinterpret_cast<double2 *>(o + 1 * D * SQ) = make_double2(f10, f11);
         }
         *reinterpret_cast<double2 *>(o + 2 * D * SQ) = make_double2(f20, f21);
      }
      __syncthreads();

      // ------------------------------------------------------------- phase B: column, q-point operator, column^T
      mbar_wait(&qbar, qphase);
      qphase ^= 1u;
      {
         double *s = sE + tid;              // column c = tid = qy * 8 + qx
         const double *qd = sQd + tid, *qm = sQm + tid;
         double f0[D], f1[D], f2[D], p0[D], p1[D], p2[D];
         B200PA_UNROLL
         for (int dz = 0; dz < D; ++dz)
         {
            if (DIFF) { f0[dz] = s[(0 * D + dz) * SQ]; f1[dz] = s[(1 * D + dz) * SQ]; }
            f2[dz] = s[(2 * D + dz) * SQ];
            p0[dz] = 0.0; p1[dz] = 0.0; p2[dz] = 0.0;
         }
         B200PA_UNROLL
         for (int qz = 0; qz < Q; ++qz)
         {
            double gX = 0.0, gY = 0.0, gZ = 0.0, val = 0.0;
            B200PA_UNROLL
            for (int dz = 0; dz < D; ++dz)
            {
               if (DIFF)
               {
                  gX = fma(Bm(qz, dz), f0[dz], gX);
                  gY = fma(Bm(qz, dz), f1[dz], gY);
                  gZ = fma(Gm(qz, dz), f2[dz], gZ);
               }
               if (MASS) { val = fma(Bm(qz, dz), f2[dz], val); }
            }
            double hX = 0.0, hY = 0.0, hZ = 0.0, hM = 0.0;
            if (DIFF)
            {
               const double *d = qd + qz * Q2;
               const double o0 = d[0], o1 = d[Q3], o2 = d[2 * Q3], o3 = d[3 * Q3], o4 = d[4 * Q3], o5 = d[5 * Q3];
               hX = o0 * gX + o1 * gY + o2 * gZ;
               hY = o1 * gX + o3 * gY + o4 * gZ;
               hZ = o2 * gX + o4 * gY + o5 * gZ;
            }
            if (MASS) { hM = qm[qz * Q2] * val; }
            B200PA_UNROLL
            for (int dz = 0; dz < D; ++dz)
            {
               if (DIFF)
               {
                  p0[dz] = fma(Bm(qz, dz), hX, p0[dz]);
                  p1[dz] = fma(Bm(qz, dz), hY, p1[dz]);
                  p2[dz] = fma(Gm(qz, dz), hZ, p2[dz]);
               }
               if (MASS) { p2[dz] = fma(Bm(qz, dz), hM, p2[dz]); }
            }
         }
         B200PA_UNROLL
         for (int dz = 0; dz < D; ++dz)
         {
            if (DIFF) { s[(0 * D + dz) * SQ] = p0[dz]; s[(1 * D + dz) * SQ] = p1[dz]; }
            s[(2 * D + dz) * SQ] = p2[dz];
         }
      }
      __syncthreads();
      // every thread is done reading the staged q-data: refill it with the next batch's
      if (next < nbatch && tid == 0) { tma_issue(next); }

      // ------------------------------------------------------------- phase C: x^T then y^T, fused per slab, to y_S[slot]
      for (int dz = warp; dz < D; dz += NT / 32)
      {
         const double *in = sE + dz * SQ + Q * l4 + t2;                    // R_f[qy = l4][qx = t2 + s]: the B-operand pairs
         const double2 r2 = *reinterpret_cast<const double2 *>(in + 2 * D * SQ);
         double s020 = 0.0, s021 = 0.0, s10 = 0.0, s11 = 0.0;
         dmma884(s020, s021, fcB[0], r2.x); dmma884(s020, s021, fcB[1], r2.y);
         if (DIFF)
         {
            const double2 r0 = *reinterpret_cast<const double2 *>(in + 0 * D * SQ);
            const double2 r1 = *reinterpret_cast<const double2 *>(in + 1 * D * SQ);
            dmma884(s020, s021, fcG[0], r0.x); dmma884(s020, s021, fcG[1], r0.y);  // S02^T[dx = l4][qy = t2 + j]
            dmma884(s10, s11, fcB[0], r1.x); dmma884(s10, s11, fcB[1], r1.y);      // S1^T
         }
         double o0 = 0.0, o1 = 0.0;
         dmma884(o0, o1, s020, fcB[0]); dmma884(o0, o1, s021, fcB[1]);             // OUT^T[dx = l4][dy = t2 + j]
         if (DIFF) { dmma884(o0, o1, s10, fcG[0]); dmma884(o0, o1, s11, fcG[1]); }
         if (l4 < D)
         {
            const int *sl = sSl + cur * D3 + dz * D2 + l4;
            const int k0 = sl[t2 * D];
            if (k0 >= 0) { P.y[k0] = o0; }
            if (t2 + 1 < D)
            {
               const int k1 = sl[(t2 + 1) * D];
               if (k1 >= 0) { P.y[k1] = o1; }
            }
         }
      }
   }
#undef Bm
#undef Gm
}

} // namespace b200pa
