// Order-6 variant of the hot kernel with the row phases on the FP64 tensor cores (DMMA, mma.sync.m8n8k4.f64).
//
// North star: "DMMA tried for the small contractions but kept only if ncu shows a win".  The probe
// (tools/probes/dmma_probe.cu, profiles/r2i_dmma_probe.json) measured on B200: DMMA.884 peak 37.1 TFLOP/s = the DFMA peak
// (33.9 measured), so there is no flop advantage - but the row contraction of this kernel at p = 6 runs 1.53x faster as
// DMMA with operands from shared memory: one DMMA replaces 8 warp-wide DFMAs AND their constant-bank operand fetches (sm_100
// ptxas never folds a constant into DFMA: every B/G entry costs an LDCU), the basis lives in 4 registers per lane, and a
// 254-register kernel becomes a ~60-register one.  At p = 6 the shapes fit: Q = 8 is exactly the M / N of the tile, D = 7
// pads to 8 (DMMA efficiency (7/8)^2); at p <= 5 the padding costs more than the issue slots saved.
//
// Per element (one element per batch, four warps, one slab = fixed dz per warp at a time):
//   phase A   X_dz [dy][dx] --(B | G along y)--> TB, TG [qy][dx] --(B | G along x)--> F0, F1, F2 [qy][qx]    10 DMMA / slab
//             the accumulator layout of the first product IS the A-operand layout of the second once the k-slots are
//             read as dx = 2 (lane % 4) + s: no shuffle, no shared-memory round trip between the two contractions
//   phase B   rows of 8 columns: G^T = F^T B^T, q-point operator on the accumulators, P^T = H^T B                        16 DMMA / row
//             (the q-point operator is the only scalar FP64 work left: 15 DFMA/DMUL per q-point)
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
   static constexpr int NT = 128;              // four warps: slabs (phases A, C) and rows of 8 columns (phase B) are dealt out to them
   static constexpr int DP = 10;               // row stride of a gathered slab [dy][dx]
   static constexpr int SXS = D * DP;          // slab stride of sX
   static constexpr int SQ = Q2 + 2;           // (field, slab) stride of sE: rows of 8; 66 = 2 (mod 8), so that the phase-B fragments
                                               // (lanes walk 4 columns x 4 slabs 2 t + s) touch 16 different banks
   static constexpr int QP2 = 2 * Q2 + 4;      // stride of a staged PAIR of q-data planes (qz = 2 t, 2 t + 1): the lanes of a phase-B
                                               // fragment walk t = 0..3, 132 = 4 (mod 16) puts the four pairs into different banks
   static constexpr int QCS = 4 * QP2;         // stride of a q-data component (8 planes)
   static constexpr int SX_DOUBLES = D * SXS;  // one x buffer
   static constexpr int SE_DOUBLES = 3 * D * SQ;
   static constexpr int IDX_OFF = (2 * SX_DOUBLES + SE_DOUBLES) * 8;
   static constexpr int QD_OFF = (IDX_OFF + 4 * D3 * 4 + 15) & ~15;
   static constexpr int SQD_DOUBLES = 6 * QCS, SQM_DOUBLES = QCS;          // one bulk copy per pair of planes, padded
   static constexpr size_t SMEM_BYTES = QD_OFF + sizeof(double) * (SQD_DOUBLES + SQM_DOUBLES);
};

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
   // not volatile: the compiler may interleave independent accumulator chains (a dependent DMMA waits for the full pipe latency)
   asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <bool DIFF, bool MASS>
__global__ void __launch_bounds__(DmmaCfg::NT, 4)   // 4 CTAs per SM (shared memory): 128 registers per thread
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
   // q-data of one element: 7 components x 4 pairs of planes (2 x 64 doubles), one bulk copy each into a padded slot.
   // The issue is serialised per lane (UBLKCP takes uniform operands): warp 3, which has one slab instead of two in phases
   // A and C, issues 16 of the 28 copies, the other warps 4 each
   auto tma_issue = [&](int b)
   {
      const int i = warp == 3 ? 12 + lane : 4 * warp + lane;
      if (lane < (warp == 3 ? 16 : 4))
      {
         const int k = i >> 2, pr = i & 3;
         if (k < 6) { if (DIFF) { tma_bulk_g2s(sQd + k * C::QCS + pr * C::QP2, P.pa_diff + (long long)b * 6 * Q3 + k * Q3 + pr * 2 * Q2, (unsigned)(2 * Q2 * sizeof(double)), &qbar, pol); } }
         else if (MASS) { tma_bulk_g2s(sQm + pr * C::QP2, P.pa_mass + (long long)b * Q3 + pr * 2 * Q2, (unsigned)(2 * Q2 * sizeof(double)), &qbar, pol); }
      }
   };
   auto tma_expect = [&]()
   {
      unsigned bytes = 0;
      if (DIFF) { bytes += (unsigned)(6 * Q3 * sizeof(double)); }
      if (MASS) { bytes += (unsigned)(Q3 * sizeof(double)); }
      mbar_expect_tx(&qbar, bytes);
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
      if (tid == 0) { tma_expect(); }
      tma_issue(batch);
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

      // ------------------------------------------------------------- phase A: the warp's slabs dz = warp, warp + 4 together
      // (independent accumulator chains issued back to back: k-step-major order, two slabs interleaved)
      {
         constexpr int NS = 2;
         const bool on1 = warp + 4 < D;     // warp 3 has one slab only
         double x0[NS], x1[NS];
         B200PA_UNROLL
         for (int i = 0; i < NS; ++i)
         {
            const int dz = warp + 4 * i;
            // B-operand of the y-contraction: X[dy = t2 + s][dx = l4]  (the pad column dx = 7 reads zeros, the row dy = 7 is skipped)
            const double *xs = sXin + (i == 0 || on1 ? dz : warp) * SXS + l4;
            x0[i] = xs[t2 * DP];
            x1[i] = (t2 + 1 < D) ? xs[(t2 + 1) * DP] : 0.0;
         }
         double tb[NS][2] = {{0.0, 0.0}, {0.0, 0.0}}, tg[NS][2] = {{0.0, 0.0}, {0.0, 0.0}};
         B200PA_UNROLL
         for (int i = 0; i < NS; ++i) { dmma884(tb[i][0], tb[i][1], faB[0], x0[i]); if (DIFF) { dmma884(tg[i][0], tg[i][1], faG[0], x0[i]); } }
         B200PA_UNROLL
         for (int i = 0; i < NS; ++i) { dmma884(tb[i][0], tb[i][1], faB[1], x1[i]); if (DIFF) { dmma884(tg[i][0], tg[i][1], faG[1], x1[i]); } }
         // x-contraction: the accumulators are the A-operand (k-slot (s, t) <-> dx = 2 t + s), the basis fragment the B-operand
         double f0[NS][2] = {{0.0, 0.0}, {0.0, 0.0}}, f1[NS][2] = {{0.0, 0.0}, {0.0, 0.0}}, f2[NS][2] = {{0.0, 0.0}, {0.0, 0.0}};
         B200PA_UNROLL
         for (int k = 0; k < 2; ++k)
         {
            B200PA_UNROLL
            for (int i = 0; i < NS; ++i)
            {
               dmma884(f2[i][0], f2[i][1], tb[i][k], faB[k]);                         // Bx By
               if (DIFF)
               {
                  dmma884(f0[i][0], f0[i][1], tb[i][k], faG[k]);                      // Gx By
                  dmma884(f1[i][0], f1[i][1], tg[i][k], faB[k]);                      // Bx Gy
               }
            }
         }
         B200PA_UNROLL
         for (int i = 0; i < NS; ++i)
         {
            if (i == 1 && !on1) { break; }
            double *o = sE + (warp + 4 * i) * SQ + Q * l4 + t2;                      // F[qy = l4][qx = t2 + j]
            if (DIFF)
            {
               *reinterpret_cast<double2 *>(o + 0 * D * SQ) = make_double2(f0[i][0], f0[i][1]);
               *reinterpret_cast<double2 *>(o + 1 * D * SQ) = make_double2(f1[i][0], f1[i][1]);
            }
            *reinterpret_cast<double2 *>(o + 2 * D * SQ) = make_double2(f2[i][0], f2[i][1]);
         }
      }
      __syncthreads();

      // ------------------------------------------------------------- phase B: z, q-point operator, z^T on rows of 8 columns
      // tile = the 8 columns (qx) of one qy; transposed products so that the columns sit on the M axis in every step:
      //   G^T[col][qz] = F^T[col][dz] B^T[dz][qz]          A-operand from sE (k-slot <-> dz = 2 t + s), B-operand = fa*
      //   q-point operator on the accumulator entries (col = lane / 4, qz = 2 (lane % 4) + j)
      //   P^T[col][dz] = H^T[col][qz] B[qz][dz]            A-operand = the accumulators, B-operand = fc*
      mbar_wait(&qbar, qphase);
      qphase ^= 1u;
      {
         constexpr int NS = 2;              // the warp's rows qy = warp, warp + 4, interleaved
         const bool ok1 = t2 + 1 < D;
         double a0[NS][2], a1[NS][2], a2[NS][2];
         B200PA_UNROLL
         for (int i = 0; i < NS; ++i)
         {
            const double *s = sE + Q * (warp + 4 * i) + l4;
            a0[i][0] = a0[i][1] = a1[i][0] = a1[i][1] = 0.0;
            if (DIFF)
            {
               a0[i][0] = s[(0 * D + t2) * SQ]; a1[i][0] = s[(1 * D + t2) * SQ];
               if (ok1) { a0[i][1] = s[(0 * D + t2 + 1) * SQ]; a1[i][1] = s[(1 * D + t2 + 1) * SQ]; }
            }
            a2[i][0] = s[(2 * D + t2) * SQ];
            a2[i][1] = ok1 ? s[(2 * D + t2 + 1) * SQ] : 0.0;
         }
         double gX[NS][2] = {{0.0, 0.0}, {0.0, 0.0}}, gY[NS][2] = {{0.0, 0.0}, {0.0, 0.0}}, gZ[NS][2] = {{0.0, 0.0}, {0.0, 0.0}}, vl[NS][2] = {{0.0, 0.0}, {0.0, 0.0}};
         B200PA_UNROLL
         for (int k = 0; k < 2; ++k)
         {
            B200PA_UNROLL
            for (int i = 0; i < NS; ++i)
            {
               if (DIFF)
               {
                  dmma884(gX[i][0], gX[i][1], a0[i][k], faB[k]);
                  dmma884(gY[i][0], gY[i][1], a1[i][k], faB[k]);
                  dmma884(gZ[i][0], gZ[i][1], a2[i][k], faG[k]);
               }
               if (MASS) { dmma884(vl[i][0], vl[i][1], a2[i][k], faB[k]); }
            }
         }
         double hX[NS][2], hY[NS][2], hZ[NS][2], hM[NS][2];
         B200PA_UNROLL
         for (int i = 0; i < NS; ++i)
         {
            B200PA_UNROLL
            for (int j = 0; j < 2; ++j)
            {
               const int qz = t2 + j, col = Q * (warp + 4 * i) + l4;
               hX[i][j] = hY[i][j] = hZ[i][j] = hM[i][j] = 0.0;
               if (DIFF)
               {
                  const double *d = sQd + (qz >> 1) * C::QP2 + (qz & 1) * Q2 + col;   // = (lane % 4) * QP2 + j * Q2 + col
                  const double o0 = d[0], o1 = d[C::QCS], o2 = d[2 * C::QCS], o3 = d[3 * C::QCS], o4 = d[4 * C::QCS], o5 = d[5 * C::QCS];
                  hX[i][j] = o0 * gX[i][j] + o1 * gY[i][j] + o2 * gZ[i][j];
                  hY[i][j] = o1 * gX[i][j] + o3 * gY[i][j] + o4 * gZ[i][j];
                  hZ[i][j] = o2 * gX[i][j] + o4 * gY[i][j] + o5 * gZ[i][j];
               }
               if (MASS) { hM[i][j] = sQm[(qz >> 1) * C::QP2 + (qz & 1) * Q2 + col] * vl[i][j]; }
            }
         }
         double p0[NS][2] = {{0.0, 0.0}, {0.0, 0.0}}, p1[NS][2] = {{0.0, 0.0}, {0.0, 0.0}}, p2[NS][2] = {{0.0, 0.0}, {0.0, 0.0}};
         B200PA_UNROLL
         for (int k = 0; k < 2; ++k)
         {
            B200PA_UNROLL
            for (int i = 0; i < NS; ++i)
            {
               if (DIFF)
               {
                  dmma884(p0[i][0], p0[i][1], hX[i][k], fcB[k]);
                  dmma884(p1[i][0], p1[i][1], hY[i][k], fcB[k]);
                  dmma884(p2[i][0], p2[i][1], hZ[i][k], fcG[k]);
               }
            }
            if (MASS)
            {
               B200PA_UNROLL
               for (int i = 0; i < NS; ++i) { dmma884(p2[i][0], p2[i][1], hM[i][k], fcB[k]); }
            }
         }
         // P^T[col = l4][dz = t2 + j]
         B200PA_UNROLL
         for (int i = 0; i < NS; ++i)
         {
            double *s = sE + Q * (warp + 4 * i) + l4;
            if (DIFF) { s[(0 * D + t2) * SQ] = p0[i][0]; s[(1 * D + t2) * SQ] = p1[i][0]; }
            s[(2 * D + t2) * SQ] = p2[i][0];
            if (ok1)
            {
               if (DIFF) { s[(0 * D + t2 + 1) * SQ] = p0[i][1]; s[(1 * D + t2 + 1) * SQ] = p1[i][1]; }
               s[(2 * D + t2 + 1) * SQ] = p2[i][1];
            }
         }
      }
      __syncthreads();
      // every thread is done reading the staged q-data: refill it with the next batch's
      if (next < nbatch) { if (tid == 0) { tma_expect(); } tma_issue(next); }

      // ------------------------------------------------------------- phase C: x^T then y^T, fused per slab, to y_S[slot]
      {
         constexpr int NS = 2;
         const bool on1 = warp + 4 < D;
         double2 r0[NS], r1[NS], r2[NS];
         int k0[NS], k1[NS];
         B200PA_UNROLL
         for (int i = 0; i < NS; ++i)
         {
            const int dz = (i == 0 || on1) ? warp + 4 * i : warp;
            const double *in = sE + dz * SQ + Q * l4 + t2;                          // R_f[qy = l4][qx = t2 + s]: the B-operand pairs
            r2[i] = *reinterpret_cast<const double2 *>(in + 2 * D * SQ);
            if (DIFF)
            {
               r0[i] = *reinterpret_cast<const double2 *>(in + 0 * D * SQ);
               r1[i] = *reinterpret_cast<const double2 *>(in + 1 * D * SQ);
            }
            // slots of (dz, dy = t2 + j, dx = l4), fetched now: their latency hides behind the products
            const int *sl = sSl + cur * D3 + dz * D2 + (l4 < D ? l4 : 0);
            k0[i] = sl[t2 * D];
            k1[i] = (t2 + 1 < D) ? sl[(t2 + 1) * D] : -1;
         }
         double s02[NS][2] = {{0.0, 0.0}, {0.0, 0.0}}, s1[NS][2] = {{0.0, 0.0}, {0.0, 0.0}};
         B200PA_UNROLL
         for (int i = 0; i < NS; ++i)
         {
            dmma884(s02[i][0], s02[i][1], fcB[0], r2[i].x);
            if (DIFF) { dmma884(s1[i][0], s1[i][1], fcB[0], r1[i].x); }
         }
         B200PA_UNROLL
         for (int i = 0; i < NS; ++i)
         {
            dmma884(s02[i][0], s02[i][1], fcB[1], r2[i].y);
            if (DIFF) { dmma884(s1[i][0], s1[i][1], fcB[1], r1[i].y); }
         }
         if (DIFF)
         {
            B200PA_UNROLL
            for (int i = 0; i < NS; ++i) { dmma884(s02[i][0], s02[i][1], fcG[0], r0[i].x); }
            B200PA_UNROLL
            for (int i = 0; i < NS; ++i) { dmma884(s02[i][0], s02[i][1], fcG[1], r0[i].y); }      // S02^T[dx = l4][qy = t2 + j]
         }
         double o[NS][2] = {{0.0, 0.0}, {0.0, 0.0}};
         B200PA_UNROLL
         for (int k = 0; k < 2; ++k)
         {
            B200PA_UNROLL
            for (int i = 0; i < NS; ++i) { dmma884(o[i][0], o[i][1], s02[i][k], fcB[k]); }         // OUT^T[dx = l4][dy = t2 + j]
            if (DIFF)
            {
               B200PA_UNROLL
               for (int i = 0; i < NS; ++i) { dmma884(o[i][0], o[i][1], s1[i][k], fcG[k]); }
            }
         }
         if (l4 < D)
         {
            B200PA_UNROLL
            for (int i = 0; i < NS; ++i)
            {
               if (i == 1 && !on1) { break; }
               if (k0[i] >= 0) { P.y[k0[i]] = o[i][0]; }
               if (k1[i] >= 0) { P.y[k1[i]] = o[i][1]; }
            }
         }
      }
   }
#undef Bm
#undef Gm
}

} // namespace b200pa
