// The hot kernel: fused L->E gather + sum-factorised diffusion (+ mass) + slot-layout E->L write,
// FP64, sm_100a.  One launch replaces the reference's K1 + K16 + K4 + K7 and the write half of K2
// (fem/restriction.cpp:109-129, fem/bilinearform_ext.cpp:543-557,
//  fem/integ/bilininteg_diffusion_kernels.hpp:989-1214, fem/integ/bilininteg_mass_kernels.hpp:807-1033).
//
// Second design, driven by the first ncu capture (profiles/r1a_*): the slab/column kernel moved the
// right bytes (DRAM traffic = algorithmic bytes) but kept only ~21 KB per SM in flight - HBM latency
// bound at 39 % of peak with 25 % occupancy.  Changes:
//   * small batches (NEB elements, NEB*Q^2 = one q-point column per thread) whose q-data is ONE
//     contiguous range of each q-data array, fetched one batch ahead: thread 0 issues two TMA bulk
//     copies (cp.async.bulk -> UBLKCP, completion on an mbarrier) right after batch b's columns are
//     done, and the copy flies during the whole of phases C/out/in/A of the next batch
//     (3-4 CTAs/SM x 29-36 KB = 100-140 KB per SM in flight, Little's law needs ~45 KB); no register
//     is tied up by the prefetch (the register-resident variant, QPF && !TMA, spilled for p >= 3);
//   * the gather is prefetched the same way (indices one batch ahead, x values half a batch ahead);
//   * fine-grained tasks - (slab,qy) rows instead of whole slabs - so the small batch still fills
//     the CTA in the x/y contractions, and all B/G operands except one row per task stay
//     compile-time kernel-parameter constants (free DFMA operands);
//   * shared-memory strides chosen so that every phase is bank-conflict free at p=2.
//
//   stage-in : x (regs)                              -> sXin[e][dz][dy][dx]
//   phase A  : task (e,dz,qy): y- then x-contraction -> sE[e][f][dz][qy][qx]        f < 3
//   phase B  : task (e,qx,qy): z, q-point op, z^T    -> sE (in place)
//   phase C1 : task (e,dz,qy): x^T                   -> sE[e][f][dz][qy][dx]        f < 2 (in place)
//   phase C2 : task (e,dz,dx): y^T                   -> sXout[e][dz][dy][dx]
//   stage-out: sXout -> y_S[slot]  (slot = position in the E->L CSR: the segmented reduction
//              that follows streams contiguously and stays atomic-free, ascending element order)
#pragma once
#include <cuda_runtime.h>

#include "pa_element_kernel.cuh"
#include "tma.cuh"

namespace b200pa
{

template <int D, int Q>
struct ApplyCfg
{
   static constexpr int D2 = D * D, D3 = D * D * D, Q2 = Q * Q, Q3 = Q * Q * Q;
   // elements per batch: one q-point column per thread
#ifdef B200PA_TUNE_NEB   // tuning builds only (tools/tune.sh)
   static constexpr int NEB = B200PA_TUNE_NEB;
   static constexpr int MINB = B200PA_TUNE_MINB;
   static constexpr bool QPF = B200PA_TUNE_QPF;
   static constexpr bool TMA = B200PA_TUNE_TMA;
#else
   // tuned on B200 with tools/tune.sh + tools/tune_run.sh (profiles/r1c_tuning.txt)
   static constexpr int NEB = (D == 2) ? 28 : (D == 3) ? 8 : (D == 4) ? 5 : (D == 5) ? 3 : 2;
   static constexpr int MINB = (D == 2) ? 3 : (D == 3) ? 4 : (D <= 6) ? 3 : 2; // resident CTAs/SM the register budget allows
   static constexpr bool QPF = true;  // q-data of batch b+1 is fetched while batch b is still being contracted ...
   static constexpr bool TMA = true;  // ... by TMA bulk copies into shared memory (false: into registers, 7Q doubles/thread)
#endif
   static constexpr int NT = ((NEB * Q2 + 31) / 32) * 32;
   static constexpr int NIO = (NEB * D3 + NT - 1) / NT;          // gather / scatter items per thread
   // shared-memory strides found by tools/smem_strides.py (fewest bank-conflict wavefronts over all phases;
   // conflict-free at p=2): SXS = slab stride of sXin/sXout, SQ = slab stride and ES = element stride of sE
   static constexpr int SXS = (D == 2) ? 6 : (D == 3) ? 9 : (D == 4) ? 20 : (D == 5) ? 25 : (D == 6) ? 38 : 55;
   static constexpr int SQ = (D == 2) ? 11 : (D == 3) ? 19 : (D == 4) ? 25 : (D == 5) ? 37 : (D == 6) ? 49 : 71;
   static constexpr int ES = (D == 2) ? 73 : (D == 3) ? 185 : (D == 4) ? 308 : (D == 5) ? 564 : (D == 6) ? 886 : 1505;
   static_assert(SXS >= D2 && SQ >= Q2 && ES >= 3 * D * SQ, "strides too small");
   static constexpr int SX_DOUBLES = NEB * D * SXS;
   static constexpr int SE_DOUBLES = NEB * ES;
   // TMA mode: one batch of q-data staged in shared memory, global layout kept ([e][6][Q^3] and [e][Q^3]);
   // +2 doubles of slack each for the 16-byte alignment of the bulk copies
   static constexpr int SQD_DOUBLES = TMA ? (NEB * 6 * Q3 + 2) : 0;
   static constexpr int SQM_DOUBLES = TMA ? (((NEB * Q3 + 2) + 1) & ~1) : 0;
   static constexpr int WORK_DOUBLES = ((2 * SX_DOUBLES + SE_DOUBLES + 2 * Q * D) + 1) & ~1;
   static constexpr size_t SMEM_BYTES = sizeof(double) * (WORK_DOUBLES + SQD_DOUBLES + SQM_DOUBLES);
};

template <int D, int Q, bool DIFF, bool MASS>
__global__ void __launch_bounds__(ApplyCfg<D, Q>::NT, ApplyCfg<D, Q>::MINB)
pa_apply_kernel(const __grid_constant__ ElemParams<D, Q> P)
{
   using C = ApplyCfg<D, Q>;
   constexpr int D2 = C::D2, D3 = C::D3, Q2 = C::Q2, Q3 = C::Q3, NEB = C::NEB, NT = C::NT, NIO = C::NIO;
   constexpr int SXS = C::SXS, SQ = C::SQ, ES = C::ES;
#define Bm(q, d) P.bg.B[(q) + Q * (d)]
#define Gm(q, d) P.bg.G[(q) + Q * (d)]
   extern __shared__ double smem[];
   double *sXin = smem;
   double *sXout = sXin + C::SX_DOUBLES;
   double *sE = sXout + C::SX_DOUBLES;
   double *sBt = sE + C::SE_DOUBLES; // sBt[qy*D + dy] = B(qy,dy): the one runtime-indexed row of phase A
   double *sGt = sBt + Q * D;
   double *sQd = smem + C::WORK_DOUBLES;      // TMA mode: this batch's diffusion q-data (16-byte aligned)
   double *sQm = sQd + C::SQD_DOUBLES;        //           and mass q-data
   __shared__ unsigned long long qbar;        // TMA mode: "q-data of the current batch has landed"
   const int tid = threadIdx.x;
   if (P.done && *P.done) { return; }
   const int nbatch = (P.NE + NEB - 1) / NEB;
   for (int i = tid; i < Q * D; i += NT)
   {
      const int q = i / D, d = i - q * D;
      sBt[i] = P.bg.B[q + Q * d];
      sGt[i] = P.bg.G[q + Q * d];
   }

   // fixed per-thread roles
   const int eB = tid / Q2, cB = tid - eB * Q2; // phase B column
   const bool actB = tid < NEB * Q2;

   double O[DIFF ? Q : 1][6], Mq[MASS ? Q : 1]; // this thread's q-data column (prefetched one batch ahead)
   double xg[NIO];                              // gathered x values (prefetched)
   int gi[NIO];

   auto load_qdata = [&](int b)
   {
      const long long eg = (long long)b * NEB + eB;
      const bool ok = actB && eg < P.NE;
      B200PA_UNROLL
      for (int qz = 0; qz < Q; ++qz)
      {
         if (DIFF)
         {
            const double *d = P.pa_diff + (eg * 6) * Q3 + qz * Q2 + cB;
            B200PA_UNROLL
            for (int k = 0; k < 6; ++k) { O[qz][k] = ok ? __ldg(d + k * Q3) : 0.0; }
         }
         if (MASS) { Mq[qz] = ok ? __ldg(P.pa_mass + eg * Q3 + qz * Q2 + cB) : 0.0; }
      }
   };
   auto load_gidx = [&](int b)
   {
      const long long base = (long long)b * NEB * D3;
      const long long lim = (long long)P.NE * D3;
      B200PA_UNROLL
      for (int r = 0; r < NIO; ++r)
      {
         const int t = tid + r * NT;
         gi[r] = (t < NEB * D3 && base + t < lim) ? __ldg(P.gmap + base + t) : -1;
      }
   };
   auto load_x = [&]()
   {
      B200PA_UNROLL
      for (int r = 0; r < NIO; ++r) { xg[r] = gi[r] >= 0 ? P.x[gi[r]] : 0.0; }
   };

   // TMA mode: two bulk copies per batch (the batch's elements are contiguous in both q-data arrays).
   // The mass array's element stride (Q^3 doubles) is odd for odd Q, so its copy starts at the
   // 16-byte boundary below the batch and `mshift` (0 or 1 doubles) finds the data again.
   int mshift = 0;
   auto tma_issue = [&](int b)
   {
      const long long e0 = (long long)b * NEB;
      const int nel = (int)(P.NE - e0 < NEB ? P.NE - e0 : NEB);
      unsigned bytes_d = 0, bytes_m = 0;
      const double *src_m = nullptr;
      if (DIFF) { bytes_d = (unsigned)(nel * 6 * Q3 * sizeof(double)); }
      if (MASS)
      {
         const double *src = P.pa_mass + e0 * Q3;
         const int sh = (int)(((unsigned long long)src >> 3) & 1ull);
         src_m = src - sh;
         int nd = sh + nel * Q3;
         if (nd & 1)
         {
            // 16-byte granularity: take one double more, except at the very end of the array, where the
            // last double is fetched with a plain load instead (never read past the caller's buffer)
            if (e0 + nel < P.NE) { nd += 1; }
            else { nd -= 1; sQm[nd] = __ldg(src_m + nd); }
         }
         bytes_m = (unsigned)(nd * sizeof(double));
      }
      mbar_expect_tx(&qbar, bytes_d + bytes_m);
      if (DIFF) { tma_bulk_g2s(sQd, P.pa_diff + e0 * 6 * Q3, bytes_d, &qbar); }
      if (MASS) { tma_bulk_g2s(sQm, src_m, bytes_m, &qbar); }
   };
   auto mass_shift = [&](int b) { return (int)(((unsigned long long)(P.pa_mass + (long long)b * NEB * Q3) >> 3) & 1ull); };
   unsigned qphase = 0;
   if (C::TMA)
   {
      if (tid == 0) { mbar_init(&qbar, 1); }
      __syncthreads();
   }

   int batch = blockIdx.x;
   if (batch < nbatch)
   {
      load_gidx(batch);
      if (C::TMA) { if (tid == 0) { tma_issue(batch); } }
      else if (C::QPF) { load_qdata(batch); }
      load_x();
   }
   for (; batch < nbatch; batch += gridDim.x)
   {
      const long long base = (long long)batch * NEB * D3;
      const long long lim = (long long)P.NE * D3;
      const int next = batch + gridDim.x;

      // ------------------------------------------------------------- stage-in
      B200PA_UNROLL
      for (int r = 0; r < NIO; ++r)
      {
         const int t = tid + r * NT;
         if (t < NEB * D3)
         {
            const int slab = t / D2, k = t - slab * D2;
            sXin[slab * SXS + k] = xg[r];
         }
      }
      __syncthreads();
      if (next < nbatch) { load_gidx(next); } // indices for the next batch fly during phase A
      int sl[NIO];
      B200PA_UNROLL
      for (int r = 0; r < NIO; ++r)
      {
         const int t = tid + r * NT;
         sl[r] = (t < NEB * D3 && base + t < lim) ? __ldg(P.slot + base + t) : -1;
      }

      // ------------------------------------ phase A: (slab, qy) rows, y then x
      for (int task = tid; task < NEB * D * Q; task += NT)
      {
         const int slab = task / Q, qy = task - slab * Q;
         const int e = slab / D, dz = slab - e * D;
         const double *xs = sXin + slab * SXS;
         double bq[D], gq[D];
         B200PA_UNROLL
         for (int dy = 0; dy < D; ++dy) { bq[dy] = sBt[qy * D + dy]; if (DIFF) { gq[dy] = sGt[qy * D + dy]; } }
         double tB[D], tG[D];
         B200PA_UNROLL
         for (int dx = 0; dx < D; ++dx) { tB[dx] = 0.0; tG[dx] = 0.0; }
         B200PA_UNROLL
         for (int dy = 0; dy < D; ++dy)
         {
            B200PA_UNROLL
            for (int dx = 0; dx < D; ++dx)
            {
               const double xv = xs[dy * D + dx];
               tB[dx] = fma(bq[dy], xv, tB[dx]);
               if (DIFF) { tG[dx] = fma(gq[dy], xv, tG[dx]); }
            }
         }
         double *o = sE + e * ES + dz * SQ + qy * Q;
         B200PA_UNROLL
         for (int qx = 0; qx < Q; ++qx)
         {
            double f0 = 0.0, f1 = 0.0, f2 = 0.0;
            B200PA_UNROLL
            for (int dx = 0; dx < D; ++dx)
            {
               if (DIFF)
               {
                  f0 = fma(Gm(qx, dx), tB[dx], f0); // Gx By
                  f1 = fma(Bm(qx, dx), tG[dx], f1); // Bx Gy
               }
               f2 = fma(Bm(qx, dx), tB[dx], f2);    // Bx By
            }
            if (DIFF) { o[0 * D * SQ + qx] = f0; o[1 * D * SQ + qx] = f1; }
            o[2 * D * SQ + qx] = f2;
         }
      }
      __syncthreads();

      // -------------------------------- phase B: column, q-point op, column^T
      if (!C::QPF && !C::TMA) { load_qdata(batch); }
      if (C::TMA)
      {
         mbar_wait(&qbar, qphase);
         qphase ^= 1u;
         if (MASS) { mshift = mass_shift(batch); }
      }
      const double *qd = sQd + (eB * 6) * Q3 + cB;         // TMA mode: this thread's column in the staged q-data
      const double *qm = sQm + mshift + eB * Q3 + cB;
      if (actB)
      {
         double *s = sE + eB * ES + cB;
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
               double o0, o1, o2, o3, o4, o5;
               if (C::TMA)
               {
                  const double *d = qd + qz * Q2;
                  o0 = d[0]; o1 = d[Q3]; o2 = d[2 * Q3]; o3 = d[3 * Q3]; o4 = d[4 * Q3]; o5 = d[5 * Q3];
               }
               else { o0 = O[qz][0]; o1 = O[qz][1]; o2 = O[qz][2]; o3 = O[qz][3]; o4 = O[qz][4]; o5 = O[qz][5]; }
               hX = o0 * gX + o1 * gY + o2 * gZ;
               hY = o1 * gX + o3 * gY + o4 * gZ;
               hZ = o2 * gX + o4 * gY + o5 * gZ;
            }
            if (MASS) { hM = (C::TMA ? qm[qz * Q2] : Mq[qz]) * val; }
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
      // prefetch for the next batch: q-data column and gathered x (indices arrived during phase A)
      if (next < nbatch)
      {
         if (C::QPF && !C::TMA) { load_qdata(next); }
         load_x();
      }
      __syncthreads();
      // every thread is done reading the staged q-data: refill the buffer with the next batch; the copy
      // flies during phases C1/C2, stage-out, stage-in and phase A of the next batch
      if (C::TMA && next < nbatch && tid == 0) { tma_issue(next); }

      // ----------------------------------------------- phase C1: (slab, qy) rows, x^T
      for (int task = tid; task < NEB * D * Q; task += NT)
      {
         const int slab = task / Q, qy = task - slab * Q;
         const int e = slab / D, dz = slab - e * D;
         double *io = sE + e * ES + dz * SQ + qy * Q;
         double r0[Q], r1[Q], r2[Q];
         B200PA_UNROLL
         for (int qx = 0; qx < Q; ++qx)
         {
            if (DIFF) { r0[qx] = io[0 * D * SQ + qx]; r1[qx] = io[1 * D * SQ + qx]; }
            r2[qx] = io[2 * D * SQ + qx];
         }
         B200PA_UNROLL
         for (int dx = 0; dx < D; ++dx)
         {
            double s02 = 0.0, s1 = 0.0;
            B200PA_UNROLL
            for (int qx = 0; qx < Q; ++qx)
            {
               if (DIFF)
               {
                  s02 = fma(Gm(qx, dx), r0[qx], s02);
                  s1 = fma(Bm(qx, dx), r1[qx], s1);
               }
               s02 = fma(Bm(qx, dx), r2[qx], s02);
            }
            io[0 * D * SQ + dx] = s02; // row qy of field 0 / 1 now holds the x^T results (D <= Q)
            if (DIFF) { io[1 * D * SQ + dx] = s1; }
         }
      }
      __syncthreads();

      // ----------------------------------------------- phase C2: (slab, dx) columns, y^T
      for (int task = tid; task < NEB * D * D; task += NT)
      {
         const int slab = task / D, dx = task - slab * D;
         const int e = slab / D, dz = slab - e * D;
         const double *in = sE + e * ES + dz * SQ + dx;
         double out[D];
         B200PA_UNROLL
         for (int dy = 0; dy < D; ++dy) { out[dy] = 0.0; }
         B200PA_UNROLL
         for (int qy = 0; qy < Q; ++qy)
         {
            const double a = in[0 * D * SQ + qy * Q];
            const double b = DIFF ? in[1 * D * SQ + qy * Q] : 0.0;
            B200PA_UNROLL
            for (int dy = 0; dy < D; ++dy)
            {
               out[dy] = fma(Bm(qy, dy), a, out[dy]);
               if (DIFF) { out[dy] = fma(Gm(qy, dy), b, out[dy]); }
            }
         }
         double *xs = sXout + slab * SXS + dx;
         B200PA_UNROLL
         for (int dy = 0; dy < D; ++dy) { xs[dy * D] = out[dy]; }
      }
      __syncthreads();

      // --------------------------------------------------------------- stage-out
      B200PA_UNROLL
      for (int r = 0; r < NIO; ++r)
      {
         const int t = tid + r * NT;
         if (t < NEB * D3 && sl[r] >= 0)
         {
            const int slab = t / D2, k = t - slab * D2;
            P.y[sl[r]] = sXout[slab * SXS + k];
         }
      }
   }
#undef Bm
#undef Gm
}

} // namespace b200pa
