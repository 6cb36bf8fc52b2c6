// The hot kernel: fused L->E gather + sum-factorised diffusion (+ mass) + slot-layout E->L write,
// FP64, sm_100a.  One launch replaces the reference's K1 + K16 + K4 + K7 and the write half of K2
// (fem/restriction.cpp:109-129, fem/bilinearform_ext.cpp:543-557,
//  fem/integ/bilininteg_diffusion_kernels.hpp:989-1214, fem/integ/bilininteg_mass_kernels.hpp:807-1033).
//
// Shape (DESIGN.md 4.1 has the history: four designs, each driven by an ncu capture):
//   * persistent CTAs walk small batches of NEB elements; NEB*Q^2 threads = one q-point column each;
//   * nothing on the memory path returns into a register:
//       - the batch's q-data is ONE contiguous range of each q-data array: thread 0 issues two TMA bulk
//         copies (cp.async.bulk -> UBLKCP, completion on an mbarrier) into a shared-memory stage right
//         after the previous batch's columns are done; the copy flies during phases C1/C2/A;
//       - gather indices (two batches ahead), slots and - factorised form - element tensors (one batch
//         ahead) are copied to shared memory by per-thread 4/8-byte cp.async (LDGSTS); every thread
//         then reads ITS OWN indices back and gathers x with 8-byte cp.async (zero-fill form for
//         constrained entries) straight into the x buffer of the next batch;
//   * fine-grained tasks - (slab,qy) rows instead of whole slabs - so the small batch still fills the
//     CTA in the x/y contractions; all B/G operands except one row per task are compile-time indices
//     into the kernel-parameter constant bank;
//   * per-order lane->task maps and shared-memory strides from a bank-conflict model
//     (tools/smem_strides.py report): every phase over the work array conflict-free at p=2 and p=3.
//
//   (cp.async) : x_L[gather]                         -> sX[next][e][dz][dy][dx]
//   phase A  : task (e,dz,qy): y- then x-contraction -> sE[e][f][dz][qy][qx]        f < 3
//   phase B  : task (e,qx,qy): z, q-point op, z^T    -> sE (in place)
//   phase C1 : task (e,dz,qy): x^T                   -> sE[e][f][dz][qy][dx]        f < 2 (in place)
//   phase C2 : task (e,dz,dx): y^T                   -> y_S[slot]   (FUSE_OUT; else via sX and a stage-out pass)
//   slot = position in the E->L CSR: the segmented reduction that follows streams contiguously and
//   stays atomic-free, summing in ascending element order as fem/restriction.cpp:163-179.
#pragma once
#include <cuda_runtime.h>

#include "pa_element_kernel.cuh"
#include "tma.cuh"

namespace b200pa
{

// AFF: diffusion q-data in factorised form (affine elements): one scalar c_q = w_q k_q per q-point and the
// per-element tensor adj(J) adj(J)^T / det J (6 doubles) instead of 6 doubles per q-point
template <int D, int Q, bool AFF = false>
struct ApplyCfg
{
   static constexpr int D2 = D * D, D3 = D * D * D, Q2 = Q * Q, Q3 = Q * Q * Q;
   // elements per batch: one q-point column per thread
   // tuned on B200 with tools/tune.sh + tools/tune_run.sh (profiles/r1c_tuning.txt, r1g_tuning.txt);
   // the B200PA_TUNE_* macros exist for those tuning builds only
#ifdef B200PA_TUNE_NEB
   static constexpr int NEB = B200PA_TUNE_NEB;
#else
   static constexpr int NEB = (D == 2) ? 28 : (D == 3) ? (AFF ? 8 : 16) : (D == 4) ? 5 : (D == 5) ? 3 : (D == 6) ? 2 : 1;
#endif
#ifdef B200PA_TUNE_MINB
   static constexpr int MINB = B200PA_TUNE_MINB;
#else
   static constexpr int MINB = (D == 3) ? (AFF ? 4 : 2) : (D == 7 && AFF) ? 5 : 3; // resident CTAs/SM the register budget allows
#endif
#ifdef B200PA_TUNE_L2HINT
   static constexpr bool L2HINT = B200PA_TUNE_L2HINT;
#else
   static constexpr bool L2HINT = (D <= 3); // q-data is read once per apply: L2 evict_first (measured: +3 % at p=2, -4 % at p=5)
#endif
   // JAM tasks per thread in the row phases A, C1, C2 (unroll-and-jam): every B/G constant fetched into a uniform
   // register (LDCU, or an R2UR pair when it was spilled) then feeds JAM DFMAs instead of one.  Measured with JAM = 2:
   // 13-25 % SLOWER at p = 3..6 (profiles/r1n_tuning_unroll_and_jam.txt) - the row phases are short of threads, not of
   // issue slots; kept for tuning builds only
#ifdef B200PA_TUNE_JAM
   static constexpr int JAM = B200PA_TUNE_JAM;
#else
   static constexpr int JAM = 1;
#endif
   static_assert(JAM == 1 || JAM == 2, "one or two tasks per thread");
   // phase C2 stores its results straight to y_S[slot] (no staging through sX, one barrier less per batch)
#ifdef B200PA_TUNE_FUSE_OUT
   static constexpr bool FUSE_OUT = B200PA_TUNE_FUSE_OUT;
#else
   static constexpr bool FUSE_OUT = (D != 4); // measured: +2.5 % at p=4, -2 % at p=3, neutral elsewhere (profiles/r1j_*)
#endif
   static constexpr int NT = ((NEB * Q2 + 31) / 32) * 32;
   static constexpr int NIDX = NEB * D3;                         // E-entries per batch
   static constexpr int NIO = (NIDX + NT - 1) / NT;              // gather / scatter items per thread
   // shared-memory strides and lane -> task maps found by tools/smem_strides.py (fewest bank-conflict wavefronts
   // over all phases; the work-array phases are conflict-free at p=2 and p=3): SXS = slab stride of sX, RQ / SQ / ES = row / slab / element
   // stride of sE, BS = row stride of the staged basis rows; MAPA_SLAB / MAPC_SLAB: consecutive lanes of the
   // row phases A, C1 / of phase C2 walk the slabs (else the rows / columns of one slab)
#ifdef B200PA_TUNE_LAYOUT0 // the layout of rounds r1c..r1f
   static constexpr bool MAPA_SLAB = false, MAPC_SLAB = false;
   static constexpr int SXS = (D == 2) ? 6 : (D == 3) ? 9 : (D == 4) ? 20 : (D == 5) ? 25 : (D == 6) ? 38 : 55;
   static constexpr int SQ = (D == 2) ? 11 : (D == 3) ? 19 : (D == 4) ? 25 : (D == 5) ? 37 : (D == 6) ? 49 : 87;
   static constexpr int ES = (D == 2) ? 73 : (D == 3) ? 185 : (D == 4) ? 308 : (D == 5) ? 564 : (D == 6) ? 886 : 1841;
   static constexpr int BS = D;
   static constexpr int QES = 6 * Q3, QMS = Q3;
#else
   static constexpr bool MAPA_SLAB = (D == 4), MAPC_SLAB = (D == 4 || D == 6);
   static constexpr int SXS = (D == 2) ? 6 : (D == 3) ? 9 : (D == 4) ? 17 : (D == 5) ? 25 : (D == 6) ? 38 : 55;
   static constexpr int SQ = (D == 2) ? 11 : (D == 3) ? 19 : (D == 4) ? 28 : (D == 5) ? 37 : (D == 6) ? 49 : 87;
   static constexpr int ES = (D == 2) ? 73 : (D == 3) ? 185 : (D == 4) ? 345 : (D == 5) ? 564 : (D == 6) ? 886 : 1841;
   static constexpr int BS = D;
   // element strides of the staged q-data: where Q^2 is not a multiple of 16 lanes the elements of a batch are
   // staged by one bulk copy each, padded so that consecutive lanes keep hitting consecutive banks across the
   // element boundary (possible for even Q only: bulk copies need 16-byte aligned ends)
   static constexpr int QES = (D == 5) ? 6 * Q3 + 4 : 6 * Q3, QMS = (D == 5) ? Q3 + 12 : Q3;
#endif
   static constexpr int RQ = (D == 7) ? 10 : Q;
   static_assert(SXS >= D2 && SQ >= (Q - 1) * RQ + Q && ES >= 3 * D * SQ && BS >= D, "strides too small");
   static_assert((QES == 6 * Q3 && QMS == Q3) || (Q % 2 == 0 && QES % 2 == 0 && QMS % 2 == 0), "padded q-data staging needs even Q");
   // shared-memory map (bytes): two x buffers (gather target of the next batch | input and output of this one),
   // the work array sE, basis rows, three index buffers, and one staged batch of q-data in its global layout
   // ([e][6][Q^3] and [e][Q^3]; +2 doubles of slack each for the 16-byte alignment of the bulk copies)
   static constexpr int SX_DOUBLES = NEB * D * SXS;
   static constexpr int SE_DOUBLES = NEB * ES;
   static constexpr int WORK_DOUBLES = 2 * SX_DOUBLES + SE_DOUBLES + 2 * Q * BS;
   static constexpr int GEO_DOUBLES = AFF ? 2 * NEB * 6 : 0;      // sGeo[2][NEB][6]
   static constexpr int IDX_OFF = (WORK_DOUBLES + GEO_DOUBLES) * 8; // int sGi[2][NIDX], sSl[2][NIDX]
   static constexpr int QD_OFF = (IDX_OFF + 4 * NIDX * 4 + 15) & ~15;
   static constexpr int SQD_DOUBLES = AFF ? (((NEB * QMS + 2) + 1) & ~1) : NEB * QES + 2;
   static constexpr int SQM_DOUBLES = ((NEB * QMS + 2) + 1) & ~1;
   // q-data stages: 2 = the next batch's q-data is requested a whole batch ahead into a second stage.  Measured on
   // the factorised form (small stages): no gain at p <= 4, -10 % at p = 6 (profiles/r1l_*): one stage everywhere
#ifdef B200PA_TUNE_QSTAGES
   static constexpr int QSTAGES = B200PA_TUNE_QSTAGES;
#else
   static constexpr int QSTAGES = 1;
#endif
   static_assert(QSTAGES == 1 || QSTAGES == 2, "one or two q-data stages");
   static constexpr size_t SMEM_BYTES = QD_OFF + sizeof(double) * QSTAGES * (SQD_DOUBLES + SQM_DOUBLES);
};

template <int D, int Q, bool DIFF, bool MASS, bool AFF = false>
__global__ void __launch_bounds__(ApplyCfg<D, Q, AFF>::NT, ApplyCfg<D, Q, AFF>::MINB)
pa_apply_kernel(const __grid_constant__ ElemParams<D, Q> P)
{
   static_assert(!AFF || DIFF, "the factorised q-data is the diffusion integrator's");
   using C = ApplyCfg<D, Q, AFF>;
   constexpr int D2 = C::D2, D3 = C::D3, Q2 = C::Q2, Q3 = C::Q3, NEB = C::NEB, NT = C::NT, NIO = C::NIO, NIDX = C::NIDX;
   constexpr int SXS = C::SXS, SQ = C::SQ, ES = C::ES, RQ = C::RQ, BS = C::BS;
#define Bm(q, d) P.bg.B[(q) + Q * (d)]
#define Gm(q, d) P.bg.G[(q) + Q * (d)]
   extern __shared__ __align__(16) unsigned char smem_raw[];
   double *sX = reinterpret_cast<double *>(smem_raw);          // sX[2][NEB*D][SXS]
   double *sE = sX + 2 * C::SX_DOUBLES;
   double *sBt = sE + C::SE_DOUBLES; // sBt[qy*BS + dy] = B(qy,dy): the one runtime-indexed row of phase A
   double *sGt = sBt + Q * BS;
   double *sGeo = sGt + Q * BS;                                // AFF: sGeo[2][NEB][6], one batch ahead
   int *sGi = reinterpret_cast<int *>(smem_raw + C::IDX_OFF);  // sGi[2][NIDX]: gather indices, two batches deep
   int *sSl = sGi + 2 * NIDX;                                  // sSl[2][NIDX]: slots, one batch ahead
   double *sQ0 = reinterpret_cast<double *>(smem_raw + C::QD_OFF); // q-data stages (16-byte aligned): diffusion, then mass
   constexpr int QST = C::SQD_DOUBLES + C::SQM_DOUBLES;
   __shared__ unsigned long long qbars[2];                     // "the q-data of stage s has landed"
   const int tid = threadIdx.x;
   if (P.done && *P.done) { return; }
   const int nbatch = (P.NE + NEB - 1) / NEB;
   for (int i = tid; i < Q * D; i += NT)
   {
      const int q = i / D, d = i - q * D;
      sBt[q * BS + d] = P.bg.B[q + Q * d];
      sGt[q * BS + d] = P.bg.G[q + Q * d];
   }

   // fixed per-thread roles
   const int eB = tid / Q2, cB = tid - eB * Q2; // phase B column
   const bool actB = tid < NEB * Q2;
   const unsigned long long pol = C::L2HINT ? l2_policy_evict_first() : 0ull;
   const long long lim = (long long)P.NE * D3;

   // The gather runs two batches ahead without tying up a register: every thread owns the E-entries
   // t = tid + r*NT of a batch; it copies their gather indices (batch b+2) and slots (batch b) into shared
   // memory with 4-byte cp.async (LDGSTS), reads its own indices of batch b+1 back and gathers x with 8-byte
   // cp.async straight into the x buffer of batch b+1 (zero-filled for constrained entries, index < 0).
   auto copy_idx = [&](int *dst, const int *src, int b)
   {
      const long long base = (long long)b * NIDX;
      B200PA_UNROLL
      for (int r = 0; r < NIO; ++r)
      {
         const int t = tid + r * NT;
         if (t < NIDX)
         {
            if (base + t < lim)
            {
               cp_async4(dst + t, src + base + t);
            }
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
         if (t < NIDX)
         {
            const int slab = t / D2, k = t - slab * D2;
            const int g = idx[t];
            cp_async8_zfill(dst + slab * SXS + k, P.x + (g >= 0 ? g : 0), g >= 0);
         }
      }
   };

   // two bulk copies per batch (the batch's elements are contiguous in both q-data arrays).
   // A scalar q-field (mass; AFF: the diffusion coefficient too) has an element stride of Q^3 doubles, odd for
   // odd Q, so its copy starts at the 16-byte boundary below the batch and a shift of 0 or 1 doubles finds the
   // data again; the very last double of the array is fetched with a plain load (never read past the buffer).
   struct FieldCopy { const double *src; unsigned bytes; };
   auto scalar_field = [&](const double *arr, double *sdst, long long e0, int nel)
   {
      const double *src = arr + e0 * Q3;
      const int sh = (int)(((unsigned long long)src >> 3) & 1ull);
      src -= sh;
      int nd = sh + nel * Q3;
      if (nd & 1)
      {
         if (e0 + nel < P.NE) { nd += 1; }
         else { nd -= 1; sdst[nd] = __ldg(src + nd); }
      }
      return FieldCopy{src, (unsigned)(nd * sizeof(double))};
   };
   auto field_shift = [&](const double *arr, int b) { return (int)(((unsigned long long)(arr + (long long)b * NEB * Q3) >> 3) & 1ull); };
   int mshift = 0, dshift = 0;
   auto tma_issue = [&](int b, int stage)
   {
      double *sQd = sQ0 + stage * QST, *sQm = sQd + C::SQD_DOUBLES;
      unsigned long long &qbar = qbars[stage];
      const long long e0 = (long long)b * NEB;
      const int nel = (int)(P.NE - e0 < NEB ? P.NE - e0 : NEB);
      constexpr bool PADDED = C::QES != 6 * Q3 || C::QMS != Q3; // even Q: every element starts 16-byte aligned, no shift
      FieldCopy cd{nullptr, 0}, cm{nullptr, 0};
      if (DIFF) { cd = AFF ? scalar_field(P.pa_diff, sQd, e0, nel) : FieldCopy{P.pa_diff + e0 * 6 * Q3, (unsigned)(nel * 6 * Q3 * sizeof(double))}; }
      if (MASS) { cm = scalar_field(P.pa_mass, sQm, e0, nel); }
      mbar_expect_tx(&qbar, cd.bytes + cm.bytes);
      if (PADDED)
      {
         for (int e = 0; e < nel; ++e)
         {
            if (DIFF && !AFF) { tma_bulk_g2s(sQd + e * C::QES, P.pa_diff + (e0 + e) * 6 * Q3, (unsigned)(6 * Q3 * sizeof(double)), &qbar, pol); }
            if (DIFF && AFF) { tma_bulk_g2s(sQd + e * C::QMS, P.pa_diff + (e0 + e) * Q3, (unsigned)(Q3 * sizeof(double)), &qbar, pol); }
            if (MASS) { tma_bulk_g2s(sQm + e * C::QMS, P.pa_mass + (e0 + e) * Q3, (unsigned)(Q3 * sizeof(double)), &qbar, pol); }
         }
      }
      else
      {
         if (DIFF) { tma_bulk_g2s(sQd, cd.src, cd.bytes, &qbar, pol); }
         if (MASS) { tma_bulk_g2s(sQm, cm.src, cm.bytes, &qbar, pol); }
      }
   };
   // AFF: the per-element tensors of a batch, 6 doubles each, by 8-byte cp.async (threads < 6 NEB)
   auto copy_geo = [&](double *dst, int b)
   {
      if (AFF && tid < NEB * 6)
      {
         const long long i = (long long)b * NEB * 6 + tid;
         cp_async8_zfill(dst + tid, P.geo + (i < (long long)P.NE * 6 ? i : 0), i < (long long)P.NE * 6);
      }
   };
   unsigned qphase = 0; // bit s: parity of stage s
   if (tid == 0) { mbar_init(&qbars[0], 1); mbar_init(&qbars[1], 1); }
   __syncthreads();

   int batch = blockIdx.x;
   if (batch < nbatch)
   {
      if (tid == 0) { tma_issue(batch, 0); }
      copy_idx(sSl, P.slot, batch);
      copy_geo(sGeo, batch);
      copy_idx(sGi, P.gmap, batch);
      if (batch + (int)gridDim.x < nbatch) { copy_idx(sGi + NIDX, P.gmap, batch + gridDim.x); }
      cp_async_commit();
      cp_async_wait_all();                 // the one exposed index latency of the CTA (own entries only: no barrier)
      gather_x(sX, sGi);
      cp_async_commit();
   }
   int cur = 0;
   for (; batch < nbatch; batch += gridDim.x, cur ^= 1)
   {
      const int next = batch + gridDim.x;
      double *sXin = sX + cur * C::SX_DOUBLES; // input of this batch; its output too once phase A is done
      double *sXout = sXin;

      // ------------------------------------------------------------- stage-in
      cp_async_wait_all();
      __syncthreads();                     // x of this batch has landed; everybody is done with the previous batch
      if (next < nbatch) { gather_x(sX + (cur ^ 1) * C::SX_DOUBLES, sGi + (cur ^ 1) * NIDX); }
      if (next + (int)gridDim.x < nbatch) { copy_idx(sGi + cur * NIDX, P.gmap, next + gridDim.x); }
      if (next < nbatch) { copy_idx(sSl + (cur ^ 1) * NIDX, P.slot, next); copy_geo(sGeo + (cur ^ 1) * NEB * 6, next); }
      cp_async_commit();
      // two q-data stages: the next batch's q-data is requested now, a whole batch ahead
      if (C::QSTAGES == 2 && next < nbatch && tid == 0) { tma_issue(next, cur ^ 1); }

      // ------------------------------------ phase A: (slab, qy) rows, y then x
      {
         constexpr int JAM = C::JAM, NTASK = NEB * D * Q, NT2 = (NTASK + JAM - 1) / JAM;
         for (int t0 = tid; t0 < NT2; t0 += NT)
         {
            const double *xs[JAM];
            double *o[JAM];
            int qy[JAM];
            bool ok[JAM];
            B200PA_UNROLL
            for (int j = 0; j < JAM; ++j)
            {
               const int task = t0 + j * NT2;
               ok[j] = task < NTASK;
               const int tk = ok[j] ? task : t0;
               const int slab = C::MAPA_SLAB ? tk % (NEB * D) : tk / Q;
               qy[j] = C::MAPA_SLAB ? tk / (NEB * D) : tk - slab * Q;
               const int e = slab / D, dz = slab - e * D;
               xs[j] = sXin + slab * SXS;
               o[j] = sE + e * ES + dz * SQ + qy[j] * RQ;
            }
            double tB[JAM][D], tG[JAM][D];
            B200PA_UNROLL
            for (int j = 0; j < JAM; ++j)
            {
               double bq[D], gq[D];
               B200PA_UNROLL
               for (int dy = 0; dy < D; ++dy) { bq[dy] = sBt[qy[j] * BS + dy]; if (DIFF) { gq[dy] = sGt[qy[j] * BS + dy]; } }
               B200PA_UNROLL
               for (int dx = 0; dx < D; ++dx) { tB[j][dx] = 0.0; tG[j][dx] = 0.0; }
               B200PA_UNROLL
               for (int dy = 0; dy < D; ++dy)
               {
                  B200PA_UNROLL
                  for (int dx = 0; dx < D; ++dx)
                  {
                     const double xv = xs[j][dy * D + dx];
                     tB[j][dx] = fma(bq[dy], xv, tB[j][dx]);
                     if (DIFF) { tG[j][dx] = fma(gq[dy], xv, tG[j][dx]); }
                  }
               }
            }
            B200PA_UNROLL
            for (int qx = 0; qx < Q; ++qx)
            {
               double f0[JAM], f1[JAM], f2[JAM];
               B200PA_UNROLL
               for (int j = 0; j < JAM; ++j) { f0[j] = 0.0; f1[j] = 0.0; f2[j] = 0.0; }
               B200PA_UNROLL
               for (int dx = 0; dx < D; ++dx)
               {
                  B200PA_UNROLL
                  for (int j = 0; j < JAM; ++j)
                  {
                     if (DIFF)
                     {
                        f0[j] = fma(Gm(qx, dx), tB[j][dx], f0[j]); // Gx By
                        f1[j] = fma(Bm(qx, dx), tG[j][dx], f1[j]); // Bx Gy
                     }
                     f2[j] = fma(Bm(qx, dx), tB[j][dx], f2[j]);    // Bx By
                  }
               }
               B200PA_UNROLL
               for (int j = 0; j < JAM; ++j)
               {
                  if (ok[j])
                  {
                     if (DIFF) { o[j][0 * D * SQ + qx] = f0[j]; o[j][1 * D * SQ + qx] = f1[j]; }
                     o[j][2 * D * SQ + qx] = f2[j];
                  }
               }
            }
         }
      }
      __syncthreads();

      // -------------------------------- phase B: column, q-point op, column^T
      const int qs = C::QSTAGES == 2 ? cur : 0;
      mbar_wait(&qbars[qs], (qphase >> qs) & 1u);
      qphase ^= 1u << qs;
      const double *sQd = sQ0 + qs * QST, *sQm = sQd + C::SQD_DOUBLES;
      if (MASS) { mshift = C::QMS != Q3 ? 0 : field_shift(P.pa_mass, batch); }
      if (AFF) { dshift = C::QMS != Q3 ? 0 : field_shift(P.pa_diff, batch); }
      // this thread's column in the staged q-data
      const double *qd = AFF ? sQd + dshift + eB * C::QMS + cB : sQd + eB * C::QES + cB;
      const double *qm = sQm + mshift + eB * C::QMS + cB;
      if (actB)
      {
         double *s = sE + eB * ES + (cB / Q) * RQ + (cB % Q);
         double f0[D], f1[D], f2[D], p0[D], p1[D], p2[D];
         double Ce[6];
         if (AFF)
         {
            const double *g = sGeo + cur * NEB * 6 + eB * 6;
            B200PA_UNROLL
            for (int k = 0; k < 6; ++k) { Ce[k] = g[k]; }
         }
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
               if (AFF)
               {
                  const double c = qd[qz * Q2];
                  hX = c * (Ce[0] * gX + Ce[1] * gY + Ce[2] * gZ);
                  hY = c * (Ce[1] * gX + Ce[3] * gY + Ce[4] * gZ);
                  hZ = c * (Ce[2] * gX + Ce[4] * gY + Ce[5] * gZ);
               }
               else
               {
                  const double *d = qd + qz * Q2;
                  const double o0 = d[0], o1 = d[Q3], o2 = d[2 * Q3], o3 = d[3 * Q3], o4 = d[4 * Q3], o5 = d[5 * Q3];
                  hX = o0 * gX + o1 * gY + o2 * gZ;
                  hY = o1 * gX + o3 * gY + o4 * gZ;
                  hZ = o2 * gX + o4 * gY + o5 * gZ;
               }
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
      // every thread is done reading the staged q-data: refill the buffer with the next batch; the copy
      // flies during phases C1/C2, stage-out, stage-in and phase A of the next batch
      if (C::QSTAGES == 1 && next < nbatch && tid == 0) { tma_issue(next, 0); }

      // ----------------------------------------------- phase C1: (slab, qy) rows, x^T
      {
         constexpr int JAM = C::JAM, NTASK = NEB * D * Q, NT2 = (NTASK + JAM - 1) / JAM;
         for (int t0 = tid; t0 < NT2; t0 += NT)
         {
            double *io[JAM];
            bool ok[JAM];
            double r0[JAM][Q], r1[JAM][Q], r2[JAM][Q];
            B200PA_UNROLL
            for (int j = 0; j < JAM; ++j)
            {
               const int task = t0 + j * NT2;
               ok[j] = task < NTASK;
               const int tk = ok[j] ? task : t0;
               const int slab = C::MAPA_SLAB ? tk % (NEB * D) : tk / Q, qy = C::MAPA_SLAB ? tk / (NEB * D) : tk - slab * Q;
               const int e = slab / D, dz = slab - e * D;
               io[j] = sE + e * ES + dz * SQ + qy * RQ;
               B200PA_UNROLL
               for (int qx = 0; qx < Q; ++qx)
               {
                  if (DIFF) { r0[j][qx] = io[j][0 * D * SQ + qx]; r1[j][qx] = io[j][1 * D * SQ + qx]; }
                  r2[j][qx] = io[j][2 * D * SQ + qx];
               }
            }
            B200PA_UNROLL
            for (int dx = 0; dx < D; ++dx)
            {
               double s02[JAM], s1[JAM];
               B200PA_UNROLL
               for (int j = 0; j < JAM; ++j) { s02[j] = 0.0; s1[j] = 0.0; }
               B200PA_UNROLL
               for (int qx = 0; qx < Q; ++qx)
               {
                  B200PA_UNROLL
                  for (int j = 0; j < JAM; ++j)
                  {
                     if (DIFF)
                     {
                        s02[j] = fma(Gm(qx, dx), r0[j][qx], s02[j]);
                        s1[j] = fma(Bm(qx, dx), r1[j][qx], s1[j]);
                     }
                     s02[j] = fma(Bm(qx, dx), r2[j][qx], s02[j]);
                  }
               }
               B200PA_UNROLL
               for (int j = 0; j < JAM; ++j)
               {
                  if (ok[j])
                  {
                     io[j][0 * D * SQ + dx] = s02[j]; // row qy of field 0 / 1 now holds the x^T results (D <= Q)
                     if (DIFF) { io[j][1 * D * SQ + dx] = s1[j]; }
                  }
               }
            }
         }
      }
      __syncthreads();

      // ----------------------------------------------- phase C2: (slab, dx) columns, y^T
      {
         constexpr int JAM = C::JAM, NTASK = NEB * D * D, NT2 = (NTASK + JAM - 1) / JAM;
         for (int t0 = tid; t0 < NT2; t0 += NT)
         {
            const double *in[JAM];
            int slab[JAM], dx[JAM];
            bool ok[JAM];
            double out[JAM][D];
            B200PA_UNROLL
            for (int j = 0; j < JAM; ++j)
            {
               const int task = t0 + j * NT2;
               ok[j] = task < NTASK;
               const int tk = ok[j] ? task : t0;
               slab[j] = C::MAPC_SLAB ? tk % (NEB * D) : tk / D;
               dx[j] = C::MAPC_SLAB ? tk / (NEB * D) : tk - slab[j] * D;
               const int e = slab[j] / D, dz = slab[j] - e * D;
               in[j] = sE + e * ES + dz * SQ + dx[j];
               B200PA_UNROLL
               for (int dy = 0; dy < D; ++dy) { out[j][dy] = 0.0; }
            }
            B200PA_UNROLL
            for (int qy = 0; qy < Q; ++qy)
            {
               double a[JAM], b[JAM];
               B200PA_UNROLL
               for (int j = 0; j < JAM; ++j)
               {
                  a[j] = in[j][0 * D * SQ + qy * RQ];
                  b[j] = DIFF ? in[j][1 * D * SQ + qy * RQ] : 0.0;
               }
               B200PA_UNROLL
               for (int dy = 0; dy < D; ++dy)
               {
                  B200PA_UNROLL
                  for (int j = 0; j < JAM; ++j)
                  {
                     out[j][dy] = fma(Bm(qy, dy), a[j], out[j][dy]);
                     if (DIFF) { out[j][dy] = fma(Gm(qy, dy), b[j], out[j][dy]); }
                  }
               }
            }
            B200PA_UNROLL
            for (int j = 0; j < JAM; ++j)
            {
               if (!ok[j]) { continue; }
               if (C::FUSE_OUT)
               {
                  // slot = position in the E->L CSR (copied into sSl a whole batch ago; < 0 beyond the last element)
                  const int *sl = sSl + cur * NIDX + slab[j] * D2 + dx[j];
                  B200PA_UNROLL
                  for (int dy = 0; dy < D; ++dy)
                  {
                     const int k = sl[dy * D];
                     if (k >= 0) { P.y[k] = out[j][dy]; }
                  }
               }
               else
               {
                  double *xs = sXout + slab[j] * SXS + dx[j];
                  B200PA_UNROLL
                  for (int dy = 0; dy < D; ++dy) { xs[dy * D] = out[j][dy]; }
               }
            }
         }
      }
      if (C::FUSE_OUT) { continue; }
      __syncthreads();

      // --------------------------------------------------------------- stage-out
      B200PA_UNROLL
      for (int r = 0; r < NIO; ++r)
      {
         const int t = tid + r * NT;
         if (t < NIDX)
         {
            const int sl = sSl[cur * NIDX + t];
            if (sl >= 0)
            {
               const int slab = t / D2, k = t - slab * D2;
               P.y[sl] = sXout[slab * SXS + k];
            }
         }
      }
   }
#undef Bm
#undef Gm
}

} // namespace b200pa
