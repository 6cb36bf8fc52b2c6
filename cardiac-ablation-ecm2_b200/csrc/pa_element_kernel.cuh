// Sum-factorised 3-D hex element kernel for sm_100a, FP64 — the "slab / column" design.
//
// One launch covers what the reference does in several (K1 gather, K16 zero,
// K4 diffusion apply, K7 mass apply, and the write half of K2; also K13, K14, K15 through the
// q-point functor).  Reference behaviour being reproduced (not its code):
//   fem/integ/bilininteg_diffusion_kernels.hpp:989-1214, fem/integ/bilininteg_mass_kernels.hpp:807-1033,
//   fem/qinterp/eval.hpp:131-193, fem/qinterp/grad.hpp:233-374, fem/integ/lininteg_domain_kernels.hpp:164-298.
//
// Why this shape on B200: per SM the chip has 64 DFMA/clk but only 128 B/clk of shared-memory
// bandwidth (16 doubles/clk) against ~23 B/clk of HBM; the classic "one thread per (qx,qy), all
// contractions through shared memory" layout moves ~4x more doubles through shared memory than
// through the FP64 pipe can afford.  Here a thread owns a whole xy-SLAB (fixed z) for the x- and
// y-contractions and a whole z-COLUMN (fixed qx,qy) for the z-contraction and the q-point
// operation, so an element crosses shared memory only twice in each direction
// (12 Q^2 D doubles), every B/G operand is a compile-time index into the kernel-parameter
// constant bank (free DFMA operand), and registers per thread stay near D^2 + 3Q.
//
//   stage-in : x_E (or x_L through the gather map)              -> sX[e][dz][dy][dx]
//   phase A  : thread (e,dz)   : y- then x-contraction          -> sE[e][f][dz][qy][qx], f<3
//   phase B  : thread (e,qx,qy): z-contraction, q-point op, z^T -> sE (in place)
//   phase C  : thread (e,dz)   : x^T then y^T                   -> sX
//   stage-out: sX -> y_E (+=), or -> y_S[slot] (slot = position in the E->L CSR, so the
//              following segmented reduction streams contiguously and stays atomic-free)
#pragma once
#include <cuda_runtime.h>

#include "geometry.cuh"

namespace b200pa
{

template <int D, int Q>
struct BasisT
{
   double B[Q * D]; // B[q + Q*d]  (fem/fe/fe_base.cpp:2654)
   double G[Q * D];
};

enum InMode { IN_E = 0, IN_GATHER = 1, IN_NONE = 2 };
enum OutMode { OUT_E_ADD = 0, OUT_E_SET = 1, OUT_SLOT = 2, OUT_NONE = 3 };
enum QOp { QOP_APPLY = 0, QOP_VALUES = 1, QOP_PHYSGRAD = 2, QOP_LF = 3, QOP_COEFF = 4, QOP_JOULE = 5 };

template <int D, int Q>
struct ElemParams
{
   BasisT<D, Q> bg;
   int NE;
   const double *__restrict__ x;       // x_E [D^3,NE] or x_L
   const int *__restrict__ gmap;       // IN_GATHER: E->L index, <0 means "constrained: read 0"
   double *__restrict__ y;             // y_E, y_S, or q-point output
   const int *__restrict__ slot;       // OUT_SLOT
   const double *__restrict__ pa_diff; // [Q^3,6,NE] or null
   const double *__restrict__ pa_mass; // [Q^3,NE]   or null
   const double *__restrict__ geo;     // pa_apply_kernel<.., AFF>: adj(J) adj(J)^T / det J per element [6,NE]; pa_diff is then [Q^3,NE]
   const double *__restrict__ J;       // QOP_PHYSGRAD: [Q^3,3,3,NE]
   const double *__restrict__ f;       // QOP_LF: f [Q^3,NE] or 1 value
   const double *__restrict__ detJ;    // QOP_LF: per q-point [Q^3,NE], or null and
   const double *__restrict__ detE;    //         one determinant per (affine) element [NE]
   const double *__restrict__ W;       // QOP_LF
   long long nf;
   const int *done;                    // PCG early-exit flag (device) or null
   double ca, cb, cT0;                 // QOP_COEFF: a (1 + b (T - T0)); QOP_JOULE: s |grad|^2 + a
   const double *__restrict__ s;       // QOP_JOULE: sigma at q-points [Q^3,NE]
   // QOP_PHYSGRAD / QOP_JOULE without stored Jacobians (J == null): trilinear geometry from the vertices
   const double *__restrict__ jinv;    // affine elements: rows of J^{-T} per ELEMENT [9,NE] (takes precedence over J / vtx)
   const double *__restrict__ vtx;     // [3,nv]
   const int *__restrict__ ev;         // [8,NE]
   double xi[Q];                       // 1-D Gauss-Legendre points
};

// launch shape per (D,Q): elements per CTA and threads per CTA
template <int D, int Q> struct ElemCfg { static constexpr int NEB = 8, NT = 128; };
template <> struct ElemCfg<2, 3> { static constexpr int NEB = 32, NT = 96; };
template <> struct ElemCfg<3, 4> { static constexpr int NEB = 32, NT = 128; };
template <> struct ElemCfg<4, 5> { static constexpr int NEB = 16, NT = 128; };
template <> struct ElemCfg<5, 6> { static constexpr int NEB = 6, NT = 128; };
template <> struct ElemCfg<6, 7> { static constexpr int NEB = 5, NT = 128; };
template <> struct ElemCfg<7, 8> { static constexpr int NEB = 4, NT = 128; };

template <int D, int Q>
struct ElemLayout
{
   static constexpr int D3 = D * D * D, Q2 = Q * Q, Q3 = Q * Q * Q;
   static constexpr int SX = (D * D) | 1; // padded slab strides (odd => conflict-free across lanes)
   static constexpr int SQ = (Q * Q) | 1;
   static constexpr int NEB = ElemCfg<D, Q>::NEB, NT = ElemCfg<D, Q>::NT;
   static constexpr int SX_DOUBLES = NEB * D * SX;
   static constexpr int SE_DOUBLES = NEB * 3 * D * SQ;
   static constexpr size_t SMEM_BYTES = sizeof(double) * (SX_DOUBLES + SE_DOUBLES);
};

#define B200PA_UNROLL _Pragma("unroll")

template <int D, int Q, bool DIFF, bool MASS, int INMODE, int OUTMODE, int QOP>
__global__ void __launch_bounds__(ElemCfg<D, Q>::NT)
pa_element_kernel(const __grid_constant__ ElemParams<D, Q> P)
{
   using L = ElemLayout<D, Q>;
   constexpr int D3 = L::D3, Q2 = L::Q2, Q3 = L::Q3, SX = L::SX, SQ = L::SQ, NEB = L::NEB, NT = L::NT;
   // which intermediate fields are live
   constexpr bool GRAD = (QOP == QOP_APPLY && DIFF) || QOP == QOP_PHYSGRAD || QOP == QOP_JOULE;
   constexpr bool VAL = (QOP == QOP_APPLY && MASS) || QOP == QOP_VALUES || QOP == QOP_LF || QOP == QOP_COEFF;
   constexpr bool FWD = QOP != QOP_LF;
   constexpr bool BWD = QOP == QOP_APPLY || QOP == QOP_LF;
#define Bm(q, d) P.bg.B[(q) + Q * (d)]
#define Gm(q, d) P.bg.G[(q) + Q * (d)]

   extern __shared__ double smem[];
   double *sX = smem;
   double *sE = smem + L::SX_DOUBLES;
   const int tid = threadIdx.x;
   const int nbatch = (P.NE + NEB - 1) / NEB;
   if (P.done && *P.done) { return; }

   for (int batch = blockIdx.x; batch < nbatch; batch += gridDim.x)
   {
      const int e0 = batch * NEB;
      const int nel = min(NEB, P.NE - e0);
      const long long base = (long long)e0 * D3;

      // ---------------------------------------------------------------- stage-in
      if (FWD)
      {
         for (int t = tid; t < nel * D3; t += NT)
         {
            const int e = t / D3, l = t - e * D3;
            const int dz = l / (D * D), k = l - dz * (D * D);
            double v;
            if (INMODE == IN_GATHER)
            {
               const int g = P.gmap[base + t];
               v = g >= 0 ? P.x[g] : 0.0;
            }
            else { v = P.x[base + t]; }
            sX[(e * D + dz) * SX + k] = v;
         }
         __syncthreads();

         // ------------------------------------------------ phase A: slab forward
         for (int task = tid; task < nel * D; task += NT)
         {
            const int e = task / D, dz = task - e * D;
            const double *xs = sX + task * SX;
            double xr[D][D];
            B200PA_UNROLL
            for (int dy = 0; dy < D; ++dy)
            {
               B200PA_UNROLL
               for (int dx = 0; dx < D; ++dx) { xr[dy][dx] = xs[dy * D + dx]; }
            }
            double *o = sE + ((e * 3) * D + dz) * SQ;
            B200PA_UNROLL
            for (int qy = 0; qy < Q; ++qy)
            {
               double tB[D], tG[D];
               B200PA_UNROLL
               for (int dx = 0; dx < D; ++dx)
               {
                  double b = 0.0, g = 0.0;
                  B200PA_UNROLL
                  for (int dy = 0; dy < D; ++dy)
                  {
                     b = fma(Bm(qy, dy), xr[dy][dx], b);
                     if (GRAD) { g = fma(Gm(qy, dy), xr[dy][dx], g); }
                  }
                  tB[dx] = b;
                  tG[dx] = g;
               }
               B200PA_UNROLL
               for (int qx = 0; qx < Q; ++qx)
               {
                  double f0 = 0.0, f1 = 0.0, f2 = 0.0;
                  B200PA_UNROLL
                  for (int dx = 0; dx < D; ++dx)
                  {
                     if (GRAD)
                     {
                        f0 = fma(Gm(qx, dx), tB[dx], f0); // Gx By
                        f1 = fma(Bm(qx, dx), tG[dx], f1); // Bx Gy
                     }
                     f2 = fma(Bm(qx, dx), tB[dx], f2);    // Bx By
                  }
                  if (GRAD)
                  {
                     o[0 * D * SQ + qy * Q + qx] = f0;
                     o[1 * D * SQ + qy * Q + qx] = f1;
                  }
                  o[2 * D * SQ + qy * Q + qx] = f2;
               }
            }
         }
         __syncthreads();
      }

      // ---------------------------------------- phase B: column, q-point, column^T
      for (int task = tid; task < nel * Q2; task += NT)
      {
         const int e = task / Q2, c = task - e * Q2;
         const long long eg = e0 + e;
         double *s = sE + (e * 3) * D * SQ + c;
         double f0[D], f1[D], f2[D], p0[D], p1[D], p2[D];
         if (FWD)
         {
            B200PA_UNROLL
            for (int dz = 0; dz < D; ++dz)
            {
               if (GRAD) { f0[dz] = s[(0 * D + dz) * SQ]; f1[dz] = s[(1 * D + dz) * SQ]; }
               f2[dz] = s[(2 * D + dz) * SQ];
            }
         }
         B200PA_UNROLL
         for (int dz = 0; dz < D; ++dz) { p0[dz] = 0.0; p1[dz] = 0.0; p2[dz] = 0.0; }

         // q-data for the whole column first: Q*(6+1) independent loads in flight per thread
         double O[Q][6], Mq[Q];
         if (QOP == QOP_APPLY)
         {
            B200PA_UNROLL
            for (int qz = 0; qz < Q; ++qz)
            {
               if (DIFF)
               {
                  const double *d = P.pa_diff + (eg * 6) * Q3 + qz * Q2 + c;
                  B200PA_UNROLL
                  for (int k = 0; k < 6; ++k) { O[qz][k] = __ldg(d + k * Q3); }
               }
               if (MASS) { Mq[qz] = __ldg(P.pa_mass + eg * Q3 + qz * Q2 + c); }
            }
         }

         // q-point inputs of the other operations, also for the whole column up front: the loads of one thread are then in
         // flight together instead of one DRAM latency per q-point (nvcc does not hoist them over the stores of the loop)
         double Sq[Q], Fq[Q], Dq[Q], Ji[9];
         if (QOP == QOP_JOULE)
         {
            B200PA_UNROLL
            for (int qz = 0; qz < Q; ++qz) { Sq[qz] = __ldg(P.s + eg * Q3 + qz * Q2 + c); }
         }
         if (QOP == QOP_LF)
         {
            B200PA_UNROLL
            for (int qz = 0; qz < Q; ++qz)
            {
               Fq[qz] = P.nf == 1 ? __ldg(P.f) : __ldg(P.f + eg * Q3 + qz * Q2 + c);
               Dq[qz] = P.detJ ? __ldg(P.detJ + eg * Q3 + qz * Q2 + c) : __ldg(P.detE + eg);
            }
         }
         if ((QOP == QOP_PHYSGRAD || QOP == QOP_JOULE) && P.jinv)
         {
            B200PA_UNROLL
            for (int k = 0; k < 9; ++k) { Ji[k] = __ldg(P.jinv + eg * 9 + k); }
         }

         B200PA_UNROLL
         for (int qz = 0; qz < Q; ++qz)
         {
            double gX = 0.0, gY = 0.0, gZ = 0.0, val = 0.0;
            if (FWD)
            {
               B200PA_UNROLL
               for (int dz = 0; dz < D; ++dz)
               {
                  if (GRAD)
                  {
                     gX = fma(Bm(qz, dz), f0[dz], gX);
                     gY = fma(Bm(qz, dz), f1[dz], gY);
                     gZ = fma(Gm(qz, dz), f2[dz], gZ);
                  }
                  if (VAL) { val = fma(Bm(qz, dz), f2[dz], val); }
               }
            }
            const long long q = qz * Q2 + c;
            if (QOP == QOP_APPLY)
            {
               double hX = 0.0, hY = 0.0, hZ = 0.0, hM = 0.0;
               if (DIFF)
               {
                  hX = O[qz][0] * gX + O[qz][1] * gY + O[qz][2] * gZ;
                  hY = O[qz][1] * gX + O[qz][3] * gY + O[qz][4] * gZ;
                  hZ = O[qz][2] * gX + O[qz][4] * gY + O[qz][5] * gZ;
               }
               if (MASS) { hM = Mq[qz] * val; }
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
            else if (QOP == QOP_VALUES) { P.y[eg * Q3 + q] = val; }
            else if (QOP == QOP_COEFF) { P.y[eg * Q3 + q] = P.ca * (1.0 + P.cb * (val - P.cT0)); }
            else if (QOP == QOP_PHYSGRAD || QOP == QOP_JOULE)
            {
               // J^{-T} (gX,gY,gZ): adjugate / det, fem/qinterp/grad.hpp:340-352
               double g0, g1, g2;
               if (P.jinv)
               {
                  // affine element: J is constant over it, its inverse was formed once per element at geometry time
                  g0 = Ji[0] * gX + Ji[1] * gY + Ji[2] * gZ;
                  g1 = Ji[3] * gX + Ji[4] * gY + Ji[5] * gZ;
                  g2 = Ji[6] * gX + Ji[7] * gY + Ji[8] * gZ;
               }
               else
               {
               double a0, a1, a2, a3, a4, a5, a6, a7, a8;
               if (P.J)
               {
                  const double *Je = P.J + eg * 9 * Q3 + q;
                  a0 = Je[0 * Q3]; a1 = Je[1 * Q3]; a2 = Je[2 * Q3];
                  a3 = Je[3 * Q3]; a4 = Je[4 * Q3]; a5 = Je[5 * Q3];
                  a6 = Je[6 * Q3]; a7 = Je[7 * Q3]; a8 = Je[8 * Q3];
               }
               else
               {
                  double Jm[9];
                  trilinear_jacobian(P.vtx, P.ev + 8 * eg, P.xi[c % Q], P.xi[c / Q], P.xi[qz], Jm);
                  a0 = Jm[0]; a1 = Jm[1]; a2 = Jm[2]; a3 = Jm[3]; a4 = Jm[4]; a5 = Jm[5]; a6 = Jm[6]; a7 = Jm[7]; a8 = Jm[8];
               }
               const double i0 = a4 * a8 - a5 * a7, i1 = a2 * a7 - a1 * a8, i2 = a1 * a5 - a2 * a4;
               const double i3 = a5 * a6 - a3 * a8, i4 = a0 * a8 - a2 * a6, i5 = a2 * a3 - a0 * a5;
               const double i6 = a3 * a7 - a4 * a6, i7 = a1 * a6 - a0 * a7, i8 = a0 * a4 - a1 * a3;
               const double idet = 1.0 / (a0 * i0 + a1 * i3 + a2 * i6);
               g0 = (i0 * gX + i1 * gY + i2 * gZ) * idet;
               g1 = (i3 * gX + i4 * gY + i5 * gZ) * idet;
               g2 = (i6 * gX + i7 * gY + i8 * gZ) * idet;
               }
               if (QOP == QOP_PHYSGRAD)
               {
                  double *g = P.y + 3 * (eg * Q3 + q);
                  g[0] = g0; g[1] = g1; g[2] = g2;
               }
               else { P.y[eg * Q3 + q] = Sq[qz] * (g0 * g0 + g1 * g1 + g2 * g2) + P.ca; }
            }
            else if (QOP == QOP_LF)
            {
               const double hM = P.W[q] * Fq[qz] * Dq[qz];
               B200PA_UNROLL
               for (int dz = 0; dz < D; ++dz) { p2[dz] = fma(Bm(qz, dz), hM, p2[dz]); }
            }
         }
         if (BWD)
         {
            B200PA_UNROLL
            for (int dz = 0; dz < D; ++dz)
            {
               if (QOP == QOP_APPLY && DIFF) { s[(0 * D + dz) * SQ] = p0[dz]; s[(1 * D + dz) * SQ] = p1[dz]; }
               s[(2 * D + dz) * SQ] = p2[dz];
            }
         }
      }
      if (!BWD) { __syncthreads(); continue; }
      __syncthreads();

      // ----------------------------------------------- phase C: slab transpose
      constexpr bool BG3 = (QOP == QOP_APPLY && DIFF); // all three fields live
      for (int task = tid; task < nel * D; task += NT)
      {
         const int e = task / D, dz = task - e * D;
         const double *in = sE + ((e * 3) * D + dz) * SQ;
         double out[D][D];
         B200PA_UNROLL
         for (int dy = 0; dy < D; ++dy)
         {
            B200PA_UNROLL
            for (int dx = 0; dx < D; ++dx) { out[dy][dx] = 0.0; }
         }
         B200PA_UNROLL
         for (int qy = 0; qy < Q; ++qy)
         {
            double r0[Q], r1[Q], r2[Q];
            B200PA_UNROLL
            for (int qx = 0; qx < Q; ++qx)
            {
               if (BG3) { r0[qx] = in[0 * D * SQ + qy * Q + qx]; r1[qx] = in[1 * D * SQ + qy * Q + qx]; }
               r2[qx] = in[2 * D * SQ + qy * Q + qx];
            }
            B200PA_UNROLL
            for (int dx = 0; dx < D; ++dx)
            {
               double s02 = 0.0, s1 = 0.0;
               B200PA_UNROLL
               for (int qx = 0; qx < Q; ++qx)
               {
                  if (BG3)
                  {
                     s02 = fma(Gm(qx, dx), r0[qx], s02);
                     s1 = fma(Bm(qx, dx), r1[qx], s1);
                  }
                  s02 = fma(Bm(qx, dx), r2[qx], s02);
               }
               B200PA_UNROLL
               for (int dy = 0; dy < D; ++dy)
               {
                  out[dy][dx] = fma(Bm(qy, dy), s02, out[dy][dx]);
                  if (BG3) { out[dy][dx] = fma(Gm(qy, dy), s1, out[dy][dx]); }
               }
            }
         }
         double *xs = sX + task * SX;
         B200PA_UNROLL
         for (int dy = 0; dy < D; ++dy)
         {
            B200PA_UNROLL
            for (int dx = 0; dx < D; ++dx) { xs[dy * D + dx] = out[dy][dx]; }
         }
      }
      __syncthreads();

      // --------------------------------------------------------------- stage-out
      for (int t = tid; t < nel * D3; t += NT)
      {
         const int e = t / D3, l = t - e * D3;
         const int dz = l / (D * D), k = l - dz * (D * D);
         const double v = sX[(e * D + dz) * SX + k];
         if (OUTMODE == OUT_E_ADD) { P.y[base + t] += v; }
         else if (OUTMODE == OUT_E_SET) { P.y[base + t] = v; }
         else if (OUTMODE == OUT_SLOT) { P.y[P.slot[base + t]] = v; }
      }
      __syncthreads();
   }
#undef Bm
#undef Gm
}

} // namespace b200pa
