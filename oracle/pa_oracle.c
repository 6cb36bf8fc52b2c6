/* TEST INFRASTRUCTURE ONLY — see pa_oracle.h.  Plain C restatement of the reference's
 * CPU algorithm, loop nest by loop nest, with the reference's summation order (the "Smem"
 * kernels are the ones the reference dispatches for H1 orders 1..7 on hexes, also on the
 * CPU — SURVEY.md §2.2).  Build with -ffp-contract=off: the reference's default x86-64
 * build has no FMA, and bit-for-bit agreement with it is what pins this file.
 * Parity status: PINNED against oracle/_ref (tests/test_oracle_golden.py).
 */
#include "pa_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define MX 10 /* max D1D / Q1D handled (reference registers up to D1D 9 / Q1D 9) */

/* ------------------------------------------------------------------ restriction */
/* fem/restriction.cpp:66-106 */
void orc_restriction_tables(int ne, int nd, int ndofs, const int *gather_map, int *offsets, int *indices)
{
   for (int i = 0; i <= ndofs; ++i) { offsets[i] = 0; }
   for (long l = 0; l < (long)ne * nd; ++l)
   {
      const int s = gather_map[l];
      const int gid = s >= 0 ? s : -1 - s;
      ++offsets[gid + 1];
   }
   for (int i = 1; i <= ndofs; ++i) { offsets[i] += offsets[i - 1]; }
   for (long l = 0; l < (long)ne * nd; ++l) /* ascending element order */
   {
      const int s = gather_map[l];
      const int gid = s >= 0 ? s : -1 - s;
      indices[offsets[gid]++] = s >= 0 ? (int)l : -1 - (int)l;
   }
   for (int i = ndofs; i > 0; --i) { offsets[i] = offsets[i - 1]; }
   offsets[0] = 0;
}

/* fem/restriction.cpp:109-129 */
void orc_restrict_mult(int ne, int nd, const int *gather_map, const double *x, double *y)
{
   for (long i = 0; i < (long)ne * nd; ++i)
   {
      const int gid = gather_map[i];
      const int j = gid >= 0 ? gid : -1 - gid;
      y[i] = gid >= 0 ? x[j] : -x[j];
   }
}

/* fem/restriction.cpp:152-186, 196-221 */
void orc_restrict_mult_transpose(int ndofs, const int *offsets, const int *indices, const double *xE,
                                 double *yL, int abs)
{
   for (int i = 0; i < ndofs; ++i)
   {
      double v = 0;
      for (int j = offsets[i]; j < offsets[i + 1]; ++j)
      {
         const int k = indices[j] >= 0 ? indices[j] : -1 - indices[j];
         v += (abs || indices[j] >= 0) ? xE[k] : -xE[k];
      }
      yL[i] = v;
   }
}

/* ------------------------------------------------------------------------ setup */
/* fem/integ/bilininteg_diffusion_kernels.cpp:243-367 (coeffDim == 1) */
void orc_diffusion_setup(int Q1D, int NE, const double *W, const double *J, const double *C, long nc, double *D)
{
   const long NQ = (long)Q1D * Q1D * Q1D;
   const int const_c = (nc == 1);
   for (long e = 0; e < NE; ++e)
   {
      for (long q = 0; q < NQ; ++q)
      {
         const double *Je = J + e * 9 * NQ + q; /* J(q,row,col,e) = Je[(row + 3*col)*NQ] */
         const double J11 = Je[0 * NQ], J21 = Je[1 * NQ], J31 = Je[2 * NQ];
         const double J12 = Je[3 * NQ], J22 = Je[4 * NQ], J32 = Je[5 * NQ];
         const double J13 = Je[6 * NQ], J23 = Je[7 * NQ], J33 = Je[8 * NQ];
         const double detJ = J11 * (J22 * J33 - J32 * J23) -
                             J21 * (J12 * J33 - J32 * J13) +
                             J31 * (J12 * J23 - J22 * J13);
         const double w_detJ = W[q] / detJ;
         const double A11 = (J22 * J33) - (J23 * J32);
         const double A12 = (J32 * J13) - (J12 * J33);
         const double A13 = (J12 * J23) - (J22 * J13);
         const double A21 = (J31 * J23) - (J21 * J33);
         const double A22 = (J11 * J33) - (J13 * J31);
         const double A23 = (J21 * J13) - (J11 * J23);
         const double A31 = (J21 * J32) - (J31 * J22);
         const double A32 = (J31 * J12) - (J11 * J32);
         const double A33 = (J11 * J22) - (J12 * J21);
         const double C1 = const_c ? C[0] : C[e * NQ + q];
         const double C2 = C1, C3 = C1;
         double *De = D + e * 6 * NQ + q;
         De[0 * NQ] = w_detJ * (C1 * A11 * A11 + C2 * A12 * A12 + C3 * A13 * A13);
         De[1 * NQ] = w_detJ * (C1 * A11 * A21 + C2 * A12 * A22 + C3 * A13 * A23);
         De[2 * NQ] = w_detJ * (C1 * A11 * A31 + C2 * A12 * A32 + C3 * A13 * A33);
         De[3 * NQ] = w_detJ * (C1 * A21 * A21 + C2 * A22 * A22 + C3 * A23 * A23);
         De[4 * NQ] = w_detJ * (C1 * A21 * A31 + C2 * A22 * A32 + C3 * A23 * A33);
         De[5 * NQ] = w_detJ * (C1 * A31 * A31 + C2 * A32 * A32 + C3 * A33 * A33);
      }
   }
}

/* fem/integ/bilininteg_mass_pa.cpp:62-78 */
void orc_mass_setup(int NQ, int NE, const double *W, const double *detJ, const double *C, long nc, double *v)
{
   const int const_c = (nc == 1);
   for (long e = 0; e < NE; ++e)
   {
      for (long q = 0; q < NQ; ++q)
      {
         const double c = const_c ? C[0] : C[e * NQ + q];
         v[e * NQ + q] = W[q] * c * detJ[e * NQ + q];
      }
   }
}

/* ------------------------------------------------------------------------ apply */
/* fem/integ/bilininteg_diffusion_kernels.hpp:989-1214 */
void orc_diffusion_apply(int NE, int D1D, int Q1D, const double *B, const double *G, const double *D,
                         const double *xE, double *yE)
{
   const long NQ = (long)Q1D * Q1D * Q1D, ND = (long)D1D * D1D * D1D;
   static double s0[3][MX][MX][MX], s1[3][MX][MX][MX];
#define b_(q, d) B[(q) + Q1D * (d)]
#define g_(q, d) G[(q) + Q1D * (d)]
   for (long e = 0; e < NE; ++e)
   {
      const double *x = xE + e * ND;
      const double *d = D + e * 6 * NQ;
      double *y = yE + e * ND;
      /* :1067-1087  X contraction */
      for (int dz = 0; dz < D1D; ++dz)
         for (int dy = 0; dy < D1D; ++dy)
            for (int qx = 0; qx < Q1D; ++qx)
            {
               double u = 0.0, v = 0.0;
               for (int dx = 0; dx < D1D; ++dx)
               {
                  const double c = x[dx + D1D * (dy + D1D * dz)];
                  u += c * b_(qx, dx);
                  v += c * g_(qx, dx);
               }
               s0[0][dz][dy][qx] = u; /* DDQ0 */
               s0[1][dz][dy][qx] = v; /* DDQ1 */
            }
      /* :1089-1110  Y contraction */
      for (int dz = 0; dz < D1D; ++dz)
         for (int qy = 0; qy < Q1D; ++qy)
            for (int qx = 0; qx < Q1D; ++qx)
            {
               double u = 0.0, v = 0.0, w = 0.0;
               for (int dy = 0; dy < D1D; ++dy)
               {
                  u += s0[1][dz][dy][qx] * b_(qy, dy);
                  v += s0[0][dz][dy][qx] * g_(qy, dy);
                  w += s0[0][dz][dy][qx] * b_(qy, dy);
               }
               s1[0][dz][qy][qx] = u;
               s1[1][dz][qy][qx] = v;
               s1[2][dz][qy][qx] = w;
            }
      /* :1112-1147  Z contraction + D at the q-point */
      for (int qz = 0; qz < Q1D; ++qz)
         for (int qy = 0; qy < Q1D; ++qy)
            for (int qx = 0; qx < Q1D; ++qx)
            {
               double u = 0.0, v = 0.0, w = 0.0;
               for (int dz = 0; dz < D1D; ++dz)
               {
                  u += s1[0][dz][qy][qx] * b_(qz, dz);
                  v += s1[1][dz][qy][qx] * b_(qz, dz);
                  w += s1[2][dz][qy][qx] * g_(qz, dz);
               }
               const long q = qx + Q1D * (qy + (long)Q1D * qz);
               const double O11 = d[q + 0 * NQ], O12 = d[q + 1 * NQ], O13 = d[q + 2 * NQ];
               const double O22 = d[q + 3 * NQ], O23 = d[q + 4 * NQ], O33 = d[q + 5 * NQ];
               s0[0][qz][qy][qx] = (O11 * u) + (O12 * v) + (O13 * w);
               s0[1][qz][qy][qx] = (O12 * u) + (O22 * v) + (O23 * w);
               s0[2][qz][qy][qx] = (O13 * u) + (O23 * v) + (O33 * w);
            }
      /* :1160-1181  X^T */
      for (int qz = 0; qz < Q1D; ++qz)
         for (int qy = 0; qy < Q1D; ++qy)
            for (int dx = 0; dx < D1D; ++dx)
            {
               double u = 0.0, v = 0.0, w = 0.0;
               for (int qx = 0; qx < Q1D; ++qx)
               {
                  u += s0[0][qz][qy][qx] * g_(qx, dx);
                  v += s0[1][qz][qy][qx] * b_(qx, dx);
                  w += s0[2][qz][qy][qx] * b_(qx, dx);
               }
               s1[0][qz][qy][dx] = u;
               s1[1][qz][qy][dx] = v;
               s1[2][qz][qy][dx] = w;
            }
      /* :1183-1204  Y^T */
      for (int qz = 0; qz < Q1D; ++qz)
         for (int dy = 0; dy < D1D; ++dy)
            for (int dx = 0; dx < D1D; ++dx)
            {
               double u = 0.0, v = 0.0, w = 0.0;
               for (int qy = 0; qy < Q1D; ++qy)
               {
                  u += s1[0][qz][qy][dx] * b_(qy, dy);
                  v += s1[1][qz][qy][dx] * g_(qy, dy);
                  w += s1[2][qz][qy][dx] * b_(qy, dy);
               }
               s0[0][qz][dy][dx] = u;
               s0[1][qz][dy][dx] = v;
               s0[2][qz][dy][dx] = w;
            }
      /* :1206-1226  Z^T, y += (u + v + w) */
      for (int dz = 0; dz < D1D; ++dz)
         for (int dy = 0; dy < D1D; ++dy)
            for (int dx = 0; dx < D1D; ++dx)
            {
               double u = 0.0, v = 0.0, w = 0.0;
               for (int qz = 0; qz < Q1D; ++qz)
               {
                  u += s0[0][qz][dy][dx] * b_(qz, dz);
                  v += s0[1][qz][dy][dx] * b_(qz, dz);
                  w += s0[2][qz][dy][dx] * g_(qz, dz);
               }
               y[dx + D1D * (dy + D1D * dz)] += (u + v + w);
            }
   }
}

/* fem/integ/bilininteg_mass_kernels.hpp:807-1033 */
void orc_mass_apply(int NE, int D1D, int Q1D, const double *B, const double *vq, const double *xE, double *yE)
{
   const long NQ = (long)Q1D * Q1D * Q1D, ND = (long)D1D * D1D * D1D;
   static double s0[MX][MX][MX], s1[MX][MX][MX];
   for (long e = 0; e < NE; ++e)
   {
      const double *x = xE + e * ND;
      const double *d = vq + e * NQ;
      double *y = yE + e * ND;
      for (int dy = 0; dy < D1D; ++dy)
         for (int qx = 0; qx < Q1D; ++qx)
         {
            double u[MX];
            for (int dz = 0; dz < D1D; ++dz) { u[dz] = 0; }
            for (int dx = 0; dx < D1D; ++dx)
               for (int dz = 0; dz < D1D; ++dz) { u[dz] += x[dx + D1D * (dy + D1D * dz)] * b_(qx, dx); }
            for (int dz = 0; dz < D1D; ++dz) { s1[dz][dy][qx] = u[dz]; } /* DDQ */
         }
      for (int qy = 0; qy < Q1D; ++qy)
         for (int qx = 0; qx < Q1D; ++qx)
         {
            double u[MX];
            for (int dz = 0; dz < D1D; ++dz) { u[dz] = 0; }
            for (int dy = 0; dy < D1D; ++dy)
               for (int dz = 0; dz < D1D; ++dz) { u[dz] += s1[dz][dy][qx] * b_(qy, dy); }
            for (int dz = 0; dz < D1D; ++dz) { s0[dz][qy][qx] = u[dz]; } /* DQQ */
         }
      for (int qy = 0; qy < Q1D; ++qy)
         for (int qx = 0; qx < Q1D; ++qx)
         {
            double u[MX];
            for (int qz = 0; qz < Q1D; ++qz) { u[qz] = 0; }
            for (int dz = 0; dz < D1D; ++dz)
               for (int qz = 0; qz < Q1D; ++qz) { u[qz] += s0[dz][qy][qx] * b_(qz, dz); }
            for (int qz = 0; qz < Q1D; ++qz) { s1[qz][qy][qx] = u[qz] * d[qx + Q1D * (qy + (long)Q1D * qz)]; } /* QQQ */
         }
      for (int qy = 0; qy < Q1D; ++qy)
         for (int dx = 0; dx < D1D; ++dx)
         {
            double u[MX];
            for (int qz = 0; qz < Q1D; ++qz) { u[qz] = 0; }
            for (int qx = 0; qx < Q1D; ++qx)
               for (int qz = 0; qz < Q1D; ++qz) { u[qz] += s1[qz][qy][qx] * b_(qx, dx); }
            for (int qz = 0; qz < Q1D; ++qz) { s0[qz][qy][dx] = u[qz]; } /* QQD */
         }
      for (int dy = 0; dy < D1D; ++dy)
         for (int dx = 0; dx < D1D; ++dx)
         {
            double u[MX];
            for (int qz = 0; qz < Q1D; ++qz) { u[qz] = 0; }
            for (int qy = 0; qy < Q1D; ++qy)
               for (int qz = 0; qz < Q1D; ++qz) { u[qz] += s0[qz][qy][dx] * b_(qy, dy); }
            for (int qz = 0; qz < Q1D; ++qz) { s1[qz][dy][dx] = u[qz]; } /* QDD */
         }
      for (int dy = 0; dy < D1D; ++dy)
         for (int dx = 0; dx < D1D; ++dx)
         {
            double u[MX];
            for (int dz = 0; dz < D1D; ++dz) { u[dz] = 0; }
            for (int qz = 0; qz < Q1D; ++qz)
               for (int dz = 0; dz < D1D; ++dz) { u[dz] += s1[qz][dy][dx] * b_(qz, dz); }
            for (int dz = 0; dz < D1D; ++dz) { y[dx + D1D * (dy + D1D * dz)] += u[dz]; }
         }
   }
}

/* fem/integ/bilininteg_diffusion_kernels.hpp:369-484 (symmetric) */
void orc_diffusion_diag(int NE, int D1D, int Q1D, const double *B, const double *G, const double *D, double *dE)
{
   const long NQ = (long)Q1D * Q1D * Q1D, ND = (long)D1D * D1D * D1D;
   static double QQD[MX][MX][MX], QDD[MX][MX][MX];
   for (long e = 0; e < NE; ++e)
   {
      const double *d = D + e * 6 * NQ;
      double *y = dE + e * ND;
      for (int i = 0; i < 3; ++i)
         for (int j = 0; j < 3; ++j)
         {
            const int k = j >= i ? 3 - (3 - i) * (2 - i) / 2 + j : 3 - (3 - j) * (2 - j) / 2 + i;
            for (int qx = 0; qx < Q1D; ++qx)
               for (int qy = 0; qy < Q1D; ++qy)
                  for (int dz = 0; dz < D1D; ++dz)
                  {
                     QQD[qx][qy][dz] = 0.0;
                     for (int qz = 0; qz < Q1D; ++qz)
                     {
                        const long q = qx + (qy + (long)qz * Q1D) * Q1D;
                        const double O = d[q + k * NQ];
                        const double Bz = b_(qz, dz), Gz = g_(qz, dz);
                        const double L = i == 2 ? Gz : Bz, R = j == 2 ? Gz : Bz;
                        QQD[qx][qy][dz] += L * O * R;
                     }
                  }
            for (int qx = 0; qx < Q1D; ++qx)
               for (int dz = 0; dz < D1D; ++dz)
                  for (int dy = 0; dy < D1D; ++dy)
                  {
                     QDD[qx][dy][dz] = 0.0;
                     for (int qy = 0; qy < Q1D; ++qy)
                     {
                        const double By = b_(qy, dy), Gy = g_(qy, dy);
                        const double L = i == 1 ? Gy : By, R = j == 1 ? Gy : By;
                        QDD[qx][dy][dz] += L * QQD[qx][qy][dz] * R;
                     }
                  }
            for (int dz = 0; dz < D1D; ++dz)
               for (int dy = 0; dy < D1D; ++dy)
                  for (int dx = 0; dx < D1D; ++dx)
                     for (int qx = 0; qx < Q1D; ++qx)
                     {
                        const double Bx = b_(qx, dx), Gx = g_(qx, dx);
                        const double L = i == 0 ? Gx : Bx, R = j == 0 ? Gx : Bx;
                        y[dx + D1D * (dy + D1D * dz)] += L * QDD[qx][dy][dz] * R;
                     }
         }
   }
}

/* fem/integ/bilininteg_mass_kernels.hpp:324-408 */
void orc_mass_diag(int NE, int D1D, int Q1D, const double *B, const double *vq, double *dE)
{
   const long NQ = (long)Q1D * Q1D * Q1D, ND = (long)D1D * D1D * D1D;
   static double QQD[MX][MX][MX], QDD[MX][MX][MX];
   for (long e = 0; e < NE; ++e)
   {
      const double *d = vq + e * NQ;
      double *y = dE + e * ND;
      for (int qx = 0; qx < Q1D; ++qx)
         for (int qy = 0; qy < Q1D; ++qy)
            for (int dz = 0; dz < D1D; ++dz)
            {
               QQD[qx][qy][dz] = 0.0;
               for (int qz = 0; qz < Q1D; ++qz)
               {
                  QQD[qx][qy][dz] += b_(qz, dz) * b_(qz, dz) * d[qx + Q1D * (qy + (long)Q1D * qz)];
               }
            }
      for (int qx = 0; qx < Q1D; ++qx)
         for (int dz = 0; dz < D1D; ++dz)
            for (int dy = 0; dy < D1D; ++dy)
            {
               QDD[qx][dy][dz] = 0.0;
               for (int qy = 0; qy < Q1D; ++qy) { QDD[qx][dy][dz] += b_(qy, dy) * b_(qy, dy) * QQD[qx][qy][dz]; }
            }
      for (int dz = 0; dz < D1D; ++dz)
         for (int dy = 0; dy < D1D; ++dy)
            for (int dx = 0; dx < D1D; ++dx)
            {
               double t = 0.0;
               for (int qx = 0; qx < Q1D; ++qx) { t += b_(qx, dx) * b_(qx, dx) * QDD[qx][dy][dz]; }
               y[dx + D1D * (dy + D1D * dz)] += t;
            }
   }
}

/* --------------------------------------------------------------------- operator */
/* fem/bilinearform_ext.cpp:487-564 */
void orc_op_mult(const orc_operator *op, const double *x, double *y, double *workE)
{
   const long nE = (long)op->NE * op->D1D * op->D1D * op->D1D;
   double *xE = workE, *yE = workE + nE;
   orc_restrict_mult(op->NE, op->D1D * op->D1D * op->D1D, op->gather_map, x, xE);
   for (long i = 0; i < nE; ++i) { yE[i] = 0.0; }
   if (op->pa_diff) { orc_diffusion_apply(op->NE, op->D1D, op->Q1D, op->B, op->G, op->pa_diff, xE, yE); }
   if (op->pa_mass) { orc_mass_apply(op->NE, op->D1D, op->Q1D, op->B, op->pa_mass, xE, yE); }
   orc_restrict_mult_transpose(op->ndofs, op->offsets, op->indices, yE, y, 0);
}

/* fem/bilinearform_ext.cpp:370-454 */
void orc_op_diag(const orc_operator *op, double *diag, double *workE)
{
   const long nE = (long)op->NE * op->D1D * op->D1D * op->D1D;
   double *dE = workE;
   for (long i = 0; i < nE; ++i) { dE[i] = 0.0; }
   if (op->pa_diff) { orc_diffusion_diag(op->NE, op->D1D, op->Q1D, op->B, op->G, op->pa_diff, dE); }
   if (op->pa_mass) { orc_mass_diag(op->NE, op->D1D, op->Q1D, op->B, op->pa_mass, dE); }
   orc_restrict_mult_transpose(op->ndofs, op->offsets, op->indices, dE, diag, 1);
}

/* linalg/operator.cpp:586-646, DIAG_ONE */
void orc_constrained_mult(const orc_operator *op, const double *x, double *y, double *work, double *workE)
{
   if (op->n_ess == 0) { orc_op_mult(op, x, y, workE); return; }
   double *z = work;
   memcpy(z, x, sizeof(double) * op->ndofs);
   for (int i = 0; i < op->n_ess; ++i) { z[op->ess[i]] = 0.0; }
   orc_op_mult(op, z, y, workE);
   for (int i = 0; i < op->n_ess; ++i) { y[op->ess[i]] = x[op->ess[i]]; }
}

/* linalg/operator.cpp:559-584 */
void orc_eliminate_rhs(const orc_operator *op, const double *x, double *b, double *work2n, double *workE)
{
   double *w = work2n, *z = work2n + op->ndofs;
   for (int i = 0; i < op->ndofs; ++i) { w[i] = 0.0; }
   for (int i = 0; i < op->n_ess; ++i) { w[op->ess[i]] = x[op->ess[i]]; }
   orc_op_mult(op, w, z, workE);
   for (int i = 0; i < op->ndofs; ++i) { b[i] -= z[i]; }
   for (int i = 0; i < op->n_ess; ++i) { b[op->ess[i]] = x[op->ess[i]]; }
}

/* ----------------------------------------------------------------- Jacobi / PCG */
/* linalg/solvers.cpp:401-425; returns nonzero on a zero diagonal entry (:410-413 aborts) */
int orc_jacobi_setup(int n, const double *diag, int n_ess, const int *ess, double damping, double *dinv)
{
   for (int i = 0; i < n; ++i)
   {
      if (diag[i] == 0.0) { return 1; }
      dinv[i] = damping / diag[i];
   }
   for (int i = 0; i < n_ess; ++i) { dinv[ess[i]] = damping; }
   return 0;
}

/* linalg/solvers.cpp:442-452: residual = x; y = 0; y += dinv*r */
void orc_jacobi_mult(int n, const double *dinv, const double *r, double *z)
{
   for (int i = 0; i < n; ++i) { z[i] = 0.0; z[i] += dinv[i] * r[i]; }
}

double orc_dot(long n, const double *a, const double *b)
{
   double r = 0;
   for (long i = 0; i < n; ++i) { r += a[i] * b[i]; }
   return r;
}


/* ------------------------------------------------------- Chebyshev smoother, power method */
/* OperatorChebyshevSmoother::Setup, linalg/solvers.cpp:557-621 (the coefficient formulas verbatim in meaning;
 * dinv is the Jacobi one: 1/diag, 1 on essential dofs, :561-569).  Returns nonzero for an order outside 1..5. */
int orc_chebyshev_coeffs(int order, double max_eig, double *coeffs)
{
   const double upper_bound = 1.2 * max_eig, lower_bound = 0.3 * max_eig;
   const double theta = 0.5 * (upper_bound + lower_bound), delta = 0.5 * (upper_bound - lower_bound);
   switch (order - 1)
   {
      case 0: coeffs[0] = 1.0 / theta; break;
      case 1:
      {
         const double tmp_0 = 1.0 / (pow(delta, 2) - 2 * pow(theta, 2));
         coeffs[0] = -4 * theta * tmp_0;
         coeffs[1] = 2 * tmp_0;
         break;
      }
      case 2:
      {
         const double tmp_0 = 3 * pow(delta, 2), tmp_1 = pow(theta, 2);
         const double tmp_2 = 1.0 / (-4 * pow(theta, 3) + theta * tmp_0);
         coeffs[0] = tmp_2 * (tmp_0 - 12 * tmp_1);
         coeffs[1] = 12 / (tmp_0 - 4 * tmp_1);
         coeffs[2] = -4 * tmp_2;
         break;
      }
      case 3:
      {
         const double tmp_0 = pow(delta, 2), tmp_1 = pow(theta, 2), tmp_2 = 8 * tmp_0;
         const double tmp_3 = 1.0 / (pow(delta, 4) + 8 * pow(theta, 4) - tmp_1 * tmp_2);
         coeffs[0] = tmp_3 * (32 * pow(theta, 3) - 16 * theta * tmp_0);
         coeffs[1] = tmp_3 * (-48 * tmp_1 + tmp_2);
         coeffs[2] = 32 * theta * tmp_3;
         coeffs[3] = -8 * tmp_3;
         break;
      }
      case 4:
      {
         const double tmp_0 = 5 * pow(delta, 4), tmp_1 = pow(theta, 4), tmp_2 = pow(theta, 2), tmp_3 = pow(delta, 2);
         const double tmp_4 = 60 * tmp_3, tmp_5 = 20 * tmp_3;
         const double tmp_6 = 1.0 / (16 * pow(theta, 5) - pow(theta, 3) * tmp_5 + theta * tmp_0);
         const double tmp_7 = 160 * tmp_2;
         const double tmp_8 = 1.0 / (tmp_0 + 16 * tmp_1 - tmp_2 * tmp_5);
         coeffs[0] = tmp_6 * (tmp_0 + 80 * tmp_1 - tmp_2 * tmp_4);
         coeffs[1] = tmp_8 * (tmp_4 - tmp_7);
         coeffs[2] = tmp_6 * (-tmp_5 + tmp_7);
         coeffs[3] = -80 * tmp_8;
         coeffs[4] = 16 * tmp_6;
         break;
      }
      default: return 1;
   }
   return 0;
}

/* OperatorChebyshevSmoother::Mult, linalg/solvers.cpp:623-657, on the constrained operator */
void orc_chebyshev_mult(const orc_operator *op, const double *dinv, int order, const double *coeffs, const double *x,
                        double *y, double *work, double *workE)
{
   const int n = op->ndofs;
   double *residual = malloc(sizeof(double) * n), *helper = malloc(sizeof(double) * n);
   memcpy(residual, x, sizeof(double) * n);
   for (int i = 0; i < n; ++i) { y[i] = 0.0; }
   for (int k = 0; k < order; ++k)
   {
      if (k > 0)
      {
         orc_constrained_mult(op, residual, helper, work, workE);
         memcpy(residual, helper, sizeof(double) * n);
      }
      for (int i = 0; i < n; ++i) { residual[i] *= dinv[i]; }
      for (int i = 0; i < n; ++i) { y[i] += coeffs[k] * residual[i]; }
   }
   free(residual); free(helper);
}

/* PowerMethod::EstimateLargestEigenvalue, linalg/operator.cpp:871-928, for opr = Dinv * A (the ProductOperator of
 * linalg/solvers.cpp:500-501).  v0: start vector on entry (the caller's Vector::Randomize(seed)), work space after. */
double orc_power_method(const orc_operator *op, const double *dinv, double *v0, int num_steps, double tolerance)
{
   const int n = op->ndofs;
   const long nE = (long)op->NE * op->D1D * op->D1D * op->D1D;
   double *v1 = malloc(sizeof(double) * n), *t = malloc(sizeof(double) * n);
   double *work = malloc(sizeof(double) * n), *workE = malloc(sizeof(double) * 2 * nE);
   double *a = v0, *b = v1;
   double eigenvalue = 1.0;
   for (int iter = 0; iter < num_steps; ++iter)
   {
      const double normV0 = orc_dot(n, a, a);
      const double s = sqrt(normV0);
      for (int i = 0; i < n; ++i) { a[i] /= s; }
      orc_constrained_mult(op, a, t, work, workE);
      orc_jacobi_mult(n, dinv, t, b);
      const double eigenvalueNew = orc_dot(n, a, b);
      const double diff = fabs((eigenvalueNew - eigenvalue) / eigenvalue);
      eigenvalue = eigenvalueNew;
      { double *sw = a; a = b; b = sw; }
      if (diff < tolerance) { break; }
   }
   if (a != v0) { memcpy(v0, a, sizeof(double) * n); }
   free(v1); free(t); free(work); free(workE);
   return eigenvalue;
}

/* linalg/solvers.cpp:869-1050; preconditioner = Jacobi (cheb_order = 0) or the Chebyshev smoother of that order */
static void orc_precond(const orc_operator *op, const double *dinv, int cheb_order, const double *coeffs, const double *r,
                        double *z, double *work, double *workE)
{
   if (cheb_order > 0) { orc_chebyshev_mult(op, dinv, cheb_order, coeffs, r, z, work, workE); }
   else { orc_jacobi_mult(op->ndofs, dinv, r, z); }
}

int orc_pcg_prec(const orc_operator *op, const double *dinv, int cheb_order, double max_eig, const double *b, double *x,
                 double rel_tol, double abs_tol, int max_iter, int *converged, double *final_norm, double *norms);

int orc_pcg(const orc_operator *op, const double *dinv, const double *b, double *x,
            double rel_tol, double abs_tol, int max_iter, int *converged, double *final_norm,
            double *norms)
{
   return orc_pcg_prec(op, dinv, 0, 0.0, b, x, rel_tol, abs_tol, max_iter, converged, final_norm, norms);
}

int orc_pcg_prec(const orc_operator *op, const double *dinv, int cheb_order, double max_eig, const double *b, double *x,
                 double rel_tol, double abs_tol, int max_iter, int *converged, double *final_norm, double *norms)
{
   double coeffs[5] = {0, 0, 0, 0, 0};
   if (cheb_order > 0 && orc_chebyshev_coeffs(cheb_order, max_eig, coeffs)) { *converged = 0; *final_norm = -1.0; return -1; }
   const int n = op->ndofs;
   const long nE = (long)op->NE * op->D1D * op->D1D * op->D1D;
   double *r = malloc(sizeof(double) * n), *d = malloc(sizeof(double) * n), *z = malloc(sizeof(double) * n);
   double *work = malloc(sizeof(double) * n), *workE = malloc(sizeof(double) * 2 * nE);
   double r0, den, nom, betanom = 0.0, alpha, beta;
   int i, final_iter;
   /* iterative_mode: r = b - A x  (:875-879; subtract(b, r, r)) */
   orc_constrained_mult(op, x, r, work, workE);
   for (int k = 0; k < n; ++k) { r[k] = b[k] - r[k]; }
   orc_precond(op, dinv, cheb_order, coeffs, r, z, work, workE);
   memcpy(d, z, sizeof(double) * n);
   nom = orc_dot(n, d, r);
   if (norms) { norms[0] = nom; }
   *converged = 0;
   if (nom < 0.0) { *final_norm = nom; final_iter = 0; goto done; }
   r0 = fmax(nom * rel_tol * rel_tol, abs_tol * abs_tol);
   if (nom <= r0) { *converged = 1; final_iter = 0; *final_norm = sqrt(nom); goto done; }
   orc_constrained_mult(op, d, z, work, workE);
   den = orc_dot(n, z, d);
   if (den <= 0.0 && den == 0.0) { final_iter = 0; *final_norm = sqrt(nom); goto done; }
   final_iter = max_iter;
   for (i = 1; 1;)
   {
      alpha = nom / den;
      for (int k = 0; k < n; ++k) { x[k] = x[k] + alpha * d[k]; }
      { const double ma = -alpha; for (int k = 0; k < n; ++k) { r[k] = r[k] + ma * z[k]; } }
      orc_precond(op, dinv, cheb_order, coeffs, r, z, work, workE);
      betanom = orc_dot(n, r, z);
      if (norms) { norms[i] = betanom; }
      if (betanom < 0.0) { final_iter = i; break; }
      if (betanom <= r0) { *converged = 1; final_iter = i; break; }
      if (++i > max_iter) { break; }
      beta = betanom / nom;
      /* add(z, beta, d, d): alpha==0 → copy, ==1 → plain add (linalg/vector.cpp:441-448) */
      if (beta == 0.0) { memcpy(d, z, sizeof(double) * n); }
      else if (beta == 1.0) { for (int k = 0; k < n; ++k) { d[k] = z[k] + d[k]; } }
      else { for (int k = 0; k < n; ++k) { d[k] = z[k] + beta * d[k]; } }
      orc_constrained_mult(op, d, z, work, workE);
      den = orc_dot(n, d, z);
      if (den <= 0.0 && den == 0.0) { final_iter = i; break; }
      nom = betanom;
   }
   *final_norm = sqrt(betanom);
done:
   free(r); free(d); free(z); free(work); free(workE);
   return final_iter;
}

/* ------------------------------------------------------------- q-point operators */
/* fem/qinterp/eval.hpp:131-193 with EvalX/EvalY/EvalZ (x, then y, then z) */
void orc_qvalues(int NE, int D1D, int Q1D, const double *B, const double *xE, double *yq)
{
   const long NQ = (long)Q1D * Q1D * Q1D, ND = (long)D1D * D1D * D1D;
   static double DDQ[MX][MX][MX], DQQ[MX][MX][MX];
   for (long e = 0; e < NE; ++e)
   {
      const double *x = xE + e * ND;
      double *y = yq + e * NQ;
      for (int dz = 0; dz < D1D; ++dz)
         for (int dy = 0; dy < D1D; ++dy)
            for (int qx = 0; qx < Q1D; ++qx)
            {
               double u = 0.0;
               for (int dx = 0; dx < D1D; ++dx) { u += b_(qx, dx) * x[dx + D1D * (dy + D1D * dz)]; }
               DDQ[dz][dy][qx] = u;
            }
      for (int dz = 0; dz < D1D; ++dz)
         for (int qy = 0; qy < Q1D; ++qy)
            for (int qx = 0; qx < Q1D; ++qx)
            {
               double u = 0.0;
               for (int dy = 0; dy < D1D; ++dy) { u += DDQ[dz][dy][qx] * b_(qy, dy); }
               DQQ[dz][qy][qx] = u;
            }
      for (int qz = 0; qz < Q1D; ++qz)
         for (int qy = 0; qy < Q1D; ++qy)
            for (int qx = 0; qx < Q1D; ++qx)
            {
               double u = 0.0;
               for (int dz = 0; dz < D1D; ++dz) { u += DQQ[dz][qy][qx] * b_(qz, dz); }
               y[qx + Q1D * (qy + (long)Q1D * qz)] = u;
            }
   }
}

/* fem/qinterp/grad.hpp:233-374 ; inverse as linalg/kernels.hpp:306-311 (adjugate * (1/det)) */
void orc_qphysgrad(int NE, int D1D, int Q1D, const double *B, const double *G, const double *J,
                   const double *xE, double *gq)
{
   const long NQ = (long)Q1D * Q1D * Q1D, ND = (long)D1D * D1D * D1D;
   static double s0[2][MX][MX][MX], s1[3][MX][MX][MX];
   for (long e = 0; e < NE; ++e)
   {
      const double *x = xE + e * ND;
      for (int dz = 0; dz < D1D; ++dz)
         for (int dy = 0; dy < D1D; ++dy)
            for (int qx = 0; qx < Q1D; ++qx)
            {
               double u = 0.0, v = 0.0;
               for (int dx = 0; dx < D1D; ++dx)
               {
                  const double in = x[dx + D1D * (dy + D1D * dz)];
                  u += in * b_(qx, dx);
                  v += in * g_(qx, dx);
               }
               s0[0][dz][dy][qx] = u;
               s0[1][dz][dy][qx] = v;
            }
      for (int dz = 0; dz < D1D; ++dz)
         for (int qy = 0; qy < Q1D; ++qy)
            for (int qx = 0; qx < Q1D; ++qx)
            {
               double u = 0.0, v = 0.0, w = 0.0;
               for (int dy = 0; dy < D1D; ++dy)
               {
                  u += s0[1][dz][dy][qx] * b_(qy, dy);
                  v += s0[0][dz][dy][qx] * g_(qy, dy);
                  w += s0[0][dz][dy][qx] * b_(qy, dy);
               }
               s1[0][dz][qy][qx] = u;
               s1[1][dz][qy][qx] = v;
               s1[2][dz][qy][qx] = w;
            }
      for (int qz = 0; qz < Q1D; ++qz)
         for (int qy = 0; qy < Q1D; ++qy)
            for (int qx = 0; qx < Q1D; ++qx)
            {
               double u = 0.0, v = 0.0, w = 0.0;
               for (int dz = 0; dz < D1D; ++dz)
               {
                  u += s1[0][dz][qy][qx] * b_(qz, dz);
                  v += s1[1][dz][qy][qx] * b_(qz, dz);
                  w += s1[2][dz][qy][qx] * g_(qz, dz);
               }
               const long q = qx + Q1D * (qy + (long)Q1D * qz);
               const double *Je = J + e * 9 * NQ + q;
               /* column-major 3x3: a[row + 3*col] */
               double a[9], inv[9];
               for (int k = 0; k < 9; ++k) { a[k] = Je[k * NQ]; }
               /* adjugate and determinant as linalg/tmatrix.hpp TAdjDetHD (3x3 specialisation) */
               inv[0] = a[4] * a[8] - a[5] * a[7];
               inv[1] = a[2] * a[7] - a[1] * a[8];
               inv[2] = a[1] * a[5] - a[2] * a[4];
               inv[3] = a[5] * a[6] - a[3] * a[8];
               inv[4] = a[0] * a[8] - a[2] * a[6];
               inv[5] = a[2] * a[3] - a[0] * a[5];
               inv[6] = a[3] * a[7] - a[4] * a[6];
               inv[7] = a[1] * a[6] - a[0] * a[7];
               inv[8] = a[0] * a[4] - a[1] * a[3];
               const double det = a[0] * inv[0] + a[1] * inv[3] + a[2] * inv[6];
               const double idet = 1.0 / det;
               for (int k = 0; k < 9; ++k) { inv[k] *= idet; }
               const double U = inv[0] * u + inv[1] * v + inv[2] * w;
               const double V = inv[3] * u + inv[4] * v + inv[5] * w;
               const double Wv = inv[6] * u + inv[7] * v + inv[8] * w;
               double *g = gq + 3 * (e * NQ + q);
               g[0] = U; g[1] = V; g[2] = Wv;
            }
   }
}

/* fem/integ/lininteg_domain_kernels.hpp:164-298 */
void orc_domain_lf(int NE, int D1D, int Q1D, const double *B, const double *detJ, const double *W,
                   const double *f, long nf, double *bE)
{
   const long NQ = (long)Q1D * Q1D * Q1D, ND = (long)D1D * D1D * D1D;
   const int cst = (nf == 1);
   static double QQQ[MX][MX][MX];
   for (long e = 0; e < NE; ++e)
   {
      double *y = bE + e * ND;
      for (int x = 0; x < Q1D; ++x)
         for (int yy = 0; yy < Q1D; ++yy)
            for (int z = 0; z < Q1D; ++z)
            {
               const long q = x + Q1D * (yy + (long)Q1D * z);
               const double c = cst ? f[0] : f[e * NQ + q];
               QQQ[z][yy][x] = W[q] * c * detJ[e * NQ + q];
            }
      for (int qx = 0; qx < Q1D; ++qx)
         for (int qy = 0; qy < Q1D; ++qy)
         {
            double u[MX];
            for (int dz = 0; dz < D1D; ++dz) { u[dz] = 0.0; }
            for (int qz = 0; qz < Q1D; ++qz)
            {
               const double ZYX = QQQ[qz][qy][qx];
               for (int dz = 0; dz < D1D; ++dz) { u[dz] += ZYX * b_(qz, dz); }
            }
            for (int dz = 0; dz < D1D; ++dz) { QQQ[dz][qy][qx] = u[dz]; }
         }
      for (int dz = 0; dz < D1D; ++dz)
         for (int qx = 0; qx < Q1D; ++qx)
         {
            double u[MX];
            for (int dy = 0; dy < D1D; ++dy) { u[dy] = 0.0; }
            for (int qy = 0; qy < Q1D; ++qy)
            {
               const double zYX = QQQ[dz][qy][qx];
               for (int dy = 0; dy < D1D; ++dy) { u[dy] += zYX * b_(qy, dy); }
            }
            for (int dy = 0; dy < D1D; ++dy) { QQQ[dz][dy][qx] = u[dy]; }
         }
      for (int dz = 0; dz < D1D; ++dz)
         for (int dy = 0; dy < D1D; ++dy)
         {
            double u[MX];
            for (int dx = 0; dx < D1D; ++dx) { u[dx] = 0.0; }
            for (int qx = 0; qx < Q1D; ++qx)
            {
               const double zyX = QQQ[dz][dy][qx];
               for (int dx = 0; dx < D1D; ++dx) { u[dx] += zyX * b_(qx, dx); }
            }
            for (int dx = 0; dx < D1D; ++dx) { y[dx + D1D * (dy + D1D * dz)] += u[dx]; }
         }
   }
}
