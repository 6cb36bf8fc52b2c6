// Untemplated launch interface of the sum-factorised element kernel.  One translation unit
// per (D1D,Q1D) (elem_inst.cu compiled with -DB200PA_D/-DB200PA_Q) keeps build time parallel.
#pragma once
#include <cuda_runtime.h>

namespace b200pa
{

// which specialisation to run (see pa_element_kernel.cuh for the mode enums)
enum ElemVariant
{
   EV_APPLY_E = 0,      // x_E -> y_E += (diffusion and/or mass)           K4 / K7
   EV_APPLY_L2S,        // x_L -(gather)-> ... -> y_S[slot]                K1+K16+K4+K7+(write half of K2)
   EV_VALUES_E,         // x_E -> values at q-points                        K13
   EV_VALUES_L,         // x_L -(gather)-> values at q-points               K1+K13
   EV_PHYSGRAD_E,       // x_E -> physical gradient at q-points             K14
   EV_PHYSGRAD_L,       // x_L -(gather)-> physical gradient at q-points    K1+K14
   EV_LF_E,             // f_q -> b_E +=                                    K15
   EV_LF_S,             // f_q -> b_S[slot]                                 K15 + (write half of K2)
   // fused q-point physics: no q-data ever stored for T_q / grad phi
   EV_COEFF_L,          // T_L -(gather)-> T_q -> out_q = a (1 + b (T_q - T0))
   EV_JOULE_L,          // phi_L -(gather)-> grad phi_q -> out_q = s_q |grad phi|^2 + a
};

struct ElemArgs
{
   const double *B = nullptr, *G = nullptr; // HOST pointers, column-major [Q,D]
   int NE = 0;
   const double *x = nullptr;
   const int *gmap = nullptr;
   double *y = nullptr;
   const int *slot = nullptr;
   const double *pa_diff = nullptr, *pa_mass = nullptr;
   const double *geo = nullptr;             // EV_APPLY_L2S: factorised diffusion q-data (pa_diff = c_q [Q^3,NE], geo [6,NE])
   const double *J = nullptr;
   const double *f = nullptr, *detJ = nullptr, *detE = nullptr, *W = nullptr; // detE: one determinant per element (detJ == nullptr)
   long long nf = 0;
   const int *done = nullptr;
   double ca = 0.0, cb = 0.0, cT0 = 0.0;    // EV_COEFF_L / EV_JOULE_L parameters
   const double *s = nullptr;               // EV_JOULE_L: sigma_q
   const double *jinv = nullptr;            // affine meshes: rows of J^{-T} per element [9,NE] (wins over J / vtx)
   const double *vtx = nullptr;             // J == nullptr: trilinear geometry from vertices
   const int *ev = nullptr;
   const double *xi = nullptr;              // HOST pointer, Q values
};

// returns cudaError_t as int; `num_sms` sizes the persistent grid
int launch_element(int d1d, int q1d, int variant, const ElemArgs &a, int num_sms, cudaStream_t stream);

#define B200PA_DECL_ELEM(D, Q) int launch_element_##D##_##Q(int variant, const ElemArgs &a, int num_sms, cudaStream_t stream);
B200PA_DECL_ELEM(2, 3)
B200PA_DECL_ELEM(3, 4)
B200PA_DECL_ELEM(4, 5)
B200PA_DECL_ELEM(5, 6)
B200PA_DECL_ELEM(6, 7)
B200PA_DECL_ELEM(7, 8)
#undef B200PA_DECL_ELEM

} // namespace b200pa
