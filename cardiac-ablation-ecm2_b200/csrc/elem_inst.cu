// One (D1D,Q1D) instantiation set of the element kernel.  Compiled six times:
//   nvcc ... -DB200PA_D=3 -DB200PA_Q=4 -c elem_inst.cu -o elem_3_4.o
#include <atomic>

#include "common.cuh"
#include "elem_launch.cuh"
#include "pa_element_kernel.cuh"
#include "pa_apply_kernel.cuh"
#if B200PA_D == 7
#include "pa_apply_dmma.cuh"
#endif
#include <cstdlib>

#ifndef B200PA_D
#error "compile with -DB200PA_D=<D1D> -DB200PA_Q=<Q1D>"
#endif

namespace b200pa
{

namespace
{
constexpr int D = B200PA_D, Q = B200PA_Q;

// Dynamic shared-memory opt-in + resident CTAs per SM of one kernel instantiation.  Both are PER DEVICE (a process
// may hold contexts on several GPUs), so they are cached per (instantiation, device); the cache entries are atomics and a
// set-once race between threads only repeats the same two idempotent runtime calls.
// Tag: a type unique to the kernel instantiation (kernels of one signature share the function-pointer type K).
template <typename Tag, typename K>
int kernel_setup(K kern, int threads, size_t smem, int *blocks_per_sm)
{
   static std::atomic<int> cache[MAX_DEVICES];
   int dev = 0;
   cudaError_t e = cudaGetDevice(&dev);
   if (e != cudaSuccess) { return (int)e; }
   const bool cached = dev >= 0 && dev < MAX_DEVICES;
   int n = cached ? cache[dev].load(std::memory_order_acquire) : 0;
   if (n == 0)
   {
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) { return (int)e; }
      e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, threads, smem);
      if (e != cudaSuccess) { return (int)e; }
      n = n > 0 ? n : 1;
      if (cached) { cache[dev].store(n, std::memory_order_release); }
   }
   *blocks_per_sm = n;
   return 0;
}

template <int DD, int QQ>
void fill_params(ElemParams<DD, QQ> &P, const ElemArgs &a)
{
   for (int i = 0; i < QQ * DD; ++i) { P.bg.B[i] = a.B[i]; P.bg.G[i] = a.G ? a.G[i] : 0.0; }
   P.NE = a.NE;
   P.x = a.x; P.gmap = a.gmap; P.y = a.y; P.slot = a.slot;
   P.pa_diff = a.pa_diff; P.pa_mass = a.pa_mass; P.geo = a.geo; P.J = a.J;
   P.f = a.f; P.detJ = a.detJ; P.detE = a.detE; P.W = a.W; P.nf = a.nf; P.done = a.done;
   P.ca = a.ca; P.cb = a.cb; P.cT0 = a.cT0; P.s = a.s;
   P.vtx = a.vtx; P.ev = a.ev; P.jinv = a.jinv;
   for (int i = 0; i < QQ; ++i) { P.xi[i] = a.xi ? a.xi[i] : 0.0; }
}

template <bool DIFF, bool MASS, int INMODE, int OUTMODE, int QOP>
int run(const ElemArgs &a, int num_sms, cudaStream_t stream)
{
   using L = ElemLayout<D, Q>;
   auto kern = pa_element_kernel<D, Q, DIFF, MASS, INMODE, OUTMODE, QOP>;
   int blocks_per_sm = 0;
   {
      struct ThisKernel {}; // local to this instantiation of run<>
      const int e = kernel_setup<ThisKernel>(kern, L::NT, L::SMEM_BYTES, &blocks_per_sm);
      if (e) { return e; }
   }
   ElemParams<D, Q> P;
   fill_params(P, a);
   if (a.NE <= 0) { return 0; }
   const int nbatch = (a.NE + L::NEB - 1) / L::NEB;
   const int grid = nbatch < num_sms * blocks_per_sm ? nbatch : num_sms * blocks_per_sm;
   kern<<<grid, L::NT, L::SMEM_BYTES, stream>>>(P);
   return (int)cudaGetLastError();
}

// the hot path: gather -> diffusion (+ mass) -> slot write (pa_apply_kernel.cuh)
template <bool DIFF, bool MASS, bool AFF>
int run_fused(const ElemArgs &a, int num_sms, cudaStream_t stream)
{
   using C = ApplyCfg<D, Q, AFF>;
   auto kern = pa_apply_kernel<D, Q, DIFF, MASS, AFF>;
   int blocks_per_sm = 0;
   {
      struct ThisKernel {}; // local to this instantiation of run_fused<>
      const int e = kernel_setup<ThisKernel>(kern, C::NT, C::SMEM_BYTES, &blocks_per_sm);
      if (e) { return e; }
   }
   if (a.NE <= 0) { return 0; }
   // TMA bulk copies need 16-byte aligned sources (cudaMalloc gives 256)
   if (((((unsigned long long)a.pa_diff) | ((unsigned long long)a.pa_mass)) & 15ull)) { return (int)cudaErrorMisalignedAddress; }
   ElemParams<D, Q> P;
   fill_params(P, a);
   const int nbatch = (a.NE + C::NEB - 1) / C::NEB;
   const int grid = nbatch < num_sms * blocks_per_sm ? nbatch : num_sms * blocks_per_sm;
   kern<<<grid, C::NT, C::SMEM_BYTES, stream>>>(P);
   return (int)cudaGetLastError();
}

#if B200PA_D == 7
// order 6, stored q-data: the variant with the row phases on the FP64 tensor cores (pa_apply_dmma.cuh).  B200PA_DMMA=0 in the
// environment selects the DFMA kernel instead (read once; for A/B measurements)
template <bool DIFF, bool MASS>
int run_fused_dmma(const ElemArgs &a, int num_sms, cudaStream_t stream)
{
   auto kern = pa_apply_dmma_kernel<DIFF, MASS>;
   int blocks_per_sm = 0;
   {
      struct ThisKernel {};
      const int e = kernel_setup<ThisKernel>(kern, DmmaCfg::NT, DmmaCfg::SMEM_BYTES, &blocks_per_sm);
      if (e) { return e; }
   }
   if (a.NE <= 0) { return 0; }
   if (((((unsigned long long)a.pa_diff) | ((unsigned long long)a.pa_mass)) & 15ull)) { return (int)cudaErrorMisalignedAddress; }
   ElemParams<D, Q> P;
   fill_params(P, a);
   const int grid = a.NE < num_sms * blocks_per_sm ? a.NE : num_sms * blocks_per_sm;
   kern<<<grid, DmmaCfg::NT, DmmaCfg::SMEM_BYTES, stream>>>(P);
   return (int)cudaGetLastError();
}
static bool use_dmma()
{
   static const bool on = [] { const char *e = getenv("B200PA_DMMA"); return !(e && e[0] == '0'); }();
   return on;
}
#endif

int run_apply_fused(const ElemArgs &a, int num_sms, cudaStream_t stream)
{
#if B200PA_D == 7
   if (!a.geo && use_dmma())
   {
      if (a.pa_diff && a.pa_mass) { return run_fused_dmma<true, true>(a, num_sms, stream); }
      if (a.pa_diff) { return run_fused_dmma<true, false>(a, num_sms, stream); }
      if (a.pa_mass) { return run_fused_dmma<false, true>(a, num_sms, stream); }
   }
#endif
   if (a.pa_diff && a.geo) { return a.pa_mass ? run_fused<true, true, true>(a, num_sms, stream) : run_fused<true, false, true>(a, num_sms, stream); }
   if (a.pa_diff && a.pa_mass) { return run_fused<true, true, false>(a, num_sms, stream); }
   if (a.pa_diff) { return run_fused<true, false, false>(a, num_sms, stream); }
   if (a.pa_mass) { return run_fused<false, true, false>(a, num_sms, stream); }
   return (int)cudaErrorInvalidValue;
}

template <int INMODE, int OUTMODE>
int run_apply(const ElemArgs &a, int num_sms, cudaStream_t stream)
{
   if (a.pa_diff && a.pa_mass) { return run<true, true, INMODE, OUTMODE, QOP_APPLY>(a, num_sms, stream); }
   if (a.pa_diff) { return run<true, false, INMODE, OUTMODE, QOP_APPLY>(a, num_sms, stream); }
   if (a.pa_mass) { return run<false, true, INMODE, OUTMODE, QOP_APPLY>(a, num_sms, stream); }
   return (int)cudaErrorInvalidValue;
}
} // namespace

#define B200PA_CAT3(a, b, c) a##b##_##c
#define B200PA_NAME(D_, Q_) B200PA_CAT3(launch_element_, D_, Q_)

int B200PA_NAME(B200PA_D, B200PA_Q)(int variant, const ElemArgs &a, int num_sms, cudaStream_t stream)
{
   switch (variant)
   {
      case EV_APPLY_E: return run_apply<IN_E, OUT_E_ADD>(a, num_sms, stream);
      case EV_APPLY_L2S: return run_apply_fused(a, num_sms, stream);
      case EV_VALUES_E: return run<false, false, IN_E, OUT_NONE, QOP_VALUES>(a, num_sms, stream);
      case EV_VALUES_L: return run<false, false, IN_GATHER, OUT_NONE, QOP_VALUES>(a, num_sms, stream);
      case EV_PHYSGRAD_E: return run<false, false, IN_E, OUT_NONE, QOP_PHYSGRAD>(a, num_sms, stream);
      case EV_PHYSGRAD_L: return run<false, false, IN_GATHER, OUT_NONE, QOP_PHYSGRAD>(a, num_sms, stream);
      case EV_LF_E: return run<false, false, IN_NONE, OUT_E_ADD, QOP_LF>(a, num_sms, stream);
      case EV_LF_S: return run<false, false, IN_NONE, OUT_SLOT, QOP_LF>(a, num_sms, stream);
      case EV_COEFF_L: return run<false, false, IN_GATHER, OUT_NONE, QOP_COEFF>(a, num_sms, stream);
      case EV_JOULE_L: return run<false, false, IN_GATHER, OUT_NONE, QOP_JOULE>(a, num_sms, stream);
      default: return (int)cudaErrorInvalidValue;
   }
}

} // namespace b200pa
