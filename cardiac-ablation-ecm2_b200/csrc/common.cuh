// Shared declarations for libb200pa.so (context, error handling, device-buffer helpers).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/b200pa.h"

namespace b200pa
{

extern thread_local std::string g_err;
extern std::atomic<long long> g_launches;

inline int fail(const std::string &msg)
{
   g_err = msg;
   return 1;
}

#define B200PA_CK(call)                                                                          \
   do                                                                                            \
   {                                                                                             \
      cudaError_t e_ = (call);                                                                   \
      if (e_ != cudaSuccess)                                                                     \
      {                                                                                          \
         return ::b200pa::fail(std::string(#call) + ": " + cudaGetErrorString(e_) + " (" +       \
                               __FILE__ + ":" + std::to_string(__LINE__) + ")");                 \
      }                                                                                          \
   } while (0)

#define B200PA_REQUIRE(cond, msg)                                                                \
   do                                                                                            \
   {                                                                                             \
      if (!(cond)) { return ::b200pa::fail(std::string("b200pa: ") + (msg)); }                   \
   } while (0)

// after every launch (≙ MFEM_GPU_CHECK(cudaGetLastError()), general/forall.hpp:620)
#define B200PA_LAUNCHED()                                                                        \
   do                                                                                            \
   {                                                                                             \
      ::b200pa::g_launches.fetch_add(1, std::memory_order_relaxed);                              \
      B200PA_CK(cudaGetLastError());                                                             \
   } while (0)

// device buffer that is either owned (allocated here) or borrowed (caller's device pointer)
struct DevBuf
{
   void *p = nullptr;
   size_t bytes = 0;
   bool owned = false;
   void release()
   {
      if (owned && p) { cudaFree(p); }
      p = nullptr; bytes = 0; owned = false;
   }
   template <typename T> T *as() const { return static_cast<T *>(p); }
};

} // namespace b200pa

struct b200pa_ctx_s
{
   int device = 0;
   cudaStream_t stream = nullptr;
   bool own_stream = false;
   int num_sms = 148;
   // reduction scratch: block partials + "last block" tickets + result slots (device), pinned mirror
   double *d_partials = nullptr; // [MAX_RED_BLOCKS * 2]
   unsigned int *d_ticket = nullptr;
   double *d_result = nullptr;   // [8]
   double *h_result = nullptr;   // pinned [8]
   cudaEvent_t ev_poll[2] = {nullptr, nullptr}; // PCG convergence-flag read-backs in flight (DonePoller)
};

namespace b200pa
{
constexpr int MAX_RED_BLOCKS = 2048;
constexpr int MAX_DEVICES = 64; // per-device caches of kernel attributes / occupancy

// Returns a device pointer for `src` (host or device).  Host data is copied into `buf`.
int to_device(b200pa_ctx ctx, const void *src, size_t bytes, DevBuf &buf, const void **out);
bool is_device_ptr(const void *p);
int alloc(DevBuf &buf, size_t bytes);
} // namespace b200pa
