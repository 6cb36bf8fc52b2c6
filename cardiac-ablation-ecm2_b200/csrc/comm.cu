// Multi-GPU plumbing: one rank per GPU, NCCL over NVLink 5 / NVSwitch.
//
// What the reference does (fem/pfespace.cpp:5394-5532, general/communication.cpp:723-1120):
// per operator apply, P (owner -> sharers, "Bcast") before the element work and P^T (sharers ->
// owner, "Reduce") after it, i.e. two neighbour message rounds, plus an MPI_Allreduce per dot.
//
// What this does instead: PCG runs on *consistent L-vectors* (every ghost copy of a shared dof
// holds the owner's value), so P is never needed inside the loop, and P^T followed by the next
// P collapses into ONE symmetric exchange: every rank sends its partial sums of the dofs it shares
// with neighbour k to k, receives k's, and each rank adds the contributions of a dof in ascending
// rank order (its own included).  The result is bit-identical on all sharers and independent of
// message arrival order, which keeps run-to-run reproducibility (SURVEY §8e "determinism note").
// Dots count every dof once through the owner mask (owner = lowest sharing rank) and are combined
// with a 1-double ncclAllReduce.
//
// NCCL is loaded with dlopen so the library has no link-time dependency on a particular
// libnccl (the host application - e.g. torch - may already have one mapped).
#include <dlfcn.h>

#include <algorithm>
#include <map>

#include "comm.cuh"
#include "pcg_state.cuh"
#include "reduce.cuh"

namespace
{
// the slice of nccl.h this file uses (ABI-stable since NCCL 2.x)
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclSuccess = 0 };
enum { ncclFloat64 = 8 };
enum { ncclSum = 0 };

struct NcclApi
{
   void *h = nullptr;
   ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
   ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
   ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
   ncclResult_t (*GroupStart)() = nullptr;
   ncclResult_t (*GroupEnd)() = nullptr;
   ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
   ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
   ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
   const char *(*GetErrorString)(ncclResult_t) = nullptr;
   std::string err;
   bool load()
   {
      if (h) { return true; }
      const char *names[] = {getenv("B200PA_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
      for (const char *n : names)
      {
         if (!n) { continue; }
         h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
         if (h) { break; }
      }
      if (!h) { err = "cannot dlopen libnccl.so.2 (set B200PA_NCCL_LIB)"; return false; }
#define SYM(field, name)                                              \
   field = (decltype(field))dlsym(h, name);                           \
   if (!field) { err = std::string("missing NCCL symbol ") + name; h = nullptr; return false; }
      SYM(GetUniqueId, "ncclGetUniqueId")
      SYM(CommInitRank, "ncclCommInitRank")
      SYM(CommDestroy, "ncclCommDestroy")
      SYM(GroupStart, "ncclGroupStart")
      SYM(GroupEnd, "ncclGroupEnd")
      SYM(Send, "ncclSend")
      SYM(Recv, "ncclRecv")
      SYM(AllReduce, "ncclAllReduce")
      SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
      return true;
   }
};
NcclApi g_nccl;

#define NCCL_CK(call)                                                                                   \
   do                                                                                                   \
   {                                                                                                    \
      ncclResult_t r_ = (call);                                                                         \
      if (r_ != ncclSuccess) { return ::b200pa::fail(std::string(#call) + ": " + g_nccl.GetErrorString(r_)); } \
   } while (0)

__global__ void k_pack(int n, const int *__restrict__ ldof, const double *__restrict__ y, double *__restrict__ buf,
                       const int *done)
{
   if (done && *done) { return; }
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { buf[i] = y[ldof[i]]; }
}

// one thread per shared L-dof; sources in ascending rank order: src < 0 -> this rank's own value,
// else position in the concatenated receive buffer.  owner_only: take the first (lowest-rank) source.
__global__ void k_unpack(int ns, const int *__restrict__ sh_ldof, const int *__restrict__ sh_off, const int *__restrict__ sh_src,
                         const double *__restrict__ recv, double *__restrict__ y, int owner_only, const int *done)
{
   if (done && *done) { return; }
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ns; i += gridDim.x * blockDim.x)
   {
      const int l = sh_ldof[i];
      const double own = y[l];
      const int j0 = sh_off[i], j1 = owner_only ? j0 + 1 : sh_off[i + 1];
      double v = 0.0;
      for (int j = j0; j < j1; ++j)
      {
         const int s = sh_src[j];
         v += s < 0 ? own : recv[s];
      }
      y[l] = v;
   }
}

// ---------------------------------------------------------------------------------------------------
// Peer-memory path (NVLink 5 / NVSwitch, one process per GPU, buffers mapped with CUDA IPC): the shared-dof
// exchange and the scalar all-reduce are done by the kernels themselves with stores into the peers'
// mailboxes and release/acquire flags - no NCCL call (and no NCCL launch latency) inside the PCG loop.
//   mailbox of rank r (device memory of r, mapped by every peer):
//     flags_x [nranks] u64   epoch of the last exchange whose data from rank s has landed in r
//     flags_ar[nranks] u64   epoch of the last all-reduce contribution of rank s
//     ar_slot [2][nranks][4] f64   all-reduce contributions, by epoch parity
//     recv    [2][n_send(r)] f64   exchange data, by epoch parity, laid out like r's NCCL receive buffer
// A rank cannot get two epochs ahead of a peer (it needs that peer's data of the current epoch to finish
// it), so two parities suffice.  Spins are bounded: on time-out the kernel raises an error word the host
// checks, instead of hanging the GPU.
constexpr int PX_AR_MAX = 4;
constexpr unsigned long long PX_SPIN_LIMIT = 1ull << 26;

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
   asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
   unsigned long long v;
   asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
   return v;
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double *p)
{
   double v;
   asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
   return v;
}
__device__ __forceinline__ bool spin_until(const unsigned long long *flag, unsigned long long epoch, int *err)
{
   for (unsigned long long it = 0; it < PX_SPIN_LIMIT; ++it)
   {
      if (ld_acquire_sys(flag) >= epoch) { return true; }
      __nanosleep(64);
   }
   atomicExch(err, 1);
   return false;
}

struct PxPeers
{
   double *recv[26];               // per neighbour: where this rank's block starts in the neighbour's recv area (parity 0)
   long long parity_stride[26];    // n_send of the neighbour (doubles)
   unsigned long long *flag[26];   // the neighbour's flags_x[my rank]
   int off[27];                    // this rank's send offsets per neighbour
   int nbr_rank[26];
   int n_nbr;
};

// pack + send: y[shared ldofs] -> straight into the neighbours' mailboxes; the last block to finish
// publishes the epoch to every neighbour
__global__ void k_px_send(int n, const int *__restrict__ ldof, const unsigned char *__restrict__ nbr_of, const double *__restrict__ y,
                          const __grid_constant__ PxPeers P, unsigned long long epoch, unsigned int *ticket, const int *done)
{
   if (done && *done) { return; }
   const int par = (int)(epoch & 1ull);
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
   {
      const int k = nbr_of[i];
      P.recv[k][par * P.parity_stride[k] + (i - P.off[k])] = y[ldof[i]];
   }
   __threadfence_system();
   __shared__ bool last;
   __syncthreads();
   if (threadIdx.x == 0)
   {
      const unsigned int t = atomicInc(ticket, gridDim.x - 1);
      last = (t == gridDim.x - 1);
   }
   __syncthreads();
   if (last)
   {
      __threadfence_system();
      if (threadIdx.x < P.n_nbr) { st_release_sys(P.flag[threadIdx.x], epoch); }
   }
}

// The same send straight from the slot-order scratch of the operator apply: the local partial sum of a shared dof is its
// segment of y_S, summed here (ascending element order, as the segmented reduction does for everybody a moment later) and
// stored into the neighbours' mailboxes BEFORE the full E->L reduction runs - the messages fly while that kernel streams
// the whole scratch, and a neighbour that is a little ahead or behind finds the data (or its slack) already there.
__global__ void k_px_sumsend(int n, const int *__restrict__ ldof, const unsigned char *__restrict__ nbr_of, const int *__restrict__ offsets,
                             const double *__restrict__ yS, const __grid_constant__ PxPeers P, unsigned long long epoch,
                             unsigned int *ticket, const int *done)
{
   if (done && *done) { return; }
   const int par = (int)(epoch & 1ull);
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
   {
      const int l = ldof[i];
      double v = 0.0;
      const int j1 = offsets[l + 1];
      for (int j = offsets[l]; j < j1; ++j) { v += yS[j]; }
      const int k = nbr_of[i];
      P.recv[k][par * P.parity_stride[k] + (i - P.off[k])] = v;
   }
   __threadfence_system();
   __shared__ bool last;
   __syncthreads();
   if (threadIdx.x == 0)
   {
      const unsigned int t = atomicInc(ticket, gridDim.x - 1);
      last = (t == gridDim.x - 1);
   }
   __syncthreads();
   if (last)
   {
      __threadfence_system();
      if (threadIdx.x < P.n_nbr) { st_release_sys(P.flag[threadIdx.x], epoch); }
   }
}

// wait for every neighbour's data of this epoch, then the ascending-rank sum (as k_unpack)
// Optional epilogue for the operator apply (x != null): ConstrainedOperator fix-up of the shared dofs and their
// owned part of the dot x.y, reduced deterministically into *dot_out.
__global__ void k_px_recv(int ns, const int *__restrict__ sh_ldof, const int *__restrict__ sh_off, const int *__restrict__ sh_src,
                          const double *recv, double *__restrict__ y, int owner_only, const unsigned long long *my_flags,
                          const __grid_constant__ PxPeers P, unsigned long long epoch, int *err, const int *done,
                          const double *__restrict__ x, const unsigned char *__restrict__ ess_mask,
                          const unsigned char *__restrict__ own_mask, double *partials, unsigned int *ticket, double *dot_out)
{
   if (done && *done) { return; }
   if (threadIdx.x < P.n_nbr && !*(volatile int *)err) { spin_until(my_flags + P.nbr_rank[threadIdx.x], epoch, err); }
   __syncthreads();
   double acc = 0.0;
   for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ns; i += gridDim.x * blockDim.x)
   {
      const int l = sh_ldof[i];
      const double own = y[l];
      const int j0 = sh_off[i], j1 = owner_only ? j0 + 1 : sh_off[i + 1];
      double v = 0.0;
      for (int j = j0; j < j1; ++j)
      {
         const int s = sh_src[j];
         v += s < 0 ? own : ld_relaxed_sys_f64(recv + s);
      }
      if (x)
      {
         if (ess_mask && ess_mask[l]) { v = x[l]; }
         if (dot_out && own_mask[l]) { acc = fma(x[l], v, acc); }
      }
      y[l] = v;
   }
   if (x && dot_out) { b200pa::grid_sum(acc, partials, ticket, dot_out); }
}

struct PxAll
{
   double *slot[8];                // rank t's ar_slot base
   unsigned long long *flag[8];    // rank t's flags_ar
   int nranks, rank;
};

// all-reduce (sum, n <= 4 doubles) by peer stores: every rank writes its contribution into everybody's slot,
// waits for everybody's, and adds them in rank order -> bit-identical result on all ranks
// epilogue: 0 none, 1/2/3 = the PCG scalar step that consumes the reduced value (init / beta / den)
__global__ void k_px_allreduce(double *vals, int n, const __grid_constant__ PxAll P, const double *my_slots,
                               const unsigned long long *my_flags, unsigned long long epoch, int *err, int epilogue,
                               b200pa::PcgState *st, double *norms, const double *extra)
{
   if (epilogue >= 2 && st->done) { return; } // every rank holds the same scalars: all skip together
   if (extra && threadIdx.x == 0) { vals[0] += extra[0]; } // second local partial (shared dofs) of the same dot
   __syncthreads();
   const int t = threadIdx.x, par = (int)(epoch & 1ull);
   if (t < P.nranks)
   {
      double *dst = P.slot[t] + ((size_t)par * P.nranks + P.rank) * PX_AR_MAX;
      for (int k = 0; k < n; ++k) { dst[k] = vals[k]; }
      __threadfence_system();
      st_release_sys(P.flag[t] + P.rank, epoch);
      if (!*(volatile int *)err) { spin_until(my_flags + t, epoch, err); }
   }
   __syncthreads();
   if (t < n)
   {
      double s = 0.0;
      for (int r = 0; r < P.nranks; ++r) { s += ld_relaxed_sys_f64(my_slots + ((size_t)par * P.nranks + r) * PX_AR_MAX + t); }
      vals[t] = s;
   }
   if (epilogue && t == 0) // n == 1: thread 0 wrote vals[0] itself
   {
      if (epilogue == 1) { b200pa::pcg_scalar_init(st, norms); }
      else if (epilogue == 2) { b200pa::pcg_scalar_beta(st, norms); }
      else { b200pa::pcg_scalar_den(st); }
   }
}
} // namespace

struct b200pa_comm_s
{
   b200pa_ctx ctx = nullptr;
   ncclComm_t nccl = nullptr;
   int rank = 0, nranks = 1;
   int ndofs = 0, n_nbr = 0, n_send = 0, n_shared = 0;
   std::vector<int> nbr_rank, nbr_off;
   b200pa::DevBuf send_ldof, sendbuf, recvbuf, sh_ldof, sh_off, sh_src, owner_mask;
   // peer-memory path
   bool px = false;
   void *mailbox = nullptr;                  // cudaMalloc'ed (IPC-exportable)
   size_t mb_flags_x = 0, mb_flags_ar = 0, mb_slots = 0, mb_recv = 0, mb_bytes = 0; // byte offsets
   std::vector<void *> peer_mb;              // opened peer mailboxes (own entry = mailbox)
   b200pa::DevBuf send_nbr, px_err, px_ticket, shared_mask;
   PxPeers peers{};
   PxAll all{};
   unsigned long long epoch_x = 0, epoch_ar = 0;
};

using namespace b200pa;

// closes the peer mappings and frees the mailbox (the peers must not be inside an exchange: collective call sites only)
static void px_teardown(b200pa_comm c)
{
   c->px = false;
   for (size_t r = 0; r < c->peer_mb.size(); ++r)
   {
      if (c->peer_mb[r] && (int)r != c->rank) { cudaIpcCloseMemHandle(c->peer_mb[r]); }
   }
   c->peer_mb.clear();
   if (c->mailbox) { cudaFree(c->mailbox); c->mailbox = nullptr; }
   c->epoch_x = c->epoch_ar = 0;
}

extern "C" int b200pa_comm_unique_id(unsigned char id_out[128])
{
   if (!g_nccl.load()) { return fail("b200pa: " + g_nccl.err); }
   ncclUniqueId id;
   NCCL_CK(g_nccl.GetUniqueId(&id));
   std::memcpy(id_out, id.internal, 128);
   return 0;
}

extern "C" int b200pa_comm_create(b200pa_ctx ctx, const unsigned char nccl_id[128], int rank, int nranks, b200pa_comm *out)
{
   B200PA_REQUIRE(ctx && out, "comm_create: NULL argument");
   B200PA_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "comm_create: bad rank / nranks");
   B200PA_CK(cudaSetDevice(ctx->device));
   b200pa_comm c = new b200pa_comm_s;
   c->ctx = ctx; c->rank = rank; c->nranks = nranks;
   // nccl_id == NULL: no NCCL communicator - the peer-memory transport (b200pa_comm_px_prepare / _connect) is then the
   // only one and every exchange before it is connected fails.  (Ranks that share one GPU need this: NCCL refuses them.)
   if (nccl_id)
   {
      if (!g_nccl.load()) { delete c; return fail("b200pa: " + g_nccl.err); }
      ncclUniqueId id;
      std::memcpy(id.internal, nccl_id, 128);
      ncclResult_t r = g_nccl.CommInitRank(&c->nccl, nranks, id, rank);
      if (r != ncclSuccess) { delete c; return fail(std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r)); }
   }
   *out = c;
   return 0;
}

extern "C" int b200pa_comm_destroy(b200pa_comm c)
{
   if (!c) { return 0; }
   cudaSetDevice(c->ctx->device);
   cudaStreamSynchronize(c->ctx->stream);
   px_teardown(c);
   c->send_nbr.release(); c->px_err.release(); c->px_ticket.release(); c->shared_mask.release();
   if (c->nccl) { g_nccl.CommDestroy(c->nccl); }
   for (DevBuf *b : {&c->send_ldof, &c->sendbuf, &c->recvbuf, &c->sh_ldof, &c->sh_off, &c->sh_src, &c->owner_mask}) { b->release(); }
   delete c;
   return 0;
}

// Host-side table construction, shared with the CPU tests through b200pa_comm_build_tables.
// Inputs: for every neighbour k (ranks ascending, none equal to `rank`) the local L-dofs shared
// with k in an order both sides agree on (the partitioner sorts them by global dof id).
// Outputs: unique shared ldofs (ascending), CSR of sources in ascending rank order (-1 = self,
// else index into the concatenated receive buffer), owner mask (1 unless a lower rank shares the dof).
extern "C" int b200pa_comm_build_tables(int rank, int ndofs, int n_nbr, const int *nbr_rank, const int *shared_offsets,
                                        const int *shared_ldofs, int *n_shared_out, int *sh_ldof, int *sh_off, int *sh_src,
                                        unsigned char *owner_mask)
{
   B200PA_REQUIRE(n_nbr == 0 || (nbr_rank && shared_offsets && shared_ldofs), "comm_build_tables: NULL argument");
   for (int k = 0; k < n_nbr; ++k)
   {
      B200PA_REQUIRE(nbr_rank[k] != rank, "comm_build_tables: a rank cannot neighbour itself");
      B200PA_REQUIRE(k == 0 || nbr_rank[k] > nbr_rank[k - 1], "comm_build_tables: neighbour ranks must be strictly ascending");
   }
   const int n_send = n_nbr ? shared_offsets[n_nbr] : 0;
   std::vector<int> count(ndofs, 0);
   for (int i = 0; i < n_send; ++i)
   {
      B200PA_REQUIRE(shared_ldofs[i] >= 0 && shared_ldofs[i] < ndofs, "comm_build_tables: shared ldof out of range");
      count[shared_ldofs[i]]++;
   }
   std::vector<int> pos(ndofs, -1);
   int ns = 0, nsrc = 0;
   for (int l = 0; l < ndofs; ++l)
   {
      if (count[l]) { pos[l] = ns++; nsrc += count[l] + 1; }
   }
   if (n_shared_out) { *n_shared_out = ns; }
   if (owner_mask) { for (int l = 0; l < ndofs; ++l) { owner_mask[l] = 1; } }
   if (!sh_ldof || !sh_off || !sh_src) { return 0; }
   {
      int o = 0;
      for (int l = 0; l < ndofs; ++l)
      {
         if (count[l]) { sh_ldof[pos[l]] = l; sh_off[pos[l]] = o; o += count[l] + 1; }
      }
      sh_off[ns] = o;
   }
   // fill in ascending rank order: neighbours below `rank`, self, neighbours above
   std::vector<int> fill(ns, 0);
   bool self_done = false;
   auto put_self = [&]()
   {
      for (int i = 0; i < ns; ++i) { sh_src[sh_off[i] + fill[i]++] = -1; }
      self_done = true;
   };
   for (int k = 0; k < n_nbr; ++k)
   {
      if (!self_done && nbr_rank[k] > rank) { put_self(); }
      for (int i = shared_offsets[k]; i < shared_offsets[k + 1]; ++i)
      {
         const int l = shared_ldofs[i], p = pos[l];
         sh_src[sh_off[p] + fill[p]++] = i;
         if (owner_mask && nbr_rank[k] < rank) { owner_mask[l] = 0; }
      }
   }
   if (!self_done) { put_self(); }
   for (int i = 0; i < ns; ++i)
   {
      B200PA_REQUIRE(sh_off[i] + fill[i] == sh_off[i + 1], "comm_build_tables: a dof is listed twice for one neighbour");
   }
   (void)nsrc;
   return 0;
}

extern "C" int b200pa_comm_set_tables(b200pa_comm c, int ndofs, int n_nbr, const int *nbr_rank, const int *shared_offsets,
                                      const int *shared_ldofs)
{
   B200PA_REQUIRE(c, "comm is NULL");
   b200pa_ctx ctx = c->ctx;
   B200PA_CK(cudaSetDevice(ctx->device));
   // new tables invalidate the peer-memory set-up (mailbox size and the peers' offsets came from the old ones):
   // back to the NCCL transport until b200pa_comm_px_prepare / _connect run again (collectively)
   B200PA_CK(cudaStreamSynchronize(ctx->stream));
   px_teardown(c);
   int ns = 0;
   if (b200pa_comm_build_tables(c->rank, ndofs, n_nbr, nbr_rank, shared_offsets, shared_ldofs, &ns, nullptr, nullptr, nullptr, nullptr)) { return 1; }
   const int n_send = n_nbr ? shared_offsets[n_nbr] : 0;
   std::vector<int> sh_ldof(std::max(ns, 1)), sh_off(ns + 1), sh_src((size_t)n_send + ns + 1);
   std::vector<unsigned char> mask(std::max(ndofs, 1));
   if (b200pa_comm_build_tables(c->rank, ndofs, n_nbr, nbr_rank, shared_offsets, shared_ldofs, &ns, sh_ldof.data(), sh_off.data(),
                                sh_src.data(), mask.data()))
   {
      return 1;
   }
   c->ndofs = ndofs; c->n_nbr = n_nbr; c->n_send = n_send; c->n_shared = ns;
   c->nbr_rank.assign(nbr_rank, nbr_rank + n_nbr);
   c->nbr_off.assign(shared_offsets, shared_offsets + n_nbr + (n_nbr ? 1 : 0));
   if (!n_nbr) { c->nbr_off.assign(1, 0); }
   auto up = [&](DevBuf &b, const void *src, size_t bytes) -> int
   {
      if (alloc(b, std::max<size_t>(bytes, 8))) { return 1; }
      if (bytes) { B200PA_CK(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream)); }
      return 0;
   };
   if (up(c->send_ldof, shared_ldofs, sizeof(int) * (size_t)n_send) || up(c->sh_ldof, sh_ldof.data(), sizeof(int) * (size_t)ns) ||
       up(c->sh_off, sh_off.data(), sizeof(int) * ((size_t)ns + 1)) || up(c->sh_src, sh_src.data(), sizeof(int) * (size_t)sh_off[ns]) ||
       up(c->owner_mask, mask.data(), (size_t)ndofs))
   {
      return 1;
   }
   if (alloc(c->sendbuf, sizeof(double) * (size_t)std::max(n_send, 1)) || alloc(c->recvbuf, sizeof(double) * (size_t)std::max(n_send, 1))) { return 1; }
   std::vector<unsigned char> shm(std::max(ndofs, 1), 0);
   for (int i = 0; i < ns; ++i) { shm[sh_ldof[i]] = 1; }
   if (up(c->shared_mask, shm.data(), (size_t)ndofs)) { return 1; }
   B200PA_CK(cudaStreamSynchronize(ctx->stream));
   return 0;
}

namespace b200pa
{
const unsigned char *comm_owner_mask(b200pa_comm c) { return c->owner_mask.as<unsigned char>(); }

struct ApplyEpilogue
{
   const double *x = nullptr;
   const unsigned char *ess_mask = nullptr;
   double *dot_out = nullptr;
};

static int exchange(b200pa_comm c, double *y, int owner_only, const int *done, const ApplyEpilogue &ep = ApplyEpilogue())
{
   b200pa_ctx ctx = c->ctx;
   if (c->n_send == 0) { return 0; }
   const int bs = 256;
   int g = (c->n_send + bs - 1) / bs;
   g = std::min(g, ctx->num_sms * 8);
   if (c->px)
   {
      const unsigned long long epoch = ++c->epoch_x;
      const char *mb = (const char *)c->mailbox;
      k_px_send<<<g, bs, 0, ctx->stream>>>(c->n_send, c->send_ldof.as<int>(), c->send_nbr.as<unsigned char>(), y, c->peers, epoch,
                                           c->px_ticket.as<unsigned int>(), done);
      B200PA_LAUNCHED();
      const int g2 = std::max(1, std::min((c->n_shared + bs - 1) / bs, ctx->num_sms * 8));
      const double *recv = (const double *)(mb + c->mb_recv) + (size_t)(epoch & 1ull) * (size_t)c->n_send;
      k_px_recv<<<g2, bs, 0, ctx->stream>>>(c->n_shared, c->sh_ldof.as<int>(), c->sh_off.as<int>(), c->sh_src.as<int>(), recv, y, owner_only,
                                            (const unsigned long long *)(mb + c->mb_flags_x), c->peers, epoch, c->px_err.as<int>(), done,
                                            ep.x, ep.ess_mask, c->owner_mask.as<unsigned char>(), ctx->d_partials, ctx->d_ticket, ep.dot_out);
      B200PA_LAUNCHED();
      return 0;
   }
   B200PA_REQUIRE(c->nccl, "shared-dof exchange: this communicator was created without NCCL and its peer-memory path is not connected");
   k_pack<<<g, bs, 0, ctx->stream>>>(c->n_send, c->send_ldof.as<int>(), y, c->sendbuf.as<double>(), done);
   B200PA_LAUNCHED();
   NCCL_CK(g_nccl.GroupStart());
   for (int k = 0; k < c->n_nbr; ++k)
   {
      const size_t cnt = (size_t)(c->nbr_off[k + 1] - c->nbr_off[k]);
      if (!cnt) { continue; }
      NCCL_CK(g_nccl.Send(c->sendbuf.as<double>() + c->nbr_off[k], cnt, ncclFloat64, c->nbr_rank[k], c->nccl, ctx->stream));
      NCCL_CK(g_nccl.Recv(c->recvbuf.as<double>() + c->nbr_off[k], cnt, ncclFloat64, c->nbr_rank[k], c->nccl, ctx->stream));
   }
   NCCL_CK(g_nccl.GroupEnd());
   g = std::min((c->n_shared + bs - 1) / bs, ctx->num_sms * 8);
   k_unpack<<<std::max(g, 1), bs, 0, ctx->stream>>>(c->n_shared, c->sh_ldof.as<int>(), c->sh_off.as<int>(), c->sh_src.as<int>(),
                                                    c->recvbuf.as<double>(), y, owner_only, done);
   B200PA_LAUNCHED();
   return 0;
}

int comm_exchange_sum(b200pa_comm c, double *y, const int *done) { return exchange(c, y, 0, done); }
bool comm_px(b200pa_comm c) { return c && c->px; }
const unsigned char *comm_shared_mask(b200pa_comm c) { return c->shared_mask.as<unsigned char>(); }
// peer path only: exchange + constraint fix-up of the shared dofs + their owned part of x.y -> *dot_out
int comm_exchange_sum_apply(b200pa_comm c, double *y, const int *done, const double *x, const unsigned char *ess_mask, double *dot_out)
{
   ApplyEpilogue ep;
   ep.x = x; ep.ess_mask = ess_mask; ep.dot_out = dot_out;
   return exchange(c, y, 0, done, ep);
}
int comm_exchange_owner(b200pa_comm c, double *x) { return exchange(c, x, 1, nullptr); }

// peer path, operator apply: first half (before the full segmented reduction) ...
int comm_px_send_from_slots(b200pa_comm c, const int *offsets, const double *yS, const int *done)
{
   b200pa_ctx ctx = c->ctx;
   if (c->n_send == 0) { return 0; }
   const int bs = 256;
   const int g = std::min((c->n_send + bs - 1) / bs, ctx->num_sms * 8);
   const unsigned long long epoch = ++c->epoch_x;
   k_px_sumsend<<<g, bs, 0, ctx->stream>>>(c->n_send, c->send_ldof.as<int>(), c->send_nbr.as<unsigned char>(), offsets, yS, c->peers, epoch,
                                           c->px_ticket.as<unsigned int>(), done);
   B200PA_LAUNCHED();
   return 0;
}
// ... and second half: wait for the neighbours' data of the epoch the send opened, finish the shared dofs of y
int comm_px_recv_apply(b200pa_comm c, double *y, const int *done, const double *x, const unsigned char *ess_mask, double *dot_out)
{
   b200pa_ctx ctx = c->ctx;
   if (c->n_send == 0) { return 0; }
   const int bs = 256;
   const unsigned long long epoch = c->epoch_x;
   const char *mb = (const char *)c->mailbox;
   const int g2 = std::max(1, std::min((c->n_shared + bs - 1) / bs, ctx->num_sms * 8));
   const double *recv = (const double *)(mb + c->mb_recv) + (size_t)(epoch & 1ull) * (size_t)c->n_send;
   k_px_recv<<<g2, bs, 0, ctx->stream>>>(c->n_shared, c->sh_ldof.as<int>(), c->sh_off.as<int>(), c->sh_src.as<int>(), recv, y, 0,
                                         (const unsigned long long *)(mb + c->mb_flags_x), c->peers, epoch, c->px_err.as<int>(), done,
                                         x, ess_mask, c->owner_mask.as<unsigned char>(), ctx->d_partials, ctx->d_ticket, dot_out);
   B200PA_LAUNCHED();
   return 0;
}
const int *comm_px_err_ptr(b200pa_comm c) { return (c && c->px) ? c->px_err.as<int>() : nullptr; }
int comm_px_check(b200pa_comm c, const char *where)
{
   if (!c || !c->px) { return 0; }
   if (b200pa_comm_px_error(c))
   {
      return fail(std::string(where) + ": a peer-memory wait timed out (a rank did not take part in the shared-dof exchange / "
                  "all-reduce); the result is invalid");
   }
   return 0;
}
int comm_validate(b200pa_comm c, int ndofs)
{
   B200PA_REQUIRE(c->nranks == 1 || c->owner_mask.p, "form_set_comm: the communicator has no neighbour tables (call b200pa_comm_set_tables first)");
   B200PA_REQUIRE(c->nranks == 1 || c->ndofs == ndofs, "form_set_comm: the communicator's tables were built for a different number of L-dofs");
   return 0;
}

// all-reduce of one PCG dot + the scalar step that follows it, in one launch (peer-memory path only;
// returns 0 and does nothing when that path is off, the caller then uses the NCCL all-reduce + a scalar kernel)
int comm_allreduce_scalar_step(b200pa_comm c, double *val, int step, void *pcg_state, double *norms, bool *handled,
                               const double *extra)
{
   *handled = false;
   if (!c->px || c->nranks == 1) { return 0; }
   const unsigned long long epoch = ++c->epoch_ar;
   const char *mb = (const char *)c->mailbox;
   k_px_allreduce<<<1, 32, 0, c->ctx->stream>>>(val, 1, c->all, (const double *)(mb + c->mb_slots),
                                               (const unsigned long long *)(mb + c->mb_flags_ar), epoch, c->px_err.as<int>(), step,
                                               (PcgState *)pcg_state, norms, extra);
   B200PA_LAUNCHED();
   *handled = true;
   return 0;
}

int comm_allreduce_sum_dev(b200pa_comm c, double *vals, int n)
{
   if (c->nranks == 1) { return 0; }
   if (c->px && n <= PX_AR_MAX)
   {
      const unsigned long long epoch = ++c->epoch_ar;
      const char *mb = (const char *)c->mailbox;
      k_px_allreduce<<<1, 32, 0, c->ctx->stream>>>(vals, n, c->all, (const double *)(mb + c->mb_slots),
                                                  (const unsigned long long *)(mb + c->mb_flags_ar), epoch, c->px_err.as<int>(), 0, nullptr,
                                                  nullptr, nullptr);
      B200PA_LAUNCHED();
      return 0;
   }
   B200PA_REQUIRE(c->nccl, "all-reduce: this communicator was created without NCCL and its peer-memory path is not connected (or n > 4)");
   NCCL_CK(g_nccl.AllReduce(vals, vals, (size_t)n, ncclFloat64, ncclSum, c->nccl, c->ctx->stream));
   return 0;
}
} // namespace b200pa

// ---- peer-memory set-up (two steps around one host-side all-gather of the handles and neighbour tables)
extern "C" int b200pa_comm_px_prepare(b200pa_comm c, unsigned char handle_out[64])
{
   B200PA_REQUIRE(c && handle_out, "comm_px_prepare: NULL argument");
   B200PA_REQUIRE(c->nranks <= 8 && c->n_nbr <= 26, "comm_px_prepare: peer path supports up to 8 ranks / 26 neighbours");
   static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
   b200pa_ctx ctx = c->ctx;
   B200PA_CK(cudaSetDevice(ctx->device));
   if (!c->mailbox)
   {
      auto up256 = [](size_t v) { return (v + 255) & ~(size_t)255; };
      c->mb_flags_x = 0;
      c->mb_flags_ar = up256(c->mb_flags_x + sizeof(unsigned long long) * c->nranks);
      c->mb_slots = up256(c->mb_flags_ar + sizeof(unsigned long long) * c->nranks);
      c->mb_recv = up256(c->mb_slots + sizeof(double) * 2 * c->nranks * PX_AR_MAX);
      c->mb_bytes = up256(c->mb_recv + sizeof(double) * 2 * (size_t)std::max(c->n_send, 1));
      B200PA_CK(cudaMalloc(&c->mailbox, c->mb_bytes));
      B200PA_CK(cudaMemset(c->mailbox, 0, c->mb_bytes));
      B200PA_CK(cudaDeviceSynchronize());
   }
   cudaIpcMemHandle_t h;
   B200PA_CK(cudaIpcGetMemHandle(&h, c->mailbox));
   std::memcpy(handle_out, &h, 64);
   return 0;
}

// handles: nranks x 64 bytes (rank order); for every neighbour k (this rank's order): remote_off[k] = where the
// neighbour expects this rank's block in its receive area, remote_nsend[k] = the neighbour's total send count
extern "C" int b200pa_comm_px_connect(b200pa_comm c, const unsigned char *handles, const long long *remote_off,
                                      const long long *remote_nsend)
{
   B200PA_REQUIRE(c && handles && (c->n_nbr == 0 || (remote_off && remote_nsend)), "comm_px_connect: NULL argument");
   B200PA_REQUIRE(c->mailbox, "comm_px_connect: call b200pa_comm_px_prepare first");
   b200pa_ctx ctx = c->ctx;
   B200PA_CK(cudaSetDevice(ctx->device));
   B200PA_REQUIRE(!c->px && c->peer_mb.empty(), "comm_px_connect: already connected (b200pa_comm_set_tables resets the peer path)");
   c->peer_mb.assign(c->nranks, nullptr);
   for (int r = 0; r < c->nranks; ++r)
   {
      if (r == c->rank) { c->peer_mb[r] = c->mailbox; continue; }
      cudaIpcMemHandle_t h;
      std::memcpy(&h, handles + 64 * (size_t)r, 64);
      const cudaError_t e = cudaIpcOpenMemHandle(&c->peer_mb[r], h, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess)
      {
         c->peer_mb[r] = nullptr;
         for (int q = 0; q < r; ++q) { if (q != c->rank && c->peer_mb[q]) { cudaIpcCloseMemHandle(c->peer_mb[q]); } }
         c->peer_mb.clear();
         cudaGetLastError();
         return fail(std::string("comm_px_connect: cudaIpcOpenMemHandle(rank ") + std::to_string(r) + "): " + cudaGetErrorString(e));
      }
   }
   // layout is the same function of (nranks, n_send) on every rank; only the recv offset depends on the peer's n_send,
   // and that offset (mb_recv) depends on nranks only
   std::vector<unsigned char> nbr_of((size_t)std::max(c->n_send, 1));
   PxPeers &P = c->peers;
   P.n_nbr = c->n_nbr;
   for (int k = 0; k < c->n_nbr; ++k)
   {
      char *pm = (char *)c->peer_mb[c->nbr_rank[k]];
      P.recv[k] = (double *)(pm + c->mb_recv) + remote_off[k];
      P.parity_stride[k] = remote_nsend[k];
      P.flag[k] = (unsigned long long *)(pm + c->mb_flags_x) + c->rank;
      P.off[k] = c->nbr_off[k];
      P.nbr_rank[k] = c->nbr_rank[k];
      for (int i = c->nbr_off[k]; i < c->nbr_off[k + 1]; ++i) { nbr_of[i] = (unsigned char)k; }
   }
   P.off[c->n_nbr] = c->n_send;
   PxAll &A = c->all;
   A.nranks = c->nranks; A.rank = c->rank;
   for (int r = 0; r < c->nranks; ++r)
   {
      char *pm = (char *)c->peer_mb[r];
      A.slot[r] = (double *)(pm + c->mb_slots);
      A.flag[r] = (unsigned long long *)(pm + c->mb_flags_ar);
   }
   if (alloc(c->send_nbr, nbr_of.size()) || alloc(c->px_err, sizeof(int)) || alloc(c->px_ticket, sizeof(unsigned int))) { return 1; }
   B200PA_CK(cudaMemcpy(c->send_nbr.p, nbr_of.data(), nbr_of.size(), cudaMemcpyHostToDevice));
   B200PA_CK(cudaMemset(c->px_err.p, 0, sizeof(int)));
   B200PA_CK(cudaMemset(c->px_ticket.p, 0, sizeof(unsigned int)));
   B200PA_CK(cudaDeviceSynchronize());
   c->px = true;
   return 0;
}

// 0 = no error; nonzero = a peer wait timed out since the last query (clears the flag).  Synchronises the stream.
extern "C" int b200pa_comm_px_error(b200pa_comm c)
{
   if (!c || !c->px) { return 0; }
   int h = 0;
   cudaSetDevice(c->ctx->device);
   cudaMemcpyAsync(&h, c->px_err.p, sizeof(int), cudaMemcpyDeviceToHost, c->ctx->stream);
   cudaStreamSynchronize(c->ctx->stream);
   if (h) { cudaMemsetAsync(c->px_err.p, 0, sizeof(int), c->ctx->stream); }
   return h;
}
extern "C" int b200pa_comm_px_enabled(b200pa_comm c) { return c && c->px ? 1 : 0; }
// back to the NCCL transport (collective decision of the host application, e.g. when one rank could not map a peer)
extern "C" int b200pa_comm_px_disable(b200pa_comm c)
{
   if (c) { c->px = false; }
   return 0;
}

extern "C" int b200pa_comm_exchange_sum(b200pa_comm c, double *yL_dev)
{
   B200PA_REQUIRE(c && yL_dev, "comm_exchange_sum: NULL argument");
   B200PA_CK(cudaSetDevice(c->ctx->device));
   return comm_exchange_sum(c, yL_dev, nullptr);
}
extern "C" int b200pa_comm_bcast(b200pa_comm c, double *xL_dev)
{
   B200PA_REQUIRE(c && xL_dev, "comm_bcast: NULL argument");
   B200PA_CK(cudaSetDevice(c->ctx->device));
   return comm_exchange_owner(c, xL_dev);
}
extern "C" int b200pa_comm_allreduce_sum(b200pa_comm c, double *vals_dev, int n)
{
   B200PA_REQUIRE(c && vals_dev && n >= 0, "comm_allreduce_sum: bad argument");
   B200PA_CK(cudaSetDevice(c->ctx->device));
   return comm_allreduce_sum_dev(c, vals_dev, n);
}
extern "C" const unsigned char *b200pa_comm_owner_mask(b200pa_comm c) { return c ? c->owner_mask.as<unsigned char>() : nullptr; }
