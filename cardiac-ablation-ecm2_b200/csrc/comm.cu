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
} // namespace

struct b200pa_comm_s
{
   b200pa_ctx ctx = nullptr;
   ncclComm_t nccl = nullptr;
   int rank = 0, nranks = 1;
   int ndofs = 0, n_nbr = 0, n_send = 0, n_shared = 0;
   std::vector<int> nbr_rank, nbr_off;
   b200pa::DevBuf send_ldof, sendbuf, recvbuf, sh_ldof, sh_off, sh_src, owner_mask;
};

using namespace b200pa;

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
   B200PA_REQUIRE(ctx && nccl_id && out, "comm_create: NULL argument");
   B200PA_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "comm_create: bad rank / nranks");
   if (!g_nccl.load()) { return fail("b200pa: " + g_nccl.err); }
   B200PA_CK(cudaSetDevice(ctx->device));
   b200pa_comm c = new b200pa_comm_s;
   c->ctx = ctx; c->rank = rank; c->nranks = nranks;
   ncclUniqueId id;
   std::memcpy(id.internal, nccl_id, 128);
   ncclResult_t r = g_nccl.CommInitRank(&c->nccl, nranks, id, rank);
   if (r != ncclSuccess) { delete c; return fail(std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r)); }
   *out = c;
   return 0;
}

extern "C" int b200pa_comm_destroy(b200pa_comm c)
{
   if (!c) { return 0; }
   cudaSetDevice(c->ctx->device);
   cudaStreamSynchronize(c->ctx->stream);
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
   B200PA_CK(cudaStreamSynchronize(ctx->stream));
   return 0;
}

namespace b200pa
{
const unsigned char *comm_owner_mask(b200pa_comm c) { return c->owner_mask.as<unsigned char>(); }

static int exchange(b200pa_comm c, double *y, int owner_only, const int *done)
{
   b200pa_ctx ctx = c->ctx;
   if (c->n_send == 0) { return 0; }
   const int bs = 256;
   int g = (c->n_send + bs - 1) / bs;
   g = std::min(g, ctx->num_sms * 8);
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
int comm_exchange_owner(b200pa_comm c, double *x) { return exchange(c, x, 1, nullptr); }

int comm_allreduce_sum_dev(b200pa_comm c, double *vals, int n)
{
   if (c->nranks == 1) { return 0; }
   NCCL_CK(g_nccl.AllReduce(vals, vals, (size_t)n, ncclFloat64, ncclSum, c->nccl, c->ctx->stream));
   return 0;
}
} // namespace b200pa

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
