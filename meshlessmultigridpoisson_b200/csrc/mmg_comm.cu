// Multi-GPU plumbing of the solve path: NCCL (resolved at run time with dlopen, so single-GPU use needs no NCCL at
// all), the contiguous row-block partition of SURVEY.md §8(e), and halo-exchange plans built from the column ranges
// each rank's rows touch.  One process per GPU; every rank holds full-length vectors and (this round) the full
// operators, computes only its own row block of every partitioned level and exchanges contiguous index ranges with
// grouped ncclSend/ncclRecv on the solver's stream.  Levels below the partition threshold are replicated: every rank
// computes them redundantly and identically, which is the trivial form of coarse-level agglomeration.
#include <dlfcn.h>

#include <algorithm>
#include <cstring>

#include "mmg_internal.hpp"

namespace mmg {

namespace {
typedef struct { char internal[128]; } ncclUniqueId;
typedef void* ncclComm_t;
enum { ncclSuccess = 0, ncclSum = 0, ncclInt32 = 2, ncclFloat64 = 8 };

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};

NcclApi& nccl() {
  static NcclApi api;
  if (api.handle) return api;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
  }
  MMG_REQUIRE(api.handle != nullptr, MMG_ERR_NCCL, std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "not found"));
  auto sym = [&](const char* s) {
    void* p = dlsym(api.handle, s);
    MMG_REQUIRE(p != nullptr, MMG_ERR_NCCL, std::string("libnccl lacks ") + s);
    return p;
  };
  api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
  api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
  api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
  api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
  api.Send = (decltype(api.Send))sym("ncclSend");
  api.Recv = (decltype(api.Recv))sym("ncclRecv");
  api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
  api.Broadcast = (decltype(api.Broadcast))sym("ncclBroadcast");
  api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
  return api;
}

#define MMG_NCCL(call)                                                                                               \
  do {                                                                                                               \
    int r__ = (call);                                                                                                \
    if (r__ != ncclSuccess) throw Error(MMG_ERR_NCCL, std::string(#call) + " failed: " + nccl().GetErrorString(r__)); \
  } while (0)
}  // namespace

// rank(i) = the block that holds i when [0,n) is cut into `world` contiguous blocks whose sizes differ by at most one
void partition_bounds(int n, int world, int* bounds) {
  const int base = n / world, rem = n % world;
  bounds[0] = 0;
  for (int r = 0; r < world; r++) bounds[r + 1] = bounds[r] + base + (r < rem ? 1 : 0);
}

void comm_unique_id(char* out128) {
  ncclUniqueId id;
  MMG_NCCL(nccl().GetUniqueId(&id));
  std::memcpy(out128, id.internal, 128);
}

void comm_init(Solver& s, int rank, int world, const char* id128) {
  MMG_REQUIRE(world >= 1 && rank >= 0 && rank < world, MMG_ERR_ARG, "init_comm: bad rank/world");
  MMG_REQUIRE(!s.grids.empty(), MMG_ERR_STATE, "init_comm: add the grids first (the communicator lives on their device)");
  s.rank = rank; s.world = world;
  if (world == 1) return;
  ncclUniqueId id;
  std::memcpy(id.internal, id128, 128);
  ncclComm_t c = nullptr;
  MMG_NCCL(nccl().CommInitRank(&c, world, id, rank));
  s.nccl_comm = c;
}

void comm_destroy(Solver& s) {
  if (s.nccl_comm) { nccl().CommDestroy((ncclComm_t)s.nccl_comm); s.nccl_comm = nullptr; }
}

// intersection of the need interval of `needer` with what `owner` owns, as (offset, count); count 0 if empty
static std::pair<int, int> overlap(int nlo, int nhi, int olo, int ohi) {
  const int a = std::max(nlo, olo), b = std::min(nhi, ohi);
  return {a, std::max(0, b - a)};
}

// need[r] = [lo, hi) index range of the vector that rank r's rows read; bounds = ownership of that vector
void plan_build(ExchangePlan& P, int rank, int world, const std::vector<std::pair<int, int>>& need, const std::vector<int>& bounds) {
  P.sends.clear(); P.recvs.clear();
  for (int r = 0; r < world; r++) {
    if (r == rank) continue;
    auto s = overlap(need[r].first, need[r].second, bounds[rank], bounds[rank + 1]);   // what r reads of mine
    if (s.second > 0) P.sends.push_back({r, s.first, s.second});
    auto v = overlap(need[rank].first, need[rank].second, bounds[r], bounds[r + 1]);   // what I read of r's
    if (v.second > 0) P.recvs.push_back({r, v.first, v.second});
  }
}

void plan_execute(Solver& s, const ExchangePlan& P, double* vec) {
  if (s.world == 1 || (P.sends.empty() && P.recvs.empty())) return;
  NcclApi& n = nccl();
  MMG_NCCL(n.GroupStart());
  for (const ExchangePlan::Msg& m : P.sends) MMG_NCCL(n.Send(vec + m.offset, (size_t)m.count, ncclFloat64, m.peer, (ncclComm_t)s.nccl_comm, s.stream));
  for (const ExchangePlan::Msg& m : P.recvs) MMG_NCCL(n.Recv(vec + m.offset, (size_t)m.count, ncclFloat64, m.peer, (ncclComm_t)s.nccl_comm, s.stream));
  MMG_NCCL(n.GroupEnd());
  s.comm_msgs += (int64_t)P.sends.size() + (int64_t)P.recvs.size();
  for (const ExchangePlan::Msg& m : P.sends) s.comm_bytes += (int64_t)m.count * 8;
}

// every rank contributes vec[bounds[r], bounds[r+1]) ; afterwards all ranks hold the whole vector
void allgather_blocks(Solver& s, double* vec, const std::vector<int>& bounds) {
  if (s.world == 1) return;
  NcclApi& n = nccl();
  MMG_NCCL(n.GroupStart());
  for (int r = 0; r < s.world; r++) {
    const int cnt = bounds[r + 1] - bounds[r];
    if (cnt > 0) MMG_NCCL(n.Broadcast(vec + bounds[r], vec + bounds[r], (size_t)cnt, ncclFloat64, r, (ncclComm_t)s.nccl_comm, s.stream));
  }
  MMG_NCCL(n.GroupEnd());
  s.comm_msgs += s.world;
  s.comm_bytes += (int64_t)(bounds[s.rank + 1] - bounds[s.rank]) * 8 * (s.world - 1);
}

void allreduce_sum(Solver& s, double* dev, int count) {
  if (s.world == 1) return;
  MMG_NCCL(nccl().AllReduce(dev, dev, (size_t)count, ncclFloat64, ncclSum, (ncclComm_t)s.nccl_comm, s.stream));
  s.comm_msgs += 1;
  s.comm_bytes += (int64_t)count * 8;
}

}  // namespace mmg
