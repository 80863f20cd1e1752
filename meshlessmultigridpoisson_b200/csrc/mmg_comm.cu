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
  const char* forced = getenv("MMG_NCCL_LIB");          // tests point this at a missing file to exercise the error path
  const char* names[] = {forced ? forced : "libnccl.so.2", forced ? forced : "libnccl.so"};
  for (const char* n : names) {
    api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
  }
  if (!api.handle) {
    const char* e = dlerror();                           // dlerror() clears the message: read it exactly once
    throw Error(MMG_ERR_NCCL, std::string("cannot load libnccl.so.2: ") + (e ? e : "not found"));
  }
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

// ncclGroupStart ... ncclGroupEnd with the group closed on every path: an exception between the two must not leave the
// communicator inside an open group
struct NcclGroup {
  NcclApi& n;
  bool open = false;
  explicit NcclGroup(NcclApi& api) : n(api) {
    const int r = n.GroupStart();
    if (r != ncclSuccess) throw Error(MMG_ERR_NCCL, std::string("ncclGroupStart failed: ") + n.GetErrorString(r));
    open = true;
  }
  void close() {
    if (!open) return;
    open = false;
    const int r = n.GroupEnd();
    if (r != ncclSuccess) throw Error(MMG_ERR_NCCL, std::string("ncclGroupEnd failed: ") + n.GetErrorString(r));
  }
  ~NcclGroup() { if (open) n.GroupEnd(); }
};

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

// The versioned-vector allocation of every partitioned level is exported with cudaIpcGetMemHandle, the 64-byte handles
// travel through NCCL (a grouped broadcast per level, so the C-ABI needs no side channel), and every rank maps the
// allocations of the ranks it sends halo values to.  Any failure leaves peer_ready false: the NCCL exchange path stays.
void peer_setup(Solver& s) {
  NcclApi& n = nccl();
  const int W = s.world;
  for (size_t l = 0; l < s.grids.size(); l++) {
    LevelDist& D = s.dist[l];
    if (!D.partitioned) continue;
    Grid& g = *s.grids[l];
    int usable = (D.x_plan.sends.size() <= 2 && g.props.iters >= 1) ? 1 : 0;
    D.peer_iters = std::max(g.props.iters, 1);
    D.peer_stride = ((size_t)g.A + 63) / 64 * 64;
    D.peer_xs.alloc((size_t)2 * (D.peer_iters + 1) * D.peer_stride);
    cudaIpcMemHandle_t mine;
    if (cudaIpcGetMemHandle(&mine, D.peer_xs.p) != cudaSuccess) { cudaGetLastError(); usable = 0; std::memset(&mine, 0, sizeof(mine)); }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    DevBuf<unsigned char> dh;
    dh.alloc((size_t)W * 72);
    std::vector<unsigned char> hh((size_t)W * 72, 0);
    std::memcpy(hh.data() + (size_t)s.rank * 72, &mine, 64);
    hh[(size_t)s.rank * 72 + 64] = (unsigned char)usable;
    dh.upload(hh, g.stream);
    {
      NcclGroup grp(n);
      for (int r = 0; r < W; r++) MMG_NCCL(n.Broadcast(dh.p + (size_t)r * 72, dh.p + (size_t)r * 72, 72, /*ncclUint8*/ 1, r, (ncclComm_t)s.nccl_comm, g.stream));
      grp.close();
    }
    hh = dh.to_host(g.stream);
    bool all = true;
    for (int r = 0; r < W; r++) all = all && hh[(size_t)r * 72 + 64] == 1;
    D.n_sends = 0;
    if (all) {
      for (const ExchangePlan::Msg& m : D.x_plan.sends) {
        cudaIpcMemHandle_t h;
        std::memcpy(&h, hh.data() + (size_t)m.peer * 72, 64);
        void* ptr = nullptr;
        if (cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); all = false; break; }
        D.send_lo[D.n_sends] = m.offset; D.send_hi[D.n_sends] = m.offset + m.count; D.send_base[D.n_sends] = (double*)ptr;
        D.n_sends++;
      }
    }
    // every rank must agree, otherwise one side would wait for stores the other never makes
    s.sums.zero(g.stream);
    if (!all) { const double one = 1.0; MMG_CUDA(cudaMemcpyAsync(s.sums.p, &one, sizeof(double), cudaMemcpyHostToDevice, g.stream)); }
    peer_init_sets(g, D);
    allreduce_sum(s, s.sums.p, 1);                          // also orders every rank's initialisation before anybody's first store
    double failed = 0;
    s.sums.download(&failed, 1, g.stream);
    D.peer_ready = failed == 0.0;
    D.peer_parity = 0;
  }
}

void peer_teardown(Solver& s) {
  for (LevelDist& D : s.dist) {
    for (int k = 0; k < D.n_sends; k++) if (D.send_base[k]) { cudaIpcCloseMemHandle(D.send_base[k]); D.send_base[k] = nullptr; }
    D.n_sends = 0; D.peer_ready = false;
  }
}

void comm_destroy(Solver& s) {
  peer_teardown(s);
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
  {
    NcclGroup grp(n);
    for (const ExchangePlan::Msg& m : P.sends) MMG_NCCL(n.Send(vec + m.offset, (size_t)m.count, ncclFloat64, m.peer, (ncclComm_t)s.nccl_comm, s.stream));
    for (const ExchangePlan::Msg& m : P.recvs) MMG_NCCL(n.Recv(vec + m.offset, (size_t)m.count, ncclFloat64, m.peer, (ncclComm_t)s.nccl_comm, s.stream));
    grp.close();
  }
  s.comm_msgs += (int64_t)P.sends.size() + (int64_t)P.recvs.size();
  for (const ExchangePlan::Msg& m : P.sends) s.comm_bytes += (int64_t)m.count * 8;
}

// every rank contributes vec[bounds[r], bounds[r+1]) ; afterwards all ranks hold the whole vector
void allgather_blocks(Solver& s, double* vec, const std::vector<int>& bounds) {
  if (s.world == 1) return;
  NcclApi& n = nccl();
  {
    NcclGroup grp(n);
    for (int r = 0; r < s.world; r++) {
      const int cnt = bounds[r + 1] - bounds[r];
      if (cnt > 0) MMG_NCCL(n.Broadcast(vec + bounds[r], vec + bounds[r], (size_t)cnt, ncclFloat64, r, (ncclComm_t)s.nccl_comm, s.stream));
    }
    grp.close();
  }
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
