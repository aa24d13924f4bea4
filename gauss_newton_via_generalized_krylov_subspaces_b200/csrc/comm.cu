// comm.cu -- the only data that crosses GPUs: one halo exchange per outer iteration and a few
// tiny all-gathers (dot products, norms, TSQR R factors).  The library owns its NCCL communicator; NCCL is
// resolved with dlopen at run time (the torch-bundled libnccl.so.2 that is already mapped into the Python
// process), so the shared object has no link-time NCCL dependency and single-GPU use needs no NCCL at all.
// Reductions are "all-gather, then the same rank-ordered reduction on every rank": results are bitwise
// identical on all ranks and independent of NCCL's internal algorithm choice.
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace {

typedef struct { char internal[128]; } NcclId;
typedef void* NcclComm;
enum { NCCL_F64 = 8 };  // ncclFloat64 (nccl.h: ncclDataType_t)

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclId, int) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
  int (*Broadcast)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;

int load_nccl() {
  if (g_nccl.handle) return 0;
  const char* cand[] = {getenv("GNK_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* c : cand) {
    if (!c || !*c) continue;
    h = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) {
    gnk_set_error(std::string("cannot dlopen NCCL (set GNK_NCCL_LIB): ") + (dlerror() ? dlerror() : "?"));
    return -3;
  }
#define SYM(field, name)                                                 \
  *(void**)(&g_nccl.field) = dlsym(h, name);                             \
  if (!g_nccl.field) {                                                   \
    gnk_set_error(std::string("NCCL symbol missing: ") + name);          \
    return -3;                                                           \
  }
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(AllGather, "ncclAllGather");
  SYM(Broadcast, "ncclBroadcast");
  SYM(Send, "ncclSend");
  SYM(Recv, "ncclRecv");
  SYM(GroupStart, "ncclGroupStart");
  SYM(GroupEnd, "ncclGroupEnd");
  SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  g_nccl.handle = h;
  return 0;
}

int nccl_fail(const char* what, int rc) {
  gnk_set_error(std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "nccl error"));
  return -4;
}
#define GNK_NCCL(expr)                          \
  do {                                          \
    int rc__ = (expr);                          \
    if (rc__ != 0) return nccl_fail(#expr, rc__); \
  } while (0)

// d_buf[i] = reduce_r gathered[r*count + i] in rank order
__global__ void rank_reduce_kernel(const double* __restrict__ gathered, int nranks, int count, int op,
                                   double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  double a = gathered[i];
  const bool is_max = (op == 1) || (op == 2 && i == 1);
  for (int r = 1; r < nranks; ++r) {
    const double b = gathered[(int64_t)r * count + i];
    a = is_max ? fmax(a, b) : a + b;
  }
  out[i] = a;
}

int ensure_gather(gnk_ctx* ctx, size_t bytes, cudaStream_t st) {
  if (bytes <= ctx->gather_bytes) return 0;
  GNK_CUDA(cudaStreamSynchronize(st));
  if (ctx->d_gather) GNK_CUDA(cudaFree(ctx->d_gather));
  ctx->d_gather = nullptr;
  GNK_CUDA(cudaMalloc(&ctx->d_gather, bytes));
  ctx->gather_bytes = bytes;
  return 0;
}


// =================================================================================================
// Peer-memory path (one node, NVLink/NVSwitch): every rank owns a "mailbox" in its HBM that all peers map with
// CUDA IPC.  A collective is ONE single-CTA kernel: write my contribution into every peer's mailbox with plain
// stores (they travel over NVLink), fence, raise my flag in every peer's mailbox (st.release.sys), wait until all
// peers' flags in MY mailbox carry this collective's sequence number (ld.acquire.sys), then reduce the slots in rank
// order out of local memory.  The messages are <= 85 KB (k doubles, 2 scalars, a (k+1)^2 triangle, two grid rows), so
// the cost is NVLink latency (~2-3 us) instead of NCCL's ~20 us launch+protocol path, and one launch instead of two.
// Slots are double-buffered on the parity of the sequence number: a rank can be at most one collective ahead of a
// peer (it cannot finish collective s without the peer's flag s), so slot parity s is never overwritten before the
// peer has left collective s-2.  Results are bitwise identical on all ranks (same rank-ordered reduction).
// The residual, Gram-Schmidt dots and update-statistics kernels run the same protocol in their own last CTA
// (p2p_tail_allreduce, common.cuh), so those reductions need no kernel of their own at all.
// =================================================================================================
inline size_t p2p_bytes(int nranks) { return p2p_hll_off(nranks, 2, 0); }

// op: 0 sum, 1 max, 2 (sum, max) pair as in gnk_comm_allreduce, 3 no reduction: out receives the nranks x count stack
__global__ void __launch_bounds__(1024) p2p_gather_kernel(void* const* __restrict__ peers, int rank, int nranks,
                                                           const double* __restrict__ src, int count, int op,
                                                           unsigned long long seq, double* __restrict__ out) {
  const int parity = (int)(seq & 1ull);
  for (int r = 0; r < nranks; ++r) {
    double* dst = reinterpret_cast<double*>(static_cast<char*>(peers[r]) + p2p_gather_off(nranks, parity, rank));
    for (int i = threadIdx.x; i < count; i += blockDim.x) dst[i] = src[i];
  }
  __threadfence_system();
  __syncthreads();
  char* mine = static_cast<char*>(peers[rank]);
  if ((int)threadIdx.x < nranks) {
    unsigned long long* theirs = reinterpret_cast<unsigned long long*>(peers[threadIdx.x]);
    st_release_sys(theirs + rank, seq);
    wait_flag(reinterpret_cast<const unsigned long long*>(mine) + threadIdx.x, seq);
  }
  __syncthreads();
  if (op == 3) {
    for (int r = 0; r < nranks; ++r) {
      const double* g = reinterpret_cast<const double*>(mine + p2p_gather_off(nranks, parity, r));
      for (int i = threadIdx.x; i < count; i += blockDim.x) out[(int64_t)r * count + i] = ld_volatile(g + i);
    }
    return;
  }
  for (int i = threadIdx.x; i < count; i += blockDim.x) {
    double a = ld_volatile(reinterpret_cast<const double*>(mine + p2p_gather_off(nranks, parity, 0)) + i);
    const bool is_max = (op == 1) || (op == 2 && i == 1);
    for (int r = 1; r < nranks; ++r) {
      const double b = ld_volatile(reinterpret_cast<const double*>(mine + p2p_gather_off(nranks, parity, r)) + i);
      a = is_max ? fmax(a, b) : a + b;
    }
    out[i] = a;
  }
}

// own0: first owned double of the stored column; cnt = depth * m doubles per message
__global__ void __launch_bounds__(1024) p2p_halo_kernel(void* const* __restrict__ peers, int rank, int nranks,
                                                         double* __restrict__ own0, int64_t cnt, int64_t rows_m,
                                                         int has_lo, int has_hi, unsigned long long seq) {
  const int parity = (int)(seq & 1ull);
  // my first rows become the lower neighbour's upper halo (its side 1), my last rows the upper neighbour's side 0
  if (has_lo) {
    double* dst = reinterpret_cast<double*>(static_cast<char*>(peers[rank - 1]) + p2p_halo_off(nranks, parity, 1));
    for (int64_t i = threadIdx.x; i < cnt; i += blockDim.x) dst[i] = own0[i];
  }
  if (has_hi) {
    double* dst = reinterpret_cast<double*>(static_cast<char*>(peers[rank + 1]) + p2p_halo_off(nranks, parity, 0));
    const double* srcp = own0 + rows_m - cnt;
    for (int64_t i = threadIdx.x; i < cnt; i += blockDim.x) dst[i] = srcp[i];
  }
  __threadfence_system();
  __syncthreads();
  char* mine = static_cast<char*>(peers[rank]);
  const unsigned long long* myflags = reinterpret_cast<const unsigned long long*>(mine + 1024);
  if (threadIdx.x == 0 && has_lo) {
    st_release_sys(reinterpret_cast<unsigned long long*>(static_cast<char*>(peers[rank - 1]) + 1024) + 1, seq);
    wait_flag(myflags + 0, seq);
  }
  if (threadIdx.x == 32 && has_hi) {
    st_release_sys(reinterpret_cast<unsigned long long*>(static_cast<char*>(peers[rank + 1]) + 1024) + 0, seq);
    wait_flag(myflags + 1, seq);
  }
  __syncthreads();
  if (has_lo) {
    const double* g = reinterpret_cast<const double*>(mine + p2p_halo_off(nranks, parity, 0));
    for (int64_t i = threadIdx.x; i < cnt; i += blockDim.x) own0[i - cnt] = ld_volatile(g + i);
  }
  if (has_hi) {
    const double* g = reinterpret_cast<const double*>(mine + p2p_halo_off(nranks, parity, 1));
    for (int64_t i = threadIdx.x; i < cnt; i += blockDim.x) own0[rows_m + i] = ld_volatile(g + i);
  }
}

int p2p_gather(gnk_ctx* ctx, const double* d_src, int64_t count, int op, double* d_out, cudaStream_t st) {
  ctx->p2p_seq++;
  const int threads = count >= 1024 ? 1024 : (count > 256 ? 512 : 256);
  p2p_gather_kernel<<<1, threads, 0, st>>>(ctx->d_p2p_peer, ctx->rank, ctx->nranks, d_src, (int)count, op, ctx->p2p_seq,
                                           d_out);
  GNK_LAUNCH_CHECK(ctx);
  return 0;
}

}  // namespace

int gnk_comm_allgather_doubles(gnk_ctx* ctx, const double* d_send, double* d_recv, int64_t count, void* stream) {
  if (ctx->p2p_ready && count <= P2P_GMAX) return p2p_gather(ctx, d_send, count, 3, d_recv, (cudaStream_t)stream);
  GNK_REQUIRE(ctx->nccl_comm, "all-gather without a communicator");
  GNK_NCCL(g_nccl.AllGather(d_send, d_recv, (size_t)count, NCCL_F64, (NcclComm)ctx->nccl_comm, (cudaStream_t)stream));
  return 0;
}

void gnk_comm_teardown(gnk_ctx* ctx) {
  if (ctx->p2p_local) {
    cudaDeviceSynchronize();
    for (int r = 0; r < ctx->nranks && r < P2P_MAXR; ++r)
      if (r != ctx->rank && ctx->p2p_peer[r]) cudaIpcCloseMemHandle(ctx->p2p_peer[r]);
    if (ctx->d_p2p_peer) cudaFree(ctx->d_p2p_peer);
    cudaFree(ctx->p2p_local);
    ctx->p2p_local = nullptr;
    ctx->d_p2p_peer = nullptr;
    ctx->p2p_ready = 0;
  }
  if (ctx->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy((NcclComm)ctx->nccl_comm);
  ctx->nccl_comm = nullptr;
}

extern "C" {

int gnk_comm_unique_id(void* out128) {
  GNK_REQUIRE(out128, "gnk_comm_unique_id: null argument");
  if (int rc = load_nccl()) return rc;
  NcclId id;
  GNK_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(out128, &id, sizeof(id));
  return 0;
}

int gnk_comm_init(gnk_ctx* ctx, const void* id128, int rank, int nranks) {
  GNK_REQUIRE(ctx && id128, "gnk_comm_init: null argument");
  GNK_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "gnk_comm_init: bad rank");
  GNK_REQUIRE(!ctx->nccl_comm, "gnk_comm_init: communicator already attached");
  if (int rc = load_nccl()) return rc;
  GNK_CUDA(cudaSetDevice(ctx->device));
  NcclId id;
  memcpy(&id, id128, sizeof(id));
  NcclComm comm = nullptr;
  GNK_NCCL(g_nccl.CommInitRank(&comm, nranks, id, rank));
  ctx->nccl_comm = comm;
  ctx->rank = rank;
  ctx->nranks = nranks;
  return 0;
}

int gnk_comm_p2p_export(gnk_ctx* ctx, void* out_handle64) {
  GNK_REQUIRE(ctx && out_handle64, "gnk_comm_p2p_export: null argument");
  GNK_REQUIRE(ctx->nranks >= 2 && ctx->nranks <= P2P_MAXR, "gnk_comm_p2p_export: needs 2..16 ranks (call gnk_comm_init first)");
  GNK_REQUIRE(!ctx->p2p_local, "gnk_comm_p2p_export: mailbox already exported");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  GNK_CUDA(cudaSetDevice(ctx->device));
  const size_t bytes = p2p_bytes(ctx->nranks);
  GNK_CUDA(cudaMalloc(&ctx->p2p_local, bytes));
  GNK_CUDA(cudaMemset(ctx->p2p_local, 0, bytes));
  GNK_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  GNK_CUDA(cudaIpcGetMemHandle(&h, ctx->p2p_local));
  memcpy(out_handle64, &h, sizeof(h));
  return 0;
}

int gnk_comm_p2p_attach(gnk_ctx* ctx, const void* handles) {
  GNK_REQUIRE(ctx && handles && ctx->p2p_local, "gnk_comm_p2p_attach: export the local mailbox first");
  GNK_CUDA(cudaSetDevice(ctx->device));
  for (int r = 0; r < ctx->nranks; ++r) {
    if (r == ctx->rank) {
      ctx->p2p_peer[r] = ctx->p2p_local;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const char*>(handles) + 64 * (size_t)r, sizeof(h));
    GNK_CUDA(cudaIpcOpenMemHandle(&ctx->p2p_peer[r], h, cudaIpcMemLazyEnablePeerAccess));
  }
  GNK_CUDA(cudaMalloc(&ctx->d_p2p_peer, sizeof(void*) * P2P_MAXR));
  GNK_CUDA(cudaMemcpy(ctx->d_p2p_peer, ctx->p2p_peer, sizeof(void*) * P2P_MAXR, cudaMemcpyHostToDevice));
  ctx->p2p_seq = 0;
  ctx->p2p_hseq = 0;
  ctx->p2p_ready = 1;
  ctx->p2p_fused = !(getenv("GNK_P2P_FUSED") && atoi(getenv("GNK_P2P_FUSED")) == 0);
  return 0;
}

int gnk_comm_p2p_enabled(gnk_ctx* ctx) { return ctx ? ctx->p2p_ready : 0; }
int gnk_comm_p2p_disable(gnk_ctx* ctx) {
  if (ctx) ctx->p2p_ready = ctx->p2p_fused = 0;
  return 0;
}
int gnk_comm_fused_reductions(gnk_ctx* ctx) { return (ctx && ctx->p2p_ready && ctx->p2p_fused) ? 1 : 0; }

int gnk_comm_size(gnk_ctx* ctx) { return ctx ? ctx->nranks : 0; }

int gnk_comm_allreduce(gnk_ctx* ctx, double* d_buf, int count, int op, void* stream) {
  GNK_REQUIRE(ctx && d_buf, "gnk_comm_allreduce: null argument");
  GNK_REQUIRE(count >= 1 && count <= 256 && op >= 0 && op <= 2, "gnk_comm_allreduce: bad count/op");
  if (ctx->nranks == 1) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (ctx->p2p_ready) return p2p_gather(ctx, d_buf, count, op, d_buf, st);
  if (int rc = ensure_gather(ctx, sizeof(double) * 256 * (size_t)ctx->nranks, st)) return rc;
  if (int rc = gnk_comm_allgather_doubles(ctx, d_buf, ctx->d_gather, count, stream)) return rc;
  rank_reduce_kernel<<<(count + 127) / 128, 128, 0, st>>>(ctx->d_gather, ctx->nranks, count, op, d_buf);
  GNK_LAUNCH_CHECK(ctx);
  return 0;
}

int gnk_comm_halo_exchange(gnk_ctx* ctx, const gnk_layout* lay, double* d_col, int depth, void* stream) {
  GNK_REQUIRE(ctx && lay && d_col, "gnk_comm_halo_exchange: null argument");
  GNK_REQUIRE(depth >= 1 && depth <= lay->halo && depth <= lay->rows, "gnk_comm_halo_exchange: bad depth");
  if (ctx->nranks == 1) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t cnt = (size_t)depth * lay->m;
  double* own0 = d_col + lay->off;
  if (ctx->p2p_ready && (int64_t)cnt <= P2P_HMAX) {
    ctx->p2p_hseq++;
    p2p_halo_kernel<<<1, 1024, 0, st>>>(ctx->d_p2p_peer, ctx->rank, ctx->nranks, own0, (int64_t)cnt,
                                        (int64_t)lay->rows * lay->m, lay->has_lo, lay->has_hi, ctx->p2p_hseq);
    GNK_LAUNCH_CHECK(ctx);
    return 0;
  }
  GNK_REQUIRE(ctx->nccl_comm, "gnk_comm_halo_exchange: no communicator");
  NcclComm comm = (NcclComm)ctx->nccl_comm;
  GNK_NCCL(g_nccl.GroupStart());
  if (lay->has_lo) {
    GNK_NCCL(g_nccl.Send(own0, cnt, NCCL_F64, ctx->rank - 1, comm, st));
    GNK_NCCL(g_nccl.Recv(own0 - cnt, cnt, NCCL_F64, ctx->rank - 1, comm, st));
  }
  if (lay->has_hi) {
    GNK_NCCL(g_nccl.Send(own0 + (size_t)(lay->rows - depth) * lay->m, cnt, NCCL_F64, ctx->rank + 1, comm, st));
    GNK_NCCL(g_nccl.Recv(own0 + (size_t)lay->rows * lay->m, cnt, NCCL_F64, ctx->rank + 1, comm, st));
  }
  GNK_NCCL(g_nccl.GroupEnd());
  return 0;
}

int gnk_comm_allgather_owned(gnk_ctx* ctx, const gnk_layout* lay, const double* d_col, double* d_full,
                             const int64_t* counts, void* stream) {
  GNK_REQUIRE(ctx && lay && d_col && d_full, "gnk_comm_allgather_owned: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (ctx->nranks == 1) {
    GNK_CUDA(cudaMemcpyAsync(d_full, d_col + lay->off, sizeof(double) * lay->n_own, cudaMemcpyDeviceToDevice, st));
    return 0;
  }
  GNK_REQUIRE(ctx->nccl_comm && counts, "gnk_comm_allgather_owned: no communicator / counts");
  NcclComm comm = (NcclComm)ctx->nccl_comm;
  GNK_NCCL(g_nccl.GroupStart());
  int64_t displ = 0;
  for (int r = 0; r < ctx->nranks; ++r) {
    GNK_NCCL(g_nccl.Broadcast(d_col + lay->off, d_full + displ, (size_t)counts[r], NCCL_F64, r, comm, st));
    displ += counts[r];
  }
  GNK_NCCL(g_nccl.GroupEnd());
  return 0;
}

}  // extern "C"
