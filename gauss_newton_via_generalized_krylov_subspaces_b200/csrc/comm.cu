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

}  // namespace

int gnk_comm_allgather_doubles(gnk_ctx* ctx, const double* d_send, double* d_recv, int64_t count, void* stream) {
  GNK_REQUIRE(ctx->nccl_comm, "all-gather without a communicator");
  GNK_NCCL(g_nccl.AllGather(d_send, d_recv, (size_t)count, NCCL_F64, (NcclComm)ctx->nccl_comm, (cudaStream_t)stream));
  return 0;
}

void gnk_comm_teardown(gnk_ctx* ctx) {
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

int gnk_comm_size(gnk_ctx* ctx) { return ctx ? ctx->nranks : 0; }

int gnk_comm_allreduce(gnk_ctx* ctx, double* d_buf, int count, int op, void* stream) {
  GNK_REQUIRE(ctx && d_buf, "gnk_comm_allreduce: null argument");
  GNK_REQUIRE(count >= 1 && count <= 256 && op >= 0 && op <= 2, "gnk_comm_allreduce: bad count/op");
  if (ctx->nranks == 1) return 0;
  cudaStream_t st = (cudaStream_t)stream;
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
  GNK_REQUIRE(ctx->nccl_comm, "gnk_comm_halo_exchange: no communicator");
  cudaStream_t st = (cudaStream_t)stream;
  NcclComm comm = (NcclComm)ctx->nccl_comm;
  const size_t cnt = (size_t)depth * lay->m;
  double* own0 = d_col + lay->off;
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
