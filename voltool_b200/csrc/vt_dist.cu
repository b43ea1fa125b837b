// vt_dist.cu -- the one collective of the path, inside the library: the per-stage CC tables of a sharded fit
// (rank r evaluates candidates r, r + world, ... of every rotate() table, TclVoltool.C:160-176) are merged with
// ONE ncclAllGather on device buffers over NVLink, and the per-shard best (cc, pose index) of an exhaustive pose
// list with one more.  NCCL is bound at run time (dlopen of libnccl.so.2: the copy already in the process when the
// caller is a torch program, the system one for a C++ VMD caller), so the library still loads where NCCL is
// absent; the unique id travels by whatever the caller has (MPI, a file, torch.distributed).
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <vector>

#include "vt_device.cuh"
#include "vt_host.h"

namespace vt {

namespace {

struct Nccl {
  void *so = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  float *d_send = nullptr, *d_recv = nullptr;   // table slices
  long cap = 0;                                 // floats per rank
  float *h_recv = nullptr;                      // pinned
  long h_cap = 0;
};
Nccl g_nccl;

int nccl_load() {
  Nccl &n = g_nccl;
  if (n.so) return 0;
  const char *names[] = {getenv("VOLTOOL_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char *nm : names) {
    if (!nm || !*nm) continue;
    n.so = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (n.so) break;
  }
  if (!n.so) { set_error("vt_nccl: cannot load libnccl.so.2 (%s)", dlerror()); return 1; }
#define VT_SYM(field, name)                                                             \
  n.field = (decltype(n.field))dlsym(n.so, name);                                       \
  if (!n.field) { set_error("vt_nccl: symbol %s missing", name); dlclose(n.so); n.so = nullptr; return 1; }
  VT_SYM(GetUniqueId, "ncclGetUniqueId")
  VT_SYM(CommInitRank, "ncclCommInitRank")
  VT_SYM(CommDestroy, "ncclCommDestroy")
  VT_SYM(AllGather, "ncclAllGather")
  VT_SYM(GetErrorString, "ncclGetErrorString")
#undef VT_SYM
  return 0;
}

int nccl_fail(ncclResult_t r, const char *what) {
  set_error("NCCL error %d (%s) in %s", (int)r, g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?", what);
  return 4;
}

int ensure_buffers(long per) {
  Nccl &n = g_nccl;
  if (per > n.cap) {
    dev_free(n.d_send); dev_free(n.d_recv);
    n.d_send = n.d_recv = nullptr;
    if (dev_alloc((void **)&n.d_send, sizeof(float) * (size_t)per)) return 2;
    if (dev_alloc((void **)&n.d_recv, sizeof(float) * (size_t)per * n.world)) return 2;
    n.cap = per;
  }
  if (per * n.world > n.h_cap) {
    if (n.h_recv) cudaFreeHost(n.h_recv);
    n.h_recv = nullptr;
    VT_CUDA(cudaMallocHost((void **)&n.h_recv, sizeof(float) * (size_t)per * n.world));
    n.h_cap = per * n.world;
  }
  return 0;
}

// deterministic arg max of a CC list on the device: NaN entries (poses whose CC is undefined) never win, ties go
// to the lowest index -- the order a sequential `if (cc > best)` scan would give (TclVoltool.C:181)
__global__ void __launch_bounds__(1024) best_pose_kernel(const float *__restrict__ cc, long n, long index_base,
                                                         float *__restrict__ out2) {
  __shared__ float sv[32];
  __shared__ long si[32];
  float bv = -INFINITY;
  long bi = -1;
  for (long i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = cc[i];
    if (v > bv || (bi < 0 && v == v)) { bv = v; bi = i; }     // ascending i per thread: first maximum kept
  }
  auto better = [](float v, long i, float w, long j) { return j < 0 ? false : (i < 0 ? true : (w > v || (w == v && j < i))); };
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float w = __shfl_xor_sync(0xffffffffu, bv, o);
    const long j = __shfl_xor_sync(0xffffffffu, bi, o);
    if (better(bv, bi, w, j)) { bv = w; bi = j; }
  }
  if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = bv; si[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x < 32) {
    bv = threadIdx.x < (blockDim.x >> 5) ? sv[threadIdx.x] : -INFINITY;
    bi = threadIdx.x < (blockDim.x >> 5) ? si[threadIdx.x] : -1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float w = __shfl_xor_sync(0xffffffffu, bv, o);
      const long j = __shfl_xor_sync(0xffffffffu, bi, o);
      if (better(bv, bi, w, j)) { bv = w; bi = j; }
    }
    if (threadIdx.x == 0) {
      out2[0] = bi < 0 ? __int_as_float(0x7fc00000) : bv;
      out2[1] = bi < 0 ? -1.0f : (float)(index_base + bi);
    }
  }
}

}  // namespace

bool nccl_active() { return g_nccl.comm != nullptr; }

// complete table[n] on every rank from the interleaved device slices: rank r holds entries r, r + world, ...
// packed in d_slice[0 .. count_r).  One ncclAllGather on the library stream, one D2H of the merged table.
int nccl_merge_table(const float *d_slice, float *table, int n) {
  Nccl &nc = g_nccl;
  Ctx &c = ctx();
  const long per = (n + nc.world - 1) / nc.world;
  int rc = ensure_buffers(per);
  if (rc) return rc;
  const long mine = (n - nc.rank + nc.world - 1) / nc.world;
  VT_CUDA(cudaMemsetAsync(nc.d_send, 0xff, sizeof(float) * (size_t)per, c.stream));   // NaN padding
  if (mine > 0) VT_CUDA(cudaMemcpyAsync(nc.d_send, d_slice, sizeof(float) * (size_t)mine, cudaMemcpyDeviceToDevice, c.stream));
  ncclResult_t r = nc.AllGather(nc.d_send, nc.d_recv, (size_t)per, ncclFloat, nc.comm, c.stream);
  if (r != ncclSuccess) return nccl_fail(r, "ncclAllGather");
  VT_CUDA(cudaMemcpyAsync(nc.h_recv, nc.d_recv, sizeof(float) * (size_t)per * nc.world, cudaMemcpyDeviceToHost, c.stream));
  VT_CUDA(cudaStreamSynchronize(c.stream));
  for (int r2 = 0; r2 < nc.world; ++r2) {
    long k = 0;
    for (long i = r2; i < n; i += nc.world) table[i] = nc.h_recv[(size_t)r2 * per + k++];
  }
  return 0;
}

}  // namespace vt

using namespace vt;

extern "C" {

int vt_nccl_unique_id(void *id_out, int len) {
  if (!id_out || len < (int)sizeof(ncclUniqueId)) { set_error("vt_nccl_unique_id: need %d bytes", (int)sizeof(ncclUniqueId)); return 1; }
  if (nccl_load()) return 1;
  ncclUniqueId id;
  ncclResult_t r = g_nccl.GetUniqueId(&id);
  if (r != ncclSuccess) return nccl_fail(r, "ncclGetUniqueId");
  memcpy(id_out, &id, sizeof(id));
  return 0;
}

int vt_nccl_init(int rank, int world, const void *id, int len) {
  VT_REQUIRE_CTX();
  if (world < 1 || rank < 0 || rank >= world || !id || len < (int)sizeof(ncclUniqueId)) { set_error("vt_nccl_init: bad arguments"); return 1; }
  if (nccl_load()) return 1;
  Nccl &n = g_nccl;
  if (n.comm) { set_error("vt_nccl_init: already initialised"); return 1; }
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof(uid));
  ncclResult_t r = n.CommInitRank(&n.comm, world, uid, rank);
  if (r != ncclSuccess) { n.comm = nullptr; return nccl_fail(r, "ncclCommInitRank"); }
  n.rank = rank; n.world = world;
  return vt_fit_set_shard(rank, world, nullptr, nullptr);   // the pose search now shards; tables merge over NCCL
}

int vt_nccl_shutdown(void) {
  Nccl &n = g_nccl;
  if (!n.comm) return 0;
  if (ctx_ready()) {
    cudaStreamSynchronize(ctx().stream);
    dev_free(n.d_send); dev_free(n.d_recv);
  }
  if (n.h_recv) cudaFreeHost(n.h_recv);
  n.d_send = n.d_recv = n.h_recv = nullptr;
  n.cap = n.h_cap = 0;
  n.CommDestroy(n.comm);
  n.comm = nullptr;
  n.rank = 0; n.world = 1;
  vt_fit_set_shard(0, 1, nullptr, nullptr);
  return 0;
}

int vt_fit_best_pose_device(const float *d_cc, long n, long index_base, float *d_best2) {
  VT_REQUIRE_CTX();
  if (!d_cc || !d_best2 || n < 0) { set_error("vt_fit_best_pose_device: bad arguments"); return 1; }
  best_pose_kernel<<<1, 1024, 0, ctx().stream>>>(d_cc, n, index_base, d_best2);
  VT_LAUNCHED();
  return 0;
}

int vt_nccl_allgather(const float *d_send, float *d_recv, long count) {
  VT_REQUIRE_CTX();
  Nccl &n = g_nccl;
  if (!n.comm) {   // one rank: the gather is a copy
    VT_CUDA(cudaMemcpyAsync(d_recv, d_send, sizeof(float) * (size_t)count, cudaMemcpyDeviceToDevice, ctx().stream));
    return 0;
  }
  ncclResult_t r = n.AllGather(d_send, d_recv, (size_t)count, ncclFloat, n.comm, ctx().stream);
  return r == ncclSuccess ? 0 : nccl_fail(r, "ncclAllGather");
}

}  // extern "C"
