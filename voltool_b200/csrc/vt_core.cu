// vt_core.cu -- process-global context, device memory, resident maps.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "vt_device.cuh"

namespace vt {

static Ctx g_ctx;
static char g_err[1024] = "";

Ctx &ctx() { return g_ctx; }
bool ctx_ready() { return g_ctx.device >= 0; }

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  if (getenv("VOLTOOL_VERBOSE")) fprintf(stderr, "[voltool_gpu] %s\n", g_err);
}

static cudaError_t g_last_cuda = cudaSuccess;
cudaError_t last_cuda_error() { return g_last_cuda; }

int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
  g_last_cuda = e;
  set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
  return 2;
}

// ---- per-kernel event timing -----------------------------------------------------------------
struct ProfSpan { cudaEvent_t a, b; int key; bool closed; };
static bool g_prof_on = false;
static std::vector<ProfSpan> g_spans;
static std::vector<cudaEvent_t> g_event_pool;
static const char *kProfNames[PK_COUNT] = {"transform", "setup", "splat", "pose_cc", "pose_finish", "cc", "stats",
                                           "unary", "binary", "blur", "resample"};

static cudaEvent_t take_event() {
  if (!g_event_pool.empty()) { cudaEvent_t e = g_event_pool.back(); g_event_pool.pop_back(); return e; }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
NvtxScope::NvtxScope(const char *name) { nvtxRangePushA(name); }
NvtxScope::~NvtxScope() { nvtxRangePop(); }

void prof_begin(int key) {
  nvtxRangePushA(kProfNames[key]);   // kernel-class range (splat, pose_cc, blur, ...) inside the entry point's range
  if (!g_prof_on) return;
  ProfSpan s{take_event(), take_event(), key, false};
  cudaEventRecord(s.a, g_ctx.stream);
  g_spans.push_back(s);
}
void prof_end(int key) {
  nvtxRangePop();
  if (!g_prof_on) return;
  for (size_t i = g_spans.size(); i-- > 0;)
    if (g_spans[i].key == key && !g_spans[i].closed) {
      cudaEventRecord(g_spans[i].b, g_ctx.stream);
      g_spans[i].closed = true;
      return;
    }
}

int dev_alloc(void **p, size_t bytes) {
  if (bytes == 0) bytes = 16;
  VT_CUDA(cudaMallocAsync(p, bytes, g_ctx.stream));
  return 0;
}
int dev_free(void *p) {
  if (p) VT_CUDA(cudaFreeAsync(p, g_ctx.stream));
  return 0;
}

}  // namespace vt

using namespace vt;

extern "C" {

const char *vt_last_error(void) { return g_err; }

int vt_init(int device) {
  if (ctx_ready()) return 0;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0) {
    set_error("voltool_gpu: no CUDA device (%s); this library has no CPU fallback",
              e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    return 1;
  }
  if (device < 0) {
    const char *s = getenv("VOLTOOL_GPU_DEVICE");
    if (!s) s = getenv("LOCAL_RANK");
    device = s ? atoi(s) : 0;
    if (device >= count) device = device % count;
  }
  if (device >= count) { set_error("voltool_gpu: device %d requested, %d present", device, count); return 1; }
  VT_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  VT_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) {
    set_error("voltool_gpu: built for sm_100a (Blackwell); device %d is sm_%d%d", device, prop.major, prop.minor);
    return 1;
  }
  Ctx &c = ctx();
  c.sm_count = prop.multiProcessorCount;
  VT_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
  // keep freed blocks in the pool: the ops allocate/free map-sized buffers per call
  cudaMemPool_t pool;
  VT_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
  unsigned long long keep = ~0ull;
  VT_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  c.scratch_bytes = (size_t)8 << 20;
  VT_CUDA(cudaMalloc((void **)&c.d_scratch, c.scratch_bytes));
  VT_CUDA(cudaMalloc((void **)&c.d_ticket, 64 * sizeof(unsigned int)));
  VT_CUDA(cudaMemset(c.d_ticket, 0, 64 * sizeof(unsigned int)));
  c.pinned_bytes = (size_t)1 << 20;
  VT_CUDA(cudaMallocHost((void **)&c.h_pinned, c.pinned_bytes));
  c.launches = 0;
  c.device = device;
  return 0;
}

int vt_shutdown(void) {
  Ctx &c = ctx();
  if (!ctx_ready()) return 0;
  cudaStreamSynchronize(c.stream);
  cudaFree(c.d_scratch);
  cudaFree(c.d_ticket);
  cudaFreeHost(c.h_pinned);
  cudaStreamDestroy(c.stream);
  c = Ctx();
  return 0;
}

int vt_sync(void) {
  VT_REQUIRE_CTX();
  VT_CUDA(cudaStreamSynchronize(ctx().stream));
  return 0;
}

long vt_launch_count(void) { return ctx().launches; }
void *vt_stream(void) { return (void *)ctx().stream; }

int vt_device_props(int *sm_count, int *clock_khz, size_t *mem_bytes, char *name, int name_len) {
  VT_REQUIRE_CTX();
  cudaDeviceProp prop;
  VT_CUDA(cudaGetDeviceProperties(&prop, ctx().device));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (clock_khz) {
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, ctx().device);
    *clock_khz = khz;
  }
  if (mem_bytes) *mem_bytes = prop.totalGlobalMem;
  if (name && name_len > 0) { strncpy(name, prop.name, name_len - 1); name[name_len - 1] = 0; }
  return 0;
}

// ------------------------------------------------------------------------------------------ maps
static int map_alloc(const vt_grid *g, vt_map **out) {
  if (!g || !out) { set_error("vt_map_create: null argument"); return 1; }
  if (g->xsize < 0 || g->ysize < 0 || g->zsize < 0) { set_error("vt_map_create: negative size"); return 1; }
  vt_map *m = new vt_map();
  m->grid = *g;
  m->n = gridsize(*g);
  memset(&m->cache, 0, sizeof(m->cache));
  if (dev_alloc((void **)&m->d, sizeof(float) * (size_t)m->n)) { delete m; return 2; }
  *out = m;
  return 0;
}

// compute_minmax(), VolumetricData.C:1063-1074
static int map_compute_minmax(vt_map *m) {
  m->cache.cmin = 0; m->cache.cmax = 0;
  if (m->n > 0) {
    int rc = k_minmax(m->d, m->n, &m->cache.cmin, &m->cache.cmax);
    if (rc) return rc;
  }
  m->cache.minmax_valid = 1;
  return 0;
}

int vt_map_create(const vt_grid *g, const float *host, vt_map **out) {
  VT_REQUIRE_CTX();
  int rc = map_alloc(g, out);
  if (rc) return rc;
  vt_map *m = *out;
  if (host) VT_CUDA(cudaMemcpyAsync(m->d, host, sizeof(float) * (size_t)m->n, cudaMemcpyHostToDevice, ctx().stream));
  else VT_CUDA(cudaMemsetAsync(m->d, 0, sizeof(float) * (size_t)(m->n > 0 ? m->n : 4), ctx().stream));
  return map_compute_minmax(m);
}

// the resident twin of an EXISTING VolumetricData: its voxels and its cached-statistics block as they stand
// (no compute_minmax: the object was constructed long ago and its caches carry history, VolumetricData.h:202-210)
int vt_map_adopt_host(const vt_grid *g, const float *host, const vt_cache *cache, vt_map **out) {
  VT_REQUIRE_CTX();
  if (!host || !cache) { set_error("vt_map_adopt_host: null argument"); return 1; }
  int rc = map_alloc(g, out);
  if (rc) return rc;
  vt_map *m = *out;
  m->cache = *cache;
  if (m->n > 0) VT_CUDA(cudaMemcpyAsync(m->d, host, sizeof(float) * (size_t)m->n, cudaMemcpyHostToDevice, ctx().stream));
  return 0;
}

int vt_map_create_from_device(const vt_grid *g, const float *dev, vt_map **out) {
  VT_REQUIRE_CTX();
  int rc = map_alloc(g, out);
  if (rc) return rc;
  vt_map *m = *out;
  VT_CUDA(cudaMemcpyAsync(m->d, dev, sizeof(float) * (size_t)m->n, cudaMemcpyDeviceToDevice, ctx().stream));
  return map_compute_minmax(m);
}

int vt_map_clone(const vt_map *src, vt_map **out) {
  VT_REQUIRE_CTX();
  int rc = map_alloc(&src->grid, out);
  if (rc) return rc;
  (*out)->cache = src->cache;
  VT_CUDA(cudaMemcpyAsync((*out)->d, src->d, sizeof(float) * (size_t)src->n, cudaMemcpyDeviceToDevice, ctx().stream));
  return 0;
}

int vt_map_destroy(vt_map *m) {
  if (!m) return 0;
  VT_REQUIRE_CTX();
  dev_free(m->d);
  delete m;
  return 0;
}

int vt_map_grid(const vt_map *m, vt_grid *g) { *g = m->grid; return 0; }
int vt_map_cache(const vt_map *m, vt_cache *c) { *c = m->cache; return 0; }
long vt_map_gridsize(const vt_map *m) { return m->n; }
float *vt_map_device_ptr(vt_map *m) { return m->d; }

int vt_map_download(const vt_map *m, float *host) {
  VT_REQUIRE_CTX();
  VT_CUDA(cudaMemcpyAsync(host, m->d, sizeof(float) * (size_t)m->n, cudaMemcpyDeviceToHost, ctx().stream));
  VT_CUDA(cudaStreamSynchronize(ctx().stream));
  return 0;
}

int vt_map_upload(vt_map *m, const float *host) {
  VT_REQUIRE_CTX();
  VT_CUDA(cudaMemcpyAsync(m->d, host, sizeof(float) * (size_t)m->n, cudaMemcpyHostToDevice, ctx().stream));
  m->cache.mean_valid = 0; m->cache.cmean = 0;
  m->cache.sigma_valid = 0; m->cache.csigma = 0;
  return map_compute_minmax(m);
}

void vt_free_host(void *p) { free(p); }

int vt_profile_enable(int on) {
  g_prof_on = on != 0;
  return 0;
}
int vt_profile_reset(void) {
  for (auto &s : g_spans) { g_event_pool.push_back(s.a); g_event_pool.push_back(s.b); }
  g_spans.clear();
  return 0;
}
int vt_profile_get(const char *name, double *ms_total, long *count) {
  VT_REQUIRE_CTX();
  VT_CUDA(cudaStreamSynchronize(ctx().stream));
  int key = -1;
  for (int k = 0; k < PK_COUNT; ++k)
    if (!strcmp(name, kProfNames[k])) key = k;
  if (key < 0) { set_error("vt_profile_get: unknown kernel class %s", name); return 1; }
  double total = 0.0;
  long n = 0;
  for (auto &s : g_spans) {
    if (s.key != key || !s.closed) continue;
    float ms = 0.f;
    VT_CUDA(cudaEventElapsedTime(&ms, s.a, s.b));
    total += ms;
    ++n;
  }
  if (ms_total) *ms_total = total;
  if (count) *count = n;
  return 0;
}

}  // extern "C"
