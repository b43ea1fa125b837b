// vt_stream.cu -- HBM-streaming kernels: statistics (min/max/mean/sigma), the in-place unary ops,
// histogram and vol_com.  All are bandwidth-bound: float4 loads/stores, grid = multiple of the
// SM count, deterministic two-level reductions (block partials + ordered finish by the last block).
#include <cfloat>
#include <cmath>

#include "vt_device.cuh"

namespace vt {

constexpr int kThreads = 256;
constexpr int kBlocksPerSM = 8;

static inline int stream_grid(long n_vec) {
  long want = (n_vec + kThreads - 1) / kThreads;
  long cap = (long)ctx().sm_count * kBlocksPerSM;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return (int)want;
}

__device__ __forceinline__ float4 ld_stream(const float4 *p) { return __ldcs(p); }

// ------------------------------------------------------------------------------------------ stats
// MODE 0: min,max   1: sum   2: min,max,sum   3: sum of (x-mean)^2 (float delta, float square,
// double accumulate -- compute_sigma, VolumetricData.C:1127-1147)
struct StatOut {
  float mn, mx;
  double sum;
};

template <int MODE>
__global__ void __launch_bounds__(kThreads) stats_kernel(const float *__restrict__ d, long n, float mean,
                                                         double *partials, unsigned int *ticket, StatOut *out) {
  __shared__ double sm[3 * 32];
  float mn = INFINITY, mx = -INFINITY;
  double s0 = 0.0, s1 = 0.0;  // two chains for ILP
  const long nv = n >> 2;
  const float4 *v = reinterpret_cast<const float4 *>(d);
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
    const float4 q = ld_stream(v + i);
    if (MODE == 0 || MODE == 2) {
      mn = fminf(fminf(mn, q.x), fminf(q.y, fminf(q.z, q.w)));
      mx = fmaxf(fmaxf(mx, q.x), fmaxf(q.y, fmaxf(q.z, q.w)));
    }
    if (MODE == 1 || MODE == 2) {
      s0 += (double)q.x; s1 += (double)q.y; s0 += (double)q.z; s1 += (double)q.w;
    }
    if (MODE == 3) {
      const float a = q.x - mean, b = q.y - mean, c = q.z - mean, e = q.w - mean;
      s0 += (double)(a * a); s1 += (double)(b * b); s0 += (double)(c * c); s1 += (double)(e * e);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {  // scalar tail
    const float x = d[(nv << 2) + threadIdx.x];
    if (MODE == 0 || MODE == 2) { mn = fminf(mn, x); mx = fmaxf(mx, x); }
    if (MODE == 1 || MODE == 2) s0 += (double)x;
    if (MODE == 3) { const float a = x - mean; s0 += (double)(a * a); }
  }
  // block reduce
  double s = s0 + s1;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  mn = warp_min(mn); mx = warp_max(mx); s = warp_sum(s);
  if (lane == 0) { sm[warp] = (double)mn; sm[32 + warp] = (double)mx; sm[64 + warp] = s; }
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) {
    double bmn = sm[0], bmx = sm[32], bs = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) { bmn = fmin(bmn, sm[w]); bmx = fmax(bmx, sm[32 + w]); bs += sm[64 + w]; }
    partials[(size_t)blockIdx.x * 3 + 0] = bmn;
    partials[(size_t)blockIdx.x * 3 + 1] = bmx;
    partials[(size_t)blockIdx.x * 3 + 2] = bs;
    __threadfence();
    last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  // ordered finish by one warp: lane l takes blocks l, l+32, ...; then a fixed shuffle tree
  if (warp == 0) {
    double fmn = INFINITY, fmx = -INFINITY, fs = 0.0;
    for (unsigned int b = lane; b < gridDim.x; b += 32) {
      fmn = fmin(fmn, partials[(size_t)b * 3 + 0]);
      fmx = fmax(fmx, partials[(size_t)b * 3 + 1]);
      fs += partials[(size_t)b * 3 + 2];
    }
    for (int o = 16; o > 0; o >>= 1) {
      fmn = fmin(fmn, __shfl_xor_sync(0xffffffffu, fmn, o));
      fmx = fmax(fmx, __shfl_xor_sync(0xffffffffu, fmx, o));
      fs += __shfl_xor_sync(0xffffffffu, fs, o);
    }
    if (lane == 0) {
      float omn = (float)fmn, omx = (float)fmx;
      // the sequential scan seeds min/max with element 0: a NaN there poisons both
      if ((MODE == 0 || MODE == 2) && n > 0 && is_nan_bits(d[0])) { omn = d[0]; omx = d[0]; }
      out->mn = omn; out->mx = omx; out->sum = fs;
      *ticket = 0u;
    }
  }
}

template <int MODE>
static int run_stats(const float *d, long n, float mean, StatOut *host_out) {
  VT_REQUIRE_CTX();
  Ctx &c = ctx();
  const int grid = stream_grid(n >> 2);
  StatOut *dout = reinterpret_cast<StatOut *>(c.d_scratch);
  double *partials = c.d_scratch + 8;
  prof_begin(PK_STATS);
  stats_kernel<MODE><<<grid, kThreads, 0, c.stream>>>(d, n, mean, partials, c.d_ticket, dout);
  prof_end(PK_STATS);
  VT_LAUNCHED();
  StatOut *hp = reinterpret_cast<StatOut *>(c.h_pinned);
  VT_CUDA(cudaMemcpyAsync(hp, dout, sizeof(StatOut), cudaMemcpyDeviceToHost, c.stream));
  VT_CUDA(cudaStreamSynchronize(c.stream));
  *host_out = *hp;
  return 0;
}

int k_minmax(const float *d, long n, float *mn, float *mx) {
  StatOut o;
  int rc = run_stats<0>(d, n, 0.f, &o);
  if (rc) return rc;
  *mn = o.mn; *mx = o.mx;
  return 0;
}
int k_sum(const float *d, long n, double *sum) {
  StatOut o;
  int rc = run_stats<1>(d, n, 0.f, &o);
  if (rc) return rc;
  *sum = o.sum;
  return 0;
}
int k_minmaxsum(const float *d, long n, float *mn, float *mx, double *sum) {
  StatOut o;
  int rc = run_stats<2>(d, n, 0.f, &o);
  if (rc) return rc;
  *mn = o.mn; *mx = o.mx; *sum = o.sum;
  return 0;
}
int k_sumsq_delta(const float *d, long n, float mean, double *sum) {
  StatOut o;
  int rc = run_stats<3>(d, n, mean, &o);
  if (rc) return rc;
  *sum = o.sum;
  return 0;
}

// ------------------------------------------------------------------------------------------ unary
struct UnaryParams {
  float p0, p1, p2;
  double pd;
};

template <int OP>
__device__ __forceinline__ float unary_apply(float x, const UnaryParams &p) {
  switch (OP) {
    case U_CLAMP:  // VolumetricData.C:637-638 (else-if: NaN passes through)
      return x < p.p0 ? p.p0 : (x > p.p1 ? p.p1 : x);
    case U_SCALE: return nan_like_x86(x * p.p0, x, p.p0);                       // :659
    case U_ADD: return nan_like_x86(x + p.p0, x, p.p0);                         // :683
    case U_RESCALE: return nan_like_x86(p.p0 + p.p1 * (x - p.p2), p.p0, p.p1, x, p.p2);  // :706  min + ratio*(x - oldmin)
    case U_SIGMA: return nan_like_x86((x - p.p0) * p.p1, x, p.p0, p.p1);              // :944-945  (x -= mean; x *= 1/sigma)
    case U_BINMASK: return x > p.p0 ? 1.0f : 0.0f;       // :960
    case U_MDFFPOT: {                                    // :1010-1016
      float v = x;
      if ((double)v < p.pd) v = (float)p.pd;
      v = v * -1.0f;
      return nan_like_x86(p.p0 * (v - p.p1), p.p0, v, p.p1);  // ratio * (v - mininvert)
    }
  }
  return x;
}

template <int OP>
__global__ void __launch_bounds__(kThreads) unary_kernel(float *__restrict__ d, long n, UnaryParams p) {
  const long nv = n >> 2;
  float4 *v = reinterpret_cast<float4 *>(d);
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
    float4 q = ld_stream(v + i);
    q.x = unary_apply<OP>(q.x, p); q.y = unary_apply<OP>(q.y, p);
    q.z = unary_apply<OP>(q.z, p); q.w = unary_apply<OP>(q.w, p);
    __stcs(v + i, q);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long i = (nv << 2) + threadIdx.x;
    d[i] = unary_apply<OP>(d[i], p);
  }
}

int k_unary(UnaryOp op, float *d, long n, float p0, float p1, float p2, double pd) {
  VT_REQUIRE_CTX();
  if (n <= 0) return 0;
  const int grid = stream_grid(n >> 2);
  UnaryParams p{p0, p1, p2, pd};
  cudaStream_t s = ctx().stream;
  prof_begin(PK_UNARY);
  switch (op) {
    case U_CLAMP: unary_kernel<U_CLAMP><<<grid, kThreads, 0, s>>>(d, n, p); break;
    case U_SCALE: unary_kernel<U_SCALE><<<grid, kThreads, 0, s>>>(d, n, p); break;
    case U_ADD: unary_kernel<U_ADD><<<grid, kThreads, 0, s>>>(d, n, p); break;
    case U_RESCALE: unary_kernel<U_RESCALE><<<grid, kThreads, 0, s>>>(d, n, p); break;
    case U_SIGMA: unary_kernel<U_SIGMA><<<grid, kThreads, 0, s>>>(d, n, p); break;
    case U_BINMASK: unary_kernel<U_BINMASK><<<grid, kThreads, 0, s>>>(d, n, p); break;
    case U_MDFFPOT: unary_kernel<U_MDFFPOT><<<grid, kThreads, 0, s>>>(d, n, p); break;
  }
  prof_end(PK_UNARY);
  VT_LAUNCHED();
  return 0;
}

__global__ void __launch_bounds__(kThreads) fill_kernel(float *__restrict__ d, long n, float v) {
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) d[i] = v;
}
int k_fill(float *d, long n, float v) {
  VT_REQUIRE_CTX();
  if (n <= 0) return 0;
  fill_kernel<<<stream_grid(n), kThreads, 0, ctx().stream>>>(d, n, v);
  VT_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------ histogram
// Voltool.C:443-461: bin = long((x - min) * binwidthinv) with a double inverse width.  The
// reference writes bins[nbins] for the maximum voxel (one past the end); that count is dropped.
__global__ void __launch_bounds__(kThreads) hist_kernel(const float *__restrict__ d, long n, float mn,
                                                        double binwidthinv, int nbins,
                                                        unsigned long long *__restrict__ bins) {
  extern __shared__ unsigned int sh[];
  for (int i = threadIdx.x; i < nbins; i += blockDim.x) sh[i] = 0u;
  __syncthreads();
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double t = (double)(d[i] - mn) * binwidthinv;
    // (long) cast semantics of x86: NaN / out of range -> LONG_MIN -> rejected below
    if (t >= 0.0 && t < (double)nbins) {
      const int k = (int)t;
      if (k >= 0 && k < nbins) atomicAdd(&sh[k], 1u);
    } else if (t > -1.0 && t < 0.0) {
      atomicAdd(&sh[0], 1u);  // truncation toward zero maps (-1,0) to bin 0
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nbins; i += blockDim.x)
    if (sh[i]) atomicAdd(&bins[i], (unsigned long long)sh[i]);
}

int k_histogram(const float *d, long n, float mn, double binwidthinv, int nbins, unsigned long long *d_bins) {
  VT_REQUIRE_CTX();
  VT_CUDA(cudaMemsetAsync(d_bins, 0, sizeof(unsigned long long) * nbins, ctx().stream));
  if (n <= 0) return 0;
  int grid = stream_grid(n);
  if (grid > ctx().sm_count * 4) grid = ctx().sm_count * 4;
  hist_kernel<<<grid, kThreads, sizeof(unsigned int) * nbins, ctx().stream>>>(d, n, mn, binwidthinv, nbins, d_bins);
  VT_LAUNCHED();
  return 0;
}

}  // namespace vt
