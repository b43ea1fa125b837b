// vt_blur.cu -- separable 3-D Gaussian blur (VolumetricData::gaussian_blur, VolumetricData.C:977-997).
//
// The reference delegates the arithmetic to GaussianBlur<float> (Segmentation.h, not in the
// snapshot); the definition held to here is the one in oracle/ (x -> y -> z passes, taps
// -R..R in order, acc = fmaf(w, v, acc), clamp-to-edge, f32 between passes).
//
// Pass structure: the x pass stages each row segment (+halo) in shared memory; the y and z
// passes march a register window along the filtered axis so each input element is read from
// L2/HBM once per pass.
#include "vt_device.cuh"

namespace vt {

constexpr int kMaxTaps = 2 * 96 + 1;
__constant__ float c_taps[kMaxTaps];

// ---- x pass: one row per block-iteration, row staged through shared memory -------------------
__global__ void __launch_bounds__(256) blur_x_kernel(const float *__restrict__ in, float *__restrict__ out, int nx,
                                                     long rows, int R) {
  extern __shared__ float srow[];  // nx + 2R
  for (long row = blockIdx.x; row < rows; row += gridDim.x) {
    const float *src = in + row * nx;
    for (int i = threadIdx.x; i < nx + 2 * R; i += blockDim.x) {
      int j = i - R;
      j = j < 0 ? 0 : (j > nx - 1 ? nx - 1 : j);
      srow[i] = __ldcs(src + j);
    }
    __syncthreads();
    for (int x = threadIdx.x; x < nx; x += blockDim.x) {
      float acc = 0.0f;
      for (int k = 0; k <= 2 * R; ++k) acc = __fmaf_rn(c_taps[k], srow[x + k], acc);
      out[row * nx + x] = acc;
    }
    __syncthreads();
  }
}

// ---- y / z pass: lines along `stride`, threads across the contiguous x (and outer) dimension ----
// Each thread owns one line position and slides along the axis in segments of SEG outputs,
// reading SEG + 2R inputs per segment (clamped at the ends).
template <int SEG>
__global__ void __launch_bounds__(256) blur_axis_kernel(const float *__restrict__ in, float *__restrict__ out,
                                                        int n_axis, long stride, long inner, long outer, int R) {
  // a "line" is identified by (o, i): base = o * (n_axis * stride) + i, i in [0, inner)
  const long lines = inner * outer;
  const int nseg = (n_axis + SEG - 1) / SEG;
  const long work = lines * nseg;
  const long step = (long)gridDim.x * blockDim.x;
  for (long w = (long)blockIdx.x * blockDim.x + threadIdx.x; w < work; w += step) {
    // consecutive threads -> consecutive i (coalesced); segments are the slowest index
    const long line = w % lines;
    const int seg = (int)(w / lines);
    const long o = line / inner, i = line - o * inner;
    const float *src = in + o * (n_axis * stride) + i;
    float *dst = out + o * (n_axis * stride) + i;
    const int a0 = seg * SEG;
    float acc[SEG];
#pragma unroll
    for (int s = 0; s < SEG; ++s) acc[s] = 0.0f;
    // input j = a0 - R + t contributes to output a0+s with tap index k = t - s
    for (int t = 0; t < SEG + 2 * R; ++t) {
      int j = a0 - R + t;
      j = j < 0 ? 0 : (j > n_axis - 1 ? n_axis - 1 : j);
      const float v = __ldg(src + (long)j * stride);
#pragma unroll
      for (int s = 0; s < SEG; ++s) {
        const int k = t - s;
        if (k >= 0 && k <= 2 * R) acc[s] = __fmaf_rn(c_taps[k], v, acc[s]);
      }
    }
#pragma unroll
    for (int s = 0; s < SEG; ++s)
      if (a0 + s < n_axis) dst[(long)(a0 + s) * stride] = acc[s];
  }
}

int k_blur(const float *src, float *dst, float *tmp, int nx, int ny, int nz, const float *w, int R) {
  VT_REQUIRE_CTX();
  Ctx &c = ctx();
  if (2 * R + 1 > kMaxTaps) { set_error("vt_gaussian_blur: sigma too large (radius %d > 96)", R); return 1; }
  VT_CUDA(cudaMemcpyToSymbolAsync(c_taps, w, sizeof(float) * (2 * R + 1), 0, cudaMemcpyHostToDevice, c.stream));
  const long rows = (long)ny * nz;
  int grid = (int)(rows < (long)c.sm_count * 8 ? rows : (long)c.sm_count * 8);
  if (grid < 1) grid = 1;
  // x: src -> dst ; y: dst -> tmp ; z: tmp -> dst
  prof_begin(PK_BLUR);
  blur_x_kernel<<<grid, 256, sizeof(float) * (nx + 2 * R), c.stream>>>(src, dst, nx, rows, R);
  VT_LAUNCHED();
  const int g2 = c.sm_count * 8;
  blur_axis_kernel<8><<<g2, 256, 0, c.stream>>>(dst, tmp, ny, (long)nx, (long)nx, (long)nz, R);
  VT_LAUNCHED();
  blur_axis_kernel<8><<<g2, 256, 0, c.stream>>>(tmp, dst, nz, (long)nx * ny, (long)nx * ny, 1L, R);
  prof_end(PK_BLUR);
  VT_LAUNCHED();
  return 0;
}

}  // namespace vt
