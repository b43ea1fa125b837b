// TMA box-load throughput for blur-like access: each CTA streams the planes of its (BXW x ROWS) tile column.
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma3(unsigned dst, const CUtensorMap *m, int x, int y, int z, unsigned bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst), "l"(m), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
}
template <int S>
__global__ void k(const __grid_constant__ CUtensorMap tm, float *out, int xt, int tiles, int nz, int bxw, int rows, int zper, int tx_step, int ty_step, int halo_x, int halo_y) {
  extern __shared__ __align__(1024) unsigned char sm[];
  const unsigned base = (smem_u32(sm) + 1023u) & ~1023u;
  const unsigned stage_bytes = ((unsigned)(bxw * rows * zper * 4) + 1023u) & ~1023u;
  const unsigned bars = base + S * stage_bytes;
  if (threadIdx.x == 0) { for (int s = 0; s < S; ++s) mbar_init(bars + 8 * s, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
  __syncthreads();
  const int t = blockIdx.x % tiles, tx = t % xt, ty = t / xt;
  const int x0 = tx * tx_step - halo_x, y0 = ty * ty_step - halo_y;
  const int nops = nz / zper;
  float acc = 0.f;
  if (threadIdx.x == 0) for (int q = 0; q < S && q < nops; ++q) { mbar_expect_tx(bars + 8 * q, bxw * rows * zper * 4); tma3(base + q * stage_bytes, &tm, x0, y0, q * zper, bars + 8 * q); }
  for (int q = 0; q < nops; ++q) {
    const unsigned st = q % S;
    mbar_wait(bars + 8 * st, (q / S) & 1);
    float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(base + st * stage_bytes + (threadIdx.x & 255) * 4));
    acc += v;
    __syncthreads();
    if (threadIdx.x == 0 && q + S < nops) { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); mbar_expect_tx(bars + 8 * st, bxw * rows * zper * 4); tma3(base + st * stage_bytes, &tm, x0, y0, (q + S) * zper, bars + 8 * st); }
  }
  if (acc == 123.456f) out[0] = acc;
}
typedef CUresult (*Enc)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
template <int S>
void run(Enc enc, float *src, float *out, int n, int bxw, int rows, int zper, int tx_step, int ty_step, int hx, int hy, CUtensorMapSwizzle sw, const char *swn, int ctas_per_sm, CUtensorMapL2promotion prom) {
  CUtensorMap m;
  cuuint64_t dims[3] = {(cuuint64_t)n, (cuuint64_t)n, (cuuint64_t)n}, str[2] = {(cuuint64_t)n * 4, (cuuint64_t)n * n * 4};
  cuuint32_t box[3] = {(cuuint32_t)bxw, (cuuint32_t)rows, (cuuint32_t)zper}, es[3] = {1, 1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, src, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, prom, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d (box %d x %d x %d %s)\n", (int)r, bxw, rows, zper, swn); return; }
  const int xt = (n + tx_step - 1) / tx_step, yt = (n + ty_step - 1) / ty_step, tiles = xt * yt;
  const size_t smem = (size_t)S * (((size_t)bxw * rows * zper * 4 + 1023) & ~1023) + 8 * S + 1024;
  cudaFuncSetAttribute(k<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int grid = tiles;  // one CTA per tile column; the hardware fills SMs in order
  (void)ctas_per_sm;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<S><<<grid, 256, smem>>>(m, out, xt, tiles, n, bxw, rows, zper, tx_step, ty_step, hx, hy);
  cudaEventRecord(e0);
  k<S><<<grid, 256, smem>>>(m, out, xt, tiles, n, bxw, rows, zper, tx_step, ty_step, hx, hy);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t e = cudaGetLastError();
  const double bytes = (double)tiles * bxw * rows * 4.0 * n;
  printf("box %3d x %3d x %d  step %2d x %3d  S=%d %-5s smem/CTA %6.1f KB  CTAs %5d: %.3f ms  staged %.2f GB  %.0f GB/s staged  (%s)\n", bxw, rows, zper, tx_step, ty_step, S, swn, smem / 1024.0, grid, ms, bytes / 1e9, bytes / ms / 1e6, cudaGetErrorString(e));
}
int main() {
  const int n = 512;
  float *src, *out; cudaMalloc(&src, sizeof(float) * (size_t)n * n * n); cudaMalloc(&out, 1024);
  cudaMemset(src, 0, sizeof(float) * (size_t)n * n * n);
  void *p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  Enc enc = (Enc)p;
  auto P128 = CU_TENSOR_MAP_L2_PROMOTION_L2_128B, P256 = CU_TENSOR_MAP_L2_PROMOTION_L2_256B, PN = CU_TENSOR_MAP_L2_PROMOTION_NONE;
  run<6>(enc, src, out, n, 32, 76, 1, 16, 64, 8, 6, CU_TENSOR_MAP_SWIZZLE_128B, "sw128", 2, P128);   // the blur kernel's shape
  run<6>(enc, src, out, n, 32, 76, 1, 16, 64, 8, 6, CU_TENSOR_MAP_SWIZZLE_NONE, "none", 2, P128);
  run<6>(enc, src, out, n, 32, 76, 1, 16, 64, 8, 6, CU_TENSOR_MAP_SWIZZLE_128B, "sw128", 2, PN);
  run<6>(enc, src, out, n, 32, 76, 1, 16, 64, 8, 6, CU_TENSOR_MAP_SWIZZLE_128B, "sw128", 2, P256);
  run<3>(enc, src, out, n, 32, 76, 2, 16, 64, 8, 6, CU_TENSOR_MAP_SWIZZLE_128B, "sw128", 2, P128);   // two planes per op
  run<2>(enc, src, out, n, 32, 76, 4, 16, 64, 8, 6, CU_TENSOR_MAP_SWIZZLE_128B, "sw128", 2, P128);
  run<12>(enc, src, out, n, 32, 76, 1, 16, 64, 8, 6, CU_TENSOR_MAP_SWIZZLE_128B, "sw128", 1, P128);
  run<6>(enc, src, out, n, 32, 64, 1, 32, 64, 0, 0, CU_TENSOR_MAP_SWIZZLE_128B, "sw128", 2, P128);   // no halo: disjoint boxes
  run<6>(enc, src, out, n, 64, 76, 1, 48, 64, 8, 6, CU_TENSOR_MAP_SWIZZLE_NONE, "none", 2, P128);    // 256-byte rows
  run<6>(enc, src, out, n, 128, 44, 1, 112, 32, 8, 6, CU_TENSOR_MAP_SWIZZLE_NONE, "none", 2, P128);  // 512-byte rows
  run<4>(enc, src, out, n, 256, 44, 1, 240, 32, 8, 6, CU_TENSOR_MAP_SWIZZLE_NONE, "none", 2, P128);  // 1 KB rows
  run<6>(enc, src, out, n, 32, 38, 1, 16, 32, 8, 3, CU_TENSOR_MAP_SWIZZLE_128B, "sw128", 2, P128);   // half the rows
  return 0;
}
