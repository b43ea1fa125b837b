#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float x, float y){ f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y)); return r;}
__device__ __forceinline__ void unpack2(f32x2 r, float &x, float &y) { asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(r)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c){ f32x2 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;}
__constant__ float c_w[32];
// MODE 0: FFMA2 acc = bcast(v) * (UR pair) + acc      (scalar reg x uniform pair)
// MODE 1: FFMA2 acc = (UR scalar) * vpair + acc        (uniform scalar x reg pair)
// MODE 2: FFMA2 acc = vpair * wpair(reg) + acc         (all registers)
// MODE 3: scalar FFMA x2
template <int MODE, int NACC>
__global__ void k(float *out, int iters, float seed) {
  f32x2 acc[NACC];
  float v = seed + threadIdx.x, u = seed * 0.5f;
  for (int i = 0; i < NACC; ++i) acc[i] = pack2(0.f, 0.f);
  f32x2 wreg = pack2(seed * 1.0001f, seed * 0.9999f);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < NACC; ++i) {
        if (MODE == 0) acc[i] = fma2(pack2(v, v), pack2(c_w[2 * (i & 7)], c_w[2 * (i & 7) + 1]), acc[i]);
        else if (MODE == 1) acc[i] = fma2(pack2(c_w[i & 15], c_w[i & 15]), pack2(v, u), acc[i]);
        else if (MODE == 2) acc[i] = fma2(pack2(v, u), wreg, acc[i]);
        else { float a, b; unpack2(acc[i], a, b); a = __fmaf_rn(v, c_w[i & 15], a); b = __fmaf_rn(u, c_w[i & 15], b); acc[i] = pack2(a, b); }
      }
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < NACC; ++i) { float a, b; unpack2(acc[i], a, b); s += a + b; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[gridDim.x * blockDim.x] = (float)(t1 - t0);
}
template <int MODE, int NACC>
void run(const char *name, int warps) {
  float *d; cudaMalloc(&d, sizeof(float) * (148 * 1024 + 1));
  const int iters = 2000;
  k<MODE, NACC><<<148, warps * 32>>>(d, iters, 1.0f);
  cudaDeviceSynchronize();
  k<MODE, NACC><<<148, warps * 32>>>(d, iters, 1.0f);
  cudaDeviceSynchronize();
  float cyc; cudaMemcpy(&cyc, d + 148 * warps * 32, 4, cudaMemcpyDeviceToHost);
  double inst = (double)iters * 8 * NACC * (MODE == 3 ? 2 : 1);
  printf("%-40s warps/SM=%2d acc=%2d: %.0f cycles, %.3f cycles per warp-instr per warp, SM rate %.2f warp-instr/clk (%.1f FMA/clk/SM)\n", name, warps, NACC, cyc,
         cyc / inst, inst * warps / cyc, inst * warps / cyc * (MODE == 3 ? 32 : 64));
  cudaFree(d);
}
int main() {
  float w[32]; for (int i = 0; i < 32; ++i) w[i] = 1.0f / (i + 3);
  cudaMemcpyToSymbol(c_w, w, sizeof(w));
  run<0, 8>("FFMA2 R.F32 x UR.F32x2", 4);  run<0, 8>("FFMA2 R.F32 x UR.F32x2", 8);  run<0, 8>("FFMA2 R.F32 x UR.F32x2", 16);
  run<1, 8>("FFMA2 UR.F32 x R.F32x2", 4);  run<1, 8>("FFMA2 UR.F32 x R.F32x2", 8);  run<1, 8>("FFMA2 UR.F32 x R.F32x2", 16);
  run<2, 8>("FFMA2 R.F32x2 x R.F32x2", 4); run<2, 8>("FFMA2 R.F32x2 x R.F32x2", 8); run<2, 8>("FFMA2 R.F32x2 x R.F32x2", 16);
  run<3, 8>("FFMA x2 (scalar)", 4);        run<3, 8>("FFMA x2 (scalar)", 8);        run<3, 8>("FFMA x2 (scalar)", 16);
  run<0, 2>("FFMA2 R.F32 x UR.F32x2 (2 chains)", 8); run<0, 2>("FFMA2 R.F32 x UR.F32x2 (2 chains)", 16);
  run<1, 2>("FFMA2 UR.F32 x R.F32x2 (2 chains)", 8);
  return 0;
}
