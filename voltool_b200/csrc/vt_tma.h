// vt_tma.h -- tensor-map TMA (cp.async.bulk.tensor) plumbing shared by the kernels that stage boxes of a map:
// the driver's cuTensorMapEncodeTiled is reached through the runtime's entry-point query (no -lcuda at link time).
#pragma once
#include <cuda.h>

namespace vt {

// 3-D fp32 tensor map over a dense map (x fastest): dims = {nx, ny, nz}, box = {bx, by, bz} elements.
// swizzle128: CU_TENSOR_MAP_SWIZZLE_128B (the box's inner extent must then be exactly 32 floats), else none.
// Returns 0 on success, 1 if the driver entry point is unavailable or the encode fails (callers fall back).
int tma_encode_map3d(CUtensorMap *map, const float *base, int nx, int ny, int nz, int bx, int by, int bz, bool swizzle128);

__device__ __forceinline__ void tma_load_box3d(unsigned int dst_smem, const CUtensorMap *map, int cx, int cy, int cz,
                                               unsigned int bar_smem) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst_smem),
      "l"(map), "r"(cx), "r"(cy), "r"(cz), "r"(bar_smem)
      : "memory");
}

}  // namespace vt
