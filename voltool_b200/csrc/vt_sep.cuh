// vt_sep.cuh -- table-driven trilinear sampling for the separable (axis-aligned) path, shared by the
// two-map CC kernel and the batched fit CC kernel.
#pragma once
#include "vt_device.cuh"
#include "vt_geom.h"

namespace vt {
using geom::AxisEntry;

// Bilinear value of one z-plane of a source map at (x entry, y entry): the x then y lerps of
// voxel_value_interpolate (VolumetricData.C:273-287) restricted to one plane.
__device__ __forceinline__ float plane_value(const float *__restrict__ plane, long row0, long row1, int x0, int x1,
                                             float xf, float yf) {
  const float a0 = __ldg(plane + row0 + x0), a1 = __ldg(plane + row0 + x1);
  const float b0 = __ldg(plane + row1 + x0), b1 = __ldg(plane + row1 + x1);
  const float xa = a0 + xf * (a1 - a0);
  const float xb = b0 + xf * (b1 - b0);
  return xa + yf * (xb - xa);
}

// Tracks the two bilinear plane values (z0, z1) of one source map for a thread marching in z.
struct PlaneCache {
  int z0 = -1, z1 = -1;
  float v0 = 0.f, v1 = 0.f;
};

__device__ __forceinline__ float sep_sample(const float *__restrict__ data, long plane_stride, long row0, long row1,
                                            int x0, int x1, float xf, float yf, const AxisEntry &ez, PlaneCache &pc) {
  float n0, n1;
  if (ez.i0 == pc.z0) n0 = pc.v0;
  else if (ez.i0 == pc.z1) n0 = pc.v1;
  else n0 = plane_value(data + (long)ez.i0 * plane_stride, row0, row1, x0, x1, xf, yf);
  if (ez.i1 == ez.i0) n1 = n0;
  else if (ez.i1 == pc.z1) n1 = pc.v1;
  else if (ez.i1 == pc.z0) n1 = pc.v0;
  else n1 = plane_value(data + (long)ez.i1 * plane_stride, row0, row1, x0, x1, xf, yf);
  pc.z0 = ez.i0; pc.v0 = n0; pc.z1 = ez.i1; pc.v1 = n1;
  return n0 + ez.f * (n1 - n0);
}


}  // namespace vt
