// vt_sim.cu -- the batched pose pipeline: rigid transform, bounding box, per-pose grid geometry,
// Gaussian splat, and the per-pose cross correlation against the resident target map.
//
// Replaces, per candidate pose of fit's rotate() loop (TclVoltool.C:160-199):
//   moveby / do_rotate x3 / moveby              TclVoltool.C:52-63, 131-136, 165-176
//   QuickSurf::calc_density_map (CPU branch)    TclVoltool.C:119-122   (restated in oracle/)
//   cc_threaded(sim, target, &cc, 0.1)          TclVoltool.C:124 -> MDFF.C:102-151
// and, with a single identity pose, voltool sim / cc (TclMDFF.C:154-157, 496-524).
//
// Splat design (no atomics, no tensor cores: nothing here is a contraction):
//   * atoms are sorted once along a Morton curve and cut into groups of 32 with a bounding sphere;
//     groups are rigid, so a pose only moves their centres.
//   * a CTA owns a 16x16x8-voxel tile.  It culls groups against the tile (thread per group), then
//     tests the atoms of surviving groups exactly (warp per group, the reference's integer box),
//     compacting ids in a fixed order through ballots + shared-memory offsets.
//   * atoms are staged 64 at a time as separable records: Ex[16] = exp2(dx^2 * a) per tile column
//     (0 outside the atom's x range), dy^2[16], dz^2[8] (+inf outside the y/z range), so the
//     reference's cutoff test dy^2+dz^2 < (gausslim*sigma)^2 is evaluated on the same floats.
//   * each thread owns 8 consecutive x voxels of one (y,z) in registers; a warp covers 8x8x4
//     voxels and skips staged atoms whose box misses that region (8-bit mask per record).
//     Per (thread, atom): one ex2 for the y-z factor and 8 FMAs.
//   * every voxel is written exactly once (no memset, no read-modify-write).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "vt_pose.cuh"
#include "vt_sep.cuh"

namespace vt {

constexpr float kLog2e = 1.44269504088896f;  // MLOG2EF of the oracle's splat

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// monotone float <-> uint key for atomicMin/atomicMax
__device__ __forceinline__ unsigned int fkey(float f) {
  const unsigned int b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(unsigned int k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// ------------------------------------------------------------------------------------------ atom set
// The atom set is prepared on the device: Morton keys, a stable radix sort of (key, index) pairs (cub), the packed
// per-atom records and constants, the group bounding spheres.  The host only scans the coordinates once (bounding
// box for the key grid, largest radius, the capacity-planning sphere) and uploads them as they are.  (Round 1 sorted
// and packed on the host: 0.24 of the 0.48 ms of a 10 k-atom `voltool cc` call.)
__host__ __device__ static inline unsigned int spread10(unsigned int v) {
  v &= 0x3ff;
  v = (v | (v << 16)) & 0x030000FF;
  v = (v | (v << 8)) & 0x0300F00F;
  v = (v | (v << 4)) & 0x030C30C3;
  v = (v | (v << 2)) & 0x09249249;
  return v;
}

// Morton order over 2^10 cells per axis; the sort is stable, so ties keep the caller's atom order
__global__ void __launch_bounds__(256) morton_keys_kernel(const float *__restrict__ pos, int n, float lox, float loy, float loz, float q,
                                                          unsigned int *__restrict__ key, int *__restrict__ idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float lo[3] = {lox, loy, loz};
  unsigned int code = 0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    int c = (int)((pos[3 * i + k] - lo[k]) * q);
    c = min(1023, max(0, c));
    code |= spread10((unsigned int)c) << k;
  }
  key[i] = code;
  idx[i] = i;
}

// sorted records: xyzr, and (first time) the per-atom constants of the oracle's splat_range
__global__ void __launch_bounds__(256) atom_pack_kernel(const float *__restrict__ pos, const float *__restrict__ radii,
                                                        const int *__restrict__ perm, int n, SimParams prm, int first,
                                                        float4 *__restrict__ xyzr, float4 *__restrict__ cst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int o = perm[i];
  const float r = radii[o];
  xyzr[i] = make_float4(pos[3 * o], pos[3 * o + 1], pos[3 * o + 2], r);
  if (first) {
    const float scaledrad = r * prm.radscale;
    const float arinv = -(1.0f / (2.0f * scaledrad * scaledrad)) * kLog2e;
    float radlim = prm.gausslim * scaledrad;
    const float radlim2 = radlim * radlim;
    radlim *= prm.invgs;
    cst[i] = make_float4(arinv, radlim2, radlim, r);
  }
}

// a warp per group of 32 atoms: bounding sphere about the box centre + slack for the float rounding of the rigid
// transform; the largest cutoff of the group is kept separately (an atom's footprint is a box in x and a disc in
// y-z, both inside the cube of half-width cutoff, so "sphere(radius) meets tile box grown by cutoff" is a tighter
// conservative test than one sphere of radius + sqrt(2) * cutoff)
__global__ void __launch_bounds__(128) group_sphere_kernel(const float4 *__restrict__ xyzr, int n, int ngroups, float gausslim,
                                                           float radscale, float4 *__restrict__ groups, float *__restrict__ cuts) {
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (g >= ngroups) return;
  const int i = g * 32 + lane;
  const bool has = i < n;
  const float4 a = has ? xyzr[i] : xyzr[g * 32];
  float lo[3] = {a.x, a.y, a.z}, hi[3] = {a.x, a.y, a.z};
  float maxcut = gausslim * a.w * radscale;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
      hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
    }
    maxcut = fmaxf(maxcut, __shfl_xor_sync(0xffffffffu, maxcut, o));
  }
  const float c[3] = {0.5f * (lo[0] + hi[0]), 0.5f * (lo[1] + hi[1]), 0.5f * (lo[2] + hi[2])};
  const float dx = a.x - c[0], dy = a.y - c[1], dz = a.z - c[2];
  float r2 = dx * dx + dy * dy + dz * dz;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) r2 = fmaxf(r2, __shfl_xor_sync(0xffffffffu, r2, o));
  if (lane == 0) {
    const float reach = sqrtf(r2) + 0.05f + 1e-4f * (fabsf(c[0]) + fabsf(c[1]) + fabsf(c[2]));
    cuts[g] = maxcut;
    groups[g] = make_float4(c[0], c[1], c[2], reach);
  }
}

}  // namespace vt
#include <cub/device/device_radix_sort.cuh>
namespace vt {

// coordinates (caller's order) -> device, then the sorted records and the group spheres
static int upload_positions(AtomSet &as, const float *pos, bool first) {
  cudaStream_t s = ctx().stream;
  // the sphere all atoms lie in (capacity planning only)
  double c[3] = {0, 0, 0};
  for (int i = 0; i < as.n; ++i) { c[0] += pos[3 * i]; c[1] += pos[3 * i + 1]; c[2] += pos[3 * i + 2]; }
  for (int k = 0; k < 3; ++k) as.bound_center[k] = (float)(c[k] / std::max(1, as.n));
  double r2 = 0;
  for (int i = 0; i < as.n; ++i) {
    const double dx = pos[3 * i] - as.bound_center[0], dy = pos[3 * i + 1] - as.bound_center[1], dz = pos[3 * i + 2] - as.bound_center[2];
    r2 = std::max(r2, dx * dx + dy * dy + dz * dz);
  }
  as.bound_radius = (float)std::sqrt(r2);
  VT_CUDA(cudaMemcpyAsync(as.d_rawpos, pos, sizeof(float) * 3 * (size_t)as.n, cudaMemcpyHostToDevice, s));   // (pageable: staged)
  atom_pack_kernel<<<(as.n + 255) / 256, 256, 0, s>>>(as.d_rawpos, as.d_rawrad, as.d_perm, as.n, as.prm, first ? 1 : 0, as.d_xyzr,
                                                      as.d_const);
  VT_LAUNCHED();
  group_sphere_kernel<<<(as.ngroups * 32 + 127) / 128, 128, 0, s>>>(as.d_xyzr, as.n, as.ngroups, as.prm.gausslim, as.prm.radscale,
                                                                    as.d_group, as.d_gcut);
  VT_LAUNCHED();
  return 0;
}

int atomset_create(const float *pos, const float *radii, long n, const SimParams &prm, AtomSet &as) {
  VT_REQUIRE_CTX();
  if (n <= 0) { set_error("atom set is empty"); return 1; }
  cudaStream_t s = ctx().stream;
  as.n = (int)n;
  as.ngroups = (as.n + 31) / 32;
  as.prm = prm;
  float lo[3] = {pos[0], pos[1], pos[2]}, hi[3] = {pos[0], pos[1], pos[2]};
  float maxrad = radii[0];
  for (long i = 0; i < n; ++i) {
    for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], pos[3 * i + k]); hi[k] = std::max(hi[k], pos[3 * i + k]); }
    maxrad = std::max(maxrad, radii[i]);
  }
  as.maxrad = maxrad;
  as.pad = sim_padding(prm.radscale, maxrad);
  const float ext = std::max(std::max(hi[0] - lo[0], hi[1] - lo[1]), std::max(hi[2] - lo[2], 1e-3f));
  const float q = 1023.0f / ext;
  if (dev_alloc((void **)&as.d_xyzr, sizeof(float4) * n)) return 2;
  if (dev_alloc((void **)&as.d_const, sizeof(float4) * n)) return 2;
  if (dev_alloc((void **)&as.d_group, sizeof(float4) * as.ngroups)) return 2;
  if (dev_alloc((void **)&as.d_gcut, sizeof(float) * as.ngroups)) return 2;
  if (dev_alloc((void **)&as.d_rawpos, sizeof(float) * 3 * n)) return 2;
  if (dev_alloc((void **)&as.d_rawrad, sizeof(float) * n)) return 2;
  if (dev_alloc((void **)&as.d_perm, sizeof(int) * n)) return 2;
  // scratch of the sort: keys in / out, indices in, cub's temporary storage
  unsigned int *d_key = nullptr, *d_key2 = nullptr;
  int *d_idx = nullptr;
  void *d_tmp = nullptr;
  size_t tmp_bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_key, d_key2, d_idx, as.d_perm, as.n, 0, 30, s);
  int rc = dev_alloc((void **)&d_key, sizeof(unsigned int) * n);
  if (!rc) rc = dev_alloc((void **)&d_key2, sizeof(unsigned int) * n);
  if (!rc) rc = dev_alloc((void **)&d_idx, sizeof(int) * n);
  if (!rc) rc = dev_alloc(&d_tmp, tmp_bytes ? tmp_bytes : 16);
  cudaError_t e = cudaSuccess;
  if (!rc) {
    e = cudaMemcpyAsync(as.d_rawpos, pos, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(as.d_rawrad, radii, sizeof(float) * n, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) {
      morton_keys_kernel<<<(as.n + 255) / 256, 256, 0, s>>>(as.d_rawpos, as.n, lo[0], lo[1], lo[2], q, d_key, d_idx);
      ctx().launches++;
      // (key, index) by key, ties by index: a stable LSD radix sort over the 30 key bits
      e = cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_key, d_key2, d_idx, as.d_perm, as.n, 0, 30, s);
      ctx().launches += 3;
    }
  }
  dev_free(d_key); dev_free(d_key2); dev_free(d_idx); dev_free(d_tmp);
  if (rc) return rc;
  if (e != cudaSuccess) return cuda_fail(e, "atom set sort", __FILE__, __LINE__);
  return upload_positions(as, pos, true);
}

int atomset_set_positions(AtomSet &as, const float *pos) { return upload_positions(as, pos, false); }

void atomset_destroy(AtomSet &as) {
  dev_free(as.d_xyzr); dev_free(as.d_const); dev_free(as.d_group); dev_free(as.d_gcut);
  dev_free(as.d_rawpos); dev_free(as.d_rawrad); dev_free(as.d_perm);
  as = AtomSet();
}

// ------------------------------------------------------------------------------------------ batch workspace
int batch_create(const AtomSet &as, const vt_grid *target, int B, PoseBatch &pb) {
  VT_REQUIRE_CTX();
  pb.B = B;
  const float gs = as.prm.gs;
  // any rigid pose keeps the atoms inside the bounding sphere: extent <= 2R + 2 pad (+ shift slack)
  const double ext = 2.0 * as.bound_radius + 2.0 * as.pad + 4.0 * gs;
  pb.cap_n = (int)std::ceil(ext / gs) + 2;
  pb.cap_grid = (long)pb.cap_n * pb.cap_n * pb.cap_n;
  if (target) {
    double dens = 1.15 / gs;
    const double la[3] = {target->xaxis[0], target->yaxis[1], target->zaxis[2]};
    const int na[3] = {target->xsize, target->ysize, target->zsize};
    double maxlen = 0;
    for (int k = 0; k < 3; ++k) {
      if (la[k] > 0) dens = std::max(dens, na[k] / la[k]);
      maxlen = std::max(maxlen, std::min(la[k], (double)pb.cap_n * gs));
    }
    pb.cap_o = (int)(maxlen * dens * 1.01) + 8;
    pb.cap_items = (long)((pb.cap_o + 31) / 32) * ((pb.cap_o + kRowsFit - 1) / kRowsFit) * ((pb.cap_o + kZChunkFit - 1) / kZChunkFit);
  } else {
    pb.cap_o = 0;
    pb.cap_items = 0;
  }
  if (dev_alloc((void **)&pb.d_tpos, sizeof(float4) * (size_t)B * as.n)) return 2;
  if (dev_alloc((void **)&pb.d_tgroup, sizeof(float4) * (size_t)B * as.ngroups)) return 2;
  if (dev_alloc((void **)&pb.d_bbox, sizeof(unsigned int) * 6 * B)) return 2;
  if (dev_alloc((void **)&pb.d_geom, sizeof(PoseGeom) * B)) return 2;
  if (dev_alloc((void **)&pb.d_totals, sizeof(int) * 8)) return 2;
  {
    const int tn = (pb.cap_n + kTX - 1) / kTX, tm = (pb.cap_n + kTY - 1) / kTY, tk = (pb.cap_n + kTZ - 1) / kTZ;
    pb.tiles_cap = tn * tm * tk;
    if (dev_alloc((void **)&pb.d_bitmap, sizeof(unsigned int) * (size_t)B * pb.tiles_cap * ((as.ngroups + 31) / 32))) return 2;
  }
  if (dev_alloc((void **)&pb.d_grid, sizeof(float) * (size_t)B * pb.cap_grid)) return 2;
  if (dev_alloc((void **)&pb.d_prep, sizeof(PrepPose) * B)) return 2;
  if (target) {
    if (dev_alloc((void **)&pb.d_tab, sizeof(geom::AxisEntry) * (size_t)B * 6 * pb.cap_o)) return 2;
    if (dev_alloc((void **)&pb.d_part, sizeof(double) * 6 * (size_t)B * pb.cap_items)) return 2;
  }
  return 0;
}

void batch_destroy(PoseBatch &pb) {
  dev_free(pb.d_tpos); dev_free(pb.d_tgroup); dev_free(pb.d_bbox); dev_free(pb.d_geom); dev_free(pb.d_totals); dev_free(pb.d_bitmap);
  dev_free(pb.d_boxes); dev_free(pb.d_gbox); dev_free(pb.d_tlist); dev_free(pb.d_tcount);
  dev_free(pb.d_grid); dev_free(pb.d_prep); dev_free(pb.d_tab); dev_free(pb.d_part); dev_free(pb.d_org0);
  pb = PoseBatch();
}

// ------------------------------------------------------------------------------------------ transform + bbox
// Matrix4::multpoint3d with the axis rotations of do_rotate (w = 1, zero translation).  Terms that
// are an exact zero (coordinate * 0, + 0) are dropped: they can only flip the sign of a zero
// result.  Every surviving product and sum rounds as in the reference (--fmad=false).
__device__ __forceinline__ float3 rigid_pose(float3 p, const PrepPose &q, float3 com) {
  // moveby(-com): TclVoltool.C:157,165
  p.x = p.x + (-1.0f * com.x); p.y = p.y + (-1.0f * com.y); p.z = p.z + (-1.0f * com.z);
  // about X (m5=c m6=s m9=-s m10=c): y' = (x*0 + y*c) + z*(-s); z' = (x*0 + y*s) + z*c
  { const float c = q.c[0], s = q.s[0];
    const float ny = p.y * c + p.z * (-s), nz = p.y * s + p.z * c;
    p.y = ny; p.z = nz; }
  // about Y (m0=c m2=-s m8=s m10=c): x' = (x*c + y*0) + z*s; z' = (x*(-s) + y*0) + z*c
  { const float c = q.c[1], s = q.s[1];
    const float nx = p.x * c + p.z * s, nz = p.x * (-s) + p.z * c;
    p.x = nx; p.z = nz; }
  // about Z (m0=c m1=s m4=-s m5=c): x' = (x*c + y*(-s)) + z*0; y' = (x*s + y*c) + z*0
  { const float c = q.c[2], s = q.s[2];
    const float nx = p.x * c + p.y * (-s), ny = p.x * s + p.y * c;
    p.x = nx; p.y = ny; }
  p.x = p.x + com.x; p.y = p.y + com.y; p.z = p.z + com.z;             // moveby(+com) :176
  p.x = p.x + q.shift[0]; p.y = p.y + q.shift[1]; p.z = p.z + q.shift[2];
  return p;
}

__global__ void bbox_reset_kernel(unsigned int *bbox, int npose) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < npose * 6) bbox[i] = (i % 6) < 3 ? 0xffffffffu : 0u;
}

__global__ void __launch_bounds__(256) transform_kernel(const float4 *__restrict__ xyzr, const float4 *__restrict__ groups,
                                                        int n, int ngroups, const PrepPose *__restrict__ prep,
                                                        float3 com, int identity, float4 *__restrict__ tpos,
                                                        float4 *__restrict__ tgroup, unsigned int *__restrict__ bbox) {
  const int pose = blockIdx.y;
  PrepPose q;
  if (!identity) q = prep[pose];
  float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float4 a = xyzr[i];
    float3 p = make_float3(a.x, a.y, a.z);
    if (!identity) p = rigid_pose(p, q, com);
    tpos[(size_t)pose * n + i] = make_float4(p.x, p.y, p.z, a.w);
    mn[0] = fminf(mn[0], p.x); mn[1] = fminf(mn[1], p.y); mn[2] = fminf(mn[2], p.z);
    mx[0] = fmaxf(mx[0], p.x); mx[1] = fmaxf(mx[1], p.y); mx[2] = fmaxf(mx[2], p.z);
  }
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += gridDim.x * blockDim.x) {
    const float4 c = groups[g];
    float3 p = make_float3(c.x, c.y, c.z);
    if (!identity) p = rigid_pose(p, q, com);
    tgroup[(size_t)pose * ngroups + g] = make_float4(p.x, p.y, p.z, c.w);
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) { mn[k] = warp_min(mn[k]); mx[k] = warp_max(mx[k]); }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      if (mn[k] <= mx[k]) {
        atomicMin(&bbox[pose * 6 + k], fkey(mn[k]));
        atomicMax(&bbox[pose * 6 + 3 + k], fkey(mx[k]));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------ per-pose geometry
struct TargetGeom {
  vt_grid grid;
  float inv[16];
  int present;
};

// the cc half of a pose's geometry: init_from_intersection(sim grid, target) (MDFF.C:132), its cell axes, the sim
// map's inverse, the work items of the batched cc kernel
__device__ __forceinline__ void setup_target(PoseGeom &G, const vt_grid &sg, bool ok, const TargetGeom &T, int cap_o) {
  G.valid = 0;
  G.cc_items = 0; G.cc_xt = 0; G.cc_zc = 0;
  for (int k = 0; k < 3; ++k) { G.no[k] = 0; G.oorg[k] = 0; G.od[k] = 0; G.sinv_d[k] = 0; G.sinv_t[k] = 0; }
  if (ok && T.present) {
    vt_grid og;
    geom::intersection_grid(sg, T.grid, og);  // MDFF.C:132 (A = sim, B = target)
    float xd[3], yd[3], zd[3], sinv[16];
    geom::cell_axes(og, xd, yd, zd);
    geom::grid_inverse(sg, false, sinv);
    const bool aligned = xd[1] == 0.f && xd[2] == 0.f && yd[0] == 0.f && yd[2] == 0.f && zd[0] == 0.f && zd[1] == 0.f;
    const bool sep = aligned && geom::diagonal_inverse(sinv) && geom::diagonal_inverse(T.inv);
    G.no[0] = og.xsize; G.no[1] = og.ysize; G.no[2] = og.zsize;
    const bool fits = og.xsize > 0 && og.ysize > 0 && og.zsize > 0 && og.xsize <= cap_o && og.ysize <= cap_o &&
                      og.zsize <= cap_o;
    if (sep && fits) {
      G.valid = 1;
      for (int k = 0; k < 3; ++k) { G.oorg[k] = og.origin[k]; G.sinv_d[k] = sinv[5 * k]; G.sinv_t[k] = sinv[12 + k]; }
      G.od[0] = xd[0]; G.od[1] = yd[1]; G.od[2] = zd[2];
      G.cc_xt = (og.xsize + 31) / 32;
      G.cc_zc = (og.zsize + kZChunkFit - 1) / kZChunkFit;
      G.cc_items = G.cc_xt * ((og.ysize + kRowsFit - 1) / kRowsFit) * G.cc_zc;
    }
  }
}

__global__ void __launch_bounds__(256) pose_setup_kernel(const unsigned int *__restrict__ bbox, int npose, float pad, float gs, int cap_n,
                                  int cap_o, TargetGeom T, PoseGeom *__restrict__ out, int *__restrict__ totals) {
  for (int p = threadIdx.x; p < npose; p += blockDim.x) {
    PoseGeom G;
    vt_grid sg;
    bool ok = true;
    for (int k = 0; k < 3; ++k) {
      // oracle vo_sim_grid: pad the box, count voxels, snap the far end to the lattice
      float mn = fkey_inv(bbox[p * 6 + k]), mx = fkey_inv(bbox[p * 6 + 3 + k]);
      mn -= pad;
      mx += pad;
      float ax = mx - mn;
      const int nv = (int)ceil((double)(ax / gs));
      ax = (float)(nv - 1) * gs;
      G.org[k] = mn;
      G.nv[k] = nv;
      if (nv < 2 || nv > cap_n) ok = false;
      sg.origin[k] = (double)mn;
      sg.xaxis[k] = sg.yaxis[k] = sg.zaxis[k] = 0.0;
      if (k == 0) sg.xaxis[0] = (double)ax;
      if (k == 1) sg.yaxis[1] = (double)ax;
      if (k == 2) sg.zaxis[2] = (double)ax;
    }
    sg.xsize = G.nv[0]; sg.ysize = G.nv[1]; sg.zsize = G.nv[2];
    G.tiles[0] = (G.nv[0] + kTX - 1) / kTX;
    G.tiles[1] = (G.nv[1] + kTY - 1) / kTY;
    G.tiles[2] = (G.nv[2] + kTZ - 1) / kTZ;
    if (!ok) { G.tiles[0] = G.tiles[1] = G.tiles[2] = 0; }
    setup_target(G, sg, ok, T, cap_o);
    out[p] = G;
  }
  // exclusive prefixes (npose <= a few thousand: one thread, after everyone has written)
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0, c = 0;
    for (int i = 0; i < npose; ++i) {
      out[i].tile_base = t;
      out[i].cc_base = c;
      t += out[i].tiles[0] * out[i].tiles[1] * out[i].tiles[2];
      c += out[i].cc_items;
    }
    totals[0] = t;
    totals[1] = c;
    totals[3] = 0;
    totals[4] = 0;   // work queues of the splat, cc and gather kernels of this batch (warps draw tasks with atomicAdd)
    totals[5] = 0;
    totals[6] = 0;
  }
}

__global__ void __launch_bounds__(128) pose_tables_kernel(const PoseGeom *__restrict__ geo, TargetGeom T, int cap_o,
                                                          geom::AxisEntry *__restrict__ tab) {
  const int p = blockIdx.x;
  const PoseGeom &G = geo[p];
  if (!G.valid) return;
  const int nsrcA[3] = {G.nv[0], G.nv[1], G.nv[2]};
  const int nsrcB[3] = {T.grid.xsize, T.grid.ysize, T.grid.zsize};
  for (int t = threadIdx.x; t < 6 * cap_o; t += blockDim.x) {
    const int which = t / cap_o, g = t - which * cap_o;
    const int ax = which % 3;
    if (g >= G.no[ax]) continue;
    geom::AxisEntry e;
    if (which < 3) e = geom::axis_entry(ax, g, G.oorg[ax], G.od[ax], G.sinv_d[ax], G.sinv_t[ax], nsrcA[ax]);
    else e = geom::axis_entry(ax, g, G.oorg[ax], G.od[ax], T.inv[5 * ax], T.inv[12 + ax], nsrcB[ax]);
    tab[((size_t)p * 6 + which) * cap_o + g] = e;
  }
}

// The density map of a rigidly TRANSLATED structure is the same array of voxels on a translated frame: QuickSurf's
// grid hangs on the atoms' bounding box (origin = min - pad), so moving every atom by t moves the origin by t and
// leaves every (atom - origin) difference -- hence every voxel -- unchanged up to the rounding of the additions.
// pose_retarget_kernel re-aims the already splatted grids of a batch at the frame shifted by `shift` from where they
// were splatted (org0, saved on the first call): new origin, new intersection with the target, new cc work items;
// pose_tables_kernel and the cc kernel then score the translated candidates without a second splat.
__global__ void __launch_bounds__(256) pose_retarget_kernel(PoseGeom *__restrict__ geo, float *__restrict__ org0, int npose, float sx,
                                                            float sy, float sz, int save, float gs, int cap_o, TargetGeom T,
                                                            int *__restrict__ totals) {
  const float shift[3] = {sx, sy, sz};
  for (int p = threadIdx.x; p < npose; p += blockDim.x) {
    PoseGeom G = geo[p];
    vt_grid sg;
    const bool ok = G.tiles[0] > 0;
    for (int k = 0; k < 3; ++k) {
      if (save) org0[3 * p + k] = G.org[k];
      G.org[k] = org0[3 * p + k] + shift[k];
      sg.origin[k] = (double)G.org[k];
      sg.xaxis[k] = sg.yaxis[k] = sg.zaxis[k] = 0.0;
      const float ax = (float)(G.nv[k] - 1) * gs;
      if (k == 0) sg.xaxis[0] = (double)ax;
      if (k == 1) sg.yaxis[1] = (double)ax;
      if (k == 2) sg.zaxis[2] = (double)ax;
    }
    sg.xsize = G.nv[0]; sg.ysize = G.nv[1]; sg.zsize = G.nv[2];
    setup_target(G, sg, ok, T, cap_o);
    geo[p] = G;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int c = 0;
    for (int i = 0; i < npose; ++i) { geo[i].cc_base = c; c += geo[i].cc_items; }
    totals[1] = c;
    totals[5] = 0;   // the cc kernel's task queue
  }
}

int batch_retarget(const AtomSet &as, PoseBatch &pb, int npose, const float shift[3], bool first, const vt_grid *target) {
  VT_REQUIRE_CTX();
  Ctx &c = ctx();
  if (!target || npose > pb.B || npose > 1024) { set_error("batch_retarget: bad arguments"); return 1; }
  if (!pb.d_org0 && dev_alloc((void **)&pb.d_org0, sizeof(float) * 3 * (size_t)pb.B)) return 2;
  TargetGeom T;
  memset(&T, 0, sizeof(T));
  T.grid = *target;
  grid_inverse(*target, false, T.inv);
  T.present = 1;
  pose_retarget_kernel<<<1, 256, 0, c.stream>>>(pb.d_geom, pb.d_org0, npose, shift[0], shift[1], shift[2], first ? 1 : 0, as.prm.gs,
                                                 pb.cap_o, T, pb.d_totals);
  VT_LAUNCHED();
  pose_tables_kernel<<<npose, 128, 0, c.stream>>>(pb.d_geom, T, pb.cap_o, pb.d_tab);
  VT_LAUNCHED();
  return 0;
}

int batch_prepare(const AtomSet &as, PoseBatch &pb, int npose, const PrepPose *d_prep, const float com[3],
                  int identity, const vt_grid *target) {
  VT_REQUIRE_CTX();
  Ctx &c = ctx();
  if (npose > pb.B) { set_error("batch_prepare: %d poses > %d slots", npose, pb.B); return 1; }
  bbox_reset_kernel<<<(npose * 6 + 127) / 128, 128, 0, c.stream>>>(pb.d_bbox, npose);
  VT_LAUNCHED();
  int gx = (as.n + 255) / 256;
  if (gx > 64) gx = 64;
  prof_begin(PK_TRANSFORM);
  transform_kernel<<<dim3(gx, npose), 256, 0, c.stream>>>(as.d_xyzr, as.d_group, as.n, as.ngroups, d_prep,
                                                           make_float3(com[0], com[1], com[2]), identity, pb.d_tpos,
                                                           pb.d_tgroup, pb.d_bbox);
  prof_end(PK_TRANSFORM);
  VT_LAUNCHED();
  TargetGeom T;
  memset(&T, 0, sizeof(T));
  if (target) {
    T.grid = *target;
    grid_inverse(*target, false, T.inv);
    T.present = 1;
  }
  if (npose > 1024) { set_error("batch_prepare: at most 1024 poses per batch"); return 1; }
  pose_setup_kernel<<<1, 256, 0, c.stream>>>(pb.d_bbox, npose, as.pad, as.prm.gs, pb.cap_n, pb.cap_o, T, pb.d_geom,
                                              pb.d_totals);
  VT_LAUNCHED();
  if (target) {
    pose_tables_kernel<<<npose, 128, 0, c.stream>>>(pb.d_geom, T, pb.cap_o, pb.d_tab);
    VT_LAUNCHED();
  }
  return 0;
}

// ------------------------------------------------------------------------------------------ splat
__device__ __forceinline__ int find_pose(const PoseGeom *__restrict__ geo, int npose, int item, bool cc) {
  int lo = 0, hi = npose - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    const int base = cc ? geo[mid].cc_base : geo[mid].tile_base;
    if (base <= item) lo = mid; else hi = mid - 1;
  }
  return lo;
}

constexpr int kWRec = 20;     // floats per warp-staged atom: Ex[8] | dy^2[8] | dz^2[4]
constexpr int kGChunk = 1024; // groups expanded from the bitmap at a time (32 words, one warp)
constexpr int kWGroups = 48;  // groups a warp gathers per batch
constexpr int kWList = kWGroups * 32;  // per-warp gathered atoms per batch

// Which groups can touch which tile: thread per (pose, group) marks the tiles its bounding sphere
// reaches in a per-tile bitmap (atomicOr: the result does not depend on the order).  The splat
// kernel then walks set bits in index order, so its accumulation order stays deterministic.
__global__ void __launch_bounds__(256) pose_bin_kernel(const PoseGeom *__restrict__ geo, int ngroups,
                                                       const float4 *__restrict__ tgroup_all,
                                                       const float *__restrict__ gcut, float gs, float invgs,
                                                       unsigned int *__restrict__ bitmap, int tiles_cap, int nwords) {
  const int p = blockIdx.y;
  const PoseGeom &G = geo[p];
  const int ntx = G.tiles[0], nty = G.tiles[1], ntz = G.tiles[2];
  if (ntx == 0) return;
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += gridDim.x * blockDim.x) {
    const float4 c = tgroup_all[(size_t)p * ngroups + g];
    const float rx = (c.x - G.org[0]) * invgs, ry = (c.y - G.org[1]) * invgs, rz = (c.z - G.org[2]) * invgs;
    const float cut = gcut[g];
    const float rr = (c.w + cut) * invgs + 2.0f;  // reach in voxels + slack; the exact test below decides
    const int tx0 = max(0, (int)floorf((rx - rr) / kTX) - 1), tx1 = min(ntx - 1, (int)floorf((rx + rr) / kTX) + 1);
    const int ty0 = max(0, (int)floorf((ry - rr) / kTY) - 1), ty1 = min(nty - 1, (int)floorf((ry + rr) / kTY) + 1);
    const int tz0 = max(0, (int)floorf((rz - rr) / kTZ) - 1), tz1 = min(ntz - 1, (int)floorf((rz + rr) / kTZ) + 1);
    for (int tz = tz0; tz <= tz1; ++tz)
      for (int ty = ty0; ty <= ty1; ++ty)
        for (int tx = tx0; tx <= tx1; ++tx) {
          // tile box in cartesian space with one voxel of slack for the integer truncations
          const int x0 = tx * kTX, y0 = ty * kTY, z0 = tz * kTZ;
          const float bx0 = G.org[0] + (x0 - 1) * gs - cut, bx1 = G.org[0] + (x0 + kTX) * gs + cut;
          const float by0 = G.org[1] + (y0 - 1) * gs - cut, by1 = G.org[1] + (y0 + kTY) * gs + cut;
          const float bz0 = G.org[2] + (z0 - 1) * gs - cut, bz1 = G.org[2] + (z0 + kTZ) * gs + cut;
          const float dx = fmaxf(fmaxf(bx0 - c.x, 0.0f), c.x - bx1);
          const float dy = fmaxf(fmaxf(by0 - c.y, 0.0f), c.y - by1);
          const float dz = fmaxf(fmaxf(bz0 - c.z, 0.0f), c.z - bz1);
          if (dx * dx + dy * dy + dz * dz <= c.w * c.w) {
            const int tile = (tz * nty + ty) * ntx + tx;
            atomicOr(&bitmap[((size_t)p * tiles_cap + tile) * nwords + (g >> 5)], 1u << (g & 31));
          }
        }
  }
}

__global__ void __launch_bounds__(32 * kSplatWarps, (kSplatWarps == 4 ? 6 : 3)) splat_kernel(const PoseGeom *__restrict__ geo, const int *__restrict__ totals,
                                                       int npose, int natoms, int ngroups,
                                                       const float4 *__restrict__ tpos_all,
                                                       const unsigned int *__restrict__ bitmap, int tiles_cap, int nwords,
                                                       const float4 *__restrict__ cst, float gs, float invgs,
                                                       float *__restrict__ grid_all, long cap_grid) {
  extern __shared__ __align__(16) unsigned char s_dyn[];
  float(*s_rec)[32 * kWRec] = reinterpret_cast<float(*)[32 * kWRec]>(s_dyn);          // warp-private staged records
  constexpr int W = kSplatWarps;
  unsigned short(*s_wid)[kWList] = reinterpret_cast<unsigned short(*)[kWList]>(s_dyn + sizeof(float) * W * 32 * kWRec);
  unsigned char(*s_wm)[kWList] = reinterpret_cast<unsigned char(*)[kWList]>(s_dyn + sizeof(float) * W * 32 * kWRec +
                                                                            sizeof(unsigned short) * W * kWList);
  // s_wid: atoms gathered by warp w whose box meets this tile, as (index in s_glist) << 5 | lane;
  // s_wm: which of the 8 warp regions they meet
  __shared__ float s_ar[W][32], s_rl2[W][32];           // arinv, (gausslim*sigma)^2 per staged atom
  __shared__ int s_pend[W][64];                         // warp-private pending ids
  __shared__ int s_wn[W];
  __shared__ unsigned short s_glist[kGChunk];           // groups (offsets inside the chunk) that can reach the tile
  __shared__ int s_scan[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned int lt = (1u << lane) - 1u;
  // region of warp w: x half (w & 1), z half ((w >> 1) & 1), y block (w >> 2)
  const int wx = warp & 1, wz = (warp >> 1) & 1, wy = warp >> 2;
  const int total = totals[0];

  for (int item = blockIdx.x; item < total; item += gridDim.x) {
    const int p = find_pose(geo, npose, item, false);
    const PoseGeom &G = geo[p];
    const int local = item - G.tile_base;
    const int tx = local % G.tiles[0], ty = (local / G.tiles[0]) % G.tiles[1], tz = local / (G.tiles[0] * G.tiles[1]);
    const int x0 = tx * kTX, y0 = ty * kTY, z0 = tz * kTZ;
    const int nv0 = G.nv[0], nv1 = G.nv[1], nv2 = G.nv[2];
    const float ox = G.org[0], oy = G.org[1], oz = G.org[2];
    const float4 *__restrict__ tpos = tpos_all + (size_t)p * natoms;
    const unsigned int *__restrict__ bits = bitmap + ((size_t)p * tiles_cap + local) * nwords;
    // this warp's 8x8x4 region and this thread's column inside it
    const int X0 = x0 + wx * 8, Y0 = y0 + wy * 8, Z0 = z0 + wz * 4;
    const int ly = lane & 7, lz = lane >> 3;

    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.0f;

    // ---- phase 2 (warp-independent): stage 32 atoms (lane per atom), accumulate ----
    auto process = [&](int n) {
      float *rec = s_rec[warp];
      if (lane < n) {
        const int id = s_pend[warp][lane];
        const float4 a4 = tpos[id];
        const float4 k4 = cst[id];
        const float ax = a4.x - ox, ay = a4.y - oy, az = a4.z - oz;  // xyzr = pos - origin
        const float arinv = k4.x, rl = k4.z;
        float t = ax * invgs;
        const int xmin = max(__float2int_rz(t - rl), 0), xmax = min(__float2int_rz(t + rl), nv0 - 1);
        t = ay * invgs;
        const int ymin = max(__float2int_rz(t - rl), 0), ymax = min(__float2int_rz(t + rl), nv1 - 1);
        t = az * invgs;
        const int zmin = max(__float2int_rz(t - rl), 0), zmax = min(__float2int_rz(t + rl), nv2 - 1);
        float *r = rec + lane * kWRec;
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int X = X0 + k;
          const float dx = (float)X * gs - ax;
          const float e = ex2_approx((dx * dx) * arinv);
          v[k] = (X >= xmin && X <= xmax) ? e : 0.0f;
        }
        *reinterpret_cast<float4 *>(r) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4 *>(r + 4) = make_float4(v[4], v[5], v[6], v[7]);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int Y = Y0 + k;
          const float dy = (float)Y * gs - ay;
          v[k] = (Y >= ymin && Y <= ymax) ? dy * dy : INFINITY;
        }
        *reinterpret_cast<float4 *>(r + 8) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4 *>(r + 12) = make_float4(v[4], v[5], v[6], v[7]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int Z = Z0 + k;
          const float dz = (float)Z * gs - az;
          v[k] = (Z >= zmin && Z <= zmax) ? dz * dz : INFINITY;
        }
        *reinterpret_cast<float4 *>(r + 16) = make_float4(v[0], v[1], v[2], v[3]);
        s_ar[warp][lane] = arinv;
        s_rl2[warp][lane] = k4.y;
      }
      __syncwarp();
      for (int a = 0; a < n; ++a) {
        const float *r = rec + a * kWRec;
        const float d2 = r[8 + ly] + r[16 + lz];               // dy*dy + dz*dz
        const bool pass = d2 < s_rl2[warp][a];                 // !(dy2dz2 >= radlim2)
        if (!__any_sync(0xffffffffu, pass)) continue;
        const float e = pass ? ex2_approx(d2 * s_ar[warp][a]) : 0.0f;
        const float4 e0 = *reinterpret_cast<const float4 *>(r);
        const float4 e1 = *reinterpret_cast<const float4 *>(r + 4);
        acc[0] = __fmaf_rn(e0.x, e, acc[0]); acc[1] = __fmaf_rn(e0.y, e, acc[1]);
        acc[2] = __fmaf_rn(e0.z, e, acc[2]); acc[3] = __fmaf_rn(e0.w, e, acc[3]);
        acc[4] = __fmaf_rn(e1.x, e, acc[4]); acc[5] = __fmaf_rn(e1.y, e, acc[5]);
        acc[6] = __fmaf_rn(e1.z, e, acc[6]); acc[7] = __fmaf_rn(e1.w, e, acc[7]);
      }
      __syncwarp();
    };

    // every warp walks the 8 gathered lists in the same order, keeping the atoms whose box meets
    // its own region
    auto drain = [&](int gbase) {
      int have = 0;
      for (int l = 0; l < W; ++l) {
        const int count = s_wn[l];
        for (int b0 = 0; b0 < count; b0 += 32) {
          const int i = b0 + lane;
          const bool mine = i < count && ((s_wm[l][i] >> warp) & 1);
          const unsigned int bal = __ballot_sync(0xffffffffu, mine);
          if (bal == 0u) continue;
          if (mine) {
            const int e = s_wid[l][i];
            s_pend[warp][have + __popc(bal & lt)] = (gbase + (int)s_glist[e >> 5]) * 32 + (e & 31);
          }
          have += __popc(bal);
          __syncwarp();
          if (have >= 32) {
            process(32);
            const int rest = have - 32;
            const int keep = lane < rest ? s_pend[warp][32 + lane] : 0;
            __syncwarp();
            if (lane < rest) s_pend[warp][lane] = keep;
            have = rest;
            __syncwarp();
          }
        }
      }
      if (have > 0) process(have);
    };

    // ---- phase 1: expand the tile's group bitmap, gather atoms (warp per group) ----
    for (int w0 = 0; w0 < nwords; w0 += kGChunk / 32) {
      // 32 words -> ordered list of set bits (warp 0)
      if (warp == 0) {
        unsigned int word = (w0 + lane < nwords) ? bits[w0 + lane] : 0u;
        const int pc = __popc(word);
        int incl = pc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += t;
        }
        if (lane == 31) s_scan[0] = incl;
        int off = incl - pc;
        while (word) {
          const int bit = __ffs(word) - 1;
          word &= word - 1;
          s_glist[off++] = (unsigned short)(lane * 32 + bit);
        }
      }
      __syncthreads();
      const int nhit = s_scan[0];
      for (int b0 = 0; b0 < nhit; b0 += W * kWGroups) {
        // each warp tests the atoms of up to kWGroups groups with the reference's truncated integer box
        int cnt = 0;
        for (int j = 0; j < kWGroups; ++j) {
          const int gi = b0 + warp + W * j;
          if (gi >= nhit) break;
          const int id = (w0 * 32 + (int)s_glist[gi]) * 32 + lane;
          bool take = false;
          int mask = 0;
          if (id < natoms) {
            const float4 a4 = tpos[id];
            const float rl = cst[id].z;
            const float ax = a4.x - ox, ay = a4.y - oy, az = a4.z - oz;
            float t = ax * invgs;
            const int xmin = max(__float2int_rz(t - rl), 0), xmax = min(__float2int_rz(t + rl), nv0 - 1);
            t = ay * invgs;
            const int ymin = max(__float2int_rz(t - rl), 0), ymax = min(__float2int_rz(t + rl), nv1 - 1);
            t = az * invgs;
            const int zmin = max(__float2int_rz(t - rl), 0), zmax = min(__float2int_rz(t + rl), nv2 - 1);
            take = xmin <= x0 + kTX - 1 && xmax >= x0 && ymin <= y0 + kTY - 1 && ymax >= y0 && zmin <= z0 + kTZ - 1 &&
                   zmax >= z0 && xmin <= xmax && ymin <= ymax && zmin <= zmax;
            if (take) {
              // bit w: box meets the region of warp w (x half w&1, z half (w>>1)&1, y block w>>2)
              const int mx = (xmin <= x0 + 7 ? 1 : 0) | (xmax >= x0 + 8 ? 2 : 0);
              const int mz = (zmin <= z0 + 3 ? 1 : 0) | (zmax >= z0 + 4 ? 2 : 0);
#pragma unroll
              for (int w = 0; w < W; ++w) {
                const int yb = y0 + 8 * (w >> 2);
                if (((mx >> (w & 1)) & 1) && ((mz >> ((w >> 1) & 1)) & 1) && ymin <= yb + 7 && ymax >= yb) mask |= 1 << w;
              }
            }
          }
          const unsigned int b2 = __ballot_sync(0xffffffffu, take);
          if (take) {
            const int slot = cnt + __popc(b2 & lt);
            s_wid[warp][slot] = (unsigned short)((gi << 5) | lane);
            s_wm[warp][slot] = (unsigned char)mask;
          }
          cnt += __popc(b2);
        }
        if (lane == 0) s_wn[warp] = cnt;
        __syncthreads();
        drain(w0 * 32);
        __syncthreads();
      }
    }

    // ---- every voxel of the tile written exactly once ----
    const int Y = Y0 + ly, Z = Z0 + lz;
    if (Y < nv1 && Z < nv2) {
      float *row = grid_all + (size_t)p * cap_grid + ((size_t)Z * nv1 + Y) * nv0;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (X0 + k < nv0) row[X0 + k] = acc[k];
    }
  }
}

int batch_bin(const AtomSet &as, PoseBatch &pb, int npose) {
  VT_REQUIRE_CTX();
  Ctx &c = ctx();
  const int nwords = (as.ngroups + 31) / 32;
  prof_begin(PK_SETUP);
  VT_CUDA(cudaMemsetAsync(pb.d_bitmap, 0, sizeof(unsigned int) * (size_t)npose * pb.tiles_cap * nwords, c.stream));
  int gx = (as.ngroups + 255) / 256;
  if (gx > 32) gx = 32;
  pose_bin_kernel<<<dim3(gx, npose), 256, 0, c.stream>>>(pb.d_geom, as.ngroups, pb.d_tgroup, as.d_gcut, as.prm.gs, as.prm.invgs,
                                                          pb.d_bitmap, pb.tiles_cap, nwords);
  prof_end(PK_SETUP);
  VT_LAUNCHED();
  return 0;
}

int batch_lists_front(const AtomSet &as, PoseBatch &pb, int npose);
int batch_lists_back(const AtomSet &as, PoseBatch &pb, int npose);

static bool lists_on(const PoseBatch &pb) { return pb.use_lists && !getenv("VOLTOOL_SPLAT_COOP"); }

int batch_splat_front(const AtomSet &as, PoseBatch &pb, int npose) {
  VT_REQUIRE_CTX();
  if (lists_on(pb)) return batch_lists_front(as, pb, npose);   // boxes -> tile bitmaps from the group boxes -> lists
  return batch_bin(as, pb, npose);                              // tile bitmaps from the group spheres
}

int batch_splat_back(const AtomSet &as, PoseBatch &pb, int npose) {
  VT_REQUIRE_CTX();
  Ctx &c = ctx();
  if (lists_on(pb)) return batch_lists_back(as, pb, npose);
  const int nwords = (as.ngroups + 31) / 32;
  const int grid = c.sm_count * (kSplatWarps == 4 ? 6 : 3) * 4;  // persistent: resident CTAs per SM x 4 rounds
  prof_begin(PK_SPLAT);
  const size_t smem = sizeof(float) * kSplatWarps * 32 * kWRec + (sizeof(unsigned short) + 1) * kSplatWarps * kWList;
  static bool attr = false;
  if (!attr) {
    VT_CUDA(cudaFuncSetAttribute(splat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  splat_kernel<<<grid, 32 * kSplatWarps, smem, c.stream>>>(pb.d_geom, pb.d_totals, npose, as.n, as.ngroups, pb.d_tpos, pb.d_bitmap,
                                           pb.tiles_cap, nwords, as.d_const, as.prm.gs, as.prm.invgs, pb.d_grid, pb.cap_grid);
  prof_end(PK_SPLAT);
  VT_LAUNCHED();
  return 0;
}

int batch_splat(const AtomSet &as, PoseBatch &pb, int npose) {
  int rc = batch_splat_front(as, pb, npose);
  if (!rc) rc = batch_splat_back(as, pb, npose);
  return rc;
}

// ------------------------------------------------------------------------------------------ batched cc
// 128 threads, 8 CTAs per SM (64 registers, 8 bytes of spill): the kernel waits on L2 gathers, and with the task
// queue 32 resident warps per SM beat the 16 that 126 registers allowed (measured 2 x 256 / 5 / 6 / 7 / 8 / 10 x
// 128 threads: 38.5 / 39.8 / 40.4 / 41.3 / 41.3 / 40.7 k poses/s)
__global__ void __launch_bounds__(128, 8) pose_cc_kernel(const PoseGeom *__restrict__ geo, const int *__restrict__ totals,
                                                      int npose, const float *__restrict__ grid_all, long cap_grid,
                                                      const float *__restrict__ tdata, int tnx, int tny,
                                                      const geom::AxisEntry *__restrict__ tab_all, int cap_o,
                                                      long cap_items, float thr_f, double *__restrict__ part, int *queues) {
  const int lane = threadIdx.x & 31;
  const int total = totals[1];
  int *queue = queues + 1;
  while (true) {
    // items cost very different amounts (empty corners vs the dense interior): warps draw them from a queue
    int item = 0;
    if (lane == 0) item = atomicAdd(queue, 1);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= total) break;
    const int p = find_pose(geo, npose, item, true);
    const PoseGeom &G = geo[p];
    const int local = item - G.cc_base;
    const int xb = local % G.cc_xt;
    const int r = local / G.cc_xt;
    const int ygroups = (G.no[1] + kRowsFit - 1) / kRowsFit;
    const int yg = r % ygroups, zc = r / ygroups;
    const geom::AxisEntry *__restrict__ T = tab_all + (size_t)p * 6 * cap_o;
    CcAcc a;
#pragma unroll
    for (int k = 0; k < 5; ++k) a.s[k] = 0.0;
    a.n = 0;
    cc_march_item<kRowsFit>(grid_all + (size_t)p * cap_grid, G.nv[0], G.nv[1], tdata, tnx, tny, T, T + cap_o, T + 2 * cap_o,
                  T + 3 * cap_o, T + 4 * cap_o, T + 5 * cap_o, G.no[0], G.no[1], xb * 32 + lane, yg * kRowsFit,
                  zc * kZChunkFit, min(G.no[2], (zc + 1) * kZChunkFit), thr_f, a);
    double v[6] = {a.s[0], a.s[1], a.s[2], a.s[3], a.s[4], (double)a.n};
#pragma unroll
    for (int k = 0; k < 6; ++k) v[k] = warp_sum(v[k]);
    if (lane == 0) {
      double *o = part + ((size_t)p * cap_items + local) * 6;
#pragma unroll
      for (int k = 0; k < 6; ++k) o[k] = v[k];
    }
  }
}

// one block per pose: ordered sum of the warp partials, then MDFF.C:143-149
__global__ void __launch_bounds__(256) pose_cc_finish_kernel(const PoseGeom *__restrict__ geo,
                                                             const double *__restrict__ part, long cap_items,
                                                             float *__restrict__ cc_out, double *__restrict__ sums_out) {
  __shared__ double sm[6 * 32];
  const int p = blockIdx.x;
  const PoseGeom &G = geo[p];
  double acc[6] = {0, 0, 0, 0, 0, 0};
  const double *src = part + (size_t)p * cap_items * 6;
  for (int i = threadIdx.x; i < G.cc_items; i += blockDim.x)
#pragma unroll
    for (int k = 0; k < 6; ++k) acc[k] += src[(size_t)i * 6 + k];
  block_sum<6>(acc, sm);
  if (threadIdx.x == 0) {
    float cc = quiet_nan();
    if (G.valid) {
      const int size = (int)acc[5];
      const double mA = acc[0] / size, mB = acc[1] / size;
      const double sdA = sqrt(acc[2] / size - mA * mA);
      const double sdB = sqrt(acc[3] / size - mB * mB);
      const double c = (acc[4] - size * mA * mB) / (size * sdA * sdB);
      cc = (float)c;  // calc_cc: float return_cc += cc (TclVoltool.C:75,126)
    }
    cc_out[p] = cc;
    if (sums_out)
#pragma unroll
      for (int k = 0; k < 6; ++k) sums_out[(size_t)p * 6 + k] = acc[k];
  }
}

int batch_cc(PoseBatch &pb, int npose, const vt_map *target, double threshold, float *d_cc, double *d_sums) {
  VT_REQUIRE_CTX();
  Ctx &c = ctx();
  prof_begin(PK_POSE_CC);
  static int cc_per_sm = getenv("VOLTOOL_CC_CTAS") ? atoi(getenv("VOLTOOL_CC_CTAS")) : 8;
  pose_cc_kernel<<<c.sm_count * cc_per_sm, 128, 0, c.stream>>>(pb.d_geom, pb.d_totals, npose, pb.d_grid, pb.cap_grid, target->d,
                                                       target->grid.xsize, target->grid.ysize, pb.d_tab, pb.cap_o,
                                                       pb.cap_items, threshold_as_float(threshold), pb.d_part, pb.d_totals + 4);
  prof_end(PK_POSE_CC);
  VT_LAUNCHED();
  pose_cc_finish_kernel<<<npose, 256, 0, c.stream>>>(pb.d_geom, pb.d_part, pb.cap_items, d_cc, d_sums);
  VT_LAUNCHED();
  return 0;
}

}  // namespace vt
