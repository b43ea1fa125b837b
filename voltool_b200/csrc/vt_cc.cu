// vt_cc.cu -- thresholded cross correlation (cc_threaded, MDFF.C:55-151) and the two-map ops
// add/subtract/multiply/average (Voltool.C:249-377).
//
// Both walk an output grid (intersection or union of the two inputs), turn each voxel index into a
// cartesian point, and sample both source maps there.  Two code paths, selected on the host:
//
//   generic   any cell: the literal per-voxel arithmetic (voxel_coord -> Matrix4 product ->
//             truncation -> 8-neighbour trilinear / nearest).
//   separable orthogonal axis-aligned cells (the only case the reference supports, Voltool.C:55-56)
//             where every off-diagonal term of the chain is an exact zero: voxel coordinate,
//             truncated index, clamped neighbours and fraction depend on one axis index each, so
//             they are tabulated per axis once (same float operations, same order => same bits)
//             and the inner loop is table-driven.  Threads run along x and march in z, reusing
//             the bilinear value of the shared plane between consecutive z steps when the index
//             advances by one.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "vt_device.cuh"
#include "vt_geom.h"
#include "vt_sep.cuh"

namespace vt {

constexpr int kThreads = 256;

static void to_dev(const vt_map *m, bool shifted, DevMap &d) {
  MapGeom g;
  map_geom(m->grid, g);
  d.d = m->d;
  for (int i = 0; i < 16; ++i) d.inv[i] = shifted ? g.inv1[i] : g.inv0[i];
  d.nx = g.nx; d.ny = g.ny; d.nz = g.nz;
}
static void to_dev(const vt_grid &grid, DevCoord &c) {
  CoordGeom g;
  coord_geom(grid, g);
  c.ox = g.origin[0]; c.oy = g.origin[1]; c.oz = g.origin[2];
  for (int k = 0; k < 3; ++k) { c.xd[k] = g.xd[k]; c.yd[k] = g.yd[k]; c.zd[k] = g.zd[k]; }
  c.nx = g.nx; c.ny = g.ny; c.nz = g.nz;
}

// ------------------------------------------------------------------------------------------ generic CC
__global__ void __launch_bounds__(kThreads) cc_generic_kernel(DevMap A, DevMap B, DevCoord C, double thr,
                                                              double *partials, unsigned int *ticket, double *out) {
  __shared__ double sm[6 * 32];
  double acc[6] = {0, 0, 0, 0, 0, 0};  // sumA sumB sumAA sumBB sumAB count
  const long rows = (long)C.ny * C.nz;
  for (long row = blockIdx.x; row < rows; row += gridDim.x) {
    const int gy = (int)(row % C.ny), gz = (int)(row / C.ny);
    for (int gx = threadIdx.x; gx < C.nx; gx += blockDim.x) {
      float x, y, z;
      voxel_coord(C, gx, gy, gz, x, y, z);
      const float vA = sample_interp<false>(A, x, y, z);
      if (is_nan_bits(vA) || !((double)vA >= thr)) continue;  // MDFF.C:73-77 (B is only needed past this)
      const float vB = sample_interp<false>(B, x, y, z);
      if (is_nan_bits(vB)) continue;
      acc[0] += (double)vA;
      acc[1] += (double)vB;
      acc[2] += (double)(vA * vA);
      acc[3] += (double)(vB * vB);
      acc[4] += (double)(vA * vB);
      acc[5] += 1.0;
    }
  }
  block_sum<6>(acc, sm);
  grid_finish<6>(acc, partials, ticket, out, sm);
}

// ------------------------------------------------------------------------------------------ separable path
using geom::AxisEntry;
using geom::diagonal_inverse;

// Host: can the chain voxel_coord -> cart_to_voxel be split per axis with identical bits?
// Requires (1) output cell axes with exact-zero off-diagonals, (2) an inverse whose off-diagonal
// 3x3 terms and projective column are exact zeros with m[15] == 1 (geom::diagonal_inverse).
static bool axis_aligned_out(const DevCoord &c) {
  return c.xd[1] == 0.f && c.xd[2] == 0.f && c.yd[0] == 0.f && c.yd[2] == 0.f && c.zd[0] == 0.f && c.zd[1] == 0.f;
}

static void build_axis(const DevCoord &c, int axis, const float *inv, int n_src, int n_out, std::vector<AxisEntry> &tab) {
  tab.resize(n_out > 0 ? n_out : 1);
  const double o = axis == 0 ? c.ox : (axis == 1 ? c.oy : c.oz);
  const float delta = axis == 0 ? c.xd[0] : (axis == 1 ? c.yd[1] : c.zd[2]);
  for (int g = 0; g < n_out; ++g) tab[g] = geom::axis_entry(axis, g, o, delta, inv[5 * axis], inv[12 + axis], n_src);
}

// The split is only bit-identical if the foreign-axis cartesian coordinates cannot leak in:
// with exact-zero off-diagonals the products are +-0 unless a coordinate is inf/NaN.
static bool finite_geom(const DevCoord &c) {
  return std::isfinite(c.ox) && std::isfinite(c.oy) && std::isfinite(c.oz) && std::isfinite(c.xd[0]) &&
         std::isfinite(c.yd[1]) && std::isfinite(c.zd[2]);
}

struct SepTables {
  AxisEntry *d = nullptr;  // [Ax | Ay | Az | Bx | By | Bz]
  int offAx, offAy, offAz, offBx, offBy, offBz;
  std::vector<AxisEntry> host;
  int off[6];
};

int k_cc_staged(const DevMap &A, const DevMap &B, int nx, int ny, int nz, const std::vector<AxisEntry> &host_tab,
                const int off[6], const AxisEntry *d_tab, float thr_f, double *partials, unsigned int *ticket,
                double *dout);

static int upload_tables(const DevCoord &C, const DevMap &A, const DevMap &B, SepTables &T) {
  std::vector<AxisEntry> &all = T.host;
  std::vector<AxisEntry> t;
  all.clear();
  const DevMap *maps[2] = {&A, &B};
  int offs[6];
  for (int m = 0; m < 2; ++m) {
    const int nsrc[3] = {maps[m]->nx, maps[m]->ny, maps[m]->nz};
    const int nout[3] = {C.nx, C.ny, C.nz};
    for (int ax = 0; ax < 3; ++ax) {
      build_axis(C, ax, maps[m]->inv, nsrc[ax], nout[ax], t);
      offs[m * 3 + ax] = (int)all.size();
      all.insert(all.end(), t.begin(), t.begin() + (nout[ax] > 0 ? nout[ax] : 0));
    }
  }
  T.offAx = offs[0]; T.offAy = offs[1]; T.offAz = offs[2];
  T.offBx = offs[3]; T.offBy = offs[4]; T.offBz = offs[5];
  for (int k = 0; k < 6; ++k) T.off[k] = offs[k];
  if (dev_alloc((void **)&T.d, sizeof(AxisEntry) * (all.size() + 1))) return 2;
  VT_CUDA(cudaMemcpyAsync(T.d, all.data(), sizeof(AxisEntry) * all.size(), cudaMemcpyHostToDevice, ctx().stream));
  VT_CUDA(cudaStreamSynchronize(ctx().stream));  // pageable source
  return 0;
}

// ------------------------------------------------------------------------------------------ identity CC
// When every table entry of both maps is "inside, fraction exactly 0, index = g + const" (two maps
// on one lattice with exactly representable spacing, or one a whole-voxel-shifted sub-box of the
// other) each trilinear sample is a0 + 0*(a1-a0) + ... = the voxel itself, provided no neighbour is
// NaN/inf and no difference overflows (0*inf would poison the sample, VolumetricData.C:273-287).
// The kernel streams both maps once with 16-byte loads and raises `bad` if it meets a value that
// could break that (non-finite or |v| >= 2^126); the host then discards the result and takes the
// interpolating path.  A -0 voxel may sample as +0 in the reference: irrelevant to the sums.
struct IdentArgs {
  const float *a, *b;           // first voxel of the common box in each map
  long a_row, a_plane, b_row, b_plane;
  int nx, ny, nz;
};

template <bool VEC>
__global__ void __launch_bounds__(kThreads) cc_identity_kernel(IdentArgs g, float thr_f, double *partials,
                                                               unsigned int *ticket, double *out, int *bad_out) {
  __shared__ double sm[6 * 32];
  double acc[6] = {0, 0, 0, 0, 0, 0};
  long long nvox = 0;
  bool bad = false;
  constexpr int W = VEC ? 4 : 1;            // voxels per lane per load
  constexpr int U = 4;                      // loads in flight per lane and map
  const unsigned int rows = (unsigned int)g.ny * (unsigned int)g.nz;
  const int lane = threadIdx.x & 31;
  const unsigned int wstride = gridDim.x * (kThreads / 32);
  // a warp owns whole rows: one 32-bit division per row, 2 x U independent 16-byte loads per trip
  for (unsigned int row = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); row < rows; row += wstride) {
    const unsigned int z = row / (unsigned int)g.ny, y = row - z * (unsigned int)g.ny;
    const float *pa = g.a + z * g.a_plane + y * g.a_row;
    const float *pb = g.b + z * g.b_plane + y * g.b_row;
    for (int x0 = lane * W; x0 < g.nx; x0 += 32 * W * U) {
      float va[U][W], vb[U][W];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int x = x0 + u * 32 * W;
        const bool in = x < g.nx;
        if (VEC) {
          float4 fa = make_float4(0.f, 0.f, 0.f, 0.f), fb = fa;
          if (in) { fa = __ldcs(reinterpret_cast<const float4 *>(pa + x)); fb = __ldcs(reinterpret_cast<const float4 *>(pb + x)); }
          va[u][0] = fa.x; va[u][W > 1 ? 1 : 0] = fa.y; va[u][W > 1 ? 2 : 0] = fa.z; va[u][W > 1 ? 3 : 0] = fa.w;
          vb[u][0] = fb.x; vb[u][W > 1 ? 1 : 0] = fb.y; vb[u][W > 1 ? 2 : 0] = fb.z; vb[u][W > 1 ? 3 : 0] = fb.w;
        } else {
          va[u][0] = in ? __ldcs(pa + x) : 0.f;
          vb[u][0] = in ? __ldcs(pb + x) : 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const bool in = x0 + u * 32 * W < g.nx;
        float fA = 0.f, fB = 0.f, fAA = 0.f, fBB = 0.f, fAB = 0.f;
        int cnt = 0;
#pragma unroll
        for (int k = 0; k < W; ++k) {
          bad = bad || !(fabsf(va[u][k]) < 8.0e37f) || !(fabsf(vb[u][k]) < 8.0e37f);
          const bool take = in && va[u][k] >= thr_f;             // MDFF.C:73-77 (NaNs raise `bad`)
          const float a = take ? va[u][k] : 0.f, b = take ? vb[u][k] : 0.f;
          fA += a; fB += b;
          fAA += a * a; fBB += b * b; fAB += a * b;              // float products, as MDFF.C:81-83
          cnt += take ? 1 : 0;
        }
        acc[0] += (double)fA; acc[1] += (double)fB; acc[2] += (double)fAA; acc[3] += (double)fBB;
        acc[4] += (double)fAB;
        nvox += cnt;
      }
    }
  }
  if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(bad_out, 1);
  acc[5] = (double)nvox;
  block_sum<6>(acc, sm);
  grid_finish<6>(acc, partials, ticket, out, sm);
}

// Host: do the six tables describe whole-voxel identity sampling?  shift[m][axis] = index offset.
static bool identity_tables(const std::vector<AxisEntry> &tab, const int off[6], const int nout[3], const DevMap &A,
                            const DevMap &B, int shift[2][3]) {
  const DevMap *maps[2] = {&A, &B};
  for (int m = 0; m < 2; ++m) {
    const int nsrc[3] = {maps[m]->nx, maps[m]->ny, maps[m]->nz};
    for (int ax = 0; ax < 3; ++ax) {
      const AxisEntry *t = &tab[off[3 * m + ax]];
      const int c0 = t[0].i0;
      for (int gi = 0; gi < nout[ax]; ++gi) {
        const AxisEntry &e = t[gi];
        if (!e.inside || e.f != 0.0f || e.i0 != c0 + gi) return false;
        if (e.i1 != e.i0 + 1 && !(e.i1 == e.i0 && e.i0 == nsrc[ax] - 1)) return false;
      }
      shift[m][ax] = c0;
    }
  }
  return true;
}

// The reference's sample of a box voxel also reads its +1 neighbours (a0 + 0*(a1-a0), VolumetricData.C:273-287):
// where the common box stops short of a map's upper face those neighbours lie OUTSIDE the box the identity
// kernels stream, and a NaN / inf / huge value there poisons the reference's sample all the same.  One-voxel-thick
// slabs just beyond the box's upper faces (edges and the corner included) are scanned for such values here.
__global__ void __launch_bounds__(kThreads) face_bad_kernel(const float *__restrict__ d, long row, long plane, int x0, int nxs,
                                                            int y0, int nys, int z0, int nzs, int *bad_out) {
  const long n = (long)nxs * nys * nzs;
  bool bad = false;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const int x = (int)(i % nxs), y = (int)((i / nxs) % nys), z = (int)(i / ((long)nxs * nys));
    const float v = d[(long)(z0 + z) * plane + (long)(y0 + y) * row + (x0 + x)];
    bad = bad || !(fabsf(v) < 8.0e37f);
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(bad_out, 1);
}

static int scan_plus_faces(const DevMap &m, const int shift[3], const int nout[3], int *d_bad) {
  Ctx &c = ctx();
  const int nsrc[3] = {m.nx, m.ny, m.nz};
  int ext[3];
  for (int ax = 0; ax < 3; ++ax) ext[ax] = shift[ax] + nout[ax] < nsrc[ax] ? 1 : 0;   // a +1 neighbour exists beyond the box
  for (int ax = 0; ax < 3; ++ax) {
    if (!ext[ax]) continue;
    int lo[3], len[3];
    for (int k = 0; k < 3; ++k) { lo[k] = shift[k]; len[k] = nout[k] + ext[k]; }
    lo[ax] = shift[ax] + nout[ax]; len[ax] = 1;
    const long n = (long)len[0] * len[1] * len[2];
    const int grid = (int)std::min<long>((n + kThreads - 1) / kThreads, (long)c.sm_count * 4);
    face_bad_kernel<<<grid, kThreads, 0, c.stream>>>(m.d, m.nx, (long)m.nx * m.ny, lo[0], len[0], lo[1], len[1], lo[2], len[2], d_bad);
    VT_LAUNCHED();
  }
  return 0;
}

constexpr int kZChunk = 32;  // output z steps per warp work item

__global__ void __launch_bounds__(kThreads, 2) cc_sep_kernel(DevMap A, DevMap B, int nx, int ny, int nz, int zchunk,
                                                          const AxisEntry *__restrict__ T, int offAx, int offAy,
                                                          int offAz, int offBx, int offBy, int offBz, float thr_f,
                                                          double *partials, unsigned int *ticket, double *out) {
  __shared__ double sm[6 * 32];
  CcAcc a;
#pragma unroll
  for (int k = 0; k < 5; ++k) a.s[k] = 0.0;
  a.n = 0;
  // work item = (32-wide x tile, kRows y rows, z chunk), dealt to warps round-robin (x tile fastest)
  const int xtiles = (nx + 31) >> 5;
  const int ygroups = (ny + kRows - 1) / kRows;
  const int zchunks = (nz + zchunk - 1) / zchunk;
  const long work = (long)xtiles * ygroups * zchunks;
  const int lane = threadIdx.x & 31;
  const long wstride = (long)gridDim.x * (kThreads / 32);
  for (long w = (long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); w < work; w += wstride) {
    const int xb = (int)(w % xtiles);
    const long r = w / xtiles;
    const int yg = (int)(r % ygroups), zc = (int)(r / ygroups);
    cc_march_item<kRows>(A.d, A.nx, A.ny, B.d, B.nx, B.ny, T + offAx, T + offAy, T + offAz, T + offBx, T + offBy, T + offBz, nx,
                  ny, xb * 32 + lane, yg * kRows, zc * zchunk, min(nz, (zc + 1) * zchunk), thr_f, a);
  }
  double acc[6] = {a.s[0], a.s[1], a.s[2], a.s[3], a.s[4], (double)a.n};
  block_sum<6>(acc, sm);
  grid_finish<6>(acc, partials, ticket, out, sm);
}

int k_cc(const vt_map *A, const vt_map *B, double thr, double sums[5], long *n) {
  VT_REQUIRE_CTX();
  Ctx &c = ctx();
  vt_grid og;
  intersection_grid(A->grid, B->grid, og);  // MDFF.C:132
  DevMap dA, dB;
  DevCoord dC;
  to_dev(A, false, dA);
  to_dev(B, false, dB);
  to_dev(og, dC);
  const long vox = (long)og.xsize * og.ysize * og.zsize;
  double *dout = c.d_scratch;
  double *partials = c.d_scratch + 8;
  if (vox <= 0 || og.xsize <= 0 || og.ysize <= 0 || og.zsize <= 0) {
    for (int k = 0; k < 5; ++k) sums[k] = 0.0;
    *n = 0;
    return 0;
  }
  // (the separable kernels address source rows with 32-bit byte offsets inside a plane)
  const bool sep = axis_aligned_out(dC) && diagonal_inverse(dA.inv) && diagonal_inverse(dB.inv) && finite_geom(dC) &&
                   (long)dA.nx * dA.ny < (1L << 30) && (long)dB.nx * dB.ny < (1L << 30) && !getenv("VOLTOOL_FORCE_GENERIC");
  if (sep) {
    SepTables T;
    int rc = upload_tables(dC, dA, dB, T);
    if (rc) return rc;
    // large maps stream through the shared-memory ring (vt_cc_staged.cu); small ones and anything
    // its preconditions reject take the direct kernel
    const char *force = getenv("VOLTOOL_CC_STAGED");
    const bool want_staged = force ? atoi(force) != 0 : vox >= (8L << 20);
    int st = 1;
    // whole-voxel identity sampling: one streaming pass, unless the data turn out to hold a value
    // that makes "sample == voxel" unsafe
    int shift[2][3];
    const int nout[3] = {og.xsize, og.ysize, og.zsize};
    // (large maps only, like the ring kernel: its sums go through float per 4 voxels)
    const char *fid = getenv("VOLTOOL_CC_IDENTITY");
    const bool want_ident = fid ? atoi(fid) != 0 : vox >= (8L << 20);
    if (want_ident && (long)og.ysize * og.zsize < (1L << 31) &&
        identity_tables(T.host, T.off, nout, dA, dB, shift)) {
      IdentArgs ia;
      ia.a_row = dA.nx; ia.a_plane = (long)dA.nx * dA.ny;
      ia.b_row = dB.nx; ia.b_plane = (long)dB.nx * dB.ny;
      ia.a = dA.d + shift[0][2] * ia.a_plane + shift[0][1] * ia.a_row + shift[0][0];
      ia.b = dB.d + shift[1][2] * ia.b_plane + shift[1][1] * ia.b_row + shift[1][0];
      ia.nx = og.xsize; ia.ny = og.ysize; ia.nz = og.zsize;
      const bool vec = (og.xsize % 4 == 0) && (dA.nx % 4 == 0) && (dB.nx % 4 == 0) &&
                       ((uintptr_t)ia.a % 16 == 0) && ((uintptr_t)ia.b % 16 == 0);
      int *d_bad = reinterpret_cast<int *>(dout + 6);
      VT_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), c.stream));
      const long wblocks = ((long)og.ysize * og.zsize + kThreads / 32 - 1) / (kThreads / 32);
      const int grid = (int)(wblocks < (long)c.sm_count * 8 ? wblocks : (long)c.sm_count * 8);
      if (int frc = scan_plus_faces(dA, shift[0], nout, d_bad)) return frc;
      if (int frc = scan_plus_faces(dB, shift[1], nout, d_bad)) return frc;
      prof_begin(PK_CC);
      if (vec) cc_identity_kernel<true><<<grid, kThreads, 0, c.stream>>>(ia, threshold_as_float(thr), partials, c.d_ticket, dout, d_bad);
      else cc_identity_kernel<false><<<grid, kThreads, 0, c.stream>>>(ia, threshold_as_float(thr), partials, c.d_ticket, dout, d_bad);
      prof_end(PK_CC);
      VT_LAUNCHED();
      VT_CUDA(cudaMemcpyAsync(c.h_pinned, dout, 7 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
      VT_CUDA(cudaStreamSynchronize(c.stream));
      int bad;
      memcpy(&bad, &c.h_pinned[6], sizeof(int));
      if (!bad) {
        dev_free(T.d);
        for (int k = 0; k < 5; ++k) sums[k] = c.h_pinned[k];
        *n = (long)c.h_pinned[5];
        return 0;
      }
    }
    if (want_staged)
      st = k_cc_staged(dA, dB, og.xsize, og.ysize, og.zsize, T.host, T.off, T.d, threshold_as_float(thr), partials,
                       c.d_ticket, dout);
    if (st > 1) return st;
    if (st == 1) {
      // z steps per warp work item: 32, shorter for small maps so that there are enough items to fill the GPU
      // (a 128^3 intersection has 512 items of 32 x 4 x 32 voxels for 2368 resident warps)
      int zchunk = kZChunk;
      const long xy_items = (long)((og.xsize + 31) / 32) * ((og.ysize + kRows - 1) / kRows);
      while (zchunk > 4 && xy_items * ((og.zsize + zchunk - 1) / zchunk) < (long)c.sm_count * 2 * (kThreads / 32)) zchunk /= 2;
      const long work = xy_items * ((og.zsize + zchunk - 1) / zchunk);
      const long wblocks = (work + kThreads / 32 - 1) / (kThreads / 32);
      int grid = (int)(wblocks < (long)c.sm_count * 8 ? wblocks : (long)c.sm_count * 8);
      prof_begin(PK_CC);
      cc_sep_kernel<<<grid, kThreads, 0, c.stream>>>(dA, dB, og.xsize, og.ysize, og.zsize, zchunk, T.d, T.offAx, T.offAy, T.offAz,
                                                     T.offBx, T.offBy, T.offBz, threshold_as_float(thr), partials, c.d_ticket, dout);
      prof_end(PK_CC);
      VT_LAUNCHED();
    }
    dev_free(T.d);
  } else {
    const long rows = (long)og.ysize * og.zsize;
    int grid = (int)(rows < (long)c.sm_count * 8 ? rows : (long)c.sm_count * 8);
    prof_begin(PK_CC);
    cc_generic_kernel<<<grid, kThreads, 0, c.stream>>>(dA, dB, dC, thr, partials, c.d_ticket, dout);
    prof_end(PK_CC);
    VT_LAUNCHED();
  }
  VT_CUDA(cudaMemcpyAsync(c.h_pinned, dout, 6 * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  VT_CUDA(cudaStreamSynchronize(c.stream));
  for (int k = 0; k < 5; ++k) sums[k] = c.h_pinned[k];
  *n = (long)c.h_pinned[5];
  return 0;
}

// ------------------------------------------------------------------------------------------ two-map ops
template <int OP, bool INTERP, bool UNION>
__global__ void __launch_bounds__(kThreads) binary_kernel(DevMap A, DevMap B, DevCoord C, float *__restrict__ out,
                                                          int zbeg, int zend) {
  const long rows = (long)C.ny * (zend - zbeg);
  for (long row0 = blockIdx.x; row0 < rows; row0 += gridDim.x) {
    const long row = row0 + (long)zbeg * C.ny;
    const int gy = (int)(row % C.ny), gz = (int)(row / C.ny);
    for (int gx = threadIdx.x; gx < C.nx; gx += blockDim.x) {
      float x, y, z;
      voxel_coord(C, gx, gy, gz, x, y, z);
      float r;
      if (OP == VT_MULTIPLY) {  // NaN-returning lookups, Voltool.C:303-350
        const float vA = INTERP ? sample_interp<false>(A, x, y, z) : sample_nearest<false>(A, x, y, z);
        const float vB = INTERP ? sample_interp<false>(B, x, y, z) : sample_nearest<false>(B, x, y, z);
        if (UNION) {
          const bool na = is_nan_bits(vA), nb = is_nan_bits(vB);
          r = (!na && !nb) ? nan_like_x86(vA * vB, vA, vB) : (!na ? vA : (!nb ? vB : 0.0f));
        } else {
          r = nan_like_x86(vA * vB, vA, vB);
        }
      } else {  // zero outside, Voltool.C:249-300, 352-377
        const float vA = INTERP ? sample_interp<true>(A, x, y, z) : sample_nearest<true>(A, x, y, z);
        const float vB = INTERP ? sample_interp<true>(B, x, y, z) : sample_nearest<true>(B, x, y, z);
        if (OP == VT_ADD) r = nan_like_x86(vA + vB, vA, vB);
        else if (OP == VT_SUBTRACT) r = nan_like_x86(vA - vB, vA, vB);
        else r = nan_like_x86((float)((double)(vA + vB) * 0.5), vA, vB);
      }
      out[row * C.nx + gx] = r;
    }
  }
}

template <int OP>
static void launch_binary(int interp, int use_union, int grid, const DevMap &A, const DevMap &B, const DevCoord &C,
                          float *out, int zbeg, int zend) {
  cudaStream_t s = ctx().stream;
  if (interp) {
    if (use_union) binary_kernel<OP, true, true><<<grid, kThreads, 0, s>>>(A, B, C, out, zbeg, zend);
    else binary_kernel<OP, true, false><<<grid, kThreads, 0, s>>>(A, B, C, out, zbeg, zend);
  } else {
    if (use_union) binary_kernel<OP, false, true><<<grid, kThreads, 0, s>>>(A, B, C, out, zbeg, zend);
    else binary_kernel<OP, false, false><<<grid, kThreads, 0, s>>>(A, B, C, out, zbeg, zend);
  }
}

int k_binary_staged(int mode, const DevMap &A, const DevMap &B, const DevCoord &C, int nx, int ny, int nz,
                    const std::vector<AxisEntry> &host_tab, const int off[6], const AxisEntry *d_tab, float *outp,
                    int *zlo, int *zhi);

static int binary_generic(int op, int interp, int use_union, const DevMap &dA, const DevMap &dB, const DevCoord &dC,
                          float *out, int zbeg, int zend) {
  const long rows = (long)dC.ny * (zend - zbeg);
  if (rows <= 0 || dC.nx <= 0) return 0;
  int grid = (int)(rows < (long)ctx().sm_count * 8 ? rows : (long)ctx().sm_count * 8);
  prof_begin(PK_BINARY);
  switch (op) {
    case VT_ADD: launch_binary<VT_ADD>(interp, use_union, grid, dA, dB, dC, out, zbeg, zend); break;
    case VT_SUBTRACT: launch_binary<VT_SUBTRACT>(interp, use_union, grid, dA, dB, dC, out, zbeg, zend); break;
    case VT_MULTIPLY: launch_binary<VT_MULTIPLY>(interp, use_union, grid, dA, dB, dC, out, zbeg, zend); break;
    case VT_AVERAGE: launch_binary<VT_AVERAGE>(interp, use_union, grid, dA, dB, dC, out, zbeg, zend); break;
    default: set_error("vt_binary: unknown op %d", op); return 1;
  }
  prof_end(PK_BINARY);
  VT_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------ identity two-map ops
// Same precondition as cc_identity_kernel: every sample is the voxel itself, so add/sub/mult/avg on
// one exact lattice stream at 12 B/voxel.  Two refinements keep the bits of the reference:
//  * a -0 voxel samples as -0 + 0*(a1-a0) = +0 or -0 depending on its neighbours, which shows in
//    the sign of a zero result: such voxels are redone with the literal per-voxel arithmetic;
//  * any non-finite or huge value raises `bad` and the host redoes the whole op.
// MODE: 1 add, 2 subtract, 3/4 multiply, 5 average (Voltool.C:249-377).
template <int MODE>
__device__ __forceinline__ float identity_op(float a, float b) {
  if (MODE == 1) return a + b;
  if (MODE == 2) return a - b;
  if (MODE == 3 || MODE == 4) return a * b;
  return (float)((double)(a + b) * 0.5);
}

template <int MODE, bool VEC>
__global__ void __launch_bounds__(kThreads) binary_identity_kernel(IdentArgs g, float *__restrict__ outp, DevMap gA,
                                                                   DevMap gB, DevCoord gC, int *bad_out) {
  constexpr int W = VEC ? 4 : 1;
  constexpr int U = 2;
  constexpr bool SAFE = (MODE != 3 && MODE != 4);
  bool bad = false;
  const unsigned int rows = (unsigned int)g.ny * (unsigned int)g.nz;
  const int lane = threadIdx.x & 31;
  const unsigned int wstride = gridDim.x * (kThreads / 32);
  for (unsigned int row = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); row < rows; row += wstride) {
    const unsigned int z = row / (unsigned int)g.ny, y = row - z * (unsigned int)g.ny;
    const float *pa = g.a + z * g.a_plane + y * g.a_row;
    const float *pb = g.b + z * g.b_plane + y * g.b_row;
    float *po = outp + (size_t)row * g.nx;
    for (int x0 = lane * W; x0 < g.nx; x0 += 32 * W * U) {
      float va[U][W], vb[U][W];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int x = x0 + u * 32 * W;
        const bool in = x < g.nx;
        if (VEC) {
          float4 fa = make_float4(0.f, 0.f, 0.f, 0.f), fb = fa;
          if (in) { fa = __ldcs(reinterpret_cast<const float4 *>(pa + x)); fb = __ldcs(reinterpret_cast<const float4 *>(pb + x)); }
          va[u][0] = fa.x; va[u][W > 1 ? 1 : 0] = fa.y; va[u][W > 1 ? 2 : 0] = fa.z; va[u][W > 1 ? 3 : 0] = fa.w;
          vb[u][0] = fb.x; vb[u][W > 1 ? 1 : 0] = fb.y; vb[u][W > 1 ? 2 : 0] = fb.z; vb[u][W > 1 ? 3 : 0] = fb.w;
        } else {
          va[u][0] = in ? __ldcs(pa + x) : 0.f;
          vb[u][0] = in ? __ldcs(pb + x) : 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int x = x0 + u * 32 * W;
        if (x >= g.nx) continue;
        float r[W];
#pragma unroll
        for (int k = 0; k < W; ++k) {
          float a = va[u][k], b = vb[u][k];
          bad = bad || !(fabsf(a) < 8.0e37f) || !(fabsf(b) < 8.0e37f);
          if (__float_as_uint(a) == 0x80000000u || __float_as_uint(b) == 0x80000000u) {
            float cx, cy, cz;
            voxel_coord(gC, x + k, (int)y, (int)z, cx, cy, cz);
            a = sample_interp<SAFE>(gA, cx, cy, cz);
            b = sample_interp<SAFE>(gB, cx, cy, cz);
          }
          r[k] = identity_op<MODE>(a, b);
        }
        if (VEC) __stcs(reinterpret_cast<float4 *>(po + x), make_float4(r[0], r[W > 1 ? 1 : 0], r[W > 1 ? 2 : 0], r[W > 1 ? 3 : 0]));
        else __stcs(po + x, r[0]);
      }
    }
  }
  if (__any_sync(0xffffffffu, bad) && lane == 0) atomicOr(bad_out, 1);
}

template <int MODE>
static int launch_binary_identity(const IdentArgs &ia, bool vec, float *out, const DevMap &dA, const DevMap &dB,
                                  const DevCoord &dC, int *d_bad) {
  Ctx &c = ctx();
  const long wblocks = ((long)ia.ny * ia.nz + kThreads / 32 - 1) / (kThreads / 32);
  const int grid = (int)(wblocks < (long)c.sm_count * 8 ? wblocks : (long)c.sm_count * 8);
  prof_begin(PK_BINARY);
  if (vec) binary_identity_kernel<MODE, true><<<grid, kThreads, 0, c.stream>>>(ia, out, dA, dB, dC, d_bad);
  else binary_identity_kernel<MODE, false><<<grid, kThreads, 0, c.stream>>>(ia, out, dA, dB, dC, d_bad);
  prof_end(PK_BINARY);
  VT_LAUNCHED();
  return 0;
}

int k_binary(int op, int interp, int use_union, const vt_map *A, const vt_map *B, const vt_grid &og, float *out) {
  VT_REQUIRE_CTX();
  DevMap dA, dB;
  DevCoord dC;
  to_dev(A, !interp, dA);
  to_dev(B, !interp, dB);
  to_dev(og, dC);
  if ((long)og.ysize * og.zsize <= 0 || og.xsize <= 0) return 0;
  const long vox = (long)og.xsize * og.ysize * og.zsize;
  // large interpolated ops on orthogonal cells stream through the shared-memory ring; voxels whose
  // result involves a NaN are redone there with the literal arithmetic, so the bits are the same
  const char *force = getenv("VOLTOOL_CC_STAGED");
  const bool want_staged = interp && (force ? atoi(force) != 0 : vox >= (8L << 20)) && axis_aligned_out(dC) &&
                           diagonal_inverse(dA.inv) && diagonal_inverse(dB.inv) && finite_geom(dC) &&
                           !getenv("VOLTOOL_FORCE_GENERIC");
  if (want_staged) {
    SepTables T;
    int rc = upload_tables(dC, dA, dB, T);
    if (rc) return rc;
    const int mode = op == VT_ADD ? 1 : op == VT_SUBTRACT ? 2 : op == VT_MULTIPLY ? (use_union ? 4 : 3) : 5;
    // both maps on the output lattice with whole-voxel offsets: stream (see binary_identity_kernel)
    int shift[2][3];
    const int nout[3] = {og.xsize, og.ysize, og.zsize};
    const char *fid = getenv("VOLTOOL_CC_IDENTITY");
    if ((!fid || atoi(fid) != 0) && (long)og.ysize * og.zsize < (1L << 31) &&
        identity_tables(T.host, T.off, nout, dA, dB, shift)) {
      Ctx &c = ctx();
      IdentArgs ia;
      ia.a_row = dA.nx; ia.a_plane = (long)dA.nx * dA.ny;
      ia.b_row = dB.nx; ia.b_plane = (long)dB.nx * dB.ny;
      ia.a = dA.d + shift[0][2] * ia.a_plane + shift[0][1] * ia.a_row + shift[0][0];
      ia.b = dB.d + shift[1][2] * ia.b_plane + shift[1][1] * ia.b_row + shift[1][0];
      ia.nx = og.xsize; ia.ny = og.ysize; ia.nz = og.zsize;
      const bool vec = (og.xsize % 4 == 0) && (dA.nx % 4 == 0) && (dB.nx % 4 == 0) && ((uintptr_t)ia.a % 16 == 0) &&
                       ((uintptr_t)ia.b % 16 == 0) && ((uintptr_t)out % 16 == 0);
      int *d_bad = reinterpret_cast<int *>(c.d_scratch);
      VT_CUDA(cudaMemsetAsync(d_bad, 0, sizeof(int), c.stream));
      if (int frc = scan_plus_faces(dA, shift[0], nout, d_bad)) return frc;
      if (int frc = scan_plus_faces(dB, shift[1], nout, d_bad)) return frc;
      switch (mode) {
        case 1: launch_binary_identity<1>(ia, vec, out, dA, dB, dC, d_bad); break;
        case 2: launch_binary_identity<2>(ia, vec, out, dA, dB, dC, d_bad); break;
        case 3: launch_binary_identity<3>(ia, vec, out, dA, dB, dC, d_bad); break;
        case 4: launch_binary_identity<4>(ia, vec, out, dA, dB, dC, d_bad); break;
        default: launch_binary_identity<5>(ia, vec, out, dA, dB, dC, d_bad); break;
      }
      int bad = 0;
      VT_CUDA(cudaMemcpyAsync(c.h_pinned, d_bad, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
      VT_CUDA(cudaStreamSynchronize(c.stream));
      memcpy(&bad, c.h_pinned, sizeof(int));
      if (!bad) { dev_free(T.d); return 0; }
    }
    int zlo = 0, zhi = 0;
    const int st = k_binary_staged(mode, dA, dB, dC, og.xsize, og.ysize, og.zsize, T.host, T.off, T.d, out, &zlo, &zhi);
    dev_free(T.d);
    if (st > 1) return st;
    if (st == 0) {
      rc = binary_generic(op, interp, use_union, dA, dB, dC, out, 0, zlo);
      if (!rc) rc = binary_generic(op, interp, use_union, dA, dB, dC, out, zhi, og.zsize);
      return rc;
    }
  }
  return binary_generic(op, interp, use_union, dA, dB, dC, out, 0, og.zsize);
}

}  // namespace vt
