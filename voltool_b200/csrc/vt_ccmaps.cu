// vt_ccmaps.cu -- `voltool cc -calcdiffmap / -calcspatialccmap` (usage TclMDFF.C:180-184, plumbing :555-594).
//
// The reference fills these two maps inside vmd_cuda_compare_sel_refmap (CUDAMDFF.cu), which is not in the
// snapshot; what they mean is stated only in the usage text.  The definition built here is written down next to
// its CPU twin (oracle/voltool_oracle.c, vo_cc_maps):
//   G          = init_from_intersection(A, B): the grid cc_threaded walks (MDFF.C:132)
//   diff[i]    = a_i - b_i, the two interpolated samples of MDFF.C:70-71 at voxel i of G
//   spatial[t] = the CC expression of MDFF.C:143-149 over the 8 x 8 x 8 tile t of G (voxels that pass MDFF.C:73-77)
//   cc         = the same expression over the tile sums added in tile order
// One CTA per tile: 256 threads sample the tile's 512 voxels (the literal per-voxel arithmetic, any cell shape),
// store the difference, and leave the samples in shared memory; six threads, one per warp, then add the tile's
// terms strictly in voxel order, so the tile sums -- and with them every spatial CC -- equal the sequential CPU
// loop bit for bit.  The per-tile CC and the global one are finished on the host in double from those sums.
#include <cmath>
#include <cstring>
#include <vector>

#include "vt_device.cuh"

namespace vt {

namespace {

void to_dev_map(const vt_map *m, DevMap &d) {
  MapGeom g;
  map_geom(m->grid, g);
  d.d = m->d;
  for (int i = 0; i < 16; ++i) d.inv[i] = g.inv0[i];
  d.nx = g.nx; d.ny = g.ny; d.nz = g.nz;
}
void to_dev_coord(const vt_grid &grid, DevCoord &c) {
  CoordGeom g;
  coord_geom(grid, g);
  c.ox = g.origin[0]; c.oy = g.origin[1]; c.oz = g.origin[2];
  for (int k = 0; k < 3; ++k) { c.xd[k] = g.xd[k]; c.yd[k] = g.yd[k]; c.zd[k] = g.zd[k]; }
  c.nx = g.nx; c.ny = g.ny; c.nz = g.nz;
}

}  // namespace

__global__ void __launch_bounds__(256) cc_tiles_kernel(DevMap A, DevMap B, DevCoord C, double thr, int tnx, int tny,
                                                       float *__restrict__ diff, double *__restrict__ tile_sums) {
  __shared__ float sa[512], sb[512];
  __shared__ unsigned char ok[512];
  const long tile = blockIdx.x;
  const int tx = (int)(tile % tnx), ty = (int)((tile / tnx) % tny), tz = (int)(tile / ((long)tnx * tny));
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int j = threadIdx.x + 256 * h;
    const int gx = tx * 8 + (j & 7), gy = ty * 8 + ((j >> 3) & 7), gz = tz * 8 + (j >> 6);
    bool pass = false;
    float vA = 0.f, vB = 0.f;
    if (gx < C.nx && gy < C.ny && gz < C.nz) {
      float x, y, z;
      voxel_coord(C, gx, gy, gz, x, y, z);
      vA = sample_interp<false>(A, x, y, z);
      vB = sample_interp<false>(B, x, y, z);
      if (diff) diff[((long)gz * C.ny + gy) * C.nx + gx] = nan_like_x86(vA - vB, vA, vB);
      pass = !is_nan_bits(vB) && !is_nan_bits(vA) && (double)vA >= thr;
    }
    sa[j] = vA; sb[j] = vB; ok[j] = pass ? 1 : 0;
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0 && threadIdx.x < 6 * 32) {   // term k of every voxel, in voxel order
    const int k = threadIdx.x >> 5;
    double s = 0.0;
    for (int j = 0; j < 512; ++j) {
      if (!ok[j]) continue;
      const float a = sa[j], b = sb[j];
      const float t = k == 0 ? a : k == 1 ? b : k == 2 ? a * a : k == 3 ? b * b : k == 4 ? a * b : 1.0f;
      s += (double)t;
    }
    tile_sums[tile * 6 + k] = s;
  }
}

}  // namespace vt

using namespace vt;

extern "C" {

int vt_cc_maps(const vt_map *A, const vt_map *B, double threshold, double *cc, vt_map **diff_out, vt_map **spatial_out) {
  VT_REQUIRE_CTX();
  if (!A || !B || !cc) { set_error("vt_cc_maps: bad arguments"); return 1; }
  Ctx &c = ctx();
  vt_grid og;
  intersection_grid(A->grid, B->grid, og);
  if (og.xsize <= 0 || og.ysize <= 0 || og.zsize <= 0) { set_error("vt_cc_maps: the maps do not overlap"); return 1; }
  // spatial grid: ceil(n/8) tiles per axis on G's origin, one cell = 8 cells of G
  vt_grid sg = og;
  {
    float xd[3], yd[3], zd[3];
    cell_axes(og, xd, yd, zd);
    sg.xsize = (og.xsize + 7) / 8; sg.ysize = (og.ysize + 7) / 8; sg.zsize = (og.zsize + 7) / 8;
    for (int k = 0; k < 3; ++k) {
      sg.xaxis[k] = 8.0 * (double)xd[k] * (double)(sg.xsize > 1 ? sg.xsize - 1 : 1);
      sg.yaxis[k] = 8.0 * (double)yd[k] * (double)(sg.ysize > 1 ? sg.ysize - 1 : 1);
      sg.zaxis[k] = 8.0 * (double)zd[k] * (double)(sg.zsize > 1 ? sg.zsize - 1 : 1);
    }
  }
  const long ntiles = (long)sg.xsize * sg.ysize * sg.zsize;
  if (ntiles > 0x7fffffffL) { set_error("vt_cc_maps: too many tiles"); return 1; }
  DevMap dA, dB;
  DevCoord dC;
  to_dev_map(A, dA);
  to_dev_map(B, dB);
  to_dev_coord(og, dC);
  vt_map *dm = nullptr;
  if (diff_out) {
    dm = new vt_map();
    dm->grid = og;
    dm->n = gridsize(og);
    memset(&dm->cache, 0, sizeof(dm->cache));
    if (dev_alloc((void **)&dm->d, sizeof(float) * (size_t)dm->n)) { delete dm; return 2; }
  }
  double *d_sums = nullptr;
  if (dev_alloc((void **)&d_sums, sizeof(double) * 6 * (size_t)ntiles)) { vt_map_destroy(dm); return 2; }
  prof_begin(PK_CC);
  cc_tiles_kernel<<<(unsigned int)ntiles, 256, 0, c.stream>>>(dA, dB, dC, threshold, sg.xsize, sg.ysize, dm ? dm->d : nullptr, d_sums);
  prof_end(PK_CC);
  c.launches++;
  std::vector<double> sums(6 * (size_t)ntiles);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(sums.data(), d_sums, sizeof(double) * sums.size(), cudaMemcpyDeviceToHost, c.stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c.stream);
  dev_free(d_sums);
  if (e != cudaSuccess) { vt_map_destroy(dm); return cuda_fail(e, "cc_tiles_kernel", __FILE__, __LINE__); }
  std::vector<float> sp((size_t)ntiles);
  double tot[5] = {0, 0, 0, 0, 0};
  long ntot = 0;
  for (long t = 0; t < ntiles; ++t) {
    const double *s = &sums[6 * (size_t)t];
    const long cnt = (long)s[5];
    sp[(size_t)t] = cnt > 0 ? (float)cc_from_sums(s, cnt) : 0.0f;
    for (int k = 0; k < 5; ++k) tot[k] += s[k];
    ntot += cnt;
  }
  *cc = cc_from_sums(tot, ntot);
  int rc = 0;
  if (dm) {   // new VolumetricData: the constructor's compute_minmax (VolumetricData.C:66-89)
    if (dm->n > 0) rc = k_minmax(dm->d, dm->n, &dm->cache.cmin, &dm->cache.cmax);
    dm->cache.minmax_valid = 1;
    if (rc) { vt_map_destroy(dm); return rc; }
    *diff_out = dm;
  }
  if (spatial_out) {
    rc = vt_map_create(&sg, sp.data(), spatial_out);
    if (rc) { if (dm) { vt_map_destroy(dm); *diff_out = nullptr; } return rc; }
  }
  return 0;
}

// mdff_cc with -calcsynthmap / -calcdiffmap / -calcspatialccmap (TclMDFF.C:440-594): the policy and the sim of
// vt_sim_cc, then the maps above.  Any of the three outputs may be NULL.
int vt_sim_cc_maps(const float *pos, const float *radii, long natoms, double resolution, double gspacing_in,
                   const vt_map *target, double threshold, double *cc, vt_map **synth_out, vt_map **diff_out,
                   vt_map **spatial_out) {
  VT_REQUIRE_CTX();
  float radscale;
  double gspacing;
  int quality;
  sim_policy(resolution, gspacing_in, radscale, gspacing, quality);
  vt_map *sim = nullptr;
  int rc = vt_sim(pos, radii, natoms, quality, radscale, (float)gspacing, &sim);
  if (rc) return rc;
  rc = vt_cc_maps(sim, target, threshold, cc, diff_out, spatial_out);
  if (synth_out && !rc) *synth_out = sim;
  else vt_map_destroy(sim);
  return rc;
}

}  // extern "C"
