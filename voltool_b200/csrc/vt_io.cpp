// vt_io.cpp -- map files at the boundary of the path: OpenDX scalar fields and MRC/CCP4.
//
// The reference saves maps with volmap_write_dx_file (call sites TclMDFF.C:146,159,538,560,585;
// `voltool ... -o` goes through the molfile plugins, TclVoltool.C:321-358).  Neither writer is in the
// snapshot, so the format is restated from the OpenDX layout VMD reads and writes ("parity
// unpinned"): gridpositions counts / origin / three delta vectors (= axis / (n-1)) / gridconnections /
// one array of nx*ny*nz values with **z fastest, x slowest**, three per line.  Maps here are x
// fastest (VolumetricData.h:35-36), so both directions transpose.  Values are written with "%.9g"
// (nine significant digits restore every float exactly; VMD prints "%g").  Host only: no CUDA.
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/voltool_gpu.h"

namespace vt {
extern void set_error(const char *fmt, ...);
}
using vt::set_error;

static void delta_of(const double axis[3], int n, double out[3]) {
  // cell_axes (VolumetricData.C:138-168): spacing = axis / (n - 1); a single sample keeps the axis
  const double d = n > 1 ? (double)(n - 1) : 1.0;
  for (int k = 0; k < 3; ++k) out[k] = axis[k] / d;
}

extern "C" {

int vt_dx_write(const char *path, const vt_grid *g, const float *data, const char *name) {
  if (!path || !g || (!data && (long)g->xsize * g->ysize * g->zsize > 0)) { set_error("vt_dx_write: null argument"); return 1; }
  if (g->xsize < 0 || g->ysize < 0 || g->zsize < 0) { set_error("vt_dx_write: negative size"); return 1; }
  FILE *f = fopen(path, "w");
  if (!f) { set_error("vt_dx_write: cannot open '%s' for writing", path); return 1; }
  const int nx = g->xsize, ny = g->ysize, nz = g->zsize;
  const long n = (long)nx * ny * nz;
  double dx[3], dy[3], dz[3];
  delta_of(g->xaxis, nx, dx);
  delta_of(g->yaxis, ny, dy);
  delta_of(g->zaxis, nz, dz);
  fprintf(f, "# Data calculated by voltool_b200 (OpenDX scalar field, VMD volmap layout)\n");
  fprintf(f, "object 1 class gridpositions counts %d %d %d\n", nx, ny, nz);
  fprintf(f, "origin %.17g %.17g %.17g\n", g->origin[0], g->origin[1], g->origin[2]);
  fprintf(f, "delta %.17g %.17g %.17g\n", dx[0], dx[1], dx[2]);
  fprintf(f, "delta %.17g %.17g %.17g\n", dy[0], dy[1], dy[2]);
  fprintf(f, "delta %.17g %.17g %.17g\n", dz[0], dz[1], dz[2]);
  fprintf(f, "object 2 class gridconnections counts %d %d %d\n", nx, ny, nz);
  fprintf(f, "object 3 class array type double rank 0 items %ld data follows\n", n);
  long col = 0;
  for (int x = 0; x < nx; ++x)
    for (int y = 0; y < ny; ++y)
      for (int z = 0; z < nz; ++z) {
        fprintf(f, "%.9g", (double)data[((long)z * ny + y) * nx + x]);
        fputc(++col % 3 == 0 ? '\n' : ' ', f);
      }
  if (col % 3 != 0) fputc('\n', f);
  fprintf(f, "\nobject \"%s\" class field\n", (name && *name) ? name : "density");
  const bool bad = ferror(f) != 0;
  if (fclose(f) != 0 || bad) { set_error("vt_dx_write: write to '%s' failed", path); return 1; }
  return 0;
}

int vt_dx_read(const char *path, vt_grid *g, float **data_out) {
  if (!path || !g || !data_out) { set_error("vt_dx_read: null argument"); return 1; }
  *data_out = nullptr;
  FILE *f = fopen(path, "r");
  if (!f) { set_error("vt_dx_read: cannot open '%s'", path); return 1; }
  int nx = -1, ny = -1, nz = -1, ndelta = 0;
  double origin[3] = {0, 0, 0}, delta[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  long items = -1;
  bool have_origin = false;
  char line[4096];
  while (fgets(line, sizeof line, f)) {
    const char *p = line;
    while (isspace((unsigned char)*p)) ++p;
    if (*p == '#' || *p == 0) continue;
    if (!strncmp(p, "object", 6) && strstr(p, "gridpositions")) {
      const char *c = strstr(p, "counts");
      if (!c || sscanf(c + 6, "%d %d %d", &nx, &ny, &nz) != 3) { fclose(f); set_error("vt_dx_read: bad gridpositions line in '%s'", path); return 1; }
    } else if (!strncmp(p, "origin", 6)) {
      if (sscanf(p + 6, "%lf %lf %lf", &origin[0], &origin[1], &origin[2]) != 3) { fclose(f); set_error("vt_dx_read: bad origin line in '%s'", path); return 1; }
      have_origin = true;
    } else if (!strncmp(p, "delta", 5)) {
      if (ndelta >= 3 || sscanf(p + 5, "%lf %lf %lf", &delta[ndelta][0], &delta[ndelta][1], &delta[ndelta][2]) != 3) { fclose(f); set_error("vt_dx_read: bad delta line in '%s'", path); return 1; }
      ++ndelta;
    } else if (!strncmp(p, "object", 6) && strstr(p, "class array")) {
      const char *c = strstr(p, "items");
      if (!c || sscanf(c + 5, "%ld", &items) != 1) { fclose(f); set_error("vt_dx_read: bad array line in '%s'", path); return 1; }
      if (!strstr(p, "data follows")) { fclose(f); set_error("vt_dx_read: only inline text data is supported ('%s')", path); return 1; }
      break;
    }
  }
  if (nx < 0 || ny < 0 || nz < 0 || !have_origin || ndelta != 3 || items < 0) { fclose(f); set_error("vt_dx_read: incomplete header in '%s'", path); return 1; }
  const long n = (long)nx * ny * nz;
  if (items != n) { fclose(f); set_error("vt_dx_read: %ld items for a %d x %d x %d grid in '%s'", items, nx, ny, nz, path); return 1; }
  float *data = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
  if (!data) { fclose(f); set_error("vt_dx_read: out of memory (%ld voxels)", n); return 1; }
  long got = 0;
  for (int x = 0; x < nx && got >= 0; ++x)
    for (int y = 0; y < ny && got >= 0; ++y)
      for (int z = 0; z < nz; ++z) {
        double v;
        char tok[64];
        if (fscanf(f, "%63s", tok) != 1) { got = -1; break; }
        char *end = nullptr;
        v = strtod(tok, &end);   // accepts nan / inf spellings as well
        if (end == tok) { got = -1; break; }
        data[((long)z * ny + y) * nx + x] = (float)v;
        ++got;
      }
  fclose(f);
  if (got != n) { free(data); set_error("vt_dx_read: '%s' ends after fewer than %ld values", path, n); return 1; }
  for (int k = 0; k < 3; ++k) {
    g->origin[k] = origin[k];
    g->xaxis[k] = delta[0][k] * (nx > 1 ? nx - 1 : 1);
    g->yaxis[k] = delta[1][k] * (ny > 1 ? ny - 1 : 1);
    g->zaxis[k] = delta[2][k] * (nz > 1 ? nz - 1 : 1);
  }
  g->xsize = nx; g->ysize = ny; g->zsize = nz;
  *data_out = data;
  return 0;
}

// ------------------------------------------------------------------------------------------ MRC / CCP4
// MRC2014 / CCP4 maps (the container real EM maps come in).  VMD reads them through its
// ccp4plugin, which is not in the snapshot, so the mapping onto VolumetricData's geometry is restated:
// voxel = CELLA / M{X,Y,Z}; origin = ORIGIN (words 50-52) + N{X,Y,Z}START * voxel; axis = voxel * (n - 1)
// (cell_axes, VolumetricData.C:138-168); MAPC/MAPR/MAPS give the axis of columns / rows / sections and
// the data are transposed to x fastest.  Orthogonal cells only (the reference's two-map code assumes
// them, Voltool.C:55-56).  Modes 0 (int8), 1 (int16), 2 (float32) and 6 (uint16); either byte order.
// The writer emits mode 2, MAPC/R/S = 1/2/3, NSTART = 0, the origin in words 50-52.
namespace {

inline unsigned int bswap32(unsigned int v) { return (v >> 24) | ((v >> 8) & 0xff00u) | ((v << 8) & 0xff0000u) | (v << 24); }
inline unsigned short bswap16(unsigned short v) { return (unsigned short)((v >> 8) | (v << 8)); }

struct MrcHeader {
  int w[256];
  int i(int word) const { return w[word - 1]; }             // 1-based word numbers of the format description
  float f(int word) const { float v; memcpy(&v, &w[word - 1], 4); return v; }
  void set_f(int word, float v) { memcpy(&w[word - 1], &v, 4); }
};

}  // namespace

int vt_mrc_write(const char *path, const vt_grid *g, const float *data) {
  if (!path || !g || !data) { set_error("vt_mrc_write: null argument"); return 1; }
  const int nx = g->xsize, ny = g->ysize, nz = g->zsize;
  if (nx < 1 || ny < 1 || nz < 1) { set_error("vt_mrc_write: empty map"); return 1; }
  if (g->xaxis[1] != 0 || g->xaxis[2] != 0 || g->yaxis[0] != 0 || g->yaxis[2] != 0 || g->zaxis[0] != 0 || g->zaxis[1] != 0) {
    set_error("vt_mrc_write: only axis-aligned orthogonal cells can be written");
    return 1;
  }
  double dx[3], dy[3], dz[3];
  delta_of(g->xaxis, nx, dx);
  delta_of(g->yaxis, ny, dy);
  delta_of(g->zaxis, nz, dz);
  const long n = (long)nx * ny * nz;
  double mn = data[0], mx = data[0], sum = 0.0, sq = 0.0;
  for (long k = 0; k < n; ++k) {
    const double v = data[k];
    if (v < mn) mn = v;
    if (v > mx) mx = v;
    sum += v;
  }
  const double mean = sum / (double)n;
  for (long k = 0; k < n; ++k) { const double d = data[k] - mean; sq += d * d; }
  MrcHeader h;
  memset(&h, 0, sizeof h);
  h.w[0] = nx; h.w[1] = ny; h.w[2] = nz; h.w[3] = 2;
  h.w[7] = nx; h.w[8] = ny; h.w[9] = nz;
  h.set_f(11, (float)(dx[0] * nx)); h.set_f(12, (float)(dy[1] * ny)); h.set_f(13, (float)(dz[2] * nz));
  h.set_f(14, 90.f); h.set_f(15, 90.f); h.set_f(16, 90.f);
  h.w[16] = 1; h.w[17] = 2; h.w[18] = 3;
  h.set_f(20, (float)mn); h.set_f(21, (float)mx); h.set_f(22, (float)mean);
  h.w[22] = 1;                                    // ISPG: P1
  h.w[27] = 20140;                                // NVERSION
  h.set_f(50, (float)g->origin[0]); h.set_f(51, (float)g->origin[1]); h.set_f(52, (float)g->origin[2]);
  memcpy(&h.w[52], "MAP ", 4);
  const unsigned char stamp[4] = {0x44, 0x44, 0x00, 0x00};   // little-endian reals and integers
  memcpy(&h.w[53], stamp, 4);
  h.set_f(55, (float)std::sqrt(sq / (double)n));
  h.w[55] = 1;
  char *label = reinterpret_cast<char *>(&h.w[56]);
  memset(label, ' ', 80);
  memcpy(label, "voltool_b200", 12);
  FILE *f = fopen(path, "wb");
  if (!f) { set_error("vt_mrc_write: cannot open '%s' for writing", path); return 1; }
  bool ok = fwrite(&h, 1, 1024, f) == 1024;
  ok = ok && fwrite(data, sizeof(float), (size_t)n, f) == (size_t)n;
  if (fclose(f) != 0 || !ok) { set_error("vt_mrc_write: write to '%s' failed", path); return 1; }
  return 0;
}

int vt_mrc_read(const char *path, vt_grid *g, float **data_out) {
  if (!path || !g || !data_out) { set_error("vt_mrc_read: null argument"); return 1; }
  *data_out = nullptr;
  FILE *f = fopen(path, "rb");
  if (!f) { set_error("vt_mrc_read: cannot open '%s'", path); return 1; }
  MrcHeader h;
  if (fread(&h, 1, 1024, f) != 1024) { fclose(f); set_error("vt_mrc_read: '%s' is shorter than an MRC header", path); return 1; }
  // byte order: the mode and the axis words are small positive numbers in the file's own order
  auto sane = [](const MrcHeader &q) {
    return q.i(4) >= 0 && q.i(4) <= 12 && q.i(1) > 0 && q.i(1) < (1 << 20) && q.i(2) > 0 && q.i(2) < (1 << 20) && q.i(3) > 0 &&
           q.i(3) < (1 << 20);
  };
  bool swapped = false;
  if (!sane(h)) {
    for (int k = 0; k < 56; ++k) h.w[k] = (int)bswap32((unsigned int)h.w[k]);
    swapped = true;
    if (!sane(h)) { fclose(f); set_error("vt_mrc_read: '%s' does not look like an MRC/CCP4 map", path); return 1; }
  }
  const int nc = h.i(1), nr = h.i(2), ns = h.i(3), mode = h.i(4);
  int mapc = h.i(17), mapr = h.i(18), maps = h.i(19);
  if (mapc == 0 && mapr == 0 && maps == 0) { mapc = 1; mapr = 2; maps = 3; }   // very old files leave them empty
  if (mapc < 1 || mapc > 3 || mapr < 1 || mapr > 3 || maps < 1 || maps > 3 || mapc == mapr || mapc == maps || mapr == maps) {
    fclose(f); set_error("vt_mrc_read: bad axis order %d %d %d in '%s'", mapc, mapr, maps, path); return 1;
  }
  const float al = h.f(14), be = h.f(15), ga = h.f(16);
  auto right = [](float a) { return a == 0.f || std::fabs(a - 90.f) < 1e-3f; };
  if (!right(al) || !right(be) || !right(ga)) { fclose(f); set_error("vt_mrc_read: non-orthogonal cell in '%s'", path); return 1; }
  size_t esize;
  switch (mode) {
    case 0: esize = 1; break;
    case 1: case 6: esize = 2; break;
    case 2: esize = 4; break;
    default: fclose(f); set_error("vt_mrc_read: unsupported mode %d in '%s'", mode, path); return 1;
  }
  const int nsymbt = h.i(24);
  if (nsymbt < 0 || fseek(f, 1024L + nsymbt, SEEK_SET) != 0) { fclose(f); set_error("vt_mrc_read: bad extended header in '%s'", path); return 1; }
  // file axes (columns, rows, sections) -> x, y, z
  int dim_of[4] = {0, 0, 0, 0};        // size along X, Y, Z (index 1..3)
  int start_of[4] = {0, 0, 0, 0};
  dim_of[mapc] = nc; dim_of[mapr] = nr; dim_of[maps] = ns;
  start_of[mapc] = h.i(5); start_of[mapr] = h.i(6); start_of[maps] = h.i(7);
  const int m[4] = {0, h.i(8), h.i(9), h.i(10)};
  const float cell[4] = {0.f, h.f(11), h.f(12), h.f(13)};
  double vox[4];
  for (int a = 1; a <= 3; ++a) vox[a] = (m[a] > 0 && cell[a] > 0.f) ? (double)cell[a] / m[a] : 1.0;
  const int nx = dim_of[1], ny = dim_of[2], nz = dim_of[3];
  const long n = (long)nx * ny * nz;
  float *data = (float *)malloc(sizeof(float) * (size_t)n);
  std::vector<unsigned char> row((size_t)nc * esize);
  if (!data) { fclose(f); set_error("vt_mrc_read: out of memory (%ld voxels)", n); return 1; }
  long stride_of[4] = {0, 1, nx, (long)nx * ny};   // memory stride along X, Y, Z
  const long sc = stride_of[mapc], sr = stride_of[mapr], ss = stride_of[maps];
  for (int s = 0; s < ns; ++s)
    for (int r = 0; r < nr; ++r) {
      if (fread(row.data(), esize, (size_t)nc, f) != (size_t)nc) {
        free(data); fclose(f); set_error("vt_mrc_read: '%s' ends before section %d row %d", path, s, r); return 1;
      }
      float *dst = data + s * ss + r * sr;
      for (int c = 0; c < nc; ++c) {
        float v;
        if (mode == 2) {
          unsigned int u; memcpy(&u, &row[4 * (size_t)c], 4);
          if (swapped) u = bswap32(u);
          memcpy(&v, &u, 4);
        } else if (mode == 0) {
          v = (float)(signed char)row[c];
        } else {
          unsigned short u; memcpy(&u, &row[2 * (size_t)c], 2);
          if (swapped) u = bswap16(u);
          v = mode == 1 ? (float)(short)u : (float)u;
        }
        dst[c * sc] = v;
      }
    }
  fclose(f);
  memset(g, 0, sizeof *g);
  const float org[4] = {0.f, h.f(50), h.f(51), h.f(52)};
  const int nxyz[4] = {0, nx, ny, nz};
  double *axis[4] = {nullptr, g->xaxis, g->yaxis, g->zaxis};
  // the origin is the ORIGIN words (MRC2014) when any of them is set, else the N*START offsets of the older CCP4
  // convention -- never both (files that carry both describe the same shift twice)
  const bool has_origin = org[1] != 0.f || org[2] != 0.f || org[3] != 0.f;
  for (int a = 1; a <= 3; ++a) {
    g->origin[a - 1] = has_origin ? (double)org[a] : (double)start_of[a] * vox[a];
    axis[a][a - 1] = vox[a] * (nxyz[a] > 1 ? nxyz[a] - 1 : 1);
  }
  g->xsize = nx; g->ysize = ny; g->zsize = nz;
  *data_out = data;
  return 0;
}

int vt_map_write_mrc(const vt_map *m, const char *path) {
  if (!m) { set_error("vt_map_write_mrc: null map"); return 1; }
  vt_grid g;
  int rc = vt_map_grid(m, &g);
  if (rc) return rc;
  const long n = (long)g.xsize * g.ysize * g.zsize;
  std::vector<float> host((size_t)(n > 0 ? n : 1));
  rc = vt_map_download(m, host.data());
  if (rc) return rc;
  return vt_mrc_write(path, &g, host.data());
}

int vt_map_read_mrc(const char *path, vt_map **out) {
  if (!out) { set_error("vt_map_read_mrc: null argument"); return 1; }
  vt_grid g;
  float *data = nullptr;
  int rc = vt_mrc_read(path, &g, &data);
  if (rc) return rc;
  rc = vt_map_create(&g, data, out);
  if (!rc) rc = vt_sync();   // the upload reads `data`
  free(data);
  return rc;
}

int vt_map_write_dx(const vt_map *m, const char *path, const char *name) {
  if (!m) { set_error("vt_map_write_dx: null map"); return 1; }
  vt_grid g;
  int rc = vt_map_grid(m, &g);
  if (rc) return rc;
  const long n = (long)g.xsize * g.ysize * g.zsize;
  std::vector<float> host((size_t)(n > 0 ? n : 1));
  rc = vt_map_download(m, host.data());
  if (rc) return rc;
  return vt_dx_write(path, &g, host.data(), name);
}

int vt_map_read_dx(const char *path, vt_map **out) {
  if (!out) { set_error("vt_map_read_dx: null argument"); return 1; }
  vt_grid g;
  float *data = nullptr;
  int rc = vt_dx_read(path, &g, &data);
  if (rc) return rc;
  rc = vt_map_create(&g, data, out);
  if (!rc) rc = vt_sync();   // the upload reads `data`
  free(data);
  return rc;
}

}  // extern "C"
