// vt_io.cpp -- OpenDX scalar-field files at the boundary of the path.
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
