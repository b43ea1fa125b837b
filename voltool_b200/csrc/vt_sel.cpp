// vt_sel.cpp -- atom selections at the boundary.  The reference walks `for (i = sel->firstsel; i <= sel->lastsel;
// i++) if (sel->on[i])` over the molecule's full coordinate / radius / mass arrays (moveby TclVoltool.C:52-59,
// the QuickSurf call sites TclMDFF.C:154-157, 496-499, measure_center :789).  The kernels work on compacted
// arrays, so these entry points compact by the flags, run the compacted call and scatter coordinates back:
// unselected atoms are never read by the arithmetic and never written.
#include <cstring>
#include <vector>

#include "vt_host.h"

namespace vt {
extern void set_error(const char *fmt, ...);

static long compact(const float *pos, const float *radii, const float *mass, const int *on, long num_atoms,
                    std::vector<float> &p, std::vector<float> &r, std::vector<float> &m, std::vector<long> &idx) {
  long first = -1, last = -1;
  for (long i = 0; i < num_atoms; ++i)
    if (on[i]) { if (first < 0) first = i; last = i; }
  if (first < 0) return 0;
  for (long i = first; i <= last; ++i) {   // firstsel .. lastsel, ascending, like every AtomSel loop of the reference
    if (!on[i]) continue;
    idx.push_back(i);
    p.push_back(pos[3 * i]); p.push_back(pos[3 * i + 1]); p.push_back(pos[3 * i + 2]);
    if (radii) r.push_back(radii[i]);
    if (mass) m.push_back(mass[i]);
  }
  return (long)idx.size();
}
}  // namespace vt

using namespace vt;

extern "C" {

int vt_compact_selection(const float *pos, const float *radii, const float *mass, const int *on, long num_atoms,
                         float *pos_out, float *radii_out, float *mass_out, long *index_out, long *nsel) {
  if (!pos || !on || !nsel || num_atoms < 0) { set_error("vt_compact_selection: bad arguments"); return 1; }
  std::vector<float> p, r, m;
  std::vector<long> idx;
  const long n = compact(pos, radii, mass, on, num_atoms, p, r, m, idx);
  *nsel = n;
  if (pos_out && n) memcpy(pos_out, p.data(), sizeof(float) * 3 * (size_t)n);
  if (radii_out && radii && n) memcpy(radii_out, r.data(), sizeof(float) * (size_t)n);
  if (mass_out && mass && n) memcpy(mass_out, m.data(), sizeof(float) * (size_t)n);
  if (index_out && n) memcpy(index_out, idx.data(), sizeof(long) * (size_t)n);
  return 0;
}

int vt_sim_sel(const float *pos, const float *radii, const int *on, long num_atoms, int quality, float radscale,
               float gspacing, vt_map **out) {
  if (!pos || !radii || !on) { set_error("vt_sim_sel: bad arguments"); return 1; }
  std::vector<float> p, r, m;
  std::vector<long> idx;
  const long n = compact(pos, radii, nullptr, on, num_atoms, p, r, m, idx);
  if (n <= 0) { set_error("vt_sim_sel: empty selection"); return 1; }
  return vt_sim(p.data(), r.data(), n, quality, radscale, gspacing, out);
}

int vt_sim_cc_sel(const float *pos, const float *radii, const int *on, long num_atoms, double resolution,
                  double gspacing_in, const vt_map *target, double threshold, double *cc, vt_map **synth_out) {
  if (!pos || !radii || !on) { set_error("vt_sim_cc_sel: bad arguments"); return 1; }
  std::vector<float> p, r, m;
  std::vector<long> idx;
  const long n = compact(pos, radii, nullptr, on, num_atoms, p, r, m, idx);
  if (n <= 0) { set_error("vt_sim_cc_sel: empty selection"); return 1; }
  return vt_sim_cc(p.data(), r.data(), n, resolution, gspacing_in, target, threshold, cc, synth_out);
}

int vt_fit_sel(float *pos, const float *radii, const float *mass, const int *on, long num_atoms, const vt_map *target,
               float resolution, vt_fit_result *res) {
  if (!pos || !radii || !mass || !on) { set_error("vt_fit_sel: bad arguments"); return 1; }
  std::vector<float> p, r, m;
  std::vector<long> idx;
  const long n = compact(pos, radii, mass, on, num_atoms, p, r, m, idx);
  if (n <= 0) { set_error("vt_fit_sel: empty selection"); return 1; }
  const int rc = vt_fit(p.data(), r.data(), m.data(), n, target, resolution, res);
  if (rc) return rc;
  for (long k = 0; k < n; ++k) {   // framepos[i] = bestpos[...] for the selected atoms only (TclVoltool.C:911-913)
    const long i = idx[(size_t)k];
    pos[3 * i] = p[3 * k]; pos[3 * i + 1] = p[3 * k + 1]; pos[3 * i + 2] = p[3 * k + 2];
  }
  return 0;
}

}  // extern "C"
