// vt_fit_host.cpp -- the fit() schedule of TclVoltool.C:784-916 as host logic with an injected
// candidate evaluator.  No CUDA here: the CUDA evaluator lives in vt_fit.cu; tests inject their
// own (and run the multi-rank merge over gloo).
#include <cmath>
#include <cstring>
#include <vector>

#include "vt_host.h"

namespace vt {

static ShardConfig g_shard;
const ShardConfig &shard_config() { return g_shard; }

extern void set_error(const char *fmt, ...);

}  // namespace vt

using namespace vt;

extern "C" {

int vt_fit_set_shard(int rank, int world, vt_table_merge_fn merge, void *user) {
  if (world < 1 || rank < 0 || rank >= world) return 1;
  if (world > 1 && !merge && !nccl_active()) return 1;   // without a callback the tables must merge over NCCL
  g_shard.rank = rank; g_shard.world = world; g_shard.merge = merge; g_shard.user = user;
  return 0;
}

int vt_rotate_replay(const float *table, int max_rot, float *returncc, int *best_idx) {
  return rotate_replay(table, max_rot, returncc, best_idx);
}

int vt_pose_apply(float *pos, long natoms, const float com[3], const vt_pose *pose) {
  pose_apply(pos, natoms, com, *pose);
  return 0;
}

int vt_measure_center(const float *pos, const float *weight, long natoms, float com[3]) {
  measure_center(pos, weight, natoms, com);
  return 0;
}

int vt_sim_policy(double resolution, double gspacing_in, float *radscale, double *gspacing, int *quality) {
  sim_policy(resolution, gspacing_in, *radscale, *gspacing, *quality);
  return 0;
}

// fit(): COM alignment, then rotate() five times on the 24 / +-4 / +-1 degree lattices with the
// running best carried across calls, origin reset after the first and third (TclVoltool.C:895-909).
int vt_fit_schedule(float *pos, const float *mass, long natoms, const float dencom[3], vt_table_eval_fn eval,
                    void *user, vt_fit_result *res) {
  if (!pos || !mass || !eval || !res || natoms <= 0) return 1;
  const size_t nfl = 3 * (size_t)natoms;
  float com[3];
  measure_center(pos, mass, natoms, com);                                   // :789
  const float mv[3] = {dencom[0] - com[0], dencom[1] - com[1], dencom[2] - com[2]};
  for (long i = 0; i < natoms; ++i) {                                       // moveby :52-59, :801
    pos[3 * i] = pos[3 * i] + mv[0];
    pos[3 * i + 1] = pos[3 * i + 1] + mv[1];
    pos[3 * i + 2] = pos[3 * i + 2] + mv[2];
  }
  measure_center(pos, mass, natoms, com);                                   // :803
  std::vector<float> origin(pos, pos + nfl), best(pos, pos + nfl);          // :812-817
  // rotate() transforms `framepos` in place and restores it from `origin` after every candidate
  // (:191-195).  reset_origin() (:314-318) rewrites `origin` only, so the FIRST candidate of the
  // rotate() call that follows a reset still starts from the previous origin: it re-scores the old
  // origin under the zero rotation.  `frame` tracks what framepos holds when a call starts, and
  // `frame_cc0` the score of its zero-rotation candidate (candidate 0 of the last table built on it).
  std::vector<float> frame(pos, pos + nfl);
  float frame_cc0 = NAN;
  float cc = -1;
  memset(res, 0, sizeof(*res));
  for (int k = 0; k < 3; ++k) { res->com[k] = com[k]; res->dencom[k] = dencom[k]; }
  res->exhaustive_cc = -INFINITY;

  static const int strides[5] = {24, 4, -4, 1, -1};
  static const int maxrots[5] = {360 / 24, 24 / 4, 24 / 4, 4 / 1, 4 / 1};
  const ShardConfig &sh = shard_config();
  for (int s = 0; s < 5; ++s) {
    const int stride = strides[s], m = maxrots[s], total = m * m * m;
    std::vector<float> table(total, NAN);
    int rc = eval(user, origin.data(), com, stride, m, sh.rank, sh.world, table.data());
    if (rc) return rc;
    if (sh.world > 1 && sh.merge) {   // (no callback: the evaluator returned the merged table, vt_dist.cu)
      rc = sh.merge(table.data(), total, sh.rank, sh.world, sh.user);
      if (rc) return rc;
    }
    res->n_computed += total;
    for (int i = 0; i < total; ++i)
      if (table[i] > res->exhaustive_cc) res->exhaustive_cc = table[i];
    const float fresh_cc0 = table[0];
    const bool stale = memcmp(frame.data(), origin.data(), sizeof(float) * nfl) != 0;
    if (stale) table[0] = frame_cc0;
    int bi = -1;
    res->n_evaluated += rotate_replay(table.data(), m, &cc, &bi);
    if (bi >= 0) {                                                           // bestpos <- framepos :181-185
      best = (bi == 0 && stale) ? frame : origin;
      vt_pose pose;
      pose.rot_deg[0] = (float)(stride * (bi / (m * m)));
      pose.rot_deg[1] = (float)(stride * ((bi / m) % m));
      pose.rot_deg[2] = (float)(stride * (bi % m));
      pose.shift[0] = pose.shift[1] = pose.shift[2] = 0.f;
      pose_apply(best.data(), natoms, com, pose);
    }
    res->stage_best[s] = bi;
    frame = origin;                                                          // restored after the last candidate
    frame_cc0 = fresh_cc0;
    if (s == 0 || s == 2) origin = best;                                     // reset_origin :898, :904
  }
  memcpy(pos, best.data(), sizeof(float) * nfl);                             // :911-913
  res->best_cc = cc;
  return 0;
}

}  // extern "C"
