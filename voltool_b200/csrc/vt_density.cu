// vt_density.cu -- the two search variants SURVEY 8f-4 names, composed from the resident-map primitives.
//
// (1) The map-rotating search: density_rotate / do_density_rotate / reset_density (TclVoltool.C:202-311; its caller
//     is commented out, :820-891).  Nothing is re-splatted: per candidate the synthetic map's FRAME (origin + axes)
//     is rotated on the host (vol_moveto / vol_move, Voltool.C:379-438), its density-weighted centre recomputed on
//     the device (vol_com with the exact sequential float sum, vt_volcom.cu) and moved back onto com, and the
//     candidate scored with cc_threaded(target, synth, 0.1) -- the generic any-cell CC kernel, since the rotated
//     map's cell is skewed against the target's.  One deliberate deviation, shared with the oracle twin
//     (vo_density_rotate): the dead caller hands the loop synthvol's own origin / axis arrays as the copies to
//     restore from (:874-877), which makes the restore at :262-273 a self-assignment; here the frame is restored
//     from true copies, as the loop is written to do.
// (2) The rotation x translation grid search of BASELINE cfg2: every rotation of a list combined with every shift of
//     a (2 t + 1)^3 lattice, scored by calc_cc (TclVoltool.C:73-129) through the batched pose evaluator, arg max in
//     candidate order.  With a communicator up (vt_nccl_init) rank r scores candidates r, r + world, ... and the
//     per-rank (cc, index) pairs are merged by one ncclAllGather.
#include <cmath>
#include <cstring>
#include <vector>

#include "vt_device.cuh"

using namespace vt;

namespace {

constexpr double kPi = 3.14159265358979323846;

// one candidate's frame: TclVoltool.C:210-226
int density_pose(vt_map *S, const float com[3], int stride, const int amt[3]) {
  const float zero[3] = {0.f, 0.f, 0.f};
  vol_moveto(S->grid, com, zero);
  for (int ax = 0; ax < 3; ++ax) {
    const double amount = (stride * amt[ax] * kPi / 180.0);   // DEGTORAD(stride * amt), :141
    Mat4 R;
    mat4_axis_rotation(R, ax, (float)amount);
    vol_move(S->grid, R.m);
  }
  float currcom[3];
  const int rc = vt_vol_com(S, currcom);
  if (rc) return rc;
  vol_moveto(S->grid, currcom, com);
  return 0;
}

// merge of the per-rank winners (one ncclAllGather of (cc, index bits) per rank; every rank then picks the same one)
template <class PoseOf>
int finish_search(const ShardConfig &sh, bool sharded, float bv, long bi, float *best_cc, long *best_index, vt_pose *best_pose,
                  PoseOf pose_of) {
  if (sharded) {
    Ctx &c = ctx();
    float *d_buf = nullptr;
    if (dev_alloc((void **)&d_buf, sizeof(float) * 2 * (size_t)(sh.world + 1))) return 2;
    float mine[2];
    const int bi32 = (int)bi;
    mine[0] = bv;
    memcpy(&mine[1], &bi32, sizeof(int));
    std::vector<float> all(2 * (size_t)sh.world);
    int rc = 0;
    cudaError_t e = cudaMemcpyAsync(d_buf, mine, sizeof(mine), cudaMemcpyHostToDevice, c.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c.stream);   // `mine` is pageable
    if (e == cudaSuccess) rc = vt_nccl_allgather(d_buf, d_buf + 2, 2);
    if (e == cudaSuccess && !rc) e = cudaMemcpyAsync(all.data(), d_buf + 2, sizeof(float) * all.size(), cudaMemcpyDeviceToHost, c.stream);
    if (e == cudaSuccess && !rc) e = cudaStreamSynchronize(c.stream);
    dev_free(d_buf);
    if (e != cudaSuccess) return cuda_fail(e, "vt_fit_search_grid merge", __FILE__, __LINE__);
    if (rc) return rc;
    bv = -INFINITY; bi = -1;
    for (int r = 0; r < sh.world; ++r) {
      int idx;
      memcpy(&idx, &all[2 * r + 1], sizeof(int));
      const float v = all[2 * r];
      if (idx >= 0 && (v > bv || (v == bv && idx < bi))) { bv = v; bi = idx; }
    }
  }
  *best_cc = bi >= 0 ? bv : NAN;
  *best_index = bi;
  if (best_pose && bi >= 0) *best_pose = pose_of(bi);
  return 0;
}

}  // namespace

extern "C" {

int vt_density_rotate(const vt_map *target, vt_map *synth, int stride, int max_rot, const float com[3], float *returncc,
                      int bestrot[3], float *table) {
  VT_REQUIRE_CTX();
  if (!target || !synth || !com || !returncc || !bestrot || max_rot < 0) { set_error("vt_density_rotate: bad arguments"); return 1; }
  const vt_grid saved = synth->grid;
  long n = 0;
  for (int x = 0; x < max_rot; ++x)
    for (int y = 0; y < max_rot; ++y)
      for (int z = 0; z < max_rot; ++z) {
        const int amt[3] = {x, y, z};
        int rc = density_pose(synth, com, stride, amt);
        double ccd = 0.0;
        if (!rc) rc = vt_cc(target, synth, 0.1, &ccd, nullptr, nullptr);   // density_calc_cc :66-71, :252
        synth->grid = saved;                                                // :262-273
        if (rc) return rc;
        const float cc = (float)ccd;
        if (table) table[n] = cc;
        ++n;
        if (cc > *returncc) { *returncc = cc; bestrot[0] = x; bestrot[1] = y; bestrot[2] = z; }   // :254-261
      }
  return 0;
}

// the map half of reset_density (TclVoltool.C:283-300); the atom half (:302-311) is vt_pose_apply
int vt_density_reset(vt_map *synth, int stride, const int bestrot[3], const float synthcom[3], const float com[3]) {
  VT_REQUIRE_CTX();
  if (!synth || !bestrot || !synthcom || !com) { set_error("vt_density_reset: bad arguments"); return 1; }
  const float move1[3] = {-1.0f * synthcom[0], -1.0f * synthcom[1], -1.0f * synthcom[2]};
  vol_moveto(synth->grid, synthcom, move1);
  for (int ax = 0; ax < 3; ++ax) {
    const double amount = (stride * bestrot[ax] * kPi / 180.0);
    Mat4 R;
    mat4_axis_rotation(R, ax, (float)amount);
    vol_move(synth->grid, R.m);
  }
  float currcom[3];
  const int rc = vt_vol_com(synth, currcom);
  if (rc) return rc;
  vol_moveto(synth->grid, currcom, com);
  return 0;
}

int vt_fit_search_grid(vt_fit_ctx *fctx, const float com[3], const float *rot_deg, long nrot, int tsteps, float tstep,
                       float *best_cc, long *best_index, vt_pose *best_pose, float *cc_all) {
  VT_REQUIRE_CTX();
  if (!fctx || !com || !rot_deg || nrot <= 0 || tsteps < 0 || !best_cc || !best_index) { set_error("vt_fit_search_grid: bad arguments"); return 1; }
  const long nt = 2L * tsteps + 1, per_rot = nt * nt * nt, total = nrot * per_rot;
  if (total >= (1L << 31)) { set_error("vt_fit_search_grid: more than 2^31 candidates"); return 1; }
  const ShardConfig &sh = shard_config();
  const bool sharded = sh.world > 1 && nccl_active();
  const long first = sharded ? sh.rank : 0, step = sharded ? sh.world : 1;
  auto pose_of = [&](long idx) {
    vt_pose p;
    const long ir = idx / per_rot, it = idx % per_rot;
    const long ix = it / (nt * nt), iy = (it / nt) % nt, iz = it % nt;
    for (int k = 0; k < 3; ++k) p.rot_deg[k] = rot_deg[3 * ir + k];
    p.shift[0] = (float)(ix - tsteps) * tstep;
    p.shift[1] = (float)(iy - tsteps) * tstep;
    p.shift[2] = (float)(iz - tsteps) * tstep;
    return p;
  };
  // translations of one rotation share that rotation's splat (fit_eval_rot_shifts): ranks then split the ROTATIONS
  const char *shared_env = getenv("VOLTOOL_GRID_SHARED");
  if (tsteps > 0 && !(shared_env && atoi(shared_env) == 0)) {
    std::vector<float> shifts(3 * (size_t)per_rot);
    for (long it = 0; it < per_rot; ++it) {
      const vt_pose q = pose_of(it);
      for (int k = 0; k < 3; ++k) shifts[3 * it + k] = q.shift[k];
    }
    std::vector<float> rots;
    std::vector<long> rot_index;
    for (long ir = first; ir < nrot; ir += step) {
      rot_index.push_back(ir);
      for (int k = 0; k < 3; ++k) rots.push_back(rot_deg[3 * ir + k]);
    }
    std::vector<float> cc(rot_index.size() * (size_t)per_rot + 1);
    int rc = fit_eval_rot_shifts(fctx, com, rots.data(), (long)rot_index.size(), shifts.data(), (int)per_rot, cc.data());
    if (rc) return rc;
    if (cc_all && sharded) for (long i = 0; i < total; ++i) cc_all[i] = NAN;
    float bv = -INFINITY;
    long bi = -1;
    for (size_t r = 0; r < rot_index.size(); ++r)   // ascending candidate index: `if (cc > best)` keeps the first maximum
      for (long it = 0; it < per_rot; ++it) {
        const float v = cc[r * per_rot + it];
        const long idx = rot_index[r] * per_rot + it;
        if (cc_all) cc_all[idx] = v;
        if (v > bv) { bv = v; bi = idx; }
      }
    return finish_search(sh, sharded, bv, bi, best_cc, best_index, best_pose, pose_of);
  }
  std::vector<vt_pose> poses;
  poses.reserve((size_t)((total - first + step - 1) / step));
  for (long idx = first; idx < total; idx += step) poses.push_back(pose_of(idx));
  std::vector<float> cc(poses.size() ? poses.size() : 1);
  int rc = poses.empty() ? 0 : vt_fit_eval_poses(fctx, com, poses.data(), (long)poses.size(), cc.data());
  if (rc) return rc;
  // arg max in candidate order: `if (cc > best)` keeps the first maximum and never takes a NaN (TclVoltool.C:181)
  float bv = -INFINITY;
  long bi = -1;
  for (size_t k = 0; k < poses.size(); ++k)
    if (cc[k] > bv) { bv = cc[k]; bi = first + (long)k * step; }
  if (cc_all) {
    if (sharded) for (long i = 0; i < total; ++i) cc_all[i] = NAN;   // a rank only knows its own candidates
    for (size_t k = 0; k < poses.size(); ++k) cc_all[first + (long)k * step] = cc[k];
  }
  return finish_search(sh, sharded, bv, bi, best_cc, best_index, best_pose, pose_of);
}

}  // extern "C"
