// vt_host.h -- host-side (CPU, double/float) metadata arithmetic of the voltool GPU path.
//
// Everything here is O(1) per call (or O(poses)): grid geometry, the loop-invariant 4x4
// inverse, parameter policy, rotation matrices, the fit schedule and the skip-rule replay.
// It is compiled by g++ with -ffp-contract=off so that every expression rounds exactly where
// the reference's C++ does; the kernels receive the finished constants.
// Citations: file:line into the ryanmcgreevy/voltool snapshot.
#pragma once
#include "../../include/voltool_gpu.h"

#include <cstdint>

namespace vt {

struct Mat4 {  // column-major, translation in m[12..14] (as used at VolumetricData.C:207-223)
  float m[16];
};

void mat4_identity(Mat4 &a);
bool mat4_invert(Mat4 &a);                                   // false if singular
void mat4_point(const Mat4 &a, const float in[3], float out[3]);
void mat4_axis_rotation(Mat4 &a, int axis, float angle_rad); // Matrix4::rotate_axis for unit axes

// per-map constants handed to kernels
struct MapGeom {
  float inv0[16];  // cart -> voxel, no half-voxel shift (VolumetricData.C:194-228, shiftflag=false)
  float inv1[16];  // with the half-voxel shift (nearest lookups)
  int nx, ny, nz;
};
// output-grid constants: VolumetricData::voxel_coord (VolumetricData.h:131-153), Voltool.h:50-77
struct CoordGeom {
  double origin[3];
  float xd[3], yd[3], zd[3];  // cell_axes, VolumetricData.C:138-168
  int nx, ny, nz;
};

void cell_axes(const vt_grid &g, float xd[3], float yd[3], float zd[3]);
void map_geom(const vt_grid &g, MapGeom &out);
void coord_geom(const vt_grid &g, CoordGeom &out);
void grid_inverse(const vt_grid &g, bool half_voxel_shift, float inv[16]);
long gridsize(const vt_grid &g);

void intersection_grid(const vt_grid &A, const vt_grid &B, vt_grid &o);  // Voltool.C:51-78
void union_grid(const vt_grid &A, const vt_grid &B, vt_grid &o);         // Voltool.C:114-147

// VolumetricData::pad geometry (VolumetricData.C:542-609); returns new sizes + copy window
struct PadPlan {
  int nx, ny, nz;           // new sizes
  int sx, sy, sz, ex, ey, ez;  // copy window in NEW index space
  int pxm, pym, pzm;
};
void pad_plan(vt_grid &g /*updated*/, int pxm, int pxp, int pym, int pyp, int pzm, int pzp, PadPlan &p);
void crop_pads(const vt_grid &g, double minx, double miny, double minz, double maxx, double maxy, double maxz,
               int pads[6]);                                              // VolumetricData.C:615-631
void downsample_geom(vt_grid &g);                                         // VolumetricData.C:744-757
void supersample_geom(vt_grid &g);                                        // VolumetricData.C:778-781, 915-917

void vol_moveto(vt_grid &g, const float com[3], const float pos[3]);      // Voltool.C:379-391
void vol_move(vt_grid &g, const float mat[16]);                           // Voltool.C:399-438
double cc_from_sums(const double s[5], long n);                           // MDFF.C:143-149
int blur_taps(double sigma, float *w, int maxw);                          // GaussianBlur restatement

float gausslim_for_quality(int quality);
float sim_padding(float radscale, float maxrad);
void sim_policy(double resolution, double gspacing_in, float &radscale, double &gspacing, int &quality);

// rotation of one do_rotate() call: cos/sin as float, from the float-rounded angle
// (TclVoltool.C:131-136, DEGTORAD utilities.h:49)
void euler_cs(double degrees, float &c, float &s);
void pose_apply(float *pos, long natoms, const float com[3], const vt_pose &pose);
void measure_center(const float *pos, const float *w, long natoms, float com[3]);
int rotate_replay(const float *table, int max_rot, float *returncc, int *best_idx);

// how the pose search is sharded (vt_fit_set_shard / vt_nccl_init): rank r evaluates candidates r, r + world, ...
struct ShardConfig {
  int rank = 0, world = 1;
  vt_table_merge_fn merge = nullptr;   // host-side merge callback; nullptr with world > 1: the evaluator merges (NCCL)
  void *user = nullptr;
};
const ShardConfig &shard_config();
bool nccl_active();                    // vt_dist.cu: a communicator is up, tables merge on the device

}  // namespace vt
