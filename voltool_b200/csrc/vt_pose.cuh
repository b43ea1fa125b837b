// vt_pose.cuh -- data structures of the batched pose pipeline (sim + cc for many rigid poses of one
// atom set against one resident target map).
#pragma once
#include "vt_device.cuh"
#include "vt_geom.h"

namespace vt {

// one candidate, host-prepared: cos/sin of the three Euler angles as the reference's do_rotate
// forms them (TclVoltool.C:131-136), plus the extension translation
struct PrepPose {
  float c[3], s[3];
  float shift[3];
};

struct SimParams {
  float radscale, gs, invgs, gausslim;
  int quality;
};

// splat tile (voxels) and staging sizes
constexpr int kSplatWarps = 8;                 // warps per splat CTA: regions are 8x8x4, 2 in x, kSplatWarps/4 in y, 2 in z
constexpr int kTX = 16, kTY = 2 * kSplatWarps, kTZ = 8;
constexpr int kChunk = 64;     // atoms staged per compute round
constexpr int kRec = 44;       // floats per staged atom record
constexpr int kMaxIds = 2048;  // gathered atom ids per flush
constexpr int kZChunkFit = 32;
constexpr int kRowsFit = 4;     // output rows per thread in the batched cc kernel

// per-pose geometry, written by pose_setup_kernel
struct PoseGeom {
  float org[3];       // sim grid origin (float, as QuickSurf keeps it)
  int nv[3];          // sim grid size
  int no[3];          // intersection grid size (cc)
  int valid;          // 0: cc undefined for this pose (empty intersection / capacity)
  double oorg[3];     // intersection origin
  float od[3];        // intersection cell deltas (diagonal)
  float sinv_d[3];    // sim map inverse: diagonal
  float sinv_t[3];    //                 translation
  int tiles[3];
  int tile_base;      // exclusive prefix of tile counts over the batch
  int cc_items;       // warp work items of the cc stage
  int cc_base;        // exclusive prefix
  int cc_xt, cc_zc;
};

// resident atom set, sorted along a Morton curve and cut into groups of 32
struct AtomSet {
  int n = 0, ngroups = 0;
  float4 *d_xyzr = nullptr;   // origin-frame coordinates + radius (sorted order)
  float4 *d_const = nullptr;  // per atom: arinv, radlim^2, radlim/gs, radius
  float4 *d_group = nullptr;  // per group: centre (origin frame) + bounding-sphere radius
  float *d_gcut = nullptr;    // per group: largest cutoff (gausslim * sigma) of its atoms
  float *d_rawpos = nullptr;  // coordinates in the caller's order (the last upload)
  float *d_rawrad = nullptr;  // radii in the caller's order
  int *d_perm = nullptr;      // sorted position -> caller's atom index
  float maxrad = 0.f, pad = 0.f;
  float bound_center[3] = {0, 0, 0};
  float bound_radius = 0.f;   // all atoms within this distance of bound_center
  SimParams prm{};
};

// workspace for a batch of B poses
struct PoseBatch {
  int B = 0;              // slots
  int cap_n = 0;          // sim grid capacity per axis
  long cap_grid = 0;      // floats per sim grid slot
  int cap_o = 0;          // intersection capacity per axis
  long cap_items = 0;     // cc work items capacity per pose
  float4 *d_tpos = nullptr;
  float4 *d_tgroup = nullptr;
  unsigned int *d_bbox = nullptr;
  PoseGeom *d_geom = nullptr;
  int *d_totals = nullptr;  // [0] tiles, [1] cc items, [2] list overflow flag, [3] calibration scratch, [4..6] task queues (splat, cc, gather)
  unsigned int *d_bitmap = nullptr;  // [B][tiles_cap][ceil(ngroups/32)]: groups that can reach a tile
  int tiles_cap = 0;
  // barrier-free splat (vt_splat2.cu): packed integer boxes, per-tile atom lists
  uint3 *d_boxes = nullptr;          // [B][natoms]
  uint3 *d_gbox = nullptr;           // [B][ngroups]: union of the boxes of a group's atoms (same packing; lo > hi: empty)
  unsigned int *d_tlist = nullptr;   // [B][tiles_cap][cap_t]: id | region mask << 24
  int *d_tcount = nullptr;           // [B][tiles_cap]
  int cap_t = 0;
  bool use_lists = false;
  int splat_ctas = 0;                // resident splat CTAs per SM (0: default)
  float *d_grid = nullptr;
  geom::AxisEntry *d_tab = nullptr;  // [B][6][cap_o]
  double *d_part = nullptr;          // [B][cap_items][6]
  PrepPose *d_prep = nullptr;        // [B]
  float *d_org0 = nullptr;           // [B][3]: sim grid origins as splatted (batch_retarget)
};

int atomset_create(const float *pos, const float *radii, long n, const SimParams &prm, AtomSet &as);
int atomset_set_positions(AtomSet &as, const float *pos);
void atomset_destroy(AtomSet &as);

int batch_create(const AtomSet &as, const vt_grid *target, int B, PoseBatch &pb);
void batch_destroy(PoseBatch &pb);

// stage 1: transform (or copy, identity != 0) + bbox + per-pose sim geometry (+ cc geometry if target)
int batch_prepare(const AtomSet &as, PoseBatch &pb, int npose, const PrepPose *d_prep, const float com[3],
                  int identity, const vt_grid *target);
// re-aim the batch's splatted grids at the frame translated by `shift` (relative to where they were splatted): new
// origin, intersection with the target, tables.  first: remember the splatted origins.  Then batch_cc as usual.
int batch_retarget(const AtomSet &as, PoseBatch &pb, int npose, const float shift[3], bool first, const vt_grid *target);
// stage 2: splat every pose's density grid
int batch_splat(const AtomSet &as, PoseBatch &pb, int npose);
// size and switch on the list-based splat from the pose(s) currently prepared in pb (needs
// batch_prepare first); leaves the cooperative kernel in place when it does not apply
int batch_calibrate_lists(const AtomSet &as, PoseBatch &pb);
// single-pose variant without the counting pass (slot sized from the mean tile fill); nv = sim grid size
int batch_estimate_lists(const AtomSet &as, PoseBatch &pb, const int nv[3]);
int batch_splat_lists(const AtomSet &as, PoseBatch &pb, int npose);
// the splat in two halves, so that a caller can run the list building of the next batch on a second
// stream while this batch accumulates: front = tile bitmaps (+ boxes and per-tile lists), back = the
// accumulation kernel.  batch_splat == front then back.
int batch_splat_front(const AtomSet &as, PoseBatch &pb, int npose);
int batch_splat_back(const AtomSet &as, PoseBatch &pb, int npose);
int batch_bin(const AtomSet &as, PoseBatch &pb, int npose);
// 1 if a tile list overflowed since the last reset (synchronises)
int batch_lists_overflowed(PoseBatch &pb, int *flag);
// stage 3: cc of every pose's grid against the target; d_cc[npose] floats
int batch_cc(PoseBatch &pb, int npose, const vt_map *target, double threshold, float *d_cc, double *d_sums);

}  // namespace vt
