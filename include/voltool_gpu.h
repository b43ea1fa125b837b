/*
 * voltool_gpu.h -- C ABI of libvoltool_gpu.so, the B200 (sm_100a) drop-in for voltool's
 * data-parallel hot path.  Plain pointers and sizes only; no C++/torch types cross this line.
 *
 * Every entry point names the reference interface it replaces (file:line into the
 * ryanmcgreevy/voltool snapshot).  INTEGRATION.md shows the replacement bodies a VMD maintainer
 * would put into MDFF.C / VolumetricData.C / Voltool.C / TclVoltool.C to call these.
 *
 * Conventions
 *   - every function returns 0 on success, nonzero on failure; vt_last_error() holds the text.
 *     There is no CPU fallback: without a usable CUDA device vt_init() fails and so does
 *     everything else.
 *   - vt_map is the resident twin of VolumetricData (VolumetricData.h:27-36): a float grid in
 *     HBM (x fastest, then y, then z) + double origin/axes + the cached min/max/mean/sigma block
 *     (VolumetricData.h:202-210) maintained by exactly the reference's invalidation rules.
 *   - "host" variants take host pointers, copy in, run, copy out (one call = one reference call).
 *   - single Tcl thread model (tcl_commands.C:291-295): the library keeps one process-global
 *     context and is not re-entrant.
 */
#ifndef VOLTOOL_GPU_H
#define VOLTOOL_GPU_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* mirrors VolumetricData.h:29-33 */
typedef struct vt_grid {
  double origin[3];
  double xaxis[3];
  double yaxis[3];
  double zaxis[3];
  int xsize, ysize, zsize;
} vt_grid;

/* mirrors the cached-statistics block, VolumetricData.h:202-210 */
typedef struct vt_cache {
  int minmax_valid, mean_valid, sigma_valid;
  float cmin, cmax, cmean, csigma;
} vt_cache;

typedef struct vt_map vt_map; /* opaque resident map */

/* ---------------------------------------------------------------- lifecycle */
/* device < 0: honour VOLTOOL_GPU_DEVICE, else LOCAL_RANK, else 0.  Replaces the
   "vmdinfo numcudadevices"/VMDCUDA bring-up the old hooks relied on (tcl_commands.C:107-118). */
int vt_init(int device);
int vt_shutdown(void);
int vt_sync(void);
const char *vt_last_error(void);
/* sm_count, max SM clock (kHz), total device memory; any pointer may be NULL */
int vt_device_props(int *sm_count, int *clock_khz, size_t *mem_bytes, char *name, int name_len);
/* number of kernels this library has launched since vt_init (bench.py's gpu_launches) */
long vt_launch_count(void);
/* optional CUDA-event timing of the library's own kernels on its stream, by kernel class:
   "transform" "setup" "splat" "pose_cc" "pose_finish" "cc" "stats" "unary" "binary" "blur" "resample".
   vt_profile_get synchronises and returns the summed duration (ms) and the number of launches. */
int vt_profile_enable(int on);
int vt_profile_reset(void);
int vt_profile_get(const char *kernel_class, double *ms_total, long *count);
/* all library work runs on one stream; torch callers use this to order their own events */
void *vt_stream(void);

/* ---------------------------------------------------------------- resident maps */
/* VolumetricData::VolumetricData (VolumetricData.C:66-89): adopts a copy of host_data and runs
   compute_minmax() like the constructor does.  host_data == NULL gives a zero-filled map. */
int vt_map_create(const vt_grid *g, const float *host_data, vt_map **out);
/* as above from a DEVICE pointer (device-to-device copy) */
int vt_map_create_from_device(const vt_grid *g, const float *dev_data, vt_map **out);
/* the twin of an EXISTING VolumetricData (shim/VolumetricDataGPU.C): uploads its voxels and takes its
   cached-statistics block as it stands instead of recomputing it */
int vt_map_adopt_host(const vt_grid *g, const float *host_data, const vt_cache *cache, vt_map **out);
int vt_map_clone(const vt_map *m, vt_map **out);
int vt_map_destroy(vt_map *m); /* VolumetricData::~VolumetricData, VolumetricData.C:93-99 */
int vt_map_grid(const vt_map *m, vt_grid *g);
int vt_map_cache(const vt_map *m, vt_cache *c);
long vt_map_gridsize(const vt_map *m); /* VolumetricData::gridsize, VolumetricData.h:78-80 */
int vt_map_download(const vt_map *m, float *host_out);
int vt_map_upload(vt_map *m, const float *host_in); /* replaces data, reruns compute_minmax */
float *vt_map_device_ptr(vt_map *m);

/* ---------------------------------------------------------------- cross correlation */
/* cc_threaded(qsVol, targetVol, &cc, threshold)  MDFF.C:102-151 (+ correlationthread :55-99).
   A is the thresholded map (MDFF.C:77).  nused/sums may be NULL; sums = {sumA, sumB, sumAA,
   sumBB, sumAB}. */
int vt_cc(const vt_map *A, const vt_map *B, double threshold, double *cc, long *nused, double sums[5]);
int vt_cc_host(const vt_grid *gA, const float *a, const vt_grid *gB, const float *b, double threshold,
               double *cc);

/* ---------------------------------------------------------------- unary streams (in place) */
int vt_clamp(vt_map *m, float min_value, float max_value);            /* VolumetricData::clamp :634-653 */
int vt_scale_by(vt_map *m, float ff);                                 /* ::scale_by :656-677 */
int vt_scalar_add(vt_map *m, float ff);                               /* ::scalar_add :680-694 */
int vt_rescale_voxel_value_range(vt_map *m, float min_value, float max_value); /* :697-714 */
int vt_sigma_scale(vt_map *m);                                        /* ::sigma_scale :937-953 */
int vt_binmask(vt_map *m, float threshold);                           /* ::binmask :957-969 */
int vt_mdff_potential(vt_map *m, double threshold);                   /* ::mdff_potential :999-1025 */

/* ---------------------------------------------------------------- statistics */
int vt_datarange(vt_map *m, float *min_value, float *max_value);      /* ::datarange :1035-1044 */
int vt_mean(vt_map *m, float *mean);                                  /* ::mean :1078-1086 */
int vt_sigma(vt_map *m, float *sigma);                                /* ::sigma :1095-1100 */
int vt_vol_com(const vt_map *m, float com[3]);                        /* vol_com, Voltool.C:229-247 */
int vt_histogram(vt_map *m, int nbins, long *bins, float *midpts);    /* histogram, Voltool.C:443-461 */
/* metadata only: translate / rigidly transform the map frame (column-major 4x4, translation in m[12..14]) */
int vt_vol_moveto(vt_map *m, const float com[3], const float pos[3]); /* vol_moveto, Voltool.C:379-391 */
int vt_vol_move(vt_map *m, const float mat[16]);                       /* vol_move, Voltool.C:399-438 */

/* ---------------------------------------------------------------- map files
 * OpenDX scalar fields as VMD's volmap_write_dx_file produces them (call sites TclMDFF.C:146,159,538;
 * the writer is not in the snapshot: layout restated -- z fastest on disk, x fastest in memory).
 * vt_dx_* work on host arrays and need no GPU; vt_dx_read mallocs *data (release: vt_free_host). */
int vt_dx_write(const char *path, const vt_grid *g, const float *data, const char *name);
int vt_dx_read(const char *path, vt_grid *g, float **data);
int vt_map_write_dx(const vt_map *m, const char *path, const char *name);  /* `-o <file>` of sim/cc, TclMDFF.C:159 */
int vt_map_read_dx(const char *path, vt_map **out);                         /* `-i <input map>`, TclMDFF.C:176 */
/* MRC2014 / CCP4 maps (what VMD's ccp4plugin hands to `mol new`; the plugin is not in the snapshot, the
 * mapping onto VolumetricData geometry is restated in vt_io.cpp).  Read: modes 0/1/2/6, any axis order,
 * either byte order, orthogonal cells.  Write: mode 2.  vt_mrc_* are host only; vt_mrc_read mallocs *data
 * (release: vt_free_host). */
int vt_mrc_write(const char *path, const vt_grid *g, const float *data);
int vt_mrc_read(const char *path, vt_grid *g, float **data);
int vt_map_write_mrc(const vt_map *m, const char *path);
int vt_map_read_mrc(const char *path, vt_map **out);

/* ---------------------------------------------------------------- resizing ops (replace data) */
int vt_pad(vt_map *m, int padxm, int padxp, int padym, int padyp, int padzm, int padzp); /* ::pad :542-609; trim = pad(-t) TclVoltool.C:1185 */
int vt_crop(vt_map *m, double minx, double miny, double minz, double maxx, double maxy, double maxz); /* ::crop :615-631 */
int vt_downsample(vt_map *m);                                         /* ::downsample :720-769 */
int vt_supersample(vt_map *m);                                        /* ::supersample :773-932 */
int vt_gaussian_blur(vt_map *m, double sigma);                        /* ::gaussian_blur :977-997 */

/* ---------------------------------------------------------------- two-map ops */
enum { VT_ADD = 0, VT_SUBTRACT = 1, VT_MULTIPLY = 2, VT_AVERAGE = 3 };
/* add/subtract/multiply/average(mapA, mapB, newvol, interp, USE_UNION)  Voltool.C:249-377,
   output grid from init_from_union / init_from_intersection (Voltool.C:51-147) */
int vt_binary(int op, int interp, int use_union, const vt_map *A, const vt_map *B, vt_map **out);
int vt_binary_host(int op, int interp, int use_union, const vt_grid *gA, const float *a, const vt_grid *gB,
                   const float *b, vt_grid *out_grid, float **out_data /* malloc()ed; vt_free_host */);
void vt_free_host(void *p);
/* geometry only (host double arithmetic, identical to the reference expressions) */
int vt_init_from_intersection(const vt_grid *A, const vt_grid *B, vt_grid *out); /* Voltool.C:51-78 */
int vt_init_from_union(const vt_grid *A, const vt_grid *B, vt_grid *out);        /* Voltool.C:114-147 */

/* ---------------------------------------------------------------- sim (density map from atoms) */
/* parameter policy of mdff_sim / mdff_cc / calc_cc: TclMDFF.C:125-135, 397-401, 492-494 */
int vt_sim_policy(double resolution, double gspacing_in, float *radscale, double *gspacing, int *quality);
/* QuickSurf::calc_density_map(sel, mol, pos, radii, quality, radscale, gspacing)
   call sites TclMDFF.C:154-157, 496-499; TclVoltool.C:119-122.  pos = 3*natoms, radii = natoms
   (selected atoms only, compacted by the caller exactly as the Tcl wrappers' AtomSel loop does). */
int vt_sim(const float *pos, const float *radii, long natoms, int quality, float radscale, float gspacing,
           vt_map **out);
/* calc_cc(sel, target, resolution, ...)  TclVoltool.C:73-129: policy + sim + cc(threshold 0.1) */
int vt_calc_cc(const float *pos, const float *radii, long natoms, const vt_map *target, float resolution,
               double *cc);
/* mdff_cc's CPU branch TclMDFF.C:486-524: sim at the given policy then cc_threaded(sim, target) */
int vt_sim_cc(const float *pos, const float *radii, long natoms, double resolution, double gspacing_in,
              const vt_map *target, double threshold, double *cc, vt_map **synth_out /* may be NULL */);

/* `voltool cc -calcsynthmap -calcdiffmap -calcspatialccmap` (usage TclMDFF.C:180-184, plumbing :555-594).  The
   reference fills the last two only inside vmd_cuda_compare_sel_refmap (CUDAMDFF.cu, not in the snapshot); the
   definition built here (vt_ccmaps.cu, oracle twin vo_cc_maps): on G = init_from_intersection(A, B), the grid
   cc_threaded walks, diff = sample(A) - sample(B); spatial = the CC expression of MDFF.C:143-149 per 8x8x8 tile
   of G (grid: ceil(n/8) tiles per axis on G's origin, one cell = 8 cells of G; 0 for a tile without voxels);
   *cc = the same expression on the tile sums.  diff_out / spatial_out may be NULL. */
int vt_cc_maps(const vt_map *A, const vt_map *B, double threshold, double *cc, vt_map **diff_out, vt_map **spatial_out);
int vt_sim_cc_maps(const float *pos, const float *radii, long natoms, double resolution, double gspacing_in,
                   const vt_map *target, double threshold, double *cc, vt_map **synth_out, vt_map **diff_out,
                   vt_map **spatial_out);

/* ---------------------------------------------------------------- fit (rigid-body pose search) */
/* one candidate of rotate() (TclVoltool.C:160-176) or of an explicit pose list: rotate about
   com by Euler angles (degrees, X then Y then Z, each a separate rounded pass like do_rotate
   :131-136), translate back by +com, then by shift (extension; {0,0,0} for the reference). */
typedef struct vt_pose {
  float rot_deg[3];
  float shift[3];
} vt_pose;

typedef struct vt_fit_result {
  float best_cc;      /* "Best CC" TclVoltool.C:926 */
  float com[3];       /* structure centre after the COM alignment, :803 */
  float dencom[3];    /* vol_com(target), :799 */
  int n_evaluated;    /* candidates the reference's sequential loop visits (skip rule :186-190) */
  int n_computed;     /* candidates this library actually evaluated (all of them) */
  int stage_best[5];  /* winning index per rotate() call, -1 if that call found nothing better */
  float exhaustive_cc;   /* argmax over every evaluated candidate of the final stage set */
} vt_fit_result;

typedef struct vt_fit_ctx vt_fit_ctx;

/* resident state for a pose search: target map, atoms (origin coordinates), radii.  The target
   is borrowed (must outlive the ctx). */
int vt_fit_create(const float *pos, const float *radii, long natoms, const vt_map *target, float resolution,
                  vt_fit_ctx **out);
int vt_fit_destroy(vt_fit_ctx *ctx);
int vt_fit_set_origin(vt_fit_ctx *ctx, const float *pos); /* reset_origin, TclVoltool.C:314-318 */
/* evaluate poses [0,nposes) of a list; cc_out[i] = (float)calc_cc for pose i (host buffers) */
int vt_fit_eval_poses(vt_fit_ctx *ctx, const float com[3], const vt_pose *poses, long nposes, float *cc_out);
/* same with the pose list already uploaded and the result left on the device (for benchmarks and
   for NCCL allgather directly from device memory); d_poses/d_cc are device pointers */
int vt_fit_eval_poses_device(vt_fit_ctx *ctx, const float com[3], const vt_pose *d_poses, long nposes,
                             float *d_cc);
/* all max_rot^3 candidates of one rotate() call; table[x*m*m+y*m+z]; only entries
   first, first+step, ... are written (sharding across ranks) */
int vt_fit_rotate_table(vt_fit_ctx *ctx, const float com[3], int stride, int max_rot, int first, int step,
                        float *table);
/* the reference's sequential loop incl. the skip rule, driven from a table (TclVoltool.C:160-199) */
int vt_rotate_replay(const float *table, int max_rot, float *returncc, int *best_idx);
/* apply one candidate transform to host coordinates with the reference's rounding (do_rotate etc.) */
int vt_pose_apply(float *pos, long natoms, const float com[3], const vt_pose *pose);
int vt_measure_center(const float *pos, const float *weight, long natoms, float com[3]); /* TclVoltool.C:789 */

/* density_rotate (TclVoltool.C:202-279, dead code: caller commented out at :820-891): the map-rotating search.
   Per candidate (x, y, z) < max_rot^3 the synthetic map's frame is moved by -com, rotated about X, Y, Z by
   stride * amt degrees (do_density_rotate :138-149 -> vol_move), its vol_com put back onto com, and
   cc_threaded(target, synth, 0.1) kept if `cc > *returncc`.  synth's frame is restored from true copies on return
   (the dead caller aliases the copies with the live arrays, :874-877; see vt_density.cu).  table (may be NULL)
   receives all max_rot^3 scores in loop order. */
int vt_density_rotate(const vt_map *target, vt_map *synth, int stride, int max_rot, const float com[3], float *returncc,
                      int bestrot[3], float *table);
/* the map half of reset_density (TclVoltool.C:283-300); its atom half (:302-311) is vt_pose_apply */
int vt_density_reset(vt_map *synth, int stride, const int bestrot[3], const float synthcom[3], const float com[3]);
/* rotation x translation grid search (BASELINE cfg2's workload): candidate
   ((ir * nt + ix) * nt + iy) * nt + iz, nt = 2 tsteps + 1, is rotation rot_deg[3 ir ..] about com followed by the
   shift ((ix, iy, iz) - tsteps) * tstep (a vt_pose), scored by calc_cc (TclVoltool.C:73-129); the winner is the
   first maximum in candidate order (`if (cc > best)`, :181; NaN never wins).  cc_all (may be NULL) receives every
   score.  After vt_nccl_init rank r scores candidates r, r + world, ... and one ncclAllGather of the per-rank
   (cc, index) merges them: every rank returns the same winner, cc_all holds NaN for other ranks' candidates. */
int vt_fit_search_grid(vt_fit_ctx *ctx, const float com[3], const float *rot_deg, long nrot, int tsteps, float tstep,
                       float *best_cc, long *best_index, vt_pose *best_pose, float *cc_all);

/* multi-GPU: the pose search is the only sharded path.  rank r evaluates candidates r, r+world, ...
   and the per-stage tables are merged by the caller-supplied allgather (torch.distributed/NCCL in
   voltool_b200.dist).  merge(table, n, rank, world, user) must leave the complete table on every rank. */
typedef int (*vt_table_merge_fn)(float *table, int n, int rank, int world, void *user);
int vt_fit_set_shard(int rank, int world, vt_table_merge_fn merge, void *user);
/* the same merge inside the library: ONE ncclAllGather per rotate() table on device buffers over NVLink (no
   host round trip before the collective).  NCCL is bound at run time (dlopen libnccl.so.2; VOLTOOL_NCCL_LIB
   overrides).  Rank 0 makes the 128-byte unique id, the caller carries it to the other ranks (MPI, a file,
   torch.distributed), every rank calls vt_nccl_init, and vt_fit / vt_fit_sel then shard by themselves: rank r
   evaluates candidates r, r + world, ... (TclVoltool.C:160-176) and all ranks replay the skip rule on the
   identical merged table, so every rank returns the same pose. */
int vt_nccl_unique_id(void *id_out, int len);
int vt_nccl_init(int rank, int world, const void *id, int len);
int vt_nccl_shutdown(void);
/* per-shard best of an exhaustive pose list, on the device: out2 = {cc, (float)(index_base + arg max)}; NaN
   entries never win, ties go to the lowest index (the order of `if (cc > best)`, TclVoltool.C:181) */
int vt_fit_best_pose_device(const float *d_cc, long n, long index_base, float *d_best2);
/* ncclAllGather of `count` floats per rank on the library stream (a device copy when no communicator is up) */
int vt_nccl_allgather(const float *d_send, float *d_recv, long count);
/* evaluator hook: by default the CUDA pose evaluator; host-logic tests inject their own */
typedef int (*vt_table_eval_fn)(void *user, const float *origin_pos, const float com[3], int stride, int max_rot,
                                int first, int step, float *table);
/* fit(): TclVoltool.C:784-916.  pos updated in place with the best pose's coordinates. */
int vt_fit(float *pos, const float *radii, const float *mass, long natoms, const vt_map *target, float resolution,
           vt_fit_result *res);
/* the schedule alone (COM alignment, 3-stage Euler schedule, replay, commit) with an injected
   evaluator; dencom is vol_com(target) supplied by the caller.  No CUDA is touched. */
int vt_fit_schedule(float *pos, const float *mass, long natoms, const float dencom[3], vt_table_eval_fn eval,
                    void *user, vt_fit_result *res);

/* ---------------------------------------------------------------- atom selections
 * The reference's AtomSel loops (`for i in firstsel..lastsel: if sel->on[i]`, moveby TclVoltool.C:52-59; the
 * QuickSurf call sites TclMDFF.C:154-157, 496-499; fit's write-back :911-913) over the molecule's FULL arrays:
 * pos = 3*num_atoms, radii/mass = num_atoms, on = num_atoms flags.  These compact by the flags (ascending atom
 * index), run the compacted entry point above and scatter coordinates back; unselected atoms are untouched. */
int vt_compact_selection(const float *pos, const float *radii, const float *mass, const int *on, long num_atoms,
                         float *pos_out, float *radii_out, float *mass_out, long *index_out, long *nsel);
int vt_sim_sel(const float *pos, const float *radii, const int *on, long num_atoms, int quality, float radscale,
               float gspacing, vt_map **out);
int vt_sim_cc_sel(const float *pos, const float *radii, const int *on, long num_atoms, double resolution,
                  double gspacing_in, const vt_map *target, double threshold, double *cc, vt_map **synth_out);
int vt_fit_sel(float *pos, const float *radii, const float *mass, const int *on, long num_atoms, const vt_map *target,
               float resolution, vt_fit_result *res);
/* what vt_fit_create chose: poses per batch (from the free device memory), workspaces (2 = two-stream
   pipeline), whether the list-based splat is active, and whether the last device-list call had to redo its
   cos/sin preparation on the host (see vt_fit_eval_poses_device).  Any pointer may be NULL. */
int vt_fit_info(const vt_fit_ctx *ctx, int *batch, int *nslots, int *use_lists, int *prep_on_host);

#ifdef __cplusplus
}
#endif
#endif
