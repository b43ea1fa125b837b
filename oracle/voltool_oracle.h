/*
 * voltool_oracle.h -- CPU restatement of voltool's hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Nothing under voltool_b200/ (the product) may include, link or call this.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, as the
 * checker.  Every function cites the reference file:line it restates (paths are relative to
 * the upstream snapshot ryanmcgreevy/voltool).
 *
 * Parity pins:
 *   - everything marked [PINNED] is validated bit-for-bit against the reference's own object
 *     code (oracle/_ref, built from VolumetricData.C / Voltool.C / MDFF.C unchanged) by
 *     tests/test_oracle_vs_ref.py and by the committed fixtures in tests/golden/.
 *   - everything marked [UNPINNED] restates VMD code that is NOT in the snapshot (QuickSurf
 *     splat, GaussianBlur, Matrix4, Measure): "parity unpinned".  The restatement here is the
 *     definition the CUDA path is held to; each choice is documented at the function.
 */
#ifndef VOLTOOL_ORACLE_H
#define VOLTOOL_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* mirrors VolumetricData.h:29-33 */
typedef struct vo_grid {
  double origin[3];
  double xaxis[3];
  double yaxis[3];
  double zaxis[3];
  int xsize, ysize, zsize;
} vo_grid;

/* cached statistics block of VolumetricData (VolumetricData.h:202-210) */
typedef struct vo_cache {
  int minmax_valid, mean_valid, sigma_valid;
  float cmin, cmax, cmean, csigma;
} vo_cache;

long vo_gridsize(const vo_grid *g);

/* ---- Matrix4 [UNPINNED] ---- */
void vo_mat4_identity(float m[16]);
int vo_mat4_inverse(float m[16]);
void vo_mat4_multpoint3d(const float m[16], const float in[3], float out[3]);
void vo_mat4_rot(float m[16], int axis, float angle_rad);

/* ---- geometry / lookup [PINNED] ---- */
void vo_cell_axes(const vo_grid *g, float xax[3], float yax[3], float zax[3]);
void vo_voxel_coord(const vo_grid *g, long i, float *x, float *y, float *z);
void vo_voxel_coord_int(const vo_grid *g, int i, float *x, float *y, float *z);
void vo_grid_inverse(const vo_grid *g, int shiftflag, float inv[16]);
void vo_voxel_coord_from_cartesian(const vo_grid *g, const float car[3], float vox[3], int shiftflag);
float vo_voxel_value_safe(const vo_grid *g, const float *d, int x, int y, int z);
float vo_voxel_value_interpolate(const vo_grid *g, const float *d, float xv, float yv, float zv);
long vo_voxel_index_from_coord(const vo_grid *g, float x, float y, float z);
float vo_value_from_coord(const vo_grid *g, const float *d, float x, float y, float z, int safe);
float vo_value_interp_from_coord(const vo_grid *g, const float *d, float x, float y, float z, int safe);
void vo_init_from_intersection(const vo_grid *A, const vo_grid *B, vo_grid *out);
void vo_init_from_union(const vo_grid *A, const vo_grid *B, vo_grid *out);

/* ---- CC [PINNED] ---- */
/* sums[5] = {sumA, sumB, sumAA, sumBB, sumAB}; returns 0 */
int vo_cc(const vo_grid *A, const float *a, const vo_grid *B, const float *b, double thr,
          double *cc, long *nused, double sums[5], int nthreads);

/* ---- two-map ops [PINNED]; op: 0 add 1 subtract 2 multiply 3 average ---- */
/* *out is malloc()ed; caller frees with vo_free */
void vo_binary(int op, int interp, int use_union, const vo_grid *A, const float *a,
               const vo_grid *B, const float *b, vo_grid *og, float **out);
void vo_free(void *p);

/* ---- unary streams [PINNED] ---- */
void vo_cache_init(vo_cache *c);
void vo_clamp(const vo_grid *g, float *d, vo_cache *c, float lo, float hi);
void vo_scale_by(const vo_grid *g, float *d, vo_cache *c, float ff);
void vo_scalar_add(const vo_grid *g, float *d, vo_cache *c, float ff);
void vo_rescale_range(const vo_grid *g, float *d, vo_cache *c, float lo, float hi);
void vo_sigma_scale(const vo_grid *g, float *d, vo_cache *c);
void vo_binmask(const vo_grid *g, float *d, vo_cache *c, float thr);
void vo_mdff_potential(const vo_grid *g, float *d, vo_cache *c, double thr);

/* ---- stats [PINNED modulo minmax stubs] ---- */
void vo_datarange(const vo_grid *g, const float *d, vo_cache *c, float *mn, float *mx);
float vo_mean(const vo_grid *g, const float *d, vo_cache *c);
float vo_sigma(const vo_grid *g, const float *d, vo_cache *c);
void vo_minmax_1fv(const float *f, long n, float *mn, float *mx);
void vo_minmaxmean_1fv(const float *f, long n, float *mn, float *mx, float *mean);
void vo_vol_com(const vo_grid *g, const float *d, float com[3]);
void vo_vol_moveto(vo_grid *g, const float com[3], const float pos[3]);
void vo_vol_move(vo_grid *g, const float mat[16]);
void vo_histogram(const vo_grid *g, const float *d, vo_cache *c, int nbins, long *bins, float *midpts);

/* ---- resizing ops [PINNED] (data reallocated with malloc; old buffer freed by caller) ---- */
void vo_pad(vo_grid *g, const float *d, float **out, int pxm, int pxp, int pym, int pyp, int pzm, int pzp);
void vo_crop(vo_grid *g, const float *d, float **out, double minx, double miny, double minz,
             double maxx, double maxy, double maxz);
void vo_downsample(vo_grid *g, const float *d, float **out);
void vo_supersample(vo_grid *g, const float *d, float **out);

/* ---- Gaussian blur [UNPINNED] ---- */
int vo_blur_kernel(double sigma, float *w, int maxw); /* returns radius R; fills w[0..2R] */
void vo_gaussian_blur(const vo_grid *g, const float *d, float **out, double sigma);

/* ---- QuickSurf density map [UNPINNED] ---- */
float vo_gausslim_for_quality(int quality);
void vo_sim_grid(const float *pos, const float *radii, long natoms, float radscale, float gspacing,
                 vo_grid *g, float origin_f[3]);
void vo_sim(const float *pos, const float *radii, long natoms, int quality, float radscale,
            float gspacing, vo_grid *g, float **out, int nthreads);
/* policy of mdff_sim / mdff_cc / calc_cc: TclMDFF.C:125-135, TclVoltool.C:81-88 */
void vo_sim_policy(double resolution, double gspacing_in, float *radscale, double *gspacing, int *quality);

/* ---- rigid transforms + fit driver [UNPINNED pieces: Matrix4/measure_*] ---- */
void vo_moveby(float *pos, long natoms, const float v[3]);
void vo_apply_mat(float *pos, long natoms, const float m[16]);
void vo_measure_center(const float *pos, const float *w, long natoms, float com[3]);
/* one candidate of rotate(): TclVoltool.C:165-176 */
void vo_pose_transform(float *pos, long natoms, const float com[3], int stride, int x, int y, int z);
void vo_pose_general(float *pos, long natoms, const float com[3], const float rot_deg[3], const float shift[3]);
double vo_calc_cc(const float *pos, const float *radii, long natoms, const vo_grid *T, const float *t,
                  float resolution, int nthreads);

typedef double (*vo_cc_fn)(const vo_grid *A, const float *a, const vo_grid *B, const float *b,
                           double thr, void *ctx);
/* plug the reference object code's cc_threaded in here (oracle/_ref) */
void vo_set_cc_backend(vo_cc_fn fn, void *ctx);

/* evaluates all max_rot^3 candidates of one rotate() call, no skip rule: cc_table[x*m*m+y*m+z] */
void vo_rotate_table(const float *origin_pos, const float *radii, long natoms, const float com[3],
                     int stride, int max_rot, const vo_grid *T, const float *t, float resolution,
                     float *cc_table, int pose_stride, int pose_first, int nthreads);
/* the reference's sequential loop incl. skip rule, driven by a table (TclVoltool.C:160-199).
   returns number of candidates actually visited; best index in *best_idx (-1 = no update) */
int vo_rotate_replay(const float *cc_table, int max_rot, float *returncc, int *best_idx);

/* cc extras (TclMDFF.C:180-184, 555-594): diff map on the intersection grid + 8x8x8-tile spatial CC map
   [RESTATEMENT, UNPINNED: definition in voltool_oracle.c].  *diff / *spatial are malloc()ed (vo_free). */
void vo_spatial_grid(const vo_grid *G, vo_grid *sg);
int vo_cc_maps(const vo_grid *A, const float *a, const vo_grid *B, const float *b, double thr, double *cc,
               vo_grid *dg, float **diff, vo_grid *sg, float **spatial);
/* density_rotate / reset_density (TclVoltool.C:202-311, dead code): the map-rotating search.  S's frame is
   restored on return; cc_table (may be NULL) receives all max_rot^3 scores; returns the candidate count. */
int vo_density_rotate(int stride, int max_rot, const float com[3], float *returncc, int bestrot[3], const vo_grid *A,
                      const float *a, vo_grid *S, const float *s, float *cc_table, int nthreads);
void vo_density_reset(int stride, const int bestrot[3], vo_grid *S, const float *s, const float synthcom[3],
                      const float com[3]);

typedef struct vo_fit_result {
  float best_cc;
  float com[3];
  float dencom[3];
  int n_evaluated;       /* candidates the reference's loop visits */
  int stage_best[5];     /* winning index per rotate() call, -1 if none */
} vo_fit_result;
/* full fit(): TclVoltool.C:784-916.  pos updated in place.  If direct!=0 runs the literal
   sequential loop (calc_cc per visited pose); else table + replay. */
void vo_fit(float *pos, const float *radii, const float *mass, long natoms, const vo_grid *T,
            const float *t, float resolution, vo_fit_result *res, int direct, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
