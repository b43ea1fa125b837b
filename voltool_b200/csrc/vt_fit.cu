// vt_fit.cu -- C ABI of sim / calc_cc / fit over the batched pose pipeline (vt_sim.cu).
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "vt_pose.cuh"

using namespace vt;

namespace vt {
int nccl_merge_table(const float *d_slice, float *table, int n);   // vt_dist.cu
}

// Two pose-batch workspaces: while batch i splats and scores on the library stream, batch i+1 is
// transformed, binned and listed on a second stream (those kernels are small and latency-bound, so
// they fit into the gaps of the two big ones).
struct vt_fit_ctx {
  AtomSet as;
  PoseBatch pb[2];
  const vt_map *target = nullptr;
  float resolution = 0.f;
  double threshold = 0.1;  // calc_cc hard-codes thresholddensity = 0.1 (TclVoltool.C:77)
  int B = 0;
  int nslots = 2;              // 1: no overlap (VOLTOOL_FIT_OVERLAP=0)
  cudaStream_t aux = nullptr;
  cudaEvent_t ev_start = nullptr, ev_front[2] = {nullptr, nullptr}, ev_back[2] = {nullptr, nullptr};
  PrepPose *h_prep = nullptr;  // pinned staging of the whole call
  long h_prep_cap = 0;
  PrepPose *d_prep = nullptr;  // prepared poses of the whole call
  long d_prep_cap = 0;
  float *d_cc = nullptr;       // results of the current call
  int *d_flag = nullptr;       // ambiguous-cast counter of prep_from_pose_kernel
  int last_prep_on_host = 0;   // the last device-list call fell back to the host preparation
  long cc_cap = 0;
};

namespace {

SimParams make_params(int quality, float radscale, float gspacing) {
  SimParams p;
  p.radscale = radscale;
  p.gs = gspacing;
  p.invgs = 1.0f / gspacing;
  p.gausslim = gausslim_for_quality(quality);
  p.quality = quality;
  return p;
}

// single structure, identity pose -> resident map (QuickSurf::calc_density_map)
int sim_to_map(const float *pos, const float *radii, long natoms, int quality, float radscale, float gspacing,
               bool want_minmax, vt_map **out) {
  VT_REQUIRE_CTX();
  if (!pos || !radii || natoms <= 0 || !(gspacing > 0.f)) { set_error("vt_sim: bad arguments"); return 1; }
  AtomSet as;
  const bool verbose = getenv("VOLTOOL_VERBOSE") != nullptr;
  const auto ts0 = std::chrono::steady_clock::now();
  int rc = atomset_create(pos, radii, natoms, make_params(quality, radscale, gspacing), as);
  if (rc) return rc;
  const auto ts1 = std::chrono::steady_clock::now();
  // exact grid size on the host (same float expressions as pose_setup_kernel) to size the slot
  float mn[3] = {pos[0], pos[1], pos[2]}, mx[3] = {pos[0], pos[1], pos[2]};
  for (long i = 0; i < natoms; ++i)
    for (int k = 0; k < 3; ++k) {
      const float v = pos[3 * i + k];
      if (v < mn[k]) mn[k] = v;
      if (v > mx[k]) mx[k] = v;
    }
  int nv[3];
  long cells = 1;
  int maxn = 0;
  for (int k = 0; k < 3; ++k) {
    const float lo = mn[k] - as.pad, hi = mx[k] + as.pad;
    const float ax = hi - lo;
    nv[k] = (int)std::ceil((double)(ax / gspacing));
    if (nv[k] < 2) { atomset_destroy(as); set_error("vt_sim: degenerate grid"); return 1; }
    cells *= nv[k];
    if (nv[k] > maxn) maxn = nv[k];
  }
  PoseBatch pb;
  pb.B = 1;
  pb.cap_n = maxn;
  pb.cap_grid = cells;
  Ctx &c = ctx();
  rc = dev_alloc((void **)&pb.d_tpos, sizeof(float4) * (size_t)as.n);
  if (!rc) rc = dev_alloc((void **)&pb.d_tgroup, sizeof(float4) * (size_t)as.ngroups);
  if (!rc) rc = dev_alloc((void **)&pb.d_bbox, sizeof(unsigned int) * 6);
  if (!rc) rc = dev_alloc((void **)&pb.d_geom, sizeof(PoseGeom));
  if (!rc) rc = dev_alloc((void **)&pb.d_totals, sizeof(int) * 8);
  if (!rc) rc = dev_alloc((void **)&pb.d_grid, sizeof(float) * (size_t)cells);
  pb.tiles_cap = ((nv[0] + kTX - 1) / kTX) * ((nv[1] + kTY - 1) / kTY) * ((nv[2] + kTZ - 1) / kTZ);
  if (!rc) rc = dev_alloc((void **)&pb.d_bitmap, sizeof(unsigned int) * (size_t)pb.tiles_cap * ((as.ngroups + 31) / 32));
  const float com0[3] = {0.f, 0.f, 0.f};
  if (!rc) rc = batch_prepare(as, pb, 1, nullptr, com0, 1, nullptr);
  if (!rc) rc = batch_estimate_lists(as, pb, nv);
  if (!rc) rc = batch_splat(as, pb, 1);
  // one round trip for the overflow flag and the geometry
  PoseGeom G;
  int totals[4] = {0, 0, 0, 0};
  auto readback = [&]() -> int {
    cudaError_t e = cudaMemcpyAsync(&G, pb.d_geom, sizeof(G), cudaMemcpyDeviceToHost, c.stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(totals, pb.d_totals, sizeof(totals), cudaMemcpyDeviceToHost, c.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c.stream);
    return e == cudaSuccess ? 0 : cuda_fail(e, "sim readback", __FILE__, __LINE__);
  };
  const auto ts2 = std::chrono::steady_clock::now();
  if (!rc) rc = readback();
  if (verbose)
    fprintf(stderr, "voltool_gpu: sim_to_map: atom set %.3f ms, allocate + enqueue %.3f ms, wait %.3f ms\n",
            std::chrono::duration<double, std::milli>(ts1 - ts0).count(), std::chrono::duration<double, std::milli>(ts2 - ts1).count(),
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - ts2).count());
  if (!rc && pb.use_lists && totals[2]) {   // a tile list did not fit: redo with the cooperative kernel
    pb.use_lists = false;
    rc = batch_splat(as, pb, 1);
    if (!rc) rc = readback();
  }
  if (!rc && (G.nv[0] != nv[0] || G.nv[1] != nv[1] || G.nv[2] != nv[2])) {
    set_error("vt_sim: host/device grid mismatch (%d %d %d vs %d %d %d)", nv[0], nv[1], nv[2], G.nv[0], G.nv[1], G.nv[2]);
    rc = 3;
  }
  if (!rc) {
    vt_map *m = new vt_map();
    memset(&m->grid, 0, sizeof(m->grid));
    for (int k = 0; k < 3; ++k) m->grid.origin[k] = (double)G.org[k];
    m->grid.xaxis[0] = (double)((float)(nv[0] - 1) * gspacing);
    m->grid.yaxis[1] = (double)((float)(nv[1] - 1) * gspacing);
    m->grid.zaxis[2] = (double)((float)(nv[2] - 1) * gspacing);
    m->grid.xsize = nv[0]; m->grid.ysize = nv[1]; m->grid.zsize = nv[2];
    m->n = cells;
    m->d = pb.d_grid;  // adopt the slot
    pb.d_grid = nullptr;
    memset(&m->cache, 0, sizeof(m->cache));
    if (want_minmax) {
      rc = k_minmax(m->d, m->n, &m->cache.cmin, &m->cache.cmax);
      m->cache.minmax_valid = 1;
    }
    *out = m;
  }
  batch_destroy(pb);
  atomset_destroy(as);
  return rc;
}

// (float) of a double whose last few ulps are uncertain: true when a float rounding boundary lies within ~9
// double ulps of d, i.e. when the cast of a differently rounded evaluation of the same function could differ
__device__ __forceinline__ bool float_cast_ambiguous(double d) {
  const float f = (float)d;
  const double fd = (double)f;
  const double up = (double)nextafterf(f, INFINITY), dn = (double)nextafterf(f, -INFINITY);
  const double tol = fabs(d) * 2.0e-15 + 1e-300;
  return fabs(d - 0.5 * (fd + up)) <= tol || fabs(d - 0.5 * (fd + dn)) <= tol;
}

// device twin of euler_cs (vt_host.cpp; do_rotate TclVoltool.C:131-136).  CUDA's double cos/sin are good to
// 2 ulp, glibc's to <1 ulp: after the cast to float the two agree unless the double lies within a few ulps of a
// float rounding boundary (about 3 in 10^8 evaluations).  Those poses raise *ambiguous and the caller redoes
// the preparation on the host, so a device pose list always scores exactly like the host list.
__global__ void prep_from_pose_kernel(const vt_pose *__restrict__ poses, int n, PrepPose *__restrict__ out,
                                      int *__restrict__ ambiguous) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const vt_pose q = poses[i];
  PrepPose r;
  bool amb = false;
  for (int k = 0; k < 3; ++k) {
    // DEGTORAD in double, rotate_axis takes a float angle, cos/sin evaluated in double
    const double amount = (double)q.rot_deg[k] * 3.14159265358979323846 / 180.0;
    const float a = (float)amount;
    const double cd = cos((double)a), sd = sin((double)a);
    amb = amb || float_cast_ambiguous(cd) || float_cast_ambiguous(sd);
    r.c[k] = (float)cd;
    r.s[k] = (float)sd;
    r.shift[k] = q.shift[k];
  }
  out[i] = r;
  if (amb) atomicAdd(ambiguous, 1);
}

int ensure_cc_buffer(vt_fit_ctx *f, long n) {
  if (n <= f->cc_cap) return 0;
  dev_free(f->d_cc);
  f->d_cc = nullptr;
  if (dev_alloc((void **)&f->d_cc, sizeof(float) * (size_t)n)) return 2;
  f->cc_cap = n;
  return 0;
}

int ensure_prep_buffers(vt_fit_ctx *f, long n, bool host) {
  if (n > f->d_prep_cap) {
    dev_free(f->d_prep);
    f->d_prep = nullptr;
    if (dev_alloc((void **)&f->d_prep, sizeof(PrepPose) * (size_t)n)) return 2;
    f->d_prep_cap = n;
  }
  if (host && n > f->h_prep_cap) {
    if (f->h_prep) cudaFreeHost(f->h_prep);
    f->h_prep = nullptr;
    cudaError_t e = cudaMallocHost((void **)&f->h_prep, sizeof(PrepPose) * (size_t)n);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMallocHost", __FILE__, __LINE__);
    f->h_prep_cap = n;
  }
  return 0;
}

struct StreamScope {   // the launch wrappers use ctx().stream: point it at the second stream for a while
  cudaStream_t saved;
  explicit StreamScope(cudaStream_t s) : saved(ctx().stream) { ctx().stream = s; }
  ~StreamScope() { ctx().stream = saved; }
};

// front half of one batch: transform + geometry + tables, tile bitmaps, boxes, per-tile lists
int batch_front(vt_fit_ctx *f, PoseBatch &pb, const float com[3], const PrepPose *d_prep, int n) {
  int rc = batch_prepare(f->as, pb, n, d_prep, com, 0, &f->target->grid);
  if (!rc) rc = batch_splat_front(f->as, pb, n);
  return rc;
}

// back half: accumulate the density grids, score them against the target
int batch_back(vt_fit_ctx *f, PoseBatch &pb, int n, float *d_cc) {
  int rc = batch_splat_back(f->as, pb, n);
  if (!rc) rc = batch_cc(pb, n, f->target, f->threshold, d_cc, nullptr);
  return rc;
}

// run the pipeline for poses already prepared on the device (d_prep[nposes]); everything is
// enqueued without a host round trip
int eval_prepared(vt_fit_ctx *f, const float com[3], const PrepPose *d_prep, long nposes, float *d_cc) {
  Ctx &c = ctx();
  const long nb = (nposes + f->B - 1) / f->B;
  auto count = [&](long i) { return (int)std::min<long>(f->B, nposes - i * f->B); };
  if (f->nslots < 2 || nb < 2) {
    for (long i = 0; i < nb; ++i) {
      int rc = batch_front(f, f->pb[0], com, d_prep + i * f->B, count(i));
      if (!rc) rc = batch_back(f, f->pb[0], count(i), d_cc + i * f->B);
      if (rc) return rc;
    }
    return 0;
  }
  VT_CUDA(cudaEventRecord(f->ev_start, c.stream));       // the second stream starts after everything queued so far
  VT_CUDA(cudaStreamWaitEvent(f->aux, f->ev_start, 0));
  auto front = [&](long i) -> int {
    const int slot = (int)(i & 1);
    if (i >= 2) VT_CUDA(cudaStreamWaitEvent(f->aux, f->ev_back[slot], 0));   // the slot's previous batch has been scored
    int rc;
    {
      StreamScope scope(f->aux);
      rc = batch_front(f, f->pb[slot], com, d_prep + i * f->B, count(i));
    }
    if (rc) return rc;
    VT_CUDA(cudaEventRecord(f->ev_front[slot], f->aux));
    return 0;
  };
  int rc = front(0);
  for (long i = 0; i < nb && !rc; ++i) {
    const int slot = (int)(i & 1);
    if (i + 1 < nb) rc = front(i + 1);
    if (rc) break;
    VT_CUDA(cudaStreamWaitEvent(c.stream, f->ev_front[slot], 0));
    rc = batch_back(f, f->pb[slot], count(i), d_cc + i * f->B);
    if (rc) break;
    VT_CUDA(cudaEventRecord(f->ev_back[slot], c.stream));
  }
  return rc;
}

__global__ void scatter_cc_kernel(const float *__restrict__ src, int n, float *__restrict__ dst, long first, int stride, int offset) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < n) dst[(first + j) * stride + offset] = src[j];
}

int reset_overflow(vt_fit_ctx *f) {
  for (int k = 0; k < f->nslots; ++k)
    if (f->pb[k].use_lists) VT_CUDA(cudaMemsetAsync(f->pb[k].d_totals + 2, 0, sizeof(int), ctx().stream));
  return 0;
}

// 1 in *flag if a tile list of either workspace overflowed since reset_overflow (synchronises)
int any_overflow(vt_fit_ctx *f, int *flag) {
  *flag = 0;
  for (int k = 0; k < f->nslots; ++k)
    if (f->pb[k].use_lists) {
      int ovf = 0;
      int rc = batch_lists_overflowed(f->pb[k], &ovf);
      if (rc) return rc;
      *flag |= ovf;
    }
  return 0;
}

int cuda_table_eval(void *user, const float *origin_pos, const float com[3], int stride, int max_rot, int first,
                    int step, float *table) {
  vt_fit_ctx *f = (vt_fit_ctx *)user;
  const auto t0 = std::chrono::steady_clock::now();
  int rc = vt_fit_set_origin(f, origin_pos);
  if (rc) return rc;
  const auto t1 = std::chrono::steady_clock::now();
  rc = vt_fit_rotate_table(f, com, stride, max_rot, first, step, table);
  if (rc) return rc;
  if (getenv("VOLTOOL_VERBOSE")) {
    const auto t2 = std::chrono::steady_clock::now();
    fprintf(stderr, "voltool_gpu: rotate table stride %d, %d^3 candidates (1 of %d here): set_origin %.2f ms, evaluate %.2f ms\n", stride,
            max_rot, step, std::chrono::duration<double, std::milli>(t1 - t0).count(),
            std::chrono::duration<double, std::milli>(t2 - t1).count());
  }
  // sharded over NCCL: this rank's slice is still in f->d_cc in evaluation order (candidates first, first + step,
  // ...): one allgather of the device slices, and every rank holds the whole table
  if (step > 1 && nccl_active() && !shard_config().merge) rc = nccl_merge_table(f->d_cc, table, max_rot * max_rot * max_rot);
  return rc;
}

}  // namespace

// Rotation x translation candidates with ONE splat per rotation (vt_fit_search_grid): every rotation of the list is
// transformed and splatted once, and its grid is then re-aimed at each of the ns translated frames and scored
// (batch_retarget + the batched cc kernel).  cc_out[ir * ns + s], host.  The translated candidates differ from a
// separately splatted pose only by the rounding of `origin + shift` against `coordinates + shift` (an ulp of the
// coordinates): CC agrees to ~1e-7, tested against the per-pose path and against the oracle.
int vt::fit_eval_rot_shifts(vt_fit_ctx *f, const float com[3], const float *rot_deg, long nrot, const float *shifts, int ns,
                            float *cc_out) {
  VT_REQUIRE_CTX();
  if (nrot <= 0 || ns <= 0) return 0;
  Ctx &c = ctx();
  int rc = ensure_cc_buffer(f, nrot * ns + f->B);
  if (!rc) rc = ensure_prep_buffers(f, nrot, true);
  if (rc) return rc;
  VT_CUDA(cudaStreamSynchronize(c.stream));
  for (long i = 0; i < nrot; ++i)
    for (int k = 0; k < 3; ++k) {
      euler_cs((double)rot_deg[3 * i + k], f->h_prep[i].c[k], f->h_prep[i].s[k]);
      f->h_prep[i].shift[k] = 0.f;
    }
  VT_CUDA(cudaMemcpyAsync(f->d_prep, f->h_prep, sizeof(PrepPose) * (size_t)nrot, cudaMemcpyHostToDevice, c.stream));
  float *d_tmp = f->d_cc + nrot * ns;   // one batch of scores before they are scattered to [rotation][shift]
  for (int attempt = 0; attempt < 2; ++attempt) {
    rc = reset_overflow(f);
    if (rc) return rc;
    PoseBatch &pb = f->pb[0];
    for (long i0 = 0; i0 < nrot && !rc; i0 += f->B) {
      const int n = (int)std::min<long>(f->B, nrot - i0);
      rc = batch_front(f, pb, com, f->d_prep + i0, n);
      if (!rc) rc = batch_splat_back(f->as, pb, n);
      for (int s = 0; s < ns && !rc; ++s) {
        rc = batch_retarget(f->as, pb, n, shifts + 3 * s, s == 0, &f->target->grid);
        if (!rc) rc = batch_cc(pb, n, f->target, f->threshold, d_tmp, nullptr);
        if (!rc) {
          scatter_cc_kernel<<<(n + 127) / 128, 128, 0, c.stream>>>(d_tmp, n, f->d_cc, i0, ns, s);
          VT_LAUNCHED();
        }
      }
    }
    if (rc) return rc;
    VT_CUDA(cudaMemcpyAsync(cc_out, f->d_cc, sizeof(float) * (size_t)(nrot * ns), cudaMemcpyDeviceToHost, c.stream));
    VT_CUDA(cudaStreamSynchronize(c.stream));
    int ovf = 0;
    rc = any_overflow(f, &ovf);
    if (rc) return rc;
    if (!ovf) return 0;
    for (int k = 0; k < 2; ++k) f->pb[k].use_lists = false;   // a tile list was too small: redo with the cooperative kernel
  }
  return 0;
}

extern "C" {

int vt_sim(const float *pos, const float *radii, long natoms, int quality, float radscale, float gspacing,
           vt_map **out) {
  return sim_to_map(pos, radii, natoms, quality, radscale, gspacing, true, out);
}

int vt_calc_cc(const float *pos, const float *radii, long natoms, const vt_map *target, float resolution, double *cc) {
  VT_REQUIRE_CTX();
  float radscale;
  double gspacing;
  int quality;
  // calc_cc: radscale = .2*res; gspacing = 1.5*radscale; threshold 0.1 (TclVoltool.C:77-88)
  radscale = (float)(.2 * resolution);
  gspacing = 1.5 * radscale;
  quality = resolution >= 9 ? 0 : 3;
  vt_map *sim = nullptr;
  int rc = sim_to_map(pos, radii, natoms, quality, radscale, (float)gspacing, false, &sim);
  if (rc) return rc;
  double c = 0.0;
  rc = vt_cc(sim, target, 0.1, &c, nullptr, nullptr);
  vt_map_destroy(sim);
  if (rc) return rc;
  float ret = 0;
  ret += c;  // float return_cc += cc (TclVoltool.C:75,126)
  *cc = ret;
  return 0;
}

int vt_sim_cc(const float *pos, const float *radii, long natoms, double resolution, double gspacing_in,
              const vt_map *target, double threshold, double *cc, vt_map **synth_out) {
  VT_REQUIRE_CTX();
  float radscale;
  double gspacing;
  int quality;
  sim_policy(resolution, gspacing_in, radscale, gspacing, quality);
  vt_map *sim = nullptr;
  const auto t0 = std::chrono::steady_clock::now();
  int rc = sim_to_map(pos, radii, natoms, quality, radscale, (float)gspacing, synth_out != nullptr, &sim);
  if (rc) return rc;
  const auto t1 = std::chrono::steady_clock::now();
  rc = vt_cc(sim, target, threshold, cc, nullptr, nullptr);
  if (getenv("VOLTOOL_VERBOSE"))
    fprintf(stderr, "voltool_gpu: vt_sim_cc: sim %.3f ms, cc %.3f ms\n", std::chrono::duration<double, std::milli>(t1 - t0).count(),
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count());
  if (synth_out && !rc) *synth_out = sim;
  else vt_map_destroy(sim);
  return rc;
}

// ------------------------------------------------------------------------------------------ fit context
int vt_fit_create(const float *pos, const float *radii, long natoms, const vt_map *target, float resolution,
                  vt_fit_ctx **out) {
  VT_REQUIRE_CTX();
  if (!pos || !radii || !target || natoms <= 0) { set_error("vt_fit_create: bad arguments"); return 1; }
  {
    // the batched cc kernel is the separable one: a skewed target cell would make every pose's CC undefined,
    // so say so here (the reference's own two-map code assumes orthogonal axis-aligned cells, Voltool.C:55-56)
    float tinv[16];
    geom::grid_inverse(target->grid, false, tinv);
    const vt_grid &tg = target->grid;
    const bool aligned = tg.xaxis[1] == 0 && tg.xaxis[2] == 0 && tg.yaxis[0] == 0 && tg.yaxis[2] == 0 && tg.zaxis[0] == 0 &&
                         tg.zaxis[1] == 0;
    if (!aligned || !geom::diagonal_inverse(tinv)) {
      set_error("vt_fit_create: the target map must have an orthogonal, axis-aligned cell");
      return 1;
    }
  }
  if ((long)target->grid.xsize * target->grid.ysize >= (1L << 30)) {
    set_error("vt_fit_create: target planes of 2^30 voxels or more are not supported");
    return 1;
  }
  vt_fit_ctx *f = new vt_fit_ctx();
  f->target = target;
  f->resolution = resolution;
  const float radscale = (float)(.2 * resolution);
  const double gspacing = 1.5 * radscale;
  const int quality = resolution >= 9 ? 0 : 3;
  int rc = atomset_create(pos, radii, natoms, make_params(quality, radscale, (float)gspacing), f->as);
  if (rc) { delete f; return rc; }
  const char *ov = getenv("VOLTOOL_FIT_OVERLAP");
  f->nslots = (ov && atoi(ov) == 0) ? 1 : 2;
  // poses per batch: 256 while a pose's density grid stays small (cfg2: 2.6 MB), 128 for big structures
  // (measured on cfg2: 128 -> 256 poses: +2.8 %), and never more than fits: the reference evaluates poses one
  // at a time, so any structure it can fit must fit here too.  The per-pose workspace is estimated from the
  // capacities batch_create will use, B is cut to 80 % of the free device memory, and an allocation that
  // still fails halves B (and finally drops the second workspace).  VOLTOOL_FIT_BATCH only overrides the start.
  {
    const double ext = 2.0 * f->as.bound_radius + 2.0 * f->as.pad + 4.0 * gspacing;
    const double cap_n = std::ceil(ext / gspacing) + 2.0;
    const double grid_bytes = cap_n * cap_n * cap_n * 4.0;
    const char *env = getenv("VOLTOOL_FIT_BATCH");
    f->B = env ? atoi(env) : (grid_bytes <= 8.0e6 ? 256 : 128);
    if (f->B < 1) f->B = 1;
    if (f->B > 1024) f->B = 1024;
    const double tiles = std::ceil(cap_n / kTX) * std::ceil(cap_n / kTY) * std::ceil(cap_n / kTZ);
    const double nwords = (f->as.ngroups + 31) / 32;
    const double h = std::ceil((double)f->as.prm.gausslim * f->as.maxrad * radscale / gspacing) + 1.0;
    const double reach = (kTX + 2 * h) * (kTY + 2 * h) * (kTZ + 2 * h);
    const double list_cap = 1.6 * 2.0 * (double)f->as.n * reach / (cap_n * cap_n * cap_n) + 256.0;   // per tile, as calibrated
    const double per_pose = grid_bytes + 28.0 * f->as.n + 28.0 * f->as.ngroups + tiles * (4.0 * nwords + 4.0 + 4.0 * list_cap) +
                            6.0 * sizeof(geom::AxisEntry) * (cap_n * 1.6 + 8) + 48.0 * (cap_n * cap_n * cap_n / (32.0 * kRowsFit * kZChunkFit) + 64);
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && per_pose > 0) {
      // memory the pool already holds but has handed back to us is reusable as well
      cudaMemPool_t pool = nullptr;
      unsigned long long reserved = 0, used = 0;
      if (cudaDeviceGetDefaultMemPool(&pool, ctx().device) == cudaSuccess &&
          cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved) == cudaSuccess &&
          cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used) == cudaSuccess && reserved > used)
        free_b += (size_t)(reserved - used);
      const double fit_b = 0.8 * (double)free_b / (f->nslots * per_pose);
      if (fit_b < f->B) f->B = fit_b < 1.0 ? 1 : (int)fit_b;
    }
  }
  while (true) {
    rc = 0;
    for (int k = 0; k < f->nslots && !rc; ++k) {
      rc = batch_create(f->as, &target->grid, f->B, f->pb[k]);
      if (!rc) {
        const float com0[3] = {0.f, 0.f, 0.f};
        rc = batch_prepare(f->as, f->pb[k], 1, nullptr, com0, 1, &target->grid);
        if (!rc) rc = batch_calibrate_lists(f->as, f->pb[k]);
        if (f->nslots > 1) f->pb[k].splat_ctas = 5;   // leave room for the second stream (vt_splat2.cu)
      }
    }
    if (rc != 2 || last_cuda_error() != cudaErrorMemoryAllocation) break;
    for (int k = 0; k < 2; ++k) batch_destroy(f->pb[k]);
    cudaGetLastError();
    if (f->B > 1) f->B = (f->B + 1) / 2;
    else if (f->nslots > 1) f->nslots = 1;
    else break;
  }
  if (getenv("VOLTOOL_VERBOSE"))
    fprintf(stderr, "voltool_gpu: fit workspace: %d poses per batch x %d slot(s), lists %s\n", f->B, f->nslots,
            f->pb[0].use_lists ? "on" : "off");
  if (!rc && f->nslots > 1) {
    cudaError_t e = cudaStreamCreateWithFlags(&f->aux, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&f->ev_start, cudaEventDisableTiming);
    for (int k = 0; k < 2 && e == cudaSuccess; ++k) {
      e = cudaEventCreateWithFlags(&f->ev_front[k], cudaEventDisableTiming);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&f->ev_back[k], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) rc = cuda_fail(e, "fit streams", __FILE__, __LINE__);
  }
  if (rc) { vt_fit_destroy(f); return rc; }
  *out = f;
  return 0;
}

int vt_fit_destroy(vt_fit_ctx *f) {
  if (!f) return 0;
  if (f->aux) cudaStreamSynchronize(f->aux);
  cudaStreamSynchronize(ctx().stream);
  for (int k = 0; k < 2; ++k) {
    batch_destroy(f->pb[k]);
    if (f->ev_front[k]) cudaEventDestroy(f->ev_front[k]);
    if (f->ev_back[k]) cudaEventDestroy(f->ev_back[k]);
  }
  if (f->ev_start) cudaEventDestroy(f->ev_start);
  if (f->aux) cudaStreamDestroy(f->aux);
  atomset_destroy(f->as);
  dev_free(f->d_cc);
  dev_free(f->d_prep);
  dev_free(f->d_flag);
  if (f->h_prep) cudaFreeHost(f->h_prep);
  delete f;
  return 0;
}

int vt_fit_info(const vt_fit_ctx *f, int *batch, int *nslots, int *use_lists, int *prep_on_host) {
  if (!f) { set_error("vt_fit_info: null context"); return 1; }
  if (batch) *batch = f->B;
  if (nslots) *nslots = f->nslots;
  if (use_lists) *use_lists = f->pb[0].use_lists ? 1 : 0;
  if (prep_on_host) *prep_on_host = f->last_prep_on_host;
  return 0;
}

int vt_fit_set_origin(vt_fit_ctx *f, const float *pos) {
  VT_REQUIRE_CTX();
  return atomset_set_positions(f->as, pos);
}

int vt_fit_eval_poses(vt_fit_ctx *f, const float com[3], const vt_pose *poses, long nposes, float *cc_out) {
  VT_REQUIRE_CTX();
  if (nposes <= 0) return 0;
  int rc = ensure_cc_buffer(f, nposes);
  if (!rc) rc = ensure_prep_buffers(f, nposes, true);
  if (rc) return rc;
  Ctx &c = ctx();
  VT_CUDA(cudaStreamSynchronize(c.stream));   // the pinned staging buffer may still feed an earlier call's upload
  for (long i = 0; i < nposes; ++i) {
    const vt_pose &q = poses[i];
    for (int k = 0; k < 3; ++k) {
      euler_cs((double)q.rot_deg[k], f->h_prep[i].c[k], f->h_prep[i].s[k]);
      f->h_prep[i].shift[k] = q.shift[k];
    }
  }
  VT_CUDA(cudaMemcpyAsync(f->d_prep, f->h_prep, sizeof(PrepPose) * (size_t)nposes, cudaMemcpyHostToDevice, c.stream));
  rc = reset_overflow(f);
  if (!rc) rc = eval_prepared(f, com, f->d_prep, nposes, f->d_cc);
  if (rc) return rc;
  VT_CUDA(cudaMemcpyAsync(cc_out, f->d_cc, sizeof(float) * (size_t)nposes, cudaMemcpyDeviceToHost, c.stream));
  VT_CUDA(cudaStreamSynchronize(c.stream));
  int ovf = 0;
  rc = any_overflow(f, &ovf);
  if (rc) return rc;
  if (ovf) {  // a tile list was too small: nothing is truncated silently, redo with the cooperative kernel
    for (int k = 0; k < 2; ++k) f->pb[k].use_lists = false;
    return vt_fit_eval_poses(f, com, poses, nposes, cc_out);
  }
  return 0;
}

int vt_fit_eval_poses_device(vt_fit_ctx *f, const float com[3], const vt_pose *d_poses, long nposes, float *d_cc) {
  VT_REQUIRE_CTX();
  if (nposes <= 0) return 0;
  Ctx &c = ctx();
  bool host_prep = getenv("VOLTOOL_PREP_HOST") != nullptr;   // tests: force the host preparation
  for (int attempt = 0; attempt < 4; ++attempt) {
    int rc = ensure_prep_buffers(f, nposes, host_prep);
    if (!rc && !f->d_flag) rc = dev_alloc((void **)&f->d_flag, sizeof(int));
    if (!rc) rc = reset_overflow(f);
    if (rc) return rc;
    VT_CUDA(cudaMemsetAsync(f->d_flag, 0, sizeof(int), c.stream));
    if (!host_prep) {
      prep_from_pose_kernel<<<(unsigned int)((nposes + 127) / 128), 128, 0, c.stream>>>(d_poses, (int)nposes, f->d_prep,
                                                                                         f->d_flag);
      VT_LAUNCHED();
    } else {   // cos/sin with the host libm, exactly as vt_fit_eval_poses forms them
      std::vector<vt_pose> hp((size_t)nposes);
      VT_CUDA(cudaMemcpyAsync(hp.data(), d_poses, sizeof(vt_pose) * (size_t)nposes, cudaMemcpyDeviceToHost, c.stream));
      VT_CUDA(cudaStreamSynchronize(c.stream));
      for (long i = 0; i < nposes; ++i)
        for (int k = 0; k < 3; ++k) {
          euler_cs((double)hp[i].rot_deg[k], f->h_prep[i].c[k], f->h_prep[i].s[k]);
          f->h_prep[i].shift[k] = hp[i].shift[k];
        }
      VT_CUDA(cudaMemcpyAsync(f->d_prep, f->h_prep, sizeof(PrepPose) * (size_t)nposes, cudaMemcpyHostToDevice, c.stream));
    }
    rc = eval_prepared(f, com, f->d_prep, nposes, d_cc);
    if (rc) return rc;
    // one small readback per call: a truncated tile list or an ambiguous cast must never go unnoticed
    int amb = 0;
    VT_CUDA(cudaMemcpyAsync(&amb, f->d_flag, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    VT_CUDA(cudaStreamSynchronize(c.stream));
    int ovf = 0;
    rc = any_overflow(f, &ovf);
    if (rc) return rc;
    if (!ovf && !amb) { f->last_prep_on_host = host_prep ? 1 : 0; return 0; }
    if (ovf) for (int k = 0; k < 2; ++k) f->pb[k].use_lists = false;
    if (amb) host_prep = true;
  }
  set_error("vt_fit_eval_poses_device: no clean pass after 4 attempts");
  return 3;
}

int vt_fit_rotate_table(vt_fit_ctx *f, const float com[3], int stride, int max_rot, int first, int step,
                        float *table) {
  VT_REQUIRE_CTX();
  if (max_rot < 1 || step < 1 || first < 0) { set_error("vt_fit_rotate_table: bad arguments"); return 1; }
  const int total = max_rot * max_rot * max_rot;
  std::vector<vt_pose> poses;
  std::vector<int> idx;
  for (int i = first; i < total; i += step) {
    vt_pose q;
    q.rot_deg[0] = (float)(stride * (i / (max_rot * max_rot)));   // x outermost, TclVoltool.C:160-174
    q.rot_deg[1] = (float)(stride * ((i / max_rot) % max_rot));
    q.rot_deg[2] = (float)(stride * (i % max_rot));
    q.shift[0] = q.shift[1] = q.shift[2] = 0.f;
    poses.push_back(q);
    idx.push_back(i);
  }
  std::vector<float> cc(poses.size());
  int rc = vt_fit_eval_poses(f, com, poses.data(), (long)poses.size(), cc.data());
  if (rc) return rc;
  for (size_t k = 0; k < idx.size(); ++k) table[idx[k]] = cc[k];
  return 0;
}

int vt_fit(float *pos, const float *radii, const float *mass, long natoms, const vt_map *target, float resolution,
           vt_fit_result *res) {
  VT_REQUIRE_CTX();
  const bool verbose = getenv("VOLTOOL_VERBOSE") != nullptr;
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
    return std::chrono::duration<double, std::milli>(b - a).count();
  };
  const auto t0 = now();
  float dencom[3];
  int rc = vt_vol_com(target, dencom);  // :799
  if (rc) return rc;
  const auto t1 = now();
  // the evaluator context is created on the COM-aligned coordinates inside the schedule: build it
  // on the input coordinates (the atom order/radii are what matter) and let set_origin refresh it
  vt_fit_ctx *f = nullptr;
  rc = vt_fit_create(pos, radii, natoms, target, resolution, &f);
  if (rc) return rc;
  const auto t2 = now();
  rc = vt_fit_schedule(pos, mass, natoms, dencom, cuda_table_eval, f, res);
  const auto t3 = now();
  vt_fit_destroy(f);
  if (verbose)
    fprintf(stderr, "voltool_gpu: vt_fit: vol_com %.2f ms, workspace %.2f ms, schedule %.2f ms, teardown %.2f ms\n", ms(t0, t1),
            ms(t1, t2), ms(t2, t3), ms(t3, now()));
  return rc;
}

}  // extern "C"
