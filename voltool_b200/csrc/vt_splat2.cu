// vt_splat2.cu -- barrier-free splat: per-tile atom lists (one warp per tile) + per-region
// accumulation (one warp per 8x8x8 region).  Same arithmetic and the same deterministic
// accumulation order per voxel as the cooperative kernel in vt_sim.cu up to how partial sums are grouped
// (group order, then lane order; the two agree to a few 1e-7 relative, tests/test_gpu_fullsize.py), but no
// block barrier anywhere, so unevenly filled regions do not stall one another:
//
//   atom_box_kernel   per (pose, atom): the reference's truncated integer box, packed 3 x (lo|hi<<16)
//   pose_bin_kernel   (vt_sim.cu) per (pose, group): tile bitmaps
//   gather_kernel     warp per tile: walk the tile's group bitmap, test the 32 boxes of each group,
//                     append id | region-mask << 24 to the tile's list in global memory (L2 resident)
//   splat_list_kernel warp per (tile, region): filter the list by region bit, stage 32 atoms at a
//                     time as separable records, accumulate 4 columns x 4 x-voxels per thread, write once
//
// A tile whose list would overflow its slot sets a flag; the caller then reruns the batch with the
// cooperative kernel (no silent truncation).
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "vt_pose.cuh"

namespace vt {

__device__ __forceinline__ float ex2_approx2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ int find_pose_tile(const PoseGeom *__restrict__ geo, int npose, int item) {
  int lo = 0, hi = npose - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (geo[mid].tile_base <= item) lo = mid; else hi = mid - 1;
  }
  return lo;
}

constexpr int kListRegions = 4;   // regions of the list path: 8 x 8 x 8 voxels, 2 x 2 x 1 per tile (bit r: x half r&1, y half r>>1)

// ------------------------------------------------------------------------------------------ boxes
__global__ void __launch_bounds__(256) atom_box_kernel(const PoseGeom *__restrict__ geo, int natoms,
                                                       const float4 *__restrict__ tpos_all,
                                                       const float4 *__restrict__ cst, float invgs,
                                                       uint3 *__restrict__ boxes, uint3 *__restrict__ gboxes, int ngroups) {
  const int p = blockIdx.y;
  const PoseGeom &G = geo[p];
  const float ox = G.org[0], oy = G.org[1], oz = G.org[2];
  const int nv0 = G.nv[0], nv1 = G.nv[1], nv2 = G.nv[2];
  const int nround = (natoms + 31) & ~31;   // whole warps: a warp trip covers exactly one group of 32 atoms
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nround; i += gridDim.x * blockDim.x) {
    int xmin = 1, xmax = 0, ymin = 1, ymax = 0, zmin = 1, zmax = 0;
    if (i < natoms) {
      const float4 a4 = tpos_all[(size_t)p * natoms + i];
      const float rl = cst[i].z;
      const float ax = a4.x - ox, ay = a4.y - oy, az = a4.z - oz;  // xyzr = pos - origin
      float t = ax * invgs;
      xmin = max(__float2int_rz(t - rl), 0); xmax = min(__float2int_rz(t + rl), nv0 - 1);
      t = ay * invgs;
      ymin = max(__float2int_rz(t - rl), 0); ymax = min(__float2int_rz(t + rl), nv1 - 1);
      t = az * invgs;
      zmin = max(__float2int_rz(t - rl), 0); zmax = min(__float2int_rz(t + rl), nv2 - 1);
      const bool empty = xmin > xmax || ymin > ymax || zmin > zmax;
      boxes[(size_t)p * natoms + i] = empty ? make_uint3(1u, 1u, 1u)   // lo 1 > hi 0: empty
                                            : make_uint3((unsigned)xmin | ((unsigned)xmax << 16), (unsigned)ymin | ((unsigned)ymax << 16),
                                                         (unsigned)zmin | ((unsigned)zmax << 16));
      if (empty) { xmin = ymin = zmin = 1; xmax = ymax = zmax = 0; }
    }
    // the group's box: union of its atoms' boxes (what pose_bin_box_kernel marks tiles with)
    const bool has = xmin <= xmax;
    int lo[3] = {has ? xmin : 0x7fffffff, has ? ymin : 0x7fffffff, has ? zmin : 0x7fffffff};
    int hi[3] = {has ? xmax : -1, has ? ymax : -1, has ? zmax : -1};
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        lo[k] = min(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
        hi[k] = max(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
      }
    if ((threadIdx.x & 31) == 0 && gboxes) {
      const int g = i >> 5;
      if (g < ngroups)
        gboxes[(size_t)p * ngroups + g] = hi[0] < 0 ? make_uint3(1u, 1u, 1u)
                                                     : make_uint3((unsigned)lo[0] | ((unsigned)hi[0] << 16), (unsigned)lo[1] | ((unsigned)hi[1] << 16),
                                                                  (unsigned)lo[2] | ((unsigned)hi[2] << 16));
    }
  }
}

// Which groups can touch which tile, from the group boxes: thread per (pose, group) marks every tile its box
// meets (atomicOr: order independent).  Integer arithmetic only, and tighter than the bounding-sphere test of
// pose_bin_kernel (vt_sim.cu), which stays for the cooperative splat path that has no boxes.
__global__ void __launch_bounds__(256) pose_bin_box_kernel(const PoseGeom *__restrict__ geo, int ngroups,
                                                           const uint3 *__restrict__ gboxes,
                                                           unsigned int *__restrict__ bitmap, int tiles_cap, int nwords) {
  const int p = blockIdx.y;
  const PoseGeom &G = geo[p];
  const int ntx = G.tiles[0], nty = G.tiles[1], ntz = G.tiles[2];
  if (ntx == 0) return;
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < ngroups; g += gridDim.x * blockDim.x) {
    const uint3 b = gboxes[(size_t)p * ngroups + g];
    const int xmin = b.x & 0xffff, xmax = b.x >> 16, ymin = b.y & 0xffff, ymax = b.y >> 16;
    const int zmin = b.z & 0xffff, zmax = b.z >> 16;
    if (xmin > xmax) continue;
    const int tx0 = xmin / kTX, tx1 = min(ntx - 1, xmax / kTX);
    const int ty0 = ymin / kTY, ty1 = min(nty - 1, ymax / kTY);
    const int tz0 = zmin / kTZ, tz1 = min(ntz - 1, zmax / kTZ);
    const unsigned int bit = 1u << (g & 31);
    for (int tz = tz0; tz <= tz1; ++tz)
      for (int ty = ty0; ty <= ty1; ++ty)
        for (int tx = tx0; tx <= tx1; ++tx)
          atomicOr(&bitmap[((size_t)p * tiles_cap + (tz * nty + ty) * ntx + tx) * nwords + (g >> 5)], bit);
  }
}

// boxes, group boxes and the tile bitmaps made from them
static int lists_boxes_and_bins(const AtomSet &as, PoseBatch &pb, int npose) {
  Ctx &c = ctx();
  const int nwords = (as.ngroups + 31) / 32;
  int gx = (as.n + 255) / 256;
  if (gx > 64) gx = 64;
  VT_CUDA(cudaMemsetAsync(pb.d_bitmap, 0, sizeof(unsigned int) * (size_t)npose * pb.tiles_cap * nwords, c.stream));
  atom_box_kernel<<<dim3(gx, npose), 256, 0, c.stream>>>(pb.d_geom, as.n, pb.d_tpos, as.d_const, as.prm.invgs, pb.d_boxes,
                                                          pb.d_gbox, as.ngroups);
  VT_LAUNCHED();
  int gg = (as.ngroups + 255) / 256;
  if (gg > 32) gg = 32;
  pose_bin_box_kernel<<<dim3(gg, npose), 256, 0, c.stream>>>(pb.d_geom, as.ngroups, pb.d_gbox, pb.d_bitmap, pb.tiles_cap, nwords);
  VT_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------ gather
constexpr int kGlist = 1024;  // hit groups expanded per round (32 bitmap words)

__global__ void __launch_bounds__(128) gather_kernel(const PoseGeom *__restrict__ geo, const int *__restrict__ totals,
                                                     int npose, int natoms, const unsigned int *__restrict__ bitmap,
                                                     int tiles_cap, int nwords, const uint3 *__restrict__ boxes,
                                                     unsigned int *__restrict__ tlist, int *__restrict__ tcount,
                                                     int cap_t, int *__restrict__ overflow, int *queues) {
  __shared__ unsigned short s_gl[4][kGlist];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned int lt = (1u << lane) - 1u;
  const int total = totals[0];
  int *queue = queues + 2;
  while (true) {
    int item = 0;
    if (lane == 0) item = atomicAdd(queue, 1);   // tiles differ a lot in how many groups reach them
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= total) break;
    const int p = find_pose_tile(geo, npose, item);
    const PoseGeom &G = geo[p];
    const int local = item - G.tile_base;
    const int tx = local % G.tiles[0], ty = (local / G.tiles[0]) % G.tiles[1], tz = local / (G.tiles[0] * G.tiles[1]);
    const int x0 = tx * kTX, y0 = ty * kTY, z0 = tz * kTZ;
    const unsigned int *__restrict__ bits = bitmap + ((size_t)p * tiles_cap + local) * nwords;
    const uint3 *__restrict__ bx = boxes + (size_t)p * natoms;
    unsigned int *__restrict__ out = tlist + ((size_t)p * tiles_cap + local) * cap_t;
    int cnt = 0;
    for (int w0 = 0; w0 < nwords; w0 += 32) {
      // 32 words -> ordered list of set bits
      unsigned int word = (w0 + lane < nwords) ? bits[w0 + lane] : 0u;
      const int pc = __popc(word);
      int incl = pc;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      const int nhit = __shfl_sync(0xffffffffu, incl, 31);
      int off = incl - pc;
      while (word) {
        const int bit = __ffs(word) - 1;
        word &= word - 1;
        s_gl[warp][off++] = (unsigned short)(lane * 32 + bit);
      }
      __syncwarp();
      // one group per iteration, lane per atom; the next group's boxes are loaded one iteration ahead
      uint3 nb = make_uint3(1u, 1u, 1u);
      if (nhit > 0) {
        const int id0 = (w0 * 32 + (int)s_gl[warp][0]) * 32 + lane;
        if (id0 < natoms) nb = bx[id0];
      }
      for (int gi = 0; gi < nhit; ++gi) {
        const int id = (w0 * 32 + (int)s_gl[warp][gi]) * 32 + lane;
        const uint3 b = nb;
        nb = make_uint3(1u, 1u, 1u);
        if (gi + 1 < nhit) {
          const int idn = (w0 * 32 + (int)s_gl[warp][gi + 1]) * 32 + lane;
          if (idn < natoms) nb = bx[idn];
        }
        const int xmin = b.x & 0xffff, xmax = b.x >> 16, ymin = b.y & 0xffff, ymax = b.y >> 16;
        const int zmin = b.z & 0xffff, zmax = b.z >> 16;
        const bool take = id < natoms && xmin <= xmax && xmin <= x0 + kTX - 1 && xmax >= x0 && ymin <= y0 + kTY - 1 &&
                          ymax >= y0 && zmin <= z0 + kTZ - 1 && zmax >= z0;
        int mask = 0;
        if (take) {
          // bit r: box meets region r (x half r&1, y half r>>1; a region spans the tile in z): the AND of two
          // nibble patterns, one per axis, each the OR of the halves the box reaches
          static_assert(kListRegions == 4 && kTX == 16 && kTY == 16, "region masks are written for 2 x 2 x 1 regions per tile");
          const int bx_ = (xmin <= x0 + 7 ? 0x5 : 0) | (xmax >= x0 + 8 ? 0xa : 0);
          const int by_ = (ymin <= y0 + 7 ? 0x3 : 0) | (ymax >= y0 + 8 ? 0xc : 0);
          mask = bx_ & by_;
        }
        const unsigned int bal = __ballot_sync(0xffffffffu, take);
        if (take) {
          const int slot = cnt + __popc(bal & lt);
          if (slot < cap_t) out[slot] = (unsigned int)id | ((unsigned int)mask << 24);
        }
        cnt += __popc(bal);
      }
      __syncwarp();
    }
    if (lane == 0) {
      tcount[(size_t)p * tiles_cap + local] = cnt;
      if (cnt > cap_t) atomicExch(overflow, 1);
    }
  }
}

// ------------------------------------------------------------------------------------------ accumulate
constexpr int kRec2 = 28;  // Ex[8] | dy^2[8] | dz^2[8] | arinv, radlim^2, pad x 2 (stride 28: the 16-byte staging stores of 8 lanes cover all 32 banks)

__global__ void __launch_bounds__(128, 6) splat_list_kernel(const PoseGeom *__restrict__ geo,
                                                            const int *__restrict__ totals, int npose, int natoms,
                                                            const float4 *__restrict__ tpos_all,
                                                            const float4 *__restrict__ cst,
                                                            const uint3 *__restrict__ boxes,
                                                            const unsigned int *__restrict__ tlist,
                                                            const int *__restrict__ tcount, int cap_t, int tiles_cap,
                                                            float gs, float *__restrict__ grid_all, long cap_grid,
                                                            int *queues) {
  __shared__ __align__(16) float s_rec[4][32 * kRec2];
  __shared__ int s_pend[4][64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned int lt = (1u << lane) - 1u;
  const long ntask = (long)totals[0] * kListRegions;
  int *queue = queues;
  // A warp owns one 8 x 8 x 8 region; lane = (x half, y, z quad) owns 4 x-voxels of the four columns (y, 4 zq + c).
  // It evaluates the y-z factor of two of them, (y, 4 zq + 2 xh + {0, 1}), and takes the other two from its
  // partner lane (same y and z quad, other x half) by shuffle.  A visit is 7 shared-memory-pipe wavefronts (half
  // of Ex: 2; dy^2, the dz^2 pair, the constants: 1 each; two shuffles) for 16 voxels and 8 FFMA2 -- half the
  // wavefronts per voxel of the 8 x 8 x 4 regions this replaced, which is what counts: the kernel saturates the
  // shared-memory pipe, not the issue slots.  A region twice as tall also meets a third fewer atoms per voxel.
  const int ly = lane & 7, xh = (lane >> 3) & 1, zq = lane >> 4;
  float *rec = s_rec[warp];
  while (true) {
    // regions cost very different amounts (empty corners vs the dense interior): warps draw them from a queue
    int task = 0;
    if (lane == 0) task = atomicAdd(queue, 1);
    task = __shfl_sync(0xffffffffu, task, 0);
    if (task >= ntask) break;
    const int item = (int)(task / kListRegions), region = (int)(task - (long)item * kListRegions);
    const int p = find_pose_tile(geo, npose, item);
    const PoseGeom &G = geo[p];
    const int local = item - G.tile_base;
    const int tx = local % G.tiles[0], ty = (local / G.tiles[0]) % G.tiles[1], tz = local / (G.tiles[0] * G.tiles[1]);
    const int nv0 = G.nv[0], nv1 = G.nv[1], nv2 = G.nv[2];
    const float ox = G.org[0], oy = G.org[1], oz = G.org[2];
    const int X0 = tx * kTX + (region & 1) * 8, Y0 = ty * kTY + (region >> 1) * 8, Z0 = tz * kTZ;
    const float4 *__restrict__ tpos = tpos_all + (size_t)p * natoms;
    const uint3 *__restrict__ bx = boxes + (size_t)p * natoms;
    const unsigned int *__restrict__ list = tlist + ((size_t)p * tiles_cap + local) * cap_t;
    const int n = min(tcount[(size_t)p * tiles_cap + local], cap_t);

    f32x2 acc[8];  // [column c][x pair]: c = 0, 1 own columns (z = 4 zq + 2 xh + c), c = 2, 3 the partner's
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = pack2(0.0f, 0.0f);

    auto process = [&](int m) {
      if (lane < m) {
        const int id = s_pend[warp][lane];
        const float4 a4 = tpos[id];
        const float4 k4 = cst[id];
        const uint3 b = bx[id];
        const float ax = a4.x - ox, ay = a4.y - oy, az = a4.z - oz;  // xyzr = pos - origin
        const float arinv = k4.x;
        const int xmin = b.x & 0xffff, xmax = b.x >> 16, ymin = b.y & 0xffff, ymax = b.y >> 16;
        const int zmin = b.z & 0xffff, zmax = b.z >> 16;
        float *r = rec + lane * kRec2;
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int X = X0 + k;
          const float dx = (float)X * gs - ax;
          const float e = ex2_approx2((dx * dx) * arinv);
          v[k] = (X >= xmin && X <= xmax) ? e : 0.0f;
        }
        *reinterpret_cast<float4 *>(r) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4 *>(r + 4) = make_float4(v[4], v[5], v[6], v[7]);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int Y = Y0 + k;
          const float dy = (float)Y * gs - ay;
          v[k] = (Y >= ymin && Y <= ymax) ? dy * dy : INFINITY;
        }
        *reinterpret_cast<float4 *>(r + 8) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4 *>(r + 12) = make_float4(v[4], v[5], v[6], v[7]);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int Z = Z0 + k;
          const float dz = (float)Z * gs - az;
          v[k] = (Z >= zmin && Z <= zmax) ? dz * dz : INFINITY;
        }
        *reinterpret_cast<float4 *>(r + 16) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4 *>(r + 20) = make_float4(v[4], v[5], v[6], v[7]);
        *reinterpret_cast<float4 *>(r + 24) = make_float4(arinv, k4.y, 0.f, 0.f);
      }
      __syncwarp();
      // branch-free: a staged atom whose disc misses every lane just adds zeros (rare: its box already
      // meets this region).  Two atoms per trip so their shared-memory loads overlap.
      auto visit = [&](const float *r) {
        const float dy2 = r[8 + ly];
        const float2 dz2 = *reinterpret_cast<const float2 *>(r + 16 + 4 * zq + 2 * xh);   // this lane's two columns
        const float2 c2 = *reinterpret_cast<const float2 *>(r + 24);                      // arinv, radlim^2
        const float d2a = dy2 + dz2.x, d2b = dy2 + dz2.y;                                 // dy*dy + dz*dz
        const float ea = (d2a < c2.y) ? ex2_approx2(d2a * c2.x) : 0.0f;                   // !(dy2dz2 >= radlim2)
        const float eb = (d2b < c2.y) ? ex2_approx2(d2b * c2.x) : 0.0f;
        const float ec = __shfl_xor_sync(0xffffffffu, ea, 8);                             // the partner's two columns
        const float ed = __shfl_xor_sync(0xffffffffu, eb, 8);
        const ulonglong2 ex = *reinterpret_cast<const ulonglong2 *>(r + 4 * xh);          // Ex of this lane's four x
        const f32x2 a2 = pack2(ea, ea), b2 = pack2(eb, eb), c22 = pack2(ec, ec), d22 = pack2(ed, ed);
        acc[0] = fma2(ex.x, a2, acc[0]); acc[1] = fma2(ex.y, a2, acc[1]);
        acc[2] = fma2(ex.x, b2, acc[2]); acc[3] = fma2(ex.y, b2, acc[3]);
        acc[4] = fma2(ex.x, c22, acc[4]); acc[5] = fma2(ex.y, c22, acc[5]);
        acc[6] = fma2(ex.x, d22, acc[6]); acc[7] = fma2(ex.y, d22, acc[7]);
      };
      int a = 0;
      for (; a + 4 <= m; a += 4) {
        visit(rec + a * kRec2);
        visit(rec + (a + 1) * kRec2);
        visit(rec + (a + 2) * kRec2);
        visit(rec + (a + 3) * kRec2);
      }
      for (; a < m; ++a) visit(rec + a * kRec2);
      __syncwarp();
    };

    int have = 0;
    unsigned int e_next = lane < n ? __ldg(list + lane) : 0u;  // the list is read one chunk ahead
    for (int b0 = 0; b0 < n; b0 += 32) {
      const int i = b0 + lane;
      const unsigned int e = e_next;
      e_next = (i + 32 < n) ? __ldg(list + i + 32) : 0u;
      const bool mine = i < n && ((e >> (24 + region)) & 1u);
      const unsigned int bal = __ballot_sync(0xffffffffu, mine);
      if (bal == 0u) continue;
      if (mine) s_pend[warp][have + __popc(bal & lt)] = (int)(e & 0xffffffu);
      have += __popc(bal);
      __syncwarp();
      if (have >= 32) {
        process(32);
        const int rest = have - 32;
        const int keep = lane < rest ? s_pend[warp][32 + lane] : 0;
        __syncwarp();
        if (lane < rest) s_pend[warp][lane] = keep;
        have = rest;
        __syncwarp();
      }
    }
    if (have > 0) process(have);

    const int Y = Y0 + ly;
    if (Y < nv1) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        // acc[2c], acc[2c+1]: c < 2 own columns, else the partner's
        const int Z = Z0 + 4 * zq + 2 * (c < 2 ? xh : (xh ^ 1)) + (c & 1);
        if (Z >= nv2) continue;
        float o[4];
        unpack2(acc[2 * c], o[0], o[1]);
        unpack2(acc[2 * c + 1], o[2], o[3]);
        float *row = grid_all + (size_t)p * cap_grid + ((size_t)Z * nv1 + Y) * nv0 + X0 + 4 * xh;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (X0 + 4 * xh + k < nv0) row[k] = o[k];
      }
    }
  }
}

int batch_lists_overflowed(PoseBatch &pb, int *flag) {
  Ctx &c = ctx();
  int h[4];
  VT_CUDA(cudaMemcpyAsync(h, pb.d_totals, sizeof(h), cudaMemcpyDeviceToHost, c.stream));
  VT_CUDA(cudaStreamSynchronize(c.stream));
  *flag = h[2];
  return 0;
}

// Size the per-tile lists from the pose(s) already prepared in pb: count-only gather on pose 0,
// capacity = 1.6 x the fullest tile (rigid poses of one structure fill tiles alike); then allocate.
int batch_calibrate_lists(const AtomSet &as, PoseBatch &pb) {
  VT_REQUIRE_CTX();
  Ctx &c = ctx();
  pb.use_lists = false;
  if (getenv("VOLTOOL_SPLAT_COOP") || as.n >= (1 << 24) || pb.cap_n > 65535) return 0;
  const int nwords = (as.ngroups + 31) / 32;
  if (dev_alloc((void **)&pb.d_boxes, sizeof(uint3) * (size_t)pb.B * as.n)) return 2;
  if (dev_alloc((void **)&pb.d_tcount, sizeof(int) * (size_t)pb.B * pb.tiles_cap)) return 2;
  if (dev_alloc((void **)&pb.d_gbox, sizeof(uint3) * (size_t)pb.B * as.ngroups)) return 2;
  VT_CUDA(cudaMemsetAsync(pb.d_totals + 2, 0, sizeof(int), c.stream));
  int rc = lists_boxes_and_bins(as, pb, 1);
  if (rc) return rc;
  VT_CUDA(cudaMemsetAsync(pb.d_tcount, 0, sizeof(int) * (size_t)pb.tiles_cap, c.stream));
  VT_CUDA(cudaMemsetAsync(pb.d_totals + 6, 0, sizeof(int), c.stream));   // the gather work queue (else reset by pose_setup)
  gather_kernel<<<c.sm_count * 8, 128, 0, c.stream>>>(pb.d_geom, pb.d_totals, 1, as.n, pb.d_bitmap, pb.tiles_cap, nwords,
                                                      pb.d_boxes, nullptr, pb.d_tcount, 0, pb.d_totals + 3, pb.d_totals + 4);
  VT_LAUNCHED();
  std::vector<int> cnt(pb.tiles_cap);
  VT_CUDA(cudaMemcpyAsync(cnt.data(), pb.d_tcount, sizeof(int) * pb.tiles_cap, cudaMemcpyDeviceToHost, c.stream));
  VT_CUDA(cudaStreamSynchronize(c.stream));
  int mx = 0;
  for (int v : cnt) mx = v > mx ? v : mx;
  int cap = (int)(1.6 * mx) + 128;
  cap = (cap + 127) & ~127;
  if (const char *force = getenv("VOLTOOL_LIST_CAP")) cap = std::max(32, atoi(force));   // tests: provoke the overflow fallback
  const size_t bytes = sizeof(unsigned int) * (size_t)pb.B * pb.tiles_cap * cap;
  if (bytes > ((size_t)12 << 30)) return 0;  // would not pay: keep the cooperative kernel
  if (dev_alloc((void **)&pb.d_tlist, bytes)) return 2;
  pb.cap_t = cap;
  pb.use_lists = true;
  return 0;
}

// Single structure (vt_sim): no counting pass.  The slot is sized from the mean number of atoms whose
// box can reach a tile (x4 for the denser interior of the structure); if a tile overflows anyway the
// caller falls back to the cooperative kernel, as everywhere.
int batch_estimate_lists(const AtomSet &as, PoseBatch &pb, const int nv[3]) {
  VT_REQUIRE_CTX();
  Ctx &c = ctx();
  pb.use_lists = false;
  if (pb.B != 1 || getenv("VOLTOOL_SPLAT_COOP") || as.n >= (1 << 24) || pb.cap_n > 65535) return 0;
  const double h = std::ceil((double)as.prm.gausslim * as.maxrad * as.prm.radscale * as.prm.invgs) + 1.0;
  const double reach = (kTX + 2 * h) * (kTY + 2 * h) * (kTZ + 2 * h);
  const double mean = (double)as.n * reach / ((double)nv[0] * nv[1] * nv[2]);
  long cap = (long)(4.0 * mean) + 256;
  cap = std::min<long>(cap, ((long)as.n + 127) & ~127L);
  cap = (cap + 127) & ~127L;
  if (const char *force = getenv("VOLTOOL_LIST_CAP")) cap = std::max(32, atoi(force));
  const size_t bytes = sizeof(unsigned int) * (size_t)pb.tiles_cap * (size_t)cap;
  if (bytes > ((size_t)4 << 30)) return 0;
  if (dev_alloc((void **)&pb.d_boxes, sizeof(uint3) * (size_t)as.n)) return 2;
  if (dev_alloc((void **)&pb.d_gbox, sizeof(uint3) * (size_t)as.ngroups)) return 2;
  if (dev_alloc((void **)&pb.d_tcount, sizeof(int) * (size_t)pb.tiles_cap)) return 2;
  if (dev_alloc((void **)&pb.d_tlist, bytes)) return 2;
  VT_CUDA(cudaMemsetAsync(pb.d_totals + 2, 0, sizeof(int), c.stream));
  pb.cap_t = (int)cap;
  pb.use_lists = true;
  return 0;
}

// returns 0 on launch; the overflow flag (d_totals[2]) must be checked by the caller after a sync
int batch_lists_front(const AtomSet &as, PoseBatch &pb, int npose) {
  VT_REQUIRE_CTX();
  Ctx &c = ctx();
  const int nwords = (as.ngroups + 31) / 32;
  prof_begin(PK_SETUP);
  int rc = lists_boxes_and_bins(as, pb, npose);
  if (rc) return rc;
  gather_kernel<<<c.sm_count * 8, 128, 0, c.stream>>>(pb.d_geom, pb.d_totals, npose, as.n, pb.d_bitmap, pb.tiles_cap, nwords,
                                                      pb.d_boxes, pb.d_tlist, pb.d_tcount, pb.cap_t, pb.d_totals + 2, pb.d_totals + 4);
  prof_end(PK_SETUP);
  VT_LAUNCHED();
  return 0;
}

int batch_lists_back(const AtomSet &as, PoseBatch &pb, int npose) {
  VT_REQUIRE_CTX();
  Ctx &c = ctx();
  prof_begin(PK_SPLAT);
  // persistent: the warps draw their tasks from a queue, so one resident set of CTAs is enough (6 fit an SM).
  // The fit pipeline asks for 5 per SM: the kernel saturates the shared-memory pipe long before it runs out
  // of warps, and the free slots are where the list building of the next batch runs beside it
  // (measured 6 / 5 / 4 per SM: 38.3 / 38.6 / 37.6 k poses/s; with the heavier front end before the group
  // boxes it was 35.0 / 35.5 / 36.9)
  static const int forced = getenv("VOLTOOL_SPLAT_CTAS") ? atoi(getenv("VOLTOOL_SPLAT_CTAS")) : 0;
  const int per_sm = forced > 0 ? forced : (pb.splat_ctas > 0 ? pb.splat_ctas : 6);
  splat_list_kernel<<<c.sm_count * per_sm, 128, 0, c.stream>>>(pb.d_geom, pb.d_totals, npose, as.n, pb.d_tpos, as.d_const,
                                                              pb.d_boxes, pb.d_tlist, pb.d_tcount, pb.cap_t, pb.tiles_cap,
                                                              as.prm.gs, pb.d_grid, pb.cap_grid, pb.d_totals + 4);
  prof_end(PK_SPLAT);
  VT_LAUNCHED();
  return 0;
}

int batch_splat_lists(const AtomSet &as, PoseBatch &pb, int npose) {
  int rc = batch_lists_front(as, pb, npose);
  if (!rc) rc = batch_lists_back(as, pb, npose);
  return rc;
}

}  // namespace vt
