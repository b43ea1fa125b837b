// vt_volcom.cu -- vol_com (Voltool.C:229-247), bit-exact and parallel.
//
// The reference accumulates m_i * coord_i into three FLOAT accumulators strictly in voxel order.
// A float running sum is order dependent (it even stagnates once its ulp outgrows the addends),
// and fit() starts from this value, so the sequential semantics must be kept -- but not the
// sequential cost.  Observation: while the running sum s stays inside one binade (ulp u), every
// step is  s <- s + rint(a_i / u) * u  (round to nearest; a tie would consult the parity of s).
// So for a chunk of addends and a given binade the whole chunk collapses to one integer
//      T_e = sum_i rint(a_i / u_e)
// valid whenever s does not leave the binade inside the chunk: the smallest and largest partial sum of the rounded
// addends (kept per chunk and binade) give the excursion exactly.
// An exact tie (a_i / u_e = h + 1/2) rounds to the EVEN neighbour of the running sum, so its outcome depends on the
// parity of the sum when it is reached -- which the parity of the sum at the start of the chunk fixes.  Ties are not
// rare (an addend within 2^-k of the sum's magnitude ties with probability ~2^-k: signed noise in a map's padding
// keeps the sum that close to its addends for millions of voxels), so the tables hold the chunk's total for both
// start parities, T_e[0] and T_e[1], and the walk picks by the parity of s.
//
//   pass 1 (parallel, one CTA per chunk): a_i for the three accumulators, and for 26 candidate
//           binades starting at ceil(log2(sum |a_i|)) (T_e[0], T_e[1] - T_e[0], excursions).  Above that
//           window the chunk is the identity (|a_i| < u/8: kBinades - kBelow = 26 binades up), below it the chunk must be replayed.
//   pass 2 (one CTA): walks the chunks in order; a table hit is O(1) -- 32 of them at a time by a warp scan -- and
//           anything else (binade crossing, zero/subnormal sum) replays that chunk add by add.
// The result equals the sequential loop bit for bit for any input.  The mass is a double sum
// (order-insensitive to ~1e-16; the reference adds it in voxel order).
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "vt_device.cuh"

namespace vt {

constexpr int kBinades = 32;   // binades tabulated per chunk and accumulator (one per lane)
constexpr int kBelow = 6;      // ... of which this many lie below ceil(log2(sum |a_i|)): sums of random sign stay far below sum |a_i|

struct ComEntry {
  long long T0;     // sum of rounded addends in units of the binade's ulp, for an even sum at the start of the chunk
  int dT;           // ... for an odd one: T0 + dT
  int lo[2], hi[2]; // smallest / largest partial sum inside the chunk (0 included), per start parity: the exact
                    // excursion of the running sum (|partial sums| < 2^23 + L inside the window)
  int pad;
};

__device__ __forceinline__ float com_addend(const DevCoord &c, int k, float m, int gx, int gy, int gz) {
  const double o = k == 0 ? c.ox : (k == 1 ? c.oy : c.oz);
  const float coord =
      (float)(((o + (double)((float)gx * c.xd[k])) + (double)((float)gy * c.yd[k])) + (double)((float)gz * c.zd[k]));
  return m * coord;
}

// pass 1: blockDim = 96 (warp k <-> accumulator k).  Dynamic smem: 3 * L floats.
__global__ void __launch_bounds__(96) com_tables_kernel(const float *__restrict__ d, long n, DevCoord c, int L,
                                                        ComEntry *__restrict__ tab, int *__restrict__ emin) {
  extern __shared__ float s_a[];  // [3][L]
  const long chunk = blockIdx.x;
  const long base = chunk * L;
  const int cnt = (int)((n - base) < L ? (n - base) : L);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long nxy = (long)c.nx * c.ny;
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
    const long idx = base + i;
    const int gz = (int)(idx / nxy);
    const long r = idx - (long)gz * nxy;
    const int gy = (int)(r / c.nx), gx = (int)(r - (long)gy * c.nx);
    const float m = d[idx];
#pragma unroll
    for (int k = 0; k < 3; ++k) s_a[k * L + i] = com_addend(c, k, m, gx, gy, gz);
  }
  __syncthreads();
  const float *a = s_a + warp * L;
  // window start: smallest binade in which the chunk cannot move s by more than ~2^23 ulps
  double tot = 0.0;
  for (int i = lane; i < cnt; i += 32) tot += fabs((double)a[i]);
  tot = warp_sum(tot);
  int e0;
  if (!(tot > 0.0)) e0 = -100000;               // all addends zero: identity in every binade
  else if (!(tot < 1.0e300)) e0 = 100000;       // inf/NaN in the chunk: always replay
  else { int ex; frexp(tot, &ex); e0 = ex - kBelow; }    // tot < 2^ex; the window is [ex - kBelow, ex - kBelow + kBinades)
  if (lane == 0) emin[chunk * 3 + warp] = e0;
  if (lane < kBinades && e0 > -100000 && e0 < 100000) {
    const int e = e0 + lane;
    const double scale = ldexp(1.0, 23 - e);    // 1 / ulp of binade e
    long long T0 = 0, T1 = 0, lo0 = 0, hi0 = 0, lo1 = 0, hi1 = 0;
    for (int i = 0; i < cnt; ++i) {
      const double x = (double)a[i] * scale;    // exact
      if (fabs(x - trunc(x)) == 0.5) {          // tie: the even one of sum + h, sum + h + 1
        const long long h = (long long)floor(x);
        T0 += h + ((T0 + h) & 1);               // parity of the running sum = start parity + total so far
        T1 += h + ((1 + T1 + h) & 1);
      } else {
        const long long ri = (long long)rint(x);
        T0 += ri; T1 += ri;
      }
      lo0 = T0 < lo0 ? T0 : lo0; hi0 = T0 > hi0 ? T0 : hi0;
      lo1 = T1 < lo1 ? T1 : lo1; hi1 = T1 > hi1 ? T1 : hi1;
    }
    ComEntry en;
    en.T0 = T0;
    en.dT = (int)(T1 - T0);
    en.lo[0] = (int)lo0; en.hi[0] = (int)hi0; en.lo[1] = (int)lo1; en.hi[1] = (int)hi1;
    en.pad = 0;
    tab[(chunk * 3 + warp) * kBinades + lane] = en;
  }
}

// pass 2: one CTA of 8 warps.  Lanes 0..2 of warp 0 own the accumulators and take the O(1) table step when they
// can.  The walk is a chain of dependent steps, so what matters is the latency of one step: the table entries of
// the next kWalkBlock chunks are fetched by all 256 threads at once into shared memory -- for the binade each
// accumulator is in when the block starts, which is the right one unless the sum crosses a binade inside the block
// (then that step reads global memory, as every step used to).  When an accumulator has to replay a chunk, all 256
// threads first rebuild that chunk's addends in shared memory (coalesced loads, parallel coordinate arithmetic) and
// the owning lane then does only the dependent float adds.
constexpr int kWalkThreads = 256;
constexpr int kWalkBlock = 256;   // chunks per prefetch (one per thread)
constexpr unsigned int kCmdExit = 0xffffffffu, kCmdPrefetch = 0xfffffffeu;

__global__ void __launch_bounds__(kWalkThreads) com_walk_kernel(const float *__restrict__ d, long n, DevCoord c, int L,
                                                               long nchunks, const ComEntry *__restrict__ tab,
                                                               const int *__restrict__ emin, float *__restrict__ out,
                                                               int *__restrict__ replayed) {
  extern __shared__ float s_a[];  // [3][L]
  __shared__ unsigned int s_need;   // accumulators that replay the chunk in s_chunk, or a command
  __shared__ long s_chunk;          // chunk to replay / first chunk of the block to prefetch
  __shared__ int s_snap[3];         // exponent bits of each accumulator when the block was prefetched
  __shared__ int s_e0[3][kWalkBlock];
  __shared__ long long s_T[3][kWalkBlock];
  __shared__ int s_dT[3][kWalkBlock];
  __shared__ int s_lo[3][2][kWalkBlock], s_hi[3][2][kWalkBlock];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long nxy = (long)c.nx * c.ny;

  // all threads: rebuild the addends of one chunk for the accumulators in `need`
  auto stage = [&](long chunk, unsigned int need) {
    const long base = chunk * L;
    const int cnt = (int)((n - base) < L ? (n - base) : L);
    for (int i = tid; i < cnt; i += kWalkThreads) {
      const long idx = base + i;
      const int gz = (int)(idx / nxy);
      const long r = idx - (long)gz * nxy;
      const int gy = (int)(r / c.nx), gx = (int)(r - (long)gy * c.nx);
      const float m = d[idx];
#pragma unroll
      for (int kk = 0; kk < 3; ++kk)
        if ((need >> kk) & 1) s_a[kk * L + i] = com_addend(c, kk, m, gx, gy, gz);
    }
  };
  // all threads: thread t fetches chunk first + t's window start and its table entry for the snapshot binades
  auto prefetch = [&](long first) {
    const long chunk = first + tid;
    if (chunk < nchunks) {
      {
        const int kk = blockIdx.x;
        const int e0 = emin[chunk * 3 + kk];
        const int j = (s_snap[kk] - 127) - e0;
        ComEntry en;
        en.T0 = 0; en.dT = 0; en.lo[0] = en.lo[1] = en.hi[0] = en.hi[1] = 0;
        if (e0 > -100000 && e0 < 100000 && j >= 0 && j < kBinades) en = tab[(chunk * 3 + kk) * kBinades + j];
        s_e0[kk][tid] = e0; s_T[kk][tid] = en.T0; s_dT[kk][tid] = en.dT;
        s_lo[kk][0][tid] = en.lo[0]; s_lo[kk][1][tid] = en.lo[1]; s_hi[kk][0][tid] = en.hi[0]; s_hi[kk][1][tid] = en.hi[1];
      }
    }
  };

  if (warp != 0) {
    // helper warps sleep on barrier 1 until warp 0 posts a command; the table fast path of warp 0 involves no barrier
    while (true) {
      asm volatile("bar.sync 1, %0;" ::"r"(kWalkThreads));
      const unsigned int need = s_need;
      if (need == kCmdExit) break;
      if (need == kCmdPrefetch) prefetch(s_chunk);
      else stage(s_chunk, need);
      asm volatile("bar.sync 2, %0;" ::"r"(kWalkThreads));
    }
    return;
  }

  // one CTA per accumulator (blockIdx.x = 0, 1, 2 for x, y, z): their walks are independent, and a chunk one of
  // them has to replay does not hold up the other two
  const int K = blockIdx.x;
  const bool owner = lane == K;
  const int k = K;
  float s = 0.0f;
  int nreplay = 0;
  int snap = 0;
  bool skip = false;   // this accumulator already jumped the current group of 32 chunks
  for (long chunk = 0; chunk < nchunks; ++chunk) {
    const int t = (int)(chunk % kWalkBlock);
    if (t == 0) {
      snap = (int)((__float_as_uint(s) >> 23) & 0xff);
      if (owner) s_snap[k] = snap;
      if (lane == 0) { s_need = kCmdPrefetch; s_chunk = chunk; }
      __syncwarp();
      asm volatile("bar.sync 1, %0;" ::"r"(kWalkThreads));
      prefetch(chunk);
      asm volatile("bar.sync 2, %0;" ::"r"(kWalkThreads));
    }
    // 32 chunks in one step: within a binade the table step is the integer addition S <- S + T, so a warp scan of
    // the T's gives every intermediate sum at once and each chunk's validity test (the sum stays inside the binade,
    // no tie, a usable table entry) is checked against its own prefix.  If all 32 hold the accumulator jumps the
    // whole group; otherwise that accumulator walks the group chunk by chunk below.
    if ((t & 31) == 0) {
      skip = false;
      if (chunk + 32 <= nchunks) {
        int npass = 0;
        {
          const int kk = K;
          const unsigned int bits = __float_as_uint(__shfl_sync(0xffffffffu, s, kk));
          const int ebits = (int)((bits >> 23) & 0xff);
          const int snap_k = __shfl_sync(0xffffffffu, snap, kk);
          const long long S = (long long)(bits & 0x7fffffu) | 0x800000ll;
          const long long Ss = (bits >> 31) ? -S : S;
          const int e0 = s_e0[kk][t + lane];
          long long T0 = 0, T1 = 0;              // this chunk's step for an even / odd sum before it
          bool tabled = false;
          bool ok = ebits != 0 && ebits != 0xff && ebits == snap_k;
          if (e0 != -100000) {
            const int j = (ebits - 127) - e0;
            if (e0 >= 100000 || j < 0) ok = false;
            else if (j < kBinades) { T0 = s_T[kk][t + lane]; T1 = T0 + s_dT[kk][t + lane]; tabled = true; }
          }
          // inclusive scan of the steps as maps (start parity -> total): (F o G)[p] = F[p] + G[(p + F[p]) & 1]
          long long P0 = T0, P1 = T1;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const long long q0 = __shfl_up_sync(0xffffffffu, P0, o), q1 = __shfl_up_sync(0xffffffffu, P1, o);
            if (lane >= o) {   // q = the chunks before this lane's span, P = this lane's span
              const long long n0 = q0 + ((q0 & 1) ? P1 : P0), n1 = q1 + (((1 + q1) & 1) ? P1 : P0);
              P0 = n0; P1 = n1;
            }
          }
          const long long P = (Ss & 1) ? P1 : P0;                                   // total through this chunk
          long long Pex = __shfl_up_sync(0xffffffffu, P, 1);                        // ... up to the chunk before
          if (lane == 0) Pex = 0;
          const long long before = Ss + Pex;
          const int par = (int)(before & 1);
          const long long lo = before + (tabled ? s_lo[kk][par][t + lane] : 0), hi = before + (tabled ? s_hi[kk][par][t + lane] : 0);
          const bool inside = Ss > 0 ? (lo >= 0x800001ll && hi <= 0xfffffell) : (hi <= -0x800001ll && lo >= -0xfffffell);
          // (a chunk without addends, e0 == -100000, changes nothing wherever the sum stands)
          const bool pass = __all_sync(0xffffffffu, ok && (inside || e0 == -100000));
          const long long R = Ss + __shfl_sync(0xffffffffu, P, 31);
          if (pass) {
            ++npass;
            if (lane == kk) {
              const long long Ra = R < 0 ? -R : R;
              s = __uint_as_float((R < 0 ? 0x80000000u : 0u) | ((unsigned int)ebits << 23) | (unsigned int)(Ra & 0x7fffffll));
              skip = true;
            }
          }
        }
        if (npass == 1) { chunk += 31; continue; }
      }
    }
    bool replay = false;
    if (owner && !skip) {
      const int e0 = s_e0[k][t];
      if (e0 != -100000) {  // else: nothing to add
        replay = true;
        const unsigned int bits = __float_as_uint(s);
        const int ebits = (int)((bits >> 23) & 0xff);
        if (ebits != 0 && ebits != 0xff && e0 < 100000) {
          const int j = (ebits - 127) - e0;                  // |s| in [2^e, 2^(e+1)), e = ebits-127
          const long long S = (long long)(bits & 0x7fffffu) | 0x800000ll;  // |s| / ulp
          const long long Ss = (bits >> 31) ? -S : S;
          long long T = 0, xlo = 0, xhi = 0;   // total and excursion of the chunk for this start parity
          bool have = false;
          if (j >= kBinades) { have = true; }                // every addend rounds to zero
          else if (j >= 0) {
            const int par = (int)(Ss & 1);
            if (ebits == snap) { T = s_T[k][t] + (par ? s_dT[k][t] : 0); xlo = s_lo[k][par][t]; xhi = s_hi[k][par][t]; }
            else {                                           // the sum left the prefetched binade inside this block
              const ComEntry en = tab[(chunk * 3 + k) * kBinades + j];
              T = en.T0 + (par ? en.dT : 0); xlo = en.lo[par]; xhi = en.hi[par];
            }
            have = true;
          }
          if (have) {
            const long long lo = Ss + xlo, hi = Ss + xhi;
            const bool inside =
                Ss > 0 ? (lo >= 0x800001ll && hi <= 0xfffffell) : (hi <= -0x800001ll && lo >= -0xfffffell);
            if (inside) {
              const long long R = Ss + T;
              const long long Ra = R < 0 ? -R : R;
              s = __uint_as_float((R < 0 ? 0x80000000u : 0u) | ((unsigned int)ebits << 23) |
                                  (unsigned int)(Ra & 0x7fffffll));
              replay = false;
            }
          }
        }
      }
    }
    const unsigned int need = __ballot_sync(0xffffffffu, replay);
    if (need == 0u) continue;
    if (lane == 0) { s_need = need; s_chunk = chunk; }
    __syncwarp();
    asm volatile("bar.sync 1, %0;" ::"r"(kWalkThreads));
    stage(chunk, need);
    asm volatile("bar.sync 2, %0;" ::"r"(kWalkThreads));
    if (replay) {
      ++nreplay;
      const long base = chunk * L;
      const int cnt = (int)((n - base) < L ? (n - base) : L);
      const float *a = s_a + k * L;
      int i = 0;
      for (; i + 16 <= cnt; i += 16) {  // loads batched ahead of the dependent add chain
        const float4 q0 = *reinterpret_cast<const float4 *>(a + i), q1 = *reinterpret_cast<const float4 *>(a + i + 4);
        const float4 q2 = *reinterpret_cast<const float4 *>(a + i + 8), q3 = *reinterpret_cast<const float4 *>(a + i + 12);
        s = s + q0.x; s = s + q0.y; s = s + q0.z; s = s + q0.w;
        s = s + q1.x; s = s + q1.y; s = s + q1.z; s = s + q1.w;
        s = s + q2.x; s = s + q2.y; s = s + q2.z; s = s + q2.w;
        s = s + q3.x; s = s + q3.y; s = s + q3.z; s = s + q3.w;
      }
      for (; i < cnt; ++i) s = s + a[i];
    }
    __syncwarp();
  }
  if (lane == 0) s_need = kCmdExit;
  __syncwarp();
  asm volatile("bar.sync 1, %0;" ::"r"(kWalkThreads));
  if (owner) {
    out[K] = s;
    replayed[K] = nreplay;
  }
}

int k_vol_com(const vt_map *m, double out[4]) {
  VT_REQUIRE_CTX();
  Ctx &c = ctx();
  CoordGeom cg;
  coord_geom(m->grid, cg);
  DevCoord dc;
  dc.ox = cg.origin[0]; dc.oy = cg.origin[1]; dc.oz = cg.origin[2];
  for (int k = 0; k < 3; ++k) { dc.xd[k] = cg.xd[k]; dc.yd[k] = cg.yd[k]; dc.zd[k] = cg.zd[k]; }
  dc.nx = cg.nx; dc.ny = cg.ny; dc.nz = cg.nz;
  const long n = m->n;
  out[0] = out[1] = out[2] = 0.0;
  out[3] = 0.0;
  if (n <= 0) return 0;
  // Chunk length: the tables cost 3 x 32 entries of 32 bytes per chunk, a replayed chunk L dependent adds.
  int L = 1024;
  if (const char *e = getenv("VOLTOOL_COM_CHUNK")) L = std::max(64, std::min(4096, atoi(e) & ~15));
  while (L < 4096 && (double)((n + L - 1) / L) * 3 * kBinades * sizeof(ComEntry) > 512.0e6) L *= 2;
  const long nchunks = (n + L - 1) / L;
  ComEntry *tab = nullptr;
  int *emin = nullptr;
  float *d_out = nullptr;
  // (every exit below releases all three: the pool never hands blocks back by itself)
  struct Scratch {
    ComEntry *&tab; int *&emin; float *&d_out;
    ~Scratch() { dev_free(tab); dev_free(emin); dev_free(d_out); }
  } scratch{tab, emin, d_out};
  if (dev_alloc((void **)&tab, sizeof(ComEntry) * (size_t)nchunks * 3 * kBinades)) return 2;
  if (dev_alloc((void **)&emin, sizeof(int) * (size_t)nchunks * 3)) return 2;
  if (dev_alloc((void **)&d_out, sizeof(float) * 8)) return 2;
  const size_t smem = sizeof(float) * 3 * (size_t)L;
  if (smem > 48 * 1024)
    VT_CUDA(cudaFuncSetAttribute(com_tables_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  prof_begin(PK_STATS);
  com_tables_kernel<<<(unsigned int)nchunks, 96, smem, c.stream>>>(m->d, n, dc, L, tab, emin);
  VT_LAUNCHED();
  if (smem > 16 * 1024)   // (the walk also holds ~27 KB of static shared memory)
    VT_CUDA(cudaFuncSetAttribute(com_walk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  com_walk_kernel<<<3, kWalkThreads, smem, c.stream>>>(m->d, n, dc, L, nchunks, tab, emin, d_out, (int *)(d_out + 4));
  prof_end(PK_STATS);
  VT_LAUNCHED();
  float h[8];
  VT_CUDA(cudaMemcpyAsync(h, d_out, sizeof(h), cudaMemcpyDeviceToHost, c.stream));
  VT_CUDA(cudaStreamSynchronize(c.stream));
  for (int k = 0; k < 3; ++k) out[k] = (double)h[k];
  double mass = 0.0;
  int rc = k_sum(m->d, n, &mass);
  if (rc) return rc;
  out[3] = mass;
  if (getenv("VOLTOOL_VERBOSE")) {
    const int *rp = (const int *)(h + 4);
    fprintf(stderr, "[voltool_gpu] vol_com: %ld chunks of %d, replayed %d/%d/%d\n", nchunks, L, rp[0], rp[1], rp[2]);
  }
  return 0;
}

}  // namespace vt
