// Device-side reset sampler: Environment.reset's draws (env.py:81-125, 150-170) on the env's own numpy legacy
// RandomState stream, evaluated by the lane group that owns the environment at the moment it resets.
//
// numpy.random.RandomState(seed) with an int seed is MT19937 seeded by init_genrand(seed); legacy
// randint(lo, hi) is lo + a masked-rejection draw in [0, hi-1-lo], one 32-bit output per try, and no draw at all when
// the range is a single value.  The 624-word generator state of every env lives in HBM (2 496 B per env); a group
// reads it LPE tempered outputs at a time (one coalesced load, then shuffles), twists it in place LPE words at a
// time when it runs out, and all lanes follow the same sequence of draws, so every drawn value is group-uniform.
// Draw order (must match the reference exactly): R robot cells from the SHRINKING row-major free
// list, then per package start, target (redrawn until != start), deadline slack, start_time (only for packages
// beyond the first min(R, 20) + 1); stable sort by start_time; package id = rank + 1.
//
// The host restatement of the same stream (sampler.cpp, mdx_sampler_* of include/mdx.h) is what tests pin this
// against, together with the reference-generated reset blocks in tests/golden/reset_blocks.npz.
//
// Included by mdx_kernels.cuh after Grp<LPE>.
#pragma once

namespace mdx {

__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}

template <int LPE>
struct DevMt {
  uint32_t* mt;    // this env's 624 state words (read / written through L2: __ldcg / __stcg)
  int pos;         // next output index
  int bbase;       // index of the first output held in `batch`, -1 = none
  uint32_t batch;  // tempered output bbase + sl

  // mt[k] <- mt[k + 397 mod 624] ^ f(mt[k], mt[k + 1 mod 624]), in place, strips of LPE words in ascending order: a strip reads
  // its three operands into registers before it writes, words k+1 / k+397 beyond the strip are still old, words
  // k-227 (and word 0 for k = 623) are already new -- the same dependences as the sequential loop.
  __device__ __forceinline__ void twist(const Grp<LPE>& g) {
#pragma unroll 1
    for (int k0 = 0; k0 < 624; k0 += LPE) {
      const int k = k0 + g.sl;
      uint32_t y = 0u, m = 0u;
      if (k < 624) {
        const uint32_t a = __ldcg(mt + k), b = __ldcg(mt + (k == 623 ? 0 : k + 1));
        m = __ldcg(mt + (k < 227 ? k + 397 : k - 227));
        y = (a & 0x80000000u) | (b & 0x7fffffffu);
      }
      g.sync();
      if (k < 624) __stcg(mt + k, m ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u));
      g.sync();
    }
  }

  __device__ __forceinline__ uint32_t next32(const Grp<LPE>& g) {
    if (pos == 624) {
      twist(g);
      pos = 0;
      bbase = -1;
    }
    if (bbase < 0 || pos >= bbase + LPE) {
      bbase = pos;
      const int idx = pos + g.sl;
      batch = idx < 624 ? mt_temper(__ldcg(mt + idx)) : 0u;
    }
    const uint32_t v = g.shfl(batch, pos - bbase);
    pos += 1;
    return v;
  }

  // legacy RandomState.randint(lo, hi), lo < hi (checked on the host at mdx_create)
  __device__ __forceinline__ int randint(const Grp<LPE>& g, int lo, int hi) {
    const uint32_t span = (uint32_t)(hi - 1 - lo);
    if (span == 0u) return lo;  // numpy does not touch the generator
    uint32_t mask = span;
    mask |= mask >> 1;
    mask |= mask >> 2;
    mask |= mask >> 4;
    mask |= mask >> 8;
    mask |= mask >> 16;
    uint32_t v;
    do {
      v = next32(g) & mask;
    } while (v > span);
    return lo + (int)v;
  }
};

// Draws the env's next reset block.  On return the group's shared-memory schedule `s_sched` ([4][Pp8] u16: start
// cell, target cell, start_time, deadline in package-id order, padding start_time = 0xFFFF) and the env's
// current-episode block p.cur hold it, the generator state in HBM is advanced, and the return value is the start
// cell of robot `sl` (undefined for sl >= R).  Every lane of the group must call it.
// `p` must be the kernel's __grid_constant__ parameter itself (its address is taken: a local copy would be forced into
// local memory for the whole kernel); the map is passed as the three scalars the sampler needs for the same reason.
template <int LPE>
__device__ __noinline__ int sample_episode(const KParams& p, const uint16_t* free_cells, int F, int map_rows, uint16_t* s_sched, int env) {
  const Grp<LPE> g((int)(threadIdx.x & 31));
  const int sl = g.sl;
  const int R = p.R, P = p.P, Pp8 = p.Pp8;
  DevMt<LPE> rng;
  rng.mt = p.mt + (size_t)env * 624;
  rng.pos = p.mtpos[env];
  rng.bbase = -1;
  rng.batch = 0u;
  // robots: index k into the shrinking free list = the k-th free-list entry not taken yet.  Lane i remembers the
  // entry robot i took; the k-th remaining entry is the fixed point of x = k + #{taken <= x} (monotone from x = k).
  int mine = -1;
#pragma unroll 1
  for (int i = 0; i < R; ++i) {
    const int k = rng.randint(g, 0, F - i);
    int x = k;
    while (true) {
      const int nx = k + __popc(g.ballot(mine >= 0 && mine <= x));
      if (nx == x) break;
      x = nx;
    }
    if (sl == i) mine = x;
  }
  const int cell = (sl < R) ? (int)__ldg(free_cells + mine) : 0;
  // packages in draw order; lane i % LPE files record i
  uint16_t* s_start = s_sched;
  uint16_t* s_target = s_sched + Pp8;
  uint16_t* s_t0 = s_sched + 2 * Pp8;
  uint16_t* s_dl = s_sched + 3 * Pp8;
  const int lim = R < 20 ? R : 20;
  const int lo_slack = map_rows / 2, hi_slack = 3 * map_rows;  // env.py:112, numpy truncates the float bound N/2
#pragma unroll 1
  for (int i = 0; i < P; ++i) {
    const int a = rng.randint(g, 0, F);
    int b;
    do {
      b = rng.randint(g, 0, F);
    } while (b == a);  // free cells are distinct, so target != start  <=>  other index
    const int slack = 10 + rng.randint(g, lo_slack, hi_slack);
    const int st = i <= lim ? 0 : rng.randint(g, 1, p.T);
    if (sl == (i & (LPE - 1))) {
      s_start[i] = __ldg(free_cells + a);
      s_target[i] = __ldg(free_cells + b);
      s_t0[i] = (uint16_t)st;
      s_dl[i] = (uint16_t)(st + slack);
    }
  }
  if (sl == 0) p.mtpos[env] = rng.pos;
  g.sync();
  // stable sort by start_time (env.py:119): rank = #{smaller} + #{equal with a lower draw index}; scatter into
  // the env's current-episode block, then reload it (sorted) into shared memory
  uint16_t* cur = p.cur + (size_t)env * 4 * Pp8;
#pragma unroll 1
  for (int j = sl; j < Pp8; j += LPE) {
    if (j < P) {
      const uint32_t key = ((uint32_t)s_t0[j] << 16) | (uint32_t)j;
      int rank = 0;
#pragma unroll 4
      for (int k = 0; k < P; ++k) rank += ((((uint32_t)s_t0[k] << 16) | (uint32_t)k) < key) ? 1 : 0;
      cur[rank] = s_start[j];
      cur[Pp8 + rank] = s_target[j];
      cur[2 * Pp8 + rank] = s_t0[j];
      cur[3 * Pp8 + rank] = s_dl[j];
    } else {  // padding never spawns
      cur[j] = 0;
      cur[Pp8 + j] = 0;
      cur[2 * Pp8 + j] = 0xffff;
      cur[3 * Pp8 + j] = 0;
    }
  }
  __threadfence_block();
  g.sync();
  {
    const uint4* src = reinterpret_cast<const uint4*>(cur);
    uint4* dst = reinterpret_cast<uint4*>(s_sched);
#pragma unroll 1
    for (int i = sl; i < (Pp8 >> 1); i += LPE) dst[i] = __ldcg(src + i);
  }
  g.sync();
  return cell;
}

}  // namespace mdx
