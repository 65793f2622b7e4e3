// Hand-written sm_100a kernels of the marl-delivery hot path.
//
// Mapping: ONE LANE GROUP PER ENVIRONMENT, robots are the group's lanes.  A group is a whole warp
// (LPE = 32 lanes per env, R <= 32), a half warp (LPE = 16, R <= 16) or a quarter warp (LPE = 8,
// R <= 8): with 5 robots most of the transition logic is warp-uniform control work, so packing 2-4
// environments into one warp amortises those instructions while the 128-bit stores still leave the
// warp 512 bytes at a time (LPE = 8: four 128-byte lines, one per env).  All warp collectives are
// issued on the group's member mask.
//
// The reference's sequential, robot-index-ordered move/collision resolution (env.py:203-250) is
// evaluated with __match_any_sync / __ballot_sync as a monotone fixed point; package scans, the shaped
// reward (MAPPO/helper.py:257-369) and the three featurisers (MAPPO/helper.py:6-255) run on the same
// group out of a shared-memory copy of the environment, and the feature tensors are written straight
// into the caller's (torch-owned) HBM tensors with 128-bit coalesced streaming stores.  0/1 planes live
// as bit masks and are expanded four cells at a time through a 16-entry nibble -> float4 table; sparse
// rows (gvec) are flushed as "data prefix + zero tail" segments; per-robot searches over packages run
// package-parallel with group reductions.
//
// Exactness: integers everywhere except (a) the env reward, accumulated in float64 in the reference's
// order (env.py:181,256,288-291), (b) the shaped reward, float32 adds in the reference's order of terms
// plus numpy's pairwise float32 sum, and (c) feature quotients, which are host-built float32 tables
// of correctly rounded float64 quotients (what `np.array([...python floats...], dtype=float32)`
// produces).  Sort keys are dense ranks of the reference's float64 keys, ties broken by the original
// list order like Python's stable sort.
#pragma once
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#include <cstdint>
#endif

#include "mdx_internal.h"

namespace mdx {

#define MDX_NONE32 0xffffffffu
// Resident CTAs per SM the featurising variants are compiled for (register cap 65536 / 128 / n).  Measured on B200
// (C3, 65 536 envs): 7 CTAs (72 regs) 251.6 M env-steps/s, 8 (64 regs) 268.4 M, 9 (56 regs, 4 B spill) 279.3 M, 10 (48 regs)
// 275.5 M -- the kernel is bound by dependent-instruction latency, so more resident warps win until spills bite.
#ifndef MDX_FEAT_BLOCKS
#define MDX_FEAT_BLOCKS 7  // generic (run-time shape) variants
#endif
#ifndef MDX_FIX_BLOCKS
#define MDX_FIX_BLOCKS 9   // shape-specialised variant (kFix = 1)
#endif
#ifndef MDX_FAT_BLOCKS
#define MDX_FAT_BLOCKS 5   // big-plane variant (kFix = 2): 96 registers.  Measured: C4 45.1 (4) / 47.9 (5) / 47.2 (6) vs 45.8 M env-steps/s
                           // for the 7-CTA build; C5 (shared memory allows 4 CTAs/SM) 13.6 / 13.7 / 13.8 vs 13.1 M
#endif
#ifndef MDX_JIT_BLOCKS
#define MDX_JIT_BLOCKS 7   // run-time-compiled variant (kFix = 3): chosen by mdx_jit.cpp from the instance's shared-memory footprint
#endif
#ifndef MDX_FIX_THREADS
#define MDX_FIX_THREADS 128  // CTA size of the shape-specialised variant (one warp per env)
#endif
constexpr int kLutBytes = 256;  // block-shared nibble -> float4 table at the start of dynamic smem

// A lane group: LPE consecutive lanes of a warp that own one environment.
template <int LPE>
struct Grp {
  int sl;          // lane index inside the group (= robot index for sl < R)
  int shift;       // first lane of the group inside the warp
  unsigned mask;   // member mask of the group
  __device__ __forceinline__ explicit Grp(int lane) {
    sl = lane & (LPE - 1);
    shift = lane & ~(LPE - 1);
    mask = LPE == 32 ? 0xffffffffu : (((1u << (LPE & 31)) - 1u) << shift);
  }
  static __device__ __forceinline__ unsigned low() { return LPE == 32 ? 0xffffffffu : ((1u << (LPE & 31)) - 1u); }
  __device__ __forceinline__ void sync() const { __syncwarp(mask); }
  __device__ __forceinline__ unsigned ballot(bool pr) const { return (__ballot_sync(mask, pr) >> shift) & low(); }  // group-relative bits
  __device__ __forceinline__ bool all(bool pr) const { return __all_sync(mask, pr); }
  __device__ __forceinline__ bool any(bool pr) const { return __any_sync(mask, pr); }
  __device__ __forceinline__ unsigned match(int v) const { return (__match_any_sync(mask, v) >> shift) & low(); }
  __device__ __forceinline__ uint32_t rmin(uint32_t v) const { return __reduce_min_sync(mask, v); }
  template <typename T>
  __device__ __forceinline__ T shfl(T v, int src) const { return __shfl_sync(mask, v, src, LPE); }  // src is group-relative
  template <typename T>
  __device__ __forceinline__ T shfl_xor(T v, int lane_mask) const { return __shfl_xor_sync(mask, v, lane_mask, LPE); }
};

}  // namespace mdx
#include "mdx_sampler.cuh"
namespace mdx {

// x / d through the host-built multiplier ceil(2^32 / d); d == 1 has no 32-bit multiplier and is encoded as 0
// (one robot, or one selection slot per robot).
__device__ __forceinline__ uint32_t div_magic(uint32_t x, uint32_t m) { return m ? __umulhi(x, m) : x; }

struct Env {
  uint16_t *start, *target, *t0, *dl;  // [Pp8] current episode schedule
  uint8_t* status;                     // [Pp16]
  uint16_t *wlist, *wrc, *wdl;         // compact waiting list: package index, start (r << 8 | c), remaining deadline
  uint16_t *alist, *sel;               // active list (gvec); selected waiting-list indices per robot (vec)
  uint32_t* wkey;                      // per-robot selection keys of the waiting list (generic path)
  uint32_t* planes;                    // [10][W32] grid, robots, waiting starts, active targets, 2 x (self, others, own target)
  uint16_t *rcell, *rcarry, *rtgt, *rdl, *rrc;  // [32] per robot
  float* stage;                        // 16-byte aligned staging for vec / gvec rows
  uint16_t *pid, *slot;                // tracker view only (kTrk): package id of every list slot; slot of every id (0xFFFF = absent)
};

constexpr uint32_t kNoSlot = 0xFFFFu;

// The featurisers and the shaped reward read packages through an Env "view".  Default (env truth): slot j IS package id j + 1.
// kTrk (MDX_CFG_STALE_TRACKER): the view is the trainer's dict in insertion order, ids are data.
template <bool kTrk>
__device__ __forceinline__ uint32_t slot_of(const Env& w, int id) { return kTrk ? (uint32_t)w.slot[id - 1] : (uint32_t)(id - 1); }
template <bool kTrk>
__device__ __forceinline__ int id_of(const Env& w, int slot) { return kTrk ? (int)w.pid[slot] : slot + 1; }

__device__ __forceinline__ Env carve(const KParams& p, unsigned char* base) {
  Env w;
  uint16_t* s = reinterpret_cast<uint16_t*>(base + p.off_sched);
  w.start = s;
  w.target = s + p.Pp8;
  w.t0 = s + 2 * p.Pp8;
  w.dl = s + 3 * p.Pp8;
  w.status = base + p.off_status;
  uint16_t* wl = reinterpret_cast<uint16_t*>(base + p.off_wlist);
  w.wlist = wl;
  w.wrc = wl + p.Pp8;
  w.wdl = wl + 2 * p.Pp8;
  w.alist = reinterpret_cast<uint16_t*>(base + p.off_alist);
  w.sel = reinterpret_cast<uint16_t*>(base + p.off_sel);
  w.wkey = reinterpret_cast<uint32_t*>(base + p.off_wkey);
  w.planes = reinterpret_cast<uint32_t*>(base + p.off_masks);
  uint16_t* ri = reinterpret_cast<uint16_t*>(base + p.off_rinfo);
  w.rcell = ri;
  w.rcarry = ri + 32;
  w.rtgt = ri + 64;
  w.rdl = ri + 96;
  w.rrc = ri + 128;
  w.stage = reinterpret_cast<float*>(base + p.off_stage);
  w.pid = nullptr;
  w.slot = nullptr;
  return w;
}

// The tracker view of the same working set: package arrays point at the tracker table, lists / planes / stage are shared.
__device__ __forceinline__ Env carve_tracker(const KParams& p, unsigned char* base, const Env& w) {
  Env f = w;
  uint16_t* s = reinterpret_cast<uint16_t*>(base + p.off_trk);
  f.pid = s;
  f.start = s + p.Pp8;
  f.target = s + 2 * p.Pp8;
  f.t0 = s + 3 * p.Pp8;
  f.dl = s + 4 * p.Pp8;
  f.status = base + p.off_trk + 10 * p.Pp8;
  f.slot = reinterpret_cast<uint16_t*>(base + p.off_trk + 10 * p.Pp8 + p.Pp16);
  return f;
}

// max(0, deadline - t) / T  and  t / T  (MAPPO/helper.py:107,155,163): table of correctly rounded
// float64 quotients, IEEE double division beyond the table (t keeps running past T).
static __device__ __noinline__ float tnorm_slow(int k, int T) { return __double2float_rn(__ddiv_rn((double)k, (double)T)); }
__device__ __forceinline__ float tnorm(const KParams& p, int k) {
  if (k < p.tq_len) return __ldg(p.tq + k);  // tq is all zeros when T <= 0
  return p.T > 0 ? tnorm_slow(k, p.T) : 0.0f;
}

// Schedule of the env's current episode -> shared memory (one LDG.128 per lane for 50 packages).
template <int LPE>
__device__ __forceinline__ void load_episode(const KParams& p, const Env& w, int env, int sl) {
  const uint4* src = reinterpret_cast<const uint4*>(p.cur + (size_t)env * 4 * p.Pp8);
  uint4* dst = reinterpret_cast<uint4*>(w.start);
  const int n16 = p.Pp8 >> 1;  // 4 rows * Pp8 u16 * 2 B / 16 B
#pragma unroll 1
  for (int i = sl; i < n16; i += LPE) dst[i] = __ldg(src + i);
}

// numpy float32 add.reduce over the R robot lanes (pairwise_sum: plain loop for n < 8, eight running
// partials for 8 <= n <= 128); every lane of the group returns the same value.
template <int LPE>
__device__ __forceinline__ float group_np_sum(const Grp<LPE>& g, float v, int R) {
  if (R < 8) {
    float s = 0.0f;
#pragma unroll 1
    for (int i = 0; i < R; ++i) s = __fadd_rn(s, g.shfl(v, i));
    return s;
  }
  float r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = g.shfl(v, j);
  int i = 8;
  for (; i < R - (R % 8); i += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], g.shfl(v, i + j));
  }
  float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                        __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
  for (; i < R; ++i) res = __fadd_rn(res, g.shfl(v, i));
  return res;
}

// Write n_total floats to global memory: the first n_data come from the staged row, the rest are
// zeros.  `stage` is 16-byte aligned, the row starts at stage[off] with off == (global float index
// & 3) so shared and global addresses are congruent modulo 16 bytes and the body moves as aligned
// float4; the stage must be zero for at least 4 floats past n_data.
template <int LPE>
__device__ __forceinline__ void flush_seg(float* gdst, const float* stage, int off, int n_data, int n_total, int sl) {
  const float* s = stage + off;
  int head = (4 - off) & 3;
  if (head > n_total) head = n_total;
  if (sl < head) __stcs(gdst + sl, sl < n_data ? s[sl] : 0.0f);
  const int body = (n_total - head) >> 2;
  const float4* s4 = reinterpret_cast<const float4*>(s + head);
  float4* g4 = reinterpret_cast<float4*>(gdst + head);
  int live = n_data > head ? (n_data - head + 3) >> 2 : 0;  // float4s that carry data
  if (live > body) live = body;
#pragma unroll 1
  for (int i = sl; i < live; i += LPE) __stcs(g4 + i, s4[i]);
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  int i = live + sl;
#pragma unroll 1
  for (; i + 3 * LPE < body; i += 4 * LPE) {  // four independent stores per trip: the zero tail is most of a gvec row
    __stcs(g4 + i, z4);
    __stcs(g4 + i + LPE, z4);
    __stcs(g4 + i + 2 * LPE, z4);
    __stcs(g4 + i + 3 * LPE, z4);
  }
#pragma unroll 1
  for (; i < body; i += LPE) __stcs(g4 + i, z4);
  const int done = head + (body << 2);
  if (sl < n_total - done) __stcs(gdst + done + sl, done + sl < n_data ? s[done + sl] : 0.0f);
}

template <int LPE>
__device__ __forceinline__ void zero_stage(float* stage, int n_floats, int sl) {
  float4* s4 = reinterpret_cast<float4*>(stage);
  const int n4 = (n_floats + 3) >> 2;
#pragma unroll 1
  for (int i = sl; i < n4; i += LPE) s4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// Compact list of packages that are 'waiting' with start_time <= t, in id order: package index and
// packed start position.  Returns the count.
template <int LPE>
__device__ __forceinline__ int build_waiting(const KParams& p, const Env& w, const Grp<LPE>& g, int t, int Pn) {
  int nw = 0;
  const unsigned lt = (1u << g.sl) - 1u;
#pragma unroll 1
  for (int base = 0; base < Pn; base += LPE) {
    const int j = base + g.sl;
    const bool wait = j < Pn && w.status[j] == ST_WAITING && (int)w.t0[j] <= t;
    const unsigned m = g.ballot(wait);
    if (wait) {
      const int k = nw + __popc(m & lt);
      const uint32_t sc = w.start[j];
      const uint32_t sr = __umulhi(sc, p.mul_wout);
      w.wlist[k] = (uint16_t)j;
      w.wrc[k] = (uint16_t)((sr << 8) | (sc - sr * p.W_out));
    }
    nw += __popc(m);
  }
  g.sync();
  return nw;
}

struct Scan {
  uint32_t best;   // (manhattan distance << 16 | list index) of the nearest waiting start, first minimum in id order
  uint32_t at_new; // lowest list index whose start is the robot's NEW cell, MDX_NONE32 if none
  bool near;       // some waiting start within manhattan distance 3 of the OLD cell
};

// For every robot: nearest waiting package from its old cell (MAPPO/helper.py:346-349), "waiting
// package within 3" (:360-365) and the lowest-id waiting package on its new cell (env.py:267-274,
// MAPPO/helper.py:312-317).  Package-parallel: lanes own list entries (64 / LPE of them in registers
// when P <= 64), one group reduction per robot.
// kBfs: distances are shortest-path lengths from the per-map BFS table (row of the robot's old cell) instead of
// Manhattan distances (opt-in shaping mode).
template <int LPE, bool kSmall, bool kBfs>
__device__ __forceinline__ Scan scan_waiting(const Env& w, const Grp<LPE>& g, int R, int nw, int r0, int c0, int r1, int c1,
                                             const uint16_t* bfs, int HW, int Wo) {
  constexpr int EPL = 64 / LPE;  // entries per lane when P <= 64
  Scan mine{MDX_NONE32, MDX_NONE32, false};
  if (nw == 0) return mine;
  const uint32_t rc1 = ((uint32_t)r1 << 8) | (uint32_t)c1;
  uint32_t e[EPL];
  if (kSmall) {
#pragma unroll
    for (int s = 0; s < EPL; ++s) e[s] = (g.sl + LPE * s < nw) ? (uint32_t)w.wrc[g.sl + LPE * s] : MDX_NONE32;
  }
#pragma unroll 1
  for (int i = 0; i < R; ++i) {
    const int ri = g.shfl(r0, i), ci = g.shfl(c0, i);
    const uint32_t rci = g.shfl(rc1, i);
    const uint16_t* brow = kBfs ? bfs + (size_t)(ri * Wo + ci) * HW : nullptr;
    uint32_t best = MDX_NONE32, at_new = MDX_NONE32;
    if (kSmall) {
#pragma unroll
      for (int s = 0; s < EPL; ++s) {
        if (e[s] != MDX_NONE32) {
          const uint32_t k = (uint32_t)(g.sl + LPE * s);
          const int d = kBfs ? (int)__ldg(brow + (e[s] >> 8) * Wo + (e[s] & 255u))
                             : abs((int)(e[s] >> 8) - ri) + abs((int)(e[s] & 255u) - ci);
          best = min(best, ((uint32_t)d << 16) | k);
          if (e[s] == rci) at_new = min(at_new, k);
        }
      }
    } else {
#pragma unroll 1
      for (int k = g.sl; k < nw; k += LPE) {
        const uint32_t rc = w.wrc[k];
        const int d = kBfs ? (int)__ldg(brow + (rc >> 8) * Wo + (rc & 255u))
                           : abs((int)(rc >> 8) - ri) + abs((int)(rc & 255u) - ci);
        best = min(best, ((uint32_t)d << 16) | (uint32_t)k);
        if (rc == rci) at_new = min(at_new, (uint32_t)k);
      }
    }
    best = g.rmin(best);
    at_new = g.rmin(at_new);
    if (g.sl == i) {
      mine.best = best;
      mine.at_new = at_new;
      mine.near = (best >> 16) <= 3u;  // the minimum distance decides "some waiting start within 3"
    }
  }
  return mine;
}

// Shaped team reward of one transition (MAPPO/helper.py:257-369).  Per lane (robot): old / new
// position and carried id, decoded action, the scan above; w.status still holds the PREVIOUS statuses.
template <int LPE, bool kBfs, bool kTrk = false>
__device__ __forceinline__ float shaped_team_reward(const KParams& p, const Env& w, const Grp<LPE>& g, int t_cur, int r0,
                                                    int c0, int k0, int r1, int c1, int k1, int mv, int op, int nw,
                                                    const Scan& sc, double env_reward, const uint16_t* bfs) {
  float acc = 0.0f;
  if (g.sl < p.R) {
    // "prev_pkg in persistent_packages_at_prev_state": env truth -> the package is waiting or in transit; tracker view -> the id has a slot
    const uint32_t s0 = k0 ? slot_of<kTrk>(w, k0) : kNoSlot;
    const uint8_t st0 = s0 != kNoSlot ? w.status[s0] : (uint8_t)ST_NONE;
    const bool tracked = s0 != kNoSlot && (st0 == ST_WAITING || st0 == ST_IN_TRANSIT);
    int gr = -1, gc = 0, dl0 = 0;
    if (s0 != kNoSlot) {
      const uint32_t tg = w.target[s0];
      gr = (int)__umulhi(tg, p.mul_wout);
      gc = (int)tg - gr * p.W_out;
      dl0 = w.dl[s0];
    }
    const bool at_target = k0 != 0 && r1 == gr && c1 == gc;
    // 1. successful pickup / delivery (:295-305)
    if (k0 == 0 && k1 != 0) {
      acc = __fadd_rn(acc, p.shaping[0]);
    } else if (k0 != 0 && k1 == 0) {
      if (tracked && at_target) acc = __fadd_rn(acc, (t_cur <= dl0) ? p.shaping[1] : p.shaping[2]);
    }
    // 2. wasted package ops (:308-327)
    if (op == 1) {
      if (k0 != 0) {
        acc = __fadd_rn(acc, p.shaping[4]);
      } else if (k1 == 0) {
        if (sc.at_new == MDX_NONE32) acc = __fadd_rn(acc, p.shaping[4]);
      }
    } else if (op == 2) {
      if (k0 == 0) {
        acc = __fadd_rn(acc, p.shaping[5]);
      } else if (k1 != 0) {
        if (tracked && !at_target) acc = __fadd_rn(acc, p.shaping[5]);
      }
    }
    // 3. movement (:330-356)
    const bool moved = r0 != r1 || c0 != c1;
    if (mv != MV_S && !moved) acc = __fadd_rn(acc, p.shaping[6]);
    bool have_goal = tracked;
    if (!tracked && nw > 0) {  // nearest waiting start, first minimum in id order
      const uint32_t rc = w.wrc[sc.best & 0xffffu];
      gr = (int)(rc >> 8);
      gc = (int)(rc & 255u);
      have_goal = true;
    }
    if (have_goal && moved) {
      int d_before, d_after;
      if (kBfs) {
        const uint16_t* brow = bfs + (size_t)(gr * p.W_out + gc) * p.HW;
        d_before = (int)__ldg(brow + r0 * p.W_out + c0);
        d_after = (int)__ldg(brow + r1 * p.W_out + c1);
      } else {
        d_before = abs(r0 - gr) + abs(c0 - gc);
        d_after = abs(r1 - gr) + abs(c1 - gc);
      }
      if (d_after < d_before)
        acc = __fadd_rn(acc, p.shaping[3]);
      else if (d_after > d_before)
        acc = __fadd_rn(acc, p.shaping[8]);
    }
    // 4. idle next to work (:359-367)
    if (!moved && mv == MV_S && k0 == 0 && sc.near) acc = __fadd_rn(acc, p.shaping[7]);
  }
  const float s = group_np_sum<LPE>(g, acc, p.R);
  // python float + np.float32 -> float32 under numpy >= 2
  return __fadd_rn(__double2float_rn(env_reward), s);
}

// IDQ's per-agent reward_shaping (IDQ/networks.py:228-349) for one robot, Python-float accumulation in the
// reference's order.  w.status / the waiting list still describe the state BEFORE the transition.
__device__ __forceinline__ double idq_agent_reward(const Env& w, int nw, bool moved, int op, int r0, int c0, int k0, int k1,
                                                   bool at_target, int t_cur, int dl0, int ops_int) {
  double acc = 0.0;
  if (!moved) acc = __dadd_rn(acc, -0.1);  // SHAPING_STAY_PENALTY (:283-284)
  if (ops_int) {  // with the trainer's string ops `package_op == 1` is never true (IDQ/trainer.py:276-291)
    if (op == 1) {  // :300-318
      if (k0 == 0 && k1 != 0) {
        acc = __dadd_rn(acc, 2.0);
      } else if (k0 != 0) {
        acc = __dadd_rn(acc, -0.1);
      } else {  // failed pickup: was a waiting package on the robot's PREVIOUS cell?
        const uint32_t rc_old = ((uint32_t)r0 << 8) | (uint32_t)c0;
        bool avail = false;
        for (int k = 0; k < nw; ++k) avail = avail || (uint32_t)w.wrc[k] == rc_old;
        if (!avail) acc = __dadd_rn(acc, -0.1);
      }
    } else if (op == 2) {  // :321-347 (a wasted drop costs 0)
      if (k0 != 0 && k1 == 0) {
        const uint8_t st0 = w.status[k0 - 1];
        if ((st0 == ST_WAITING || st0 == ST_IN_TRANSIT) && at_target) {
          acc = __dadd_rn(acc, 10.0);
          if (t_cur > dl0) acc = __dadd_rn(acc, -5.0);
        }
      }
    }
  }
  return acc;
}

__device__ __forceinline__ void decode_action(int a, int& mv, int& op) {
  if (a < 15) {
    mv = a % 5;
    op = a / 5;
  } else if (a < 18) {
    mv = MV_BAD;
    op = a - 15;
  } else {
    mv = MV_BAD;
    op = 0;
  }
}

// ------------------------------------------------------------------------------------------------
// Featurisation of the env held in shared memory (statuses, schedule) and registers (cell / carry of
// robot `sl`), MAPPO/helper.py:6-255.  `row` selects the output row.
//   kS     > 0: a plane is at most kS slices of LPE float4 (H*W % 4 == 0, H*W <= 4*LPE*kS); the four
//               shared planes' nibbles are hoisted into registers and every robot costs 6 table loads +
//               6 stores per slice.
//          = 0: generic flat loops over per-robot bit planes (any H*W).
//   kSmall:     P <= 64, so a lane owns at most 64/LPE waiting-list entries and keeps their selection
//               keys in registers.
//   kTrk:       `w` is the tracker view (dict order, Pn entries, ids are data); otherwise the env truth (Pn = P, slot j = id j + 1).
template <int LPE, int kS, bool kSmall, int kFix, bool kTrk = false>
__device__ __forceinline__ void featurize_env(const KParams& p, const Env& w, const Grp<LPE>& g, const float4* lut,
                                              const MapDev& map, int row, int t, int cell, int carry, int Pn) {
  constexpr int EPL = 64 / LPE;
  const int sl = g.sl;
  const int R = p.R, P = Pn, W32 = p.W32, HW = p.HW, Wo = p.W_out;
  const int Hm1 = p.H_out - 1, Wm1 = p.W_out - 1, NC = 2 * p.W_out - 1;
  const unsigned lt = (1u << sl) - 1u;
  const bool active = sl < R;
  const int r = active ? (int)__umulhi((uint32_t)cell, p.mul_wout) : 0;
  const int c = active ? cell - r * Wo : 0;
  const float* qr = map.qr;
  const float* qc = map.qc;

  // ---- per-robot info, shared planes, waiting / active lists -----------------------------------
  bool carrying_it = false;
  int own_tgt = 0xFFFF, own_rem = 0, tr_own = 0, tc_own = 0;
  if (active) {
    const uint32_t sc_ = carry != 0 ? slot_of<kTrk>(w, carry) : kNoSlot;
    if (sc_ != kNoSlot && w.status[sc_] == ST_IN_TRANSIT) {
      carrying_it = true;
      own_tgt = w.target[sc_];
      own_rem = max(0, (int)w.dl[sc_] - t);
      tr_own = (int)__umulhi((uint32_t)own_tgt, p.mul_wout);
      tc_own = own_tgt - tr_own * Wo;
    }
    w.rcell[sl] = (uint16_t)cell;
    w.rcarry[sl] = (uint16_t)carry;
    w.rtgt[sl] = (uint16_t)own_tgt;
    w.rdl[sl] = (uint16_t)own_rem;
    w.rrc[sl] = (uint16_t)((r << 8) | c);
  }
#pragma unroll 1
  for (int i = sl; i < W32; i += LPE) {
    w.planes[i] = __ldg(map.grid_bits + i);
    w.planes[W32 + i] = 0u;
    w.planes[2 * W32 + i] = 0u;
    w.planes[3 * W32 + i] = 0u;
  }
  g.sync();
  if (active) atomicOr(&w.planes[W32 + (cell >> 5)], 1u << (cell & 31));
  int nw = 0, na = 0;
#pragma unroll 1
  for (int base = 0; base < P; base += LPE) {
    const int j = base + sl;
    const uint8_t st = j < P ? w.status[j] : (uint8_t)ST_NONE;
    const bool wait = j < P && st == ST_WAITING && (int)w.t0[j] <= t;
    const bool act = wait || (j < P && st == ST_IN_TRANSIT);
    uint32_t sc = 0;
    if (wait) {
      sc = w.start[j];
      atomicOr(&w.planes[2 * W32 + (sc >> 5)], 1u << (sc & 31));
    }
    if (act) {
      const int tc = w.target[j];
      atomicOr(&w.planes[3 * W32 + (tc >> 5)], 1u << (tc & 31));
    }
    const unsigned wm = g.ballot(wait), am = g.ballot(act);
    if (wait) {
      const int k = nw + __popc(wm & lt);
      const uint32_t sr = __umulhi(sc, p.mul_wout);
      w.wlist[k] = (uint16_t)j;
      w.wrc[k] = (uint16_t)((sr << 8) | (sc - sr * Wo));
      w.wdl[k] = (uint16_t)max(0, (int)w.dl[j] - t);
    }
    if (!kTrk && act) {
      const int k = na + __popc(am & lt);
      if (k < p.MP) w.alist[k] = (uint16_t)j;
    }
    nw += __popc(wm);
    na += __popc(am);
  }
  if (kTrk) {
    // convert_global_state lists the active packages sorted by ID (MAPPO/helper.py:229), not in dict order: walk the ids
    g.sync();
    na = 0;
#pragma unroll 1
    for (int base = 0; base < p.P; base += LPE) {
      const int id0 = base + sl;
      const uint32_t s_ = id0 < p.P ? (uint32_t)w.slot[id0] : kNoSlot;
      const bool act = s_ != kNoSlot && ((w.status[s_] == ST_WAITING && (int)w.t0[s_] <= t) || w.status[s_] == ST_IN_TRANSIT);
      const unsigned am = g.ballot(act);
      if (act) {
        const int k = na + __popc(am & lt);
        if (k < p.MP) w.alist[k] = (uint16_t)s_;
      }
      na += __popc(am);
    }
  }
  g.sync();

  if (kS > 0) {
    // ---- plane-structured path: lane owns float4 number (sl + LPE*s) of every plane -------------
    constexpr int NS = kS > 0 ? kS : 1;
    const int n4p = HW >> 2;
    uint32_t ng[NS], nrb[NS], nws[NS], nat[NS];
#pragma unroll
    for (int s_ = 0; s_ < NS; ++s_) {
      const int q = sl + LPE * s_;
      const int wi = q >> 3;
      const uint32_t sh = (uint32_t)(q & 7) << 2;
      const bool ok = q < n4p;
      ng[s_] = ok ? (w.planes[wi] >> sh) & 15u : 0u;
      nrb[s_] = ok ? (w.planes[W32 + wi] >> sh) & 15u : 0u;
      nws[s_] = ok ? (w.planes[2 * W32 + wi] >> sh) & 15u : 0u;
      nat[s_] = ok ? (w.planes[3 * W32 + wi] >> sh) & 15u : 0u;
    }
    // (Round 2 tried the obvious alternative -- expand the planes all robots share once per slice into registers and build the three
    // per-robot planes with compares / selects, so that no store waits for a table load: 12-15 % SLOWER on C4 / C5 / K = 100
    // (profiles/r2_experiments.txt).  Stores hold their data AND address registers until the LSU has read them; a loop whose
    // next iteration rewrites those registers with ALU results stalls on that release (long scoreboard) every iteration, whereas
    // the table loads below are themselves scoreboarded and overlap it.)
    if (p.gstate != nullptr) {
      float4* out4 = reinterpret_cast<float4*>(p.gstate + (size_t)row * 4 * HW);
#pragma unroll
      for (int s_ = 0; s_ < NS; ++s_) {
        const int q = sl + LPE * s_;
        if (q < n4p) {
          __stcs(out4 + q, lut[ng[s_]]);
          __stcs(out4 + n4p + q, lut[nrb[s_]]);
          __stcs(out4 + 2 * n4p + q, lut[nws[s_]]);
          __stcs(out4 + 3 * n4p + q, lut[nat[s_]]);
        }
      }
    }
    if (p.obs != nullptr) {
      float4* base4 = reinterpret_cast<float4*>(p.obs + (size_t)row * R * 6 * HW);
#pragma unroll 1
      for (int i = 0; i < R; ++i) {
        const uint32_t rc = w.rcell[i], tg = w.rtgt[i];  // tg = 0xFFFF (none) never matches: tg >> 2 > any q
        float4* out4 = base4 + (size_t)i * 6 * n4p;
#pragma unroll
        for (int s_ = 0; s_ < NS; ++s_) {
          const int q = sl + LPE * s_;
          if (q < n4p) {
            const uint32_t oh_r = ((rc >> 2) == (uint32_t)q) ? (1u << (rc & 3u)) : 0u;
            const uint32_t oh_t = ((tg >> 2) == (uint32_t)q) ? (1u << (tg & 3u)) : 0u;
            __stcs(out4 + q, lut[ng[s_]]);                        // 0: obstacles
            __stcs(out4 + n4p + q, lut[oh_r]);                    // 1: this robot
            __stcs(out4 + 2 * n4p + q, lut[nrb[s_] & ~oh_r]);     // 2: other robots
            __stcs(out4 + 3 * n4p + q, lut[nws[s_]]);             // 3: waiting starts
            __stcs(out4 + 4 * n4p + q, lut[nat[s_]]);             // 4: active targets
            __stcs(out4 + 5 * n4p + q, lut[oh_t]);                // 5: own carried target
          }
        }
      }
    }
  } else {
    // ---- generic path: flat loops, per robot three private planes (double-buffered) ---------------
    if (p.gstate != nullptr) {
      float* out = p.gstate + (size_t)row * 4 * HW;
      if ((HW & 3) == 0) {
        float4* out4 = reinterpret_cast<float4*>(out);
#pragma unroll 1
        for (int q = sl; q < HW; q += LPE) {  // 4 planes * HW / 4 float4s
          const uint32_t e = (uint32_t)q << 2;
          const uint32_t ch = __umulhi(e, p.mul_plane);
          const uint32_t cl = e - ch * HW;
          const uint32_t nib = (w.planes[ch * W32 + (cl >> 5)] >> (cl & 31u)) & 15u;
          __stcs(out4 + q, lut[nib]);
        }
      } else {
#pragma unroll 1
        for (int e = sl; e < 4 * HW; e += LPE) {
          const uint32_t ch = __umulhi((uint32_t)e, p.mul_plane);
          const uint32_t cl = (uint32_t)e - ch * HW;
          __stcs(out + e, ((w.planes[ch * W32 + (cl >> 5)] >> (cl & 31u)) & 1u) ? 1.0f : 0.0f);
        }
      }
    }
    if (p.obs != nullptr) {
#pragma unroll 1
      for (int i = 0; i < R; ++i) {
        const int b = 4 + 3 * (i & 1);
        const uint32_t rc = w.rcell[i], tg = w.rtgt[i];
#pragma unroll 1
        for (int x = sl; x < W32; x += LPE) {
          const uint32_t self = ((rc >> 5) == (uint32_t)x) ? (1u << (rc & 31u)) : 0u;
          w.planes[b * W32 + x] = self;
          w.planes[(b + 1) * W32 + x] = w.planes[W32 + x] & ~self;
          w.planes[(b + 2) * W32 + x] = ((tg >> 5) == (uint32_t)x) ? (1u << (tg & 31u)) : 0u;  // tg = 0xFFFF: no word matches
        }
        g.sync();
        // channel -> plane: 0 grid, 1 self, 2 others, 3 waiting starts, 4 active targets, 5 own target
        const uint32_t pmap = 0u | ((uint32_t)b << 4) | ((uint32_t)(b + 1) << 8) | (2u << 12) | (3u << 16) | ((uint32_t)(b + 2) << 20);
        float* out = p.obs + ((size_t)row * R + i) * 6 * HW;
        if ((HW & 3) == 0) {
          float4* out4 = reinterpret_cast<float4*>(out);
          const int n4 = (6 * HW) >> 2;
#pragma unroll 1
          for (int q = sl; q < n4; q += LPE) {
            const uint32_t e = (uint32_t)q << 2;
            const uint32_t ch = __umulhi(e, p.mul_plane);
            const uint32_t cl = e - ch * HW;
            const uint32_t pl = (pmap >> (ch << 2)) & 15u;
            const uint32_t nib = (w.planes[pl * W32 + (cl >> 5)] >> (cl & 31u)) & 15u;
            __stcs(out4 + q, lut[nib]);
          }
        } else {
#pragma unroll 1
          for (int e = sl; e < 6 * HW; e += LPE) {
            const uint32_t ch = __umulhi((uint32_t)e, p.mul_plane);
            const uint32_t cl = (uint32_t)e - ch * HW;
            const uint32_t pl = (pmap >> (ch << 2)) & 15u;
            __stcs(out + e, ((w.planes[pl * W32 + (cl >> 5)] >> (cl & 31u)) & 1u) ? 1.0f : 0.0f);
          }
        }
      }
    }
  }

  // ---- vec [R, D]  (generate_vector_features, MAPPO/helper.py:68-165) -------------------------
  if (p.vec != nullptr) {
    const int D = p.D, K_o = p.K_o, Ksel = p.sel_k;
    const float tfrac = tnorm(p, t);
    // an other-robot row ends with "target relative to the carrier itself" (:125-136)
    float ot_r = 0.0f, ot_c = 0.0f;
    if (carrying_it) {
      ot_r = __ldg(qr + (tr_own - r + Hm1));
      ot_c = __ldg(qc + (tc_own - c + Wm1));
    }
    const float cflag = carry != 0 ? 1.0f : 0.0f;
    const uint32_t wmask = (1u << p.key_wbits) - 1u;
    const int nsel = min(Ksel, nw);
    const int rpf = p.vec_rows_per_flush;
    // lanes as (robot, other) pairs for the stable sort of the other robots
    const int gi = (int)div_magic((uint32_t)sl, p.mul_r), gj = sl - gi * R;
    const int G = (int)div_magic((uint32_t)LPE, p.mul_r);  // robots per pass = LPE / R
    // winners of the package selection are kept in lanes when a robot's slots fit the group
    const bool lane_sel = kSmall && Ksel <= LPE;
    const int RP = lane_sel ? (int)div_magic((uint32_t)LPE, p.mul_ksel) : 1;  // robots per pass = LPE / Ksel
    const int sel_ii = (int)div_magic((uint32_t)sl, p.mul_ksel), sel_q = sl - sel_ii * Ksel;
    // Few slots on a long waiting list (C4 / C5: K_p <= 5, P > 64).  The selection order is (remaining deadline, distance rank,
    // list index) and the deadline does not depend on the robot, so only entries whose remaining deadline is <= the nsel-th
    // smallest one (with multiplicity) can ever be selected, for ANY robot.  That candidate set is found once per env; while it
    // fits the group (it nearly always does: it only grows with ties, i.e. many packages past their deadline), every robot costs
    // one key per lane and nsel arg-min rounds instead of a pass over the whole list.
    bool use_cand = false;
    uint32_t c_k = MDX_NONE32, c_dlk = 0u;
    int c_off = 0;
    if constexpr (!kSmall && kFix != 1) {
      if (Ksel <= 5 && nsel > 0 && nw > LPE) {
        uint32_t d0 = MDX_NONE32, d1 = MDX_NONE32, d2 = MDX_NONE32, d3 = MDX_NONE32, d4 = MDX_NONE32;
#pragma unroll 2
        for (int k = sl; k < nw; k += LPE) {
          uint32_t x = w.wdl[k], lo;
          lo = min(d0, x), x = max(d0, x), d0 = lo;
          lo = min(d1, x), x = max(d1, x), d1 = lo;
          lo = min(d2, x), x = max(d2, x), d2 = lo;
          lo = min(d3, x), x = max(d3, x), d3 = lo;
          d4 = min(d4, x);
        }
        uint32_t thr = 0u;
#pragma unroll 1
        for (int q = 0; q < nsel; ++q) {  // pop the nsel smallest deadlines one at a time (ties: the lowest lane pops)
          thr = g.rmin(d0);
          const unsigned m = g.ballot(d0 == thr);
          if (sl == __ffs(m) - 1) d0 = d1, d1 = d2, d2 = d3, d3 = d4, d4 = MDX_NONE32;
        }
        uint16_t* cand = w.sel + ((rpf * Ksel + 7) & ~7);
        int ncand = 0;
#pragma unroll 1
        for (int base = 0; base < nw; base += LPE) {
          const int k = base + sl;
          const bool is = k < nw && (uint32_t)w.wdl[k] <= thr;
          const unsigned m = g.ballot(is);
          if (is) {
            const int pos = ncand + __popc(m & lt);
            if (pos < LPE) cand[pos] = (uint16_t)k;
          }
          ncand += __popc(m);
        }
        g.sync();
        use_cand = ncand <= LPE;
        if (use_cand && sl < ncand) {
          c_k = cand[sl];
          const uint32_t rc = w.wrc[c_k];
          c_off = (int)(rc >> 8) * NC + (int)(rc & 255u);
          c_dlk = (uint32_t)w.wdl[c_k] * (uint32_t)p.nranks_max;
        }
      }
    }
#pragma unroll 1
    for (int i0 = 0; i0 < R; i0 += rpf) {
      const int nrows = min(rpf, R - i0);
      const size_t g0 = ((size_t)row * R + i0) * D;
      const int off = (int)(g0 & 3);
      zero_stage<LPE>(w.stage, nrows * D + 4, sl);
      g.sync();
      float* chunk = w.stage + off;
      float* pk_base = chunk + 6 + K_o * 5;  // first package slot of the first staged row
      // own block (:93-112) and global time (:163)
      if (sl >= i0 && sl < i0 + nrows) {
        float* rowp = chunk + (sl - i0) * D;
        rowp[0] = __ldg(qr + (r + Hm1));
        rowp[1] = __ldg(qc + (c + Wm1));
        rowp[2] = cflag;
        if (carrying_it) {
          rowp[3] = ot_r;  // (tr - my_r) / n_rows: same quotient as in the other-robot rows
          rowp[4] = ot_c;
          rowp[5] = tnorm(p, own_rem);
        }
        rowp[D - 1] = tfrac;
      }
      // other robots, stable-sorted by the float64 key rank (:115-141)
      if (kFix != 1 && LPE == 32 && K_o > 0 && G == 1) {
        // More than 16 robots: one robot per pass, lane j holds the key of "other robot j".  Counting the smaller keys costs 32
        // shuffles + compares per pass (20 % of C5's instructions); a bitonic network sorts the 32 keys in 15 shuffle + min/max steps
        // instead, after which lane p holds the p-th nearest other robot (its index sits in the key's low 5 bits) and writes row p.
        uint32_t dirmask = 0u;  // bit s: this lane keeps the MIN in compare-exchange step s
        {
          int st_ = 0;
#pragma unroll
          for (int k = 2; k <= 32; k <<= 1)
#pragma unroll
            for (int j = k >> 1; j > 0; j >>= 1, ++st_) dirmask |= ((((sl & j) == 0) == ((sl & k) == 0)) ? 1u : 0u) << st_;
        }
        // Two robots per pass: their networks are independent chains of shuffles, issued back to back (each step waits ~25 cycles
        // for its shuffle; C5 spent 11 % of its warp time here with one chain).
        const uint32_t rcj_own = sl < R ? w.rrc[sl] : 0u;
        auto key_of = [&](int i, uint32_t rci) -> uint32_t {
          if (sl >= R || sl == i) return MDX_NONE32;
          const int dr = (int)(rcj_own >> 8) - (int)(rci >> 8), dc = (int)(rcj_own & 255u) - (int)(rci & 255u);
          return ((uint32_t)__ldg(map.rank + (dr + Hm1) * NC + (dc + Wm1)) << 5) | (uint32_t)sl;
        };
        auto emit = [&](int i, uint32_t rci, uint32_t sk) {
          const int jj = (int)(sk & 31u);
          const float cf_j = g.shfl(cflag, jj);
          const float otr_j = g.shfl(ot_r, jj), otc_j = g.shfl(ot_c, jj);
          if (sk != MDX_NONE32 && sl < K_o) {
            const uint32_t rcj = w.rrc[jj];
            const int dr = (int)(rcj >> 8) - (int)(rci >> 8), dc = (int)(rcj & 255u) - (int)(rci & 255u);
            float* o = chunk + (i - i0) * D + 6 + sl * 5;
            o[0] = __ldg(qr + (dr + Hm1));
            o[1] = __ldg(qc + (dc + Wm1));
            o[2] = cf_j;
            o[3] = otr_j;
            o[4] = otc_j;
          }
        };
#pragma unroll 1
        for (int i = i0; i < i0 + nrows; i += 2) {
          const bool two = i + 1 < i0 + nrows;
          const uint32_t rci = w.rrc[i], rci1 = two ? w.rrc[i + 1] : 0u;
          uint32_t sk = key_of(i, rci), sk1 = two ? key_of(i + 1, rci1) : MDX_NONE32;
          {
            int st_ = 0;
#pragma unroll
            for (int k = 2; k <= 32; k <<= 1)
#pragma unroll
              for (int j = k >> 1; j > 0; j >>= 1, ++st_) {
                const uint32_t other = g.shfl_xor(sk, j), other1 = g.shfl_xor(sk1, j);
                const bool keep_min = (dirmask >> st_) & 1u;
                sk = keep_min ? min(sk, other) : max(sk, other);
                sk1 = keep_min ? min(sk1, other1) : max(sk1, other1);
              }
          }
          emit(i, rci, sk);
          if (two) emit(i + 1, rci1, sk1);
        }
      } else if (K_o > 0 && R > 1) {
#pragma unroll 1
        for (int ib = i0; ib < i0 + nrows; ib += G) {
          const int i = ib + gi;
          const bool valid = gi < G && i < i0 + nrows && gj != i;
          uint32_t okey = MDX_NONE32;
          int dr = 0, dc = 0;
          if (valid) {
            const uint32_t rci = w.rrc[i], rcj = w.rrc[gj];
            dr = (int)(rcj >> 8) - (int)(rci >> 8);
            dc = (int)(rcj & 255u) - (int)(rci & 255u);
            okey = ((uint32_t)__ldg(map.rank + (dr + Hm1) * NC + (dc + Wm1)) << 5) | (uint32_t)gj;
          }
          int pos = 0;
          const int kbase = gi * R;
#pragma unroll 8
          for (int k = 0; k < R; ++k) pos += (g.shfl(okey, kbase + k) < okey) ? 1 : 0;
          const float cf_j = g.shfl(cflag, gj);
          const float otr_j = g.shfl(ot_r, gj), otc_j = g.shfl(ot_c, gj);
          if (valid && pos < K_o) {
            float* o = chunk + (i - i0) * D + 6 + pos * 5;
            o[0] = __ldg(qr + (dr + Hm1));
            o[1] = __ldg(qc + (dc + Wm1));
            o[2] = cf_j;
            o[3] = otr_j;
            o[4] = otc_j;
          }
        }
      }
      // waiting packages, stable-sorted by (remaining deadline, key rank) (:143-160): per robot,
      // min(K_p, n_waiting) rounds of a group arg-min over the unique keys
      //   (remaining deadline * nranks + rank of the float64 distance key) << wbits | list index;
      // the winner of round q for the ii-th robot of a pass is remembered by lane ii*Ksel+q, which then
      // writes that row.
#ifndef MDX_J_INTERLEAVE
#define MDX_J_INTERLEAVE 0  // run-time-compiled variant: 1 = this shape qualifies for the interleaved selection (mdx_jit.cpp decides)
#define MDX_J_NR 5
#endif
      if ((kFix == 1 || (kFix == 3 && MDX_J_INTERLEAVE)) && nsel > 0) {
        // Shape-specialised selection (baked C3: 5 robots, 5 slots; run-time-compiled: any R <= 8 robots with R * slots <= 32, P <= 64,
        // all rows staged at once): the rounds of all robots run interleaved -- R independent group reductions in flight per round
        // instead of R * slots dependent ones -- and a lane holds ONE key per robot while the waiting list fits the group (nw <= LPE).
        constexpr int NR = kFix == 1 ? 5 : MDX_J_NR;
        int bi[NR];
#pragma unroll
        for (int ii = 0; ii < NR; ++ii) {
          const uint32_t rci = w.rrc[ii];
          bi[ii] = (Hm1 - (int)(rci >> 8)) * NC + (Wm1 - (int)(rci & 255u));
        }
        uint32_t mysel = 0;
        if (nw <= LPE) {
          uint32_t key[NR];
          const bool has = sl < nw;
          int off = 0;
          uint32_t dlk = 0;
          if (has) {
            const uint32_t rc = w.wrc[sl];
            off = (int)(rc >> 8) * NC + (int)(rc & 255u);
            dlk = (uint32_t)w.wdl[sl] * (uint32_t)p.nranks_max;
          }
#pragma unroll
          for (int ii = 0; ii < NR; ++ii)
            key[ii] = has ? (((dlk + (uint32_t)__ldg(map.rank + off + bi[ii])) << p.key_wbits) | (uint32_t)sl) : MDX_NONE32;
#pragma unroll 1
          for (int q = 0; q < nsel; ++q) {
            uint32_t b[NR];
#pragma unroll
            for (int ii = 0; ii < NR; ++ii) b[ii] = g.rmin(key[ii]);
#pragma unroll
            for (int ii = 0; ii < NR; ++ii)
              if (key[ii] == b[ii]) key[ii] = MDX_NONE32;  // keys are unique: exactly one entry is retired
            uint32_t pick = b[0];
#pragma unroll
            for (int ii = 1; ii < NR; ++ii)
              if (sel_ii == ii) pick = b[ii];
            if (sel_q == q) mysel = pick & wmask;
          }
        } else {
          uint32_t k0[NR], k1[NR];
          const bool has1 = sl + LPE < nw;  // nw > LPE: every lane owns entry sl
          const uint32_t rc0 = w.wrc[sl];
          const int off0 = (int)(rc0 >> 8) * NC + (int)(rc0 & 255u);
          const uint32_t dl0k = (uint32_t)w.wdl[sl] * (uint32_t)p.nranks_max;
          int off1 = 0;
          uint32_t dl1k = 0;
          if (has1) {
            const uint32_t rc1 = w.wrc[sl + LPE];
            off1 = (int)(rc1 >> 8) * NC + (int)(rc1 & 255u);
            dl1k = (uint32_t)w.wdl[sl + LPE] * (uint32_t)p.nranks_max;
          }
#pragma unroll
          for (int ii = 0; ii < NR; ++ii) {
            k0[ii] = ((dl0k + (uint32_t)__ldg(map.rank + off0 + bi[ii])) << p.key_wbits) | (uint32_t)sl;
            k1[ii] = has1 ? (((dl1k + (uint32_t)__ldg(map.rank + off1 + bi[ii])) << p.key_wbits) | (uint32_t)(sl + LPE)) : MDX_NONE32;
          }
#pragma unroll 1
          for (int q = 0; q < nsel; ++q) {
            uint32_t b[NR];
#pragma unroll
            for (int ii = 0; ii < NR; ++ii) b[ii] = g.rmin(min(k0[ii], k1[ii]));
#pragma unroll
            for (int ii = 0; ii < NR; ++ii) {
              if (k0[ii] == b[ii]) k0[ii] = MDX_NONE32;
              if (k1[ii] == b[ii]) k1[ii] = MDX_NONE32;
            }
            uint32_t pick = b[0];
#pragma unroll
            for (int ii = 1; ii < NR; ++ii)
              if (sel_ii == ii) pick = b[ii];
            if (sel_q == q) mysel = pick & wmask;
          }
        }
        if (sel_ii < NR && sel_q < nsel) {
          const uint32_t rci = w.rrc[sel_ii];
          const int ri = (int)(rci >> 8), ci = (int)(rci & 255u);
          const int k = (int)mysel;
          const int j = w.wlist[k];
          const uint32_t rc = w.wrc[k];
          const uint32_t tc_ = w.target[j];
          const int tr = (int)__umulhi(tc_, p.mul_wout), tcol = (int)tc_ - tr * Wo;
          float* o = pk_base + sel_ii * D + sel_q * 5;
          o[0] = __ldg(qr + ((int)(rc >> 8) - ri + Hm1));
          o[1] = __ldg(qc + ((int)(rc & 255u) - ci + Wm1));
          o[2] = __ldg(qr + (tr - ri + Hm1));
          o[3] = __ldg(qc + (tcol - ci + Wm1));
          o[4] = tnorm(p, (int)w.wdl[k]);
        }
      } else if (nsel > 0 && lane_sel) {
#pragma unroll 1
        for (int ib = i0; ib < i0 + nrows; ib += RP) {
          const int npass = min(RP, i0 + nrows - ib);
          uint32_t mysel = 0;
#pragma unroll 1
          for (int ii = 0; ii < npass; ++ii) {
            const uint32_t rci = w.rrc[ib + ii];
            const int bi = (Hm1 - (int)(rci >> 8)) * NC + (Wm1 - (int)(rci & 255u));
            uint32_t key[EPL];
#pragma unroll
            for (int s = 0; s < EPL; ++s) {
              const int k = sl + LPE * s;
              key[s] = MDX_NONE32;
              if (k < nw) {
                const uint32_t rc = w.wrc[k];
                const uint32_t rk = __ldg(map.rank + ((int)(rc >> 8) * NC + (int)(rc & 255u) + bi));
                key[s] = (((uint32_t)w.wdl[k] * (uint32_t)p.nranks_max + rk) << p.key_wbits) | (uint32_t)k;
              }
            }
            const int x0 = ii * Ksel;
#pragma unroll 1
            for (int q = 0; q < nsel; ++q) {
              uint32_t lm = key[0];
#pragma unroll
              for (int s = 1; s < EPL; ++s) lm = min(lm, key[s]);
              const uint32_t best = g.rmin(lm);
              if (sl == x0 + q) mysel = best & wmask;
#pragma unroll
              for (int s = 0; s < EPL; ++s)
                if (key[s] == best) key[s] = MDX_NONE32;  // keys are unique: exactly one entry is retired
            }
          }
          if (sel_ii < npass && sel_q < nsel) {
            const uint32_t rci = w.rrc[ib + sel_ii];
            const int ri = (int)(rci >> 8), ci = (int)(rci & 255u);
            const int k = (int)mysel;
            const int j = w.wlist[k];
            const uint32_t rc = w.wrc[k];
            const uint32_t tc_ = w.target[j];
            const int tr = (int)__umulhi(tc_, p.mul_wout), tcol = (int)tc_ - tr * Wo;
            float* o = pk_base + (ib - i0 + sel_ii) * D + sel_q * 5;
            o[0] = __ldg(qr + ((int)(rc >> 8) - ri + Hm1));
            o[1] = __ldg(qc + ((int)(rc & 255u) - ci + Wm1));
            o[2] = __ldg(qr + (tr - ri + Hm1));
            o[3] = __ldg(qc + (tcol - ci + Wm1));
            o[4] = tnorm(p, (int)w.wdl[k]);
          }
        }
      } else if (nsel > 0 && 2 * nsel >= nw && (kSmall || Ksel > 5)) {
        // Most (or all) of the waiting list is selected -- the K = 100 defaults of generate_vector_features
        // (MAPPO/helper.py:70): rank every entry by counting the smaller keys (O(nw) per entry, no rounds) and let
        // the entry's owner write its row at that position.
#pragma unroll 1
        for (int i = i0; i < i0 + nrows; ++i) {
          const uint32_t rci = w.rrc[i];
          const int ri = (int)(rci >> 8), ci = (int)(rci & 255u);
#pragma unroll 1
          for (int k = sl; k < nw; k += LPE) {
            const uint32_t rc = w.wrc[k];
            const uint32_t rk = __ldg(map.rank + ((int)(rc >> 8) - ri + Hm1) * NC + ((int)(rc & 255u) - ci + Wm1));
            w.wkey[k] = (((uint32_t)w.wdl[k] * (uint32_t)p.nranks_max + rk) << p.key_wbits) | (uint32_t)k;
          }
          g.sync();
#pragma unroll 1
          for (int k = sl; k < nw; k += LPE) {
            const uint32_t key = w.wkey[k];
            int pos = 0;
#pragma unroll 4
            for (int m = 0; m < nw; ++m) pos += (w.wkey[m] < key) ? 1 : 0;  // same address for all lanes: broadcast
            if (pos < nsel) {
              const int j = w.wlist[k];
              const uint32_t rc = w.wrc[k];
              const uint32_t tc_ = w.target[j];
              const int tr = (int)__umulhi(tc_, p.mul_wout), tcol = (int)tc_ - tr * Wo;
              float* o = pk_base + (i - i0) * D + pos * 5;
              o[0] = __ldg(qr + ((int)(rc >> 8) - ri + Hm1));
              o[1] = __ldg(qc + ((int)(rc & 255u) - ci + Wm1));
              o[2] = __ldg(qr + (tr - ri + Hm1));
              o[3] = __ldg(qc + (tcol - ci + Wm1));
              o[4] = tnorm(p, (int)w.wdl[k]);
            }
          }
          g.sync();
        }
      } else if (nsel > 0 && Ksel <= 5) {
        // Few slots, long waiting list (C4 / C5): while it computes the keys of its entries, every lane keeps its
        // own five smallest in sorted registers; a round is then one group arg-min over the lanes' heads and a
        // register shift in the winning lane -- no key array in shared memory, no rescans.
        if (use_cand) {
#pragma unroll 1
          for (int i = i0; i < i0 + nrows; i += 2) {
            const bool two = i + 1 < i0 + nrows;
            const uint32_t rci = w.rrc[i], rci1 = w.rrc[two ? i + 1 : i];
            const int bi = (Hm1 - (int)(rci >> 8)) * NC + (Wm1 - (int)(rci & 255u));
            const int bi1 = (Hm1 - (int)(rci1 >> 8)) * NC + (Wm1 - (int)(rci1 & 255u));
            uint16_t* selp = w.sel + (i - i0) * Ksel;
            uint32_t key = MDX_NONE32, key1 = MDX_NONE32;
            if (c_k != MDX_NONE32) {
              key = ((c_dlk + (uint32_t)__ldg(map.rank + c_off + bi)) << p.key_wbits) | c_k;
              key1 = ((c_dlk + (uint32_t)__ldg(map.rank + c_off + bi1)) << p.key_wbits) | c_k;
            }
#pragma unroll 1
            for (int q = 0; q < nsel; ++q) {
              const uint32_t best = g.rmin(key), best1 = g.rmin(key1);
              if (key == best) {  // exactly one lane: keys are unique
                selp[q] = (uint16_t)(best & wmask);
                key = MDX_NONE32;
              }
              if (key1 == best1) {
                if (two) selp[Ksel + q] = (uint16_t)(best1 & wmask);
                key1 = MDX_NONE32;
              }
            }
          }
        } else {
        // two robots per pass over the waiting list: one read of an entry's cell / deadline feeds both robots' keys, and the two
        // arg-min chains of the pop rounds are independent (the second robot of an odd tail repeats the first and is not written)
#pragma unroll 1
        for (int i = i0; i < i0 + nrows; i += 2) {
          const bool two = i + 1 < i0 + nrows;
          const uint32_t rci = w.rrc[i], rci1 = w.rrc[two ? i + 1 : i];
          const int bi = (Hm1 - (int)(rci >> 8)) * NC + (Wm1 - (int)(rci & 255u));
          const int bi1 = (Hm1 - (int)(rci1 >> 8)) * NC + (Wm1 - (int)(rci1 & 255u));
          uint16_t* selp = w.sel + (i - i0) * Ksel;
          uint32_t a0 = MDX_NONE32, a1 = MDX_NONE32, a2 = MDX_NONE32, a3 = MDX_NONE32, a4 = MDX_NONE32;
          uint32_t b0 = MDX_NONE32, b1 = MDX_NONE32, b2 = MDX_NONE32, b3 = MDX_NONE32, b4 = MDX_NONE32;
#pragma unroll 1
          for (int k = sl; k < nw; k += LPE) {
            const uint32_t rc = w.wrc[k];
            const int co = (int)(rc >> 8) * NC + (int)(rc & 255u);
            const uint32_t rk = __ldg(map.rank + (co + bi)), rk1 = __ldg(map.rank + (co + bi1));
            const uint32_t dk = (uint32_t)w.wdl[k] * (uint32_t)p.nranks_max;
            uint32_t x = ((dk + rk) << p.key_wbits) | (uint32_t)k, y = ((dk + rk1) << p.key_wbits) | (uint32_t)k;
            uint32_t lo;
            lo = min(a0, x), x = max(a0, x), a0 = lo;
            lo = min(b0, y), y = max(b0, y), b0 = lo;
            lo = min(a1, x), x = max(a1, x), a1 = lo;
            lo = min(b1, y), y = max(b1, y), b1 = lo;
            lo = min(a2, x), x = max(a2, x), a2 = lo;
            lo = min(b2, y), y = max(b2, y), b2 = lo;
            lo = min(a3, x), x = max(a3, x), a3 = lo;
            lo = min(b3, y), y = max(b3, y), b3 = lo;
            a4 = min(a4, x);
            b4 = min(b4, y);
          }
#pragma unroll 1
          for (int q = 0; q < nsel; ++q) {
            const uint32_t best = g.rmin(a0), best1 = g.rmin(b0);
            if (a0 == best) {  // exactly one lane: keys are unique
              selp[q] = (uint16_t)(best & wmask);
              a0 = a1, a1 = a2, a2 = a3, a3 = a4, a4 = MDX_NONE32;
            }
            if (b0 == best1) {
              if (two) selp[Ksel + q] = (uint16_t)(best1 & wmask);
              b0 = b1, b1 = b2, b2 = b3, b3 = b4, b4 = MDX_NONE32;
            }
          }
        }
        }
        g.sync();
        const int nslot = nrows * Ksel;
#pragma unroll 1
        for (int x = sl; x < nslot; x += LPE) {
          const int ii = (int)div_magic((uint32_t)x, p.mul_ksel), q = x - ii * Ksel;
          if (q < nsel) {
            const uint32_t rci = w.rrc[i0 + ii];
            const int ri = (int)(rci >> 8), ci = (int)(rci & 255u);
            const int k = w.sel[ii * Ksel + q];
            const int j = w.wlist[k];
            const uint32_t rc = w.wrc[k];
            const uint32_t tc_ = w.target[j];
            const int tr = (int)__umulhi(tc_, p.mul_wout), tcol = (int)tc_ - tr * Wo;
            float* o = pk_base + ii * D + q * 5;
            o[0] = __ldg(qr + ((int)(rc >> 8) - ri + Hm1));
            o[1] = __ldg(qc + ((int)(rc & 255u) - ci + Wm1));
            o[2] = __ldg(qr + (tr - ri + Hm1));
            o[3] = __ldg(qc + (tcol - ci + Wm1));
            o[4] = tnorm(p, (int)w.wdl[k]);
          }
        }
      } else if (nsel > 0) {
#pragma unroll 1
        for (int i = i0; i < i0 + nrows; ++i) {
          const uint32_t rci = w.rrc[i];
          const int ri = (int)(rci >> 8), ci = (int)(rci & 255u);
          uint16_t* selp = w.sel + (i - i0) * Ksel;
          uint32_t lm = MDX_NONE32;
#pragma unroll 1
          for (int k = sl; k < nw; k += LPE) {
            const uint32_t rc = w.wrc[k];
            const uint32_t rk = __ldg(map.rank + ((int)(rc >> 8) - ri + Hm1) * NC + ((int)(rc & 255u) - ci + Wm1));
            const uint32_t key = (((uint32_t)w.wdl[k] * (uint32_t)p.nranks_max + rk) << p.key_wbits) | (uint32_t)k;
            w.wkey[k] = key;
            lm = min(lm, key);
          }
#pragma unroll 1
          for (int q = 0; q < nsel; ++q) {
            const uint32_t best = g.rmin(lm);
            if (lm == best) {  // exactly one lane: keys are unique
              const uint32_t k = best & wmask;
              selp[q] = (uint16_t)k;
              w.wkey[k] = MDX_NONE32;
              lm = MDX_NONE32;
#pragma unroll 1
              for (int kk = sl; kk < nw; kk += LPE) lm = min(lm, w.wkey[kk]);
            }
          }
        }
        g.sync();
        const int nslot = nrows * Ksel;
#pragma unroll 1
        for (int x = sl; x < nslot; x += LPE) {
          const int ii = (int)div_magic((uint32_t)x, p.mul_ksel), q = x - ii * Ksel;
          if (q < nsel) {
            const uint32_t rci = w.rrc[i0 + ii];
            const int ri = (int)(rci >> 8), ci = (int)(rci & 255u);
            const int k = w.sel[ii * Ksel + q];
            const int j = w.wlist[k];
            const uint32_t rc = w.wrc[k];
            const uint32_t tc_ = w.target[j];
            const int tr = (int)__umulhi(tc_, p.mul_wout), tcol = (int)tc_ - tr * Wo;
            float* o = pk_base + ii * D + q * 5;
            o[0] = __ldg(qr + ((int)(rc >> 8) - ri + Hm1));
            o[1] = __ldg(qc + ((int)(rc & 255u) - ci + Wm1));
            o[2] = __ldg(qr + (tr - ri + Hm1));
            o[3] = __ldg(qc + (tcol - ci + Wm1));
            o[4] = tnorm(p, (int)w.wdl[k]);
          }
        }
      }
      g.sync();
      flush_seg<LPE>(p.vec + g0, w.stage, off, nrows * D, nrows * D, sl);
      g.sync();
    }
  }

  // ---- gvec [Dg]  (convert_global_state vector part, MAPPO/helper.py:199-255) ------------------
  // three segments: robots (6*MR floats, first 6*min(R,MR) carry data), packages (7*MP floats, first
  // 7*min(n_active,MP) carry data), global time.  Only the data prefixes are staged.
  if (p.gvec != nullptr) {
    const int Dg = p.Dg, MR = p.MR, MP = p.MP;
    const size_t gA = (size_t)row * Dg, gB = gA + 6 * (size_t)MR;
    const int offA = (int)(gA & 3);
    if constexpr (kFix == 1) {  // the baked C3 shape stages all its (<= 56) package rows at once: round-1 code, kept verbatim (56 registers)
    const int nA = 6 * min(R, MR), nrow = min(na, MP), nB = 7 * nrow;
    const int offB = (int)(gB & 3);
    float* stA = w.stage;
    float* stB = w.stage + p.stage_b;
    zero_stage<LPE>(stA, nA + 8, sl);
    zero_stage<LPE>(stB, nB + 8, sl);
    g.sync();
    if (active && sl < MR) {
      float* o = stA + offA + sl * 6;
      o[0] = __ldg(qr + (r + Hm1));
      o[1] = __ldg(qc + (c + Wm1));
      o[2] = carry != 0 ? 1.0f : 0.0f;
      if (carrying_it) {
        o[3] = __ldg(qr + (tr_own + Hm1));
        o[4] = __ldg(qc + (tc_own + Wm1));
        o[5] = tnorm(p, own_rem);
      }
    }
#pragma unroll 1
    for (int k = sl; k < nrow; k += LPE) {
      const int j = w.alist[k];
      const bool wait = w.status[j] == ST_WAITING;
      const uint32_t sc = w.start[j], tc_ = w.target[j];
      const int sr = (int)__umulhi(sc, p.mul_wout), scol = (int)sc - sr * Wo;
      const int tr = (int)__umulhi(tc_, p.mul_wout), tcol = (int)tc_ - tr * Wo;
      float* o = stB + offB + 7 * k;
      if (wait) {
        o[0] = __ldg(qr + (sr + Hm1));
        o[1] = __ldg(qc + (scol + Wm1));
      }
      o[2] = __ldg(qr + (tr + Hm1));
      o[3] = __ldg(qc + (tcol + Wm1));
      o[4] = tnorm(p, max(0, (int)w.dl[j] - t));
      o[5] = wait ? 0.0f : 1.0f;
      float cn = -1.0f;
      if (!wait) {
        for (int ridx = 0; ridx < R; ++ridx)
          if ((int)w.rcarry[ridx] == id_of<kTrk>(w, j)) {
            cn = __ldg(p.cq + ridx);  // ridx / (MR - 1), or 0 when MR <= 1
            break;
          }
      }
      o[6] = cn;
    }
    g.sync();
    flush_seg<LPE>(p.gvec + gA, stA, offA, nA, 6 * MR, sl);
    flush_seg<LPE>(p.gvec + gB, stB, offB, nB, 7 * MP, sl);
    } else {
    const int nA = 6 * min(R, MR), nrow = min(na, MP);
    float* stA = w.stage;
    float* stB = w.stage + p.stage_b;
    // package rows go out p.gvec_chunk at a time (all at once unless a smaller stage lets one more CTA fit the SM); the first chunk
    // is staged together with the robot rows so that the usual single-chunk case leaves as one burst; the last chunk also writes
    // the segment's zero tail
    const int CH = p.gvec_chunk;
    auto fill_rows = [&](int r0, int nr, int offC) {
#pragma unroll 1
      for (int k = r0 + sl; k < r0 + nr; k += LPE) {
        const int j = w.alist[k];
        const bool wait = w.status[j] == ST_WAITING;
        const uint32_t sc = w.start[j], tc_ = w.target[j];
        const int sr = (int)__umulhi(sc, p.mul_wout), scol = (int)sc - sr * Wo;
        const int tr = (int)__umulhi(tc_, p.mul_wout), tcol = (int)tc_ - tr * Wo;
        float* o = stB + offC + 7 * (k - r0);
        if (wait) {
          o[0] = __ldg(qr + (sr + Hm1));
          o[1] = __ldg(qc + (scol + Wm1));
        }
        o[2] = __ldg(qr + (tr + Hm1));
        o[3] = __ldg(qc + (tcol + Wm1));
        o[4] = tnorm(p, max(0, (int)w.dl[j] - t));
        o[5] = wait ? 0.0f : 1.0f;
        float cn = -1.0f;
        if (!wait) {
          for (int ridx = 0; ridx < R; ++ridx)
            if ((int)w.rcarry[ridx] == id_of<kTrk>(w, j)) {
              cn = __ldg(p.cq + ridx);  // ridx / (MR - 1), or 0 when MR <= 1
              break;
            }
        }
        o[6] = cn;
      }
    };
    const int offB = (int)(gB & 3), nr0 = min(CH, nrow);
    zero_stage<LPE>(stA, nA + 8, sl);
    zero_stage<LPE>(stB, 7 * nr0 + 8, sl);
    g.sync();
    if (active && sl < MR) {
      float* o = stA + offA + sl * 6;
      o[0] = __ldg(qr + (r + Hm1));
      o[1] = __ldg(qc + (c + Wm1));
      o[2] = carry != 0 ? 1.0f : 0.0f;
      if (carrying_it) {
        o[3] = __ldg(qr + (tr_own + Hm1));
        o[4] = __ldg(qc + (tc_own + Wm1));
        o[5] = tnorm(p, own_rem);
      }
    }
    fill_rows(0, nr0, offB);
    g.sync();
    flush_seg<LPE>(p.gvec + gA, stA, offA, nA, 6 * MR, sl);
    flush_seg<LPE>(p.gvec + gB, stB, offB, 7 * nr0, CH >= nrow ? 7 * MP : 7 * nr0, sl);
#pragma unroll 1
    for (int r0 = CH; r0 < nrow; r0 += CH) {
      const int nr = min(CH, nrow - r0);
      const size_t gC = gB + 7 * (size_t)r0;
      const int offC = (int)(gC & 3);
      g.sync();
      zero_stage<LPE>(stB, 7 * nr + 8, sl);
      g.sync();
      fill_rows(r0, nr, offC);
      g.sync();
      flush_seg<LPE>(p.gvec + gC, stB, offC, 7 * nr, r0 + CH >= nrow ? 7 * (MP - r0) : 7 * nr, sl);
    }
    }
    if (sl == 0) __stcs(p.gvec + gA + Dg - 1, tnorm(p, t));
    g.sync();
  }
}

// ------------------------------------------------------------------------------------------------
// IDQ / legacy-qmix featurisers (IDQ/networks.py:112-217, qmix/networks.py:350-468) of the env held in shared memory,
// fused behind the step (kMode bit 2): obs6 [R, 6, H, W] per robot and the 7-channel global tensor.  Plane-sliced like
// featurize_env: a lane owns float4 number (sl + 32*s) of every plane; six bit planes live in w.planes, the urgency
// float plane in w.stage.
// urgency of a waiting package whose start_time has passed (IDQ/networks.py:172-179): float64 quotient,
// clamped, then stored into a float32 array
__device__ __forceinline__ float alt_urgency(int t, int st, int dl) {
  if (dl > st) {
    double u = __ddiv_rn((double)(t - st), (double)(dl - st));
    u = u > 1.0 ? 1.0 : u;
    return __double2float_rn(u);
  }
  return dl == st ? 1.0f : 0.0f;
}

template <int kS>
__device__ __forceinline__ void featurize_alt_env(const KParams& p, const Env& w, const Grp<32>& g, const float4* lut, const MapDev& map,
                                                  int row, int t, int cell, int carry) {
  constexpr int NS = kS > 0 ? kS : 1;
  const int sl = g.sl;
  const int R = p.R, P = p.P, HW = p.HW, W32 = p.W32;
  const bool active = sl < R;
  float* urg = w.stage;
  uint32_t* planes = w.planes;
  const int n4p = HW >> 2;
#pragma unroll 1
  for (int i = sl; i < n4p; i += 32) reinterpret_cast<float4*>(urg)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
  for (int i = sl; i < W32; i += 32) {
    planes[i] = __ldg(map.grid_bits + i);
#pragma unroll
    for (int k = 1; k < 6; ++k) planes[k * W32 + i] = 0u;
  }
  g.sync();
  // planes: 0 grid, 1 robots, 2 robots that carry, 3 waiting starts, 4 waiting targets, 5 in-transit targets
  int own_tgt = 0xFFFF;
  if (active) {
    atomicOr(&planes[W32 + (cell >> 5)], 1u << (cell & 31));
    if (carry != 0) {
      atomicOr(&planes[2 * W32 + (cell >> 5)], 1u << (cell & 31));
      const uint8_t st = w.status[carry - 1];
      if (st == ST_WAITING || st == ST_IN_TRANSIT) own_tgt = w.target[carry - 1];  // "id in persistent_packages"
    }
  }
#pragma unroll 1
  for (int j = sl; j < P; j += 32) {
    const uint8_t st = w.status[j];
    if (st != ST_WAITING && st != ST_IN_TRANSIT) continue;
    const int t0 = w.t0[j];
    if (t < t0) continue;
    const int tc = w.target[j];
    if (st == ST_WAITING) {
      const int sc = w.start[j];
      atomicOr(&planes[3 * W32 + (sc >> 5)], 1u << (sc & 31));
      atomicOr(&planes[4 * W32 + (tc >> 5)], 1u << (tc & 31));
      atomicMax(reinterpret_cast<int*>(urg + sc), __float_as_int(alt_urgency(t, t0, (int)w.dl[j])));
    } else {
      atomicOr(&planes[5 * W32 + (tc >> 5)], 1u << (tc & 31));
    }
  }
  g.sync();
  uint32_t nib[NS][6];
  float4 u4[NS];
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    const int q = sl + 32 * s;
    const bool ok = q < n4p;
    const int wi = q >> 3;
    const uint32_t sh = (uint32_t)(q & 7) << 2;
#pragma unroll
    for (int k = 0; k < 6; ++k) nib[s][k] = ok ? (planes[k * W32 + wi] >> sh) & 15u : 0u;
    u4[s] = ok ? reinterpret_cast<const float4*>(urg)[q] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p.alt_g7 != nullptr) {
    float4* out4 = reinterpret_cast<float4*>(p.alt_g7 + (size_t)row * 7 * HW);
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      const int q = sl + 32 * s;
      if (q < n4p) {
#pragma unroll
        for (int k = 0; k < 6; ++k) __stcs(out4 + k * n4p + q, lut[nib[s][k]]);
        __stcs(out4 + 6 * n4p + q, u4[s]);
      }
    }
  }
  if (p.alt_obs6 != nullptr) {
    float4* base4 = reinterpret_cast<float4*>(p.alt_obs6 + (size_t)row * R * 6 * HW);
#pragma unroll 1
    for (int i = 0; i < R; ++i) {
      const uint32_t ci = (uint32_t)g.shfl(cell, i), ti = (uint32_t)g.shfl(own_tgt, i);
      const bool free_i = g.shfl(carry, i) == 0;
      float4* out4 = base4 + (size_t)i * 6 * n4p;
#pragma unroll
      for (int s = 0; s < NS; ++s) {
        const int q = sl + 32 * s;
        if (q < n4p) {
          const uint32_t oh_r = ((ci >> 2) == (uint32_t)q) ? (1u << (ci & 3u)) : 0u;
          const uint32_t oh_t = ((ti >> 2) == (uint32_t)q) ? (1u << (ti & 3u)) : 0u;  // 0xFFFF (none) matches no q
          __stcs(out4 + q, lut[nib[s][0]]);
          __stcs(out4 + n4p + q, free_i ? u4[s] : z4);
          __stcs(out4 + 2 * n4p + q, lut[free_i ? nib[s][3] : 0u]);
          __stcs(out4 + 3 * n4p + q, lut[nib[s][1] & ~oh_r]);
          __stcs(out4 + 4 * n4p + q, lut[oh_r]);
          __stcs(out4 + 5 * n4p + q, lut[oh_t]);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// kMode bit 0: step (+ shaped reward, + optional auto-reset);  bit 1: featurise (MAPPO);  bit 2: featurise (IDQ / legacy, whole warps only).
// LPE lanes per environment: a 128-thread CTA holds 128 / LPE environments.
// kFix: 2 = like 0, compiled with a looser register cap (5 CTAs/SM instead of 7) for the four-slice planes with long package lists
// (C4, C5: 20x20 maps, > 64 packages), whose time goes into the plane stores: with registers to spare the compiler keeps six
// LDS.128 -> STG.128 pairs in flight per warp instead of recycling one register quad (ncu: the capped build stalls on the
// write-after-read of that quad).
// kFix: shape specialisation.  0 = every shape parameter is read from KParams at run time; 1 = the
// headline MAPPO shape (C3: map1 10x10, 5 robots, <= 56 packages, K_o = 4, K_p = 5, MR = MP = 100) is baked
// in as compile-time constants, so loops over robots / planes / rows unroll and the index arithmetic folds.
// MAPPO/trainer.py:95-130 `_update_persistent_packages_for_env(env_idx, state)` on the tracker table in shared memory; `state` is
// the env's current one: time t_now, robot i carries `carry` (lane i), packages of the env truth `w`.  Returns the new dict size.
//   1. every package the state lists as newly spawned (start_time == t_now, env.py:134-137) joins the END of the dict unless its
//      id is already a key -- a stale entry of an earlier episode shadows it;
//   2. ids some robot carries become 'in_transit'; 'in_transit' ids nobody carries any more are deleted (the others keep their order).
template <int LPE>
__device__ __forceinline__ int tracker_update(const KParams& p, const Env& w, const Env& f, const Grp<LPE>& g, int t_now, int carry, int n) {
  const unsigned lt = (1u << g.sl) - 1u;
#pragma unroll 1
  for (int base = 0; base < p.P; base += LPE) {
    const int j = base + g.sl;
    const bool add = j < p.P && (int)w.t0[j] == t_now && f.slot[j] == kNoSlot;
    const unsigned m = g.ballot(add);
    if (add) {
      const int k = n + __popc(m & lt);
      f.pid[k] = (uint16_t)(j + 1);
      f.start[k] = w.start[j];
      f.target[k] = w.target[j];
      f.t0[k] = w.t0[j];
      f.dl[k] = w.dl[j];
      f.status[k] = ST_WAITING;
      f.slot[j] = (uint16_t)k;
    }
    n += __popc(m);
  }
  g.sync();
  int out = 0;
#pragma unroll 1
  for (int base = 0; base < n; base += LPE) {
    const int k = base + g.sl;
    const bool valid = k < n;
    uint32_t id = 0u, a = 0u, b = 0u, c = 0u, d = 0u;
    uint8_t st = ST_NONE;
    if (valid) {
      id = f.pid[k], a = f.start[k], b = f.target[k], c = f.t0[k], d = f.dl[k];
      st = f.status[k];
    }
    bool carried = false;
#pragma unroll 1
    for (int i = 0; i < p.R; ++i) {
      const uint32_t cid = (uint32_t)g.shfl(carry, i);  // (every lane of the group executes every shuffle: no short-circuit)
      carried = carried | (cid == id);                  // id 0 (invalid lane) matches "carries nothing": masked below
    }
    carried = carried && valid;
    if (carried) st = ST_IN_TRANSIT;
    const bool keep = valid && (carried || st != ST_IN_TRANSIT);
    const unsigned m = g.ballot(keep);
    g.sync();  // every lane holds its entry in registers: the slots below `base` may be overwritten
    if (keep) {
      const int q = out + __popc(m & lt);
      f.pid[q] = (uint16_t)id, f.start[q] = (uint16_t)a, f.target[q] = (uint16_t)b, f.t0[q] = (uint16_t)c, f.dl[q] = (uint16_t)d;
      f.status[q] = st;
      f.slot[id - 1] = (uint16_t)q;
    } else if (valid) {
      f.slot[id - 1] = (uint16_t)kNoSlot;
    }
    out += __popc(m);
    g.sync();
  }
  return out;
}

template <int kFix>
__device__ __forceinline__ void specialise(KParams& p) {
#ifdef MDX_JIT_SHAPE
  // kFix = 3: this translation unit is being compiled AT RUN TIME (mdx_jit.cpp, NVRTC) for the shape of one mdx_env instance; every
  // shape parameter arrives as a macro, so loops unroll and index arithmetic folds exactly like in the baked C3 variant (kFix = 1) --
  // for ANY (H, W, R, P, K_o, K_p, MR, MP).  The algorithms are the run-time-shape ones (kFix = 0 code paths).
  if (kFix == 3) {
    p.R = MDX_J_R, p.P = MDX_J_P, p.T = MDX_J_T, p.Pp8 = MDX_J_PP8, p.Pp16 = MDX_J_PP16;
    p.H_out = MDX_J_H, p.W_out = MDX_J_W, p.HW = MDX_J_H * MDX_J_W, p.W32 = (MDX_J_H * MDX_J_W + 31) / 32;
    p.K_o = MDX_J_KO, p.K_p = MDX_J_KP, p.sel_k = MDX_J_SELK, p.MR = MDX_J_MR, p.MP = MDX_J_MP;
    p.D = 6 + 5 * MDX_J_KO + 5 * MDX_J_KP + 1, p.Dg = 6 * MDX_J_MR + 7 * MDX_J_MP + 1;
    p.vec_rows_per_flush = MDX_J_RPF, p.gvec_chunk = MDX_J_GCH, p.stage_b = MDX_J_STAGEB, p.key_wbits = MDX_J_WBITS, p.nranks_max = MDX_J_NRANKS;
    p.mul_wout = MDX_J_MUL_WOUT, p.mul_r = MDX_J_MUL_R, p.mul_ksel = MDX_J_MUL_KSEL, p.mul_plane = MDX_J_MUL_PLANE;
    p.pair_g = 32 / MDX_J_R, p.sel_rp = MDX_J_SELRP, p.tq_len = MDX_J_TQLEN;
    p.off_sched = MDX_J_OFF_SCHED, p.off_status = MDX_J_OFF_STATUS, p.off_wlist = MDX_J_OFF_WLIST, p.off_wkey = MDX_J_OFF_WKEY;
    p.off_rinfo = MDX_J_OFF_RINFO, p.off_alist = MDX_J_OFF_ALIST, p.off_sel = MDX_J_OFF_SEL, p.off_masks = MDX_J_OFF_MASKS;
    p.off_stage = MDX_J_OFF_STAGE, p.smem_per_warp = MDX_J_SPW;
  }
#endif
  if (kFix == 1) {
    p.R = 5;
    p.Pp8 = 56;
    p.Pp16 = 64;
    p.H_out = 10;
    p.W_out = 10;
    p.HW = 100;
    p.W32 = 4;
    p.K_o = 4;
    p.K_p = 5;
    p.sel_k = 5;
    p.MR = 100;
    p.MP = 100;
    p.D = 52;
    p.Dg = 1301;
    p.vec_rows_per_flush = 5;
    p.mul_wout = 429496730u;   // ceil(2^32 / 10)
    p.mul_r = 858993460u;      // ceil(2^32 / 5)
    p.mul_ksel = 858993460u;
    p.mul_plane = 42949673u;   // ceil(2^32 / 100)
    p.pair_g = 6;
    p.stage_b = 44;            // round_up(6 * 5 + 12, 4)
    // shared-memory carve-up of this shape (mdx_create's formula; the launcher checks it): constant offsets fold
    // into the LDS/STS immediates instead of living in registers
    p.off_sched = 0;
    p.off_status = 448;
    p.off_wlist = 512;
    p.off_wkey = 848;
    p.off_rinfo = 1072;
    p.off_alist = 1392;
    p.off_sel = 1504;
    p.off_masks = 1568;
    p.off_stage = 1728;
    p.key_wbits = 6;           // ceil(log2(P)) for 33..64 packages
  }
}

// kTrk: MDX_CFG_STALE_TRACKER instances -- shaped reward and featurisers read the trainer's own (never cleared) package dict, which
// the kernel also maintains; flags bit 2 (internal): apply the tracker update for the CURRENT state first (after mdx_reset);
// flags bit 3 (internal): empty the dict first (mdx_tracker_clear).
template <int kMode, int LPE, int kS, bool kSmall, int kFix, bool kTrk = false>
__global__ void __launch_bounds__((kFix == 1 && (kMode & 2)) ? MDX_FIX_THREADS : 128,
                                   (kMode & 2) ? (kFix == 1 ? MDX_FIX_BLOCKS : kFix == 2 ? MDX_FAT_BLOCKS : kFix == 3 ? MDX_JIT_BLOCKS : MDX_FEAT_BLOCKS)
                                               : ((kMode & 4) && kS > 2) ? 6 : 8) mdx_env_kernel(const __grid_constant__ KParams pg) {
  // Programmatic dependent launch (sm_90+): when the launch carries the programmatic-stream-serialisation attribute (launch-bound
  // step-only batches, mdx_launch.inl), this grid may be scheduled while its predecessor on the stream is still draining; it lets ITS
  // successor do the same right away and then waits here until the predecessor's writes (the env state) are complete and visible.
  // Without the attribute both instructions are no-ops.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  KParams p = pg;
  specialise<kFix>(p);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const float4* lut = reinterpret_cast<const float4*>(smem_raw);
  if (kMode & 6) {
    if (threadIdx.x < 16) {
      const unsigned n = threadIdx.x;
      reinterpret_cast<float4*>(smem_raw)[n] =
          make_float4((n & 1u) ? 1.f : 0.f, (n & 2u) ? 1.f : 0.f, (n & 4u) ? 1.f : 0.f, (n & 8u) ? 1.f : 0.f);
    }
    __syncthreads();
  }
  const int gib = threadIdx.x / LPE;  // group index inside the CTA
  const int widx = blockIdx.x * (blockDim.x / LPE) + gib;
  if (widx >= p.n) return;  // whole groups leave; every collective below names only the group's lanes
  const Grp<LPE> g(threadIdx.x & 31);
  const int sl = g.sl;
  const Env w = carve(p, smem_raw + kLutBytes + (size_t)gib * p.smem_per_warp);
  const int env = p.env_ids ? __ldg(p.env_ids + widx) : p.env_base + widx;
  if ((unsigned)env >= (unsigned)p.n_envs) return;  // a bad caller-supplied id never becomes an out-of-bounds access (row left unwritten)
  const int R = p.R, P = p.P, Wo = p.W_out;
  const bool active = sl < R;
  // single-map instances read the map descriptor from the kernel parameters (no pointer chase)
  MapDev map = p.map0;
  if (p.env_map) map = p.maps[__ldg(p.env_map + env)];

  // Issue every global load of the env's state up front so that their latencies overlap; no address depends
  // on a loaded value (the current episode's schedule has its own block, the ring is touched on resets only).
  int t = p.t[env];
  int ep = p.ep[env];
  const uint4* st_src = reinterpret_cast<const uint4*>(p.status + (size_t)env * p.Pp16);
  const int n_st16 = p.Pp16 >> 4;
  uint4 st0 = make_uint4(0u, 0u, 0u, 0u);
  if (sl < n_st16) st0 = st_src[sl];
  uint32_t rob = active ? p.robots[(size_t)env * R + sl] : 0u;
  int action = 3;
  double total0 = 0.0;
  if (kMode & 1) {
    if (active) {
      if (p.flags & 2u) {  // MDX_F_ACTIONS_I64: the policy's int64 samples, decoded like MAPPO/trainer.py:195-204
        const long long a64 = __ldg(reinterpret_cast<const long long*>(p.actions) + (size_t)widx * R + sl);
        action = a64 < 0 ? 18 : (a64 < 15 ? (int)a64 : (int)(a64 % 5));  // package op index >= 3 is clamped to '0'
      } else {
        action = (int)__ldg(p.actions + (size_t)widx * R + sl);
      }
    }
    total0 = p.total_reward[env];
  }
  // first block of the schedule into registers with the other state loads, then the uses
  const uint4* ep_src = reinterpret_cast<const uint4*>(p.cur + (size_t)env * 4 * p.Pp8);
  const int n_ep16 = p.Pp8 >> 1;
  uint4 ep0 = make_uint4(0u, 0u, 0u, 0u);
  if (sl < n_ep16) ep0 = __ldg(ep_src + sl);
  {
    uint4* dst = reinterpret_cast<uint4*>(w.start);
    if (sl < n_ep16) dst[sl] = ep0;
#pragma unroll 1
    for (int i = sl + LPE; i < n_ep16; i += LPE) dst[i] = __ldg(ep_src + i);
  }
  {
    uint4* dst = reinterpret_cast<uint4*>(w.status);
    if (sl < n_st16) dst[sl] = st0;
#pragma unroll 1
    for (int i = sl + LPE; i < n_st16; i += LPE) dst[i] = st_src[i];
  }
  int cell = active ? (int)(rob & 0xffffu) : (0x10000 + sl);
  int carry = active ? (int)(rob >> 16) : 0;
  Env wf = w;  // what the shaped reward and the featurisers read
  int ntrk = 0;
  if constexpr (kTrk) {
    unsigned char* base = smem_raw + kLutBytes + (size_t)gib * p.smem_per_warp;
    wf = carve_tracker(p, base, w);
    ntrk = (p.flags & 8u) ? 0 : p.trk_n[env];
    const uint4* src = reinterpret_cast<const uint4*>(p.trk + (size_t)env * 5 * p.Pp8);
    uint4* dst = reinterpret_cast<uint4*>(wf.pid);
#pragma unroll 1
    for (int i = sl; i < (5 * p.Pp8) >> 3; i += LPE) dst[i] = src[i];
    const uint4* ssrc = reinterpret_cast<const uint4*>(p.trk_status + (size_t)env * p.Pp16);
    uint4* sdst = reinterpret_cast<uint4*>(wf.status);
#pragma unroll 1
    for (int i = sl; i < (p.Pp16 >> 4); i += LPE) sdst[i] = ssrc[i];
#pragma unroll 1
    for (int j = sl; j < p.Pp8; j += LPE) wf.slot[j] = (uint16_t)kNoSlot;
    g.sync();
#pragma unroll 1
    for (int k = sl; k < ntrk; k += LPE) wf.slot[wf.pid[k] - 1] = (uint16_t)k;
  }
  g.sync();
  [[maybe_unused]] bool trk_dirty = false;
  if constexpr (kTrk) {
    if (p.flags & 12u) {
      if (p.flags & 4u) ntrk = tracker_update<LPE>(p, w, wf, g, t, carry, ntrk);
      trk_dirty = true;
    }
  }

  if (kMode & 1) {
    // ---------------- decode ----------------
    int mv, op;
    decode_action(action, mv, op);
    // ---------------- proposals (env.py:193-201) ----------------
    const int r = active ? (int)__umulhi((uint32_t)cell, p.mul_wout) : 0;
    const int c = active ? cell - r * Wo : 0;
    const int nr = r + (mv == MV_D) - (mv == MV_U);
    const int nc = c + (mv == MV_R) - (mv == MV_L);
    int tgt = cell;
    if (active && nr >= 0 && nr < map.H && nc >= 0 && nc < map.W) {
      const int ncell = nr * Wo + nc;
      if (((__ldg(map.grid_bits + (ncell >> 5)) >> (ncell & 31)) & 1u) == 0u) tgt = ncell;
    }
    // ---------------- resolution (env.py:203-250) as a monotone fixed point ----------------
    // robot i moves  <=>  it is the lowest-index contender for its target cell AND that cell was
    // empty at step start or its holder moves away.  Cycles/swaps never enter the fixed point.
    int holder = -1;
#pragma unroll 1
    for (int j = 0; j < R; ++j) {
      const int cj = g.shfl(cell, j);
      if (cj == tgt && j != sl) holder = j;
    }
    const bool wants = active && tgt != cell;
    const unsigned grp = g.match(tgt);
    const unsigned wants_mask = g.ballot(wants);
    const bool winner = wants && (((grp & wants_mask) & ((1u << sl) - 1u)) == 0u);
    unsigned mm = 0u;
    while (true) {
      const bool m = winner && (holder < 0 || ((mm >> holder) & 1u));
      const unsigned nm = g.ballot(m);
      if (nm == mm) break;
      mm = nm;
    }
    const bool moved = (mm >> sl) & 1u;
    const int cell1 = moved ? tgt : cell;
    const int r1 = moved ? nr : r, c1 = moved ? nc : c;
    const int n_moved = __popc(mm);  // every mover used one of L/R/U/D (env.py:255)

    // ---------------- package ops in robot order (env.py:260-292) ----------------
    // robots stand on distinct cells, so pickups never contend across robots: each takes the
    // lowest-id waiting package on its own cell
    const int nw_prev = build_waiting<LPE>(p, w, g, t, P);
    const bool bfs_mode = kFix != 1 && p.shaping_bfs != 0;
    const Scan sc = bfs_mode ? scan_waiting<LPE, kSmall, true>(w, g, R, nw_prev, r, c, r1, c1, map.bfs, p.HW, Wo)
                             : scan_waiting<LPE, kSmall, false>(w, g, R, nw_prev, r, c, r1, c1, nullptr, 0, 0);
    const bool want_pick = active && op == 1 && carry == 0;
    const int picked = (want_pick && sc.at_new != MDX_NONE32) ? (int)w.wlist[sc.at_new] : -1;
    const int tgt0 = carry ? (int)w.target[carry - 1] : -1;
    const int dl0 = carry ? (int)w.dl[carry - 1] : 0;
    const bool delivered = active && op == 2 && carry != 0 && cell1 == tgt0;
    const bool on_time = delivered && t <= dl0;
    const int carry1 = delivered ? 0 : (picked >= 0 ? picked + 1 : carry);
    // env reward in float64, reference order: move costs first, then deliveries by robot index
    const unsigned dmask = g.ballot(delivered), omask = g.ballot(on_time);
    double rew = 0.0;
#pragma unroll 1
    for (int k = 0; k < n_moved; ++k) rew = __dadd_rn(rew, p.move_cost);
    for (unsigned mleft = dmask; mleft; mleft &= mleft - 1u) {
      const int b = __ffs(mleft) - 1;
      rew = __dadd_rn(rew, ((omask >> b) & 1u) ? p.delivery_reward : p.delay_reward);
    }
    // shaped reward reads the PREVIOUS statuses, so evaluate it before they change
    float shaped = 0.0f;
    if constexpr (kTrk) {
      // persistent_packages_at_prev_state is the trainer's dict, not the env truth: second waiting list + scan on the tracker view
      // (`picked` above already consumed the truth list)
      if (p.out_shaped != nullptr) {
        g.sync();
        const int nw_t = build_waiting<LPE>(p, wf, g, t, ntrk);
        const Scan sct = scan_waiting<LPE, kSmall, false>(wf, g, R, nw_t, r, c, r1, c1, nullptr, 0, 0);
        shaped = shaped_team_reward<LPE, false, true>(p, wf, g, t + 1, r, c, carry, r1, c1, carry1, mv, op, nw_t, sct, rew, nullptr);
      }
    } else {
    if (p.out_shaped != nullptr)
      shaped = bfs_mode ? shaped_team_reward<LPE, true>(p, w, g, t + 1, r, c, carry, r1, c1, carry1, mv, op, nw_prev, sc, rew, map.bfs)
                        : shaped_team_reward<LPE, false>(p, w, g, t + 1, r, c, carry, r1, c1, carry1, mv, op, nw_prev, sc, rew, nullptr);
    }
    if (p.out_agent != nullptr && active)
      p.out_agent[(size_t)widx * R + sl] = idq_agent_reward(w, nw_prev, moved, op, r, c, carry, carry1, cell1 == tgt0, t + 1, dl0, p.idq_ops_int);
    g.sync();
    if (delivered) w.status[carry - 1] = ST_DELIVERED;
    if (picked >= 0) w.status[picked] = ST_IN_TRANSIT;
    g.sync();
    // ---------------- time, spawn, termination (env.py:295-306, 127-137) ----------------
    const int t1 = t + 1;
    bool all_delivered = true;
#pragma unroll 1
    for (int base = 0; base < P; base += LPE) {
      const int j = base + sl;
      bool ok = true;
      if (j < P) {
        if ((int)w.t0[j] == t1) w.status[j] = ST_WAITING;
        ok = w.status[j] == ST_DELIVERED;
      }
      all_delivered = g.all(ok) && all_delivered;
    }
    const bool done = (t1 == p.T) || all_delivered;  // not latched
    double total = total0 + rew;
    if (sl == 0) {
      if (p.out_reward) p.out_reward[widx] = rew;
      if (p.out_shaped) p.out_shaped[widx] = shaped;
      if (p.out_done) p.out_done[widx] = done ? 1 : 0;
    }
    // cumulative per-env counters: fire-and-forget reductions, lane = counter index (no load, no wait)
    {
      const int n_pick = __popc(g.ballot(picked >= 0));
      uint32_t delta = 0u;
      if (sl == C_EPISODES) delta = done ? 1u : 0u;
      else if (sl == C_MOVES) delta = (uint32_t)n_moved;
      else if (sl == C_ONTIME) delta = (uint32_t)__popc(omask);
      else if (sl == C_LATE) delta = (uint32_t)__popc(dmask & ~omask);
      else if (sl == C_PICKUPS) delta = (uint32_t)n_pick;
      else if (sl == C_STEPS) delta = 1u;
      else if (sl == C_EPLEN) delta = done ? (uint32_t)t1 : 0u;
      if (sl < 7 && delta != 0u) atomicAdd(p.counters + (size_t)env * 8 + sl, delta);
    }
    cell = active ? cell1 : cell;
    carry = carry1;
    t = t1;
    g.sync();
    // ---------------- on-device auto-reset (MAPPO/trainer.py:229-234) ----------------
    // the env's next reset block is drawn right here, by the group that owns it, from the env's MT19937 state in HBM
    if (done && (p.flags & 1u)) {
      ep += 1;
      const int c0 = sample_episode<LPE>(pg, map.free_cells, map.n_free, map.H, w.start, env);
      for (int j = sl; j < p.Pp16; j += LPE)
        w.status[j] = (j < P && w.t0[j] == 0) ? (uint8_t)ST_WAITING : (uint8_t)ST_NONE;
      if (active) {
        cell = c0;
        carry = 0;
      }
      t = 0;
      total = 0.0;
      g.sync();
    }
    // ---------------- the trainer's tracker (MAPPO/trainer.py:229-280) ----------------
    // a done env is reset first and the dict sees only the RESET state (it is never cleared); a done env that is
    // not auto-reset waits for the caller's mdx_reset, which applies the update; otherwise the dict sees the next state
    if constexpr (kTrk) {
      if (!done || (p.flags & 1u)) {
        ntrk = tracker_update<LPE>(p, w, wf, g, t, carry, ntrk);
        trk_dirty = true;
      }
    }
    // ---------------- write the state back ----------------
    if (active) p.robots[(size_t)env * R + sl] = (uint32_t)cell | ((uint32_t)carry << 16);
    {
      uint4* dst = reinterpret_cast<uint4*>(p.status + (size_t)env * p.Pp16);
      const uint4* src = reinterpret_cast<const uint4*>(w.status);
#pragma unroll 1
      for (int i = sl; i < (p.Pp16 >> 4); i += LPE) dst[i] = src[i];
    }
    if (sl == 0) {
      p.t[env] = t;
      p.ep[env] = ep;
      p.total_reward[env] = total;
    }
    if constexpr (kMode == 1) {
      // mdx_step_host: the packed state row of the env (what the reference's get_state / attribute reads are served from) leaves with
      // the step instead of a second launch; the schedule and the statuses are still in shared memory
      if (p.rows != nullptr) {
        uint32_t* rowp = p.rows + (size_t)widx * p.row_words;
        if (sl == 0) {
          rowp[0] = (uint32_t)t;
          rowp[1] = (uint32_t)ep;
          rowp[2] = p.env_map ? (uint32_t)__ldg(p.env_map + env) : 0u;
          rowp[3] = 0u;
          const unsigned long long tr_ = (unsigned long long)__double_as_longlong(total);
          rowp[4] = (uint32_t)tr_;
          rowp[5] = (uint32_t)(tr_ >> 32);
        }
        if (active) rowp[p.row_off_robots + sl] = (uint32_t)cell | ((uint32_t)carry << 16);
        const uint32_t* s32 = reinterpret_cast<const uint32_t*>(w.start);
        for (int k = sl; k < 2 * p.Pp8; k += LPE) rowp[p.row_off_sched + k] = s32[k];
        const uint32_t* st32 = reinterpret_cast<const uint32_t*>(w.status);
        for (int k = sl; k < (p.Pp16 >> 2); k += LPE) rowp[p.row_off_status + k] = st32[k];
        for (int k = p.row_off_robots + R + sl; k < p.row_off_sched; k += LPE) rowp[k] = 0u;
        for (int k = p.row_off_status + (p.Pp16 >> 2) + sl; k < p.row_words; k += LPE) rowp[k] = 0u;
      }
    }
  }

  if constexpr (kTrk) {
    if (trk_dirty) {
      g.sync();
      uint4* dst = reinterpret_cast<uint4*>(p.trk + (size_t)env * 5 * p.Pp8);
      const uint4* src = reinterpret_cast<const uint4*>(wf.pid);
#pragma unroll 1
      for (int i = sl; i < (5 * p.Pp8) >> 3; i += LPE) dst[i] = src[i];
      uint4* sdst = reinterpret_cast<uint4*>(p.trk_status + (size_t)env * p.Pp16);
      const uint4* ssrc = reinterpret_cast<const uint4*>(wf.status);
#pragma unroll 1
      for (int i = sl; i < (p.Pp16 >> 4); i += LPE) sdst[i] = ssrc[i];
      if (sl == 0) p.trk_n[env] = ntrk;
    }
  }
  if (kMode & 2) featurize_env<LPE, kS, kSmall, kFix, kTrk>(p, kTrk ? wf : w, g, lut, map, widx, t, cell, carry, kTrk ? ntrk : P);
  if constexpr ((kMode & 4) != 0 && LPE == 32 && kS > 0) featurize_alt_env<kS>(p, w, g, lut, map, widx, t, cell, carry);
}

}  // namespace mdx
