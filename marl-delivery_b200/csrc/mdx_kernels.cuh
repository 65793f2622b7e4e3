// Hand-written sm_100a kernels of the marl-delivery hot path.
//
// Mapping: ONE WARP PER ENVIRONMENT, robots are lanes (R <= 32).  The reference's sequential,
// robot-index-ordered move/collision resolution (env.py:203-250) is evaluated with
// __match_any_sync / __ballot_sync as a monotone fixed point; package scans, the shaped reward
// (MAPPO/helper.py:257-369) and the three featurisers (MAPPO/helper.py:6-255) run on the same warp
// out of a shared-memory copy of the environment, and the feature tensors are written straight into
// the caller's (torch-owned) HBM tensors with 128-bit coalesced streaming stores.
//
// Exactness: integers everywhere except (a) the env reward, accumulated in float64 in the reference's
// order (env.py:181,256,288-291), (b) the shaped reward, float32 adds in the reference's order of terms
// plus numpy's pairwise float32 sum, and (c) feature quotients, which are host-built float32 tables
// of correctly rounded float64 quotients (what `np.array([...python floats...], dtype=float32)`
// produces).  Sort keys are dense ranks of the reference's float64 keys, ties broken by the original
// list order like Python's stable sort.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "mdx_internal.h"

namespace mdx {

#define MDX_FULL 0xffffffffu

struct Warp {
  uint16_t *start, *target, *t0, *dl;  // [Pp8] current episode schedule
  uint8_t* status;                     // [Pp16]
  uint16_t *wlist, *wdl, *alist;       // compact waiting list (+ remaining-deadline), active list
  uint32_t* wkey;                      // per-robot selection keys of the waiting list
  uint32_t* masks;                     // [3][W32] robots / waiting starts / active targets
  uint16_t *rcell, *rcarry, *rtgt, *rdl;  // [32] per robot
  float* stage;                        // 16-byte aligned staging for vec / gvec rows
};

__device__ __forceinline__ Warp carve(const KParams& p, unsigned char* base) {
  Warp w;
  uint16_t* s = reinterpret_cast<uint16_t*>(base + p.off_sched);
  w.start = s;
  w.target = s + p.Pp8;
  w.t0 = s + 2 * p.Pp8;
  w.dl = s + 3 * p.Pp8;
  w.status = base + p.off_status;
  w.wlist = reinterpret_cast<uint16_t*>(base + p.off_wlist);
  w.wdl = reinterpret_cast<uint16_t*>(base + p.off_wdl);
  w.alist = reinterpret_cast<uint16_t*>(base + p.off_alist);
  w.wkey = reinterpret_cast<uint32_t*>(base + p.off_wkey);
  w.masks = reinterpret_cast<uint32_t*>(base + p.off_masks);
  uint16_t* ri = reinterpret_cast<uint16_t*>(base + p.off_rinfo);
  w.rcell = ri;
  w.rcarry = ri + 32;
  w.rtgt = ri + 64;
  w.rdl = ri + 96;
  w.stage = reinterpret_cast<float*>(base + p.off_stage);
  return w;
}

__device__ __forceinline__ float one_if(uint32_t bit) { return __uint_as_float((0u - (bit & 1u)) & 0x3f800000u); }

__device__ __forceinline__ void st_stream4(float* ptr, float a, float b, float c, float d) {
  __stcs(reinterpret_cast<float4*>(ptr), make_float4(a, b, c, d));
}

// max(0, deadline - t) / T  and  t / T  (MAPPO/helper.py:107,155,163): table of correctly rounded
// float64 quotients, IEEE double division beyond the table (t keeps running past T).
__device__ __forceinline__ float tnorm(const KParams& p, int k) {
  if (p.T <= 0) return 0.0f;
  if (k < p.tq_len) return __ldg(p.tq + k);
  return __double2float_rn(__ddiv_rn((double)k, (double)p.T));
}

__device__ __forceinline__ void load_episode(const KParams& p, const Warp& w, int env, int slot, int lane) {
  const uint4* src = reinterpret_cast<const uint4*>(p.sched + ((size_t)slot * p.n_envs + env) * 4 * p.Pp8);
  uint4* dst = reinterpret_cast<uint4*>(w.start);
  const int n16 = p.Pp8 / 2;  // 4 rows * Pp8 u16 * 2 B / 16 B
  for (int i = lane; i < n16; i += 32) dst[i] = __ldg(src + i);
}

// numpy float32 add.reduce over R lane values (pairwise_sum: plain loop for n < 8, eight running
// partials for 8 <= n <= 128); every lane returns the same value.
__device__ __forceinline__ float warp_np_sum(float v, int R) {
  if (R < 8) {
    float s = 0.0f;
    for (int i = 0; i < R; ++i) s = __fadd_rn(s, __shfl_sync(MDX_FULL, v, i));
    return s;
  }
  float r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = __shfl_sync(MDX_FULL, v, j);
  int i = 8;
  for (; i < R - (R % 8); i += 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], __shfl_sync(MDX_FULL, v, i + j));
  }
  float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                        __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
  for (; i < R; ++i) res = __fadd_rn(res, __shfl_sync(MDX_FULL, v, i));
  return res;
}

// Copy n staged floats to global memory.  `stage` is 16-byte aligned and the row starts at
// stage[off] with off == (global float index & 3), so shared and global addresses are congruent
// modulo 16 bytes and the body moves as aligned float4.
__device__ __forceinline__ void flush_row(float* gdst, const float* stage, int off, int n, int lane) {
  const float* s = stage + off;
  int head = (4 - off) & 3;
  if (head > n) head = n;
  if (lane < head) __stcs(gdst + lane, s[lane]);
  const int body = (n - head) >> 2;
  const float4* s4 = reinterpret_cast<const float4*>(s + head);
  float4* g4 = reinterpret_cast<float4*>(gdst + head);
  for (int i = lane; i < body; i += 32) __stcs(g4 + i, s4[i]);
  const int done = head + (body << 2);
  if (lane < n - done) __stcs(gdst + done + lane, s[done + lane]);
}

// ------------------------------------------------------------------------------------------------
// Shaped team reward of one transition (MAPPO/helper.py:257-369), evaluated by the warp that owns
// the env.  Inputs per lane (robot): previous cell / carried id, new cell / carried id, decoded
// action.  `nw` entries of w.wlist hold the packages that were 'waiting' (start_time <= t) in the
// PREVIOUS state, in id order; w.status still holds the PREVIOUS statuses.
__device__ __forceinline__ float shaped_team_reward(const KParams& p, const Warp& w, int lane, int t_prev,
                                                    int t_cur, int cell0, int k0, int cell1, int k1, int mv,
                                                    int op, int nw, double env_reward) {
  const int R = p.R;
  const bool active = lane < R;
  float acc = 0.0f;
  const int r0 = __umulhi((uint32_t)cell0, p.mul_wout), c0 = cell0 - r0 * p.W_out;
  const int r1 = __umulhi((uint32_t)cell1, p.mul_wout), c1 = cell1 - r1 * p.W_out;
  // one pass over the previous waiting list, lanes = robots
  int best_d = 0x7fffffff, best_cell = -1;
  bool idle_near = false, at_new = false;
  for (int k = 0; k < nw; ++k) {
    const int sc = w.start[w.wlist[k]];
    const int sr = __umulhi((uint32_t)sc, p.mul_wout), scol = sc - sr * p.W_out;
    const int d = abs(sr - r0) + abs(scol - c0);
    if (d < best_d) {  // first minimum in id order, MAPPO/helper.py:346-349
      best_d = d;
      best_cell = sc;
    }
    idle_near |= (d <= 3);
    at_new |= (sc == cell1);
  }
  if (active) {
    const uint8_t st0 = k0 ? w.status[k0 - 1] : (uint8_t)ST_NONE;
    const bool tracked = k0 != 0 && (st0 == ST_WAITING || st0 == ST_IN_TRANSIT);
    const int tgt0 = k0 ? (int)w.target[k0 - 1] : -1;
    const int dl0 = k0 ? (int)w.dl[k0 - 1] : 0;
    // 1. successful pickup / delivery (:295-305)
    if (k0 == 0 && k1 != 0) {
      acc = __fadd_rn(acc, p.shaping[0]);
    } else if (k0 != 0 && k1 == 0) {
      if (tracked && cell1 == tgt0) acc = __fadd_rn(acc, (t_cur <= dl0) ? p.shaping[1] : p.shaping[2]);
    }
    // 2. wasted package ops (:308-327)
    if (op == 1) {
      if (k0 != 0) {
        acc = __fadd_rn(acc, p.shaping[4]);
      } else if (k1 == 0) {
        if (!at_new) acc = __fadd_rn(acc, p.shaping[4]);
      }
    } else if (op == 2) {
      if (k0 == 0) {
        acc = __fadd_rn(acc, p.shaping[5]);
      } else if (k1 != 0) {
        if (tracked && cell1 != tgt0) acc = __fadd_rn(acc, p.shaping[5]);
      }
    }
    // 3. movement (:330-356)
    const bool moved = cell0 != cell1;
    if (mv != MV_S && !moved) acc = __fadd_rn(acc, p.shaping[6]);
    int goal = -1;
    if (tracked)
      goal = tgt0;
    else if (nw > 0)
      goal = best_cell;
    if (goal >= 0 && moved) {
      const int gr = __umulhi((uint32_t)goal, p.mul_wout), gc = goal - gr * p.W_out;
      const int d_before = abs(r0 - gr) + abs(c0 - gc);
      const int d_after = abs(r1 - gr) + abs(c1 - gc);
      if (d_after < d_before)
        acc = __fadd_rn(acc, p.shaping[3]);
      else if (d_after > d_before)
        acc = __fadd_rn(acc, p.shaping[8]);
    }
    // 4. idle next to work (:359-367)
    if (!moved && mv == MV_S && k0 == 0 && idle_near) acc = __fadd_rn(acc, p.shaping[7]);
  }
  (void)t_prev;
  const float s = warp_np_sum(acc, R);
  // python float + np.float32 -> float32 under numpy >= 2
  return __fadd_rn(__double2float_rn(env_reward), s);
}

// Compact list of packages that are 'waiting' with start_time <= t, in id order.  Returns the count.
__device__ __forceinline__ int build_waiting(const KParams& p, const Warp& w, int lane, int t) {
  int nw = 0;
  const unsigned lt = (1u << lane) - 1u;
  for (int base = 0; base < p.P; base += 32) {
    const int j = base + lane;
    const bool wait = j < p.P && w.status[j] == ST_WAITING && (int)w.t0[j] <= t;
    const unsigned m = __ballot_sync(MDX_FULL, wait);
    if (wait) w.wlist[nw + __popc(m & lt)] = (uint16_t)j;
    nw += __popc(m);
  }
  __syncwarp();
  return nw;
}

// ------------------------------------------------------------------------------------------------
// Featurisation of the env held in shared memory (statuses, schedule) and registers (cell / carry of
// robot `lane`), MAPPO/helper.py:6-255.  `row` selects the output row.
__device__ __forceinline__ void featurize_env(const KParams& p, const Warp& w, const MapDev& map, int lane,
                                              int row, int t, int cell, int carry) {
  const int R = p.R, P = p.P, W32 = p.W32, HW = p.HW, Wo = p.W_out;
  const int Hm1 = p.H_out - 1, Wm1 = p.W_out - 1, NC = 2 * p.W_out - 1;
  const unsigned lt = (1u << lane) - 1u;
  const bool active = lane < R;
  const int r = active ? (int)__umulhi((uint32_t)cell, p.mul_wout) : 0;
  const int c = active ? cell - r * Wo : 0;

  // ---- per-robot info + plane masks ----------------------------------------------------------
  bool carrying_it = false;
  int own_tgt = 0xFFFF, own_rem = 0;
  if (active) {
    if (carry != 0 && w.status[carry - 1] == ST_IN_TRANSIT) {
      carrying_it = true;
      own_tgt = w.target[carry - 1];
      own_rem = max(0, (int)w.dl[carry - 1] - t);
    }
    w.rcell[lane] = (uint16_t)cell;
    w.rcarry[lane] = (uint16_t)carry;
    w.rtgt[lane] = (uint16_t)own_tgt;
    w.rdl[lane] = (uint16_t)own_rem;
  }
  for (int i = lane; i < 3 * W32; i += 32) w.masks[i] = 0u;
  __syncwarp();
  if (active) atomicOr(&w.masks[cell >> 5], 1u << (cell & 31));
  int nw = 0, na = 0;
  for (int base = 0; base < P; base += 32) {
    const int j = base + lane;
    const uint8_t st = j < P ? w.status[j] : (uint8_t)ST_NONE;
    const bool wait = j < P && st == ST_WAITING && (int)w.t0[j] <= t;
    const bool act = wait || (j < P && st == ST_IN_TRANSIT);
    if (wait) {
      const int sc = w.start[j];
      atomicOr(&w.masks[W32 + (sc >> 5)], 1u << (sc & 31));
    }
    if (act) {
      const int tc = w.target[j];
      atomicOr(&w.masks[2 * W32 + (tc >> 5)], 1u << (tc & 31));
    }
    const unsigned wm = __ballot_sync(MDX_FULL, wait), am = __ballot_sync(MDX_FULL, act);
    if (wait) {
      const int k = nw + __popc(wm & lt);
      w.wlist[k] = (uint16_t)j;
      w.wdl[k] = (uint16_t)max(0, (int)w.dl[j] - t);
    }
    if (act) {
      const int k = na + __popc(am & lt);
      if (k < p.MP) w.alist[k] = (uint16_t)j;
    }
    nw += __popc(wm);
    na += __popc(am);
  }
  __syncwarp();

  const uint32_t* gridb = map.grid_bits;
  const uint32_t* m_rob = w.masks;
  const uint32_t* m_ws = w.masks + W32;
  const uint32_t* m_at = w.masks + 2 * W32;

  // ---- obs [R,6,H,W] and gstate [4,H,W] ------------------------------------------------------
  if ((HW & 3) == 0) {
    // fast path: planes are whole float4s; 4 consecutive cells never straddle a mask word
    if (p.obs != nullptr) {
      const int n4 = (6 * HW) >> 2;
      for (int i = 0; i < R; ++i) {
        float* out = p.obs + ((size_t)row * R + i) * 6 * HW;
        const uint32_t rc = w.rcell[i], tg = w.rtgt[i];
        for (int q = lane; q < n4; q += 32) {
          const uint32_t e = (uint32_t)q << 2;
          const uint32_t ch = __umulhi(e, p.mul_plane);
          const uint32_t cl = e - ch * HW;
          const uint32_t wi = cl >> 5, sh = cl & 31u;
          const uint32_t oh_r = ((rc >> 2) == (cl >> 2)) ? (1u << (rc & 3u)) : 0u;
          const uint32_t oh_t = ((tg >> 2) == (cl >> 2)) ? (1u << (tg & 3u)) : 0u;
          uint32_t nib;
          if (ch == 0)
            nib = (__ldg(gridb + wi) >> sh) & 15u;
          else if (ch == 1)
            nib = oh_r;
          else if (ch == 2)
            nib = ((m_rob[wi] >> sh) & 15u) & ~oh_r;
          else if (ch == 3)
            nib = (m_ws[wi] >> sh) & 15u;
          else if (ch == 4)
            nib = (m_at[wi] >> sh) & 15u;
          else
            nib = oh_t;
          st_stream4(out + e, one_if(nib), one_if(nib >> 1), one_if(nib >> 2), one_if(nib >> 3));
        }
      }
    }
    if (p.gstate != nullptr) {
      float* out = p.gstate + (size_t)row * 4 * HW;
      const int n4 = HW;  // 4 planes * HW / 4
      for (int q = lane; q < n4; q += 32) {
        const uint32_t e = (uint32_t)q << 2;
        const uint32_t ch = __umulhi(e, p.mul_plane);
        const uint32_t cl = e - ch * HW;
        const uint32_t wi = cl >> 5, sh = cl & 31u;
        uint32_t word;
        if (ch == 0)
          word = __ldg(gridb + wi);
        else if (ch == 1)
          word = m_rob[wi];
        else if (ch == 2)
          word = m_ws[wi];
        else
          word = m_at[wi];
        const uint32_t nib = (word >> sh) & 15u;
        st_stream4(out + e, one_if(nib), one_if(nib >> 1), one_if(nib >> 2), one_if(nib >> 3));
      }
    }
  } else {
    // generic path (e.g. the 7x7 map): one coalesced scalar store per element
    if (p.obs != nullptr) {
      const int n = 6 * HW;
      for (int i = 0; i < R; ++i) {
        float* out = p.obs + ((size_t)row * R + i) * 6 * HW;
        const uint32_t rc = w.rcell[i], tg = w.rtgt[i];
        for (int e = lane; e < n; e += 32) {
          const uint32_t ch = __umulhi((uint32_t)e, p.mul_plane);
          const uint32_t cl = (uint32_t)e - ch * HW;
          const uint32_t wi = cl >> 5, sh = cl & 31u;
          uint32_t b;
          if (ch == 0)
            b = __ldg(gridb + wi) >> sh;
          else if (ch == 1)
            b = (rc == cl);
          else if (ch == 2)
            b = (m_rob[wi] >> sh) & (rc != cl ? 1u : 0u);
          else if (ch == 3)
            b = m_ws[wi] >> sh;
          else if (ch == 4)
            b = m_at[wi] >> sh;
          else
            b = (tg == cl);
          __stcs(out + e, one_if(b));
        }
      }
    }
    if (p.gstate != nullptr) {
      float* out = p.gstate + (size_t)row * 4 * HW;
      const int n = 4 * HW;
      for (int e = lane; e < n; e += 32) {
        const uint32_t ch = __umulhi((uint32_t)e, p.mul_plane);
        const uint32_t cl = (uint32_t)e - ch * HW;
        const uint32_t wi = cl >> 5, sh = cl & 31u;
        uint32_t word;
        if (ch == 0)
          word = __ldg(gridb + wi);
        else if (ch == 1)
          word = m_rob[wi];
        else if (ch == 2)
          word = m_ws[wi];
        else
          word = m_at[wi];
        __stcs(out + e, one_if(word >> sh));
      }
    }
  }

  const float* qr = map.qr;
  const float* qc = map.qc;

  // ---- vec [R, D]  (generate_vector_features, MAPPO/helper.py:68-165) -------------------------
  if (p.vec != nullptr) {
    const int D = p.D, K_o = p.K_o, K_p = p.K_p;
    const float tfrac = tnorm(p, t);
    // the two "target relative to the carrier itself" entries of an other-robot row (:125-136)
    float ot_r = 0.0f, ot_c = 0.0f;
    if (carrying_it) {
      const int tr = __umulhi((uint32_t)own_tgt, p.mul_wout), tc = own_tgt - tr * Wo;
      ot_r = __ldg(qr + (tr - r + Hm1));
      ot_c = __ldg(qc + (tc - c + Wm1));
    }
    const uint32_t wmask = (1u << p.key_wbits) - 1u;
    const int rpf = p.vec_rows_per_flush;
    for (int i0 = 0; i0 < R; i0 += rpf) {
      const int nrows = min(rpf, R - i0);
      const size_t g0 = ((size_t)row * R + i0) * D;
      const int off = (int)(g0 & 3);
      for (int x = lane; x < nrows * D + 4; x += 32) w.stage[x] = 0.0f;
      __syncwarp();
      for (int i = i0; i < i0 + nrows; ++i) {
        float* rowp = w.stage + off + (i - i0) * D;
        const int ri = __shfl_sync(MDX_FULL, r, i), ci = __shfl_sync(MDX_FULL, c, i);
        if (lane == i) {  // own block (:93-112) and global time (:163)
          rowp[0] = __ldg(qr + (r + Hm1));
          rowp[1] = __ldg(qc + (c + Wm1));
          rowp[2] = carry != 0 ? 1.0f : 0.0f;
          if (carrying_it) {
            const int tr = __umulhi((uint32_t)own_tgt, p.mul_wout), tc = own_tgt - tr * Wo;
            rowp[3] = __ldg(qr + (tr - r + Hm1));
            rowp[4] = __ldg(qc + (tc - c + Wm1));
            rowp[5] = tnorm(p, own_rem);
          }
          rowp[D - 1] = tfrac;
        }
        // other robots, stable-sorted by the float64 key rank (:115-141)
        const int dr = r - ri, dc = c - ci;
        uint32_t okey = 0xffffffffu;
        if (active && lane != i) okey = ((uint32_t)__ldg(map.rank + (dr + Hm1) * NC + (dc + Wm1)) << 5) | (uint32_t)lane;
        int pos = 0;
        for (int k = 0; k < R; ++k) pos += (__shfl_sync(MDX_FULL, okey, k) < okey) ? 1 : 0;
        if (active && lane != i && pos < K_o) {
          float* o = rowp + 6 + pos * 5;
          o[0] = __ldg(qr + (dr + Hm1));
          o[1] = __ldg(qc + (dc + Wm1));
          o[2] = carry != 0 ? 1.0f : 0.0f;
          o[3] = ot_r;
          o[4] = ot_c;
        }
        // waiting packages, stable-sorted by (remaining deadline, key rank) (:143-160): repeated
        // warp arg-min over 32-bit keys (rem * nranks + rank) << wbits | list index
        const int nsel = min(K_p, nw);
        if (nsel > 0) {
          uint32_t lm = 0xffffffffu;
          for (int k = lane; k < nw; k += 32) {
            const int sc = w.start[w.wlist[k]];
            const int sr = __umulhi((uint32_t)sc, p.mul_wout), scol = sc - sr * Wo;
            const uint32_t rk = __ldg(map.rank + (sr - ri + Hm1) * NC + (scol - ci + Wm1));
            const uint32_t key = (((uint32_t)w.wdl[k] * (uint32_t)p.nranks_max + rk) << p.key_wbits) | (uint32_t)k;
            w.wkey[k] = key;
            lm = min(lm, key);
          }
          for (int q = 0; q < nsel; ++q) {
            const uint32_t best = __reduce_min_sync(MDX_FULL, lm);
            if (lm == best) {  // exactly one lane: keys are unique
              const int k = (int)(best & wmask);
              const int j = w.wlist[k];
              const int sc = w.start[j], tc_ = w.target[j];
              const int sr = __umulhi((uint32_t)sc, p.mul_wout), scol = sc - sr * Wo;
              const int tr = __umulhi((uint32_t)tc_, p.mul_wout), tcol = tc_ - tr * Wo;
              float* o = rowp + 6 + K_o * 5 + q * 5;
              o[0] = __ldg(qr + (sr - ri + Hm1));
              o[1] = __ldg(qc + (scol - ci + Wm1));
              o[2] = __ldg(qr + (tr - ri + Hm1));
              o[3] = __ldg(qc + (tcol - ci + Wm1));
              o[4] = tnorm(p, (int)w.wdl[k]);
              w.wkey[k] = 0xffffffffu;
              lm = 0xffffffffu;
              for (int kk = lane; kk < nw; kk += 32) lm = min(lm, w.wkey[kk]);
            }
          }
        }
      }
      __syncwarp();
      flush_row(p.vec + g0, w.stage, off, nrows * D, lane);
      __syncwarp();
    }
  }

  // ---- gvec [Dg]  (convert_global_state vector part, MAPPO/helper.py:199-255) ------------------
  if (p.gvec != nullptr) {
    const int Dg = p.Dg, MR = p.MR, MP = p.MP;
    const size_t g0 = (size_t)row * Dg;
    const int off = (int)(g0 & 3);
    for (int x = lane; x < Dg + 4; x += 32) w.stage[x] = 0.0f;
    __syncwarp();
    float* g = w.stage + off;
    if (active && lane < MR) {
      float* o = g + lane * 6;
      o[0] = __ldg(qr + (r + Hm1));
      o[1] = __ldg(qc + (c + Wm1));
      o[2] = carry != 0 ? 1.0f : 0.0f;
      if (carrying_it) {
        const int tr = __umulhi((uint32_t)own_tgt, p.mul_wout), tc = own_tgt - tr * Wo;
        o[3] = __ldg(qr + (tr + Hm1));
        o[4] = __ldg(qc + (tc + Wm1));
        o[5] = tnorm(p, own_rem);
      }
    }
    const int nrow = min(na, MP);
    for (int k = lane; k < nrow; k += 32) {
      const int j = w.alist[k];
      const bool wait = w.status[j] == ST_WAITING;
      const int sc = w.start[j], tc_ = w.target[j];
      const int sr = __umulhi((uint32_t)sc, p.mul_wout), scol = sc - sr * Wo;
      const int tr = __umulhi((uint32_t)tc_, p.mul_wout), tcol = tc_ - tr * Wo;
      float* o = g + 6 * MR + 7 * k;
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
          if ((int)w.rcarry[ridx] == j + 1) {
            cn = MR > 1 ? __double2float_rn(__ddiv_rn((double)ridx, (double)(MR - 1))) : 0.0f;
            break;
          }
      }
      o[6] = cn;
    }
    if (lane == 0) g[Dg - 1] = tnorm(p, t);
    __syncwarp();
    flush_row(p.gvec + g0, w.stage, off, Dg, lane);
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// kMode bit 0: step (+ shaped reward, + optional auto-reset);  bit 1: featurise.
template <int kMode>
__global__ void __launch_bounds__(256) mdx_env_kernel(const __grid_constant__ KParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int widx = blockIdx.x * (blockDim.x >> 5) + wib;
  if (widx >= p.n) return;
  const Warp w = carve(p, smem_raw + (size_t)wib * p.smem_per_warp);
  const int env = p.env_ids ? __ldg(p.env_ids + widx) : widx;
  const int R = p.R, P = p.P, Wo = p.W_out;
  const bool active = lane < R;
  const MapDev& map = p.maps[p.env_map ? __ldg(p.env_map + env) : 0];

  int t = p.t[env];
  int ep = p.ep[env];
  load_episode(p, w, env, ep % p.E, lane);
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.status + (size_t)env * p.Pp16);
    uint4* dst = reinterpret_cast<uint4*>(w.status);
    for (int i = lane; i < (p.Pp16 >> 4); i += 32) dst[i] = src[i];
  }
  uint32_t rob = active ? p.robots[(size_t)env * R + lane] : 0u;
  int cell = active ? (int)(rob & 0xffffu) : (0x10000 + lane);
  int carry = active ? (int)(rob >> 16) : 0;
  __syncwarp();

  if (kMode & 1) {
    // ---------------- decode ----------------
    const int a = active ? (int)__ldg(p.actions + (size_t)widx * R + lane) : 3;
    int mv, op;
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
    for (int j = 0; j < R; ++j) {
      const int cj = __shfl_sync(MDX_FULL, cell, j);
      if (cj == tgt && j != lane) holder = j;
    }
    const bool wants = active && tgt != cell;
    const unsigned grp = __match_any_sync(MDX_FULL, tgt);
    const unsigned wants_mask = __ballot_sync(MDX_FULL, wants);
    const bool winner = wants && (((grp & wants_mask) & ((1u << lane) - 1u)) == 0u);
    unsigned mm = 0u;
    while (true) {
      const bool m = winner && (holder < 0 || ((mm >> holder) & 1u));
      const unsigned nm = __ballot_sync(MDX_FULL, m);
      if (nm == mm) break;
      mm = nm;
    }
    const bool moved = (mm >> lane) & 1u;
    const int cell1 = moved ? tgt : cell;
    const int n_moved = __popc(mm);  // every mover used one of L/R/U/D (env.py:255)

    // ---------------- package ops in robot order (env.py:260-292) ----------------
    const int nw_prev = build_waiting(p, w, lane, t);
    const bool want_pick = active && op == 1 && carry == 0;
    int picked = -1;
    if (__any_sync(MDX_FULL, want_pick)) {
      for (int k = 0; k < nw_prev; ++k) {
        const int j = w.wlist[k];
        if (want_pick && picked < 0 && (int)w.start[j] == cell1) picked = j;  // lowest id first
      }
    }
    const int tgt0 = carry ? (int)w.target[carry - 1] : -1;
    const int dl0 = carry ? (int)w.dl[carry - 1] : 0;
    const bool delivered = active && op == 2 && carry != 0 && cell1 == tgt0;
    const bool on_time = delivered && t <= dl0;
    const int carry1 = delivered ? 0 : (picked >= 0 ? picked + 1 : carry);
    // env reward in float64, reference order: move costs first, then deliveries by robot index
    const unsigned dmask = __ballot_sync(MDX_FULL, delivered), omask = __ballot_sync(MDX_FULL, on_time);
    double rew = 0.0;
    for (int k = 0; k < n_moved; ++k) rew = __dadd_rn(rew, p.move_cost);
    for (unsigned mleft = dmask; mleft; mleft &= mleft - 1u) {
      const int b = __ffs(mleft) - 1;
      rew = __dadd_rn(rew, ((omask >> b) & 1u) ? p.delivery_reward : p.delay_reward);
    }
    // shaped reward reads the PREVIOUS statuses, so evaluate it before they change
    float shaped = 0.0f;
    if (p.out_shaped != nullptr)
      shaped = shaped_team_reward(p, w, lane, t, t + 1, cell, carry, cell1, carry1, mv, op, nw_prev, rew);
    __syncwarp();
    if (delivered) w.status[carry - 1] = ST_DELIVERED;
    if (picked >= 0) w.status[picked] = ST_IN_TRANSIT;
    __syncwarp();
    // ---------------- time, spawn, termination (env.py:295-306, 127-137) ----------------
    const int t1 = t + 1;
    bool all_delivered = true;
    for (int base = 0; base < P; base += 32) {
      const int j = base + lane;
      bool ok = true;
      if (j < P) {
        if ((int)w.t0[j] == t1) w.status[j] = ST_WAITING;
        ok = w.status[j] == ST_DELIVERED;
      }
      all_delivered = all_delivered && __all_sync(MDX_FULL, ok);
    }
    const bool done = (t1 == p.T) || all_delivered;  // not latched
    double total = p.total_reward[env] + rew;
    if (lane == 0) {
      if (p.out_reward) p.out_reward[widx] = rew;
      if (p.out_shaped) p.out_shaped[widx] = shaped;
      if (p.out_done) p.out_done[widx] = done ? 1 : 0;
    }
    // cumulative per-env counters
    {
      const int n_pick = __popc(__ballot_sync(MDX_FULL, picked >= 0));
      uint32_t* cn = p.counters + (size_t)env * 8;
      if (lane == 0) {
        cn[C_MOVES] += (uint32_t)n_moved;
        cn[C_ONTIME] += (uint32_t)__popc(omask);
        cn[C_LATE] += (uint32_t)__popc(dmask & ~omask);
        cn[C_PICKUPS] += (uint32_t)n_pick;
        cn[C_STEPS] += 1u;
        if (done) {
          cn[C_EPISODES] += 1u;
          cn[C_EPLEN] += (uint32_t)t1;
        }
      }
    }
    cell = active ? cell1 : cell;
    carry = carry1;
    t = t1;
    __syncwarp();
    // ---------------- on-device auto-reset (MAPPO/trainer.py:229-234) ----------------
    if (done && (p.flags & 1u)) {
      if (ep + 1 < __ldg(p.ep_avail + env)) {
        ep += 1;
        const int slot = ep % p.E;
        load_episode(p, w, env, slot, lane);
        __syncwarp();
        for (int j = lane; j < p.Pp16; j += 32)
          w.status[j] = (j < P && w.t0[j] == 0) ? (uint8_t)ST_WAITING : (uint8_t)ST_NONE;
        if (active) {
          cell = __ldg(p.sched_robots + ((size_t)slot * p.n_envs + env) * 32 + lane);
          carry = 0;
        }
        t = 0;
        total = 0.0;
        __syncwarp();
      } else if (lane == 0) {
        p.counters[(size_t)env * 8 + C_STARVED] = 1u;
      }
    }
    // ---------------- write the state back ----------------
    if (active) p.robots[(size_t)env * R + lane] = (uint32_t)cell | ((uint32_t)carry << 16);
    {
      uint4* dst = reinterpret_cast<uint4*>(p.status + (size_t)env * p.Pp16);
      const uint4* src = reinterpret_cast<const uint4*>(w.status);
      for (int i = lane; i < (p.Pp16 >> 4); i += 32) dst[i] = src[i];
    }
    if (lane == 0) {
      p.t[env] = t;
      p.ep[env] = ep;
      p.total_reward[env] = total;
    }
  }

  if (kMode & 2) featurize_env(p, w, map, lane, widx, t, cell, carry);
}

// compute_shaped_rewards on injected "current" robots against the stored state as "previous".
__global__ void __launch_bounds__(256) mdx_shape_kernel(const __grid_constant__ KParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int env = blockIdx.x * (blockDim.x >> 5) + wib;
  if (env >= p.n) return;
  const Warp w = carve(p, smem_raw + (size_t)wib * p.smem_per_warp);
  const int R = p.R;
  const bool active = lane < R;
  const int t = p.t[env];
  load_episode(p, w, env, p.ep[env] % p.E, lane);
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.status + (size_t)env * p.Pp16);
    uint4* dst = reinterpret_cast<uint4*>(w.status);
    for (int i = lane; i < (p.Pp16 >> 4); i += 32) dst[i] = src[i];
  }
  __syncwarp();
  const uint32_t rob0 = active ? p.robots[(size_t)env * R + lane] : 0u;
  const uint32_t rob1 = active ? p.cur_robots[(size_t)env * R + lane] : 0u;
  const int a = active ? (int)p.actions[(size_t)env * R + lane] : 3;
  int mv, op;
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
  const int nw = build_waiting(p, w, lane, t);
  const float s = shaped_team_reward(p, w, lane, t, p.cur_t[env], (int)(rob0 & 0xffffu), (int)(rob0 >> 16),
                                     (int)(rob1 & 0xffffu), (int)(rob1 >> 16), mv, op, nw, p.global_reward[env]);
  if (lane == 0) p.out_shaped[env] = s;
}

// Apply uploaded reset blocks: env ids[i] advances to its next episode (env.py:81-125 minus the draws,
// which happened on the host).  One thread per (env, package slot) would be overkill; one warp per env.
__global__ void __launch_bounds__(256) mdx_reset_kernel(const __grid_constant__ KParams p) {
  const int lane = threadIdx.x & 31;
  const int widx = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (widx >= p.n) return;
  const int env = p.env_ids ? p.env_ids[widx] : widx;
  const int ep = p.ep[env] + (int)(p.flags & 1u);  // flags bit 0: advance (0 = re-initialise in place)
  const int slot = ep % p.E;
  const uint16_t* t0 = p.sched + ((size_t)slot * p.n_envs + env) * 4 * p.Pp8 + 2 * p.Pp8;
  for (int j = lane; j < p.Pp16; j += 32)
    p.status[(size_t)env * p.Pp16 + j] = (j < p.P && t0[j] == 0) ? (uint8_t)ST_WAITING : (uint8_t)ST_NONE;
  if (lane < p.R) p.robots[(size_t)env * p.R + lane] = p.sched_robots[((size_t)slot * p.n_envs + env) * 32 + lane];
  if (lane == 0) {
    p.t[env] = 0;
    p.ep[env] = ep;
    p.total_reward[env] = 0.0;
  }
}

// Scatter freshly sampled reset blocks from a packed staging buffer into the ring.
// staging record i: [4*Pp8 u16 schedule][32 u16 robot cells]; dst env = ids[i], slot = slots[i].
__global__ void mdx_upload_kernel(uint16_t* sched, uint16_t* sched_robots, const uint16_t* staging,
                                  const int32_t* ids, const int32_t* slots, int n, int n_envs, int Pp8) {
  const int rec = 4 * Pp8 + 32;
  const int per = rec / 8;  // uint4 per record (rec is a multiple of 8 u16)
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t i = gid / per;
  const int x = (int)(gid % per);
  if (i >= (size_t)n) return;
  const int env = ids[i], slot = slots[i];
  const uint4 v = reinterpret_cast<const uint4*>(staging + i * rec)[x];
  const int nsched = (4 * Pp8) / 8;
  if (x < nsched)
    reinterpret_cast<uint4*>(sched + ((size_t)slot * n_envs + env) * 4 * Pp8)[x] = v;
  else
    reinterpret_cast<uint4*>(sched_robots + ((size_t)slot * n_envs + env) * 32)[x - nsched] = v;
}

// sum the per-env counters into 8 int64 (then all-reduced across ranks by NCCL).
__global__ void mdx_stats_kernel(const uint32_t* counters, int n_envs, unsigned long long* out8) {
  __shared__ unsigned long long acc[8];
  if (threadIdx.x < 8) acc[threadIdx.x] = 0ull;
  __syncthreads();
  const int k = threadIdx.x & 7;
  unsigned long long s = 0ull;
  for (size_t e = (size_t)blockIdx.x * (blockDim.x >> 3) + (threadIdx.x >> 3); e < (size_t)n_envs;
       e += (size_t)gridDim.x * (blockDim.x >> 3))
    s += counters[e * 8 + k];
  atomicAdd(&acc[k], s);
  __syncthreads();
  if (threadIdx.x < 8) atomicAdd(&out8[threadIdx.x], acc[threadIdx.x]);
}

}  // namespace mdx
