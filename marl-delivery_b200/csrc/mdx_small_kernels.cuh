// The non-templated kernels of libmdx (shaped reward on injected states, reset, generator seeding, counter
// reduction).  Included by mdx_api.cu only.
#pragma once
#include "mdx_kernels.cuh"

namespace mdx {

// compute_shaped_rewards on injected "current" robots against the stored state as "previous".
// kTrk: persistent_packages_at_prev_state is the env's tracker table (MDX_CFG_STALE_TRACKER), not the env truth.
template <bool kTrk>
__global__ void __launch_bounds__(128) mdx_shape_kernel(const __grid_constant__ KParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int wib = threadIdx.x >> 5;
  const int env = blockIdx.x * (blockDim.x >> 5) + wib;
  if (env >= p.n) return;
  const Grp<32> g(threadIdx.x & 31);
  const int lane = g.sl;
  unsigned char* base = smem_raw + kLutBytes + (size_t)wib * p.smem_per_warp;
  Env w = carve(p, base);
  const int R = p.R;
  const bool active = lane < R;
  const int t = p.t[env];
  int Pn = p.P;
  if constexpr (kTrk) {
    w = carve_tracker(p, base, w);
    Pn = p.trk_n[env];
    const uint4* src = reinterpret_cast<const uint4*>(p.trk + (size_t)env * 5 * p.Pp8);
    uint4* dst = reinterpret_cast<uint4*>(w.pid);
    for (int i = lane; i < (5 * p.Pp8) >> 3; i += 32) dst[i] = src[i];
    const uint4* ssrc = reinterpret_cast<const uint4*>(p.trk_status + (size_t)env * p.Pp16);
    uint4* sdst = reinterpret_cast<uint4*>(w.status);
    for (int i = lane; i < (p.Pp16 >> 4); i += 32) sdst[i] = ssrc[i];
    for (int j = lane; j < p.Pp8; j += 32) w.slot[j] = (uint16_t)kNoSlot;
    g.sync();
    for (int k = lane; k < Pn; k += 32) w.slot[w.pid[k] - 1] = (uint16_t)k;
  } else {
    load_episode<32>(p, w, env, lane);
    const uint4* src = reinterpret_cast<const uint4*>(p.status + (size_t)env * p.Pp16);
    uint4* dst = reinterpret_cast<uint4*>(w.status);
#pragma unroll 1
    for (int i = lane; i < (p.Pp16 >> 4); i += 32) dst[i] = src[i];
  }
  g.sync();
  const uint32_t rob0 = active ? p.robots[(size_t)env * R + lane] : 0u;
  const uint32_t rob1 = active ? p.cur_robots[(size_t)env * R + lane] : 0u;
  int mv, op;
  decode_action(active ? (int)p.actions[(size_t)env * R + lane] : 3, mv, op);
  const int r0 = (int)__umulhi(rob0 & 0xffffu, p.mul_wout), c0 = (int)(rob0 & 0xffffu) - r0 * p.W_out;
  const int r1 = (int)__umulhi(rob1 & 0xffffu, p.mul_wout), c1 = (int)(rob1 & 0xffffu) - r1 * p.W_out;
  const int nw = build_waiting<32>(p, w, g, t, Pn);
  const MapDev map = p.env_map ? p.maps[p.env_map[env]] : p.map0;
  if (!kTrk && p.out_agent != nullptr && active) {  // IDQ reward_shaping on the injected transition
    const int k0 = (int)(rob0 >> 16);
    const int tg = k0 ? (int)w.target[k0 - 1] : -1;
    p.out_agent[(size_t)env * R + lane] = idq_agent_reward(w, nw, (rob0 & 0xffffu) != (rob1 & 0xffffu), op, r0, c0, k0, (int)(rob1 >> 16),
                                                           (int)(rob1 & 0xffffu) == tg, p.cur_t[env], k0 ? (int)w.dl[k0 - 1] : 0, p.idq_ops_int);
  }
  if (p.out_shaped == nullptr) return;
  float s;
  if (!kTrk && p.shaping_bfs) {
    const Scan sc = scan_waiting<32, false, true>(w, g, R, nw, r0, c0, r1, c1, map.bfs, p.HW, p.W_out);
    s = shaped_team_reward<32, true>(p, w, g, p.cur_t[env], r0, c0, (int)(rob0 >> 16), r1, c1, (int)(rob1 >> 16), mv, op, nw, sc,
                                     p.global_reward[env], map.bfs);
  } else {
    const Scan sc = scan_waiting<32, false, false>(w, g, R, nw, r0, c0, r1, c1, nullptr, 0, 0);
    s = shaped_team_reward<32, false, kTrk>(p, w, g, p.cur_t[env], r0, c0, (int)(rob0 >> 16), r1, c1, (int)(rob1 >> 16), mv, op, nw, sc,
                                            p.global_reward[env], nullptr);
  }
  if (lane == 0) p.out_shaped[env] = s;
}

// Environment.reset (env.py:81-125) for the listed envs: the env's next reset block is drawn on the device from its own
// MT19937 stream (mdx_sampler.cuh), then applied.  One warp per env; dynamic shared memory = 8 * Pp8 bytes per warp.
// flags bit 0: count the reset in the episode cursor (0 = the constructor's own reset, env.py:41).
__global__ void __launch_bounds__(256) mdx_reset_kernel(const __grid_constant__ KParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int widx = blockIdx.x * (blockDim.x >> 5) + wib;
  if (widx >= p.n) return;
  const int env = p.env_ids ? p.env_ids[widx] : p.env_base + widx;
  const MapDev map = p.env_map ? p.maps[p.env_map[env]] : p.map0;
  uint16_t* s_sched = reinterpret_cast<uint16_t*>(smem_raw + (size_t)wib * 8 * p.Pp8);
  const int cell = sample_episode<32>(p, map.free_cells, map.n_free, map.H, s_sched, env);
  const uint16_t* t0 = s_sched + 2 * p.Pp8;
  for (int j = lane; j < p.Pp16; j += 32)
    p.status[(size_t)env * p.Pp16 + j] = (j < p.P && t0[j] == 0) ? (uint8_t)ST_WAITING : (uint8_t)ST_NONE;
  if (lane < p.R) p.robots[(size_t)env * p.R + lane] = (uint32_t)cell;
  if (lane == 0) {
    p.t[env] = 0;
    p.ep[env] += (int)(p.flags & 1u);
    p.total_reward[env] = 0.0;
  }
}

// init_genrand(seed) for every env (numpy: RandomState(int seed)); one thread per env.
__global__ void mdx_seed_kernel(uint32_t* mt, int32_t* mtpos, const uint32_t* seeds, int n) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= n) return;
  uint32_t s = seeds[env];
  uint32_t* m = mt + (size_t)env * 624;
#pragma unroll 1
  for (int i = 0; i < 624; ++i) {
    m[i] = s;
    s = 1812433253u * (s ^ (s >> 30)) + (uint32_t)(i + 1);
  }
  mtpos[env] = 624;
}

// Packed host <-> device state rows (include/mdx.h: mdx_get_states / mdx_set_states / mdx_step_host).  Row layout, in 4-byte
// words: [t, episode, map_id, 0, total_reward lo, hi][robots: R words][schedule: 2*Pp8 words][status: Pp16/4 words], padded
// to a multiple of 4 words.  One warp per listed env; env = ids ? ids[i] : base + i.
struct RowArgs {
  uint32_t* rows;       // [n][row_words]
  const int32_t* ids;   // device ids or nullptr
  int32_t base, n, row_words, off_robots, off_sched, off_status;
};

__global__ void mdx_gather_rows_kernel(const __grid_constant__ KParams p, const RowArgs a) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= a.n) return;
  const int env = a.ids ? a.ids[i] : a.base + i;
  uint32_t* row = a.rows + (size_t)i * a.row_words;
  if (lane == 0) {
    row[0] = (uint32_t)p.t[env];
    row[1] = (uint32_t)p.ep[env];
    row[2] = p.env_map ? (uint32_t)p.env_map[env] : 0u;
    row[3] = 0u;
    const unsigned long long tr = (unsigned long long)__double_as_longlong(p.total_reward[env]);
    row[4] = (uint32_t)tr;
    row[5] = (uint32_t)(tr >> 32);
  }
  if (lane < p.R) row[a.off_robots + lane] = p.robots[(size_t)env * p.R + lane];
  const uint32_t* sched = reinterpret_cast<const uint32_t*>(p.cur + (size_t)env * 4 * p.Pp8);
  for (int k = lane; k < 2 * p.Pp8; k += 32) row[a.off_sched + k] = sched[k];
  const uint32_t* st = reinterpret_cast<const uint32_t*>(p.status + (size_t)env * p.Pp16);
  for (int k = lane; k < (p.Pp16 >> 2); k += 32) row[a.off_status + k] = st[k];
  // padding words are zero, so that equal states give byte-identical rows
  for (int k = a.off_robots + p.R + lane; k < a.off_sched; k += 32) row[k] = 0u;
  for (int k = a.off_status + (p.Pp16 >> 2) + lane; k < a.row_words; k += 32) row[k] = 0u;
}

// The inverse: rows (validated on the host) -> env state.  The episode cursor, the map id and the generator are not touched.
__global__ void mdx_scatter_rows_kernel(const __grid_constant__ KParams p, const RowArgs a) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= a.n) return;
  const int env = a.ids ? a.ids[i] : a.base + i;
  const uint32_t* row = a.rows + (size_t)i * a.row_words;
  if (lane == 0) {
    p.t[env] = (int32_t)row[0];
    p.total_reward[env] = __longlong_as_double((long long)(((unsigned long long)row[5] << 32) | row[4]));
  }
  if (lane < p.R) p.robots[(size_t)env * p.R + lane] = row[a.off_robots + lane];
  uint32_t* sched = reinterpret_cast<uint32_t*>(p.cur + (size_t)env * 4 * p.Pp8);
  for (int k = lane; k < 2 * p.Pp8; k += 32) sched[k] = row[a.off_sched + k];
  uint32_t* st = reinterpret_cast<uint32_t*>(p.status + (size_t)env * p.Pp16);
  for (int k = lane; k < (p.Pp16 >> 2); k += 32) st[k] = row[a.off_status + k];
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
