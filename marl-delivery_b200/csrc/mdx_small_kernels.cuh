// The non-templated kernels of libmdx (shaped reward on injected states, reset apply, ring upload, counter
// reduction).  Included by mdx_api.cu only.
#pragma once
#include "mdx_kernels.cuh"

namespace mdx {

// compute_shaped_rewards on injected "current" robots against the stored state as "previous".
__global__ void __launch_bounds__(128) mdx_shape_kernel(const __grid_constant__ KParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int wib = threadIdx.x >> 5;
  const int env = blockIdx.x * (blockDim.x >> 5) + wib;
  if (env >= p.n) return;
  const Grp<32> g(threadIdx.x & 31);
  const int lane = g.sl;
  const Env w = carve(p, smem_raw + kLutBytes + (size_t)wib * p.smem_per_warp);
  const int R = p.R;
  const bool active = lane < R;
  const int t = p.t[env];
  load_episode<32>(p, w, env, lane);
  {
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
  const int nw = build_waiting<32>(p, w, g, t);
  const MapDev map = p.env_map ? p.maps[p.env_map[env]] : p.map0;
  if (p.out_agent != nullptr && active) {  // IDQ reward_shaping on the injected transition
    const int k0 = (int)(rob0 >> 16);
    const int tg = k0 ? (int)w.target[k0 - 1] : -1;
    p.out_agent[(size_t)env * R + lane] = idq_agent_reward(w, nw, (rob0 & 0xffffu) != (rob1 & 0xffffu), op, r0, c0, k0, (int)(rob1 >> 16),
                                                           (int)(rob1 & 0xffffu) == tg, p.cur_t[env], k0 ? (int)w.dl[k0 - 1] : 0, p.idq_ops_int);
  }
  if (p.out_shaped == nullptr) return;
  float s;
  if (p.shaping_bfs) {
    const Scan sc = scan_waiting<32, false, true>(w, g, R, nw, r0, c0, r1, c1, map.bfs, p.HW, p.W_out);
    s = shaped_team_reward<32, true>(p, w, g, p.cur_t[env], r0, c0, (int)(rob0 >> 16), r1, c1, (int)(rob1 >> 16), mv, op, nw, sc,
                                     p.global_reward[env], map.bfs);
  } else {
    const Scan sc = scan_waiting<32, false, false>(w, g, R, nw, r0, c0, r1, c1, nullptr, 0, 0);
    s = shaped_team_reward<32, false>(p, w, g, p.cur_t[env], r0, c0, (int)(rob0 >> 16), r1, c1, (int)(rob1 >> 16), mv, op, nw, sc,
                                      p.global_reward[env], nullptr);
  }
  if (lane == 0) p.out_shaped[env] = s;
}

// Apply uploaded reset blocks: env ids[i] advances to its next episode (env.py:81-125 minus the draws,
// which happened on the host).  One warp per env.
__global__ void __launch_bounds__(256) mdx_reset_kernel(const __grid_constant__ KParams p) {
  const int lane = threadIdx.x & 31;
  const int widx = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (widx >= p.n) return;
  const int env = p.env_ids ? p.env_ids[widx] : widx;
  const int ep = p.ep[env] + (int)(p.flags & 1u);  // flags bit 0: advance (0 = re-initialise in place)
  const int slot = ep % p.E;
  const uint16_t* blk = p.sched + ((size_t)slot * p.n_envs + env) * 4 * p.Pp8;
  const uint16_t* t0 = blk + 2 * p.Pp8;
  {  // the block becomes the env's current-episode schedule
    const uint4* src = reinterpret_cast<const uint4*>(blk);
    uint4* cur = reinterpret_cast<uint4*>(p.cur + (size_t)env * 4 * p.Pp8);
    for (int i = lane; i < (p.Pp8 >> 1); i += 32) cur[i] = src[i];
  }
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
