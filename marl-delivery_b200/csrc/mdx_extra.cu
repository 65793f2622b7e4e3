// Callers either side of the hot path: the batched greedy policy on per-map
// all-pairs BFS tables (f2), GAE over a device-resident rollout (f4) and env-state checkpointing (f4).
// See include/mdx.h for the contract of every entry point.
#include <cstring>
#include <deque>
#include <vector>

#include "mdx_host.h"
#include "mdx_alt.cuh"

using namespace mdx;

namespace mdx {

// ------------------------------------------------------------------------------------------------
// All-pairs BFS (greedyagent.py:6-29 run from every free cell), output-plane cell ids.
// ------------------------------------------------------------------------------------------------
static void bfs_all_pairs(const MapHost& m, int Ho, int Wo, uint16_t* out) {
  const int HW = Ho * Wo;
  std::fill(out, out + (size_t)HW * HW, (uint16_t)0xFFFF);
  std::vector<int> queue((size_t)m.rows * m.cols);
  for (int gr = 0; gr < m.rows; ++gr)
    for (int gc = 0; gc < m.cols; ++gc) {
      if (m.cells[(size_t)gr * m.cols + gc] != 0) continue;
      uint16_t* d = out + (size_t)(gr * Wo + gc) * HW;
      int head = 0, tail = 0;
      d[gr * Wo + gc] = 0;
      queue[(size_t)tail++] = gr * m.cols + gc;
      while (head < tail) {
        const int cur = queue[(size_t)head++];
        const int r = cur / m.cols, c = cur % m.cols;
        const uint16_t dc = d[r * Wo + c];
        const int nr[4] = {r - 1, r + 1, r, r}, nc[4] = {c, c, c - 1, c + 1};
        for (int k = 0; k < 4; ++k) {
          if (nr[k] < 0 || nr[k] >= m.rows || nc[k] < 0 || nc[k] >= m.cols) continue;
          if (m.cells[(size_t)nr[k] * m.cols + nc[k]] != 0) continue;
          uint16_t& dn = d[nr[k] * Wo + nc[k]];
          if (dn != 0xFFFF) continue;
          dn = (uint16_t)(dc + 1);
          queue[(size_t)tail++] = nr[k] * m.cols + nc[k];
        }
      }
    }
}

int ensure_bfs(mdx_env* e) {
  if (e->d_bfs) return MDX_OK;
  const int HW = e->base.HW;
  if (HW > 4096) return fail(MDX_E_INVALID, "all-pairs BFS tables are limited to maps of at most 4096 cells (have %d)", HW);
  DeviceGuard g(e->device);
  const size_t per = (size_t)HW * HW;
  const int nm = (int)e->maps.size();
  std::vector<uint16_t> host(per * (size_t)nm);
  for (int m = 0; m < nm; ++m) bfs_all_pairs(e->maps[(size_t)m], e->H_out, e->W_out, host.data() + per * (size_t)m);
  CU(cudaMalloc(&e->d_bfs, sizeof(uint16_t) * host.size()));
  CU(cudaMemcpy(e->d_bfs, host.data(), sizeof(uint16_t) * host.size(), cudaMemcpyHostToDevice));
  std::vector<MapDev> md((size_t)nm);
  CU(cudaMemcpy(md.data(), e->d_maps, sizeof(MapDev) * (size_t)nm, cudaMemcpyDeviceToHost));
  for (int m = 0; m < nm; ++m) md[(size_t)m].bfs = e->d_bfs + per * (size_t)m;
  CU(cudaMemcpy(e->d_maps, md.data(), sizeof(MapDev) * (size_t)nm, cudaMemcpyHostToDevice));
  e->base.map0 = md[0];
  return MDX_OK;
}

// ------------------------------------------------------------------------------------------------
// GreedyAgents.get_actions, one warp per environment, robots are lanes (greedyagent.py:42-174).
//
// The agent's package list is reconstructed in closed form from the env's schedule: init_agents stores the
// packages visible at t = 0 and the first get_actions call appends the same packages again (is_init is only
// set inside get_actions, :52-62,121-129), later calls append the packages spawning at that step.  With n0
// packages at t = 0 the list entry j therefore holds package id (j < n0 ? j + 1 : j - n0 + 1), and the list
// has n0 + #{start_time <= t} entries at time t.  The reference indexes that list by `package_id - 1`
// everywhere (:143,147,165-168) -- for ids beyond the t = 0 packages that is a different entry -- and so
// does this kernel.
// ------------------------------------------------------------------------------------------------
struct GreedyArgs {
  uint32_t* gstate;   // [n_envs][32]
  uint32_t* gfree;    // [n_envs][fw]
  int32_t fw;
  uint8_t* out;       // [n][R]
};

__global__ void __launch_bounds__(128) mdx_greedy_kernel(const __grid_constant__ KParams p, const GreedyArgs a) {
  extern __shared__ uint32_t sm_free[];
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int widx = blockIdx.x * (blockDim.x >> 5) + wib;
  if (widx >= p.n) return;
  const int env = p.env_ids ? __ldg(p.env_ids + widx) : p.env_base + widx;
  if ((unsigned)env >= (unsigned)p.n_envs) return;  // a bad caller-supplied id never becomes an out-of-bounds access (row left unwritten)
  uint32_t* fr = sm_free + (size_t)wib * a.fw;
  MapDev map = p.map0;
  if (p.env_map) map = p.maps[__ldg(p.env_map + env)];
  const int R = p.R, P = p.P, Wo = p.W_out, HW = p.HW;
  const int t = p.t[env];
  const uint16_t* S = p.cur + (size_t)env * 4 * p.Pp8;  // start cells of the current episode
  const uint16_t* G = S + p.Pp8;                                                // target cells
  const uint16_t* T0 = S + 2 * p.Pp8;                                           // start times
  int n0 = 0, cnt = 0;
  for (int base = 0; base < P; base += 32) {
    const int j = base + lane;
    const int t0 = j < P ? (int)__ldg(T0 + j) : 0x7fffffff;
    n0 += __popc(__ballot_sync(FULL, t0 == 0));
    cnt += __popc(__ballot_sync(FULL, t0 <= t));
  }
  const int L = n0 + cnt;  // entries of the agent's package list after this call's update_inner_state
  const bool active = lane < R;
  const uint32_t rob = active ? p.robots[(size_t)env * R + lane] : 0u;
  const int cell = (int)(rob & 0xffffu), carry = (int)(rob >> 16);
  uint32_t gs = 0u;
  if (t == 0) {  // GreedyAgents() + init_agents(state): every robot 'free', every package free
    for (int i = lane; i < a.fw; i += 32) fr[i] = 0xffffffffu;
  } else {
    for (int i = lane; i < a.fw; i += 32) fr[i] = a.gfree[(size_t)env * a.fw + i];
    if (active) gs = a.gstate[(size_t)env * 32 + lane];
  }
  int tgt = (int)(gs & 0xffffu);
  const int prev_carry = (int)(gs >> 16);
  // update_inner_state (:102-117): only a robot that WAS carrying refreshes its target from the env
  if (prev_carry != 0) tgt = carry;
  __syncwarp();
  const int r = (int)__umulhi((uint32_t)cell, p.mul_wout), c = cell - r * Wo;
  bool fresh = false;
  // robots in index order: a 'free' robot takes the Manhattan-nearest entry whose flag is still set
  // (first minimum in list order, :149-160) and clears the flag AT INDEX id - 1 (:163)
  for (int i = 0; i < R; ++i) {
    if (__shfl_sync(FULL, tgt, i) != 0) continue;
    const int ri = __shfl_sync(FULL, r, i), ci = __shfl_sync(FULL, c, i);
    uint32_t best = 0xffffffffu;
    for (int j = lane; j < L; j += 32) {
      if ((fr[j >> 5] >> (j & 31)) & 1u) {
        const int id = j < n0 ? j + 1 : j - n0 + 1;
        const uint32_t sc = __ldg(S + id - 1);
        const int sr = (int)__umulhi(sc, p.mul_wout);
        const int d = abs(sr - ri) + abs((int)sc - sr * Wo - ci);
        best = min(best, ((uint32_t)d << 16) | (uint32_t)j);
      }
    }
    best = __reduce_min_sync(FULL, best);
    if (best != 0xffffffffu) {
      const int j = (int)(best & 0xffffu);
      const int id = j < n0 ? j + 1 : j - n0 + 1;
      if (lane == 0) fr[(id - 1) >> 5] &= ~(1u << ((id - 1) & 31));
      if (lane == i) {
        tgt = id;
        fresh = true;
      }
      __syncwarp();
    }
  }
  // update_move_to_target (:64-100) for every robot that has a target
  int code = MV_S;  // ('S', '0')
  if (active && tgt != 0) {
    const int idx = tgt - 1;                                 // list index used by the reference
    const int id = idx < n0 ? idx + 1 : idx - n0 + 1;        // package that entry really holds
    const bool to_target = !fresh && carry != 0;
    const int goal = (int)__ldg((to_target ? G : S) + id - 1);
    const int gr = (int)__umulhi((uint32_t)goal, p.mul_wout), gc = goal - gr * Wo;
    int dist = abs(gr - r) + abs(gc - c);
    int mv = MV_S, act = 0;
    if (dist >= 1) {
      // run_bfs (:6-41): first of U, D, L, R whose BFS distance to the goal is one less than ours
      const uint16_t* d = map.bfs + (size_t)goal * HW;
      const int ds = (int)__ldg(d + cell);
      if (ds == 0xFFFF) {
        dist = 100000;
      } else {
        dist = ds;
        if (r > 0 && (int)__ldg(d + cell - Wo) == ds - 1) {
          mv = MV_U;
          dist = ds - 1;
        } else if (r + 1 < map.H && (int)__ldg(d + cell + Wo) == ds - 1) {
          mv = MV_D;
          dist = ds - 1;
        } else if (c > 0 && (int)__ldg(d + cell - 1) == ds - 1) {
          mv = MV_L;
          dist = ds - 1;
        } else if (c + 1 < map.W && (int)__ldg(d + cell + 1) == ds - 1) {
          mv = MV_R;
          dist = ds - 1;
        }
      }
      if (dist == 0) act = to_target ? 2 : 1;
    } else {
      act = to_target ? 2 : 1;
    }
    code = mv + 5 * act;
  }
  if (active) {
    a.out[(size_t)widx * R + lane] = (uint8_t)code;
    a.gstate[(size_t)env * 32 + lane] = (uint32_t)tgt | ((uint32_t)carry << 16);
  }
  for (int i = lane; i < a.fw; i += 32) a.gfree[(size_t)env * a.fw + i] = fr[i];
}

// ------------------------------------------------------------------------------------------------
// GAE (MAPPO/trainer.py:288-301): one thread per env walks its column of the [T, N] rollout backwards;
// every load/store is coalesced across envs.  float32, no FMA contraction (-fmad=false), the reference's
// order of operations.
// ------------------------------------------------------------------------------------------------
#ifndef MDX_GAE_CHUNK
#define MDX_GAE_CHUNK 16
#endif
#ifndef MDX_GAE_THREADS
#define MDX_GAE_THREADS 128
#endif
#ifndef MDX_GAE_DB
#define MDX_GAE_DB 1
#endif
constexpr int kGaeChunk = MDX_GAE_CHUNK;  // time steps whose loads are issued together before the dependent recurrence runs
constexpr int kGaeThreads = MDX_GAE_THREADS;

struct GaeChunk {
  float r[kGaeChunk], v[kGaeChunk];
  uint8_t d[kGaeChunk];
};

__device__ __forceinline__ void gae_load(GaeChunk& c, const float* __restrict__ rew, const float* __restrict__ val,
                                         const uint8_t* __restrict__ done, int t1, int N, int n) {
#pragma unroll
  for (int u = 0; u < kGaeChunk; ++u) {  // independent, coalesced loads: the recurrence never waits on DRAM
    const int t = t1 - 1 - u;
    if (t >= 0) {
      const size_t k = (size_t)t * N + n;
      c.r[u] = __ldcs(rew + k);
      c.v[u] = __ldcs(val + k);
      c.d[u] = __ldcs(done + k);
    }
  }
}

__device__ __forceinline__ void gae_scan(const GaeChunk& c, int t1, int N, int n, float gamma, float gl, float& next_v, float& last,
                                         float* __restrict__ adv, float* __restrict__ ret) {
#pragma unroll
  for (int u = 0; u < kGaeChunk; ++u) {
    const int t = t1 - 1 - u;
    if (t >= 0) {
      const size_t k = (size_t)t * N + n;
      const float nnt = __fsub_rn(1.0f, c.d[u] ? 1.0f : 0.0f);
      // delta = r + GAMMA * next_values * next_non_terminal - v
      const float delta = __fsub_rn(__fadd_rn(c.r[u], __fmul_rn(__fmul_rn(gamma, next_v), nnt)), c.v[u]);
      // last = delta + GAMMA * GAE_LAMBDA * next_non_terminal * last   (GAMMA * GAE_LAMBDA is one Python float)
      last = __fadd_rn(delta, __fmul_rn(__fmul_rn(gl, nnt), last));
      __stcs(adv + k, last);
      __stcs(ret + k, __fadd_rn(last, c.v[u]));
      next_v = c.v[u];
    }
  }
}

// Two chunks are in flight: the loads of the next 16 time steps are issued before the recurrence of the current 16 runs, so
// the (few: one thread per env) warps always have DRAM requests outstanding.
__global__ void __launch_bounds__(kGaeThreads) mdx_gae_kernel(const float* __restrict__ rew, const float* __restrict__ val,
                                                      const uint8_t* __restrict__ done, const float* __restrict__ last_value,
                                                      int T, int N, float gamma, float gl, float* __restrict__ adv,
                                                      float* __restrict__ ret) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float next_v = last_value[n];
  float last = 0.0f;
#if MDX_GAE_DB
  GaeChunk a, b;
  gae_load(a, rew, val, done, T, N, n);
  for (int t1 = T; t1 > 0; t1 -= 2 * kGaeChunk) {
    gae_load(b, rew, val, done, t1 - kGaeChunk, N, n);
    gae_scan(a, t1, N, n, gamma, gl, next_v, last, adv, ret);
    gae_load(a, rew, val, done, t1 - 2 * kGaeChunk, N, n);
    gae_scan(b, t1 - kGaeChunk, N, n, gamma, gl, next_v, last, adv, ret);
  }
#else
  GaeChunk a;
  for (int t1 = T; t1 > 0; t1 -= kGaeChunk) {
    gae_load(a, rew, val, done, t1, N, n);
    gae_scan(a, t1, N, n, gamma, gl, next_v, last, adv, ret);
  }
#endif
}

// ---- TMA variant -----------------------------------------------------------------------------------------------------
// One thread per env leaves only ~14 warps per SM, and what they can keep in flight from registers (two chunks) is ~9 MB
// chip-wide -- short of what HBM3e needs.  Here the loads do not occupy registers at all: a CTA owns 128 adjacent columns and
// streams them through a ring of shared-memory stages filled by bulk async copies (cp.async.bulk, one 512-byte row segment
// per array and time step, completion counted on an mbarrier per stage); kGaeTmaStages - 1 stages are always in flight while
// the threads run the recurrence on the landed one.  Needs N % 16 == 0 and 16-byte aligned arrays (row segments of the uint8
// dones must be 16-byte aligned); anything else takes the register kernel above.
// Measured on B200, T = 500, N = 65 536 (rows x stages): 20x2 96.5 us (88.1 % of the HBM copy peak), 16x3 97.8, 10x4 99.5, 8x6 103.0,
// 5x8 108.6; 10x5 (57 KB: only 3 CTAs/SM, the 512 CTAs no longer fit one wave) 175 us.  Register kernel: 116.7 us (72.9 %).
#ifndef MDX_GAE_TMA_ROWS
#define MDX_GAE_TMA_ROWS 20
#endif
#ifndef MDX_GAE_TMA_STAGES
#define MDX_GAE_TMA_STAGES 2
#endif
constexpr int kGaeTmaRows = MDX_GAE_TMA_ROWS, kGaeTmaStages = MDX_GAE_TMA_STAGES, kGaeTmaCols = 128;
constexpr int kGaeStageBytes = kGaeTmaRows * kGaeTmaCols * 9;  // rewards f32 + values f32 + dones u8
constexpr int kGaeTmaSmem = 128 + kGaeTmaStages * kGaeStageBytes;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* b, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(b)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(b))
               : "memory");
}

__global__ void __launch_bounds__(kGaeTmaCols) mdx_gae_tma_kernel(const float* __restrict__ rew, const float* __restrict__ val,
                                                                 const uint8_t* __restrict__ done, const float* __restrict__ last_value,
                                                                 int T, int N, float gamma, float gl, float* __restrict__ adv,
                                                                 float* __restrict__ ret) {
  extern __shared__ __align__(128) unsigned char gae_smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(gae_smem);  // [kGaeTmaStages]
  unsigned char* ring = gae_smem + 128;
  const int tid = threadIdx.x;
  const int c0 = blockIdx.x * kGaeTmaCols;
  const int cols = min(kGaeTmaCols, N - c0);  // a multiple of 16
  const int n = c0 + tid;
  const bool act = tid < cols;
  const int nchunks = (T + kGaeTmaRows - 1) / kGaeTmaRows;
  if (tid == 0) {
    for (int s = 0; s < kGaeTmaStages; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  // chunk k = time steps [T - k*rows - rows, T - k*rows), newest first; warp 0 issues it, one row per lane
  auto issue = [&](int k) {
    const int s = k % kGaeTmaStages;
    const int t1 = T - k * kGaeTmaRows;
    const int rows = min(kGaeTmaRows, t1);
    float* s_rew = reinterpret_cast<float*>(ring + (size_t)s * kGaeStageBytes);
    float* s_val = s_rew + kGaeTmaRows * kGaeTmaCols;
    unsigned char* s_done = reinterpret_cast<unsigned char*>(s_val + kGaeTmaRows * kGaeTmaCols);
    if (tid == 0) mbar_expect_tx(&full[s], (unsigned)(rows * cols * 9));
    __syncwarp();
    for (int u = tid; u < rows; u += 32) {
      const size_t g = (size_t)(t1 - 1 - u) * N + c0;
      bulk_g2s(s_rew + u * kGaeTmaCols, rew + g, (unsigned)cols * 4u, &full[s]);
      bulk_g2s(s_val + u * kGaeTmaCols, val + g, (unsigned)cols * 4u, &full[s]);
      bulk_g2s(s_done + u * kGaeTmaCols, done + g, (unsigned)cols, &full[s]);
    }
  };
  if (tid < 32)
    for (int k = 0; k < min(kGaeTmaStages, nchunks); ++k) issue(k);
  float next_v = act ? last_value[n] : 0.0f;
  float last = 0.0f;
  for (int k = 0; k < nchunks; ++k) {
    const int s = k % kGaeTmaStages;
    const unsigned parity = (unsigned)(k / kGaeTmaStages) & 1u;
    unsigned spins = 0;
    while (!mbar_try_wait(&full[s], parity))
      if (++spins > (1u << 24)) __trap();  // a lost copy must not hang the GPU
    const int t1 = T - k * kGaeTmaRows;
    const int rows = min(kGaeTmaRows, t1);
    const float* s_rew = reinterpret_cast<const float*>(ring + (size_t)s * kGaeStageBytes);
    const float* s_val = s_rew + kGaeTmaRows * kGaeTmaCols;
    const unsigned char* s_done = reinterpret_cast<const unsigned char*>(s_val + kGaeTmaRows * kGaeTmaCols);
    if (act) {
#pragma unroll
      for (int u = 0; u < kGaeTmaRows; ++u) {
        if (u < rows) {
          const size_t g = (size_t)(t1 - 1 - u) * N + n;
          const float r = s_rew[u * kGaeTmaCols + tid], v = s_val[u * kGaeTmaCols + tid];
          const float nnt = __fsub_rn(1.0f, s_done[u * kGaeTmaCols + tid] ? 1.0f : 0.0f);
          const float delta = __fsub_rn(__fadd_rn(r, __fmul_rn(__fmul_rn(gamma, next_v), nnt)), v);
          last = __fadd_rn(delta, __fmul_rn(__fmul_rn(gl, nnt), last));
          __stcs(adv + g, last);
          __stcs(ret + g, __fadd_rn(last, v));
          next_v = v;
        }
      }
    }
    __syncthreads();  // every thread has consumed stage s: it can be refilled
    if (tid < 32 && k + kGaeTmaStages < nchunks) issue(k + kGaeTmaStages);
  }
}

}  // namespace mdx

// ------------------------------------------------------------------------------------------------
// checkpoint blob: header, then every array in a fixed order
// ------------------------------------------------------------------------------------------------
namespace {

struct CkptHeader {
  uint64_t magic;  // "MDXCKPT2"
  int32_t n, R, P, T, Pp8, Pp16, H_out, W_out, n_maps, has_greedy, gfree_words, reserved;
  uint64_t config_hash;  // maps, env -> map assignment, rewards, shaping constants, feature sizes
  int64_t bytes;
};
constexpr uint64_t kMagic = 0x3254504b4358444dull;

struct Seg {
  void* dev;
  size_t bytes;
};

// FNV-1a over everything that decides how the stored integers are to be read (cell ids depend on W_out and the maps, rewards on the
// constants): a blob of an instance configured differently must not load.
uint64_t config_hash(const mdx_env* e) {
  uint64_t h = 1469598103934665603ull;
  auto mix = [&](const void* p, size_t nbytes) {
    const unsigned char* c = static_cast<const unsigned char*>(p);
    for (size_t i = 0; i < nbytes; ++i) h = (h ^ c[i]) * 1099511628211ull;
  };
  for (const MapHost& m : e->maps) {
    mix(&m.rows, sizeof(m.rows));
    mix(&m.cols, sizeof(m.cols));
    mix(m.cells.data(), m.cells.size());
  }
  mix(e->env_map.data(), sizeof(int32_t) * e->env_map.size());
  const KParams& b = e->base;
  mix(&b.move_cost, sizeof(double));
  mix(&b.delivery_reward, sizeof(double));
  mix(&b.delay_reward, sizeof(double));
  mix(b.shaping, sizeof(b.shaping));
  const int32_t sizes[6] = {b.K_o, b.K_p, b.MR, b.MP, b.shaping_bfs, e->stale ? 1 : 0};
  mix(sizes, sizeof(sizes));
  return h;
}

std::vector<Seg> device_segments(mdx_env* e) {
  const size_t n = (size_t)e->n;
  const KParams& b = e->base;
  std::vector<Seg> s = {
      {e->d_t, sizeof(int32_t) * n},
      {e->d_ep, sizeof(int32_t) * n},
      {e->d_mtpos, sizeof(int32_t) * n},  // (index 2: validated on load)
      {e->d_total, sizeof(double) * n},
      {e->d_robots, sizeof(uint32_t) * n * e->R},
      {e->d_status, n * b.Pp16},
      {e->d_cur, sizeof(uint16_t) * n * 4 * b.Pp8},
      {e->d_counters, sizeof(uint32_t) * n * 8},
      {e->d_mt, sizeof(uint32_t) * n * 624},
  };
  if (e->stale) {  // the trainer's package dicts are part of the state (MDX_CFG_STALE_TRACKER)
    s.push_back({e->d_trk, sizeof(uint16_t) * n * 5 * b.Pp8});
    s.push_back({e->d_trk_status, n * b.Pp16});
    s.push_back({e->d_trk_n, sizeof(int32_t) * n});
  }
  if (e->d_greedy) {
    s.push_back({e->d_greedy, sizeof(uint32_t) * n * 32});
    s.push_back({e->d_gfree, sizeof(uint32_t) * n * (size_t)e->gfree_words});
  }
  return s;
}

int ensure_greedy(mdx_env* e) {
  if (e->d_greedy) return MDX_OK;
  DeviceGuard g(e->device);
  e->gfree_words = (2 * std::max(1, e->P) + 31) / 32;
  const size_t n = (size_t)e->n;
  CU(cudaMalloc(&e->d_greedy, sizeof(uint32_t) * n * 32));
  CU(cudaMalloc(&e->d_gfree, sizeof(uint32_t) * n * (size_t)e->gfree_words));
  CU(cudaMemset(e->d_greedy, 0, sizeof(uint32_t) * n * 32));
  CU(cudaMemset(e->d_gfree, 0xff, sizeof(uint32_t) * n * (size_t)e->gfree_words));
  return MDX_OK;
}

}  // namespace

extern "C" {

int32_t mdx_greedy_actions(mdx_env* e, const int32_t* d_env_ids, int32_t n, uint8_t* d_actions, void* stream) {
  if (!e || !d_actions) return fail(MDX_E_INVALID, "null argument");
  if (d_env_ids == nullptr && n != e->n) return fail(MDX_E_INVALID, "n must equal n_envs when env ids are not given (got %d, have %d)", n, e->n);
  if (n < 0 || n > e->n) return fail(MDX_E_INVALID, "bad env count %d", n);
  if (n == 0) return MDX_OK;
  int rc = ensure_bfs(e);
  if (rc != MDX_OK) return rc;
  if ((rc = ensure_greedy(e)) != MDX_OK) return rc;
  DeviceGuard g(e->device);
  KParams kp = e->base;
  kp.env_ids = d_env_ids;
  kp.n = n;
  GreedyArgs a{e->d_greedy, e->d_gfree, e->gfree_words, d_actions};
  const size_t smem = sizeof(uint32_t) * 4 * (size_t)e->gfree_words;
  mdx_greedy_kernel<<<(unsigned)((n + 3) / 4), 128, smem, static_cast<cudaStream_t>(stream)>>>(kp, a);
  CU(cudaGetLastError());
  return MDX_OK;
}

int32_t mdx_featurize_alt(mdx_env* e, const int32_t* d_env_ids, int32_t n, float* d_obs6, float* d_gstate7, void* stream) {
  if (!e) return fail(MDX_E_INVALID, "null env");
  if (d_env_ids == nullptr && n != e->n) return fail(MDX_E_INVALID, "n must equal n_envs when env ids are not given (got %d, have %d)", n, e->n);
  if (n < 0 || n > e->n) return fail(MDX_E_INVALID, "bad env count %d", n);
  if (n == 0) return MDX_OK;
  if ((reinterpret_cast<uintptr_t>(d_obs6) | reinterpret_cast<uintptr_t>(d_gstate7)) & 15u)
    return fail(MDX_E_INVALID, "feature tensors must be 16-byte aligned");
  DeviceGuard g(e->device);
  KParams kp = e->base;
  kp.env_ids = d_env_ids;
  kp.n = n;
  const int HW = e->base.HW;
  AltArgs a{d_obs6, d_gstate7, (int32_t)(sizeof(float) * ((HW + 3) & ~3) + sizeof(uint32_t) * 6 * e->base.W32 + 15) / 16 * 16};
  const cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned)((n + 3) / 4);
  if ((HW & 3) == 0 && HW <= 4 * 32 * 4 && !getenv("MDX_ALT_GENERIC")) {
    // plane-sliced variants: a plane is at most NS slices of 32 float4
    const size_t smem = (size_t)kLutBytes + (size_t)4 * a.smem_per_warp;
    const int ns = (HW / 4 + 31) / 32;
    auto go = [&](auto kern) -> int {
      if (smem > 48 * 1024) CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kern<<<grid, 128, smem, st>>>(kp, a);
      return MDX_OK;
    };
    int rc = ns == 1 ? go(mdx_alt_featurize_sliced<1>) : ns == 2 ? go(mdx_alt_featurize_sliced<2>)
           : ns == 3 ? go(mdx_alt_featurize_sliced<3>) : go(mdx_alt_featurize_sliced<4>);
    if (rc != MDX_OK) return rc;
  } else {
    const size_t smem = (size_t)4 * a.smem_per_warp;
    if (smem > 48 * 1024) CU(cudaFuncSetAttribute(mdx_alt_featurize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mdx_alt_featurize_kernel<<<grid, 128, smem, st>>>(kp, a);
  }
  CU(cudaGetLastError());
  return MDX_OK;
}

int32_t mdx_bfs_table(mdx_env* e, int32_t map_id, uint16_t* h_out) {
  if (!e || !h_out) return fail(MDX_E_INVALID, "null argument");
  if (map_id < 0 || map_id >= (int)e->maps.size()) return fail(MDX_E_INVALID, "map id %d out of range", map_id);
  int rc = ensure_bfs(e);
  if (rc != MDX_OK) return rc;
  DeviceGuard g(e->device);
  const size_t per = (size_t)e->base.HW * e->base.HW;
  CU(cudaMemcpy(h_out, e->d_bfs + per * (size_t)map_id, sizeof(uint16_t) * per, cudaMemcpyDeviceToHost));
  return MDX_OK;
}

int32_t mdx_gae(const float* d_rewards, const float* d_values, const uint8_t* d_dones, const float* d_last_value, int32_t T,
                int32_t N, double gamma, double gae_lambda, float* d_advantages, float* d_returns, void* stream) {
  if (!d_rewards || !d_values || !d_dones || !d_last_value || !d_advantages || !d_returns) return fail(MDX_E_INVALID, "null argument");
  if (T < 0 || N < 0) return fail(MDX_E_INVALID, "negative rollout shape");
  if (T == 0 || N == 0) return MDX_OK;
  // a Python float scalar times a float32 tensor is evaluated in float32 with the scalar rounded once
  const float g32 = (float)gamma, gl32 = (float)(gamma * gae_lambda);
  const cudaStream_t st = static_cast<cudaStream_t>(stream);
  const auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  if (N % 16 == 0 && al16(d_rewards) && al16(d_values) && al16(d_dones) && !getenv("MDX_GAE_NO_TMA")) {
    if (mdx::kGaeTmaSmem > 48 * 1024)
      CU(cudaFuncSetAttribute(mdx::mdx_gae_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, mdx::kGaeTmaSmem));
    mdx::mdx_gae_tma_kernel<<<(unsigned)((N + mdx::kGaeTmaCols - 1) / mdx::kGaeTmaCols), mdx::kGaeTmaCols, mdx::kGaeTmaSmem, st>>>(
        d_rewards, d_values, d_dones, d_last_value, T, N, g32, gl32, d_advantages, d_returns);
  } else {
    mdx_gae_kernel<<<(unsigned)((N + mdx::kGaeThreads - 1) / mdx::kGaeThreads), mdx::kGaeThreads, 0, st>>>(d_rewards, d_values, d_dones,
                                                                                                  d_last_value, T, N, g32, gl32,
                                                                                                  d_advantages, d_returns);
  }
  CU(cudaGetLastError());
  return MDX_OK;
}

int64_t mdx_checkpoint_bytes(const mdx_env* e) {
  if (!e) return 0;
  size_t total = sizeof(CkptHeader);
  for (const Seg& s : device_segments(const_cast<mdx_env*>(e))) total += s.bytes;
  return (int64_t)total;
}

int32_t mdx_checkpoint_save(mdx_env* e, void* h_buf, int64_t bytes, void* stream) {
  if (!e || !h_buf) return fail(MDX_E_INVALID, "null argument");
  const int64_t need = mdx_checkpoint_bytes(e);
  if (bytes < need) return fail(MDX_E_INVALID, "checkpoint buffer too small (%lld < %lld bytes)", (long long)bytes, (long long)need);
  DeviceGuard g(e->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CU(cudaStreamSynchronize(st));
  unsigned char* p = static_cast<unsigned char*>(h_buf);
  CkptHeader h{};
  h.magic = kMagic;
  h.n = e->n, h.R = e->R, h.P = e->P, h.T = e->T, h.Pp8 = e->base.Pp8, h.Pp16 = e->base.Pp16;
  h.H_out = e->H_out, h.W_out = e->W_out, h.n_maps = (int32_t)e->maps.size();
  h.has_greedy = e->d_greedy ? 1 : 0;
  h.gfree_words = e->gfree_words;
  h.config_hash = config_hash(e);
  h.bytes = need;
  std::memcpy(p, &h, sizeof(h));
  p += sizeof(h);
  for (const Seg& s : device_segments(e)) {
    CU(cudaMemcpy(p, s.dev, s.bytes, cudaMemcpyDeviceToHost));
    p += s.bytes;
  }
  return MDX_OK;
}

int32_t mdx_checkpoint_load(mdx_env* e, const void* h_buf, int64_t bytes, void* stream) {
  if (!e || !h_buf) return fail(MDX_E_INVALID, "null argument");
  if (bytes < (int64_t)sizeof(CkptHeader)) return fail(MDX_E_INVALID, "not an mdx checkpoint");
  CkptHeader h;
  std::memcpy(&h, h_buf, sizeof(h));
  if (h.magic != kMagic) return fail(MDX_E_INVALID, "not an mdx checkpoint (or one written by another version of the library)");
  if (h.n != e->n || h.R != e->R || h.P != e->P || h.T != e->T || h.Pp8 != e->base.Pp8 || h.Pp16 != e->base.Pp16 ||
      h.H_out != e->H_out || h.W_out != e->W_out || h.n_maps != (int32_t)e->maps.size())
    return fail(MDX_E_INVALID, "checkpoint shape (n=%d R=%d P=%d T=%d plane %dx%d, %d maps) does not match this instance (n=%d R=%d P=%d T=%d "
                "plane %dx%d, %d maps)", h.n, h.R, h.P, h.T, h.H_out, h.W_out, h.n_maps, e->n, e->R, e->P, e->T, e->H_out, e->W_out,
                (int)e->maps.size());
  if (h.config_hash != config_hash(e))
    return fail(MDX_E_INVALID, "checkpoint was written by an instance with other maps, env -> map assignment, rewards, shaping or feature sizes");
  // ---- validate the whole blob before the first byte reaches the device (a refused blob leaves the instance untouched)
  const size_t n = (size_t)e->n;
  const size_t gw = (size_t)(h.has_greedy ? h.gfree_words : 0);
  if (h.has_greedy && h.gfree_words != (2 * std::max(1, e->P) + 31) / 32) return fail(MDX_E_INVALID, "checkpoint greedy state does not match");
  std::vector<Seg> base = device_segments(e);
  if (e->d_greedy) base.resize(base.size() - 2);
  size_t need = sizeof(CkptHeader);
  for (const Seg& s : base) need += s.bytes;
  if (h.has_greedy) need += sizeof(uint32_t) * n * 32 + sizeof(uint32_t) * n * gw;
  if ((int64_t)need != h.bytes || bytes < h.bytes) return fail(MDX_E_INVALID, "truncated or inconsistent checkpoint");
  {
    const unsigned char* q = static_cast<const unsigned char*>(h_buf) + sizeof(h) + base[0].bytes + base[1].bytes;
    for (size_t i = 0; i < n; ++i) {
      int32_t pos;
      std::memcpy(&pos, q + sizeof(int32_t) * i, sizeof(pos));
      if (pos < 0 || pos > 624) return fail(MDX_E_INVALID, "corrupt generator state in checkpoint (env %zu)", i);
    }
  }
  if (h.has_greedy) {
    int rc = ensure_greedy(e);
    if (rc != MDX_OK) return rc;
  }
  DeviceGuard g(e->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CU(cudaStreamSynchronize(st));
  const unsigned char* p = static_cast<const unsigned char*>(h_buf) + sizeof(h);
  std::vector<Seg> segs = device_segments(e);
  if (!h.has_greedy && e->d_greedy) segs.resize(segs.size() - 2);
  for (const Seg& s : segs) {
    CU(cudaMemcpy(s.dev, p, s.bytes, cudaMemcpyHostToDevice));
    p += s.bytes;
  }
  if (!h.has_greedy && e->d_greedy) {  // the blob carries no agent bookkeeping: back to a freshly created agent
    CU(cudaMemset(e->d_greedy, 0, sizeof(uint32_t) * n * 32));
    CU(cudaMemset(e->d_gfree, 0xff, sizeof(uint32_t) * n * (size_t)e->gfree_words));
  }
  return MDX_OK;
}

}  // extern "C"
