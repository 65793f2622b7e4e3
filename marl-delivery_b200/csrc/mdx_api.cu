// C-ABI glue of libmdx: configuration, device state, the host reset sampler's ring management and
// kernel launches.  See include/mdx.h for the contract of every entry point.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "../../include/mdx.h"
#include "mdx_host.h"
#include "mdx_small_kernels.cuh"

using namespace mdx;

namespace mdx {
thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
}  // namespace mdx

struct mdx_sampler {
  EpisodeSampler s;
};

namespace {

int pick_wpb(const mdx_env* e, int smem_per_warp) {
  if (e->wpb_override > 0) return std::min(4, e->wpb_override);  // kernels are compiled for <= 128 threads
  int wpb = 4;
  while (wpb > 1 && wpb * smem_per_warp > 200 * 1024) wpb >>= 1;
  return wpb;
}

// Draw block `sampled[env]` for env and pack it into staging record `rec` (output-plane cells).
bool sample_into(mdx_env* e, int env, uint16_t* rec, std::vector<int32_t>& rc, std::vector<int32_t>& pk,
                 EpisodeSampler& scratch) {
  const MapHost& m = e->maps[(size_t)e->env_map[(size_t)env]];
  // `scratch` only lends its work vectors; the generator is the env's own stream
  if (!scratch.draw(e->rng[(size_t)env], m.free_cells, m.rows, e->R, e->P, e->T, rc.data(), pk.data())) return false;
  const int Pp8 = e->base.Pp8, P = e->P, Wo = e->W_out;
  std::memset(rec, 0, sizeof(uint16_t) * (size_t)e->rec_u16);
  auto oc = [&](int cell) { return (uint16_t)((cell / m.cols) * Wo + (cell % m.cols)); };
  for (int j = 0; j < P; ++j) {
    rec[0 * Pp8 + j] = oc(pk[(size_t)(0 * P + j)]);
    rec[1 * Pp8 + j] = oc(pk[(size_t)(1 * P + j)]);
    rec[2 * Pp8 + j] = (uint16_t)pk[(size_t)(2 * P + j)];
    rec[3 * Pp8 + j] = (uint16_t)pk[(size_t)(3 * P + j)];
  }
  for (int j = P; j < Pp8; ++j) rec[2 * Pp8 + j] = 0xffff;  // padding never spawns
  for (int i = 0; i < e->R; ++i) rec[4 * Pp8 + i] = oc(rc[(size_t)i]);
  return true;
}

// Sample + upload the blocks listed in (ids[k], count[k]) pairs, at most one block per env per wave.
// `want[env]` = number of blocks the env must have after the call.
int top_up(mdx_env* e, const std::vector<int32_t>& envs, const std::vector<int32_t>& want, cudaStream_t st) {
  std::vector<int32_t> todo(envs);
  while (true) {
    std::vector<int32_t> wave;
    for (int32_t env : todo)
      if (e->sampled[(size_t)env] < want[(size_t)env]) wave.push_back(env);
    if (wave.empty()) break;
    const int m = (int)wave.size();
    std::vector<char> bad((size_t)e->threads + 1, 0);
    parallel_for(m, e->threads, [&](int tid, int lo, int hi) {
      EpisodeSampler scratch;
      std::vector<int32_t> rc((size_t)e->R + 1), pk((size_t)4 * e->P + 1);
      for (int k = lo; k < hi; ++k) {
        const int env = wave[(size_t)k];
        if (!sample_into(e, env, e->h_stage + (size_t)k * e->rec_u16, rc, pk, scratch)) bad[(size_t)tid] = 1;
        e->h_ids[k] = env;
        e->h_ids[e->n + k] = e->sampled[(size_t)env] % e->E;
        e->sampled[(size_t)env] += 1;
      }
    });
    for (char b : bad)
      if (b)
        return fail(MDX_E_INVALID,
                    "reset sampling failed: not enough free cells for the robots/packages, or max_time_steps < 2 "
                    "with more than min(n_robots, 20) + 1 packages (the reference raises ValueError here)");
    CU(cudaMemcpyAsync(e->d_stage, e->h_stage, sizeof(uint16_t) * (size_t)m * e->rec_u16, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(e->d_ids, e->h_ids, sizeof(int32_t) * (size_t)m, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(e->d_ids + e->n, e->h_ids + e->n, sizeof(int32_t) * (size_t)m, cudaMemcpyHostToDevice, st));
    const size_t total = (size_t)m * (e->rec_u16 / 8);
    mdx_upload_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(e->d_sched, e->d_sched_robots, e->d_stage, e->d_ids,
                                                                       e->d_ids + e->n, m, e->n, e->base.Pp8);
    CU(cudaGetLastError());
    // the pinned staging buffers are reused by the next wave
    CU(cudaStreamSynchronize(st));
    todo.swap(wave);
  }
  // publish the new availability
  std::memcpy(e->h_i32, e->sampled.data(), sizeof(int32_t) * (size_t)e->n);
  CU(cudaMemcpyAsync(e->d_ep_avail, e->h_i32, sizeof(int32_t) * (size_t)e->n, cudaMemcpyHostToDevice, st));
  CU(cudaStreamSynchronize(st));
  return MDX_OK;
}

int refresh_ep(mdx_env* e, cudaStream_t st) {
  CU(cudaMemcpyAsync(e->h_i32, e->d_ep, sizeof(int32_t) * (size_t)e->n, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  std::memcpy(e->ep_host.data(), e->h_i32, sizeof(int32_t) * (size_t)e->n);
  return MDX_OK;
}

int check_ptr16(const void* p, const char* name) {
  if (p != nullptr && (reinterpret_cast<uintptr_t>(p) & 15u) != 0)
    return fail(MDX_E_INVALID, "%s must be 16-byte aligned", name);
  return MDX_OK;
}

}  // namespace

extern "C" {

const char* mdx_last_error(void) { return g_err.c_str(); }
int32_t mdx_version(void) { return MDX_VERSION; }

int32_t mdx_destroy(mdx_env* e) {
  if (!e) return MDX_OK;
  DeviceGuard g(e->device);
  cudaFree(e->d_blob);
  cudaFree(e->d_t);
  cudaFree(e->d_ep);
  cudaFree(e->d_ep_avail);
  cudaFree(e->d_total);
  cudaFree(e->d_robots);
  cudaFree(e->d_status);
  cudaFree(e->d_sched);
  cudaFree(e->d_cur);
  cudaFree(e->d_sched_robots);
  cudaFree(e->d_counters);
  cudaFree(e->d_env_map);
  cudaFree(e->d_stage);
  cudaFree(e->d_ids);
  cudaFree(e->d_bfs);
  cudaFree(e->d_greedy);
  cudaFree(e->d_gfree);
  cudaFreeHost(e->h_stage);
  cudaFreeHost(e->h_ids);
  cudaFreeHost(e->h_i32);
  delete e;
  return MDX_OK;
}

int32_t mdx_create(const mdx_config* cfg, mdx_env** out) {
  if (!cfg || !out) return fail(MDX_E_INVALID, "null argument");
  *out = nullptr;
  if (cfg->n_envs <= 0) return fail(MDX_E_INVALID, "n_envs must be positive");
  if (cfg->n_robots <= 0 || cfg->n_robots > MDX_MAX_ROBOTS)
    return fail(MDX_E_INVALID, "n_robots must be in [1, %d] (one warp per environment, robots are lanes)", MDX_MAX_ROBOTS);
  if (cfg->n_packages < 0 || cfg->n_packages > 4096) return fail(MDX_E_INVALID, "n_packages must be in [0, 4096]");
  if (cfg->max_time_steps < 0 || cfg->max_time_steps > 60000) return fail(MDX_E_INVALID, "max_time_steps must be in [0, 60000]");
  if (cfg->n_maps <= 0 || !cfg->map_rows || !cfg->map_cols || !cfg->map_cells) return fail(MDX_E_INVALID, "at least one map is required");
  if (cfg->k_others < 0 || cfg->k_packages < 0 || cfg->max_robots_in_state < 0 || cfg->max_packages_in_state < 0)
    return fail(MDX_E_INVALID, "feature sizes must be non-negative");
  if (cfg->ring_episodes < 2 || cfg->ring_episodes > 64) return fail(MDX_E_INVALID, "ring_episodes must be in [2, 64]");

  std::unique_ptr<mdx_env> e(new mdx_env());
  e->cfg = *cfg;
  e->device = cfg->device;
  e->n = cfg->n_envs;
  e->R = cfg->n_robots;
  e->P = cfg->n_packages;
  e->T = cfg->max_time_steps;
  e->E = cfg->ring_episodes;
  e->threads = cfg->sampler_threads > 0 ? cfg->sampler_threads : (int)std::max(1u, std::thread::hardware_concurrency());
  if (const char* s = getenv("MDX_WPB")) e->wpb_override = atoi(s);
  if (const char* s = getenv("MDX_GENERIC")) e->force_generic = atoi(s);
  if (const char* s = getenv("MDX_LPE")) e->lpe_override = atoi(s);
  if (const char* s = getenv("MDX_NOFIX")) e->no_fixed = atoi(s);
  if (const char* s = getenv("MDX_NOFAT")) e->no_fat = atoi(s);
  if (const char* s = getenv("MDX_CARVEOUT")) e->carveout = atoi(s);

  // ---- maps
  int Ho = 0, Wo = 0;
  e->maps.resize((size_t)cfg->n_maps);
  for (int m = 0; m < cfg->n_maps; ++m) {
    MapHost& mh = e->maps[(size_t)m];
    mh.rows = cfg->map_rows[m];
    mh.cols = cfg->map_cols[m];
    if (mh.rows < 2 || mh.cols < 2 || mh.rows > 128 || mh.cols > 128) return fail(MDX_E_INVALID, "map %d: rows/cols must be in [2, 128]", m);
    mh.cells.assign(cfg->map_cells[m], cfg->map_cells[m] + (size_t)mh.rows * mh.cols);
    for (int i = 0; i < mh.rows * mh.cols; ++i)
      if (mh.cells[(size_t)i] == 0) mh.free_cells.push_back(i);
    if ((int)mh.free_cells.size() < e->R) return fail(MDX_E_INVALID, "map %d has fewer free cells than robots", m);
    Ho = std::max(Ho, mh.rows);
    Wo = std::max(Wo, mh.cols);
  }
  e->H_out = Ho;
  e->W_out = Wo;
  e->env_map.assign((size_t)e->n, 0);
  if (cfg->env_map)
    for (int i = 0; i < e->n; ++i) {
      if (cfg->env_map[i] < 0 || cfg->env_map[i] >= cfg->n_maps) return fail(MDX_E_INVALID, "env_map[%d] out of range", i);
      e->env_map[(size_t)i] = cfg->env_map[i];
    }

  // ---- shape-derived parameters
  KParams& b = e->base;
  b.n_envs = e->n;
  b.R = e->R;
  b.P = e->P;
  b.T = e->T;
  b.Pp8 = std::max(8, round_up(e->P, 8));
  b.Pp16 = std::max(16, round_up(e->P, 16));
  b.H_out = Ho;
  b.W_out = Wo;
  b.HW = Ho * Wo;
  b.W32 = (b.HW + 31) / 32;
  b.K_o = cfg->k_others;
  b.K_p = cfg->k_packages;
  b.MR = cfg->max_robots_in_state;
  b.MP = cfg->max_packages_in_state;
  b.D = 6 + 5 * b.K_o + 5 * b.K_p + 1;
  b.Dg = 6 * b.MR + 7 * b.MP + 1;
  b.E = e->E;
  b.mul_plane = magic((uint32_t)b.HW);
  b.mul_wout = magic((uint32_t)Wo);
  b.mul_r = magic((uint32_t)e->R);
  b.pair_g = 32 / e->R;
  b.move_cost = cfg->move_cost;
  b.delivery_reward = cfg->delivery_reward;
  b.delay_reward = cfg->delay_reward;
  std::memcpy(b.shaping, &cfg->shaping, sizeof(float) * 9);
  e->rec_u16 = 4 * b.Pp8 + 32;

  // ---- LUTs into one host blob
  const int nqr = 2 * Ho - 1, nqc = 2 * Wo - 1;
  struct MapOff {
    size_t grid, qr, qc, rank;
  };
  std::vector<MapOff> offs((size_t)cfg->n_maps);
  std::vector<unsigned char> blob;
  auto reserve = [&](size_t bytes) {
    size_t at = (blob.size() + 255) / 256 * 256;
    blob.resize(at + bytes, 0);
    return at;
  };
  int nranks_max = 1;
  std::vector<int> nranks((size_t)cfg->n_maps);
  for (int m = 0; m < cfg->n_maps; ++m) {
    const MapHost& mh = e->maps[(size_t)m];
    offs[(size_t)m].grid = reserve(sizeof(uint32_t) * (size_t)b.W32);
    offs[(size_t)m].qr = reserve(sizeof(float) * (size_t)nqr);
    offs[(size_t)m].qc = reserve(sizeof(float) * (size_t)nqc);
    offs[(size_t)m].rank = reserve(sizeof(uint16_t) * (size_t)nqr * nqc);
    uint32_t* gb = reinterpret_cast<uint32_t*>(blob.data() + offs[(size_t)m].grid);
    for (int r = 0; r < mh.rows; ++r)
      for (int c = 0; c < mh.cols; ++c)
        if (mh.cells[(size_t)r * mh.cols + c] != 0) {
          const int oc = r * Wo + c;
          gb[oc >> 5] |= 1u << (oc & 31);
        }
    build_quotient_lut(mh.rows, Ho, reinterpret_cast<float*>(blob.data() + offs[(size_t)m].qr));
    build_quotient_lut(mh.cols, Wo, reinterpret_cast<float*>(blob.data() + offs[(size_t)m].qc));
    nranks[(size_t)m] = build_rank_lut(mh.rows, mh.cols, Ho, Wo, reinterpret_cast<uint16_t*>(blob.data() + offs[(size_t)m].rank));
    nranks_max = std::max(nranks_max, nranks[(size_t)m]);
  }
  int max_rows = Ho;
  b.tq_len = std::min(65536, e->T + 10 + 3 * max_rows + 64);
  const size_t off_tq = reserve(sizeof(float) * (size_t)b.tq_len);
  {
    float* tq = reinterpret_cast<float*>(blob.data() + off_tq);
    for (int k = 0; k < b.tq_len; ++k) tq[k] = e->T > 0 ? (float)((double)k / (double)e->T) : 0.0f;
  }
  const size_t off_cq = reserve(sizeof(float) * 32);
  {
    float* cq = reinterpret_cast<float*>(blob.data() + off_cq);
    for (int i = 0; i < 32; ++i) cq[i] = b.MR > 1 ? (float)((double)i / (double)(b.MR - 1)) : 0.0f;
  }
  const size_t off_maps = reserve(sizeof(MapDev) * (size_t)cfg->n_maps);

  // selection key: (remaining * nranks + rank) << wbits | waiting-list index must fit 32 bits
  int wbits = 1;
  while ((1 << wbits) < std::max(2, e->P)) ++wbits;
  b.key_wbits = wbits;
  b.nranks_max = nranks_max;
  const long long max_deadline = (long long)e->T + 10 + 3LL * max_rows;  // env.py:112-117 upper bound
  if ((max_deadline + 1) * nranks_max >= (1LL << (32 - wbits)) - 1 || max_deadline > 65535)
    return fail(MDX_E_INVALID, "configuration too large for the 32-bit package selection key (T=%d, P=%d, map %dx%d)", e->T, e->P, Ho, Wo);

  // ---- shared-memory carve-up per warp
  // `rpf` = vec rows staged per flush.  Rows longer than 2 KB (the K = 100 defaults: 4 KB per row) are flushed one at a time (a
  // smaller stage means more resident CTAs: K=100 80 -> 94 M env-steps/s); otherwise up to 8 KB of rows per flush, reduced as
  // long as that lets one more CTA fit an SM (C5: 10 rows = 73 KB per CTA = 3 CTAs/SM, 4 rows = 49 KB = 4 CTAs/SM: 12.5 -> 13.5 M
  // env-steps/s).  MDX_STAGE_BYTES overrides the cap and disables the search.
  b.sel_k = std::min(b.K_p, e->P);
  b.mul_ksel = magic((uint32_t)std::max(1, b.sel_k));
  b.sel_rp = std::max(1, 32 / std::max(1, b.sel_k));
  b.stage_b = round_up(6 * std::min(e->R, b.MR) + 12, 4);
  // the key array is only touched by the selection paths for more than 5 slots per robot on a long waiting list (P > 64);
  // with <= 5 slots those environments keep their candidates in registers (featurize_env), so C4 / C5 do not pay for it
  const bool need_wkey = e->P <= 64 || b.sel_k > 5;
  auto carve = [&](int rpf, bool commit) {
    int off = 0;
    auto take = [&](int bytes) {
      const int at = off;
      off = round_up(off + bytes, 16);
      return at;
    };
    const int o_sched = take(8 * b.Pp8);
    const int o_status = take(b.Pp16);
    const int o_wlist = take(3 * 2 * b.Pp8);  // wlist, wrc, wdl
    const int o_wkey = take(need_wkey ? 4 * b.Pp8 : 0);
    const int o_rinfo = take(2 * 5 * 32);
    const int o_alist = take(2 * std::max(8, round_up(std::min(b.MP, std::max(1, e->P)), 8)));
    const int o_sel = take(2 * std::max(8, round_up(rpf * std::max(1, b.sel_k), 8)));
    const int o_masks = take(4 * 10 * b.W32);
    const int o_stage = off;
    // (the fused IDQ featuriser keeps its H*W urgency plane in the stage)
    const int stage_floats = std::max(std::max(rpf * b.D + 8, b.stage_b + 7 * std::min(e->P, b.MP) + 12), b.HW + 4);
    off = round_up(off + 4 * stage_floats, 16);
    if (commit) {
      b.off_sched = o_sched, b.off_status = o_status, b.off_wlist = o_wlist, b.off_wkey = o_wkey, b.off_rinfo = o_rinfo;
      b.off_alist = o_alist, b.off_sel = o_sel, b.off_masks = o_masks, b.off_stage = o_stage;
      b.vec_rows_per_flush = rpf;
      b.smem_per_warp = off;
    }
    return off;
  };
  auto ctas_per_sm = [&](int spw) { return std::min(7, (int)((227 * 1024) / ((size_t)kLutBytes + 4 * (size_t)spw + 1024))); };  // + 1 KB the driver reserves per CTA
  int rpf = e->R;
  int stage_cap = b.D * 4 > 2048 ? 4096 : 8192;
  const char* sc = getenv("MDX_STAGE_BYTES");
  if (sc) stage_cap = std::max(1024, atoi(sc));
  while (rpf > 1 && rpf * b.D * 4 > stage_cap) --rpf;
  if (!sc) {
    int best = rpf, best_ctas = ctas_per_sm(carve(rpf, false));
    for (int r = rpf - 1; r >= 1; --r) {
      const int c = ctas_per_sm(carve(r, false));
      if (c > best_ctas) best = r, best_ctas = c;
    }
    rpf = best;
  }
  carve(rpf, true);
  if (b.smem_per_warp > 220 * 1024)
    return fail(MDX_E_INVALID, "per-environment working set (%d bytes) exceeds shared memory", b.smem_per_warp);

  // ---- seeds
  e->rng.resize((size_t)e->n);
  for (int i = 0; i < e->n; ++i) {
    const int64_t s = cfg->env_seed ? cfg->env_seed[i] : cfg->base_seed + i;
    if (s < 0 || s > 0xFFFFFFFFll) return fail(MDX_E_INVALID, "Seed must be between 0 and 2**32 - 1");
    e->rng[(size_t)i].seed((uint32_t)s);
  }
  e->sampled.assign((size_t)e->n, 0);
  e->ep_host.assign((size_t)e->n, 0);

  // ---- device
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(MDX_E_CUDA, "no CUDA device: libmdx has no CPU fallback");
  if (cfg->device < 0 || cfg->device >= ndev) return fail(MDX_E_INVALID, "device ordinal %d out of range", cfg->device);
  DeviceGuard guard(e->device);
  const size_t n = (size_t)e->n;
  mdx_env* raw = e.get();
#define CUX(call)                                                                      \
  do {                                                                                 \
    cudaError_t e_ = (call);                                                           \
    if (e_ != cudaSuccess) {                                                           \
      fail(MDX_E_CUDA, "%s: %s", #call, cudaGetErrorString(e_));                       \
      mdx_destroy(e.release());                                                        \
      return MDX_E_CUDA;                                                               \
    }                                                                                  \
  } while (0)
  CUX(cudaMalloc(&raw->d_blob, blob.size()));
  {
    MapDev* md = reinterpret_cast<MapDev*>(blob.data() + off_maps);
    unsigned char* dbase = static_cast<unsigned char*>(raw->d_blob);
    for (int m = 0; m < cfg->n_maps; ++m) {
      md[m].H = e->maps[(size_t)m].rows;
      md[m].W = e->maps[(size_t)m].cols;
      md[m].nranks = nranks[(size_t)m];
      md[m].pad = 0;
      md[m].grid_bits = reinterpret_cast<const uint32_t*>(dbase + offs[(size_t)m].grid);
      md[m].qr = reinterpret_cast<const float*>(dbase + offs[(size_t)m].qr);
      md[m].qc = reinterpret_cast<const float*>(dbase + offs[(size_t)m].qc);
      md[m].rank = reinterpret_cast<const uint16_t*>(dbase + offs[(size_t)m].rank);
      md[m].bfs = nullptr;
    }
    CUX(cudaMemcpy(raw->d_blob, blob.data(), blob.size(), cudaMemcpyHostToDevice));
    raw->d_maps = reinterpret_cast<MapDev*>(dbase + off_maps);
    b.maps = raw->d_maps;
    b.map0 = md[0];
    b.tq = reinterpret_cast<const float*>(dbase + off_tq);
    b.cq = reinterpret_cast<const float*>(dbase + off_cq);
  }
  CUX(cudaMalloc(&raw->d_t, sizeof(int32_t) * n));
  CUX(cudaMalloc(&raw->d_ep, sizeof(int32_t) * n));
  CUX(cudaMalloc(&raw->d_ep_avail, sizeof(int32_t) * n));
  CUX(cudaMalloc(&raw->d_total, sizeof(double) * n));
  CUX(cudaMalloc(&raw->d_robots, sizeof(uint32_t) * n * e->R));
  CUX(cudaMalloc(&raw->d_status, n * b.Pp16));
  CUX(cudaMalloc(&raw->d_sched, sizeof(uint16_t) * (size_t)e->E * n * 4 * b.Pp8));
  CUX(cudaMalloc(&raw->d_cur, sizeof(uint16_t) * n * 4 * b.Pp8));
  CUX(cudaMemset(raw->d_cur, 0, sizeof(uint16_t) * n * 4 * b.Pp8));
  CUX(cudaMalloc(&raw->d_sched_robots, sizeof(uint16_t) * (size_t)e->E * n * 32));
  CUX(cudaMalloc(&raw->d_counters, sizeof(uint32_t) * n * 8));
  CUX(cudaMalloc(&raw->d_env_map, sizeof(int32_t) * n));
  CUX(cudaMalloc(&raw->d_stage, sizeof(uint16_t) * n * e->rec_u16));
  CUX(cudaMalloc(&raw->d_ids, sizeof(int32_t) * 2 * n));
  CUX(cudaMallocHost(&raw->h_stage, sizeof(uint16_t) * n * e->rec_u16));
  CUX(cudaMallocHost(&raw->h_ids, sizeof(int32_t) * 2 * n));
  CUX(cudaMallocHost(&raw->h_i32, sizeof(int32_t) * n));
  CUX(cudaMemset(raw->d_t, 0, sizeof(int32_t) * n));
  CUX(cudaMemset(raw->d_ep, 0, sizeof(int32_t) * n));
  CUX(cudaMemset(raw->d_ep_avail, 0, sizeof(int32_t) * n));
  CUX(cudaMemset(raw->d_total, 0, sizeof(double) * n));
  CUX(cudaMemset(raw->d_counters, 0, sizeof(uint32_t) * n * 8));
  CUX(cudaMemset(raw->d_sched, 0, sizeof(uint16_t) * (size_t)e->E * n * 4 * b.Pp8));
  CUX(cudaMemset(raw->d_sched_robots, 0, sizeof(uint16_t) * (size_t)e->E * n * 32));
  CUX(cudaMemcpy(raw->d_env_map, e->env_map.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice));
#undef CUX
#define CUX2(call)                                                   \
  do {                                                               \
    cudaError_t e_ = (call);                                         \
    if (e_ != cudaSuccess) {                                         \
      fail(MDX_E_CUDA, "%s: %s", #call, cudaGetErrorString(e_));     \
      mdx_destroy(e.release());                                      \
      return MDX_E_CUDA;                                             \
    }                                                                \
  } while (0)
  b.t = raw->d_t;
  b.ep = raw->d_ep;
  b.ep_avail = raw->d_ep_avail;
  b.total_reward = raw->d_total;
  b.robots = raw->d_robots;
  b.status = raw->d_status;
  b.sched = raw->d_sched;
  b.cur = raw->d_cur;
  b.sched_robots = raw->d_sched_robots;
  b.counters = raw->d_counters;
  b.env_map = cfg->n_maps > 1 ? raw->d_env_map : nullptr;

  CUX2(cudaMemset(raw->d_robots, 0, sizeof(uint32_t) * n * e->R));
  CUX2(cudaMemset(raw->d_status, 0, n * b.Pp16));
  if (cfg->flags & MDX_CFG_BFS_SHAPING) {
    b.shaping_bfs = 1;
    int rcb = ensure_bfs(raw);
    if (rcb != MDX_OK) {
      mdx_destroy(e.release());
      return rcb;
    }
  }
  if (cfg->flags & MDX_CFG_BLANK) {
    // blank envs for state injection; their first mdx_reset draws block 0.  ep = -1 is never read
    // by a kernel before that reset because slot arithmetic uses the sched of slot 0 (all zero).
    *out = e.release();
    return MDX_OK;
  }
  // ---- the constructor's own reset (env.py:41): block 0 for every env
  std::vector<int32_t> all((size_t)e->n), want((size_t)e->n, 1);
  for (int i = 0; i < e->n; ++i) all[(size_t)i] = i;
  int rc = top_up(raw, all, want, 0);
  if (rc != MDX_OK) {
    mdx_destroy(e.release());
    return rc;
  }
  KParams kp = b;
  kp.env_ids = nullptr;
  kp.n = e->n;
  kp.flags = 0;  // initialise in place (episode 0)
  mdx_reset_kernel<<<(unsigned)((e->n + 7) / 8), 256>>>(kp);
  if (cudaDeviceSynchronize() != cudaSuccess) {
    fail(MDX_E_CUDA, "reset kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
    mdx_destroy(e.release());
    return MDX_E_CUDA;
  }
  *out = e.release();
  return MDX_OK;
}

int32_t mdx_get_dims(const mdx_env* e, mdx_dims* out) {
  if (!e || !out) return fail(MDX_E_INVALID, "null argument");
  const KParams& b = e->base;
  out->n_envs = e->n;
  out->n_robots = e->R;
  out->n_packages = e->P;
  out->max_time_steps = e->T;
  out->rows_out = e->H_out;
  out->cols_out = e->W_out;
  out->vec_dim = b.D;
  out->gvec_dim = b.Dg;
  out->bytes_per_env_step = 4LL * ((long long)e->R * 6 * b.HW + (long long)e->R * b.D + 4LL * b.HW + b.Dg) + 4 + 1 + e->R;
  out->ring_episodes = e->E;
  out->reserved = 0;
  return MDX_OK;
}

int32_t mdx_reset(mdx_env* e, const int32_t* h_env_ids, int32_t n, void* stream) {
  if (!e) return fail(MDX_E_INVALID, "null env");
  DeviceGuard g(e->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  std::vector<int32_t> ids;
  if (h_env_ids == nullptr) {
    ids.resize((size_t)e->n);
    for (int i = 0; i < e->n; ++i) ids[(size_t)i] = i;
  } else {
    if (n < 0) return fail(MDX_E_INVALID, "negative count");
    ids.assign(h_env_ids, h_env_ids + n);
    for (int32_t id : ids)
      if (id < 0 || id >= e->n) return fail(MDX_E_INVALID, "env id %d out of range", id);
  }
  if (ids.empty()) return MDX_OK;
  int rc = refresh_ep(e, st);
  if (rc != MDX_OK) return rc;
  // an env listed k times is reset k times (the reference would draw k blocks); handle one pass per
  // repetition so that ring slots are never skipped
  std::vector<int32_t> remaining(ids);
  while (!remaining.empty()) {
    std::vector<int32_t> pass, later;
    std::vector<char> seen((size_t)e->n, 0);
    for (int32_t id : remaining) {
      if (seen[(size_t)id]) later.push_back(id);
      else {
        seen[(size_t)id] = 1;
        pass.push_back(id);
      }
    }
    std::vector<int32_t> want(e->sampled);
    for (int32_t id : pass) want[(size_t)id] = std::max(want[(size_t)id], e->ep_host[(size_t)id] + 2);
    rc = top_up(e, pass, want, st);
    if (rc != MDX_OK) return rc;
    for (size_t k = 0; k < pass.size(); ++k) e->h_ids[k] = pass[k];
    CU(cudaMemcpyAsync(e->d_ids, e->h_ids, sizeof(int32_t) * pass.size(), cudaMemcpyHostToDevice, st));
    KParams kp = e->base;
    kp.env_ids = e->d_ids;
    kp.n = (int)pass.size();
    kp.flags = 1;
    mdx_reset_kernel<<<(unsigned)((pass.size() + 7) / 8), 256, 0, st>>>(kp);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(st));
    for (int32_t id : pass) e->ep_host[(size_t)id] += 1;
    remaining.swap(later);
  }
  return MDX_OK;
}

int32_t mdx_prefetch(mdx_env* e, int32_t ahead, void* stream) {
  if (!e) return fail(MDX_E_INVALID, "null env");
  if (ahead < 0 || ahead > e->E - 1) return fail(MDX_E_INVALID, "episodes_ahead must be in [0, ring_episodes - 1]");
  DeviceGuard g(e->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = refresh_ep(e, st);
  if (rc != MDX_OK) return rc;
  std::vector<int32_t> all((size_t)e->n), want((size_t)e->n);
  for (int i = 0; i < e->n; ++i) {
    all[(size_t)i] = i;
    want[(size_t)i] = std::max(e->sampled[(size_t)i], e->ep_host[(size_t)i] + 1 + ahead);
  }
  return top_up(e, all, want, st);
}

static int step_common(mdx_env* e, int mode, const uint8_t* d_actions, const int32_t* d_env_ids, int32_t n, double* d_reward,
                       float* d_shaped, uint8_t* d_done, float* d_obs, float* d_vec, float* d_gstate, float* d_gvec,
                       uint32_t flags, void* stream, double* d_agent = nullptr, int ops_as_int = 0, float* d_obs6 = nullptr,
                       float* d_g7 = nullptr) {
  if (!e) return fail(MDX_E_INVALID, "null env");
  if ((mode & 1) && !d_actions) return fail(MDX_E_INVALID, "actions are required");
  if (d_env_ids == nullptr && n != e->n) return fail(MDX_E_INVALID, "n must equal n_envs when env ids are not given (got %d, have %d)", n, e->n);
  if (n < 0 || n > e->n) return fail(MDX_E_INVALID, "bad env count %d", n);
  if (n == 0) return MDX_OK;
  int rc;
  if ((rc = check_ptr16(d_obs, "obs")) || (rc = check_ptr16(d_vec, "vec")) || (rc = check_ptr16(d_gstate, "gstate")) ||
      (rc = check_ptr16(d_gvec, "gvec")))
    return rc;
  DeviceGuard g(e->device);
  KParams kp = e->base;
  kp.actions = d_actions;
  kp.env_ids = d_env_ids;
  kp.n = n;
  kp.flags = flags & 3u;
  kp.out_reward = d_reward;
  kp.out_shaped = d_shaped;
  kp.out_done = d_done;
  kp.obs = d_obs;
  kp.vec = d_vec;
  kp.gstate = d_gstate;
  kp.gvec = d_gvec;
  kp.out_agent = d_agent;
  kp.idq_ops_int = ops_as_int;
  kp.alt_obs6 = d_obs6;
  kp.alt_g7 = d_g7;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (mode == 5) return launch_env_kernel_5(e, kp, st);
  if (mode == 1) return launch_env_kernel_1(e, kp, st);
  if (mode == 2) return launch_env_kernel_2(e, kp, st);
  return launch_env_kernel_3(e, kp, st);
}

int32_t mdx_step(mdx_env* e, const uint8_t* d_actions, const int32_t* d_env_ids, int32_t n, double* d_reward, float* d_shaped,
                 uint8_t* d_done, uint32_t flags, void* stream) {
  return step_common(e, 1, d_actions, d_env_ids, n, d_reward, d_shaped, d_done, nullptr, nullptr, nullptr, nullptr, flags, stream);
}

int32_t mdx_step_idq(mdx_env* e, const uint8_t* d_actions, const int32_t* d_env_ids, int32_t n, double* d_reward, double* d_agent_rewards,
                     uint8_t* d_done, int32_t ops_as_int, uint32_t flags, void* stream) {
  if (!d_agent_rewards) return fail(MDX_E_INVALID, "agent reward output is required");
  return step_common(e, 1, d_actions, d_env_ids, n, d_reward, nullptr, d_done, nullptr, nullptr, nullptr, nullptr, flags, stream,
                     d_agent_rewards, ops_as_int ? 1 : 0);
}

int32_t mdx_step_idq_featurize(mdx_env* e, const uint8_t* d_actions, const int32_t* d_env_ids, int32_t n, double* d_reward,
                               double* d_agent_rewards, uint8_t* d_done, float* d_obs6, float* d_gstate7, int32_t ops_as_int,
                               uint32_t flags, void* stream) {
  if (!d_agent_rewards) return fail(MDX_E_INVALID, "agent reward output is required");
  int rc;
  if ((rc = check_ptr16(d_obs6, "obs6")) || (rc = check_ptr16(d_gstate7, "gstate7"))) return rc;
  rc = getenv("MDX_IDQ_UNFUSED") ? MDX_E_UNSUPPORTED_SHAPE
                                 : step_common(e, 5, d_actions, d_env_ids, n, d_reward, nullptr, d_done, nullptr, nullptr, nullptr, nullptr,
                                               flags, stream, d_agent_rewards, ops_as_int ? 1 : 0, d_obs6, d_gstate7);
  if (rc != MDX_E_UNSUPPORTED_SHAPE) return rc;
  // maps whose planes are not whole float4s (or larger than 4 x 32 float4): the same result from two launches
  rc = mdx_step_idq(e, d_actions, d_env_ids, n, d_reward, d_agent_rewards, d_done, ops_as_int, flags, stream);
  if (rc != MDX_OK) return rc;
  return mdx_featurize_alt(e, d_env_ids, n, d_obs6, d_gstate7, stream);
}

int32_t mdx_featurize(mdx_env* e, const int32_t* d_env_ids, int32_t n, float* d_obs, float* d_vec, float* d_gstate, float* d_gvec,
                      void* stream) {
  return step_common(e, 2, nullptr, d_env_ids, n, nullptr, nullptr, nullptr, d_obs, d_vec, d_gstate, d_gvec, 0, stream);
}

int32_t mdx_step_featurize(mdx_env* e, const uint8_t* d_actions, const int32_t* d_env_ids, int32_t n, double* d_reward,
                           float* d_shaped, uint8_t* d_done, float* d_obs, float* d_vec, float* d_gstate, float* d_gvec,
                           uint32_t flags, void* stream) {
  return step_common(e, 3, d_actions, d_env_ids, n, d_reward, d_shaped, d_done, d_obs, d_vec, d_gstate, d_gvec, flags, stream);
}

static int shape_common(mdx_env* e, const uint32_t* d_cur_robots, const int32_t* d_cur_t, const uint8_t* d_actions,
                        const double* d_global_reward, float* d_out, double* d_agent, int ops_as_int, void* stream);

int32_t mdx_shaped_reward(mdx_env* e, const uint32_t* d_cur_robots, const int32_t* d_cur_t, const uint8_t* d_actions,
                          const double* d_global_reward, float* d_out, void* stream) {
  if (!e || !d_cur_robots || !d_cur_t || !d_actions || !d_global_reward || !d_out) return fail(MDX_E_INVALID, "null argument");
  return shape_common(e, d_cur_robots, d_cur_t, d_actions, d_global_reward, d_out, nullptr, 0, stream);
}

int32_t mdx_reward_shaping_idq(mdx_env* e, const uint32_t* d_cur_robots, const int32_t* d_cur_t, const uint8_t* d_actions,
                               int32_t ops_as_int, double* d_out, void* stream) {
  if (!e || !d_cur_robots || !d_cur_t || !d_actions || !d_out) return fail(MDX_E_INVALID, "null argument");
  return shape_common(e, d_cur_robots, d_cur_t, d_actions, nullptr, nullptr, d_out, ops_as_int ? 1 : 0, stream);
}

static int shape_common(mdx_env* e, const uint32_t* d_cur_robots, const int32_t* d_cur_t, const uint8_t* d_actions,
                        const double* d_global_reward, float* d_out, double* d_agent, int ops_as_int, void* stream) {
  DeviceGuard g(e->device);
  KParams kp = e->base;
  kp.out_agent = d_agent;
  kp.idq_ops_int = ops_as_int;
  kp.cur_robots = d_cur_robots;
  kp.cur_t = d_cur_t;
  kp.actions = d_actions;
  kp.global_reward = d_global_reward;
  kp.out_shaped = d_out;
  kp.n = e->n;
  const int spw = e->base.off_alist;
  const int wpb = pick_wpb(e, spw);
  kp.smem_per_warp = spw;
  const size_t smem = (size_t)kLutBytes + (size_t)wpb * spw;
  CU(cudaFuncSetAttribute(mdx_shape_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 48 * 1024)));
  mdx_shape_kernel<<<(unsigned)((e->n + wpb - 1) / wpb), wpb * 32, smem, static_cast<cudaStream_t>(stream)>>>(kp);
  CU(cudaGetLastError());
  return MDX_OK;
}

int32_t mdx_get_state(mdx_env* e, int32_t id, mdx_state_view* out, void* stream) {
  if (!e || !out) return fail(MDX_E_INVALID, "null argument");
  if (id < 0 || id >= e->n) return fail(MDX_E_INVALID, "env id %d out of range", id);
  DeviceGuard g(e->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const KParams& b = e->base;
  int32_t t = 0, ep = 0;
  double total = 0;
  std::vector<uint32_t> rob((size_t)e->R);
  std::vector<uint8_t> status((size_t)b.Pp16);
  std::vector<uint16_t> sched((size_t)4 * b.Pp8);
  CU(cudaStreamSynchronize(st));
  CU(cudaMemcpy(&t, e->d_t + id, sizeof(t), cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(&ep, e->d_ep + id, sizeof(ep), cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(&total, e->d_total + id, sizeof(total), cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(rob.data(), e->d_robots + (size_t)id * e->R, sizeof(uint32_t) * e->R, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(status.data(), e->d_status + (size_t)id * b.Pp16, (size_t)b.Pp16, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(sched.data(), e->d_cur + (size_t)id * 4 * b.Pp8, sizeof(uint16_t) * 4 * b.Pp8, cudaMemcpyDeviceToHost));
  out->t = t;
  out->episode = ep;
  out->map_id = e->env_map[(size_t)id];
  out->total_reward = total;
  const int Wo = e->W_out;
  for (int i = 0; i < e->R; ++i) {
    const int cell = (int)(rob[(size_t)i] & 0xffffu);
    if (out->robot_row) out->robot_row[i] = cell / Wo;
    if (out->robot_col) out->robot_col[i] = cell % Wo;
    if (out->robot_carry) out->robot_carry[i] = (int)(rob[(size_t)i] >> 16);
  }
  for (int j = 0; j < e->P; ++j) {
    const int s = sched[(size_t)(0 * b.Pp8 + j)], g2 = sched[(size_t)(1 * b.Pp8 + j)];
    if (out->pkg_start_row) out->pkg_start_row[j] = s / Wo;
    if (out->pkg_start_col) out->pkg_start_col[j] = s % Wo;
    if (out->pkg_target_row) out->pkg_target_row[j] = g2 / Wo;
    if (out->pkg_target_col) out->pkg_target_col[j] = g2 % Wo;
    if (out->pkg_start_time) out->pkg_start_time[j] = sched[(size_t)(2 * b.Pp8 + j)];
    if (out->pkg_deadline) out->pkg_deadline[j] = sched[(size_t)(3 * b.Pp8 + j)];
    if (out->pkg_status) out->pkg_status[j] = status[(size_t)j];
  }
  return MDX_OK;
}

int32_t mdx_set_state(mdx_env* e, int32_t id, const mdx_state_view* in, void* stream) {
  if (!e || !in) return fail(MDX_E_INVALID, "null argument");
  if (id < 0 || id >= e->n) return fail(MDX_E_INVALID, "env id %d out of range", id);
  if (!in->robot_row || !in->robot_col || !in->robot_carry) return fail(MDX_E_INVALID, "robot arrays are required");
  if (e->P > 0 && (!in->pkg_start_row || !in->pkg_start_col || !in->pkg_target_row || !in->pkg_target_col || !in->pkg_start_time ||
                   !in->pkg_deadline || !in->pkg_status))
    return fail(MDX_E_INVALID, "package arrays are required");
  DeviceGuard g(e->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const KParams& b = e->base;
  const MapHost& m = e->maps[(size_t)e->env_map[(size_t)id]];
  const int Wo = e->W_out;
  const long long dl_cap = std::min<long long>(65535, ((1LL << (32 - b.key_wbits)) - 2) / b.nranks_max - 1);
  std::vector<uint32_t> rob((size_t)e->R);
  for (int i = 0; i < e->R; ++i) {
    const int r = in->robot_row[i], c = in->robot_col[i], k = in->robot_carry[i];
    if (r < 0 || r >= m.rows || c < 0 || c >= m.cols) return fail(MDX_E_INVALID, "robot %d outside the map", i);
    if (k < 0 || k > e->P) return fail(MDX_E_INVALID, "robot %d carries an unknown package id %d", i, k);
    rob[(size_t)i] = (uint32_t)(r * Wo + c) | ((uint32_t)k << 16);
  }
  std::vector<uint8_t> status((size_t)b.Pp16, 0);
  std::vector<uint16_t> sched((size_t)4 * b.Pp8, 0);
  for (int j = b.P; j < b.Pp8; ++j) sched[(size_t)(2 * b.Pp8 + j)] = 0xffff;
  for (int j = 0; j < e->P; ++j) {
    const int sr = in->pkg_start_row[j], sc = in->pkg_start_col[j], tr = in->pkg_target_row[j], tc = in->pkg_target_col[j];
    if (sr < 0 || sr >= m.rows || sc < 0 || sc >= m.cols || tr < 0 || tr >= m.rows || tc < 0 || tc >= m.cols)
      return fail(MDX_E_INVALID, "package %d outside the map", j + 1);
    if (in->pkg_start_time[j] < 0 || in->pkg_start_time[j] > 65534) return fail(MDX_E_INVALID, "package %d: bad start_time", j + 1);
    if (in->pkg_deadline[j] < 0 || in->pkg_deadline[j] > dl_cap) return fail(MDX_E_INVALID, "package %d: deadline out of range", j + 1);
    if (in->pkg_status[j] < 0 || in->pkg_status[j] > 3) return fail(MDX_E_INVALID, "package %d: bad status", j + 1);
    sched[(size_t)(0 * b.Pp8 + j)] = (uint16_t)(sr * Wo + sc);
    sched[(size_t)(1 * b.Pp8 + j)] = (uint16_t)(tr * Wo + tc);
    sched[(size_t)(2 * b.Pp8 + j)] = (uint16_t)in->pkg_start_time[j];
    sched[(size_t)(3 * b.Pp8 + j)] = (uint16_t)in->pkg_deadline[j];
    status[(size_t)j] = (uint8_t)in->pkg_status[j];
  }
  CU(cudaStreamSynchronize(st));
  const int32_t t = in->t;
  const double total = in->total_reward;
  CU(cudaMemcpy(e->d_t + id, &t, sizeof(t), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(e->d_total + id, &total, sizeof(total), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(e->d_robots + (size_t)id * e->R, rob.data(), sizeof(uint32_t) * e->R, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(e->d_status + (size_t)id * b.Pp16, status.data(), (size_t)b.Pp16, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(e->d_cur + (size_t)id * 4 * b.Pp8, sched.data(), sizeof(uint16_t) * 4 * b.Pp8, cudaMemcpyHostToDevice));
  return MDX_OK;
}

int32_t mdx_episode_stats(mdx_env* e, int64_t* d_out8, void* stream) {
  if (!e || !d_out8) return fail(MDX_E_INVALID, "null argument");
  DeviceGuard g(e->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CU(cudaMemsetAsync(d_out8, 0, sizeof(int64_t) * 8, st));
  const int blocks = std::max(1, std::min(148 * 4, (e->n + 31) / 32));
  mdx_stats_kernel<<<blocks, 256, 0, st>>>(e->d_counters, e->n, reinterpret_cast<unsigned long long*>(d_out8));
  CU(cudaGetLastError());
  return MDX_OK;
}

// ---- host-only entry points --------------------------------------------------------------------

int32_t mdx_sampler_create(uint32_t seed, const uint8_t* cells, int32_t rows, int32_t cols, int32_t n_robots, int32_t n_packages,
                           int32_t max_time_steps, mdx_sampler** out) {
  if (!cells || !out || rows <= 0 || cols <= 0 || n_robots < 0 || n_packages < 0) return fail(MDX_E_INVALID, "bad sampler argument");
  mdx_sampler* s = new mdx_sampler();
  s->s.init(seed, cells, rows, cols, n_robots, n_packages, max_time_steps);
  *out = s;
  return MDX_OK;
}

int32_t mdx_sampler_next(mdx_sampler* s, int32_t* robot_cell, int32_t* pkg4) {
  if (!s || !robot_cell || !pkg4) return fail(MDX_E_INVALID, "null argument");
  if (!s->s.next(robot_cell, pkg4)) return fail(MDX_E_INVALID, "the reference would raise ValueError (empty randint range)");
  return MDX_OK;
}

int32_t mdx_sampler_destroy(mdx_sampler* s) {
  delete s;
  return MDX_OK;
}

int32_t mdx_rng_randint(uint32_t seed, int64_t lo, int64_t hi, int32_t n, int64_t* out) {
  if (!out || n < 0) return fail(MDX_E_INVALID, "bad argument");
  Mt19937 g;
  g.seed(seed);
  for (int i = 0; i < n; ++i)
    if (!g.randint(lo, hi, out + i)) return fail(MDX_E_INVALID, "low >= high");
  return MDX_OK;
}

int32_t mdx_lut_quotient(int32_t n, int32_t n_out, float* out) {
  if (!out || n <= 0 || n_out <= 0) return fail(MDX_E_INVALID, "bad argument");
  build_quotient_lut(n, n_out, out);
  return MDX_OK;
}

int32_t mdx_lut_rank(int32_t rows, int32_t cols, int32_t rows_out, int32_t cols_out, uint16_t* out, int32_t* n_ranks) {
  if (!out || rows <= 0 || cols <= 0 || rows_out < rows || cols_out < cols) return fail(MDX_E_INVALID, "bad argument");
  const int nr = build_rank_lut(rows, cols, rows_out, cols_out, out);
  if (n_ranks) *n_ranks = nr;
  return MDX_OK;
}

float mdx_np_sum_f32(const float* x, int32_t n) { return np_sum_f32(x, n); }

}  // extern "C"
