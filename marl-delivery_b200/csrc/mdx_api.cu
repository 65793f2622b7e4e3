// C-ABI glue of libmdx: configuration, device state, the host reset sampler's ring management and
// kernel launches.  See include/mdx.h for the contract of every entry point.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/mdx.h"
#include "mdx_host.h"
#include "mdx_small_kernels.cuh"

using namespace mdx;

namespace mdx {
thread_local std::string g_err;

int ensure_dynamic_smem(const void* func, size_t bytes, int carveout) {
  struct Mark {
    const void* func;
    int device;
    size_t bytes;
  };
  static std::mutex mu;
  static std::vector<Mark> marks;
  int dev = 0;
  CU(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  Mark* m = nullptr;
  for (Mark& x : marks)
    if (x.func == func && x.device == dev) m = &x;
  if (m == nullptr) {
    marks.push_back(Mark{func, dev, 0});
    m = &marks.back();
  }
  if (bytes > m->bytes || m->bytes == 0) {
    CU(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(bytes, 48 * 1024)));
    if (carveout >= 0) CU(cudaFuncSetAttribute(func, cudaFuncAttributePreferredSharedMemoryCarveout, carveout));
    m->bytes = std::max<size_t>(bytes, 48 * 1024);
  }
  return MDX_OK;
}

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

// Launches the device reset (draw the next block on the env's own stream, apply it) for `n` envs; d_ids == nullptr: all.
int launch_reset(mdx_env* e, const int32_t* d_ids, int n, uint32_t advance, cudaStream_t st) {
  KParams kp = e->base;
  kp.env_ids = d_ids;
  kp.n = n;
  kp.flags = advance;
  const size_t per_warp = (size_t)8 * e->base.Pp8;
  int wpb = 8;
  while (wpb > 1 && wpb * per_warp > 160 * 1024) wpb >>= 1;
  const size_t smem = wpb * per_warp;
  int rc = ensure_dynamic_smem(reinterpret_cast<const void*>(mdx_reset_kernel), smem, -1);
  if (rc != MDX_OK) return rc;
  mdx_reset_kernel<<<(unsigned)((n + wpb - 1) / wpb), wpb * 32, smem, st>>>(kp);
  CU(cudaGetLastError());
  return MDX_OK;
}

inline size_t up16(size_t x) { return (x + 15) & ~(size_t)15; }

// Makes the pinned host buffer and its device mirror at least `bytes` large and safe to overwrite.
int staging(mdx_env* e, size_t bytes) {
  if (e->pin_busy == nullptr) CU(cudaEventCreateWithFlags(&e->pin_busy, cudaEventDisableTiming));
  else CU(cudaEventSynchronize(e->pin_busy));
  if (bytes <= e->pin_bytes) return MDX_OK;
  size_t cap = 64 * 1024;
  while (cap < bytes) cap <<= 1;
  CU(cudaDeviceSynchronize());
  cudaFree(e->d_pin);
  cudaFreeHost(e->h_pin);
  e->d_pin = nullptr, e->h_pin = nullptr, e->pin_bytes = 0;
  CU(cudaMalloc(&e->d_pin, cap));
  CU(cudaHostAlloc(&e->h_pin, cap, cudaHostAllocMapped | cudaHostAllocPortable));  // device-addressable (unified addressing)
  e->pin_bytes = cap;
  return MDX_OK;
}

int check_ids(const mdx_env* e, const int32_t* h_ids, int32_t n, bool unique = false) {
  if (n < 0) return fail(MDX_E_INVALID, "negative count");
  if (unique && h_ids != nullptr && n > 1) {  // two lane groups working on one env in the same launch would race
    if (n <= 32) {  // (the dict drop-ins step a handful of envs per call: no allocation on that path)
      for (int k = 1; k < n; ++k)
        for (int j = 0; j < k; ++j)
          if (h_ids[j] == h_ids[k]) return fail(MDX_E_INVALID, "env id %d is listed twice", h_ids[k]);
    } else {
      std::vector<char> seen((size_t)e->n, 0);
      for (int k = 0; k < n; ++k) {
        if (h_ids[k] < 0 || h_ids[k] >= e->n) return fail(MDX_E_INVALID, "env id %d out of range", h_ids[k]);
        if (seen[(size_t)h_ids[k]]) return fail(MDX_E_INVALID, "env id %d is listed twice", h_ids[k]);
        seen[(size_t)h_ids[k]] = 1;
      }
    }
  }
  if (h_ids == nullptr) {
    if (n != e->n) return fail(MDX_E_INVALID, "n must equal n_envs when env ids are not given (got %d, have %d)", n, e->n);
    return MDX_OK;
  }
  for (int k = 0; k < n; ++k)
    if (h_ids[k] < 0 || h_ids[k] >= e->n) return fail(MDX_E_INVALID, "env id %d out of range", h_ids[k]);
  return MDX_OK;
}

RowArgs row_args(const mdx_env* e, uint32_t* d_rows, const int32_t* d_ids, int n) {
  RowArgs a;
  a.rows = d_rows;
  a.ids = d_ids;
  a.base = 0;
  a.n = n;
  a.row_words = e->row_words;
  a.off_robots = e->off_robots;
  a.off_sched = e->off_sched;
  a.off_status = e->off_status;
  return a;
}

// What Environment.add_robot / the package constructor would refuse, plus the ranges the u16 state can hold.
int validate_row(const mdx_env* e, int env, const uint32_t* row) {
  const KParams& b = e->base;
  const MapHost& m = e->maps[(size_t)e->env_map[(size_t)env]];
  const int Wo = e->W_out;
  const int32_t t = (int32_t)row[0];
  if (t < 0 || t > 65534) return fail(MDX_E_INVALID, "time step %d out of range [0, 65534]", t);
  const long long dl_cap = std::min<long long>(65535, ((1LL << (32 - b.key_wbits)) - 2) / b.nranks_max - 1);
  auto on_map = [&](int cell) { return cell / Wo < m.rows && cell % Wo < m.cols; };
  for (int i = 0; i < e->R; ++i) {
    const int cell = (int)(row[e->off_robots + i] & 0xffffu), k = (int)(row[e->off_robots + i] >> 16);
    if (!on_map(cell)) return fail(MDX_E_INVALID, "robot %d outside the map", i);
    // Environment.add_robot raises ValueError for an occupied or blocked cell (env.py:73-79)
    if (k > e->P) return fail(MDX_E_INVALID, "robot %d carries an unknown package id %d", i, k);
    if (e->cfg.flags & MDX_CFG_LAX_ROBOTS) continue;
    if (m.cells[(size_t)(cell / Wo) * m.cols + cell % Wo] != 0) return fail(MDX_E_INVALID, "Invalid robot position: robot %d stands on an obstacle", i);
    for (int j = 0; j < i; ++j)
      if ((row[e->off_robots + j] & 0xffffu) == (uint32_t)cell) return fail(MDX_E_INVALID, "Invalid robot position: robots %d and %d share a cell", j, i);
  }
  const uint16_t* sched = reinterpret_cast<const uint16_t*>(row + e->off_sched);
  const uint8_t* status = reinterpret_cast<const uint8_t*>(row + e->off_status);
  for (int j = 0; j < e->P; ++j) {
    if (!on_map(sched[j]) || !on_map(sched[b.Pp8 + j])) return fail(MDX_E_INVALID, "package %d outside the map", j + 1);
    if (sched[2 * b.Pp8 + j] > 65534) return fail(MDX_E_INVALID, "package %d: bad start_time", j + 1);
    if ((long long)sched[3 * b.Pp8 + j] > dl_cap) return fail(MDX_E_INVALID, "package %d: deadline out of range", j + 1);
    if (status[j] > 3) return fail(MDX_E_INVALID, "package %d: bad status", j + 1);
  }
  for (int j = e->P; j < b.Pp8; ++j)
    if (sched[2 * b.Pp8 + j] != 0xffff) return fail(MDX_E_INVALID, "padding package %d must have start_time 0xFFFF", j + 1);
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
  jit_release(e);
  cudaFree(e->d_blob);
  cudaFree(e->d_t);
  cudaFree(e->d_ep);
  cudaFree(e->d_mt);
  cudaFree(e->d_mtpos);
  cudaFree(e->d_total);
  cudaFree(e->d_robots);
  cudaFree(e->d_status);
  cudaFree(e->d_cur);
  cudaFree(e->d_counters);
  cudaFree(e->d_env_map);
  cudaFree(e->d_ids);
  cudaFree(e->d_trk);
  cudaFree(e->d_trk_status);
  cudaFree(e->d_trk_n);
  cudaFree(e->d_pin);
  cudaFreeHost(e->h_pin);
  if (e->pin_busy) cudaEventDestroy(e->pin_busy);
  cudaFree(e->d_bfs);
  cudaFree(e->d_greedy);
  cudaFree(e->d_gfree);
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

  if ((cfg->flags & MDX_CFG_STALE_TRACKER) && (cfg->flags & MDX_CFG_BFS_SHAPING))
    return fail(MDX_E_INVALID, "MDX_CFG_STALE_TRACKER and MDX_CFG_BFS_SHAPING cannot be combined");
  std::unique_ptr<mdx_env> e(new mdx_env());
  e->cfg = *cfg;
  e->stale = (cfg->flags & MDX_CFG_STALE_TRACKER) != 0;
  e->device = cfg->device;
  e->n = cfg->n_envs;
  e->R = cfg->n_robots;
  e->P = cfg->n_packages;
  e->T = cfg->max_time_steps;
  e->E = 0;
  if (const char* s = getenv("MDX_WPB")) e->wpb_override = atoi(s);
  if (const char* s = getenv("MDX_PDL")) e->pdl = atoi(s);
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
    // where the reference's reset would raise or never return (env.py:106-116): a single free cell cannot hold a start and a
    // different target; randint(1, T) needs T >= 2 once there are more than min(n_robots, 20) + 1 packages
    if (e->P > 0 && mh.free_cells.size() < 2)
      return fail(MDX_E_INVALID, "map %d: packages need at least two free cells (start != target, env.py:107-110)", m);
    Ho = std::max(Ho, mh.rows);
    Wo = std::max(Wo, mh.cols);
  }
  if (e->P > std::min(e->R, 20) + 1 && e->T < 2 && !(cfg->flags & MDX_CFG_BLANK))
    return fail(MDX_E_INVALID, "max_time_steps < 2 with more than min(n_robots, 20) + 1 packages: the reference raises ValueError "
                               "in reset (randint(1, max_time_steps), env.py:116)");
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

  e->off_robots = 6;
  e->off_sched = round_up(6 + e->R, 2);
  e->off_status = e->off_sched + 2 * b.Pp8;
  e->row_words = round_up(e->off_status + b.Pp16 / 4, 4);

  // ---- LUTs into one host blob
  const int nqr = 2 * Ho - 1, nqc = 2 * Wo - 1;
  struct MapOff {
    size_t grid, qr, qc, rank, free_cells;
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
    offs[(size_t)m].free_cells = reserve(sizeof(uint16_t) * mh.free_cells.size());
    {
      uint16_t* fc = reinterpret_cast<uint16_t*>(blob.data() + offs[(size_t)m].free_cells);
      for (size_t k = 0; k < mh.free_cells.size(); ++k)
        fc[k] = (uint16_t)((mh.free_cells[k] / mh.cols) * Wo + (mh.free_cells[k] % mh.cols));
    }
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
  auto carve = [&](int rpf, int gch, bool commit) {
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
    // (+ 32 candidate slots behind the selected indices for the few-slots / long-list selection path: P > 64, <= 5 slots)
    const int o_sel = take(2 * (std::max(8, round_up(rpf * std::max(1, b.sel_k), 8)) + ((e->P > 64 && b.sel_k <= 5) ? 32 : 0)));
    const int o_masks = take(4 * 10 * b.W32);
    const int o_stage = off;
    // (the fused IDQ featuriser keeps its H*W urgency plane in the stage)
    const int stage_floats = std::max(std::max(rpf * b.D + 8, b.stage_b + 7 * gch + 12), b.HW + 4);
    off = round_up(off + 4 * stage_floats, 16);
    const int o_trk = off;
    if (e->stale) off = round_up(off + 10 * b.Pp8 + b.Pp16 + 2 * b.Pp8, 16);
    if (commit) {
      b.off_trk = o_trk;
      b.off_sched = o_sched, b.off_status = o_status, b.off_wlist = o_wlist, b.off_wkey = o_wkey, b.off_rinfo = o_rinfo;
      b.off_alist = o_alist, b.off_sel = o_sel, b.off_masks = o_masks, b.off_stage = o_stage;
      b.vec_rows_per_flush = rpf;
      b.gvec_chunk = gch;
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
  // `gch` = package rows of gvec staged per flush (all min(P, MP) of them unless a smaller chunk lets one more CTA fit: C5 14.0 -> 10.6 KB
  // per warp = 4 -> 5 CTAs/SM)
  const int gch_full = std::max(1, std::min(e->P, b.MP));
  int gch = gch_full;
  if (!sc) {
    int best = rpf, best_g = gch_full, best_ctas = ctas_per_sm(carve(rpf, gch_full, false));
    for (int gc : {gch_full, 64, 48, 32, 24, 16}) {
      if (gc > gch_full) continue;
      for (int r = rpf; r >= 1; --r) {
        const int c = ctas_per_sm(carve(r, gc, false));
        if (c > best_ctas) best = r, best_g = gc, best_ctas = c;
      }
    }
    rpf = best, gch = best_g;
  }
  if (const char* gs = getenv("MDX_GVEC_CHUNK")) gch = std::max(1, std::min(gch_full, atoi(gs)));
  carve(rpf, gch, true);
  if (b.smem_per_warp > 220 * 1024)
    return fail(MDX_E_INVALID, "per-environment working set (%d bytes) exceeds shared memory", b.smem_per_warp);

  // ---- seeds (numpy: RandomState(seed) takes 0 <= seed < 2**32)
  std::vector<uint32_t> seeds((size_t)e->n);
  for (int i = 0; i < e->n; ++i) {
    const int64_t s = cfg->env_seed ? cfg->env_seed[i] : cfg->base_seed + i;
    if (s < 0 || s > 0xFFFFFFFFll) return fail(MDX_E_INVALID, "Seed must be between 0 and 2**32 - 1");
    seeds[(size_t)i] = (uint32_t)s;
  }

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
      md[m].free_cells = reinterpret_cast<const uint16_t*>(dbase + offs[(size_t)m].free_cells);
      md[m].n_free = (int32_t)e->maps[(size_t)m].free_cells.size();
      md[m].pad2 = 0;
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
  CUX(cudaMalloc(&raw->d_mt, sizeof(uint32_t) * n * 624));
  CUX(cudaMalloc(&raw->d_mtpos, sizeof(int32_t) * n));
  CUX(cudaMalloc(&raw->d_total, sizeof(double) * n));
  CUX(cudaMalloc(&raw->d_robots, sizeof(uint32_t) * n * e->R));
  CUX(cudaMalloc(&raw->d_status, n * b.Pp16));
  CUX(cudaMalloc(&raw->d_cur, sizeof(uint16_t) * n * 4 * b.Pp8));
  CUX(cudaMemset(raw->d_cur, 0, sizeof(uint16_t) * n * 4 * b.Pp8));
  CUX(cudaMalloc(&raw->d_counters, sizeof(uint32_t) * n * 8));
  CUX(cudaMalloc(&raw->d_env_map, sizeof(int32_t) * n));
  CUX(cudaMalloc(&raw->d_ids, sizeof(int32_t) * n));
  if (e->stale) {
    CUX(cudaMalloc(&raw->d_trk, sizeof(uint16_t) * n * 5 * b.Pp8));
    CUX(cudaMalloc(&raw->d_trk_status, n * b.Pp16));
    CUX(cudaMalloc(&raw->d_trk_n, sizeof(int32_t) * n));
    CUX(cudaMemset(raw->d_trk, 0, sizeof(uint16_t) * n * 5 * b.Pp8));
    CUX(cudaMemset(raw->d_trk_status, 0, n * b.Pp16));
    CUX(cudaMemset(raw->d_trk_n, 0, sizeof(int32_t) * n));
  }
  CUX(cudaMemset(raw->d_t, 0, sizeof(int32_t) * n));
  CUX(cudaMemset(raw->d_ep, 0, sizeof(int32_t) * n));
  CUX(cudaMemset(raw->d_total, 0, sizeof(double) * n));
  CUX(cudaMemset(raw->d_counters, 0, sizeof(uint32_t) * n * 8));
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
  b.mt = raw->d_mt;
  b.mtpos = raw->d_mtpos;
  b.total_reward = raw->d_total;
  b.robots = raw->d_robots;
  b.status = raw->d_status;
  b.cur = raw->d_cur;
  b.counters = raw->d_counters;
  b.trk = raw->d_trk;
  b.trk_status = raw->d_trk_status;
  b.trk_n = raw->d_trk_n;
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
  // every env's generator: init_genrand(seed) on the device (d_ids doubles as the seed staging area)
  CUX2(cudaMemcpy(raw->d_ids, seeds.data(), sizeof(uint32_t) * n, cudaMemcpyHostToDevice));
  mdx_seed_kernel<<<(unsigned)((e->n + 127) / 128), 128>>>(raw->d_mt, raw->d_mtpos, reinterpret_cast<const uint32_t*>(raw->d_ids), e->n);
  CUX2(cudaGetLastError());
  if (cfg->flags & MDX_CFG_BLANK) {
    // blank envs for state injection; their first mdx_reset draws block 0
    CUX2(cudaDeviceSynchronize());
    *out = e.release();
    return MDX_OK;
  }
  // ---- the constructor's own reset (env.py:41): block 0 for every env
  int rc = launch_reset(raw, nullptr, e->n, 0u, 0);
  if (rc != MDX_OK) {
    mdx_destroy(e.release());
    return rc;
  }
  if (cudaDeviceSynchronize() != cudaSuccess) {
    fail(MDX_E_CUDA, "reset kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
    mdx_destroy(e.release());
    return MDX_E_CUDA;
  }
  {
    // MDX_JIT=1 / 0 overrides the flag for every instance of the process (e.g. the whole parity suite on run-time-compiled kernels)
    bool want = (cfg->flags & MDX_CFG_JIT_SHAPE) != 0;
    if (const char* s = getenv("MDX_JIT")) want = atoi(s) != 0 && !raw->stale && !b.shaping_bfs;
    if (want && jit_build(raw) != MDX_OK) raw->jit_log = g_err;  // the precompiled variants run (mdx_jit_status tells why)
  }
  *out = e.release();
  return MDX_OK;
}

int32_t mdx_jit_status(const mdx_env* e, char* msg, int32_t msg_bytes) {
  if (!e) return fail(MDX_E_INVALID, "null env");
  if (msg && msg_bytes > 0) {
    const std::string m = e->jit_state > 0 ? std::string("specialised kernel loaded") : e->jit_state < 0 ? e->jit_log : std::string("not requested");
    snprintf(msg, (size_t)msg_bytes, "%s", m.c_str());
  }
  return e->jit_state;
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
  out->ring_episodes = 0;
  out->reserved = 0;
  return MDX_OK;
}

static int step_common(mdx_env* e, int mode, const uint8_t* d_actions, const int32_t* d_env_ids, int32_t n, double* d_reward,
                       float* d_shaped, uint8_t* d_done, float* d_obs, float* d_vec, float* d_gstate, float* d_gvec,
                       uint32_t flags, void* stream, double* d_agent = nullptr, int ops_as_int = 0, float* d_obs6 = nullptr,
                       float* d_g7 = nullptr, uint32_t internal_flags = 0, int32_t first_env = -1);

// MDX_CFG_STALE_TRACKER: the trainer's dict sees the reset state (MAPPO/trainer.py:232-233) -- internal flag 4 of the
// featurise-mode kernel with no outputs; internal flag 8 empties the dict instead.
static int tracker_pass(mdx_env* e, const int32_t* d_ids, int n, uint32_t internal_flags, void* stream) {
  return step_common(e, 2, nullptr, d_ids, n, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, stream, nullptr, 0, nullptr,
                     nullptr, internal_flags);
}

int32_t mdx_reset(mdx_env* e, const int32_t* h_env_ids, int32_t n, void* stream) {
  if (!e) return fail(MDX_E_INVALID, "null env");
  DeviceGuard g(e->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (h_env_ids == nullptr) {
    int rc = launch_reset(e, nullptr, e->n, 1u, st);
    if (rc == MDX_OK && e->stale) rc = tracker_pass(e, nullptr, e->n, 4u, stream);
    return rc;
  }
  if (n < 0) return fail(MDX_E_INVALID, "negative count");
  for (int k = 0; k < n; ++k)
    if (h_env_ids[k] < 0 || h_env_ids[k] >= e->n) return fail(MDX_E_INVALID, "env id %d out of range", h_env_ids[k]);
  // an env listed k times is reset k times (the reference would draw k blocks): one launch per repetition, so that no two
  // warps of a launch work on the same generator
  std::vector<int32_t> remaining(h_env_ids, h_env_ids + n);
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
    // pageable source: the copy is staged by the driver before the call returns, so `pass` may die and d_ids is reused in
    // stream order by the next pass
    CU(cudaMemcpyAsync(e->d_ids, pass.data(), sizeof(int32_t) * pass.size(), cudaMemcpyHostToDevice, st));
    int rc = launch_reset(e, e->d_ids, (int)pass.size(), 1u, st);
    if (rc == MDX_OK && e->stale) rc = tracker_pass(e, e->d_ids, (int)pass.size(), 4u, stream);
    if (rc != MDX_OK) return rc;
    remaining.swap(later);
  }
  return MDX_OK;
}

int32_t mdx_prefetch(mdx_env* e, int32_t ahead, void* stream) {
  (void)stream;
  if (!e) return fail(MDX_E_INVALID, "null env");
  if (ahead < 0) return fail(MDX_E_INVALID, "episodes_ahead must be non-negative");
  return MDX_OK;  // nothing to pre-sample: reset blocks are drawn on the device at the moment an env resets
}

static int step_common(mdx_env* e, int mode, const uint8_t* d_actions, const int32_t* d_env_ids, int32_t n, double* d_reward,
                       float* d_shaped, uint8_t* d_done, float* d_obs, float* d_vec, float* d_gstate, float* d_gvec,
                       uint32_t flags, void* stream, double* d_agent, int ops_as_int, float* d_obs6, float* d_g7, uint32_t internal_flags,
                       int32_t first_env) {
  if (!e) return fail(MDX_E_INVALID, "null env");
  if (e->stale && (d_agent || mode == 5))
    return fail(MDX_E_INVALID, "the IDQ entry points read the env truth (IDQ/trainer.py:236 clears its tracker); not available on an "
                               "MDX_CFG_STALE_TRACKER instance");
  if ((mode & 1) && !d_actions) return fail(MDX_E_INVALID, "actions are required");
  if ((mode & 1) && (e->cfg.flags & MDX_CFG_LAX_ROBOTS)) return fail(MDX_E_INVALID, "an MDX_CFG_LAX_ROBOTS instance cannot be stepped");
  if (first_env < 0) {  // (the plain entry points: all envs, or the listed ones)
    first_env = 0;
    if (d_env_ids == nullptr && n != e->n) return fail(MDX_E_INVALID, "n must equal n_envs when env ids are not given (got %d, have %d)", n, e->n);
  } else if (d_env_ids != nullptr || first_env > e->n - n) {
    return fail(MDX_E_INVALID, "bad env range [%d, %d + %d) of %d envs", first_env, first_env, n, e->n);
  }
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
  kp.env_base = first_env;
  kp.n = n;
  kp.flags = (flags & 3u) | internal_flags;
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
  kp.rows = mode == 1 ? e->rows_for_next_step : nullptr;
  kp.row_words = e->row_words, kp.row_off_robots = e->off_robots, kp.row_off_sched = e->off_sched, kp.row_off_status = e->off_status;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if ((mode == 3 || (mode == 2 && internal_flags == 0)) && e->jit_func != nullptr) return jit_launch(e, kp, st, mode);
  if (mode == 5 && e->jit_func5 != nullptr) return jit_launch(e, kp, st, 5);
  if (e->stale) return mode == 1 ? launch_env_kernel_trk_1(e, kp, st) : mode == 2 ? launch_env_kernel_trk_2(e, kp, st) : launch_env_kernel_trk_3(e, kp, st);
  if (mode == 5) return launch_env_kernel_5(e, kp, st);
  if (mode == 1) return launch_env_kernel_1(e, kp, st);
  if (mode == 2) return launch_env_kernel_2(e, kp, st);
  return launch_env_kernel_3(e, kp, st);
}

int32_t mdx_step(mdx_env* e, const uint8_t* d_actions, const int32_t* d_env_ids, int32_t n, double* d_reward, float* d_shaped,
                 uint8_t* d_done, uint32_t flags, void* stream) {
  return step_common(e, 1, d_actions, d_env_ids, n, d_reward, d_shaped, d_done, nullptr, nullptr, nullptr, nullptr, flags, stream);
}

int32_t mdx_step_n(mdx_env* e, const uint8_t* d_actions, int32_t n_steps, double* d_reward, float* d_shaped, uint8_t* d_done,
                   uint32_t flags, void* stream) {
  if (!e) return fail(MDX_E_INVALID, "null env");
  if (n_steps < 0) return fail(MDX_E_INVALID, "negative step count");
  const size_t n = (size_t)e->n;
  const size_t a_stride = n * e->R * ((flags & MDX_F_ACTIONS_I64) ? 8 : 1);
  for (int s = 0; s < n_steps; ++s) {
    int rc = step_common(e, 1, d_actions + (size_t)s * a_stride, nullptr, e->n, d_reward ? d_reward + (size_t)s * n : nullptr,
                         d_shaped ? d_shaped + (size_t)s * n : nullptr, d_done ? d_done + (size_t)s * n : nullptr, nullptr, nullptr, nullptr,
                         nullptr, flags, stream);
    if (rc != MDX_OK) return rc;
  }
  return MDX_OK;
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

int32_t mdx_step_idq_featurize_range(mdx_env* e, const uint8_t* d_actions, int32_t first_env, int32_t n, double* d_reward,
                                     double* d_agent_rewards, uint8_t* d_done, float* d_obs6, float* d_gstate7, int32_t ops_as_int, uint32_t flags,
                                     void* stream) {
  if (!d_agent_rewards) return fail(MDX_E_INVALID, "agent reward output is required");
  if (first_env < 0) return fail(MDX_E_INVALID, "negative first env");
  int rc;
  if ((rc = check_ptr16(d_obs6, "obs6")) || (rc = check_ptr16(d_gstate7, "gstate7"))) return rc;
  rc = step_common(e, 5, d_actions, nullptr, n, d_reward, nullptr, d_done, nullptr, nullptr, nullptr, nullptr, flags, stream, d_agent_rewards,
                   ops_as_int ? 1 : 0, d_obs6, d_gstate7, 0, first_env);
  if (rc == MDX_E_UNSUPPORTED_SHAPE) return fail(MDX_E_INVALID, "the fused IDQ iteration needs planes of whole float4s (H*W %% 4 == 0, <= 512 cells)");
  return rc;
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

int32_t mdx_step_range(mdx_env* e, const uint8_t* d_actions, int32_t first_env, int32_t n, double* d_reward, float* d_shaped,
                       uint8_t* d_done, uint32_t flags, void* stream) {
  if (first_env < 0) return fail(MDX_E_INVALID, "negative first env");
  return step_common(e, 1, d_actions, nullptr, n, d_reward, d_shaped, d_done, nullptr, nullptr, nullptr, nullptr, flags, stream, nullptr, 0,
                     nullptr, nullptr, 0, first_env);
}

int32_t mdx_step_featurize_range(mdx_env* e, const uint8_t* d_actions, int32_t first_env, int32_t n, double* d_reward, float* d_shaped,
                                 uint8_t* d_done, float* d_obs, float* d_vec, float* d_gstate, float* d_gvec, uint32_t flags, void* stream) {
  if (first_env < 0) return fail(MDX_E_INVALID, "negative first env");
  return step_common(e, 3, d_actions, nullptr, n, d_reward, d_shaped, d_done, d_obs, d_vec, d_gstate, d_gvec, flags, stream, nullptr, 0,
                     nullptr, nullptr, 0, first_env);
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
  if (e->stale && d_agent) return fail(MDX_E_INVALID, "IDQ reward_shaping reads the env truth; not available on an MDX_CFG_STALE_TRACKER instance");
  const int spw = e->stale ? e->base.smem_per_warp : e->base.off_alist;
  const int wpb = pick_wpb(e, spw);
  kp.smem_per_warp = spw;
  const size_t smem = (size_t)kLutBytes + (size_t)wpb * spw;
  const void* fn = e->stale ? reinterpret_cast<const void*>(mdx_shape_kernel<true>) : reinterpret_cast<const void*>(mdx_shape_kernel<false>);
  {
    int rc = ensure_dynamic_smem(fn, smem, -1);
    if (rc != MDX_OK) return rc;
  }
  const unsigned grid = (unsigned)((e->n + wpb - 1) / wpb);
  if (e->stale) mdx_shape_kernel<true><<<grid, wpb * 32, smem, static_cast<cudaStream_t>(stream)>>>(kp);
  else mdx_shape_kernel<false><<<grid, wpb * 32, smem, static_cast<cudaStream_t>(stream)>>>(kp);
  CU(cudaGetLastError());
  return MDX_OK;
}

int64_t mdx_state_row_bytes(const mdx_env* e) { return e ? (int64_t)e->row_words * 4 : 0; }

// Small transfers skip the copy engine altogether: the pinned staging buffer is host memory the device can address directly
// (unified addressing), so the kernels read their inputs from it and write their results into it over PCIe, and the call is one or
// two launches plus one synchronisation.  Larger ones go through the device mirror with one copy each way.
constexpr size_t kZeroCopyBytes = 32 * 1024;

int32_t mdx_get_states(mdx_env* e, const int32_t* h_env_ids, int32_t n, void* h_rows, void* stream) {
  if (!e || !h_rows) return fail(MDX_E_INVALID, "null argument");
  int rc = check_ids(e, h_env_ids, n);
  if (rc != MDX_OK) return rc;
  if (n == 0) return MDX_OK;
  DeviceGuard g(e->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t ids_bytes = h_env_ids ? up16(sizeof(int32_t) * (size_t)n) : 0, rows_bytes = (size_t)n * e->row_words * 4;
  if ((rc = staging(e, ids_bytes + rows_bytes)) != MDX_OK) return rc;
  const bool zc = ids_bytes + rows_bytes <= kZeroCopyBytes;
  unsigned char* dv = zc ? e->h_pin : e->d_pin;  // what the kernels address
  RowArgs a = row_args(e, reinterpret_cast<uint32_t*>(dv + ids_bytes), nullptr, n);
  if (h_env_ids != nullptr) {
    if (n == 1) {
      a.base = h_env_ids[0];  // one env: its id travels in the kernel parameters
    } else {
      std::memcpy(e->h_pin, h_env_ids, sizeof(int32_t) * (size_t)n);
      if (!zc) CU(cudaMemcpyAsync(e->d_pin, e->h_pin, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, st));
      a.ids = reinterpret_cast<const int32_t*>(dv);
    }
  }
  mdx_gather_rows_kernel<<<(unsigned)((n + 3) / 4), 128, 0, st>>>(e->base, a);
  CU(cudaGetLastError());
  if (!zc) CU(cudaMemcpyAsync(e->h_pin + ids_bytes, e->d_pin + ids_bytes, rows_bytes, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  std::memcpy(h_rows, e->h_pin + ids_bytes, rows_bytes);
  return MDX_OK;
}

int32_t mdx_set_states(mdx_env* e, const int32_t* h_env_ids, int32_t n, const void* h_rows, void* stream) {
  if (!e || !h_rows) return fail(MDX_E_INVALID, "null argument");
  int rc = check_ids(e, h_env_ids, n, true);
  if (rc != MDX_OK) return rc;
  if (n == 0) return MDX_OK;
  const uint32_t* rows = static_cast<const uint32_t*>(h_rows);
  for (int k = 0; k < n; ++k)
    if ((rc = validate_row(e, h_env_ids ? h_env_ids[k] : k, rows + (size_t)k * e->row_words)) != MDX_OK) return rc;
  DeviceGuard g(e->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t ids_bytes = h_env_ids ? up16(sizeof(int32_t) * (size_t)n) : 0, rows_bytes = (size_t)n * e->row_words * 4;
  if ((rc = staging(e, ids_bytes + rows_bytes)) != MDX_OK) return rc;
  const bool zc = ids_bytes + rows_bytes <= kZeroCopyBytes;
  unsigned char* dv = zc ? e->h_pin : e->d_pin;
  if (h_env_ids) std::memcpy(e->h_pin, h_env_ids, sizeof(int32_t) * (size_t)n);
  std::memcpy(e->h_pin + ids_bytes, h_rows, rows_bytes);
  if (!zc) CU(cudaMemcpyAsync(e->d_pin, e->h_pin, ids_bytes + rows_bytes, cudaMemcpyHostToDevice, st));  // ONE copy: ids + rows
  RowArgs a = row_args(e, reinterpret_cast<uint32_t*>(dv + ids_bytes), h_env_ids ? reinterpret_cast<const int32_t*>(dv) : nullptr, n);
  mdx_scatter_rows_kernel<<<(unsigned)((n + 3) / 4), 128, 0, st>>>(e->base, a);
  CU(cudaGetLastError());
  CU(cudaEventRecord(e->pin_busy, st));  // the host buffer is read until here (by the copy, or by the kernel itself)
  return MDX_OK;
}

int32_t mdx_shaped_reward_host(mdx_env* e, const uint32_t* h_cur_robots, const int32_t* h_cur_t, const uint8_t* h_actions,
                               const double* h_global_reward, float* h_out, void* stream) {
  if (!e || !h_cur_robots || !h_cur_t || !h_actions || !h_global_reward || !h_out) return fail(MDX_E_INVALID, "null argument");
  DeviceGuard g(e->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t N = (size_t)e->n;
  const size_t o_rob = 0, o_t = o_rob + up16(4 * N * e->R), o_act = o_t + up16(4 * N), o_gr = o_act + up16(N * e->R), o_out = o_gr + up16(8 * N);
  const size_t total = o_out + up16(4 * N);
  int rc = staging(e, total);
  if (rc != MDX_OK) return rc;
  const bool zc = total <= kZeroCopyBytes;
  unsigned char* dv = zc ? e->h_pin : e->d_pin;
  std::memcpy(e->h_pin + o_rob, h_cur_robots, 4 * N * e->R);
  std::memcpy(e->h_pin + o_t, h_cur_t, 4 * N);
  std::memcpy(e->h_pin + o_act, h_actions, N * e->R);
  std::memcpy(e->h_pin + o_gr, h_global_reward, 8 * N);
  if (!zc) CU(cudaMemcpyAsync(e->d_pin, e->h_pin, o_out, cudaMemcpyHostToDevice, st));
  rc = mdx_shaped_reward(e, reinterpret_cast<const uint32_t*>(dv + o_rob), reinterpret_cast<const int32_t*>(dv + o_t), dv + o_act,
                         reinterpret_cast<const double*>(dv + o_gr), reinterpret_cast<float*>(dv + o_out), stream);
  if (rc != MDX_OK) return rc;
  if (!zc) CU(cudaMemcpyAsync(e->h_pin + o_out, e->d_pin + o_out, 4 * N, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  std::memcpy(h_out, e->h_pin + o_out, 4 * N);
  return MDX_OK;
}

int32_t mdx_featurize_host(mdx_env* e, const int32_t* h_env_ids, int32_t n, float* h_obs, float* h_vec, float* h_gstate, float* h_gvec,
                           void* stream) {
  if (!e) return fail(MDX_E_INVALID, "null env");
  int rc = check_ids(e, h_env_ids, n);
  if (rc != MDX_OK) return rc;
  if (n == 0) return MDX_OK;
  DeviceGuard g(e->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const KParams& b = e->base;
  const size_t N = (size_t)n;
  const size_t ids_bytes = h_env_ids ? up16(4 * N) : 0;
  const size_t b_obs = h_obs ? up16(4 * N * e->R * 6 * b.HW) : 0, b_vec = h_vec ? up16(4 * N * e->R * b.D) : 0;
  const size_t b_gs = h_gstate ? up16(4 * N * 4 * b.HW) : 0, b_gv = h_gvec ? up16(4 * N * b.Dg) : 0;
  const size_t o_obs = ids_bytes, o_vec = o_obs + b_obs, o_gs = o_vec + b_vec, o_gv = o_gs + b_gs, total = o_gv + b_gv;
  if ((rc = staging(e, total)) != MDX_OK) return rc;
  const bool zc = total <= kZeroCopyBytes;
  unsigned char* dv = zc ? e->h_pin : e->d_pin;
  if (h_env_ids) {
    std::memcpy(e->h_pin, h_env_ids, 4 * N);
    if (!zc) CU(cudaMemcpyAsync(e->d_pin, e->h_pin, 4 * N, cudaMemcpyHostToDevice, st));
  }
  rc = mdx_featurize(e, h_env_ids ? reinterpret_cast<const int32_t*>(dv) : nullptr, n, h_obs ? reinterpret_cast<float*>(dv + o_obs) : nullptr,
                     h_vec ? reinterpret_cast<float*>(dv + o_vec) : nullptr, h_gstate ? reinterpret_cast<float*>(dv + o_gs) : nullptr,
                     h_gvec ? reinterpret_cast<float*>(dv + o_gv) : nullptr, stream);
  if (rc != MDX_OK) return rc;
  if (!zc) CU(cudaMemcpyAsync(e->h_pin + o_obs, e->d_pin + o_obs, total - o_obs, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if (h_obs) std::memcpy(h_obs, e->h_pin + o_obs, 4 * N * e->R * 6 * b.HW);
  if (h_vec) std::memcpy(h_vec, e->h_pin + o_vec, 4 * N * e->R * b.D);
  if (h_gstate) std::memcpy(h_gstate, e->h_pin + o_gs, 4 * N * 4 * b.HW);
  if (h_gvec) std::memcpy(h_gvec, e->h_pin + o_gv, 4 * N * b.Dg);
  return MDX_OK;
}

int32_t mdx_step_host(mdx_env* e, const uint8_t* h_actions, const int32_t* h_env_ids, int32_t n, double* h_reward, float* h_shaped,
                      uint8_t* h_done, void* h_rows, uint32_t flags, void* stream) {
  if (!e || !h_actions) return fail(MDX_E_INVALID, "null argument");
  int rc = check_ids(e, h_env_ids, n, true);
  if (rc != MDX_OK) return rc;
  if (n == 0) return MDX_OK;
  if (flags & MDX_F_ACTIONS_I64) return fail(MDX_E_INVALID, "mdx_step_host takes uint8 action codes");
  DeviceGuard g(e->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t N = (size_t)n;
  // in:  [ids][actions]      out: [reward f64][shaped f32][done u8][rows]
  const size_t in_ids = h_env_ids ? up16(4 * N) : 0, in_bytes = in_ids + up16(N * e->R);
  const size_t o_rew = in_bytes, o_sh = o_rew + up16(8 * N), o_dn = o_sh + up16(4 * N), o_rows = o_dn + up16(N);
  const size_t total = o_rows + (h_rows ? N * e->row_words * 4 : 0);
  if ((rc = staging(e, total)) != MDX_OK) return rc;
  const bool zc = total <= kZeroCopyBytes;
  unsigned char* dv = zc ? e->h_pin : e->d_pin;  // what the kernels address
  if (h_env_ids) std::memcpy(e->h_pin, h_env_ids, 4 * N);
  std::memcpy(e->h_pin + in_ids, h_actions, N * e->R);
  if (!zc) CU(cudaMemcpyAsync(e->d_pin, e->h_pin, in_bytes, cudaMemcpyHostToDevice, st));
  const int32_t* d_ids = h_env_ids ? reinterpret_cast<const int32_t*>(dv) : nullptr;
  // the step kernel writes the packed rows of the stepped envs itself (no second launch)
  const bool fused_rows = h_rows != nullptr;
  e->rows_for_next_step = fused_rows ? reinterpret_cast<uint32_t*>(dv + o_rows) : nullptr;
  rc = mdx_step(e, dv + in_ids, d_ids, n, reinterpret_cast<double*>(dv + o_rew), h_shaped ? reinterpret_cast<float*>(dv + o_sh) : nullptr,
                dv + o_dn, flags & MDX_F_AUTO_RESET, stream);
  e->rows_for_next_step = nullptr;
  if (rc != MDX_OK) return rc;
  if (!zc) CU(cudaMemcpyAsync(e->h_pin + o_rew, e->d_pin + o_rew, total - o_rew, cudaMemcpyDeviceToHost, st));  // ONE copy back
  CU(cudaStreamSynchronize(st));
  if (h_reward) std::memcpy(h_reward, e->h_pin + o_rew, 8 * N);
  if (h_shaped) std::memcpy(h_shaped, e->h_pin + o_sh, 4 * N);
  if (h_done) std::memcpy(h_done, e->h_pin + o_dn, N);
  if (h_rows) std::memcpy(h_rows, e->h_pin + o_rows, N * e->row_words * 4);
  return MDX_OK;
}

int32_t mdx_get_state(mdx_env* e, int32_t id, mdx_state_view* out, void* stream) {
  if (!e || !out) return fail(MDX_E_INVALID, "null argument");
  if (id < 0 || id >= e->n) return fail(MDX_E_INVALID, "env id %d out of range", id);
  const KParams& b = e->base;
  std::vector<uint32_t> row((size_t)e->row_words);
  int rc = mdx_get_states(e, &id, 1, row.data(), stream);
  if (rc != MDX_OK) return rc;
  out->t = (int32_t)row[0];
  out->episode = (int32_t)row[1];
  out->map_id = e->env_map[(size_t)id];
  std::memcpy(&out->total_reward, &row[4], sizeof(double));
  const uint16_t* sched = reinterpret_cast<const uint16_t*>(row.data() + e->off_sched);
  const uint8_t* status = reinterpret_cast<const uint8_t*>(row.data() + e->off_status);
  const int Wo = e->W_out;
  for (int i = 0; i < e->R; ++i) {
    const uint32_t rob = row[(size_t)e->off_robots + i];
    const int cell = (int)(rob & 0xffffu);
    if (out->robot_row) out->robot_row[i] = cell / Wo;
    if (out->robot_col) out->robot_col[i] = cell % Wo;
    if (out->robot_carry) out->robot_carry[i] = (int)(rob >> 16);
  }
  for (int j = 0; j < e->P; ++j) {
    const int s = sched[0 * b.Pp8 + j], g2 = sched[1 * b.Pp8 + j];
    if (out->pkg_start_row) out->pkg_start_row[j] = s / Wo;
    if (out->pkg_start_col) out->pkg_start_col[j] = s % Wo;
    if (out->pkg_target_row) out->pkg_target_row[j] = g2 / Wo;
    if (out->pkg_target_col) out->pkg_target_col[j] = g2 % Wo;
    if (out->pkg_start_time) out->pkg_start_time[j] = sched[2 * b.Pp8 + j];
    if (out->pkg_deadline) out->pkg_deadline[j] = sched[3 * b.Pp8 + j];
    if (out->pkg_status) out->pkg_status[j] = status[j];
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
  const KParams& b = e->base;
  const MapHost& m = e->maps[(size_t)e->env_map[(size_t)id]];
  const int Wo = e->W_out;
  std::vector<uint32_t> row((size_t)e->row_words, 0u);
  row[0] = (uint32_t)in->t;
  std::memcpy(&row[4], &in->total_reward, sizeof(double));
  for (int i = 0; i < e->R; ++i) {
    const int r = in->robot_row[i], c = in->robot_col[i], k = in->robot_carry[i];
    if (r < 0 || r >= m.rows || c < 0 || c >= m.cols) return fail(MDX_E_INVALID, "robot %d outside the map", i);
    if (k < 0 || k > e->P) return fail(MDX_E_INVALID, "robot %d carries an unknown package id %d", i, k);
    row[(size_t)e->off_robots + i] = (uint32_t)(r * Wo + c) | ((uint32_t)k << 16);
  }
  uint16_t* sched = reinterpret_cast<uint16_t*>(row.data() + e->off_sched);
  uint8_t* status = reinterpret_cast<uint8_t*>(row.data() + e->off_status);
  for (int j = b.P; j < b.Pp8; ++j) sched[2 * b.Pp8 + j] = 0xffff;
  for (int j = 0; j < e->P; ++j) {
    const int sr = in->pkg_start_row[j], sc = in->pkg_start_col[j], tr = in->pkg_target_row[j], tc = in->pkg_target_col[j];
    if (sr < 0 || sr >= m.rows || sc < 0 || sc >= m.cols || tr < 0 || tr >= m.rows || tc < 0 || tc >= m.cols)
      return fail(MDX_E_INVALID, "package %d outside the map", j + 1);
    if (in->pkg_start_time[j] < 0 || in->pkg_start_time[j] > 65534) return fail(MDX_E_INVALID, "package %d: bad start_time", j + 1);
    if (in->pkg_deadline[j] < 0 || in->pkg_deadline[j] > 65535) return fail(MDX_E_INVALID, "package %d: deadline out of range", j + 1);
    if (in->pkg_status[j] < 0 || in->pkg_status[j] > 3) return fail(MDX_E_INVALID, "package %d: bad status", j + 1);
    sched[0 * b.Pp8 + j] = (uint16_t)(sr * Wo + sc);
    sched[1 * b.Pp8 + j] = (uint16_t)(tr * Wo + tc);
    sched[2 * b.Pp8 + j] = (uint16_t)in->pkg_start_time[j];
    sched[3 * b.Pp8 + j] = (uint16_t)in->pkg_deadline[j];
    status[j] = (uint8_t)in->pkg_status[j];
  }
  return mdx_set_states(e, &id, 1, row.data(), stream);  // (validates time, free / distinct robot cells, deadline range)
}

int32_t mdx_tracker_clear(mdx_env* e, const int32_t* h_env_ids, int32_t n, void* stream) {
  if (!e) return fail(MDX_E_INVALID, "null env");
  if (!e->stale) return fail(MDX_E_INVALID, "not an MDX_CFG_STALE_TRACKER instance");
  DeviceGuard g(e->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (h_env_ids == nullptr) {
    CU(cudaMemsetAsync(e->d_trk_n, 0, sizeof(int32_t) * (size_t)e->n, st));
    return MDX_OK;
  }
  int rc = check_ids(e, h_env_ids, n, true);
  if (rc != MDX_OK || n == 0) return rc;
  CU(cudaMemcpyAsync(e->d_ids, h_env_ids, sizeof(int32_t) * (size_t)n, cudaMemcpyHostToDevice, st));  // pageable source: staged before return
  return tracker_pass(e, e->d_ids, n, 8u, stream);
}

int32_t mdx_tracker_get(mdx_env* e, int32_t id, int32_t* h_n, uint16_t* h_rows5, uint8_t* h_status, void* stream) {
  if (!e || !h_n || !h_rows5 || !h_status) return fail(MDX_E_INVALID, "null argument");
  if (!e->stale) return fail(MDX_E_INVALID, "not an MDX_CFG_STALE_TRACKER instance");
  if (id < 0 || id >= e->n) return fail(MDX_E_INVALID, "env id %d out of range", id);
  DeviceGuard g(e->device);
  const KParams& b = e->base;
  CU(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  CU(cudaMemcpy(h_n, e->d_trk_n + id, sizeof(int32_t), cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(h_rows5, e->d_trk + (size_t)id * 5 * b.Pp8, sizeof(uint16_t) * 5 * b.Pp8, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(h_status, e->d_trk_status + (size_t)id * b.Pp16, (size_t)b.Pp16, cudaMemcpyDeviceToHost));
  return MDX_OK;
}

int32_t mdx_tracker_set(mdx_env* e, int32_t id, int32_t n, const uint16_t* h_rows5, const uint8_t* h_status, void* stream) {
  if (!e || !h_rows5 || !h_status) return fail(MDX_E_INVALID, "null argument");
  if (!e->stale) return fail(MDX_E_INVALID, "not an MDX_CFG_STALE_TRACKER instance");
  if (id < 0 || id >= e->n) return fail(MDX_E_INVALID, "env id %d out of range", id);
  if (n < 0 || n > e->P) return fail(MDX_E_INVALID, "a tracker holds at most n_packages entries");
  const KParams& b = e->base;
  const MapHost& m = e->maps[(size_t)e->env_map[(size_t)id]];
  const int Wo = e->W_out;
  const long long dl_cap = std::min<long long>(65535, ((1LL << (32 - b.key_wbits)) - 2) / b.nranks_max - 1);
  std::vector<char> seen((size_t)e->P + 1, 0);
  for (int k = 0; k < n; ++k) {
    const int pid = h_rows5[k], s = h_rows5[b.Pp8 + k], t = h_rows5[2 * b.Pp8 + k];
    if (pid < 1 || pid > e->P || seen[(size_t)pid]) return fail(MDX_E_INVALID, "tracker entry %d: package ids must be distinct and in 1..n_packages", k);
    seen[(size_t)pid] = 1;
    if (s / Wo >= m.rows || s % Wo >= m.cols || t / Wo >= m.rows || t % Wo >= m.cols) return fail(MDX_E_INVALID, "tracker entry %d outside the map", k);
    if ((long long)h_rows5[4 * b.Pp8 + k] > dl_cap) return fail(MDX_E_INVALID, "tracker entry %d: deadline out of range", k);
    if (h_status[k] != ST_WAITING && h_status[k] != ST_IN_TRANSIT) return fail(MDX_E_INVALID, "tracker entry %d: status must be waiting or in_transit", k);
  }
  DeviceGuard g(e->device);
  CU(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  CU(cudaMemcpy(e->d_trk_n + id, &n, sizeof(int32_t), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(e->d_trk + (size_t)id * 5 * b.Pp8, h_rows5, sizeof(uint16_t) * 5 * b.Pp8, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(e->d_trk_status + (size_t)id * b.Pp16, h_status, (size_t)b.Pp16, cudaMemcpyHostToDevice));
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
