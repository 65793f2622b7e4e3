// Host-side state of one mdx_env and the helpers shared by the translation units of libmdx.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <string>
#include <thread>
#include <vector>

#include "../../include/mdx.h"
#include "mdx_internal.h"

namespace mdx {

// records the calling thread's error message (mdx_last_error) and returns `code`
int fail(int code, const char* fmt, ...);

#define CU(call)                                                                                  \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess) return ::mdx::fail(MDX_E_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

struct MapHost {
  int rows = 0, cols = 0;
  std::vector<uint8_t> cells;
  std::vector<int> free_cells;
};

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
// multiplier for x / d as __umulhi(x, magic(d)); d == 1 is encoded as 0 (div_magic in the kernels), d >= 2 is exact for x < 2^32 / d
inline uint32_t magic(uint32_t d) { return d <= 1 ? 0u : (uint32_t)((0x100000000ull + d - 1) / d); }

template <typename F>
void parallel_for(int n, int n_threads, F&& fn) {
  if (n <= 0) return;
  n_threads = std::max(1, std::min(n_threads, n / 64 + 1));
  if (n_threads == 1) {
    fn(0, 0, n);
    return;
  }
  std::vector<std::thread> th;
  const int chunk = (n + n_threads - 1) / n_threads;
  for (int k = 0; k < n_threads; ++k) {
    const int lo = k * chunk, hi = std::min(n, lo + chunk);
    if (lo >= hi) break;
    th.emplace_back([&fn, k, lo, hi] { fn(k, lo, hi); });
  }
  for (auto& t : th) t.join();
}


}  // namespace mdx

struct mdx_env {
  using MapHost = mdx::MapHost;
  using Mt19937 = mdx::Mt19937;
  using KParams = mdx::KParams;
  using MapDev = mdx::MapDev;
  mdx_config cfg{};
  int device = 0;
  int n = 0, R = 0, P = 0, T = 0, E = 0;
  int H_out = 0, W_out = 0;
  std::vector<MapHost> maps;
  std::vector<int32_t> env_map;
  KParams base{};                  // everything but the per-call arguments
  // device memory
  void* d_blob = nullptr;          // maps + LUTs
  MapDev* d_maps = nullptr;
  int32_t* d_t = nullptr;
  int32_t* d_ep = nullptr;
  uint32_t* d_mt = nullptr;        // [n][624] MT19937 state per env (the device reset sampler's streams)
  int32_t* d_mtpos = nullptr;      // [n]
  double* d_total = nullptr;
  uint32_t* d_robots = nullptr;
  uint8_t* d_status = nullptr;
  uint16_t* d_cur = nullptr;
  uint32_t* d_counters = nullptr;
  // MDX_CFG_JIT_SHAPE: the fused kernel compiled at run time for this instance's shape (mdx_jit.cpp)
  void* jit_module = nullptr;      // CUmodule
  void* jit_func = nullptr;        // CUfunction: fused step + featurise (mode 3)
  void* jit_func2 = nullptr;       // CUfunction: featurise only (mode 2)
  void* jit_func5 = nullptr;       // CUfunction: fused IDQ iteration (mode 5; plane-sliced maps only)
  int jit_state = 0;               // 0 not tried, 1 loaded, -1 unavailable (the precompiled variants run)
  int jit_blocks = 0;
  std::string jit_log;
  bool stale = false;              // MDX_CFG_STALE_TRACKER
  uint16_t* d_trk = nullptr;       // [n][5][Pp8]
  uint8_t* d_trk_status = nullptr; // [n][Pp16]
  int32_t* d_trk_n = nullptr;      // [n]
  int32_t* d_env_map = nullptr;
  int32_t* d_ids = nullptr;     // [n] env ids of an mdx_reset call
  // packed-row entry points (mdx_get_states / mdx_set_states / mdx_step_host): one pinned host buffer and its device
  // mirror, grown on demand; `pin_busy` marks the last asynchronous H2D copy that read the host buffer
  unsigned char* h_pin = nullptr;
  unsigned char* d_pin = nullptr;
  size_t pin_bytes = 0;
  cudaEvent_t pin_busy = nullptr;
  int row_words = 0, off_robots = 0, off_sched = 0, off_status = 0;  // row layout in 4-byte words
  uint32_t* rows_for_next_step = nullptr;  // set by mdx_step_host around its mdx_step call: the step kernel writes the packed rows itself
  // f2: greedy policy + BFS tables (allocated on first use)
  uint16_t* d_bfs = nullptr;       // [n_maps][HW][HW]
  uint32_t* d_greedy = nullptr;    // [n][32]  target id | carried id seen at the previous call << 16
  uint32_t* d_gfree = nullptr;     // [n][gfree_words] packages_free bits over the agent's own list
  int gfree_words = 0;
  int pdl = 1;            // MDX_PDL: 0 never, 1 programmatic dependent launch for step-only launches (default), 2 for every mode
  int wpb_override = 0;
  int carveout = -1;      // MDX_CARVEOUT=0..100: preferred shared-memory carve-out of the env kernels (-1 = driver default)
  int no_fixed = 0;       // MDX_NOFIX=1: never use the shape-specialised instantiation
  int no_fat = 0;         // MDX_NOFAT=1: never use the uncapped-register instantiation for big environments
  int lpe_override = 0;   // MDX_LPE=8|16|32: minimum lanes per environment (experiments / tests)
  int force_generic = 0;  // MDX_GENERIC=1: always take the generic featuriser path (tests)
};


namespace mdx {
// The one shape with a BAKED specialised variant (kFix = 1 of mdx_kernels.cuh: the C3 benchmark / MAPPO trainer call site -- map1 10 x 10,
// 5 robots, <= 56 packages, K_o = 4, K_p = 5, MR = MP = 100) incl. the shared-memory carve-up that variant hard-codes.
inline bool baked_c3_shape(const mdx_env* e) {
  const KParams& b = e->base;
  return !e->no_fixed && !b.shaping_bfs && b.R == 5 && b.Pp8 == 56 && b.Pp16 == 64 && b.H_out == 10 && b.W_out == 10 && b.K_o == 4 && b.K_p == 5 &&
         b.sel_k == 5 && b.MR == 100 && b.MP == 100 && b.vec_rows_per_flush == 5 && b.stage_b == 44 && b.off_sched == 0 && b.off_status == 448 &&
         b.off_wlist == 512 && b.off_wkey == 848 && b.off_rinfo == 1072 && b.off_alist == 1392 && b.off_sel == 1504 && b.off_masks == 1568 &&
         b.off_stage == 1728 && b.key_wbits == 6;
}
// cudaFuncAttributeMaxDynamicSharedMemorySize is a property of the kernel FUNCTION on a device, shared by every mdx_env of the
// process: raise it to `bytes` if the high-water mark recorded for (function, current device) is lower.
int ensure_dynamic_smem(const void* func, size_t bytes, int carveout);
// one translation unit per kernel mode (mdx_launch{1,2,3}.cu): 1 step, 2 featurise, 3 fused
int launch_env_kernel_1(mdx_env* e, const KParams& kp, cudaStream_t st);
int launch_env_kernel_2(mdx_env* e, const KParams& kp, cudaStream_t st);
int launch_env_kernel_3(mdx_env* e, const KParams& kp, cudaStream_t st);
int launch_env_kernel_trk_1(mdx_env* e, const KParams& kp, cudaStream_t st);  // MDX_CFG_STALE_TRACKER instances
int launch_env_kernel_trk_2(mdx_env* e, const KParams& kp, cudaStream_t st);
int launch_env_kernel_trk_3(mdx_env* e, const KParams& kp, cudaStream_t st);
int launch_env_kernel_5(mdx_env* e, const KParams& kp, cudaStream_t st);  // step + IDQ featurisers; MDX_E_UNSUPPORTED_SHAPE: use two launches
constexpr int MDX_E_UNSUPPORTED_SHAPE = -100;  // internal: never crosses the C ABI
// run-time shape specialisation (mdx_jit.cpp)
int jit_build(mdx_env* e);
void jit_release(mdx_env* e);
int jit_launch(mdx_env* e, const KParams& kp, cudaStream_t st, int mode);
// builds the all-pairs BFS tables and publishes them in the map descriptors (idempotent)
int ensure_bfs(mdx_env* e);
}  // namespace mdx
