// Internal declarations shared by the host sampler, the C-ABI glue and the kernels.
#pragma once
#ifdef __CUDACC_RTC__
// run-time compilation of a shape-specialised kernel (mdx_jit.cpp): NVRTC has no host headers
typedef unsigned char uint8_t;
typedef unsigned short uint16_t;
typedef unsigned int uint32_t;
typedef int int32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
typedef unsigned long long uintptr_t;
#else
#include <cstdint>
#include <vector>
#endif

namespace mdx {

#ifndef __CUDACC_RTC__
struct Mt19937 {
  uint32_t mt[624];
  int pos;
  void seed(uint32_t s);
  void refill();
  inline uint32_t next32() {
    if (pos == 624) refill();
    uint32_t y = mt[pos++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
  }
  bool randint(int64_t lo, int64_t hi, int64_t* out);
};

struct EpisodeSampler {
  struct Rec {
    int start, target, start_time, deadline;
  };
  Mt19937 rng;
  int rows = 0, cols = 0, R = 0, P = 0, T = 0;
  std::vector<int> free_cells;
  std::vector<int> scratch_avail;
  std::vector<Rec> scratch_rec;
  bool init(uint32_t seed, const uint8_t* cells, int rows, int cols, int R, int P, int T);
  bool next(int32_t* robot_cell, int32_t* pkg4);
  bool draw(Mt19937& g, const std::vector<int>& free_cells, int rows, int R, int P, int T, int32_t* robot_cell,
            int32_t* pkg4);
};

void build_quotient_lut(int n, int n_out, float* out);
int build_rank_lut(int rows, int cols, int rows_out, int cols_out, uint16_t* out);
float np_sum_f32(const float* a, int n);
#endif  // !__CUDACC_RTC__

// ---- device-side views ------------------------------------------------------------------------

enum : uint8_t { ST_NONE = 0, ST_WAITING = 1, ST_IN_TRANSIT = 2, ST_DELIVERED = 3 };
enum : int { MV_D = 0, MV_L = 1, MV_R = 2, MV_S = 3, MV_U = 4, MV_BAD = 5 };

// counters[env][8], cumulative since creation
enum : int { C_EPISODES = 0, C_MOVES = 1, C_ONTIME = 2, C_LATE = 3, C_PICKUPS = 4, C_STEPS = 5, C_EPLEN = 6, C_STARVED = 7 };

struct MapDev {
  int32_t H, W;               // true map size (rows, cols)
  int32_t nranks;
  int32_t pad;
  const uint32_t* grid_bits;  // [W32] obstacle bits in output-plane cell coordinates
  const float* qr;            // [2*H_out-1]  qr[d + H_out-1] = (float)((double)d / H)
  const float* qc;            // [2*W_out-1]
  const uint16_t* rank;       // [(2*H_out-1) * (2*W_out-1)] dense rank of the float64 sort key
  const uint16_t* bfs;        // [HW][HW] all-pairs shortest-path length over free cells (output-plane cell ids,
                              // 0xFFFF = unreachable / obstacle); nullptr until the greedy policy or BFS shaping needs it
  const uint16_t* free_cells; // [n_free] output-plane cell ids of the free cells in the reference's row-major order
                              // (env.py:155,166): what the device reset sampler indexes with its draws
  int32_t n_free;
  int32_t pad2;
};

struct KParams {
  // shape
  int32_t n_envs, R, P, T;
  int32_t Pp8;    // P rounded up to 8  (u16 rows of the schedule block are 16-byte aligned)
  int32_t Pp16;   // P rounded up to 16 (status row)
  int32_t H_out, W_out, HW, W32;
  int32_t K_o, K_p, MR, MP, D, Dg;
  int32_t E;      // (unused since v2: episodes are drawn on the device when an env resets)
  int32_t key_wbits;     // bits reserved for the waiting-list index inside the 32-bit selection key
  int32_t nranks_max;
  int32_t tq_len;
  int32_t vec_rows_per_flush;
  int32_t gvec_chunk;    // package rows of gvec staged per flush
  uint32_t mul_plane;    // ceil(2^32 / HW)       : e / HW        for e < 2^32 / HW
  uint32_t mul_wout;     // ceil(2^32 / W_out)    : cell / W_out
  uint32_t mul_r;        // ceil(2^32 / R)        : lane / R
  uint32_t mul_ksel;     // ceil(2^32 / sel_k)
  int32_t pair_g;        // 32 / R: robots ranked per pass of the (robot, other) pair layout
  int32_t sel_rp;        // 32 / sel_k: robots whose package selection fits one pass of lane-held winners
  // smem carve-up (bytes from the warp's base)
  // step-only launches need [0, off_alist); the featuriser adds alist, sel, planes and the stage
  int32_t off_sched, off_status, off_wlist, off_wkey, off_rinfo, off_alist, off_sel, off_masks, off_stage;
  int32_t off_trk;        // stale-tracker instances: tracker table (ids, start, target, start_time, deadline: u16 [5][Pp8]; status u8
                          // [Pp16]; id -> slot map u16 [Pp8]) behind the stage
  int32_t smem_per_warp;
  int32_t sel_k;      // min(K_p, P): selection slots per robot
  int32_t stage_b;    // float offset of the gvec package segment inside the stage
  // rewards
  double move_cost, delivery_reward, delay_reward;
  float shaping[9];
  int32_t shaping_bfs;     // 1: shaped distance terms use the BFS table instead of Manhattan distance (opt-in)
  // state
  int32_t* t;
  int32_t* ep;             // [n_envs] resets applied since construction (= index of the env's current draw block)
  uint32_t* mt;            // [n_envs][624] MT19937 state of the env's numpy legacy RandomState stream (env.py:40)
  int32_t* mtpos;          // [n_envs] next output index, 0..624
  double* total_reward;
  uint32_t* robots;        // [n_envs][R]  cell | carry << 16
  uint8_t* status;         // [n_envs][Pp16]
  uint16_t* cur;           // [n_envs][4][Pp8]      schedule of the env's CURRENT episode: start cell, target cell, start_time,
                           //                       deadline in package-id order; rewritten by the device sampler on every reset
  uint32_t* counters;      // [n_envs][8]
  // MDX_CFG_STALE_TRACKER: the MAPPO trainer's package dict per env, in dict (insertion) order (MAPPO/trainer.py:91-130)
  uint16_t* trk;           // [n_envs][5][Pp8]  package id (1-based), start cell, target cell, start_time, deadline
  uint8_t* trk_status;     // [n_envs][Pp16]    ST_WAITING / ST_IN_TRANSIT
  int32_t* trk_n;          // [n_envs]          entries in the dict
  const int32_t* env_map;  // [n_envs]
  const MapDev* maps;
  MapDev map0;             // copy of maps[0] (used directly when every env shares one map)
  const float* tq;         // tq[k] = (float)((double)k / T)
  const float* cq;         // cq[i] = (float)((double)i / (MR - 1)), i < 32 (MAPPO/helper.py:247)
  // call arguments
  const uint8_t* actions;
  const int32_t* env_ids;
  int32_t env_base;         // env_ids == nullptr: launch row i works on env env_base + i (contiguous range)
  int32_t n;
  uint32_t flags;
  double* out_reward;
  float* out_shaped;
  uint8_t* out_done;
  float* obs;
  float* vec;
  float* gstate;
  float* gvec;
  double* out_agent;        // [n, R] IDQ per-agent reward_shaping (IDQ/networks.py:228-349), nullptr = off
  int32_t idq_ops_int;      // 0: package ops arrive as strings like in IDQ/trainer.py (only the stay penalty can fire); 1: ints
  int32_t pad1;
  float* alt_obs6;          // [n, R, 6, H, W] IDQ convert_state, fused behind the step (kernel mode 5), nullptr = off
  float* alt_g7;            // [n, 7, H, W] legacy global tensor
  // step-only launches of mdx_step_host: the stepped envs' packed state rows (mdx_get_states layout) written by the step kernel itself
  uint32_t* rows;           // [n][row_words], nullptr = off
  int32_t row_words, row_off_robots, row_off_sched, row_off_status;
  // shaped-reward-only call
  const uint32_t* cur_robots;
  const int32_t* cur_t;
  const double* global_reward;
};

}  // namespace mdx
