/*
 * mdx.h -- C ABI of the B200-native batched marl-delivery environment path ("mdx").
 *
 * Plain C: pointers, sizes and PODs only (no torch / C++ types).  Every entry point returns 0 on
 * success or a negative MDX_E_* code; mdx_last_error() gives the message of the calling thread's
 * last failure.  No C++ exception crosses this boundary.  All "d_" pointers are DEVICE pointers owned
 * by the caller (torch tensors); the library borrows them for the duration of the enqueue and never
 * frees or reallocates them.  Kernels are enqueued on the caller's stream (`stream` is a
 * cudaStream_t passed as void*; NULL = legacy default stream) and the call returns without
 * synchronising unless stated otherwise.
 *
 * Each function names the reference interface (paths relative to nguyenduchuyiu/marl-delivery) it
 * replaces.  One mdx_env owns the state of `n_envs` environments on ONE device; it is not
 * thread-safe.  The host restatement of the reset sampler (mdx_sampler_*, what the device sampler is
 * pinned against) is re-entrant and needs no GPU.
 *
 * Encodings
 *   cell        : r * W_out + c in the (possibly padded) output plane, 0-based; W_out = max map width
 *   action (u8) : 0..14  -> move "DLRSU"[a % 5], package op a / 5   (sorted LabelEncoder order,
 *                 MAPPO/trainer.py:84-89,199-204);  15..17 -> a move string that is none of
 *                 S/L/R/U/D (env.py:333-334 leaves the robot in place) with package op a - 15;
 *                 anything else -> stay, no package op
 *   pkg status  : 0 'None' (not spawned), 1 'waiting', 2 'in_transit', 3 'delivered'  (env.py:16)
 */
#ifndef MDX_H_
#define MDX_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDX_VERSION 2
#define MDX_MAX_ROBOTS 32 /* one warp per environment, robots are lanes */

enum {
  MDX_OK = 0,
  MDX_E_INVALID = -1, /* bad argument / unsupported configuration (ValueError in the Python mirror) */
  MDX_E_CUDA = -2,    /* CUDA runtime failure */
  MDX_E_STARVED = -3, /* (v1 only: the pre-sampled episode ring ran dry; cannot happen since v2) */
  MDX_E_NOMEM = -4
};

/* flags for mdx_step / mdx_step_featurize */
#define MDX_F_AUTO_RESET 1u /* done envs reset inside the kernel: the lane group that owns the env draws its next    \
                               reset block from the env's MT19937 stream in HBM and applies it                     \
                               (MAPPO/trainer.py:229-234: reset, then featurise the reset state) */
#define MDX_F_ACTIONS_I64 2u /* d_actions is int64 [n, n_robots]: the policy's raw samples a in [0, 15), decoded   \
                                in the kernel like MAPPO/trainer.py:195-204 (move a % 5 in sorted LabelEncoder order, \
                                package op a / 5, clamped to '0' when >= 3); negative values mean "stay, no op" */

/* mdx_config.flags */
#define MDX_CFG_BLANK 1 /* skip the constructor's reset draw: envs start empty (t = 0, robots on cell 0, no    \
                           packages) and are filled with mdx_set_state -- used by the per-call helper drop-ins */
#define MDX_CFG_BFS_SHAPING 2 /* OPT-IN, not reference behaviour: the distance terms of the shaped reward (nearest     \
                                 waiting package, closer / away, "work within 3") use shortest-path lengths from the   \
                                 per-map all-pairs BFS table instead of MAPPO/helper.py:284-285's Manhattan distance   \
                                 (BASELINE.json north star vs SURVEY.md Q9).  Identical on maps without inner walls. */

#define MDX_CFG_LAX_ROBOTS 4 /* state injection accepts robots on obstacles or sharing a cell (hand-built state dicts passed to the  \
                                per-call converter drop-ins, which the reference's converters accept too); such an instance must only \
                                be featurised / shaped, never stepped */

#define MDX_CFG_STALE_TRACKER 8 /* OPT-IN: reproduce what MAPPO/trainer.py REALLY feeds compute_shaped_rewards and the three        \
                                   converters.  The trainer keeps one package dict per env (`persistent_packages_list`, :91-130)   \
                                   and never clears it when the env auto-resets (:229-234; the QMIX / IDQ / legacy trainers do):   \
                                   'waiting' entries of earlier episodes survive, shadow the new episode's packages with the same \
                                   id and change the dict (= iteration) order.  With this flag every env carries that dict on the   \
                                   device, in insertion order; mdx_step* maintain it exactly like the trainer's loop (a done env    \
                                   that is auto-reset shows it only the RESET state), mdx_reset applies the update for the reset     \
                                   state without clearing, and the shaped reward / featurisers read IT instead of the env truth.    \
                                   The dict starts empty: call mdx_reset before the first step, as the trainer does.  Default (flag \
                                   off) = env truth = a tracker cleared on reset. */

#define MDX_CFG_JIT_SHAPE 16 /* compile the fused step + featurise kernel AT RUN TIME (NVRTC, sm_100a) for this instance's shape --     \
                                (H, W, R, P, T, K_o, K_p, MR, MP) and the shared-memory carve-up become compile-time constants, loops    \
                                unroll and index arithmetic folds, as in the one shape round 1 baked in (C3) -- and let               \
                                mdx_step_featurize / mdx_featurize launch it.  The cubin is cached on disk ($MDX_JIT_CACHE, default                   \
                                /tmp/mdx_jit_cache).  Needs libnvrtc.so.12 and the kernel sources next to libmdx.so; if either is      \
                                missing the instance keeps the precompiled run-time-shape kernels (see mdx_jit_status).  Same results, \
                                bit for bit. */

/* Shaping constants of compute_shaped_rewards, in this order (MAPPO/helper.py:271-279,
 * QMIX/helper.py:270-278). */
typedef struct mdx_shaping {
  float pickup, delivery_on_time, delivery_late, moved_closer, wasted_pickup, wasted_drop,
      collision_or_stuck, idle_with_tasks, moved_away;
} mdx_shaping;

/* Everything Environment.__init__ (env.py:20-43), VectorizedEnv.__init__
 * (MAPPO/env_vectorized.py:2-11) and the featuriser call sites (MAPPO/trainer.py:237-280) fix. */
typedef struct mdx_config {
  int32_t n_envs;         /* environments owned by this instance (this GPU's shard) */
  int32_t n_robots;       /* <= MDX_MAX_ROBOTS */
  int32_t n_packages;
  int32_t max_time_steps; /* T */
  double move_cost, delivery_reward, delay_reward; /* env.py:21 */
  mdx_shaping shaping;
  int32_t k_others;              /* max_other_robots_to_observe  (MAPPO/helper.py:70) */
  int32_t k_packages;            /* max_packages_to_observe */
  int32_t max_robots_in_state;   /* MAPPO/helper.py:169 */
  int32_t max_packages_in_state;
  int32_t n_maps;                /* >= 1 */
  const int32_t* map_rows;       /* [n_maps] */
  const int32_t* map_cols;       /* [n_maps] */
  const uint8_t* const* map_cells; /* [n_maps] row-major rows*cols bytes, 0 free, 1 obstacle */
  const int32_t* env_map;        /* [n_envs] map id per env, NULL = all map 0 (4-map notebook: idx % n_maps) */
  const int64_t* env_seed;       /* [n_envs] RandomState seed per env; NULL = base_seed + i */
  int64_t base_seed;             /* env i uses seed base_seed + i  (MAPPO/env_vectorized.py:7-9) */
  int32_t device;                /* CUDA ordinal */
  int32_t ring_episodes;         /* ignored since v2 (v1: depth of the host-fed ring of pre-sampled episodes) */
  int32_t sampler_threads;       /* ignored since v2 (v1: host threads for reset sampling) */
  int32_t flags;                 /* MDX_CFG_* */
} mdx_config;

typedef struct mdx_dims {
  int32_t n_envs, n_robots, n_packages, max_time_steps;
  int32_t rows_out, cols_out; /* padded plane size of obs / gstate */
  int32_t vec_dim;            /* 6 + 5*k_others + 5*k_packages + 1 */
  int32_t gvec_dim;           /* 6*max_robots_in_state + 7*max_packages_in_state + 1 */
  int64_t bytes_per_env_step; /* 4*(R*6*H*W + R*D + 4*H*W + Dg) + 4 + 1 + R */
  int32_t ring_episodes;      /* 0 since v2 */
  int32_t reserved;
} mdx_dims;

/* Host-side copy of one environment (the lazy dict / attribute views of the Python mirror:
 * env.py:127-147 get_state, .robots[i].position/.carrying, .packages[j].*, .t, .total_reward). */
typedef struct mdx_state_view {
  int32_t t;
  int32_t episode;          /* number of resets applied since construction (0 = the constructor's) */
  int32_t map_id;
  int32_t reserved;
  double total_reward;
  int32_t* robot_row;       /* [n_robots] 0-based, caller-allocated */
  int32_t* robot_col;
  int32_t* robot_carry;     /* 0 or package id (1-based) */
  int32_t* pkg_start_row;   /* [n_packages] */
  int32_t* pkg_start_col;
  int32_t* pkg_target_row;
  int32_t* pkg_target_col;
  int32_t* pkg_start_time;
  int32_t* pkg_deadline;
  int32_t* pkg_status;
} mdx_state_view;

typedef struct mdx_env mdx_env;
typedef struct mdx_sampler mdx_sampler;

const char* mdx_last_error(void);
int32_t mdx_version(void);

/* Environment.__init__ x n_envs / VectorizedEnv.__init__ (env.py:20-43, MAPPO/env_vectorized.py:2-11).
 * Like the reference constructor, this draws reset block 0 for every env and puts the envs in it. */
int32_t mdx_create(const mdx_config* cfg, mdx_env** out);
int32_t mdx_destroy(mdx_env* env);
int32_t mdx_get_dims(const mdx_env* env, mdx_dims* out);
/* 1: the run-time-specialised kernel (MDX_CFG_JIT_SHAPE) is loaded and used; 0: not requested; -1: requested but unavailable
 * (msg, if not NULL, receives the reason) -- the precompiled kernels run. */
int32_t mdx_jit_status(const mdx_env* env, char* msg, int32_t msg_bytes);

/* Environment.reset / VectorizedEnv.reset(indices) (env.py:81-125, QMIX/env_vectorized.py:13-22):
 * every listed env (HOST int32 ids, NULL = all) advances to its next reset block, drawn ON THE DEVICE
 * from the env's own numpy-legacy MT19937 stream (624 state words per env in HBM) in the reference's
 * draw order (env.py:96-123, 150-170): robots from the shrinking row-major free list, then per package
 * start, target != start, deadline slack, start_time; stable sort by start_time.  Enqueued on `stream`,
 * no host synchronisation.  An env listed k times is reset k times. */
int32_t mdx_reset(mdx_env* env, const int32_t* h_env_ids, int32_t n, void* stream);

/* v1 pre-sampled reset blocks on the host into a per-env ring and this call refilled it.  Since v2 an
 * env draws its next block on the device at the moment it resets (mdx_reset, MDX_F_AUTO_RESET), so there
 * is nothing to prefetch and an env can never starve: kept as a no-op for v1 callers. */
int32_t mdx_prefetch(mdx_env* env, int32_t episodes_ahead, void* stream);

/* Environment.step / VectorizedEnv.step(actions, indices) (env.py:173-306,
 * QMIX/env_vectorized.py:24-37) fused with compute_shaped_rewards (MAPPO/helper.py:257-369).
 * d_actions [n, n_robots] u8; d_env_ids [n] int32 DEVICE ids or NULL (= all, n must be n_envs).  Device ids cannot be
 * checked on the host without a synchronisation: they must be pairwise distinct (two lane groups stepping one env in the
 * same launch would race; the reference steps a repeated index twice in sequence), and an id outside [0, n_envs) is
 * skipped by the kernel (its output row is left unwritten) rather than dereferenced.  The Python mirror validates id
 * lists on the host and id tensors once per tensor version.
 * Outputs (each may be NULL): d_reward [n] f64 env reward, d_shaped [n] f32 shaped team reward,
 * d_done [n] u8.  Rows follow the order of d_env_ids. */
int32_t mdx_step(mdx_env* env, const uint8_t* d_actions, const int32_t* d_env_ids, int32_t n,
                 double* d_reward, float* d_shaped, uint8_t* d_done, uint32_t flags, void* stream);

/* n_steps consecutive mdx_step calls over ALL envs in one C call (the trainers' hot loops call vec_env.step once per
 * step: MAPPO/trainer.py:208, QMIX/trainer.py:435): step s reads d_actions[s] ([n_steps, n_envs, n_robots]) and writes
 * row s of d_reward / d_shaped / d_done ([n_steps, n_envs], each may be NULL).  For launch-bound batches (a few
 * thousand envs, kernel ~7 us) this removes the per-step host overhead of the binding; the call only enqueues, so it can
 * also be captured into a CUDA graph by the caller (VectorizedEnv.rollout_graph does). */
int32_t mdx_step_n(mdx_env* env, const uint8_t* d_actions, int32_t n_steps, double* d_reward, float* d_shaped,
                   uint8_t* d_done, uint32_t flags, void* stream);

/* convert_observation / generate_vector_features / convert_global_state for every listed env and
 * every robot (MAPPO/helper.py:6-255), written straight into the caller's tensors:
 * d_obs [n, R, 6, rows_out, cols_out], d_vec [n, R, vec_dim], d_gstate [n, 4, rows_out, cols_out],
 * d_gvec [n, gvec_dim], all f32, any may be NULL. */
int32_t mdx_featurize(mdx_env* env, const int32_t* d_env_ids, int32_t n, float* d_obs, float* d_vec,
                      float* d_gstate, float* d_gvec, void* stream);

/* One launch for one iteration of MAPPO/trainer.py:172-286: step + shaped reward + (auto-reset) +
 * featurisation of the state the next action is taken from. */
int32_t mdx_step_featurize(mdx_env* env, const uint8_t* d_actions, const int32_t* d_env_ids, int32_t n,
                           double* d_reward, float* d_shaped, uint8_t* d_done, float* d_obs,
                           float* d_vec, float* d_gstate, float* d_gvec, uint32_t flags, void* stream);

/* mdx_step / mdx_step_featurize for the CONTIGUOUS env range [first_env, first_env + n): every pointer addresses row 0 of
 * the range (the caller offsets its full-batch tensors).  Ranges of one instance are independent, so a caller may keep
 * several in flight on different streams -- VectorizedEnv.split() steps two half-batches on two streams, which hides one
 * half's launch ramp and drain behind the other half's steady state and lets a trainer overlap policy inference on one half
 * with env stepping of the other (MAPPO/trainer.py:172-286 is one such loop per env slice). */
int32_t mdx_step_range(mdx_env* env, const uint8_t* d_actions, int32_t first_env, int32_t n, double* d_reward, float* d_shaped,
                       uint8_t* d_done, uint32_t flags, void* stream);
int32_t mdx_step_featurize_range(mdx_env* env, const uint8_t* d_actions, int32_t first_env, int32_t n, double* d_reward,
                                 float* d_shaped, uint8_t* d_done, float* d_obs, float* d_vec, float* d_gstate, float* d_gvec,
                                 uint32_t flags, void* stream);

/* compute_shaped_rewards on caller-supplied "current" robots (MAPPO/helper.py:257-369) against the
 * stored state as "previous": d_cur_robots [n, R] u32 = cell | carry << 16, d_cur_t [n] int32,
 * d_global_reward [n] f64 -> d_out [n] f32.  Does not modify the environments. */
int32_t mdx_shaped_reward(mdx_env* env, const uint32_t* d_cur_robots, const int32_t* d_cur_t,
                          const uint8_t* d_actions, const double* d_global_reward, float* d_out,
                          void* stream);

/* Introspection (evaluation.py:50, main.py:86-103, IDQ/trainer.py:92): synchronous copies. */
int32_t mdx_get_state(mdx_env* env, int32_t env_id, mdx_state_view* out, void* stream);
/* State injection for hand-built cases (Group-Project.pdf worked example, helper drop-ins). */
int32_t mdx_set_state(mdx_env* env, int32_t env_id, const mdx_state_view* in, void* stream);

/* The same for many envs at once, as packed rows (what the dict drop-ins VectorizedEnv.reset / .step and the per-call helper
 * functions use: ONE launch + ONE copy + ONE synchronisation per call instead of one round trip per field and env).
 * A row is mdx_state_row_bytes() bytes of little-endian 4-byte words:
 *   [0] t  [1] episode  [2] map id  [3] 0  [4..5] total_reward (f64)
 *   [6 .. 6+R)                      robots: cell | carried id << 16
 *   [S .. S+2*Pp8), S = 6+R rounded up to 2:  schedule u16 [4][Pp8] = start cell, target cell, start_time, deadline by
 *                                   package id, Pp8 = max(8, P rounded up to 8); padding start_time = 0xFFFF
 *   [S+2*Pp8 .. +Pp16/4)            status u8 [Pp16], Pp16 = max(16, P rounded up to 16)
 * padded to a multiple of 4 words.  h_env_ids: HOST ids, NULL = all envs (n = n_envs).
 * mdx_get_states returns when the rows are in h_rows (it synchronises `stream` once). */
int64_t mdx_state_row_bytes(const mdx_env* env);
int32_t mdx_get_states(mdx_env* env, const int32_t* h_env_ids, int32_t n, void* h_rows, void* stream);
/* Rows are validated on the host like mdx_set_state (time 0..65534, robots on free, pairwise distinct cells as
 * Environment.add_robot demands (env.py:73-79), known carried ids, packages inside the map), then written with one
 * asynchronous copy + one kernel.  Episode cursor, map id and generator state are not touched. */
int32_t mdx_set_states(mdx_env* env, const int32_t* h_env_ids, int32_t n, const void* h_rows, void* stream);

/* Environment.step / VectorizedEnv.step(actions, indices) for callers whose actions and results live in HOST memory
 * (env.py:173-306 behind the reference's list/dict API): one H2D copy of the action codes [n, n_robots], the step
 * kernel, a packed snapshot of the stepped envs, ONE copy back and ONE synchronisation.  h_reward [n] f64, h_shaped [n]
 * f32 (NULL: the shaped reward is not evaluated), h_done [n] u8 and h_rows [n rows] may each be NULL. */
int32_t mdx_step_host(mdx_env* env, const uint8_t* h_actions, const int32_t* h_env_ids, int32_t n, double* h_reward,
                      float* h_shaped, uint8_t* h_done, void* h_rows, uint32_t flags, void* stream);

/* mdx_shaped_reward for HOST callers (the per-call compute_shaped_rewards drop-in): all arguments and the result are host
 * arrays of the shapes mdx_shaped_reward documents, for ALL envs of the instance; one launch + one synchronisation. */
int32_t mdx_shaped_reward_host(mdx_env* env, const uint32_t* h_cur_robots, const int32_t* h_cur_t, const uint8_t* h_actions,
                               const double* h_global_reward, float* h_out, void* stream);

/* mdx_featurize for HOST callers (the per-call helper drop-ins convert_observation / generate_vector_features /
 * convert_global_state return numpy, MAPPO/helper.py:6-255): the listed envs' feature tensors straight into host arrays
 * (any may be NULL), one launch + one synchronisation; small results are written by the kernel directly into pinned host
 * memory, larger ones take one copy back. */
int32_t mdx_featurize_host(mdx_env* env, const int32_t* h_env_ids, int32_t n, float* h_obs, float* h_vec, float* h_gstate,
                           float* h_gvec, void* stream);

/* MDX_CFG_STALE_TRACKER instances only.  mdx_tracker_clear empties the dict of the listed envs (HOST ids, NULL = all): what the
 * QMIX / IDQ trainers do on reset (QMIX/trainer.py:523-526).  mdx_tracker_get / mdx_tracker_set copy one env's dict to / from the
 * host in dict order (synchronous; tests, and the per-call converter drop-ins, which receive arbitrary dicts): h_rows5 is u16
 * [5][Pp8] = package id, start cell, target cell, start_time, deadline; h_status u8 [Pp16] = 1 waiting / 2 in_transit;
 * Pp8 = max(8, P rounded up to 8), Pp16 = max(16, P rounded up to 16); ids must be distinct and in 1..P. */
int32_t mdx_tracker_clear(mdx_env* env, const int32_t* h_env_ids, int32_t n, void* stream);
int32_t mdx_tracker_get(mdx_env* env, int32_t env_id, int32_t* h_n, uint16_t* h_rows5, uint8_t* h_status, void* stream);
int32_t mdx_tracker_set(mdx_env* env, int32_t env_id, int32_t n, const uint16_t* h_rows5, const uint8_t* h_status, void* stream);

/* d_out8 [8] int64 = {episodes_done, robot_moves, on_time, late, pickups,
 * env_steps, sum_episode_len, starved_envs} summed over this instance's envs (cumulative since
 * creation).  The caller all-reduces it across ranks. */
int32_t mdx_episode_stats(mdx_env* env, int64_t* d_out8, void* stream);

/* ---- callers either side of the path -------------------------------------- */

/* GreedyAgents.get_actions for every listed env (greedyagent.py:42-174, incl. run_bfs :6-41 and
 * update_move_to_target :64-100): one action code per robot into d_actions [n, n_robots] u8, ready for
 * mdx_step.  The agent's bookkeeping (per-robot target, the `packages_free` flags over its own package
 * list -- which holds every t = 0 package twice and is indexed by package_id - 1, exactly like the
 * reference) lives on the device per env; it is re-initialised whenever the env is at t = 0, i.e. a new
 * GreedyAgents() + init_agents(state) per episode as in evaluation.py:9-66.  Call it once per env step.
 * run_bfs is a lookup in the per-map all-pairs BFS table (uint16, built on first use). */
int32_t mdx_greedy_actions(mdx_env* env, const int32_t* d_env_ids, int32_t n, uint8_t* d_actions, void* stream);

/* The all-pairs BFS table of map `map_id` (greedyagent.py:6-29 run from every free cell): h_out
 * [rows_out*cols_out][rows_out*cols_out] uint16 in output-plane cell ids, 0xFFFF = unreachable. */
int32_t mdx_bfs_table(mdx_env* env, int32_t map_id, uint16_t* h_out);

/* The other featurisers of the same boundary, for every listed env and robot in one launch:
 * d_obs6 [n, R, 6, rows_out, cols_out] = convert_state of IDQ/networks.py:112-217 (identical code in
 * qmix/networks.py:243-348): map, urgency of waiting packages and their start cells (both only when the
 * robot carries nothing), other robots, this robot, target of the carried package;
 * d_gstate7 [n, 7, rows_out, cols_out] = convert_global_state_to_tensor of qmix/networks.py:350-468 with the
 * trainer's (7, H, W) shape: map, robots, carrying robots, waiting starts, waiting targets, in-transit
 * targets, urgency.  float32, either may be NULL. */
int32_t mdx_featurize_alt(mdx_env* env, const int32_t* d_env_ids, int32_t n, float* d_obs6, float* d_gstate7, void* stream);

/* mdx_step that also evaluates IDQ's per-agent reward_shaping (IDQ/networks.py:228-349) of the transition:
 * d_agent_rewards [n, n_robots] f64 (the reference returns Python floats).  ops_as_int = 0 reproduces what
 * IDQ/trainer.py:276-291 does -- it passes (move_str, op_str) tuples, so `package_op == 1` never holds and
 * only the stay penalty fires; ops_as_int = 1 evaluates every branch as for integer package ops. */
int32_t mdx_step_idq(mdx_env* env, const uint8_t* d_actions, const int32_t* d_env_ids, int32_t n, double* d_reward,
                     double* d_agent_rewards, uint8_t* d_done, int32_t ops_as_int, uint32_t flags, void* stream);

/* One IDQ trainer iteration in ONE launch (IDQ/trainer.py:276-300): mdx_step_idq, then -- on the state after the step and
 * after any on-device auto-reset, like the trainer's convert_state calls on the next state -- mdx_featurize_alt into d_obs6 /
 * d_gstate7 (either may be NULL).  Maps whose planes are not whole float4s fall back to the two launches internally. */
int32_t mdx_step_idq_featurize(mdx_env* env, const uint8_t* d_actions, const int32_t* d_env_ids, int32_t n, double* d_reward,
                               double* d_agent_rewards, uint8_t* d_done, float* d_obs6, float* d_gstate7, int32_t ops_as_int,
                               uint32_t flags, void* stream);

/* mdx_step_idq_featurize for the CONTIGUOUS env range [first_env, first_env + n) (pointers address row 0 of the range; see
 * mdx_step_featurize_range): several ranges of one instance may be in flight on different streams.  Fused kernel only (maps whose
 * planes are whole float4s). */
int32_t mdx_step_idq_featurize_range(mdx_env* env, const uint8_t* d_actions, int32_t first_env, int32_t n, double* d_reward,
                                     double* d_agent_rewards, uint8_t* d_done, float* d_obs6, float* d_gstate7, int32_t ops_as_int,
                                     uint32_t flags, void* stream);

/* IDQ reward_shaping on caller-supplied "current" robots against the stored state as "previous" (the per-call
 * drop-in of IDQ/networks.py:228-349; same inputs as mdx_shaped_reward) -> d_out [n_envs, n_robots] f64. */
int32_t mdx_reward_shaping_idq(mdx_env* env, const uint32_t* d_cur_robots, const int32_t* d_cur_t, const uint8_t* d_actions,
                               int32_t ops_as_int, double* d_out, void* stream);

/* GAE + returns of MAPPO/trainer.py:288-301 for a [T, N] rollout, float32 with the reference's order of
 * operations: delta = r + gamma*v_next*nonterminal - v; adv = last = delta + (gamma*lambda)*nonterminal*last;
 * ret = adv + v.  d_rewards/d_values [T, N] f32, d_dones [T, N] u8, d_last_value [N] f32 ->
 * d_advantages, d_returns [T, N] f32.  Needs no mdx_env. */
int32_t mdx_gae(const float* d_rewards, const float* d_values, const uint8_t* d_dones, const float* d_last_value,
                int32_t T, int32_t N, double gamma, double gae_lambda, float* d_advantages, float* d_returns,
                void* stream);

/* Env-state checkpointing (the reference checkpoints only network weights, MAPPO/helper.py:371-389): the
 * complete state of the instance -- device arrays, counters, greedy bookkeeping and every env's MT19937
 * state -- as one host blob.  A loaded instance continues bit-identically.  The header carries the shape and a
 * hash of the maps, env -> map assignment, rewards, shaping constants and feature sizes; a blob that does not
 * match, or is corrupt, is refused BEFORE anything is written (the instance stays as it was). */
int64_t mdx_checkpoint_bytes(const mdx_env* env);
int32_t mdx_checkpoint_save(mdx_env* env, void* h_buf, int64_t bytes, void* stream);
int32_t mdx_checkpoint_load(mdx_env* env, const void* h_buf, int64_t bytes, void* stream);

/* ---- host-only pieces (no GPU needed) ------------------------------------------------------- */

/* The reference's reset draw order on numpy's legacy RandomState(seed) stream (env.py:96-123,
 * 150-170; numpy: init_genrand seeding + masked-rejection bounded integers, one 32-bit output per
 * try).  Each mdx_sampler_next call produces the next reset block of ONE environment:
 * robot_cell [R], pkg4 [4][P] = start cell, target cell, start_time, deadline in package-id order
 * (cells are r * cols + c of the map itself). */
int32_t mdx_sampler_create(uint32_t seed, const uint8_t* cells, int32_t rows, int32_t cols,
                           int32_t n_robots, int32_t n_packages, int32_t max_time_steps,
                           mdx_sampler** out);
int32_t mdx_sampler_next(mdx_sampler* s, int32_t* robot_cell, int32_t* pkg4);
int32_t mdx_sampler_destroy(mdx_sampler* s);
/* raw generator access, for pinning against numpy: n draws of RandomState(seed).randint(lo, hi) */
int32_t mdx_rng_randint(uint32_t seed, int64_t lo, int64_t hi, int32_t n, int64_t* out);

/* Lookup tables the kernels read, exposed for parity tests against plain Python arithmetic:
 * quotient q[d + (n_out-1)] = (float)((double)d / n) for d in [-(n_out-1), n_out-1], and the dense
 * rank of the float64 sort key (dr/rows)**2 + (dc/cols)**2 (MAPPO/helper.py:139,158) over
 * dr in [-(rows_out-1), rows_out-1] x dc in [-(cols_out-1), cols_out-1]. */
int32_t mdx_lut_quotient(int32_t n, int32_t n_out, float* out /* [2*n_out-1] */);
int32_t mdx_lut_rank(int32_t rows, int32_t cols, int32_t rows_out, int32_t cols_out,
                     uint16_t* out /* [(2*rows_out-1)*(2*cols_out-1)] */, int32_t* n_ranks);
/* numpy's float32 pairwise `.sum()` (MAPPO/helper.py:369) restated on the host, same routine the
 * kernel implements with shuffles. */
float mdx_np_sum_f32(const float* x, int32_t n);

#ifdef __cplusplus
}
#endif
#endif /* MDX_H_ */
