/*
 * CPU ORACLE in plain C (test infrastructure, NOT product code).
 *
 * A sequential restatement of the reference's environment transition, reward, MAPPO featurisation and
 * shaped reward, fast enough to check the CUDA path at BASELINE sizes (thousands of environments,
 * whole episodes) where the Python oracle (oracle/oracle_env.py) would take hours.  It deliberately
 * follows the reference's own control flow -- the restart-scan collision loop, per-robot package
 * scans, float64 sort keys with a stable sort -- rather than the warp-parallel formulation of the
 * kernels, and carries its own MT19937 / legacy-randint so that it shares no code with the product.
 * Every function cites the reference lines (relative to nguyenduchuyiu/marl-delivery) it follows.
 *
 * Pinned by tests/test_oracle_c.py against every golden trajectory in tests/golden/ and against the
 * Python oracle on random configurations.  Only tests/ may load the resulting liboracle.so.
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off: the float64 keys must not be FMA-fused)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { ST_NONE = 0, ST_WAITING = 1, ST_IN_TRANSIT = 2, ST_DELIVERED = 3 };
enum { MV_D = 0, MV_L = 1, MV_R = 2, MV_S = 3, MV_U = 4, MV_BAD = 5 };

/* ---- numpy.random.RandomState(seed): MT19937 + legacy bounded integers ------------------------ */
typedef struct {
  uint32_t s[624];
  int idx;
} mt_t;

static void mt_seed(mt_t* g, uint32_t seed) { /* init_genrand */
  g->s[0] = seed;
  for (int i = 1; i < 624; i++) g->s[i] = 1812433253u * (g->s[i - 1] ^ (g->s[i - 1] >> 30)) + (uint32_t)i;
  g->idx = 624;
}

static uint32_t mt_u32(mt_t* g) {
  if (g->idx >= 624) {
    for (int i = 0; i < 624; i++) {
      uint32_t y = (g->s[i] & 0x80000000u) | (g->s[(i + 1) % 624] & 0x7fffffffu);
      uint32_t v = g->s[(i + 397) % 624] ^ (y >> 1);
      if (y & 1u) v ^= 0x9908b0dfu;
      g->s[i] = v;
    }
    g->idx = 0;
  }
  uint32_t y = g->s[g->idx++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}

/* rng.randint(lo, hi): masked rejection on one 32-bit output per try; no draw for a single value */
static long mt_randint(mt_t* g, long lo, long hi) {
  uint32_t span = (uint32_t)(hi - 1 - lo);
  if (span == 0) return lo;
  uint32_t mask = span;
  for (int sh = 1; sh < 32; sh <<= 1) mask |= mask >> sh;
  uint32_t v;
  do v = mt_u32(g) & mask;
  while (v > span);
  return lo + (long)v;
}

/* ---- environment (env.py:18-345) ------------------------------------------------------------- */
typedef struct {
  int H, W, R, P, T;
  double move_cost, delivery_reward, delay_reward;
  uint8_t* grid;
  int n_free;
  int* free_cells; /* row-major, env.py:155 */
  mt_t rng;
  int t;
  double total_reward;
  int *rcell, *carry;
  int *p_start, *p_target, *p_t0, *p_deadline;
  uint8_t* p_status;
  /* snapshot of the state before the last step, for the shaped reward (MAPPO/trainer.py:211-218) */
  int prev_t;
  int *prev_rcell, *prev_carry;
  uint8_t* prev_status;
} orc_env;

static void spawn(orc_env* e) { /* get_state's side effect, env.py:134-137 */
  for (int j = 0; j < e->P; j++)
    if (e->p_t0[j] == e->t) e->p_status[j] = ST_WAITING;
}

void orc_reset(orc_env* e) { /* env.py:81-125 */
  e->t = 0;
  e->total_reward = 0;
  int* avail = (int*)malloc(sizeof(int) * (size_t)(e->n_free > 0 ? e->n_free : 1));
  memcpy(avail, e->free_cells, sizeof(int) * (size_t)e->n_free);
  int na = e->n_free;
  for (int i = 0; i < e->R; i++) { /* shrinking free list, env.py:161-170 */
    long k = mt_randint(&e->rng, 0, na);
    e->rcell[i] = avail[k];
    memmove(avail + k, avail + k + 1, sizeof(int) * (size_t)(na - k - 1));
    na--;
    e->carry[i] = 0;
  }
  free(avail);
  int lim = e->R < 20 ? e->R : 20;
  int P = e->P;
  int* rec = (int*)malloc(sizeof(int) * 4 * (size_t)(P > 0 ? P : 1));
  for (int i = 0; i < P; i++) {
    int s = e->free_cells[mt_randint(&e->rng, 0, e->n_free)];
    int g;
    do g = e->free_cells[mt_randint(&e->rng, 0, e->n_free)];
    while (g == s);
    int slack = 10 + (int)mt_randint(&e->rng, e->H / 2, 3L * e->H); /* env.py:112, float bound truncated */
    int st = (i <= lim) ? 0 : (int)mt_randint(&e->rng, 1, e->T);  /* env.py:113-116 */
    rec[4 * i] = st;
    rec[4 * i + 1] = s;
    rec[4 * i + 2] = g;
    rec[4 * i + 3] = st + slack;
  }
  /* stable sort by start_time (env.py:119): counting pass over the possible start times */
  int pos = 0;
  for (int pass_t = 0; pass_t <= e->T && pos < P; pass_t++)
    for (int i = 0; i < P; i++)
      if (rec[4 * i] == pass_t) {
        e->p_t0[pos] = rec[4 * i];
        e->p_start[pos] = rec[4 * i + 1];
        e->p_target[pos] = rec[4 * i + 2];
        e->p_deadline[pos] = rec[4 * i + 3];
        e->p_status[pos] = ST_NONE;
        pos++;
      }
  free(rec);
  spawn(e);
}

orc_env* orc_create(const uint8_t* grid, int H, int W, int R, int P, int T, double move_cost, double delivery_reward,
                    double delay_reward, uint32_t seed) {
  orc_env* e = (orc_env*)calloc(1, sizeof(orc_env));
  e->H = H; e->W = W; e->R = R; e->P = P; e->T = T;
  e->move_cost = move_cost; e->delivery_reward = delivery_reward; e->delay_reward = delay_reward;
  e->grid = (uint8_t*)malloc((size_t)H * W);
  memcpy(e->grid, grid, (size_t)H * W);
  e->free_cells = (int*)malloc(sizeof(int) * (size_t)H * W);
  for (int i = 0; i < H * W; i++)
    if (grid[i] == 0) e->free_cells[e->n_free++] = i;
  size_t nr = (size_t)(R > 0 ? R : 1), np_ = (size_t)(P > 0 ? P : 1);
  e->rcell = (int*)calloc(nr, sizeof(int)); e->carry = (int*)calloc(nr, sizeof(int));
  e->prev_rcell = (int*)calloc(nr, sizeof(int)); e->prev_carry = (int*)calloc(nr, sizeof(int));
  e->p_start = (int*)calloc(np_, sizeof(int)); e->p_target = (int*)calloc(np_, sizeof(int));
  e->p_t0 = (int*)calloc(np_, sizeof(int)); e->p_deadline = (int*)calloc(np_, sizeof(int));
  e->p_status = (uint8_t*)calloc(np_, 1); e->prev_status = (uint8_t*)calloc(np_, 1);
  mt_seed(&e->rng, seed);
  orc_reset(e); /* env.py:41 */
  return e;
}

void orc_destroy(orc_env* e) {
  if (!e) return;
  free(e->grid); free(e->free_cells); free(e->rcell); free(e->carry); free(e->prev_rcell); free(e->prev_carry);
  free(e->p_start); free(e->p_target); free(e->p_t0); free(e->p_deadline); free(e->p_status); free(e->prev_status);
  free(e);
}

static void split_action(int a, int* mv, int* op) {
  if (a < 15) { *mv = a % 5; *op = a / 5; }
  else if (a < 18) { *mv = MV_BAD; *op = a - 15; }
  else { *mv = MV_BAD; *op = 0; }
}

/* env.py:173-306.  Returns done; *reward is the step's env reward. */
int orc_step(orc_env* e, const uint8_t* actions, double* reward) {
  const int R = e->R, W = e->W, H = e->H;
  e->prev_t = e->t;
  memcpy(e->prev_rcell, e->rcell, sizeof(int) * (size_t)R);
  memcpy(e->prev_carry, e->carry, sizeof(int) * (size_t)R);
  memcpy(e->prev_status, e->p_status, (size_t)e->P);
  int want[32], settled[32], fin[32], mv[32], op[32];
  for (int i = 0; i < R; i++) { /* proposals, env.py:193-201 */
    split_action(actions[i], &mv[i], &op[i]);
    int r = e->rcell[i] / W, c = e->rcell[i] % W, nr = r, nc = c;
    if (mv[i] == MV_D) nr++; else if (mv[i] == MV_U) nr--; else if (mv[i] == MV_L) nc--; else if (mv[i] == MV_R) nc++;
    if (nr < 0 || nr >= H || nc < 0 || nc >= W || e->grid[nr * W + nc] == 1) { nr = r; nc = c; }
    want[i] = nr * W + nc;
    settled[i] = 0;
    fin[i] = -1;
  }
  /* restart-scan resolution, env.py:203-250: `claimed` plays the role of the `occupied` dict */
  int progress = 1;
  while (progress) {
    progress = 0;
    for (int i = 0; i < R; i++) {
      if (settled[i]) continue;
      int tgt = want[i], holder = -1;
      for (int j = 0; j < R; j++)
        if (e->rcell[j] == tgt) holder = j;
      if (holder >= 0 && holder != i && !settled[holder]) continue;
      for (int j = 0; j < R; j++)
        if (settled[j] && fin[j] == tgt) { tgt = e->rcell[i]; break; }
      fin[i] = tgt;
      settled[i] = 1;
      progress = 1;
      break;
    }
  }
  double rew = 0;
  for (int i = 0; i < R; i++) { /* env.py:248-257 */
    if (!settled[i]) fin[i] = e->rcell[i];
    if ((mv[i] == MV_L || mv[i] == MV_R || mv[i] == MV_U || mv[i] == MV_D) && fin[i] != e->rcell[i]) rew += e->move_cost;
    e->rcell[i] = fin[i];
  }
  for (int i = 0; i < R; i++) { /* package ops in robot order, env.py:260-292 */
    if (op[i] == 1) {
      if (e->carry[i] == 0)
        for (int j = 0; j < e->P; j++)
          if (e->p_status[j] == ST_WAITING && e->p_start[j] == e->rcell[i] && e->p_t0[j] <= e->t) {
            e->carry[i] = j + 1;
            e->p_status[j] = ST_IN_TRANSIT;
            break;
          }
    } else if (op[i] == 2) {
      if (e->carry[i] != 0) {
        int j = e->carry[i] - 1;
        if (e->rcell[i] == e->p_target[j]) {
          e->p_status[j] = ST_DELIVERED;
          rew += (e->t <= e->p_deadline[j]) ? e->delivery_reward : e->delay_reward;
          e->carry[i] = 0;
        }
      }
    }
  }
  e->t += 1;
  e->total_reward += rew;
  int done = (e->t == e->T);
  if (!done) { /* env.py:308-316 */
    done = 1;
    for (int j = 0; j < e->P; j++)
      if (e->p_status[j] != ST_DELIVERED) { done = 0; break; }
  }
  spawn(e);
  *reward = rew;
  return done;
}

/* numpy float32 pairwise add.reduce (n <= 128 here) */
static float np_sum_f32(const float* a, int n) {
  if (n < 8) {
    volatile float s = 0.f;
    for (int i = 0; i < n; i++) s = s + a[i];
    return s;
  }
  volatile float r[8];
  for (int j = 0; j < 8; j++) r[j] = a[j];
  int i = 8;
  for (; i < n - (n % 8); i += 8)
    for (int j = 0; j < 8; j++) r[j] = r[j] + a[i + j];
  volatile float a01 = r[0] + r[1], a23 = r[2] + r[3], a45 = r[4] + r[5], a67 = r[6] + r[7];
  volatile float lo = a01 + a23, hi = a45 + a67;
  volatile float res = lo + hi;
  for (; i < n; i++) res = res + a[i];
  return res;
}

static int manhattan(int a, int b, int W) { return abs(a / W - b / W) + abs(a % W - b % W); }

/* compute_shaped_rewards (MAPPO/helper.py:257-369) for the transition of the LAST orc_step call,
 * evaluated before any reset.  consts: pickup, on-time, late, closer, wasted pick, wasted drop, stuck,
 * idle, away. */
float orc_shaped(const orc_env* e, const uint8_t* actions, const float* k, double reward) {
  const int R = e->R, W = e->W, P = e->P;
  float acc[32];
  for (int i = 0; i < R; i++) {
    volatile float s = 0.f;
    int p0 = e->prev_rcell[i], p1 = e->rcell[i], k0 = e->prev_carry[i], k1 = e->carry[i], mv, op;
    split_action(actions[i], &mv, &op);
    int tracked = k0 != 0 && (e->prev_status[k0 - 1] == ST_WAITING || e->prev_status[k0 - 1] == ST_IN_TRANSIT);
    if (k0 == 0 && k1 != 0) s = s + k[0];
    else if (k0 != 0 && k1 == 0) {
      if (tracked && p1 == e->p_target[k0 - 1]) s = s + ((e->t <= e->p_deadline[k0 - 1]) ? k[1] : k[2]);
    }
    if (op == 1) {
      if (k0 != 0) s = s + k[4];
      else if (k1 == 0) {
        int can = 0;
        for (int j = 0; j < P; j++)
          if (e->prev_status[j] == ST_WAITING && e->p_t0[j] <= e->prev_t && e->p_start[j] == p1) can = 1;
        if (!can) s = s + k[4];
      }
    } else if (op == 2) {
      if (k0 == 0) s = s + k[5];
      else if (k1 != 0) {
        if (tracked && p1 != e->p_target[k0 - 1]) s = s + k[5];
      }
    }
    int moved = p0 != p1;
    if (mv != MV_S && !moved) s = s + k[6];
    int goal = -1;
    if (tracked) goal = e->p_target[k0 - 1];
    else {
      int best = 1 << 30;
      for (int j = 0; j < P; j++)
        if (e->prev_status[j] == ST_WAITING && e->p_t0[j] <= e->prev_t) {
          int d = manhattan(p0, e->p_start[j], W);
          if (d < best) { best = d; goal = e->p_start[j]; }
        }
    }
    if (goal >= 0 && moved) {
      int d0 = manhattan(p0, goal, W), d1 = manhattan(p1, goal, W);
      if (d1 < d0) s = s + k[3];
      else if (d1 > d0) s = s + k[8];
    }
    if (!moved && mv == MV_S && k0 == 0) {
      int near = 0;
      for (int j = 0; j < P; j++)
        if (e->prev_status[j] == ST_WAITING && e->p_t0[j] <= e->prev_t && manhattan(p0, e->p_start[j], W) <= 3) near = 1;
      if (near) s = s + k[7];
    }
    acc[i] = s;
  }
  volatile float r32 = (float)reward; /* python float + np.float32 -> float32 (numpy >= 2) */
  return r32 + np_sum_f32(acc, R);
}

/* ---- featurisers (MAPPO/helper.py:6-255) ------------------------------------------------------- */
static int active_pkg(const orc_env* e, int j) {
  return (e->p_status[j] == ST_WAITING && e->p_t0[j] <= e->t) || e->p_status[j] == ST_IN_TRANSIT;
}

typedef struct {
  double k1, k2;
  float row[5];
} sort_rec;

static void stable_sort(sort_rec* a, sort_rec* tmp, int n) { /* bottom-up merge sort, lexicographic (k1, k2) */
  for (int w = 1; w < n; w *= 2) {
    for (int lo = 0; lo < n; lo += 2 * w) {
      int mid = lo + w < n ? lo + w : n, hi = lo + 2 * w < n ? lo + 2 * w : n, i = lo, j = mid, o = lo;
      while (i < mid && j < hi) {
        int take_right = (a[j].k1 < a[i].k1) || (a[j].k1 == a[i].k1 && a[j].k2 < a[i].k2);
        tmp[o++] = take_right ? a[j++] : a[i++];
      }
      while (i < mid) tmp[o++] = a[i++];
      while (j < hi) tmp[o++] = a[j++];
    }
    memcpy(a, tmp, sizeof(sort_rec) * (size_t)n);
  }
}

static double dq(int a, int b) { return (double)a / (double)b; }
static double sq_key(double x0, double x1) { volatile double a = x0 * x0, b = x1 * x1; return a + b; }

/* obs [R,6,H,W], vec [R,D], gstate [4,H,W], gvec [Dg]; any pointer may be NULL */
void orc_features(const orc_env* e, int K_o, int K_p, int MR, int MP, float* obs, float* vec, float* gstate, float* gvec) {
  const int R = e->R, P = e->P, H = e->H, W = e->W, HW = H * W, T = e->T, t = e->t;
  if (obs) { /* convert_observation, MAPPO/helper.py:6-66 */
    memset(obs, 0, sizeof(float) * (size_t)R * 6 * HW);
    for (int i = 0; i < R; i++) {
      float* o = obs + (size_t)i * 6 * HW;
      for (int x = 0; x < HW; x++) o[x] = (float)e->grid[x];
      o[HW + e->rcell[i]] = 1.f;
      for (int k = 0; k < R; k++)
        if (k != i) o[2 * HW + e->rcell[k]] = 1.f;
      for (int j = 0; j < P; j++) {
        if (e->p_status[j] == ST_WAITING && e->p_t0[j] <= t) o[3 * HW + e->p_start[j]] = 1.f;
        if (active_pkg(e, j)) o[4 * HW + e->p_target[j]] = 1.f;
      }
      if (e->carry[i] != 0 && e->p_status[e->carry[i] - 1] == ST_IN_TRANSIT) o[5 * HW + e->p_target[e->carry[i] - 1]] = 1.f;
    }
  }
  if (vec) { /* generate_vector_features, MAPPO/helper.py:68-165 */
    const int D = 6 + 5 * K_o + 5 * K_p + 1;
    int nmax = (R > P ? R : P) + 1;
    sort_rec* recs = (sort_rec*)malloc(sizeof(sort_rec) * (size_t)nmax);
    sort_rec* tmp = (sort_rec*)malloc(sizeof(sort_rec) * (size_t)nmax);
    for (int i = 0; i < R; i++) {
      float* v = vec + (size_t)i * D;
      memset(v, 0, sizeof(float) * (size_t)D);
      int r0 = e->rcell[i] / W, c0 = e->rcell[i] % W;
      v[0] = (float)dq(r0, H);
      v[1] = (float)dq(c0, W);
      v[2] = e->carry[i] ? 1.f : 0.f;
      if (e->carry[i] && e->p_status[e->carry[i] - 1] == ST_IN_TRANSIT) {
        int j = e->carry[i] - 1, d = e->p_deadline[j] - t;
        v[3] = (float)dq(e->p_target[j] / W - r0, H);
        v[4] = (float)dq(e->p_target[j] % W - c0, W);
        v[5] = T > 0 ? (float)dq(d > 0 ? d : 0, T) : 0.f;
      }
      int n = 0;
      for (int k = 0; k < R; k++) {
        if (k == i) continue;
        int r = e->rcell[k] / W, c = e->rcell[k] % W;
        double x0 = dq(r - r0, H), x1 = dq(c - c0, W);
        recs[n].k1 = 0;
        recs[n].k2 = sq_key(x0, x1);
        recs[n].row[0] = (float)x0;
        recs[n].row[1] = (float)x1;
        recs[n].row[2] = e->carry[k] ? 1.f : 0.f;
        recs[n].row[3] = recs[n].row[4] = 0.f;
        if (e->carry[k] && e->p_status[e->carry[k] - 1] == ST_IN_TRANSIT) {
          int j = e->carry[k] - 1;
          recs[n].row[3] = (float)dq(e->p_target[j] / W - r, H);
          recs[n].row[4] = (float)dq(e->p_target[j] % W - c, W);
        }
        n++;
      }
      stable_sort(recs, tmp, n); /* key x0**2 + x1**2, MAPPO/helper.py:139 */
      for (int k = 0; k < K_o && k < n; k++) memcpy(v + 6 + 5 * k, recs[k].row, sizeof(float) * 5);
      n = 0;
      for (int j = 0; j < P; j++)
        if (e->p_status[j] == ST_WAITING && e->p_t0[j] <= t) {
          int d = e->p_deadline[j] - t;
          double x0 = dq(e->p_start[j] / W - r0, H), x1 = dq(e->p_start[j] % W - c0, W);
          double x4 = T > 0 ? dq(d > 0 ? d : 0, T) : 0.0;
          recs[n].k1 = x4;
          recs[n].k2 = sq_key(x0, x1);
          recs[n].row[0] = (float)x0;
          recs[n].row[1] = (float)x1;
          recs[n].row[2] = (float)dq(e->p_target[j] / W - r0, H);
          recs[n].row[3] = (float)dq(e->p_target[j] % W - c0, W);
          recs[n].row[4] = (float)x4;
          n++;
        }
      stable_sort(recs, tmp, n); /* key (x4, x0**2 + x1**2), MAPPO/helper.py:158 */
      for (int k = 0; k < K_p && k < n; k++) memcpy(v + 6 + 5 * K_o + 5 * k, recs[k].row, sizeof(float) * 5);
      v[D - 1] = T > 0 ? (float)dq(t, T) : 0.f;
    }
    free(recs);
    free(tmp);
  }
  if (gstate) { /* convert_global_state spatial part, MAPPO/helper.py:173-197 */
    memset(gstate, 0, sizeof(float) * 4 * (size_t)HW);
    for (int x = 0; x < HW; x++) gstate[x] = (float)e->grid[x];
    for (int k = 0; k < R; k++) gstate[HW + e->rcell[k]] = 1.f;
    for (int j = 0; j < P; j++) {
      if (e->p_status[j] == ST_WAITING && e->p_t0[j] <= t) gstate[2 * HW + e->p_start[j]] = 1.f;
      if (active_pkg(e, j)) gstate[3 * HW + e->p_target[j]] = 1.f;
    }
  }
  if (gvec) { /* vector part, MAPPO/helper.py:199-255 */
    const int Dg = 6 * MR + 7 * MP + 1;
    memset(gvec, 0, sizeof(float) * (size_t)Dg);
    for (int k = 0; k < MR && k < R; k++) {
      float* o = gvec + 6 * k;
      o[0] = (float)dq(e->rcell[k] / W, H);
      o[1] = (float)dq(e->rcell[k] % W, W);
      o[2] = e->carry[k] ? 1.f : 0.f;
      if (e->carry[k] && e->p_status[e->carry[k] - 1] == ST_IN_TRANSIT) {
        int j = e->carry[k] - 1, d = e->p_deadline[j] - t;
        o[3] = (float)dq(e->p_target[j] / W, H);
        o[4] = (float)dq(e->p_target[j] % W, W);
        o[5] = T > 0 ? (float)dq(d > 0 ? d : 0, T) : 0.f;
      }
    }
    int n = 0;
    for (int j = 0; j < P && n < MP; j++) {
      if (!active_pkg(e, j)) continue;
      float* o = gvec + 6 * MR + 7 * n;
      int wait = e->p_status[j] == ST_WAITING, d = e->p_deadline[j] - t;
      if (wait) {
        o[0] = (float)dq(e->p_start[j] / W, H);
        o[1] = (float)dq(e->p_start[j] % W, W);
      }
      o[2] = (float)dq(e->p_target[j] / W, H);
      o[3] = (float)dq(e->p_target[j] % W, W);
      o[4] = T > 0 ? (float)dq(d > 0 ? d : 0, T) : 0.f;
      o[5] = wait ? 0.f : 1.f;
      o[6] = -1.f;
      if (!wait)
        for (int r = 0; r < R; r++)
          if (e->carry[r] == j + 1) {
            o[6] = MR > 1 ? (float)dq(r, MR - 1) : 0.f;
            break;
          }
      n++;
    }
    gvec[Dg - 1] = T > 0 ? (float)dq(t, T) : 0.f;
  }
}

void orc_state(const orc_env* e, int32_t* robots3, uint8_t* status, int32_t* t, double* total) {
  for (int i = 0; i < e->R; i++) {
    robots3[3 * i] = e->rcell[i] / e->W;
    robots3[3 * i + 1] = e->rcell[i] % e->W;
    robots3[3 * i + 2] = e->carry[i];
  }
  if (status) memcpy(status, e->p_status, (size_t)e->P);
  *t = e->t;
  *total = e->total_reward;
}

void orc_packages(const orc_env* e, int32_t* pk4) { /* start cell, target cell, start_time, deadline per package */
  for (int j = 0; j < e->P; j++) {
    pk4[4 * j] = e->p_start[j];
    pk4[4 * j + 1] = e->p_target[j];
    pk4[4 * j + 2] = e->p_t0[j];
    pk4[4 * j + 3] = e->p_deadline[j];
  }
}

/* One iteration of MAPPO/trainer.py:172-286 (minus the networks) for N environments: step, shaped
 * reward from (before, after-before-reset), auto-reset on done, features of the state the next action
 * is taken from.  Output rows are per env; feature pointers may be NULL. */
void orc_rollout_step(orc_env** envs, int N, const uint8_t* actions, const float* consts, int K_o, int K_p, int MR, int MP,
                      double* reward, float* shaped, uint8_t* done, float* obs, float* vec, float* gstate, float* gvec) {
  for (int n = 0; n < N; n++) {
    orc_env* e = envs[n];
    const int R = e->R, HW = e->H * e->W, D = 6 + 5 * K_o + 5 * K_p + 1, Dg = 6 * MR + 7 * MP + 1;
    const uint8_t* a = actions + (size_t)n * R;
    double r;
    int d = orc_step(e, a, &r);
    reward[n] = r;
    if (shaped) shaped[n] = orc_shaped(e, a, consts, r);
    done[n] = (uint8_t)d;
    if (d) orc_reset(e);
    orc_features(e, K_o, K_p, MR, MP, obs ? obs + (size_t)n * R * 6 * HW : 0, vec ? vec + (size_t)n * R * D : 0,
                 gstate ? gstate + (size_t)n * 4 * HW : 0, gvec ? gvec + (size_t)n * Dg : 0);
  }
}
