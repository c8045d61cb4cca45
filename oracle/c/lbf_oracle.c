/* LevelBasedForaging-v3 CPU oracle in C (TEST INFRASTRUCTURE ONLY -- never linked into the product).
 *
 * A second restatement, independent in form, of posggym's LevelBasedForagingModel (posggym's port
 * of lb-foraging; un-vendored dependency posggym>=0.5.0, /root/reference/pyproject.toml:36),
 * following oracle/lbf.py function by function: _get_next_state (move validity, clashes, loading
 * in player index order), _adjacent_food, _get_tuple_obs, initial_state_from_noise.  It works on
 * player / food structs with explicit loops, not on the kernels' packed words and tables, and is
 * pinned against the Python oracle's golden vectors (tests/test_oracle_native.py).  What the
 * Python oracle documents as guesses (static layout, level sampling) is inherited unchanged.
 *
 * Scope: the reference's experiment kwargs (2 players, 8 foods, static layout, tuple observations,
 * baseline_exps/env_data/LevelBasedForaging-v3/env_kwargs.yaml).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "oracle_common.h"

#define NPL 2
#define NFOOD 8
#define NFLD 15 /* p0x p0y p0lvl p1x p1y p1lvl f0..f7 spawned */
enum { NONE = 0, NORTH = 1, SOUTH = 2, WEST = 3, EAST = 4, LOAD = 5 };

typedef struct { int x[NPL], y[NPL], lvl[NPL]; int food[NFOOD]; int spawned; } LBF;
typedef struct {
  int size, sight, max_agent_level, force_coop, normalize_reward;
  int fx[NFOOD], fy[NFOOD], px[NPL], py[NPL];
} Cfg;
static Cfg C;

int lbf_init(int size, int sight, int max_agent_level, int force_coop, int normalize_reward) {
  memset(&C, 0, sizeof(C));
  C.size = size; C.sight = sight; C.max_agent_level = max_agent_level; C.force_coop = force_coop;
  C.normalize_reward = normalize_reward;
  /* static_food_coords: LAST max_food cells of sorted product(range(1, size - 1, 2), repeat=2) */
  int lx[64], ly[64], nl = 0;
  for (int x = 1; x < size - 1; x += 2)
    for (int y = 1; y < size - 1; y += 2) { lx[nl] = x; ly[nl++] = y; }
  int k = 0;
  for (int j = nl - NFOOD; j < nl; ++j) { C.fx[k] = lx[j]; C.fy[k] = ly[j]; ++k; }
  /* static_player_coords: evenly spaced along the border, from the top-right corner, walking down
   * the right side, along the bottom, up the left side and back along the top */
  int bx[64], by[64], nb = 0;
  for (int y = 0; y < size; ++y) { bx[nb] = size - 1; by[nb++] = y; }
  for (int x = size - 2; x >= 0; --x) { bx[nb] = x; by[nb++] = size - 1; }
  for (int y = size - 2; y >= 0; --y) { bx[nb] = 0; by[nb++] = y; }
  for (int x = 1; x < size - 1; ++x) { bx[nb] = x; by[nb++] = 0; }
  const double step = (double)nb / (double)NPL;
  for (int i = 0; i < NPL; ++i) { const int j = (int)((double)i * step); C.px[i] = bx[j]; C.py[i] = by[j]; }
  return k * 100 + nb;
}

static int food_at(const int* food, int x, int y) {
  for (int k = 0; k < NFOOD; ++k) if (food[k] > 0 && C.fx[k] == x && C.fy[k] == y) return k;
  return -1;
}
/* first uncollected food next to (x, y) in the order N, S, W, E */
static int adjacent_food(const int* food, int x, int y) {
  static const int dx[4] = {0, 0, -1, 1}, dy[4] = {-1, 1, 0, 0};
  for (int d = 0; d < 4; ++d) { const int k = food_at(food, x + dx[d], y + dy[d]); if (k >= 0) return k; }
  return -1;
}

/* model.step (deterministic) */
static int lbf_step1(const LBF* s, const int* act, LBF* ns, double* rew) {
  *ns = *s;
  int tx[NPL], ty[NPL], loading[NPL], nload = 0;
  for (int i = 0; i < NPL; ++i) {
    tx[i] = s->x[i]; ty[i] = s->y[i];
    const int a = act[i];
    if (a >= NORTH && a <= EAST) {
      const int cx = s->x[i] + (a == EAST) - (a == WEST), cy = s->y[i] + (a == SOUTH) - (a == NORTH);
      if (cx >= 0 && cx < C.size && cy >= 0 && cy < C.size && food_at(s->food, cx, cy) < 0) { tx[i] = cx; ty[i] = cy; }
    } else if (a == LOAD && adjacent_food(s->food, s->x[i], s->y[i]) >= 0) {
      loading[nload++] = i;
    }
  }
  for (int i = 0; i < NPL; ++i) {
    int claims = 0;
    for (int j = 0; j < NPL; ++j) claims += (tx[j] == tx[i] && ty[j] == ty[i]);
    if (claims > 1) { ns->x[i] = s->x[i]; ns->y[i] = s->y[i]; } else { ns->x[i] = tx[i]; ns->y[i] = ty[i]; }
  }
  for (int j = 0; j < NPL; ++j) rew[j] = 0.0;
  int pending[NPL], np = nload;
  memcpy(pending, loading, sizeof(int) * nload);
  while (np > 0) {
    const int i = pending[0];
    memmove(pending, pending + 1, sizeof(int) * (np - 1)); --np;
    const int k = adjacent_food(ns->food, ns->x[i], ns->y[i]);
    if (k < 0) continue;
    int group[NPL], ng = 0, keep[NPL], nk = 0;
    group[ng++] = i;
    for (int q = 0; q < np; ++q) {
      const int j = pending[q];
      if (abs(ns->x[j] - C.fx[k]) + abs(ns->y[j] - C.fy[k]) == 1) group[ng++] = j; else keep[nk++] = j;
    }
    memcpy(pending, keep, sizeof(int) * nk); np = nk;
    int level_sum = 0;
    for (int q = 0; q < ng; ++q) level_sum += ns->lvl[group[q]];
    if (level_sum < ns->food[k]) continue;
    for (int q = 0; q < ng; ++q) {
      const int j = group[q];
      double r = (double)(ns->lvl[j] * ns->food[k]);
      if (C.normalize_reward) r = r / (double)(level_sum * s->spawned);
      rew[j] += r;
    }
    ns->food[k] = 0;
  }
  for (int k = 0; k < NFOOD; ++k) if (ns->food[k] > 0) return 0;
  return 1;
}

/* _get_tuple_obs: 30 entries */
static void observe(const LBF* s, int idx, int* out) {
  const int sg = C.sight;
  int n = 0;
  out[n++] = sg; out[n++] = sg; out[n++] = s->lvl[idx];
  for (int p = 0; p < NPL; ++p) {
    if (p == idx) continue;
    const int dx = s->x[p] - s->x[idx], dy = s->y[p] - s->y[idx];
    if (abs(dx) <= sg && abs(dy) <= sg) { out[n++] = dx + sg; out[n++] = dy + sg; out[n++] = s->lvl[p]; }
    else { out[n++] = -1; out[n++] = -1; out[n++] = 0; }
  }
  int seen = 0;
  for (int k = 0; k < NFOOD; ++k) {
    const int dx = C.fx[k] - s->x[idx], dy = C.fy[k] - s->y[idx];
    if (s->food[k] > 0 && abs(dx) <= sg && abs(dy) <= sg) { out[n++] = dx + sg; out[n++] = dy + sg; out[n++] = s->food[k]; ++seen; }
  }
  for (; seen < NFOOD; ++seen) { out[n++] = -1; out[n++] = -1; out[n++] = 0; }
}

static void load(const int8_t* f, LBF* s) {
  for (int i = 0; i < NPL; ++i) { s->x[i] = f[3 * i]; s->y[i] = f[3 * i + 1]; s->lvl[i] = f[3 * i + 2]; }
  for (int k = 0; k < NFOOD; ++k) s->food[k] = f[6 + k];
  s->spawned = f[14];
}
static void store(const LBF* s, int8_t* f) {
  for (int i = 0; i < NPL; ++i) { f[3 * i] = (int8_t)s->x[i]; f[3 * i + 1] = (int8_t)s->y[i]; f[3 * i + 2] = (int8_t)s->lvl[i]; }
  for (int k = 0; k < NFOOD; ++k) f[6 + k] = (int8_t)s->food[k];
  f[14] = (int8_t)s->spawned;
}

/* batch step on field arrays: state [n][15] i8, actions [n][2] u8 -> next [n][15], obs [n][2][30] i8,
 * rewards [n][2], all_done [n] */
void lbf_step(int64_t n, const int8_t* state, const uint8_t* actions, int8_t* next, int8_t* obs, double* rewards,
              uint8_t* all_done) {
  for (int64_t i = 0; i < n; ++i) {
    LBF s, ns;
    load(state + i * NFLD, &s);
    const int act[NPL] = {actions[2 * i], actions[2 * i + 1]};
    all_done[i] = (uint8_t)lbf_step1(&s, act, &ns, rewards + 2 * i);
    store(&ns, next + i * NFLD);
    int o[30];
    for (int a = 0; a < NPL; ++a) { observe(&ns, a, o); for (int k = 0; k < 30; ++k) obs[(i * NPL + a) * 30 + k] = (int8_t)o[k]; }
  }
}

/* ---- the kernels' Philox conventions (oracle/rollout.py lbf_step_noise / lbf_initial_picks) ---- */
static void initial_state(uint64_t seed, uint64_t particle, LBF* s) {
  uint32_t a[4], b[4], c[4];
  philox_block(seed, particle, 0u, STREAM_INIT, a);
  philox_block(seed, particle, 0u, STREAM_INIT2, b);
  philox_block(seed, particle, 0u, 0xB2u, c);
  int max_level = 0;
  for (int i = 0; i < NPL; ++i) {
    s->x[i] = C.px[i]; s->y[i] = C.py[i];
    s->lvl[i] = (int)below(a[i], (uint32_t)C.max_agent_level) + 1;
    max_level += s->lvl[i];
  }
  const uint32_t fw[NFOOD] = {a[2], a[3], b[0], b[1], b[2], b[3], c[0], c[1]};
  s->spawned = 0;
  for (int k = 0; k < NFOOD; ++k) {
    s->food[k] = C.force_coop ? max_level : (int)below(fw[k], (uint32_t)max_level) + 1;
    s->spawned += s->food[k];
  }
}
static void step_noise(uint64_t seed, uint64_t particle, uint32_t step, int* act) {
  uint32_t w[4];
  philox_block(seed, particle >> 2, step, STREAM_STEP, w);
  const uint32_t idx = below(w[particle & 3], 36u);
  act[1] = (int)(idx / 6u);
  act[0] = (int)(idx - 6u * (uint32_t)act[1]);
}

/* ---- packed words of csrc/lbf.cuh / envs/lbf.py ---- */
static void pack_state(const LBF* s, uint32_t* w) {
  w[0] = (uint32_t)s->spawned << 20;
  for (int i = 0; i < NPL; ++i) w[0] |= (uint32_t)(s->x[i] | (s->y[i] << 4) | (s->lvl[i] << 8)) << (10 * i);
  w[1] = 0;
  for (int k = 0; k < NFOOD; ++k) w[1] |= (uint32_t)s->food[k] << (3 * k);
}
static void pack_obs(const int* o, uint64_t* w) {
  uint32_t h[4] = {0, 0, 0, 0};
  h[0] = (uint32_t)(o[0] + 1) | ((uint32_t)(o[1] + 1) << 3) | ((uint32_t)o[2] << 6) | ((uint32_t)(o[3] + 1) << 8) |
         ((uint32_t)(o[4] + 1) << 11) | ((uint32_t)o[5] << 14);
  for (int slot = 0; slot < NFOOD; ++slot) {
    const int x = o[6 + 3 * slot], y = o[7 + 3 * slot], l = o[8 + 3 * slot];
    const int half = slot == 0 ? 0 : (slot + 2) / 3, off = slot == 0 ? 16 : 9 * ((slot - 1) % 3);
    h[half] |= ((uint32_t)(x + 1) | ((uint32_t)(y + 1) << 3) | ((uint32_t)l << 6)) << off;
  }
  w[0] = (uint64_t)h[0] | ((uint64_t)h[1] << 32);
  w[1] = (uint64_t)h[2] | ((uint64_t)h[3] << 32);
}

/* random-action model.step sweep, see drv_trace (oracle/c/driving_oracle.c) */
int64_t lbf_trace(uint64_t seed, uint64_t p0, int64_t n, int depth, uint64_t* hash_out, uint32_t* state_tr,
                  uint64_t* obs_tr, double* rew_tr, uint8_t* flags_tr, uint8_t* act_tr) {
  const TraceOut tr = {state_tr, obs_tr, rew_tr, flags_tr, act_tr};
  int64_t live = 0;
  for (int64_t i = 0; i < n; ++i) {
    const uint64_t particle = p0 + (uint64_t)i;
    LBF s, ns;
    initial_state(seed, particle, &s);
    uint64_t h = trace_hash_init(particle);
    for (int k = 0; k < depth; ++k) {
      int act[NPL], o[30];
      uint32_t sw[2];
      uint64_t ow[2 * NPL];
      double rew[NPL];
      int any = 0;
      for (int f = 0; f < NFOOD; ++f) any |= s.food[f] > 0;
      live += any;
      step_noise(seed, particle, (uint32_t)k, act);
      const int done = lbf_step1(&s, act, &ns, rew);
      for (int a = 0; a < NPL; ++a) { observe(&ns, a, o); pack_obs(o, ow + 2 * a); }
      pack_state(&ns, sw);
      const uint8_t flags = done ? 0x83 : 0;
      const uint8_t a8[NPL] = {(uint8_t)act[0], (uint8_t)act[1]};
      h = trace_hash_step(h, sw, 2, ow, 2 * NPL, rew, NPL, flags, a8);
      trace_record(&tr, n, i, k, sw, 2, ow, 2 * NPL, rew, NPL, flags, a8);
      s = ns;
    }
    hash_out[i] = h;
  }
  return live;
}

/* MCTS._rollout with Philox-driven uniform actions (oracle/rollout.py lbf_rollout) */
int64_t lbf_rollouts(uint64_t seed, uint64_t p0, int64_t n, int ego, int max_steps, double discount, double* returns,
                     uint8_t* steps, int8_t* final_state) {
  int64_t executed = 0;
  for (int64_t i = 0; i < n; ++i) {
    LBF s, ns;
    initial_state(seed, p0 + (uint64_t)i, &s);
    double ret = 0.0;
    int k = 0;
    while (k < max_steps) {
      int act[NPL]; double rew[NPL];
      step_noise(seed, p0 + (uint64_t)i, (uint32_t)k, act);
      const int done = lbf_step1(&s, act, &ns, rew);
      ret += pow(discount, (double)k) * rew[ego];
      s = ns; ++k; ++executed;
      if (done) break;
    }
    returns[i] = ret;
    if (steps) steps[i] = (uint8_t)k;
    if (final_state) store(&s, final_state + i * NFLD);
  }
  return executed;
}
