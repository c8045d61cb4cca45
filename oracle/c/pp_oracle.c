/* PredatorPrey-v0 CPU oracle in C (TEST INFRASTRUCTURE ONLY -- never linked into the product).
 *
 * A second restatement, independent in form, of posggym's PredatorPreyModel (the published
 * algorithm of the un-vendored dependency posggym>=0.5.0, /root/reference/pyproject.toml:36),
 * following oracle/predator_prey.py function by function: _get_next_prey_state, _move_away,
 * _get_next_predator_state, _get_next_prey_caught, _get_rewards, _get_local_cell_obs,
 * initial_state_from_noise.  It works on coordinate structs with explicit candidate lists (as the
 * Python oracle does), not on the kernels' packed words and tables, and is pinned against the
 * Python oracle's golden vectors (tests/test_oracle_native.py).  Used for parity checks at sizes
 * the Python oracle cannot reach: full per-step traces for 2^16 particles x 50 steps and the
 * rolling trace hash (oracle_common.h) for millions.
 *
 * Scope: the block-free "NxN" grids, 2 predators, 3 prey (the reference's experiment kwargs,
 * baseline_exps/env_data/PredatorPrey-v0/env_kwargs.yaml).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "oracle_common.h"

#define NP 2
#define NQ 3
#define NFLD 13 /* p0x p0y p1x p1y q0x q0y q1x q1y q2x q2y c0 c1 c2 */
enum { DO_NOTHING = 0, UP = 1, DOWN = 2, LEFT = 3, RIGHT = 4 };
enum { EMPTY = 0, WALL = 1, PREDATOR = 2, PREY = 3 };

typedef struct { int x, y; } Cell;
typedef struct { Cell pred[NP]; Cell prey[NQ]; int caught[NQ]; } PP;

typedef struct {
  int size, cooperative, prey_strength, obs_dim;
  int n_pred_start, n_prey_start;
  Cell pred_start[16], prey_start[64];
  double per_prey_reward;
} Cfg;
static Cfg C;

static int in_bounds(Cell c) { return c.x >= 0 && c.y >= 0 && c.x < C.size && c.y < C.size; }
static int same(Cell a, Cell b) { return a.x == b.x && a.y == b.y; }
static int mdist(Cell a, Cell b) { return abs(a.x - b.x) + abs(a.y - b.y); }

int pp_init(int size, int cooperative, int prey_strength, int obs_dim) {
  memset(&C, 0, sizeof(C));
  C.size = size; C.cooperative = cooperative; C.prey_strength = prey_strength; C.obs_dim = obs_dim;
  C.per_prey_reward = 1.0 / (double)NQ;
  /* PPGrid: predators start on the corners and half-way along each side (itertools.product order) */
  const int v[3] = {0, size / 2, size - 1};
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) {
      const int x = v[a], y = v[b];
      if (x == 0 || x == size - 1 || y == 0 || y == size - 1) { C.pred_start[C.n_pred_start].x = x; C.pred_start[C.n_pred_start++].y = y; }
    }
  /* get_unblocked_center_coords(num_prey): whole rings around the centre until enough cells */
  const int cx = size / 2, cy = size / 2;
  for (int ring = 0; C.n_prey_start < NQ; ++ring)
    for (int x = cx - ring; x <= cx + ring; ++x)
      for (int y = cy - ring; y <= cy + ring; ++y) {
        const int ax = abs(x - cx), ay = abs(y - cy);
        if ((ax > ay ? ax : ay) != ring) continue;
        const Cell c = {x, y};
        if (in_bounds(c)) C.prey_start[C.n_prey_start++] = c;
      }
  return C.n_pred_start * 100 + C.n_prey_start;
}

/* grid.get_neighbours: N, E, S, W, inside the grid */
static int neighbours(Cell c, Cell* out) {
  int n = 0;
  if (c.y > 0) { out[n].x = c.x; out[n++].y = c.y - 1; }
  if (c.x < C.size - 1) { out[n].x = c.x + 1; out[n++].y = c.y; }
  if (c.y < C.size - 1) { out[n].x = c.x; out[n++].y = c.y + 1; }
  if (c.x > 0) { out[n].x = c.x - 1; out[n++].y = c.y; }
  return n;
}

typedef struct { Cell c[NP + NQ]; int n; } Occ;
static int occ_has(const Occ* o, Cell c) { for (int i = 0; i < o->n; ++i) if (same(o->c[i], c)) return 1; return 0; }
static void occ_move(Occ* o, Cell from, Cell to) {
  for (int i = 0; i < o->n; ++i) if (same(o->c[i], from)) { o->c[i] = to; return; }
}

/* _move_away: 1 and *out set when some threat is within obs_dim, else 0 */
static int move_away(Cell coord, const Cell* threats, int nt, const Occ* occ, uint32_t word, Cell* out) {
  if (nt == 0) return 0;
  int dmin = 1 << 20;
  for (int i = 0; i < nt; ++i) { const int d = mdist(coord, threats[i]); if (d < dmin) dmin = d; }
  if (dmin > C.obs_dim) return 0;
  Cell closest[NP + NQ]; int nc = 0;
  for (int i = 0; i < nt; ++i) if (mdist(coord, threats[i]) == dmin) closest[nc++] = threats[i];
  const Cell threat = closest[below(word, (uint32_t)nc)];
  /* largest (distance, (x, y)) among staying and the free neighbours */
  Cell best = coord; int bd = mdist(coord, threat);
  Cell nb[4];
  const int nn = neighbours(coord, nb);
  for (int k = 0; k < nn; ++k) {
    if (occ_has(occ, nb[k])) continue;
    const int d = mdist(nb[k], threat);
    if (d > bd || (d == bd && (nb[k].x > best.x || (nb[k].x == best.x && nb[k].y > best.y)))) { best = nb[k]; bd = d; }
  }
  *out = best;
  return 1;
}

static void next_prey(const PP* s, const uint32_t* words, Cell* next) {
  int set[NQ] = {0, 0, 0};
  Occ occ; occ.n = 0;
  for (int i = 0; i < NP; ++i) occ.c[occ.n++] = s->pred[i];
  for (int i = 0; i < NQ; ++i) if (!s->caught[i]) occ.c[occ.n++] = s->prey[i];
  int n_unc = 0;
  for (int i = 0; i < NQ; ++i) n_unc += !s->caught[i];
  /* 1. flee from the closest predator in sight */
  for (int i = 0; i < NQ; ++i) {
    if (s->caught[i]) { next[i] = s->prey[i]; set[i] = 1; continue; }
    Cell nx;
    if (move_away(s->prey[i], s->pred, NP, &occ, words[i], &nx)) { next[i] = nx; set[i] = 1; occ_move(&occ, s->prey[i], nx); }
  }
  /* 2. keep away from the closest other prey in sight */
  if (n_unc > 1)
    for (int i = 0; i < NQ; ++i) {
      if (set[i]) continue;
      Cell others[NQ]; int no = 0;
      for (int j = 0; j < NQ; ++j) if (j != i && !s->caught[j]) others[no++] = s->prey[j];
      Cell nx;
      if (move_away(s->prey[i], others, no, &occ, words[i], &nx)) { next[i] = nx; set[i] = 1; occ_move(&occ, s->prey[i], nx); }
    }
  /* 3. nobody in sight: uniformly among the free neighbours and staying */
  for (int i = 0; i < NQ; ++i) {
    if (set[i]) continue;
    Cell nb[4], cands[5]; int nc = 0;
    const int nn = neighbours(s->prey[i], nb);
    for (int k = 0; k < nn; ++k) if (!occ_has(&occ, nb[k])) cands[nc++] = nb[k];
    cands[nc++] = s->prey[i];
    const Cell nx = cands[below(words[i], (uint32_t)nc)];
    next[i] = nx; set[i] = 1; occ_move(&occ, s->prey[i], nx);
  }
}

static void next_pred(const PP* s, const int* act, const Cell* nprey, Cell* out) {
  Cell pot[NP];
  for (int i = 0; i < NP; ++i) {
    Cell c = s->pred[i];
    const int a = act[i];
    if (a == UP) c.y = c.y - 1 < 0 ? 0 : c.y - 1;
    else if (a == DOWN) c.y = c.y + 1 > C.size - 1 ? C.size - 1 : c.y + 1;
    else if (a == LEFT) c.x = c.x - 1 < 0 ? 0 : c.x - 1;
    else if (a == RIGHT) c.x = c.x + 1 > C.size - 1 ? C.size - 1 : c.x + 1;
    for (int j = 0; j < NQ; ++j) if (!s->caught[j] && same(nprey[j], c)) { c = s->pred[i]; break; }
    pot[i] = c;
  }
  for (int i = 0; i < NP; ++i) {
    int clash = 0;
    for (int j = 0; j < NP; ++j) if (j != i && same(pot[j], pot[i])) clash = 1;
    out[i] = clash ? s->pred[i] : pot[i];
  }
}

/* model.step_with_noise */
static int pp_step1(const PP* s, const int* act, const uint32_t* words, PP* ns, double* rew) {
  next_prey(s, words, ns->prey);
  next_pred(s, act, ns->prey, ns->pred);
  for (int i = 0; i < NQ; ++i) {
    if (s->caught[i]) { ns->caught[i] = 1; continue; }
    int adj = 0;
    for (int j = 0; j < NP; ++j) adj += mdist(ns->prey[i], ns->pred[j]) <= 1;
    ns->caught[i] = adj >= C.prey_strength;
  }
  for (int j = 0; j < NP; ++j) rew[j] = 0.0;
  int n_new = 0;
  for (int i = 0; i < NQ; ++i) n_new += (!s->caught[i] && ns->caught[i]);
  if (n_new) {
    if (C.cooperative) {
      const double r = (double)n_new * C.per_prey_reward;
      for (int j = 0; j < NP; ++j) rew[j] = r;
    } else {
      for (int i = 0; i < NQ; ++i) {
        if (s->caught[i] || !ns->caught[i]) continue;
        int inv = 0;
        for (int j = 0; j < NP; ++j) inv += mdist(ns->prey[i], ns->pred[j]) <= 1;
        for (int j = 0; j < NP; ++j) if (mdist(ns->prey[i], ns->pred[j]) <= 1) rew[j] += C.per_prey_reward / (double)inv;
      }
    }
  }
  return ns->caught[0] && ns->caught[1] && ns->caught[2];
}

/* _get_local_cell_obs: (2 obs_dim + 1)^2 cells, row-major (index = row * w + col, row = y offset) */
static void observe(const PP* s, int idx, int* out) {
  const int w = 2 * C.obs_dim + 1;
  int k = 0;
  for (int row = 0; row < w; ++row)
    for (int col = 0; col < w; ++col) {
      const Cell c = {s->pred[idx].x + col - C.obs_dim, s->pred[idx].y + row - C.obs_dim};
      int code = EMPTY;
      if (!in_bounds(c)) code = WALL;
      else {
        for (int j = 0; j < NQ; ++j) if (!s->caught[j] && same(s->prey[j], c)) code = PREY;
        for (int j = 0; j < NP; ++j) if (same(s->pred[j], c)) code = PREDATOR;
      }
      out[k++] = code;
    }
}

static void load(const int8_t* f, PP* s) {
  for (int i = 0; i < NP; ++i) { s->pred[i].x = f[2 * i]; s->pred[i].y = f[2 * i + 1]; }
  for (int i = 0; i < NQ; ++i) { s->prey[i].x = f[4 + 2 * i]; s->prey[i].y = f[5 + 2 * i]; s->caught[i] = f[10 + i]; }
}
static void store(const PP* s, int8_t* f) {
  for (int i = 0; i < NP; ++i) { f[2 * i] = (int8_t)s->pred[i].x; f[2 * i + 1] = (int8_t)s->pred[i].y; }
  for (int i = 0; i < NQ; ++i) { f[4 + 2 * i] = (int8_t)s->prey[i].x; f[5 + 2 * i] = (int8_t)s->prey[i].y; f[10 + i] = (int8_t)s->caught[i]; }
}

/* batch step on field arrays: state [n][13] i8, actions [n][2] u8, noise [n][3] u32 ->
 * next [n][13], obs [n][2][25] i8, rewards [n][2], all_done [n] */
void pp_step(int64_t n, const int8_t* state, const uint8_t* actions, const uint32_t* noise, int8_t* next, int8_t* obs,
             double* rewards, uint8_t* all_done) {
  for (int64_t i = 0; i < n; ++i) {
    PP s, ns;
    load(state + i * NFLD, &s);
    const int act[NP] = {actions[2 * i], actions[2 * i + 1]};
    all_done[i] = (uint8_t)pp_step1(&s, act, noise + 3 * i, &ns, rewards + 2 * i);
    store(&ns, next + i * NFLD);
    int o[25];
    for (int a = 0; a < NP; ++a) { observe(&ns, a, o); for (int k = 0; k < 25; ++k) obs[(i * NP + a) * 25 + k] = (int8_t)o[k]; }
  }
}

/* ---- the kernels' Philox conventions (oracle/rollout.py pp_step_noise / pp_initial_picks) ---- */
static void initial_state(uint64_t seed, uint64_t particle, PP* s) {
  uint32_t a[4], b[4];
  philox_block(seed, particle, 0u, STREAM_INIT, a);
  philox_block(seed, particle, 0u, STREAM_INIT2, b);
  Cell pool[64]; int np = C.n_pred_start;
  memcpy(pool, C.pred_start, sizeof(Cell) * np);
  const uint32_t pw[NP] = {a[0], a[1]};
  for (int i = 0; i < NP; ++i) {
    const int k = (int)below(pw[i], (uint32_t)np);
    s->pred[i] = pool[k];
    memmove(pool + k, pool + k + 1, sizeof(Cell) * (np - k - 1)); --np;     /* list.pop(k) */
  }
  int nq = C.n_prey_start;
  memcpy(pool, C.prey_start, sizeof(Cell) * nq);
  const uint32_t qw[NQ] = {b[0], b[1], b[2]};
  for (int i = 0; i < NQ; ++i) {
    const int k = (int)below(qw[i], (uint32_t)nq);
    s->prey[i] = pool[k];
    memmove(pool + k, pool + k + 1, sizeof(Cell) * (nq - k - 1)); --nq;
    s->caught[i] = 0;
  }
}
/* half a Philox block per (particle, step): particles g and g + 32 (bit 5 of g clear) share the block with counter
 * ((g >> 6) << 5 | (g & 31)); see oracle/rollout.py:pp_step_noise */
static void step_noise(uint64_t seed, uint64_t particle, uint32_t step, int* act, uint32_t* words) {
  uint32_t w[4];
  philox_block(seed, ((particle >> 6) << 5) | (particle & 31u), step, STREAM_STEP, w);
  const int hi = (int)((particle >> 5) & 1u);
  const uint32_t a = hi ? w[2] : w[0], b = hi ? w[3] : w[1];
  words[0] = a; words[1] = a * 0x10001u; words[2] = b;
  const uint32_t idx = below(b * 0x10001u, 25u);
  act[1] = (int)(idx / 5u);
  act[0] = (int)(idx - 5u * (uint32_t)act[1]);
}

/* ---- packed words of include/posg_b200.h ---- */
static void pack_state(const PP* s, uint32_t* w) {
  w[0] = (uint32_t)(s->pred[0].x | (s->pred[0].y << 4)) | ((uint32_t)(s->pred[1].x | (s->pred[1].y << 4)) << 8) |
         ((uint32_t)(s->caught[0] | (s->caught[1] << 1) | (s->caught[2] << 2)) << 16);
  w[1] = (uint32_t)(s->prey[0].x | (s->prey[0].y << 4)) | ((uint32_t)(s->prey[1].x | (s->prey[1].y << 4)) << 8) |
         ((uint32_t)(s->prey[2].x | (s->prey[2].y << 4)) << 16);
}
static uint64_t pack_obs(const int* o) {
  uint64_t w = 0;
  for (int k = 0; k < 25; ++k) w |= (uint64_t)o[k] << (2 * k);
  return w;
}

/* random-action model.step sweep, see drv_trace (oracle/c/driving_oracle.c) */
int64_t pp_trace(uint64_t seed, uint64_t p0, int64_t n, int depth, uint64_t* hash_out, uint32_t* state_tr,
                 uint64_t* obs_tr, double* rew_tr, uint8_t* flags_tr, uint8_t* act_tr) {
  const TraceOut tr = {state_tr, obs_tr, rew_tr, flags_tr, act_tr};
  int64_t live = 0;
  for (int64_t i = 0; i < n; ++i) {
    const uint64_t particle = p0 + (uint64_t)i;
    PP s, ns;
    initial_state(seed, particle, &s);
    uint64_t h = trace_hash_init(particle);
    for (int k = 0; k < depth; ++k) {
      int act[NP], o[25];
      uint32_t words[NQ], sw[2];
      uint64_t ow[NP];
      double rew[NP];
      live += !(s.caught[0] && s.caught[1] && s.caught[2]);
      step_noise(seed, particle, (uint32_t)k, act, words);
      const int done = pp_step1(&s, act, words, &ns, rew);
      for (int a = 0; a < NP; ++a) { observe(&ns, a, o); ow[a] = pack_obs(o); }
      pack_state(&ns, sw);
      const uint8_t flags = done ? 0x83 : 0;
      const uint8_t a8[NP] = {(uint8_t)act[0], (uint8_t)act[1]};
      h = trace_hash_step(h, sw, 2, ow, NP, rew, NP, flags, a8);
      trace_record(&tr, n, i, k, sw, 2, ow, NP, rew, NP, flags, a8);
      s = ns;
    }
    hash_out[i] = h;
  }
  return live;
}

/* MCTS._rollout with Philox-driven uniform actions (oracle/rollout.py pp_rollout) from Philox-drawn
 * initial states: returns, steps, final state fields [n][13]; result = env-steps executed */
int64_t pp_rollouts(uint64_t seed, uint64_t p0, int64_t n, int ego, int max_steps, double discount, double* returns,
                    uint8_t* steps, int8_t* final_state) {
  int64_t executed = 0;
  for (int64_t i = 0; i < n; ++i) {
    PP s, ns;
    initial_state(seed, p0 + (uint64_t)i, &s);
    double ret = 0.0;
    int k = 0;
    while (k < max_steps) {
      int act[NP]; uint32_t words[NQ]; double rew[NP];
      step_noise(seed, p0 + (uint64_t)i, (uint32_t)k, act, words);
      const int done = pp_step1(&s, act, words, &ns, rew);
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
