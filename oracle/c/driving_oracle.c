/* Driving-v1 CPU oracle in C (TEST INFRASTRUCTURE ONLY -- never linked into the product).
 *
 * A second, independent-in-form restatement of posggym's DrivingModel.step /
 * sample_initial_state / observation / reward (the published algorithm of the un-vendored
 * dependency posggym>=0.5.0, /root/reference/pyproject.toml:36), following oracle/driving.py
 * function by function: _get_next_state, _get_next_speed, _get_move_direction,
 * _get_next_coord, _get_rewards, _get_local_cell_obs, initial_state_from_noise.  It works on
 * unpacked field arrays (10 ints per vehicle: x, y, facing, speed, dest x, dest y,
 * dest_reached, crashed, min_dest_dist, init_dest_dist), not on the kernels' bit-packed
 * words, and is pinned against the Python oracle's golden vectors
 * (tests/test_oracle_native.py).  Used for (a) parity checks at sizes the Python oracle
 * cannot reach (millions of Philox-driven rollouts, compared bit for bit with the GPU) and
 * (b) a compiled-code CPU baseline beside the pure-Python one.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "oracle_common.h"

#define MAXC 8
#define NF 10 /* fields per vehicle */
enum { REVERSE = 0, STOPPED = 1, FORWARD_SLOW = 2, FORWARD_FAST = 3 };
enum { DO_NOTHING = 0, ACCELERATE = 1, DECELERATE = 2, TURN_RIGHT = 3, TURN_LEFT = 4 };
enum { VEHICLE = 0, WALL = 1, EMPTY = 2, DESTINATION = 3 };
enum { NO_COLLISION = 0, OBSTACLE_COLLISION = 1, VEHICLE_COLLISION = 2 };

typedef struct {
  int W, H, n_start, n_dest, obstacle_collisions;
  unsigned char blocked[16][16];
  int start[MAXC][2], dest[MAXC][2];
  int dist[MAXC][16][16]; /* shortest path distance to each destination, -1 unreachable */
} Grid;

typedef struct { int x, y, dir, speed, dx, dy, reached, crashed, mind, d0; } Veh;

static Grid G;

int drv_init(const char* rows, int obstacle_collisions) {
  memset(&G, 0, sizeof(G));
  int x = 0, y = 0;
  G.obstacle_collisions = obstacle_collisions;
  for (const char* p = rows; *p; ++p) {
    if (*p == '\n') { if (x > 0) { G.W = x; ++y; } x = 0; continue; }
    if (*p == '#') G.blocked[y][x] = 1;
    else if (*p == '+') { G.start[G.n_start][0] = x; G.start[G.n_start][1] = y; ++G.n_start; }
    else if (*p == '-') { G.dest[G.n_dest][0] = x; G.dest[G.n_dest][1] = y; ++G.n_dest; }
    ++x;
  }
  if (x > 0) { G.W = x; ++y; }
  G.H = y;
  for (int d = 0; d < G.n_dest; ++d) {
    for (int yy = 0; yy < 16; ++yy) for (int xx = 0; xx < 16; ++xx) G.dist[d][yy][xx] = -1;
    int qx[256], qy[256], head = 0, tail = 0;
    qx[tail] = G.dest[d][0]; qy[tail] = G.dest[d][1]; ++tail;
    G.dist[d][G.dest[d][1]][G.dest[d][0]] = 0;
    while (head < tail) {
      const int cx = qx[head], cy = qy[head]; ++head;
      const int nb[4][2] = {{cx, cy - 1}, {cx + 1, cy}, {cx, cy + 1}, {cx - 1, cy}};
      for (int k = 0; k < 4; ++k) {
        const int nx = nb[k][0], ny = nb[k][1];
        if (nx < 0 || ny < 0 || nx >= G.W || ny >= G.H || G.blocked[ny][nx] || G.dist[d][ny][nx] >= 0) continue;
        G.dist[d][ny][nx] = G.dist[d][cy][cx] + 1;
        qx[tail] = nx; qy[tail] = ny; ++tail;
      }
    }
  }
  return G.W * 100 + G.H;
}

static int dest_index(int dx, int dy) {
  for (int d = 0; d < G.n_dest; ++d) if (G.dest[d][0] == dx && G.dest[d][1] == dy) return d;
  return 0;
}

/* grid.get_next_coord: stay when leaving the grid or entering a block */
static void next_coord(int x, int y, int dir, int* ox, int* oy) {
  int nx = x, ny = y;
  if (dir == 0) ny = y - 1 < 0 ? 0 : y - 1;
  else if (dir == 1) nx = x + 1 > G.W - 1 ? G.W - 1 : x + 1;
  else if (dir == 2) ny = y + 1 > G.H - 1 ? G.H - 1 : y + 1;
  else nx = x - 1 < 0 ? 0 : x - 1;
  if (G.blocked[ny][nx]) { nx = x; ny = y; }
  *ox = nx; *oy = ny;
}

static int next_speed(int a, int sp) {
  if (a == DO_NOTHING) return sp;
  if (a == TURN_LEFT || a == TURN_RIGHT) return sp == FORWARD_FAST ? FORWARD_SLOW : sp;
  if (a == ACCELERATE) return sp + 1 > FORWARD_FAST ? FORWARD_FAST : sp + 1;
  return sp - 1 < REVERSE ? REVERSE : sp - 1;
}
static int next_dir(int a, int dir) {
  if (a == TURN_RIGHT) return (dir + 1) % 4;
  if (a == TURN_LEFT) return (dir + 3) % 4;
  return dir;
}

/* model.step_with_noise for 2 vehicles; order_idx 0 -> (0, 1), 1 -> (1, 0) */
static void step2(const Veh* s, const int* act, int order_idx, Veh* ns, double* rew, int* coll) {
  ns[0] = s[0]; ns[1] = s[1];
  coll[0] = coll[1] = NO_COLLISION;
  /* vehicle_coords: cells of all vehicles; an entry is replaced when its vehicle moves */
  int vx[2] = {s[0].x, s[1].x}, vy[2] = {s[0].y, s[1].y};
  for (int k = 0; k < 2; ++k) {
    const int idx = order_idx ? 1 - k : k;
    const Veh si = s[idx];
    if (si.reached || ns[idx].crashed) continue;
    const int a = act[idx];
    int nsp = next_speed(a, si.speed);
    const int mdir = nsp == REVERSE ? (si.dir + 2) % 4 : next_dir(a, si.dir);
    const int ndir = nsp == REVERSE ? si.dir : next_dir(a, si.dir);
    int cx = si.x, cy = si.y, c = NO_COLLISION, hit = -1;
    const int moves = abs(nsp - STOPPED);
    for (int m = 0; m < moves; ++m) {
      int nx, ny;
      next_coord(cx, cy, mdir, &nx, &ny);
      if (nx == cx && ny == cy) { c = OBSTACLE_COLLISION; break; }
      int found = -1;
      for (int j = 0; j < 2; ++j) if (j != idx && vx[j] == nx && vy[j] == ny) found = j;
      if (found >= 0) { c = VEHICLE_COLLISION; hit = found; break; }
      cx = nx; cy = ny;
    }
    if (c == VEHICLE_COLLISION) { coll[hit] = VEHICLE_COLLISION; ns[hit].speed = STOPPED; ns[hit].crashed = 1; }
    const int crashed = c == VEHICLE_COLLISION || (c == OBSTACLE_COLLISION && G.obstacle_collisions);
    coll[idx] = c;
    if (crashed) nsp = STOPPED;
    const int d = G.dist[dest_index(si.dx, si.dy)][cy][cx];
    Veh o = si;
    o.x = cx; o.y = cy; o.dir = ndir; o.speed = nsp;
    o.reached = (cx == si.dx && cy == si.dy);
    o.crashed = crashed;
    o.mind = d < si.mind ? d : si.mind;
    ns[idx] = o;
    vx[idx] = cx; vy[idx] = cy;
  }
  for (int i = 0; i < 2; ++i) {
    if (s[i].crashed || s[i].reached) { rew[i] = 0.0; continue; }
    double r = 0.0;
    if (coll[i] == VEHICLE_COLLISION || (G.obstacle_collisions && coll[i] == OBSTACLE_COLLISION)) r = -1.0;
    else if (ns[i].reached) r = 0.5;
    int p = s[i].mind - ns[i].mind;
    if (p < 0) p = 0;
    const int d0 = s[i].d0 < 1 ? 1 : s[i].d0;
    r += 0.5 * ((double)p / (double)d0);
    rew[i] = r;
  }
}

/* _get_local_cell_obs: 5x3 window, row-major, rows from 3 ahead to 1 behind */
static void observe(const Veh* s, int i, int* out /* 22 ints */) {
  const Veh v = s[i];
  int k = 0;
  for (int row = 0; row < 5; ++row)
    for (int col = 0; col < 3; ++col) {
      int gx, gy;
      if (v.dir == 0) { gy = v.y + row - 3; gx = v.x + col - 1; }
      else if (v.dir == 1) { gy = v.y + col - 1; gx = v.x - row + 3; }
      else if (v.dir == 2) { gy = v.y - row + 3; gx = v.x - col + 1; }
      else { gy = v.y - col + 1; gx = v.x + row - 3; }
      int code;
      if (gx < 0 || gy < 0 || gx >= G.W || gy >= G.H || G.blocked[gy][gx]) code = WALL;
      else if ((s[0].x == gx && s[0].y == gy) || (s[1].x == gx && s[1].y == gy)) code = VEHICLE;
      else if (gx == v.dx && gy == v.dy) code = DESTINATION;
      else code = EMPTY;
      out[k++] = code;
    }
  out[15] = v.speed; out[16] = v.x; out[17] = v.y; out[18] = v.dx; out[19] = v.dy;
  out[20] = v.reached; out[21] = v.crashed;
}

static void load(const int* f, Veh* v) {
  v->x = f[0]; v->y = f[1]; v->dir = f[2]; v->speed = f[3]; v->dx = f[4]; v->dy = f[5];
  v->reached = f[6]; v->crashed = f[7]; v->mind = f[8]; v->d0 = f[9];
}
static void store(const Veh* v, int* f) {
  f[0] = v->x; f[1] = v->y; f[2] = v->dir; f[3] = v->speed; f[4] = v->dx; f[5] = v->dy;
  f[6] = v->reached; f[7] = v->crashed; f[8] = v->mind; f[9] = v->d0;
}

/* batch step on field arrays: state [n][2][10], actions [n][2], order [n] ->
 * next [n][2][10], obs [n][2][22], rewards [n][2], flags [n] (bit0/1 terminated, bit7 all done) */
void drv_step(int64_t n, const int* state, const uint8_t* actions, const uint8_t* order, int* next, int* obs,
              double* rewards, uint8_t* flags) {
  for (int64_t i = 0; i < n; ++i) {
    Veh s[2], ns[2];
    load(state + i * 2 * NF, &s[0]); load(state + i * 2 * NF + NF, &s[1]);
    const int act[2] = {actions[2 * i], actions[2 * i + 1]};
    int coll[2];
    step2(s, act, order[i], ns, rewards + 2 * i, coll);
    store(&ns[0], next + i * 2 * NF); store(&ns[1], next + i * 2 * NF + NF);
    observe(ns, 0, obs + i * 44); observe(ns, 1, obs + i * 44 + 22);
    const int t0 = ns[0].reached || ns[0].crashed, t1 = ns[1].reached || ns[1].crashed;
    flags[i] = (uint8_t)(t0 | (t1 << 1) | ((t0 && t1) ? 0x80 : 0));
  }
}

/* ---- the kernels' noise conventions (include/posg_b200.h); Philox in oracle_common.h ---- */
static void initial_state(uint64_t seed, uint64_t particle, Veh* s) {
  uint32_t w[4];
  philox_block(seed, particle, 0u, STREAM_INIT, w);
  int used_s[MAXC] = {0}, used_d[MAXC] = {0};
  for (int i = 0; i < 2; ++i) {
    int navail = 0, sel = 0, k = 0;
    for (int c = 0; c < G.n_start; ++c) navail += !used_s[c];
    const int ps = (int)below(w[2 * i], (uint32_t)navail);
    for (int c = 0; c < G.n_start; ++c) { if (used_s[c]) continue; if (k == ps) sel = c; ++k; }
    used_s[sel] = 1;
    const int sx = G.start[sel][0], sy = G.start[sel][1];
    navail = 0;
    for (int c = 0; c < G.n_dest; ++c) navail += (!used_d[c] && !(G.dest[c][0] == sx && G.dest[c][1] == sy));
    const int pd = (int)below(w[2 * i + 1], (uint32_t)navail);
    int dsel = 0; k = 0;
    for (int c = 0; c < G.n_dest; ++c) {
      if (used_d[c] || (G.dest[c][0] == sx && G.dest[c][1] == sy)) continue;
      if (k == pd) dsel = c;
      ++k;
    }
    used_d[dsel] = 1;
    Veh v = {sx, sy, 0, STOPPED, G.dest[dsel][0], G.dest[dsel][1], 0, 0, G.dist[dsel][sy][sx], G.dist[dsel][sy][sx]};
    s[i] = v;
  }
}

static void step_noise(uint64_t seed, uint64_t particle, uint32_t step, int* act, int* order) {
  uint32_t w[4];
  const uint64_t blk = particle >> 2;
  philox_block(seed, blk, step, STREAM_STEP, w);
  const uint32_t idx = below(w[particle & 3], 50u);
  *order = (int)(idx & 1u);
  const uint32_t r = idx >> 1;
  act[1] = (int)(r / 5u);
  act[0] = (int)(r - 5u * (uint32_t)act[1]);
}

/* MCTS._rollout with Philox-driven uniform actions from Philox-drawn initial states, for
 * particles [p0, p0 + n): returns, steps, final state fields [n][2][10]; result = env-steps */
int64_t drv_rollouts(uint64_t seed, uint64_t p0, int64_t n, int ego, int max_steps, double discount,
                     double* returns, uint8_t* steps, int* final_state) {
  int64_t executed = 0;
  for (int64_t i = 0; i < n; ++i) {
    Veh s[2], ns[2];
    initial_state(seed, p0 + (uint64_t)i, s);
    double ret = 0.0;
    int k = 0;
    while (k < max_steps) {
      int act[2], order, coll[2];
      double rew[2];
      step_noise(seed, p0 + (uint64_t)i, (uint32_t)k, act, &order);
      step2(s, act, order, ns, rew, coll);
      ret += pow(discount, (double)k) * rew[ego];
      s[0] = ns[0]; s[1] = ns[1];
      ++k; ++executed;
      const int te = ns[ego].reached || ns[ego].crashed;
      const int all = (ns[0].reached || ns[0].crashed) && (ns[1].reached || ns[1].crashed);
      if (te || all) break;
    }
    returns[i] = ret;
    if (steps) steps[i] = (uint8_t)k;
    if (final_state) { store(&s[0], final_state + i * 2 * NF); store(&s[1], final_state + i * 2 * NF + NF); }
  }
  return executed;
}

/* random-action sweep for the compiled CPU baseline: episodes of up to `depth` steps from
 * fresh initial states until about `budget_steps` env-steps have been executed */
int64_t drv_sweep(uint64_t seed, int depth, int64_t budget_steps) {
  int64_t executed = 0;
  uint64_t particle = 0;
  volatile int sink = 0;
  while (executed < budget_steps) {
    Veh s[2], ns[2];
    initial_state(seed, particle, s);
    for (int k = 0; k < depth; ++k) {
      int act[2], order, coll[2], o[44];
      double rew[2];
      step_noise(seed, particle, (uint32_t)k, act, &order);
      step2(s, act, order, ns, rew, coll);
      observe(ns, 0, o); observe(ns, 1, o + 22);   /* model.step returns observations too */
      sink += o[7] + o[29];
      s[0] = ns[0]; s[1] = ns[1];
      ++executed;
      if ((ns[0].reached || ns[0].crashed) && (ns[1].reached || ns[1].crashed)) break;
    }
    ++particle;
  }
  return executed;
}

/* ---- packed words of include/posg_b200.h (what the kernels read and write) and the trace hash ---- */
static uint32_t pack_vehicle(const Veh* v) {
  return (uint32_t)v->x | ((uint32_t)v->y << 4) | ((uint32_t)v->dir << 8) | ((uint32_t)v->speed << 10) |
         ((uint32_t)dest_index(v->dx, v->dy) << 12) | ((uint32_t)v->reached << 15) | ((uint32_t)v->crashed << 16) |
         ((uint32_t)v->mind << 17) | ((uint32_t)v->d0 << 24);
}
static uint64_t pack_obs(const int* o) {
  uint64_t w = 0;
  for (int c = 0; c < 15; ++c) w |= (uint64_t)o[c] << (2 * c);
  w |= (uint64_t)o[15] << 30;
  w |= (uint64_t)o[16] << 32; w |= (uint64_t)o[17] << 36; w |= (uint64_t)o[18] << 40; w |= (uint64_t)o[19] << 44;
  w |= (uint64_t)o[20] << 48; w |= (uint64_t)o[21] << 49;
  return w;
}

/* BASELINE.json configs[1]: the random-action model.step sweep from Philox-drawn initial states, every
 * particle stepped `depth` times whether or not it is done (terminal states are absorbing), for
 * particles [p0, p0 + n).  hash_out[i] = rolling hash over every step's packed outputs
 * (oracle_common.h); the optional arrays receive the full per-step trace, [depth][n][..].
 * Returns the env-steps executed on particles that were not yet all-done. */
int64_t drv_trace(uint64_t seed, uint64_t p0, int64_t n, int depth, uint64_t* hash_out, uint32_t* state_tr,
                  uint64_t* obs_tr, double* rew_tr, uint8_t* flags_tr, uint8_t* act_tr) {
  const TraceOut tr = {state_tr, obs_tr, rew_tr, flags_tr, act_tr};
  int64_t live = 0;
  for (int64_t i = 0; i < n; ++i) {
    const uint64_t particle = p0 + (uint64_t)i;
    Veh s[2], ns[2];
    initial_state(seed, particle, s);
    uint64_t h = trace_hash_init(particle);
    for (int k = 0; k < depth; ++k) {
      int act[2], order, coll[2], o[44];
      double rew[2];
      live += !((s[0].reached || s[0].crashed) && (s[1].reached || s[1].crashed));
      step_noise(seed, particle, (uint32_t)k, act, &order);
      step2(s, act, order, ns, rew, coll);
      observe(ns, 0, o); observe(ns, 1, o + 22);
      const int t0 = ns[0].reached || ns[0].crashed, t1 = ns[1].reached || ns[1].crashed;
      const uint8_t flags = (uint8_t)(t0 | (t1 << 1) | ((t0 && t1) ? 0x80 : 0));
      const uint32_t sw[2] = {pack_vehicle(&ns[0]), pack_vehicle(&ns[1])};
      const uint64_t ow[2] = {pack_obs(o), pack_obs(o + 22)};
      const uint8_t a8[2] = {(uint8_t)act[0], (uint8_t)act[1]};
      h = trace_hash_step(h, sw, 2, ow, 2, rew, 2, flags, a8);
      trace_record(&tr, n, i, k, sw, 2, ow, 2, rew, 2, flags, a8);
      s[0] = ns[0]; s[1] = ns[1];
    }
    hash_out[i] = h;
  }
  return live;
}
