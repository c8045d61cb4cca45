// Driving-v1 on packed integer state: device-side transition, collision, reward and
// observation logic (one thread per env / particle).  Semantics follow posggym's
// DrivingModel.step as consumed at planning/mcts.py:333,418 and planning/belief.py:165
// of the reference; bit layouts are documented in include/posg_b200.h.
#pragma once
#include <stdint.h>

#include "philox.cuh"

namespace posg {
namespace driving {

constexpr int PAD = 4;            // wall padding around the grid in the bitmaps
constexpr int LINES = 24;         // 16 + 2*PAD
constexpr int MAX_CELLS = 8;      // max start / destination cells
constexpr int MAX_AGENTS = 4;
constexpr int MAX_D0 = 64;

enum : int { REVERSE = 0, STOPPED = 1, FORWARD_SLOW = 2, FORWARD_FAST = 3 };
enum : int { DO_NOTHING = 0, ACCELERATE = 1, DECELERATE = 2, TURN_RIGHT = 3, TURN_LEFT = 4 };
enum : int { VEHICLE = 0, WALL = 1, EMPTY = 2, DESTINATION = 3 };
enum : int { NO_COLLISION = 0, OBSTACLE_COLLISION = 1, VEHICLE_COLLISION = 2 };

// Immutable per-model data; staged into shared memory by one TMA bulk copy per CTA.
struct alignas(16) Blob {
  uint32_t rows[LINES];              // rows[y+PAD] bit (x+PAD): blocked or outside
  uint32_t cols[LINES];              // transposed: cols[x+PAD] bit (y+PAD)
  uint8_t dest_xy[MAX_CELLS];        // x | y << 4, parse (row-major) order
  uint8_t start_xy[MAX_CELLS];
  uint8_t dest_agents[MAX_CELLS];    // bit i: destination usable by agent i
  uint8_t start_agents[MAX_CELLS];
  uint8_t n_dest, n_start, width, height;
  uint8_t n_agents, obs_front, obs_back, obs_side;
  uint8_t obstacle_collisions, pad_[7];
  uint8_t dist[MAX_CELLS][256];      // dist[dest][y*16+x]; 255 = unreachable
  double progress[MAX_D0 * 3];       // 0.5 * (p / D0), p in 0..2
};
static_assert(sizeof(Blob) % 16 == 0, "blob must be a multiple of 16 B for cp.async.bulk");

// ---- field access ---------------------------------------------------------------
__host__ __device__ __forceinline__ int vx(uint32_t v) { return v & 15; }
__host__ __device__ __forceinline__ int vy(uint32_t v) { return (v >> 4) & 15; }
__host__ __device__ __forceinline__ int vxy(uint32_t v) { return v & 255; }
__host__ __device__ __forceinline__ int vdir(uint32_t v) { return (v >> 8) & 3; }
__host__ __device__ __forceinline__ int vspeed(uint32_t v) { return (v >> 10) & 3; }
__host__ __device__ __forceinline__ int vdest(uint32_t v) { return (v >> 12) & 7; }
__host__ __device__ __forceinline__ int vreached(uint32_t v) { return (v >> 15) & 1; }
__host__ __device__ __forceinline__ int vcrashed(uint32_t v) { return (v >> 16) & 1; }
__host__ __device__ __forceinline__ int vmindist(uint32_t v) { return (v >> 17) & 127; }
__host__ __device__ __forceinline__ int vinitdist(uint32_t v) { return (v >> 24) & 127; }
__host__ __device__ __forceinline__ bool vdone(uint32_t v) { return (v >> 15) & 3; }

__host__ __device__ __forceinline__ uint32_t vpack(int x, int y, int dir, int speed, int dest,
                                                   int reached, int crashed, int mind, int d0) {
  return (uint32_t)x | ((uint32_t)y << 4) | ((uint32_t)dir << 8) | ((uint32_t)speed << 10) |
         ((uint32_t)dest << 12) | ((uint32_t)reached << 15) | ((uint32_t)crashed << 16) |
         ((uint32_t)mind << 17) | ((uint32_t)d0 << 24);
}

__device__ __forceinline__ bool blocked(const Blob& B, int x, int y) {
  return (B.rows[y + PAD] >> (x + PAD)) & 1u;
}

// i-th permutation of 0..NA-1 in lexicographic order, one element per nibble
template <int NA>
__device__ __forceinline__ uint32_t order_nibbles(int idx) {
  if (NA == 1) return 0u;
  if (NA == 2) return idx ? 0x01u : 0x10u;  // nibble k = k-th vehicle processed
  // factorial number system
  uint32_t avail = 0x3210u, out = 0u;
  int f = 1;
#pragma unroll
  for (int k = 2; k < NA; ++k) f *= k;  // (NA-1)!
#pragma unroll
  for (int k = NA; k > 0; --k) {
    int q = idx / f;
    idx -= q * f;
    uint32_t pick = (avail >> (4 * q)) & 15u;
    out |= pick << (4 * (NA - k));
    // remove nibble q from avail
    uint32_t low = avail & ((1u << (4 * q)) - 1u);
    avail = low | ((avail >> (4 * (q + 1))) << (4 * q));
    if (k > 1) f /= (k - 1);
  }
  return out;
}

// ---- transition -------------------------------------------------------------------
// veh[] is updated in place; coll[] receives the per-vehicle collision type.
template <int NA>
__device__ __forceinline__ void transition(const Blob& B, uint32_t (&veh)[NA], const int (&act)[NA],
                                           int order_idx, int (&coll)[NA]) {
#pragma unroll
  for (int i = 0; i < NA; ++i) coll[i] = NO_COLLISION;
  const uint32_t order = order_nibbles<NA>(order_idx);
#pragma unroll
  for (int k = 0; k < NA; ++k) {
    const int idx = (NA == 1) ? 0 : (int)((order >> (4 * k)) & 15u);
#pragma unroll
    for (int i = 0; i < NA; ++i) {
      if (i != idx) continue;  // static indexing keeps veh[] in registers
      const uint32_t s = veh[i];
      if (vdone(s)) continue;  // reached earlier, crashed earlier, or hit earlier this step
      const int a = act[i];
      const int sp = vspeed(s), dir = vdir(s);
      int nsp = sp;
      if (a == ACCELERATE) nsp = min(sp + 1, (int)FORWARD_FAST);
      else if (a == DECELERATE) nsp = max(sp - 1, (int)REVERSE);
      else if (a >= TURN_RIGHT && sp == FORWARD_FAST) nsp = FORWARD_SLOW;
      int tdir = dir;
      if (a == TURN_RIGHT) tdir = (dir + 1) & 3;
      else if (a == TURN_LEFT) tdir = (dir + 3) & 3;
      const bool rev = (nsp == REVERSE);
      const int mdir = rev ? ((dir + 2) & 3) : tdir;
      const int ndir = rev ? dir : tdir;  // a reversing vehicle keeps its facing
      const int dx = (mdir == 1) - (mdir == 3);
      const int dy = (mdir == 2) - (mdir == 0);
      const int nmoves = nsp >= STOPPED ? nsp - STOPPED : 1;
      int cx = vx(s), cy = vy(s);
      int c = NO_COLLISION;
      for (int mv = 0; mv < nmoves; ++mv) {
        const int nx = cx + dx, ny = cy + dy;
        if (blocked(B, nx, ny)) { c = OBSTACLE_COLLISION; break; }
        const int nxy = nx | (ny << 4);
        bool hit = false;
#pragma unroll
        for (int j = 0; j < NA; ++j) {
          if (j != i && !hit && vxy(veh[j]) == nxy) {
            hit = true;
            coll[j] = VEHICLE_COLLISION;
            // hit vehicle: stopped + crashed
            veh[j] = (veh[j] & ~(3u << 10)) | ((uint32_t)STOPPED << 10) | (1u << 16);
          }
        }
        if (hit) { c = VEHICLE_COLLISION; break; }
        cx = nx; cy = ny;
      }
      const bool crashed = (c == VEHICLE_COLLISION) || (c == OBSTACLE_COLLISION && B.obstacle_collisions);
      coll[i] = c;
      if (crashed) nsp = STOPPED;
      const int dest = vdest(s);
      const int d = B.dist[dest][cx | (cy << 4)];
      const int mind = min(vmindist(s), d);
      const int reached = ((cx | (cy << 4)) == B.dest_xy[dest]);
      veh[i] = vpack(cx, cy, ndir, nsp, dest, reached, crashed ? 1 : 0, mind, vinitdist(s));
    }
  }
}

// reward of vehicle `i` given its state before / after the transition
__device__ __forceinline__ double reward(const Blob& B, uint32_t before, uint32_t after, int coll) {
  if (vdone(before)) return 0.0;
  double r = 0.0;
  if (coll == VEHICLE_COLLISION || (B.obstacle_collisions && coll == OBSTACLE_COLLISION)) r = -1.0;
  else if (vreached(after)) r = 0.5;
  int p = vmindist(before) - vmindist(after);
  p = p < 0 ? 0 : (p > 2 ? 2 : p);
  int d0 = vinitdist(before);
  d0 = d0 < 1 ? 1 : (d0 >= MAX_D0 ? MAX_D0 - 1 : d0);
  return __dadd_rn(r, B.progress[d0 * 3 + p]);
}

// ---- observation --------------------------------------------------------------------
// 3-bit wall mask -> three 2-bit cell codes (WALL=1, EMPTY=2); entry w at bits [6w, 6w+6).
// index bit 3 selects the lane-reversed variant (facing SOUTH / WEST).
__device__ __forceinline__ uint32_t expand3(uint32_t w, bool reversed) {
  // forward: cell c takes bit c of w
  constexpr unsigned long long FWD =
      (0x2AULL << 0) | (0x29ULL << 6) | (0x26ULL << 12) | (0x25ULL << 18) |
      (0x1AULL << 24) | (0x19ULL << 30) | (0x16ULL << 36) | (0x15ULL << 42);
  // reversed: cell c takes bit (2 - c) of w
  constexpr unsigned long long REV =
      (0x2AULL << 0) | (0x1AULL << 6) | (0x26ULL << 12) | (0x16ULL << 18) |
      (0x29ULL << 24) | (0x19ULL << 30) | (0x25ULL << 36) | (0x15ULL << 42);
  const unsigned long long lut = reversed ? REV : FWD;
  return (uint32_t)(lut >> (6u * w)) & 63u;
}

// window coordinates (row r: 0 = farthest ahead, col c: 0 = left) of grid offset
// (dx, dy) from the ego vehicle facing `dir`; returns cell index or -1 (FRONT/BACK/SIDE generic)
template <int FRONT, int BACK, int SIDE>
__device__ __forceinline__ int window_cell(int dx, int dy, int dir) {
  int r, c;
  if (dir == 0) { c = dx + SIDE; r = dy + FRONT; }
  else if (dir == 1) { r = FRONT - dx; c = dy + SIDE; }
  else if (dir == 2) { r = FRONT - dy; c = SIDE - dx; }
  else { r = dx + FRONT; c = SIDE - dy; }
  constexpr int DEPTH = FRONT + BACK + 1, WIDTH = 2 * SIDE + 1;
  return ((unsigned)r < (unsigned)DEPTH && (unsigned)c < (unsigned)WIDTH) ? r * WIDTH + c : -1;
}

// packed observation of vehicle `i` (see posg_b200.h); specialised for the 5x3 window
template <int NA>
__device__ __forceinline__ uint64_t observe(const Blob& B, const uint32_t (&veh)[NA], int i) {
  constexpr int FRONT = 3, BACK = 1, SIDE = 1, DEPTH = 5, WIDTH = 3;
  uint32_t me = 0;
#pragma unroll
  for (int j = 0; j < NA; ++j) if (j == i) me = veh[j];
  const int x = vx(me), y = vy(me), dir = vdir(me);
  const bool odd = dir & 1;
  const uint32_t* lines = odd ? B.cols : B.rows;
  const int p = (odd ? x : y) + PAD;          // centre line
  const int q = (odd ? y : x) + PAD - SIDE;   // lowest bit of the 3-bit slice
  const int s = (dir == 0 || dir == 3) ? 1 : -1;
  const bool rev = (dir >= 2);
  uint32_t cells = 0;
#pragma unroll
  for (int r = 0; r < DEPTH; ++r) {
    const uint32_t w = (lines[p + s * (r - FRONT)] >> q) & 7u;
    cells |= expand3(w, rev) << (2 * WIDTH * r);
  }
  // destination (never a wall), then vehicles (take priority)
  {
    const int d = B.dest_xy[vdest(me)];
    const int cidx = window_cell<FRONT, BACK, SIDE>((d & 15) - x, (d >> 4) - y, dir);
    if (cidx >= 0) cells |= 3u << (2 * cidx);
  }
#pragma unroll
  for (int j = 0; j < NA; ++j) {
    const int cidx = window_cell<FRONT, BACK, SIDE>(vx(veh[j]) - x, vy(veh[j]) - y, dir);
    if (cidx >= 0) cells &= ~(3u << (2 * cidx));
  }
  const uint32_t d = B.dest_xy[vdest(me)];
  uint64_t o = cells | ((uint64_t)vspeed(me) << 30);
  uint32_t hi = (uint32_t)x | ((uint32_t)y << 4) | ((d & 15u) << 8) | ((d >> 4) << 12) |
                ((uint32_t)vreached(me) << 16) | ((uint32_t)vcrashed(me) << 17);
  return o | ((uint64_t)hi << 32);
}

// ---- initial state --------------------------------------------------------------------
// picks[2*i], picks[2*i+1]: index into the still-available start / destination cells
template <int NA>
__device__ __forceinline__ void initial_state(const Blob& B, const int (&picks)[2 * NA],
                                              uint32_t (&veh)[NA]) {
  uint32_t used_start = 0, used_dest = 0;
#pragma unroll
  for (int i = 0; i < NA; ++i) {
    int sidx = 0, didx = 0, k = 0;
    for (int c = 0; c < B.n_start; ++c) {
      const bool ok = ((B.start_agents[c] >> i) & 1) && !((used_start >> c) & 1);
      if (ok && k == picks[2 * i]) sidx = c;
      k += ok;
    }
    used_start |= 1u << sidx;
    const int sxy = B.start_xy[sidx];
    k = 0;
    for (int c = 0; c < B.n_dest; ++c) {
      const bool ok = ((B.dest_agents[c] >> i) & 1) && !((used_dest >> c) & 1) && B.dest_xy[c] != sxy;
      if (ok && k == picks[2 * i + 1]) didx = c;
      k += ok;
    }
    used_dest |= 1u << didx;
    const int d0 = B.dist[didx][sxy];
    veh[i] = vpack(sxy & 15, sxy >> 4, 0 /*NORTH*/, STOPPED, didx, 0, 0, d0, d0);
  }
}

// number of choices for each pick given the earlier ones (mirrors initial_state)
template <int NA>
__device__ __forceinline__ void initial_state_random(const Blob& B, uint64_t seed, uint64_t particle,
                                                     uint32_t (&veh)[NA]) {
  uint32_t used_start = 0, used_dest = 0;
  uint4 w = philox_block(seed, particle, 0u, STREAM_INIT);
  uint4 w2 = w;
  if (NA > 2) w2 = philox_block(seed, particle, 0u, STREAM_INIT2);
#pragma unroll
  for (int i = 0; i < NA; ++i) {
    const uint32_t ws = (i == 0) ? w.x : (i == 1) ? w.z : (i == 2) ? w2.x : w2.z;
    const uint32_t wd = (i == 0) ? w.y : (i == 1) ? w.w : (i == 2) ? w2.y : w2.w;
    int navail = 0;
    for (int c = 0; c < B.n_start; ++c)
      navail += ((B.start_agents[c] >> i) & 1) && !((used_start >> c) & 1);
    const int ps = (int)uniform_below(ws, (uint32_t)navail);
    int sidx = 0, k = 0;
    for (int c = 0; c < B.n_start; ++c) {
      const bool ok = ((B.start_agents[c] >> i) & 1) && !((used_start >> c) & 1);
      if (ok && k == ps) sidx = c;
      k += ok;
    }
    used_start |= 1u << sidx;
    const int sxy = B.start_xy[sidx];
    navail = 0;
    for (int c = 0; c < B.n_dest; ++c)
      navail += ((B.dest_agents[c] >> i) & 1) && !((used_dest >> c) & 1) && B.dest_xy[c] != sxy;
    const int pd = (int)uniform_below(wd, (uint32_t)navail);
    int didx = 0;
    k = 0;
    for (int c = 0; c < B.n_dest; ++c) {
      const bool ok = ((B.dest_agents[c] >> i) & 1) && !((used_dest >> c) & 1) && B.dest_xy[c] != sxy;
      if (ok && k == pd) didx = c;
      k += ok;
    }
    used_dest |= 1u << didx;
    const int d0 = B.dist[didx][sxy];
    veh[i] = vpack(sxy & 15, sxy >> 4, 0, STOPPED, didx, 0, 0, d0, d0);
  }
}

// joint action + order noise of one step from one Philox block
template <int NA>
__device__ __forceinline__ void sample_step_noise(uint64_t seed, uint64_t particle, uint32_t step,
                                                  int (&act)[NA], int& order_idx) {
  const uint4 w = philox_block(seed, particle, step, STREAM_STEP);
  act[0] = (int)uniform_below(w.x, 5u);
  if (NA > 1) act[1] = (int)uniform_below(w.y, 5u);
  int nperm = 1;
#pragma unroll
  for (int k = 2; k <= NA; ++k) nperm *= k;
  order_idx = (int)uniform_below(w.z, (uint32_t)nperm);
  if (NA > 2) {
    const uint4 w2 = philox_block(seed, particle, step, STREAM_STEP2);
    act[2] = (int)uniform_below(w2.x, 5u);
    if (NA > 3) act[NA - 1] = (int)uniform_below(w2.y, 5u);
  }
}

}  // namespace driving
}  // namespace posg
