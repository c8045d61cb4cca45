// Driving-v1 on packed integer state: device-side transition, collision, reward and
// observation logic, two vehicles per env, branch-free (every lane of a warp executes
// the same instruction stream regardless of actions, speeds and collisions).
// Semantics follow posggym's DrivingModel.step as consumed at planning/mcts.py:333,418
// and planning/belief.py:165 of the reference; bit layouts: include/posg_b200.h.
#pragma once
#include <stdint.h>

#include "philox.cuh"

namespace posg {
namespace driving {

constexpr int PAD = 4;            // wall padding around the grid in the row bitmaps
constexpr int LINES = 24;         // 16 + 2*PAD
constexpr int MAX_CELLS = 8;      // max start / destination cells
constexpr int MAX_D0 = 64;
constexpr int NA = 2;             // vehicles per env handled by the kernels

enum : int { REVERSE = 0, STOPPED = 1, FORWARD_SLOW = 2, FORWARD_FAST = 3 };
enum : int { DO_NOTHING = 0, ACCELERATE = 1, DECELERATE = 2, TURN_RIGHT = 3, TURN_LEFT = 4 };
enum : int { VEHICLE = 0, WALL = 1, EMPTY = 2, DESTINATION = 3 };
enum : int { NO_COLLISION = 0, OBSTACLE_COLLISION = 1, VEHICLE_COLLISION = 2 };

// Immutable per-model data; staged into shared memory by one TMA bulk copy per CTA.
// Fixed header followed by the observation base table uint32 obs_base[n_dest][4][256].
struct alignas(16) Blob {
  uint32_t rows[LINES];              // rows[y+PAD] bit (x+PAD): blocked or outside
  uint8_t dest_xy[MAX_CELLS];        // x | y << 4, parse (row-major) order
  uint8_t start_xy[MAX_CELLS];
  uint8_t dest_agents[MAX_CELLS];    // bit i: destination usable by agent i
  uint8_t start_agents[MAX_CELLS];
  uint8_t n_dest, n_start, width, height;
  uint8_t n_agents, obs_front, obs_back, obs_side;
  uint8_t obstacle_collisions, pad_[7];
  uint8_t win_lut[4][64];            // [facing][(dy+3)*8 + (dx+3)] -> window cell 0..14, 15 = outside
  // control table [action*16 + speed*4 + facing] -> next facing | next speed << 2 | move2 row
  // (move dir * 3 + cells moved) << 4.  80 bytes = 20 words: bank-conflict free (a 16-bit table with
  // the fields pre-shifted saves two instructions per vehicle but spans 40 words: measured 71.8 %
  // instead of 76.9 % of the HBM roofline -- the L1 data pipe is the co-bottleneck).
  uint8_t ctrl2[80];
  // step noise table [floor(word * 50 / 2^32)] -> first mover's action | second mover's action
  // << 3 | (vehicle 1 moves first) << 6 | a0 << 7 | a1 << 10.  25 words: bank-conflict free.
  uint16_t noise_lut[56];
  // reward table [(D0 | progress << 6) * 4 + kind], kind 0 plain, 1 destination reached, 2 crash,
  // 3 already done (0.0): base + 0.5 * (progress / D0); index = (raw state word >> 24) * 4 + kind
  double reward_tab[MAX_D0 * 4 * 4];
  uint8_t dist[MAX_CELLS][256];      // dist[dest][y*16+x]; 255 = unreachable
  // move table [(move dir * 3 + cells moved) * 256 + y*16 + x], walls already resolved:
  //   bits 0..7   cell reached if no vehicle is in the way
  //   bits 8..15  cell after the first step (= start cell if that step is not made)
  //   bit 16 / 17 step 1 / step 2 enters a free road cell (a vehicle may be standing there)
  //   bit 18 / 19 step 1 / step 2 runs into a wall AND obstacle collisions count as crashes
  uint32_t move2[12 * 256];
  // 30-bit local-cell codes of a lone vehicle (walls, own destination, ego cell):
  // obs_base[(dest*4 + facing) * 256 + y*16 + x].  (A 64-bit entry that also carried the static half
  // of the observation's high word was measured: 68 % instead of 77 % of the HBM roofline -- the
  // 8-byte random lookups double the shared-memory wavefronts.)
  uint32_t obs_base[1];
};
constexpr size_t BLOB_HEADER_BYTES = offsetof(Blob, obs_base);
static_assert(BLOB_HEADER_BYTES % 16 == 0, "header must keep obs_base 16-byte aligned");
__host__ __device__ inline size_t blob_bytes(int n_dest) { return BLOB_HEADER_BYTES + (size_t)n_dest * 4 * 256 * 4; }

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
__host__ __device__ __forceinline__ bool vdone(uint32_t v) { return ((v >> 15) & 3u) != 0; }

__host__ __device__ __forceinline__ uint32_t vpack(int x, int y, int dir, int speed, int dest,
                                                   int reached, int crashed, int mind, int d0) {
  return (uint32_t)x | ((uint32_t)y << 4) | ((uint32_t)dir << 8) | ((uint32_t)speed << 10) |
         ((uint32_t)dest << 12) | ((uint32_t)reached << 15) | ((uint32_t)crashed << 16) |
         ((uint32_t)mind << 17) | ((uint32_t)d0 << 24);
}

// ---- transition -------------------------------------------------------------------
// One vehicle's move.  `other` is the only other vehicle (its current cell blocks).
// Speed / heading arithmetic and the wall tests are two table lookups (shared memory); what
// remains is the interaction with the other vehicle.  a16 = action << 4.
__device__ __forceinline__ void move_vehicle(const Blob& B, uint32_t& me, uint32_t& other, uint32_t a16) {
  const uint32_t c = B.ctrl2[a16 | ((me >> 8) & 15u)];
  const uint32_t xy0 = me & 255u, oxy = other & 255u;
  const uint32_t e = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const unsigned char*>(B.move2) +
                                                        (((c & 0xF0u) << 6) + (xy0 << 2)));
  const uint32_t cf = e & 255u, c1 = (e >> 8) & 255u;
  const bool active = (me & 0x18000u) == 0u;   // not reached / crashed earlier, nor hit earlier this step
  const bool h1 = (e & 0x10000u) && (c1 == oxy);
  const bool h2 = !h1 && (e & 0x20000u) && (cf == oxy);
  const bool hit_any = h1 || h2;
  const uint32_t xy = h1 ? xy0 : (h2 ? c1 : cf);
  const bool crashed = hit_any || (e & 0x40000u) || (!h1 && (e & 0x80000u));
  const uint32_t dirsp = crashed ? (((c & 3u) << 8) | ((uint32_t)STOPPED << 10)) : ((c & 15u) << 8);
  const uint32_t d = *(reinterpret_cast<const unsigned char*>(B.dist) + (((me >> 4) & 0x700u) | xy));
  const uint32_t mind0 = (me >> 17) & 127u, mind = min(mind0, d);
  // bits 30..31 (transient, see STATE_MASK): progress made in this step, 0..2 cells
  uint32_t moved = xy | dirsp | (me & 0x3F007000u) | (mind << 17) | ((mind0 - mind) << 30);
  moved |= (d == 0u) ? 0x8000u : 0u;           // distance 0 <=> on the own destination
  moved |= crashed ? 0x10000u : 0u;
  me = active ? moved : me;
  // the vehicle that was run into: stopped + crashed
  other = (active && hit_any) ? ((other & ~(3u << 10)) | ((uint32_t)STOPPED << 10) | (1u << 16)) : other;
}

// The words transition16 leaves behind carry the progress of the step in bits 30..31 (read by
// reward_raw); the stored state is `word & STATE_MASK`.
constexpr uint32_t STATE_MASK = 0x3FFFFFFFu;

// joint transition with pre-shifted actions: af16 / as16 = action << 4 of the vehicle that moves
// first / second, sw = vehicle 1 moves first.  Leaves RAW words (see STATE_MASK).
__device__ __forceinline__ void transition16(const Blob& B, uint32_t& v0, uint32_t& v1, uint32_t af16, uint32_t as16,
                                             bool sw) {
  uint32_t first = sw ? v1 : v0, second = sw ? v0 : v1;
  move_vehicle(B, first, second, af16);
  move_vehicle(B, second, first, as16);
  v0 = sw ? second : first; v1 = sw ? first : second;
}

// joint transition of both vehicles; order_idx 0 -> vehicle 0 first, 1 -> vehicle 1 first.
// Leaves RAW words (see STATE_MASK).
__device__ __forceinline__ void transition_raw(const Blob& B, uint32_t& v0, uint32_t& v1, int a0, int a1,
                                               int order_idx) {
  const bool sw = order_idx != 0;
  transition16(B, v0, v1, (uint32_t)(sw ? a1 : a0) << 4, (uint32_t)(sw ? a0 : a1) << 4, sw);
}

// Reward of a vehicle given its state before the transition and its RAW word after it.  For a
// vehicle that was not done before, the crashed bit after the step is set exactly when it
// crashed in this step -- as the mover (vehicle collision, or obstacle collision if those
// count) or as the vehicle that was run into -- which is the reference's crash-reward condition.
__device__ __forceinline__ double reward_raw(const Blob& B, uint32_t before, uint32_t after_raw) {
  const uint32_t fl = min((after_raw >> 15) & 3u, 2u);            // 0 plain, 1 reached, 2 crashed (wins)
  const uint32_t kind = (before & 0x18000u) ? 3u : fl;
  return B.reward_tab[((after_raw >> 24) << 2) | kind];
}

// the part of reward_raw's table index that is not in the stored state word: progress | kind << 2
__device__ __forceinline__ uint32_t reward_code(uint32_t before, uint32_t after_raw) {
  const uint32_t fl = min((after_raw >> 15) & 3u, 2u);
  const uint32_t kind = (before & 0x18000u) ? 3u : fl;
  return (after_raw >> 30) | (kind << 2);
}

// ---- observation --------------------------------------------------------------------
// `cidx` = window cell (0..14) the other vehicle occupies, 15 = outside the window
__device__ __forceinline__ uint64_t observe_with(const Blob& B, uint32_t me, int cidx) {
  const int dest = vdest(me);
  uint32_t cells = B.obs_base[(((me >> 10) & 0x1Cu) | ((me >> 8) & 3u)) << 8 | (me & 255u)];
  cells &= ~(3u << (2 * cidx));                                    // VEHICLE = 0 (shift 30: no-op)
  const uint32_t lo = cells | ((uint32_t)vspeed(me) << 30);
  const uint32_t hi = (uint32_t)vxy(me) | ((uint32_t)B.dest_xy[dest] << 8) | (((me >> 15) & 3u) << 16);
  return (uint64_t)lo | ((uint64_t)hi << 32);
}

__device__ __forceinline__ uint64_t observe(const Blob& B, uint32_t me, uint32_t other) {
  const int dx = vx(other) - vx(me) + 3, dy = vy(other) - vy(me) + 3;
  const bool inside = ((unsigned)dx < 7u) && ((unsigned)dy < 7u);
  return observe_with(B, me, B.win_lut[vdir(me)][inside ? (dy * 8 + dx) : 63]);
}

// both vehicles' observations: the grid offset is computed once (vehicle 1 sees the negated one)
__device__ __forceinline__ void observe_pair(const Blob& B, uint32_t v0, uint32_t v1, uint64_t& o0, uint64_t& o1) {
  const int dx = vx(v1) - vx(v0), dy = vy(v1) - vy(v0);
  const bool inside = ((unsigned)(dx + 3) < 7u) && ((unsigned)(dy + 3) < 7u);
  const int t = dy * 8 + dx;
  // outside the window: every lane reads the SAME slot (facing 0, offset (-3, -3): never inside the
  // window, so it holds 15) -- one broadcast wavefront instead of one per facing
  const unsigned char* wl = &B.win_lut[0][0];
  o0 = observe_with(B, v0, wl[inside ? (vdir(v0) << 6) + 27 + t : 0]);
  o1 = observe_with(B, v1, wl[inside ? (vdir(v1) << 6) + 27 - t : 0]);
}

// ---- initial state --------------------------------------------------------------------
// picks[2*i], picks[2*i+1]: index into the still-available start / destination cells
__device__ __forceinline__ void initial_state(const Blob& B, const int (&picks)[2 * NA], uint32_t (&veh)[NA]) {
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

// the same, picks drawn from one Philox block: counter (particle, 0, STREAM_INIT);
// words x,y -> agent 0 (start, dest), z,w -> agent 1
__device__ __forceinline__ void initial_state_random(const Blob& B, uint64_t seed, uint64_t particle,
                                                     uint32_t (&veh)[NA]) {
  uint32_t used_start = 0, used_dest = 0;
  const uint4 w = philox_block(seed, particle, 0u, STREAM_INIT);
#pragma unroll
  for (int i = 0; i < NA; ++i) {
    const uint32_t ws = (i == 0) ? w.x : w.z;
    const uint32_t wd = (i == 0) ? w.y : w.w;
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

// ---- step noise -------------------------------------------------------------------------
// One 32-bit Philox word per env-step: block = counter (particle >> 2, step, STREAM_STEP),
// word = particle & 3.  idx = floor(word * 50 / 2^32): order = idx % 2, a0 = (idx / 2) % 5,
// a1 = idx / 10 -- uniform on the 50 (order, a0, a1) combinations (bias < 1.2e-8).
__host__ __device__ __forceinline__ uint32_t pick_word(const uint4& blk, unsigned which) {
  return which == 0 ? blk.x : which == 1 ? blk.y : which == 2 ? blk.z : blk.w;
}
__host__ __device__ __forceinline__ void decode_step_noise(uint32_t word, int& a0, int& a1, int& order_idx) {
  const uint32_t idx = uniform_below(word, 50u);
  order_idx = (int)(idx & 1u);
  const uint32_t r = idx >> 1;      // 0..24
  a1 = (int)(r / 5u);
  a0 = (int)(r - 5u * (uint32_t)a1);
}
template <class Seed>   // uint64_t seed or the PhiloxKeys of one
__device__ __forceinline__ void sample_step_noise(const Seed& seed, uint64_t particle, uint32_t step,
                                                  int& a0, int& a1, int& order_idx) {
  const uint4 blk = philox_block(seed, particle >> 2, step, STREAM_STEP);
  decode_step_noise(pick_word(blk, (unsigned)(particle & 3ULL)), a0, a1, order_idx);
}

}  // namespace driving
}  // namespace posg
