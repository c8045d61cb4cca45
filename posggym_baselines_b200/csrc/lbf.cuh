// LevelBasedForaging-v3 (2 players, 8 foods, static layout, tuple observations) on packed
// integer state.  Semantics follow posggym's LevelBasedForagingModel.step (a port of
// lb-foraging) as consumed at planning/mcts.py:333,418, planning/belief.py:165 and by the
// PPO vec-env (ppo/worker.py:176-178) of the reference; see oracle/lbf.py for what the
// reference's fixtures pin and what they do not.
//   state word 0: player 0 (x | y<<4 | level<<8) | player 1 << 10 | food_spawned << 20
//   state word 1: food levels, 3 bits each (0 = collected); food cells are static
//   observation : 2 x uint64, every 3-bit field inside one 32-bit half
//     w0 lo: self x+1 @0, y+1 @3, level @6 | other x+1 @8, y+1 @11, level @14 | food slot 0 @16
//     w0 hi: food slots 1..3 (9 bits each: x+1, y+1, level) | w1 lo: slots 4..6 | w1 hi: slot 7
#pragma once
#include <stdint.h>

#include "philox.cuh"

namespace posg {
namespace lbf {

constexpr int NPL = 2, NF = 8;
enum : int { NONE = 0, NORTH = 1, SOUTH = 2, WEST = 3, EAST = 4, LOAD = 5 };

struct alignas(16) Blob {
  uint8_t food_xy[NF];        // x | y << 4 (static layout)
  uint8_t player_xy[4];       // start cells
  uint8_t size, sight, max_agent_level, force_coop, normalize_reward, pad_[11];
  uint8_t food_at[256];       // food index at cell y*16+x, 255 = none
};
static_assert(sizeof(Blob) % 16 == 0, "blob must be a multiple of 16 B");

struct Ents {
  int px[NPL], py[NPL], pl[NPL];
  int fl[NF];
  int spawned;
};

__device__ __forceinline__ void unpack(uint32_t w0, uint32_t w1, Ents& e) {
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const uint32_t p = (w0 >> (10 * i)) & 1023u;
    e.px[i] = p & 15; e.py[i] = (p >> 4) & 15; e.pl[i] = (p >> 8) & 3;
  }
  e.spawned = (w0 >> 20) & 63;
#pragma unroll
  for (int k = 0; k < NF; ++k) e.fl[k] = (w1 >> (3 * k)) & 7;
}
__device__ __forceinline__ void pack(const Ents& e, uint32_t& w0, uint32_t& w1) {
  w0 = (uint32_t)e.spawned << 20; w1 = 0;
#pragma unroll
  for (int i = 0; i < NPL; ++i)
    w0 |= ((uint32_t)e.px[i] | ((uint32_t)e.py[i] << 4) | ((uint32_t)e.pl[i] << 8)) << (10 * i);
#pragma unroll
  for (int k = 0; k < NF; ++k) w1 |= (uint32_t)e.fl[k] << (3 * k);
}

// index of the (uncollected) food at (x, y), -1 if none / outside
__device__ __forceinline__ int food_at(const Blob& B, const Ents& e, int x, int y) {
  if ((unsigned)x >= (unsigned)B.size || (unsigned)y >= (unsigned)B.size) return -1;
  const int k = B.food_at[x | (y << 4)];
  int lvl = 0;
#pragma unroll
  for (int j = 0; j < NF; ++j) lvl = (j == k) ? e.fl[j] : lvl;
  return (k != 255 && lvl > 0) ? k : -1;
}
// first adjacent food in the order N, S, W, E
__device__ __forceinline__ int adjacent_food(const Blob& B, const Ents& e, int x, int y) {
  int k = food_at(B, e, x, y - 1);
  if (k < 0) k = food_at(B, e, x, y + 1);
  if (k < 0) k = food_at(B, e, x - 1, y);
  if (k < 0) k = food_at(B, e, x + 1, y);
  return k;
}

// joint transition; rew[] per player; returns all_done
__device__ __forceinline__ bool transition(const Blob& B, Ents& e, const int (&act)[NPL], double (&rew)[NPL]) {
  int tx[NPL], ty[NPL];
  bool loading[NPL];
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const int a = act[i];
    tx[i] = e.px[i]; ty[i] = e.py[i]; loading[i] = false;
    if (a >= NORTH && a <= EAST) {
      const int nx = e.px[i] + (a == EAST) - (a == WEST), ny = e.py[i] + (a == SOUTH) - (a == NORTH);
      const bool inside = (unsigned)nx < (unsigned)B.size && (unsigned)ny < (unsigned)B.size;
      if (inside && food_at(B, e, nx, ny) < 0) { tx[i] = nx; ty[i] = ny; }
    } else if (a == LOAD) {
      loading[i] = adjacent_food(B, e, e.px[i], e.py[i]) >= 0;
    }
  }
  // players whose target cell is claimed by more than one player stay
  const bool clash = (tx[0] == tx[1]) && (ty[0] == ty[1]);
  if (!clash) {
#pragma unroll
    for (int i = 0; i < NPL; ++i) { e.px[i] = tx[i]; e.py[i] = ty[i]; }
  }
  rew[0] = 0.0; rew[1] = 0.0;
  // loading, in player index order; a group = the pending loaders adjacent to the same food
  bool pending[NPL] = {loading[0], loading[1]};
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    if (!pending[i]) continue;
    pending[i] = false;
    const int k = adjacent_food(B, e, e.px[i], e.py[i]);
    if (k < 0) continue;
    const int fx = B.food_xy[k] & 15, fy = B.food_xy[k] >> 4;
    bool in_group[NPL] = {false, false};
    in_group[i] = true;
    int level_sum = e.pl[i];
#pragma unroll
    for (int j = 0; j < NPL; ++j) {
      if (j == i || !pending[j]) continue;
      if (abs(e.px[j] - fx) + abs(e.py[j] - fy) == 1) { in_group[j] = true; pending[j] = false; level_sum += e.pl[j]; }
    }
    int flevel = 0;
#pragma unroll
    for (int f = 0; f < NF; ++f) flevel = (f == k) ? e.fl[f] : flevel;
    if (level_sum < flevel) continue;
#pragma unroll
    for (int j = 0; j < NPL; ++j) {
      if (!in_group[j]) continue;
      double r = (double)(e.pl[j] * flevel);
      if (B.normalize_reward) r = __ddiv_rn(r, (double)(level_sum * e.spawned));
      rew[j] = __dadd_rn(rew[j], r);
    }
#pragma unroll
    for (int f = 0; f < NF; ++f) e.fl[f] = (f == k) ? 0 : e.fl[f];
  }
  bool all_done = true;
#pragma unroll
  for (int f = 0; f < NF; ++f) all_done &= (e.fl[f] == 0);
  return all_done;
}

// tuple observation of player i as two packed words
__device__ __forceinline__ void observe(const Blob& B, const Ents& e, int i, unsigned long long (&obs)[2]) {
  const int s = B.sight, o = 1 - i;
  uint32_t h[4] = {0, 0, 0, 0};   // w0 lo, w0 hi, w1 lo, w1 hi
  h[0] = (uint32_t)(s + 1) | ((uint32_t)(s + 1) << 3) | ((uint32_t)e.pl[i] << 6);
  {
    const int dx = e.px[o] - e.px[i], dy = e.py[o] - e.py[i];
    if (abs(dx) <= s && abs(dy) <= s)
      h[0] |= ((uint32_t)(dx + s + 1) << 8) | ((uint32_t)(dy + s + 1) << 11) | ((uint32_t)e.pl[o] << 14);
  }
  int slot = 0;
#pragma unroll
  for (int k = 0; k < NF; ++k) {
    const int dx = (B.food_xy[k] & 15) - e.px[i], dy = (B.food_xy[k] >> 4) - e.py[i];
    if (e.fl[k] > 0 && abs(dx) <= s && abs(dy) <= s) {
      const uint32_t f = (uint32_t)(dx + s + 1) | ((uint32_t)(dy + s + 1) << 3) | ((uint32_t)e.fl[k] << 6);
      // slot 0 -> h[0] @16; slots 1..3 -> h[1] @0,9,18; 4..6 -> h[2]; 7 -> h[3] @0
      const int half = slot == 0 ? 0 : (slot + 2) / 3, off = slot == 0 ? 16 : 9 * ((slot - 1) % 3);
#pragma unroll
      for (int q = 0; q < 4; ++q) h[q] |= (q == half) ? (f << off) : 0u;
      ++slot;
    }
  }
  obs[0] = (unsigned long long)h[0] | ((unsigned long long)h[1] << 32);
  obs[1] = (unsigned long long)h[2] | ((unsigned long long)h[3] << 32);
}

__device__ __forceinline__ void initial_from_picks(const Blob& B, const int (&picks)[NPL + NF], Ents& e) {
  int max_level = 0;
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    e.px[i] = B.player_xy[i] & 15; e.py[i] = B.player_xy[i] >> 4; e.pl[i] = picks[i] + 1;
    max_level += e.pl[i];
  }
  e.spawned = 0;
#pragma unroll
  for (int k = 0; k < NF; ++k) {
    e.fl[k] = B.force_coop ? max_level : picks[NPL + k] + 1;
    e.spawned += e.fl[k];
  }
}

}  // namespace lbf

// Adapter for the generic batch / search / belief / vec-env kernels.
struct LBFSim {
  using Blob = lbf::Blob;
  using Ctx = NoCtx;
  static constexpr int A = 6, OW = 2, SW = 2, NAG = 2, NW = 1, NPICK = 10;
  struct State { uint32_t w[SW]; };

  // flattened-observation fields: (x, y, level) of self, the other player, then the eight
  // food slots; coordinate fields hold value + 1 = the one-hot index of Discrete(.., start=-1)
  // (see DrivingSim::obs_field for the encoding of the return value)
  static constexpr int NFIELD = 30;
  __host__ __device__ static constexpr uint32_t obs_field(int f) {
    if (f < 6) {
      const int sh = f == 0 ? 0 : f == 1 ? 3 : f == 2 ? 6 : f == 3 ? 8 : f == 4 ? 11 : 14;
      return 0u | ((uint32_t)sh << 8) | ((f == 2 || f == 5 ? 2u : 3u) << 16);
    }
    const int slot = (f - 6) / 3, c = (f - 6) % 3;
    const int half = slot == 0 ? 0 : (slot + 2) / 3, off = slot == 0 ? 16 : 9 * ((slot - 1) % 3);
    return (uint32_t)half | ((uint32_t)(off + 3 * c) << 8) | (3u << 16);
  }

  __device__ static __forceinline__ void step_all(const Blob& B, State& s, const int (&act)[NAG],
                                                  const uint32_t (&)[NW], unsigned long long (&obs)[NAG][OW],
                                                  double (&rew)[NAG], uint32_t& flags, bool& was_live) {
    lbf::Ents e;
    lbf::unpack(s.w[0], s.w[1], e);
    was_live = s.w[1] != 0u;
    const bool all_done = lbf::transition(B, e, act, rew);
    lbf::pack(e, s.w[0], s.w[1]);
    lbf::observe(B, e, 0, obs[0]);
    lbf::observe(B, e, 1, obs[1]);
    flags = all_done ? (0x80u | 3u) : 0u;
  }
  __device__ static __forceinline__ void step_noobs(const Blob& B, State& s, const int (&act)[NAG],
                                                    const uint32_t (&)[NW], double (&rew)[NAG], uint32_t& term,
                                                    bool& all_done) {
    lbf::Ents e;
    lbf::unpack(s.w[0], s.w[1], e);
    all_done = lbf::transition(B, e, act, rew);
    lbf::pack(e, s.w[0], s.w[1]);
    term = all_done ? 3u : 0u;
  }
  // uniform joint action from one Philox word: idx = floor(w * 36 / 2^32), a0 = idx % 6, a1 = idx / 6
  __device__ static __forceinline__ void actions_from_word(uint32_t word, int (&act)[NAG]) {
    const uint32_t idx = uniform_below(word, 36u);
    act[1] = (int)(idx / 6u);
    act[0] = (int)(idx - 6u * (uint32_t)act[1]);
  }
  __device__ static __forceinline__ void noise_from_philox(uint64_t seed, uint64_t particle, uint32_t step,
                                                           int (&act)[NAG], uint32_t (&noise)[NW]) {
    const uint4 w = philox_block(seed, particle, step, STREAM_STEP);
    actions_from_word(w.x, act);
    noise[0] = 0u;
  }
  __device__ static __forceinline__ void noise_from_word(uint32_t, uint32_t (&noise)[NW]) { noise[0] = 0u; }
  __device__ static __forceinline__ void observe_all(const Blob& B, const State& s, unsigned long long (&obs)[NAG][OW]) {
    lbf::Ents e;
    lbf::unpack(s.w[0], s.w[1], e);
    lbf::observe(B, e, 0, obs[0]);
    lbf::observe(B, e, 1, obs[1]);
  }
  __device__ static __forceinline__ void initial_from_picks(const Blob& B, const int (&picks)[NPICK], State& s) {
    lbf::Ents e;
    lbf::initial_from_picks(B, picks, e);
    lbf::pack(e, s.w[0], s.w[1]);
  }
  // picks from Philox: block (particle, 0, 0xB0): x, y -> player levels, z, w -> foods 0, 1;
  // block (.., 0xB1): foods 2..5; block (.., 0xB2): x, y -> foods 6, 7
  __device__ static __forceinline__ void initial_state(const Blob& B, uint64_t seed, uint64_t particle, State& s) {
    const uint4 a = philox_block(seed, particle, 0u, STREAM_INIT);
    const uint4 b = philox_block(seed, particle, 0u, STREAM_INIT2);
    const uint4 c = philox_block(seed, particle, 0u, 0xB2u);
    int picks[NPICK];
    picks[0] = (int)uniform_below(a.x, (uint32_t)B.max_agent_level);
    picks[1] = (int)uniform_below(a.y, (uint32_t)B.max_agent_level);
    const uint32_t bound = B.force_coop ? 1u : (uint32_t)(picks[0] + picks[1] + 2);
    const uint32_t fw[8] = {a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y};
#pragma unroll
    for (int k = 0; k < 8; ++k) picks[2 + k] = (int)uniform_below(fw[k], bound);
    initial_from_picks(B, picks, s);
  }

  // ---- search / belief adapter ------------------------------------------------------------
  __device__ static __forceinline__ void step(const Blob& B, const Ctx&, State& s, int ego, int ego_action,
                                              uint32_t word, uint32_t /*word2*/, unsigned long long (&obs)[OW],
                                              double& r, bool& done) {
    int act[NAG];
    const int other = (int)uniform_below(word, 6u);
    act[0] = ego == 0 ? ego_action : other;
    act[1] = ego == 1 ? ego_action : other;
    lbf::Ents e;
    lbf::unpack(s.w[0], s.w[1], e);
    double rew[NAG];
    done = lbf::transition(B, e, act, rew);
    lbf::pack(e, s.w[0], s.w[1]);
    r = rew[ego];
    lbf::observe(B, e, ego, obs);
  }
  __device__ static __forceinline__ void rollout_step(const Blob& B, const Ctx&, State& s, int ego, uint32_t word,
                                                      double& r, bool& done) {
    int act[NAG];
    actions_from_word(word, act);
    lbf::Ents e;
    lbf::unpack(s.w[0], s.w[1], e);
    double rew[NAG];
    done = lbf::transition(B, e, act, rew);
    lbf::pack(e, s.w[0], s.w[1]);
    r = rew[ego];
  }
  __device__ static __forceinline__ void observe(const Blob& B, const State& s, int ego, unsigned long long (&obs)[OW]) {
    lbf::Ents e;
    lbf::unpack(s.w[0], s.w[1], e);
    lbf::observe(B, e, ego, obs);
  }
};

}  // namespace posg
