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
  uint8_t size, sight, max_agent_level, force_coop, normalize_reward;
  uint8_t vis_max;            // most foods any cell can see (build_tables); <= 7: the observation stream fits 64 bits
  uint8_t adj_fast;           // 1: N, S, W, E order of a cell's adjacent foods = ascending food index (adj3 + ffs)
  uint8_t pad_[1];
  int8_t move_delta[8];       // cell-index delta (x | y << 4 arithmetic) of NONE, N, S, W, E, LOAD
  // 3 * k of the food whose cell is x | y << 4, 24 if there is none: level = (w1 >> shift) & 7
  // (w1 holds 24 level bits, so shift 24 reads 0)
  uint8_t food_sh[256];
  // food_sh of the N, S, W, E neighbours of cell x | y << 4 (one byte each; 24 = none / outside)
  uint32_t adj_sh[256];
  // what a player standing on cell x | y << 4 can see: bits 0..7 food k within sight;
  // bits 8 + 6 k .. 13 + 6 k: (dx + sight + 1) | (dy + sight + 1) << 3 of food k
  unsigned long long vis[256];
  uint32_t vis3[256];         // bit 3 k: food k within sight of cell x | y << 4 (the level bits' positions in w1)
  uint32_t adj3[256];         // bit 3 k: food k is the N, S, W or E neighbour of cell x | y << 4; bit 24 always set
  // normalised rewards (double)(level * food_level) / (double)(loaders' level sum * total level spawned), every
  // quotient the state can ask for, divided once on the host (IEEE, the oracle's float division): DEVICE pointer
  // to REWARD_DEN x REWARD_NUM doubles in global memory (a load rarely executed instead of a ~50-instruction
  // division the whole warp waits for); NULL on the host copy
  const double* reward_div;
  uint64_t pad2_;
};
constexpr int REWARD_NUM = 3 * 7 + 1, REWARD_DEN = 6 * 63 + 1;
static_assert(sizeof(Blob) % 16 == 0, "blob must be a multiple of 16 B");

// host side: fill the tables from food_xy / size / sight
inline void build_tables(Blob* b) {
  const int8_t md[8] = {0, -16, 16, -1, 1, 0, 0, 0};
  for (int k = 0; k < 8; ++k) b->move_delta[k] = md[k];
  for (int c = 0; c < 256; ++c) b->food_sh[c] = 24;
  for (int k = 0; k < NF; ++k) b->food_sh[b->food_xy[k]] = (uint8_t)(3 * k);
  const int size = b->size, s = b->sight;
  int vis_max = 0;
  for (int y = 0; y < 16; ++y)
    for (int x = 0; x < 16; ++x) {
      auto sh = [&](int cx, int cy) -> uint32_t {
        return (cx < 0 || cy < 0 || cx >= size || cy >= size) ? 24u : (uint32_t)b->food_sh[cx | (cy << 4)];
      };
      b->adj_sh[x | (y << 4)] = sh(x, y - 1) | (sh(x, y + 1) << 8) | (sh(x - 1, y) << 16) | (sh(x + 1, y) << 24);
      unsigned long long v = 0;
      for (int k = 0; k < NF; ++k) {
        const int dx = (b->food_xy[k] & 15) - x, dy = (b->food_xy[k] >> 4) - y;
        if (dx >= -s && dx <= s && dy >= -s && dy <= s) {
          v |= 1ULL << k;
          v |= (unsigned long long)((dx + s + 1) | ((dy + s + 1) << 3)) << (8 + 6 * k);
        }
      }
      b->vis[x | (y << 4)] = v;
      uint32_t v3 = 0;
      for (int k = 0; k < NF; ++k) v3 |= (uint32_t)((v >> k) & 1ULL) << (3 * k);
      b->vis3[x | (y << 4)] = v3;
      int seen = 0;
      for (int k = 0; k < NF; ++k) seen += (int)((v >> k) & 1ULL);
      if (x < size && y < size && seen > vis_max) vis_max = seen;
    }
  b->vis_max = (uint8_t)vis_max;
  // adjacent foods as a mask over the level bits of w1; the lowest set bit is the N, S, W, E choice when the foods
  // next to every cell come in ascending index order (true for the lattice layouts: two foods are never
  // adjacent to one cell across a corner)
  bool ordered = true;
  for (int c = 0; c < 256; ++c) {
    const uint32_t a = b->adj_sh[c];
    const uint32_t in_order[4] = {a & 255u, (a >> 8) & 255u, (a >> 16) & 255u, a >> 24};   // N, S, W, E
    uint32_t m = 1u << 24, last = 0;
    bool first = true;
    for (int q = 0; q < 4; ++q) {
      if (in_order[q] == 24u) continue;
      m |= 1u << in_order[q];
      if (!first && in_order[q] < last) ordered = false;
      last = in_order[q]; first = false;
    }
    b->adj3[c] = m;
  }
  b->adj_fast = ordered ? 1 : 0;
}

// shift (3 * food index) of the first uncollected food next to `cell` in the order N, S, W, E; 24 if none
__device__ __forceinline__ uint32_t adjacent_food(const Blob& B, uint32_t w1, uint32_t cell) {
  if (B.adj_fast) {   // (uniform) bit 3 k of nz: food k still has a level; bit 24 of adj3 yields "none"
    const uint32_t nz = w1 | (w1 >> 1) | (w1 >> 2) | (1u << 24);
    return (uint32_t)__ffs((int)(B.adj3[cell] & nz & 0x1249249u)) - 1u;
  }
  const uint32_t a = B.adj_sh[cell];
  const uint32_t sn = a & 255u, ss = (a >> 8) & 255u, sw = (a >> 16) & 255u, se = a >> 24;
  uint32_t sh = ((w1 >> se) & 7u) ? se : 24u;
  sh = ((w1 >> sw) & 7u) ? sw : sh;
  sh = ((w1 >> ss) & 7u) ? ss : sh;
  sh = ((w1 >> sn) & 7u) ? sn : sh;
  return sh;
}

// Joint transition on the packed words (see the header of this file); rew[] per player; returns
// all_done.  Moves are branch-free table lookups; only the (rare) loading resolution branches.
__device__ __forceinline__ bool transition(const Blob& B, uint32_t& w0, uint32_t& w1, const int (&act)[NPL],
                                           double (&rew)[NPL]) {
  const uint32_t size = B.size;
  uint32_t cell[NPL], lvl[NPL], tgt[NPL], load_sh[NPL];
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const uint32_t p = (w0 >> (10 * i)) & 1023u;
    cell[i] = p & 255u; lvl[i] = p >> 8;
    const int a = act[i];
    const uint32_t nc = cell[i] + (uint32_t)(int)B.move_delta[a & 7];
    // inside the field: row in range (a step off the top wraps to a huge value) and column in range
    // (a step off the left edge borrows: column nibble 15)
    const bool inside = (nc < (size << 4)) && ((nc & 15u) < size);
    const bool food = ((w1 >> B.food_sh[nc & 255u]) & 7u) != 0u;
    tgt[i] = (inside && !food) ? nc : cell[i];     // NONE / LOAD: nc == cell
    load_sh[i] = (a == LOAD) ? adjacent_food(B, w1, cell[i]) : 24u;
  }
  // players whose target cell is claimed by more than one player stay
  const bool clash = tgt[0] == tgt[1];
  cell[0] = clash ? cell[0] : tgt[0];
  cell[1] = clash ? cell[1] : tgt[1];
  w0 = (w0 & 0xFFF00000u) | cell[0] | (lvl[0] << 8) | ((cell[1] | (lvl[1] << 8)) << 10);
  rew[0] = 0.0; rew[1] = 0.0;
  if (load_sh[0] != 24u || load_sh[1] != 24u) {   // somebody loads
    const uint32_t spawned = (w0 >> 20) & 63u;
    bool pending[NPL] = {load_sh[0] != 24u, load_sh[1] != 24u};
    // loading, in player index order; a group = the pending loaders adjacent to the same food
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      if (!pending[i]) continue;
      pending[i] = false;
      // a loader did not move; the food it faced may have been taken by the previous group
      const uint32_t sh = (i == 0) ? load_sh[0] : adjacent_food(B, w1, cell[i]);
      if (sh == 24u) continue;
      const uint32_t fxy = B.food_xy[(sh * 11u) >> 5];   // sh / 3 for sh in {0, 3, .., 21}
      const int fx = fxy & 15, fy = fxy >> 4;
      bool in_group[NPL] = {false, false};
      in_group[i] = true;
      uint32_t level_sum = lvl[i];
#pragma unroll
      for (int j = 0; j < NPL; ++j) {
        if (j == i || !pending[j]) continue;
        if (abs((int)(cell[j] & 15u) - fx) + abs((int)(cell[j] >> 4) - fy) == 1) {
          in_group[j] = true; pending[j] = false; level_sum += lvl[j];
        }
      }
      const uint32_t flevel = (w1 >> sh) & 7u;
      if (level_sum < flevel) continue;
#pragma unroll
      for (int j = 0; j < NPL; ++j) {
        if (!in_group[j]) continue;
        double r = (double)(lvl[j] * flevel);
        if (B.normalize_reward) r = __ldg(B.reward_div + (level_sum * spawned) * (uint32_t)REWARD_NUM + lvl[j] * flevel);
        rew[j] = __dadd_rn(rew[j], r);
      }
      w1 &= ~(7u << sh);
    }
  }
  return w1 == 0u;
}

// tuple observation of player i as two packed words
__device__ __forceinline__ void observe(const Blob& B, uint32_t w0, uint32_t w1, int i, unsigned long long (&obs)[2]) {
  const uint32_t s = B.sight;
  const uint32_t me = (w0 >> (10 * i)) & 1023u, ot = (w0 >> (10 * (1 - i))) & 1023u;
  uint32_t h0 = (s + 1u) | ((s + 1u) << 3) | ((me >> 8) << 6);
  {
    const uint32_t dx = (ot & 15u) - (me & 15u) + s, dy = ((ot >> 4) & 15u) - ((me >> 4) & 15u) + s;
    if (dx <= 2u * s && dy <= 2u * s) h0 |= ((dx + 1u) << 8) | ((dy + 1u) << 11) | ((ot >> 8) << 14);
  }
  // visible foods, packed into consecutive 9-bit slots of a 72-bit stream (slot 7 straddles)
  const unsigned long long V = B.vis[me & 255u];
  unsigned long long lo = 0;
  uint32_t hi = 0;
  if (B.vis_max <= 7) {
    // (uniform branch) at most seven visible foods = 63 bits: build the stream from the LAST food down, every
    // visible food shifting the ones after it up by one slot -- constant shifts, no running slot index
    // seen = visible from here and not loaded yet (level != 0): bit 3 k of both masks
    const uint32_t seen3 = B.vis3[me & 255u] & (w1 | (w1 >> 1) | (w1 >> 2));
#pragma unroll
    for (int k = NF - 1; k >= 0; --k) {
      // level << 6 of food k, straight from the state word (3 k - 6 may be negative: shift the other way)
      const uint32_t lv6 = (3 * k >= 6 ? (w1 >> (3 * k - 6)) : (w1 << (6 - 3 * k))) & 0x1C0u;
      const uint32_t code = ((uint32_t)(V >> (8 + 6 * k)) & 63u) | lv6;
      if (seen3 & (1u << (3 * k))) lo = (lo << 9) | code;
    }
  } else {
    uint32_t slot9 = 0;   // slot9 = 9 * slot
#pragma unroll
    for (int k = 0; k < NF; ++k) {
      const uint32_t lv = (w1 >> (3 * k)) & 7u;
      const bool seen = ((V >> k) & 1ULL) && lv;
      const unsigned long long code = (unsigned long long)(((uint32_t)(V >> (8 + 6 * k)) & 63u) | (lv << 6));
      const unsigned long long put = seen ? code : 0ULL;
      lo |= put << slot9;
      if (k == NF - 1) hi = (slot9 == 63u) ? (uint32_t)(put >> 1) : 0u;
      slot9 += seen ? 9u : 0u;
    }
  }
  // stream -> layout: slot 0 at w0 lo bit 16; slots 1..3 in w0 hi; 4..6 in w1 lo; 7 in w1 hi
  h0 |= ((uint32_t)lo & 511u) << 16;
  const uint32_t h1 = (uint32_t)(lo >> 9) & 0x7FFFFFFu, h2 = (uint32_t)(lo >> 36) & 0x7FFFFFFu;
  const uint32_t h3 = ((uint32_t)(lo >> 63) | (hi << 1)) & 511u;
  obs[0] = (unsigned long long)h0 | ((unsigned long long)h1 << 32);
  obs[1] = (unsigned long long)h2 | ((unsigned long long)h3 << 32);
}

__device__ __forceinline__ void initial_from_picks(const Blob& B, const int (&picks)[NPL + NF], uint32_t& w0,
                                                   uint32_t& w1) {
  uint32_t max_level = 0;
  w0 = 0; w1 = 0;
#pragma unroll
  for (int i = 0; i < NPL; ++i) {
    const uint32_t l = (uint32_t)picks[i] + 1u;
    w0 |= ((uint32_t)B.player_xy[i] | (l << 8)) << (10 * i);
    max_level += l;
  }
  uint32_t spawned = 0;
#pragma unroll
  for (int k = 0; k < NF; ++k) {
    const uint32_t l = B.force_coop ? max_level : (uint32_t)picks[NPL + k] + 1u;
    w1 |= l << (3 * k);
    spawned += l;
  }
  w0 |= spawned << 20;
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
    was_live = s.w[1] != 0u;
    const bool all_done = lbf::transition(B, s.w[0], s.w[1], act, rew);
    lbf::observe(B, s.w[0], s.w[1], 0, obs[0]);
    lbf::observe(B, s.w[0], s.w[1], 1, obs[1]);
    flags = all_done ? (0x80u | 3u) : 0u;
  }
  __device__ static __forceinline__ void step_noobs(const Blob& B, State& s, const int (&act)[NAG],
                                                    const uint32_t (&)[NW], double (&rew)[NAG], uint32_t& term,
                                                    bool& all_done) {
    all_done = lbf::transition(B, s.w[0], s.w[1], act, rew);
    term = all_done ? 3u : 0u;
  }
  static constexpr bool ONE_WORD_NOISE = true;   // the whole step noise is one Philox word (step_fast_kernel)
  static constexpr bool NOISE_PAIR = false;
  // uniform joint action from one Philox word: idx = floor(w * 36 / 2^32), a0 = idx % 6, a1 = idx / 6
  __device__ static __forceinline__ void actions_from_word(uint32_t word, int (&act)[NAG]) {
    const uint32_t idx = uniform_below(word, 36u);
    act[1] = (int)(idx / 6u);
    act[0] = (int)(idx - 6u * (uint32_t)act[1]);
  }
  template <class Seed>   // uint64_t seed or the PhiloxKeys of one
  __device__ static __forceinline__ void noise_from_philox(const Seed& seed, uint64_t particle, uint32_t step,
                                                           int (&act)[NAG], uint32_t (&noise)[NW]) {
    // one 32-bit word per env-step: block (particle >> 2, step), word particle & 3 (four consecutive
    // envs share a block; generic::step_fast_kernel computes it once for the four)
    const uint4 w = philox_block(seed, particle >> 2, step, STREAM_STEP);
    const unsigned which = (unsigned)(particle & 3ULL);
    actions_from_word(which == 0 ? w.x : which == 1 ? w.y : which == 2 ? w.z : w.w, act);
    noise[0] = 0u;
  }
  __device__ static __forceinline__ void noise_from_word(uint32_t, uint32_t (&noise)[NW]) { noise[0] = 0u; }
  __device__ static __forceinline__ void observe_all(const Blob& B, const State& s, unsigned long long (&obs)[NAG][OW]) {
    lbf::observe(B, s.w[0], s.w[1], 0, obs[0]);
    lbf::observe(B, s.w[0], s.w[1], 1, obs[1]);
  }
  __device__ static __forceinline__ void initial_from_picks(const Blob& B, const int (&picks)[NPICK], State& s) {
    lbf::initial_from_picks(B, picks, s.w[0], s.w[1]);
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
    double rew[NAG];
    done = lbf::transition(B, s.w[0], s.w[1], act, rew);
    r = rew[ego];
    lbf::observe(B, s.w[0], s.w[1], ego, obs);
  }
  __device__ static __forceinline__ void rollout_step(const Blob& B, const Ctx&, State& s, int ego, uint32_t word,
                                                      double& r, bool& done) {
    int act[NAG];
    actions_from_word(word, act);
    double rew[NAG];
    done = lbf::transition(B, s.w[0], s.w[1], act, rew);
    r = rew[ego];
  }
  __device__ static __forceinline__ void observe(const Blob& B, const State& s, int ego, unsigned long long (&obs)[OW]) {
    lbf::observe(B, s.w[0], s.w[1], ego, obs);
  }
};

}  // namespace posg
