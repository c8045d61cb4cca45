// Driving-v1 adapter for the generic search / belief kernels (search.cuh, belief.cuh).
#pragma once
#include "driving.cuh"

namespace posg {

struct DrivingSim {
  using Blob = driving::Blob;
  using Ctx = NoCtx;
  static constexpr int A = 5;    // ego actions
  static constexpr int OW = 1;   // 64-bit words per observation key
  static constexpr int SW = 2;   // 32-bit words per state
  static constexpr int NAG = 2;  // agents
  static constexpr int NW = 1;   // explicit noise words per step (processing order)
  struct State { uint32_t w[SW]; };

  // flattened-observation fields (gymnasium.spaces.flatten order of the observation tuple):
  // 15 local cells, speed, x, y, dest x, dest y, dest_reached, crashed.  Returns
  // 32-bit half of the packed observation | shift inside the half << 8 | bits << 16.
  static constexpr int NFIELD = 22;
  __host__ __device__ static constexpr uint32_t obs_field(int f) {
    return f < 16 ? (0u | ((uint32_t)(2 * f) << 8) | (2u << 16))
         : f < 20 ? (1u | ((uint32_t)(4 * (f - 16)) << 8) | (4u << 16))
                  : (1u | ((uint32_t)(16 + (f - 20)) << 8) | (1u << 16));
  }

  // ---- batch concept (generic_kernels.cuh / vec_env.cuh) ---------------------------
  __device__ static __forceinline__ void step_all(const Blob& B, State& s, const int (&act)[NAG],
                                                  const uint32_t (&noise)[NW], unsigned long long (&obs)[NAG][OW],
                                                  double (&rew)[NAG], uint32_t& flags, bool& was_live) {
    const uint32_t b0 = s.w[0], b1 = s.w[1];
    was_live = !(driving::vdone(b0) && driving::vdone(b1));
    driving::transition_raw(B, s.w[0], s.w[1], act[0], act[1], (int)(noise[0] & 1u));
    rew[0] = driving::reward_raw(B, b0, s.w[0]);
    rew[1] = driving::reward_raw(B, b1, s.w[1]);
    s.w[0] &= driving::STATE_MASK; s.w[1] &= driving::STATE_MASK;
    uint64_t o0, o1;
    driving::observe_pair(B, s.w[0], s.w[1], o0, o1);
    obs[0][0] = o0; obs[1][0] = o1;
    const uint32_t t0 = driving::vdone(s.w[0]) ? 1u : 0u, t1 = driving::vdone(s.w[1]) ? 1u : 0u;
    flags = t0 | (t1 << 1) | ((t0 & t1) ? 0x80u : 0u);
  }
  template <class Seed>   // uint64_t seed or the PhiloxKeys of one
  __device__ static __forceinline__ void noise_from_philox(const Seed& seed, uint64_t particle, uint32_t step,
                                                           int (&act)[NAG], uint32_t (&noise)[NW]) {
    int order;
    driving::sample_step_noise(seed, particle, step, act[0], act[1], order);
    noise[0] = (uint32_t)order;
  }
  // step noise from one 32-bit word (nested search): the processing order
  __device__ static __forceinline__ void noise_from_word(uint32_t word, uint32_t (&noise)[NW]) {
    noise[0] = word >> 31;
  }
  __device__ static __forceinline__ void observe_all(const Blob& B, const State& s, unsigned long long (&obs)[NAG][OW]) {
    uint64_t o0, o1;
    driving::observe_pair(B, s.w[0], s.w[1], o0, o1);
    obs[0][0] = o0; obs[1][0] = o1;
  }

  // model.step with the ego action fixed and the other vehicle uniform random
  // (RandomOtherAgentPolicy, other_policy.py:150-151); `word` = one Philox word
  __device__ static __forceinline__ void step(const Blob& B, const Ctx&, State& s, int ego, int ego_action,
                                              uint32_t word, uint32_t /*word2*/, unsigned long long (&obs)[OW],
                                              double& r, bool& done) {
    int a0, a1, ord;
    driving::decode_step_noise(word, a0, a1, ord);
    if (ego == 0) a0 = ego_action; else a1 = ego_action;
    const uint32_t before = ego ? s.w[1] : s.w[0];
    driving::transition_raw(B, s.w[0], s.w[1], a0, a1, ord);
    r = driving::reward_raw(B, before, ego ? s.w[1] : s.w[0]);
    s.w[0] &= driving::STATE_MASK; s.w[1] &= driving::STATE_MASK;
    const uint32_t after = ego ? s.w[1] : s.w[0];
    const uint32_t other = ego ? s.w[0] : s.w[1];
    obs[0] = driving::observe(B, after, other);
    done = driving::vdone(after) || (driving::vdone(s.w[0]) && driving::vdone(s.w[1]));
  }

  // one step of MCTS._rollout: every agent uniform random (RandomSearchPolicy)
  __device__ static __forceinline__ void rollout_step(const Blob& B, const Ctx&, State& s, int ego, uint32_t word,
                                                      double& r, bool& done) {
    int a0, a1, ord;
    driving::decode_step_noise(word, a0, a1, ord);
    const uint32_t before = ego ? s.w[1] : s.w[0];
    driving::transition_raw(B, s.w[0], s.w[1], a0, a1, ord);
    r = driving::reward_raw(B, before, ego ? s.w[1] : s.w[0]);
    s.w[0] &= driving::STATE_MASK; s.w[1] &= driving::STATE_MASK;
    const uint32_t after = ego ? s.w[1] : s.w[0];
    done = driving::vdone(after) || (driving::vdone(s.w[0]) && driving::vdone(s.w[1]));
  }

  __device__ static __forceinline__ void initial_state(const Blob& B, uint64_t seed, uint64_t particle, State& s) {
    uint32_t veh[2];
    driving::initial_state_random(B, seed, particle, veh);
    s.w[0] = veh[0]; s.w[1] = veh[1];
  }

  __device__ static __forceinline__ void observe(const Blob& B, const State& s, int ego, unsigned long long (&obs)[OW]) {
    obs[0] = ego ? driving::observe(B, s.w[1], s.w[0]) : driving::observe(B, s.w[0], s.w[1]);
  }
};

}  // namespace posg
