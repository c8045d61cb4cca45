// Driving-v1 adapter for the generic search / belief kernels (search.cuh, belief.cuh).
#pragma once
#include "driving.cuh"

namespace posg {

struct DrivingSim {
  using Blob = driving::Blob;
  static constexpr int A = 5;    // ego actions
  static constexpr int OW = 1;   // 64-bit words per observation key
  static constexpr int SW = 2;   // 32-bit words per state
  struct State { uint32_t w[SW]; };

  // model.step with the ego action fixed and the other vehicle uniform random
  // (RandomOtherAgentPolicy, other_policy.py:150-151); `word` = one Philox word
  __device__ static __forceinline__ void step(const Blob& B, State& s, int ego, int ego_action, uint32_t word,
                                              unsigned long long (&obs)[OW], double& r, bool& done) {
    int a0, a1, ord, c0, c1;
    driving::decode_step_noise(word, a0, a1, ord);
    if (ego == 0) a0 = ego_action; else a1 = ego_action;
    const uint32_t before = ego ? s.w[1] : s.w[0];
    driving::transition(B, s.w[0], s.w[1], a0, a1, ord, c0, c1);
    const uint32_t after = ego ? s.w[1] : s.w[0];
    const uint32_t other = ego ? s.w[0] : s.w[1];
    r = driving::reward(B, before, after, ego ? c1 : c0);
    obs[0] = driving::observe(B, after, other);
    done = driving::vdone(after) || (driving::vdone(s.w[0]) && driving::vdone(s.w[1]));
  }

  // one step of MCTS._rollout: every agent uniform random (RandomSearchPolicy)
  __device__ static __forceinline__ void rollout_step(const Blob& B, State& s, int ego, uint32_t word, double& r,
                                                      bool& done) {
    int a0, a1, ord, c0, c1;
    driving::decode_step_noise(word, a0, a1, ord);
    const uint32_t before = ego ? s.w[1] : s.w[0];
    driving::transition(B, s.w[0], s.w[1], a0, a1, ord, c0, c1);
    const uint32_t after = ego ? s.w[1] : s.w[0];
    r = driving::reward(B, before, after, ego ? c1 : c0);
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
