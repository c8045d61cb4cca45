// I-NTMCP nesting level 1 adapter (planning/intmcp.py of the reference): the ego agent's
// tree is searched with the OTHER agent acting by its own level-0 search tree.  A particle
// is (base model state, node of the other agent's tree that its action-observation history
// leads to) -- equivalent to the reference's (state, JointHistory) particle because
// INTMCP.traverse(history) is a pure function of the history (intmcp.py:797-809).
//
//   other agent's action  = INTMCP.sample_action (intmcp.py:763-791): softmax over the
//                           children's visits / sqrt(node visits); the search policy
//                           (uniform) when the node is unvisited / unexpanded / missing;
//   history extension     = traverse one more (action, observation) step in the other tree,
//                           creating the node with zero visits when missing (797-809).
#pragma once
#include "search.cuh"

namespace posg {

template <class Base>
struct NestedSim {
  using Blob = typename Base::Blob;
  struct Ctx {
    search::Tree other_tree;   // level-0 tree of the other agent
    int other;                 // index of the other agent
  };
  static constexpr int A = Base::A, OW = Base::OW, SW = Base::SW + 1, NAG = Base::NAG;
  struct State { uint32_t w[SW]; };

  __device__ static __forceinline__ int sample_other_action(const search::Tree& T, int node, uint32_t word) {
    const int uniform = (int)uniform_below(word, (uint32_t)A);
    if (node < 0) return uniform;
    const int n = *((volatile int*)&T.node_visits[node]);
    if (n == 0 || !(T.node_flags[node] & 1u)) return uniform;
    const float inv = rsqrtf((float)n);
    float wgt[A], sum = 0.0f, mx = 0.0f;
#pragma unroll
    for (int a = 0; a < A; ++a) {
      wgt[a] = (float)(*((volatile int*)&T.an_done[node * A + a])) * inv;
      mx = fmaxf(mx, wgt[a]);
    }
#pragma unroll
    for (int a = 0; a < A; ++a) {
      wgt[a] = __expf(wgt[a] - mx);   // softmax is shift-invariant; avoids overflow at large counts
      sum += wgt[a];
    }
    float u = (float)(word >> 8) * (1.0f / 16777216.0f) * sum;
    int pick = A - 1;
#pragma unroll
    for (int a = A - 1; a >= 0; --a) {
      float below = 0.0f;
#pragma unroll
      for (int b = 0; b < A; ++b) below += (b < a) ? wgt[b] : 0.0f;
      if (u >= below && u < below + wgt[a]) pick = a;
    }
    return pick;
  }

  __device__ static __forceinline__ void step(const Blob& B, const Ctx& ctx, State& s, int ego, int ego_action,
                                              uint32_t word, uint32_t word2, unsigned long long (&obs)[OW],
                                              double& r, bool& done) {
    typename Base::State bs;
#pragma unroll
    for (int k = 0; k < Base::SW; ++k) bs.w[k] = s.w[k];
    int node = (int)s.w[Base::SW];
    const int other = ctx.other;
    const int a_other = sample_other_action(ctx.other_tree, node, word2);
    int act[NAG];
#pragma unroll
    for (int a = 0; a < NAG; ++a) act[a] = (a == ego) ? ego_action : a_other;
    uint32_t noise[Base::NW];
    Base::noise_from_word(word, noise);
    unsigned long long obs_all[NAG][OW];
    double rew[NAG];
    uint32_t flags; bool was_live;
    Base::step_all(B, bs, act, noise, obs_all, rew, flags, was_live);
    unsigned long long o_other[OW];
#pragma unroll
    for (int a = 0; a < NAG; ++a) {
      if (a == ego) {
        r = rew[a];
#pragma unroll
        for (int k = 0; k < OW; ++k) obs[k] = obs_all[a][k];
      }
      if (a == other) {
#pragma unroll
        for (int k = 0; k < OW; ++k) o_other[k] = obs_all[a][k];
      }
    }
    done = ((flags >> ego) & 1u) || (flags & 0x80u);
    // extend the other agent's history: one traverse step in its tree
    if (node >= 0) {
      bool created;
      node = search::find_or_insert<A, OW>(ctx.other_tree, node * A + a_other, o_other,
                                           (int)ctx.other_tree.node_t[node] + 1, created);
    }
#pragma unroll
    for (int k = 0; k < Base::SW; ++k) s.w[k] = bs.w[k];
    s.w[Base::SW] = (uint32_t)node;
  }

  // MCTS-style rollout: every agent by its search policy (uniform), intmcp.py:547-593
  __device__ static __forceinline__ void rollout_step(const Blob& B, const Ctx&, State& s, int ego, uint32_t word,
                                                      double& r, bool& done) {
    typename Base::State bs;
#pragma unroll
    for (int k = 0; k < Base::SW; ++k) bs.w[k] = s.w[k];
    Base::rollout_step(B, NoCtx{}, bs, ego, word, r, done);
#pragma unroll
    for (int k = 0; k < Base::SW; ++k) s.w[k] = bs.w[k];
  }
};

}  // namespace posg
