// POMCP search tree on the device (SoA arena owned by the caller) and the wave-parallel
// simulation kernel.  Restates, for a batch of simulations per launch, the reference's
//   MCTS.get_action / _simulate / _evaluate / _rollout / ucb_action_selection /
//   pucb_action_selection / min_visit_action_selection
// (planning/mcts.py:269-306, 308-382, 384-403, 405-452, 494-561) with
// RandomSearchPolicy / RandomOtherAgentPolicy plugins and state_belief_only=True.
//
// One thread = one simulation (root-to-leaf descent, leaf rollout, backup).  Lanes of a
// warp that sit at the same tree node choose their actions by "water-filling": lane with
// rank r takes the action sequential UCB would take on the r-th of those visits with the
// value estimates frozen, so a wave of 1 is exactly the reference's sequential search and
// larger waves spread over the actions the way consecutive reference simulations would.
#pragma once
#include <stdint.h>

#include "philox.cuh"

namespace posg {
namespace search {

constexpr int MAX_DEPTH = 64;      // path stack (the in-scope envs stop at 50 steps)
constexpr int SEL_UCB = 0, SEL_PUCB = 1, SEL_UNIFORM = 2;
constexpr unsigned long long RESERVE = 1024;  // node slots search never takes

struct Counters {               // first 64 bytes of the arena
  unsigned long long node_count;
  unsigned long long log_count;
  unsigned long long min_key;   // order-preserving keys of the min / max action value
  unsigned long long max_key;
  unsigned int max_depth;
  unsigned int overflow;        // simulations that could not expand (arena full)
  unsigned long long sims_done;
  unsigned long long nodes_created;   // obs nodes the searches linked into the tree (node_count also counts reserved ids)
  // order-preserving keys of the smallest / largest single return backed up so far: a node's mean lies between
  // them, which bounds what a torn (total, visits) pair of a concurrent backup can feed to min_key / max_key
  unsigned long long gmin_key, gmax_key;
  unsigned long long pad_[7];
};
static_assert(sizeof(Counters) == 128, "counter block is 128 bytes");

struct Params {
  double discount;
  double c;
  double known_min, known_max;  // known_min > known_max: no known bounds
  double pucb_exploration_fraction;
  int depth_limit;
  int step_limit;               // max_episode_steps (node.t may not exceed it)
  int selection;                // SEL_*
  int ego;
  int intmcp_depth_rule;        // 1: stop when node.t + depth > step_limit (intmcp.py:450-453)
  uint64_t seed;
  double disc_pow[MAX_DEPTH];   // discount**k from libm pow (mcts.py:421)
};

// Arena view.  OW = 64-bit words per observation key, SW = 32-bit words per state.
struct Tree {
  Counters* ctr;
  int* node_visits;
  int* node_next;
  int* node_parent;             // parent action node (node * |A| + a), -1 for roots
  unsigned char* node_flags;    // bit 0 expanded, bit 1 absorbing
  unsigned char* node_t;
  unsigned long long* node_obs;
  int* an_first;
  int* an_visits;               // selections, in-flight included (virtual visits)
  int* an_done;                 // completed backups
  double* an_total;
  unsigned int* log_node;       // particle log: node id ...
  unsigned int* log_state;      // ... and state words
  long long cap_nodes, cap_log;
};

__host__ __device__ inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// Lays the arrays out inside `base` (may be null to only compute the size).
__host__ __device__ inline size_t layout(Tree* t, unsigned char* base, long long cap_nodes, long long cap_log,
                                         int n_actions, int obs_words, int state_words) {
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align256(off + bytes); return base ? base + o : nullptr; };
  unsigned char* p;
  p = take(sizeof(Counters)); if (t) t->ctr = (Counters*)p;
  p = take((size_t)cap_nodes * 4); if (t) t->node_visits = (int*)p;
  p = take((size_t)cap_nodes * 4); if (t) t->node_next = (int*)p;
  p = take((size_t)cap_nodes * 4); if (t) t->node_parent = (int*)p;
  p = take((size_t)cap_nodes); if (t) t->node_flags = p;
  p = take((size_t)cap_nodes); if (t) t->node_t = p;
  p = take((size_t)cap_nodes * 8 * obs_words); if (t) t->node_obs = (unsigned long long*)p;
  p = take((size_t)cap_nodes * n_actions * 4); if (t) t->an_first = (int*)p;
  p = take((size_t)cap_nodes * n_actions * 4); if (t) t->an_visits = (int*)p;
  p = take((size_t)cap_nodes * n_actions * 4); if (t) t->an_done = (int*)p;
  p = take((size_t)cap_nodes * n_actions * 8); if (t) t->an_total = (double*)p;
  p = take((size_t)cap_log * 4); if (t) t->log_node = (unsigned int*)p;
  p = take((size_t)cap_log * 4 * state_words); if (t) t->log_state = (unsigned int*)p;
  if (t) { t->cap_nodes = cap_nodes; t->cap_log = cap_log; }
  return off;
}

// order-preserving map double -> uint64 (for atomicMin / atomicMax)
__device__ __forceinline__ unsigned long long dkey(double d) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(d);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ULL);
}
__device__ __forceinline__ double dunkey(unsigned long long k) {
  const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffULL) : ~k;
  return __longlong_as_double((long long)b);
}
constexpr unsigned long long KEY_POS_INF = 0xFFF0000000000000ULL;  // dkey(+inf)
constexpr unsigned long long KEY_NEG_INF = 0x000FFFFFFFFFFFFFULL;  // dkey(-inf)

template <int OW>
__device__ __forceinline__ bool obs_eq(const unsigned long long* a, const unsigned long long (&b)[OW]) {
  bool e = true;
#pragma unroll
  for (int w = 0; w < OW; ++w) e &= (a[w] == b[w]);
  return e;
}

// Child obs node of action node `anode` with observation `obs`; created when missing
// (action_node.has_child / _add_obs_node, mcts.py:358-369).  Lock-free prepend to the
// action node's child list; returns -1 when the arena is full.
// `reserved` >= 0: node id set aside for the caller (one per simulation, see simulate) -- the new
// node is then born EXPANDED (flag bit 0 set before it is linked): its creator evaluates it.
// `reserved` == -2: the caller's reservation failed (arena full).  -1: allocate here.
template <int A, int OW>
__device__ int find_or_insert(const Tree& T, int anode, const unsigned long long (&obs)[OW], int t, bool& created,
                              int reserved = -1) {
  created = false;
  int head = *((volatile int*)&T.an_first[anode]);
  int stop = -1, mine = -1;
  while (true) {
    for (int c = head; c >= 0 && c != stop; c = *((volatile int*)&T.node_next[c]))
      if (obs_eq<OW>(T.node_obs + (size_t)c * OW, obs)) return c;  // (a spare `mine` stays unused)
    if (mine < 0 && reserved != -1) {
      if (reserved < 0) { atomicAdd(&T.ctr->overflow, 1u); return -1; }
      mine = reserved;
#pragma unroll
      for (int w = 0; w < OW; ++w) T.node_obs[(size_t)mine * OW + w] = obs[w];
      T.node_t[mine] = (unsigned char)t;
      T.node_parent[mine] = anode;
      atomicOr((unsigned int*)(T.node_flags + (mine & ~3)), 1u << (8 * (mine & 3)));
    }
    if (mine < 0) {
      // the last RESERVE slots are kept for the re-root (posg_tree_advance)
      if (*((volatile unsigned long long*)&T.ctr->node_count) + RESERVE >= (unsigned long long)T.cap_nodes) {
        atomicAdd(&T.ctr->overflow, 1u);
        return -1;
      }
      const unsigned long long id = atomicAdd(&T.ctr->node_count, 1ULL);
      if (id + RESERVE / 2 >= (unsigned long long)T.cap_nodes) { atomicAdd(&T.ctr->overflow, 1u); return -1; }
      mine = (int)id;
#pragma unroll
      for (int w = 0; w < OW; ++w) T.node_obs[(size_t)mine * OW + w] = obs[w];
      T.node_t[mine] = (unsigned char)t;
      T.node_parent[mine] = anode;
      // visits / flags / action-node arrays were zeroed (an_first: -1) by posg_tree_reset
    }
    T.node_next[mine] = head;
    // publish the node's fields before it becomes reachable (release; lighter than the
    // sequentially consistent fence of __threadfence())
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    const int old = atomicCAS(&T.an_first[anode], head, mine);
    if (old == head) { created = true; return mine; }
    stop = head;
    head = old;
  }
}

__device__ __forceinline__ double normalize(double v, double mn, double mx) {
  return (mx > mn) ? (v - mn) / (mx - mn) : v;
}

// Action for the lane with rank `rank` among the `gsize` lanes sharing `node` this level.
// `grp` = mask of those lanes.  Every lane of a group runs the same arithmetic.
template <int A>
__device__ int select_action(const Tree& T, const Params& P, int node, bool is_root, unsigned grp, int rank,
                             int max_gsize, int n_base, uint32_t rnd_word, unsigned desc_mask) {
  float n[A], q[A];
  // MinMaxStats.normalize (utils.py:34-39) in single precision: the scores are compared as floats
  // anyway, and ten double divisions per tree level cost more than the rest of the selection
  const float mn = (float)dunkey(*((volatile unsigned long long*)&T.ctr->min_key));
  const float mx = (float)dunkey(*((volatile unsigned long long*)&T.ctr->max_key));
  const bool ranged = mx > mn;
  const float inv_range = ranged ? __frcp_rn(mx - mn) : 1.0f;
  const float shift = ranged ? mn : 0.0f;
#pragma unroll
  for (int a = 0; a < A; ++a) {
    const int an = node * A + a;
    const int vis = *((volatile int*)&T.an_visits[an]);
    const int done = *((volatile int*)&T.an_done[an]);
    const double tot = *((volatile double*)&T.an_total[an]);
    n[a] = (float)vis;
    q[a] = done > 0 ? (__fdividef((float)tot, (float)done) - shift) * inv_range : 0.0f;
  }
  const float c = (float)P.c;
  const float prior = 1.0f / (float)A;  // RandomSearchPolicy.get_pi is uniform; so is the noise mix
  // per-action exploration factors kept incrementally: only the chosen action's count changes
  // between two consecutive virtual visits, so one rsqrt / one reciprocal per iteration
  float rs[A];   // ucb: 1 / sqrt(n_a)      pucb: 1 / (1 + n_a)
#pragma unroll
  for (int a = 0; a < A; ++a)
    rs[a] = (P.selection == SEL_PUCB) ? __frcp_rn(1.0f + n[a]) : (n[a] > 0.0f ? rsqrtf(n[a]) : 0.0f);
  // Large groups (near the root a whole warp shares a node): the g sequential UCB picks with frozen
  // values are the g largest "marginals" m_a(i) = Q_a + cs / sqrt(n_a + i) (m_a decreases in i, so
  // greedy == top-g; ties go to the lower action index).  The lane of rank r evaluates the
  // candidates (a, r) and counts in closed form how many marginals precede each: (a, r) is taken
  // iff fewer than g do.  O(A^2) per lane instead of O(g * A); cs is frozen at the group's middle
  // visit count.  PUCB (the default action_selection) has the same structure with 1 / (1 + n).  Only the per-action COUNTS matter (lanes are exchangeable), so lane r then takes
  // the action whose cumulative count first exceeds r.  Warp-uniform switch; max_gsize < 8 (and
  // wave size 1, the reference's sequential search) keeps the exact loop below.
  if ((P.selection == SEL_UCB || P.selection == SEL_PUCB) && max_gsize >= 8) {
    const bool pucb = P.selection == SEL_PUCB;
    const int g = __popc(grp);
    const int Nmid = n_base + (is_root ? 0 : 1) + (g >> 1);
    const float fN = (float)(Nmid > 1 ? Nmid : 1);
    // exploration scale: ucb  m_a(i) = Q_a + csm / sqrt(n_a + i);  pucb  m_a(i) = Q_a + csm / (1 + n_a + i)
    const float csm = pucb ? c * prior * sqrtf(fN) : c * sqrtf(__logf(fN));
    int k[A];
#pragma unroll
    for (int a = 0; a < A; ++a) {
      const float na = n[a] + (float)rank;
      const bool inf_a = !pucb && na <= 0.0f;        // ucb: an unvisited action's first visit is +inf
      const float v = q[a] + csm * (pucb ? __frcp_rn(1.0f + na) : rsqrtf(fmaxf(na, 1.0f)));
      float before = (float)rank;                    // (a, 0..rank-1)
#pragma unroll
      for (int b = 0; b < A; ++b) {
        if (b == a) continue;
        float cnt;
        if (inf_a) {
          cnt = (n[b] <= 0.0f && b < a) ? 1.0f : 0.0f;
        } else {
          const float diff = v - q[b];
          const float ratio = __fdividef(csm, diff);
          // m_b(i) > v  <=>  n_b + i < X (ucb: X = ratio^2; pucb: X = ratio - 1)
          const float X = pucb ? ratio - 1.0f : ratio * ratio;
          const float t = X - n[b], ti = rintf(t);
          const bool tie = fabsf(t - ti) <= fmaxf(1e-3f, 4e-5f * fabsf(X));
          cnt = fmaxf(tie ? ti + (b < a ? 1.0f : 0.0f) : ceilf(t), 0.0f);
          cnt = diff > 0.0f ? cnt : 1e9f;            // Q_b >= v: every marginal of b precedes
        }
        before += cnt;
      }
      k[a] = __popc(__ballot_sync(desc_mask, before < (float)g) & grp);
    }
    int cum = 0, mine = -1;
#pragma unroll
    for (int a = 0; a < A; ++a) {
      mine = (mine < 0 && rank < cum + k[a]) ? a : mine;
      cum += k[a];
    }
    if (mine < 0) {
      // rounding left fewer than g candidates: the next best marginal
      float best = -3.0e38f;
#pragma unroll
      for (int a = 0; a < A; ++a) {
        const float nn = n[a] + (float)k[a];
        const float v = pucb ? q[a] + csm * __frcp_rn(1.0f + nn) : (nn > 0.0f ? q[a] + csm * rsqrtf(nn) : 3.0e38f);
        if (v > best) { best = v; mine = a; }
      }
    }
    return mine;
  }
  // c * sqrt(ln N) over the group's consecutive visit counts N0 .. N0 + max_gsize - 1: exact per
  // iteration while N is small, linear in between the two ends once N0 >= 256 (relative error of the
  // interpolation < (max_gsize / N0)^2 / 8 < 2e-3 of a term that itself changes by < 1 %)
  const int N0 = n_base + (is_root ? 0 : 1);
  const bool interp = (P.selection == SEL_UCB) && N0 >= 256;
  float cs_lo = 0.0f, cs_step = 0.0f;
  if (interp) {
    cs_lo = c * sqrtf(__logf((float)N0));
    const float cs_hi = c * sqrtf(__logf((float)(N0 + max_gsize)));
    cs_step = (cs_hi - cs_lo) * __frcp_rn((float)max_gsize);
  }
  int mine = 0;
  for (int s = 0; s < max_gsize; ++s) {
    const int N = N0 + s;
    int choice = 0;
    if (N == 0) {
      // obs_node.visits == 0: random action (mcts.py:496,531); derived from the node so the
      // whole group agrees on it
      choice = (int)uniform_below(rnd_word * 0x9E3779B1u + (uint32_t)s * 0x85EBCA77u, (uint32_t)A);
    } else if (P.selection == SEL_UCB) {
      // argmax_a  norm(Q_a) + c * sqrt(ln N / n_a), first unvisited action first (mcts.py:529-546)
      const float cs = interp ? fmaf((float)s, cs_step, cs_lo) : c * sqrtf(__logf((float)N));
      float best = -3.0e38f;
      bool unvisited = false;
#pragma unroll
      for (int a = 0; a < A; ++a) {
        const bool zero = (n[a] == 0.0f);
        const float v = q[a] + cs * rs[a];
        const bool take = !unvisited && (zero || v > best);
        choice = take ? a : choice;
        best = take ? v : best;
        unvisited = unvisited || zero;
      }
    } else if (P.selection == SEL_PUCB) {
      const float cs = c * prior * sqrtf((float)N);
      float best = -3.0e38f;
#pragma unroll
      for (int a = 0; a < A; ++a) {
        const float v = (n[a] > 0.0f ? q[a] : 0.0f) + cs * rs[a];
        if (v > best) { best = v; choice = a; }
      }
    } else {
      float best = (float)N + 1.0f;
#pragma unroll
      for (int a = 0; a < A; ++a)
        if (n[a] < best) { best = n[a]; choice = a; }
    }
    if (s == rank) mine = choice;
    float nc = 0.0f;
#pragma unroll
    for (int a = 0; a < A; ++a) nc = (a == choice) ? n[a] : nc;
    nc += 1.0f;
    const float rc = (P.selection == SEL_PUCB) ? __frcp_rn(1.0f + nc) : rsqrtf(nc);
#pragma unroll
    for (int a = 0; a < A; ++a) {
      n[a] = (a == choice) ? nc : n[a];
      rs[a] = (a == choice) ? rc : rs[a];
    }
  }
  return mine;
}

// Optional per-simulation root selection (I-NTMCP level-0 phase, intmcp.py:410-442): the
// simulation first draws entry i of [0, n_sel) uniformly -- a particle of the PARENT tree's
// root belief -- and then searches this tree from node sel_root[i] with a particle drawn
// from particles[sel_off[i] .. sel_off[i] + sel_cnt[i]) (that node's belief).
struct RootSelect {
  const int* sel_root;
  const long long* sel_off;
  const int* sel_cnt;
  long long n_sel;   // 0: single root
};

// Model concept (see driving_sim.cuh):
//   using Blob; static constexpr int A, OW, SW;
//   struct State { uint32_t w[SW]; };
//   static void step(const Blob&, State&, int ego, int ego_action, uint32_t word,
//                    unsigned long long (&obs)[OW], double& reward, bool& done);
//   static void rollout_step(const Blob&, State&, int ego, uint32_t word, double& reward, bool& done);
template <class M>
__device__ void simulate(const typename M::Blob& B, const typename M::Ctx& ctx, const Tree& T, const Params& P,
                         int root, int root_t, const uint32_t* __restrict__ root_particles, long long n_particles,
                         const RootSelect& sel, unsigned long long sim_id, bool active) {
  constexpr int A = M::A, OW = M::OW, SW = M::SW;
  const unsigned full = 0xffffffffu;
  const unsigned lane = threadIdx.x & 31u;
  typename M::State st;
  int path_an[MAX_DEPTH];
  double path_r[MAX_DEPTH];
  int depth = 0, node = root, t = root_t, npath = 0;
  double leaf = 0.0;
  bool descending = active;
  bool virtual_leaf = false;  // arena full: the state just reached is evaluated but not stored
  bool need_rollout = false;  // this lane expanded a leaf: evaluate it after the warp's descent
  int roll_depth = 0, roll_t = 0;
  if (active) {
    const uint4 w = philox_block(P.seed, sim_id, 0xFFFFu, STREAM_BELIEF);
    long long pi;
    if (sel.n_sel > 0) {
      const long long e = (long long)(((unsigned long long)w.x * (unsigned long long)sel.n_sel) >> 32);
      root = sel.sel_root[e];
      pi = sel.sel_off[e] + (long long)(((unsigned long long)w.y * (unsigned long long)sel.sel_cnt[e]) >> 32);
    } else {
      pi = (long long)(((unsigned long long)w.x * (unsigned long long)n_particles) >> 32);
    }
#pragma unroll
    for (int k = 0; k < SW; ++k) st.w[k] = __ldg(root_particles + pi * SW + k);
    node = root;
  }
  // A simulation adds at most one node (the node it creates is born expanded and evaluated by its
  // creator), so the warp reserves one id per simulation with ONE atomic: a wave of 2^18 simulations
  // bumping the node counter one by one serialises in L2.  Ids of simulations that end without
  // creating a node stay unlinked (zeroed) holes; the re-root compaction drops them.
  int reserved = -2;
  {
    const unsigned am = __ballot_sync(full, active);
    if (am) {
      const int ldr = __ffs(am) - 1, cnt = __popc(am);
      unsigned long long base = 0;
      if ((int)lane == ldr) {
        base = ~0ULL;   // arena full (the last RESERVE slots are kept for the re-root)
        if (*((volatile unsigned long long*)&T.ctr->node_count) + RESERVE + 32 < (unsigned long long)T.cap_nodes)
          base = atomicAdd(&T.ctr->node_count, (unsigned long long)cnt);
      }
      base = __shfl_sync(full, base, ldr);
      const unsigned long long id = base + (unsigned long long)__popc(am & ((1u << lane) - 1u));
      if (active && base != ~0ULL && id + RESERVE / 2 < (unsigned long long)T.cap_nodes) reserved = (int)id;
    }
  }
  bool created_here = false;  // this lane created (and therefore expands) the node it stands on

  while (__any_sync(full, descending)) {
    // ---- stop conditions / leaf expansion ------------------------------------------
    bool expand = false;
    if (descending) {
      const int t_rule = P.intmcp_depth_rule ? t + depth : t;
      if (depth > P.depth_limit || t_rule > P.step_limit || depth >= MAX_DEPTH - 1) {
        descending = false;  // returns 0 (mcts.py:315-316)
      } else if (virtual_leaf) {
        expand = true;
        descending = false;
      } else if (created_here) {
        expand = true;
        descending = false;
      } else {
        const unsigned char old = T.node_flags[node];
        if (!(old & 1u)) {
          // first simulation to arrive expands the node and evaluates it by a rollout
          const unsigned int word = atomicOr((unsigned int*)(T.node_flags + (node & ~3)), 1u << (8 * (node & 3)));
          expand = !((word >> (8 * (node & 3))) & 1u);
        }
        if (expand) descending = false;
      }
    }
    // The leaf rollout is deferred until every lane of the warp has finished its descent, so that
    // the lanes roll out side by side instead of one tree level after the other (the rollout only
    // reads the lane's own state and its own Philox counters: the result is the same).
    if (expand) { need_rollout = true; roll_depth = depth; roll_t = t; }
    if (!__any_sync(full, descending)) break;

    // ---- action selection (water-filling among lanes sharing the node) ---------------
    const unsigned grp = __match_any_sync(full, descending ? node : -1 - (int)lane);
    const int rank = __popc(grp & ((1u << lane) - 1u));
    const int gsize = descending ? __popc(grp) : 0;
    const int max_g = __reduce_max_sync(full, gsize);
    uint4 w = make_uint4(0, 0, 0, 0);
    int action = 0, n_base = 0;
    if (descending) {
      w = philox_block(P.seed, sim_id, (uint32_t)depth, STREAM_TREE);
      n_base = *((volatile int*)&T.node_visits[node]);
      if (node != root) n_base -= gsize;  // the group's own entries were already counted
      if (n_base < 0) n_base = 0;
    }
    // group-consistent random word: the group leader's
    const uint32_t rnd = __shfl_sync(full, w.z, __ffs(grp) - 1);
    const unsigned desc_mask = __ballot_sync(full, descending);
    if (descending)
      action = select_action<A>(T, P, node, node == root, grp, rank, max_g, n_base, rnd ^ (uint32_t)node, desc_mask);
    {
      // virtual visit, one atomic per distinct action node of the warp: near the root thousands of
      // simulations per wave hit the same few counters, and same-address atomics serialise in L2
      const unsigned va = __match_any_sync(full, descending ? node * A + action : -1 - (int)lane);
      if (descending && (int)lane == __ffs(va) - 1) atomicAdd(&T.an_visits[node * A + action], __popc(va));
    }

    // ---- model step, child lookup, particle log ---------------------------------------
    unsigned long long obs[OW];
#pragma unroll
    for (int k = 0; k < OW; ++k) obs[k] = 0ULL;
    double r = 0.0;
    bool done = false;
    int anode = -1;
    if (descending) {
      M::step(B, ctx, st, P.ego, action, w.y, w.w, obs, r, done);
      anode = node * A + action;
    }
    // lanes with the same (action node, observation) resolve the child once
    unsigned same = __match_any_sync(full, descending ? anode : -1 - (int)lane);
#pragma unroll
    for (int k = 0; k < OW; ++k) same &= __match_any_sync(full, obs[k]);
    const int leader = __ffs(same) - 1;
    int child = -1;
    if (descending && (int)lane == leader) {
      bool created;
      child = find_or_insert<A, OW>(T, anode, obs, t + 1, created, reserved);
      created_here = created;
      if (child >= 0) {
        atomicAdd(&T.node_visits[child], __popc(same));
        // child_obs_node.is_absorbing = ego_done (mcts.py:370)
        if (done) atomicOr((unsigned int*)(T.node_flags + (child & ~3)), 2u << (8 * (child & 3)));
        else atomicAnd((unsigned int*)(T.node_flags + (child & ~3)), ~(2u << (8 * (child & 3))));
      }
    }
    child = __shfl_sync(full, child, leader < 0 ? 0 : leader);
    // child_obs_node.belief.add_particle(next_hps) (mcts.py:371): one counter bump per warp
    const unsigned app = __ballot_sync(full, descending && child >= 0);
    unsigned long long log_base = 0;
    if (app) {
      const int app_leader = __ffs(app) - 1;
      if ((int)lane == app_leader) log_base = atomicAdd(&T.ctr->log_count, (unsigned long long)__popc(app));
      log_base = __shfl_sync(full, log_base, app_leader);
    }
    if (descending) {
      path_an[npath] = anode;
      path_r[npath] = r;
      ++npath;
      if (child < 0) {
        // arena full: no node for this history -- evaluate it by a rollout, store nothing
        ++depth; ++t;
        virtual_leaf = true;
        if (done) descending = false;
      } else {
        const unsigned long long li = log_base + (unsigned long long)__popc(app & ((1u << lane) - 1u));
        if (li < (unsigned long long)T.cap_log) {
          T.log_node[li] = (unsigned int)child;
#pragma unroll
          for (int k = 0; k < SW; ++k) T.log_state[li * SW + k] = st.w[k];
        }
        node = child; ++depth; ++t;
        if (done) descending = false;  // future return not added (mcts.py:373-378)
      }
    }
  }

  // ---- leaf evaluation: MCTS._rollout (mcts.py:405-452), uniform random joint actions until
  // done / limits; all expanding lanes of the warp together
  if (need_rollout) {
    int d = roll_depth, tt = roll_t, k = 0;
    double ret = 0.0;
    uint4 blk = make_uint4(0, 0, 0, 0);
    while (d <= P.depth_limit && tt <= P.step_limit && k < MAX_DEPTH) {
      if ((k & 3) == 0) blk = philox_block(P.seed, sim_id, 0x10000u + (uint32_t)(k >> 2), STREAM_TREE);
      const uint32_t word = (k & 3) == 0 ? blk.x : (k & 3) == 1 ? blk.y : (k & 3) == 2 ? blk.z : blk.w;
      double r; bool done;
      M::rollout_step(B, ctx, st, P.ego, word, r, done);
      ret = __dadd_rn(ret, __dmul_rn(P.disc_pow[k], r));
      if (done) break;
      ++d; ++tt; ++k;
    }
    leaf = ret;
  }

  // ---- backup (mcts.py:378-381; node.py:166-178) --------------------------------------
  // Warp-synchronous and aggregated: at path position i the lanes that share an action node add
  // their returns with ONE atomicAdd (same-address atomics serialise in L2, and near the root a wave
  // of simulations shares a handful of nodes).  MinMaxStats sees the node's mean after the group's
  // update (the reference: after each single update).
  {
    double G = leaf;
    for (int i = npath - 1; i >= 0; --i) {
      G = __dadd_rn(path_r[i], __dmul_rn(P.discount, G));
      path_r[i] = G;
    }
    const int max_np = __reduce_max_sync(full, active ? npath : 0);
    // MinMaxStats bounds, read once: the lane's own updates are folded into the local copy (another
    // warp's concurrent update can only make the atomics below redundant, never wrong)
    unsigned long long kmin = *((volatile unsigned long long*)&T.ctr->min_key);
    unsigned long long kmax = *((volatile unsigned long long*)&T.ctr->max_key);
    unsigned long long gmin = *((volatile unsigned long long*)&T.ctr->gmin_key);
    unsigned long long gmax = *((volatile unsigned long long*)&T.ctr->gmax_key);
    for (int i = 0; i < max_np; ++i) {
      const bool on = active && i < npath;
      const int an = on ? path_an[i] : -1 - (int)lane;
      const double g = on ? path_r[i] : 0.0;
      if (on) {
        const unsigned long long gk = dkey(g);
        if (gk < gmin) { atomicMin(&T.ctr->gmin_key, gk); gmin = gk; }
        if (gk > gmax) { atomicMax(&T.ctr->gmax_key, gk); gmax = gk; }
      }
      const unsigned m = __match_any_sync(full, an);
      const int cnt = __popc(m);
      const int max_cnt = __reduce_max_sync(full, on ? cnt : 1);
      double sum = g;
      if (max_cnt > 1) {
        // every lane sums its own group (itself included) in lane order
        sum = 0.0;
        unsigned rest = m;
        for (int k = 0; k < max_cnt; ++k) {
          const int src = rest ? __ffs(rest) - 1 : (int)lane;
          const double v = __shfl_sync(full, g, src);
          if (rest) sum += v;
          rest &= rest - 1u;
        }
      }
      if (on && (int)lane == __ffs(m) - 1) {
        const double old_tot = atomicAdd(&T.an_total[an], sum);
        const int old_done = atomicAdd(&T.an_done[an], cnt);
        // (old_tot, old_done) come from two atomics: another warp's backup of the same node can fall between
        // them.  A mean is never outside the range of the single returns, so clamp to it (a no-op for wave 1).
        unsigned long long vk = dkey((old_tot + sum) / (double)(old_done + cnt));
        const unsigned long long mk = dkey(sum / (double)cnt);   // the group's own mean is inside that range too
        const unsigned long long lo = gmin < mk ? gmin : mk, hi = gmax > mk ? gmax : mk;
        vk = vk < lo ? lo : (vk > hi ? hi : vk);
        // MinMaxStats.update (utils.py:29-32); the atomics fire only when the bound moves
        if (vk < kmin) { atomicMin(&T.ctr->min_key, vk); kmin = vk; }
        if (vk > kmax) { atomicMax(&T.ctr->max_key, vk); kmax = vk; }
      }
    }
  }
  const unsigned act = __ballot_sync(full, active);
  const unsigned int dmax = __reduce_max_sync(full, active ? (unsigned int)depth : 0u);
  // self.root.visits += 1 per simulation (mcts.py:288); lanes may have different roots
  const unsigned same_root = __match_any_sync(full, active ? root : -1 - (int)lane);
  if (active && (int)lane == __ffs(same_root) - 1) atomicAdd(&T.node_visits[root], __popc(same_root));
  const unsigned cr = __ballot_sync(full, active && created_here);
  if (cr && (int)lane == __ffs(cr) - 1) atomicAdd(&T.ctr->nodes_created, (unsigned long long)__popc(cr));
  if (act && (int)lane == __ffs(act) - 1) {
    atomicAdd(&T.ctr->sims_done, (unsigned long long)__popc(act));
    if (dmax > *((volatile unsigned int*)&T.ctr->max_depth)) atomicMax(&T.ctr->max_depth, dmax);
  }
}

}  // namespace search
}  // namespace posg
