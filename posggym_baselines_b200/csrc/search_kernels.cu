// Search-tree and particle-belief entry points of the C-ABI (model-generic; dispatched
// on the model kind).  See search.cuh / belief.cuh for the algorithms.
#include <chrono>
#include <cooperative_groups.h>

#include <cmath>
#include <cstring>

#include "belief.cuh"
#include "common.h"
#include "driving_sim.cuh"
#include "pp.cuh"
#include "lbf.cuh"
#include "nested_sim.cuh"
#include "search.cuh"
#include "stage.cuh"

namespace cg = cooperative_groups;

namespace posg {

#ifndef POSG_SEARCH_THREADS
#define POSG_SEARCH_THREADS 256
#endif
#ifndef POSG_SEARCH_MIN_CTAS
#define POSG_SEARCH_MIN_CTAS 4
#endif
constexpr int SEARCH_THREADS = POSG_SEARCH_THREADS;

template <class M>
__global__ void __launch_bounds__(SEARCH_THREADS, POSG_SEARCH_MIN_CTAS)
search_kernel(const typename M::Blob* __restrict__ gblob, uint32_t blob_size, search::Tree T,
              const __grid_constant__ search::Params P, const __grid_constant__ typename M::Ctx ctx, int root,
              int root_t, const uint32_t* __restrict__ particles, long long n_particles, search::RootSelect sel,
              long long n_sims, unsigned long long sim_id0) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ alignas(8) uint64_t mbar;
  stage_blob(smem_raw, gblob, blob_size, &mbar);
  const typename M::Blob& B = *reinterpret_cast<const typename M::Blob*>(smem_raw);
  const long long i = (long long)blockIdx.x * SEARCH_THREADS + threadIdx.x;
  // whole warps enter (warp collectives inside); lanes past n_sims are inactive
  search::simulate<M>(B, ctx, T, P, root, root_t, particles, n_particles, sel, sim_id0 + (unsigned long long)i, i < n_sims);
}

// n_waves consecutive waves of up to 1024 simulations in ONE launch: a single CTA runs wave after wave with a
// block barrier in between, so wave w + 1 sees every backup of wave w exactly as separate launches would --
// without the launch + host synchronisation per wave that dominates small waves (a 32-simulation wave is one
// warp and ~40 us of dependent work).  Small waves are what keeps the search equivalent to the reference's
// sequential loop (tests/test_pomcp_gpu.py: a wave may be ~1/2048 of the root's visits).
template <class M>
__global__ void __launch_bounds__(1024, 1)
search_waves_kernel(const typename M::Blob* __restrict__ gblob, uint32_t blob_size, search::Tree T,
                    const __grid_constant__ search::Params P, const __grid_constant__ typename M::Ctx ctx, int root,
                    int root_t, const uint32_t* __restrict__ particles, long long n_particles, search::RootSelect sel,
                    int wave_size, int n_waves, unsigned long long sim_id0) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ alignas(8) uint64_t mbar;
  stage_blob(smem_raw, gblob, blob_size, &mbar);
  const typename M::Blob& B = *reinterpret_cast<const typename M::Blob*>(smem_raw);
  for (int w = 0; w < n_waves; ++w) {
    search::simulate<M>(B, ctx, T, P, root, root_t, particles, n_particles, sel,
                        sim_id0 + (unsigned long long)w * (unsigned long long)wave_size + threadIdx.x,
                        (int)threadIdx.x < wave_size);
    __threadfence();
    __syncthreads();
  }
}

// Root parallelisation INSIDE one GPU: CTA k runs search_waves_kernel's loop on its own tree (entries[k]: arena,
// root belief, root node), with its own Philox sub-sequence.  Every tree is the reference's search on its own
// (one tree per SM instead of one tree per GPU); the root statistics are summed afterwards
// (posg_forest_stats + the host's final selection), exactly like the cross-GPU allreduce of root-parallel POMCP.
struct ForestEntry {
  unsigned char* arena;
  const uint32_t* particles;
  long long n_particles;
  int root;
  int wave;   // simulations per wave of THIS tree (0: the launch's wave_size) -- every tree follows its own schedule
};
static_assert(sizeof(ForestEntry) == sizeof(posg_forest_entry), "posg_forest_entry layout");

template <class M>
__global__ void __launch_bounds__(1024, 1)
search_forest_kernel(const typename M::Blob* __restrict__ gblob, uint32_t blob_size,
                     const ForestEntry* __restrict__ entries, long long cap_nodes, long long cap_log,
                     const __grid_constant__ search::Params P, const __grid_constant__ typename M::Ctx ctx, int root_t,
                     int wave_size, int n_waves, unsigned long long sim_id0) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ alignas(8) uint64_t mbar;
  stage_blob(smem_raw, gblob, blob_size, &mbar);
  const typename M::Blob& B = *reinterpret_cast<const typename M::Blob*>(smem_raw);
  const ForestEntry en = entries[blockIdx.x];
  search::Tree T;
  search::layout(&T, en.arena, cap_nodes, cap_log, M::A, M::OW, M::SW);
  const search::RootSelect sel = {nullptr, nullptr, nullptr, 0};
  const unsigned long long base = sim_id0 + ((unsigned long long)blockIdx.x << 32);
  const int wave = en.wave > 0 ? (en.wave < wave_size ? en.wave : wave_size) : wave_size;
  for (int w = 0; w < n_waves; ++w) {
    // (counters advance by the launch's wave_size, the largest wave of the forest: unique whatever `wave` is)
    search::simulate<M>(B, ctx, T, P, en.root, root_t, en.particles, en.n_particles, sel,
                        base + (unsigned long long)w * (unsigned long long)wave_size + threadIdx.x,
                        (int)threadIdx.x < wave);
    __threadfence();
    __syncthreads();
  }
}

// per tree: what posg_tree_node_stats returns for the tree's root (3 * A + 10 doubles)
__global__ void forest_stats_kernel(const ForestEntry* __restrict__ entries, long long cap_nodes, long long cap_log,
                                    int A, int OW, int SW, double* __restrict__ out) {
  const ForestEntry en = entries[blockIdx.x];
  search::Tree T;
  search::layout(&T, en.arena, cap_nodes, cap_log, A, OW, SW);
  double* o = out + (size_t)blockIdx.x * (3 * A + 10);
  const int a = threadIdx.x;
  if (a < A) {
    o[a] = (double)T.an_visits[(size_t)en.root * A + a];
    o[A + a] = (double)T.an_done[(size_t)en.root * A + a];
    o[2 * A + a] = T.an_total[(size_t)en.root * A + a];
  }
  if (a == 0) {
    double* e = o + 3 * A;
    const search::Counters c = *T.ctr;
    e[0] = (double)T.node_visits[en.root]; e[1] = (double)c.node_count; e[2] = (double)c.log_count;
    e[3] = (double)c.max_depth; e[4] = search::dunkey(c.min_key); e[5] = search::dunkey(c.max_key);
    e[6] = (double)c.overflow; e[7] = (double)c.sims_done; e[8] = (double)T.node_flags[en.root];
    e[9] = (double)c.nodes_created;
  }
}

template <class M>
__global__ void __launch_bounds__(belief::CTA)
reinvigorate_kernel(const typename M::Blob* __restrict__ gblob, uint32_t blob_size,
                    const __grid_constant__ typename M::Ctx ctx, const uint32_t* __restrict__ parent, long long n_parent, int ego, int action,
                    const unsigned long long* __restrict__ target_obs, long long n_needed, long long max_attempts,
                    uint64_t seed, uint64_t draw_id, uint32_t* __restrict__ out, uint32_t* __restrict__ rej,
                    long long* __restrict__ result) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ alignas(8) uint64_t mbar;
  stage_blob(smem_raw, gblob, blob_size, &mbar);
  const typename M::Blob& B = *reinterpret_cast<const typename M::Blob*>(smem_raw);
  belief::reinvigorate<M>(B, ctx, parent, n_parent, ego, action, target_obs, n_needed, max_attempts, seed, draw_id, out,
                          rej, result);
}

template <class M>
__global__ void __launch_bounds__(belief::CTA)
initial_belief_kernel(const typename M::Blob* __restrict__ gblob, uint32_t blob_size, int ego,
                      const unsigned long long* __restrict__ target_obs, long long n_needed, long long max_attempts,
                      uint64_t seed, uint64_t draw_id, uint32_t* __restrict__ out, long long* __restrict__ result) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ alignas(8) uint64_t mbar;
  stage_blob(smem_raw, gblob, blob_size, &mbar);
  const typename M::Blob& B = *reinterpret_cast<const typename M::Blob*>(smem_raw);
  belief::initial_belief<M>(B, ego, target_obs, n_needed, max_attempts, seed, draw_id, out, result);
}

// arena initialisation: zero everything, an_first = -1, root node 0, min/max keys
__global__ void tree_init_kernel(search::Tree T, int n_actions, int root_t, double known_min, double known_max) {
  const long long n = T.cap_nodes * n_actions;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    T.an_first[i] = -1;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    T.ctr->node_count = 1;  // node 0 = root
    T.ctr->log_count = 0;
    const bool known = known_min <= known_max;
    T.ctr->min_key = known ? search::dkey(known_min) : search::KEY_POS_INF;
    T.ctr->max_key = known ? search::dkey(known_max) : search::KEY_NEG_INF;
    T.ctr->max_depth = 0; T.ctr->overflow = 0; T.ctr->sims_done = 0; T.ctr->nodes_created = 0;
    T.ctr->gmin_key = search::KEY_POS_INF; T.ctr->gmax_key = search::KEY_NEG_INF;
    T.node_t[0] = (unsigned char)root_t;
    T.node_parent[0] = -1;
  }
}


// ---- re-rooting: keep the subtree under the new root (MCTS._update, mcts.py:229-263: "remove
// reference to parent, effectively pruning dead branches") by copying it, renumbered breadth-first,
// into a second arena.  Everything is sized by the KEPT subtree, not by the arena:
//  * reroot_bfs_kernel: ONE cooperative launch walks the subtree level by level (grid barrier between
//    levels, no host round trips); order[new id] = old id, mark[old id] = new id; the appends to the
//    frontier are aggregated per warp (one atomic per warp and list step, not one per node);
//  * reroot_copy_kernel: new-id major, so every write of the destination arrays is coalesced;
//  * reroot_clear_kernel: the destination's stale slots [kept, its previous high-water mark) are
//    re-initialised (the arenas ping-pong; a slot past the high-water mark is still clean).
__global__ void __launch_bounds__(256)
reroot_bfs_kernel(search::Tree S, int n_actions, int new_root, int* __restrict__ mark, int* __restrict__ order,
                  unsigned int* __restrict__ counters /* [0] nodes kept */) {
  cg::grid_group grid = cg::this_grid();
  const unsigned lane = threadIdx.x & 31u;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long long)gridDim.x * blockDim.x;
  if (tid == 0) { order[0] = new_root; mark[new_root] = 0; counters[0] = 1u; }
  grid.sync();
  unsigned lo = 0, hi = 1;
  while (lo < hi) {
    const long long work = (long long)(hi - lo) * n_actions;
    // whole warps iterate together (the ballots below need every lane of the warp)
    for (long long i0 = tid - lane; i0 < work; i0 += nthr) {
      const long long i = i0 + lane;
      int c = -1;
      if (i < work) c = S.an_first[(long long)order[lo + i / n_actions] * n_actions + (int)(i % n_actions)];
      while (true) {
        const unsigned has = __ballot_sync(0xffffffffu, c >= 0);
        if (!has) break;
        unsigned base = 0;
        const int leader = __ffs(has) - 1;
        if ((int)lane == leader) base = atomicAdd(&counters[0], (unsigned)__popc(has));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (c >= 0) {
          const unsigned id = base + __popc(has & ((1u << lane) - 1u));
          order[id] = c;
          mark[c] = (int)id;
          c = S.node_next[c];
        }
      }
    }
    grid.sync();
    lo = hi;
    hi = *((volatile unsigned int*)&counters[0]);
    grid.sync();   // every thread has read the level's end before the next level appends
  }
}

// the same walk by ONE CTA (block barrier between levels) for small trees -- the trees of a forest: a plain launch,
// so the re-roots that several host threads issue on their own streams run side by side (cooperative launches from
// several streams were seen to stall each other, and every other CUDA call of the process, for tens of milliseconds)
__global__ void __launch_bounds__(1024)
reroot_bfs_cta_kernel(search::Tree S, int n_actions, int new_root, int* __restrict__ mark, int* __restrict__ order,
                      unsigned int* __restrict__ counters /* [0] nodes kept */) {
  __shared__ unsigned s_hi;
  const unsigned lane = threadIdx.x & 31u;
  const long long tid = threadIdx.x, nthr = blockDim.x;
  if (tid == 0) { order[0] = new_root; mark[new_root] = 0; counters[0] = 1u; }
  __syncthreads();
  unsigned lo = 0, hi = 1;
  while (lo < hi) {
    const long long work = (long long)(hi - lo) * n_actions;
    for (long long i0 = tid - lane; i0 < work; i0 += nthr) {
      const long long i = i0 + lane;
      int c = -1;
      if (i < work) c = S.an_first[(long long)order[lo + i / n_actions] * n_actions + (int)(i % n_actions)];
      while (true) {
        const unsigned has = __ballot_sync(0xffffffffu, c >= 0);
        if (!has) break;
        unsigned base = 0;
        const int leader = __ffs(has) - 1;
        if ((int)lane == leader) base = atomicAdd(&counters[0], (unsigned)__popc(has));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (c >= 0) {
          const unsigned id = base + __popc(has & ((1u << lane) - 1u));
          order[id] = c;
          mark[c] = (int)id;
          c = S.node_next[c];
        }
      }
    }
    __syncthreads();            // (block-scope barrier: the CTA's global writes above are visible to the CTA)
    if (tid == 0) s_hi = *((volatile unsigned int*)&counters[0]);
    __syncthreads();
    lo = hi;
    hi = s_hi;
  }
}

__global__ void reroot_copy_kernel(search::Tree S, search::Tree D, int n_actions, int obs_words,
                                   const int* __restrict__ mark, const int* __restrict__ order,
                                   const unsigned int* __restrict__ counters) {
  const long long n_new = counters[0];
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n_new;
       j += (long long)gridDim.x * blockDim.x) {
    const long long i = order[j];
    D.node_visits[j] = S.node_visits[i];
    D.node_flags[j] = S.node_flags[i];
    D.node_t[j] = S.node_t[i];
    for (int k = 0; k < obs_words; ++k) D.node_obs[(size_t)j * obs_words + k] = S.node_obs[(size_t)i * obs_words + k];
    const int nx = S.node_next[i], pa = S.node_parent[i];
    D.node_next[j] = (j == 0 || nx < 0) ? -1 : mark[nx];
    D.node_parent[j] = (j == 0 || pa < 0) ? -1 : mark[pa / n_actions] * n_actions + pa % n_actions;
  }
  // action-node arrays: one thread per (new node, action), new-id major
  const long long n_an = n_new * n_actions;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n_an; e += (long long)gridDim.x * blockDim.x) {
    const long long j = e / n_actions;
    const long long so = (long long)order[j] * n_actions + (e - j * n_actions);
    const int first = S.an_first[so];
    D.an_first[e] = first >= 0 ? mark[first] : -1;
    D.an_visits[e] = S.an_visits[so];
    D.an_done[e] = S.an_done[so];
    D.an_total[e] = S.an_total[so];
  }
}

// slots [kept, dirty) of the destination back to their initial state (posg_tree_reset's)
__global__ void reroot_clear_kernel(search::Tree D, int n_actions, int obs_words, const unsigned int* __restrict__ counters,
                                    long long dirty) {
  const long long n_new = counters[0];
  for (long long j = n_new + (long long)blockIdx.x * blockDim.x + threadIdx.x; j < dirty;
       j += (long long)gridDim.x * blockDim.x) {
    D.node_visits[j] = 0; D.node_next[j] = 0; D.node_parent[j] = 0; D.node_flags[j] = 0; D.node_t[j] = 0;
    for (int k = 0; k < obs_words; ++k) D.node_obs[(size_t)j * obs_words + k] = 0ULL;
  }
  for (long long e = n_new * n_actions + (long long)blockIdx.x * blockDim.x + threadIdx.x; e < dirty * n_actions;
       e += (long long)gridDim.x * blockDim.x) {
    D.an_first[e] = -1; D.an_visits[e] = 0; D.an_done[e] = 0; D.an_total[e] = 0.0;
  }
}

__global__ void reroot_log_kernel(search::Tree S, search::Tree D, int state_words, const int* __restrict__ mark,
                                  unsigned long long n_log, unsigned long long* __restrict__ out_count) {
  const unsigned lane = threadIdx.x & 31u;
  for (unsigned long long i0 = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) - lane; i0 < n_log;
       i0 += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long i = i0 + lane;
    const int j = i < n_log ? mark[S.log_node[i]] : -1;
    const unsigned b = __ballot_sync(0xffffffffu, j >= 0);
    if (!b) continue;
    unsigned long long base = 0;
    const int leader = __ffs(b) - 1;
    if ((int)lane == leader) base = atomicAdd(out_count, (unsigned long long)__popc(b));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (j >= 0) {
      const unsigned long long pos = base + __popc(b & ((1u << lane) - 1u));
      D.log_node[pos] = (unsigned int)j;
      for (int k = 0; k < state_words; ++k) D.log_state[pos * state_words + k] = S.log_state[i * state_words + k];
    }
  }
}

__global__ void reroot_finish_kernel(search::Tree S, search::Tree D, const unsigned int* __restrict__ counters,
                                     const unsigned long long* __restrict__ log_count) {
  D.ctr->node_count = counters[0];
  D.ctr->log_count = *log_count;
  D.ctr->min_key = S.ctr->min_key;
  D.ctr->max_key = S.ctr->max_key;
  D.ctr->max_depth = 0; D.ctr->overflow = 0; D.ctr->nodes_created = 0;
  D.ctr->gmin_key = S.ctr->gmin_key; D.ctr->gmax_key = S.ctr->gmax_key;
  D.ctr->sims_done = S.ctr->sims_done;
}

// ActionNode.visits / total_value of a node's children as doubles, left on the device (the buffer of
// the root-parallel allreduce)
__global__ void root_stats_kernel(search::Tree T, int n_actions, int node, double* __restrict__ out) {
  const int a = threadIdx.x;
  if (a < n_actions) {
    out[a] = (double)T.an_done[(size_t)node * n_actions + a];
    out[n_actions + a] = T.an_total[(size_t)node * n_actions + a];
  }
}

template <int OW>
__global__ void tree_advance_kernel(search::Tree T, int n_actions, int root, int action,
                                    const unsigned long long* __restrict__ obs_in, int create, int* __restrict__ out) {
  unsigned long long obs[OW];
#pragma unroll
  for (int k = 0; k < OW; ++k) obs[k] = obs_in[k];
  const int anode = root * n_actions + action;
  int found = -1;
  for (int c = T.an_first[anode]; c >= 0; c = T.node_next[c])
    if (search::obs_eq<OW>(T.node_obs + (size_t)c * OW, obs)) { found = c; break; }
  int created = 0;
  if (found < 0 && create) {
    unsigned long long id = atomicAdd(&T.ctr->node_count, 1ULL);
    // search never takes the last RESERVE/2 slots; when it has filled the arena the last
    // slot serves as the (transient) new root -- the caller rebuilds the tree right after
    if (id + search::RESERVE / 2 >= (unsigned long long)T.cap_nodes) id = (unsigned long long)T.cap_nodes - 1;
    found = (int)id;
#pragma unroll
    for (int k = 0; k < OW; ++k) T.node_obs[(size_t)found * OW + k] = obs[k];
    T.node_t[found] = (unsigned char)(T.node_t[root] + 1);
    T.node_parent[found] = anode;
    T.node_visits[found] = 0;
    T.node_flags[found] = 0;
    for (int a = 0; a < n_actions; ++a) {
      T.an_first[found * n_actions + a] = -1;
      T.an_visits[found * n_actions + a] = 0;
      T.an_done[found * n_actions + a] = 0;
      T.an_total[found * n_actions + a] = 0.0;
    }
    T.node_next[found] = T.an_first[anode];
    T.an_first[anode] = found;
    // obs_node.is_absorbing = self.root.is_absorbing (mcts.py:250)
    if (T.node_flags[root] & 2u) T.node_flags[found] |= 2u;
    created = 1;
  }
  out[0] = found;
  out[1] = created;
  out[2] = found >= 0 ? (int)T.node_flags[found] : 0;
  out[3] = found >= 0 ? (int)T.node_t[found] : 0;
}

// traverse one (action, observation) step for a batch of (parent node, action, obs) triples,
// creating missing nodes with zero visits (INTMCP.traverse, intmcp.py:797-809)
template <int A, int OW>
__global__ void tree_advance_batch_kernel(search::Tree T, long long n, const int* __restrict__ parents,
                                          const int* __restrict__ actions, const unsigned long long* __restrict__ obs_in,
                                          int* __restrict__ children) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long obs[OW];
#pragma unroll
  for (int k = 0; k < OW; ++k) obs[k] = obs_in[i * OW + k];
  const int parent = parents[i];
  int child = -1;
  if (parent >= 0) {
    bool created;
    child = search::find_or_insert<A, OW>(T, parent * A + actions[i], obs, (int)T.node_t[parent] + 1, created);
  }
  children[i] = child;
}

// parent obs node, action and observation of a batch of nodes (for nested belief updates)
template <int OW>
__global__ void tree_node_info_kernel(search::Tree T, int n_actions, long long n, const int* __restrict__ nodes,
                                      int* __restrict__ parent_out, int* __restrict__ action_out,
                                      unsigned long long* __restrict__ obs_out, int* __restrict__ t_out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int node = nodes[i];
  const int an = T.node_parent[node];
  parent_out[i] = an / n_actions;
  action_out[i] = an % n_actions;
#pragma unroll
  for (int k = 0; k < OW; ++k) obs_out[i * OW + k] = T.node_obs[(size_t)node * OW + k];
  t_out[i] = T.node_t[node];
}

// belief of `node`: every logged particle that passed through it (order irrelevant)
__global__ void gather_belief_kernel(search::Tree T, int state_words, int node, uint32_t* __restrict__ out,
                                     long long max_out, unsigned long long* __restrict__ count) {
  const unsigned long long n = T.ctr->log_count < (unsigned long long)T.cap_log ? T.ctr->log_count
                                                                                 : (unsigned long long)T.cap_log;
  const unsigned lane = threadIdx.x & 31u;
  for (unsigned long long i0 = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) - lane; i0 < n;
       i0 += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long i = i0 + lane;
    const bool hit = i < n && T.log_node[i] == (unsigned int)node;
    const unsigned b = __ballot_sync(0xffffffffu, hit);
    if (!b) continue;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(count, (unsigned long long)__popc(b));
    base = __shfl_sync(0xffffffffu, base, 0);
    const unsigned long long pos = base + __popc(b & ((1u << lane) - 1u));
    if (hit && pos < (unsigned long long)max_out)
      for (int k = 0; k < state_words; ++k) out[pos * state_words + k] = T.log_state[i * state_words + k];
  }
}

struct KindInfo { int n_actions, obs_words, state_words; };

static bool kind_info(const posg_model* m, KindInfo* k) {
  if (!m) return false;
  switch (m->kind) {
    case KIND_DRIVING: *k = {DrivingSim::A, DrivingSim::OW, DrivingSim::SW}; return true;
    case KIND_PREDATOR_PREY: *k = {PPSim::A, PPSim::OW, PPSim::SW}; return true;
    case KIND_LBF: *k = {LBFSim::A, LBFSim::OW, LBFSim::SW}; return true;
    case KIND_NESTED | KIND_DRIVING: *k = {DrivingSim::A, DrivingSim::OW, DrivingSim::SW + 1}; return true;
    case KIND_NESTED | KIND_PREDATOR_PREY: *k = {PPSim::A, PPSim::OW, PPSim::SW + 1}; return true;
    default: return false;
  }
}

static int make_tree(const posg_model* m, void* arena, long long cap_nodes, long long cap_log, search::Tree* T,
                     KindInfo* k) {
  if (!kind_info(m, k)) { set_error("model kind has no search support"); return POSG_E_UNSUPPORTED; }
  if (!arena || cap_nodes < 2 || cap_log < 1 || cap_nodes > 0x7fffffffLL / (k->n_actions + 1)) {
    set_error("bad tree arena arguments");
    return POSG_E_INVALID;
  }
  search::layout(T, (unsigned char*)arena, cap_nodes, cap_log, k->n_actions, k->obs_words, k->state_words);
  return POSG_OK;
}

}  // namespace posg

using namespace posg;

extern "C" {

int posg_model_info(posg_model_t m, int* n_actions, int* obs_words, int* state_words) {
  KindInfo k;
  if (!kind_info(m, &k)) { set_error("model kind has no search support"); return POSG_E_UNSUPPORTED; }
  if (n_actions) *n_actions = k.n_actions;
  if (obs_words) *obs_words = k.obs_words;
  if (state_words) *state_words = k.state_words;
  return POSG_OK;
}

int64_t posg_tree_arena_bytes(posg_model_t m, int64_t cap_nodes, int64_t cap_log) {
  KindInfo k;
  if (!kind_info(m, &k)) { set_error("model kind has no search support"); return POSG_E_UNSUPPORTED; }
  return (int64_t)search::layout(nullptr, nullptr, cap_nodes, cap_log, k.n_actions, k.obs_words, k.state_words);
}

int posg_tree_reset(posg_model_t m, void* arena, int64_t cap_nodes, int64_t cap_log, int root_t, double known_min,
                    double known_max, void* stream) {
  search::Tree T; KindInfo k;
  int rc = make_tree(m, arena, cap_nodes, cap_log, &T, &k); if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  // zero the node / action-node arrays (the particle log needs no clearing)
  const size_t upto = (size_t)((unsigned char*)T.log_node - (unsigned char*)arena);
  POSG_CUDA(cudaMemsetAsync(arena, 0, upto, s));
  tree_init_kernel<<<sm_count() * 4, 256, 0, s>>>(T, k.n_actions, root_t, known_min, known_max);
  POSG_CUDA(cudaGetLastError());
  return POSG_OK;
}

int posg_tree_reroot(posg_model_t m, void* src_arena, void* dst_arena, int64_t cap_nodes, int64_t cap_log,
                     int new_root, int32_t* scratch, int64_t* result_host, void* stream) {
  search::Tree S, D; KindInfo k;
  int rc = make_tree(m, src_arena, cap_nodes, cap_log, &S, &k); if (rc) return rc;
  rc = make_tree(m, dst_arena, cap_nodes, cap_log, &D, &k); if (rc) return rc;
  POSG_REQUIRE(src_arena != dst_arena && scratch && result_host && new_root >= 0 && new_root < cap_nodes, "bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  // source sizes and the destination's high-water mark (its counter block still holds the node count of
  // the last time it was the live arena: posg_tree_reset or an earlier re-root)
  search::Counters c[2];
  POSG_CUDA(cudaMemcpyAsync(&c[0], S.ctr, sizeof(search::Counters), cudaMemcpyDeviceToHost, s));
  POSG_CUDA(cudaMemcpyAsync(&c[1], D.ctr, sizeof(search::Counters), cudaMemcpyDeviceToHost, s));
  POSG_CUDA(cudaStreamSynchronize(s));
  const long long n_old = (long long)(c[0].node_count < (unsigned long long)cap_nodes ? c[0].node_count : (unsigned long long)cap_nodes);
  const unsigned long long n_log = c[0].log_count < (unsigned long long)cap_log ? c[0].log_count : (unsigned long long)cap_log;
  // (a nearly full arena may also have used its last slot as a transient root: posg_tree_advance)
  const long long dirty = (c[1].node_count + search::RESERVE >= (unsigned long long)cap_nodes) ? (long long)cap_nodes
                                                                                               : (long long)c[1].node_count;
  int* mark = scratch;
  int* order = scratch + cap_nodes;
  unsigned int* counters = nullptr;   // [0] nodes kept, then the 64-bit log counter
  POSG_CUDA(cudaMallocAsync((void**)&counters, 32, s));
  unsigned long long* d_log = (unsigned long long*)(counters + 4);
  POSG_CUDA(cudaMemsetAsync(counters, 0, 32, s));
  POSG_CUDA(cudaMemsetAsync(mark, 0xFF, (size_t)n_old * 4, s));
  {
    static int bfs_grid[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (!bfs_grid[dev]) {
      int per_sm = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, reroot_bfs_kernel, 256, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
      bfs_grid[dev] = sm_count() * (per_sm > 4 ? 4 : per_sm);
    }
    int na = k.n_actions;
    void* args[] = {(void*)&S, (void*)&na, (void*)&new_root, (void*)&mark, (void*)&order, (void*)&counters};
    // a small tree gets a small grid: the level barriers are cheaper, and the re-roots of a forest's trees (issued
    // from several streams) fit on the GPU side by side instead of each waiting for all of it
    long long want = (n_old + 2047) / 2048;
    if (want < 8) want = 8;
    const int grid = (int)(want < (long long)bfs_grid[dev] ? want : (long long)bfs_grid[dev]);
    if (n_old <= (1LL << 19)) {
      reroot_bfs_cta_kernel<<<1, 1024, 0, s>>>(S, na, new_root, mark, order, counters);
      POSG_CUDA(cudaGetLastError());
    } else {
      POSG_CUDA(cudaLaunchCooperativeKernel((const void*)reroot_bfs_kernel, dim3(grid), dim3(256), args, 0, s));
    }
  }
  reroot_copy_kernel<<<sm_count() * 8, 256, 0, s>>>(S, D, k.n_actions, k.obs_words, mark, order, counters);
  POSG_CUDA(cudaGetLastError());
  if (dirty > 0) {
    reroot_clear_kernel<<<sm_count() * 8, 256, 0, s>>>(D, k.n_actions, k.obs_words, counters, dirty);
    POSG_CUDA(cudaGetLastError());
  }
  if (n_log > 0) {
    reroot_log_kernel<<<sm_count() * 8, 256, 0, s>>>(S, D, k.state_words, mark, n_log, d_log);
    POSG_CUDA(cudaGetLastError());
  }
  reroot_finish_kernel<<<1, 1, 0, s>>>(S, D, counters, d_log);
  POSG_CUDA(cudaGetLastError());
  unsigned int h_cnt[4]; unsigned long long h_log = 0;
  POSG_CUDA(cudaMemcpyAsync(h_cnt, counters, 16, cudaMemcpyDeviceToHost, s));
  POSG_CUDA(cudaMemcpyAsync(&h_log, d_log, 8, cudaMemcpyDeviceToHost, s));
  POSG_CUDA(cudaStreamSynchronize(s));
  cudaFreeAsync(counters, s);
  result_host[0] = h_cnt[0];
  result_host[1] = (int64_t)h_log;
  return POSG_OK;
}

}  // extern "C" (helpers and templates below have C++ linkage)

static int fill_params(const posg_model* m, const posg_search_params* p, search::Params* P) {
  POSG_REQUIRE(p, "null search params");
  POSG_REQUIRE(p->ego_agent >= 0 && p->ego_agent < m->num_agents, "ego_agent out of range");
  POSG_REQUIRE(p->selection >= 0 && p->selection <= 2, "selection must be 0 (ucb), 1 (pucb) or 2 (uniform)");
  P->discount = p->discount; P->c = p->c; P->known_min = p->known_min; P->known_max = p->known_max;
  P->pucb_exploration_fraction = p->pucb_exploration_fraction;
  P->depth_limit = p->depth_limit; P->step_limit = p->step_limit; P->selection = p->selection; P->ego = p->ego_agent;
  P->intmcp_depth_rule = p->intmcp_depth_rule;
  P->seed = p->seed;
  for (int i = 0; i < search::MAX_DEPTH; ++i) P->disc_pow[i] = pow(p->discount, (double)i);
  return POSG_OK;
}

template <class M>
static void launch_search(const posg_model* m, const search::Tree& T, const search::Params& P,
                          const typename M::Ctx& ctx, int root, int root_t, const uint32_t* particles,
                          long long n_particles, const search::RootSelect& sel, long long n_sims,
                          unsigned long long sim_id0, cudaStream_t s, int n_waves = 1) {
  if (n_waves > 1 && n_sims <= 1024) {
    // consecutive small waves inside one CTA
    const int threads = (int)((n_sims + 31) / 32) * 32;
    smem_optin(search_waves_kernel<M>, m->blob_bytes);
    search_waves_kernel<M><<<1, threads, m->blob_bytes, s>>>(
        (const typename M::Blob*)m->d_blob, (uint32_t)m->blob_bytes, T, P, ctx, root, root_t, particles, n_particles,
        sel, (int)n_sims, n_waves, sim_id0);
    return;
  }
  const int grid = (int)((n_sims + SEARCH_THREADS - 1) / SEARCH_THREADS);
  smem_optin(search_kernel<M>, m->blob_bytes);
  for (int w = 0; w < n_waves; ++w)   // stream order = wave order
    search_kernel<M><<<grid, SEARCH_THREADS, m->blob_bytes, s>>>(
        (const typename M::Blob*)m->d_blob, (uint32_t)m->blob_bytes, T, P, ctx, root, root_t, particles, n_particles,
        sel, n_sims, sim_id0 + (unsigned long long)w * (unsigned long long)n_sims);
}

template <class M>
static void launch_forest(const posg_model* m, const ForestEntry* entries, int n_trees, long long cap_nodes,
                          long long cap_log, const search::Params& P, int root_t, int wave_size, int n_waves,
                          unsigned long long sim_id0, cudaStream_t s) {
  const int threads = ((wave_size + 31) / 32) * 32;
  smem_optin(search_forest_kernel<M>, m->blob_bytes);
  search_forest_kernel<M><<<n_trees, threads, m->blob_bytes, s>>>(
      (const typename M::Blob*)m->d_blob, (uint32_t)m->blob_bytes, entries, cap_nodes, cap_log, P, NoCtx{}, root_t,
      wave_size, n_waves, sim_id0);
}

template <class M>
static void warm_model_kernels() {
  cudaFuncAttributes a;
  cudaFuncGetAttributes(&a, search_kernel<M>);
  cudaFuncGetAttributes(&a, search_waves_kernel<M>);
  cudaFuncGetAttributes(&a, search_forest_kernel<M>);
  cudaFuncGetAttributes(&a, reinvigorate_kernel<M>);
  cudaFuncGetAttributes(&a, initial_belief_kernel<M>);
}

static int search_common(posg_model_t m, void* arena, int64_t cap_nodes, int64_t cap_log, int root, int root_t,
                         const uint32_t* particles, int64_t n_particles, const search::RootSelect& sel, int64_t n_sims,
                         uint64_t sim_id0, const posg_search_params* p, void* stream, int n_waves = 1) {
  search::Tree T; KindInfo k;
  int rc = make_tree(m, arena, cap_nodes, cap_log, &T, &k); if (rc) return rc;
  POSG_REQUIRE(particles && n_particles > 0 && n_sims >= 0 && root >= 0 && root < cap_nodes, "bad search arguments");
  search::Params P;
  rc = fill_params(m, p, &P); if (rc) return rc;
  if (n_sims == 0) return POSG_OK;
  cudaStream_t s = (cudaStream_t)stream;
  switch (m->kind) {
    case KIND_DRIVING:
      launch_search<DrivingSim>(m, T, P, NoCtx{}, root, root_t, particles, n_particles, sel, n_sims, sim_id0, s, n_waves); break;
    case KIND_PREDATOR_PREY:   // the table-driven adapter where its tables are exact (pp.cuh)
      if (((const pp::Blob*)m->h_blob)->lut_ok)
        launch_search<PPSimLut>(m, T, P, NoCtx{}, root, root_t, particles, n_particles, sel, n_sims, sim_id0, s, n_waves);
      else
        launch_search<PPSim>(m, T, P, NoCtx{}, root, root_t, particles, n_particles, sel, n_sims, sim_id0, s, n_waves);
      break;
    case KIND_LBF:
      launch_search<LBFSim>(m, T, P, NoCtx{}, root, root_t, particles, n_particles, sel, n_sims, sim_id0, s, n_waves); break;
    default: set_error("unsupported model kind"); return POSG_E_UNSUPPORTED;
  }
  POSG_CUDA(cudaGetLastError());
  return POSG_OK;
}

// context of a nested (level-1) model: the other agent's level-0 tree
template <class Base>
static int nested_ctx(posg_model_t nested, void* arena0, int64_t cap_nodes0, int64_t cap_log0, int other_agent,
                      typename NestedSim<Base>::Ctx* ctx) {
  posg_model base = *nested;
  base.kind = nested->kind & ~KIND_NESTED;
  KindInfo k0;
  int rc = make_tree(&base, arena0, cap_nodes0, cap_log0, &ctx->other_tree, &k0); if (rc) return rc;
  POSG_REQUIRE(other_agent >= 0 && other_agent < nested->num_agents, "other_agent out of range");
  ctx->other = other_agent;
  return POSG_OK;
}

extern "C" {

int posg_search(posg_model_t m, void* arena, int64_t cap_nodes, int64_t cap_log, int root, int root_t,
                const uint32_t* particles, int64_t n_particles, int64_t n_sims, uint64_t sim_id0,
                const posg_search_params* p, void* stream) {
  POSG_REQUIRE(m, "null model");
  search::RootSelect sel = {nullptr, nullptr, nullptr, 0};
  return search_common(m, arena, cap_nodes, cap_log, root, root_t, particles, n_particles, sel, n_sims, sim_id0, p,
                       stream);
}

int posg_search_waves(posg_model_t m, void* arena, int64_t cap_nodes, int64_t cap_log, int root, int root_t,
                      const uint32_t* particles, int64_t n_particles, int64_t wave_size, int n_waves, uint64_t sim_id0,
                      const posg_search_params* p, void* stream) {
  POSG_REQUIRE(n_waves >= 0 && n_waves <= (1 << 20) && wave_size >= 0, "bad wave arguments");
  if (n_waves == 0) return POSG_OK;
  search::RootSelect sel = {nullptr, nullptr, nullptr, 0};
  return search_common(m, arena, cap_nodes, cap_log, root, root_t, particles, n_particles, sel, wave_size, sim_id0, p,
                       stream, n_waves);
}

int posg_search_forest(posg_model_t m, const posg_forest_entry* entries_dev, int n_trees, int64_t cap_nodes,
                       int64_t cap_log, int root_t, int wave_size, int n_waves, uint64_t sim_id0,
                       const posg_search_params* p, void* stream) {
  KindInfo k;
  if (!kind_info(m, &k)) { set_error("model kind has no search support"); return POSG_E_UNSUPPORTED; }
  POSG_REQUIRE(entries_dev && n_trees >= 1 && n_trees <= 4096 && cap_nodes >= 2 && cap_log >= 1 &&
               cap_nodes <= 0x7fffffffLL / (k.n_actions + 1), "bad forest arguments");
  POSG_REQUIRE(wave_size >= 1 && wave_size <= 1024 && n_waves >= 0 && n_waves <= (1 << 20), "bad wave arguments");
  if (n_waves == 0) return POSG_OK;
  search::Params P;
  int rc = fill_params(m, p, &P); if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const ForestEntry* en = (const ForestEntry*)entries_dev;
  switch (m->kind) {
    case KIND_DRIVING: launch_forest<DrivingSim>(m, en, n_trees, cap_nodes, cap_log, P, root_t, wave_size, n_waves, sim_id0, s); break;
    case KIND_PREDATOR_PREY:
      if (((const pp::Blob*)m->h_blob)->lut_ok)
        launch_forest<PPSimLut>(m, en, n_trees, cap_nodes, cap_log, P, root_t, wave_size, n_waves, sim_id0, s);
      else
        launch_forest<PPSim>(m, en, n_trees, cap_nodes, cap_log, P, root_t, wave_size, n_waves, sim_id0, s);
      break;
    case KIND_LBF: launch_forest<LBFSim>(m, en, n_trees, cap_nodes, cap_log, P, root_t, wave_size, n_waves, sim_id0, s); break;
    default: set_error("unsupported model kind"); return POSG_E_UNSUPPORTED;
  }
  POSG_CUDA(cudaGetLastError());
  return POSG_OK;
}

int posg_forest_stats(posg_model_t m, const posg_forest_entry* entries_dev, int n_trees, int64_t cap_nodes,
                      int64_t cap_log, double* out_host, void* stream) {
  KindInfo k;
  if (!kind_info(m, &k)) { set_error("model kind has no search support"); return POSG_E_UNSUPPORTED; }
  POSG_REQUIRE(entries_dev && out_host && n_trees >= 1 && n_trees <= 4096 && cap_nodes >= 2 && cap_log >= 1, "bad forest arguments");
  cudaStream_t s = (cudaStream_t)stream;
  const size_t bytes = (size_t)n_trees * (3 * k.n_actions + 10) * sizeof(double);
  double* d_out = nullptr;
  POSG_CUDA(cudaMallocAsync((void**)&d_out, bytes, s));
  forest_stats_kernel<<<n_trees, 32, 0, s>>>((const ForestEntry*)entries_dev, cap_nodes, cap_log, k.n_actions, k.obs_words,
                                             k.state_words, d_out);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(out_host, d_out, bytes, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  cudaFreeAsync(d_out, s);
  POSG_CUDA(e);
  return POSG_OK;
}

int posg_search_multi(posg_model_t m, void* arena, int64_t cap_nodes, int64_t cap_log, int root_t,
                      const uint32_t* particles, int64_t n_particles, const int32_t* sel_root,
                      const int64_t* sel_off, const int32_t* sel_cnt, int64_t n_sel, int64_t n_sims,
                      uint64_t sim_id0, const posg_search_params* p, void* stream) {
  POSG_REQUIRE(m && sel_root && sel_off && sel_cnt && n_sel > 0, "bad selection arguments");
  search::RootSelect sel = {sel_root, (const long long*)sel_off, sel_cnt, (long long)n_sel};
  return search_common(m, arena, cap_nodes, cap_log, 0, root_t, particles, n_particles, sel, n_sims, sim_id0, p,
                       stream, (p && p->n_waves > 1) ? p->n_waves : 1);
}

int posg_search_nested(posg_model_t m, void* arena, int64_t cap_nodes, int64_t cap_log, void* other_arena,
                       int64_t other_cap_nodes, int64_t other_cap_log, int other_agent, int root, int root_t,
                       const uint32_t* particles, int64_t n_particles, int64_t n_sims, uint64_t sim_id0,
                       const posg_search_params* p, void* stream) {
  POSG_REQUIRE(m && (m->kind & KIND_NESTED), "posg_search_nested needs a handle from posg_nested_create");
  search::Tree T; KindInfo k;
  int rc = make_tree(m, arena, cap_nodes, cap_log, &T, &k); if (rc) return rc;
  POSG_REQUIRE(particles && n_particles > 0 && n_sims >= 0 && root >= 0 && root < cap_nodes, "bad search arguments");
  search::Params P;
  rc = fill_params(m, p, &P); if (rc) return rc;
  POSG_REQUIRE(other_agent != p->ego_agent, "other_agent must differ from the ego agent");
  if (n_sims == 0) return POSG_OK;
  cudaStream_t s = (cudaStream_t)stream;
  search::RootSelect sel = {nullptr, nullptr, nullptr, 0};
  switch (m->kind & ~KIND_NESTED) {
    case KIND_DRIVING: {
      NestedSim<DrivingSim>::Ctx ctx;
      rc = nested_ctx<DrivingSim>(m, other_arena, other_cap_nodes, other_cap_log, other_agent, &ctx); if (rc) return rc;
      launch_search<NestedSim<DrivingSim>>(m, T, P, ctx, root, root_t, particles, n_particles, sel, n_sims, sim_id0, s,
                                           p->n_waves > 1 ? p->n_waves : 1);
    } break;
    case KIND_PREDATOR_PREY: {
      NestedSim<PPSim>::Ctx ctx;
      rc = nested_ctx<PPSim>(m, other_arena, other_cap_nodes, other_cap_log, other_agent, &ctx); if (rc) return rc;
      launch_search<NestedSim<PPSim>>(m, T, P, ctx, root, root_t, particles, n_particles, sel, n_sims, sim_id0, s,
                                      p->n_waves > 1 ? p->n_waves : 1);
    } break;
    default: set_error("unsupported model kind"); return POSG_E_UNSUPPORTED;
  }
  POSG_CUDA(cudaGetLastError());
  return POSG_OK;
}

int posg_nested_create(posg_model_t base, posg_model_t* out) {
  POSG_REQUIRE(base && out && !(base->kind & KIND_NESTED), "bad arguments");
  KindInfo k;
  if (!kind_info(base, &k)) { set_error("model kind has no search support"); return POSG_E_UNSUPPORTED; }
  posg_model* m = new posg_model(*base);
  m->kind = base->kind | KIND_NESTED;   // same blob (borrowed), one extra state word per particle
  m->borrowed = 1;
  m->d_flat_desc = nullptr; m->flat_dim = 0;
  // the copy must not share (and later free) the base handle's lazily created staging contexts
  m->host_stage = nullptr; m->host_stage_free = nullptr;
  m->step_stage = nullptr; m->step_stage_free = nullptr;
  *out = m;
  return POSG_OK;
}

int posg_tree_advance_batch(posg_model_t m, void* arena, int64_t cap_nodes, int64_t cap_log, int64_t n,
                            const int32_t* parents, const int32_t* actions, const uint64_t* obs,
                            int32_t* children_out, void* stream) {
  search::Tree T; KindInfo k;
  int rc = make_tree(m, arena, cap_nodes, cap_log, &T, &k); if (rc) return rc;
  if (n == 0) return POSG_OK;
  POSG_REQUIRE(n > 0 && parents && actions && obs && children_out, "bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = (int)((n + 127) / 128);
  if (k.n_actions == 5 && k.obs_words == 1)
    tree_advance_batch_kernel<5, 1><<<grid, 128, 0, s>>>(T, n, parents, actions, (const unsigned long long*)obs, children_out);
  else { set_error("unsupported (actions, obs words) combination"); return POSG_E_UNSUPPORTED; }
  POSG_CUDA(cudaGetLastError());
  return POSG_OK;
}

int posg_tree_node_info(posg_model_t m, void* arena, int64_t cap_nodes, int64_t cap_log, int64_t n,
                        const int32_t* nodes, int32_t* parent_out, int32_t* action_out, uint64_t* obs_out,
                        int32_t* t_out, void* stream) {
  search::Tree T; KindInfo k;
  int rc = make_tree(m, arena, cap_nodes, cap_log, &T, &k); if (rc) return rc;
  if (n == 0) return POSG_OK;
  POSG_REQUIRE(n > 0 && nodes && parent_out && action_out && obs_out && t_out, "bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = (int)((n + 127) / 128);
  if (k.obs_words == 1)
    tree_node_info_kernel<1><<<grid, 128, 0, s>>>(T, k.n_actions, n, nodes, parent_out, action_out,
                                                  (unsigned long long*)obs_out, t_out);
  else
    tree_node_info_kernel<2><<<grid, 128, 0, s>>>(T, k.n_actions, n, nodes, parent_out, action_out,
                                                  (unsigned long long*)obs_out, t_out);
  POSG_CUDA(cudaGetLastError());
  return POSG_OK;
}

int posg_tree_node_stats(posg_model_t m, void* arena, int64_t cap_nodes, int64_t cap_log, int node,
                         double* out_host, void* stream) {
  search::Tree T; KindInfo k;
  int rc = make_tree(m, arena, cap_nodes, cap_log, &T, &k); if (rc) return rc;
  POSG_REQUIRE(out_host && node >= 0 && node < cap_nodes, "bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  const int A = k.n_actions;
  int vis[16], done[16], nv = 0; double tot[16]; search::Counters c; unsigned char fl = 0;
  POSG_CUDA(cudaMemcpyAsync(vis, T.an_visits + (size_t)node * A, A * 4, cudaMemcpyDeviceToHost, s));
  POSG_CUDA(cudaMemcpyAsync(done, T.an_done + (size_t)node * A, A * 4, cudaMemcpyDeviceToHost, s));
  POSG_CUDA(cudaMemcpyAsync(tot, T.an_total + (size_t)node * A, A * 8, cudaMemcpyDeviceToHost, s));
  POSG_CUDA(cudaMemcpyAsync(&nv, T.node_visits + node, 4, cudaMemcpyDeviceToHost, s));
  POSG_CUDA(cudaMemcpyAsync(&fl, T.node_flags + node, 1, cudaMemcpyDeviceToHost, s));
  POSG_CUDA(cudaMemcpyAsync(&c, T.ctr, sizeof(c), cudaMemcpyDeviceToHost, s));
  POSG_CUDA(cudaStreamSynchronize(s));
  for (int a = 0; a < A; ++a) { out_host[a] = vis[a]; out_host[A + a] = done[a]; out_host[2 * A + a] = tot[a]; }
  double* e = out_host + 3 * A;
  auto unkey = [](unsigned long long kk) {
    unsigned long long b = (kk >> 63) ? (kk & 0x7fffffffffffffffULL) : ~kk;
    double d; memcpy(&d, &b, 8); return d;
  };
  e[0] = nv; e[1] = (double)c.node_count; e[2] = (double)c.log_count; e[3] = c.max_depth;
  e[4] = unkey(c.min_key); e[5] = unkey(c.max_key); e[6] = c.overflow; e[7] = (double)c.sims_done; e[8] = fl;
  e[9] = (double)c.nodes_created;
  return POSG_OK;
}

int posg_tree_root_stats_device(posg_model_t m, void* arena, int64_t cap_nodes, int64_t cap_log, int node,
                                double* out_device, void* stream) {
  search::Tree T; KindInfo k;
  int rc = make_tree(m, arena, cap_nodes, cap_log, &T, &k); if (rc) return rc;
  POSG_REQUIRE(out_device && node >= 0 && node < cap_nodes, "bad arguments");
  root_stats_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(T, k.n_actions, node, out_device);
  POSG_CUDA(cudaGetLastError());
  return POSG_OK;
}

int posg_tree_advance(posg_model_t m, void* arena, int64_t cap_nodes, int64_t cap_log, int root, int action,
                      const uint64_t* obs_host, int create_if_missing, int* out_host, void* stream) {
  search::Tree T; KindInfo k;
  int rc = make_tree(m, arena, cap_nodes, cap_log, &T, &k); if (rc) return rc;
  POSG_REQUIRE(obs_host && out_host && root >= 0 && root < cap_nodes && action >= 0 && action < k.n_actions, "bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  unsigned long long* d_obs = nullptr; int* d_out = nullptr;
  POSG_CUDA(cudaMallocAsync((void**)&d_obs, 8 * k.obs_words + 16, s));
  d_out = (int*)(d_obs + k.obs_words);
  POSG_CUDA(cudaMemcpyAsync(d_obs, obs_host, 8 * k.obs_words, cudaMemcpyHostToDevice, s));
  if (k.obs_words == 1) tree_advance_kernel<1><<<1, 1, 0, s>>>(T, k.n_actions, root, action, d_obs, create_if_missing, d_out);
  else tree_advance_kernel<2><<<1, 1, 0, s>>>(T, k.n_actions, root, action, d_obs, create_if_missing, d_out);
  POSG_CUDA(cudaGetLastError());
  POSG_CUDA(cudaMemcpyAsync(out_host, d_out, 16, cudaMemcpyDeviceToHost, s));
  POSG_CUDA(cudaStreamSynchronize(s));
  cudaFreeAsync(d_obs, s);
  return POSG_OK;
}

int posg_tree_gather_belief(posg_model_t m, void* arena, int64_t cap_nodes, int64_t cap_log, int node,
                            uint32_t* states_out, int64_t max_out, int64_t* count_host, void* stream) {
  search::Tree T; KindInfo k;
  int rc = make_tree(m, arena, cap_nodes, cap_log, &T, &k); if (rc) return rc;
  POSG_REQUIRE(states_out && count_host && max_out >= 0, "bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  unsigned long long* d_cnt = nullptr;
  POSG_CUDA(cudaMallocAsync((void**)&d_cnt, 8, s));
  POSG_CUDA(cudaMemsetAsync(d_cnt, 0, 8, s));
  gather_belief_kernel<<<sm_count() * 8, 256, 0, s>>>(T, k.state_words, node, states_out, (long long)max_out, d_cnt);
  POSG_CUDA(cudaGetLastError());
  unsigned long long cnt = 0;
  POSG_CUDA(cudaMemcpyAsync(&cnt, d_cnt, 8, cudaMemcpyDeviceToHost, s));
  POSG_CUDA(cudaStreamSynchronize(s));
  cudaFreeAsync(d_cnt, s);
  *count_host = (int64_t)cnt;
  return POSG_OK;
}

static int belief_common(posg_model_t m, int obs_words, const uint64_t* obs_host, cudaStream_t s,
                         unsigned long long** d_obs, long long** d_res) {
  POSG_CUDA(cudaMallocAsync((void**)d_obs, 8 * obs_words + 32, s));
  *d_res = (long long*)(*d_obs + obs_words);
  POSG_CUDA(cudaMemcpyAsync(*d_obs, obs_host, 8 * obs_words, cudaMemcpyHostToDevice, s));
  (void)m;
  return POSG_OK;
}

static int reinvigorate_common(posg_model_t m, void* other_arena, int64_t other_cap_nodes, int64_t other_cap_log,
                               int other_agent, const uint32_t* parent_particles, int64_t n_parent, int ego_agent,
                               int action, const uint64_t* target_obs_host, int64_t n_needed, int64_t max_attempts,
                               uint64_t seed, uint64_t draw_id, uint32_t* out_states, uint32_t* reject_scratch,
                               int64_t* result_host, void* stream) {
  KindInfo k;
  if (!kind_info(m, &k)) { set_error("model kind has no belief support"); return POSG_E_UNSUPPORTED; }
  POSG_REQUIRE(parent_particles && n_parent > 0 && target_obs_host && out_states && reject_scratch && result_host,
               "bad arguments");
  POSG_REQUIRE(ego_agent >= 0 && ego_agent < m->num_agents && action >= 0 && action < k.n_actions, "bad agent/action");
  result_host[0] = result_host[1] = result_host[2] = 0;
  if (n_needed <= 0) return POSG_OK;
  cudaStream_t s = (cudaStream_t)stream;
  unsigned long long* d_obs; long long* d_res;
  int rc = belief_common(m, k.obs_words, target_obs_host, s, &d_obs, &d_res); if (rc) return rc;
#define REINV(MODEL, CTX)                                                                                   \
  smem_optin(reinvigorate_kernel<MODEL>, m->blob_bytes);                                                    \
  reinvigorate_kernel<MODEL><<<1, belief::CTA, m->blob_bytes, s>>>(                                         \
      (const MODEL::Blob*)m->d_blob, (uint32_t)m->blob_bytes, CTX, parent_particles, (long long)n_parent,   \
      ego_agent, action, d_obs, (long long)n_needed, (long long)max_attempts, seed, draw_id, out_states,    \
      reject_scratch, d_res)
  switch (m->kind) {
    case KIND_DRIVING: REINV(DrivingSim, NoCtx{}); break;
    case KIND_PREDATOR_PREY: REINV(PPSim, NoCtx{}); break;
    case KIND_LBF: REINV(LBFSim, NoCtx{}); break;
    case KIND_NESTED | KIND_DRIVING: {
      NestedSim<DrivingSim>::Ctx ctx;
      rc = nested_ctx<DrivingSim>(m, other_arena, other_cap_nodes, other_cap_log, other_agent, &ctx); if (rc) return rc;
      REINV(NestedSim<DrivingSim>, ctx);
    } break;
    case KIND_NESTED | KIND_PREDATOR_PREY: {
      NestedSim<PPSim>::Ctx ctx;
      rc = nested_ctx<PPSim>(m, other_arena, other_cap_nodes, other_cap_log, other_agent, &ctx); if (rc) return rc;
      REINV(NestedSim<PPSim>, ctx);
    } break;
    default: set_error("unsupported model kind"); return POSG_E_UNSUPPORTED;
  }
#undef REINV
  POSG_CUDA(cudaGetLastError());
  POSG_CUDA(cudaMemcpyAsync(result_host, d_res, 24, cudaMemcpyDeviceToHost, s));
  POSG_CUDA(cudaStreamSynchronize(s));
  cudaFreeAsync(d_obs, s);
  return POSG_OK;
}

int posg_belief_reinvigorate(posg_model_t m, const uint32_t* parent_particles, int64_t n_parent, int ego_agent,
                             int action, const uint64_t* target_obs_host, int64_t n_needed, int64_t max_attempts,
                             uint64_t seed, uint64_t draw_id, uint32_t* out_states, uint32_t* reject_scratch,
                             int64_t* result_host, void* stream) {
  POSG_REQUIRE(m && !(m->kind & KIND_NESTED), "use posg_belief_reinvigorate_nested for nested handles");
  return reinvigorate_common(m, nullptr, 0, 0, 0, parent_particles, n_parent, ego_agent, action, target_obs_host,
                             n_needed, max_attempts, seed, draw_id, out_states, reject_scratch, result_host, stream);
}

int posg_belief_reinvigorate_nested(posg_model_t m, void* other_arena, int64_t other_cap_nodes,
                                    int64_t other_cap_log, int other_agent, const uint32_t* parent_particles,
                                    int64_t n_parent, int ego_agent, int action, const uint64_t* target_obs_host,
                                    int64_t n_needed, int64_t max_attempts, uint64_t seed, uint64_t draw_id,
                                    uint32_t* out_states, uint32_t* reject_scratch, int64_t* result_host,
                                    void* stream) {
  POSG_REQUIRE(m && (m->kind & KIND_NESTED), "needs a handle from posg_nested_create");
  return reinvigorate_common(m, other_arena, other_cap_nodes, other_cap_log, other_agent, parent_particles, n_parent,
                             ego_agent, action, target_obs_host, n_needed, max_attempts, seed, draw_id, out_states,
                             reject_scratch, result_host, stream);
}

int posg_belief_initial(posg_model_t m, int ego_agent, const uint64_t* target_obs_host, int64_t n_needed,
                        int64_t max_attempts, uint64_t seed, uint64_t draw_id, uint32_t* out_states,
                        int64_t* result_host, void* stream) {
  KindInfo k;
  if (!kind_info(m, &k)) { set_error("model kind has no belief support"); return POSG_E_UNSUPPORTED; }
  POSG_REQUIRE(target_obs_host && out_states && result_host && ego_agent >= 0 && ego_agent < m->num_agents,
               "bad arguments");
  result_host[0] = result_host[1] = result_host[2] = 0;
  if (n_needed <= 0) return POSG_OK;
  cudaStream_t s = (cudaStream_t)stream;
  unsigned long long* d_obs; long long* d_res;
  int rc = belief_common(m, k.obs_words, target_obs_host, s, &d_obs, &d_res); if (rc) return rc;
  switch (m->kind) {
    case KIND_DRIVING:
      smem_optin(initial_belief_kernel<DrivingSim>, m->blob_bytes);
      initial_belief_kernel<DrivingSim><<<1, belief::CTA, m->blob_bytes, s>>>(
          (const DrivingSim::Blob*)m->d_blob, (uint32_t)m->blob_bytes, ego_agent, d_obs, (long long)n_needed,
          (long long)max_attempts, seed, draw_id, out_states, d_res);
      break;
    case KIND_PREDATOR_PREY:
      initial_belief_kernel<PPSim><<<1, belief::CTA, m->blob_bytes, s>>>(
          (const PPSim::Blob*)m->d_blob, (uint32_t)m->blob_bytes, ego_agent, d_obs, (long long)n_needed,
          (long long)max_attempts, seed, draw_id, out_states, d_res);
      break;
    case KIND_LBF:
      initial_belief_kernel<LBFSim><<<1, belief::CTA, m->blob_bytes, s>>>(
          (const LBFSim::Blob*)m->d_blob, (uint32_t)m->blob_bytes, ego_agent, d_obs, (long long)n_needed,
          (long long)max_attempts, seed, draw_id, out_states, d_res);
      break;
    default: set_error("unsupported model kind"); return POSG_E_UNSUPPORTED;
  }
  POSG_CUDA(cudaGetLastError());
  POSG_CUDA(cudaMemcpyAsync(result_host, d_res, 24, cudaMemcpyDeviceToHost, s));
  POSG_CUDA(cudaStreamSynchronize(s));
  cudaFreeAsync(d_obs, s);
  return POSG_OK;
}

// MCTS._update (mcts.py:229-263) for one tree in ONE call: advance to the (action, observation) child, collect the
// particles the searches left there, top the belief up by rejection sampling from the old root belief, re-root.
// The steps are the entry points above (each ends with the read-back its successor's arguments need); doing them
// behind one call lets a host thread per stream update the trees of a forest without taking the interpreter lock
// between the steps.
int posg_tree_update(posg_model_t m, void* arena, void* spare_arena, int64_t cap_nodes, int64_t cap_log, int root,
                     int action, const uint64_t* obs_host, int ego_agent, const uint32_t* parent_particles,
                     int64_t n_parent, uint32_t* belief_out, int64_t belief_capacity, int64_t log_bound, int64_t target,
                     double sample_limit_factor, uint64_t seed, uint64_t draw_id, uint32_t* reject_scratch,
                     int32_t* reroot_scratch, int64_t* result_host, void* stream) {
  POSG_REQUIRE(result_host && belief_out && target >= 0 && log_bound >= 0 && belief_capacity >= log_bound + target,
               "bad arguments");
  for (int i = 0; i < 8; ++i) result_host[i] = 0;
  int out4[4] = {0, 0, 0, 0};
  auto now_us = []() { return std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  long long t0 = now_us(), t_adv = 0, t_gat = 0, t_rei = 0;
  int rc = posg_tree_advance(m, arena, cap_nodes, cap_log, root, action, obs_host, 1, out4, stream); if (rc) return rc;
  t_adv = now_us() - t0; t0 = now_us();
  const int child = out4[0], flags = out4[2];
  result_host[0] = child; result_host[1] = flags;
  if (child < 0) return POSG_OK;   // arena full: the caller reports it
  int64_t have = 0;
  if (!(flags & 2)) {
    rc = posg_tree_gather_belief(m, arena, cap_nodes, cap_log, child, belief_out, log_bound, &have, stream); if (rc) return rc;
    if (have > log_bound) have = log_bound;
    t_gat = now_us() - t0; t0 = now_us();
    const int64_t need = target - have;
    if (need > 0) {
      const int64_t limit = (int64_t)(sample_limit_factor * (double)need);
      int64_t res[3] = {0, 0, 0};
      KindInfo k;
      if (!kind_info(m, &k)) { set_error("model kind has no search support"); return POSG_E_UNSUPPORTED; }
      rc = posg_belief_reinvigorate(m, parent_particles, n_parent, ego_agent, action, obs_host, need, limit, seed, draw_id,
                                    belief_out + have * k.state_words, reject_scratch, res, stream);
      if (rc) return rc;
      have += res[2];
      result_host[5] = limit;     // draw counters consumed
      result_host[6] = 1;         // reinvigoration ran
    }
  }
  result_host[2] = have;
  t_rei = now_us() - t0; t0 = now_us();
  int64_t rr[2] = {0, 0};
  rc = posg_tree_reroot(m, arena, spare_arena, cap_nodes, cap_log, child, reroot_scratch, rr, stream); if (rc) return rc;
  result_host[3] = rr[0]; result_host[4] = rr[1];
  // diagnostics: microseconds spent in advance / gather / reinvigorate / re-root, 16 bits each (saturating)
  auto sat = [](long long v) { return (unsigned long long)(v < 0 ? 0 : v > 65535 * 16 ? 65535 : v / 16); };
  result_host[7] = (int64_t)(sat(t_adv) | (sat(t_gat) << 16) | (sat(t_rei) << 32) | (sat(now_us() - t0) << 48));
  return POSG_OK;
}

// Loads every kernel of the planner's search / update path for this model NOW (the CUDA runtime loads a kernel
// lazily at its first launch, under a process-wide lock: a forest that updates its trees from several host threads
// otherwise meets those loads in the middle of a decision, all threads at once -- stalls of 0.1-2 s were measured).
int posg_search_warmup(posg_model_t m) {
  KindInfo k;
  if (!kind_info(m, &k)) { set_error("model kind has no search support"); return POSG_E_UNSUPPORTED; }
  cudaFuncAttributes a;
  cudaFuncGetAttributes(&a, forest_stats_kernel);
  cudaFuncGetAttributes(&a, tree_init_kernel);
  cudaFuncGetAttributes(&a, reroot_bfs_kernel);
  cudaFuncGetAttributes(&a, reroot_bfs_cta_kernel);
  cudaFuncGetAttributes(&a, reroot_copy_kernel);
  cudaFuncGetAttributes(&a, reroot_clear_kernel);
  cudaFuncGetAttributes(&a, reroot_log_kernel);
  cudaFuncGetAttributes(&a, reroot_finish_kernel);
  cudaFuncGetAttributes(&a, root_stats_kernel);
  cudaFuncGetAttributes(&a, gather_belief_kernel);
  if (k.obs_words == 1) cudaFuncGetAttributes(&a, tree_advance_kernel<1>);
  else cudaFuncGetAttributes(&a, tree_advance_kernel<2>);
  switch (m->kind) {
    case KIND_DRIVING: warm_model_kernels<DrivingSim>(); break;
    case KIND_PREDATOR_PREY:
      if (((const pp::Blob*)m->h_blob)->lut_ok) warm_model_kernels<PPSimLut>(); else warm_model_kernels<PPSim>();
      break;
    case KIND_LBF: warm_model_kernels<LBFSim>(); break;
    default: break;
  }
  cudaGetLastError();
  return POSG_OK;
}

}  // extern "C"
