/*
 * posg_b200.h -- C-ABI of libposg_b200.so: batched posggym grid-world models on B200.
 *
 * The reference (RDLLab/posggym-baselines) has no FFI: its hot path calls the Python
 * methods of a posggym.model.POSGModel one state at a time.  Each entry point below is
 * the batched replacement of one of those call sites (cited per function, paths
 * relative to /root/reference).  Conventions:
 *
 *   - plain pointers and sizes only; every data pointer is a DEVICE pointer unless the
 *     function name ends in _host (then every data pointer is a HOST pointer and the
 *     call does its own H2D/D2H copies and synchronises before returning);
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream);
 *   - return value: 0 = ok, negative = POSG_E_* below; no exceptions cross the ABI;
 *     posg_last_error() returns a thread-local message for the last failure;
 *   - no ownership transfer: the caller (PyTorch) owns all buffers; the library owns
 *     only the immutable per-model map blob created by *_create / freed by *_destroy;
 *   - thread-compatible: handles are immutable after creation.
 *
 * Packed layouts (SoA, one element per env / particle; see DESIGN.md "Data layout"):
 *
 *   Driving state   : uint32 per vehicle, vehicles of one env adjacent
 *                     (n_agents == 2 -> one uint64 per env, vehicle 0 in the low word)
 *       bits  0-3 x | 4-7 y | 8-9 facing (N,E,S,W) | 10-11 speed (REV,STOP,SLOW,FAST)
 *       12-14 destination index | 15 dest_reached | 16 crashed
 *       17-23 min_dest_dist | 24-29 init_dest_dist (1..63) | 30-31 zero (inputs with these
 *       bits set are invalid; the kernels use them internally for the step's progress)
 *   Driving obs     : uint64 per agent, agents of one env adjacent
 *       bits 0-29 fifteen 2-bit cells (row-major 5x3 window, VEHICLE=0 WALL=1 EMPTY=2
 *       DESTINATION=3) | 30-31 speed | 32-35 x | 36-39 y | 40-43 dest x | 44-47 dest y
 *       48 dest_reached | 49 crashed
 *   rewards         : double per agent, agents of one env adjacent
 *   flags           : uint8 per env: bit a = terminated[a] (a < 4); bits 4-6 reserved (0:
 *                     truncation belongs to the TimeLimit wrapper, not the model);
 *                     bit 7 = all_done (POSG_FLAG_ALL_DONE)
 *   actions         : uint8 per agent, agents of one env adjacent
 */
#ifndef POSG_B200_H
#define POSG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define POSG_OK 0
#define POSG_E_INVALID (-1)   /* bad argument */
#define POSG_E_UNSUPPORTED (-2)
#define POSG_E_CUDA (-3)      /* a CUDA runtime call failed */
#define POSG_E_NOGPU (-4)     /* no CUDA device: there is no CPU fallback */

#define POSG_FLAG_ALL_DONE 0x80u

typedef struct posg_model* posg_model_t;
typedef struct posg_policy* posg_policy_t;

const char* posg_last_error(void);
int posg_version(void);
/* number of CUDA devices visible, or POSG_E_NOGPU */
int posg_device_count(void);

/* ------------------------------------------------------------------ Driving-v1 */

/* posggym DrivingModel.__init__ (called through posggym.make("Driving-v1", ...),
 * README.md:52, tests/planning/test_pomcp.py:52-58).  `grid_rows` is the
 * newline-separated ASCII layout ('.', '#', '+', '-', digits, 'a'-'j').  Builds the
 * padded wall bitmaps, destination/start tables and all shortest-path distances on
 * the host and uploads one immutable blob to the current device. */
int posg_driving_create(const char* grid_rows, int num_agents, int obs_front, int obs_back,
                        int obs_side, int obstacle_collisions, posg_model_t* out);
int posg_model_destroy(posg_model_t m);
/* bytes of the device blob (what each CTA stages into shared memory) */
int posg_model_blob_bytes(posg_model_t m);
int posg_model_num_agents(posg_model_t m);

/* model.sample_initial_state() (planning/mcts.py:191, planning/intmcp.py:237) for n
 * envs.  With picks == NULL the picks come from Philox4x32-10, key = seed, counter =
 * (particle_offset + i, stream 0xB0); otherwise picks[i*2*n_agents + k] are the
 * explicit noise words (start_pick, dest_pick per agent, see oracle/driving.py). */
int posg_driving_sample_initial_state(posg_model_t m, int64_t n, const uint8_t* picks,
                                      uint64_t seed, uint64_t particle_offset,
                                      uint32_t* state_out, void* stream);

/* model.sample_initial_obs(state) (planning/mcts.py:192, planning/intmcp.py:238). */
int posg_driving_observe(posg_model_t m, int64_t n, const uint32_t* state, uint64_t* obs_out,
                         void* stream);

/* model.step(state, joint_action) (planning/mcts.py:333,418; planning/belief.py:165;
 * planning/intmcp.py:468,562) for n envs.
 *   actions_in : n*n_agents uint8, or NULL -> uniform actions from Philox (counter =
 *                (particle_offset + i, step_idx, stream 0xA0)), written to actions_out
 *                when that is non-NULL;
 *   order_in   : n uint8 explicit processing-order noise (permutation index), or NULL
 *                -> from the same Philox block;
 *   live_count : optional device uint64 incremented by the number of envs whose input
 *                state was not all-done (the "executed" env-steps of a sweep).
 * state_out may alias state_in. */
int posg_driving_step(posg_model_t m, int64_t n, const uint32_t* state_in,
                      const uint8_t* actions_in, const uint8_t* order_in, uint64_t seed,
                      uint32_t step_idx, uint64_t particle_offset, uint32_t* state_out,
                      uint64_t* obs_out, double* rewards_out, uint8_t* flags_out,
                      uint8_t* actions_out, unsigned long long* live_count, void* stream);

/* MCTS._rollout (planning/mcts.py:405-452) with RandomSearchPolicy
 * (planning/search_policy.py:176-177) for the ego agent and RandomOtherAgentPolicy
 * (planning/other_policy.py:150-151) for the others, fused over `max_steps` model
 * steps: returns_out[i] = sum_k discount^k * r_ego,k, stopping after the step on which
 * terminated[ego] || all_done.  steps_out[i] (optional) = model steps executed,
 * state_out (optional) = final state.  Randomness: Philox counter (particle_offset + i,
 * step0 + k, stream 0xA0) -- identical to calling posg_driving_step k times with
 * step_idx = step0 + k.  max_steps is min(depth_limit - depth + 1, step_limit - t + 1)
 * evaluated by the caller (mcts.py:414). */
int posg_driving_rollout(posg_model_t m, int64_t n, const uint32_t* state_in, int ego_agent,
                         int max_steps, double discount, uint64_t seed, uint32_t step0,
                         uint64_t particle_offset, double* returns_out, uint8_t* steps_out,
                         uint32_t* state_out, unsigned long long* live_count, void* stream);

/* Host-buffer variants (the e2e path of bench.py): copy in, run, copy out, synchronise. */
int posg_driving_rollout_host(posg_model_t m, int64_t n, const uint32_t* state_in_host,
                              int ego_agent, int max_steps, double discount, uint64_t seed,
                              uint32_t step0, uint64_t particle_offset, double* returns_out_host,
                              uint8_t* steps_out_host, uint32_t* state_out_host,
                              unsigned long long* live_count_host);
/* depth `num_steps` sweep through posg_driving_step (every step writes state, obs,
 * rewards, flags, actions to device scratch owned by the call); copies the LAST step's
 * outputs back. */
int posg_driving_sweep_host(posg_model_t m, int64_t n, const uint32_t* state_in_host,
                            int num_steps, uint64_t seed, uint64_t particle_offset,
                            uint32_t* state_out_host, uint64_t* obs_out_host,
                            double* rewards_out_host, uint8_t* flags_out_host,
                            unsigned long long* live_count_host);


/* The same sweep with the rewards delivered as ONE byte per env instead of two float64 (every step of the sweep
 * writes the code, the last step's is copied back): per vehicle, progress made in the step (2 bits) | reward kind
 * (2 bits: plain, destination reached, crashed, already done), vehicle 1 in the high nibble.  Together with
 * init_dest_dist in the returned state word that is the index of the model's reward table;
 * posg_driving_decode_rewards_host expands it on the host to exactly the doubles posg_driving_sweep_host returns
 * (26 instead of 41 bytes per env cross PCIe).  n must be a multiple of 128 and particle_offset a multiple of 4. */
int posg_driving_sweep_host_compact(posg_model_t m, int64_t n, const uint32_t* state_in_host, int num_steps,
                                    uint64_t seed, uint64_t particle_offset, uint32_t* state_out_host,
                                    uint64_t* obs_out_host, uint8_t* reward_code_out_host, uint8_t* flags_out_host,
                                    unsigned long long* live_count_host);
int posg_driving_decode_rewards_host(posg_model_t m, int64_t n, const uint32_t* state_host,
                                     const uint8_t* reward_code_host, double* rewards_out_host);

/* ------------------------------------------------------------------ PredatorPrey-v0 and
 * the model-generic batch entry points (every model kind except Driving-v1, which has
 * the hand-tuned posg_driving_* kernels above).
 *
 *   PredatorPrey state : 2 x uint32 -- word 0: predator 0 (x | y<<4) | predator 1 << 8 |
 *                        caught bits << 16; word 1: prey 0 | prey 1 << 8 | prey 2 << 16
 *   PredatorPrey obs   : uint64 per agent: 25 cells x 2 bits (EMPTY=0 WALL=1 PREDATOR=2
 *                        PREY=3), cell index col*5 + row of the window centred on the agent
 *   explicit noise     : PredatorPrey: 3 x uint32 per env (one word per prey: the single
 *                        uniform choice that prey makes in the step) */

/* posggym PredatorPreyModel.__init__ (posggym.make("PredatorPrey-v0", ...),
 * baseline_exps/env_data/PredatorPrey-v0/env_kwargs.yaml).  block_rows: optional
 * newline-separated '#'/'.' rows (NULL = no blocks, the "NxN" grids). */
int posg_pp_create(int grid_size, const char* block_rows, int num_predators, int num_prey,
                   int cooperative, int prey_strength, int obs_dim, posg_model_t* out);

/* posggym LevelBasedForagingModel.__init__ (posggym.make("LevelBasedForaging-v3", ...),
 * baseline_exps/env_data/LevelBasedForaging-v3/env_kwargs.yaml): static layout given as
 * food cells and player start cells (x | y << 4).  State: 2 x uint32 (players + total food
 * level spawned; eight 3-bit food levels); observation: 2 x uint64 per agent (the 30-entry
 * tuple observation, 3-bit fields, see csrc/lbf.cuh); no step noise (one dummy word). */
int posg_lbf_create(int size, int num_agents, int max_agent_level, int max_food, int sight,
                    int force_coop, int normalize_reward, const uint8_t* food_xy,
                    const uint8_t* player_xy, posg_model_t* out);

/* model.sample_initial_state (mcts.py:191).  picks: explicit noise (uint8 per pick). */
int posg_model_sample_initial_state(posg_model_t m, int64_t n, const uint8_t* picks, uint64_t seed,
                                    uint64_t particle_offset, uint32_t* state_out, void* stream);
/* model.sample_initial_obs / observation of a state (mcts.py:192). */
int posg_model_observe(posg_model_t m, int64_t n, const uint32_t* state, uint64_t* obs_out,
                       void* stream);
/* model.step (mcts.py:333,418; belief.py:165); NULL actions_in / noise_in -> Philox block
 * (particle_offset + i, step_idx, stream 0xA0). */
int posg_model_step(posg_model_t m, int64_t n, const uint32_t* state_in, const uint8_t* actions_in,
                    const uint32_t* noise_in, uint64_t seed, uint32_t step_idx,
                    uint64_t particle_offset, uint32_t* state_out, uint64_t* obs_out,
                    double* rewards_out, uint8_t* flags_out, uint8_t* actions_out,
                    unsigned long long* live_count, void* stream);
/* model.step for a handful of rows in HOST memory, any model kind (Driving included): the scalar call of a
 * planner whose policies are Python plugins -- planning/mcts.py:333,418 and planning/intmcp.py:468,562 reached
 * through planning/host_search.py / host_intmcp.py, one call per tree level of a simulation.  Explicit actions
 * [n][A] and explicit noise [n][noise_row_bytes] (Driving: 1 byte, the processing-order index; other models:
 * 4 bytes per noise word).  One staged copy in, the model's step kernel, one copy out, one synchronisation on a
 * stream the handle owns; one call per handle at a time. */
int posg_model_step_host(posg_model_t m, int64_t n, const uint32_t* state_in_host,
                         const uint8_t* actions_host, const void* noise_host, int noise_row_bytes,
                         uint32_t* state_out_host, uint64_t* obs_out_host, double* rewards_out_host,
                         uint8_t* flags_out_host);
/* MCTS._rollout with uniform-random policies (mcts.py:405-452), see posg_driving_rollout. */
int posg_model_rollout(posg_model_t m, int64_t n, const uint32_t* state_in, int ego_agent,
                       int max_steps, double discount, uint64_t seed, uint32_t step0,
                       uint64_t particle_offset, double* returns_out, uint8_t* steps_out,
                       uint32_t* state_out, unsigned long long* live_count, void* stream);

/* ------------------------------------------------------------------ sweep checksum
 * Per-particle 64-bit rolling hash over the outputs of a model.step sweep ("every
 * step's (state, obs, r, term, trunc, all_done) compared bit-exact ... by 64-bit rolling hash for larger
 * N").  One call folds ONE step's outputs of n particles into hash_inout[i]:
 *     h <- init ? (particle_offset + i + 1) * 0xD6E8FEB86659FD93 : hash_inout[i]
 *     for w in [state words (each zero-extended), obs words (agent-major), reward bit patterns,
 *               flags | a0 << 8 | a1 << 16 ...]:   h <- (h ^ w) * 0x9E3779B97F4A7C15;  h <- h ^ (h >> 32)
 * state == NULL: only the initialisation is stored.  The CPU oracles (oracle/c/oracle_common.h) compute the
 * same value from their own restatement of the model; equal hashes certify the whole trace. */
int posg_trace_hash(int64_t n, uint64_t* hash_inout, int init, uint64_t particle_offset, const uint32_t* state,
                    int state_words, const uint64_t* obs, int obs_words, const double* rewards, int num_agents,
                    const uint8_t* flags, const uint8_t* actions, void* stream);

/* ------------------------------------------------------------------ POMCP search tree
 * The tree (obs nodes, per-action statistics, child lists, particle log) lives in one
 * caller-owned device arena; see posggym_baselines_b200/csrc/search.cuh for the layout.
 * These entry points replace the reference's MCTS hot loops for RandomSearchPolicy /
 * RandomOtherAgentPolicy and state_belief_only=True (the POMCP configuration,
 * planning/pomcp.py:20-35). */

typedef struct posg_search_params {
  double discount;                  /* MCTSConfig.discount */
  double c;                         /* MCTSConfig.c */
  double known_min, known_max;      /* MCTSConfig.known_bounds; min > max = none */
  double pucb_exploration_fraction; /* MCTSConfig.pucb_exploration_fraction */
  int depth_limit;                  /* MCTSConfig.depth_limit (derived, config.py:50-55) */
  int step_limit;                   /* MCTS.step_limit (mcts.py:53-58) */
  int selection;                    /* 0 ucb, 1 pucb, 2 uniform (mcts.py:66-74) */
  int ego_agent;
  int intmcp_depth_rule;            /* 1: INTMCP._simulate's stop rule node.t + depth > step_limit (intmcp.py:450-453) */
  uint64_t seed;
  int n_waves;                      /* posg_search_multi / posg_search_nested: > 1 = that many consecutive waves of n_sims
                                       simulations each in one call (see posg_search_waves); 0 or 1 = one wave */
  int pad_;
} posg_search_params;

/* ego action count, 64-bit words per observation key, 32-bit words per packed state */
int posg_model_info(posg_model_t m, int* n_actions, int* obs_words, int* state_words);
int64_t posg_tree_arena_bytes(posg_model_t m, int64_t cap_nodes, int64_t cap_log);
/* MCTS.reset (mcts.py:123-138): empty tree whose root is node 0 at time root_t. */
int posg_tree_reset(posg_model_t m, void* arena, int64_t cap_nodes, int64_t cap_log, int root_t,
                    double known_min, double known_max, void* stream);
/* n_sims simulations of MCTS.get_action's loop body (mcts.py:285-290 -> _simulate
 * 308-382 -> _evaluate/_rollout 384-452) from the belief `particles` (n_particles packed
 * states) of node `root`; simulation i uses Philox counter sim_id0 + i. */
int posg_search(posg_model_t m, void* arena, int64_t cap_nodes, int64_t cap_log, int root, int root_t,
                const uint32_t* particles, int64_t n_particles, int64_t n_sims, uint64_t sim_id0,
                const posg_search_params* p, void* stream);
/* n_waves consecutive waves of wave_size simulations each (wave w uses Philox counters sim_id0 + w * wave_size ..),
 * every wave seeing all backups of the one before -- the same as n_waves posg_search calls, in one launch for
 * waves of up to 1024 simulations.  Simulations of ONE wave see each other's visit counts but not each other's
 * returns, so equivalence with the reference's sequential loop (mcts.py:285-290) needs waves that are small
 * beside the root's visit count (posggym_baselines_b200/planning/mcts.py: wave_growth). */
int posg_search_waves(posg_model_t m, void* arena, int64_t cap_nodes, int64_t cap_log, int root, int root_t,
                      const uint32_t* particles, int64_t n_particles, int64_t wave_size, int n_waves, uint64_t sim_id0,
                      const posg_search_params* p, void* stream);
/* Loads the kernels of the search / update path for this model now instead of at their first launch (a lazily
 * loaded kernel is loaded under a process-wide lock; a forest updating from several host threads should not meet
 * that in the middle of a decision). */
int posg_search_warmup(posg_model_t m);
/* MCTS._update (mcts.py:229-263) for one tree in one call: posg_tree_advance (creating the child if the searches
 * never reached it), posg_tree_gather_belief into belief_out (capacity >= log_bound + target rows; log_bound = an
 * upper bound of the tree's particle-log length), posg_belief_reinvigorate from parent_particles for the particles
 * still missing to `target` (at most sample_limit_factor * missing attempts, Philox draws from draw_id), then
 * posg_tree_reroot into spare_arena.  result_host[8]: child's old node id (-1: arena full, nothing else done),
 * its flags (bit 1 = absorbing: no belief is built), particles now in belief_out, nodes kept, log entries kept,
 * draw counters consumed, 1 if reinvigoration ran, host microseconds / 16 spent in the four steps (16 bits each:
 * advance, gather, reinvigorate, re-root; diagnostics). */
int posg_tree_update(posg_model_t m, void* arena, void* spare_arena, int64_t cap_nodes, int64_t cap_log, int root,
                     int action, const uint64_t* obs_host, int ego_agent, const uint32_t* parent_particles,
                     int64_t n_parent, uint32_t* belief_out, int64_t belief_capacity, int64_t log_bound, int64_t target,
                     double sample_limit_factor, uint64_t seed, uint64_t draw_id, uint32_t* reject_scratch,
                     int32_t* reroot_scratch, int64_t* result_host, void* stream);
/* Root parallelisation inside one GPU: n_trees independent searches in ONE launch, CTA k on the tree
 * entries_dev[k] (a DEVICE array; all trees share cap_nodes / cap_log / root_t and the Philox key, tree k uses the
 * counters sim_id0 + (k << 32) + w * wave_size + thread; wave_size = the largest wave of the launch).  Each tree runs posg_search_waves' loop -- the reference's MCTS.get_action
 * loop body (mcts.py:285-290) on its own tree; the caller sums the roots' visits / total values afterwards, the
 * same merge as root-parallel POMCP across GPUs. */
typedef struct posg_forest_entry {
  void* arena;               /* the tree's arena (posg_tree_arena_bytes(cap_nodes, cap_log)) */
  const uint32_t* particles; /* its root belief: n_particles packed states */
  int64_t n_particles;
  int32_t root;              /* its root node */
  int32_t wave;              /* simulations per wave of this tree, <= wave_size (0: wave_size): a tree whose root has
                                more visits may run larger waves than its neighbours */
} posg_forest_entry;
int posg_search_forest(posg_model_t m, const posg_forest_entry* entries_dev, int n_trees, int64_t cap_nodes,
                       int64_t cap_log, int root_t, int wave_size, int n_waves, uint64_t sim_id0,
                       const posg_search_params* p, void* stream);
/* posg_tree_node_stats of every tree's root in one launch and one copy: out_host[n_trees][3 * A + 10]. */
int posg_forest_stats(posg_model_t m, const posg_forest_entry* entries_dev, int n_trees, int64_t cap_nodes,
                      int64_t cap_log, double* out_host, void* stream);
/* Copies to HOST: visits[A], completed[A], total_value[A] of the node's action children
 * (ActionNode.visits / total_value, node.py:138-140) then node visits, node count, log
 * count, max search depth, min value, max value, overflow count, sims done, node flags, obs nodes created
 * since the last reset / re-root (3 * A + 10 doubles). */
int posg_tree_node_stats(posg_model_t m, void* arena, int64_t cap_nodes, int64_t cap_log, int node,
                         double* out_host, void* stream);
/* The same completed[A], total_value[A] as doubles written to DEVICE memory (2 * A doubles): the buffer
 * of the root-parallel allreduce (ActionNode.visits / total_value summed over the ranks' trees, one
 * collective per decision), so the merge needs no host round trip before the collective. */
int posg_tree_root_stats_device(posg_model_t m, void* arena, int64_t cap_nodes, int64_t cap_log, int node,
                                double* out_device, void* stream);
/* MCTS._update re-rooting (mcts.py:229-263): child of (root, action) with observation
 * obs (HOST words); out_host = {node or -1, created, flags, t}. */
int posg_tree_advance(posg_model_t m, void* arena, int64_t cap_nodes, int64_t cap_log, int root,
                      int action, const uint64_t* obs_host, int create_if_missing, int* out_host,
                      void* stream);
/* MCTS._update (mcts.py:261-263): the subtree under `new_root` (statistics, child lists and
 * every logged particle of its nodes) is copied, renumbered breadth-first so that the new
 * root is node 0, into `dst_arena`; everything else is dropped.  `scratch`: 3 * cap_nodes
 * int32 (device).  result_host = {nodes kept, particles kept}. */
int posg_tree_reroot(posg_model_t m, void* src_arena, void* dst_arena, int64_t cap_nodes,
                     int64_t cap_log, int new_root, int32_t* scratch, int64_t* result_host,
                     void* stream);
/* obs_node.belief.particles: every particle logged at `node` -> states_out (device). */
int posg_tree_gather_belief(posg_model_t m, void* arena, int64_t cap_nodes, int64_t cap_log, int node,
                            uint32_t* states_out, int64_t max_out, int64_t* count_host, void* stream);
/* BeliefRejectionSampler._rejection_sample (planning/belief.py:145-194) with
 * use_rejected_samples=True: result_host = {accepted, attempts, written}. */
int posg_belief_reinvigorate(posg_model_t m, const uint32_t* parent_particles, int64_t n_parent,
                             int ego_agent, int action, const uint64_t* target_obs_host,
                             int64_t n_needed, int64_t max_attempts, uint64_t seed, uint64_t draw_id,
                             uint32_t* out_states, uint32_t* reject_scratch, int64_t* result_host,
                             void* stream);
/* MCTS._initial_update rejection loop (mcts.py:186-194): result_host as above. */
int posg_belief_initial(posg_model_t m, int ego_agent, const uint64_t* target_obs_host,
                        int64_t n_needed, int64_t max_attempts, uint64_t seed, uint64_t draw_id,
                        uint32_t* out_states, int64_t* result_host, void* stream);

/* ------------------------------------------------------------------ I-NTMCP (nesting level 1)
 * The ego agent's tree is searched with the other agent acting by its own level-0 tree
 * (planning/intmcp.py).  A nested handle has the base model's blob and one extra state word
 * per particle: the node of the other agent's tree its history leads to. */
int posg_nested_create(posg_model_t base, posg_model_t* out);
/* INTMCP._nested_sim for search_level == nesting_level (intmcp.py:437-442 -> _simulate
 * 444-517): other agent's actions from INTMCP.sample_action (763-791) on `other_arena`,
 * histories extended by INTMCP.traverse (797-809). */
int posg_search_nested(posg_model_t nested, void* arena, int64_t cap_nodes, int64_t cap_log,
                       void* other_arena, int64_t other_cap_nodes, int64_t other_cap_log,
                       int other_agent, int root, int root_t, const uint32_t* particles,
                       int64_t n_particles, int64_t n_sims, uint64_t sim_id0,
                       const posg_search_params* p, void* stream);
/* INTMCP._nested_sim for search_level < nesting_level (intmcp.py:433-436): simulation i draws
 * entry e of [0, n_sel) -- a particle of the parent tree's root belief -- and searches this
 * tree from node sel_root[e] with a particle of particles[sel_off[e] .. +sel_cnt[e]). */
int posg_search_multi(posg_model_t m, void* arena, int64_t cap_nodes, int64_t cap_log, int root_t,
                      const uint32_t* particles, int64_t n_particles, const int32_t* sel_root,
                      const int64_t* sel_off, const int32_t* sel_cnt, int64_t n_sel, int64_t n_sims,
                      uint64_t sim_id0, const posg_search_params* p, void* stream);
/* INTMCP.traverse, one step, for a batch (device arrays): child of (parent, action, obs). */
int posg_tree_advance_batch(posg_model_t m, void* arena, int64_t cap_nodes, int64_t cap_log,
                            int64_t n, const int32_t* parents, const int32_t* actions,
                            const uint64_t* obs, int32_t* children_out, void* stream);
/* (parent obs node, action, observation, t) of a batch of nodes = the last step of the
 * histories they stand for (hist.get_last_step(), intmcp.py:293). Device arrays. */
int posg_tree_node_info(posg_model_t m, void* arena, int64_t cap_nodes, int64_t cap_log, int64_t n,
                        const int32_t* nodes, int32_t* parent_out, int32_t* action_out,
                        uint64_t* obs_out, int32_t* t_out, void* stream);
/* INTMCP._reinvigorate with _reinvigorate_action_fn (intmcp.py:843-915): the other agent acts
 * by its tree; new particles carry the extended history node. */
int posg_belief_reinvigorate_nested(posg_model_t nested, void* other_arena, int64_t other_cap_nodes,
                                    int64_t other_cap_log, int other_agent,
                                    const uint32_t* parent_particles, int64_t n_parent, int ego_agent,
                                    int action, const uint64_t* target_obs_host, int64_t n_needed,
                                    int64_t max_attempts, uint64_t seed, uint64_t draw_id,
                                    uint32_t* out_states, uint32_t* reject_scratch,
                                    int64_t* result_host, void* stream);

/* ------------------------------------------------------------------ vector env (PPO collector)
 * One call = envs.step(actions) of StackEnv(RecordEpisodeStatistics(SyncVectorEnv(
 * [FlattenObservations(env)] * n))) as built by PPOConfig.load_vec_env
 * (ppo/config.py:223-232) and consumed by run_rollout_worker (ppo/worker.py:82-86,
 * 176-216): model.step, TimeLimit truncation, episode statistics, auto-reset (the reset
 * observation is returned), one-hot flattened float32 observations, rows env-major
 * (row = env * num_agents + agent).  All buffers are device memory owned by the caller:
 *   state [n, SW] u32, ep_t [n] u16, ep_ret [n, A] f64, ep_count [n] u32  (carried state)
 *   actions [n * A] i64 (worker.py:177)
 *   obs_out [n * A, flat_dim] f32, rew_out [n * A] f32, term_out / trunc_out [n * A] u8,
 *   done_out [n] u8, stat_r [n, A] f32 and stat_l [n] i32 (episode return / length, valid
 *   where done_out is set: infos[agent]["episode"]["r"/"l"], worker.py:208-216).
 *   Optional episode wall time (RecordEpisodeStatistics' "t"): `now` = the caller's clock in
 *   seconds, ep_t0 [n] f64 (carried: start time of the running episode), stat_t [n] f64
 *   (elapsed time of the episode that just ended, 0 elsewhere); pass NULL for both to skip.
 *   step_counter (device, 2 x u32, optional): when not NULL the step index of the Philox counter is
 *   step_counter[0] + 1 instead of global_step, and the launch stores it back when its last CTA
 *   finishes (step_counter[1] = scratch, must be 0) -- a CUDA graph replays the launch with its
 *   scalar arguments baked in, so a captured collector loop keeps its step index on the device.
 * Step noise: Philox counter (env_offset + i, global_step); reset_only != 0 = envs.reset(). */
int posg_vec_flat_dim(posg_model_t m);
int posg_vec_step(posg_model_t m, int64_t n_envs, uint32_t* state, uint16_t* ep_t, double* ep_ret,
                  uint32_t* ep_count, const int64_t* actions, uint64_t seed, uint32_t global_step,
                  uint64_t env_offset, int max_episode_steps, int reset_only, float* obs_out,
                  float* rew_out, uint8_t* term_out, uint8_t* trunc_out, uint8_t* done_out,
                  float* stat_r, int32_t* stat_l, double now, double* ep_t0, double* stat_t,
                  uint32_t* step_counter, void* stream);

/* First trunk layer of the PPO policy on the PACKED observation: what
 * PPOLSTMModel.trunk[0:2] = nn.Linear(input_size, trunk_sizes[0]) + nn.Tanh (ppo/network.py:171-178)
 * computes from the one-hot rows the worker feeds it (ppo/worker.py:143-163), without writing or
 * reading those rows: out[env * A + agent, :] = act(bias + sum over observation fields f of
 * wT[offset_f + value_f, :]).  state [n, SW] u32 (the vec env's carried state: the observation is
 * recomputed from it); wT [flat_dim, hidden] f32 row-major (= trunk.0.weight transposed); bias
 * [hidden]; hidden <= 128; out [n * A, hidden] f32.  Pass obs_out = NULL to posg_vec_step to skip
 * the one-hot rows altogether. */
int posg_vec_embed(posg_model_t m, int64_t n_envs, const uint32_t* state, const float* wT, const float* bias,
                   int hidden, int apply_tanh, float* out, void* stream);

/* envs.step FUSED with posg_vec_embed of the observations it returns (the collector's loop, worker.py:143-178:
 * the rows envs.step returns are only ever fed to trunk layer 0): posg_vec_step's step (no one-hot rows, no wall
 * clock) and, in the same launch, x_out [n * A, 64] = act(bias + sum of weight rows) for the NEXT policy step;
 * state_copy_out (optional) also receives the new packed state (the rollout storage's row).  hidden must be 64. */
int posg_vec_step_embed(posg_model_t m, int64_t n_envs, uint32_t* state, uint16_t* ep_t, double* ep_ret,
                        uint32_t* ep_count, const int64_t* actions, uint64_t seed, uint32_t global_step,
                        uint64_t env_offset, int max_episode_steps, float* rew_out, uint8_t* term_out,
                        uint8_t* trunc_out, uint8_t* done_out, float* stat_r, int32_t* stat_l, uint32_t* step_counter,
                        const float* wT, const float* bias, int hidden, int apply_tanh, float* x_out,
                        uint32_t* state_copy_out, void* stream);

/* The rest of the collector's policy step: PPOLSTMModel.get_action_and_value
 * (ppo/network.py:212-268,305-330) after trunk[0:2] for trunk_sizes [64, 64], lstm_size 64, one LSTM layer,
 * head_sizes [64] (baseline_exps/exp_utils.py:101-104), fused into one tcgen05 tensor-core kernel
 * (csrc/policy_kernels.cu).  posg_policy_create takes HOST pointers to the fp32 parameters in PyTorch's
 * layouts -- trunk.2.weight [64, 64] / bias, lstm.weight_ih_l0 [256, 64], lstm.weight_hh_l0 [256, 64],
 * lstm.bias_ih_l0 / bias_hh_l0 [256], actor.0.weight [64, 64] / bias, actor.2.weight [A, 64] / bias,
 * critic.0.weight [64, 64] / bias, critic.2.weight [1, 64] / bias -- and packs them (fp16, tensor-core operand
 * layout) on the current device.
 * posg_policy_step, all DEVICE pointers: x0 [n, 64] f32 = tanh(trunk.0(obs)) (posg_vec_embed's output),
 * h_in / c_in [n, 64] f32 (the LSTM state), done [n] f32 (1 = reset the state first: get_states' (1 - d) mask) or,
 * with done == NULL, terminated / truncated [n] u8 as posg_vec_step writes them (done = their OR, worker.py:191-195),
 * action_in [n] i64 or NULL (NULL: sample from Categorical(logits) with Philox counter (row_offset + row, step);
 * step_device != NULL: the step index is read from that device word -- posg_vec_step's step_counter -- so that a
 * collector loop captured in a CUDA graph draws fresh numbers at every replay);
 * outputs h_out / c_out [n, 64] f32 (may alias the inputs), action_out [n] i64, logprob_out / entropy_out /
 * value_out [n] f32, logits_out [n, A] f32 or NULL.  fp16 operands, fp32 accumulation and epilogues: results
 * agree with the fp32 PyTorch module within 5e-3 (tests/test_policy_gpu.py). */
int posg_policy_create(int hidden, int n_actions, const float* trunk1_w, const float* trunk1_b, const float* w_ih,
                       const float* w_hh, const float* b_ih, const float* b_hh, const float* actor0_w,
                       const float* actor0_b, const float* actor2_w, const float* actor2_b, const float* critic0_w,
                       const float* critic0_b, const float* critic2_w, const float* critic2_b, posg_policy_t* out);
int posg_policy_destroy(posg_policy_t p);
/* PPOLSTMSearchPolicy.get_next_state (planning/search_policy.py:254-276) -- the network evaluated for the node a
 * simulation adds -- as ONE call with HOST buffers: posg_policy_set_trunk0 uploads trunk.0.weight [64, in_dim] /
 * bias [64] (host pointers); posg_policy_eval_host takes, per row, the K indices of the active inputs of the
 * one-hot-flattened observation (FlattenProcessor's layout; out-of-range indices are ignored) and the LSTM state
 * h / c [n, 64], and returns the new h / c, the actor's logits [n, A], their softmax probs [n, A]
 * (Categorical(logits).probs) and the critic's value [n].  Internally: one
 * staged copy in, the index-row sum + tanh of trunk layer 0, posg_policy_step (done = 0, nothing sampled), a row
 * softmax, one copy
 * out, one synchronisation on a stream the handle owns.  One call per handle at a time. */
int posg_policy_set_trunk0(posg_policy_t p, int in_dim, const float* trunk0_w, const float* trunk0_b);
int posg_policy_eval_host(posg_policy_t p, int64_t n, int K, const int32_t* obs_index_host, const float* h_in_host,
                          const float* c_in_host, float* h_out_host, float* c_out_host, float* logits_out_host,
                          float* probs_out_host, float* value_out_host);
int posg_policy_step(posg_policy_t p, int64_t n_rows, const float* x0, const float* h_in, const float* c_in,
                     const float* done, const uint8_t* terminated, const uint8_t* truncated, uint64_t seed,
                     uint32_t step, const uint32_t* step_device, uint64_t row_offset, const int64_t* action_in,
                     float* h_out, float* c_out, int64_t* action_out, float* logprob_out, float* entropy_out,
                     float* value_out, float* logits_out, void* stream);

/* Generalised advantage estimation of the PPO learner (ppo/core.py:226-244): the
 * backward loop over seq_len steps of PPOTrainer.update in one launch.  rewards [T, n_cols] f32,
 * dones / values [T + 1, n_cols] f32 (the extra final row as the worker collects it), outputs
 * advantages / returns [T, n_cols] f32 (returns = advantages + values[:-1]).  Bit-identical to the
 * reference's float32 tensor expressions.  Advantage normalisation (norm_adv) stays with the caller. */
int posg_gae(int T, int64_t n_cols, const float* rewards, const float* dones, const float* values, double gamma,
             double gae_lambda, float* advantages, float* returns, void* stream);

/* Sequence chunks of a rollout batch for BPTT (ppo/utils.py:240-292 get_seq_idxs): dones [T, B] f32
 * (device; non-zero where the step starts a new episode).  A piece starts at step 0 and at every done step and is
 * cut every max_seq_len steps; chunk i = (chunk_env[i], chunk_start[i], chunk_len[i]) (device, capacity entries
 * each), env-major then by start.  result_host = {number of chunks, longest chunk}.  One launch. */
int posg_seq_chunks(int T, int64_t B, const float* dones, int max_seq_len, int32_t* chunk_env, int32_t* chunk_start,
                    int32_t* chunk_len, int64_t capacity, int64_t* result_host, void* stream);
/* split_and_pad_batch (ppo/utils.py:295-344) for up to 16 batch entries in one launch: entry k is src[k]
 * [T (+ 1 if overlapping[k]), B, row_bytes[k]] -> dst[k] [max_seq_len (+ 1), n_chunks, row_bytes[k]] (device
 * pointers in HOST arrays); rows past a chunk's end get pad_bits[k] (an elem_bytes[k]-wide pattern) up to the
 * longest chunk and zeros after it, as pad_sequence + F.pad leave them. */
int posg_seq_gather(int n_entries, const void* const* src, void* const* dst, const int64_t* row_bytes,
                    const int32_t* overlapping, const int32_t* elem_bytes, const uint64_t* pad_bits,
                    const int32_t* chunk_env, const int32_t* chunk_start, const int32_t* chunk_len, int64_t n_chunks,
                    int64_t B, int max_seq_len, int longest, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* POSG_B200_H */
