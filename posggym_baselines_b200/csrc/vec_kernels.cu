// Vector-env entry points of the C-ABI (PPO rollout collector path).
#include <vector>

#include "common.h"
#include "driving_sim.cuh"
#include "pp.cuh"
#include "lbf.cuh"
#include "vec_env.cuh"

namespace posg {

// gymnasium.spaces.flatten order of each model's per-agent observation (one-hot blocks)
static std::vector<vec::FlatDesc> build_desc(const posg_model* m) {
  std::vector<vec::FlatDesc> d;
  auto block = [&](int word, int shift, int bits, int width) {
    for (int v = 0; v < width; ++v)
      d.push_back({(uint8_t)word, (uint8_t)shift, (uint8_t)((1u << bits) - 1u), (uint8_t)v});
  };
  if (m->kind == KIND_DRIVING) {
    const driving::Blob* b = (const driving::Blob*)m->h_blob;
    for (int c = 0; c < 15; ++c) block(0, 2 * c, 2, 4);  // local cells
    block(0, 30, 2, 4);                                   // speed
    block(0, 32, 4, b->width); block(0, 36, 4, b->height);   // current coord
    block(0, 40, 4, b->width); block(0, 44, 4, b->height);   // destination coord
    block(0, 48, 1, 2); block(0, 49, 1, 2);               // dest_reached, crashed
  } else if (m->kind == KIND_PREDATOR_PREY) {
    for (int c = 0; c < 25; ++c) block(0, 2 * c, 2, 4);
  } else if (m->kind == KIND_LBF) {
    const lbf::Blob* b = (const lbf::Blob*)m->h_blob;
    const int pw = b->size + 1, fw = b->size + 2;    // Discrete(size+1, start=-1), Discrete(size+2, start=-1)
    const int plv = b->max_agent_level + 1, flv = 2 * b->max_agent_level + 1;
    // players: self, then the other agent (fields hold value + 1)
    block(0, 0, 3, pw); block(0, 3, 3, pw); block(0, 6, 2, plv);
    block(0, 8, 3, pw); block(0, 11, 3, pw); block(0, 14, 2, plv);
    // food slots: 0 in w0 low half @16; 1..3 in w0 high half; 4..6 in w1 low half; 7 in w1 high half
    for (int slot = 0; slot < 8; ++slot) {
      const int word = slot < 4 ? 0 : 1;
      const int base = slot == 0 ? 16 : ((slot >= 1 && slot <= 3) ? 32 + 9 * (slot - 1)
                                   : (slot <= 6 ? 9 * (slot - 4) : 32));
      block(word, base, 3, fw); block(word, base + 3, 3, fw); block(word, base + 6, 3, flv);
    }
  }
  return d;
}

static int ensure_desc(posg_model* m) {
  if (m->d_flat_desc) return POSG_OK;
  std::vector<vec::FlatDesc> d = build_desc(m);
  if (d.empty() || d.size() > (size_t)vec::MAX_FLAT) { set_error("model kind has no vector-env support"); return POSG_E_UNSUPPORTED; }
  POSG_CUDA(cudaMalloc(&m->d_flat_desc, d.size() * sizeof(vec::FlatDesc)));
  POSG_CUDA(cudaMemcpy(m->d_flat_desc, d.data(), d.size() * sizeof(vec::FlatDesc), cudaMemcpyHostToDevice));
  m->flat_dim = (int)d.size();
  return POSG_OK;
}

template <class M>
static int launch_vec(posg_model* m, int64_t n, uint32_t* state, uint16_t* ep_t, double* ep_ret, uint32_t* ep_count,
                      const long long* actions, uint64_t seed, uint32_t global_step, uint64_t env_offset,
                      int max_steps, int reset_only, float* obs_out, float* rew_out, uint8_t* term_out,
                      uint8_t* trunc_out, uint8_t* done_out, float* stat_r, int* stat_l, cudaStream_t s) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, vec::vec_step_kernel<M>, vec::THREADS, m->blob_bytes) !=
          cudaSuccess || per_sm < 1) { cudaGetLastError(); per_sm = 1; }
  const int64_t need = (n + vec::THREADS - 1) / vec::THREADS;
  const int64_t cap = (int64_t)sm_count() * per_sm;
  const int grid = (int)(need < cap ? need : cap);
  vec::vec_step_kernel<M><<<grid, vec::THREADS, m->blob_bytes, s>>>(
      (const typename M::Blob*)m->d_blob, (uint32_t)m->blob_bytes, n, state, ep_t, ep_ret, ep_count, actions,
      (const vec::FlatDesc*)m->d_flat_desc, m->flat_dim, seed, global_step, env_offset, max_steps, reset_only,
      obs_out, rew_out, term_out, trunc_out, done_out, stat_r, stat_l);
  POSG_CUDA(cudaGetLastError());
  return POSG_OK;
}

}  // namespace posg

using namespace posg;

extern "C" {

int posg_vec_flat_dim(posg_model_t m) {
  POSG_REQUIRE(m, "null model");
  int rc = ensure_desc(m); if (rc) return rc;
  return m->flat_dim;
}

int posg_vec_step(posg_model_t m, int64_t n_envs, uint32_t* state, uint16_t* ep_t, double* ep_ret,
                  uint32_t* ep_count, const int64_t* actions, uint64_t seed, uint32_t global_step,
                  uint64_t env_offset, int max_episode_steps, int reset_only, float* obs_out, float* rew_out,
                  uint8_t* term_out, uint8_t* trunc_out, uint8_t* done_out, float* stat_r, int32_t* stat_l,
                  void* stream) {
  POSG_REQUIRE(m, "null model");
  int rc = ensure_desc(m); if (rc) return rc;
  if (n_envs == 0) return POSG_OK;
  POSG_REQUIRE(n_envs > 0 && state && ep_t && ep_ret && ep_count && obs_out, "bad arguments");
  POSG_REQUIRE(reset_only || (actions && rew_out && term_out && trunc_out && done_out && stat_r && stat_l),
               "step needs actions and every output buffer");
  POSG_REQUIRE(max_episode_steps > 0 && max_episode_steps < 65535, "bad max_episode_steps");
  cudaStream_t s = (cudaStream_t)stream;
  switch (m->kind) {
    case KIND_DRIVING:
      return launch_vec<DrivingSim>(m, n_envs, state, ep_t, ep_ret, ep_count, (const long long*)actions, seed,
                                    global_step, env_offset, max_episode_steps, reset_only, obs_out, rew_out, term_out,
                                    trunc_out, done_out, stat_r, stat_l, s);
    case KIND_LBF:
      return launch_vec<LBFSim>(m, n_envs, state, ep_t, ep_ret, ep_count, (const long long*)actions, seed, global_step,
                                env_offset, max_episode_steps, reset_only, obs_out, rew_out, term_out, trunc_out,
                                done_out, stat_r, stat_l, s);
    case KIND_PREDATOR_PREY:
      return launch_vec<PPSim>(m, n_envs, state, ep_t, ep_ret, ep_count, (const long long*)actions, seed, global_step,
                               env_offset, max_episode_steps, reset_only, obs_out, rew_out, term_out, trunc_out,
                               done_out, stat_r, stat_l, s);
    default: set_error("model kind has no vector-env support"); return POSG_E_UNSUPPORTED;
  }
}

}  // extern "C"
