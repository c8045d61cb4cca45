// Vector-env entry points of the C-ABI (PPO rollout collector path).
#include <cstdlib>
#include <vector>

#include "common.h"
#include "driving_sim.cuh"
#include "pp.cuh"
#include "lbf.cuh"
#include "vec_env.cuh"

namespace posg {

// One-hot widths of the flattened-observation fields (gymnasium.spaces.flatten order of each
// model's per-agent observation tuple); field f's bits inside the packed observation are
// M::obs_field(f).
static std::vector<int> field_widths(const posg_model* m) {
  std::vector<int> w;
  if (m->kind == KIND_DRIVING) {
    const driving::Blob* b = (const driving::Blob*)m->h_blob;
    for (int c = 0; c < 15; ++c) w.push_back(4);          // local cells
    w.push_back(4);                                        // speed
    w.push_back(b->width); w.push_back(b->height);         // current coord
    w.push_back(b->width); w.push_back(b->height);         // destination coord
    w.push_back(2); w.push_back(2);                        // dest_reached, crashed
  } else if (m->kind == KIND_PREDATOR_PREY) {
    for (int c = 0; c < 25; ++c) w.push_back(4);
  } else if (m->kind == KIND_LBF) {
    const lbf::Blob* b = (const lbf::Blob*)m->h_blob;
    // players: Discrete(size+1, start=-1) x 2, Discrete(max_agent_level+1);
    // foods: Discrete(size+2, start=-1) x 2, Discrete(2*max_agent_level+1)
    for (int i = 0; i < lbf::NPL; ++i) { w.push_back(b->size + 1); w.push_back(b->size + 1); w.push_back(b->max_agent_level + 1); }
    for (int k = 0; k < lbf::NF; ++k) { w.push_back(b->size + 2); w.push_back(b->size + 2); w.push_back(2 * b->max_agent_level + 1); }
  }
  return w;
}

static int ensure_desc(posg_model* m) {
  if (m->d_flat_desc) return POSG_OK;
  const std::vector<int> widths = field_widths(m);
  std::vector<uint8_t> field, offs;    // per flattened position: its field and the one-hot offset in it
  for (size_t f = 0; f < widths.size(); ++f)
    for (int v = 0; v < widths[f]; ++v) { field.push_back((uint8_t)f); offs.push_back((uint8_t)v); }
  const size_t flat = field.size();
  if (flat == 0 || widths.size() > (size_t)vec::MAX_FIELDS) { set_error("model kind has no vector-env support"); return POSG_E_UNSUPPORTED; }
  if (flat % 4 != 0 || flat > (size_t)vec::MAX_FLAT) {
    set_error("flattened observation width %d is not a multiple of 4 (or > %d): unsupported by the vec-env kernel",
              (int)flat, vec::MAX_FLAT);
    return POSG_E_UNSUPPORTED;
  }
  std::vector<vec::Quad> d(flat / 4);
  for (size_t q = 0; q < d.size(); ++q) {
    const uint32_t w0 = field[4 * q] >> 2;
    uint32_t sel = 0, o = 0;
    for (int k = 0; k < 4; ++k) {
      sel |= (uint32_t)(field[4 * q + k] - 4 * w0) << (4 * k);   // PRMT nibble: byte index in {fr[0], fr[1]}
      o |= (uint32_t)offs[4 * q + k] << (8 * k);
    }
    d[q] = {o, w0 | (sel << 8)};
  }
  // one allocation: the quad descriptors, then the flattened offset of every field (uint16)
  std::vector<uint16_t> off(vec::MAX_FIELDS, 0);
  for (size_t f = 0, o = 0; f < widths.size(); o += (size_t)widths[f], ++f) off[f] = (uint16_t)o;
  const size_t qbytes = d.size() * sizeof(vec::Quad);
  POSG_CUDA(cudaMalloc(&m->d_flat_desc, qbytes + off.size() * sizeof(uint16_t)));
  POSG_CUDA(cudaMemcpy(m->d_flat_desc, d.data(), qbytes, cudaMemcpyHostToDevice));
  POSG_CUDA(cudaMemcpy((char*)m->d_flat_desc + qbytes, off.data(), off.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
  m->flat_dim = (int)flat;
  return POSG_OK;
}

template <class M, int LPR>
static int launch_embed_lpr(posg_model* m, int64_t n, const uint32_t* state, const float* wT, const float* bias, int H,
                            int apply_tanh, float* out, cudaStream_t s) {
  const uint16_t* off = (const uint16_t*)((const char*)m->d_flat_desc + (size_t)(m->flat_dim / 4) * sizeof(vec::Quad));
  if (n < (int64_t)sm_count() * 2 * vec::EMB_EPC) {   // small batch: more, smaller CTAs
    constexpr int EPC = 32;
    smem_optin(vec::vec_embed_kernel<M, LPR, EPC>, m->blob_bytes);
    vec::vec_embed_kernel<M, LPR, EPC><<<(int)((n + EPC - 1) / EPC), vec::EMB_THREADS, m->blob_bytes, s>>>(
        (const typename M::Blob*)m->d_blob, (uint32_t)m->blob_bytes, n, state, off, wT, bias, H, apply_tanh, out);
    POSG_CUDA(cudaGetLastError());
    return POSG_OK;
  }
  smem_optin(vec::vec_embed_kernel<M, LPR, vec::EMB_EPC>, m->blob_bytes);
  const int64_t need = (n + vec::EMB_EPC - 1) / vec::EMB_EPC;
  vec::vec_embed_kernel<M, LPR, vec::EMB_EPC><<<(int)(need < (1 << 30) ? need : (1 << 30)), vec::EMB_THREADS, m->blob_bytes, s>>>(
      (const typename M::Blob*)m->d_blob, (uint32_t)m->blob_bytes, n, state, off, wT, bias, H, apply_tanh, out);
  POSG_CUDA(cudaGetLastError());
  return POSG_OK;
}

template <class M>
static int launch_embed(posg_model* m, int64_t n, const uint32_t* state, const float* wT, const float* bias, int H,
                        int apply_tanh, float* out, cudaStream_t s) {
  const bool vec4 = ((((uintptr_t)wT | (uintptr_t)bias | (uintptr_t)out) & 15u) == 0);
  if (vec4 && H == 64) return launch_embed_lpr<M, 16>(m, n, state, wT, bias, H, apply_tanh, out, s);
  if (vec4 && H == 32) return launch_embed_lpr<M, 8>(m, n, state, wT, bias, H, apply_tanh, out, s);
  if (vec4 && H == 128) return launch_embed_lpr<M, 32>(m, n, state, wT, bias, H, apply_tanh, out, s);
  return launch_embed_lpr<M, 0>(m, n, state, wT, bias, H, apply_tanh, out, s);
}

template <class M, int EPC, int NT, int MINB>
static int launch_vec_epc(posg_model* m, int64_t n, uint32_t* state, uint16_t* ep_t, double* ep_ret, uint32_t* ep_count,
                          const long long* actions, uint64_t seed, uint32_t global_step, uint64_t env_offset,
                          int max_steps, int reset_only, float* obs_out, float* rew_out, uint8_t* term_out,
                          uint8_t* trunc_out, uint8_t* done_out, float* stat_r, int* stat_l, double now, double* ep_t0,
                          double* stat_t, uint32_t* step_counter, cudaStream_t s) {
  static int per_sm_cache[64] = {0};
  int& per_sm = per_sm_cache[m->device & 63];
  if (per_sm == 0) {
    // static row buffer + the staged model blob can exceed the 48 KB default
    POSG_CUDA(cudaFuncSetAttribute(vec::vec_step_kernel<M, EPC, NT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   m->blob_bytes > 65536 ? m->blob_bytes : 65536));
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, vec::vec_step_kernel<M, EPC, NT, MINB>, NT,
                                                      m->blob_bytes) != cudaSuccess || per_sm < 1) {
      cudaGetLastError(); per_sm = 1;
    }
  }
  const int64_t need = (n + EPC - 1) / EPC;
  static const bool persistent = getenv("POSG_VEC_PERSISTENT") != nullptr;   // dev A/B switch
  const int64_t cap = persistent ? (int64_t)sm_count() * per_sm : (int64_t)1 << 30;
  const int grid = (int)(need < cap ? need : cap);
  vec::vec_step_kernel<M, EPC, NT, MINB><<<grid, NT, m->blob_bytes, s>>>(
      (const typename M::Blob*)m->d_blob, (uint32_t)m->blob_bytes, n, state, ep_t, ep_ret, ep_count, actions,
      (const vec::Quad*)m->d_flat_desc, m->flat_dim, seed, global_step, env_offset, max_steps, reset_only,
      obs_out, rew_out, term_out, trunc_out, done_out, stat_r, stat_l, now, ep_t0, stat_t, step_counter);
  POSG_CUDA(cudaGetLastError());
  return POSG_OK;
}

// Small batches (the 8 192 envs per GPU of the reference-scale collector) are latency bound:
// 16 / 32 envs per chunk spread them over every SM; large batches step one env per thread.
template <class M, class... Args>
static int launch_vec(posg_model* m, int64_t n, Args... args) {
  const int64_t sms = sm_count();
  if (n <= sms * 16 * 4) return launch_vec_epc<M, 16, 256, 4>(m, n, args...);
  if (n <= sms * 32 * 8) return launch_vec_epc<M, 32, 256, 4>(m, n, args...);
  return launch_vec_epc<M, 256, 256, 4>(m, n, args...);
}

// envs.step + the next policy step's trunk layer 0 in one launch (vec_step_embed_kernel); H = 64 (the reference's
// trunk size, baseline_exps/exp_utils.py:101-104).  32 envs per CTA at collector-scale batches, 128 above.
template <class M>
static int launch_step_embed(posg_model* m, int64_t n, uint32_t* state, uint16_t* ep_t, double* ep_ret,
                             uint32_t* ep_count, const long long* actions, uint64_t seed, uint32_t global_step,
                             uint64_t env_offset, int max_steps, float* rew_out, uint8_t* term_out, uint8_t* trunc_out,
                             uint8_t* done_out, float* stat_r, int* stat_l, uint32_t* step_counter, const float* wT,
                             const float* bias, int apply_tanh, float* x_out, uint32_t* state_copy, cudaStream_t s) {
  const uint16_t* off = (const uint16_t*)((const char*)m->d_flat_desc + (size_t)(m->flat_dim / 4) * sizeof(vec::Quad));
#define POSG_STEP_EMBED(EPC)                                                                                        \
  do {                                                                                                              \
    smem_optin(vec::vec_step_embed_kernel<M, 16, EPC>, m->blob_bytes);                                             \
    const int64_t need = (n + EPC - 1) / EPC;                                                                       \
    vec::vec_step_embed_kernel<M, 16, EPC><<<(int)(need < (1 << 30) ? need : (1 << 30)), vec::EMB_THREADS,          \
                                             m->blob_bytes, s>>>(                                                   \
        (const typename M::Blob*)m->d_blob, (uint32_t)m->blob_bytes, n, state, ep_t, ep_ret, ep_count, actions, seed, \
        global_step, env_offset, max_steps, rew_out, term_out, trunc_out, done_out, stat_r, stat_l, step_counter, off, \
        wT, bias, apply_tanh, x_out, state_copy);                                                                   \
  } while (0)
  if (n < (int64_t)sm_count() * 2 * vec::EMB_EPC) POSG_STEP_EMBED(32);
  else POSG_STEP_EMBED(128);
#undef POSG_STEP_EMBED
  POSG_CUDA(cudaGetLastError());
  return POSG_OK;
}

// Generalised advantage estimation of the PPO learner (ppo/core.py:226-244) for a whole rollout
// batch in one launch: one thread per (env, agent) column scans the T steps backwards.  The
// float32 operations and their order are those of the reference's tensor expressions (the Python
// scalars gamma and gamma * gae_lambda are rounded to float32 once, as PyTorch does for float32
// tensors; the library is compiled with --fmad=false), so the result is bit-identical.
__global__ void gae_kernel(int T, int64_t N, const float* __restrict__ rewards, const float* __restrict__ dones,
                           const float* __restrict__ values, float gamma, float gamma_lambda,
                           float* __restrict__ advantages, float* __restrict__ returns) {
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < N; c += (int64_t)gridDim.x * blockDim.x) {
    float last = 0.0f;
    float next_value = values[(int64_t)T * N + c], next_done = dones[(int64_t)T * N + c];
    for (int t = T - 1; t >= 0; --t) {
      const int64_t i = (int64_t)t * N + c;
      const float v = values[i];
      const float nonterminal = __fsub_rn(1.0f, next_done);
      // delta = rewards[t] + gamma * nextvalues * nextnonterminal - values[t]
      const float delta = __fsub_rn(__fadd_rn(rewards[i], __fmul_rn(__fmul_rn(gamma, next_value), nonterminal)), v);
      // lastgaelam = delta + gamma * gae_lambda * nextnonterminal * lastgaelam
      last = __fadd_rn(delta, __fmul_rn(__fmul_rn(gamma_lambda, nonterminal), last));
      advantages[i] = last;
      returns[i] = __fadd_rn(last, v);
      next_value = v;
      next_done = dones[i];
    }
  }
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
                  double now, double* ep_t0, double* stat_t, uint32_t* step_counter, void* stream) {
  POSG_REQUIRE(m, "null model");
  int rc = ensure_desc(m); if (rc) return rc;
  if (n_envs == 0) return POSG_OK;
  POSG_REQUIRE(n_envs > 0 && state && ep_t && ep_ret && ep_count, "bad arguments");   // obs_out may be NULL (posg_vec_embed)
  POSG_REQUIRE(reset_only || (actions && rew_out && term_out && trunc_out && done_out && stat_r && stat_l),
               "step needs actions and every output buffer");
  POSG_REQUIRE(max_episode_steps > 0 && max_episode_steps < 65535, "bad max_episode_steps");
  POSG_REQUIRE((ep_t0 == nullptr) == (stat_t == nullptr) || reset_only, "ep_t0 and stat_t go together");
  POSG_REQUIRE(((uintptr_t)obs_out & 15u) == 0, "obs_out must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  switch (m->kind) {
    case KIND_DRIVING:
      return launch_vec<DrivingSim>(m, n_envs, state, ep_t, ep_ret, ep_count, (const long long*)actions, seed,
                                    global_step, env_offset, max_episode_steps, reset_only, obs_out, rew_out, term_out,
                                    trunc_out, done_out, stat_r, stat_l, now, ep_t0, stat_t, step_counter, s);
    case KIND_LBF:
      return launch_vec<LBFSim>(m, n_envs, state, ep_t, ep_ret, ep_count, (const long long*)actions, seed, global_step,
                                env_offset, max_episode_steps, reset_only, obs_out, rew_out, term_out, trunc_out,
                                done_out, stat_r, stat_l, now, ep_t0, stat_t, step_counter, s);
    case KIND_PREDATOR_PREY:
      return launch_vec<PPSim>(m, n_envs, state, ep_t, ep_ret, ep_count, (const long long*)actions, seed, global_step,
                               env_offset, max_episode_steps, reset_only, obs_out, rew_out, term_out, trunc_out,
                               done_out, stat_r, stat_l, now, ep_t0, stat_t, step_counter, s);
    default: set_error("model kind has no vector-env support"); return POSG_E_UNSUPPORTED;
  }
}

int posg_vec_embed(posg_model_t m, int64_t n_envs, const uint32_t* state, const float* wT, const float* bias,
                   int hidden, int apply_tanh, float* out, void* stream) {
  POSG_REQUIRE(m, "null model");
  int rc = ensure_desc(m); if (rc) return rc;
  if (n_envs == 0) return POSG_OK;
  POSG_REQUIRE(n_envs > 0 && state && wT && bias && out, "bad arguments");
  POSG_REQUIRE(hidden >= 1 && hidden <= vec::EMB_MAX_H, "hidden size outside [1, 128]");
  cudaStream_t s = (cudaStream_t)stream;
  switch (m->kind) {
    case KIND_DRIVING: return launch_embed<DrivingSim>(m, n_envs, state, wT, bias, hidden, apply_tanh, out, s);
    case KIND_LBF: return launch_embed<LBFSim>(m, n_envs, state, wT, bias, hidden, apply_tanh, out, s);
    case KIND_PREDATOR_PREY: return launch_embed<PPSim>(m, n_envs, state, wT, bias, hidden, apply_tanh, out, s);
    default: set_error("model kind has no vector-env support"); return POSG_E_UNSUPPORTED;
  }
}

int posg_vec_step_embed(posg_model_t m, int64_t n_envs, uint32_t* state, uint16_t* ep_t, double* ep_ret,
                        uint32_t* ep_count, const int64_t* actions, uint64_t seed, uint32_t global_step,
                        uint64_t env_offset, int max_episode_steps, float* rew_out, uint8_t* term_out,
                        uint8_t* trunc_out, uint8_t* done_out, float* stat_r, int32_t* stat_l, uint32_t* step_counter,
                        const float* wT, const float* bias, int hidden, int apply_tanh, float* x_out,
                        uint32_t* state_copy_out, void* stream) {
  POSG_REQUIRE(m, "null model");
  int rc = ensure_desc(m); if (rc) return rc;
  if (n_envs == 0) return POSG_OK;
  POSG_REQUIRE(n_envs > 0 && state && ep_t && ep_ret && ep_count && actions && rew_out && term_out && trunc_out &&
               done_out && stat_r && stat_l && wT && bias && x_out, "bad arguments");
  POSG_REQUIRE(max_episode_steps > 0 && max_episode_steps < 65535, "bad max_episode_steps");
  if (hidden != 64 || ((((uintptr_t)wT | (uintptr_t)bias | (uintptr_t)x_out) & 15u) != 0)) {
    set_error("posg_vec_step_embed: hidden size 64 and 16-byte aligned weights / output (use posg_vec_step + posg_vec_embed)");
    return POSG_E_UNSUPPORTED;
  }
  cudaStream_t s = (cudaStream_t)stream;
#define POSG_SE(MODEL)                                                                                            \
  return launch_step_embed<MODEL>(m, n_envs, state, ep_t, ep_ret, ep_count, (const long long*)actions, seed,      \
                                  global_step, env_offset, max_episode_steps, rew_out, term_out, trunc_out,       \
                                  done_out, stat_r, stat_l, step_counter, wT, bias, apply_tanh, x_out,            \
                                  state_copy_out, s)
  switch (m->kind) {
    case KIND_DRIVING: POSG_SE(DrivingSim);
    case KIND_LBF: POSG_SE(LBFSim);
    case KIND_PREDATOR_PREY: POSG_SE(PPSim);
    default: set_error("model kind has no vector-env support"); return POSG_E_UNSUPPORTED;
  }
#undef POSG_SE
}

int posg_gae(int T, int64_t n_cols, const float* rewards, const float* dones, const float* values, double gamma,
             double gae_lambda, float* advantages, float* returns, void* stream) {
  if (T <= 0 || n_cols == 0) return POSG_OK;
  POSG_REQUIRE(n_cols > 0 && rewards && dones && values && advantages && returns, "bad arguments");
  int ndev = posg_device_count();
  if (ndev < 0) return ndev;
  const int64_t need = (n_cols + 127) / 128;
  const int64_t cap = (int64_t)sm_count() * 16;
  gae_kernel<<<(int)(need < cap ? need : cap), 128, 0, (cudaStream_t)stream>>>(
      T, n_cols, rewards, dones, values, (float)gamma, (float)(gamma * gae_lambda), advantages, returns);
  POSG_CUDA(cudaGetLastError());
  return POSG_OK;
}

}  // extern "C"
