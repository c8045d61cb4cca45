// posg_model_step_host: POSGModel.step for a handful of rows that live in HOST memory -- the call the planners'
// plugin path makes once per tree level of a simulation (reference: planning/mcts.py:333,418 `self.model.step(
// hps.state, joint_action)`; planning/intmcp.py:468,562), where a stateful Python policy decides the next action
// from the previous step's result, so the rows cannot be batched.  One pinned block and one device block per
// handle, one copy in, the model's step kernel, one copy out, one stream synchronisation (up to 64 rows: no copies,
// the kernel works on the pinned block directly).
#include "common.h"

#include <cstdlib>
#include <cstring>
#include <mutex>

namespace posg {
namespace {

struct StepStage {
  std::mutex mu;
  unsigned char* h = nullptr;   // pinned
  unsigned char* d = nullptr;
  size_t bytes = 0;
  cudaStream_t st = nullptr;
};

void step_stage_free(void* p) {
  StepStage* s = (StepStage*)p;
  if (s->st) { cudaStreamSynchronize(s->st); cudaStreamDestroy(s->st); }
  cudaFreeHost(s->h);
  cudaFree(s->d);
  delete s;
}

StepStage* step_stage_of(posg_model* m) {
  static std::mutex create_mu;
  std::lock_guard<std::mutex> lock(create_mu);
  if (!m->step_stage) { m->step_stage = new StepStage(); m->step_stage_free = step_stage_free; }
  return (StepStage*)m->step_stage;
}

inline size_t up16(size_t x) { return (x + 15) & ~(size_t)15; }

}  // namespace
}  // namespace posg

using namespace posg;

extern "C" int posg_model_step_host(posg_model_t m, int64_t n, const uint32_t* state_in_host,
                                    const uint8_t* actions_host, const void* noise_host, int noise_row_bytes,
                                    uint32_t* state_out_host, uint64_t* obs_out_host, double* rewards_out_host,
                                    uint8_t* flags_out_host) {
  POSG_REQUIRE(m, "null model");
  if (n == 0) return POSG_OK;
  POSG_REQUIRE(n > 0 && n <= (1 << 20), "posg_model_step_host is for small batches (n <= 2^20)");
  POSG_REQUIRE(state_in_host && actions_host && noise_host && state_out_host && obs_out_host && rewards_out_host &&
               flags_out_host, "bad arguments");
  int n_act = 0, ow = 0, sw = 0;
  int rc = posg_model_info(m, &n_act, &ow, &sw);
  if (rc) return rc;
  const int na = m->num_agents;
  const bool driving = m->kind == KIND_DRIVING;
  POSG_REQUIRE(noise_row_bytes > 0 && (driving ? noise_row_bytes == 1 : noise_row_bytes % 4 == 0),
               "noise_row_bytes: 1 (Driving: processing-order index) or 4 * noise words");
  // one block, 16-byte aligned sections: [state_in | actions | noise] -> [state_out | obs | rewards | flags]
  const size_t b_sin = (size_t)n * sw * 4, b_act = (size_t)n * na, b_noise = (size_t)n * noise_row_bytes;
  const size_t b_obs = (size_t)n * na * ow * 8, b_rew = (size_t)n * na * 8, b_flags = (size_t)n;
  const size_t o_act = up16(b_sin), o_noise = o_act + up16(b_act), o_out = o_noise + up16(b_noise);
  const size_t o_sout = o_out, o_obs = o_sout + up16(b_sin), o_rew = o_obs + up16(b_obs), o_flags = o_rew + up16(b_rew);
  const size_t total = o_flags + up16(b_flags);

  StepStage* ss = step_stage_of(m);
  std::lock_guard<std::mutex> lock(ss->mu);
  POSG_CUDA(cudaSetDevice(m->device));
  if (!ss->st) POSG_CUDA(cudaStreamCreateWithFlags(&ss->st, cudaStreamNonBlocking));
  if (ss->bytes < total) {
    POSG_CUDA(cudaStreamSynchronize(ss->st));
    if (ss->h) cudaFreeHost(ss->h);
    if (ss->d) cudaFree(ss->d);
    ss->h = nullptr; ss->d = nullptr; ss->bytes = 0;
    const size_t cap = total < 4096 ? 4096 : total + total / 2;
    POSG_CUDA(cudaMallocHost((void**)&ss->h, cap));
    POSG_CUDA(cudaMalloc((void**)&ss->d, cap));
    ss->bytes = cap;
  }
  memcpy(ss->h, state_in_host, b_sin);
  memcpy(ss->h + o_act, actions_host, b_act);
  memcpy(ss->h + o_noise, noise_host, b_noise);
  // a handful of rows: the kernel reads and writes the pinned block itself (unified addressing maps it into the
  // device's address space) -- a few hundred bytes over PCIe inside the launch instead of two copy operations
  // around it.  POSG_STEP_HOST_STAGED=1 forces the copies (dev A/B switch).
  static const bool staged_only = getenv("POSG_STEP_HOST_STAGED") != nullptr;
  const bool direct = !staged_only && n <= 64;
  if (!direct) POSG_CUDA(cudaMemcpyAsync(ss->d, ss->h, o_out, cudaMemcpyHostToDevice, ss->st));
  unsigned char* d = direct ? ss->h : ss->d;
  if (driving) {
    rc = posg_driving_step(m, n, (const uint32_t*)d, d + o_act, d + o_noise, 0, 0, 0, (uint32_t*)(d + o_sout),
                           (uint64_t*)(d + o_obs), (double*)(d + o_rew), d + o_flags, nullptr, nullptr, ss->st);
  } else {
    rc = posg_model_step(m, n, (const uint32_t*)d, d + o_act, (const uint32_t*)(d + o_noise), 0, 0, 0,
                         (uint32_t*)(d + o_sout), (uint64_t*)(d + o_obs), (double*)(d + o_rew), d + o_flags, nullptr,
                         nullptr, ss->st);
  }
  if (rc) return rc;
  if (!direct) POSG_CUDA(cudaMemcpyAsync(ss->h + o_out, d + o_out, total - o_out, cudaMemcpyDeviceToHost, ss->st));
  POSG_CUDA(cudaStreamSynchronize(ss->st));
  memcpy(state_out_host, ss->h + o_sout, b_sin);
  memcpy(obs_out_host, ss->h + o_obs, b_obs);
  memcpy(rewards_out_host, ss->h + o_rew, b_rew);
  memcpy(flags_out_host, ss->h + o_flags, b_flags);
  return POSG_OK;
}
