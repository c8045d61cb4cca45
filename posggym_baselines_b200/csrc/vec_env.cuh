// Vector-env step for the PPO rollout collector: what
//   StackEnv(RecordEpisodeStatistics(SyncVectorEnv([FlattenObservations(env)] * n)))
// does per call of envs.step(actions) in the reference (ppo/config.py:223-232,
// ppo/worker.py:176-216), for a whole batch of envs in one launch:
//   model.step -> TimeLimit truncation -> episode statistics -> auto-reset of finished
//   envs (the reset observation is returned) -> one-hot flattened float32 observations
//   stacked env-major (row = env * num_agents + agent).
// One thread steps one env; the CTA then expands its packed observations to one-hot
// float rows cooperatively so that every global store is a coalesced 128-byte line.
#pragma once
#include <stdint.h>

#include "philox.cuh"
#include "stage.cuh"

namespace posg {
namespace vec {

constexpr int THREADS = 256;

// one-hot descriptor of a flattened position: which packed word, shift, mask, hot value
struct FlatDesc { uint8_t word, shift, mask, value; };

template <class M>
__global__ void __launch_bounds__(THREADS)
vec_step_kernel(const typename M::Blob* __restrict__ gblob, uint32_t blob_size, int64_t n,
                uint32_t* __restrict__ state, uint16_t* __restrict__ ep_t, double* __restrict__ ep_ret,
                uint32_t* __restrict__ ep_count, const long long* __restrict__ actions,
                const FlatDesc* __restrict__ desc, int flat_dim, uint64_t seed, uint32_t global_step,
                uint64_t env_offset, int max_episode_steps, int reset_only,
                float* __restrict__ obs_out, float* __restrict__ rew_out, uint8_t* __restrict__ term_out,
                uint8_t* __restrict__ trunc_out, uint8_t* __restrict__ done_out, float* __restrict__ stat_r,
                int* __restrict__ stat_l) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ alignas(8) uint64_t mbar;
  constexpr int SW = M::SW, OW = M::OW, NAG = M::NAG, NW = M::NW;
  __shared__ unsigned long long s_obs[THREADS * NAG * OW];
  stage_blob(smem_raw, gblob, blob_size, &mbar);
  const typename M::Blob& B = *reinterpret_cast<const typename M::Blob*>(smem_raw);

  for (int64_t base = (int64_t)blockIdx.x * THREADS; base < n; base += (int64_t)gridDim.x * THREADS) {
    const int64_t i = base + threadIdx.x;
    if (i < n) {
      typename M::State s;
      unsigned long long obs[NAG][OW];
      const uint64_t env = env_offset + (uint64_t)i;
      if (reset_only) {
        // envs.reset(): fresh episode for every env
        M::initial_state(B, seed, env, s);  // episode 0 of this env: Philox step field 0
        M::observe_all(B, s, obs);
        ep_t[i] = 0; ep_count[i] = 0;
#pragma unroll
        for (int a = 0; a < NAG; ++a) ep_ret[i * NAG + a] = 0.0;
      } else {
#pragma unroll
        for (int k = 0; k < SW; ++k) s.w[k] = state[i * SW + k];
        int act[NAG], ract[NAG];
        uint32_t noise[NW];
        M::noise_from_philox(seed, env, global_step, ract, noise);
#pragma unroll
        for (int a = 0; a < NAG; ++a) act[a] = (int)actions[i * NAG + a];
        double rew[NAG];
        uint32_t flags; bool was_live;
        M::step_all(B, s, act, noise, obs, rew, flags, was_live);
        const int t = (int)ep_t[i] + 1;
        const bool trunc = t >= max_episode_steps;
        const bool done = (flags & 0x80u) || trunc;
#pragma unroll
        for (int a = 0; a < NAG; ++a) {
          const double tot = ep_ret[i * NAG + a] + rew[a];
          rew_out[i * NAG + a] = (float)rew[a];
          term_out[i * NAG + a] = (uint8_t)((flags >> a) & 1u);
          trunc_out[i * NAG + a] = trunc ? 1 : 0;
          stat_r[i * NAG + a] = done ? (float)tot : 0.0f;
          ep_ret[i * NAG + a] = done ? 0.0 : tot;
        }
        stat_l[i] = done ? t : 0;
        done_out[i] = done ? 1 : 0;
        if (done) {
          // auto-reset (SyncVectorEnv): the observation returned is the new episode's first
          const uint32_t ep = ep_count[i] + 1;
          ep_count[i] = ep;
          M::initial_state(B, seed ^ (0x9E3779B97F4A7C15ULL * (uint64_t)ep), env, s);
          M::observe_all(B, s, obs);
          ep_t[i] = 0;
        } else {
          ep_t[i] = (uint16_t)t;
        }
      }
#pragma unroll
      for (int k = 0; k < SW; ++k) state[i * SW + k] = s.w[k];
#pragma unroll
      for (int a = 0; a < NAG; ++a)
#pragma unroll
        for (int k = 0; k < OW; ++k) s_obs[(threadIdx.x * NAG + a) * OW + k] = obs[a][k];
    }
    __syncthreads();
    // cooperative one-hot expansion: rows of this CTA are contiguous in obs_out
    const int64_t rows = ((n - base) < THREADS ? (n - base) : THREADS) * NAG;
    const int64_t total = rows * flat_dim;
    float* out = obs_out + base * NAG * flat_dim;
    for (int64_t q = threadIdx.x; q < total; q += THREADS) {
      const int row = (int)(q / flat_dim), p = (int)(q - (int64_t)row * flat_dim);
      const FlatDesc d = desc[p];
      const unsigned long long w = s_obs[row * OW + d.word];
      out[q] = (((unsigned)(w >> d.shift) & d.mask) == d.value) ? 1.0f : 0.0f;
    }
    __syncthreads();
  }
}

}  // namespace vec
}  // namespace posg
