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

constexpr int THREADS = 128;
constexpr int MAX_FLAT = 320;   // widest flattened per-agent observation supported

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
  __shared__ FlatDesc s_desc[MAX_FLAT];
  for (int p = threadIdx.x; p < flat_dim; p += THREADS) s_desc[p] = desc[p];
  stage_blob(smem_raw, gblob, blob_size, &mbar);  // contains a __syncthreads()
  const typename M::Blob& B = *reinterpret_cast<const typename M::Blob*>(smem_raw);
  // per-lane descriptors of positions lane, lane + 32, ...: shift (mod 32) | mask << 8 |
  // value << 16 | 32-bit half index (2 * word + (shift >= 32)) << 24
  constexpr int MAXK = MAX_FLAT / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int K = (flat_dim + 31) / 32;
  uint32_t ldesc[MAXK];
#pragma unroll
  for (int k = 0; k < MAXK; ++k) {
    const int p = lane + 32 * k;
    const FlatDesc d = s_desc[p < flat_dim ? p : 0];
    ldesc[k] = (uint32_t)(d.shift & 31) | ((uint32_t)d.mask << 8) | ((uint32_t)d.value << 16) |
               ((uint32_t)(2 * d.word + (d.shift >> 5)) << 24);
  }

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
    // cooperative one-hot expansion: one warp per (env, agent) row, lane L owns positions
    // L, L+32, ... (their descriptors are decoded once, above) -> every store instruction
    // writes 32 consecutive floats
    const int rows = (int)(((n - base) < THREADS ? (n - base) : THREADS) * NAG);
    float* out = obs_out + base * NAG * flat_dim;
    for (int row = warp; row < rows; row += THREADS / 32) {
      uint32_t w32[2 * OW];
#pragma unroll
      for (int k = 0; k < OW; ++k) {
        const unsigned long long w = s_obs[row * OW + k];
        w32[2 * k] = (uint32_t)w; w32[2 * k + 1] = (uint32_t)(w >> 32);
      }
      float* orow = out + (int64_t)row * flat_dim + lane;
#pragma unroll
      for (int k = 0; k < MAXK; ++k) {
        if (k < K && lane + 32 * k < flat_dim) {
          const uint32_t d = ldesc[k];
          uint32_t ww = w32[0];
#pragma unroll
          for (int h = 1; h < 2 * OW; ++h) ww = ((d >> 24) == (uint32_t)h) ? w32[h] : ww;
          orow[32 * k] = (((ww >> (d & 31u)) & ((d >> 8) & 255u)) == ((d >> 16) & 255u)) ? 1.0f : 0.0f;
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace vec
}  // namespace posg
