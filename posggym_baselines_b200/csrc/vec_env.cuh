// Vector-env step for the PPO rollout collector: what
//   StackEnv(RecordEpisodeStatistics(SyncVectorEnv([FlattenObservations(env)] * n)))
// does per call of envs.step(actions) in the reference (ppo/config.py:223-232,
// ppo/worker.py:176-216), for a whole batch of envs in one launch:
//   model.step -> TimeLimit truncation -> episode statistics -> auto-reset of finished
//   envs (the reset observation is returned) -> one-hot flattened float32 observations
//   stacked env-major (row = env * num_agents + agent).
// A CTA of NT threads works on chunks of EPC envs.  Phase 1: one thread steps one env and
// leaves the observation of each agent as a row of field bytes in shared memory.  Phase 2: all
// threads expand the chunk's rows to one-hot floats, one 16-byte store per thread per iteration
// (consecutive threads -> consecutive float4s: every warp store covers 512 contiguous bytes).
// Large batches run <EPC = NT = 256> (every thread steps an env, no idle lanes in phase 1) with
// one CTA per chunk, so the hardware block scheduler balances the SMs; small, latency-bound
// batches run <16 or 32 envs, 256 threads> so that every SM gets work.
// The output (2 x 300 floats per env for LevelBasedForaging-v3) is a pure write stream and the
// kernel is HBM-write bound; the stores are streaming (st.global.cs).
#pragma once
#include <stdint.h>

#include "philox.cuh"
#include "stage.cuh"

namespace posg {
namespace vec {

constexpr int MAX_FLAT = 512;     // widest flattened per-agent observation supported (floats)
constexpr int MAX_FIELDS = 32;    // observation fields (one byte each) per row
constexpr int ROW_WORDS = 10;     // 32 field bytes + 8 bytes of padding (the gather reads two words)

// Descriptor of one float4 (4 consecutive flattened positions of a row):
//   x = the four one-hot offsets inside their fields, one byte each
//   y = first field word (field >> 2 of position 0) | PRMT selector of the four field bytes << 8
struct Quad { uint32_t offs, ctrl; };

template <class M>
__device__ __forceinline__ void field_bytes(const unsigned long long (&obs)[M::OW], uint32_t (&fw)[MAX_FIELDS / 4]) {
  uint32_t h[2 * M::OW];
#pragma unroll
  for (int k = 0; k < M::OW; ++k) { h[2 * k] = (uint32_t)obs[k]; h[2 * k + 1] = (uint32_t)(obs[k] >> 32); }
#pragma unroll
  for (int w = 0; w < MAX_FIELDS / 4; ++w) {
    uint32_t v = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int f = 4 * w + b;
      if (f < M::NFIELD) {
        const uint32_t d = M::obs_field(f);
        v |= ((h[d & 255u] >> ((d >> 8) & 255u)) & ((1u << (d >> 16)) - 1u)) << (8 * b);
      }
    }
    fw[w] = v;
  }
}

// envs.step / envs.reset of ONE env (ppo/worker.py:176-216 through the wrapper stack of ppo/config.py:223-232): model
// step, TimeLimit, RecordEpisodeStatistics, auto-reset; leaves the observations the call returns in `obs`.
template <class M>
__device__ __forceinline__ void step_one_env(
    const typename M::Blob& B, int64_t i, uint32_t* __restrict__ state, uint16_t* __restrict__ ep_t,
    double* __restrict__ ep_ret, uint32_t* __restrict__ ep_count, const long long* __restrict__ actions, uint64_t seed,
    uint32_t global_step, uint64_t env_offset, int max_episode_steps, int reset_only, float* __restrict__ rew_out,
    uint8_t* __restrict__ term_out, uint8_t* __restrict__ trunc_out, uint8_t* __restrict__ done_out,
    float* __restrict__ stat_r, int* __restrict__ stat_l, double now, double* __restrict__ ep_t0,
    double* __restrict__ stat_t, unsigned long long (&obs)[M::NAG][M::OW]) {
  constexpr int SW = M::SW, NAG = M::NAG, NW = M::NW;
  {
      typename M::State s;
      const uint64_t env = env_offset + (uint64_t)i;
      if (reset_only) {
        // envs.reset(): fresh episode for every env
        M::initial_state(B, seed, env, s);  // episode 0 of this env: Philox step field 0
        M::observe_all(B, s, obs);
        ep_t[i] = 0; ep_count[i] = 0;
        if (ep_t0 != nullptr) ep_t0[i] = now;
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
        if (ep_t0 != nullptr) {
          // RecordEpisodeStatistics "t": wall time of the episode that ends with this step
          const double t0 = ep_t0[i];
          stat_t[i] = done ? now - t0 : 0.0;
          if (done) ep_t0[i] = now;
        }
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
  }
}

template <class M, int EPC, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB)
vec_step_kernel(const typename M::Blob* __restrict__ gblob, uint32_t blob_size, int64_t n,
                uint32_t* __restrict__ state, uint16_t* __restrict__ ep_t, double* __restrict__ ep_ret,
                uint32_t* __restrict__ ep_count, const long long* __restrict__ actions,
                const Quad* __restrict__ desc, int flat_dim, uint64_t seed, uint32_t global_step,
                uint64_t env_offset, int max_episode_steps, int reset_only,
                float* __restrict__ obs_out, float* __restrict__ rew_out, uint8_t* __restrict__ term_out,
                uint8_t* __restrict__ trunc_out, uint8_t* __restrict__ done_out, float* __restrict__ stat_r,
                int* __restrict__ stat_l, double now, double* __restrict__ ep_t0, double* __restrict__ stat_t,
                uint32_t* __restrict__ step_counter) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ alignas(8) uint64_t mbar;
  constexpr int SW = M::SW, OW = M::OW, NAG = M::NAG, NW = M::NW;
  static_assert(M::NFIELD <= MAX_FIELDS, "too many observation fields");
  __shared__ uint32_t s_fields[EPC * NAG * ROW_WORDS];
  __shared__ Quad s_desc[MAX_FLAT / 4];
  const int Q = flat_dim >> 2;                 // float4s per row
  for (int p = threadIdx.x; p < Q; p += NT) s_desc[p] = desc[p];
  stage_blob(smem_raw, gblob, blob_size, &mbar);  // contains a __syncthreads()
  const typename M::Blob& B = *reinterpret_cast<const typename M::Blob*>(smem_raw);
  // CUDA-graph replays bake the scalar arguments: the step index then comes from device memory,
  // step_counter[0] + 1; the last CTA to finish publishes it for the next replay (every CTA has read
  // the counter before it can finish; step_counter[1] counts the finished CTAs)
  if (step_counter != nullptr) global_step = *((volatile uint32_t*)step_counter) + 1u;
  const int tid = threadIdx.x;
  const int row0 = tid / Q, q0 = tid - row0 * Q;      // this thread's first float4 of a chunk
  const int dr = NT / Q, dq = NT - dr * Q;  // and its stride

  for (int64_t base = (int64_t)blockIdx.x * EPC; base < n; base += (int64_t)gridDim.x * EPC) {
    const int64_t i = base + tid;
    if (tid < EPC && i < n) {
      unsigned long long obs[NAG][OW];
      step_one_env<M>(B, i, state, ep_t, ep_ret, ep_count, actions, seed, global_step, env_offset, max_episode_steps,
                      reset_only, rew_out, term_out, trunc_out, done_out, stat_r, stat_l, now, ep_t0, stat_t, obs);
#pragma unroll
      for (int a = 0; a < NAG; ++a) {
        uint32_t fw[MAX_FIELDS / 4];
        field_bytes<M>(obs[a], fw);
        uint32_t* dst = s_fields + (tid * NAG + a) * ROW_WORDS;
#pragma unroll
        for (int w = 0; w < (M::NFIELD + 3) / 4; ++w) dst[w] = fw[w];
      }
    }
    __syncthreads();
    // one-hot expansion of the chunk: float4 j of the chunk = (row j / Q, quad j % Q)
    const int64_t left = n - base;
    const int total = obs_out == nullptr ? 0 : (int)(left < EPC ? left : EPC) * NAG * Q;   // no one-hot rows wanted
    float4* out4 = reinterpret_cast<float4*>(obs_out + base * NAG * flat_dim);
    int row = row0, q = q0;
    for (int j = tid; j < total; j += NT) {
      const Quad d = s_desc[q];
      const uint32_t* fr = s_fields + row * ROW_WORDS + (d.ctrl & 255u);
      const uint32_t x = __byte_perm(fr[0], fr[1], d.ctrl >> 8) ^ d.offs;   // byte k == 0 <=> position k is hot
      float4 o;
      o.x = (x & 0x000000FFu) ? 0.0f : 1.0f;
      o.y = (x & 0x0000FF00u) ? 0.0f : 1.0f;
      o.z = (x & 0x00FF0000u) ? 0.0f : 1.0f;
      o.w = (x & 0xFF000000u) ? 0.0f : 1.0f;
      __stcs(out4 + j, o);
      q += dq; row += dr;
      if (q >= Q) { q -= Q; ++row; }
    }
    __syncthreads();
  }
  if (step_counter != nullptr && !reset_only && threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(step_counter + 1, 1u) == gridDim.x - 1u) {
      step_counter[1] = 0u;
      step_counter[0] = global_step;
      __threadfence();
    }
  }
}

// First trunk layer of the PPO network applied to the flattened one-hot observation WITHOUT
// materialising it (ppo/network.py:171-178: nn.Linear(input_size, trunk_sizes[0]) + nn.Tanh, fed by
// the worker at ppo/worker.py:143-163): a one-hot row times W is the sum of one weight column per
// observation field, so  hidden[row] = act(bias + sum_f wT[offset_f + value_f]).  Per env-step this
// reads 30 x 256 B of (L1-resident) weights and writes H floats per agent instead of writing and
// re-reading 300 floats per agent.  Phase 1 as in vec_step_kernel (observation -> field bytes in
// shared memory), phase 2: one warp per row, lane j owns hidden units j, j + 32, ...
constexpr int EMB_THREADS = 256, EMB_EPC = 128, EMB_MAX_H = 128;

// LPR = lanes per row (each lane owns four consecutive hidden units: one 16-byte weight load per
// field); H = 4 * LPR in {32, 64, 128}.  LPR == 0: any H <= 128, one warp per row, scalar loads.
// EPC = envs per CTA (128 for large batches; 32 at the few-thousand-env batches of the reference-scale collector,
// where 128 would leave most SMs without a CTA)
template <class M, int LPR, int EPC>
__global__ void __launch_bounds__(EMB_THREADS)
vec_embed_kernel(const typename M::Blob* __restrict__ gblob, uint32_t blob_size, int64_t n,
                 const uint32_t* __restrict__ state, const uint16_t* __restrict__ field_offset,
                 const float* __restrict__ wT, const float* __restrict__ bias, int H, int apply_tanh,
                 float* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ alignas(8) uint64_t mbar;
  constexpr int SW = M::SW, OW = M::OW, NAG = M::NAG, NF = M::NFIELD;
  // per row: the weight-row index offset_f + value_f of every field.  Row stride 17 words: the env
  // threads of phase 1 store to consecutive rows, which then fall into consecutive banks
  __shared__ uint16_t s_idx[EPC * NAG][MAX_FIELDS + 2];
  __shared__ uint16_t s_off[MAX_FIELDS];
  if (threadIdx.x < MAX_FIELDS) s_off[threadIdx.x] = threadIdx.x < NF ? field_offset[threadIdx.x] : 0;
  stage_blob(smem_raw, gblob, blob_size, &mbar);
  const typename M::Blob& B = *reinterpret_cast<const typename M::Blob*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int64_t base = (int64_t)blockIdx.x * EPC; base < n; base += (int64_t)gridDim.x * EPC) {
    const int64_t i = base + tid;
    if (tid < EPC && i < n) {
      typename M::State s;
#pragma unroll
      for (int k = 0; k < SW; ++k) s.w[k] = __ldg(state + i * SW + k);
      unsigned long long obs[NAG][OW];
      M::observe_all(B, s, obs);
#pragma unroll
      for (int a = 0; a < NAG; ++a) {
        uint32_t fw[MAX_FIELDS / 4];
        field_bytes<M>(obs[a], fw);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&s_idx[tid * NAG + a][0]);
#pragma unroll
        for (int f = 0; f < NF; f += 2) {
          const uint32_t v0 = (fw[f >> 2] >> (8 * (f & 3))) & 255u;
          const uint32_t v1 = f + 1 < NF ? (fw[(f + 1) >> 2] >> (8 * ((f + 1) & 3))) & 255u : 0u;
          dst[f >> 1] = (v0 + s_off[f]) | ((v1 + (f + 1 < NF ? s_off[f + 1] : 0u)) << 16);
        }
      }
    }
    __syncthreads();
    const int64_t left = n - base;
    const int rows = (int)(left < EPC ? left : EPC) * NAG;
    if constexpr (LPR > 0) {
      constexpr int RPW = 32 / LPR;                 // rows per warp iteration
      const int sub = lane / LPR, col = (lane % LPR) * 4;
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + col));
      for (int row = warp * RPW + sub; row < rows; row += (EMB_THREADS / 32) * RPW) {
        float4 acc = b4;
#pragma unroll
        for (int f = 0; f < NF; ++f) {
          const float4 w = __ldg(reinterpret_cast<const float4*>(wT + (int)s_idx[row][f] * (4 * LPR) + col));
          acc.x = __fadd_rn(acc.x, w.x); acc.y = __fadd_rn(acc.y, w.y);
          acc.z = __fadd_rn(acc.z, w.z); acc.w = __fadd_rn(acc.w, w.w);
        }
        if (apply_tanh) { acc.x = tanhf(acc.x); acc.y = tanhf(acc.y); acc.z = tanhf(acc.z); acc.w = tanhf(acc.w); }
        *reinterpret_cast<float4*>(out + (base * NAG + row) * (4 * LPR) + col) = acc;
      }
    } else {
      const int KH = (H + 31) >> 5;
      for (int row = warp; row < rows; row += EMB_THREADS / 32) {
        float acc[EMB_MAX_H / 32];
#pragma unroll
        for (int k = 0; k < EMB_MAX_H / 32; ++k) acc[k] = (k < KH && lane + 32 * k < H) ? __ldg(bias + lane + 32 * k) : 0.0f;
#pragma unroll
        for (int f = 0; f < NF; ++f) {
          const float* wr = wT + (int64_t)s_idx[row][f] * H + lane;
#pragma unroll
          for (int k = 0; k < EMB_MAX_H / 32; ++k)
            if (k < KH && lane + 32 * k < H) acc[k] = __fadd_rn(acc[k], __ldg(wr + 32 * k));
        }
        float* orow = out + (base * NAG + row) * H + lane;
#pragma unroll
        for (int k = 0; k < EMB_MAX_H / 32; ++k)
          if (k < KH && lane + 32 * k < H) orow[32 * k] = apply_tanh ? tanhf(acc[k]) : acc[k];
      }
    }
    __syncthreads();
  }
}

// envs.step FUSED with the next policy step's trunk layer 0 (the collector's loop: worker.py:143-178): phase 1 is
// vec_step_kernel's (step_one_env), phase 2 is vec_embed_kernel's embedding-sum on the observations the step just
// produced -- no one-hot rows, no second kernel that recomputes the observation from the state, and the new state is
// also written to `state_copy` (the rollout storage's row for the next step) when that is not NULL.  H = 4 * LPR.
template <class M, int LPR, int EPC>
__global__ void __launch_bounds__(EMB_THREADS)
vec_step_embed_kernel(const typename M::Blob* __restrict__ gblob, uint32_t blob_size, int64_t n,
                      uint32_t* __restrict__ state, uint16_t* __restrict__ ep_t, double* __restrict__ ep_ret,
                      uint32_t* __restrict__ ep_count, const long long* __restrict__ actions, uint64_t seed,
                      uint32_t global_step, uint64_t env_offset, int max_episode_steps, float* __restrict__ rew_out,
                      uint8_t* __restrict__ term_out, uint8_t* __restrict__ trunc_out, uint8_t* __restrict__ done_out,
                      float* __restrict__ stat_r, int* __restrict__ stat_l, uint32_t* __restrict__ step_counter,
                      const uint16_t* __restrict__ field_offset, const float* __restrict__ wT,
                      const float* __restrict__ bias, int apply_tanh, float* __restrict__ x_out,
                      uint32_t* __restrict__ state_copy) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ alignas(8) uint64_t mbar;
  constexpr int SW = M::SW, OW = M::OW, NAG = M::NAG, NF = M::NFIELD;
  __shared__ uint16_t s_idx[EPC * NAG][MAX_FIELDS + 2];
  __shared__ uint16_t s_off[MAX_FIELDS];
  if (threadIdx.x < MAX_FIELDS) s_off[threadIdx.x] = threadIdx.x < NF ? field_offset[threadIdx.x] : 0;
  stage_blob(smem_raw, gblob, blob_size, &mbar);
  const typename M::Blob& B = *reinterpret_cast<const typename M::Blob*>(smem_raw);
  if (step_counter != nullptr) global_step = *((volatile uint32_t*)step_counter) + 1u;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int64_t base = (int64_t)blockIdx.x * EPC; base < n; base += (int64_t)gridDim.x * EPC) {
    const int64_t i = base + tid;
    if (tid < EPC && i < n) {
      unsigned long long obs[NAG][OW];
      step_one_env<M>(B, i, state, ep_t, ep_ret, ep_count, actions, seed, global_step, env_offset, max_episode_steps, 0,
                      rew_out, term_out, trunc_out, done_out, stat_r, stat_l, 0.0, nullptr, nullptr, obs);
      if (state_copy != nullptr) {
#pragma unroll
        for (int k = 0; k < SW; ++k) state_copy[i * SW + k] = state[i * SW + k];
      }
#pragma unroll
      for (int a = 0; a < NAG; ++a) {
        uint32_t fw[MAX_FIELDS / 4];
        field_bytes<M>(obs[a], fw);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&s_idx[tid * NAG + a][0]);
#pragma unroll
        for (int f = 0; f < NF; f += 2) {
          const uint32_t v0 = (fw[f >> 2] >> (8 * (f & 3))) & 255u;
          const uint32_t v1 = f + 1 < NF ? (fw[(f + 1) >> 2] >> (8 * ((f + 1) & 3))) & 255u : 0u;
          dst[f >> 1] = (v0 + s_off[f]) | ((v1 + (f + 1 < NF ? s_off[f + 1] : 0u)) << 16);
        }
      }
    }
    __syncthreads();
    const int64_t left = n - base;
    const int rows = (int)(left < EPC ? left : EPC) * NAG;
    constexpr int RPW = 32 / LPR;
    const int sub = lane / LPR, col = (lane % LPR) * 4;
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + col));
    for (int row = warp * RPW + sub; row < rows; row += (EMB_THREADS / 32) * RPW) {
      float4 acc = b4;
#pragma unroll
      for (int f = 0; f < NF; ++f) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(wT + (int)s_idx[row][f] * (4 * LPR) + col));
        acc.x = __fadd_rn(acc.x, w.x); acc.y = __fadd_rn(acc.y, w.y);
        acc.z = __fadd_rn(acc.z, w.z); acc.w = __fadd_rn(acc.w, w.w);
      }
      if (apply_tanh) { acc.x = tanhf(acc.x); acc.y = tanhf(acc.y); acc.z = tanhf(acc.z); acc.w = tanhf(acc.w); }
      *reinterpret_cast<float4*>(x_out + (base * NAG + row) * (4 * LPR) + col) = acc;
    }
    __syncthreads();
  }
  if (step_counter != nullptr && threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(step_counter + 1, 1u) == gridDim.x - 1u) {
      step_counter[1] = 0u;
      step_counter[0] = global_step;
      __threadfence();
    }
  }
}

}  // namespace vec
}  // namespace posg
