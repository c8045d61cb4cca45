// Model-generic batch kernels (one thread per env): model.step, observation, initial
// state and the fused random rollout, for adapters that implement the batch concept
// (see PPSim in pp.cuh).  Driving-v1 has its own hand-tuned kernels (driving_kernels.cu).
#pragma once
#include <stdint.h>

#include "philox.cuh"
#include "stage.cuh"

namespace posg {
namespace generic {

constexpr int THREADS = 256;

template <class M>
__global__ void __launch_bounds__(THREADS)
step_kernel(const typename M::Blob* __restrict__ gblob, uint32_t blob_size, int64_t n,
            const uint32_t* __restrict__ state_in, const uint8_t* __restrict__ actions_in,
            const uint32_t* __restrict__ noise_in, uint64_t seed, uint32_t step_idx, uint64_t particle_offset,
            uint32_t* __restrict__ state_out, uint64_t* __restrict__ obs_out, double* __restrict__ rewards_out,
            uint8_t* __restrict__ flags_out, uint8_t* __restrict__ actions_out,
            unsigned long long* __restrict__ live_count) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ alignas(8) uint64_t mbar;
  stage_blob(smem_raw, gblob, blob_size, &mbar);
  const typename M::Blob& B = *reinterpret_cast<const typename M::Blob*>(smem_raw);
  constexpr int SW = M::SW, OW = M::OW, NAG = M::NAG, NW = M::NW;
  unsigned live = 0;
  const int64_t stride = (int64_t)gridDim.x * THREADS;
  for (int64_t i = (int64_t)blockIdx.x * THREADS + threadIdx.x; i < n; i += stride) {
    typename M::State s;
    if constexpr (SW == 2) {   // one 8-byte load / store per env (cudaMalloc'ed arrays are aligned)
      const uint2 v = __ldg(reinterpret_cast<const uint2*>(state_in) + i);
      s.w[0] = v.x; s.w[1] = v.y;
    } else {
#pragma unroll
      for (int k = 0; k < SW; ++k) s.w[k] = __ldg(state_in + i * SW + k);
    }
    int act[NAG];
    uint32_t noise[NW];
    if (actions_in == nullptr || noise_in == nullptr)
      M::noise_from_philox(seed, particle_offset + (uint64_t)i, step_idx, act, noise);
    if (actions_in != nullptr) {
#pragma unroll
      for (int a = 0; a < NAG; ++a) act[a] = __ldg(actions_in + i * NAG + a);
    }
    if (noise_in != nullptr) {
#pragma unroll
      for (int k = 0; k < NW; ++k) noise[k] = __ldg(noise_in + i * NW + k);
    }
    unsigned long long obs[NAG][OW];
    double rew[NAG];
    uint32_t flags;
    bool was_live;
    M::step_all(B, s, act, noise, obs, rew, flags, was_live);
    live += was_live ? 1u : 0u;
    if constexpr (SW == 2) {
      reinterpret_cast<uint2*>(state_out)[i] = make_uint2(s.w[0], s.w[1]);
    } else {
#pragma unroll
      for (int k = 0; k < SW; ++k) state_out[i * SW + k] = s.w[k];
    }
    if constexpr (NAG == 2 && (OW == 1 || OW == 2)) {
      // 16-byte stores: both agents' words (OW == 1) or one agent's two words (OW == 2) at a time
      ulonglong2* o = reinterpret_cast<ulonglong2*>(obs_out) + i * OW;
      if constexpr (OW == 1) { o[0] = make_ulonglong2(obs[0][0], obs[1][0]); }
      else { o[0] = make_ulonglong2(obs[0][0], obs[0][1]); o[1] = make_ulonglong2(obs[1][0], obs[1][1]); }
      reinterpret_cast<double2*>(rewards_out)[i] = make_double2(rew[0], rew[1]);
      if (actions_out) reinterpret_cast<uchar2*>(actions_out)[i] = make_uchar2((unsigned char)act[0], (unsigned char)act[1]);
    } else {
#pragma unroll
      for (int a = 0; a < NAG; ++a) {
#pragma unroll
        for (int k = 0; k < OW; ++k) obs_out[(i * NAG + a) * OW + k] = obs[a][k];
        rewards_out[i * NAG + a] = rew[a];
        if (actions_out) actions_out[i * NAG + a] = (uint8_t)act[a];
      }
    }
    flags_out[i] = (uint8_t)flags;
  }
  if (live_count != nullptr) {
    for (int o = 16; o > 0; o >>= 1) live += __shfl_down_sync(0xffffffffu, live, o);
    if ((threadIdx.x & 31) == 0 && live) atomicAdd(live_count, (unsigned long long)live);
  }
}

// Fast path of model.step for random actions, particle_offset a multiple of four and whole blocks of 128
// envs.  Models whose step noise is ONE Philox word per env (M::ONE_WORD_NOISE: block = particle >> 2,
// word = particle & 3) share a block between four envs; the others (PredatorPrey: four words per env)
// compute one block per env in place -- the design of driving::step_fast_kernel: a warp owns 128 consecutive
// envs, lane L steps envs 32 j + L (every warp-level global access is one contiguous run), lane L
// computes the Philox block of envs 4 L .. 4 L + 3 and the words are transposed through shared
// memory; the next block's states and Philox rounds are issued one iteration ahead.
constexpr int FAST_WARP_ENVS = 128, FAST_EPT = 4;

#ifndef GENERIC_FAST_MIN_CTAS
#define GENERIC_FAST_MIN_CTAS 3
#endif
template <class M, bool WRITE_ACTIONS, bool COUNT_LIVE>
__global__ void __launch_bounds__(THREADS, GENERIC_FAST_MIN_CTAS)
step_fast_kernel(const typename M::Blob* __restrict__ gblob, uint32_t blob_size, uint32_t nblocks,
                 const uint2* __restrict__ state_in, const __grid_constant__ PhiloxKeys seed, uint32_t step_idx,
                 uint64_t group_offset,
                 uint2* __restrict__ state_out, ulonglong2* __restrict__ obs_out, double2* __restrict__ rewards_out,
                 uint8_t* __restrict__ flags_out, uchar2* __restrict__ actions_out,
                 unsigned long long* __restrict__ live_count) {
  static_assert(M::SW == 2 && M::NAG == 2 && (M::OW == 1 || M::OW == 2), "unsupported model");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ alignas(8) uint64_t mbar;
  __shared__ alignas(16) uint32_t s_words[THREADS / 32][FAST_WARP_ENVS];
  stage_blob(smem_raw, gblob, blob_size, &mbar);
  const typename M::Blob& B = *reinterpret_cast<const typename M::Blob*>(smem_raw);
  constexpr int OW = M::OW, NAG = M::NAG, NW = M::NW;
  unsigned live = 0;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t wstride = gridDim.x * (THREADS / 32);
  uint32_t wb = blockIdx.x * (THREADS / 32) + warp;
  uint2 nxt[FAST_EPT];
#pragma unroll
  for (int j = 0; j < FAST_EPT; ++j)
    nxt[j] = wb < nblocks ? __ldg(state_in + wb * FAST_WARP_ENVS + lane + 32 * j) : make_uint2(0u, 0u);
  uint4 blk_next = make_uint4(0u, 0u, 0u, 0u);
  if constexpr (M::ONE_WORD_NOISE) blk_next = philox_block(seed, group_offset + (uint64_t)wb * 32u + lane, step_idx, STREAM_STEP);
  for (; wb < nblocks; wb += wstride) {
    const uint32_t e0 = wb * FAST_WARP_ENVS + lane;
    uint2 sv[FAST_EPT];
#pragma unroll
    for (int j = 0; j < FAST_EPT; ++j) sv[j] = nxt[j];
    if (wb + wstride < nblocks) {
#pragma unroll
      for (int j = 0; j < FAST_EPT; ++j) nxt[j] = __ldg(state_in + (wb + wstride) * FAST_WARP_ENVS + lane + 32 * j);
    }
    uint32_t word[FAST_EPT];
    if constexpr (M::ONE_WORD_NOISE) {
      __syncwarp();
      *reinterpret_cast<uint4*>(&s_words[warp][4 * lane]) = blk_next;
      __syncwarp();
#pragma unroll
      for (int j = 0; j < FAST_EPT; ++j) word[j] = s_words[warp][32 * j + lane];
      blk_next = philox_block(seed, group_offset + (uint64_t)(wb + wstride) * 32u + lane, step_idx, STREAM_STEP);
    }
    uint4 pair_blk = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int j = 0; j < FAST_EPT; ++j) {
      typename M::State s;
      s.w[0] = sv[j].x; s.w[1] = sv[j].y;
      int act[NAG];
      uint32_t noise[NW];
      if constexpr (M::ONE_WORD_NOISE) {
        M::actions_from_word(word[j], act);
#pragma unroll
        for (int k = 0; k < NW; ++k) noise[k] = 0u;
      } else if constexpr (M::NOISE_PAIR) {
        // half a block per env: envs e and e + 32 of this thread share one (M::noise_counter; the launch
        // guarantees that the global index of e0 has bit 5 clear)
        if ((j & 1) == 0)
          pair_blk = philox_block(seed, M::noise_counter(group_offset * 4u + (uint64_t)(e0 + 32 * j)), step_idx, STREAM_STEP);
        M::noise_from_pair((j & 1) ? pair_blk.z : pair_blk.x, (j & 1) ? pair_blk.w : pair_blk.y, act, noise);
      } else {
        // one Philox block per env: counter = global particle index (the general kernel's convention)
        M::noise_from_philox(seed, group_offset * 4u + (uint64_t)(e0 + 32 * j), step_idx, act, noise);
      }
      unsigned long long obs[NAG][OW];
      double rew[NAG];
      uint32_t flags;
      bool was_live;
      M::step_all(B, s, act, noise, obs, rew, flags, was_live);
      if (COUNT_LIVE) live += was_live ? 1u : 0u;
      const uint32_t e = e0 + 32 * j;
      state_out[e] = make_uint2(s.w[0], s.w[1]);
      flags_out[e] = (uint8_t)flags;
      if (WRITE_ACTIONS) actions_out[e] = make_uchar2((unsigned char)act[0], (unsigned char)act[1]);
      if constexpr (OW == 1) {
        obs_out[e] = make_ulonglong2(obs[0][0], obs[1][0]);
      } else {
        obs_out[2 * e] = make_ulonglong2(obs[0][0], obs[0][1]);
        obs_out[2 * e + 1] = make_ulonglong2(obs[1][0], obs[1][1]);
      }
      rewards_out[e] = make_double2(rew[0], rew[1]);
    }
  }
  if (COUNT_LIVE) {
    for (int o = 16; o > 0; o >>= 1) live += __shfl_down_sync(0xffffffffu, live, o);
    if (lane == 0 && live) atomicAdd(live_count, (unsigned long long)live);
  }
}

template <class M>
__global__ void __launch_bounds__(THREADS)
observe_kernel(const typename M::Blob* __restrict__ gblob, uint32_t blob_size, int64_t n,
               const uint32_t* __restrict__ state, uint64_t* __restrict__ obs_out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ alignas(8) uint64_t mbar;
  stage_blob(smem_raw, gblob, blob_size, &mbar);
  const typename M::Blob& B = *reinterpret_cast<const typename M::Blob*>(smem_raw);
  constexpr int SW = M::SW, OW = M::OW, NAG = M::NAG;
  const int64_t stride = (int64_t)gridDim.x * THREADS;
  for (int64_t i = (int64_t)blockIdx.x * THREADS + threadIdx.x; i < n; i += stride) {
    typename M::State s;
#pragma unroll
    for (int k = 0; k < SW; ++k) s.w[k] = __ldg(state + i * SW + k);
    unsigned long long obs[NAG][OW];
    M::observe_all(B, s, obs);
#pragma unroll
    for (int a = 0; a < NAG; ++a)
#pragma unroll
      for (int k = 0; k < OW; ++k) obs_out[(i * NAG + a) * OW + k] = obs[a][k];
  }
}

template <class M>
__global__ void __launch_bounds__(THREADS)
initial_state_kernel(const typename M::Blob* __restrict__ gblob, uint32_t blob_size, int64_t n,
                     const uint8_t* __restrict__ picks, uint64_t seed, uint64_t particle_offset,
                     uint32_t* __restrict__ state_out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ alignas(8) uint64_t mbar;
  stage_blob(smem_raw, gblob, blob_size, &mbar);
  const typename M::Blob& B = *reinterpret_cast<const typename M::Blob*>(smem_raw);
  constexpr int SW = M::SW, NPICK = M::NPICK;
  const int64_t stride = (int64_t)gridDim.x * THREADS;
  for (int64_t i = (int64_t)blockIdx.x * THREADS + threadIdx.x; i < n; i += stride) {
    typename M::State s;
    if (picks != nullptr) {
      int p[NPICK];
#pragma unroll
      for (int k = 0; k < NPICK; ++k) p[k] = __ldg(picks + i * NPICK + k);
      M::initial_from_picks(B, p, s);
    } else {
      M::initial_state(B, seed, particle_offset + (uint64_t)i, s);
    }
#pragma unroll
    for (int k = 0; k < SW; ++k) state_out[i * SW + k] = s.w[k];
  }
}

struct DiscountTable { double v[256]; };

// MCTS._rollout with uniform-random policies; lanes refill from a work counter when
// their rollout ends (warp ballot compaction), as in the Driving rollout kernel.
template <class M>
__global__ void __launch_bounds__(THREADS)
rollout_kernel(const typename M::Blob* __restrict__ gblob, uint32_t blob_size, int64_t n,
               const uint32_t* __restrict__ state_in, int ego, int max_steps,
               const __grid_constant__ DiscountTable dtab, const __grid_constant__ PhiloxKeys seed, uint32_t step0,
               uint64_t particle_offset, double* __restrict__ returns_out, uint8_t* __restrict__ steps_out,
               uint32_t* __restrict__ state_out, unsigned long long* __restrict__ work_counter,
               unsigned long long* __restrict__ live_count) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ alignas(8) uint64_t mbar;
  __shared__ double sdisc[256];
  for (int t = threadIdx.x; t < 256; t += THREADS) sdisc[t] = dtab.v[t];
  stage_blob(smem_raw, gblob, blob_size, &mbar);
  const typename M::Blob& B = *reinterpret_cast<const typename M::Blob*>(smem_raw);
  constexpr int SW = M::SW, NAG = M::NAG, NW = M::NW;
  const unsigned lane = threadIdx.x & 31u;
  typename M::State s;
  int64_t item = -1;
  int k = 0;
  double ret = 0.0;
  unsigned long long executed = 0;
  bool exhausted = false;
  while (true) {
    const unsigned idle = __ballot_sync(0xffffffffu, item < 0);
    if (idle && !exhausted) {
      unsigned long long base = 0;
      const int cnt = __popc(idle), leader = __ffs(idle) - 1;
      if ((int)lane == leader) base = atomicAdd(work_counter, (unsigned long long)cnt);
      base = __shfl_sync(0xffffffffu, base, leader);
      if (item < 0) {
        const int64_t mine = (int64_t)base + __popc(idle & ((1u << lane) - 1u));
        if (mine < n) {
          item = mine;
#pragma unroll
          for (int w = 0; w < SW; ++w) s.w[w] = __ldg(state_in + item * SW + w);
          k = 0; ret = 0.0;
        }
      }
      if ((int64_t)base + cnt >= n) exhausted = true;
    }
    if (__ballot_sync(0xffffffffu, item >= 0) == 0u) break;
    if (item >= 0) {
      bool finished = (k >= max_steps);
      if (!finished) {
        int act[NAG];
        uint32_t noise[NW];
        M::noise_from_philox(seed, particle_offset + (uint64_t)item, step0 + (uint32_t)k, act, noise);
        double rew[NAG];
        uint32_t term; bool all_done;
        M::step_noobs(B, s, act, noise, rew, term, all_done);
        double r = rew[0];
#pragma unroll
        for (int a = 1; a < NAG; ++a) r = (a == ego) ? rew[a] : r;
        ret = __dadd_rn(ret, __dmul_rn(sdisc[k], r));
        ++k; ++executed;
        finished = ((term >> ego) & 1u) || all_done || (k >= max_steps);
      }
      if (finished) {
        returns_out[item] = ret;
        if (steps_out) steps_out[item] = (uint8_t)k;
        if (state_out) {
#pragma unroll
          for (int w = 0; w < SW; ++w) state_out[item * SW + w] = s.w[w];
        }
        item = -1;
      }
    }
  }
  if (live_count != nullptr) {
    for (int o = 16; o > 0; o >>= 1) executed += __shfl_down_sync(0xffffffffu, executed, o);
    if (lane == 0 && executed) atomicAdd(live_count, executed);
  }
}

}  // namespace generic
}  // namespace posg
