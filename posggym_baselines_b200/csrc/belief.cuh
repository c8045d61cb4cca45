// Particle-belief kernels: initial belief by rejection (MCTS._initial_update,
// planning/mcts.py:175-227), reinvigoration by rejection with reject padding
// (BeliefRejectionSampler._rejection_sample, planning/belief.py:145-194, driven by
// MCTS._reinvigorate, mcts.py:651-700) and re-rooting helpers (MCTS._update,
// mcts.py:229-263).  The reference's loops are strictly sequential ("first N accepted in
// attempt order, then the first missing rejects in attempt order"); here one CTA runs
// 1024 attempts per round and an order-preserving block scan reproduces that ordering.
#pragma once
#include <stdint.h>

#include "philox.cuh"
#include "search.cuh"

namespace posg {
namespace belief {

constexpr int CTA = 1024;

// exclusive prefix sum of a 0/1 flag over the CTA; returns this thread's offset, total in `total`
__device__ __forceinline__ int block_scan(bool flag, int* warp_sums, int& total) {
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const unsigned b = __ballot_sync(0xffffffffu, flag);
  const int within = __popc(b & ((1u << lane) - 1u));
  if (lane == 0) warp_sums[warp] = __popc(b);
  __syncthreads();
  int base = 0, tot = 0;
  for (int w = 0; w < CTA / 32; ++w) {
    const int v = warp_sums[w];
    if (w < (int)warp) base += v;
    tot += v;
  }
  __syncthreads();
  total = tot;
  return base + within;
}

// out[0 .. n_needed) <- accepted next-states in attempt order, padded with rejected ones.
// result[0] = accepted count (<= n_needed), result[1] = attempts made.
template <class M>
__device__ void reinvigorate(const typename M::Blob& B, const typename M::Ctx& ctx,
                             const uint32_t* __restrict__ parent, long long n_parent,
                             int ego, int action, const unsigned long long* __restrict__ target_obs,
                             long long n_needed, long long max_attempts, uint64_t seed, uint64_t draw_id,
                             uint32_t* __restrict__ out, uint32_t* __restrict__ rej_scratch,
                             long long* __restrict__ result) {
  constexpr int OW = M::OW, SW = M::SW;
  __shared__ int warp_sums[CTA / 32];
  unsigned long long tgt[OW];
#pragma unroll
  for (int k = 0; k < OW; ++k) tgt[k] = target_obs[k];
  long long n_acc = 0, n_rej = 0, attempts = 0;
  while (n_acc < n_needed && attempts < max_attempts) {
    const long long att = attempts + threadIdx.x;
    const bool live = att < max_attempts;
    typename M::State st;
    bool ok = false;
    if (live) {
      const uint4 w = philox_block(seed, draw_id + (uint64_t)att, 0xFFFEu, STREAM_BELIEF);
      const long long pi = (long long)(((unsigned long long)w.x * (unsigned long long)n_parent) >> 32);
#pragma unroll
      for (int k = 0; k < SW; ++k) st.w[k] = parent[pi * SW + k];
      unsigned long long obs[OW];
      double r; bool done;
      M::step(B, ctx, st, ego, action, w.y, w.z, obs, r, done);
      ok = true;
#pragma unroll
      for (int k = 0; k < OW; ++k) ok &= (obs[k] == tgt[k]);
    }
    int tot_a, tot_r;
    const int pa = block_scan(live && ok, warp_sums, tot_a);
    const int pr = block_scan(live && !ok, warp_sums, tot_r);
    if (live && ok && n_acc + pa < n_needed) {
#pragma unroll
      for (int k = 0; k < SW; ++k) out[(n_acc + pa) * SW + k] = st.w[k];
    }
    if (live && !ok && n_rej + pr < n_needed) {
#pragma unroll
      for (int k = 0; k < SW; ++k) rej_scratch[(n_rej + pr) * SW + k] = st.w[k];
    }
    n_acc += tot_a; n_rej += tot_r;
    attempts += CTA;
  }
  if (n_acc > n_needed) n_acc = n_needed;
  if (attempts > max_attempts) attempts = max_attempts;
  __syncthreads();
  // pad with the first rejected samples (use_rejected_samples=True, mcts.py:696)
  const long long missing = n_needed - n_acc;
  const long long avail = n_rej < missing ? n_rej : missing;
  for (long long j = threadIdx.x; j < avail; j += CTA)
#pragma unroll
    for (int k = 0; k < SW; ++k) out[(n_acc + j) * SW + k] = rej_scratch[j * SW + k];
  if (threadIdx.x == 0) { result[0] = n_acc; result[1] = attempts; result[2] = n_acc + avail; }
}

// rejection sampling of the initial belief: s0 ~ b0 until `n_needed` match the first obs
template <class M>
__device__ void initial_belief(const typename M::Blob& B, int ego, const unsigned long long* __restrict__ target_obs,
                               long long n_needed, long long max_attempts, uint64_t seed, uint64_t draw_id,
                               uint32_t* __restrict__ out, long long* __restrict__ result) {
  constexpr int OW = M::OW, SW = M::SW;
  __shared__ int warp_sums[CTA / 32];
  unsigned long long tgt[OW];
#pragma unroll
  for (int k = 0; k < OW; ++k) tgt[k] = target_obs[k];
  long long n_acc = 0, attempts = 0;
  while (n_acc < n_needed && attempts < max_attempts) {
    const long long att = attempts + threadIdx.x;
    typename M::State st;
    M::initial_state(B, seed, draw_id + (uint64_t)att, st);
    unsigned long long obs[OW];
    M::observe(B, st, ego, obs);
    bool ok = att < max_attempts;
#pragma unroll
    for (int k = 0; k < OW; ++k) ok &= (obs[k] == tgt[k]);
    int tot;
    const int p = block_scan(ok, warp_sums, tot);
    if (ok && n_acc + p < n_needed) {
#pragma unroll
      for (int k = 0; k < SW; ++k) out[(n_acc + p) * SW + k] = st.w[k];
    }
    n_acc += tot;
    attempts += CTA;
  }
  if (threadIdx.x == 0) { result[0] = n_acc < n_needed ? n_acc : n_needed; result[1] = attempts; result[2] = result[0]; }
}

}  // namespace belief
}  // namespace posg
