// Driving-v1 kernels + C-ABI.  Persistent CTAs (grid = k x SM count) stage the model
// blob (wall bitmap, distance / move / observation base / reward tables; ~40 KB)
// into shared memory once via one TMA bulk copy and then grid-stride over the batch
// with coalesced 16-byte loads and stores.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <queue>
#include <string>
#include <vector>

#include "common.h"
#include "driving.cuh"
#include "stage.cuh"

namespace posg {
namespace driving {

constexpr int THREADS = 256;

// ------------------------------------------------------------------ step (model.step)
// Each thread owns EPT = 4 consecutive envs (one Philox block serves the four), loads
// their states with two 16-byte loads and writes every output with 16-byte stores
// (obs, rewards: one 16-byte store per env; states: two per thread; flags: one 4-byte
// store; actions: one 8-byte store) -- every warp-level access is fully coalesced.
constexpr int EPT = 4;
#ifndef STEP_MIN_CTAS
#define STEP_MIN_CTAS 3  // 72 registers, no spills: measured best (3: 51.9 %, 4: 48 %, 5: 39 % of HBM roofline)
#endif

struct EnvOut { uint32_t v0, v1; uint64_t o0, o1; double r0, r1; uint32_t flags; };

__device__ __forceinline__ EnvOut step_env(const Blob& B, uint32_t v0, uint32_t v1, int a0, int a1, int order_idx) {
  const uint32_t b0 = v0, b1 = v1;
  transition_raw(B, v0, v1, a0, a1, order_idx);
  EnvOut e;
  e.r0 = reward_raw(B, b0, v0);
  e.r1 = reward_raw(B, b1, v1);
  v0 &= STATE_MASK; v1 &= STATE_MASK;
  e.v0 = v0; e.v1 = v1;
  observe_pair(B, v0, v1, e.o0, e.o1);
  const uint32_t t0 = vdone(v0) ? 1u : 0u, t1 = vdone(v1) ? 1u : 0u;
  e.flags = t0 | (t1 << 1) | ((t0 & t1) ? POSG_FLAG_ALL_DONE : 0u);
  return e;
}

__global__ void __launch_bounds__(THREADS, STEP_MIN_CTAS)
step_kernel(const Blob* __restrict__ gblob, uint32_t blob_size, int64_t n,
            const uint32_t* __restrict__ state_in, const uint8_t* __restrict__ actions_in,
            const uint8_t* __restrict__ order_in, uint64_t seed, uint32_t step_idx,
            uint64_t particle_offset, uint32_t* __restrict__ state_out, uint64_t* __restrict__ obs_out,
            double* __restrict__ rewards_out, uint8_t* __restrict__ flags_out,
            uint8_t* __restrict__ actions_out, unsigned long long* __restrict__ live_count) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ alignas(8) uint64_t mbar;
  stage_blob(smem_raw, gblob, blob_size, &mbar);
  const Blob& B = *reinterpret_cast<const Blob*>(smem_raw);

  unsigned live = 0;
  const int64_t ngroups = (n + EPT - 1) / EPT;
  const int64_t stride = (int64_t)gridDim.x * THREADS;
  // the fast path needs 16-byte aligned group accesses: particle_offset only shifts the RNG
  for (int64_t g = (int64_t)blockIdx.x * THREADS + threadIdx.x; g < ngroups; g += stride) {
    const int64_t i0 = g * EPT;
    const bool full = (i0 + EPT <= n);
    uint32_t sv[2 * EPT];
    if (full) {
      const uint4 s01 = __ldg(reinterpret_cast<const uint4*>(state_in) + 2 * g);
      const uint4 s23 = __ldg(reinterpret_cast<const uint4*>(state_in) + 2 * g + 1);
      sv[0] = s01.x; sv[1] = s01.y; sv[2] = s01.z; sv[3] = s01.w;
      sv[4] = s23.x; sv[5] = s23.y; sv[6] = s23.z; sv[7] = s23.w;
    } else {
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        const bool ok = i0 + j < n;
        sv[2 * j] = ok ? __ldg(state_in + 2 * (i0 + j)) : 0u;
        sv[2 * j + 1] = ok ? __ldg(state_in + 2 * (i0 + j) + 1) : 0u;
      }
    }
    int act[2 * EPT], ord[EPT];
    if (actions_in == nullptr || order_in == nullptr) {
      // particle_offset + i0 is a multiple of 4 whenever particle_offset is; handle the general case
      const uint64_t p0 = particle_offset + (uint64_t)i0;
      uint4 blk = philox_block(seed, p0 >> 2, step_idx, STREAM_STEP);
      uint4 blk2 = blk;
      if (p0 & 3ULL) blk2 = philox_block(seed, (p0 >> 2) + 1, step_idx, STREAM_STEP);
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        const uint64_t p = p0 + j;
        const uint32_t w = pick_word(((p >> 2) == (p0 >> 2)) ? blk : blk2, (unsigned)(p & 3ULL));
        decode_step_noise(w, act[2 * j], act[2 * j + 1], ord[j]);
      }
    }
    if (actions_in != nullptr) {
      if (full) {
        const uint2 a = __ldg(reinterpret_cast<const uint2*>(actions_in) + g);
#pragma unroll
        for (int j = 0; j < 4; ++j) { act[j] = (a.x >> (8 * j)) & 255; act[4 + j] = (a.y >> (8 * j)) & 255; }
      } else {
#pragma unroll
        for (int j = 0; j < 2 * EPT; ++j) act[j] = (i0 + j / 2 < n) ? __ldg(actions_in + 2 * i0 + j) : 0;
      }
    }
    if (order_in != nullptr) {
      if (full) {
        const uint32_t o = __ldg(reinterpret_cast<const uint32_t*>(order_in) + g);
#pragma unroll
        for (int j = 0; j < EPT; ++j) ord[j] = (o >> (8 * j)) & 255;
      } else {
#pragma unroll
        for (int j = 0; j < EPT; ++j) ord[j] = (i0 + j < n) ? __ldg(order_in + i0 + j) : 0;
      }
    }

    uint32_t so[2 * EPT], fl = 0, a_lo = 0, a_hi = 0;
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
      const bool ok = full || (i0 + j < n);
      live += (ok && !(vdone(sv[2 * j]) && vdone(sv[2 * j + 1]))) ? 1u : 0u;
      const EnvOut e = step_env(B, sv[2 * j], sv[2 * j + 1], act[2 * j], act[2 * j + 1], ord[j]);
      so[2 * j] = e.v0; so[2 * j + 1] = e.v1;
      fl |= e.flags << (8 * j);
      const uint32_t aa = (uint32_t)act[2 * j] | ((uint32_t)act[2 * j + 1] << 8);
      if (j < 2) a_lo |= aa << (16 * j); else a_hi |= aa << (16 * (j - 2));
      if (ok) {
        reinterpret_cast<ulonglong2*>(obs_out)[i0 + j] = make_ulonglong2(e.o0, e.o1);
        reinterpret_cast<double2*>(rewards_out)[i0 + j] = make_double2(e.r0, e.r1);
      }
    }
    if (full) {
      reinterpret_cast<uint4*>(state_out)[2 * g] = make_uint4(so[0], so[1], so[2], so[3]);
      reinterpret_cast<uint4*>(state_out)[2 * g + 1] = make_uint4(so[4], so[5], so[6], so[7]);
      reinterpret_cast<uint32_t*>(flags_out)[g] = fl;
      if (actions_out) reinterpret_cast<uint2*>(actions_out)[g] = make_uint2(a_lo, a_hi);
    } else {
#pragma unroll
      for (int j = 0; j < EPT; ++j) {
        if (i0 + j < n) {
          state_out[2 * (i0 + j)] = so[2 * j]; state_out[2 * (i0 + j) + 1] = so[2 * j + 1];
          flags_out[i0 + j] = (uint8_t)(fl >> (8 * j));
          if (actions_out) {
            actions_out[2 * (i0 + j)] = (uint8_t)act[2 * j];
            actions_out[2 * (i0 + j) + 1] = (uint8_t)act[2 * j + 1];
          }
        }
      }
    }
  }
  if (live_count != nullptr) {
    for (int o = 16; o > 0; o >>= 1) live += __shfl_down_sync(0xffffffffu, live, o);
    if ((threadIdx.x & 31) == 0 && live) atomicAdd(live_count, (unsigned long long)live);
  }
}

// Fast path of model.step for the random-action sweep (BASELINE.json configs[1]): actions and
// processing order from Philox, particle_offset a multiple of four, whole blocks of 128 envs.
// A warp owns 128 consecutive envs; lane L steps envs 32 j + L (j = 0..3), so every warp-level
// global access is one contiguous run (state 256 B, observations / rewards 512 B per
// instruction: the minimum number of L1 wavefronts).  The step noise keeps the convention of
// the general kernel -- Philox block (particle >> 2), word (particle & 3): lane L computes the
// block of envs 4 L .. 4 L + 3 and the words are transposed through 512 B of shared memory per
// warp.  Same results as step_kernel (tests/test_driving_gpu.py).
constexpr int WARP_ENVS = 128;

// Software pipeline (A/B switch): the next block's states are loaded and its Philox rounds
// issued one iteration ahead, so neither the DRAM latency nor the ten dependent multiply rounds
// sit in front of the first env of a block.  Measured share of the HBM roofline, CTAs/SM:
// without the pipeline 3: 71.7 %, 4: 72.4 %, 5: 69.5 %; with it 3: 76.9 % (80 registers), 4: 72.8 %.
#ifndef STEP_FAST_PREFETCH
#define STEP_FAST_PREFETCH 1
#endif
#ifndef STEP_FAST_MIN_CTAS
#define STEP_FAST_MIN_CTAS 3
#endif

// REWARD_CODE: instead of the two float64 rewards (16 B) the kernel writes ONE byte per env to
// `reward_code_out` -- per vehicle the progress made (2 bits) and the reward kind (2 bits), vehicle 1 in the
// high nibble: the index of reward_tab is that code plus init_dest_dist, which the state word carries, so the
// host decodes the exact double (posg_driving_decode_rewards_host).  The compact result of the host-buffer
// sweep: 26 instead of 41 bytes per env cross PCIe.
template <bool WRITE_ACTIONS, bool COUNT_LIVE, bool REWARD_CODE = false>
__global__ void __launch_bounds__(THREADS, STEP_FAST_MIN_CTAS)
step_fast_kernel(const Blob* __restrict__ gblob, uint32_t blob_size, uint32_t nblocks,
                 const uint2* __restrict__ state_in, const __grid_constant__ PhiloxKeys seed, uint32_t step_idx,
                 uint64_t group_offset,
                 uint2* __restrict__ state_out, ulonglong2* __restrict__ obs_out, double2* __restrict__ rewards_out,
                 uint8_t* __restrict__ flags_out, uchar2* __restrict__ actions_out,
                 unsigned long long* __restrict__ live_count, uint8_t* __restrict__ reward_code_out = nullptr) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ alignas(8) uint64_t mbar;
  __shared__ alignas(16) uint32_t s_words[THREADS / 32][WARP_ENVS];
  stage_blob(smem_raw, gblob, blob_size, &mbar);
  const Blob& B = *reinterpret_cast<const Blob*>(smem_raw);
  unsigned live = 0;
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t wstride = gridDim.x * (THREADS / 32);
  uint32_t wb = blockIdx.x * (THREADS / 32) + warp;
#if STEP_FAST_PREFETCH
  // software pipeline: the next block's states are requested before this block is computed, so
  // the DRAM latency of the only global loads of the loop is hidden behind ~1000 instructions
  uint2 nxt[EPT];
#pragma unroll
  for (int j = 0; j < EPT; ++j) nxt[j] = wb < nblocks ? __ldg(state_in + wb * WARP_ENVS + lane + 32 * j) : make_uint2(0u, 0u);
#endif
#if STEP_FAST_PREFETCH
  uint4 blk_next = philox_block(seed, group_offset + (uint64_t)wb * 32u + lane, step_idx, STREAM_STEP);
#endif
  for (; wb < nblocks; wb += wstride) {
    const uint32_t e0 = wb * WARP_ENVS + lane;     // this lane's first env
    uint2 sv[EPT];
#if STEP_FAST_PREFETCH
#pragma unroll
    for (int j = 0; j < EPT; ++j) sv[j] = nxt[j];
    if (wb + wstride < nblocks) {
#pragma unroll
      for (int j = 0; j < EPT; ++j) nxt[j] = __ldg(state_in + (wb + wstride) * WARP_ENVS + lane + 32 * j);
    }
#else
#pragma unroll
    for (int j = 0; j < EPT; ++j) sv[j] = __ldg(state_in + e0 + 32 * j);
#endif
#if STEP_FAST_PREFETCH
    __syncwarp();
    *reinterpret_cast<uint4*>(&s_words[warp][4 * lane]) = blk_next;
    __syncwarp();
    uint32_t nzw[EPT];
#pragma unroll
    for (int j = 0; j < EPT; ++j) nzw[j] = s_words[warp][32 * j + lane];
    // the next block's Philox rounds are independent of everything below: the scheduler overlaps them
    blk_next = philox_block(seed, group_offset + (uint64_t)(wb + wstride) * 32u + lane, step_idx, STREAM_STEP);
#else
    const uint4 blk = philox_block(seed, group_offset + (uint64_t)wb * 32u + lane, step_idx, STREAM_STEP);
    __syncwarp();
    *reinterpret_cast<uint4*>(&s_words[warp][4 * lane]) = blk;
    __syncwarp();
    uint32_t nzw[EPT];
#pragma unroll
    for (int j = 0; j < EPT; ++j) nzw[j] = s_words[warp][32 * j + lane];
#endif
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
      const uint32_t nz = B.noise_lut[__umulhi(nzw[j], 50u)];
      uint32_t v0 = sv[j].x, v1 = sv[j].y;
      const uint32_t b0 = v0, b1 = v1;
      if (COUNT_LIVE) live += ((b0 & 0x18000u) && (b1 & 0x18000u)) ? 0u : 1u;
      transition16(B, v0, v1, (nz << 4) & 0x70u, (nz << 1) & 0x70u, (nz & 0x40u) != 0u);
      const uint32_t e = e0 + 32 * j;
      if constexpr (REWARD_CODE)
        reward_code_out[e] = (uint8_t)(reward_code(b0, v0) | (reward_code(b1, v1) << 4));
      else
        rewards_out[e] = make_double2(reward_raw(B, b0, v0), reward_raw(B, b1, v1));
      v0 &= STATE_MASK; v1 &= STATE_MASK;
      state_out[e] = make_uint2(v0, v1);
      const uint32_t t0 = (v0 & 0x18000u) ? 1u : 0u, t1 = (v1 & 0x18000u) ? 1u : 0u;
      flags_out[e] = (uint8_t)(t0 | (t1 << 1) | ((t0 & t1) ? POSG_FLAG_ALL_DONE : 0u));
      if (WRITE_ACTIONS) actions_out[e] = make_uchar2((unsigned char)((nz >> 7) & 7u), (unsigned char)((nz >> 10) & 7u));
      uint64_t o0, o1;
      observe_pair(B, v0, v1, o0, o1);
      obs_out[e] = make_ulonglong2(o0, o1);
    }
  }
  if (COUNT_LIVE) {
    for (int o = 16; o > 0; o >>= 1) live += __shfl_down_sync(0xffffffffu, live, o);
    if (lane == 0 && live) atomicAdd(live_count, (unsigned long long)live);
  }
}

// ------------------------------------------------------------------ observe / initial state
__global__ void __launch_bounds__(THREADS)
observe_kernel(const Blob* __restrict__ gblob, uint32_t blob_size, int64_t n,
               const uint32_t* __restrict__ state, uint64_t* __restrict__ obs_out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ alignas(8) uint64_t mbar;
  stage_blob(smem_raw, gblob, blob_size, &mbar);
  const Blob& B = *reinterpret_cast<const Blob*>(smem_raw);
  const int64_t stride = (int64_t)gridDim.x * THREADS;
  for (int64_t i = (int64_t)blockIdx.x * THREADS + threadIdx.x; i < n; i += stride) {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(state) + i);
    uint64_t o0, o1;
    observe_pair(B, v.x, v.y, o0, o1);
    reinterpret_cast<ulonglong2*>(obs_out)[i] = make_ulonglong2(o0, o1);
  }
}

__global__ void __launch_bounds__(THREADS)
initial_state_kernel(const Blob* __restrict__ gblob, uint32_t blob_size, int64_t n,
                     const uint8_t* __restrict__ picks, uint64_t seed, uint64_t particle_offset,
                     uint32_t* __restrict__ state_out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ alignas(8) uint64_t mbar;
  stage_blob(smem_raw, gblob, blob_size, &mbar);
  const Blob& B = *reinterpret_cast<const Blob*>(smem_raw);
  const int64_t stride = (int64_t)gridDim.x * THREADS;
  for (int64_t i = (int64_t)blockIdx.x * THREADS + threadIdx.x; i < n; i += stride) {
    uint32_t veh[NA];
    if (picks != nullptr) {
      const uint32_t pk = __ldg(reinterpret_cast<const uint32_t*>(picks) + i);
      int p[2 * NA];
#pragma unroll
      for (int k = 0; k < 2 * NA; ++k) p[k] = (pk >> (8 * k)) & 255;
      initial_state(B, p, veh);
    } else {
      initial_state_random(B, seed, particle_offset + (uint64_t)i, veh);
    }
    reinterpret_cast<uint2*>(state_out)[i] = make_uint2(veh[0], veh[1]);
  }
}

// ------------------------------------------------------------------ fused rollout
// MCTS._rollout with uniform-random policies for every agent.  Each lane owns one
// rollout at a time; when lanes of a warp finish (ego terminated / all done / horizon)
// the warp refills them from a global work counter with one ballot + one atomic
// (warp-level compaction of terminated rollouts), so long rollouts never leave the
// other 31 lanes idle.
// discount**k for k < 256, evaluated on the host with libm pow() -- the same function
// CPython's ``discount**rollout_t`` (mcts.py:421) calls -- so returns match bit for bit.
struct DiscountTable { double v[256]; };

__global__ void __launch_bounds__(THREADS)
rollout_kernel(const Blob* __restrict__ gblob, uint32_t blob_size, int64_t n,
               const uint32_t* __restrict__ state_in, int ego, int max_steps,
               const __grid_constant__ DiscountTable dtab, const __grid_constant__ PhiloxKeys seed, uint32_t step0,
               uint64_t particle_offset, double* __restrict__ returns_out,
               uint8_t* __restrict__ steps_out, uint32_t* __restrict__ state_out,
               unsigned long long* __restrict__ work_counter,
               unsigned long long* __restrict__ live_count) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ alignas(8) uint64_t mbar;
  __shared__ double sdisc[256];
  for (int t = threadIdx.x; t < 256; t += THREADS) sdisc[t] = dtab.v[t];
  stage_blob(smem_raw, gblob, blob_size, &mbar);  // contains a __syncthreads()
  const Blob& B = *reinterpret_cast<const Blob*>(smem_raw);

  const unsigned lane = threadIdx.x & 31u;
  uint32_t v0 = 0, v1 = 0;
  int64_t item = -1;       // rollout owned by this lane (-1: none)
  int k = 0;               // steps taken
  double ret = 0.0;
  unsigned long long executed = 0;
  bool exhausted = false;  // the work counter ran past n

  while (true) {
    // ---- refill idle lanes (warp-synchronous) ----
    const unsigned idle = __ballot_sync(0xffffffffu, item < 0);
    if (idle && !exhausted) {
      unsigned long long base = 0;
      const int cnt = __popc(idle);
      const int leader = __ffs(idle) - 1;
      if ((int)lane == leader) base = atomicAdd(work_counter, (unsigned long long)cnt);
      base = __shfl_sync(0xffffffffu, base, leader);
      if (item < 0) {
        const int64_t mine = (int64_t)base + __popc(idle & ((1u << lane) - 1u));
        if (mine < n) {
          item = mine;
          const uint2 v = __ldg(reinterpret_cast<const uint2*>(state_in) + item);
          v0 = v.x; v1 = v.y;
          k = 0; ret = 0.0;
        }
      }
      if ((int64_t)base + cnt >= n) exhausted = true;
    }
    if (__ballot_sync(0xffffffffu, item >= 0) == 0u) break;

    // ---- one model step for every active lane ----
    if (item >= 0) {
      bool finished = (k >= max_steps);
      if (!finished) {
        int a0, a1, order_idx;
        sample_step_noise(seed, particle_offset + (uint64_t)item, step0 + (uint32_t)k, a0, a1, order_idx);
        const uint32_t before = ego ? v1 : v0;
        transition_raw(B, v0, v1, a0, a1, order_idx);
        const double r = reward_raw(B, before, ego ? v1 : v0);
        v0 &= STATE_MASK; v1 &= STATE_MASK;
        const uint32_t after = ego ? v1 : v0;
        ret = __dadd_rn(ret, __dmul_rn(sdisc[k], r));
        ++k; ++executed;
        finished = vdone(after) || (vdone(v0) && vdone(v1)) || (k >= max_steps);
      }
      if (finished) {
        returns_out[item] = ret;
        if (steps_out) steps_out[item] = (uint8_t)k;
        if (state_out) reinterpret_cast<uint2*>(state_out)[item] = make_uint2(v0, v1);
        item = -1;
      }
    }
  }
  if (live_count != nullptr) {
    for (int o = 16; o > 0; o >>= 1) executed += __shfl_down_sync(0xffffffffu, executed, o);
    if (lane == 0 && executed) atomicAdd(live_count, executed);
  }
}

// ------------------------------------------------------------------ host: model build
// grid cell seen at window position (row r: 0 = farthest ahead, col c: 0 = left) by a
// vehicle at (x, y) facing `dir`, for the (front, back, side) = (3, 1, 1) window
static void window_to_grid(int x, int y, int dir, int r, int c, int* gx, int* gy) {
  const int FRONT = 3, SIDE = 1;
  switch (dir) {
    case 0: *gy = y + r - FRONT; *gx = x + c - SIDE; break;   // NORTH
    case 1: *gy = y + c - SIDE;  *gx = x - r + FRONT; break;  // EAST
    case 2: *gy = y - r + FRONT; *gx = x - c + SIDE; break;   // SOUTH
    default: *gy = y - c + SIDE; *gx = x + r - FRONT; break;  // WEST
  }
}

static Blob* build_blob(const char* grid_rows, int num_agents, int obs_front, int obs_back,
                        int obs_side, int obstacle_collisions, size_t* bytes_out) {
  std::vector<std::string> rows;
  {
    std::string cur;
    for (const char* p = grid_rows; *p; ++p) {
      if (*p == '\n') { if (!cur.empty()) rows.push_back(cur); cur.clear(); }
      else if (*p != '\r' && *p != ' ') cur.push_back(*p);
    }
    if (!cur.empty()) rows.push_back(cur);
  }
  const int H = (int)rows.size();
  if (H < 1 || H > 16) { set_error("grid height %d outside [1,16]", H); return nullptr; }
  const int W = (int)rows[0].size();
  if (W < 1 || W > 16) { set_error("grid width %d outside [1,16]", W); return nullptr; }
  for (auto& r : rows) if ((int)r.size() != W) { set_error("ragged grid rows"); return nullptr; }

  // first pass: count destinations (sizes the blob)
  int n_dest_total = 0;
  for (auto& r : rows) for (char ch : r) n_dest_total += (ch == '-' || (ch >= 'a' && ch <= 'j'));
  if (n_dest_total > MAX_CELLS) { set_error("more than %d destination cells", MAX_CELLS); return nullptr; }
  const size_t bytes = blob_bytes(n_dest_total);
  Blob* out = (Blob*)calloc(1, bytes);
  for (int l = 0; l < LINES; ++l) out->rows[l] = 0xffffffffu;
  bool blk[16][16];
  memset(blk, 1, sizeof(blk));
  int n_start = 0, n_dest = 0;
  for (int y = 0; y < H; ++y) {
    for (int x = 0; x < W; ++x) {
      const char ch = rows[y][x];
      blk[y][x] = (ch == '#');
      if (!blk[y][x]) out->rows[y + PAD] &= ~(1u << (x + PAD));
      uint8_t smask = 0, dmask = 0;
      if (ch == '+') smask = 0x0f;
      else if (ch >= '0' && ch <= '9') smask = (uint8_t)(((ch - '0') < 4) ? (1u << (ch - '0')) : 0x80u);
      else if (ch == '-') dmask = 0x0f;
      else if (ch >= 'a' && ch <= 'j') dmask = (uint8_t)(((ch - 'a') < 4) ? (1u << (ch - 'a')) : 0x80u);
      else if (ch != '.' && ch != '#') { set_error("bad grid character '%c'", ch); free(out); return nullptr; }
      if (smask) {
        if (n_start >= MAX_CELLS) { set_error("more than %d start cells", MAX_CELLS); free(out); return nullptr; }
        out->start_xy[n_start] = (uint8_t)(x | (y << 4)); out->start_agents[n_start++] = smask;
      }
      if (dmask) { out->dest_xy[n_dest] = (uint8_t)(x | (y << 4)); out->dest_agents[n_dest++] = dmask; }
    }
  }
  for (int i = 0; i < num_agents; ++i) {
    int ns = 0, nd = 0;
    for (int c = 0; c < n_start; ++c) ns += (out->start_agents[c] >> i) & 1;
    for (int c = 0; c < n_dest; ++c) nd += (out->dest_agents[c] >> i) & 1;
    if (ns < 1 || nd < 1 || n_start < num_agents || n_dest < num_agents) {
      set_error("grid does not support %d agents", num_agents); free(out); return nullptr;
    }
  }
  out->n_start = (uint8_t)n_start; out->n_dest = (uint8_t)n_dest;
  out->width = (uint8_t)W; out->height = (uint8_t)H;
  out->n_agents = (uint8_t)num_agents;
  out->obs_front = (uint8_t)obs_front; out->obs_back = (uint8_t)obs_back; out->obs_side = (uint8_t)obs_side;
  out->obstacle_collisions = obstacle_collisions ? 1 : 0;
  // unit-cost shortest paths from every destination (breadth-first)
  for (int d = 0; d < n_dest; ++d) {
    memset(out->dist[d], 255, 256);
    std::queue<int> q;
    const int o = out->dest_xy[d];
    out->dist[d][o] = 0; q.push(o);
    while (!q.empty()) {
      const int c = q.front(); q.pop();
      const int cx = c & 15, cy = c >> 4;
      const int nb[4][2] = {{cx, cy - 1}, {cx + 1, cy}, {cx, cy + 1}, {cx - 1, cy}};
      for (auto& nbr : nb) {
        const int nx = nbr[0], ny = nbr[1];
        if (nx < 0 || ny < 0 || nx >= W || ny >= H || blk[ny][nx]) continue;
        const int k = nx | (ny << 4);
        if (out->dist[d][k] != 255) continue;
        if (out->dist[d][c] + 1 >= MAX_D0) { set_error("shortest path longer than %d", MAX_D0 - 1); free(out); return nullptr; }
        out->dist[d][k] = (uint8_t)(out->dist[d][c] + 1); q.push(k);
      }
    }
  }
  for (int d0 = 0; d0 < MAX_D0; ++d0)
    for (int p = 0; p < 4; ++p) {
      // r = base; r += R_PROGRESS * (progress / D0)   (base: step cost 0, destination 0.5, crash -1)
      const double prog = 0.5 * ((double)(p > 2 ? 2 : p) / (double)(d0 < 1 ? 1 : d0));
      double* r = out->reward_tab + (d0 | (p << 6)) * 4;
      r[0] = 0.0 + prog; r[1] = 0.5 + prog; r[2] = -1.0 + prog; r[3] = 0.0;
    }
  // control table: (action, speed, facing) -> next speed, next facing, move direction, cells moved
  for (int a = 0; a < 5; ++a)
    for (int sp = 0; sp < 4; ++sp)
      for (int dir = 0; dir < 4; ++dir) {
        int nsp = sp;
        if (a == ACCELERATE) nsp = sp + 1 > FORWARD_FAST ? (int)FORWARD_FAST : sp + 1;
        else if (a == DECELERATE) nsp = sp - 1 < REVERSE ? (int)REVERSE : sp - 1;
        else if ((a == TURN_RIGHT || a == TURN_LEFT) && sp == FORWARD_FAST) nsp = FORWARD_SLOW;
        int tdir = dir;
        if (a == TURN_RIGHT) tdir = (dir + 1) & 3;
        else if (a == TURN_LEFT) tdir = (dir + 3) & 3;
        const bool rev = nsp == REVERSE;
        const int mdir = rev ? ((dir + 2) & 3) : tdir;
        const int ndir = rev ? dir : tdir;   // a reversing vehicle keeps its facing
        const int nmoves = rev ? 1 : nsp - STOPPED;
        out->ctrl2[a * 16 + sp * 4 + dir] = (uint8_t)(ndir | (nsp << 2) | ((mdir * 3 + nmoves) << 4));
      }
  // move table: up to two cells in a direction, stopping in front of walls / the grid edge
  for (int dir = 0; dir < 4; ++dir)
    for (int nmoves = 0; nmoves < 3; ++nmoves)
      for (int y = 0; y < 16; ++y)
        for (int x = 0; x < 16; ++x) {
          const int dx = (dir == 1) - (dir == 3), dy = (dir == 2) - (dir == 0);
          auto wall = [&](int cx, int cy) { return cx < 0 || cy < 0 || cx >= W || cy >= H || blk[cy][cx]; };
          const uint32_t c0 = (uint32_t)(x | (y << 4));
          const int x1 = x + dx, y1 = y + dy, x2 = x1 + dx, y2 = y1 + dy;
          const bool try1 = nmoves >= 1, b1 = wall(x1, y1);
          const bool s1 = try1 && !b1, o1 = try1 && b1;
          const bool try2 = s1 && nmoves == 2, b2 = wall(x2, y2);
          const bool s2 = try2 && !b2, o2 = try2 && b2;
          const uint32_t c1 = s1 ? (uint32_t)(x1 | (y1 << 4)) : c0;
          const uint32_t cf = s2 ? (uint32_t)(x2 | (y2 << 4)) : c1;
          const uint32_t oc = obstacle_collisions ? 1u : 0u;
          out->move2[((dir * 3 + nmoves) << 8) | (y << 4) | x] =
              cf | (c1 << 8) | ((s1 ? 1u : 0u) << 16) | ((s2 ? 1u : 0u) << 17) | ((o1 ? oc : 0u) << 18) |
              ((o2 ? oc : 0u) << 19);
        }
  // step noise: the 50 (order, a0, a1) combinations (decode_step_noise)
  for (uint32_t idx = 0; idx < 56; ++idx) {
    const uint32_t i = idx < 50 ? idx : 49;
    const uint32_t sw = i & 1u, r = i >> 1, a1 = r / 5u, a0 = r - 5u * a1;
    const uint32_t af = sw ? a1 : a0, as = sw ? a0 : a1;
    out->noise_lut[idx] = (uint16_t)(af | (as << 3) | (sw << 6) | (a0 << 7) | (a1 << 10));
  }
  // window lookup: where inside the 5x3 window does the grid offset (dx, dy) fall?
  memset(out->win_lut, 15, sizeof(out->win_lut));
  for (int dir = 0; dir < 4; ++dir)
    for (int r = 0; r < 5; ++r)
      for (int c = 0; c < 3; ++c) {
        int gx, gy;
        window_to_grid(0, 0, dir, r, c, &gx, &gy);
        out->win_lut[dir][(gy + 3) * 8 + (gx + 3)] = (uint8_t)(r * 3 + c);
      }
  // local-cell codes of a lone vehicle: walls / own destination / ego cell
  for (int d = 0; d < n_dest; ++d)
    for (int dir = 0; dir < 4; ++dir)
      for (int y = 0; y < 16; ++y)
        for (int x = 0; x < 16; ++x) {
          uint32_t cells = 0;
          for (int r = 0; r < 5; ++r)
            for (int c = 0; c < 3; ++c) {
              int gx, gy;
              window_to_grid(x, y, dir, r, c, &gx, &gy);
              uint32_t code;
              if (gx < 0 || gy < 0 || gx >= W || gy >= H || blk[gy][gx]) code = WALL;
              else if (gx == x && gy == y) code = VEHICLE;
              else if ((gx | (gy << 4)) == out->dest_xy[d]) code = DESTINATION;
              else code = EMPTY;
              cells |= code << (2 * (r * 3 + c));
            }
          out->obs_base[((d * 4 + dir) << 8) | (y << 4) | x] = cells;
        }
  *bytes_out = bytes;
  return out;
}

// persistent grid: every resident CTA slot of every SM gets exactly one CTA
template <typename K>
static int grid_for(K kernel, size_t smem, int64_t work_items) {
  smem_optin(kernel, (int)smem);
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, THREADS, smem) != cudaSuccess || per_sm < 1) {
    cudaGetLastError();
    per_sm = 1;
  }
  const int64_t need = (work_items + THREADS - 1) / THREADS;
  const int64_t cap = (int64_t)sm_count() * per_sm;
  return (int)(need < cap ? (need < 1 ? 1 : need) : cap);
}

}  // namespace driving
}  // namespace posg

using namespace posg;
using namespace posg::driving;

// posg_driving_step restricted to whole fast-kernel blocks with the reward CODE as output (compact sweep)
static int step_fast_coded(posg_model_t m, int64_t n, const uint32_t* state_in, uint64_t seed, uint32_t step_idx,
                           uint64_t particle_offset, uint32_t* state_out, uint64_t* obs_out, uint8_t* code_out,
                           uint8_t* flags_out, uint8_t* actions_out, unsigned long long* live_count, cudaStream_t s) {
  const uint32_t nb = (uint32_t)(n / WARP_ENVS);
  step_fast_kernel<true, true, true><<<grid_for(step_fast_kernel<true, true, true>, m->blob_bytes, (int64_t)nb * 32), THREADS,
                                        m->blob_bytes, s>>>(
      (const Blob*)m->d_blob, (uint32_t)m->blob_bytes, nb, (const uint2*)state_in, PhiloxKeys(seed), step_idx,
      particle_offset >> 2, (uint2*)state_out, (ulonglong2*)obs_out, nullptr, flags_out, (uchar2*)actions_out, live_count,
      code_out);
  POSG_CUDA(cudaGetLastError());
  return POSG_OK;
}

// staging context of posg_driving_sweep_host (see posg_model::host_stage)
struct HostStage {
  static constexpr int NSLOT = 3;
  std::mutex mu;
  unsigned char* d_base[NSLOT] = {nullptr, nullptr, nullptr};
  size_t bytes[NSLOT] = {0, 0, 0};
  cudaStream_t st[NSLOT] = {nullptr, nullptr, nullptr};
  unsigned long long* d_live = nullptr;
};
static void host_stage_free(void* p) {
  HostStage* hs = (HostStage*)p;
  for (int i = 0; i < HostStage::NSLOT; ++i) {
    if (hs->st[i]) { cudaStreamSynchronize(hs->st[i]); cudaStreamDestroy(hs->st[i]); }
    cudaFree(hs->d_base[i]);
  }
  cudaFree(hs->d_live);
  delete hs;
}
static HostStage* host_stage_of(posg_model* m) {
  static std::mutex create_mu;
  std::lock_guard<std::mutex> lock(create_mu);
  if (!m->host_stage) { m->host_stage = new HostStage(); m->host_stage_free = host_stage_free; }
  return (HostStage*)m->host_stage;
}

static int check_driving(posg_model_t m) {
  if (!m || m->kind != KIND_DRIVING) { set_error("not a Driving model handle"); return POSG_E_INVALID; }
  return POSG_OK;
}

extern "C" {

int posg_driving_create(const char* grid_rows, int num_agents, int obs_front, int obs_back,
                        int obs_side, int obstacle_collisions, posg_model_t* out) {
  POSG_REQUIRE(grid_rows && out, "null argument");
  if (!(obs_front == 3 && obs_back == 1 && obs_side == 1)) {
    set_error("obs_dim (%d,%d,%d) unsupported: kernels are specialised for (3,1,1)", obs_front, obs_back, obs_side);
    return POSG_E_UNSUPPORTED;
  }
  int ndev = posg_device_count();
  if (ndev < 0) return ndev;
  async_pool_init();
  if (num_agents != NA) {
    set_error("num_agents %d unsupported: the sm_100a kernels are specialised for 2 vehicles", num_agents);
    return POSG_E_UNSUPPORTED;
  }
  size_t bytes = 0;
  Blob* h = build_blob(grid_rows, num_agents, obs_front, obs_back, obs_side, obstacle_collisions, &bytes);
  if (!h) return POSG_E_INVALID;
  posg_model* m = new posg_model();
  m->kind = KIND_DRIVING; m->num_agents = num_agents; m->blob_bytes = (int)bytes;
  m->h_blob = h; m->d_blob = nullptr;
  cudaError_t e = cudaGetDevice(&m->device);
  if (e == cudaSuccess) e = cudaMalloc(&m->d_blob, bytes);
  if (e == cudaSuccess) e = cudaMemcpy(m->d_blob, h, bytes, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { int rc = cuda_fail(e, "model upload"); posg_model_destroy(m); return rc; }
  *out = m;
  return POSG_OK;
}

int posg_driving_sample_initial_state(posg_model_t m, int64_t n, const uint8_t* picks, uint64_t seed,
                                      uint64_t particle_offset, uint32_t* state_out, void* stream) {
  int rc = check_driving(m); if (rc) return rc;
  if (n == 0) return POSG_OK;
  POSG_REQUIRE(n > 0 && state_out, "bad arguments");
  initial_state_kernel<<<grid_for(initial_state_kernel, m->blob_bytes, n), THREADS, m->blob_bytes, (cudaStream_t)stream>>>(
      (const Blob*)m->d_blob, (uint32_t)m->blob_bytes, n, picks, seed, particle_offset, state_out);
  POSG_CUDA(cudaGetLastError());
  return POSG_OK;
}

int posg_driving_observe(posg_model_t m, int64_t n, const uint32_t* state, uint64_t* obs_out, void* stream) {
  int rc = check_driving(m); if (rc) return rc;
  if (n == 0) return POSG_OK;
  POSG_REQUIRE(n > 0 && state && obs_out, "bad arguments");
  observe_kernel<<<grid_for(observe_kernel, m->blob_bytes, n), THREADS, m->blob_bytes, (cudaStream_t)stream>>>(
      (const Blob*)m->d_blob, (uint32_t)m->blob_bytes, n, state, obs_out);
  POSG_CUDA(cudaGetLastError());
  return POSG_OK;
}

int posg_driving_step(posg_model_t m, int64_t n, const uint32_t* state_in, const uint8_t* actions_in,
                      const uint8_t* order_in, uint64_t seed, uint32_t step_idx, uint64_t particle_offset,
                      uint32_t* state_out, uint64_t* obs_out, double* rewards_out, uint8_t* flags_out,
                      uint8_t* actions_out, unsigned long long* live_count, void* stream) {
  int rc = check_driving(m); if (rc) return rc;
  if (n == 0) return POSG_OK;
  POSG_REQUIRE(n > 0 && state_in && state_out && obs_out && rewards_out && flags_out, "bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  const Blob* blob = (const Blob*)m->d_blob;
  const uint32_t bb = (uint32_t)m->blob_bytes;
  const bool aligned = (((uintptr_t)state_in | (uintptr_t)state_out | (uintptr_t)obs_out | (uintptr_t)rewards_out) & 15u) == 0 &&
                       ((uintptr_t)flags_out & 3u) == 0 && ((uintptr_t)actions_out & 7u) == 0;
  int64_t done = 0;
  static const bool nofast = getenv("POSG_STEP_NOFAST") != nullptr;   // dev A/B switch
  if (!nofast && !actions_in && !order_in && (particle_offset & 3ULL) == 0 && n >= WARP_ENVS && n < ((int64_t)1 << 31) && aligned) {
    // whole blocks of 128 envs through the fast kernel; a ragged tail through the general one
    const uint32_t nb = (uint32_t)(n / WARP_ENVS);
    const int64_t items = (int64_t)nb * 32;   // threads needed
#define POSG_STEP_FAST(WA, CL)                                                                                      \
    step_fast_kernel<WA, CL><<<grid_for(step_fast_kernel<WA, CL>, m->blob_bytes, items), THREADS, m->blob_bytes, s>>>( \
        blob, bb, nb, (const uint2*)state_in, PhiloxKeys(seed), step_idx, particle_offset >> 2, (uint2*)state_out,   \
        (ulonglong2*)obs_out, (double2*)rewards_out, flags_out, (uchar2*)actions_out, live_count)
    if (actions_out && live_count) POSG_STEP_FAST(true, true);
    else if (actions_out) POSG_STEP_FAST(true, false);
    else if (live_count) POSG_STEP_FAST(false, true);
    else POSG_STEP_FAST(false, false);
#undef POSG_STEP_FAST
    POSG_CUDA(cudaGetLastError());
    done = (int64_t)nb * WARP_ENVS;
  }
  if (done < n) {
    const int64_t r = n - done;
    step_kernel<<<grid_for(step_kernel, m->blob_bytes, (r + EPT - 1) / EPT), THREADS, m->blob_bytes, s>>>(
        blob, bb, r, state_in + 2 * done, actions_in ? actions_in + 2 * done : nullptr,
        order_in ? order_in + done : nullptr, seed, step_idx, particle_offset + (uint64_t)done, state_out + 2 * done,
        obs_out + 2 * done, rewards_out + 2 * done, flags_out + done, actions_out ? actions_out + 2 * done : nullptr,
        live_count);
  }
  POSG_CUDA(cudaGetLastError());
  return POSG_OK;
}

static int rollout_launch(posg_model_t m, int64_t n, const uint32_t* state_in, int ego, int max_steps,
                          double discount, uint64_t seed, uint32_t step0, uint64_t particle_offset,
                          double* returns_out, uint8_t* steps_out, uint32_t* state_out,
                          unsigned long long* work_counter, unsigned long long* live_count,
                          cudaStream_t stream) {
  DiscountTable dtab;
  for (int k = 0; k < 256; ++k) dtab.v[k] = pow(discount, (double)k);
  rollout_kernel<<<grid_for(rollout_kernel, m->blob_bytes, n), THREADS, m->blob_bytes, stream>>>(
      (const Blob*)m->d_blob, (uint32_t)m->blob_bytes, n, state_in, ego, max_steps, dtab, PhiloxKeys(seed), step0,
      particle_offset, returns_out, steps_out, state_out, work_counter, live_count);
  POSG_CUDA(cudaGetLastError());
  return POSG_OK;
}

int posg_driving_rollout(posg_model_t m, int64_t n, const uint32_t* state_in, int ego_agent, int max_steps,
                         double discount, uint64_t seed, uint32_t step0, uint64_t particle_offset,
                         double* returns_out, uint8_t* steps_out, uint32_t* state_out,
                         unsigned long long* live_count, void* stream) {
  int rc = check_driving(m); if (rc) return rc;
  POSG_REQUIRE(ego_agent >= 0 && ego_agent < m->num_agents, "ego_agent out of range");
  POSG_REQUIRE(max_steps >= 0 && max_steps <= 255, "max_steps outside [0,255]");
  if (n == 0) return POSG_OK;
  POSG_REQUIRE(n > 0 && state_in && returns_out, "bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  // per-call work counter (stream-ordered allocation: no global mutable state)
  unsigned long long* counter = nullptr;
  POSG_CUDA(cudaMallocAsync((void**)&counter, sizeof(unsigned long long), s));
  POSG_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), s));
  rc = rollout_launch(m, n, state_in, ego_agent, max_steps, discount, seed, step0, particle_offset,
                      returns_out, steps_out, state_out, counter, live_count, s);
  cudaFreeAsync(counter, s);
  return rc;
}

int posg_driving_rollout_host(posg_model_t m, int64_t n, const uint32_t* state_in_host, int ego_agent,
                              int max_steps, double discount, uint64_t seed, uint32_t step0,
                              uint64_t particle_offset, double* returns_out_host, uint8_t* steps_out_host,
                              uint32_t* state_out_host, unsigned long long* live_count_host) {
  int rc = check_driving(m); if (rc) return rc;
  if (n == 0) return POSG_OK;
  POSG_REQUIRE(n > 0 && state_in_host && returns_out_host, "bad arguments");
  const size_t na = (size_t)m->num_agents;
  uint32_t *d_in = nullptr, *d_out = nullptr; double* d_ret = nullptr; uint8_t* d_steps = nullptr;
  unsigned long long* d_live = nullptr;
  cudaStream_t s = nullptr;
  POSG_CUDA(cudaMalloc((void**)&d_in, n * na * 4));
  POSG_CUDA(cudaMalloc((void**)&d_ret, n * 8));
  POSG_CUDA(cudaMalloc((void**)&d_live, 8));
  if (steps_out_host) POSG_CUDA(cudaMalloc((void**)&d_steps, n));
  if (state_out_host) POSG_CUDA(cudaMalloc((void**)&d_out, n * na * 4));
  POSG_CUDA(cudaMemcpyAsync(d_in, state_in_host, n * na * 4, cudaMemcpyHostToDevice, s));
  POSG_CUDA(cudaMemsetAsync(d_live, 0, 8, s));
  rc = posg_driving_rollout(m, n, d_in, ego_agent, max_steps, discount, seed, step0, particle_offset,
                            d_ret, d_steps, d_out, d_live, s);
  if (rc == POSG_OK) {
    POSG_CUDA(cudaMemcpyAsync(returns_out_host, d_ret, n * 8, cudaMemcpyDeviceToHost, s));
    if (steps_out_host) POSG_CUDA(cudaMemcpyAsync(steps_out_host, d_steps, n, cudaMemcpyDeviceToHost, s));
    if (state_out_host) POSG_CUDA(cudaMemcpyAsync(state_out_host, d_out, n * na * 4, cudaMemcpyDeviceToHost, s));
    if (live_count_host) POSG_CUDA(cudaMemcpyAsync(live_count_host, d_live, 8, cudaMemcpyDeviceToHost, s));
    POSG_CUDA(cudaStreamSynchronize(s));
  }
  cudaFree(d_in); cudaFree(d_ret); cudaFree(d_live); cudaFree(d_steps); cudaFree(d_out);
  return rc;
}

// Pipelined over CHUNKS slices of the batch on two streams so that, after the first
// slice, the H2D copy of slice k+1, the 50 step launches of slice k and the D2H copy of
// slice k-1 overlap (PCIe is full duplex; the copy engines run beside the SMs).
}  // extern "C"

static int sweep_host_impl(posg_model_t m, int64_t n, const uint32_t* state_in_host, int num_steps,
                           uint64_t seed, uint64_t particle_offset, uint32_t* state_out_host,
                           uint64_t* obs_out_host, double* rewards_out_host, uint8_t* code_out_host,
                           uint8_t* flags_out_host, unsigned long long* live_count_host) {
  int rc = check_driving(m); if (rc) return rc;
  if (n == 0) return POSG_OK;
  POSG_REQUIRE(n > 0 && state_in_host && num_steps >= 1, "bad arguments");
  const bool coded = code_out_host != nullptr;
  POSG_REQUIRE(!coded || ((n % WARP_ENVS) == 0 && (particle_offset & 3ULL) == 0),
               "the compact sweep needs a multiple of 128 envs and a particle offset that is a multiple of 4");
  // Slices are multiples of 128 envs (whole fast-kernel blocks; also keeps the Philox word
  // grouping of the monolithic call).  Three slots: while one slice computes, the previous one
  // drains to the host and the next one is uploaded (PCIe is full duplex).
  HostStage* hs = host_stage_of(m);
  std::lock_guard<std::mutex> lock(hs->mu);
  int64_t chunk = ((n + 15) / 16 + 127) & ~(int64_t)127;
  if (chunk < (1 << 16)) chunk = (n + 127) & ~(int64_t)127;
  const size_t per_env = 8 + 16 + 16 + 1 + 2;
  const size_t slot_bytes = (size_t)chunk * per_env + 1024;
#define SWEEP_CUDA(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) return cuda_fail(_e, #call); } while (0)
  for (int i = 0; i < HostStage::NSLOT; ++i) {
    if (!hs->st[i]) SWEEP_CUDA(cudaStreamCreateWithFlags(&hs->st[i], cudaStreamNonBlocking));
    if (hs->bytes[i] < slot_bytes) {
      if (hs->d_base[i]) { SWEEP_CUDA(cudaStreamSynchronize(hs->st[i])); cudaFree(hs->d_base[i]); hs->d_base[i] = nullptr; hs->bytes[i] = 0; }
      SWEEP_CUDA(cudaMalloc((void**)&hs->d_base[i], slot_bytes));
      hs->bytes[i] = slot_bytes;
    }
  }
  if (!hs->d_live) SWEEP_CUDA(cudaMalloc((void**)&hs->d_live, 8));
  SWEEP_CUDA(cudaMemsetAsync(hs->d_live, 0, 8, hs->st[0]));
  SWEEP_CUDA(cudaStreamSynchronize(hs->st[0]));
  int slot = 0;
  for (int64_t off = 0; off < n && rc == POSG_OK; off += chunk, slot = (slot + 1) % HostStage::NSLOT) {
    const int64_t c = (n - off < chunk) ? (n - off) : chunk;
    cudaStream_t s = hs->st[slot];
    unsigned char* base = hs->d_base[slot];
    uint32_t* d_state = (uint32_t*)base;
    uint64_t* d_obs = (uint64_t*)(base + (size_t)chunk * 8);
    double* d_rew = (double*)(base + (size_t)chunk * 24);
    uint8_t* d_flags = base + (size_t)chunk * 40;
    uint8_t* d_act = base + (((size_t)chunk * 41 + 255) & ~(size_t)255);
    SWEEP_CUDA(cudaMemcpyAsync(d_state, state_in_host + 2 * off, (size_t)c * 8, cudaMemcpyHostToDevice, s));
    for (int k = 0; k < num_steps && rc == POSG_OK; ++k)
      rc = coded ? step_fast_coded(m, c, d_state, seed, (uint32_t)k, particle_offset + (uint64_t)off, d_state, d_obs,
                                   (uint8_t*)d_rew, d_flags, d_act, hs->d_live, s)
                 : posg_driving_step(m, c, d_state, nullptr, nullptr, seed, (uint32_t)k, particle_offset + (uint64_t)off,
                                     d_state, d_obs, d_rew, d_flags, d_act, hs->d_live, s);
    if (rc != POSG_OK) break;
    if (coded) SWEEP_CUDA(cudaMemcpyAsync(code_out_host + off, d_rew, (size_t)c, cudaMemcpyDeviceToHost, s));
    if (state_out_host) SWEEP_CUDA(cudaMemcpyAsync(state_out_host + 2 * off, d_state, (size_t)c * 8, cudaMemcpyDeviceToHost, s));
    if (obs_out_host) SWEEP_CUDA(cudaMemcpyAsync(obs_out_host + 2 * off, d_obs, (size_t)c * 16, cudaMemcpyDeviceToHost, s));
    if (rewards_out_host) SWEEP_CUDA(cudaMemcpyAsync(rewards_out_host + 2 * off, d_rew, (size_t)c * 16, cudaMemcpyDeviceToHost, s));
    if (flags_out_host) SWEEP_CUDA(cudaMemcpyAsync(flags_out_host + off, d_flags, (size_t)c, cudaMemcpyDeviceToHost, s));
  }
  for (int i = 0; i < HostStage::NSLOT; ++i) {
    const cudaError_t e = cudaStreamSynchronize(hs->st[i]);
    if (e != cudaSuccess && rc == POSG_OK) rc = cuda_fail(e, "cudaStreamSynchronize");
  }
  if (rc == POSG_OK && live_count_host) SWEEP_CUDA(cudaMemcpy(live_count_host, hs->d_live, 8, cudaMemcpyDeviceToHost));
#undef SWEEP_CUDA
  return rc;
}

extern "C" {

int posg_driving_sweep_host(posg_model_t m, int64_t n, const uint32_t* state_in_host, int num_steps,
                            uint64_t seed, uint64_t particle_offset, uint32_t* state_out_host,
                            uint64_t* obs_out_host, double* rewards_out_host, uint8_t* flags_out_host,
                            unsigned long long* live_count_host) {
  return sweep_host_impl(m, n, state_in_host, num_steps, seed, particle_offset, state_out_host, obs_out_host,
                         rewards_out_host, nullptr, flags_out_host, live_count_host);
}

int posg_driving_sweep_host_compact(posg_model_t m, int64_t n, const uint32_t* state_in_host, int num_steps,
                                    uint64_t seed, uint64_t particle_offset, uint32_t* state_out_host,
                                    uint64_t* obs_out_host, uint8_t* reward_code_out_host, uint8_t* flags_out_host,
                                    unsigned long long* live_count_host) {
  POSG_REQUIRE(reward_code_out_host, "null reward code buffer");
  return sweep_host_impl(m, n, state_in_host, num_steps, seed, particle_offset, state_out_host, obs_out_host, nullptr,
                         reward_code_out_host, flags_out_host, live_count_host);
}

int posg_driving_decode_rewards_host(posg_model_t m, int64_t n, const uint32_t* state_host,
                                     const uint8_t* reward_code_host, double* rewards_out_host) {
  int rc = check_driving(m); if (rc) return rc;
  POSG_REQUIRE(n >= 0 && (n == 0 || (state_host && reward_code_host && rewards_out_host)), "bad arguments");
  const Blob* B = (const Blob*)m->h_blob;
  for (int64_t i = 0; i < n; ++i) {
    const uint32_t code = reward_code_host[i];
#pragma GCC unroll 2
    for (int a = 0; a < 2; ++a) {
      const uint32_t c4 = (code >> (4 * a)) & 15u, d0 = (state_host[2 * i + a] >> 24) & 63u;
      rewards_out_host[2 * i + a] = B->reward_tab[((d0 | ((c4 & 3u) << 6)) << 2) | (c4 >> 2)];
    }
  }
  return POSG_OK;
}

}  // extern "C"
