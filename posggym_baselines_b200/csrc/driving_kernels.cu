// Driving-v1 kernels + C-ABI.  One thread per env / particle; persistent CTAs
// (grid = k x SM count) stage the model blob into shared memory once via TMA and
// then grid-stride over the batch with coalesced 8/16-byte loads and stores.
#include <cmath>
#include <cstdlib>
#include <cstring>
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
template <int NA>
__global__ void __launch_bounds__(THREADS)
step_kernel(const Blob* __restrict__ gblob, int64_t n, const uint32_t* __restrict__ state_in,
            const uint8_t* __restrict__ actions_in, const uint8_t* __restrict__ order_in,
            uint64_t seed, uint32_t step_idx, uint64_t particle_offset,
            uint32_t* __restrict__ state_out, uint64_t* __restrict__ obs_out,
            double* __restrict__ rewards_out, uint8_t* __restrict__ flags_out,
            uint8_t* __restrict__ actions_out, unsigned long long* __restrict__ live_count) {
  __shared__ Blob B;
  __shared__ alignas(8) uint64_t mbar;
  stage_blob(&B, gblob, (uint32_t)sizeof(Blob), &mbar);

  unsigned live = 0;
  const int64_t stride = (int64_t)gridDim.x * THREADS;
  for (int64_t i = (int64_t)blockIdx.x * THREADS + threadIdx.x; i < n; i += stride) {
    uint32_t veh[NA], before[NA];
    if (NA == 2) {
      const uint2 v = __ldg(reinterpret_cast<const uint2*>(state_in) + i);
      veh[0] = v.x; veh[1] = v.y;
    } else {
#pragma unroll
      for (int a = 0; a < NA; ++a) veh[a] = __ldg(state_in + i * NA + a);
    }
    bool all_done_before = true;
#pragma unroll
    for (int a = 0; a < NA; ++a) { before[a] = veh[a]; all_done_before &= vdone(veh[a]); }
    live += all_done_before ? 0u : 1u;

    int act[NA], order_idx;
    if (actions_in == nullptr || order_in == nullptr) {
      int ract[NA], rorder;
      sample_step_noise<NA>(seed, particle_offset + (uint64_t)i, step_idx, ract, rorder);
#pragma unroll
      for (int a = 0; a < NA; ++a) act[a] = ract[a];
      order_idx = rorder;
    }
    if (actions_in != nullptr) {
      if (NA == 2) {
        const uint16_t aa = __ldg(reinterpret_cast<const uint16_t*>(actions_in) + i);
        act[0] = aa & 255; act[1] = aa >> 8;
      } else {
#pragma unroll
        for (int a = 0; a < NA; ++a) act[a] = __ldg(actions_in + i * NA + a);
      }
    }
    if (order_in != nullptr) order_idx = __ldg(order_in + i);

    int coll[NA];
    transition<NA>(B, veh, act, order_idx, coll);

    uint64_t obs[NA];
    double rew[NA];
    unsigned flags = 0;
    bool all_done = true;
#pragma unroll
    for (int a = 0; a < NA; ++a) {
      obs[a] = observe<NA>(B, veh, a);
      rew[a] = reward(B, before[a], veh[a], coll[a]);
      const bool t = vdone(veh[a]);
      flags |= (t ? 1u : 0u) << a;
      all_done &= t;
    }
    if (all_done) flags |= POSG_FLAG_ALL_DONE;

    if (NA == 2) {
      reinterpret_cast<uint2*>(state_out)[i] = make_uint2(veh[0], veh[1]);
      reinterpret_cast<ulonglong2*>(obs_out)[i] = make_ulonglong2(obs[0], obs[1]);
      reinterpret_cast<double2*>(rewards_out)[i] = make_double2(rew[0], rew[1]);
      if (actions_out) reinterpret_cast<uint16_t*>(actions_out)[i] = (uint16_t)(act[0] | (act[1] << 8));
    } else {
#pragma unroll
      for (int a = 0; a < NA; ++a) {
        state_out[i * NA + a] = veh[a];
        obs_out[i * NA + a] = obs[a];
        rewards_out[i * NA + a] = rew[a];
        if (actions_out) actions_out[i * NA + a] = (uint8_t)act[a];
      }
    }
    flags_out[i] = (uint8_t)flags;
  }
  if (live_count != nullptr) {
    // warp-aggregated, one atomic per warp
    for (int o = 16; o > 0; o >>= 1) live += __shfl_down_sync(0xffffffffu, live, o);
    if ((threadIdx.x & 31) == 0 && live) atomicAdd(live_count, (unsigned long long)live);
  }
}

// ------------------------------------------------------------------ observe / initial state
template <int NA>
__global__ void __launch_bounds__(THREADS)
observe_kernel(const Blob* __restrict__ gblob, int64_t n, const uint32_t* __restrict__ state,
               uint64_t* __restrict__ obs_out) {
  __shared__ Blob B;
  __shared__ alignas(8) uint64_t mbar;
  stage_blob(&B, gblob, (uint32_t)sizeof(Blob), &mbar);
  const int64_t stride = (int64_t)gridDim.x * THREADS;
  for (int64_t i = (int64_t)blockIdx.x * THREADS + threadIdx.x; i < n; i += stride) {
    uint32_t veh[NA];
#pragma unroll
    for (int a = 0; a < NA; ++a) veh[a] = __ldg(state + i * NA + a);
#pragma unroll
    for (int a = 0; a < NA; ++a) obs_out[i * NA + a] = observe<NA>(B, veh, a);
  }
}

template <int NA>
__global__ void __launch_bounds__(THREADS)
initial_state_kernel(const Blob* __restrict__ gblob, int64_t n, const uint8_t* __restrict__ picks,
                     uint64_t seed, uint64_t particle_offset, uint32_t* __restrict__ state_out) {
  __shared__ Blob B;
  __shared__ alignas(8) uint64_t mbar;
  stage_blob(&B, gblob, (uint32_t)sizeof(Blob), &mbar);
  const int64_t stride = (int64_t)gridDim.x * THREADS;
  for (int64_t i = (int64_t)blockIdx.x * THREADS + threadIdx.x; i < n; i += stride) {
    uint32_t veh[NA];
    if (picks != nullptr) {
      int p[2 * NA];
#pragma unroll
      for (int k = 0; k < 2 * NA; ++k) p[k] = __ldg(picks + i * 2 * NA + k);
      initial_state<NA>(B, p, veh);
    } else {
      initial_state_random<NA>(B, seed, particle_offset + (uint64_t)i, veh);
    }
#pragma unroll
    for (int a = 0; a < NA; ++a) state_out[i * NA + a] = veh[a];
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

template <int NA>
__global__ void __launch_bounds__(THREADS)
rollout_kernel(const Blob* __restrict__ gblob, int64_t n, const uint32_t* __restrict__ state_in,
               int ego, int max_steps, const __grid_constant__ DiscountTable dtab, uint64_t seed, uint32_t step0,
               uint64_t particle_offset, double* __restrict__ returns_out,
               uint8_t* __restrict__ steps_out, uint32_t* __restrict__ state_out,
               unsigned long long* __restrict__ work_counter,
               unsigned long long* __restrict__ live_count) {
  __shared__ Blob B;
  __shared__ alignas(8) uint64_t mbar;
  __shared__ double sdisc[256];
  for (int t = threadIdx.x; t < 256; t += THREADS) sdisc[t] = dtab.v[t];
  stage_blob(&B, gblob, (uint32_t)sizeof(Blob), &mbar);  // contains a __syncthreads()

  const unsigned lane = threadIdx.x & 31u;
  uint32_t veh[NA];
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
      if (lane == (unsigned)(__ffs(idle) - 1)) base = atomicAdd(work_counter, (unsigned long long)cnt);
      base = __shfl_sync(0xffffffffu, base, __ffs(idle) - 1);
      if (item < 0) {
        const int64_t mine = (int64_t)base + __popc(idle & ((1u << lane) - 1u));
        if (mine < n) {
          item = mine;
          if (NA == 2) {
            const uint2 v = __ldg(reinterpret_cast<const uint2*>(state_in) + item);
            veh[0] = v.x; veh[1] = v.y;
          } else {
#pragma unroll
            for (int a = 0; a < NA; ++a) veh[a] = __ldg(state_in + item * NA + a);
          }
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
        int act[NA], order_idx, coll[NA];
        sample_step_noise<NA>(seed, particle_offset + (uint64_t)item, step0 + (uint32_t)k, act, order_idx);
        uint32_t before = 0;
#pragma unroll
        for (int a = 0; a < NA; ++a) if (a == ego) before = veh[a];
        transition<NA>(B, veh, act, order_idx, coll);
        uint32_t after = 0; int c = 0; bool all_done = true;
#pragma unroll
        for (int a = 0; a < NA; ++a) {
          if (a == ego) { after = veh[a]; c = coll[a]; }
          all_done &= vdone(veh[a]);
        }
        const double r = reward(B, before, after, c);
        ret = __dadd_rn(ret, __dmul_rn(sdisc[k], r));
        ++k; ++executed;
        finished = vdone(after) || all_done || (k >= max_steps);
      }
      if (finished) {
        returns_out[item] = ret;
        if (steps_out) steps_out[item] = (uint8_t)k;
        if (state_out) {
#pragma unroll
          for (int a = 0; a < NA; ++a) state_out[item * NA + a] = veh[a];
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

// ------------------------------------------------------------------ host: model build
static bool build_blob(const char* grid_rows, int num_agents, int obs_front, int obs_back,
                       int obs_side, int obstacle_collisions, Blob* out) {
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
  if (H < 1 || H > 16) { set_error("grid height %d outside [1,16]", H); return false; }
  const int W = (int)rows[0].size();
  if (W < 1 || W > 16) { set_error("grid width %d outside [1,16]", W); return false; }
  for (auto& r : rows) if ((int)r.size() != W) { set_error("ragged grid rows"); return false; }
  if (num_agents < 1 || num_agents > MAX_AGENTS) { set_error("num_agents %d outside [1,%d]", num_agents, MAX_AGENTS); return false; }

  memset(out, 0, sizeof(Blob));
  for (int l = 0; l < LINES; ++l) { out->rows[l] = 0xffffffffu; out->cols[l] = 0xffffffffu; }
  bool blk[16][16];
  int n_start = 0, n_dest = 0;
  for (int y = 0; y < H; ++y) {
    for (int x = 0; x < W; ++x) {
      const char ch = rows[y][x];
      blk[y][x] = (ch == '#');
      if (!blk[y][x]) {
        out->rows[y + PAD] &= ~(1u << (x + PAD));
        out->cols[x + PAD] &= ~(1u << (y + PAD));
      }
      uint8_t smask = 0, dmask = 0;
      if (ch == '+') smask = 0x0f;
      else if (ch >= '0' && ch <= '9') smask = (ch - '0') < MAX_AGENTS ? (uint8_t)(1u << (ch - '0')) : 0;
      else if (ch == '-') dmask = 0x0f;
      else if (ch >= 'a' && ch <= 'j') dmask = (ch - 'a') < MAX_AGENTS ? (uint8_t)(1u << (ch - 'a')) : 0;
      else if (ch != '.' && ch != '#') { set_error("bad grid character '%c'", ch); return false; }
      if (smask) {
        if (n_start >= MAX_CELLS) { set_error("more than %d start cells", MAX_CELLS); return false; }
        out->start_xy[n_start] = (uint8_t)(x | (y << 4)); out->start_agents[n_start++] = smask;
      }
      if (dmask) {
        if (n_dest >= MAX_CELLS) { set_error("more than %d destination cells", MAX_CELLS); return false; }
        out->dest_xy[n_dest] = (uint8_t)(x | (y << 4)); out->dest_agents[n_dest++] = dmask;
      }
    }
  }
  if (n_start < num_agents || n_dest < num_agents) { set_error("grid supports fewer than %d agents", num_agents); return false; }
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
        if (out->dist[d][c] + 1 >= 127) { set_error("shortest path longer than 126"); return false; }
        out->dist[d][k] = (uint8_t)(out->dist[d][c] + 1); q.push(k);
      }
    }
  }
  for (int d0 = 0; d0 < MAX_D0; ++d0)
    for (int p = 0; p < 3; ++p)
      out->progress[d0 * 3 + p] = 0.5 * ((double)p / (double)(d0 < 1 ? 1 : d0));
  return true;
}

static int grid_for(int64_t n, int ctas_per_sm) {
  const int64_t need = (n + THREADS - 1) / THREADS;
  const int64_t cap = (int64_t)sm_count() * ctas_per_sm;
  return (int)(need < cap ? (need < 1 ? 1 : need) : cap);
}

}  // namespace driving
}  // namespace posg

using namespace posg;
using namespace posg::driving;

#define DISPATCH_NA(m, ...)                                         \
  switch ((m)->num_agents) {                                        \
    case 1: { constexpr int NA = 1; __VA_ARGS__; } break;           \
    case 2: { constexpr int NA = 2; __VA_ARGS__; } break;           \
    case 3: { constexpr int NA = 3; __VA_ARGS__; } break;           \
    case 4: { constexpr int NA = 4; __VA_ARGS__; } break;           \
    default: set_error("unsupported num_agents"); return POSG_E_UNSUPPORTED; \
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
  Blob* h = (Blob*)malloc(sizeof(Blob));
  if (!build_blob(grid_rows, num_agents, obs_front, obs_back, obs_side, obstacle_collisions, h)) {
    free(h);
    return POSG_E_INVALID;
  }
  posg_model* m = new posg_model();
  m->kind = KIND_DRIVING; m->num_agents = num_agents; m->blob_bytes = (int)sizeof(Blob);
  m->h_blob = h; m->d_blob = nullptr;
  cudaError_t e = cudaGetDevice(&m->device);
  if (e == cudaSuccess) e = cudaMalloc(&m->d_blob, sizeof(Blob));
  if (e == cudaSuccess) e = cudaMemcpy(m->d_blob, h, sizeof(Blob), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { int rc = cuda_fail(e, "model upload"); posg_model_destroy(m); return rc; }
  *out = m;
  return POSG_OK;
}

int posg_driving_sample_initial_state(posg_model_t m, int64_t n, const uint8_t* picks, uint64_t seed,
                                      uint64_t particle_offset, uint32_t* state_out, void* stream) {
  int rc = check_driving(m); if (rc) return rc;
  POSG_REQUIRE(n >= 0 && state_out, "bad arguments");
  if (n == 0) return POSG_OK;
  DISPATCH_NA(m, initial_state_kernel<NA><<<grid_for(n, 4), THREADS, 0, (cudaStream_t)stream>>>(
      (const Blob*)m->d_blob, n, picks, seed, particle_offset, state_out));
  POSG_CUDA(cudaGetLastError());
  return POSG_OK;
}

int posg_driving_observe(posg_model_t m, int64_t n, const uint32_t* state, uint64_t* obs_out, void* stream) {
  int rc = check_driving(m); if (rc) return rc;
  POSG_REQUIRE(n >= 0 && state && obs_out, "bad arguments");
  if (n == 0) return POSG_OK;
  DISPATCH_NA(m, observe_kernel<NA><<<grid_for(n, 4), THREADS, 0, (cudaStream_t)stream>>>(
      (const Blob*)m->d_blob, n, state, obs_out));
  POSG_CUDA(cudaGetLastError());
  return POSG_OK;
}

int posg_driving_step(posg_model_t m, int64_t n, const uint32_t* state_in, const uint8_t* actions_in,
                      const uint8_t* order_in, uint64_t seed, uint32_t step_idx, uint64_t particle_offset,
                      uint32_t* state_out, uint64_t* obs_out, double* rewards_out, uint8_t* flags_out,
                      uint8_t* actions_out, unsigned long long* live_count, void* stream) {
  int rc = check_driving(m); if (rc) return rc;
  POSG_REQUIRE(n >= 0 && state_in && state_out && obs_out && rewards_out && flags_out, "bad arguments");
  if (n == 0) return POSG_OK;
  DISPATCH_NA(m, step_kernel<NA><<<grid_for(n, 8), THREADS, 0, (cudaStream_t)stream>>>(
      (const Blob*)m->d_blob, n, state_in, actions_in, order_in, seed, step_idx, particle_offset,
      state_out, obs_out, rewards_out, flags_out, actions_out, live_count));
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
  DISPATCH_NA(m, rollout_kernel<NA><<<grid_for(n, 8), THREADS, 0, stream>>>(
      (const Blob*)m->d_blob, n, state_in, ego, max_steps, dtab, seed, step0, particle_offset,
      returns_out, steps_out, state_out, work_counter, live_count));
  POSG_CUDA(cudaGetLastError());
  return POSG_OK;
}

int posg_driving_rollout(posg_model_t m, int64_t n, const uint32_t* state_in, int ego_agent, int max_steps,
                         double discount, uint64_t seed, uint32_t step0, uint64_t particle_offset,
                         double* returns_out, uint8_t* steps_out, uint32_t* state_out,
                         unsigned long long* live_count, void* stream) {
  int rc = check_driving(m); if (rc) return rc;
  POSG_REQUIRE(n >= 0 && state_in && returns_out, "bad arguments");
  POSG_REQUIRE(ego_agent >= 0 && ego_agent < m->num_agents, "ego_agent out of range");
  POSG_REQUIRE(max_steps >= 0 && max_steps <= 255, "max_steps outside [0,255]");
  if (n == 0) return POSG_OK;
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
  POSG_REQUIRE(n >= 0 && state_in_host && returns_out_host, "bad arguments");
  if (n == 0) return POSG_OK;
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

int posg_driving_sweep_host(posg_model_t m, int64_t n, const uint32_t* state_in_host, int num_steps,
                            uint64_t seed, uint64_t particle_offset, uint32_t* state_out_host,
                            uint64_t* obs_out_host, double* rewards_out_host, uint8_t* flags_out_host,
                            unsigned long long* live_count_host) {
  int rc = check_driving(m); if (rc) return rc;
  POSG_REQUIRE(n >= 0 && state_in_host && num_steps >= 1, "bad arguments");
  if (n == 0) return POSG_OK;
  const size_t na = (size_t)m->num_agents;
  uint32_t* d_state = nullptr; uint64_t* d_obs = nullptr; double* d_rew = nullptr;
  uint8_t *d_flags = nullptr, *d_act = nullptr; unsigned long long* d_live = nullptr;
  cudaStream_t s = nullptr;
  POSG_CUDA(cudaMalloc((void**)&d_state, n * na * 4));
  POSG_CUDA(cudaMalloc((void**)&d_obs, n * na * 8));
  POSG_CUDA(cudaMalloc((void**)&d_rew, n * na * 8));
  POSG_CUDA(cudaMalloc((void**)&d_flags, n));
  POSG_CUDA(cudaMalloc((void**)&d_act, n * na));
  POSG_CUDA(cudaMalloc((void**)&d_live, 8));
  POSG_CUDA(cudaMemcpyAsync(d_state, state_in_host, n * na * 4, cudaMemcpyHostToDevice, s));
  POSG_CUDA(cudaMemsetAsync(d_live, 0, 8, s));
  for (int k = 0; k < num_steps && rc == POSG_OK; ++k)
    rc = posg_driving_step(m, n, d_state, nullptr, nullptr, seed, (uint32_t)k, particle_offset, d_state,
                           d_obs, d_rew, d_flags, d_act, d_live, s);
  if (rc == POSG_OK) {
    if (state_out_host) POSG_CUDA(cudaMemcpyAsync(state_out_host, d_state, n * na * 4, cudaMemcpyDeviceToHost, s));
    if (obs_out_host) POSG_CUDA(cudaMemcpyAsync(obs_out_host, d_obs, n * na * 8, cudaMemcpyDeviceToHost, s));
    if (rewards_out_host) POSG_CUDA(cudaMemcpyAsync(rewards_out_host, d_rew, n * na * 8, cudaMemcpyDeviceToHost, s));
    if (flags_out_host) POSG_CUDA(cudaMemcpyAsync(flags_out_host, d_flags, n, cudaMemcpyDeviceToHost, s));
    if (live_count_host) POSG_CUDA(cudaMemcpyAsync(live_count_host, d_live, 8, cudaMemcpyDeviceToHost, s));
    POSG_CUDA(cudaStreamSynchronize(s));
  }
  cudaFree(d_state); cudaFree(d_obs); cudaFree(d_rew); cudaFree(d_flags); cudaFree(d_act); cudaFree(d_live);
  return rc;
}

}  // extern "C"
