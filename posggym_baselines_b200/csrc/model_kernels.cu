// Model-generic batch entry points of the C-ABI (posg_model_*), dispatched on the model
// kind, and the PredatorPrey-v0 model builder.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "common.h"
#include "generic_kernels.cuh"
#include "pp.cuh"
#include "lbf.cuh"

namespace posg {

template <typename K>
static int persistent_grid(K kernel, size_t smem, int64_t work_items) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, generic::THREADS, smem) != cudaSuccess || per_sm < 1) {
    cudaGetLastError();
    per_sm = 1;
  }
  const int64_t need = (work_items + generic::THREADS - 1) / generic::THREADS;
  const int64_t cap = (int64_t)sm_count() * per_sm;
  return (int)(need < cap ? (need < 1 ? 1 : need) : cap);
}

static pp::Blob* build_pp_blob(int size, const char* blocks, int num_predators, int num_prey, int cooperative,
                               int prey_strength, int obs_dim) {
  if (size < 3 || size > 16) { set_error("grid size %d outside [3,16]", size); return nullptr; }
  if (num_predators != pp::NP || num_prey != pp::NQ) {
    set_error("the sm_100a kernels are specialised for 2 predators and 3 prey"); return nullptr;
  }
  if (obs_dim != 2) { set_error("obs_dim %d unsupported: kernels are specialised for obs_dim 2", obs_dim); return nullptr; }
  if (prey_strength < 1 || prey_strength > num_predators) { set_error("bad prey_strength"); return nullptr; }
  pp::Blob* b = (pp::Blob*)calloc(1, sizeof(pp::Blob));
  bool blk[16][16];
  memset(blk, 0, sizeof(blk));
  if (blocks) {
    // optional '#'/'.' rows
    int x = 0, y = 0;
    for (const char* p = blocks; *p; ++p) {
      if (*p == '\n') { ++y; x = 0; continue; }
      if (x < 16 && y < 16 && *p == '#') blk[y][x] = true;
      ++x;
    }
  }
  for (int l = 0; l < pp::LINES; ++l) b->rows[l] = 0xffffffffu;
  for (int y = 0; y < size; ++y)
    for (int x = 0; x < size; ++x)
      if (!blk[y][x]) b->rows[y + pp::PAD] &= ~(1u << (x + pp::PAD));
  // predator starts: corners and half-way along each side (itertools.product order: x outer)
  const int v[3] = {0, size / 2, size - 1};
  int n = 0;
  for (int a = 0; a < 3; ++a)
    for (int c = 0; c < 3; ++c) {
      const int x = v[a], y = v[c];
      if (x == 0 || x == size - 1 || y == 0 || y == size - 1) b->pred_start[n++] = (uint8_t)(x | (y << 4));
    }
  b->n_pred_start = (uint8_t)n;
  // prey starts: whole rings around the centre until at least num_prey unblocked cells
  const int cx = size / 2, cy = size / 2;
  n = 0;
  for (int ring = 0; n < num_prey && ring < size; ++ring)
    for (int x = cx - ring; x <= cx + ring; ++x)
      for (int y = cy - ring; y <= cy + ring; ++y) {
        const int d = std::max(abs(x - cx), abs(y - cy));
        if (d != ring || x < 0 || y < 0 || x >= size || y >= size || blk[y][x] || n >= 32) continue;
        b->prey_start[n++] = (uint8_t)(x | (y << 4));
      }
  b->n_prey_start = (uint8_t)n;
  b->size = (uint8_t)size; b->obs_dim = (uint8_t)obs_dim; b->prey_strength = (uint8_t)prey_strength;
  b->cooperative = cooperative ? 1 : 0;
  b->per_prey_reward = 1.0 / (double)num_prey;
  for (int k = 0; k < 4; ++k) b->reward_k[k] = (double)k * b->per_prey_reward;
  const int w = 2 * obs_dim + 1;
  for (int y = 0; y < 16; ++y)
    for (int x = 0; x < 16; ++x) {
      unsigned long long cells = 0;
      for (int col = 0; col < w; ++col)
        for (int row = 0; row < w; ++row) {
          const int gx = x + col - obs_dim, gy = y + row - obs_dim;
          const bool wall = gx < 0 || gy < 0 || gx >= size || gy >= size || blk[gy][gx];
          if (wall) cells |= (unsigned long long)pp::WALL << (2 * (row * w + col));
        }
      b->obs_base[x | (y << 4)] = cells;
    }
  pp::build_tables(b);
  return b;
}

}  // namespace posg

using namespace posg;

// PredatorPrey: the table-driven adapter where its tables are exact (grids up to 10x10), else the general one
static inline bool pp_lut(const posg_model* m) { return ((const pp::Blob*)m->h_blob)->lut_ok != 0; }

#define MODEL_DISPATCH(m, CALL)                                                   \
  switch ((m)->kind) {                                                            \
    case KIND_PREDATOR_PREY:                                                      \
      if (pp_lut(m)) { using M = PPSimLut; CALL; } else { using M = PPSim; CALL; } \
      break;                                                                      \
    case KIND_LBF: { using M = LBFSim; CALL; } break;                             \
    default: set_error("posg_model_*: unsupported model kind %d", (m)->kind);     \
             return POSG_E_UNSUPPORTED;                                           \
  }

extern "C" {

int posg_pp_create(int grid_size, const char* block_rows, int num_predators, int num_prey, int cooperative,
                   int prey_strength, int obs_dim, posg_model_t* out) {
  POSG_REQUIRE(out, "null argument");
  int ndev = posg_device_count();
  if (ndev < 0) return ndev;
  async_pool_init();
  pp::Blob* h = build_pp_blob(grid_size, block_rows, num_predators, num_prey, cooperative, prey_strength, obs_dim);
  if (!h) return POSG_E_UNSUPPORTED;
  posg_model* m = new posg_model();
  m->kind = KIND_PREDATOR_PREY; m->num_agents = num_predators; m->blob_bytes = (int)sizeof(pp::Blob);
  m->h_blob = h; m->d_blob = nullptr;
  cudaError_t e = cudaGetDevice(&m->device);
  if (e == cudaSuccess) e = cudaMalloc(&m->d_blob, sizeof(pp::Blob));
  if (e == cudaSuccess) e = cudaMemcpy(m->d_blob, h, sizeof(pp::Blob), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { int rc = cuda_fail(e, "model upload"); posg_model_destroy(m); return rc; }
  *out = m;
  return POSG_OK;
}

int posg_lbf_create(int size, int num_agents, int max_agent_level, int max_food, int sight, int force_coop,
                    int normalize_reward, const uint8_t* food_xy, const uint8_t* player_xy, posg_model_t* out) {
  POSG_REQUIRE(out && food_xy && player_xy, "null argument");
  if (num_agents != lbf::NPL || max_food != lbf::NF) {
    set_error("the sm_100a kernels are specialised for 2 players and 8 foods");
    return POSG_E_UNSUPPORTED;
  }
  if (size < 3 || size > 15 || sight != 2 || max_agent_level < 1 || max_agent_level > 3) {
    set_error("unsupported LBF geometry (size 3..15, sight 2, max_agent_level 1..3)");
    return POSG_E_UNSUPPORTED;
  }
  int ndev = posg_device_count();
  if (ndev < 0) return ndev;
  async_pool_init();
  lbf::Blob* h = (lbf::Blob*)calloc(1, sizeof(lbf::Blob));
  for (int k = 0; k < lbf::NF; ++k) h->food_xy[k] = food_xy[k];
  for (int i = 0; i < lbf::NPL; ++i) h->player_xy[i] = player_xy[i];
  h->size = (uint8_t)size; h->sight = (uint8_t)sight; h->max_agent_level = (uint8_t)max_agent_level;
  h->force_coop = force_coop ? 1 : 0; h->normalize_reward = normalize_reward ? 1 : 0;
  lbf::build_tables(h);
  posg_model* m = new posg_model();
  m->kind = KIND_LBF; m->num_agents = num_agents; m->blob_bytes = (int)sizeof(lbf::Blob);
  m->h_blob = h; m->d_blob = nullptr;
  // the reward quotient table follows the staged blob in the same allocation (only the blob is staged into shared
  // memory; the table stays in global memory)
  std::vector<double> div((size_t)lbf::REWARD_DEN * lbf::REWARD_NUM, 0.0);
  for (int d = 1; d < lbf::REWARD_DEN; ++d)
    for (int n = 0; n < lbf::REWARD_NUM; ++n) div[(size_t)d * lbf::REWARD_NUM + n] = (double)n / (double)d;
  cudaError_t e = cudaGetDevice(&m->device);
  if (e == cudaSuccess) e = cudaMalloc(&m->d_blob, sizeof(lbf::Blob) + div.size() * sizeof(double));
  if (e == cudaSuccess) {
    lbf::Blob up = *h;
    up.reward_div = (const double*)((const unsigned char*)m->d_blob + sizeof(lbf::Blob));
    e = cudaMemcpy(m->d_blob, &up, sizeof(lbf::Blob), cudaMemcpyHostToDevice);
  }
  if (e == cudaSuccess)
    e = cudaMemcpy((unsigned char*)m->d_blob + sizeof(lbf::Blob), div.data(), div.size() * sizeof(double), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { int rc = cuda_fail(e, "model upload"); posg_model_destroy(m); return rc; }
  *out = m;
  return POSG_OK;
}

int posg_model_sample_initial_state(posg_model_t m, int64_t n, const uint8_t* picks, uint64_t seed,
                                    uint64_t particle_offset, uint32_t* state_out, void* stream) {
  POSG_REQUIRE(m, "null model");
  if (n == 0) return POSG_OK;
  POSG_REQUIRE(n > 0 && state_out, "bad arguments");
  MODEL_DISPATCH(m, (generic::initial_state_kernel<M><<<persistent_grid(generic::initial_state_kernel<M>, m->blob_bytes, n),
                     generic::THREADS, m->blob_bytes, (cudaStream_t)stream>>>(
                     (const M::Blob*)m->d_blob, (uint32_t)m->blob_bytes, n, picks, seed, particle_offset, state_out)));
  POSG_CUDA(cudaGetLastError());
  return POSG_OK;
}

int posg_model_observe(posg_model_t m, int64_t n, const uint32_t* state, uint64_t* obs_out, void* stream) {
  POSG_REQUIRE(m, "null model");
  if (n == 0) return POSG_OK;
  POSG_REQUIRE(n > 0 && state && obs_out, "bad arguments");
  MODEL_DISPATCH(m, (generic::observe_kernel<M><<<persistent_grid(generic::observe_kernel<M>, m->blob_bytes, n),
                     generic::THREADS, m->blob_bytes, (cudaStream_t)stream>>>(
                     (const M::Blob*)m->d_blob, (uint32_t)m->blob_bytes, n, state, obs_out)));
  POSG_CUDA(cudaGetLastError());
  return POSG_OK;
}

int posg_model_step(posg_model_t m, int64_t n, const uint32_t* state_in, const uint8_t* actions_in,
                    const uint32_t* noise_in, uint64_t seed, uint32_t step_idx, uint64_t particle_offset,
                    uint32_t* state_out, uint64_t* obs_out, double* rewards_out, uint8_t* flags_out,
                    uint8_t* actions_out, unsigned long long* live_count, void* stream) {
  POSG_REQUIRE(m, "null model");
  if (n == 0) return POSG_OK;
  POSG_REQUIRE(n > 0 && state_in && state_out && obs_out && rewards_out && flags_out, "bad arguments");
  POSG_REQUIRE((((uintptr_t)state_in | (uintptr_t)state_out) & 7u) == 0 &&
               (((uintptr_t)obs_out | (uintptr_t)rewards_out) & 15u) == 0 && ((uintptr_t)actions_out & 1u) == 0,
               "state buffers must be 8-byte, obs / rewards buffers 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  int64_t done = 0;
  static const bool nofast = getenv("POSG_STEP_NOFAST") != nullptr;   // dev A/B switch
  const bool fast_kind = m->kind == KIND_LBF || (m->kind == KIND_PREDATOR_PREY && pp_lut(m));
  // (PredatorPrey: particles g and g + 32 share a Philox block, so its blocks of 128 must start at a multiple of 64)
  if (!nofast && fast_kind && !actions_in && !noise_in && (particle_offset & (m->kind == KIND_LBF ? 3ULL : 63ULL)) == 0 &&
      n >= generic::FAST_WARP_ENVS && n < ((int64_t)1 << 31)) {
    // whole blocks of 128 envs through the fast kernel; a ragged tail through the general one
    const uint32_t nb = (uint32_t)(n / generic::FAST_WARP_ENVS);
#define POSG_FAST(WA, CL)                                                                                        \
    generic::step_fast_kernel<M, WA, CL><<<persistent_grid(generic::step_fast_kernel<M, WA, CL>, m->blob_bytes,   \
                                                           (int64_t)nb * 32), generic::THREADS, m->blob_bytes, s>>>( \
        (const M::Blob*)m->d_blob, (uint32_t)m->blob_bytes, nb, (const uint2*)state_in, PhiloxKeys(seed), step_idx, \
        particle_offset >> 2, (uint2*)state_out, (ulonglong2*)obs_out, (double2*)rewards_out, flags_out,          \
        (uchar2*)actions_out, live_count)
#define POSG_FAST_ALL()                                 \
    if (actions_out && live_count) POSG_FAST(true, true); \
    else if (actions_out) POSG_FAST(true, false);         \
    else if (live_count) POSG_FAST(false, true);          \
    else POSG_FAST(false, false)
    if (m->kind == KIND_LBF) { using M = LBFSim; POSG_FAST_ALL(); }
    else { using M = PPSimLut; POSG_FAST_ALL(); }
#undef POSG_FAST_ALL
#undef POSG_FAST
    POSG_CUDA(cudaGetLastError());
    done = (int64_t)nb * generic::FAST_WARP_ENVS;
  }
  if (done < n) {
    const int64_t r = n - done;
    int n_act = 0, ow = 0, sw = 0;
    int rc = posg_model_info(m, &n_act, &ow, &sw); if (rc) return rc;
    const int na = m->num_agents;
    const uint32_t* sin = state_in + done * sw;
    uint32_t* sout = state_out + done * sw;
    uint64_t* oout = obs_out + done * na * ow;
    double* rout = rewards_out + done * na;
    uint8_t* fout = flags_out + done;
    const uint8_t* ain = actions_in ? actions_in + done * na : nullptr;
    uint8_t* aout = actions_out ? actions_out + done * na : nullptr;
    const uint64_t off = particle_offset + (uint64_t)done;
    MODEL_DISPATCH(m, (generic::step_kernel<M><<<persistent_grid(generic::step_kernel<M>, m->blob_bytes, r),
                       generic::THREADS, m->blob_bytes, s>>>(
                       (const M::Blob*)m->d_blob, (uint32_t)m->blob_bytes, r, sin, ain,
                       noise_in /* explicit noise excludes the fast path: done == 0 */, seed, step_idx, off, sout, oout,
                       rout, fout, aout, live_count)));
    POSG_CUDA(cudaGetLastError());
  }
  return POSG_OK;
}

int posg_model_rollout(posg_model_t m, int64_t n, const uint32_t* state_in, int ego_agent, int max_steps,
                       double discount, uint64_t seed, uint32_t step0, uint64_t particle_offset,
                       double* returns_out, uint8_t* steps_out, uint32_t* state_out,
                       unsigned long long* live_count, void* stream) {
  POSG_REQUIRE(m, "null model");
  POSG_REQUIRE(ego_agent >= 0 && ego_agent < m->num_agents, "ego_agent out of range");
  POSG_REQUIRE(max_steps >= 0 && max_steps <= 255, "max_steps outside [0,255]");
  if (n == 0) return POSG_OK;
  POSG_REQUIRE(n > 0 && state_in && returns_out, "bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  unsigned long long* counter = nullptr;
  POSG_CUDA(cudaMallocAsync((void**)&counter, sizeof(unsigned long long), s));
  POSG_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), s));
  generic::DiscountTable dtab;
  for (int k = 0; k < 256; ++k) dtab.v[k] = pow(discount, (double)k);
  MODEL_DISPATCH(m, (generic::rollout_kernel<M><<<persistent_grid(generic::rollout_kernel<M>, m->blob_bytes, n),
                     generic::THREADS, m->blob_bytes, s>>>(
                     (const M::Blob*)m->d_blob, (uint32_t)m->blob_bytes, n, state_in, ego_agent, max_steps, dtab,
                     PhiloxKeys(seed), step0, particle_offset, returns_out, steps_out, state_out, counter, live_count)));
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(counter, s);
  POSG_CUDA(e);
  return POSG_OK;
}

}  // extern "C"
