// Shared host-side helpers of libposg_b200 (error reporting, launch geometry).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "../../include/posg_b200.h"

namespace posg {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define POSG_CUDA(call)                                   \
  do {                                                    \
    cudaError_t _e = (call);                              \
    if (_e != cudaSuccess) return posg::cuda_fail(_e, #call); \
  } while (0)

#define POSG_REQUIRE(cond, ...)          \
  do {                                   \
    if (!(cond)) {                       \
      posg::set_error(__VA_ARGS__);      \
      return POSG_E_INVALID;             \
    }                                    \
  } while (0)

// SM count of the current device (cached per device).
int sm_count();

// Configure (once per device) the stream-ordered allocator used for the host entry points' scratch.
void async_pool_init();

// Dynamic shared memory above the 48 KB default needs a per-kernel opt-in (the Driving blob with
// its 64-bit observation table is ~56 KB).  Remembered per (kernel, device); call before a launch.
void smem_optin_impl(const void* kernel, int bytes);
template <typename K>
inline void smem_optin(K kernel, int bytes) {
  // static + dynamic shared memory together must stay under 48 KB without the opt-in; the kernels
  // here hold up to ~22 KB of static arrays, so opt in from 24 KB of dynamic memory on
  if (bytes > 24 * 1024) smem_optin_impl(reinterpret_cast<const void*>(kernel), bytes);
}

enum ModelKind : int { KIND_DRIVING = 1, KIND_PREDATOR_PREY = 2, KIND_LBF = 3, KIND_NESTED = 0x100 };

}  // namespace posg

struct posg_model {
  int kind;
  int device;
  int num_agents;
  int blob_bytes;
  void* d_blob;   // device copy of the immutable model blob
  void* h_blob;   // host copy (for the *_host entry points / introspection)
  void* d_flat_desc = nullptr;  // one-hot descriptors of the flattened observation (vec env)
  int flat_dim = 0;
  int borrowed = 0;             // blob pointers belong to another handle (posg_nested_create)
  // staging of the *_host entry points: device slices + streams, created on first use and
  // reused by later calls (guarded by a mutex: one host-buffer call per handle at a time)
  void* host_stage = nullptr;
  void (*host_stage_free)(void*) = nullptr;
  // staging of posg_model_step_host (small batches from host memory: one pinned block, one device block, one stream)
  void* step_stage = nullptr;
  void (*step_stage_free)(void*) = nullptr;
};
