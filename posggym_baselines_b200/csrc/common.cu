// libposg_b200: error reporting, device queries, model handle lifetime.
#include "common.h"

#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <utility>

namespace posg {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  // clear the sticky-free error so the next call reports its own failure
  cudaGetLastError();
  return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? POSG_E_NOGPU : POSG_E_CUDA;
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// The planner's host entry points take a few bytes of scratch with cudaMallocAsync.  By default the
// pool gives its memory back to the driver at every synchronisation and maps it again at the next
// allocation (milliseconds each; the very first allocation creates the pool: ~0.5 s).  Keep freed
// blocks cached and pay the pool creation when the model is built, not inside planner.update().
void async_pool_init() {
  static std::mutex mu;
  static bool done[64] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { cudaGetLastError(); return; }
  std::lock_guard<std::mutex> lock(mu);
  if (done[dev]) return;
  done[dev] = true;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    unsigned long long keep = ~0ULL;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    void* p = nullptr;
    if (cudaMallocAsync(&p, 256, 0) == cudaSuccess) { cudaFreeAsync(p, 0); cudaStreamSynchronize(0); }
  }
  cudaGetLastError();
}

void smem_optin_impl(const void* kernel, int bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, int> granted;   // (kernel, device) -> bytes opted in
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  int& have = granted[std::make_pair(kernel, dev)];
  if (have >= bytes) return;
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) == cudaSuccess) have = bytes;
  else cudaGetLastError();   // the launch that follows reports the failure
}

// posg_trace_hash: one thread per particle; every pointer is read once, coalesced.
__global__ void __launch_bounds__(256)
trace_hash_kernel(int64_t n, unsigned long long* __restrict__ hash, int init, uint64_t particle_offset,
                  const uint32_t* __restrict__ state, int sw, const unsigned long long* __restrict__ obs, int ow,
                  const unsigned long long* __restrict__ rew_bits, int nag, const uint8_t* __restrict__ flags,
                  const uint8_t* __restrict__ actions) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    unsigned long long h = init ? (particle_offset + (uint64_t)i + 1ULL) * 0xD6E8FEB86659FD93ULL : hash[i];
    auto mix = [&h](unsigned long long w) { h = (h ^ w) * 0x9E3779B97F4A7C15ULL; h ^= h >> 32; };
    if (state != nullptr) {
      for (int k = 0; k < sw; ++k) mix((unsigned long long)__ldg(state + i * sw + k));
      for (int k = 0; k < ow; ++k) mix(__ldg(obs + i * ow + k));
      for (int k = 0; k < nag; ++k) mix(__ldg(rew_bits + i * nag + k));
      unsigned long long tail = __ldg(flags + i);
      for (int k = 0; k < nag; ++k) tail |= (unsigned long long)__ldg(actions + i * nag + k) << (8 * (k + 1));
      mix(tail);
    }
    hash[i] = h;
  }
}

}  // namespace posg

extern "C" {

int posg_trace_hash(int64_t n, uint64_t* hash_inout, int init, uint64_t particle_offset, const uint32_t* state,
                    int state_words, const uint64_t* obs, int obs_words, const double* rewards, int num_agents,
                    const uint8_t* flags, const uint8_t* actions, void* stream) {
  if (n == 0) return POSG_OK;
  POSG_REQUIRE(n > 0 && hash_inout, "bad arguments");
  POSG_REQUIRE(state == nullptr || (obs && rewards && flags && actions && state_words > 0 && obs_words > 0 &&
                                    num_agents > 0 && num_agents <= 7),
               "posg_trace_hash: every output array of the step is needed");
  const int64_t blocks = (n + 255) / 256, cap = (int64_t)posg::sm_count() * 8;
  posg::trace_hash_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(
      n, (unsigned long long*)hash_inout, init, particle_offset, state, state_words, (const unsigned long long*)obs,
      obs_words, (const unsigned long long*)rewards, num_agents, flags, actions);
  POSG_CUDA(cudaGetLastError());
  return POSG_OK;
}

const char* posg_last_error(void) { return posg::g_err; }

int posg_version(void) { return 100; }

int posg_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    posg::set_error("no CUDA device visible (%s); libposg_b200 has no CPU fallback",
                    e == cudaSuccess ? "count=0" : cudaGetErrorString(e));
    cudaGetLastError();
    return POSG_E_NOGPU;
  }
  return n;
}

int posg_model_destroy(posg_model_t m) {
  if (!m) return POSG_OK;
  if (!m->borrowed) {
    if (m->d_blob) cudaFree(m->d_blob);
    free(m->h_blob);
  }
  if (m->d_flat_desc) cudaFree(m->d_flat_desc);
  if (m->host_stage && m->host_stage_free) m->host_stage_free(m->host_stage);
  if (m->step_stage && m->step_stage_free) m->step_stage_free(m->step_stage);
  delete m;
  return POSG_OK;
}

int posg_model_blob_bytes(posg_model_t m) { return m ? m->blob_bytes : POSG_E_INVALID; }
int posg_model_num_agents(posg_model_t m) { return m ? m->num_agents : POSG_E_INVALID; }

}  // extern "C"
