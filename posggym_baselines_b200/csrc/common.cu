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

}  // namespace posg

extern "C" {

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
  delete m;
  return POSG_OK;
}

int posg_model_blob_bytes(posg_model_t m) { return m ? m->blob_bytes : POSG_E_INVALID; }
int posg_model_num_agents(posg_model_t m) { return m ? m->num_agents : POSG_E_INVALID; }

}  // extern "C"
