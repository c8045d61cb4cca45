// Counter-based Philox4x32-10 (Salmon et al., SC'11) -- one 128-bit block per
// (particle, step, stream); no RNG state is ever stored, so results do not depend on
// the launch geometry or the number of GPUs.
#pragma once
#include <stdint.h>

namespace posg {

struct NoCtx {};   // context of model adapters that need none (see search.cuh)

constexpr uint32_t PHILOX_M0 = 0xD2511F53u;
constexpr uint32_t PHILOX_M1 = 0xCD9E8D57u;
constexpr uint32_t PHILOX_W0 = 0x9E3779B9u;
constexpr uint32_t PHILOX_W1 = 0xBB67AE85u;

// streams (4th counter word)
constexpr uint32_t STREAM_STEP = 0xA0u;      // joint action + step noise
constexpr uint32_t STREAM_STEP2 = 0xA1u;     // extra step noise (agents 2,3 / prey moves)
constexpr uint32_t STREAM_INIT = 0xB0u;      // initial-state picks
constexpr uint32_t STREAM_INIT2 = 0xB1u;
constexpr uint32_t STREAM_BELIEF = 0xC0u;    // particle index draws (belief.sample)
constexpr uint32_t STREAM_TREE = 0xD0u;      // tie-breaks inside the search tree

__host__ __device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                        uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
    uint32_t hi0 = __umulhi(PHILOX_M0, c0), lo0 = PHILOX_M0 * c0;
    uint32_t hi1 = __umulhi(PHILOX_M1, c2), lo1 = PHILOX_M1 * c2;
#else
    uint64_t p0 = (uint64_t)PHILOX_M0 * c0, p1 = (uint64_t)PHILOX_M1 * c2;
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
    uint32_t n0 = hi1 ^ c1 ^ k0;
    uint32_t n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += PHILOX_W0; k1 += PHILOX_W1;
  }
  uint4 out; out.x = c0; out.y = c1; out.z = c2; out.w = c3;
  return out;
}

__host__ __device__ __forceinline__ uint4 philox_block(uint64_t seed, uint64_t particle,
                                                       uint32_t step, uint32_t stream) {
  return philox4x32_10((uint32_t)particle, (uint32_t)(particle >> 32), step, stream,
                       (uint32_t)seed, (uint32_t)(seed >> 32));
}

// The ten round keys of one seed, computed once on the host and passed as a (__grid_constant__) kernel parameter:
// the key schedule is the same for every block of a launch, and as constant-bank operands the keys cost the
// integer ALU pipe nothing (two additions per round and block otherwise).
struct PhiloxKeys {
  uint32_t k0[10], k1[10];
  PhiloxKeys() = default;
  __host__ __device__ explicit PhiloxKeys(uint64_t seed) {
    uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) { k0[r] = a; k1[r] = b; a += PHILOX_W0; b += PHILOX_W1; }
  }
};

__device__ __forceinline__ uint4 philox_block(const PhiloxKeys& K, uint64_t particle, uint32_t step, uint32_t stream) {
  uint32_t c0 = (uint32_t)particle, c1 = (uint32_t)(particle >> 32), c2 = step, c3 = stream;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(PHILOX_M0, c0), lo0 = PHILOX_M0 * c0;
    const uint32_t hi1 = __umulhi(PHILOX_M1, c2), lo1 = PHILOX_M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ K.k0[r];
    const uint32_t n2 = hi0 ^ c3 ^ K.k1[r];
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
  }
  uint4 out; out.x = c0; out.y = c1; out.z = c2; out.w = c3;
  return out;
}

// uniform integer in [0, n) by multiply-shift (bias < n / 2^32)
__host__ __device__ __forceinline__ uint32_t uniform_below(uint32_t word, uint32_t n) {
#ifdef __CUDA_ARCH__
  return __umulhi(word, n);
#else
  return (uint32_t)(((uint64_t)word * n) >> 32);
#endif
}

}  // namespace posg
