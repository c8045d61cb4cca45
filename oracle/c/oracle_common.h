/* Shared pieces of the C oracles (TEST INFRASTRUCTURE ONLY -- never linked into the product):
 * Philox4x32-10 with the kernels' counter conventions (include/posg_b200.h), the bounded
 * integer draw, and the per-particle rolling trace hash of posg_trace_hash
 * (include/posg_b200.h): both sides hash the same documented packed words in the same order,
 * so one 64-bit value per particle certifies every (state, obs, rewards, flags, actions) of a
 * whole sweep. */
#ifndef ORACLE_COMMON_H
#define ORACLE_COMMON_H
#include <stdint.h>
#include <string.h>

#define STREAM_STEP 0xA0u
#define STREAM_INIT 0xB0u
#define STREAM_INIT2 0xB1u

static inline void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                 uint32_t* out) {
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* counter = (index lo, index hi, step, stream), key = seed */
static inline void philox_block(uint64_t seed, uint64_t index, uint32_t step, uint32_t stream, uint32_t* out) {
  philox4x32_10((uint32_t)index, (uint32_t)(index >> 32), step, stream, (uint32_t)seed, (uint32_t)(seed >> 32), out);
}

static inline uint32_t below(uint32_t w, uint32_t n) { return (uint32_t)(((uint64_t)w * n) >> 32); }

/* ---- rolling trace hash (posg_trace_hash) ---- */
static inline uint64_t trace_hash_init(uint64_t particle) { return (particle + 1ULL) * 0xD6E8FEB86659FD93ULL; }
static inline uint64_t trace_hash_mix(uint64_t h, uint64_t w) {
  h = (h ^ w) * 0x9E3779B97F4A7C15ULL;
  return h ^ (h >> 32);
}
static inline uint64_t double_bits(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }

/* one step of one particle: state words (each zero-extended), observation words (agent-major),
 * reward bit patterns, then flags | a0 << 8 | a1 << 16 | ... */
static inline uint64_t trace_hash_step(uint64_t h, const uint32_t* state, int sw, const uint64_t* obs, int ow_total,
                                       const double* rew, int nag, uint8_t flags, const uint8_t* act) {
  for (int k = 0; k < sw; ++k) h = trace_hash_mix(h, (uint64_t)state[k]);
  for (int k = 0; k < ow_total; ++k) h = trace_hash_mix(h, obs[k]);
  for (int k = 0; k < nag; ++k) h = trace_hash_mix(h, double_bits(rew[k]));
  uint64_t tail = flags;
  for (int k = 0; k < nag; ++k) tail |= (uint64_t)act[k] << (8 * (k + 1));
  return trace_hash_mix(h, tail);
}

/* optional full per-step trace buffers, [depth][n][..] (NULL = not recorded) */
typedef struct {
  uint32_t* state;   /* [depth][n][sw] */
  uint64_t* obs;     /* [depth][n][ow_total] */
  double* rew;       /* [depth][n][nag] */
  uint8_t* flags;    /* [depth][n] */
  uint8_t* act;      /* [depth][n][nag] */
} TraceOut;

static inline void trace_record(const TraceOut* t, int64_t n, int64_t i, int k, const uint32_t* state, int sw,
                                const uint64_t* obs, int ow_total, const double* rew, int nag, uint8_t flags,
                                const uint8_t* act) {
  const int64_t row = (int64_t)k * n + i;
  if (t->state) memcpy(t->state + row * sw, state, sizeof(uint32_t) * sw);
  if (t->obs) memcpy(t->obs + row * ow_total, obs, sizeof(uint64_t) * ow_total);
  if (t->rew) memcpy(t->rew + row * nag, rew, sizeof(double) * nag);
  if (t->flags) t->flags[row] = flags;
  if (t->act) memcpy(t->act + row * nag, act, nag);
}
#endif
