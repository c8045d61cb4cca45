// The collector's policy step after the first trunk layer: what
// PPOLSTMModel.get_action_and_value (ppo/network.py:212-268,305-330) computes per row from the output of
// trunk[0:2] -- trunk[2:4] = Linear(64, 64) + Tanh, one LSTM(64) cell with the done-reset of get_states,
// actor / critic heads Linear(64, 64) + Tanh + Linear(64, A | 1), Categorical sample, log-prob, entropy --
// as ONE kernel on the 5th-generation tensor cores:
//
//   * a CTA owns 128 rows (= the 128 TMEM lanes); two of its 256 threads share a row in the epilogues (half the
//     columns each);
//   * all weights (fp16, 94 KB) are staged once per CTA by one TMA bulk copy and stay in shared memory
//     in the UMMA K-major no-swizzle "core matrix" layout; activations are written back to shared
//     memory in that layout by the epilogues, so no activation ever leaves the SM between the layers;
//   * four tcgen05.mma groups (kind::f16, fp32 accumulate in TMEM): [128x64]x[64x64], [128x128]x[128x256]
//     (input and recurrent weights side by side: the four gates in one accumulator), [128x64]x[64x128]
//     (both heads), [128x128]x[128x16] (logits and value from a block-diagonal final layer), each
//     committed to an mbarrier the epilogue threads wait on; accumulators are read with tcgen05.ld.
//
// Operands are fp16 (11 significant bits, as tf32 has; weights and tanh-bounded activations are far inside
// its range), accumulation and every epilogue (bias, tanh, sigmoid, cell update, softmax) are fp32.
#include <cuda_fp16.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.h"
#include "philox.cuh"
#include "stage.cuh"

namespace posg {
namespace policy {
#ifdef POSG_POLICY_TRACE
__device__ unsigned long long g_policy_trace[16];
#define PTRACE(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); g_policy_trace[i] = t_; } } while (0)
#else
#define PTRACE(i) do {} while (0)
#endif

constexpr int TILE = 128;   // rows per CTA = TMEM lanes
#ifndef POSG_POLICY_THREADS
#define POSG_POLICY_THREADS 256
#endif
constexpr int THREADS = POSG_POLICY_THREADS;   // NPART threads per row: warps w, w + 4, ... read the same 32 TMEM lanes and split the columns
constexpr int NPART = THREADS / TILE;          // (two or four: more warps hide the epilogues' MUFU / TMEM-load latency)
static_assert(NPART == 2 || NPART == 4, "two or four threads per row");
constexpr int H = 64;       // trunk / LSTM / head width (baseline_exps/exp_utils.py:101-104)
constexpr int NF = 16;      // final layer: logits in columns 0..A-1 (A <= 8), value in column VCOL
constexpr int VCOL = 8;
constexpr uint32_t STREAM_POLICY = 0xC0u;

// byte offset of element (r, k) of an [R x K] fp16 operand, K-major: 8 x 8-element core matrices of 128
// contiguous bytes; cores adjacent in K are 128 B apart (LBO), 8-row groups (K / 8) * 128 B apart (SBO)
__host__ __device__ constexpr uint32_t core_off(int r, int k, int K) {
  return (uint32_t)((r >> 3) * (K >> 3) * 128 + (k >> 3) * 128 + (r & 7) * 16 + (k & 7) * 2);
}

struct alignas(16) Weights {
  __half w1[H * H];            // trunk[2]:            N = 64,  K = 64
  __half wl[4 * H * 2 * H];    // [W_ih | W_hh]:       N = 256 (i, f, g, o), K = 128 ([x | h])
  __half wh[2 * H * H];        // actor[0] | critic[0]: N = 128, K = 64
  __half wf[NF * 2 * H];       // block-diagonal actor[2] / critic[2]: N = 16, K = 128 ([actor hidden | critic hidden])
  float b1[H], bl[4 * H], bh[2 * H], bf[NF];
};
static_assert(sizeof(Weights) % 16 == 0, "bulk copy granularity");

constexpr uint32_t SM_A0 = sizeof(Weights);                 // x0            [128 x 64]
constexpr uint32_t SM_AL = SM_A0 + TILE * H * 2;            // [x1 | h]      [128 x 128]
constexpr uint32_t SM_AH = SM_AL + TILE * 2 * H * 2;        // h'            [128 x 64]
constexpr uint32_t SM_AF = SM_AH + TILE * H * 2;            // [ha | hc]     [128 x 128]
constexpr uint32_t SM_TOTAL = SM_AF + TILE * 2 * H * 2;
// TMEM columns (fp32 accumulators, 128 lanes each)
constexpr uint32_t TM_D1 = 0, TM_G = 64, TM_HD = 320, TM_F = 448, TM_COLS = 512;

__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t sbo_bytes) {
  // UMMA shared-memory descriptor: start address >> 4 [0,14), leading byte offset >> 4 [16,30) (K-adjacent
  // cores: 128 B), stride byte offset >> 4 [32,46) (8-row groups), version 1 [46,48), no swizzle
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
         (1ULL << 46);
}
__host__ __device__ constexpr uint32_t instr_desc(int M, int N) {
  // kind::f16 instruction descriptor: D fp32 (1 << 4), A / B fp16 K-major (zeros), N >> 3 at [17,23), M >> 4 at [24,29)
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] = A[smem, 128 x K] . B[smem, N x K]^T, K in steps of 16 (two cores = 256 B), then commit to the barrier
__device__ __forceinline__ void gemm(uint32_t tmem_d, uint32_t a_addr, uint32_t b_addr, int K, int N, uint32_t bar) {
  const uint32_t sbo = (uint32_t)(K >> 3) * 128u, idesc = instr_desc(TILE, N);
  for (int k = 0; k < K; k += 16)
    mma_f16(tmem_d, smem_desc(a_addr + (uint32_t)(k >> 3) * 128u, sbo), smem_desc(b_addr + (uint32_t)(k >> 3) * 128u, sbo),
            idesc, k > 0 ? 1u : 0u);
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
// 16 consecutive accumulator columns of this thread's lane (row)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ uint4 pack8(const float* v) {
  __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
  __half2 c = __floats2half2_rn(v[4], v[5]), d = __floats2half2_rn(v[6], v[7]);
  return make_uint4(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b), *reinterpret_cast<uint32_t*>(&c),
                    *reinterpret_cast<uint32_t*>(&d));
}
// one MUFU.TANH each (max relative error 2^-11, the precision of the fp16 operands)
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoidf_(float x) { return 0.5f * tanh_fast(0.5f * x) + 0.5f; }

// generic-proxy writes to shared memory -> visible to the tensor core (async proxy), then CTA barrier
__device__ __forceinline__ void publish_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

__global__ void __launch_bounds__(THREADS, 1)
policy_step_kernel(const Weights* __restrict__ gw, long long n_rows, const float* __restrict__ x0,
                   const float* h_in, const float* c_in, const float* __restrict__ done,
                   const uint8_t* __restrict__ term, const uint8_t* __restrict__ trunc, int n_actions, uint64_t seed,
                   uint32_t step, const uint32_t* __restrict__ step_device, uint64_t row_offset,
                   const long long* __restrict__ action_in, float* h_out, float* c_out,
                   long long* __restrict__ action_out, float* __restrict__ logprob_out, float* __restrict__ entropy_out,
                   float* __restrict__ value_out, float* __restrict__ logits_out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ alignas(8) uint64_t stage_bar, mma_bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ float s_keep[TILE];   // 1 - done per row
  const int tid = threadIdx.x, warp = tid >> 5;
  const int r_ = ((warp & 3) << 5) | (tid & 31);   // this thread's row of the tile (= TMEM lane)
  const int half = warp >> 2;                       // which part (of NPART) of the columns it owns
  const Weights& W = *reinterpret_cast<const Weights*>(smem);
  const uint32_t sbase = smem_u32(smem), bar = smem_u32(&mma_bar);

  PTRACE(0);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(TM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // the 94 KB of weights travel while the first tile's activations are loaded and converted: waited for just
  // before the first MMA
  stage_blob_issue(smem, gw, (uint32_t)sizeof(Weights), &stage_bar);
  bool weights_ready = false;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);   // this warp's 32 lanes
  uint32_t parity = 0;

  const long long n_tiles = (n_rows + TILE - 1) / TILE;
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long row0 = tile * TILE, row = row0 + r_;
    const bool live = row < n_rows;
    if (half == 0)
      s_keep[r_] = !live ? 0.0f : (done != nullptr ? 1.0f - done[row] : ((term[row] | trunc[row]) ? 0.0f : 1.0f));
    // this thread's 32 cell-state values: requested now, needed after the second MMA group
    float4 cprev[H / NPART / 4];
#pragma unroll
    for (int q = 0; q < H / NPART / 4; ++q)
      cprev[q] = live ? *reinterpret_cast<const float4*>(c_in + row * H + half * (H / NPART) + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    PTRACE(1);
    // ---- x0 -> A0, (1 - done) * h -> AL[:, 64:128): coalesced float4 loads, fp16 core-matrix stores
#pragma unroll
    for (int it = 0; it < (TILE * H) / (4 * THREADS); ++it) {
      const int e = 4 * (it * THREADS + tid), r = e / H, k = e % H;
      float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), hv = xv;
      if (row0 + r < n_rows) {
        xv = __ldg(reinterpret_cast<const float4*>(x0 + (row0 + r) * H + k));
        hv = *reinterpret_cast<const float4*>(h_in + (row0 + r) * H + k);   // (h_out may alias h_in: no read-only path)
      }
      const float keep = s_keep[r];
      const __half2 x01 = __floats2half2_rn(xv.x, xv.y), x23 = __floats2half2_rn(xv.z, xv.w);
      const __half2 h01 = __floats2half2_rn(hv.x * keep, hv.y * keep), h23 = __floats2half2_rn(hv.z * keep, hv.w * keep);
      *reinterpret_cast<uint2*>(smem + SM_A0 + core_off(r, k, H)) =
          make_uint2(*reinterpret_cast<const uint32_t*>(&x01), *reinterpret_cast<const uint32_t*>(&x23));
      *reinterpret_cast<uint2*>(smem + SM_AL + core_off(r, H + k, 2 * H)) =
          make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
    }
    publish_smem();
    PTRACE(2);
    if (!weights_ready) { stage_blob_wait(&stage_bar); weights_ready = true; }
    PTRACE(3);

    // ---- trunk[2:4]: x1 = tanh(x0 W1^T + b1) -> AL[:, 0:64)
    if (tid == 0) gemm(tmem + TM_D1, sbase + SM_A0, sbase + (uint32_t)offsetof(Weights, w1), H, H, bar);
    mbar_wait(bar, parity); parity ^= 1u;
    PTRACE(4);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
    for (int jj = 0; jj < H / NPART / 16; ++jj) {
      const int jc = half * (H / NPART / 16) + jj;
      float v[16];
      tmem_ld16(trow + TM_D1 + jc * 16, v);
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = tanh_fast(v[j] + W.b1[jc * 16 + j]);
      *reinterpret_cast<uint4*>(smem + SM_AL + core_off(r_, jc * 16, 2 * H)) = pack8(v);
      *reinterpret_cast<uint4*>(smem + SM_AL + core_off(r_, jc * 16 + 8, 2 * H)) = pack8(v + 8);
    }
    publish_smem();

    PTRACE(5);
    // ---- LSTM cell: gates = [x1 | h] [W_ih | W_hh]^T + b (i, f, g, o)
    if (tid == 0) gemm(tmem + TM_G, sbase + SM_AL, sbase + (uint32_t)offsetof(Weights, wl), 2 * H, 4 * H, bar);
    mbar_wait(bar, parity); parity ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    PTRACE(6);
    {
      const float keep = s_keep[r_];
#pragma unroll
      for (int jj = 0; jj < H / NPART / 16; ++jj) {
        const int jc = half * (H / NPART / 16) + jj;
        float gi[16], gf[16], gg[16], go[16], hn[16], cn[16];
        tmem_ld16(trow + TM_G + 0 * H + jc * 16, gi);
        tmem_ld16(trow + TM_G + 1 * H + jc * 16, gf);
        tmem_ld16(trow + TM_G + 2 * H + jc * 16, gg);
        tmem_ld16(trow + TM_G + 3 * H + jc * 16, go);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 cv = cprev[jj * 4 + q];
          const float cp[4] = {cv.x * keep, cv.y * keep, cv.z * keep, cv.w * keep};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int j = 4 * q + u, col = jc * 16 + j;
            const float i_ = sigmoidf_(gi[j] + W.bl[col]), f_ = sigmoidf_(gf[j] + W.bl[H + col]);
            const float g_ = tanh_fast(gg[j] + W.bl[2 * H + col]), o_ = sigmoidf_(go[j] + W.bl[3 * H + col]);
            cn[j] = f_ * cp[u] + i_ * g_;
            hn[j] = o_ * tanh_fast(cn[j]);
          }
          if (live) {
            *reinterpret_cast<float4*>(c_out + row * H + jc * 16 + 4 * q) = make_float4(cn[4 * q], cn[4 * q + 1], cn[4 * q + 2], cn[4 * q + 3]);
            *reinterpret_cast<float4*>(h_out + row * H + jc * 16 + 4 * q) = make_float4(hn[4 * q], hn[4 * q + 1], hn[4 * q + 2], hn[4 * q + 3]);
          }
        }
        *reinterpret_cast<uint4*>(smem + SM_AH + core_off(r_, jc * 16, H)) = pack8(hn);
        *reinterpret_cast<uint4*>(smem + SM_AH + core_off(r_, jc * 16 + 8, H)) = pack8(hn + 8);
      }
    }
    publish_smem();

    PTRACE(7);
    // ---- heads, first layer: [ha | hc] = tanh(h' [Wa1; Wc1]^T + b) -> AF
    if (tid == 0) gemm(tmem + TM_HD, sbase + SM_AH, sbase + (uint32_t)offsetof(Weights, wh), H, 2 * H, bar);
    mbar_wait(bar, parity); parity ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
    for (int jj = 0; jj < 2 * H / NPART / 16; ++jj) {
      const int jc = half * (2 * H / NPART / 16) + jj;
      float v[16];
      tmem_ld16(trow + TM_HD + jc * 16, v);
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = tanh_fast(v[j] + W.bh[jc * 16 + j]);
      *reinterpret_cast<uint4*>(smem + SM_AF + core_off(r_, jc * 16, 2 * H)) = pack8(v);
      *reinterpret_cast<uint4*>(smem + SM_AF + core_off(r_, jc * 16 + 8, 2 * H)) = pack8(v + 8);
    }
    publish_smem();

    PTRACE(8);
    // ---- final layer (block diagonal): logits and value, then Categorical
    if (tid == 0) gemm(tmem + TM_F, sbase + SM_AF, sbase + (uint32_t)offsetof(Weights, wf), 2 * H, NF, bar);
    mbar_wait(bar, parity); parity ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    PTRACE(9);
    if (half == 0) {
      float v[16];
      tmem_ld16(trow + TM_F, v);
      float mx = -3.0e38f;
#pragma unroll
      for (int a = 0; a < VCOL; ++a) { v[a] += W.bf[a]; mx = a < n_actions ? fmaxf(mx, v[a]) : mx; }
      float z = 0.f;
#pragma unroll
      for (int a = 0; a < VCOL; ++a) z += a < n_actions ? __expf(v[a] - mx) : 0.f;
      const float lse = mx + __logf(z);
      float ent = 0.f;
#pragma unroll
      for (int a = 0; a < VCOL; ++a) {
        const float lp = v[a] - lse;
        ent -= a < n_actions ? __expf(lp) * lp : 0.f;
      }
      if (live) {
        long long act;
        if (action_in != nullptr) {
          act = action_in[row];
        } else {
          // inverse CDF on one Philox word per (row, step)
          const uint4 w = philox_block(seed, row_offset + (uint64_t)row, step_device != nullptr ? __ldg(step_device) : step,
                                       STREAM_POLICY);
          const float u = (float)(w.x >> 8) * (1.0f / 16777216.0f) * z;
          float acc = 0.f;
          int pick = n_actions - 1;
          for (int a = 0; a < n_actions; ++a) {   // the first action whose cumulative mass exceeds u
            acc += __expf(v[a] - mx);
            if (u < acc) { pick = a; break; }
          }
          act = pick;
        }
        float lp_act = 0.f;
#pragma unroll
        for (int a = 0; a < VCOL; ++a) lp_act = (a == (int)act) ? v[a] - lse : lp_act;
        action_out[row] = act;
        logprob_out[row] = lp_act;
        entropy_out[row] = ent;
        value_out[row] = v[VCOL] + W.bf[VCOL];
        if (logits_out != nullptr)
          for (int a = 0; a < n_actions; ++a) logits_out[row * n_actions + a] = v[a];
      }
    }
    PTRACE(10);
    // the next tile overwrites shared memory and TMEM the tensor core may still be reading: all accumulator
    // loads of this tile are complete (tcgen05.wait::ld), order them before the next tile's MMAs
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TM_COLS) : "memory");
}

}  // namespace policy
}  // namespace posg

using namespace posg;

struct posg_policy {
  policy::Weights* d_w;
  int n_actions, device;
  // posg_policy_set_trunk0 / posg_policy_eval_host: trunk layer 0 as an index-row sum, and the host staging
  float* d_w0t = nullptr;   // [in_dim, H] = trunk.0.weight transposed
  float* d_b0 = nullptr;    // [H]
  int in_dim = 0;
  unsigned char* h_stage = nullptr;   // pinned
  unsigned char* d_stage = nullptr;
  size_t stage_bytes = 0;
  cudaStream_t st = nullptr;
};

namespace posg {
namespace policy {

// x0[row] = tanh(b0 + sum_k w0t[idx[row][k]]): trunk layer 0 on a one-hot-flattened observation given as the K row
// indices of its active inputs (ppo/network.py:171-178 on FlattenProcessor's output).  One warp per row.
__global__ void embed_rows_kernel(const float* __restrict__ w0t, const float* __restrict__ b0,
                                  const int* __restrict__ idx, int n, int K, int in_dim, float* __restrict__ x0) {
  const int row = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (row >= n) return;
  float a0 = b0[lane], a1 = b0[lane + 32];
  for (int k = 0; k < K; ++k) {
    const int r = idx[row * K + k];
    if (r >= 0 && r < in_dim) { a0 += w0t[r * H + lane]; a1 += w0t[r * H + lane + 32]; }
  }
  x0[row * H + lane] = tanhf(a0);
  x0[row * H + lane + 32] = tanhf(a1);
}

// probs[row] = softmax(logits[row]) over A <= 8 actions (Categorical(logits).probs, planning/search_policy.py:270)
__global__ void softmax_rows_kernel(const float* __restrict__ logits, int n, int A, float* __restrict__ probs) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n) return;
  float m = -INFINITY;
  for (int a = 0; a < A; ++a) m = fmaxf(m, logits[row * A + a]);
  float e[VCOL], z = 0.f;
  for (int a = 0; a < A; ++a) { e[a] = expf(logits[row * A + a] - m); z += e[a]; }
  for (int a = 0; a < A; ++a) probs[row * A + a] = e[a] / z;
}

}  // namespace policy
}  // namespace posg

#ifdef POSG_POLICY_TRACE
extern "C" int posg_policy_trace(unsigned long long* out_host) {
  return (int)cudaMemcpyFromSymbol(out_host, posg::policy::g_policy_trace, 16 * 8);
}
#endif

extern "C" {

int posg_policy_create(int hidden, int n_actions, const float* trunk1_w, const float* trunk1_b, const float* w_ih,
                       const float* w_hh, const float* b_ih, const float* b_hh, const float* actor0_w,
                       const float* actor0_b, const float* actor2_w, const float* actor2_b, const float* critic0_w,
                       const float* critic0_b, const float* critic2_w, const float* critic2_b, posg_policy_t* out) {
  POSG_REQUIRE(out && trunk1_w && trunk1_b && w_ih && w_hh && b_ih && b_hh && actor0_w && actor0_b && actor2_w &&
               actor2_b && critic0_w && critic0_b && critic2_w && critic2_b, "null argument");
  if (hidden != policy::H || n_actions < 1 || n_actions > policy::VCOL) {
    set_error("posg_policy_create: the fused kernel is built for hidden size 64 and at most 8 actions");
    return POSG_E_UNSUPPORTED;
  }
  int ndev = posg_device_count();
  if (ndev < 0) return ndev;
  using namespace policy;
  std::vector<unsigned char> buf(sizeof(Weights), 0);
  Weights* w = reinterpret_cast<Weights*>(buf.data());
  auto put = [](__half* dst, int r, int k, int K, float v) {
    *reinterpret_cast<__half*>(reinterpret_cast<unsigned char*>(dst) + core_off(r, k, K)) = __float2half_rn(v);
  };
  for (int n = 0; n < H; ++n)
    for (int k = 0; k < H; ++k) put(w->w1, n, k, H, trunk1_w[n * H + k]);
  for (int n = 0; n < 4 * H; ++n)
    for (int k = 0; k < H; ++k) { put(w->wl, n, k, 2 * H, w_ih[n * H + k]); put(w->wl, n, H + k, 2 * H, w_hh[n * H + k]); }
  for (int n = 0; n < H; ++n)
    for (int k = 0; k < H; ++k) { put(w->wh, n, k, H, actor0_w[n * H + k]); put(w->wh, H + n, k, H, critic0_w[n * H + k]); }
  for (int a = 0; a < n_actions; ++a)
    for (int k = 0; k < H; ++k) put(w->wf, a, k, 2 * H, actor2_w[a * H + k]);
  for (int k = 0; k < H; ++k) put(w->wf, VCOL, H + k, 2 * H, critic2_w[k]);
  for (int n = 0; n < H; ++n) { w->b1[n] = trunk1_b[n]; w->bh[n] = actor0_b[n]; w->bh[H + n] = critic0_b[n]; }
  for (int n = 0; n < 4 * H; ++n) w->bl[n] = b_ih[n] + b_hh[n];
  for (int a = 0; a < n_actions; ++a) w->bf[a] = actor2_b[a];
  w->bf[VCOL] = critic2_b[0];
  posg_policy* p = new posg_policy();
  p->n_actions = n_actions; p->d_w = nullptr;
  cudaError_t e = cudaGetDevice(&p->device);
  if (e == cudaSuccess) e = cudaMalloc((void**)&p->d_w, sizeof(Weights));
  if (e == cudaSuccess) e = cudaMemcpy(p->d_w, w, sizeof(Weights), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(policy_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL);
  if (e != cudaSuccess) { int rc = cuda_fail(e, "policy upload"); cudaFree(p->d_w); delete p; return rc; }
  *out = p;
  return POSG_OK;
}

int posg_policy_destroy(posg_policy_t p) {
  if (!p) return POSG_OK;
  if (p->st) { cudaStreamSynchronize(p->st); cudaStreamDestroy(p->st); }
  cudaFree(p->d_w);
  cudaFree(p->d_w0t);
  cudaFree(p->d_b0);
  cudaFreeHost(p->h_stage);
  cudaFree(p->d_stage);
  delete p;
  return POSG_OK;
}

int posg_policy_set_trunk0(posg_policy_t p, int in_dim, const float* trunk0_w, const float* trunk0_b) {
  POSG_REQUIRE(p && trunk0_w && trunk0_b && in_dim > 0, "bad arguments");
  using namespace policy;
  std::vector<float> wt((size_t)in_dim * H);
  for (int n = 0; n < H; ++n)
    for (int k = 0; k < in_dim; ++k) wt[(size_t)k * H + n] = trunk0_w[(size_t)n * in_dim + k];
  POSG_CUDA(cudaSetDevice(p->device));
  cudaFree(p->d_w0t); cudaFree(p->d_b0);
  p->d_w0t = nullptr; p->d_b0 = nullptr; p->in_dim = 0;
  POSG_CUDA(cudaMalloc((void**)&p->d_w0t, wt.size() * sizeof(float)));
  POSG_CUDA(cudaMalloc((void**)&p->d_b0, H * sizeof(float)));
  POSG_CUDA(cudaMemcpy(p->d_w0t, wt.data(), wt.size() * sizeof(float), cudaMemcpyHostToDevice));
  POSG_CUDA(cudaMemcpy(p->d_b0, trunk0_b, H * sizeof(float), cudaMemcpyHostToDevice));
  p->in_dim = in_dim;
  return POSG_OK;
}

int posg_policy_eval_host(posg_policy_t p, int64_t n, int K, const int32_t* obs_index_host, const float* h_in_host,
                          const float* c_in_host, float* h_out_host, float* c_out_host, float* logits_out_host,
                          float* probs_out_host, float* value_out_host) {
  POSG_REQUIRE(p, "null policy");
  if (n == 0) return POSG_OK;
  POSG_REQUIRE(p->d_w0t, "posg_policy_set_trunk0 first");
  POSG_REQUIRE(n > 0 && n <= (1 << 16) && K > 0 && obs_index_host && h_in_host && c_in_host && h_out_host && c_out_host &&
               logits_out_host && probs_out_host && value_out_host, "bad arguments");
  using namespace policy;
  const int A = p->n_actions;
  auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
  // in:  [idx | h | c]   scratch (device only): [x0 | done | action_in | action | logprob | entropy]   out: [h | c | logits | value]
  const size_t b_idx = (size_t)n * K * 4, b_hc = (size_t)n * H * 4;
  const size_t o_h = up(b_idx), o_c = o_h + up(b_hc), o_scr = o_c + up(b_hc);
  const size_t o_x0 = o_scr, o_done = o_x0 + up(b_hc), o_ain = o_done + up((size_t)n * 4), o_act = o_ain + up((size_t)n * 8);
  const size_t o_lp = o_act + up((size_t)n * 8), o_ent = o_lp + up((size_t)n * 4), o_out = o_ent + up((size_t)n * 4);
  const size_t o_ho = o_out, o_co = o_ho + up(b_hc), o_lg = o_co + up(b_hc), o_pr = o_lg + up((size_t)n * A * 4);
  const size_t o_val = o_pr + up((size_t)n * A * 4);
  const size_t total = o_val + up((size_t)n * 4);
  POSG_CUDA(cudaSetDevice(p->device));
  if (!p->st) POSG_CUDA(cudaStreamCreateWithFlags(&p->st, cudaStreamNonBlocking));
  if (p->stage_bytes < total) {
    POSG_CUDA(cudaStreamSynchronize(p->st));
    cudaFreeHost(p->h_stage); cudaFree(p->d_stage);
    p->h_stage = nullptr; p->d_stage = nullptr; p->stage_bytes = 0;
    const size_t cap = total < 16384 ? 16384 : total + total / 2;
    POSG_CUDA(cudaMallocHost((void**)&p->h_stage, cap));
    POSG_CUDA(cudaMalloc((void**)&p->d_stage, cap));
    p->stage_bytes = cap;
  }
  unsigned char *h = p->h_stage, *d = p->d_stage;
  memcpy(h, obs_index_host, b_idx);
  memcpy(h + o_h, h_in_host, b_hc);
  memcpy(h + o_c, c_in_host, b_hc);
  POSG_CUDA(cudaMemcpyAsync(d, h, o_scr, cudaMemcpyHostToDevice, p->st));
  POSG_CUDA(cudaMemsetAsync(d + o_done, 0, o_act - o_done, p->st));        // done = 0, action_in = 0 (nothing sampled)
  const int warps_per_block = 8;
  embed_rows_kernel<<<(unsigned)((n + warps_per_block - 1) / warps_per_block), warps_per_block * 32, 0, p->st>>>(
      p->d_w0t, p->d_b0, (const int*)d, (int)n, K, p->in_dim, (float*)(d + o_x0));
  POSG_CUDA(cudaGetLastError());
  int rc = posg_policy_step(p, n, (const float*)(d + o_x0), (const float*)(d + o_h), (const float*)(d + o_c),
                            (const float*)(d + o_done), nullptr, nullptr, 0, 0, nullptr, 0, (const int64_t*)(d + o_ain),
                            (float*)(d + o_ho), (float*)(d + o_co), (int64_t*)(d + o_act), (float*)(d + o_lp),
                            (float*)(d + o_ent), (float*)(d + o_val), (float*)(d + o_lg), p->st);
  if (rc) return rc;
  softmax_rows_kernel<<<(unsigned)((n + 127) / 128), 128, 0, p->st>>>((const float*)(d + o_lg), (int)n, A, (float*)(d + o_pr));
  POSG_CUDA(cudaGetLastError());
  POSG_CUDA(cudaMemcpyAsync(h + o_out, d + o_out, total - o_out, cudaMemcpyDeviceToHost, p->st));
  POSG_CUDA(cudaStreamSynchronize(p->st));
  memcpy(h_out_host, h + o_ho, b_hc);
  memcpy(c_out_host, h + o_co, b_hc);
  memcpy(logits_out_host, h + o_lg, (size_t)n * A * 4);
  memcpy(probs_out_host, h + o_pr, (size_t)n * A * 4);
  memcpy(value_out_host, h + o_val, (size_t)n * 4);
  return POSG_OK;
}

int posg_policy_step(posg_policy_t p, int64_t n_rows, const float* x0, const float* h_in, const float* c_in,
                     const float* done, const uint8_t* terminated, const uint8_t* truncated, uint64_t seed,
                     uint32_t step, const uint32_t* step_device, uint64_t row_offset, const int64_t* action_in,
                     float* h_out, float* c_out, int64_t* action_out, float* logprob_out, float* entropy_out,
                     float* value_out, float* logits_out, void* stream) {
  POSG_REQUIRE(p, "null policy");
  if (n_rows == 0) return POSG_OK;
  POSG_REQUIRE(n_rows > 0 && x0 && h_in && c_in && (done || (terminated && truncated)) && h_out && c_out && action_out &&
               logprob_out && entropy_out && value_out, "bad arguments");
  POSG_REQUIRE((((uintptr_t)x0 | (uintptr_t)h_in | (uintptr_t)c_in | (uintptr_t)h_out | (uintptr_t)c_out) & 15u) == 0,
               "activation buffers must be 16-byte aligned");
  const long long tiles = (n_rows + policy::TILE - 1) / policy::TILE;
  const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
  policy::policy_step_kernel<<<grid, policy::THREADS, policy::SM_TOTAL, (cudaStream_t)stream>>>(
      p->d_w, (long long)n_rows, x0, h_in, c_in, done, terminated, truncated, p->n_actions, seed, step, step_device,
      row_offset,
      (const long long*)action_in, h_out, c_out, (long long*)action_out, logprob_out, entropy_out, value_out, logits_out);
  POSG_CUDA(cudaGetLastError());
  return POSG_OK;
}

}  // extern "C"
