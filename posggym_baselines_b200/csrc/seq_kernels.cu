// Sequence chunking of a rollout batch for BPTT: the reference's get_seq_idxs /
// split_and_pad_batch / split_lstm_state_batch (ppo/utils.py:240-383) walk every env in Python.  Here:
//
//  * seq_chunks_kernel -- ONE launch: a thread per env walks its column of the dones buffer, counts the chunks
//    (an episode piece starts at step 0 and at every done step; pieces are cut every max_seq_len steps), the
//    counts are scanned inside the CTA and chained across CTAs (each CTA takes a ticket, waits for the prefix of
//    the ticket before it and publishes its own), and the thread then writes its (env, start, length) rows at its
//    offset -- env-major, by start: the order of the reference's nested lists flattened;
//  * seq_gather_kernel -- ONE launch pads and gathers EVERY entry of the batch (obs, actions, log-probs, values,
//    ... up to 16 tensors with their own row sizes) into [rows, chunks, ...] with the reference's padding
//    (padding_value up to the longest chunk of the batch, zeros after it).
#include <cstring>

#include "common.h"

namespace posg {
namespace seq {

constexpr int CTA = 256;
constexpr int MAX_ENTRIES = 16;

struct ScanState {            // device scratch: 16 bytes + one 8-byte slot per CTA
  unsigned int ticket;
  unsigned int max_len;
  unsigned long long total;
};

// count (and, in the second pass, write) the chunks of one env column
template <bool WRITE>
__device__ __forceinline__ int walk(const float* __restrict__ dones, int T, long long B, long long e, int L,
                                    int* __restrict__ c_env, int* __restrict__ c_start, int* __restrict__ c_len,
                                    long long off, unsigned int* max_len) {
  int n = 0, start = 0, piece = 0;   // current chunk start, current piece start
  unsigned int longest = 0;
  for (int t = 1; t <= T; ++t) {
    const bool cut = (t == T) || (dones[(long long)t * B + e] != 0.0f) || ((t - piece) % L == 0);
    if (cut) {
      if (WRITE) { c_env[off + n] = (int)e; c_start[off + n] = start; c_len[off + n] = t - start; }
      longest = max(longest, (unsigned int)(t - start));
      ++n;
      if (t < T && dones[(long long)t * B + e] != 0.0f) piece = t;
      start = t;
    }
  }
  if (WRITE && longest) atomicMax(max_len, longest);
  return n;
}

__global__ void __launch_bounds__(CTA)
seq_chunks_kernel(const float* __restrict__ dones, int T, long long B, int L, int* __restrict__ c_env,
                  int* __restrict__ c_start, int* __restrict__ c_len, long long cap, ScanState* __restrict__ st,
                  unsigned long long* __restrict__ prefix /* [gridDim.x + 1], zeroed; bit 63 = ready */) {
  __shared__ unsigned int s_ticket;
  __shared__ long long s_warp[CTA / 32];
  __shared__ long long s_base;
  if (threadIdx.x == 0) s_ticket = atomicAdd(&st->ticket, 1u);
  __syncthreads();
  const unsigned int cta = s_ticket;                 // CTAs are numbered in the order they start running
  const long long e = (long long)cta * CTA + threadIdx.x;
  const int n = e < B ? walk<false>(dones, T, B, e, L, nullptr, nullptr, nullptr, 0, nullptr) : 0;
  // block-wide exclusive scan of the counts
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  long long incl = n;
  for (int o = 1; o < 32; o <<= 1) {
    const long long v = __shfl_up_sync(0xffffffffu, incl, o);
    if ((int)lane >= o) incl += v;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long run = 0;
    for (int w = 0; w < CTA / 32; ++w) { const long long v = s_warp[w]; s_warp[w] = run; run += v; }
    // chained scan: wait for the inclusive prefix of the CTA before this one, publish ours
    unsigned long long before = 0;
    if (cta > 0) {
      volatile unsigned long long* p = prefix + cta;
      unsigned long long v;
      while (!((v = *p) >> 63)) {}
      before = v & ~(1ULL << 63);
    }
    __threadfence();
    *((volatile unsigned long long*)(prefix + cta + 1)) = (before + (unsigned long long)run) | (1ULL << 63);
    if (cta == gridDim.x - 1) st->total = before + (unsigned long long)run;
    s_base = (long long)before;
  }
  __syncthreads();
  const long long off = s_base + s_warp[warp] + (incl - n);
  if (e < B && off + n <= cap) walk<true>(dones, T, B, e, L, c_env, c_start, c_len, off, &st->max_len);
}

struct Entry {
  const unsigned char* src;   // [T(+1), B, row_bytes]
  unsigned char* dst;         // [rows, C, row_bytes]
  long long row_bytes;
  int overlapping;            // 1: chunk + the first step after it (rows = L + 1)
  int elem_bytes;             // size of one element (for the padding pattern)
  unsigned long long pad;     // padding value's bit pattern (elem_bytes wide)
};
struct Entries { Entry e[MAX_ENTRIES]; int n; };

// one warp per (entry, row s, chunk c): copies row_bytes from src[start + s, env] or writes the padding
__global__ void __launch_bounds__(256)
seq_gather_kernel(const __grid_constant__ Entries E, const int* __restrict__ c_env, const int* __restrict__ c_start,
                  const int* __restrict__ c_len, long long C, long long B, int L, const ScanState* __restrict__ st) {
  const unsigned lane = threadIdx.x & 31u;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int longest = (int)st->max_len;
  for (int k = 0; k < E.n; ++k) {
    const Entry en = E.e[k];
    const int rows = L + en.overlapping;
    const long long items = (long long)rows * C;
    for (long long it = warp; it < items; it += nwarps) {
      const int s = (int)(it / C);
      const long long c = it - (long long)s * C;
      const int len = c_len[c] + en.overlapping;
      unsigned char* d = en.dst + it * en.row_bytes;
      if (s < len) {
        const unsigned char* src = en.src + ((long long)(c_start[c] + s) * B + c_env[c]) * en.row_bytes;
        if (((en.row_bytes | (uintptr_t)src | (uintptr_t)d) & 15) == 0) {
          for (long long b = (long long)lane * 16; b < en.row_bytes; b += 32 * 16)
            *reinterpret_cast<uint4*>(d + b) = *reinterpret_cast<const uint4*>(src + b);
        } else if (((en.row_bytes | (uintptr_t)src | (uintptr_t)d) & 3) == 0) {
          for (long long b = (long long)lane * 4; b < en.row_bytes; b += 32 * 4)
            *reinterpret_cast<uint32_t*>(d + b) = *reinterpret_cast<const uint32_t*>(src + b);
        } else {
          for (long long b = lane; b < en.row_bytes; b += 32) d[b] = src[b];
        }
      } else {
        // pad_sequence pads with padding_value up to the longest chunk, F.pad with zeros after it
        const unsigned long long pat = s < longest + en.overlapping ? en.pad : 0ULL;
        for (long long b = lane; b < en.row_bytes; b += 32) d[b] = (unsigned char)(pat >> (8 * (b % en.elem_bytes)));
      }
    }
  }
}

}  // namespace seq
}  // namespace posg

using namespace posg;

extern "C" {

int posg_seq_chunks(int T, int64_t B, const float* dones, int max_seq_len, int32_t* chunk_env, int32_t* chunk_start,
                    int32_t* chunk_len, int64_t capacity, int64_t* result_host, void* stream) {
  POSG_REQUIRE(T >= 1 && B >= 0 && max_seq_len >= 1 && result_host, "bad arguments");
  result_host[0] = 0; result_host[1] = 0;
  if (B == 0) return POSG_OK;
  POSG_REQUIRE(dones && chunk_env && chunk_start && chunk_len && capacity >= 0, "bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  const int grid = (int)((B + seq::CTA - 1) / seq::CTA);
  unsigned char* scratch = nullptr;
  const size_t bytes = sizeof(seq::ScanState) + (size_t)(grid + 1) * 8;
  POSG_CUDA(cudaMallocAsync((void**)&scratch, bytes, s));
  POSG_CUDA(cudaMemsetAsync(scratch, 0, bytes, s));
  seq::ScanState* st = (seq::ScanState*)scratch;
  unsigned long long* prefix = (unsigned long long*)(scratch + sizeof(seq::ScanState));
  seq::seq_chunks_kernel<<<grid, seq::CTA, 0, s>>>(dones, T, (long long)B, max_seq_len, chunk_env, chunk_start, chunk_len,
                                                   (long long)capacity, st, prefix);
  POSG_CUDA(cudaGetLastError());
  seq::ScanState h;
  POSG_CUDA(cudaMemcpyAsync(&h, st, sizeof(h), cudaMemcpyDeviceToHost, s));
  POSG_CUDA(cudaStreamSynchronize(s));
  cudaFreeAsync(scratch, s);
  result_host[0] = (int64_t)h.total;
  result_host[1] = (int64_t)h.max_len;
  if ((int64_t)h.total > capacity) { set_error("posg_seq_chunks: %lld chunks exceed the capacity %lld", (long long)h.total, (long long)capacity); return POSG_E_INVALID; }
  return POSG_OK;
}

int posg_seq_gather(int n_entries, const void* const* src, void* const* dst, const int64_t* row_bytes,
                    const int32_t* overlapping, const int32_t* elem_bytes, const uint64_t* pad_bits,
                    const int32_t* chunk_env, const int32_t* chunk_start, const int32_t* chunk_len, int64_t n_chunks,
                    int64_t B, int max_seq_len, int longest, void* stream) {
  POSG_REQUIRE(n_entries >= 0 && n_entries <= seq::MAX_ENTRIES, "at most 16 batch entries per call");
  if (n_entries == 0 || n_chunks == 0) return POSG_OK;
  POSG_REQUIRE(src && dst && row_bytes && overlapping && elem_bytes && pad_bits && chunk_env && chunk_start && chunk_len,
               "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  seq::Entries E;
  E.n = n_entries;
  for (int k = 0; k < n_entries; ++k) {
    POSG_REQUIRE(elem_bytes[k] >= 1 && elem_bytes[k] <= 8 && row_bytes[k] % elem_bytes[k] == 0, "bad element size");
    E.e[k] = {(const unsigned char*)src[k], (unsigned char*)dst[k], (long long)row_bytes[k], overlapping[k], elem_bytes[k],
              (unsigned long long)pad_bits[k]};
  }
  seq::ScanState h{};
  h.max_len = (unsigned int)longest;
  seq::ScanState* d_st = nullptr;
  POSG_CUDA(cudaMallocAsync((void**)&d_st, sizeof(h), s));
  POSG_CUDA(cudaMemcpyAsync(d_st, &h, sizeof(h), cudaMemcpyHostToDevice, s));
  seq::seq_gather_kernel<<<sm_count() * 8, 256, 0, s>>>(E, chunk_env, chunk_start, chunk_len, (long long)n_chunks,
                                                        (long long)B, max_seq_len, d_st);
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(d_st, s);
  POSG_CUDA(e);
  return POSG_OK;
}

}  // extern "C"
