// Per-row top-k with the reference's ranking order.
// run_retrieval.py:487,506 rank with np.argsort(sim)[::-1]; made deterministic as: descending score,
// ties broken by DESCENDING index (the reversal of a stable ascending sort), see oracle topk_desc.
// One CTA per row: the row is staged in shared memory as order-preserving 32-bit keys, then k
// block-wide arg-max passes over (key << 32 | index) select the winners in order.  For the COCO-5k
// shapes (25 000 scores, k=128 / 5 000 scores, k=64) this is ~0.5 ms per orientation on one B200,
// against minutes of host-side numpy argsort loops in the reference.
#include "common.cuh"

namespace mvptr {

__device__ __forceinline__ uint32_t float_key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void __launch_bounds__(1024)
topk_rows_kernel(const float* __restrict__ x, long long ld, int n, int k, int64_t* __restrict__ idx_out,
                 float* __restrict__ val_out) {
  extern __shared__ uint32_t keys[];
  __shared__ unsigned long long red[32];
  __shared__ unsigned long long winner;
  const float* row = x + (size_t)blockIdx.x * ld;
  for (int i = threadIdx.x; i < n; i += blockDim.x) keys[i] = float_key(row[i]);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int t = 0; t < k; ++t) {
    unsigned long long best = 0ull;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const unsigned long long c = ((unsigned long long)keys[i] << 32) | (unsigned)i;
      best = c > best ? c : best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
      best = other > best ? other : best;
    }
    if (lane == 0) red[warp] = best;
    __syncthreads();
    if (warp == 0) {
      unsigned long long b = lane < nw ? red[lane] : 0ull;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, b, o);
        b = other > b ? other : b;
      }
      if (lane == 0) {
        winner = b;
        const int wi = (int)(b & 0xffffffffu);
        idx_out[(size_t)blockIdx.x * k + t] = wi;
        if (val_out) val_out[(size_t)blockIdx.x * k + t] = row[wi];
        keys[wi] = 0u;  // below every real key
      }
    }
    __syncthreads();
  }
}

// prob[i] = softmax(logits[i, :])[1]  (run_retrieval.py:776-777, 818-820), C == 2
__global__ void match_prob_kernel(const float* __restrict__ logits, float* __restrict__ prob, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float a = logits[2 * i], b = logits[2 * i + 1];
  const float m = fmaxf(a, b);
  const float ea = __expf(a - m), eb = __expf(b - m);
  prob[i] = eb / (ea + eb);
}

}  // namespace mvptr

using namespace mvptr;

extern "C" int mvptr_topk_rows(const float* x, long long ld, int rows, int n, int k, int64_t* idx_out, float* val_out,
                               void* stream) {
  MVPTR_PROF("topk_rows", 4.0 * rows * n, stream);
  if (rows <= 0) return 0;
  if (k <= 0 || k > n) MVPTR_FAIL(MVPTR_ERR_ARG, "topk: need 0 < k <= n (k=%d n=%d)", k, n);
  const int smem = n * 4;
  if (smem > 200 * 1024) MVPTR_FAIL(MVPTR_ERR_UNSUPPORTED, "topk: row of %d scores exceeds the shared-memory stage", n);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(topk_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) MVPTR_FAIL(MVPTR_ERR_CUDA, "topk smem attr: %s", cudaGetErrorString(e));
    configured = true;
  }
  const int threads = n >= 4096 ? 1024 : (n >= 512 ? 256 : 64);
  topk_rows_kernel<<<rows, threads, smem, (cudaStream_t)stream>>>(x, ld, n, k, idx_out, val_out);
  MVPTR_CHECK_LAUNCH("topk_rows");
  return 0;
}

extern "C" int mvptr_match_prob(const float* logits, float* prob, int n, void* stream) {
  if (n <= 0) return 0;
  match_prob_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(logits, prob, n);
  MVPTR_CHECK_LAUNCH("match_prob");
  return 0;
}
