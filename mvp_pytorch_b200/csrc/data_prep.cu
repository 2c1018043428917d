// SURVEY f-2: the input pipeline's per-sample byte / integer work on the GPU.
//
// (1) Region-feature decode.  The reference's loaders keep VinVL features as base64 text in TSV files and decode them
//     per sample on DataLoader workers: np.frombuffer(base64.b64decode(arr[-1]), dtype=np.float32).reshape(num_boxes,
//     img_feature_dim) -> torch.tensor(..., dtype=args.dtype) (oscar_datasets_ml/oscar_tsv4.py:696-727; same in
//     run_retrieval.py / run_vqa.py datasets), then zero-pad to max_img_seq_length.  Here the raw base64 bytes of a
//     whole batch are copied to the device and ONE kernel produces the padded [B, R, K] feature tensor in the model
//     dtype: 16 base64 characters = 12 bytes = 3 floats per thread (the two alignments meet every 12 bytes).  Pure
//     byte / integer work, bit-exact; HBM bound (4/3 bytes read + 2 or 4 written per float).
// (2) BERT / phrase masking (random_word :782-820, random_phrases :822-850) on token ids: 15 % of the positions,
//     of which 80 % -> [MASK], 10 % -> random id, 10 % kept; a phrase whose linked caption token was selected is
//     masked too (phrase_mask_map).  The random numbers are either supplied (bit-exact replay of recorded draws,
//     tests) or drawn from the stateless hash of common.cuh.
#include "common.cuh"

namespace mvptr {

__constant__ uint8_t kB64[256];
static uint8_t g_b64_host[256];
static bool g_b64_ready[16] = {false};

static int ensure_b64_table() {
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 15;
  if (g_b64_ready[dev]) return 0;
  for (int i = 0; i < 256; ++i) g_b64_host[i] = 0xFF;
  const char* abc = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
  for (int i = 0; i < 64; ++i) g_b64_host[(uint8_t)abc[i]] = (uint8_t)i;
  g_b64_host[(uint8_t)'='] = 0xFE;  // padding: decodes as zero bits, only legal in the last two characters
  cudaError_t e = cudaMemcpyToSymbol(kB64, g_b64_host, 256);
  if (e != cudaSuccess) MVPTR_FAIL(MVPTR_ERR_CUDA, "b64 table: %s", cudaGetErrorString(e));
  g_b64_ready[dev] = true;
  return 0;
}

// grid: (ceil(groups / 256), B); one thread = 16 characters -> 3 floats
template <typename T>
__global__ void __launch_bounds__(256)
b64_decode_kernel(const uint8_t* __restrict__ src, const int64_t* __restrict__ offsets, const int32_t* __restrict__ num_boxes,
                  T* __restrict__ dst, int R, int K, int ld_dst, int* __restrict__ error) {
  const int b = blockIdx.y;
  const int nb = num_boxes[b];
  const long long n_float = (long long)nb * K;
  const long long n_char = offsets[b + 1] - offsets[b];
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // group of 16 characters
  if (g * 3 >= n_float) return;
  // every sample's text must hold exactly ceil(4 n / 3) * 4 ... characters for n = 4 * n_float bytes
  if (g == 0 && n_char != ((4 * n_float + 2) / 3) * 4) atomicOr(error, 1);
  const uint8_t* s = src + offsets[b] + g * 16;
  uint32_t w[3] = {0, 0, 0};  // 12 decoded bytes, little endian floats
  uint8_t bytes[12];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t v = 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const long long pos = g * 16 + q * 4 + c;
      uint8_t d = 0;
      if (pos < n_char) {
        d = kB64[__ldg(s + q * 4 + c)];
        if (d == 0xFF) atomicOr(error, 2);                        // not a base64 character
        if (d == 0xFE) { if (pos < n_char - 2) atomicOr(error, 4); d = 0; }  // '=' before the tail
      }
      v = (v << 6) | d;
    }
    bytes[q * 3] = (uint8_t)(v >> 16);
    bytes[q * 3 + 1] = (uint8_t)(v >> 8);
    bytes[q * 3 + 2] = (uint8_t)v;
  }
#pragma unroll
  for (int i = 0; i < 3; ++i)
    w[i] = (uint32_t)bytes[4 * i] | ((uint32_t)bytes[4 * i + 1] << 8) | ((uint32_t)bytes[4 * i + 2] << 16) |
           ((uint32_t)bytes[4 * i + 3] << 24);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const long long e = g * 3 + i;
    if (e < n_float) {
      const int row = (int)(e / K), col = (int)(e - (long long)row * K);
      if (row < R) {
        const float f = __uint_as_float(w[i]);
        T* o = dst + ((size_t)b * R + row) * ld_dst + col;
        if constexpr (sizeof(T) == 4) *o = f;
        else *o = __float2bfloat16(f);
      }
    }
  }
}
// zero rows >= num_boxes and the pitch padding columns
template <typename T>
__global__ void pad_rows_kernel(const int32_t* __restrict__ num_boxes, T* __restrict__ dst, int R, int K, int ld_dst) {
  const int b = blockIdx.y, row = blockIdx.x;
  const int nb = min(num_boxes[b], R);
  T* o = dst + ((size_t)b * R + row) * ld_dst;
  const int c0 = row < nb ? K : 0;
  for (int c = c0 + threadIdx.x; c < ld_dst; c += blockDim.x) {
    if constexpr (sizeof(T) == 4) o[c] = 0.f;
    else o[c] = __float2bfloat16(0.f);
  }
}

// One CTA per sample.  Positions [tok_first, tok_first + tok_count) are caption / tag tokens (random_word), positions
// [phr_first, phr_first + phr_count) phrase concepts (random_phrases).  links[b, i, :] lists the phrase indexes tied
// to caption token i (phrase_mask_map), -1 padded.
__global__ void __launch_bounds__(128)
mlm_mask_kernel(int64_t* __restrict__ ids, int64_t* __restrict__ labels, const int32_t* __restrict__ tok_first,
                const int32_t* __restrict__ tok_count, const int32_t* __restrict__ phr_first,
                const int32_t* __restrict__ phr_count, const int32_t* __restrict__ links, int max_links,
                const float* __restrict__ u, const int64_t* __restrict__ r, int L, long long mask_id, long long word_vocab,
                long long phrase_vocab, long long vocab_size, uint32_t seed) {
  __shared__ int forced[128];
  const int b = blockIdx.x;
  const int t0 = tok_first[b], tn = tok_count[b];
  const int p0 = phr_first ? phr_first[b] : 0, pn = phr_count ? phr_count[b] : 0;
  for (int i = threadIdx.x; i < 128; i += blockDim.x) forced[i] = 0;
  for (int i = threadIdx.x; i < L; i += blockDim.x) labels[(size_t)b * L + i] = -1;
  __syncthreads();
  auto uniform = [&](int pos) -> float {
    if (u) return u[(size_t)b * L + pos];
    return (hash32(seed + (uint32_t)(b * L + pos) * 0x9E3779B9u) >> 8) * (1.0f / 16777216.0f);
  };
  auto draw = [&](int pos, long long n) -> long long {
    if (r) return r[(size_t)b * L + pos] % n;
    return (long long)(hash32(seed * 0x85EBCA6Bu + (uint32_t)(b * L + pos) * 0xC2B2AE35u + 1u) % (uint32_t)n);
  };
  for (int i = threadIdx.x; i < tn; i += blockDim.x) {  // random_word, oscar_tsv4.py:782-820
    const int pos = t0 + i;
    double prob = (double)uniform(pos);  // the reference's arithmetic is Python double precision
    if (prob < 0.15) {
      prob /= 0.15;
      const long long orig = ids[(size_t)b * L + pos];
      if (prob < 0.8) ids[(size_t)b * L + pos] = mask_id;
      else if (prob < 0.9) ids[(size_t)b * L + pos] = draw(pos, word_vocab);
      labels[(size_t)b * L + pos] = orig;
      if (links)
        for (int k = 0; k < max_links; ++k) {
          const int ph = links[((size_t)b * L + i) * max_links + k];
          if (ph >= 0 && ph < 128) forced[ph] = 1;
        }
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < pn; j += blockDim.x) {  // random_phrases, :822-850 (its labels are dropped, :960)
    const int pos = p0 + j;
    if (j < 128 && forced[j]) {
      ids[(size_t)b * L + pos] = mask_id;
    } else {
      double prob = (double)uniform(pos);
      if (prob < 0.15) {
        prob /= 0.15;
        if (prob < 0.8) ids[(size_t)b * L + pos] = mask_id;
        else if (prob < 0.9) ids[(size_t)b * L + pos] = draw(pos, phrase_vocab) + vocab_size;
      }
    }
  }
}

}  // namespace mvptr

using namespace mvptr;

extern "C" int mvptr_b64_decode_features(const void* src, const int64_t* offsets, const int32_t* num_boxes, void* dst,
                                         int dst_is_f32, int B, int R, int K, int ld_dst, int max_boxes, int* error_flag,
                                         void* stream) {
  if (B <= 0) return 0;
  if (!src || !offsets || !num_boxes || !dst || !error_flag) MVPTR_FAIL(MVPTR_ERR_ARG, "b64_decode_features: null argument");
  if (ld_dst < K || R <= 0 || K <= 0 || max_boxes <= 0) MVPTR_FAIL(MVPTR_ERR_ARG, "b64_decode_features: bad shape");
  if (int rc = ensure_b64_table()) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const long long groups = ((long long)max_boxes * K + 2) / 3;
  dim3 grid((unsigned)((groups + 255) / 256), B);
  if (dst_is_f32) {
    pad_rows_kernel<float><<<dim3(R, B), 128, 0, s>>>(num_boxes, (float*)dst, R, K, ld_dst);
    b64_decode_kernel<float><<<grid, 256, 0, s>>>((const uint8_t*)src, offsets, num_boxes, (float*)dst, R, K, ld_dst, error_flag);
  } else {
    pad_rows_kernel<bf16><<<dim3(R, B), 128, 0, s>>>(num_boxes, (bf16*)dst, R, K, ld_dst);
    b64_decode_kernel<bf16><<<grid, 256, 0, s>>>((const uint8_t*)src, offsets, num_boxes, (bf16*)dst, R, K, ld_dst, error_flag);
  }
  ++g_launch_count;
  MVPTR_CHECK_LAUNCH("b64_decode_features");
  return 0;
}

extern "C" int mvptr_mlm_mask(int64_t* ids, int64_t* labels, const int32_t* tok_first, const int32_t* tok_count,
                              const int32_t* phr_first, const int32_t* phr_count, const int32_t* links, int max_links,
                              const float* u, const int64_t* r, int B, int L, long long mask_id, long long word_vocab,
                              long long phrase_vocab, long long vocab_size, uint32_t seed, void* stream) {
  if (B <= 0) return 0;
  if (!ids || !labels || !tok_first || !tok_count) MVPTR_FAIL(MVPTR_ERR_ARG, "mlm_mask: null argument");
  if (word_vocab <= 0 || (phr_count && phrase_vocab <= 0)) MVPTR_FAIL(MVPTR_ERR_ARG, "mlm_mask: vocabulary sizes must be positive");
  mlm_mask_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(ids, labels, tok_first, tok_count, phr_first, phr_count, links,
                                                       max_links, u, r, L, mask_id, word_vocab, phrase_vocab, vocab_size, seed);
  MVPTR_CHECK_LAUNCH("mlm_mask");
  return 0;
}
