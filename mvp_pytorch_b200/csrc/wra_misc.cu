// Weakly-supervised phrase grounding (WRA) as one batched kernel, plus the VQA BCE loss and
// the GELU-backward elementwise kernel of the head transforms.
//
// WRA replaces the per-sample Python loops of modeling_vlbert.py:1288-1300 with helpers
// mask_slice_and_stack :1502-1508, t2i_sim :1543-1550, get_pos_neg_sims :1553-1596
// (O(B) host iterations, ~10 launches each).  The host still draws the random choices
// (negative image per sample, one of the top-3 per phrase) so the RNG stream order stays
// the reference's; the kernel consumes them as index tensors.
#include "common.cuh"

namespace mvptr {

constexpr int kMaxPhrases = 16;
constexpr int kMaxRegions = 128;

// fp32 rows (verification tier, csrc/fp32_tier.cu): same kernels instantiated on float storage
__device__ __forceinline__ float row_dot(const float* a, const float* b, int H, int lane) {
  float s = 0.f;
  for (int i = lane; i < H; i += 32) s = fmaf(a[i], b[i], s);
  return warp_sum(s);
}
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(bf16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float row_dot(const bf16* a, const bf16* b, int H, int lane) {
  float s = 0.f;
  for (int ch = lane; ch < (H >> 3); ch += 32) {
    float x[8], y[8];
    unpack8(*reinterpret_cast<const bf16x8*>(a + ch * 8), x);
    unpack8(*reinterpret_cast<const bf16x8*>(b + ch * 8), y);
#pragma unroll
    for (int j = 0; j < 8; ++j) s += x[j] * y[j];
  }
  return warp_sum(s);
}

// one CTA per sample; outputs pos_sim[b], neg_sim[b] and the selected (phrase -> region token) pairs
template <typename T>
__global__ void __launch_bounds__(256)
wra_fwd_kernel(const T* __restrict__ seq, int Ltot, int H, const int64_t* __restrict__ phrase_index,
               const int64_t* __restrict__ img_index, const int64_t* __restrict__ neg_img,
               const int64_t* __restrict__ rand_pos, const int64_t* __restrict__ rand_neg, int P,
               float* __restrict__ pos_out, float* __restrict__ neg_out, int* __restrict__ sel_pos,
               int* __restrict__ sel_neg) {
  __shared__ float sims[2][kMaxPhrases][kMaxRegions];
  __shared__ float nph[kMaxPhrases], nreg[2][kMaxRegions];
  __shared__ float picked[2][kMaxPhrases];
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int p0 = (int)phrase_index[2 * b], p1 = (int)phrase_index[2 * b + 1];
  const int n_ph = min(max(p1 - p0, 0), kMaxPhrases);
  const int nb = (int)neg_img[b];
  const int r0[2] = {(int)img_index[2 * b], (int)img_index[2 * nb]};
  const int nr[2] = {min(max((int)img_index[2 * b + 1] - r0[0], 0), kMaxRegions),
                     min(max((int)img_index[2 * nb + 1] - r0[1], 0), kMaxRegions)};
  const T* own = seq + (size_t)b * Ltot * H;
  const T* oth[2] = {own, seq + (size_t)nb * Ltot * H};
  if (n_ph == 0) {  // t2i_sim of an empty phrase set is 0 (modeling_vlbert.py:1544-1545)
    if (threadIdx.x == 0) pos_out[b] = neg_out[b] = 0.f;
    return;
  }
  for (int i = warp; i < n_ph; i += nw) {
    const float s = row_dot(own + (size_t)(p0 + i) * H, own + (size_t)(p0 + i) * H, H, lane);
    if (lane == 0) nph[i] = fmaxf(sqrtf(s), 1e-12f);
  }
  for (int w = 0; w < 2; ++w)
    for (int i = warp; i < nr[w]; i += nw) {
      const T* r = oth[w] + (size_t)(r0[w] + i) * H;
      const float s = row_dot(r, r, H, lane);
      if (lane == 0) nreg[w][i] = fmaxf(sqrtf(s), 1e-12f);
    }
  __syncthreads();
  for (int w = 0; w < 2; ++w)
    for (int i = warp; i < n_ph * nr[w]; i += nw) {
      const int ph = i / nr[w], rg = i - ph * nr[w];
      const float s = row_dot(own + (size_t)(p0 + ph) * H, oth[w] + (size_t)(r0[w] + rg) * H, H, lane);
      if (lane == 0) sims[w][ph][rg] = s / (nph[ph] * nreg[w][rg]);
    }
  __syncthreads();
  // per phrase: top-3 regions, keep the rand-th (torch.topk(3) then f_sim[arange, rand_index])
  if (threadIdx.x < 2 * n_ph) {
    const int w = threadIdx.x / n_ph, ph = threadIdx.x - w * n_ph;
    float v[3] = {-INFINITY, -INFINITY, -INFINITY};
    int ix[3] = {-1, -1, -1};
    for (int r = 0; r < nr[w]; ++r) {
      const float s = sims[w][ph][r];
      if (s > v[0]) { v[2] = v[1]; ix[2] = ix[1]; v[1] = v[0]; ix[1] = ix[0]; v[0] = s; ix[0] = r; }
      else if (s > v[1]) { v[2] = v[1]; ix[2] = ix[1]; v[1] = s; ix[1] = r; }
      else if (s > v[2]) { v[2] = s; ix[2] = r; }
    }
    int k = (int)(w == 0 ? rand_pos : rand_neg)[(size_t)b * P + ph];
    k = k < 0 ? 0 : (k > 2 ? 2 : k);
    picked[w][ph] = v[k];
    (w == 0 ? sel_pos : sel_neg)[(size_t)b * kMaxPhrases + ph] = ix[k] >= 0 ? r0[w] + ix[k] : -1;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    float s = 0.f;
    for (int i = 0; i < n_ph; ++i) s += picked[threadIdx.x][i];
    (threadIdx.x == 0 ? pos_out : neg_out)[b] = s / n_ph;
  }
}

// dseq (fp32, atomics) += d pos_sim / d neg_sim through the selected cosine similarities
template <typename T>
__global__ void __launch_bounds__(256)
wra_bwd_kernel(const T* __restrict__ seq, int Ltot, int H, const int64_t* __restrict__ phrase_index,
               const int64_t* __restrict__ neg_img, const int* __restrict__ sel_pos, const int* __restrict__ sel_neg,
               const float* __restrict__ dpos, const float* __restrict__ dneg, float* __restrict__ dseq) {
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int p0 = (int)phrase_index[2 * b], p1 = (int)phrase_index[2 * b + 1];
  const int n_ph = min(max(p1 - p0, 0), kMaxPhrases);
  if (n_ph == 0) return;
  const int nb = (int)neg_img[b];
  for (int item = warp; item < 2 * n_ph; item += nw) {
    const int w = item / n_ph, ph = item - w * n_ph;
    const int tok = (w == 0 ? sel_pos : sel_neg)[(size_t)b * kMaxPhrases + ph];
    if (tok < 0) continue;
    const float g = (w == 0 ? dpos[b] : dneg[b]) / n_ph;
    if (g == 0.f) continue;
    const size_t urow = ((size_t)b * Ltot + p0 + ph) * H;
    const size_t vrow = ((size_t)(w == 0 ? b : nb) * Ltot + tok) * H;
    const T* u = seq + urow;
    const T* v = seq + vrow;
    const float nu = fmaxf(sqrtf(row_dot(u, u, H, lane)), 1e-12f);
    const float nv = fmaxf(sqrtf(row_dot(v, v, H, lane)), 1e-12f);
    const float s = row_dot(u, v, H, lane) / (nu * nv);
    for (int i = lane; i < H; i += 32) {
      const float uh = to_f32(u[i]) / nu, vh = to_f32(v[i]) / nv;
      atomicAdd(dseq + urow + i, g * (vh - s * uh) / nu);
      atomicAdd(dseq + vrow + i, g * (uh - s * vh) / nv);
    }
  }
}

// dx = dy * gelu'(pre)    (backward of BertPredictionHeadTransform's activation, modeling_bert.py:489)
__global__ void gelu_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ pre, bf16* __restrict__ dx, size_t n) {
  size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  const size_t stride = (size_t)gridDim.x * blockDim.x * 8;
  for (; i + 8 <= n; i += stride) {
    float a[8], x[8];
    unpack8(*reinterpret_cast<const bf16x8*>(dy + i), a);
    unpack8(*reinterpret_cast<const bf16x8*>(pre + i), x);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] *= gelu_erf_grad(x[j]);
    *reinterpret_cast<bf16x8*>(dx + i) = pack8(a);
  }
}

// instance_bce_with_logits (modeling_vlbert.py:878-883): mean BCE * C == sum / n
__global__ void __launch_bounds__(256)
bce_fwd_kernel(const float* __restrict__ logits, int ld, const float* __restrict__ labels, int n, int C,
               float* __restrict__ loss) {
  __shared__ float red[8];
  const int r = blockIdx.x;
  float s = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float x = logits[(size_t)r * ld + c], y = labels[(size_t)r * C + c];
    s += fmaxf(x, 0.f) - x * y + log1pf(__expf(-fabsf(x)));
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    atomicAdd(loss, t / n);
  }
}
__global__ void __launch_bounds__(256)
bce_bwd_kernel(const float* __restrict__ logits, int ld, const float* __restrict__ labels, int n, int C,
               const float* __restrict__ gscale, bf16* __restrict__ dlogits, int ld_d) {
  const int r = blockIdx.x;
  const float g = (gscale ? *gscale : 1.f) / n;
  for (int c = threadIdx.x; c < ld_d; c += blockDim.x) {
    float d = 0.f;
    if (c < C) {
      const float x = logits[(size_t)r * ld + c];
      d = (1.f / (1.f + __expf(-x)) - labels[(size_t)r * C + c]) * g;
    }
    dlogits[(size_t)r * ld_d + c] = __float2bfloat16(d);
  }
}

// ---------------------------------------------------------------------------------
// Referring-expression head (BiImageBertForRE, modeling_vlbert.py:1936-1956): score of every region
// token against the [CLS] token of its own sequence -- cosine similarity (mod 1) or raw dot product
// (mod 2).  seq [B, Ltot, H] bf16 (an optional dropout between encoder and head is applied in place on
// the fly, same hash as the row kernels); regions are tokens [first, first + R).
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void load_row_dropped(const bf16* row, int H, int lane, uint32_t seed, uint32_t base,
                                                 uint32_t keep_thr, float inv_keep, float (&x)[4][8]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int ch = lane + 32 * c;
    if (ch < (H >> 3)) {
      unpack8(*reinterpret_cast<const bf16x8*>(row + ch * 8), x[c]);
      if (keep_thr != 0xffffffffu) dropout8(x[c], seed, base + ch * 8, keep_thr, inv_keep);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) x[c][j] = 0.f;
    }
  }
}
// one warp per (b, region): logits[b, j], plus 1/|cls|, 1/|region| for backward
__global__ void __launch_bounds__(128)
cls_region_score_fwd_kernel(const bf16* __restrict__ seq, int B, int Ltot, int H, int first, int R, int normalize,
                            float* __restrict__ logits, float* __restrict__ inv_norm, uint32_t keep_thr,
                            float inv_keep, uint32_t seed_) {
  const uint32_t seed = site_seed(seed_);
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (w >= B * R) return;
  const int b = w / R, j = w - b * R;
  const size_t rc = (size_t)b * Ltot, rr = rc + first + j;
  float c[4][8], r[4][8];
  load_row_dropped(seq + rc * H, H, lane, seed, (uint32_t)rc * (uint32_t)H, keep_thr, inv_keep, c);
  load_row_dropped(seq + rr * H, H, lane, seed, (uint32_t)rr * (uint32_t)H, keep_thr, inv_keep, r);
  float dot = 0.f, cc = 0.f, r2 = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      dot += c[k][i] * r[k][i];
      cc += c[k][i] * c[k][i];
      r2 += r[k][i] * r[k][i];
    }
  dot = warp_sum(dot); cc = warp_sum(cc); r2 = warp_sum(r2);
  if (lane == 0) {
    float ic = 1.f, ir = 1.f;
    if (normalize) {  // F.normalize: x / max(|x|, 1e-12)
      ic = 1.f / fmaxf(sqrtf(cc), 1e-12f);
      ir = 1.f / fmaxf(sqrtf(r2), 1e-12f);
    }
    logits[w] = dot * ic * ir;
    inv_norm[2 * w] = ic;
    inv_norm[2 * w + 1] = ir;
  }
}
// one CTA per sample: dseq rows of the regions and of [CLS] (sum over regions); other rows stay zero
__global__ void __launch_bounds__(256)
cls_region_score_bwd_kernel(const bf16* __restrict__ seq, int B, int Ltot, int H, int first, int R, int normalize,
                            const float* __restrict__ logits, const float* __restrict__ inv_norm,
                            const float* __restrict__ dlogits, bf16* __restrict__ dseq, uint32_t keep_thr,
                            float inv_keep, uint32_t seed_) {
  extern __shared__ float dcls[];  // [8 warps][H]
  const uint32_t seed = site_seed(seed_);
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t rc = (size_t)b * Ltot;
  float c[4][8], acc[4][8];
  load_row_dropped(seq + rc * H, H, lane, seed, (uint32_t)rc * (uint32_t)H, keep_thr, inv_keep, c);
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[k][i] = 0.f;
  for (int j = warp; j < R; j += 8) {
    const int w = b * R + j;
    const size_t rr = rc + first + j;
    const float g = dlogits[w], s = logits[w], ic = inv_norm[2 * w], ir = inv_norm[2 * w + 1];
    float r[4][8];
    load_row_dropped(seq + rr * H, H, lane, seed, (uint32_t)rr * (uint32_t)H, keep_thr, inv_keep, r);
    // s = <c, r> ic ir ;  ds/dr = ic ir c - (normalize ? s ir^2 r : 0) ;  ds/dc symmetric
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int ch = lane + 32 * k;
      if (ch < (H >> 3)) {
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          o[i] = g * (ic * ir * c[k][i] - (normalize ? s * ir * ir * r[k][i] : 0.f));
          acc[k][i] += g * (ic * ir * r[k][i] - (normalize ? s * ic * ic * c[k][i] : 0.f));
        }
        if (keep_thr != 0xffffffffu) dropout8(o, seed, (uint32_t)rr * (uint32_t)H + ch * 8, keep_thr, inv_keep);
        *reinterpret_cast<bf16x8*>(dseq + rr * H + ch * 8) = pack8(o);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int ch = lane + 32 * k;
    if (ch < (H >> 3))
#pragma unroll
      for (int i = 0; i < 8; ++i) dcls[warp * H + ch * 8 + i] = acc[k][i];
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < (H >> 3); ch += blockDim.x) {
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float t = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) t += dcls[w8 * H + ch * 8 + i];
      o[i] = t;
    }
    if (keep_thr != 0xffffffffu) dropout8(o, seed, (uint32_t)rc * (uint32_t)H + ch * 8, keep_thr, inv_keep);
    *reinterpret_cast<bf16x8*>(dseq + rc * H + ch * 8) = pack8(o);
  }
}

}  // namespace mvptr

using namespace mvptr;

extern "C" int mvptr_wra_max_phrases(void) { return kMaxPhrases; }

extern "C" int mvptr_wra_fwd(const void* seq, int B, int Ltot, int H, const int64_t* phrase_index,
                             const int64_t* img_index, const int64_t* neg_img, const int64_t* rand_pos,
                             const int64_t* rand_neg, int P, float* pos_out, float* neg_out, int* sel_pos,
                             int* sel_neg, void* stream) {
  if (B <= 0) return 0;
  if (H & 7) MVPTR_FAIL(MVPTR_ERR_ARG, "wra: H must be a multiple of 8");
  wra_fwd_kernel<bf16><<<B, 256, 0, (cudaStream_t)stream>>>((const bf16*)seq, Ltot, H, phrase_index, img_index, neg_img,
                                                            rand_pos, rand_neg, P, pos_out, neg_out, sel_pos, sel_neg);
  MVPTR_CHECK_LAUNCH("wra_fwd");
  return 0;
}
/* fp32 verification tier: seq [B, Ltot, H] fp32 */
extern "C" int mvptr_f32_wra_fwd(const float* seq, int B, int Ltot, int H, const int64_t* phrase_index,
                                 const int64_t* img_index, const int64_t* neg_img, const int64_t* rand_pos,
                                 const int64_t* rand_neg, int P, float* pos_out, float* neg_out, int* sel_pos,
                                 int* sel_neg, void* stream) {
  if (B <= 0) return 0;
  wra_fwd_kernel<float><<<B, 256, 0, (cudaStream_t)stream>>>(seq, Ltot, H, phrase_index, img_index, neg_img, rand_pos,
                                                             rand_neg, P, pos_out, neg_out, sel_pos, sel_neg);
  MVPTR_CHECK_LAUNCH("f32_wra_fwd");
  return 0;
}
extern "C" int mvptr_f32_wra_bwd(const float* seq, int B, int Ltot, int H, const int64_t* phrase_index,
                                 const int64_t* neg_img, const int* sel_pos, const int* sel_neg, const float* dpos,
                                 const float* dneg, float* dseq, void* stream) {
  if (B <= 0) return 0;
  wra_bwd_kernel<float><<<B, 256, 0, (cudaStream_t)stream>>>(seq, Ltot, H, phrase_index, neg_img, sel_pos, sel_neg, dpos,
                                                             dneg, dseq);
  MVPTR_CHECK_LAUNCH("f32_wra_bwd");
  return 0;
}
extern "C" int mvptr_wra_bwd(const void* seq, int B, int Ltot, int H, const int64_t* phrase_index,
                             const int64_t* neg_img, const int* sel_pos, const int* sel_neg, const float* dpos,
                             const float* dneg, float* dseq, void* stream) {
  if (B <= 0) return 0;
  wra_bwd_kernel<bf16><<<B, 256, 0, (cudaStream_t)stream>>>((const bf16*)seq, Ltot, H, phrase_index, neg_img, sel_pos,
                                                            sel_neg, dpos, dneg, dseq);
  MVPTR_CHECK_LAUNCH("wra_bwd");
  return 0;
}
extern "C" int mvptr_gelu_bwd(const void* dy, const void* pre, void* dx, size_t n, void* stream) {
  if (n == 0) return 0;
  if (n & 7) MVPTR_FAIL(MVPTR_ERR_ARG, "gelu_bwd: n must be a multiple of 8");
  size_t blocks = (n / 8 + 255) / 256;
  if (blocks > (size_t)kNumSMs * 8) blocks = (size_t)kNumSMs * 8;
  gelu_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const bf16*)dy, (const bf16*)pre, (bf16*)dx, n);
  MVPTR_CHECK_LAUNCH("gelu_bwd");
  return 0;
}
extern "C" int mvptr_bce_fwd(const float* logits, int ld, const float* labels, int n, int C, float* loss, void* stream) {
  if (n <= 0) return 0;
  bce_fwd_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(logits, ld, labels, n, C, loss);
  MVPTR_CHECK_LAUNCH("bce_fwd");
  return 0;
}
extern "C" int mvptr_bce_bwd(const float* logits, int ld, const float* labels, int n, int C, const float* gscale,
                             void* dlogits, int ld_d, void* stream) {
  if (n <= 0) return 0;
  bce_bwd_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(logits, ld, labels, n, C, gscale, (bf16*)dlogits, ld_d);
  MVPTR_CHECK_LAUNCH("bce_bwd");
  return 0;
}

extern "C" int mvptr_cls_region_score_fwd(const void* seq, int B, int Ltot, int H, int first, int R, int normalize,
                                          float* logits, float* inv_norm, float p_drop, uint32_t seed, void* stream) {
  if (B <= 0 || R <= 0) return 0;
  if ((H & 7) || H > 1024) MVPTR_FAIL(MVPTR_ERR_UNSUPPORTED, "cls_region_score: hidden size %d", H);
  if (first < 1 || first + R > Ltot) MVPTR_FAIL(MVPTR_ERR_ARG, "cls_region_score: regions [%d, %d) outside 1..%d", first, first + R, Ltot);
  cls_region_score_fwd_kernel<<<(B * R + 3) / 4, 128, 0, (cudaStream_t)stream>>>(
      (const bf16*)seq, B, Ltot, H, first, R, normalize, logits, inv_norm, keep_threshold(p_drop),
      p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f, seed);
  MVPTR_CHECK_LAUNCH("cls_region_score_fwd");
  return 0;
}
/* dseq must be zero-initialised by the caller: only the [CLS] row and the region rows are written */
extern "C" int mvptr_cls_region_score_bwd(const void* seq, int B, int Ltot, int H, int first, int R, int normalize,
                                          const float* logits, const float* inv_norm, const float* dlogits, void* dseq,
                                          float p_drop, uint32_t seed, void* stream) {
  if (B <= 0 || R <= 0) return 0;
  if ((H & 7) || H > 1024) MVPTR_FAIL(MVPTR_ERR_UNSUPPORTED, "cls_region_score: hidden size %d", H);
  cls_region_score_bwd_kernel<<<B, 256, 8 * H * sizeof(float), (cudaStream_t)stream>>>(
      (const bf16*)seq, B, Ltot, H, first, R, normalize, logits, inv_norm, dlogits, (bf16*)dseq, keep_threshold(p_drop),
      p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f, seed);
  MVPTR_CHECK_LAUNCH("cls_region_score_bwd");
  return 0;
}

MVPTR_DEFINE_EPOCH_SETTER(mvptr_set_epoch_wra)
