// fp32 VERIFICATION tier (BASELINE.json north_star: "bit-exact ... masking and top-k ranking order under fp32",
// "1e-4 in fp32").  The reference runs fp32 by default (oscar/tmp_config_FP32.json; run_retrieval.py:1047 only
// halves on a flag); this tier reproduces it on the GPU with fp32 STORAGE everywhere and fp32-accurate
// contractions, so that arg-max / top-k decisions can be compared with the reference without injecting picks.
//
// Contractions stay on the tcgen05 GEMM (gemm_tcgen05.cu): every fp32 operand is split into three bf16 terms
// x = hi + mid + lo (8 + 8 + 8 significand bits, split3 below) and the six products hi.hi, hi.mid, mid.hi, hi.lo,
// lo.hi, mid.mid are accumulated in fp32 by TMA reduce-add -- dropped terms are <= 2^-24 of |a||b|.  Everything
// else in this file is a deliberately plain fp32 kernel (one warp per row, libm erff / expf / tanhf): written
// independently of the tuned bf16 row kernels, it doubles as a cross-check of them.  Not a performance path.
#include <math.h>

#include "common.cuh"

namespace mvptr {
namespace f32 {

// ---------------------------------------------------------------------------------------------------
// x[rows, K] fp32 (pitch ld_src)  ->  hi | mid | lo bf16 [rows, ld_dst], zero padded in [K, ld_dst)
// ---------------------------------------------------------------------------------------------------
__global__ void split3_kernel(const float* __restrict__ src, long long ld_src, int rows, int K, bf16* __restrict__ hi,
                              bf16* __restrict__ mid, bf16* __restrict__ lo, int ld_dst) {
  const int r = blockIdx.x;
  for (int c = blockIdx.y * blockDim.x + threadIdx.x; c < ld_dst; c += gridDim.y * blockDim.x) {
    float x = c < K ? src[(size_t)r * ld_src + c] : 0.f;
    const bf16 h = __float2bfloat16(x);
    const float r1 = x - __bfloat162float(h);
    const bf16 m = __float2bfloat16(r1);
    const float r2 = r1 - __bfloat162float(m);
    const size_t o = (size_t)r * ld_dst + c;
    hi[o] = h;
    mid[o] = m;
    lo[o] = __float2bfloat16(r2);
  }
}

__device__ __forceinline__ void row_stats(const float* x, int H, int lane, float eps, float& mean, float& rstd) {
  float s = 0.f;
  for (int i = lane; i < H; i += 32) s += x[i];
  mean = warp_sum(s) / H;
  float v = 0.f;
  for (int i = lane; i < H; i += 32) {
    const float d = x[i] - mean;
    v += d * d;
  }
  rstd = 1.0f / sqrtf(warp_sum(v) / H + eps);  // TF-style LayerNorm, eps inside the sqrt (modeling_bert.py:242-246)
}

// y[row map] = LN(x (+ residual)) * gamma + beta; optionally saves the pre-LN sum, mean and rstd (for backward)
__global__ void __launch_bounds__(128)
ln_kernel(const float* __restrict__ x, const float* __restrict__ residual, const float* __restrict__ gamma,
          const float* __restrict__ beta, float* __restrict__ y, int y_rows_per_batch, long long y_batch_stride,
          float* __restrict__ pre_out, float* __restrict__ mean_out, float* __restrict__ rstd_out, int rows, int H,
          float eps) {
  extern __shared__ float sm[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int r = blockIdx.x * 4 + w;
  if (r >= rows) return;
  float* buf = sm + w * H;
  for (int i = lane; i < H; i += 32) {
    float v = x[(size_t)r * H + i];
    if (residual) v += residual[(size_t)r * H + i];
    buf[i] = v;
    if (pre_out) pre_out[(size_t)r * H + i] = v;
  }
  __syncwarp();
  float mean, rstd;
  row_stats(buf, H, lane, eps, mean, rstd);
  if (mean_out && lane == 0) {
    mean_out[r] = mean;
    rstd_out[r] = rstd;
  }
  size_t off = (size_t)r * H;
  if (y_rows_per_batch > 0) {
    const int b = r / y_rows_per_batch;
    off = (size_t)b * y_batch_stride + (size_t)(r - b * y_rows_per_batch) * H;
  }
  for (int i = lane; i < H; i += 32) y[off + i] = gamma[i] * ((buf[i] - mean) * rstd) + beta[i];
}

// BertEmbeddings.forward (modeling_bert.py:262-277): LN(word[ids] + pos[p] + type[seg]), fp32 tables
__global__ void __launch_bounds__(128)
embed_ln_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ type_ids, const int64_t* __restrict__ pos_ids,
                const float* __restrict__ word, const float* __restrict__ pos, const float* __restrict__ type,
                const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ y,
                int y_rows_per_batch, long long y_batch_stride, float* __restrict__ pre_out, float* __restrict__ mean_out,
                float* __restrict__ rstd_out, int rows, int L, int H, float eps, int vocab, int max_pos, int n_types) {
  extern __shared__ float sm[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int r = blockIdx.x * 4 + w;
  if (r >= rows) return;
  long long id = ids[r], ty = type_ids ? type_ids[r] : 0, ps = pos_ids ? pos_ids[r] : (r % L);
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  ty = ty < 0 ? 0 : (ty >= n_types ? n_types - 1 : ty);
  ps = ps < 0 ? 0 : (ps >= max_pos ? max_pos - 1 : ps);
  float* buf = sm + w * H;
  for (int i = lane; i < H; i += 32) {
    // the reference adds words + position + token_type in this order (modeling_bert.py:274)
    const float v = (word[(size_t)id * H + i] + pos[(size_t)ps * H + i]) + type[(size_t)ty * H + i];
    buf[i] = v;
    if (pre_out) pre_out[(size_t)r * H + i] = v;
  }
  __syncwarp();
  float mean, rstd;
  row_stats(buf, H, lane, eps, mean, rstd);
  if (mean_out && lane == 0) {
    mean_out[r] = mean;
    rstd_out[r] = rstd;
  }
  size_t off = (size_t)r * H;
  if (y_rows_per_batch > 0) {
    const int b = r / y_rows_per_batch;
    off = (size_t)b * y_batch_stride + (size_t)(r - b * y_rows_per_batch) * H;
  }
  for (int i = lane; i < H; i += 32) y[off + i] = gamma[i] * ((buf[i] - mean) * rstd) + beta[i];
}

// in place: 1 = erf-GELU (modeling_bert.py:142-148), 2 = tanh, 3 = relu
__global__ void act_kernel(float* __restrict__ x, size_t n, int act) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const float v = x[i];
    x[i] = act == 1 ? v * 0.5f * (1.0f + erff(v * 0.70710678118654752f)) : act == 2 ? tanhf(v) : fmaxf(v, 0.f);
  }
}
// dx = dy * act'(.)  with `saved` = pre-activation (GELU) or output (tanh / relu)
__global__ void act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ saved, float* __restrict__ dx,
                               size_t n, int act) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const float s = saved[i];
    float d;
    if (act == 1) d = 0.5f * (1.0f + erff(s * 0.70710678118654752f)) + s * 0.39894228040143268f * expf(-0.5f * s * s);
    else if (act == 2) d = 1.0f - s * s;
    else d = s > 0.f ? 1.f : 0.f;
    dx[i] = dy[i] * d;
  }
}

// ---------------------------------------------------------------------------------------------------
// Attention (modeling_vlbert.py:63-103), fp32: one CTA per (batch, head); K and V of the head in shared
// memory; one warp per query row: scores -> softmax -> context.  lse saved for backward.
// ---------------------------------------------------------------------------------------------------
constexpr int D = 64;

__global__ void __launch_bounds__(256)
attn_fwd_kernel(const float* __restrict__ qkv, int ld, const float* __restrict__ maskadd, float* __restrict__ ctx,
                int ld_ctx, float* __restrict__ probs, int L, int nh, int H) {
  extern __shared__ float sm[];
  float* Ks = sm;                 // [L][D + 1]
  float* Vs = Ks + L * (D + 1);   // [L][D]
  float* Ps = Vs + L * D;         // [8 warps][L]
  float* Qw = Ps + 8 * L;         // [8 warps][D]
  const int b = blockIdx.y, h = blockIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const float* base = qkv + (size_t)b * L * ld + h * D;
  for (int i = threadIdx.x; i < L * D; i += blockDim.x) {
    const int r = i / D, c = i - r * D;
    Ks[r * (D + 1) + c] = base[(size_t)r * ld + H + c];
    Vs[r * D + c] = base[(size_t)r * ld + 2 * H + c];
  }
  __syncthreads();
  float* p = Ps + w * L;
  float* qs = Qw + w * D;
  for (int q = w; q < L; q += 8) {
    const float* qr = base + (size_t)q * ld;
    qs[lane] = qr[lane];
    qs[lane + 32] = qr[lane + 32];
    __syncwarp();
    // scores: lane owns keys lane, lane + 32, ...
    float mx = -INFINITY;
    for (int k = lane; k < L; k += 32) {
      float s = 0.f;
      for (int d = 0; d < D; ++d) s = fmaf(qs[d], Ks[k * (D + 1) + d], s);
      s = s * 0.125f + maskadd[(size_t)b * L + k];  // / sqrt(64), + (1 - mask) * -10000
      p[k] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int k = lane; k < L; k += 32) {
      const float e = expf(p[k] - mx);
      p[k] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int k = 0; k < L; ++k) {
      const float pk = p[k] * inv;
      o0 = fmaf(pk, Vs[k * D + lane], o0);
      o1 = fmaf(pk, Vs[k * D + lane + 32], o1);
    }
    float* out = ctx + ((size_t)b * L + q) * ld_ctx + h * D;
    out[lane] = o0;
    out[lane + 32] = o1;
    if (probs) {
      float* pr = probs + (((size_t)b * nh + h) * L + q) * L;
      for (int k = lane; k < L; k += 32) pr[k] = p[k] * inv;
    }
    __syncwarp();
  }
}

// backward from the saved probabilities P [B, nh, L, L]:
//   dV = P^T dO ; dP = dO V^T ; dS = P * (dP - rowsum(dP * P)) ; dQ = dS K / 8 ; dK = dS^T Q / 8
// one CTA per (batch, head); Q, K, V, dO and dS live in shared memory (L <= 128 in the verification tier)
__global__ void __launch_bounds__(256)
attn_bwd_kernel(const float* __restrict__ qkv, int ld, const float* __restrict__ probs, const float* __restrict__ dctx,
                int ld_ctx, float* __restrict__ dqkv, int L, int nh, int H) {
  extern __shared__ float sm[];
  float* Qs = sm;               // [L][D]
  float* Ks = Qs + L * D;
  float* Vs = Ks + L * D;
  float* dOs = Vs + L * D;
  float* dSs = dOs + L * D;     // [L][L]
  const int b = blockIdx.y, h = blockIdx.x;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const float* base = qkv + (size_t)b * L * ld + h * D;
  for (int i = threadIdx.x; i < L * D; i += blockDim.x) {
    const int r = i / D, c = i - r * D;
    Qs[i] = base[(size_t)r * ld + c];
    Ks[i] = base[(size_t)r * ld + H + c];
    Vs[i] = base[(size_t)r * ld + 2 * H + c];
    dOs[i] = dctx[((size_t)b * L + r) * ld_ctx + h * D + c];
  }
  __syncthreads();
  const float* P = probs + ((size_t)b * nh + h) * L * L;
  // dS rows: one warp per query
  for (int q = w; q < L; q += 8) {
    float dot = 0.f;
    for (int k = lane; k < L; k += 32) {
      float dp = 0.f;
      for (int d = 0; d < D; ++d) dp = fmaf(dOs[q * D + d], Vs[k * D + d], dp);
      dSs[q * L + k] = dp;
      dot = fmaf(dp, P[(size_t)q * L + k], dot);
    }
    dot = warp_sum(dot);
    for (int k = lane; k < L; k += 32) dSs[q * L + k] = P[(size_t)q * L + k] * (dSs[q * L + k] - dot);
  }
  __syncthreads();
  float* out = dqkv + (size_t)b * L * ld + h * D;
  for (int r = w; r < L; r += 8) {
    // dQ[r] = sum_k dS[r,k] K[k] / 8 ; dK[r] = sum_q dS[q,r] Q[q] / 8 ; dV[r] = sum_q P[q,r] dO[q]
    float a0 = 0.f, a1 = 0.f, k0 = 0.f, k1 = 0.f, v0 = 0.f, v1 = 0.f;
    for (int j = 0; j < L; ++j) {
      const float ds = dSs[r * L + j];
      a0 = fmaf(ds, Ks[j * D + lane], a0);
      a1 = fmaf(ds, Ks[j * D + lane + 32], a1);
      const float dst = dSs[j * L + r];
      k0 = fmaf(dst, Qs[j * D + lane], k0);
      k1 = fmaf(dst, Qs[j * D + lane + 32], k1);
      const float pt = P[(size_t)j * L + r];
      v0 = fmaf(pt, dOs[j * D + lane], v0);
      v1 = fmaf(pt, dOs[j * D + lane + 32], v1);
    }
    float* o = out + (size_t)r * ld;
    o[lane] = a0 * 0.125f;
    o[lane + 32] = a1 * 0.125f;
    o[H + lane] = k0 * 0.125f;
    o[H + lane + 32] = k1 * 0.125f;
    o[2 * H + lane] = v0;
    o[2 * H + lane + 32] = v1;
  }
}

// LayerNorm backward, fp32: dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma;
// dgamma += sum dy * xhat, dbeta += sum dy (atomics).  dy rows may be strided like the forward's output rows.
__global__ void __launch_bounds__(128)
ln_bwd_kernel(const float* __restrict__ dy, int dy_rows_per_batch, long long dy_batch_stride, const float* __restrict__ pre,
              const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
              float* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta, int rows, int H) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int r = blockIdx.x * 4 + w;
  if (r >= rows) return;
  size_t off = (size_t)r * H;
  if (dy_rows_per_batch > 0) {
    const int b = r / dy_rows_per_batch;
    off = (size_t)b * dy_batch_stride + (size_t)(r - b * dy_rows_per_batch) * H;
  }
  const float m = mean[r], rs = rstd[r];
  float s1 = 0.f, s2 = 0.f;
  for (int i = lane; i < H; i += 32) {
    const float xh = (pre[(size_t)r * H + i] - m) * rs;
    const float g = dy[off + i] * gamma[i];
    s1 += g;
    s2 = fmaf(g, xh, s2);
  }
  s1 = warp_sum(s1) / H;
  s2 = warp_sum(s2) / H;
  for (int i = lane; i < H; i += 32) {
    const float xh = (pre[(size_t)r * H + i] - m) * rs;
    const float d = dy[off + i];
    dx[(size_t)r * H + i] = rs * (d * gamma[i] - s1 - xh * s2);
    atomicAdd(dgamma + i, d * xh);
    atomicAdd(dbeta + i, d);
  }
}

// out[n] += sum_m x[m, n]
__global__ void colsum_kernel(const float* __restrict__ x, int ldx, float* __restrict__ out, int M, int N) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s = 0.f;
  for (int m = blockIdx.y; m < M; m += gridDim.y) s += x[(size_t)m * ldx + n];
  atomicAdd(out + n, s);
}

// embedding-table gradients: dword[ids[r]] += dpre[r] (padding row 0 skipped like nn.Embedding(padding_idx=0)),
// dpos[r % L] += dpre[r], dtype[seg[r]] += dpre[r]
__global__ void embed_bwd_kernel(const float* __restrict__ dpre, const int64_t* __restrict__ ids,
                                 const int64_t* __restrict__ type_ids, float* __restrict__ dword, float* __restrict__ dpos,
                                 float* __restrict__ dtype, int rows, int L, int H) {
  const int r = blockIdx.x;
  const long long id = ids[r], ty = type_ids ? type_ids[r] : 0;
  for (int i = threadIdx.x; i < H; i += blockDim.x) {
    const float g = dpre[(size_t)r * H + i];
    if (id != 0) atomicAdd(dword + (size_t)id * H + i, g);
    atomicAdd(dpos + (size_t)(r % L) * H + i, g);
    atomicAdd(dtype + (size_t)ty * H + i, g);
  }
}

// logits[n, C] = x[n, :] . W[C, :]^T + b, all fp32, C small (ITM / retrieval classifier)
__global__ void __launch_bounds__(128)
small_head_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ W, const float* __restrict__ bias,
                  float* __restrict__ logits, int n, int H, int C) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (r >= n) return;
  for (int c = 0; c < C; ++c) {
    float s = 0.f;
    for (int i = lane; i < H; i += 32) s = fmaf(x[(size_t)r * ldx + i], W[(size_t)c * H + i], s);
    s = warp_sum(s);
    if (lane == 0) logits[(size_t)r * C + c] = s + (bias ? bias[c] : 0.f);
  }
}
// dx[n, H] = dlogits . W ; dW[C, H] += dlogits^T x ; db[C] += colsum(dlogits)
__global__ void small_head_bwd_kernel(const float* __restrict__ dl, const float* __restrict__ x, long long ldx,
                                      const float* __restrict__ W, float* __restrict__ dx, float* __restrict__ dW,
                                      float* __restrict__ db, int n, int H, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < H) {
    for (int c = 0; c < C; ++c) {
      float s = 0.f;
      for (int r = 0; r < n; ++r) s = fmaf(dl[(size_t)r * C + c], x[(size_t)r * ldx + i], s);
      atomicAdd(dW + (size_t)c * H + i, s);
    }
    for (int r = 0; r < n; ++r) {
      float s = 0.f;
      for (int c = 0; c < C; ++c) s = fmaf(dl[(size_t)r * C + c], W[(size_t)c * H + i], s);
      dx[(size_t)r * H + i] = s;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < C) {
    float s = 0.f;
    for (int r = 0; r < n; ++r) s += dl[(size_t)r * C + threadIdx.x];
    atomicAdd(db + threadIdx.x, s);
  }
}

// backward of the concat gather (modeling_vlbert.py:542-566, 586): da[row_a[r], t] += dout[r, t] ; db[row_b[r], ...] += ...
__global__ void concat_rows_bwd_kernel(const float* __restrict__ dout, int La, int Lb, int b_col0,
                                       const int64_t* __restrict__ row_a, const int64_t* __restrict__ row_b,
                                       float* __restrict__ da, float* __restrict__ db, int rows, int H) {
  const int Lout = La + (Lb - b_col0);
  const int i = blockIdx.x;
  const int r = i / Lout, t = i - r * Lout;
  float* dst;
  if (t < La) dst = da + ((size_t)(row_a ? row_a[r] : r) * La + t) * H;
  else dst = db + ((size_t)(row_b ? row_b[r] : r) * Lb + b_col0 + (t - La)) * H;
  for (int c = threadIdx.x; c < H; c += blockDim.x) atomicAdd(dst + c, dout[(size_t)i * H + c]);
}
// dst[idx[i], :] += src[i, :]
__global__ void scatter_rows_add_kernel(const float* __restrict__ src, const int64_t* __restrict__ idx,
                                        float* __restrict__ dst, int n, int H) {
  const int i = blockIdx.x;
  for (int c = threadIdx.x; c < H; c += blockDim.x) atomicAdd(dst + (size_t)idx[i] * H + c, src[(size_t)i * H + c]);
}
// dx = (dy - y <dy, y>) / |x|  for y = x / max(|x|, 1e-12)   (F.normalize backward)
__global__ void __launch_bounds__(128)
l2norm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ norm,
                  float* __restrict__ dx, int n, int H) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (r >= n) return;
  float d = 0.f;
  for (int i = lane; i < H; i += 32) d = fmaf(dy[(size_t)r * H + i], y[(size_t)r * H + i], d);
  d = warp_sum(d);
  const float inv = 1.0f / fmaxf(norm[r], 1e-12f);
  for (int i = lane; i < H; i += 32) dx[(size_t)r * H + i] = (dy[(size_t)r * H + i] - y[(size_t)r * H + i] * d) * inv;
}
// dlogits = (softmax(logits) - onehot) * g / n_valid, fp32 out (rows with an ignored label: 0)
__global__ void __launch_bounds__(256)
ce_bwd_kernel(const float* __restrict__ logits, int ld, const int64_t* __restrict__ labels, int V, int ignore_index,
              const float* __restrict__ row_lse, const float* __restrict__ n_valid, const float* __restrict__ gscale,
              float* __restrict__ dlogits, int ld_d) {
  const int r = blockIdx.x;
  const int64_t lab = labels[r];
  const bool valid = lab != ignore_index && lab >= 0 && lab < V;
  const float g = valid ? (gscale ? *gscale : 1.f) / fmaxf(*n_valid, 1.f) : 0.f;
  const float lse = row_lse[r];
  for (int c = threadIdx.x; c < ld_d; c += blockDim.x) {
    float d = 0.f;
    if (c < V && valid) d = (expf(logits[(size_t)r * ld + c] - lse) - (c == lab ? 1.f : 0.f)) * g;
    dlogits[(size_t)r * ld_d + c] = d;
  }
}
// BCE-with-logits backward (modeling_vlbert.py:878-883): (sigmoid(x) - y) * g / n, fp32 out
__global__ void __launch_bounds__(256)
bce_bwd_kernel(const float* __restrict__ logits, int ld, const float* __restrict__ labels, int n, int C,
               const float* __restrict__ gscale, float* __restrict__ dlogits, int ld_d) {
  const int r = blockIdx.x;
  const float g = (gscale ? *gscale : 1.f) / n;
  for (int c = threadIdx.x; c < ld_d; c += blockDim.x) {
    float d = 0.f;
    if (c < C) d = (1.f / (1.f + expf(-logits[(size_t)r * ld + c])) - labels[(size_t)r * C + c]) * g;
    dlogits[(size_t)r * ld_d + c] = d;
  }
}

}  // namespace f32
}  // namespace mvptr

using namespace mvptr;

extern "C" int mvptr_f32_concat_rows_bwd(const float* dout, int La, int Lb, int b_col0, const int64_t* row_a,
                                         const int64_t* row_b, float* da, float* db, int rows, int H, void* stream) {
  const int n = rows * (La + Lb - b_col0);
  if (n <= 0) return 0;
  f32::concat_rows_bwd_kernel<<<n, 128, 0, (cudaStream_t)stream>>>(dout, La, Lb, b_col0, row_a, row_b, da, db, rows, H);
  MVPTR_CHECK_LAUNCH("f32_concat_rows_bwd");
  return 0;
}
extern "C" int mvptr_f32_scatter_rows_add(const float* src, const int64_t* idx, float* dst, int n, int H, void* stream) {
  if (n <= 0) return 0;
  f32::scatter_rows_add_kernel<<<n, 128, 0, (cudaStream_t)stream>>>(src, idx, dst, n, H);
  MVPTR_CHECK_LAUNCH("f32_scatter_rows_add");
  return 0;
}
extern "C" int mvptr_f32_l2norm_bwd(const float* dy, const float* y, const float* norm, float* dx, int n, int H,
                                    void* stream) {
  if (n <= 0) return 0;
  f32::l2norm_bwd_kernel<<<(n + 3) / 4, 128, 0, (cudaStream_t)stream>>>(dy, y, norm, dx, n, H);
  MVPTR_CHECK_LAUNCH("f32_l2norm_bwd");
  return 0;
}
extern "C" int mvptr_f32_ce_bwd(const float* logits, int ld, const int64_t* labels, int n, int V, int ignore_index,
                                const float* row_lse, const float* n_valid, const float* gscale, float* dlogits, int ld_d,
                                void* stream) {
  if (n <= 0) return 0;
  f32::ce_bwd_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(logits, ld, labels, V, ignore_index, row_lse, n_valid, gscale,
                                                         dlogits, ld_d);
  MVPTR_CHECK_LAUNCH("f32_ce_bwd");
  return 0;
}
extern "C" int mvptr_f32_bce_bwd(const float* logits, int ld, const float* labels, int n, int C, const float* gscale,
                                 float* dlogits, int ld_d, void* stream) {
  if (n <= 0) return 0;
  f32::bce_bwd_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(logits, ld, labels, n, C, gscale, dlogits, ld_d);
  MVPTR_CHECK_LAUNCH("f32_bce_bwd");
  return 0;
}

extern "C" int mvptr_f32_split3(const float* src, long long ld_src, int rows, int K, void* hi, void* mid, void* lo,
                                int ld_dst, void* stream) {
  if (rows <= 0) return 0;
  if (ld_dst < K || (ld_dst & 7)) MVPTR_FAIL(MVPTR_ERR_ARG, "f32_split3: ld_dst=%d must be >= K=%d and a multiple of 8", ld_dst, K);
  int gy = (ld_dst + 255) / 256;
  if (gy > 16) gy = 16;
  f32::split3_kernel<<<dim3(rows, gy), 256, 0, (cudaStream_t)stream>>>(src, ld_src, rows, K, (bf16*)hi, (bf16*)mid,
                                                                        (bf16*)lo, ld_dst);
  MVPTR_CHECK_LAUNCH("f32_split3");
  return 0;
}

extern "C" int mvptr_f32_ln_fwd(const float* x, const float* residual, const float* gamma, const float* beta, float* y,
                                int y_rows_per_batch, long long y_batch_stride, float* pre_out, float* mean_out,
                                float* rstd_out, int rows, int H, float eps, void* stream) {
  if (rows <= 0) return 0;
  f32::ln_kernel<<<(rows + 3) / 4, 128, 4 * H * sizeof(float), (cudaStream_t)stream>>>(
      x, residual, gamma, beta, y, y_rows_per_batch, y_batch_stride, pre_out, mean_out, rstd_out, rows, H, eps);
  MVPTR_CHECK_LAUNCH("f32_ln_fwd");
  return 0;
}

extern "C" int mvptr_f32_embed_ln_fwd(const int64_t* ids, const int64_t* type_ids, const int64_t* pos_ids,
                                      const float* word, const float* pos, const float* type, const float* gamma,
                                      const float* beta, float* y, int y_rows_per_batch, long long y_batch_stride,
                                      float* pre_out, float* mean_out, float* rstd_out, int B, int L, int H, float eps,
                                      int vocab, int max_pos, int n_types, void* stream) {
  const int rows = B * L;
  if (rows <= 0) return 0;
  f32::embed_ln_kernel<<<(rows + 3) / 4, 128, 4 * H * sizeof(float), (cudaStream_t)stream>>>(
      ids, type_ids, pos_ids, word, pos, type, gamma, beta, y, y_rows_per_batch, y_batch_stride, pre_out, mean_out,
      rstd_out, rows, L, H, eps, vocab, max_pos, n_types);
  MVPTR_CHECK_LAUNCH("f32_embed_ln_fwd");
  return 0;
}

extern "C" int mvptr_f32_act(float* x, size_t n, int act, void* stream) {
  if (n == 0) return 0;
  if (act < 1 || act > 3) MVPTR_FAIL(MVPTR_ERR_ARG, "f32_act: act %d (1 gelu, 2 tanh, 3 relu)", act);
  size_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  f32::act_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, n, act);
  MVPTR_CHECK_LAUNCH("f32_act");
  return 0;
}
extern "C" int mvptr_f32_act_bwd(const float* dy, const float* saved, float* dx, size_t n, int act, void* stream) {
  if (n == 0) return 0;
  if (act < 1 || act > 3) MVPTR_FAIL(MVPTR_ERR_ARG, "f32_act_bwd: act %d (1 gelu, 2 tanh, 3 relu)", act);
  size_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  f32::act_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(dy, saved, dx, n, act);
  MVPTR_CHECK_LAUNCH("f32_act_bwd");
  return 0;
}

extern "C" int mvptr_f32_attn_fwd(const float* qkv, int ld_qkv, const float* maskadd, float* ctx, int ld_ctx,
                                  float* probs, int B, int L, int nh, int H, void* stream) {
  if (H != nh * f32::D) MVPTR_FAIL(MVPTR_ERR_UNSUPPORTED, "f32 attention: head size must be 64");
  if (B <= 0 || L <= 0 || L > 256) MVPTR_FAIL(MVPTR_ERR_UNSUPPORTED, "f32 attention: B=%d L=%d unsupported (L in 1..256)", B, L);
  const int smem = (L * (f32::D + 1) + L * f32::D + 8 * L + 8 * f32::D) * (int)sizeof(float);
  static int configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(f32::attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    if (e != cudaSuccess) MVPTR_FAIL(MVPTR_ERR_CUDA, "f32 attn smem attr: %s", cudaGetErrorString(e));
    configured = 160 * 1024;
  }
  f32::attn_fwd_kernel<<<dim3(nh, B), 256, smem, (cudaStream_t)stream>>>(qkv, ld_qkv, maskadd, ctx, ld_ctx, probs, L, nh, H);
  MVPTR_CHECK_LAUNCH("f32_attn_fwd");
  return 0;
}
extern "C" int mvptr_f32_attn_bwd(const float* qkv, int ld_qkv, const float* probs, const float* dctx, int ld_ctx,
                                  float* dqkv, int B, int L, int nh, int H, void* stream) {
  if (H != nh * f32::D) MVPTR_FAIL(MVPTR_ERR_UNSUPPORTED, "f32 attention: head size must be 64");
  const int smem = (4 * L * f32::D + L * L) * (int)sizeof(float);
  if (B <= 0 || L <= 0 || smem > 200 * 1024)
    MVPTR_FAIL(MVPTR_ERR_UNSUPPORTED, "f32 attention backward: B=%d L=%d unsupported (verification tier: L <= 128)", B, L);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(f32::attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) MVPTR_FAIL(MVPTR_ERR_CUDA, "f32 attn bwd smem attr: %s", cudaGetErrorString(e));
    configured = true;
  }
  f32::attn_bwd_kernel<<<dim3(nh, B), 256, smem, (cudaStream_t)stream>>>(qkv, ld_qkv, probs, dctx, ld_ctx, dqkv, L, nh, H);
  MVPTR_CHECK_LAUNCH("f32_attn_bwd");
  return 0;
}

extern "C" int mvptr_f32_ln_bwd(const float* dy, int dy_rows_per_batch, long long dy_batch_stride, const float* pre,
                                const float* mean, const float* rstd, const float* gamma, float* dx, float* dgamma,
                                float* dbeta, int rows, int H, void* stream) {
  if (rows <= 0) return 0;
  f32::ln_bwd_kernel<<<(rows + 3) / 4, 128, 0, (cudaStream_t)stream>>>(dy, dy_rows_per_batch, dy_batch_stride, pre, mean,
                                                                        rstd, gamma, dx, dgamma, dbeta, rows, H);
  MVPTR_CHECK_LAUNCH("f32_ln_bwd");
  return 0;
}
extern "C" int mvptr_f32_colsum(const float* x, int ldx, float* out, int M, int N, void* stream) {
  if (M <= 0 || N <= 0) return 0;
  int gy = (M + 63) / 64;
  if (gy > 64) gy = 64;
  f32::colsum_kernel<<<dim3((N + 127) / 128, gy), 128, 0, (cudaStream_t)stream>>>(x, ldx, out, M, N);
  MVPTR_CHECK_LAUNCH("f32_colsum");
  return 0;
}
extern "C" int mvptr_f32_embed_bwd(const float* dpre, const int64_t* ids, const int64_t* type_ids, float* dword,
                                   float* dpos, float* dtype, int B, int L, int H, void* stream) {
  if (B * L <= 0) return 0;
  f32::embed_bwd_kernel<<<B * L, 128, 0, (cudaStream_t)stream>>>(dpre, ids, type_ids, dword, dpos, dtype, B * L, L, H);
  MVPTR_CHECK_LAUNCH("f32_embed_bwd");
  return 0;
}
extern "C" int mvptr_f32_small_head_fwd(const float* x, long long ldx, const float* W, const float* bias, float* logits,
                                        int n, int H, int C, void* stream) {
  if (n <= 0) return 0;
  f32::small_head_kernel<<<(n + 3) / 4, 128, 0, (cudaStream_t)stream>>>(x, ldx, W, bias, logits, n, H, C);
  MVPTR_CHECK_LAUNCH("f32_small_head_fwd");
  return 0;
}
extern "C" int mvptr_f32_small_head_bwd(const float* dlogits, const float* x, long long ldx, const float* W, float* dx,
                                        float* dW, float* db, int n, int H, int C, void* stream) {
  if (n <= 0) return 0;
  if (C > 128) MVPTR_FAIL(MVPTR_ERR_ARG, "f32_small_head_bwd: C=%d too large", C);
  f32::small_head_bwd_kernel<<<(H + 127) / 128, 128, 0, (cudaStream_t)stream>>>(dlogits, x, ldx, W, dx, dW, db, n, H, C);
  MVPTR_CHECK_LAUNCH("f32_small_head_bwd");
  return 0;
}
