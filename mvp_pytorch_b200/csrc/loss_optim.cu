// Loss and optimizer kernels (HBM-bound): softmax cross-entropy over vocabulary logits,
// in-batch contrastive (VSC) loss with hard-negative argmax, ITM head, L2 normalisation,
// fused flat-arena AdamW and gradient norm.
#include "common.cuh"

namespace mvptr {

__device__ __forceinline__ float block_reduce_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float s = 0.f;
  for (int i = 0; i < nw; ++i) s += red[i];
  return s;
}
__device__ __forceinline__ float block_reduce_max(float v, float* red) {
  v = warp_max(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float s = -INFINITY;
  for (int i = 0; i < nw; ++i) s = fmaxf(s, red[i]);
  return s;
}

// ---------------------------------------------------------------------------------
// CrossEntropyLoss(ignore_index) over fp32 logits [n, V] (pitch ld)
// modeling_vlbert.py:1229,1235,1249 ; one block per row
// ---------------------------------------------------------------------------------
// Forward: the whole row is staged in registers by 16-byte loads (512 threads x NV float4), so the logits are
// read from HBM exactly once and every load of the row is in flight at the same time.  Rows whose label is
// ignored (the padding slots of the fixed-capacity masked-LM selection) are skipped.
template <int NV>
__global__ void __launch_bounds__(512)
ce_fwd_kernel(const float* __restrict__ logits, int ld, const int64_t* __restrict__ labels, int V, int ignore_index,
              float* __restrict__ row_lse, float* __restrict__ loss_sum, float* __restrict__ n_valid) {
  __shared__ float red[16];
  const int r = blockIdx.x;
  const long long lab = labels[r];
  if (lab == ignore_index || lab < 0 || lab >= V) {  // block-uniform
    if (threadIdx.x == 0) row_lse[r] = 0.f;
    return;
  }
  const float* x = logits + (size_t)r * ld;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  const int nvec = V >> 2;
  float4 buf[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int i = threadIdx.x + k * 512;
    buf[k] = i < nvec ? x4[i] : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  }
  const int ti = 4 * nvec + threadIdx.x;
  const float tail = ti < V ? x[ti] : -INFINITY;  // V % 4 leftover elements
  float mx = tail;
#pragma unroll
  for (int k = 0; k < NV; ++k) mx = fmaxf(mx, fmaxf(fmaxf(buf[k].x, buf[k].y), fmaxf(buf[k].z, buf[k].w)));
  mx = block_reduce_max(mx, red);
  const float m2 = mx * 1.4426950408889634f;
  float s = ex2_approx(fmaf(tail, 1.4426950408889634f, -m2));
#pragma unroll
  for (int k = 0; k < NV; ++k)
    s += ex2_approx(fmaf(buf[k].x, 1.4426950408889634f, -m2)) + ex2_approx(fmaf(buf[k].y, 1.4426950408889634f, -m2)) +
         ex2_approx(fmaf(buf[k].z, 1.4426950408889634f, -m2)) + ex2_approx(fmaf(buf[k].w, 1.4426950408889634f, -m2));
  s = block_reduce_sum(s, red);
  if (threadIdx.x == 0) {
    const float lse = mx + logf(s);
    row_lse[r] = lse;
    atomicAdd(loss_sum, lse - x[lab]);
    atomicAdd(n_valid, 1.0f);
  }
}

// Backward: dlogits = (softmax - onehot) * g / n_valid as bf16, 8 columns (two 16-byte loads, one 16-byte
// store) per thread and step.
__global__ void __launch_bounds__(256)
ce_bwd_kernel(const float* __restrict__ logits, int ld, const int64_t* __restrict__ labels, int V, int ignore_index,
              const float* __restrict__ row_lse, const float* __restrict__ n_valid, const float* __restrict__ gscale,
              bf16* __restrict__ dlogits, int ld_d) {
  const int r = blockIdx.x;
  const float* x = logits + (size_t)r * ld;
  bf16* d = dlogits + (size_t)r * ld_d;
  const long long lab = labels[r];
  const bool valid = lab != ignore_index && lab >= 0 && lab < V;
  const float nv = *n_valid;
  const float sc = valid && nv > 0.f ? (gscale ? *gscale : 1.f) / nv : 0.f;
  const float l2 = row_lse[r] * 1.4426950408889634f;
  const bool vec = (ld & 3) == 0 && (ld_d & 7) == 0 && ld >= ld_d;  // whole 8-column groups are readable
  if (vec) {
    for (int i = threadIdx.x * 8; i < ld_d; i += blockDim.x * 8) {
      float g[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (valid) {
        const float4 a = *reinterpret_cast<const float4*>(x + i), b = *reinterpret_cast<const float4*>(x + i + 4);
        const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int j = 0; j < 8; ++j)
          g[j] = i + j < V ? (ex2_approx(fmaf(v[j], 1.4426950408889634f, -l2)) - (i + j == lab ? 1.f : 0.f)) * sc : 0.f;
      }
      *reinterpret_cast<bf16x8*>(d + i) = pack8(g);
    }
  } else {
    for (int i = threadIdx.x; i < ld_d; i += blockDim.x) {
      float g = 0.f;
      if (i < V && valid) g = (__expf(x[i] - row_lse[r]) - (i == lab ? 1.f : 0.f)) * sc;
      d[i] = __float2bfloat16(g);
    }
  }
}

// ---------------------------------------------------------------------------------
// y = x / max(||x||_2, 1e-12)   (F.normalize, modeling_vlbert.py:525-526); x fp32 [n, H] -> y fp32 + bf16
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
l2norm_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, bf16* __restrict__ y16, float* __restrict__ norm,
                  int n, int H) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (r >= n) return;
  float s = 0.f;
  for (int i = lane; i < H; i += 32) {
    const float v = x[(size_t)r * H + i];
    s += v * v;
  }
  const float nr = fmaxf(sqrtf(warp_sum(s)), 1e-12f);
  if (lane == 0 && norm) norm[r] = nr;
  for (int i = lane; i < H; i += 32) {
    const float v = x[(size_t)r * H + i] / nr;
    if (y) y[(size_t)r * H + i] = v;
    if (y16) y16[(size_t)r * H + i] = __float2bfloat16(v);
  }
}
// dx = (dy - y * <dy, y>) / norm
__global__ void __launch_bounds__(128)
l2norm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ norm,
                  bf16* __restrict__ dx16, int n, int H) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (r >= n) return;
  float s = 0.f;
  for (int i = lane; i < H; i += 32) s += dy[(size_t)r * H + i] * y[(size_t)r * H + i];
  s = warp_sum(s);
  const float inv = 1.f / norm[r];
  for (int i = lane; i < H; i += 32)
    dx16[(size_t)r * H + i] = __float2bfloat16((dy[(size_t)r * H + i] - y[(size_t)r * H + i] * s) * inv);
}

// ---------------------------------------------------------------------------------
// VSC loss + hard negatives on sim [B,B] fp32 (modeling_vlbert.py:527-534, 1238-1241).
// Block r handles row r and column r.  scale = exp(logit_scale).
//   loss = ( CE(scale*sim, arange) + CE(scale*sim^T, arange) ) / 2
//   hard_img[r] = argmax_j (sim[r,j] - 2*[j==r]) ; hard_txt[r] = argmax_i (sim[i,r] - 2*[i==r])
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
vsc_fwd_kernel(const float* __restrict__ sim, int B, const float* __restrict__ logit_scale, float* __restrict__ row_lse,
               float* __restrict__ col_lse, float* __restrict__ loss, int64_t* __restrict__ hard_img,
               int64_t* __restrict__ hard_txt) {
  __shared__ float red[8];
  __shared__ float bestv[256];
  __shared__ int besti[256];
  const int r = blockIdx.x;
  const float sc = __expf(*logit_scale);
  float lse2[2];
  for (int pass = 0; pass < 2; ++pass) {
    const size_t stride = pass == 0 ? 1 : (size_t)B;
    const float* x = sim + (pass == 0 ? (size_t)r * B : (size_t)r);
    float mx = -INFINITY, bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < B; i += blockDim.x) {
      const float v = x[i * stride];
      mx = fmaxf(mx, v * sc);
      const float mv = v - (i == r ? 2.f : 0.f);
      if (mv > bv) {
        bv = mv;
        bi = i;
      }
    }
    mx = block_reduce_max(mx, red);
    float s = 0.f;
    for (int i = threadIdx.x; i < B; i += blockDim.x) s += __expf(x[i * stride] * sc - mx);
    s = block_reduce_sum(s, red);
    lse2[pass] = mx + logf(s);
    // first-index argmax (torch.max semantics on ties)
    bestv[threadIdx.x] = bv;
    besti[threadIdx.x] = bi;
    __syncthreads();
    if (threadIdx.x == 0) {
      float v = bestv[0];
      int idx = besti[0];
      for (int i = 1; i < blockDim.x; ++i)
        if (bestv[i] > v || (bestv[i] == v && besti[i] < idx)) {
          v = bestv[i];
          idx = besti[i];
        }
      if (pass == 0) hard_img[r] = idx;
      else hard_txt[r] = idx;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    row_lse[r] = lse2[0];
    col_lse[r] = lse2[1];
    const float d = sim[(size_t)r * B + r] * sc;
    atomicAdd(loss, ((lse2[0] - d) + (lse2[1] - d)) / (2.f * B));
  }
}
// dsim[i,j] = g*scale/(2B) * (softmax_row + softmax_col - 2*delta) ; dlogit_scale += sum dsim[i,j]*sim[i,j]
__global__ void __launch_bounds__(256)
vsc_bwd_kernel(const float* __restrict__ sim, int B, const float* __restrict__ logit_scale,
               const float* __restrict__ row_lse, const float* __restrict__ col_lse, const float* __restrict__ gscale,
               float* __restrict__ dsim, float* __restrict__ dlogit_scale) {
  __shared__ float red[8];
  const int r = blockIdx.x;
  const float sc = __expf(*logit_scale);
  const float g = (gscale ? *gscale : 1.f) / (2.f * B);
  float acc = 0.f;
  for (int j = threadIdx.x; j < B; j += blockDim.x) {
    const float v = sim[(size_t)r * B + j];
    const float m = v * sc;
    const float dm = g * (__expf(m - row_lse[r]) + __expf(m - col_lse[j]) - (j == r ? 2.f : 0.f));
    dsim[(size_t)r * B + j] = dm * sc;
    acc += dm * m;  // d/d(logit_scale) of m = sim*exp(ls) is m
  }
  acc = block_reduce_sum(acc, red);
  if (threadIdx.x == 0 && dlogit_scale) atomicAdd(dlogit_scale, acc);
}

// ---------------------------------------------------------------------------------
// small-N head: logits[n, C] = x[n, H] . W[C, H]^T + b  (ITM / retrieval classifier,
// modeling_vlbert.py:1247,1680,1708), C <= 8; one warp per row.  Optional CE + grads.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
small_head_fwd_kernel(const bf16* __restrict__ x, int ldx, const bf16* __restrict__ W, const bf16* __restrict__ bias,
                      float* __restrict__ logits, int n, int H, int C) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (r >= n) return;
  for (int c = 0; c < C; ++c) {
    float s = 0.f;
    for (int i = lane; i < H; i += 32) s += __bfloat162float(x[(size_t)r * ldx + i]) * __bfloat162float(W[(size_t)c * H + i]);
    s = warp_sum(s);
    if (lane == 0) logits[(size_t)r * C + c] = s + (bias ? __bfloat162float(bias[c]) : 0.f);
  }
}
// dx[n,H] = dlogits[n,C] . W ; dW[C,H] += dlogits^T . x ; db[C] += colsum(dlogits)
__global__ void __launch_bounds__(128)
small_head_bwd_dx_kernel(const float* __restrict__ dlogits, const bf16* __restrict__ W, bf16* __restrict__ dx, int ld_dx,
                         int n, int H, int C) {
  const int r = blockIdx.x;
  for (int i = threadIdx.x; i < H; i += blockDim.x) {
    float s = 0.f;
    for (int c = 0; c < C; ++c) s += dlogits[(size_t)r * C + c] * __bfloat162float(W[(size_t)c * H + i]);
    dx[(size_t)r * ld_dx + i] = __float2bfloat16(s);
  }
}
__global__ void __launch_bounds__(128)
small_head_bwd_dw_kernel(const float* __restrict__ dlogits, const bf16* __restrict__ x, int ldx, float* __restrict__ dW,
                         float* __restrict__ db, int n, int H, int C) {
  // grid (ceil(H/128), C, row slices): every CTA reduces one slice of the rows and adds it atomically (a single
  // CTA per column block walking all n rows serially took 0.13 ms for the 512 x 768 ITM head)
  const int c = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int per = (n + gridDim.z - 1) / gridDim.z;
  const int r0 = blockIdx.z * per, r1 = min(n, r0 + per);
  if (i < H) {
    float s = 0.f;
    for (int r = r0; r < r1; ++r) s += dlogits[(size_t)r * C + c] * __bfloat162float(x[(size_t)r * ldx + i]);
    atomicAdd(dW + (size_t)c * H + i, s);
  }
  if (blockIdx.x == 0 && threadIdx.x < 32 && db) {
    float s = 0.f;
    for (int r = r0 + threadIdx.x; r < r1; r += 32) s += dlogits[(size_t)r * C + c];
    s = warp_sum(s);
    if (threadIdx.x == 0) atomicAdd(db + c, s);
  }
}
// CE over tiny logits [n, C] fp32 with int64 labels (CrossEntropyLoss(ignore_index), modeling_vlbert.py:1251,
// :1262-1264, :1682).  Rows whose label is ignore_index -- or outside [0, C) -- contribute neither loss nor
// gradient and the mean runs over the valid rows only, as torch does.
// Forward (dlogits == null): acc[0] += sum of row losses, acc[1] += valid rows.  Backward (dlogits != null):
// dlogits = (softmax - onehot) * gscale / acc[1] with acc[1] left by the forward; acc is not touched.
__global__ void small_ce_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, int n, int C,
                                float* __restrict__ acc, float* __restrict__ dlogits, const float* __restrict__ gscale) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int64_t lab64 = labels[r];
  const bool valid = lab64 >= 0 && lab64 < C;
  const int lab = valid ? (int)lab64 : 0;
  float mx = -INFINITY;
  for (int c = 0; c < C; ++c) mx = fmaxf(mx, logits[(size_t)r * C + c]);
  float s = 0.f;
  for (int c = 0; c < C; ++c) s += __expf(logits[(size_t)r * C + c] - mx);
  const float lse = mx + logf(s);
  if (!dlogits) {
    if (valid) {
      atomicAdd(acc, lse - logits[(size_t)r * C + lab]);
      atomicAdd(acc + 1, 1.f);
    }
    return;
  }
  const float g = valid ? (gscale ? *gscale : 1.f) / fmaxf(acc[1], 1.f) : 0.f;
  for (int c = 0; c < C; ++c)
    dlogits[(size_t)r * C + c] = (__expf(logits[(size_t)r * C + c] - lse) - (c == lab ? 1.f : 0.f)) * g;
}

// ---------------------------------------------------------------------------------
// Fused AdamW over a flat parameter arena (optimization.py:130-189), fp32 master +
// moments, also refreshes the bf16 compute copy.  Elements [0, decay_end) get weight
// decay (the arena is laid out decay-first).  grad_scale folds gradient clipping.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             bf16* __restrict__ p16, size_t n, size_t decay_end, float lr, float step_size, float beta1, float beta2,
             float eps, float weight_decay, const float* __restrict__ grad_norm, float max_norm,
             const float* __restrict__ dyn, int correct_bias) {
  if (dyn) {  // {lr, step} read at execution time (CUDA-graph replays)
    lr = dyn[0];
    const float st = dyn[1];
    step_size = correct_bias ? lr * sqrtf(1.f - powf(beta2, st)) / (1.f - powf(beta1, st)) : lr;
  }
  float gs = 1.f;
  if (grad_norm && max_norm > 0.f) {
    const float nrm = sqrtf(*grad_norm);
    gs = fminf(1.f, max_norm / (nrm + 1e-6f));
  }
  size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const size_t stride = (size_t)gridDim.x * blockDim.x * 4;
  for (; i < n; i += stride) {
    float4 pv = *reinterpret_cast<float4*>(p + i);
    const float4 gv = *reinterpret_cast<const float4*>(g + i);
    float4 mv = *reinterpret_cast<float4*>(m + i);
    float4 vv = *reinterpret_cast<float4*>(v + i);
    float* pp = &pv.x;
    const float* gg = &gv.x;
    float* mm = &mv.x;
    float* vq = &vv.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gr = gg[j] * gs;
      mm[j] = mm[j] * beta1 + (1.f - beta1) * gr;
      vq[j] = vq[j] * beta2 + (1.f - beta2) * gr * gr;
      float x = pp[j] - step_size * mm[j] / (sqrtf(vq[j]) + eps);
      if (i + j < decay_end) x -= lr * weight_decay * x;
      pp[j] = x;
    }
    *reinterpret_cast<float4*>(p + i) = pv;
    *reinterpret_cast<float4*>(m + i) = mv;
    *reinterpret_cast<float4*>(v + i) = vv;
    if (p16) {
      *reinterpret_cast<uint2*>(p16 + i) = make_uint2(pack2(pv.x, pv.y), pack2(pv.z, pv.w));
    }
  }
}
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, size_t n, float* __restrict__ out) {
  __shared__ float red[8];
  float s = 0.f;
  size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const size_t stride = (size_t)gridDim.x * blockDim.x * 4;
  for (; i < n; i += stride) {
    const float4 v = *reinterpret_cast<const float4*>(g + i);
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  s = block_reduce_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(out, s);
}

}  // namespace mvptr

using namespace mvptr;

extern "C" int mvptr_ce_fwd(const float* logits, int ld, const int64_t* labels, int n, int V, int ignore_index,
                            float* row_lse, float* loss_sum, float* n_valid, void* stream) {
  MVPTR_PROF("ce_fwd", 4.0*n*V, stream);
  if (n <= 0) return 0;
  if ((ld & 3) || (reinterpret_cast<uintptr_t>(logits) & 15))
    MVPTR_FAIL(MVPTR_ERR_ARG, "ce_fwd: logits must be 16-byte aligned with a pitch that is a multiple of 4");
  const int per_thread = ((V >> 2) + 511) / 512;  // float4 per thread
  if (per_thread <= 4)
    ce_fwd_kernel<4><<<n, 512, 0, (cudaStream_t)stream>>>(logits, ld, labels, V, ignore_index, row_lse, loss_sum, n_valid);
  else if (per_thread <= 16)
    ce_fwd_kernel<16><<<n, 512, 0, (cudaStream_t)stream>>>(logits, ld, labels, V, ignore_index, row_lse, loss_sum, n_valid);
  else if (per_thread <= 48)
    ce_fwd_kernel<48><<<n, 512, 0, (cudaStream_t)stream>>>(logits, ld, labels, V, ignore_index, row_lse, loss_sum, n_valid);
  else
    MVPTR_FAIL(MVPTR_ERR_UNSUPPORTED, "ce_fwd: more than 98304 classes per row");
  MVPTR_CHECK_LAUNCH("ce_fwd");
  return 0;
}
extern "C" int mvptr_ce_bwd(const float* logits, int ld, const int64_t* labels, int n, int V, int ignore_index,
                            const float* row_lse, const float* n_valid, const float* gscale, void* dlogits, int ld_d,
                            void* stream) {
  MVPTR_PROF("ce_bwd", 6.0*n*V, stream);
  if (n <= 0) return 0;
  ce_bwd_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(logits, ld, labels, V, ignore_index, row_lse, n_valid, gscale,
                                                     (bf16*)dlogits, ld_d);
  MVPTR_CHECK_LAUNCH("ce_bwd");
  return 0;
}
extern "C" int mvptr_l2norm_fwd(const float* x, float* y, void* y16, float* norm, int n, int H, void* stream) {
  if (n <= 0) return 0;
  l2norm_fwd_kernel<<<(n + 3) / 4, 128, 0, (cudaStream_t)stream>>>(x, y, (bf16*)y16, norm, n, H);
  MVPTR_CHECK_LAUNCH("l2norm_fwd");
  return 0;
}
extern "C" int mvptr_l2norm_bwd(const float* dy, const float* y, const float* norm, void* dx16, int n, int H,
                                void* stream) {
  if (n <= 0) return 0;
  l2norm_bwd_kernel<<<(n + 3) / 4, 128, 0, (cudaStream_t)stream>>>(dy, y, norm, (bf16*)dx16, n, H);
  MVPTR_CHECK_LAUNCH("l2norm_bwd");
  return 0;
}
extern "C" int mvptr_vsc_fwd(const float* sim, int B, const float* logit_scale, float* row_lse, float* col_lse,
                             float* loss, int64_t* hard_img, int64_t* hard_txt, void* stream) {
  if (B <= 0) return 0;
  vsc_fwd_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(sim, B, logit_scale, row_lse, col_lse, loss, hard_img, hard_txt);
  MVPTR_CHECK_LAUNCH("vsc_fwd");
  return 0;
}
extern "C" int mvptr_vsc_bwd(const float* sim, int B, const float* logit_scale, const float* row_lse,
                             const float* col_lse, const float* gscale, float* dsim, float* dlogit_scale, void* stream) {
  if (B <= 0) return 0;
  vsc_bwd_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(sim, B, logit_scale, row_lse, col_lse, gscale, dsim, dlogit_scale);
  MVPTR_CHECK_LAUNCH("vsc_bwd");
  return 0;
}
extern "C" int mvptr_small_head_fwd(const void* x, int ldx, const void* W, const void* bias, float* logits, int n, int H,
                                    int C, void* stream) {
  if (n <= 0) return 0;
  if (C > 64) MVPTR_FAIL(MVPTR_ERR_ARG, "small_head: C=%d too large (use mvptr_gemm)", C);
  small_head_fwd_kernel<<<(n + 3) / 4, 128, 0, (cudaStream_t)stream>>>((const bf16*)x, ldx, (const bf16*)W,
                                                                       (const bf16*)bias, logits, n, H, C);
  MVPTR_CHECK_LAUNCH("small_head_fwd");
  return 0;
}
extern "C" int mvptr_small_head_bwd(const float* dlogits, const void* x, int ldx, const void* W, void* dx, int ld_dx,
                                    float* dW, float* db, int n, int H, int C, void* stream) {
  if (n <= 0) return 0;
  if (dx) {
    small_head_bwd_dx_kernel<<<n, 128, 0, (cudaStream_t)stream>>>(dlogits, (const bf16*)W, (bf16*)dx, ld_dx, n, H, C);
    MVPTR_CHECK_LAUNCH("small_head_bwd_dx");
  }
  if (dW) {
    const int slices = n >= 512 ? 32 : (n + 15) / 16;
    small_head_bwd_dw_kernel<<<dim3((H + 127) / 128, C, slices), 128, 0, (cudaStream_t)stream>>>(dlogits, (const bf16*)x,
                                                                                                  ldx, dW, db, n, H, C);
    MVPTR_CHECK_LAUNCH("small_head_bwd_dw");
  }
  return 0;
}
extern "C" int mvptr_small_ce(const float* logits, const int64_t* labels, int n, int C, float* acc, float* dlogits,
                              const float* gscale, void* stream) {
  if (n <= 0) return 0;
  if (!acc) MVPTR_FAIL(MVPTR_ERR_ARG, "small_ce: acc (float[2]: loss sum, valid rows) is required");
  small_ce_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(logits, labels, n, C, acc, dlogits, gscale);
  MVPTR_CHECK_LAUNCH("small_ce");
  return 0;
}
extern "C" int mvptr_adamw(float* p, const float* g, float* m, float* v, void* p16, size_t n, size_t decay_end, float lr,
                           float beta1, float beta2, float eps, float weight_decay, int step, int correct_bias,
                           const float* grad_sumsq, float max_norm, const float* dyn_lr_step, void* stream) {
  MVPTR_PROF("adamw", 30.0*n, stream);
  if (n == 0) return 0;
  if (n & 3) MVPTR_FAIL(MVPTR_ERR_ARG, "adamw: arena size must be a multiple of 4");
  float step_size = lr;
  if (correct_bias) step_size = lr * sqrtf(1.f - powf(beta2, (float)step)) / (1.f - powf(beta1, (float)step));
  size_t blocks = (n / 4 + 255) / 256;
  if (blocks > (size_t)kNumSMs * 8) blocks = (size_t)kNumSMs * 8;
  adamw_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, (bf16*)p16, n, decay_end, lr, step_size,
                                                                   beta1, beta2, eps, weight_decay, grad_sumsq, max_norm,
                                                                   dyn_lr_step, correct_bias);
  MVPTR_CHECK_LAUNCH("adamw");
  return 0;
}
extern "C" int mvptr_sumsq(const float* g, size_t n, float* out, void* stream) {
  MVPTR_PROF("sumsq", 4.0*n, stream);
  if (n == 0) return 0;
  if (n & 3) MVPTR_FAIL(MVPTR_ERR_ARG, "sumsq: n must be a multiple of 4");
  size_t blocks = (n / 4 + 255) / 256;
  if (blocks > (size_t)kNumSMs * 8) blocks = (size_t)kNumSMs * 8;
  sumsq_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(g, n, out);
  MVPTR_CHECK_LAUNCH("sumsq");
  return 0;
}
