// HBM-bound row kernels: embedding gather + LayerNorm, LayerNorm fwd/bwd, column sums,
// embedding scatter-add, region-feature pad/cast, mask preparation, row gathers.
// One warp per row, 16-byte vector accesses, warp-shuffle reductions.
#include <cuda_fp16.h>

#include "common.cuh"

namespace mvptr {

constexpr int kMaxChunks = 4;  // hidden <= 32 lanes * 4 chunks * 8 = 1024

struct RowMap {
  // logical row r = b * rows_per_batch + t  ->  element offset b * batch_stride + t * H
  int rows_per_batch;
  long long batch_stride;
  __device__ __forceinline__ size_t off(int r, int H) const {
    const int b = r / rows_per_batch;
    return (size_t)b * batch_stride + (size_t)(r - b * rows_per_batch) * H;
  }
};

__device__ __forceinline__ void ln_row(float (&x)[kMaxChunks][8], int nchunk, int lane, int H, float eps, float& mean,
                                       float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < kMaxChunks; ++c)
    if (lane + 32 * c < nchunk)
#pragma unroll
      for (int j = 0; j < 8; ++j) s += x[c][j];
  mean = warp_sum(s) / H;
  float v = 0.f;
#pragma unroll
  for (int c = 0; c < kMaxChunks; ++c)
    if (lane + 32 * c < nchunk)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = x[c][j] - mean;
        v += d * d;
      }
  rstd = rsqrtf(warp_sum(v) / H + eps);  // TF-style: eps inside the sqrt (modeling_bert.py:245)
}

// ---------------------------------------------------------------------------------
// y = dropout(LN(word[ids] + pos[p] + type[seg]))      modeling_bert.py:262-277
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
embed_ln_fwd_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ type_ids,
                    const int64_t* __restrict__ pos_ids, const bf16* __restrict__ word,
                    const bf16* __restrict__ pos, const bf16* __restrict__ type, const bf16* __restrict__ gamma,
                    const bf16* __restrict__ beta, bf16* __restrict__ y, RowMap ymap, bf16* __restrict__ pre,
                    float* __restrict__ mean_out, float* __restrict__ rstd_out, int rows, int L, int H, float eps,
                    int vocab, int max_pos, int n_types, uint32_t keep_thr, float inv_keep, uint32_t seed_) {
  const uint32_t seed = site_seed(seed_);
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int nchunk = H >> 3;
  long long id = ids[r];
  long long ty = type_ids ? type_ids[r] : 0;
  long long ps = pos_ids ? pos_ids[r] : (r % L);
  // out-of-range indices would be a hard device fault in the reference; clamp so the
  // kernel stays memory-safe (the host wrapper validates ranges in debug mode)
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  ty = ty < 0 ? 0 : (ty >= n_types ? n_types - 1 : ty);
  ps = ps < 0 ? 0 : (ps >= max_pos ? max_pos - 1 : ps);
  const bf16* w = word + (size_t)id * H;
  const bf16* p = pos + (size_t)ps * H;
  const bf16* t = type + (size_t)ty * H;
  float x[kMaxChunks][8];
#pragma unroll
  for (int c = 0; c < kMaxChunks; ++c) {
    const int ch = lane + 32 * c;
    if (ch < nchunk) {
      float a[8], b2[8], c2[8];
      unpack8(*reinterpret_cast<const bf16x8*>(w + ch * 8), a);
      unpack8(*reinterpret_cast<const bf16x8*>(p + ch * 8), b2);
      unpack8(*reinterpret_cast<const bf16x8*>(t + ch * 8), c2);
#pragma unroll
      for (int j = 0; j < 8; ++j) x[c][j] = a[j] + b2[j] + c2[j];
      if (pre) *reinterpret_cast<bf16x8*>(pre + (size_t)r * H + ch * 8) = pack8(x[c]);
    }
  }
  float mean, rstd;
  ln_row(x, nchunk, lane, H, eps, mean, rstd);
  if (mean_out && lane == 0) {
    mean_out[r] = mean;
    rstd_out[r] = rstd;
  }
  bf16* yo = y + ymap.off(r, H);
#pragma unroll
  for (int c = 0; c < kMaxChunks; ++c) {
    const int ch = lane + 32 * c;
    if (ch < nchunk) {
      float g[8], b2[8], o[8];
      unpack8(*reinterpret_cast<const bf16x8*>(gamma + ch * 8), g);
      unpack8(*reinterpret_cast<const bf16x8*>(beta + ch * 8), b2);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = g[j] * ((x[c][j] - mean) * rstd) + b2[j];
      if (keep_thr != 0xffffffffu) dropout8(o, seed, (uint32_t)r * (uint32_t)H + ch * 8, keep_thr, inv_keep);
      *reinterpret_cast<bf16x8*>(yo + ch * 8) = pack8(o);
    }
  }
}

// ---------------------------------------------------------------------------------
// y = dropout(LN(x))                    modeling_bert.py:242-246 (+ modeling_vlbert.py:499-503)
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
ln_fwd_kernel(const bf16* __restrict__ x_in, const bf16* __restrict__ residual, bf16* __restrict__ pre_out,
              const bf16* __restrict__ gamma, const bf16* __restrict__ beta, bf16* __restrict__ y, RowMap ymap,
              float* __restrict__ mean_out, float* __restrict__ rstd_out, int rows, int H, float eps,
              uint32_t in_keep_thr, float in_inv_keep, uint32_t in_seed_, uint32_t keep_thr, float inv_keep,
              uint32_t seed_) {
  const uint32_t in_seed = site_seed(in_seed_), seed = site_seed(seed_);
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int nchunk = H >> 3;
  float x[kMaxChunks][8];
#pragma unroll
  for (int c = 0; c < kMaxChunks; ++c) {
    const int ch = lane + 32 * c;
    if (ch < nchunk) {
      unpack8(*reinterpret_cast<const bf16x8*>(x_in + (size_t)r * H + ch * 8), x[c]);
      if (in_keep_thr != 0xffffffffu) dropout8(x[c], in_seed, (uint32_t)r * (uint32_t)H + ch * 8, in_keep_thr, in_inv_keep);
      if (residual) {
        float rr[8];
        unpack8(*reinterpret_cast<const bf16x8*>(residual + (size_t)r * H + ch * 8), rr);
#pragma unroll
        for (int j = 0; j < 8; ++j) x[c][j] += rr[j];
      }
      if (pre_out) {
        // the saved pre-LN activation is what backward normalises, so LN runs on its bf16 rounding
        const bf16x8 pk = pack8(x[c]);
        *reinterpret_cast<bf16x8*>(pre_out + (size_t)r * H + ch * 8) = pk;
        unpack8(pk, x[c]);
      }
    }
  }
  float mean, rstd;
  ln_row(x, nchunk, lane, H, eps, mean, rstd);
  if (mean_out && lane == 0) {
    mean_out[r] = mean;
    rstd_out[r] = rstd;
  }
  bf16* yo = y + ymap.off(r, H);
#pragma unroll
  for (int c = 0; c < kMaxChunks; ++c) {
    const int ch = lane + 32 * c;
    if (ch < nchunk) {
      float g[8], b2[8], o[8];
      unpack8(*reinterpret_cast<const bf16x8*>(gamma + ch * 8), g);
      unpack8(*reinterpret_cast<const bf16x8*>(beta + ch * 8), b2);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = g[j] * ((x[c][j] - mean) * rstd) + b2[j];
      if (keep_thr != 0xffffffffu) dropout8(o, seed, (uint32_t)r * (uint32_t)H + ch * 8, keep_thr, inv_keep);
      *reinterpret_cast<bf16x8*>(yo + ch * 8) = pack8(o);
    }
  }
}

// y = gelu(x) elementwise (BertIntermediate activation, modeling_bert.py:396) -- run as its own
// fully-occupied coalesced pass: cheaper than ~25 instructions/element in the 8-warp GEMM epilogue
__global__ void __launch_bounds__(256) gelu_fwd_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, size_t n) {
  size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  const size_t stride = (size_t)gridDim.x * blockDim.x * 8;
  for (; i + stride + 8 <= n; i += 2 * stride) {
    const bf16x8 q0 = *reinterpret_cast<const bf16x8*>(x + i), q1 = *reinterpret_cast<const bf16x8*>(x + i + stride);
    float a[8], b[8];
    unpack8(q0, a); unpack8(q1, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) { a[j] = gelu_erf(a[j]); b[j] = gelu_erf(b[j]); }
    *reinterpret_cast<bf16x8*>(y + i) = pack8(a);
    *reinterpret_cast<bf16x8*>(y + i + stride) = pack8(b);
  }
  for (; i + 8 <= n; i += stride) {
    float a[8];
    unpack8(*reinterpret_cast<const bf16x8*>(x + i), a);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = gelu_erf(a[j]);
    *reinterpret_cast<bf16x8*>(y + i) = pack8(a);
  }
}

// dx = dy * gelu'(pre) and dbias[n] += sum_m dx[m,n] in one pass (thread = 8 columns, loops rows)
__global__ void __launch_bounds__(256)
gelu_bwd_colsum_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ pre, bf16* __restrict__ dx,
                       float* __restrict__ dbias, int M, int N, int rows_per_block) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + tx) * 8;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(M, r0 + rows_per_block);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (col < N) {
    int r = r0 + ty;
    for (; r + 8 < r1; r += 16) {  // two rows per iteration: 4 independent loads in flight
      const size_t o0 = (size_t)r * N + col, o1 = (size_t)(r + 8) * N + col;
      const bf16x8 d0 = *reinterpret_cast<const bf16x8*>(dy + o0), d1 = *reinterpret_cast<const bf16x8*>(dy + o1);
      const bf16x8 p0 = *reinterpret_cast<const bf16x8*>(pre + o0), p1 = *reinterpret_cast<const bf16x8*>(pre + o1);
      float a0[8], a1[8], x0[8], x1[8];
      unpack8(d0, a0); unpack8(d1, a1); unpack8(p0, x0); unpack8(p1, x1);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a0[j] *= gelu_erf_grad(x0[j]);
        a1[j] *= gelu_erf_grad(x1[j]);
      }
      const bf16x8 k0 = pack8(a0), k1 = pack8(a1);
      *reinterpret_cast<bf16x8*>(dx + o0) = k0;
      *reinterpret_cast<bf16x8*>(dx + o1) = k1;
      unpack8(k0, a0); unpack8(k1, a1);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += a0[j] + a1[j];
    }
    for (; r < r1; r += 8) {
      float a[8], x[8];
      const size_t off = (size_t)r * N + col;
      unpack8(*reinterpret_cast<const bf16x8*>(dy + off), a);
      unpack8(*reinterpret_cast<const bf16x8*>(pre + off), x);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] *= gelu_erf_grad(x[j]);
      const bf16x8 pk = pack8(a);
      *reinterpret_cast<bf16x8*>(dx + off) = pk;
      unpack8(pk, a);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += a[j];
    }
  }
  if (dbias == nullptr) return;
  __shared__ float red[8][32 * 8 + 1];
#pragma unroll
  for (int j = 0; j < 8; ++j) red[ty][tx * 8 + j] = acc[j];
  __syncthreads();
  if (ty == 0 && col < N) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += red[w][tx * 8 + j];
      atomicAdd(dbias + col + j, s);
    }
  }
}

// ---------------------------------------------------------------------------------
// LayerNorm backward.  dy may live in a strided (batch-mapped) buffer.
//   g = dy (* output-dropout mask) ; dx = rstd * (g*gamma - mean(g*gamma) - xhat*mean(g*gamma*xhat))
//   dgamma += sum_r g*xhat ; dbeta += sum_r g
//   dx_drop = dx * input-dropout mask (the dropout that sat between the dense and the
//             residual add, modeling_bert.py:350-351) ; dbias += sum_r dx_drop
// ---------------------------------------------------------------------------------
// Single pass over the rows: a persistent grid (2 CTAs of 8 warps per SM) walks the rows warp by
// warp, writes dx / dx_drop and keeps the three column sums (dgamma, dbeta, dbias) of the columns a
// lane owns in registers; they are folded warp -> CTA (shared atomics) -> global (one atomicAdd per
// column per CTA) at the end.  HBM traffic = read dy, x + write dx (+ dx_drop): 3-4 tensors instead of
// the 7 of a dx pass followed by a column-sum pass.
template <int NC>
__global__ void __launch_bounds__(256, 2)
ln_bwd_kernel(const bf16* __restrict__ dy, RowMap dymap, const bf16* __restrict__ x_in,
              const float* __restrict__ mean_in, const float* __restrict__ rstd_in, const bf16* __restrict__ gamma,
              bf16* __restrict__ dx, bf16* __restrict__ dx_drop, float* __restrict__ dgamma,
              float* __restrict__ dbeta, float* __restrict__ dbias, int rows, int H, uint32_t out_keep_thr,
              float out_inv_keep, uint32_t out_seed_, uint32_t in_keep_thr, float in_inv_keep, uint32_t in_seed_) {
  const uint32_t out_seed = site_seed(out_seed_), in_seed = site_seed(in_seed_);
  extern __shared__ float cta_acc[];  // [3][H]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nchunk = H >> 3;
  for (int i = threadIdx.x; i < 3 * H; i += 256) cta_acc[i] = 0.f;
  __syncthreads();
  float ag[NC][8], ab[NC][8], ad[NC][8];
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) ag[c][j] = ab[c][j] = ad[c][j] = 0.f;
  const float inv_h = 1.f / H;
  for (int r = blockIdx.x * 8 + warp; r < rows; r += gridDim.x * 8) {
    const float mean = mean_in[r], rstd = rstd_in[r];
    const bf16* dyr = dy + dymap.off(r, H);
    bf16x8 pg[NC], px[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int ch = lane + 32 * c;
      if (ch < nchunk) {
        pg[c] = *reinterpret_cast<const bf16x8*>(dyr + ch * 8);
        px[c] = *reinterpret_cast<const bf16x8*>(x_in + (size_t)r * H + ch * 8);
      }
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int ch = lane + 32 * c;
      if (ch < nchunk) {
        float g[8], xv[8], gm[8];
        unpack8(pg[c], g);
        unpack8(px[c], xv);
        unpack8(__ldg(reinterpret_cast<const uint4*>(gamma + ch * 8)), gm);
        if (out_keep_thr != 0xffffffffu) {
          dropout8(g, out_seed, (uint32_t)r * (uint32_t)H + ch * 8, out_keep_thr, out_inv_keep);
          pg[c] = pack8(g);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (xv[j] - mean) * rstd;
          const float gg = g[j] * gm[j];
          s1 += gg;
          s2 += gg * xh;
          ag[c][j] += g[j] * xh;
          ab[c][j] += g[j];
        }
      }
    }
    s1 = warp_sum(s1) * inv_h;
    s2 = warp_sum(s2) * inv_h;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int ch = lane + 32 * c;
      if (ch < nchunk) {
        float g[8], xv[8], o[8], gm[8];
        unpack8(pg[c], g);
        unpack8(px[c], xv);
        unpack8(__ldg(reinterpret_cast<const uint4*>(gamma + ch * 8)), gm);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = rstd * (g[j] * gm[j] - s1 - (xv[j] - mean) * rstd * s2);
        bf16x8 pk = pack8(o);
        if (dx) *reinterpret_cast<bf16x8*>(dx + (size_t)r * H + ch * 8) = pk;
        if (dx_drop) {
          dropout8(o, in_seed, (uint32_t)r * (uint32_t)H + ch * 8, in_keep_thr, in_inv_keep);
          pk = pack8(o);
          *reinterpret_cast<bf16x8*>(dx_drop + (size_t)r * H + ch * 8) = pk;
        }
        if (dbias) {  // column sum of exactly what the dense layer's backward reads (the bf16 rounding)
          unpack8(pk, o);
#pragma unroll
          for (int j = 0; j < 8; ++j) ad[c][j] += o[j];
        }
      }
    }
  }
  // j outer: the 32 lanes of a warp then hit 32 addresses 8 words apart (4-way bank conflict at worst)
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int ch = lane + 32 * c;
    if (ch < nchunk) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        atomicAdd(&cta_acc[ch * 8 + j], ag[c][j]);
        atomicAdd(&cta_acc[H + ch * 8 + j], ab[c][j]);
        if (dbias) atomicAdd(&cta_acc[2 * H + ch * 8 + j], ad[c][j]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < H; i += 256) {
    atomicAdd(dgamma + i, cta_acc[i]);
    atomicAdd(dbeta + i, cta_acc[H + i]);
    if (dbias) atomicAdd(dbias + i, cta_acc[2 * H + i]);
  }
}

// out[n] += sum_m X[m,n]   (bias gradients of a dense layer: column sum of dY)
__global__ void __launch_bounds__(256)
colsum_kernel(const bf16* __restrict__ x, int ldx, float* __restrict__ out, int M, int N, int rows_per_block) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = (blockIdx.x * 32 + tx) * 8;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(M, r0 + rows_per_block);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (col < N) {
    int r = r0 + ty;
    for (; r + 24 < r1; r += 32) {  // 4 independent 16-byte loads in flight per thread
      bf16x8 q[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) q[u] = *reinterpret_cast<const bf16x8*>(x + (size_t)(r + 8 * u) * ldx + col);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float v[8];
        unpack8(q[u], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += v[j];
      }
    }
    for (; r < r1; r += 8) {
      float v[8];
      unpack8(*reinterpret_cast<const bf16x8*>(x + (size_t)r * ldx + col), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
  }
  __shared__ float red[8][32 * 8 + 1];
#pragma unroll
  for (int j = 0; j < 8; ++j) red[ty][tx * 8 + j] = acc[j];
  __syncthreads();
  if (ty == 0 && col < N) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += red[w][tx * 8 + j];
      if (col + j < N) atomicAdd(out + col + j, s);
    }
  }
}

// Embedding backward (modeling_bert.py:262-277): word rows by atomics (padding_idx row
// gets no gradient, as nn.Embedding(padding_idx=0)), position / type rows by a
// per-position block reduction over the batch.
__global__ void __launch_bounds__(128)
embed_word_bwd_kernel(const bf16* __restrict__ dpre, const int64_t* __restrict__ ids, float* __restrict__ dword,
                      int rows, int H, int vocab, int padding_idx) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const long long id = ids[r];
  if (id == padding_idx || id < 0 || id >= vocab) return;
  float* dst = dword + (size_t)id * H;
  for (int ch = lane; ch < (H >> 3); ch += 32) {
    float v[8];
    unpack8(*reinterpret_cast<const bf16x8*>(dpre + (size_t)r * H + ch * 8), v);
    if ((reinterpret_cast<uintptr_t>(dword) & 15) == 0) {  // vector reductions: a quarter of the atomic operations
      atomicAdd(reinterpret_cast<float4*>(dst + ch * 8), make_float4(v[0], v[1], v[2], v[3]));
      atomicAdd(reinterpret_cast<float4*>(dst + ch * 8 + 4), make_float4(v[4], v[5], v[6], v[7]));
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(dst + ch * 8 + j, v[j]);
    }
  }
}
__global__ void __launch_bounds__(128)
embed_pos_type_bwd_kernel(const bf16* __restrict__ dpre, const int64_t* __restrict__ type_ids,
                          float* __restrict__ dpos, float* __restrict__ dtype, int B, int L, int H, int n_types) {
  // block = one position p and one slice of the batch (grid.y); thread = 8 columns
  const int p = blockIdx.x;
  const int b_per = (B + gridDim.y - 1) / gridDim.y;
  const int b_lo = blockIdx.y * b_per, b_hi = min(B, b_lo + b_per);
  for (int col = threadIdx.x * 8; col < H; col += blockDim.x * 8) {
    float accp[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float acct[2][8] = {{0, 0, 0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0, 0, 0}};
    for (int b = b_lo; b < b_hi; ++b) {
      const int r = b * L + p;
      float v[8];
      unpack8(*reinterpret_cast<const bf16x8*>(dpre + (size_t)r * H + col), v);
      const long long ty = type_ids ? type_ids[r] : 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        accp[j] += v[j];
        if (ty == 0) acct[0][j] += v[j];
        else if (ty == 1) acct[1][j] += v[j];
      }
      if (ty > 1 && ty < n_types)
#pragma unroll
        for (int j = 0; j < 8; ++j) atomicAdd(dtype + (size_t)ty * H + col + j, v[j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(dpos + (size_t)p * H + col + j, accp[j]);
      atomicAdd(dtype + col + j, acct[0][j]);
      if (n_types > 1) atomicAdd(dtype + H + col + j, acct[1][j]);
    }
  }
}

template <typename T>
constexpr bool kIsHalf = false;
template <>
constexpr bool kIsHalf<__half> = true;

// fp32/bf16/fp16 [rows, K] (any pitch) -> bf16 [rows, Kp] zero padded  (region features, K=2054)
template <typename T>
__global__ void pad_cast_kernel(const T* __restrict__ src, long long ld_src, bf16* __restrict__ dst, int ld_dst,
                                int rows, int K) {
  const int r = blockIdx.x;  // rows on grid.x: config 5 has 102 400 region rows (> the 65 535 limit of grid.y)
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  if (c >= ld_dst) return;
  float v = 0.f;
  if (c < K) v = (float)src[(size_t)r * ld_src + c];
  dst[(size_t)r * ld_dst + c] = __float2bfloat16(v);
}
// Same, 8 output columns per thread: one 16-byte store fed by four PAIR loads (a 2054-element row is only
// 4-byte (bf16) / 8-byte (fp32) aligned, so pairs are the widest aligned source access).  Needs even K and
// ld_src, ld_dst % 8 == 0 and a pair-aligned source.
template <typename T>
__global__ void __launch_bounds__(256)
pad_cast8_kernel(const T* __restrict__ src, long long ld_src, bf16* __restrict__ dst, int ld_dst, int rows, int K) {
  const int r = blockIdx.x;
  const int c0 = (blockIdx.y * blockDim.x + threadIdx.x) * 8;
  if (c0 >= ld_dst) return;
  const T* s = src + (size_t)r * ld_src + c0;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; j += 2) {
    v[j] = v[j + 1] = 0.f;
    if (c0 + j < K) {  // K even: a pair is either fully inside or fully outside
      if constexpr (sizeof(T) == 4) {
        const float2 t = __ldg(reinterpret_cast<const float2*>(s + j));
        v[j] = t.x;
        v[j + 1] = t.y;
      } else if constexpr (sizeof(T) == 2 && !kIsHalf<T>) {
        const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(s + j));
        v[j] = __uint_as_float(w << 16);
        v[j + 1] = __uint_as_float(w & 0xffff0000u);
      } else {  // fp16 features (the reference's --half_evaluation / DeepSpeed fp16 loaders)
        const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(s + j));
        const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&w));
        v[j] = t.x;
        v[j + 1] = t.y;
      }
    }
  }
  *reinterpret_cast<bf16x8*>(dst + (size_t)r * ld_dst + c0) = pack8(v);
}

// additive attention mask: (1 - mask) * -10000   (modeling_vlbert.py:430-460), with an
// optional per-row source remap and column window so the joint / hard-negative masks
// are assembled without torch.cat / index_select (modeling_vlbert.py:542-566, 587)
__global__ void mask_prepare_kernel(const int64_t* __restrict__ mask_a, int La, const int64_t* __restrict__ mask_b,
                                    int Lb, int b_col0, const int64_t* __restrict__ row_a,
                                    const int64_t* __restrict__ row_b, float* __restrict__ out, int rows) {
  const int Lout = La + (Lb - b_col0);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * Lout) return;
  const int r = i / Lout, c = i - r * Lout;
  long long m;
  if (c < La) m = mask_a[(size_t)(row_a ? row_a[r] : r) * La + c];
  else m = mask_b[(size_t)(row_b ? row_b[r] : r) * Lb + b_col0 + (c - La)];
  out[i] = (1.0f - (float)m) * -10000.0f;
}

// out[r, t, :] = t < La ? a[row_a[r], t, :] : b[row_b[r], b_col0 + t - La, :]
// (torch.cat + index_select of modeling_vlbert.py:542-566, 586 as one gather)
__global__ void __launch_bounds__(128)
concat_rows_kernel(const bf16* __restrict__ a, int La, const bf16* __restrict__ b, int Lb, int b_col0,
                   const int64_t* __restrict__ row_a, const int64_t* __restrict__ row_b, bf16* __restrict__ out,
                   int rows, int H) {
  const int Lout = La + (Lb - b_col0);
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (i >= rows * Lout) return;
  const int r = i / Lout, t = i - r * Lout;
  const bf16* src;
  if (t < La) src = a + ((size_t)(row_a ? row_a[r] : r) * La + t) * H;
  else src = b + ((size_t)(row_b ? row_b[r] : r) * Lb + b_col0 + (t - La)) * H;
  bf16* dst = out + (size_t)i * H;
  for (int ch = lane; ch < (H >> 3); ch += 32)
    *reinterpret_cast<uint4*>(dst + ch * 8) = *reinterpret_cast<const uint4*>(src + ch * 8);
}
// backward of the gather: da[row_a[r], t] += dout[r, t] ; db[row_b[r], b_col0 + t - La] += ...
__global__ void __launch_bounds__(128)
concat_rows_bwd_kernel(const bf16* __restrict__ dout, int La, int Lb, int b_col0, const int64_t* __restrict__ row_a,
                       const int64_t* __restrict__ row_b, float* __restrict__ da, float* __restrict__ db, int rows,
                       int H) {
  const int Lout = La + (Lb - b_col0);
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (i >= rows * Lout) return;
  const int r = i / Lout, t = i - r * Lout;
  float* dst;
  if (t < La) dst = da + ((size_t)(row_a ? row_a[r] : r) * La + t) * H;
  else dst = db + ((size_t)(row_b ? row_b[r] : r) * Lb + b_col0 + (t - La)) * H;
  const bf16* src = dout + (size_t)i * H;
  // 16-byte vector reductions (red.global.add.v4.f32, sm_90+): a quarter of the atomic operations of the scalar form
  const bool vec = ((reinterpret_cast<uintptr_t>(da) | reinterpret_cast<uintptr_t>(db)) & 15) == 0;
  for (int ch = lane; ch < (H >> 3); ch += 32) {
    float v[8];
    unpack8(*reinterpret_cast<const bf16x8*>(src + ch * 8), v);
    if (vec) {
      atomicAdd(reinterpret_cast<float4*>(dst + ch * 8), make_float4(v[0], v[1], v[2], v[3]));
      atomicAdd(reinterpret_cast<float4*>(dst + ch * 8 + 4), make_float4(v[4], v[5], v[6], v[7]));
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(dst + ch * 8 + j, v[j]);
    }
  }
}

// out[i, :] = src[rows[i], :]   (masked_select of MLM rows, modeling_vlbert.py:1232,1246)
__global__ void __launch_bounds__(128)
gather_rows_kernel(const bf16* __restrict__ src, const int64_t* __restrict__ idx, bf16* __restrict__ out, int n, int H) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (i >= n) return;
  const bf16* s = src + (size_t)idx[i] * H;
  for (int ch = lane; ch < (H >> 3); ch += 32)
    *reinterpret_cast<uint4*>(out + (size_t)i * H + ch * 8) = *reinterpret_cast<const uint4*>(s + ch * 8);
}
// dst[rows[i], :] += src[i, :]  (fp32 accumulate; indexes are unique for masked_select)
__global__ void __launch_bounds__(128)
scatter_rows_add_kernel(const bf16* __restrict__ src, const int64_t* __restrict__ idx, float* __restrict__ dst, int n,
                        int H) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (i >= n) return;
  float* d = dst + (size_t)idx[i] * H;
  for (int ch = lane; ch < (H >> 3); ch += 32) {
    float v[8];
    unpack8(*reinterpret_cast<const bf16x8*>(src + (size_t)i * H + ch * 8), v);
    if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
      atomicAdd(reinterpret_cast<float4*>(d + ch * 8), make_float4(v[0], v[1], v[2], v[3]));
      atomicAdd(reinterpret_cast<float4*>(d + ch * 8 + 4), make_float4(v[4], v[5], v[6], v[7]));
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(d + ch * 8 + j, v[j]);
    }
  }
}

// elementwise helpers on flat buffers
__global__ void cast_f32_to_bf16_kernel(const float* __restrict__ s, bf16* __restrict__ d, size_t n) {
  size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  const size_t stride = (size_t)gridDim.x * blockDim.x * 8;
  for (; i + 8 <= n; i += stride) {
    const float4 a = *reinterpret_cast<const float4*>(s + i);
    const float4 b = *reinterpret_cast<const float4*>(s + i + 4);
    float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    *reinterpret_cast<bf16x8*>(d + i) = pack8(v);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (size_t k = n & ~size_t(7); k < n; ++k) d[k] = __float2bfloat16(s[k]);
}
// d (+)= float(s): the data-parallel gradient path reduces bf16 copies of the arena slices over NVLink and
// widens the averaged result back into the fp32 arena (parallel.GradientSync)
__global__ void cast_bf16_to_f32_kernel(const bf16* __restrict__ s, float* __restrict__ d, size_t n) {
  size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  const size_t stride = (size_t)gridDim.x * blockDim.x * 8;
  for (; i + 8 <= n; i += stride) {
    float v[8];
    unpack8(*reinterpret_cast<const bf16x8*>(s + i), v);
    *reinterpret_cast<float4*>(d + i) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(d + i + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (size_t k = n & ~size_t(7); k < n; ++k) d[k] = __bfloat162float(s[k]);
}
// d = bf16(a + b) for two fp32 gradient buffers (or b == nullptr)
__global__ void add_cast_kernel(const float* __restrict__ a, const bf16* __restrict__ b, bf16* __restrict__ d, size_t n) {
  size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  const size_t stride = (size_t)gridDim.x * blockDim.x * 8;
  for (; i + 8 <= n; i += stride) {
    const float4 x = *reinterpret_cast<const float4*>(a + i);
    const float4 y = *reinterpret_cast<const float4*>(a + i + 4);
    float v[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
    if (b) {
      float w[8];
      unpack8(*reinterpret_cast<const bf16x8*>(b + i), w);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += w[j];
    }
    *reinterpret_cast<bf16x8*>(d + i) = pack8(v);
  }
}

static inline uint32_t thr(float p) { return keep_threshold(p); }
static inline float invk(float p) { return p > 0.f ? 1.f / (1.f - p) : 1.f; }

}  // namespace mvptr

using namespace mvptr;

#define CHECK_H(H)                                                                         \
  if ((H) <= 0 || ((H)&7) || (H) > 32 * kMaxChunks * 8) MVPTR_FAIL(MVPTR_ERR_ARG, "hidden size %d unsupported (multiple of 8, <= 1024)", (H))

extern "C" int mvptr_embed_ln_fwd(const int64_t* ids, const int64_t* type_ids, const int64_t* pos_ids,
                                  const void* word, const void* pos, const void* type, const void* gamma,
                                  const void* beta, void* y, int y_rows_per_batch, long long y_batch_stride, void* pre,
                                  float* mean, float* rstd, int B, int L, int H, float eps, int vocab, int max_pos,
                                  int n_types, float p_drop, uint32_t seed, void* stream) {
  MVPTR_PROF("embed_ln_fwd", 0, stream);
  CHECK_H(H);
  if (L > max_pos && !pos_ids) MVPTR_FAIL(MVPTR_ERR_ARG, "sequence length %d exceeds position table %d", L, max_pos);
  const int rows = B * L;
  if (rows == 0) return 0;
  RowMap ym{y_rows_per_batch > 0 ? y_rows_per_batch : L, y_rows_per_batch > 0 ? y_batch_stride : (long long)L * H};
  embed_ln_fwd_kernel<<<(rows + 3) / 4, 128, 0, (cudaStream_t)stream>>>(
      ids, type_ids, pos_ids, (const bf16*)word, (const bf16*)pos, (const bf16*)type, (const bf16*)gamma,
      (const bf16*)beta, (bf16*)y, ym, (bf16*)pre, mean, rstd, rows, L, H, eps, vocab, max_pos, n_types, thr(p_drop),
      invk(p_drop), seed);
  MVPTR_CHECK_LAUNCH("embed_ln_fwd");
  return 0;
}

extern "C" int mvptr_ln_fwd(const void* x, const void* gamma, const void* beta, void* y, int y_rows_per_batch,
                            long long y_batch_stride, float* mean, float* rstd, int rows, int H, float eps,
                            float p_drop, uint32_t seed, void* stream) {
  MVPTR_PROF("ln_fwd", 4.0*rows*H, stream);
  CHECK_H(H);
  if (rows == 0) return 0;
  RowMap ym{y_rows_per_batch > 0 ? y_rows_per_batch : rows, y_rows_per_batch > 0 ? y_batch_stride : 0};
  ln_fwd_kernel<<<(rows + 3) / 4, 128, 0, (cudaStream_t)stream>>>(
      (const bf16*)x, nullptr, nullptr, (const bf16*)gamma, (const bf16*)beta, (bf16*)y, ym, mean, rstd, rows, H, eps,
      0xffffffffu, 1.f, 0, thr(p_drop), invk(p_drop), seed);
  MVPTR_CHECK_LAUNCH("ln_fwd");
  return 0;
}

extern "C" int mvptr_add_ln_fwd(const void* x, const void* residual, float in_p_drop, uint32_t in_seed, void* pre_out,
                                const void* gamma, const void* beta, void* y, float* mean, float* rstd, int rows,
                                int H, float eps, void* stream) {
  MVPTR_PROF("add_ln_fwd", 8.0*rows*H, stream);
  CHECK_H(H);
  if (rows == 0) return 0;
  RowMap ym{rows, 0};
  ln_fwd_kernel<<<(rows + 3) / 4, 128, 0, (cudaStream_t)stream>>>(
      (const bf16*)x, (const bf16*)residual, (bf16*)pre_out, (const bf16*)gamma, (const bf16*)beta, (bf16*)y, ym, mean,
      rstd, rows, H, eps, thr(in_p_drop), invk(in_p_drop), in_seed, 0xffffffffu, 1.f, 0);
  MVPTR_CHECK_LAUNCH("add_ln_fwd");
  return 0;
}

extern "C" int mvptr_gelu_fwd(const void* x, void* y, size_t n, void* stream) {
  MVPTR_PROF("gelu_fwd", 4.0*n, stream);
  if (n == 0) return 0;
  if (n & 7) MVPTR_FAIL(MVPTR_ERR_ARG, "gelu_fwd: n must be a multiple of 8");
  size_t blocks = (n / 8 + 255) / 256;
  if (blocks > (size_t)kNumSMs * 16) blocks = (size_t)kNumSMs * 16;
  gelu_fwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)y, n);
  MVPTR_CHECK_LAUNCH("gelu_fwd");
  return 0;
}

extern "C" int mvptr_gelu_bwd_colsum(const void* dy, const void* pre, void* dx, float* dbias, int M, int N,
                                     void* stream) {
  MVPTR_PROF("gelu_bwd_colsum", 6.0*M*N, stream);
  if (M <= 0 || N <= 0) return 0;
  if (N & 7) MVPTR_FAIL(MVPTR_ERR_ARG, "gelu_bwd_colsum: N must be a multiple of 8");
  const int rpb = 128;
  dim3 grid((N + 255) / 256, (M + rpb - 1) / rpb);
  gelu_bwd_colsum_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)dy, (const bf16*)pre, (bf16*)dx, dbias, M,
                                                                 N, rpb);
  MVPTR_CHECK_LAUNCH("gelu_bwd_colsum");
  return 0;
}

extern "C" int mvptr_ln_bwd(const void* dy, int dy_rows_per_batch, long long dy_batch_stride, const void* x,
                            const float* mean, const float* rstd, const void* gamma, void* dx, void* dx_drop,
                            float* dgamma, float* dbeta, float* dbias, int rows, int H, float out_p_drop,
                            uint32_t out_seed, float in_p_drop, uint32_t in_seed, void* stream) {
  MVPTR_PROF("ln_bwd", 6.0*rows*H, stream);
  CHECK_H(H);
  if (rows == 0) return 0;
  if (dx_drop && in_p_drop <= 0.f) MVPTR_FAIL(MVPTR_ERR_ARG, "ln_bwd: dx_drop given without in_p_drop");
  RowMap dm{dy_rows_per_batch > 0 ? dy_rows_per_batch : rows, dy_rows_per_batch > 0 ? dy_batch_stride : 0};
  if (!dgamma || !dbeta) MVPTR_FAIL(MVPTR_ERR_ARG, "ln_bwd: dgamma/dbeta accumulators are required");
  if (dbias && !dx && !dx_drop) MVPTR_FAIL(MVPTR_ERR_ARG, "ln_bwd: dbias needs dx or dx_drop to be written");
  const int nc = (H + 255) / 256;
  int grid = (rows + 7) / 8;
  if (grid > 2 * kNumSMs) grid = 2 * kNumSMs;  // persistent: 2 CTAs per SM
  const int smem = 3 * H * (int)sizeof(float);
#define LN_BWD_LAUNCH(NC)                                                                                          \
  ln_bwd_kernel<NC><<<grid, 256, smem, (cudaStream_t)stream>>>(                                                    \
      (const bf16*)dy, dm, (const bf16*)x, mean, rstd, (const bf16*)gamma, (bf16*)dx, (bf16*)dx_drop, dgamma, dbeta, \
      dbias, rows, H, thr(out_p_drop), invk(out_p_drop), out_seed, thr(in_p_drop), invk(in_p_drop), in_seed)
  switch (nc) {
    case 1: LN_BWD_LAUNCH(1); break;
    case 2: LN_BWD_LAUNCH(2); break;
    case 3: LN_BWD_LAUNCH(3); break;
    default: LN_BWD_LAUNCH(4); break;
  }
#undef LN_BWD_LAUNCH
  MVPTR_CHECK_LAUNCH("ln_bwd");
  return 0;
}

extern "C" int mvptr_colsum(const void* x, int ldx, float* out, int M, int N, void* stream) {
  MVPTR_PROF("colsum", 2.0*M*N, stream);
  if (M <= 0 || N <= 0) return 0;
  if ((N & 7) || (ldx & 7)) MVPTR_FAIL(MVPTR_ERR_ARG, "colsum: N and ldx must be multiples of 8");
  const int rpb = 512;
  dim3 grid((N + 255) / 256, (M + rpb - 1) / rpb);
  colsum_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, ldx, out, M, N, rpb);
  MVPTR_CHECK_LAUNCH("colsum");
  return 0;
}

extern "C" int mvptr_embed_bwd(const void* dpre, const int64_t* ids, const int64_t* type_ids, float* dword,
                               float* dpos, float* dtype, int B, int L, int H, int vocab, int n_types,
                               int padding_idx, void* stream) {
  MVPTR_PROF("embed_bwd", 0, stream);
  CHECK_H(H);
  const int rows = B * L;
  if (rows == 0) return 0;
  embed_word_bwd_kernel<<<(rows + 3) / 4, 128, 0, (cudaStream_t)stream>>>((const bf16*)dpre, ids, dword, rows, H,
                                                                          vocab, padding_idx);
  MVPTR_CHECK_LAUNCH("embed_word_bwd");
  // L alone (40 / 20 blocks) would leave most SMs idle: also split the batch
  const int bsplit = B >= 64 ? 16 : (B >= 8 ? 4 : 1);
  embed_pos_type_bwd_kernel<<<dim3(L, bsplit), 128, 0, (cudaStream_t)stream>>>((const bf16*)dpre, type_ids, dpos, dtype, B, L, H,
                                                                 n_types);
  MVPTR_CHECK_LAUNCH("embed_pos_type_bwd");
  return 0;
}

// src_kind: 0 = bf16, 1 = fp32, 2 = fp16
extern "C" int mvptr_pad_cast(const void* src, int src_kind, long long ld_src, void* dst, int ld_dst, int rows, int K,
                              void* stream) {
  MVPTR_PROF("pad_cast", 0, stream);
  if (rows <= 0) return 0;
  if (src_kind < 0 || src_kind > 2) MVPTR_FAIL(MVPTR_ERR_ARG, "pad_cast: src_kind %d (0 bf16, 1 fp32, 2 fp16)", src_kind);
  const bool src_is_f32 = src_kind == 1;
  const size_t pair_bytes = src_is_f32 ? 8 : 4;
  const bool wide = (K % 2 == 0) && (ld_src % 2 == 0) && (ld_dst % 8 == 0) &&
                    (reinterpret_cast<uintptr_t>(src) % pair_bytes == 0) && (reinterpret_cast<uintptr_t>(dst) % 16 == 0);
  if (wide) {
    dim3 grid(rows, (ld_dst / 8 + 255) / 256);
    if (src_is_f32)
      pad_cast8_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)src, ld_src, (bf16*)dst, ld_dst, rows, K);
    else if (src_kind == 2)
      pad_cast8_kernel<__half><<<grid, 256, 0, (cudaStream_t)stream>>>((const __half*)src, ld_src, (bf16*)dst, ld_dst, rows, K);
    else
      pad_cast8_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)src, ld_src, (bf16*)dst, ld_dst, rows, K);
    MVPTR_CHECK_LAUNCH("pad_cast");
    return 0;
  }
  dim3 grid(rows, (ld_dst + 255) / 256);
  if (src_is_f32)
    pad_cast_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)src, ld_src, (bf16*)dst, ld_dst, rows, K);
  else if (src_kind == 2)
    pad_cast_kernel<__half><<<grid, 256, 0, (cudaStream_t)stream>>>((const __half*)src, ld_src, (bf16*)dst, ld_dst, rows, K);
  else
    pad_cast_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)src, ld_src, (bf16*)dst, ld_dst, rows, K);
  MVPTR_CHECK_LAUNCH("pad_cast");
  return 0;
}

extern "C" int mvptr_mask_prepare(const int64_t* mask_a, int La, const int64_t* mask_b, int Lb, int b_col0,
                                  const int64_t* row_a, const int64_t* row_b, float* out, int rows, void* stream) {
  if (rows <= 0) return 0;
  if (Lb > 0 && !mask_b) MVPTR_FAIL(MVPTR_ERR_ARG, "mask_prepare: mask_b missing");
  const int n = rows * (La + (Lb - b_col0));
  mask_prepare_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(mask_a, La, mask_b, Lb, b_col0, row_a, row_b,
                                                                         out, rows);
  MVPTR_CHECK_LAUNCH("mask_prepare");
  return 0;
}

extern "C" int mvptr_concat_rows(const void* a, int La, const void* b, int Lb, int b_col0, const int64_t* row_a,
                                 const int64_t* row_b, void* out, int rows, int H, void* stream) {
  MVPTR_PROF("concat_rows", 0, stream);
  if (rows <= 0) return 0;
  if (H & 7) MVPTR_FAIL(MVPTR_ERR_ARG, "concat_rows: H must be a multiple of 8");
  const int n = rows * (La + (Lb - b_col0));
  concat_rows_kernel<<<(n + 3) / 4, 128, 0, (cudaStream_t)stream>>>((const bf16*)a, La, (const bf16*)b, Lb, b_col0,
                                                                    row_a, row_b, (bf16*)out, rows, H);
  MVPTR_CHECK_LAUNCH("concat_rows");
  return 0;
}

extern "C" int mvptr_concat_rows_bwd(const void* dout, int La, int Lb, int b_col0, const int64_t* row_a,
                                     const int64_t* row_b, float* da, float* db, int rows, int H, void* stream) {
  MVPTR_PROF("concat_rows_bwd", 0, stream);
  if (rows <= 0) return 0;
  const int n = rows * (La + (Lb - b_col0));
  concat_rows_bwd_kernel<<<(n + 3) / 4, 128, 0, (cudaStream_t)stream>>>((const bf16*)dout, La, Lb, b_col0, row_a,
                                                                        row_b, da, db, rows, H);
  MVPTR_CHECK_LAUNCH("concat_rows_bwd");
  return 0;
}

extern "C" int mvptr_gather_rows(const void* src, const int64_t* idx, void* out, int n, int H, void* stream) {
  MVPTR_PROF("gather_rows", 0, stream);
  if (n <= 0) return 0;
  gather_rows_kernel<<<(n + 3) / 4, 128, 0, (cudaStream_t)stream>>>((const bf16*)src, idx, (bf16*)out, n, H);
  MVPTR_CHECK_LAUNCH("gather_rows");
  return 0;
}

extern "C" int mvptr_scatter_rows_add(const void* src, const int64_t* idx, float* dst, int n, int H, void* stream) {
  if (n <= 0) return 0;
  scatter_rows_add_kernel<<<(n + 3) / 4, 128, 0, (cudaStream_t)stream>>>((const bf16*)src, idx, dst, n, H);
  MVPTR_CHECK_LAUNCH("scatter_rows_add");
  return 0;
}

extern "C" int mvptr_cast_f32_bf16(const float* src, void* dst, size_t n, void* stream) {
  MVPTR_PROF("cast_f32_bf16", 6.0*n, stream);
  if (n == 0) return 0;
  size_t blocks = (n / 8 + 255) / 256;
  if (blocks > (size_t)kNumSMs * 8) blocks = (size_t)kNumSMs * 8;
  if (blocks == 0) blocks = 1;
  cast_f32_to_bf16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, (bf16*)dst, n);
  MVPTR_CHECK_LAUNCH("cast_f32_bf16");
  return 0;
}

/* max_ctas > 0 caps the grid: the gradient casts run on the communication stream NEXT TO the backward GEMMs and
 * should not take more than a few SMs' worth of issue slots at a time */
extern "C" int mvptr_cast_bf16_f32(const void* src, float* dst, size_t n, int max_ctas, void* stream) {
  MVPTR_PROF("cast_bf16_f32", 6.0*n, stream);
  if (n == 0) return 0;
  if ((reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(dst) & 15))
    MVPTR_FAIL(MVPTR_ERR_ARG, "cast_bf16_f32: buffers must be 16-byte aligned");
  size_t blocks = (n / 8 + 255) / 256;
  const size_t cap = max_ctas > 0 ? (size_t)max_ctas : (size_t)kNumSMs * 8;
  if (blocks > cap) blocks = cap;
  if (blocks == 0) blocks = 1;
  cast_bf16_to_f32_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const bf16*)src, dst, n);
  MVPTR_CHECK_LAUNCH("cast_bf16_f32");
  return 0;
}

extern "C" int mvptr_add_cast(const float* a, const void* b, void* d, size_t n, void* stream) {
  MVPTR_PROF("add_cast", 6.0*n, stream);
  if (n == 0) return 0;
  if (n & 7) MVPTR_FAIL(MVPTR_ERR_ARG, "add_cast: n must be a multiple of 8");
  size_t blocks = (n / 8 + 255) / 256;
  if (blocks > (size_t)kNumSMs * 8) blocks = (size_t)kNumSMs * 8;
  add_cast_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a, (const bf16*)b, (bf16*)d, n);
  MVPTR_CHECK_LAUNCH("add_cast");
  return 0;
}

MVPTR_DEFINE_EPOCH_SETTER(mvptr_set_epoch_rows)
