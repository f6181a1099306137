// Post-norm transformer encoder layer pieces for CNNTransformer (reference src/cnn_transformer.py:25-31, i.e.
// nn.TransformerEncoderLayer(batch_first=True): x = LN1(x + drop(MHA(x))); x = LN2(x + drop(W2 drop(relu(W1 x))))).
// The linear layers run on the tcgen05 1x1 path (pcm_conv1x1_tc / pcm_wgrad1x1_tc); here: residual + LayerNorm and
// multi-head attention over L = 216 tokens (12 x 18 grid) with head dim 32, fp32 math on bf16/fp32 storage.
#include "common.cuh"

namespace pcm {

__device__ __forceinline__ float hash_uniform_t(unsigned long long seed, unsigned long long idx) {
  return dropout_uniform(seed, idx);      // common.cuh: one definition for every mask-drawing kernel
}

constexpr int kLnMaxVec = 4;   // E <= 32 lanes * 4 vectors * 8 = 1024

// y = LN(a + b) * gamma + beta ; sum_out = a + b (saved for backward) ; stat[m] = (mean, rstd).  One warp per row.
// LayerNorm(a + dropout(b)) over the last dim.  A token's E/8 vectors are spread over `tl` = min(32, E/8) lanes, so a
// warp processes 32/tl tokens at once (E = 128: two tokens per warp, every lane busy).  drop_p > 0: b goes through
// nn.Dropout's counter-based mask (element index m*E + c of stream `seed`, the same mask pcm_dropout draws), so
// the transformer layer's dropout -> residual add -> LayerNorm is ONE pass.
template <typename T>
__global__ void __launch_bounds__(256)
add_layernorm_fwd_kernel(const T* __restrict__ a, const T* __restrict__ b, const float* __restrict__ gamma,
                         const float* __restrict__ beta, T* __restrict__ sum_out, T* __restrict__ y,
                         float* __restrict__ stat, int M, int E, float eps, float drop_p, unsigned long long seed,
                         const unsigned long long* __restrict__ epoch) {
  PCM_PDL_ENTRY();
  const int nv = E / 8;
  const int tl = nv < 32 ? nv : 32;                 // lanes per token (power of two)
  const int tpw = 32 / tl;                          // tokens per warp
  const int lane = threadIdx.x & 31, sub = lane / tl, l = lane - sub * tl;
  const int wpb = blockDim.x >> 5;
  const float keep_sc = 1.f / (1.f - drop_p);
  if (drop_p > 0.f) seed = mix_epoch(seed, epoch);
  for (int m0 = (blockIdx.x * wpb + (threadIdx.x >> 5)) * tpw; m0 < M; m0 += gridDim.x * wpb * tpw) {
    const int m = m0 + sub;
    const bool tok = m < M;
    float v[kLnMaxVec][8];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kLnMaxVec; ++k) {
      const int vi = l + tl * k;
      if (vi < nv && tok) {
        load8(a + (long long)m * E + vi * 8, v[k]);
        if (b != nullptr) {
          float t[8];
          load8(b + (long long)m * E + vi * 8, t);
          if (drop_p > 0.f) {
            const unsigned long long i0 = (unsigned long long)m * E + vi * 8;
#pragma unroll
            for (int j = 0; j < 8; ++j) t[j] = dropout_uniform(seed, i0 + j) >= drop_p ? round_to<T>(t[j] * keep_sc) : 0.f;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) v[k][j] = round_to<T>(v[k][j] + t[j]);
        }
        if (sum_out != nullptr) store8(sum_out + (long long)m * E + vi * 8, v[k]);
#pragma unroll
        for (int j = 0; j < 8; ++j) s += v[k][j];
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[k][j] = 0.f;
      }
    }
    for (int o = tl >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)E;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < kLnMaxVec; ++k) {
      if (l + tl * k < nv) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = v[k][j] - mean; q = fmaf(d, d, q); }
      }
    }
    for (int o = tl >> 1; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / (float)E + eps);
    if (l == 0 && tok) { stat[2 * m] = mean; stat[2 * m + 1] = rstd; }
#pragma unroll
    for (int k = 0; k < kLnMaxVec; ++k) {
      const int vi = l + tl * k;
      if (vi < nv && tok) {
        float g[8], bt[8], o[8];
        load8(gamma + vi * 8, g);
        load8(beta + vi * 8, bt);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf((v[k][j] - mean) * rstd, g[j], bt[j]);
        store8(y + (long long)m * E + vi * 8, o);
      }
    }
  }
}

// ds = rstd*(g - mean(g) - xhat*mean(g*xhat)), g = dy*gamma ; dgamma += sum dy*xhat ; dbeta += sum dy.
// drop_p > 0: additionally db = dropout(ds) with the forward's mask (the gradient of the dropped sub-layer output), so
// the separate dropout launch of the backward disappears too.  Same token-per-sub-warp mapping as the forward; the
// per-channel sums are reduced through per-warp shared-memory rows (no shared-memory atomics).
template <typename T>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ dy2, const T* __restrict__ sum_in, const float* __restrict__ stat,
                     const float* __restrict__ gamma, T* __restrict__ ds, T* __restrict__ db, float* __restrict__ dgamma,
                     float* __restrict__ dbeta, int M, int E, float drop_p, unsigned long long seed,
                     const unsigned long long* __restrict__ epoch) {
  PCM_PDL_ENTRY();
  extern __shared__ float sh[];     // [warps][2][E]
  const int nv = E / 8;
  const int tl = nv < 32 ? nv : 32, tpw = 32 / tl;
  const int lane = threadIdx.x & 31, sub = lane / tl, l = lane - sub * tl;
  const int warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const float keep_sc = 1.f / (1.f - drop_p);
  if (drop_p > 0.f) seed = mix_epoch(seed, epoch);
  float ag[kLnMaxVec][8], ab[kLnMaxVec][8], gm[kLnMaxVec][8];
#pragma unroll
  for (int k = 0; k < kLnMaxVec; ++k) {
    const int vi = l + tl * k;
    if (vi < nv) load8(gamma + vi * 8, gm[k]);
#pragma unroll
    for (int j = 0; j < 8; ++j) ag[k][j] = ab[k][j] = 0.f;
  }
  for (int m0 = (blockIdx.x * wpb + warp) * tpw; m0 < M; m0 += gridDim.x * wpb * tpw) {
    const int m = m0 + sub;
    const bool tok = m < M;
    const float mean = tok ? stat[2 * m] : 0.f, rstd = tok ? stat[2 * m + 1] : 0.f;
    float g[kLnMaxVec][8], xh[kLnMaxVec][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < kLnMaxVec; ++k) {
      const int vi = l + tl * k;
#pragma unroll
      for (int j = 0; j < 8; ++j) g[k][j] = xh[k][j] = 0.f;
      if (vi < nv && tok) {
        float d[8], x[8];
        load8(dy + (long long)m * E + vi * 8, d);
        if (dy2 != nullptr) {                 // second addend of the incoming gradient (a residual fork's other branch),
          float d2[8];                        // summed here instead of by a launch of its own; rounded like pcm_add's output
          load8(dy2 + (long long)m * E + vi * 8, d2);
#pragma unroll
          for (int j = 0; j < 8; ++j) d[j] = round_to<T>(d[j] + d2[j]);
        }
        load8(sum_in + (long long)m * E + vi * 8, x);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          xh[k][j] = (x[j] - mean) * rstd;
          g[k][j] = d[j] * gm[k][j];
          s1 += g[k][j];
          s2 = fmaf(g[k][j], xh[k][j], s2);
          ag[k][j] = fmaf(d[j], xh[k][j], ag[k][j]);
          ab[k][j] += d[j];
        }
      }
    }
    for (int o = tl >> 1; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    s1 /= (float)E;
    s2 /= (float)E;
#pragma unroll
    for (int k = 0; k < kLnMaxVec; ++k) {
      const int vi = l + tl * k;
      if (vi < nv && tok) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = rstd * (g[k][j] - s1 - xh[k][j] * s2);
        store8(ds + (long long)m * E + vi * 8, o);
        if (db != nullptr) {
          const unsigned long long i0 = (unsigned long long)m * E + vi * 8;
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = dropout_uniform(seed, i0 + j) >= drop_p ? round_to<T>(o[j]) * keep_sc : 0.f;
          store8(db + (long long)m * E + vi * 8, o);
        }
      }
    }
  }
  // per-channel sums: combine the sub-warp token groups with shuffles, one shared-memory row per warp, then one
  // thread per channel adds the rows and issues a single global atomic
  for (int o = tl; o < 32; o <<= 1) {
#pragma unroll
    for (int k = 0; k < kLnMaxVec; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        ag[k][j] += __shfl_xor_sync(0xffffffffu, ag[k][j], o);
        ab[k][j] += __shfl_xor_sync(0xffffffffu, ab[k][j], o);
      }
  }
  if (sub == 0) {
#pragma unroll
    for (int k = 0; k < kLnMaxVec; ++k) {
      const int vi = l + tl * k;
      if (vi < nv) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          sh[(warp * 2 + 0) * E + vi * 8 + j] = ag[k][j];
          sh[(warp * 2 + 1) * E + vi * 8 + j] = ab[k][j];
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * E; i += blockDim.x) {
    const int which = i / E, c = i - which * E;
    float t = 0.f;
    for (int w = 0; w < wpb; ++w) t += sh[(w * 2 + which) * E + c];
    atomicAdd((which ? dbeta : dgamma) + c, t);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Multi-head self-attention, one CTA per (batch, head); every thread owns one query row (forward / dq) or one key
// row (dk, dv); K/V (or Q/dO) of the head are staged once in shared memory as fp32 and read as warp broadcasts.
// qkv: [B][L][3E] (q | k | v, head h = columns h*D..), out: [B][L][E], lse: [B][nh][L] fp32 (log-sum-exp of the
// scaled scores).  Dropout on the attention probabilities (nn.MultiheadAttention dropout) uses the counter-based
// mask keyed by ((b*nh+h)*L+i)*L+j, regenerated in backward.
// ---------------------------------------------------------------------------------------------------------------
template <typename T, int D>
__global__ void __launch_bounds__(256)
mha_fwd_kernel(const T* __restrict__ qkv, T* __restrict__ out, float* __restrict__ lse, int L, int nh, float scale,
               float drop_p, unsigned long long seed, const unsigned long long* __restrict__ epoch) {
  PCM_PDL_ENTRY();
  seed = mix_epoch(seed, epoch);
  extern __shared__ float sm[];
  float* sK = sm;                 // [L][D]
  float* sV = sm + (size_t)L * D; // [L][D]
  const int bh = blockIdx.x, b = bh / nh, h = bh % nh;
  const int E = nh * D;
  const T* base = qkv + (long long)b * L * 3 * E + h * D;
  for (int idx = threadIdx.x; idx < L * (D / 8); idx += blockDim.x) {
    const int j = idx / (D / 8), c = (idx % (D / 8)) * 8;
    float t[8];
    load8(base + (long long)j * 3 * E + E + c, t);
#pragma unroll
    for (int k = 0; k < 8; ++k) sK[j * D + c + k] = t[k];
    load8(base + (long long)j * 3 * E + 2 * E + c, t);
#pragma unroll
    for (int k = 0; k < 8; ++k) sV[j * D + c + k] = t[k];
  }
  __syncthreads();
  const float keep_sc = 1.f / (1.f - drop_p);
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    float q[D], acc[D];
#pragma unroll
    for (int c = 0; c < D; c += 8) {
      float t[8];
      load8(base + (long long)i * 3 * E + c, t);
#pragma unroll
      for (int k = 0; k < 8; ++k) { q[c + k] = t[k] * scale; acc[c + k] = 0.f; }
    }
    float mx = -INFINITY, l = 0.f;
    for (int j0 = 0; j0 < L; j0 += 8) {
      float s[8];
      float cm = -INFINITY;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const int j = j0 + jj;
        float d = 0.f;
        if (j < L) {
          const float4* kp = reinterpret_cast<const float4*>(sK + j * D);
#pragma unroll
          for (int c = 0; c < D / 4; ++c) {
            const float4 kk = kp[c];
            d = fmaf(q[4 * c], kk.x, d); d = fmaf(q[4 * c + 1], kk.y, d);
            d = fmaf(q[4 * c + 2], kk.z, d); d = fmaf(q[4 * c + 3], kk.w, d);
          }
        } else {
          d = -INFINITY;
        }
        s[jj] = d;
        cm = fmaxf(cm, d);
      }
      const float mn = fmaxf(mx, cm);
      const float corr = __expf(mx - mn);
      l *= corr;
#pragma unroll
      for (int c = 0; c < D; ++c) acc[c] *= corr;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const int j = j0 + jj;
        if (j < L) {
          float p = __expf(s[jj] - mn);
          l += p;
          if (drop_p > 0.f)
            p = hash_uniform_t(seed, ((unsigned long long)bh * L + i) * L + j) >= drop_p ? p * keep_sc : 0.f;
          const float4* vp = reinterpret_cast<const float4*>(sV + j * D);
#pragma unroll
          for (int c = 0; c < D / 4; ++c) {
            const float4 vv = vp[c];
            acc[4 * c] = fmaf(p, vv.x, acc[4 * c]); acc[4 * c + 1] = fmaf(p, vv.y, acc[4 * c + 1]);
            acc[4 * c + 2] = fmaf(p, vv.z, acc[4 * c + 2]); acc[4 * c + 3] = fmaf(p, vv.w, acc[4 * c + 3]);
          }
        }
      }
      mx = mn;
    }
    const float inv = 1.f / l;
    T* op = out + ((long long)b * L + i) * E + h * D;
#pragma unroll
    for (int c = 0; c < D; c += 8) {
      float t[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) t[k] = acc[c + k] * inv;
      store8(op + c, t);
    }
    lse[(long long)bh * L + i] = mx + __logf(l);
  }
}

// backward: dq (thread = query row), then dk/dv (thread = key row); all five operands staged in smem as fp32
template <typename T, int D>
__global__ void __launch_bounds__(256)
mha_bwd_kernel(const T* __restrict__ qkv, const T* __restrict__ out, const T* __restrict__ dout,
               const float* __restrict__ lse, T* __restrict__ dqkv, int L, int nh, float scale, float drop_p,
               unsigned long long seed, const unsigned long long* __restrict__ epoch) {
  PCM_PDL_ENTRY();
  seed = mix_epoch(seed, epoch);
  extern __shared__ float sm[];
  float* sQ = sm;                       // [L][D]  (pre-scaled by `scale`)
  float* sK = sQ + (size_t)L * D;
  float* sV = sK + (size_t)L * D;
  float* sdO = sV + (size_t)L * D;
  float* sLse = sdO + (size_t)L * D;    // [L]
  float* sDl = sLse + L;                // [L]  D_i = dO_i . O_i
  const int bh = blockIdx.x, b = bh / nh, h = bh % nh;
  const int E = nh * D;
  const T* base = qkv + (long long)b * L * 3 * E + h * D;
  for (int idx = threadIdx.x; idx < L * (D / 8); idx += blockDim.x) {
    const int j = idx / (D / 8), c = (idx % (D / 8)) * 8;
    float t[8];
    load8(base + (long long)j * 3 * E + c, t);
#pragma unroll
    for (int k = 0; k < 8; ++k) sQ[j * D + c + k] = t[k] * scale;
    load8(base + (long long)j * 3 * E + E + c, t);
#pragma unroll
    for (int k = 0; k < 8; ++k) sK[j * D + c + k] = t[k];
    load8(base + (long long)j * 3 * E + 2 * E + c, t);
#pragma unroll
    for (int k = 0; k < 8; ++k) sV[j * D + c + k] = t[k];
    load8(dout + ((long long)b * L + j) * E + h * D + c, t);
#pragma unroll
    for (int k = 0; k < 8; ++k) sdO[j * D + c + k] = t[k];
  }
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    sLse[i] = lse[(long long)bh * L + i];
    const T* op = out + ((long long)b * L + i) * E + h * D;
    const T* dp = dout + ((long long)b * L + i) * E + h * D;
    float d = 0.f;
#pragma unroll
    for (int c = 0; c < D; c += 8) {
      float o[8], g[8];
      load8(op + c, o); load8(dp + c, g);
#pragma unroll
      for (int k = 0; k < 8; ++k) d = fmaf(o[k], g[k], d);
    }
    sDl[i] = d;
  }
  __syncthreads();
  const float keep_sc = 1.f / (1.f - drop_p);
  T* dbase = dqkv + (long long)b * L * 3 * E + h * D;
  // ---- dq_i = scale * sum_j ds_ij k_j
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    float q[D], g[D], acc[D];
#pragma unroll
    for (int c = 0; c < D; ++c) { q[c] = sQ[i * D + c]; g[c] = sdO[i * D + c]; acc[c] = 0.f; }
    const float li = sLse[i], Di = sDl[i];
    for (int j = 0; j < L; ++j) {
      const float4* kp = reinterpret_cast<const float4*>(sK + j * D);
      const float4* vp = reinterpret_cast<const float4*>(sV + j * D);
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int c = 0; c < D / 4; ++c) {
        const float4 kk = kp[c], vv = vp[c];
        s = fmaf(q[4 * c], kk.x, s); s = fmaf(q[4 * c + 1], kk.y, s);
        s = fmaf(q[4 * c + 2], kk.z, s); s = fmaf(q[4 * c + 3], kk.w, s);
        dp = fmaf(g[4 * c], vv.x, dp); dp = fmaf(g[4 * c + 1], vv.y, dp);
        dp = fmaf(g[4 * c + 2], vv.z, dp); dp = fmaf(g[4 * c + 3], vv.w, dp);
      }
      const float p = __expf(s - li);
      if (drop_p > 0.f)
        dp = hash_uniform_t(seed, ((unsigned long long)bh * L + i) * L + j) >= drop_p ? dp * keep_sc : 0.f;
      const float ds = p * (dp - Di);
#pragma unroll
      for (int c = 0; c < D / 4; ++c) {
        const float4 kk = kp[c];
        acc[4 * c] = fmaf(ds, kk.x, acc[4 * c]); acc[4 * c + 1] = fmaf(ds, kk.y, acc[4 * c + 1]);
        acc[4 * c + 2] = fmaf(ds, kk.z, acc[4 * c + 2]); acc[4 * c + 3] = fmaf(ds, kk.w, acc[4 * c + 3]);
      }
    }
#pragma unroll
    for (int c = 0; c < D; c += 8) {
      float t[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) t[k] = acc[c + k] * scale;
      store8(dbase + (long long)i * 3 * E + c, t);
    }
  }
  // ---- dk_j = sum_i ds_ij (scale q_i) ; dv_j = sum_i p~_ij dO_i
  for (int j = threadIdx.x; j < L; j += blockDim.x) {
    float kx[D], vx[D], dk[D], dv[D];
#pragma unroll
    for (int c = 0; c < D; ++c) { kx[c] = sK[j * D + c]; vx[c] = sV[j * D + c]; dk[c] = 0.f; dv[c] = 0.f; }
    for (int i = 0; i < L; ++i) {
      const float4* qp = reinterpret_cast<const float4*>(sQ + i * D);
      const float4* gp = reinterpret_cast<const float4*>(sdO + i * D);
      float s = 0.f, dp = 0.f;
#pragma unroll
      for (int c = 0; c < D / 4; ++c) {
        const float4 qq = qp[c], gg = gp[c];
        s = fmaf(qq.x, kx[4 * c], s); s = fmaf(qq.y, kx[4 * c + 1], s);
        s = fmaf(qq.z, kx[4 * c + 2], s); s = fmaf(qq.w, kx[4 * c + 3], s);
        dp = fmaf(gg.x, vx[4 * c], dp); dp = fmaf(gg.y, vx[4 * c + 1], dp);
        dp = fmaf(gg.z, vx[4 * c + 2], dp); dp = fmaf(gg.w, vx[4 * c + 3], dp);
      }
      const float p = __expf(s - sLse[i]);
      float pt = p;
      if (drop_p > 0.f) {
        const float ks = hash_uniform_t(seed, ((unsigned long long)bh * L + i) * L + j) >= drop_p ? keep_sc : 0.f;
        pt = p * ks;
        dp *= ks;
      }
      const float ds = p * (dp - sDl[i]);
#pragma unroll
      for (int c = 0; c < D / 4; ++c) {
        const float4 qq = qp[c], gg = gp[c];
        dk[4 * c] = fmaf(ds, qq.x, dk[4 * c]); dk[4 * c + 1] = fmaf(ds, qq.y, dk[4 * c + 1]);
        dk[4 * c + 2] = fmaf(ds, qq.z, dk[4 * c + 2]); dk[4 * c + 3] = fmaf(ds, qq.w, dk[4 * c + 3]);
        dv[4 * c] = fmaf(pt, gg.x, dv[4 * c]); dv[4 * c + 1] = fmaf(pt, gg.y, dv[4 * c + 1]);
        dv[4 * c + 2] = fmaf(pt, gg.z, dv[4 * c + 2]); dv[4 * c + 3] = fmaf(pt, gg.w, dv[4 * c + 3]);
      }
    }
#pragma unroll
    for (int c = 0; c < D; c += 8) {
      store8(dbase + (long long)j * 3 * E + E + c, dk + c);       // sQ is pre-scaled: dk already carries `scale`
      store8(dbase + (long long)j * 3 * E + 2 * E + c, dv + c);
    }
  }
}

template <typename K>
static int set_smem(K kern, size_t bytes, const char* what) {
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) { set_error("%s: smem attribute (%zu B): %s", what, bytes, cudaGetErrorString(e)); return PCM_ERR_CUDA; }
  }
  return PCM_OK;
}

}  // namespace pcm

using namespace pcm;

static bool ln_shape_ok(int E) {
  const int nv = E / 8;
  return E % 8 == 0 && E >= 8 && E <= 32 * kLnMaxVec * 8 && (nv >= 32 || (nv & (nv - 1)) == 0);
}

extern "C" int pcm_add_layernorm_fwd(const void* a, const void* b, const float* gamma, const float* beta, void* sum_out,
                                     void* y, float* stat, int M, int E, float eps, float drop_p, long long seed,
                                     int dtype, pcm_stream_t s) {
  PCM_REQUIRE(ln_shape_ok(E), "add_layernorm_fwd: E must be 8..256 (power of two) or a multiple of 256 up to 1024 (got %d)", E);
  PCM_REQUIRE(drop_p >= 0.f && drop_p < 1.f && (drop_p == 0.f || b != nullptr), "add_layernorm_fwd: 0 <= drop_p < 1, dropout needs b");
  if (M == 0) return PCM_OK;
  const int nv = E / 8, tpw = nv < 32 ? 32 / nv : 1;
  int grid = ceil_div(M, 8 * tpw);
  if (grid > 148 * 8) grid = 148 * 8;
  const unsigned long long* epoch = dropout_epoch_cell();
  PCM_REQUIRE(epoch != nullptr, "add_layernorm_fwd: could not allocate the dropout epoch cell");
  PCM_DISPATCH_DTYPE(dtype, T, (pcm::launch(add_layernorm_fwd_kernel<T>, grid, 256, 0, (cudaStream_t)s,
                                   static_cast<const T*>(a), static_cast<const T*>(b), gamma, beta, static_cast<T*>(sum_out),
                                   static_cast<T*>(y), stat, M, E, eps, drop_p, (unsigned long long)seed, epoch)));
  return check_launch("add_layernorm_fwd");
}

extern "C" int pcm_layernorm_bwd2(const void* dy, const void* dy2, const void* sum_in, const float* stat, const float* gamma,
                                  void* ds, void* db, float* dgamma, float* dbeta, int M, int E, float drop_p, long long seed,
                                  int dtype, pcm_stream_t s);
extern "C" int pcm_layernorm_bwd(const void* dy, const void* sum_in, const float* stat, const float* gamma, void* ds,
                                 void* db, float* dgamma, float* dbeta, int M, int E, float drop_p, long long seed,
                                 int dtype, pcm_stream_t s) {
  return pcm_layernorm_bwd2(dy, nullptr, sum_in, stat, gamma, ds, db, dgamma, dbeta, M, E, drop_p, seed, dtype, s);
}

extern "C" int pcm_layernorm_bwd2(const void* dy, const void* dy2, const void* sum_in, const float* stat, const float* gamma,
                                  void* ds, void* db, float* dgamma, float* dbeta, int M, int E, float drop_p, long long seed,
                                  int dtype, pcm_stream_t s) {
  PCM_REQUIRE(ln_shape_ok(E), "layernorm_bwd: E must be 8..256 (power of two) or a multiple of 256 up to 1024 (got %d)", E);
  PCM_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "layernorm_bwd: 0 <= drop_p < 1");
  PCM_REQUIRE(E <= 768, "layernorm_bwd: E up to 768 (per-warp reduction rows in shared memory), got %d", E);
  if (M == 0) return PCM_OK;
  const int nv = E / 8, tpw = nv < 32 ? 32 / nv : 1;
  // two token groups per warp on average: enough CTAs to fill the machine, few enough global atomics (2E per CTA)
  int grid = ceil_div(M, 8 * tpw * 2);
  if (grid > 148 * 4) grid = 148 * 4;
  const unsigned long long* epoch = dropout_epoch_cell();
  PCM_REQUIRE(epoch != nullptr, "layernorm_bwd: could not allocate the dropout epoch cell");
  PCM_DISPATCH_DTYPE(dtype, T, (pcm::launch(layernorm_bwd_kernel<T>, grid, 256, 8 * 2 * E * sizeof(float), (cudaStream_t)s,
                                   static_cast<const T*>(dy), static_cast<const T*>(dy2), static_cast<const T*>(sum_in), stat, gamma,
                                   static_cast<T*>(ds),
                                   drop_p > 0.f ? static_cast<T*>(db) : nullptr, dgamma, dbeta, M, E, drop_p,
                                   (unsigned long long)seed, epoch)));
  return check_launch("layernorm_bwd");
}

template <typename T, int D>
static int mha_fwd_launch(const void* qkv, void* out, float* lse, int B, int L, int nh, float scale, float p,
                          long long seed, pcm_stream_t s) {
  const size_t smem = (size_t)2 * L * D * sizeof(float);
  int rc = set_smem(mha_fwd_kernel<T, D>, smem, "mha_fwd");
  if (rc != PCM_OK) return rc;
  pcm::launch(mha_fwd_kernel<T, D>, B * nh, 256, smem, (cudaStream_t)s, static_cast<const T*>(qkv), static_cast<T*>(out), lse, L, nh,
                                                               scale, p, (unsigned long long)seed, dropout_epoch_cell());
  return check_launch("mha_fwd");
}

template <typename T, int D>
static int mha_bwd_launch(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int B, int L,
                          int nh, float scale, float p, long long seed, pcm_stream_t s) {
  const size_t smem = ((size_t)4 * L * D + 2 * L) * sizeof(float);
  int rc = set_smem(mha_bwd_kernel<T, D>, smem, "mha_bwd");
  if (rc != PCM_OK) return rc;
  pcm::launch(mha_bwd_kernel<T, D>, B * nh, 256, smem, (cudaStream_t)s, static_cast<const T*>(qkv), static_cast<const T*>(out),
                                                               static_cast<const T*>(dout), lse, static_cast<T*>(dqkv), L,
                                                               nh, scale, p, (unsigned long long)seed, dropout_epoch_cell());
  return check_launch("mha_bwd");
}

#define MHA_DISPATCH(D, CALL32, CALL16, CALL64, CALL8)                                               \
  do {                                                                                               \
    if ((D) == 32) { CALL32; } else if ((D) == 16) { CALL16; } else if ((D) == 64) { CALL64; }        \
    else if ((D) == 8) { CALL8; }                                                                    \
    else { set_error("mha: head dim must be 8, 16, 32 or 64 (got %d)", (int)(D)); return PCM_ERR_INVALID; } \
  } while (0)

extern "C" int pcm_mha_fwd(const void* qkv, void* out, float* lse, int B, int L, int nh, int D, float scale, float drop_p,
                           long long seed, int dtype, pcm_stream_t s) {
  PCM_REQUIRE(B >= 0 && L >= 1 && nh >= 1 && drop_p >= 0.f && drop_p < 1.f, "mha_fwd: bad arguments");
  PCM_REQUIRE((size_t)2 * L * D * 4 <= 200 * 1024, "mha_fwd: L*D too large for shared memory (L=%d D=%d)", L, D);
  if (B == 0) return PCM_OK;
  int rc = PCM_OK;
  PCM_DISPATCH_DTYPE(dtype, T, MHA_DISPATCH(D, (rc = mha_fwd_launch<T, 32>(qkv, out, lse, B, L, nh, scale, drop_p, seed, s)),
                                            (rc = mha_fwd_launch<T, 16>(qkv, out, lse, B, L, nh, scale, drop_p, seed, s)),
                                            (rc = mha_fwd_launch<T, 64>(qkv, out, lse, B, L, nh, scale, drop_p, seed, s)),
                                            (rc = mha_fwd_launch<T, 8>(qkv, out, lse, B, L, nh, scale, drop_p, seed, s))));
  return rc;
}

extern "C" int pcm_mha_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int B, int L,
                           int nh, int D, float scale, float drop_p, long long seed, int dtype, pcm_stream_t s) {
  PCM_REQUIRE(B >= 0 && L >= 1 && nh >= 1 && drop_p >= 0.f && drop_p < 1.f, "mha_bwd: bad arguments");
  PCM_REQUIRE(((size_t)4 * L * D + 2 * L) * 4 <= 220 * 1024, "mha_bwd: L*D too large for shared memory (L=%d D=%d)", L, D);
  if (B == 0) return PCM_OK;
  int rc = PCM_OK;
  PCM_DISPATCH_DTYPE(dtype, T,
                     MHA_DISPATCH(D, (rc = mha_bwd_launch<T, 32>(qkv, out, dout, lse, dqkv, B, L, nh, scale, drop_p, seed, s)),
                                  (rc = mha_bwd_launch<T, 16>(qkv, out, dout, lse, dqkv, B, L, nh, scale, drop_p, seed, s)),
                                  (rc = mha_bwd_launch<T, 64>(qkv, out, dout, lse, dqkv, B, L, nh, scale, drop_p, seed, s)),
                                  (rc = mha_bwd_launch<T, 8>(qkv, out, dout, lse, dqkv, B, L, nh, scale, drop_p, seed, s))));
  return rc;
}
