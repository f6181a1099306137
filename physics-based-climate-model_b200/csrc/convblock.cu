// ConvBlock tail kernels: GroupNorm(8)+SiLU, squeeze-excitation, CBAM spatial gate — forward and
// backward (reference: src/unet.py:6-29, 35-49).  All memory-bound; NHWC, 8-channel vectors.
#include "common.cuh"

namespace pcm {

// Thread decomposition shared by the per-channel kernels: a block covers pixels [p0,p1) of one
// image; thread -> (cb = tid % cv, lane = tid / cv) so consecutive threads read consecutive 16 B.
struct PixSplit {
  int p0, p1, cb, lane, lanes;
  bool active;
};
__device__ __forceinline__ PixSplit pix_split(int P, int cv) {
  PixSplit s;
  const int per = (P + gridDim.x - 1) / gridDim.x;
  s.p0 = blockIdx.x * per;
  s.p1 = min(P, s.p0 + per);
  s.lanes = blockDim.x / cv;
  s.cb = threadIdx.x % cv;
  s.lane = threadIdx.x / cv;
  s.active = s.lane < s.lanes;
  return s;
}

__device__ __forceinline__ void group_mean_rstd(const float* __restrict__ stats, int n, int G, int c, int cg,
                                                float cnt, float eps, float& mu, float& rs) {
  const int g = c / cg;
  const float s = __ldg(stats + ((long long)n * G + g) * 2);
  const float ss = __ldg(stats + ((long long)n * G + g) * 2 + 1);
  mu = s / cnt;
  const float var = fmaxf(ss / cnt - mu * mu, 0.f);
  rs = rsqrtf(var + eps);
}

// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
gn_stats_kernel(const T* __restrict__ x, float* __restrict__ stats, int P, int C, int G) {
  __shared__ float sg[64 * 2];
  const int n = blockIdx.y, cv = C / 8, cg = C / G;
  for (int i = threadIdx.x; i < G * 2; i += blockDim.x) sg[i] = 0.f;
  __syncthreads();
  const PixSplit ps = pix_split(P, cv);
  if (ps.active) {
    float s[8], ss[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = ss[j] = 0.f;
    const T* xn = x + (long long)n * P * C + ps.cb * 8;
    for (int p = ps.p0 + ps.lane; p < ps.p1; p += ps.lanes) {
      float v[8];
      load8(xn + (long long)p * C, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) { s[j] += v[j]; ss[j] = fmaf(v[j], v[j], ss[j]); }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = (ps.cb * 8 + j) / cg;
      atomicAdd(&sg[g * 2], s[j]);
      atomicAdd(&sg[g * 2 + 1], ss[j]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < G * 2; i += blockDim.x) atomicAdd(stats + (long long)n * G * 2 + i, sg[i]);
}

template <typename T>
__global__ void __launch_bounds__(256)
gn_silu_fwd_kernel(const T* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ gamma,
                   const float* __restrict__ beta, T* __restrict__ y, float* __restrict__ pool, int P, int C, int G,
                   float eps) {
  extern __shared__ float spool[];   // [C]
  const int n = blockIdx.y, cv = C / 8, cg = C / G;
  if (pool) {
    for (int i = threadIdx.x; i < C; i += blockDim.x) spool[i] = 0.f;
    __syncthreads();
  }
  const PixSplit ps = pix_split(P, cv);
  if (ps.active) {
    float mu[8], rs[8], ga[8], be[8], acc[8];
    const float cnt = (float)cg * (float)P;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = ps.cb * 8 + j;
      group_mean_rstd(stats, n, G, c, cg, cnt, eps, mu[j], rs[j]);
      ga[j] = __ldg(gamma + c); be[j] = __ldg(beta + c); acc[j] = 0.f;
    }
    const long long base = (long long)n * P * C + ps.cb * 8;
    for (int p = ps.p0 + ps.lane; p < ps.p1; p += ps.lanes) {
      float v[8];
      load8(x + base + (long long)p * C, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float z = fmaf(ga[j], (v[j] - mu[j]) * rs[j], be[j]);
        v[j] = round_to<T>(z * sigmoidf_(z));
        acc[j] += v[j];
      }
      store8(y + base + (long long)p * C, v);
    }
    if (pool) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&spool[ps.cb * 8 + j], acc[j]);
    }
  }
  if (pool) {
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(pool + (long long)n * C + i, spool[i]);
  }
}

// SE excitation (recomputed per block; split 0 publishes it) + per-pixel mean/max of a*se
template <typename T>
__global__ void __launch_bounds__(256)
se_chanstat_fwd_kernel(const T* __restrict__ a, const float* __restrict__ pool, const float* __restrict__ w1,
                       const float* __restrict__ w2, float* __restrict__ se, float* __restrict__ hid,
                       float* __restrict__ cmap, int P, int C, int Cr) {
  extern __shared__ float sm[];   // pm[C] | sh[Cr] | sse[C]
  float* pm = sm;
  float* sh = sm + C;
  float* sse = sh + Cr;
  const int n = blockIdx.y;
  const float invP = 1.f / (float)P;
  if (w1 == nullptr) {            // no excitation: se = 1 (stand-alone SpatialGate)
    for (int i = threadIdx.x; i < C; i += blockDim.x) sse[i] = 1.f;
    for (int j = threadIdx.x; j < Cr; j += blockDim.x) sh[j] = 0.f;
    __syncthreads();
  } else {
    for (int i = threadIdx.x; i < C; i += blockDim.x) pm[i] = __ldg(pool + (long long)n * C + i) * invP;
    __syncthreads();
    for (int j = threadIdx.x; j < Cr; j += blockDim.x) {
      float acc = 0.f;
      for (int c = 0; c < C; ++c) acc = fmaf(__ldg(w1 + (long long)j * C + c), pm[c], acc);
      sh[j] = fmaxf(acc, 0.f);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float acc = 0.f;
      for (int j = 0; j < Cr; ++j) acc = fmaf(__ldg(w2 + (long long)c * Cr + j), sh[j], acc);
      sse[c] = sigmoidf_(acc);
    }
    __syncthreads();
  }
  if (blockIdx.x == 0) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) se[(long long)n * C + c] = sse[c];
    for (int j = threadIdx.x; j < Cr; j += blockDim.x) hid[(long long)n * Cr + j] = sh[j];
  }
  const int per = (P + gridDim.x - 1) / gridDim.x;
  const int p0 = blockIdx.x * per, p1 = min(P, p0 + per);
  for (int p = p0 + threadIdx.x; p < p1; p += blockDim.x) {
    const T* ap = a + ((long long)n * P + p) * C;
    float sum = 0.f, mx = -INFINITY;
    for (int cb = 0; cb < C; cb += 8) {
      float v[8];
      load8(ap + cb, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float u = v[j] * sse[cb + j];
        sum += u;
        mx = fmaxf(mx, u);
      }
    }
    reinterpret_cast<float2*>(cmap)[(long long)n * P + p] = make_float2(sum / (float)C, mx);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
spatial_gate_fwd_kernel(const T* __restrict__ a, const float* __restrict__ se, const float* __restrict__ cmap,
                        const float* __restrict__ wsp, float* __restrict__ gate, T* __restrict__ out, int H, int W,
                        int C) {
  extern __shared__ float sm[];   // wsp[98] | sse[C]
  float* sw = sm;
  float* sse = sm + 98;
  const int n = blockIdx.y, P = H * W;
  for (int i = threadIdx.x; i < 98; i += blockDim.x) sw[i] = __ldg(wsp + i);
  for (int i = threadIdx.x; i < C; i += blockDim.x) sse[i] = __ldg(se + (long long)n * C + i);
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const int h = p / W, w = p % W;
  const float2* cm = reinterpret_cast<const float2*>(cmap) + (long long)n * P;
  float q = 0.f;
#pragma unroll
  for (int dy = 0; dy < 7; ++dy) {
    const int hh = h + dy - 3;
    if (hh < 0 || hh >= H) continue;
#pragma unroll
    for (int dx = 0; dx < 7; ++dx) {
      const int ww = w + dx - 3;
      if (ww < 0 || ww >= W) continue;
      const float2 m = __ldg(cm + hh * W + ww);
      q = fmaf(sw[dy * 7 + dx], m.x, q);
      q = fmaf(sw[49 + dy * 7 + dx], m.y, q);
    }
  }
  const float gt = sigmoidf_(q);
  gate[(long long)n * P + p] = gt;
  const T* ap = a + ((long long)n * P + p) * C;
  T* op = out + ((long long)n * P + p) * C;
  for (int cb = 0; cb < C; cb += 8) {
    float v[8];
    load8(ap + cb, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = v[j] * sse[cb + j] * gt;
    store8(op + cb, v);
  }
}

// ---- backward ---------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
spatial_gate_bwd_dq_kernel(const T* __restrict__ dout, const T* __restrict__ a, const float* __restrict__ se,
                           const float* __restrict__ gate, float* __restrict__ dq, int P, int C) {
  extern __shared__ float sse[];
  const int n = blockIdx.y;
  for (int i = threadIdx.x; i < C; i += blockDim.x) sse[i] = __ldg(se + (long long)n * C + i);
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const long long off = ((long long)n * P + p) * C;
  float acc = 0.f;
  for (int cb = 0; cb < C; cb += 8) {
    float d[8], v[8];
    load8(dout + off + cb, d);
    load8(a + off + cb, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc = fmaf(d[j], v[j] * sse[cb + j], acc);
  }
  const float gt = __ldg(gate + (long long)n * P + p);
  dq[(long long)n * P + p] = acc * gt * (1.f - gt);
}

__global__ void __launch_bounds__(256)
spatial_gate_bwd_dw_kernel(const float* __restrict__ dq, const float* __restrict__ cmap, float* __restrict__ dwsp,
                           int N, int H, int W) {
  __shared__ float sred[98];
  for (int i = threadIdx.x; i < 98; i += blockDim.x) sred[i] = 0.f;
  __syncthreads();
  float acc[98];
#pragma unroll
  for (int i = 0; i < 98; ++i) acc[i] = 0.f;
  const int P = H * W;
  const long long total = (long long)N * P;
  const float2* cm = reinterpret_cast<const float2*>(cmap);
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(idx % P);
    const long long nb = idx - p;
    const int h = p / W, w = p % W;
    const float d = __ldg(dq + idx);
#pragma unroll
    for (int dy = 0; dy < 7; ++dy) {
      const int hh = h + dy - 3;
#pragma unroll
      for (int dx = 0; dx < 7; ++dx) {
        const int ww = w + dx - 3;
        if (hh >= 0 && hh < H && ww >= 0 && ww < W) {
          const float2 m = __ldg(cm + nb + hh * W + ww);
          acc[dy * 7 + dx] = fmaf(d, m.x, acc[dy * 7 + dx]);
          acc[49 + dy * 7 + dx] = fmaf(d, m.y, acc[49 + dy * 7 + dx]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 98; ++i) {
    const float v = warp_sum(acc[i]);
    if ((threadIdx.x & 31) == 0) atomicAdd(&sred[i], v);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 98; i += blockDim.x) atomicAdd(dwsp + i, sred[i]);
}

template <typename T>
__global__ void __launch_bounds__(256)
spatial_gate_bwd_da_kernel(const T* __restrict__ dout, const T* __restrict__ a, const float* __restrict__ se,
                           const float* __restrict__ gate, const float* __restrict__ cmap,
                           const float* __restrict__ dq, const float* __restrict__ wsp, T* __restrict__ da,
                           float* __restrict__ dse, int H, int W, int C) {
  extern __shared__ float sm[];   // wsp[98] | sse[C] | sdse[C]
  float* sw = sm;
  float* sse = sm + 98;
  float* sdse = sse + C;
  const int n = blockIdx.y, P = H * W;
  for (int i = threadIdx.x; i < 98; i += blockDim.x) sw[i] = __ldg(wsp + i);
  for (int i = threadIdx.x; i < C; i += blockDim.x) { sse[i] = __ldg(se + (long long)n * C + i); sdse[i] = 0.f; }
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = p < P;
  float dm0 = 0.f, dm1 = 0.f, gt = 0.f, mx = 0.f;
  int nties = 1;
  long long off = 0;
  if (valid) {
    const int h = p / W, w = p % W;
    const float* dqn = dq + (long long)n * P;
    // cmap[p] feeds q[p'] with p' = p - (dy-3, dx-3), weight w[dy][dx]
#pragma unroll
    for (int dy = 0; dy < 7; ++dy) {
      const int hh = h - (dy - 3);
      if (hh < 0 || hh >= H) continue;
#pragma unroll
      for (int dx = 0; dx < 7; ++dx) {
        const int ww = w - (dx - 3);
        if (ww < 0 || ww >= W) continue;
        const float d = __ldg(dqn + hh * W + ww);
        dm0 = fmaf(sw[dy * 7 + dx], d, dm0);
        dm1 = fmaf(sw[49 + dy * 7 + dx], d, dm1);
      }
    }
    gt = __ldg(gate + (long long)n * P + p);
    mx = __ldg(cmap + ((long long)n * P + p) * 2 + 1);
    off = ((long long)n * P + p) * C;
    nties = 0;
    for (int cb = 0; cb < C; cb += 8) {
      float v[8];
      load8(a + off + cb, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) nties += (v[j] * sse[cb + j] == mx) ? 1 : 0;
    }
    if (nties < 1) nties = 1;
  }
  const float dmean = dm0 / (float)C;
  const float dmax = dm1 / (float)nties;
  for (int cb = 0; cb < C; cb += 8) {
    float d[8], v[8], r[8];
    if (valid) {
      load8(dout + off + cb, d);
      load8(a + off + cb, v);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) d[j] = v[j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float s = sse[cb + j];
      const float u = v[j] * s;
      float du = d[j] * gt + dmean + ((u == mx) ? dmax : 0.f);
      if (!valid) du = 0.f;
      r[j] = du * s;
      const float part = warp_sum(du * v[j]);
      if ((threadIdx.x & 31) == 0) atomicAdd(&sdse[cb + j], part);
    }
    if (valid) store8(da + off + cb, r);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(dse + (long long)n * C + i, sdse[i]);
}

__global__ void __launch_bounds__(128)
se_bwd_kernel(const float* __restrict__ dse, const float* __restrict__ se, const float* __restrict__ hid,
              const float* __restrict__ pool, const float* __restrict__ w1, const float* __restrict__ w2,
              float* __restrict__ dpool, float* __restrict__ dw1, float* __restrict__ dw2, int P, int C, int Cr) {
  extern __shared__ float sm[];   // dpre2[C] | dpre1[Cr] | sh[Cr] | pm[C]
  float* dpre2 = sm;
  float* dpre1 = sm + C;
  float* sh = dpre1 + Cr;
  float* pm = sh + Cr;
  const int n = blockIdx.x;
  const float invP = 1.f / (float)P;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float s = __ldg(se + (long long)n * C + c);
    dpre2[c] = __ldg(dse + (long long)n * C + c) * s * (1.f - s);
    pm[c] = __ldg(pool + (long long)n * C + c) * invP;
  }
  for (int j = threadIdx.x; j < Cr; j += blockDim.x) sh[j] = __ldg(hid + (long long)n * Cr + j);
  __syncthreads();
  for (int j = threadIdx.x; j < Cr; j += blockDim.x) {
    float acc = 0.f;
    for (int c = 0; c < C; ++c) acc = fmaf(__ldg(w2 + (long long)c * Cr + j), dpre2[c], acc);
    dpre1[j] = sh[j] > 0.f ? acc : 0.f;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < Cr; ++j) acc = fmaf(__ldg(w1 + (long long)j * C + c), dpre1[j], acc);
    dpool[(long long)n * C + c] = acc * invP;
  }
  for (int i = threadIdx.x; i < C * Cr; i += blockDim.x) {
    {  // dw2[c][j]
      const int c = i / Cr, j = i % Cr;
      const float v = dpre2[c] * sh[j];
      if (v != 0.f) atomicAdd(dw2 + i, v);
    }
    {  // dw1[j][c]
      const int j = i / C, c = i % C;
      const float v = dpre1[j] * pm[c];
      if (v != 0.f) atomicAdd(dw1 + i, v);
    }
  }
}

template <typename T, bool APPLY>
__global__ void __launch_bounds__(256)
gn_silu_bwd_kernel(const T* __restrict__ da, const float* __restrict__ dpool, const T* __restrict__ x,
                   const float* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta,
                   float* __restrict__ gsum, float* __restrict__ dgamma, float* __restrict__ dbeta,
                   T* __restrict__ dx, int P, int C, int G, float eps) {
  extern __shared__ float sm[];   // reduce: sgs[G*2] | sdg[C] | sdb[C]
  float* sgs = sm;
  float* sdg = sm + G * 2;
  float* sdb = sdg + C;
  const int n = blockIdx.y, cv = C / 8, cg = C / G;
  if (!APPLY) {
    for (int i = threadIdx.x; i < G * 2 + 2 * C; i += blockDim.x) sm[i] = 0.f;
    __syncthreads();
  }
  const PixSplit ps = pix_split(P, cv);
  if (ps.active) {
    float mu[8], rs[8], ga[8], be[8], dp[8], m1[8], m2[8];
    float a_dg[8], a_db[8], a_s1[8], a_s2[8];
    const float cnt = (float)cg * (float)P;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = ps.cb * 8 + j;
      group_mean_rstd(stats, n, G, c, cg, cnt, eps, mu[j], rs[j]);
      ga[j] = __ldg(gamma + c); be[j] = __ldg(beta + c);
      dp[j] = dpool ? __ldg(dpool + (long long)n * C + c) : 0.f;
      a_dg[j] = a_db[j] = a_s1[j] = a_s2[j] = 0.f;
      if (APPLY) {
        const int g = c / cg;
        m1[j] = __ldg(gsum + ((long long)n * G + g) * 2) / cnt;
        m2[j] = __ldg(gsum + ((long long)n * G + g) * 2 + 1) / cnt;
      }
    }
    const long long base = (long long)n * P * C + ps.cb * 8;
    for (int p = ps.p0 + ps.lane; p < ps.p1; p += ps.lanes) {
      float v[8], d[8];
      load8(x + base + (long long)p * C, v);
      load8(da + base + (long long)p * C, d);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float xh = (v[j] - mu[j]) * rs[j];
        const float z = fmaf(ga[j], xh, be[j]);
        const float sg = sigmoidf_(z);
        const float dz = (d[j] + dp[j]) * sg * (1.f + z * (1.f - sg));
        const float dxh = dz * ga[j];
        if (APPLY) {
          v[j] = rs[j] * (dxh - m1[j] - xh * m2[j]);
        } else {
          a_dg[j] = fmaf(dz, xh, a_dg[j]);
          a_db[j] += dz;
          a_s1[j] += dxh;
          a_s2[j] = fmaf(dxh, xh, a_s2[j]);
        }
      }
      if (APPLY) store8(dx + base + (long long)p * C, v);
    }
    if (!APPLY) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = ps.cb * 8 + j, g = c / cg;
        atomicAdd(&sdg[c], a_dg[j]);
        atomicAdd(&sdb[c], a_db[j]);
        atomicAdd(&sgs[g * 2], a_s1[j]);
        atomicAdd(&sgs[g * 2 + 1], a_s2[j]);
      }
    }
  }
  if (!APPLY) {
    __syncthreads();
    for (int i = threadIdx.x; i < G * 2; i += blockDim.x) atomicAdd(gsum + (long long)n * G * 2 + i, sgs[i]);
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
      atomicAdd(dgamma + i, sdg[i]);
      atomicAdd(dbeta + i, sdb[i]);
    }
  }
}

// out = x * scale[n][c] + add[n][c]  (either nullable)
template <typename T>
__global__ void __launch_bounds__(256)
scale_channels_kernel(const T* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ add,
                      T* __restrict__ out, int N, int P, int C) {
  const int cv = C / 8;
  const long long total = (long long)N * P * cv;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int cb = (int)(idx % cv);
    const int n = (int)(idx / ((long long)cv * P));
    float v[8];
    load8(x + idx * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const long long k = (long long)n * C + cb * 8 + j;
      v[j] = v[j] * (scale ? __ldg(scale + k) : 1.f) + (add ? __ldg(add + k) : 0.f);
    }
    store8(out + idx * 8, v);
  }
}

static inline int splits_for(int N, int P, int cv) {
  long long work = (long long)P * cv;               // 16-byte vectors per image
  int s = (592 + N - 1) / N;                         // ~4 blocks per SM overall
  int smax = (int)((work + 1023) / 1024);            // >= 4 vectors per thread
  if (s > smax) s = smax;
  if (s < 1) s = 1;
  return s;
}

}  // namespace pcm

using namespace pcm;

extern "C" int pcm_gn_stats(const void* x, float* stats, int N, int P, int C, int G, int dtype, pcm_stream_t s) {
  PCM_REQUIRE(C % 8 == 0 && C % G == 0 && G <= 64 && C / 8 <= 256, "gn_stats: unsupported C=%d G=%d", C, G);
  if (N == 0) return PCM_OK;
  dim3 grid(splits_for(N, P, C / 8), N);
  PCM_DISPATCH_DTYPE(dtype, T, (gn_stats_kernel<T><<<grid, 256, 0, (cudaStream_t)s>>>((const T*)x, stats, P, C, G)));
  return check_launch("gn_stats");
}

extern "C" int pcm_gn_silu_fwd(const void* x, const float* stats, const float* gamma, const float* beta, void* y,
                               float* pool, int N, int P, int C, int G, float eps, int dtype, pcm_stream_t s) {
  PCM_REQUIRE(C % 8 == 0 && C % G == 0 && G <= 64 && C / 8 <= 256, "gn_silu_fwd: unsupported C=%d G=%d", C, G);
  if (N == 0) return PCM_OK;
  dim3 grid(splits_for(N, P, C / 8), N);
  PCM_DISPATCH_DTYPE(dtype, T, (gn_silu_fwd_kernel<T><<<grid, 256, C * sizeof(float), (cudaStream_t)s>>>(
                                   (const T*)x, stats, gamma, beta, (T*)y, pool, P, C, G, eps)));
  return check_launch("gn_silu_fwd");
}

extern "C" int pcm_se_chanstat_fwd(const void* a, const float* pool, const float* w1, const float* w2, float* se,
                                   float* hid, float* cmap, int N, int P, int C, int Cr, int dtype, pcm_stream_t s) {
  PCM_REQUIRE(C % 8 == 0 && Cr >= 1, "se_chanstat_fwd: unsupported C=%d Cr=%d", C, Cr);
  if (N == 0) return PCM_OK;
  int splits = (592 + N - 1) / N;
  const int smax = (P + 255) / 256;
  if (splits > smax) splits = smax;
  dim3 grid(splits, N);
  const size_t smem = (2 * C + Cr) * sizeof(float);
  PCM_DISPATCH_DTYPE(dtype, T, (se_chanstat_fwd_kernel<T><<<grid, 256, smem, (cudaStream_t)s>>>(
                                   (const T*)a, pool, w1, w2, se, hid, cmap, P, C, Cr)));
  return check_launch("se_chanstat_fwd");
}

extern "C" int pcm_scale_channels(const void* x, const float* scale, const float* add, void* out, int N, int P, int C,
                                  int dtype, pcm_stream_t s) {
  PCM_REQUIRE(C % 8 == 0, "scale_channels: C must be a multiple of 8");
  if (N == 0) return PCM_OK;
  const long long total = (long long)N * P * (C / 8);
  long long blocks = (total + 255) / 256;
  if (blocks > 1184) blocks = 1184;
  PCM_DISPATCH_DTYPE(dtype, T, (scale_channels_kernel<T><<<(int)blocks, 256, 0, (cudaStream_t)s>>>(
                                   (const T*)x, scale, add, (T*)out, N, P, C)));
  return check_launch("scale_channels");
}

extern "C" int pcm_spatial_gate_fwd(const void* a, const float* se, const float* cmap, const float* wsp, float* gate,
                                    void* out, int N, int H, int W, int C, int dtype, pcm_stream_t s) {
  PCM_REQUIRE(C % 8 == 0, "spatial_gate_fwd: C must be a multiple of 8");
  if (N == 0) return PCM_OK;
  dim3 grid(ceil_div(H * W, 256), N);
  const size_t smem = (98 + C) * sizeof(float);
  PCM_DISPATCH_DTYPE(dtype, T, (spatial_gate_fwd_kernel<T><<<grid, 256, smem, (cudaStream_t)s>>>(
                                   (const T*)a, se, cmap, wsp, gate, (T*)out, H, W, C)));
  return check_launch("spatial_gate_fwd");
}

extern "C" int pcm_spatial_gate_bwd_dq(const void* dout, const void* a, const float* se, const float* gate, float* dq,
                                       int N, int P, int C, int dtype, pcm_stream_t s) {
  PCM_REQUIRE(C % 8 == 0, "spatial_gate_bwd_dq: C must be a multiple of 8");
  if (N == 0) return PCM_OK;
  dim3 grid(ceil_div(P, 256), N);
  PCM_DISPATCH_DTYPE(dtype, T, (spatial_gate_bwd_dq_kernel<T><<<grid, 256, C * sizeof(float), (cudaStream_t)s>>>(
                                   (const T*)dout, (const T*)a, se, gate, dq, P, C)));
  return check_launch("spatial_gate_bwd_dq");
}

extern "C" int pcm_spatial_gate_bwd_dw(const float* dq, const float* cmap, float* dwsp, int N, int H, int W,
                                       pcm_stream_t s) {
  if (N == 0) return PCM_OK;
  const long long total = (long long)N * H * W;
  const int blocks = (int)min((long long)296, (total + 255) / 256);
  spatial_gate_bwd_dw_kernel<<<blocks, 256, 0, (cudaStream_t)s>>>(dq, cmap, dwsp, N, H, W);
  return check_launch("spatial_gate_bwd_dw");
}

extern "C" int pcm_spatial_gate_bwd_da(const void* dout, const void* a, const float* se, const float* gate,
                                       const float* cmap, const float* dq, const float* wsp, void* da, float* dse,
                                       int N, int H, int W, int C, int dtype, pcm_stream_t s) {
  PCM_REQUIRE(C % 8 == 0, "spatial_gate_bwd_da: C must be a multiple of 8");
  if (N == 0) return PCM_OK;
  dim3 grid(ceil_div(H * W, 256), N);
  const size_t smem = (98 + 2 * C) * sizeof(float);
  PCM_DISPATCH_DTYPE(dtype, T, (spatial_gate_bwd_da_kernel<T><<<grid, 256, smem, (cudaStream_t)s>>>(
                                   (const T*)dout, (const T*)a, se, gate, cmap, dq, wsp, (T*)da, dse, H, W, C)));
  return check_launch("spatial_gate_bwd_da");
}

extern "C" int pcm_se_bwd(const float* dse, const float* se, const float* hid, const float* pool, const float* w1,
                          const float* w2, float* dpool, float* dw1, float* dw2, int N, int P, int C, int Cr,
                          pcm_stream_t s) {
  if (N == 0) return PCM_OK;
  const size_t smem = (2 * C + 2 * Cr) * sizeof(float);
  se_bwd_kernel<<<N, 128, smem, (cudaStream_t)s>>>(dse, se, hid, pool, w1, w2, dpool, dw1, dw2, P, C, Cr);
  return check_launch("se_bwd");
}

extern "C" int pcm_gn_silu_bwd_reduce(const void* da, const float* dpool, const void* x, const float* stats,
                                      const float* gamma, const float* beta, float* gsum, float* dgamma,
                                      float* dbeta, int N, int P, int C, int G, float eps, int dtype,
                                      pcm_stream_t s) {
  PCM_REQUIRE(C % 8 == 0 && C % G == 0 && G <= 64 && C / 8 <= 256, "gn_silu_bwd: unsupported C=%d G=%d", C, G);
  if (N == 0) return PCM_OK;
  dim3 grid(splits_for(N, P, C / 8), N);
  const size_t smem = (G * 2 + 2 * C) * sizeof(float);
  PCM_DISPATCH_DTYPE(dtype, T, (gn_silu_bwd_kernel<T, false><<<grid, 256, smem, (cudaStream_t)s>>>(
                                   (const T*)da, dpool, (const T*)x, stats, gamma, beta, gsum, dgamma, dbeta,
                                   (T*)nullptr, P, C, G, eps)));
  return check_launch("gn_silu_bwd_reduce");
}

extern "C" int pcm_gn_silu_bwd_apply(const void* da, const float* dpool, const void* x, const float* stats,
                                     const float* gamma, const float* beta, const float* gsum, void* dx, int N, int P,
                                     int C, int G, float eps, int dtype, pcm_stream_t s) {
  PCM_REQUIRE(C % 8 == 0 && C % G == 0 && G <= 64 && C / 8 <= 256, "gn_silu_bwd: unsupported C=%d G=%d", C, G);
  if (N == 0) return PCM_OK;
  dim3 grid(splits_for(N, P, C / 8), N);
  PCM_DISPATCH_DTYPE(dtype, T, (gn_silu_bwd_kernel<T, true><<<grid, 256, 0, (cudaStream_t)s>>>(
                                   (const T*)da, dpool, (const T*)x, stats, gamma, beta, (float*)gsum, nullptr,
                                   nullptr, (T*)dx, P, C, G, eps)));
  return check_launch("gn_silu_bwd_apply");
}
