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

// Lanes of a warp that hold the same channel block (stride cv, cv a power of two < 32) are summed with
// xor-shuffles; afterwards lanes [0, cv) (or every lane when cv >= 32) carry distinct channels, so the
// shared-memory atomics that follow are conflict-free inside a warp.
template <int NV>
__device__ __forceinline__ bool warp_reduce_same_cb(float (&v)[NV], int cv) {
  if (cv < 32) {
    for (int off = cv; off < 32; off <<= 1) {
#pragma unroll
      for (int j = 0; j < NV; ++j) v[j] += __shfl_xor_sync(0xffffffffu, v[j], off);
    }
    return (threadIdx.x & 31) < cv;
  }
  return true;
}
__device__ __forceinline__ bool pow2(int x) { return (x & (x - 1)) == 0; }

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
  PCM_PDL_ENTRY();
  __shared__ float sg[64 * 2];
  const int n = blockIdx.y, cv = C / 8, cg = C / G;
  for (int i = threadIdx.x; i < G * 2; i += blockDim.x) sg[i] = 0.f;
  __syncthreads();
  const PixSplit ps = pix_split(P, cv);
  float acc[16];     // [0..8) sums, [8..16) sums of squares
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = 0.f;
  if (ps.active) {
    const T* xn = x + (long long)n * P * C + ps.cb * 8;
    int p = ps.p0 + ps.lane;
    for (; p + ps.lanes < ps.p1; p += 2 * ps.lanes) {      // two independent loads in flight
      float v[8], u[8];
      load8(xn + (long long)p * C, v);
      load8(xn + (long long)(p + ps.lanes) * C, u);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[j] += v[j] + u[j];
        acc[8 + j] = fmaf(v[j], v[j], fmaf(u[j], u[j], acc[8 + j]));
      }
    }
    if (p < ps.p1) {
      float v[8];
      load8(xn + (long long)p * C, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) { acc[j] += v[j]; acc[8 + j] = fmaf(v[j], v[j], acc[8 + j]); }
    }
  }
  const bool writer = (pow2(cv) ? warp_reduce_same_cb(acc, cv) : true) && ps.active;
  if (writer) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = (ps.cb * 8 + j) / cg;
      atomicAdd(&sg[g * 2], acc[j]);
      atomicAdd(&sg[g * 2 + 1], acc[8 + j]);
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
  PCM_PDL_ENTRY();
  extern __shared__ float spool[];   // [C]
  const int n = blockIdx.y, cv = C / 8, cg = C / G;
  if (pool) {
    for (int i = threadIdx.x; i < C; i += blockDim.x) spool[i] = 0.f;
    __syncthreads();
  }
  const PixSplit ps = pix_split(P, cv);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (ps.active) {
    float mu[8], rs[8], ga[8], be[8];
    const float cnt = (float)cg * (float)P;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = ps.cb * 8 + j;
      group_mean_rstd(stats, n, G, c, cg, cnt, eps, mu[j], rs[j]);
      ga[j] = __ldg(gamma + c); be[j] = __ldg(beta + c);
      // fold the normalisation into one FMA: z = a*x + b
      ga[j] *= rs[j];
      be[j] = fmaf(-mu[j], ga[j], be[j]);
    }
    const long long base = (long long)n * P * C + ps.cb * 8;
    int p = ps.p0 + ps.lane;
    for (; p + ps.lanes < ps.p1; p += 2 * ps.lanes) {
      float v[8], u[8];
      load8(x + base + (long long)p * C, v);
      load8(x + base + (long long)(p + ps.lanes) * C, u);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float z = fmaf(ga[j], v[j], be[j]), z2 = fmaf(ga[j], u[j], be[j]);
        v[j] = round_to<T>(z * sigmoidf_(z));
        u[j] = round_to<T>(z2 * sigmoidf_(z2));
        acc[j] += v[j] + u[j];
      }
      store8(y + base + (long long)p * C, v);
      store8(y + base + (long long)(p + ps.lanes) * C, u);
    }
    if (p < ps.p1) {
      float v[8];
      load8(x + base + (long long)p * C, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float z = fmaf(ga[j], v[j], be[j]);
        v[j] = round_to<T>(z * sigmoidf_(z));
        acc[j] += v[j];
      }
      store8(y + base + (long long)p * C, v);
    }
    if (pool && !pow2(cv)) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&spool[ps.cb * 8 + j], acc[j]);
    }
  }
  if (pool && pow2(cv)) {          // (all threads are active when cv is a power of two <= 256)
    if (warp_reduce_same_cb(acc, cv)) {
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
  PCM_PDL_ENTRY();
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

// 7x7 stencil over the 2-channel (mean, max) map, tiled through shared memory: a block owns a band of TH
// image rows (+3 halo rows/cols, zero padded), computes the gate per pixel, then applies out = a*se*gate
// with the coalesced (pixel, 8-channel vector) mapping.
constexpr int kGateTH = 8;

template <typename T>
__global__ void __launch_bounds__(256)
spatial_gate_fwd_kernel(const T* __restrict__ a, const float* __restrict__ se, const float* __restrict__ cmap,
                        const float* __restrict__ wsp, float* __restrict__ gate, T* __restrict__ out, int H, int W,
                        int C) {
  PCM_PDL_ENTRY();
  extern __shared__ float sm[];   // wsp[98] | pad | sse[C] | sgate[TH*W] | tile float2 [(TH+6)*(W+6)]
  float* sw = sm;
  float* sse = sm + 100;
  float* sgate = sse + C;
  float2* tile = reinterpret_cast<float2*>(sgate + kGateTH * W + ((kGateTH * W + C) & 1));
  const int n = blockIdx.y, P = H * W, Wt = W + 6;
  const int h0 = blockIdx.x * kGateTH;
  const int th = min(kGateTH, H - h0);
  for (int i = threadIdx.x; i < 98; i += blockDim.x) sw[i] = __ldg(wsp + i);
  for (int i = threadIdx.x; i < C; i += blockDim.x) sse[i] = __ldg(se + (long long)n * C + i);
  const float2* cm = reinterpret_cast<const float2*>(cmap) + (long long)n * P;
  for (int i = threadIdx.x; i < (th + 6) * Wt; i += blockDim.x) {
    const int hh = h0 - 3 + i / Wt, ww = i % Wt - 3;
    tile[i] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? __ldg(cm + hh * W + ww) : make_float2(0.f, 0.f);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < th * W; i += blockDim.x) {
    const int hl = i / W, wl = i % W;
    float q = 0.f;
#pragma unroll
    for (int dy = 0; dy < 7; ++dy) {
#pragma unroll
      for (int dx = 0; dx < 7; ++dx) {
        const float2 m = tile[(hl + dy) * Wt + wl + dx];
        q = fmaf(sw[dy * 7 + dx], m.x, q);
        q = fmaf(sw[49 + dy * 7 + dx], m.y, q);
      }
    }
    const float gt = sigmoidf_(q);
    sgate[i] = gt;
    gate[(long long)n * P + h0 * W + i] = gt;
  }
  __syncthreads();
  const int cv = C / 8;
  const long long base = ((long long)n * P + (long long)h0 * W) * C;
  for (int v = threadIdx.x; v < th * W * cv; v += blockDim.x) {
    const int pix = v / cv, cb = v % cv;
    float x[8];
    load8(a + base + (long long)v * 8, x);
    const float gt = sgate[pix];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = x[j] * sse[cb * 8 + j] * gt;
    store8(out + base + (long long)v * 8, x);
  }
}

// ---- backward ---------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
spatial_gate_bwd_dq_kernel(const T* __restrict__ dout, const T* __restrict__ a, const float* __restrict__ se,
                           const float* __restrict__ gate, float* __restrict__ dq, int P, int C) {
  PCM_PDL_ENTRY();
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

// dwsp[k][dy][dx] += sum_p dq[p] * cmap_k[p + (dy-3, dx-3)].  One thread per tap (2 groups of 98 threads
// split the pixels), operands staged per (image, row band) in shared memory; blocks loop over items and
// keep their 98 sums in registers, so there is no per-pixel reduction at all.
__global__ void __launch_bounds__(256)
spatial_gate_bwd_dw_kernel(const float* __restrict__ dq, const float* __restrict__ cmap, float* __restrict__ dwsp,
                           int N, int H, int W) {
  PCM_PDL_ENTRY();
  extern __shared__ float sm[];   // sdq[TH*W] | tile float2 [(TH+6)*(W+6)]
  float* sdq = sm;
  float2* tile = reinterpret_cast<float2*>(sdq + kGateTH * W + ((kGateTH * W) & 1));
  const int P = H * W, Wt = W + 6;
  const int bands = (H + kGateTH - 1) / kGateTH;
  const int tap = threadIdx.x % 98, grp = threadIdx.x / 98;       // grp 2 (threads 196..255) only helps loading
  const int k = tap / 49, dy = (tap % 49) / 7, dx = tap % 7;
  float acc = 0.f;
  for (int item = blockIdx.x; item < N * bands; item += gridDim.x) {
    const int n = item / bands, h0 = (item % bands) * kGateTH;
    const int th = min(kGateTH, H - h0);
    const float2* cm = reinterpret_cast<const float2*>(cmap) + (long long)n * P;
    __syncthreads();
    for (int i = threadIdx.x; i < (th + 6) * Wt; i += blockDim.x) {
      const int hh = h0 - 3 + i / Wt, ww = i % Wt - 3;
      tile[i] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? __ldg(cm + hh * W + ww) : make_float2(0.f, 0.f);
    }
    for (int i = threadIdx.x; i < th * W; i += blockDim.x) sdq[i] = __ldg(dq + (long long)n * P + h0 * W + i);
    __syncthreads();
    if (grp < 2) {
      // rows of the band alternate between the two groups; 4 independent accumulators hide the LDS latency
      const float* tk = reinterpret_cast<const float*>(tile) + k + 2 * dx;
      for (int hl = grp; hl < th; hl += 2) {
        const float* trow = tk + 2 * (hl + dy) * Wt;
        const float* drow = sdq + hl * W;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        int wl = 0;
        for (; wl + 3 < W; wl += 4) {
          a0 = fmaf(drow[wl], trow[2 * wl], a0);
          a1 = fmaf(drow[wl + 1], trow[2 * wl + 2], a1);
          a2 = fmaf(drow[wl + 2], trow[2 * wl + 4], a2);
          a3 = fmaf(drow[wl + 3], trow[2 * wl + 6], a3);
        }
        for (; wl < W; ++wl) a0 = fmaf(drow[wl], trow[2 * wl], a0);
        acc += (a0 + a1) + (a2 + a3);
      }
    }
  }
  if (grp < 2) atomicAdd(dwsp + tap, acc);
}

template <typename T>
__global__ void __launch_bounds__(256)
spatial_gate_bwd_da_kernel(const T* __restrict__ dout, const T* __restrict__ a, const float* __restrict__ se,
                           const float* __restrict__ gate, const float* __restrict__ cmap,
                           const float* __restrict__ dq, const float* __restrict__ wsp, T* __restrict__ da,
                           float* __restrict__ dse, int H, int W, int C) {
  PCM_PDL_ENTRY();
  extern __shared__ float sm[];
  // wsp[98] pad | sse[C] | sdse[C] | s_gate[THW] | s_mx[THW] | s_dmean[THW] | s_dmax[THW] | s_cnt[THW] | dq tile[(TH+6)*(W+6)]
  const int THW = kGateTH * W, Wt = W + 6;
  float* sw = sm;
  float* sse = sm + 100;
  float* sdse = sse + C;
  float* s_gate = sdse + C;
  float* s_mx = s_gate + THW;
  float* s_dmean = s_mx + THW;
  float* s_dmax = s_dmean + THW;
  int* s_cnt = reinterpret_cast<int*>(s_dmax + THW);
  float* tile = reinterpret_cast<float*>(s_cnt + THW);
  const int n = blockIdx.y, P = H * W, cv = C / 8;
  const int h0 = blockIdx.x * kGateTH;
  const int th = min(kGateTH, H - h0);
  for (int i = threadIdx.x; i < 98; i += blockDim.x) sw[i] = __ldg(wsp + i);
  for (int i = threadIdx.x; i < C; i += blockDim.x) { sse[i] = __ldg(se + (long long)n * C + i); sdse[i] = 0.f; }
  const float* dqn = dq + (long long)n * P;
  for (int i = threadIdx.x; i < (th + 6) * Wt; i += blockDim.x) {
    const int hh = h0 - 3 + i / Wt, ww = i % Wt - 3;
    tile[i] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? __ldg(dqn + hh * W + ww) : 0.f;
  }
  __syncthreads();
  // per pixel: gradient reaching (mean, max) through the transposed stencil
  // (cmap[p] feeds q[p'] with p' = p - (dy-3, dx-3), weight w[dy][dx])
  for (int i = threadIdx.x; i < th * W; i += blockDim.x) {
    const int hl = i / W, wl = i % W;
    float dm0 = 0.f, dm1 = 0.f;
#pragma unroll
    for (int dy = 0; dy < 7; ++dy) {
#pragma unroll
      for (int dx = 0; dx < 7; ++dx) {
        const float d = tile[(hl + 6 - dy) * Wt + wl + 6 - dx];
        dm0 = fmaf(sw[dy * 7 + dx], d, dm0);
        dm1 = fmaf(sw[49 + dy * 7 + dx], d, dm1);
      }
    }
    const long long pg = (long long)n * P + h0 * W + i;
    s_gate[i] = __ldg(gate + pg);
    s_mx[i] = __ldg(cmap + pg * 2 + 1);
    s_dmean[i] = dm0 / (float)C;
    s_dmax[i] = dm1;
    s_cnt[i] = 0;
  }
  __syncthreads();
  const long long base = ((long long)n * P + (long long)h0 * W) * C;
  const int nvec = th * W * cv;
  // torch.amax splits the gradient evenly among ties: count them (almost always 1)
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    const int pix = v / cv, cb = v % cv;
    float x[8];
    load8(a + base + (long long)v * 8, x);
    const float mx = s_mx[pix];
    int c = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) c += (x[j] * sse[cb * 8 + j] == mx) ? 1 : 0;
    if (c) atomicAdd(&s_cnt[pix], c);
  }
  __syncthreads();
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  const int cb = threadIdx.x % cv;               // blockDim % cv == 0 -> fixed channel block per thread
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    const int pix = v / cv;
    float d[8], x[8], r[8];
    load8(dout + base + (long long)v * 8, d);
    load8(a + base + (long long)v * 8, x);
    const float gt = s_gate[pix], mx = s_mx[pix], dmean = s_dmean[pix];
    const float dmax = s_dmax[pix] / (float)max(s_cnt[pix], 1);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float sc = sse[cb * 8 + j];
      const float u = x[j] * sc;
      const float du = d[j] * gt + dmean + ((u == mx) ? dmax : 0.f);
      r[j] = du * sc;
      acc[j] = fmaf(du, x[j], acc[j]);
    }
    store8(da + base + (long long)v * 8, r);
  }
  if (warp_reduce_same_cb(acc, cv)) {
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&sdse[cb * 8 + j], acc[j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(dse + (long long)n * C + i, sdse[i]);
}

__global__ void __launch_bounds__(128)
se_bwd_kernel(const float* __restrict__ dse, const float* __restrict__ se, const float* __restrict__ hid,
              const float* __restrict__ pool, const float* __restrict__ w1, const float* __restrict__ w2,
              float* __restrict__ dpool, float* __restrict__ dw1, float* __restrict__ dw2, int P, int C, int Cr) {
  PCM_PDL_ENTRY();
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
  PCM_PDL_ENTRY();
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
  float red[32];     // reduce pass: [0..8) dgamma, [8..16) dbeta, [16..24) sum dxhat, [24..32) sum dxhat*xhat
#pragma unroll
  for (int j = 0; j < 32; ++j) red[j] = 0.f;
  if (ps.active) {
    float mu[8], rs[8], ga[8], be[8], dp[8], m1[8], m2[8];
    const float cnt = (float)cg * (float)P;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = ps.cb * 8 + j;
      group_mean_rstd(stats, n, G, c, cg, cnt, eps, mu[j], rs[j]);
      ga[j] = __ldg(gamma + c); be[j] = __ldg(beta + c);
      dp[j] = dpool ? __ldg(dpool + (long long)n * C + c) : 0.f;
      m1[j] = m2[j] = 0.f;
      if (APPLY) {
        const int g = c / cg;
        m1[j] = __ldg(gsum + ((long long)n * G + g) * 2) / cnt;
        m2[j] = __ldg(gsum + ((long long)n * G + g) * 2 + 1) / cnt;
      }
    }
    const long long base = (long long)n * P * C + ps.cb * 8;
    for (int p = ps.p0 + ps.lane; p < ps.p1; p += 2 * ps.lanes) {
      const bool two = p + ps.lanes < ps.p1;
      float v[2][8], d[2][8];
      load8(x + base + (long long)p * C, v[0]);
      load8(da + base + (long long)p * C, d[0]);
      if (two) {
        load8(x + base + (long long)(p + ps.lanes) * C, v[1]);
        load8(da + base + (long long)(p + ps.lanes) * C, d[1]);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (u == 1 && !two) break;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xh = (v[u][j] - mu[j]) * rs[j];
          const float z = fmaf(ga[j], xh, be[j]);
          const float sg = sigmoidf_(z);
          const float dz = (d[u][j] + dp[j]) * sg * (1.f + z * (1.f - sg));
          const float dxh = dz * ga[j];
          if (APPLY) {
            v[u][j] = rs[j] * (dxh - m1[j] - xh * m2[j]);
          } else {
            red[j] = fmaf(dz, xh, red[j]);
            red[8 + j] += dz;
            red[16 + j] += dxh;
            red[24 + j] = fmaf(dxh, xh, red[24 + j]);
          }
        }
        if (APPLY) store8(dx + base + (long long)(p + u * ps.lanes) * C, v[u]);
      }
    }
  }
  if (!APPLY) {
    const bool writer = (pow2(cv) ? warp_reduce_same_cb(red, cv) : true) && ps.active;
    if (writer) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = ps.cb * 8 + j, g = c / cg;
        atomicAdd(&sdg[c], red[j]);
        atomicAdd(&sdb[c], red[8 + j]);
        atomicAdd(&sgs[g * 2], red[16 + j]);
        atomicAdd(&sgs[g * 2 + 1], red[24 + j]);
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
  PCM_PDL_ENTRY();
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

static inline bool pow2h(int x) { return x > 0 && (x & (x - 1)) == 0; }

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
  PCM_DISPATCH_DTYPE(dtype, T, (pcm::launch(gn_stats_kernel<T>, grid, 256, 0, (cudaStream_t)s, (const T*)x, stats, P, C, G)));
  return check_launch("gn_stats");
}

extern "C" int pcm_gn_silu_fwd(const void* x, const float* stats, const float* gamma, const float* beta, void* y,
                               float* pool, int N, int P, int C, int G, float eps, int dtype, pcm_stream_t s) {
  PCM_REQUIRE(C % 8 == 0 && C % G == 0 && G <= 64 && C / 8 <= 256, "gn_silu_fwd: unsupported C=%d G=%d", C, G);
  if (N == 0) return PCM_OK;
  dim3 grid(splits_for(N, P, C / 8), N);
  PCM_DISPATCH_DTYPE(dtype, T, (pcm::launch(gn_silu_fwd_kernel<T>, grid, 256, C * sizeof(float), (cudaStream_t)s, 
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
  PCM_DISPATCH_DTYPE(dtype, T, (pcm::launch(se_chanstat_fwd_kernel<T>, grid, 256, smem, (cudaStream_t)s, 
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
  PCM_DISPATCH_DTYPE(dtype, T, (pcm::launch(scale_channels_kernel<T>, (int)blocks, 256, 0, (cudaStream_t)s, 
                                   (const T*)x, scale, add, (T*)out, N, P, C)));
  return check_launch("scale_channels");
}

extern "C" int pcm_spatial_gate_fwd(const void* a, const float* se, const float* cmap, const float* wsp, float* gate,
                                    void* out, int N, int H, int W, int C, int dtype, pcm_stream_t s) {
  PCM_REQUIRE(C % 8 == 0, "spatial_gate_fwd: C must be a multiple of 8");
  if (N == 0) return PCM_OK;
  PCM_REQUIRE(C / 8 <= 256 && pow2h(C / 8), "spatial_gate_fwd: C/8 must be a power of two <= 256");
  dim3 grid(ceil_div(H, kGateTH), N);
  const size_t smem = (100 + C + kGateTH * W + 2 + 2 * (kGateTH + 6) * (W + 6)) * sizeof(float);
  if (smem > 48 * 1024)
    PCM_DISPATCH_DTYPE(dtype, T, cudaFuncSetAttribute(spatial_gate_fwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  PCM_DISPATCH_DTYPE(dtype, T, (pcm::launch(spatial_gate_fwd_kernel<T>, grid, 256, smem, (cudaStream_t)s, 
                                   (const T*)a, se, cmap, wsp, gate, (T*)out, H, W, C)));
  return check_launch("spatial_gate_fwd");
}

extern "C" int pcm_spatial_gate_bwd_dq(const void* dout, const void* a, const float* se, const float* gate, float* dq,
                                       int N, int P, int C, int dtype, pcm_stream_t s) {
  PCM_REQUIRE(C % 8 == 0, "spatial_gate_bwd_dq: C must be a multiple of 8");
  if (N == 0) return PCM_OK;
  dim3 grid(ceil_div(P, 256), N);
  PCM_DISPATCH_DTYPE(dtype, T, (pcm::launch(spatial_gate_bwd_dq_kernel<T>, grid, 256, C * sizeof(float), (cudaStream_t)s, 
                                   (const T*)dout, (const T*)a, se, gate, dq, P, C)));
  return check_launch("spatial_gate_bwd_dq");
}

extern "C" int pcm_spatial_gate_bwd_dw(const float* dq, const float* cmap, float* dwsp, int N, int H, int W,
                                       pcm_stream_t s) {
  if (N == 0) return PCM_OK;
  const int items = N * ceil_div(H, kGateTH);
  const int blocks = items < 148 * 6 ? items : 148 * 6;
  const size_t smem = (kGateTH * W + 2 + 2 * (kGateTH + 6) * (W + 6)) * sizeof(float);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(spatial_gate_bwd_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  pcm::launch(spatial_gate_bwd_dw_kernel, blocks, 256, smem, (cudaStream_t)s, dq, cmap, dwsp, N, H, W);
  return check_launch("spatial_gate_bwd_dw");
}

extern "C" int pcm_spatial_gate_bwd_da(const void* dout, const void* a, const float* se, const float* gate,
                                       const float* cmap, const float* dq, const float* wsp, void* da, float* dse,
                                       int N, int H, int W, int C, int dtype, pcm_stream_t s) {
  PCM_REQUIRE(C % 8 == 0, "spatial_gate_bwd_da: C must be a multiple of 8");
  if (N == 0) return PCM_OK;
  PCM_REQUIRE(C / 8 <= 256 && pow2h(C / 8), "spatial_gate_bwd_da: C/8 must be a power of two <= 256");
  dim3 grid(ceil_div(H, kGateTH), N);
  const size_t smem = (100 + 2 * C + 5 * kGateTH * W + (kGateTH + 6) * (W + 6)) * sizeof(float);
  if (smem > 48 * 1024)
    PCM_DISPATCH_DTYPE(dtype, T, cudaFuncSetAttribute(spatial_gate_bwd_da_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  PCM_DISPATCH_DTYPE(dtype, T, (pcm::launch(spatial_gate_bwd_da_kernel<T>, grid, 256, smem, (cudaStream_t)s, 
                                   (const T*)dout, (const T*)a, se, gate, cmap, dq, wsp, (T*)da, dse, H, W, C)));
  return check_launch("spatial_gate_bwd_da");
}

extern "C" int pcm_se_bwd(const float* dse, const float* se, const float* hid, const float* pool, const float* w1,
                          const float* w2, float* dpool, float* dw1, float* dw2, int N, int P, int C, int Cr,
                          pcm_stream_t s) {
  if (N == 0) return PCM_OK;
  const size_t smem = (2 * C + 2 * Cr) * sizeof(float);
  pcm::launch(se_bwd_kernel, N, 128, smem, (cudaStream_t)s, dse, se, hid, pool, w1, w2, dpool, dw1, dw2, P, C, Cr);
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
  PCM_DISPATCH_DTYPE(dtype, T, (pcm::launch(gn_silu_bwd_kernel<T, false>, grid, 256, smem, (cudaStream_t)s, 
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
  PCM_DISPATCH_DTYPE(dtype, T, (pcm::launch(gn_silu_bwd_kernel<T, true>, grid, 256, 0, (cudaStream_t)s, 
                                   (const T*)da, dpool, (const T*)x, stats, gamma, beta, (float*)gsum, nullptr,
                                   nullptr, (T*)dx, P, C, G, eps)));
  return check_launch("gn_silu_bwd_apply");
}
