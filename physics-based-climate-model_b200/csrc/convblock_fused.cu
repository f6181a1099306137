// Per-image fused ConvBlock tails (reference src/unet.py:35-49 with SEBlock :6-17 and SpatialGate :19-29).
//
// At the emulator's grid sizes one image of one level fits in a single SM's shared memory (48x72x16 bf16 =
// 108 KB, 24x36x32 = 54 KB, ...), so everything that follows a 3x3 convolution of the block runs as ONE kernel
// with one CTA per image and the image resident in shared memory:
//
//   tail 1  (after conv1):  GroupNorm(8) statistics -> normalise -> SiLU                      1 read, 1 write
//   tail 2  (after conv2):  GroupNorm(8) -> SiLU -> SE squeeze/excite -> channel mean/max map
//                           -> 7x7 gate conv -> sigmoid -> out = a*se*gate                    1 read, 1 write
//   and the two backward tails (dout -> dy2, da1 -> dy1), which recompute a2 / gate from the saved
//   pre-normalisation tensor instead of reading saved activations.
//
// The multi-kernel path in convblock.cu (grid-wide passes, 6 forward + 8 backward launches per block) remains for
// images that do not fit (config 5, fp32 at full size) — see pcm_convblock_fused_supported.
#include "common.cuh"

namespace pcm {

constexpr int kFT = 512;       // threads per CTA (one CTA per image)
constexpr int kGroups = 8;     // nn.GroupNorm(8, c)

template <typename T> __device__ __forceinline__ float sigmoid_t(float z);
template <> __device__ __forceinline__ float sigmoid_t<float>(float z) { return sigmoidf_(z); }
template <> __device__ __forceinline__ float sigmoid_t<__nv_bfloat16>(float z) {
  // one SFU op (tanh.approx, rel. error ~2^-11 < bf16 resolution) instead of ex2 + rcp
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * z));
  return fmaf(0.5f, t, 0.5f);
}

// add this thread's 8 per-channel partials (channel block cb = tid % cv) into dst[cb*8 + j]; lanes of a warp that
// hold the same cb are combined with xor-shuffles first.  Must be called by every thread of the CTA.
__device__ __forceinline__ void chan_add(float (&v)[8], float* dst, int cb, int cv) {
  if (cv < 32) {
    for (int off = cv; off < 32; off <<= 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += __shfl_xor_sync(0xffffffffu, v[j], off);
    }
    if ((int)(threadIdx.x & 31) < cv) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&dst[cb * 8 + j], v[j]);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&dst[cb * 8 + j], v[j]);
  }
}

// plain (generic-address) 8-element load: shared-memory reads and re-reads of this thread's own global writes
// (load8 in common.cuh goes through the read-only ld.global.nc path, which is neither)
template <typename T>
__device__ __forceinline__ void load8_rw(const T* p, float d[8]) {
  if (sizeof(T) == 2) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); d[2 * i] = f.x; d[2 * i + 1] = f.y; }
  } else {
    const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
    d[0] = a.x; d[1] = a.y; d[2] = a.z; d[3] = a.w; d[4] = b.x; d[5] = b.y; d[6] = b.z; d[7] = b.w;
  }
}

struct TailSmem {
  size_t img, cmap, dq, gate, dm, cnt, fl, total;
};
// bwd: 0 = forward tails, 1 = backward tails (needs dq / dm / cnt as well)
__host__ __device__ inline TailSmem tail_smem_layout(int H, int W, int C, int elt, int full, int bwd) {
  TailSmem L;
  const size_t P = (size_t)H * W, Pp = (size_t)(H + 6) * (W + 6);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 15) & ~(size_t)15; return o; };
  L.img = take(P * C * elt);
  L.cmap = full ? take(Pp * 8) : 0;
  L.dq = (full && bwd) ? take(Pp * 4) : 0;
  L.gate = full ? take(P * 4) : 0;
  L.dm = (full && bwd) ? take(P * 8) : 0;
  L.cnt = (full && bwd) ? take(P) : 0;
  // floats: 5 channel arrays | a[C] b[C] | se[C] pool[C] dpool[C] dpre2[C] | hid[64] dpre1[64] | mu[8] rs[8] m1[8] m2[8] | w[100] dw[100]
  L.fl = take((size_t)(11 * C + 128 + 32 + 200) * 4);
  L.total = off;
  return L;
}

struct TailPtrs {
  float *ch0, *ch1, *ch2, *ch3, *ch4, *ca, *cb_, *se, *pool, *dpool, *dpre2, *hid, *dpre1, *mu, *rs, *m1, *m2, *w, *dw;
};
__device__ __forceinline__ TailPtrs tail_ptrs(uint8_t* smem, const TailSmem& L, int C) {
  float* f = reinterpret_cast<float*>(smem + L.fl);
  TailPtrs p;
  p.ch0 = f; p.ch1 = f + C; p.ch2 = f + 2 * C; p.ch3 = f + 3 * C; p.ch4 = f + 4 * C;
  p.ca = f + 5 * C; p.cb_ = f + 6 * C;
  p.se = f + 7 * C; p.pool = f + 8 * C; p.dpool = f + 9 * C; p.dpre2 = f + 10 * C;
  float* g = f + 11 * C;
  p.hid = g; p.dpre1 = g + 64;
  p.mu = g + 128; p.rs = g + 136; p.m1 = g + 144; p.m2 = g + 152;
  p.w = g + 160; p.dw = g + 260;
  return p;
}

// GroupNorm statistics of the image in shared memory -> mu/rs per group (+ raw sums to `stats_out` when non-null)
template <typename T>
__device__ __forceinline__ void image_group_stats(const T* s_img, int nvec, int cv, int cg, int P, float eps,
                                                  const TailPtrs& sp, float* stats_out) {
  const int cb = threadIdx.x % cv;
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  for (int v = threadIdx.x; v < nvec; v += kFT) {
    float x[8];
    load8_rw(s_img + (size_t)v * 8, x);
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] += x[j]; q[j] = fmaf(x[j], x[j], q[j]); }
  }
  chan_add(s, sp.ch0, cb, cv);
  chan_add(q, sp.ch1, cb, cv);
  __syncthreads();
  if (threadIdx.x < kGroups) {
    const int g = threadIdx.x;
    float S = 0.f, Q = 0.f;
    for (int k = 0; k < cg; ++k) { S += sp.ch0[g * cg + k]; Q += sp.ch1[g * cg + k]; }
    const float cnt = (float)cg * (float)P;
    const float mu = S / cnt;
    const float var = fmaxf(Q / cnt - mu * mu, 0.f);
    sp.mu[g] = mu;
    sp.rs[g] = rsqrtf(var + eps);
    if (stats_out != nullptr) { stats_out[2 * g] = S; stats_out[2 * g + 1] = Q; }
  }
  __syncthreads();
}

__device__ __forceinline__ void group_mu_rs_from_stats(const float* stats_n, int cg, int P, float eps, const TailPtrs& sp) {
  if (threadIdx.x < kGroups) {
    const int g = threadIdx.x;
    const float cnt = (float)cg * (float)P;
    const float mu = stats_n[2 * g] / cnt;
    const float var = fmaxf(stats_n[2 * g + 1] / cnt - mu * mu, 0.f);
    sp.mu[g] = mu;
    sp.rs[g] = rsqrtf(var + eps);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// forward tails.  FULL = false: y = silu(GN(x)).  FULL = true: out = a*se*gate with a = silu(GN(x)).
// ---------------------------------------------------------------------------------------------------------------
template <typename T, bool FULL>
__global__ void __launch_bounds__(kFT, 1)
convblock_tail_fwd_kernel(const T* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                          const float* __restrict__ w1, const float* __restrict__ w2, const float* __restrict__ wsp,
                          float* __restrict__ stats, float* __restrict__ pool_g, float* __restrict__ se_g,
                          float* __restrict__ hid_g, T* __restrict__ out, int H, int W, int C, int Cr, float eps) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int n = blockIdx.x, P = H * W, cv = C / 8, cg = C / kGroups, nvec = P * cv, Wp = W + 6;
  const TailSmem L = tail_smem_layout(H, W, C, (int)sizeof(T), FULL ? 1 : 0, 0);
  T* s_img = reinterpret_cast<T*>(smem + L.img);
  const TailPtrs sp = tail_ptrs(smem, L, C);
  const T* xn = x + (size_t)n * P * C;
  T* on = out + (size_t)n * P * C;
  const int cb = threadIdx.x % cv;

  for (int v = threadIdx.x; v < nvec; v += kFT) {
    float t[8];
    load8(xn + (size_t)v * 8, t);
    store8(s_img + (size_t)v * 8, t);
  }
  for (int i = threadIdx.x; i < 5 * C; i += kFT) sp.ch0[i] = 0.f;
  if (FULL) {
    float2* s_cmap = reinterpret_cast<float2*>(smem + L.cmap);
    for (int i = threadIdx.x; i < (H + 6) * Wp; i += kFT) s_cmap[i] = make_float2(0.f, 0.f);
    for (int i = threadIdx.x; i < 98; i += kFT) sp.w[i] = __ldg(wsp + i);
  }
  __syncthreads();
  image_group_stats<T>(s_img, nvec, cv, cg, P, eps, sp, stats + (size_t)n * kGroups * 2);

  float ga[8], be[8], acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cb * 8 + j, g = c / cg;
    ga[j] = __ldg(gamma + c) * sp.rs[g];
    be[j] = fmaf(-sp.mu[g], ga[j], __ldg(beta + c));
    acc[j] = 0.f;
  }
  for (int v = threadIdx.x; v < nvec; v += kFT) {
    float t[8];
    load8_rw(s_img + (size_t)v * 8, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float z = fmaf(ga[j], t[j], be[j]);
      t[j] = round_to<T>(z * sigmoid_t<T>(z));
      acc[j] += t[j];
    }
    if (FULL) store8(s_img + (size_t)v * 8, t);
    else store8(on + (size_t)v * 8, t);
  }
  if (!FULL) return;

  // ---- squeeze / excite (SEBlock.forward, src/unet.py:16-17)
  chan_add(acc, sp.ch2, cb, cv);
  __syncthreads();
  const float invP = 1.f / (float)P;
  for (int c = threadIdx.x; c < C; c += kFT) pool_g[(size_t)n * C + c] = sp.ch2[c];
  for (int j = threadIdx.x; j < Cr; j += kFT) {
    float a = 0.f;
    for (int c = 0; c < C; ++c) a = fmaf(__ldg(w1 + (size_t)j * C + c), sp.ch2[c] * invP, a);
    a = fmaxf(a, 0.f);
    sp.hid[j] = a;
    hid_g[(size_t)n * Cr + j] = a;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += kFT) {
    float a = 0.f;
    for (int j = 0; j < Cr; ++j) a = fmaf(__ldg(w2 + (size_t)c * Cr + j), sp.hid[j], a);
    a = sigmoidf_(a);
    sp.se[c] = a;
    se_g[(size_t)n * C + c] = a;
  }
  __syncthreads();

  // ---- channel mean / max of u = a*se (SpatialGate.forward, src/unet.py:26-27), zero-padded by 3
  float2* s_cmap = reinterpret_cast<float2*>(smem + L.cmap);
  float* s_gate = reinterpret_cast<float*>(smem + L.gate);
  for (int p = threadIdx.x; p < P; p += kFT) {
    float sum = 0.f, mx = -INFINITY;
    for (int k = 0; k < cv; ++k) {
      float t[8];
      load8_rw(s_img + ((size_t)p * cv + k) * 8, t);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float u = t[j] * sp.se[k * 8 + j];
        sum += u;
        mx = fmaxf(mx, u);
      }
    }
    s_cmap[(p / W + 3) * Wp + (p % W) + 3] = make_float2(sum / (float)C, mx);
  }
  __syncthreads();
  // ---- gate = sigmoid(conv7x7([mean, max]))  (:28)
  for (int p = threadIdx.x; p < P; p += kFT) {
    const float2* t0 = s_cmap + (p / W) * Wp + (p % W);
    float q = 0.f;
#pragma unroll
    for (int dy = 0; dy < 7; ++dy) {
#pragma unroll
      for (int dx = 0; dx < 7; ++dx) {
        const float2 m = t0[dy * Wp + dx];
        q = fmaf(sp.w[dy * 7 + dx], m.x, q);
        q = fmaf(sp.w[49 + dy * 7 + dx], m.y, q);
      }
    }
    s_gate[p] = sigmoidf_(q);
  }
  __syncthreads();
  float sc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) sc[j] = sp.se[cb * 8 + j];
  for (int v = threadIdx.x; v < nvec; v += kFT) {
    float t[8];
    load8_rw(s_img + (size_t)v * 8, t);
    const float gt = s_gate[v / cv];
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = t[j] * sc[j] * gt;
    store8(on + (size_t)v * 8, t);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// backward of tail 1: da (gradient w.r.t. a = silu(GN(x))) -> dx ; dgamma, dbeta accumulate
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kFT, 1)
gn_silu_img_bwd_kernel(const T* __restrict__ da, const T* __restrict__ x, const float* __restrict__ stats,
                       const float* __restrict__ gamma, const float* __restrict__ beta, T* __restrict__ dx,
                       float* __restrict__ dgamma, float* __restrict__ dbeta, int H, int W, int C, float eps) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int n = blockIdx.x, P = H * W, cv = C / 8, cg = C / kGroups, nvec = P * cv;
  const TailSmem L = tail_smem_layout(H, W, C, (int)sizeof(T), 0, 1);
  T* s_img = reinterpret_cast<T*>(smem + L.img);
  const TailPtrs sp = tail_ptrs(smem, L, C);
  const T* xn = x + (size_t)n * P * C;
  const T* dan = da + (size_t)n * P * C;
  T* dxn = dx + (size_t)n * P * C;
  const int cb = threadIdx.x % cv;
  for (int v = threadIdx.x; v < nvec; v += kFT) {
    float t[8];
    load8(xn + (size_t)v * 8, t);
    store8(s_img + (size_t)v * 8, t);
  }
  for (int i = threadIdx.x; i < 5 * C; i += kFT) sp.ch0[i] = 0.f;
  group_mu_rs_from_stats(stats + (size_t)n * kGroups * 2, cg, P, eps, sp);
  __syncthreads();
  float mu[8], rs[8], ga[8], be[8], r0[8], r1[8], r2[8], r3[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cb * 8 + j, g = c / cg;
    mu[j] = sp.mu[g]; rs[j] = sp.rs[g];
    ga[j] = __ldg(gamma + c); be[j] = __ldg(beta + c);
    r0[j] = r1[j] = r2[j] = r3[j] = 0.f;
  }
  // pass 1: dxhat (stored to dx as scratch) and the reductions
  for (int v = threadIdx.x; v < nvec; v += kFT) {
    float t[8], d[8];
    load8_rw(s_img + (size_t)v * 8, t);
    load8(dan + (size_t)v * 8, d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = (t[j] - mu[j]) * rs[j];
      const float z = fmaf(ga[j], xh, be[j]);
      const float sg = sigmoid_t<T>(z);
      const float dz = d[j] * sg * (1.f + z * (1.f - sg));
      const float dxh = round_to<T>(dz * ga[j]);
      r0[j] = fmaf(dz, xh, r0[j]);
      r1[j] += dz;
      r2[j] += dxh;
      r3[j] = fmaf(dxh, xh, r3[j]);
      d[j] = dxh;
    }
    store8(dxn + (size_t)v * 8, d);
  }
  chan_add(r0, sp.ch0, cb, cv);
  chan_add(r1, sp.ch1, cb, cv);
  chan_add(r2, sp.ch2, cb, cv);
  chan_add(r3, sp.ch3, cb, cv);
  __syncthreads();
  if (threadIdx.x < kGroups) {
    const int g = threadIdx.x;
    float a = 0.f, b = 0.f;
    for (int k = 0; k < cg; ++k) { a += sp.ch2[g * cg + k]; b += sp.ch3[g * cg + k]; }
    const float cnt = (float)cg * (float)P;
    sp.m1[g] = a / cnt;
    sp.m2[g] = b / cnt;
  }
  for (int c = threadIdx.x; c < C; c += kFT) {
    atomicAdd(dgamma + c, sp.ch0[c]);
    atomicAdd(dbeta + c, sp.ch1[c]);
  }
  __syncthreads();
  float m1[8], m2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { const int g = (cb * 8 + j) / cg; m1[j] = sp.m1[g]; m2[j] = sp.m2[g]; }
  // pass 2: every thread re-reads exactly the vectors it wrote in pass 1
  for (int v = threadIdx.x; v < nvec; v += kFT) {
    float t[8], d[8];
    load8_rw(s_img + (size_t)v * 8, t);
    load8_rw(dxn + (size_t)v * 8, d);     // plain loads: written by this same thread in pass 1
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = (t[j] - mu[j]) * rs[j];
      d[j] = rs[j] * (d[j] - m1[j] - xh * m2[j]);
    }
    store8(dxn + (size_t)v * 8, d);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// backward of tail 2: dout -> dx (gradient w.r.t. the conv2 output), all parameter gradients of GN2 / SE / gate
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kFT, 1)
convblock_tail_bwd_kernel(const T* __restrict__ dout, const T* __restrict__ x, const float* __restrict__ stats,
                          const float* __restrict__ gamma, const float* __restrict__ beta,
                          const float* __restrict__ w1, const float* __restrict__ w2, const float* __restrict__ wsp,
                          const float* __restrict__ pool_g, const float* __restrict__ se_g,
                          const float* __restrict__ hid_g, T* __restrict__ dx, float* __restrict__ dgamma,
                          float* __restrict__ dbeta, float* __restrict__ dw1, float* __restrict__ dw2,
                          float* __restrict__ dwsp, int H, int W, int C, int Cr, float eps) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int n = blockIdx.x, P = H * W, cv = C / 8, cg = C / kGroups, nvec = P * cv, Wp = W + 6;
  const TailSmem L = tail_smem_layout(H, W, C, (int)sizeof(T), 1, 1);
  T* s_img = reinterpret_cast<T*>(smem + L.img);
  float2* s_cmap = reinterpret_cast<float2*>(smem + L.cmap);
  float* s_dq = reinterpret_cast<float*>(smem + L.dq);
  float* s_gate = reinterpret_cast<float*>(smem + L.gate);
  float2* s_dm = reinterpret_cast<float2*>(smem + L.dm);
  uint8_t* s_cnt = smem + L.cnt;
  const TailPtrs sp = tail_ptrs(smem, L, C);
  const T* xn = x + (size_t)n * P * C;
  const T* don = dout + (size_t)n * P * C;
  T* dxn = dx + (size_t)n * P * C;
  const int cb = threadIdx.x % cv;
  const float invP = 1.f / (float)P;

  for (int v = threadIdx.x; v < nvec; v += kFT) {
    float t[8];
    load8(xn + (size_t)v * 8, t);
    store8(s_img + (size_t)v * 8, t);
  }
  for (int i = threadIdx.x; i < 5 * C; i += kFT) sp.ch0[i] = 0.f;
  for (int i = threadIdx.x; i < (H + 6) * Wp; i += kFT) { s_cmap[i] = make_float2(0.f, 0.f); s_dq[i] = 0.f; }
  for (int i = threadIdx.x; i < 98; i += kFT) { sp.w[i] = __ldg(wsp + i); sp.dw[i] = 0.f; }
  for (int c = threadIdx.x; c < C; c += kFT) {
    sp.se[c] = __ldg(se_g + (size_t)n * C + c);
    sp.pool[c] = __ldg(pool_g + (size_t)n * C + c);
  }
  for (int j = threadIdx.x; j < Cr; j += kFT) sp.hid[j] = __ldg(hid_g + (size_t)n * Cr + j);
  group_mu_rs_from_stats(stats + (size_t)n * kGroups * 2, cg, P, eps, sp);
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += kFT) {       // z = ca*x + cb_  per channel (for the per-pixel passes)
    const int g = c / cg;
    const float a = __ldg(gamma + c) * sp.rs[g];
    sp.ca[c] = a;
    sp.cb_[c] = fmaf(-sp.mu[g], a, __ldg(beta + c));
  }
  __syncthreads();

  // ---- per pixel: u = a*se -> (mean, max, #ties) ; acc = sum_c dout*a*se   (sigmoid #1)
  for (int p = threadIdx.x; p < P; p += kFT) {
    float sum = 0.f, mx = -INFINITY, acc = 0.f;
    int cnt = 0;
    for (int k = 0; k < cv; ++k) {
      float t[8], d[8];
      load8_rw(s_img + ((size_t)p * cv + k) * 8, t);
      load8(don + ((size_t)p * cv + k) * 8, d);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = k * 8 + j;
        const float z = fmaf(sp.ca[c], t[j], sp.cb_[c]);
        const float a = round_to<T>(z * sigmoid_t<T>(z));
        const float u = a * sp.se[c];
        sum += u;
        if (u > mx) { mx = u; cnt = 1; } else if (u == mx) { ++cnt; }
        acc = fmaf(d[j], u, acc);
      }
    }
    const int ip = (p / W + 3) * Wp + (p % W) + 3;
    s_cmap[ip] = make_float2(sum / (float)C, mx);
    s_cnt[p] = (uint8_t)min(cnt, 255);
    s_dq[ip] = acc;
  }
  __syncthreads();
  // ---- gate and dq = acc * gate * (1 - gate)
  for (int p = threadIdx.x; p < P; p += kFT) {
    const float2* t0 = s_cmap + (p / W) * Wp + (p % W);
    float q = 0.f;
#pragma unroll
    for (int dy = 0; dy < 7; ++dy) {
#pragma unroll
      for (int dxx = 0; dxx < 7; ++dxx) {
        const float2 m = t0[dy * Wp + dxx];
        q = fmaf(sp.w[dy * 7 + dxx], m.x, q);
        q = fmaf(sp.w[49 + dy * 7 + dxx], m.y, q);
      }
    }
    const float gt = sigmoidf_(q);
    s_gate[p] = gt;
    const int ip = (p / W + 3) * Wp + (p % W) + 3;
    s_dq[ip] *= gt * (1.f - gt);
  }
  __syncthreads();
  // ---- dwsp[k][dy][dx] += sum_p dq[p] * cmap_k[p + (dy-3, dx-3)] : 98 taps x 5 row partitions
  if (threadIdx.x < 490) {
    const int tap = threadIdx.x % 98, part = threadIdx.x / 98;
    const int k = tap / 49, dy = (tap % 49) / 7, dxx = tap % 7;
    const float* cm = reinterpret_cast<const float*>(s_cmap) + k;
    float a0 = 0.f, a1 = 0.f;
    for (int h = part; h < H; h += 5) {
      const float* drow = s_dq + (h + 3) * Wp + 3;
      const float* crow = cm + 2 * ((h + dy) * Wp + dxx);
      int w = 0;
      for (; w + 1 < W; w += 2) {
        a0 = fmaf(drow[w], crow[2 * w], a0);
        a1 = fmaf(drow[w + 1], crow[2 * w + 2], a1);
      }
      if (w < W) a0 = fmaf(drow[w], crow[2 * w], a0);
    }
    atomicAdd(&sp.dw[tap], a0 + a1);
  }
  // ---- gradient reaching (mean, max) through the transposed stencil
  for (int p = threadIdx.x; p < P; p += kFT) {
    const float* t0 = s_dq + (p / W + 6) * Wp + (p % W) + 6;
    float d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int dy = 0; dy < 7; ++dy) {
#pragma unroll
      for (int dxx = 0; dxx < 7; ++dxx) {
        const float d = t0[-(dy * Wp + dxx)];
        d0 = fmaf(sp.w[dy * 7 + dxx], d, d0);
        d1 = fmaf(sp.w[49 + dy * 7 + dxx], d, d1);
      }
    }
    s_dm[p] = make_float2(d0 / (float)C, d1 / (float)max((int)s_cnt[p], 1));
  }
  __syncthreads();
  if (threadIdx.x < 98) atomicAdd(dwsp + threadIdx.x, sp.dw[threadIdx.x]);

  // ---- du = dout*gate + dmean + [u == max]*dmax/ties ; r = du*se (scratch in dx) ; dse = sum_p du*a  (sigmoid #2)
  float ga[8], be[8], sc[8], acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cb * 8 + j;
    ga[j] = sp.ca[c]; be[j] = sp.cb_[c]; sc[j] = sp.se[c]; acc[j] = 0.f;
  }
  for (int v = threadIdx.x; v < nvec; v += kFT) {
    const int pix = v / cv;
    float t[8], d[8];
    load8_rw(s_img + (size_t)v * 8, t);
    load8(don + (size_t)v * 8, d);
    const float gt = s_gate[pix];
    const float2 dm = s_dm[pix];
    const float mx = s_cmap[(pix / W + 3) * Wp + (pix % W) + 3].y;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float z = fmaf(ga[j], t[j], be[j]);
      const float a = round_to<T>(z * sigmoid_t<T>(z));
      const float u = a * sc[j];
      const float du = d[j] * gt + dm.x + ((u == mx) ? dm.y : 0.f);
      acc[j] = fmaf(du, a, acc[j]);
      d[j] = du * sc[j];
    }
    store8(dxn + (size_t)v * 8, d);
  }
  chan_add(acc, sp.ch0, cb, cv);
  __syncthreads();
  // ---- SE backward (tiny): dpool, dw1, dw2
  for (int c = threadIdx.x; c < C; c += kFT) {
    const float s = sp.se[c];
    sp.dpre2[c] = sp.ch0[c] * s * (1.f - s);
  }
  __syncthreads();
  for (int j = threadIdx.x; j < Cr; j += kFT) {
    float a = 0.f;
    for (int c = 0; c < C; ++c) a = fmaf(__ldg(w2 + (size_t)c * Cr + j), sp.dpre2[c], a);
    sp.dpre1[j] = sp.hid[j] > 0.f ? a : 0.f;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += kFT) {
    float a = 0.f;
    for (int j = 0; j < Cr; ++j) a = fmaf(__ldg(w1 + (size_t)j * C + c), sp.dpre1[j], a);
    sp.dpool[c] = a * invP;
  }
  for (int i = threadIdx.x; i < C * Cr; i += kFT) {
    {
      const int c = i / Cr, j = i % Cr;
      const float v = sp.dpre2[c] * sp.hid[j];
      if (v != 0.f) atomicAdd(dw2 + i, v);
    }
    {
      const int j = i / C, c = i % C;
      const float v = sp.dpre1[j] * sp.pool[c] * invP;
      if (v != 0.f) atomicAdd(dw1 + i, v);
    }
  }
  __syncthreads();
  // ---- GroupNorm + SiLU backward, pass 1 (sigmoid #3): dxhat -> scratch, reductions
  float mu[8], rs[8], gm[8], bt[8], dp[8], r0[8], r1[8], r2[8], r3[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cb * 8 + j, g = c / cg;
    mu[j] = sp.mu[g]; rs[j] = sp.rs[g];
    gm[j] = __ldg(gamma + c); bt[j] = __ldg(beta + c);
    dp[j] = sp.dpool[c];
    r0[j] = r1[j] = r2[j] = r3[j] = 0.f;
  }
  for (int v = threadIdx.x; v < nvec; v += kFT) {
    float t[8], d[8];
    load8_rw(s_img + (size_t)v * 8, t);
    load8_rw(dxn + (size_t)v * 8, d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = (t[j] - mu[j]) * rs[j];
      const float z = fmaf(gm[j], xh, bt[j]);
      const float sg = sigmoid_t<T>(z);
      const float dz = (d[j] + dp[j]) * sg * (1.f + z * (1.f - sg));
      const float dxh = round_to<T>(dz * gm[j]);
      r0[j] = fmaf(dz, xh, r0[j]);
      r1[j] += dz;
      r2[j] += dxh;
      r3[j] = fmaf(dxh, xh, r3[j]);
      d[j] = dxh;
    }
    store8(dxn + (size_t)v * 8, d);
  }
  chan_add(r0, sp.ch1, cb, cv);
  chan_add(r1, sp.ch2, cb, cv);
  chan_add(r2, sp.ch3, cb, cv);
  chan_add(r3, sp.ch4, cb, cv);
  __syncthreads();
  if (threadIdx.x < kGroups) {
    const int g = threadIdx.x;
    float a = 0.f, b = 0.f;
    for (int k = 0; k < cg; ++k) { a += sp.ch3[g * cg + k]; b += sp.ch4[g * cg + k]; }
    const float cnt = (float)cg * (float)P;
    sp.m1[g] = a / cnt;
    sp.m2[g] = b / cnt;
  }
  for (int c = threadIdx.x; c < C; c += kFT) {
    atomicAdd(dgamma + c, sp.ch1[c]);
    atomicAdd(dbeta + c, sp.ch2[c]);
  }
  __syncthreads();
  float m1[8], m2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { const int g = (cb * 8 + j) / cg; m1[j] = sp.m1[g]; m2[j] = sp.m2[g]; }
  for (int v = threadIdx.x; v < nvec; v += kFT) {
    float t[8], d[8];
    load8_rw(s_img + (size_t)v * 8, t);
    load8_rw(dxn + (size_t)v * 8, d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xh = (t[j] - mu[j]) * rs[j];
      d[j] = rs[j] * (d[j] - m1[j] - xh * m2[j]);
    }
    store8(dxn + (size_t)v * 8, d);
  }
}

static bool fused_shape_ok(int H, int W, int C, int Cr) {
  const int cv = C / 8;
  return C % 8 == 0 && C >= 8 && cv <= 32 && (cv & (cv - 1)) == 0 && Cr >= 1 && Cr <= 64 && H >= 1 && W >= 1;
}

template <typename K>
static int tail_set_smem(K kern, size_t bytes, const char* what) {
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) { set_error("%s: smem attribute (%zu B): %s", what, bytes, cudaGetErrorString(e)); return PCM_ERR_CUDA; }
  }
  return PCM_OK;
}

}  // namespace pcm

using namespace pcm;

extern "C" int pcm_convblock_fused_supported(int H, int W, int C, int Cr, int dtype) {
  if (!fused_shape_ok(H, W, C, Cr)) return 0;
  const int elt = dtype == PCM_BF16 ? 2 : 4;
  return tail_smem_layout(H, W, C, elt, 1, 1).total <= 227 * 1024 ? 1 : 0;
}

extern "C" int pcm_gn_silu_img_fwd(const void* x, const float* gamma, const float* beta, float* stats, void* y, int N,
                                   int H, int W, int C, float eps, int dtype, pcm_stream_t s) {
  PCM_REQUIRE(fused_shape_ok(H, W, C, 1), "gn_silu_img_fwd: unsupported shape H=%d W=%d C=%d", H, W, C);
  if (N == 0) return PCM_OK;
  const size_t smem = tail_smem_layout(H, W, C, dtype == PCM_BF16 ? 2 : 4, 0, 0).total;
  PCM_REQUIRE(smem <= 227 * 1024, "gn_silu_img_fwd: image does not fit shared memory (%zu B)", smem);
  int rc = PCM_OK;
  PCM_DISPATCH_DTYPE(dtype, T, {
    rc = tail_set_smem(convblock_tail_fwd_kernel<T, false>, smem, "gn_silu_img_fwd");
    if (rc == PCM_OK)
      convblock_tail_fwd_kernel<T, false><<<N, kFT, smem, (cudaStream_t)s>>>(
          (const T*)x, gamma, beta, nullptr, nullptr, nullptr, stats, nullptr, nullptr, nullptr, (T*)y, H, W, C, 1, eps);
  });
  if (rc != PCM_OK) return rc;
  return check_launch("gn_silu_img_fwd");
}

extern "C" int pcm_convblock_tail_fwd(const void* x, const float* gamma, const float* beta, const float* w1,
                                      const float* w2, const float* wsp, float* stats, float* pool, float* se,
                                      float* hid, void* out, int N, int H, int W, int C, int Cr, float eps, int dtype,
                                      pcm_stream_t s) {
  PCM_REQUIRE(fused_shape_ok(H, W, C, Cr), "convblock_tail_fwd: unsupported shape H=%d W=%d C=%d Cr=%d", H, W, C, Cr);
  if (N == 0) return PCM_OK;
  const size_t smem = tail_smem_layout(H, W, C, dtype == PCM_BF16 ? 2 : 4, 1, 0).total;
  PCM_REQUIRE(smem <= 227 * 1024, "convblock_tail_fwd: image does not fit shared memory (%zu B)", smem);
  int rc = PCM_OK;
  PCM_DISPATCH_DTYPE(dtype, T, {
    rc = tail_set_smem(convblock_tail_fwd_kernel<T, true>, smem, "convblock_tail_fwd");
    if (rc == PCM_OK)
      convblock_tail_fwd_kernel<T, true><<<N, kFT, smem, (cudaStream_t)s>>>(
          (const T*)x, gamma, beta, w1, w2, wsp, stats, pool, se, hid, (T*)out, H, W, C, Cr, eps);
  });
  if (rc != PCM_OK) return rc;
  return check_launch("convblock_tail_fwd");
}

extern "C" int pcm_gn_silu_img_bwd(const void* da, const void* x, const float* stats, const float* gamma,
                                   const float* beta, void* dx, float* dgamma, float* dbeta, int N, int H, int W, int C,
                                   float eps, int dtype, pcm_stream_t s) {
  PCM_REQUIRE(fused_shape_ok(H, W, C, 1), "gn_silu_img_bwd: unsupported shape H=%d W=%d C=%d", H, W, C);
  if (N == 0) return PCM_OK;
  const size_t smem = tail_smem_layout(H, W, C, dtype == PCM_BF16 ? 2 : 4, 0, 1).total;
  PCM_REQUIRE(smem <= 227 * 1024, "gn_silu_img_bwd: image does not fit shared memory (%zu B)", smem);
  int rc = PCM_OK;
  PCM_DISPATCH_DTYPE(dtype, T, {
    rc = tail_set_smem(gn_silu_img_bwd_kernel<T>, smem, "gn_silu_img_bwd");
    if (rc == PCM_OK)
      gn_silu_img_bwd_kernel<T><<<N, kFT, smem, (cudaStream_t)s>>>((const T*)da, (const T*)x, stats, gamma, beta, (T*)dx,
                                                                   dgamma, dbeta, H, W, C, eps);
  });
  if (rc != PCM_OK) return rc;
  return check_launch("gn_silu_img_bwd");
}

extern "C" int pcm_convblock_tail_bwd(const void* dout, const void* x, const float* stats, const float* gamma,
                                      const float* beta, const float* w1, const float* w2, const float* wsp,
                                      const float* pool, const float* se, const float* hid, void* dx, float* dgamma,
                                      float* dbeta, float* dw1, float* dw2, float* dwsp, int N, int H, int W, int C,
                                      int Cr, float eps, int dtype, pcm_stream_t s) {
  PCM_REQUIRE(fused_shape_ok(H, W, C, Cr), "convblock_tail_bwd: unsupported shape H=%d W=%d C=%d Cr=%d", H, W, C, Cr);
  if (N == 0) return PCM_OK;
  const size_t smem = tail_smem_layout(H, W, C, dtype == PCM_BF16 ? 2 : 4, 1, 1).total;
  PCM_REQUIRE(smem <= 227 * 1024, "convblock_tail_bwd: image does not fit shared memory (%zu B)", smem);
  int rc = PCM_OK;
  PCM_DISPATCH_DTYPE(dtype, T, {
    rc = tail_set_smem(convblock_tail_bwd_kernel<T>, smem, "convblock_tail_bwd");
    if (rc == PCM_OK)
      convblock_tail_bwd_kernel<T><<<N, kFT, smem, (cudaStream_t)s>>>(
          (const T*)dout, (const T*)x, stats, gamma, beta, w1, w2, wsp, pool, se, hid, (T*)dx, dgamma, dbeta, dw1, dw2,
          dwsp, H, W, C, Cr, eps);
  });
  if (rc != PCM_OK) return rc;
  return check_launch("convblock_tail_bwd");
}
