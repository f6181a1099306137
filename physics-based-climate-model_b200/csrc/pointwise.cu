// Layout staging, pooling / skip aggregation, ConvLSTM cell pointwise, head + loss, Adam.
// All HBM-bound: 16-byte vector accesses, grids sized in multiples of the SM count.
#include "common.cuh"

namespace pcm {

constexpr int kGridCap = 148 * 8;

// ---- layout --------------------------------------------------------------------------------
// NCHW fp32 -> NHWC T (pad channels with 0). One thread = one pixel x 8 output channels; reads
// are coalesced across pixels (w fastest), writes are 16 B per thread.
template <typename T>
__global__ void __launch_bounds__(256)
nchw_to_nhwc_kernel(const float* __restrict__ x, T* __restrict__ y, int N, int C, int P, int Cp,
                    const int* __restrict__ month, int Tp) {
  PCM_PDL_ENTRY();
  const int cv = Cp / 8;
  const int Bn = Tp > 1 ? N / Tp : N;
  const long long total = (long long)N * cv * P;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(idx % P);
    const int cb = (int)((idx / P) % cv);
    const int n = (int)(idx / ((long long)P * cv));              // NHWC image index (t-major when Tp > 1)
    const int ni = Tp > 1 ? (n % Bn) * Tp + n / Bn : n;         // NCHW image index b*T + t
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = cb * 8 + j;
      v[j] = (c < C) ? __ldg(x + ((long long)ni * C + c) * P + p) : 0.f;
    }
    if (month != nullptr) {   // seasonal channels C, C+1 (main_final.py:188-196)
      const float ang = 6.283185307179586f * (float)__ldg(month + ni) / 12.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = cb * 8 + j;
        if (c == C) v[j] = sinf(ang);
        if (c == C + 1) v[j] = cosf(ang);
      }
    }
    store8(y + ((long long)n * P + p) * Cp + cb * 8, v);
  }
}

// Same, four consecutive pixels per thread: one 16-byte load per input channel keeps 4x the bytes in flight (the scalar
// version ran at 2 TB/s on the 74 MB fp32 input of a training batch).  Requires P % 4 == 0.
template <typename T>
__global__ void __launch_bounds__(256)
nchw_to_nhwc_x4_kernel(const float* __restrict__ x, T* __restrict__ y, int N, int C, int P, int Cp,
                       const int* __restrict__ month, int Tp) {
  PCM_PDL_ENTRY();
  const int cv = Cp / 8, P4 = P / 4;
  const int Bn = Tp > 1 ? N / Tp : N;
  const long long total = (long long)N * cv * P4;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int p4 = (int)(idx % P4);
    const int cb = (int)((idx / P4) % cv);
    const int n = (int)(idx / ((long long)P4 * cv));
    const int ni = Tp > 1 ? (n % Bn) * Tp + n / Bn : n;
    float v[4][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = cb * 8 + j;
      float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < C) q = __ldg(reinterpret_cast<const float4*>(x + ((long long)ni * C + c) * P) + p4);
      v[0][j] = q.x; v[1][j] = q.y; v[2][j] = q.z; v[3][j] = q.w;
    }
    if (month != nullptr) {
      const float ang = 6.283185307179586f * (float)__ldg(month + ni) / 12.f;
      const float sn = sinf(ang), cs = cosf(ang);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = cb * 8 + j;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (c == C) v[k][j] = sn;
          if (c == C + 1) v[k][j] = cs;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) store8(y + ((long long)n * P + 4 * p4 + k) * Cp + cb * 8, v[k]);
  }
}

// Sliding-window staging from a device-resident series (main_final.py:97-154): NHWC image n <- NCHW frame frames[n]
// of `series` (frames[n] < 0: the zero left-pad of windows that start before the record).  Four pixels per thread.
// Optional, fused while staging:
//   * norm != nullptr: Normalizer.normalize(data, "input") of src/utils_final.py:45-128 per channel, in fp64 like
//     numpy on the reference's arrays: norm[c] = (kind, a, b, lambda), x_n = (g(x) - b) * a with g = identity
//     (kind 0: zscore a = 1/(std+1e-8), b = mean; minimax a = 1/range, b = min), log1p (1), sqrt (2), x^lambda (3);
//     kind < 0: pass through (no config for that channel).
//   * month != nullptr: channels C, C+1 = sin / cos(2 pi month[frame] / 12), the seasonal channels of
//     main_final.py:186-216 (not normalised: the reference has no statistics entry for them).
// Left-pad frames stay all-zero (the reference pads with zeros of the already normalised tensor).
__device__ __forceinline__ float norm_apply(float x, int kind, double a, double b, double lam) {
  if (kind < 0) return x;
  double u = (double)x;
  u = kind == 0 ? u : kind == 1 ? log1p(u) : kind == 2 ? sqrt(u) : pow(u, lam);
  return (float)((u - b) * a);
}

template <typename T>
__global__ void __launch_bounds__(256)
window_stage_kernel(const float* __restrict__ series, const int* __restrict__ frames, T* __restrict__ y, int N, int C,
                    int P, int Cp, const double* __restrict__ norm, const int* __restrict__ month) {
  PCM_PDL_ENTRY();
  const int cv = Cp / 8, P4 = P / 4;
  const long long total = (long long)N * cv * P4;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int p4 = (int)(idx % P4);
    const int cb = (int)((idx / P4) % cv);
    const int n = (int)(idx / ((long long)P4 * cv));
    const int f = __ldg(frames + n);
    float v[4][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = cb * 8 + j;
      float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < C && f >= 0) {
        q = __ldg(reinterpret_cast<const float4*>(series + ((long long)f * C + c) * P) + p4);
        if (norm != nullptr) {
          const int kind = (int)__ldg(norm + 4 * c);
          const double a = __ldg(norm + 4 * c + 1), b = __ldg(norm + 4 * c + 2), lam = __ldg(norm + 4 * c + 3);
          q.x = norm_apply(q.x, kind, a, b, lam); q.y = norm_apply(q.y, kind, a, b, lam);
          q.z = norm_apply(q.z, kind, a, b, lam); q.w = norm_apply(q.w, kind, a, b, lam);
        }
      }
      v[0][j] = q.x; v[1][j] = q.y; v[2][j] = q.z; v[3][j] = q.w;
    }
    if (month != nullptr && f >= 0) {
      const float ang = 6.283185307179586f * (float)__ldg(month + f) / 12.f;
      const float sn = sinf(ang), cs = cosf(ang);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = cb * 8 + j;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (c == C) v[k][j] = sn;
          if (c == C + 1) v[k][j] = cs;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) store8(y + ((long long)n * P + 4 * p4 + k) * Cp + cb * 8, v[k]);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
nhwc_to_nchw_kernel(const T* __restrict__ x, float* __restrict__ y, int N, int C, int P, int Cp, int Tp) {
  PCM_PDL_ENTRY();
  const int cv = Cp / 8;
  const int Bn = Tp > 1 ? N / Tp : N;
  const long long total = (long long)N * cv * P;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(idx % P);
    const int cb = (int)((idx / P) % cv);
    const int n = (int)(idx / ((long long)P * cv));
    const int no = Tp > 1 ? (n % Bn) * Tp + n / Bn : n;
    float v[8];
    load8(x + ((long long)n * P + p) * Cp + cb * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = cb * 8 + j;
      if (c < C) y[((long long)no * C + c) * P + p] = v[j];
    }
  }
}

// ---- pooling / skips -------------------------------------------------------------------------
template <typename T, typename I>          // I: index type (int whenever the tensor allows it)
__global__ void __launch_bounds__(256)
maxpool2_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int N, int H, int W, int C) {
  PCM_PDL_ENTRY();
  const int Ho = H / 2, Wo = W / 2, cv = C / 8;
  const I total = (I)N * Ho * Wo * cv;
  for (I idx = (I)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (I)gridDim.x * blockDim.x) {
    const int cb = (int)(idx % cv);
    I r = idx / cv;
    const int wo = (int)(r % Wo); r /= Wo;
    const int ho = (int)(r % Ho);
    const int n = (int)(r / Ho);
    const T* xp = x + (((long long)n * H + 2 * ho) * W + 2 * wo) * C + cb * 8;
    float a[8], b[8], c[8], d[8];
    load8(xp, a); load8(xp + C, b); load8(xp + (long long)W * C, c); load8(xp + (long long)W * C + C, d);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = fmaxf(fmaxf(a[j], b[j]), fmaxf(c[j], d[j]));
    store8(y + (size_t)idx * 8, a);
  }
}

// dx = [x is the FIRST max of its window (scan order)] * dy + dskip/T.  One thread per 2x2 window and channel block:
// the four x vectors are read once (not once per output pixel), nine independent 16-byte loads are in flight per
// thread, and the index arithmetic is 32-bit whenever the tensor allows it (I = int).
template <typename T, typename I, bool DOT>
__global__ void __launch_bounds__(256, 4)
maxpool2_bwd_skip_kernel(const T* __restrict__ x, const T* __restrict__ dy, const T* __restrict__ dskip,
                         long long dskip_ns, int dskip_ps, T* __restrict__ dx, float* __restrict__ sdot, int N, int H, int W,
                         int C, int Tn, int t_major) {
  // DOT: also sdot[n][h][w] = sum_c dx(n,h,w,c) * x(n,h,w,c), with dx as the storage type rounds it — the sum the next
  // kernel of the backward chain (the ConvBlock tail whose output x is) needs for its gate gradient; both operands pass
  // through this thread, so that kernel does not have to stream dx and x a second time.  x is RE-LOADED per pixel for the
  // product (an L1 hit: this thread read the same 16 bytes for the arg-max a moment ago): holding the 2x2 window of x in
  // registers until the end doubled the register count and cost more than the re-load (37 vs 29 us at 48x72x16).
  PCM_PDL_ENTRY();
  const int Ho = H / 2, Wo = W / 2, Hc = (H + 1) / 2, Wc = (W + 1) / 2, cv = C / 8;
  const float invT = 1.f / (float)Tn;
  const I total = (I)N * Hc * Wc * cv;
  for (I idx = (I)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (I)gridDim.x * blockDim.x) {
    const int cb = (int)(idx % cv);
    I r = idx / cv;
    const int wc = (int)(r % Wc); r /= Wc;
    const int hc = (int)(r % Hc);
    const int n = (int)(r / Hc);
    const int h0 = 2 * hc, w0 = 2 * wc;
    const bool has_h1 = h0 + 1 < H, has_w1 = w0 + 1 < W, pooled = dy != nullptr && hc < Ho && wc < Wo;
    const size_t base = (((size_t)n * H + h0) * W + w0) * C + (size_t)cb * 8;
    const size_t offs[4] = {0, (size_t)C, (size_t)W * C, (size_t)W * C + C};
    const bool live[4] = {true, has_w1, has_h1, has_h1 && has_w1};
    float o[4][8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[k][j] = 0.f;
    }
    if (pooled) {
      float q[4][8], g[8];
#pragma unroll
      for (int k = 0; k < 4; ++k) load8(x + base + offs[k], q[k]);
      load8(dy + (((size_t)n * Ho + hc) * Wo + wc) * C + (size_t)cb * 8, g);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        int arg = 0;
        float m = q[0][j];
#pragma unroll
        for (int k = 1; k < 4; ++k)
          if (q[k][j] > m) { m = q[k][j]; arg = k; }
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k][j] = (arg == k) ? g[j] : 0.f;
      }
    }
    if (dskip != nullptr) {
      const int bi = t_major ? n % (N / Tn) : n / Tn;
      const T* sp = dskip + (size_t)bi * dskip_ns + (size_t)cb * 8;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (live[k]) {
          float sk[8];
          load8(sp + ((size_t)(h0 + (k >> 1)) * W + (w0 + (k & 1))) * dskip_ps, sk);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[k][j] = fmaf(sk[j], invT, o[k][j]);
        }
      }
    }
    float sd[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (live[k]) {
        store8(dx + base + offs[k], o[k]);
        if (DOT) {
          float qk[8];
          load8(x + base + offs[k], qk);
#pragma unroll
          for (int j = 0; j < 8; ++j) sd[k] = fmaf(round_to<T>(o[k][j]), qk[j], sd[k]);
        }
      }
    }
    if (DOT) {
      // the cv threads of a 2x2 window are adjacent lanes (cb = idx % cv, cv a power of two <= 32 dividing the grid
      // stride) and run the same iterations, so the shuffles are convergent within the group
      const unsigned lane = threadIdx.x & 31u, grp = (lane / (unsigned)cv) * (unsigned)cv;
      const unsigned mask = cv == 32 ? 0xffffffffu : ((1u << cv) - 1u) << grp;
      for (int off = 1; off < cv; off <<= 1) {
#pragma unroll
        for (int k = 0; k < 4; ++k) sd[k] += __shfl_xor_sync(mask, sd[k], off);
      }
      if (cb == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (live[k]) sdot[((size_t)n * H + h0 + (k >> 1)) * W + w0 + (k & 1)] = sd[k];
      }
    }
  }
}

template <typename T, typename I>
__global__ void __launch_bounds__(256)
time_mean_kernel(const T* __restrict__ src, T* __restrict__ dst, long long dst_ns, int dst_ps, int B, int Tn, int P,
                 int C, int t_major) {
  PCM_PDL_ENTRY();
  const int cv = C / 8;
  const float invT = 1.f / (float)Tn;
  const I total = (I)B * P * cv;
  for (I idx = (I)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (I)gridDim.x * blockDim.x) {
    const int cb = (int)(idx % cv);
    const int p = (int)((idx / cv) % P);
    const int b = (int)(idx / ((I)cv * P));
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll 6
    for (int t = 0; t < Tn; ++t) {
      float v[8];
      const long long img = t_major ? (long long)t * B + b : (long long)b * Tn + t;
      load8(src + (img * P + p) * C + cb * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] *= invT;
    store8(dst + b * dst_ns + (long long)p * dst_ps + cb * 8, acc);
  }
}

// ---- ConvLSTM cell (src/convlstm.py:14-18) ------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
lstm_cell_fwd_kernel(const float* __restrict__ gates, const float* __restrict__ c_prev, T* __restrict__ acts,
                     float* __restrict__ c, T* __restrict__ h, int M, int Ch) {
  PCM_PDL_ENTRY();
  const int cv = Ch / 8;
  const long long total = (long long)M * cv;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int cb = (int)(idx % cv);
    const long long m = idx / cv;
    const float* gp = gates + m * 4 * Ch + cb * 8;
    float gi[8], gf[8], go[8], gg[8], cp[8], cn[8], hn[8];
    load8(gp, gi); load8(gp + Ch, gf); load8(gp + 2 * Ch, go); load8(gp + 3 * Ch, gg);
    if (c_prev) load8(c_prev + m * Ch + cb * 8, cp);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      gi[j] = round_to<T>(sigmoidf_(gi[j]));
      gf[j] = round_to<T>(sigmoidf_(gf[j]));
      go[j] = round_to<T>(sigmoidf_(go[j]));
      gg[j] = round_to<T>(tanhf(gg[j]));
      cn[j] = fmaf(gf[j], c_prev ? cp[j] : 0.f, gi[j] * gg[j]);
      hn[j] = go[j] * tanhf(cn[j]);
    }
    T* ap = acts + m * 4 * Ch + cb * 8;
    store8(ap, gi); store8(ap + Ch, gf); store8(ap + 2 * Ch, go); store8(ap + 3 * Ch, gg);
    store8(c + m * Ch + cb * 8, cn);
    store8(h + m * Ch + cb * 8, hn);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
lstm_cell_bwd_kernel(const T* __restrict__ dh_a, const T* __restrict__ dh_b, const float* __restrict__ dc_in,
                     const T* __restrict__ acts, const float* __restrict__ c_prev, const float* __restrict__ c,
                     T* __restrict__ dgates, float* __restrict__ dc_prev, int M, int Ch) {
  PCM_PDL_ENTRY();
  const int cv = Ch / 8;
  const long long total = (long long)M * cv;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int cb = (int)(idx % cv);
    const long long m = idx / cv;
    const long long o1 = m * Ch + cb * 8;
    float dh[8], t[8], dc[8], gi[8], gf[8], go[8], gg[8], cp[8], cc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) dh[j] = dc[j] = cp[j] = 0.f;
    if (dh_a) load8(dh_a + o1, dh);
    if (dh_b) {
      load8(dh_b + o1, t);
#pragma unroll
      for (int j = 0; j < 8; ++j) dh[j] += t[j];
    }
    if (dc_in) load8(dc_in + o1, dc);
    if (c_prev) load8(c_prev + o1, cp);
    load8(c + o1, cc);
    const T* ap = acts + m * 4 * Ch + cb * 8;
    load8(ap, gi); load8(ap + Ch, gf); load8(ap + 2 * Ch, go); load8(ap + 3 * Ch, gg);
    float di[8], df[8], dO[8], dg[8], dcp[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float tc = tanhf(cc[j]);
      const float dct = fmaf(dh[j] * go[j], 1.f - tc * tc, dc[j]);
      dO[j] = dh[j] * tc * go[j] * (1.f - go[j]);
      di[j] = dct * gg[j] * gi[j] * (1.f - gi[j]);
      df[j] = dct * cp[j] * gf[j] * (1.f - gf[j]);
      dg[j] = dct * gi[j] * (1.f - gg[j] * gg[j]);
      dcp[j] = dct * gf[j];
    }
    T* dp = dgates + m * 4 * Ch + cb * 8;
    store8(dp, di); store8(dp + Ch, df); store8(dp + 2 * Ch, dO); store8(dp + 3 * Ch, dg);
    store8(dc_prev + o1, dcp);
  }
}

// ---- head 1x1 (+bias) to NCHW fp32 and its backward ------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
head_fwd_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                float* __restrict__ out, int N, int P, int C, int K) {
  PCM_PDL_ENTRY();
  extern __shared__ float sw[];   // w[K][C] | b[K]
  for (int i = threadIdx.x; i < K * C; i += blockDim.x) sw[i] = __ldg(w + i);
  for (int i = threadIdx.x; i < K; i += blockDim.x) sw[K * C + i] = __ldg(b + i);
  __syncthreads();
  const long long total = (long long)N * P;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(idx % P);
    const int n = (int)(idx / P);
    for (int k = 0; k < K; ++k) {
      float acc = sw[K * C + k];
      for (int cb = 0; cb < C; cb += 8) {
        float v[8];
        load8(x + idx * C + cb, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc = fmaf(v[j], sw[k * C + cb + j], acc);
      }
      out[((long long)n * K + k) * P + p] = acc;
    }
  }
}

// dx[p][c] = sum_k w[k][c]*dout[n][k][p]; dw[k][c] += sum dout*x; db[k] += sum dout   (K <= 8)
template <typename T>
__global__ void __launch_bounds__(256)
head_bwd_kernel(const float* __restrict__ dout, const T* __restrict__ x, const float* __restrict__ w,
                T* __restrict__ dx, float* __restrict__ dw, float* __restrict__ db, int N, int P, int C, int K) {
  PCM_PDL_ENTRY();
  extern __shared__ float sm[];   // w[K][C] | accw[K][C] | accb[K]
  float* sw = sm;
  float* accw = sm + K * C;
  float* accb = accw + K * C;
  for (int i = threadIdx.x; i < K * C; i += blockDim.x) { sw[i] = __ldg(w + i); accw[i] = 0.f; }
  for (int i = threadIdx.x; i < K; i += blockDim.x) accb[i] = 0.f;
  __syncthreads();
  const long long total = (long long)N * P;
  const long long iters = (total + (long long)gridDim.x * blockDim.x - 1) / ((long long)gridDim.x * blockDim.x);
  for (long long it = 0; it < iters; ++it) {
    const long long idx = (it * gridDim.x + blockIdx.x) * (long long)blockDim.x + threadIdx.x;
    const bool valid = idx < total;
    float d[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) d[k] = 0.f;
    if (valid) {
      const int p = (int)(idx % P);
      const int n = (int)(idx / P);
      for (int k = 0; k < K; ++k) d[k] = __ldg(dout + ((long long)n * K + k) * P + p);
    }
    for (int k = 0; k < K; ++k) {
      const float s = warp_sum(d[k]);
      if ((threadIdx.x & 31) == 0) atomicAdd(&accb[k], s);
    }
    for (int cb = 0; cb < C; cb += 8) {
      float v[8], r[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = r[j] = 0.f;
      if (valid) load8(x + idx * C + cb, v);
      for (int k = 0; k < K; ++k) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          r[j] = fmaf(sw[k * C + cb + j], d[k], r[j]);
          const float s = warp_sum(d[k] * v[j]);
          if ((threadIdx.x & 31) == 0) atomicAdd(&accw[k * C + cb + j], s);
        }
      }
      if (valid) store8(dx + idx * C + cb, r);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < K * C; i += blockDim.x) atomicAdd(dw + i, accw[i]);
  for (int i = threadIdx.x; i < K; i += blockDim.x) atomicAdd(db + i, accb[i]);
}

// ---- head 1x1 + MSE loss fused (src/unet_convlstm_attention.py:104 + main_final.py:559): the training step needs the
// prediction only inside the loss, so forward is ONE pass over the last activation (pred optional) and backward ONE pass
// that recomputes pred_k - y_k per pixel (C*K FMAs) instead of reading a stored gradient: 4 launches -> 2.
//   loss += sum_{n,k,p} (pred - y)^2 / (N*K*P);   dpred = 2*(pred - y)*g/(N*K*P);  dx = W^T dpred;  dw += dpred x^T;  db += dpred
// Per-thread register accumulators for dw / db over the thread's pixels, ONE shuffle + shared reduction per block.
template <typename T, int CV, int KK>
__global__ void __launch_bounds__(256)
head_mse_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                const float* __restrict__ target, const float* __restrict__ gscale, float* __restrict__ pred,
                float* __restrict__ loss, T* __restrict__ dx, float* __restrict__ dw, float* __restrict__ db, int N, int P) {
  PCM_PDL_ENTRY();
  constexpr int C = CV * 8;
  __shared__ float sw[KK * C + KK];
  __shared__ float sacc[KK * C + KK + 1];
  for (int i = threadIdx.x; i < KK * C; i += blockDim.x) sw[i] = __ldg(w + i);
  for (int i = threadIdx.x; i < KK; i += blockDim.x) sw[KK * C + i] = __ldg(b + i);
  for (int i = threadIdx.x; i < KK * C + KK + 1; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const bool bwd = dx != nullptr;
  const long long total = (long long)N * P;
  const float inv_n = 1.f / ((float)total * (float)KK);
  const float gs = bwd ? 2.f * inv_n * __ldg(gscale) : 0.f;
  float aw[KK][C], ab[KK], al = 0.f;
#pragma unroll
  for (int k = 0; k < KK; ++k) {
    ab[k] = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) aw[k][c] = 0.f;
  }
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(idx % P);
    const long long n = idx / P;
    float v[C], d[KK];
#pragma unroll
    for (int cb = 0; cb < CV; ++cb) load8(x + idx * C + cb * 8, v + cb * 8);
#pragma unroll
    for (int k = 0; k < KK; ++k) {
      float acc = sw[KK * C + k];
#pragma unroll
      for (int c = 0; c < C; ++c) acc = fmaf(v[c], sw[k * C + c], acc);
      const long long o = (n * KK + k) * P + p;
      if (pred != nullptr) pred[o] = acc;
      d[k] = acc - __ldg(target + o);
      al = fmaf(d[k], d[k], al);
    }
    if (bwd) {
      float r[C];
#pragma unroll
      for (int c = 0; c < C; ++c) r[c] = 0.f;
#pragma unroll
      for (int k = 0; k < KK; ++k) {
        const float dk = d[k] * gs;
        ab[k] += dk;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          r[c] = fmaf(sw[k * C + c], dk, r[c]);
          aw[k][c] = fmaf(dk, v[c], aw[k][c]);
        }
      }
#pragma unroll
      for (int cb = 0; cb < CV; ++cb) store8(dx + idx * C + cb * 8, r + cb * 8);
    }
  }
  const bool lead = (threadIdx.x & 31) == 0;
  if (bwd) {
#pragma unroll
    for (int k = 0; k < KK; ++k) {
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float t = warp_sum(aw[k][c]);
        if (lead) atomicAdd(&sacc[k * C + c], t);
      }
      const float t = warp_sum(ab[k]);
      if (lead) atomicAdd(&sacc[KK * C + k], t);
    }
  } else {
    const float t = warp_sum(al);
    if (lead) atomicAdd(&sacc[KK * C + KK], t);
  }
  __syncthreads();
  if (bwd) {
    for (int i = threadIdx.x; i < KK * C; i += blockDim.x) atomicAdd(dw + i, sacc[i]);
    for (int i = threadIdx.x; i < KK; i += blockDim.x) atomicAdd(db + i, sacc[KK * C + i]);
  } else if (threadIdx.x == 0) {
    atomicAdd(loss, sacc[KK * C + KK] * inv_n);
  }
}

__global__ void __launch_bounds__(256)
mse_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ loss, long long n) {
  PCM_PDL_ENTRY();
  __shared__ float red[32];
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float d = __ldg(a + i) - __ldg(b + i);
    acc = fmaf(d, d, acc);
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(loss, acc / (float)n);
}

__global__ void __launch_bounds__(256)
mse_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ gscale,
               float* __restrict__ da, long long n) {
  PCM_PDL_ENTRY();
  const float k = 2.f / (float)n * (gscale ? __ldg(gscale) : 1.f);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    da[i] = k * (__ldg(a + i) - __ldg(b + i));
}

// ---- Adam (torch.optim.Adam semantics, main_final.py:742-746) -----------------------------------------
__global__ void adam_tick_kernel(float* state, float b1, float b2) {
  PCM_PDL_ENTRY();
  // state: [0]=step, [1]=1-b1^step, [2]=1-b2^step
  const float step = state[0] + 1.f;
  state[0] = step;
  state[1] = 1.f - powf(b1, step);
  state[2] = 1.f - powf(b2, step);
}

__global__ void __launch_bounds__(256)
adam_apply_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                  const float* __restrict__ state, long long n, float lr, float b1, float b2, float eps, float wd,
                  float grad_scale) {
  PCM_PDL_ENTRY();
  const float bc1 = state[1], bc2s = sqrtf(state[2]);
  const float step_size = lr / bc1;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float gi = g[i] * grad_scale;
    const float pi = p[i];
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    const float mi = fmaf(b1, m[i], (1.f - b1) * gi);
    const float vi = fmaf(b2, v[i], (1.f - b2) * gi * gi);
    m[i] = mi;
    v[i] = vi;
    p[i] = pi - step_size * mi / (sqrtf(vi) / bc2s + eps);
  }
}

static inline int grid_for(long long total) {
  long long b = (total + 255) / 256;
  if (b > kGridCap) b = kGridCap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace pcm

using namespace pcm;

extern "C" int pcm_nchw_to_nhwc(const float* x, void* y, int N, int C, int H, int W, int Cp, int T_, int dtype,
                                pcm_stream_t s) {
  PCM_REQUIRE(Cp % 8 == 0 && Cp >= C, "nchw_to_nhwc: bad channel padding C=%d Cp=%d", C, Cp);
  PCM_REQUIRE(T_ >= 1 && N % T_ == 0, "nchw_to_nhwc: N must be a multiple of T");
  if (N == 0) return PCM_OK;
  const long long total = (long long)N * (Cp / 8) * H * W;
  if ((H * W) % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    PCM_DISPATCH_DTYPE(dtype, T, (pcm::launch(nchw_to_nhwc_x4_kernel<T>, grid_for(total / 4), 256, 0, (cudaStream_t)s, 
                                     x, (T*)y, N, C, H * W, Cp, nullptr, T_)));
  } else {
    PCM_DISPATCH_DTYPE(dtype, T, (pcm::launch(nchw_to_nhwc_kernel<T>, grid_for(total), 256, 0, (cudaStream_t)s, 
                                     x, (T*)y, N, C, H * W, Cp, nullptr, T_)));
  }
  return check_launch("nchw_to_nhwc");
}

extern "C" int pcm_window_stage(const float* series, const int* frames, void* y, int N, int C, int H, int W, int Cp,
                                const double* norm, const int* month, int dtype, pcm_stream_t s) {
  PCM_REQUIRE(Cp % 8 == 0 && Cp >= C + (month ? 2 : 0), "window_stage: bad channel padding C=%d Cp=%d", C, Cp);
  PCM_REQUIRE((H * W) % 4 == 0 && (reinterpret_cast<uintptr_t>(series) & 15) == 0, "window_stage: H*W must be a multiple of 4");
  if (N == 0) return PCM_OK;
  const long long total = (long long)N * (Cp / 8) * H * W / 4;
  PCM_DISPATCH_DTYPE(dtype, T, (pcm::launch(window_stage_kernel<T>, grid_for(total), 256, 0, (cudaStream_t)s,
                                   series, frames, (T*)y, N, C, H * W, Cp, norm, month)));
  return check_launch("window_stage");
}

extern "C" int pcm_season_embed_stage(const float* x5, const int* month, void* y, int N, int H, int W, int Cp,
                                      int T_, int dtype, pcm_stream_t s) {
  PCM_REQUIRE(Cp % 8 == 0 && Cp >= 7, "season_embed_stage: Cp must be >= 7 and a multiple of 8");
  PCM_REQUIRE(T_ >= 1 && N % T_ == 0, "season_embed_stage: N must be a multiple of T");
  if (N == 0) return PCM_OK;
  const long long total = (long long)N * (Cp / 8) * H * W;
  if ((H * W) % 4 == 0 && (reinterpret_cast<uintptr_t>(x5) & 15) == 0) {
    PCM_DISPATCH_DTYPE(dtype, T, (pcm::launch(nchw_to_nhwc_x4_kernel<T>, grid_for(total / 4), 256, 0, (cudaStream_t)s, 
                                     x5, (T*)y, N, 5, H * W, Cp, month, T_)));
  } else {
    PCM_DISPATCH_DTYPE(dtype, T, (pcm::launch(nchw_to_nhwc_kernel<T>, grid_for(total), 256, 0, (cudaStream_t)s, 
                                     x5, (T*)y, N, 5, H * W, Cp, month, T_)));
  }
  return check_launch("season_embed_stage");
}

extern "C" int pcm_nhwc_to_nchw(const void* x, float* y, int N, int C, int H, int W, int Cp, int T_, int dtype,
                                pcm_stream_t s) {
  PCM_REQUIRE(Cp % 8 == 0 && Cp >= C, "nhwc_to_nchw: bad channel padding C=%d Cp=%d", C, Cp);
  PCM_REQUIRE(T_ >= 1 && N % T_ == 0, "nhwc_to_nchw: N must be a multiple of T");
  if (N == 0) return PCM_OK;
  const long long total = (long long)N * (Cp / 8) * H * W;
  PCM_DISPATCH_DTYPE(dtype, T, (pcm::launch(nhwc_to_nchw_kernel<T>, grid_for(total), 256, 0, (cudaStream_t)s, 
                                   (const T*)x, y, N, C, H * W, Cp, T_)));
  return check_launch("nhwc_to_nchw");
}

extern "C" int pcm_maxpool2_fwd(const void* x, void* y, int N, int H, int W, int C, int dtype, pcm_stream_t s) {
  PCM_REQUIRE(C % 8 == 0 && H >= 2 && W >= 2, "maxpool2_fwd: bad shape");
  if (N == 0) return PCM_OK;
  const long long total = (long long)N * (H / 2) * (W / 2) * (C / 8);
  if (total + (long long)kGridCap * 256 < 0x7fffffffLL) {
    PCM_DISPATCH_DTYPE(dtype, T, (pcm::launch(maxpool2_fwd_kernel<T, int>, grid_for(total), 256, 0, (cudaStream_t)s,
                                     (const T*)x, (T*)y, N, H, W, C)));
  } else {
    PCM_DISPATCH_DTYPE(dtype, T, (pcm::launch(maxpool2_fwd_kernel<T, long long>, grid_for(total), 256, 0, (cudaStream_t)s,
                                     (const T*)x, (T*)y, N, H, W, C)));
  }
  return check_launch("maxpool2_fwd");
}

extern "C" int pcm_maxpool2_bwd_skip_dot(const void* x, const void* dy, const void* dskip, long long dskip_ns,
                                         int dskip_ps, void* dx, float* sdot, int N, int H, int W, int C, int T_,
                                         int t_major, int dtype, pcm_stream_t s);
extern "C" int pcm_maxpool2_bwd_skip(const void* x, const void* dy, const void* dskip, long long dskip_ns,
                                     int dskip_ps, void* dx, int N, int H, int W, int C, int T_, int t_major,
                                     int dtype, pcm_stream_t s) {
  return pcm_maxpool2_bwd_skip_dot(x, dy, dskip, dskip_ns, dskip_ps, dx, nullptr, N, H, W, C, T_, t_major, dtype, s);
}

extern "C" int pcm_maxpool2_bwd_skip_dot(const void* x, const void* dy, const void* dskip, long long dskip_ns,
                                         int dskip_ps, void* dx, float* sdot, int N, int H, int W, int C, int T_,
                                         int t_major, int dtype, pcm_stream_t s) {
  PCM_REQUIRE(C % 8 == 0 && T_ >= 1 && N % T_ == 0, "maxpool2_bwd_skip: bad shape");
  if (sdot != nullptr) {
    const int cv = C / 8;
    PCM_REQUIRE(cv <= 32 && (cv & (cv - 1)) == 0, "maxpool2_bwd_skip_dot: C/8 must be a power of two <= 32 (C=%d)", C);
  }
  if (N == 0) return PCM_OK;
  const long long total = (long long)N * ((H + 1) / 2) * ((W + 1) / 2) * (C / 8);
  if (total + (long long)kGridCap * 256 < 0x7fffffffLL) {
    PCM_DISPATCH_DTYPE(dtype, T, {
      auto kern = sdot != nullptr ? maxpool2_bwd_skip_kernel<T, int, true> : maxpool2_bwd_skip_kernel<T, int, false>;
      pcm::launch(kern, grid_for(total), 256, 0, (cudaStream_t)s, (const T*)x, (const T*)dy, (const T*)dskip, dskip_ns, dskip_ps,
                  (T*)dx, sdot, N, H, W, C, T_, t_major);
    });
  } else {
    PCM_DISPATCH_DTYPE(dtype, T, {
      auto kern = sdot != nullptr ? maxpool2_bwd_skip_kernel<T, long long, true> : maxpool2_bwd_skip_kernel<T, long long, false>;
      pcm::launch(kern, grid_for(total), 256, 0, (cudaStream_t)s, (const T*)x, (const T*)dy, (const T*)dskip, dskip_ns, dskip_ps,
                  (T*)dx, sdot, N, H, W, C, T_, t_major);
    });
  }
  return check_launch("maxpool2_bwd_skip");
}

extern "C" int pcm_time_mean(const void* src, void* dst, long long dst_ns, int dst_ps, int B, int T_, int P, int C,
                             int t_major, int dtype, pcm_stream_t s) {
  PCM_REQUIRE(C % 8 == 0 && T_ >= 1, "time_mean: bad shape");
  if (B == 0) return PCM_OK;
  const long long total = (long long)B * P * (C / 8);
  if (total + (long long)kGridCap * 256 < 0x7fffffffLL) {
    PCM_DISPATCH_DTYPE(dtype, T, (pcm::launch(time_mean_kernel<T, int>, grid_for(total), 256, 0, (cudaStream_t)s,
                                     (const T*)src, (T*)dst, dst_ns, dst_ps, B, T_, P, C, t_major)));
  } else {
    PCM_DISPATCH_DTYPE(dtype, T, (pcm::launch(time_mean_kernel<T, long long>, grid_for(total), 256, 0, (cudaStream_t)s,
                                     (const T*)src, (T*)dst, dst_ns, dst_ps, B, T_, P, C, t_major)));
  }
  return check_launch("time_mean");
}

extern "C" int pcm_lstm_cell_fwd(const float* gates, const float* c_prev, void* acts, float* c, void* h, int M,
                                 int Ch, int dtype, pcm_stream_t s) {
  PCM_REQUIRE(Ch % 8 == 0, "lstm_cell_fwd: Ch must be a multiple of 8");
  if (M == 0) return PCM_OK;
  const long long total = (long long)M * (Ch / 8);
  PCM_DISPATCH_DTYPE(dtype, T, (pcm::launch(lstm_cell_fwd_kernel<T>, grid_for(total), 256, 0, (cudaStream_t)s, 
                                   gates, c_prev, (T*)acts, c, (T*)h, M, Ch)));
  return check_launch("lstm_cell_fwd");
}

extern "C" int pcm_lstm_cell_bwd(const void* dh_a, const void* dh_b, const float* dc_in, const void* acts,
                                 const float* c_prev, const float* c, void* dgates, float* dc_prev, int M, int Ch,
                                 int dtype, pcm_stream_t s) {
  PCM_REQUIRE(Ch % 8 == 0, "lstm_cell_bwd: Ch must be a multiple of 8");
  if (M == 0) return PCM_OK;
  const long long total = (long long)M * (Ch / 8);
  PCM_DISPATCH_DTYPE(dtype, T, (pcm::launch(lstm_cell_bwd_kernel<T>, grid_for(total), 256, 0, (cudaStream_t)s, 
                                   (const T*)dh_a, (const T*)dh_b, dc_in, (const T*)acts, c_prev, c, (T*)dgates,
                                   dc_prev, M, Ch)));
  return check_launch("lstm_cell_bwd");
}

extern "C" int pcm_head_fwd(const void* x, const float* w, const float* b, float* out_nchw, int N, int P, int C, int K,
                            int dtype, pcm_stream_t s) {
  PCM_REQUIRE(C % 8 == 0 && K >= 1, "head_fwd: bad shape");
  if (N == 0) return PCM_OK;
  const size_t smem = (size_t)(K * C + K) * sizeof(float);
  PCM_DISPATCH_DTYPE(dtype, T, (pcm::launch(head_fwd_kernel<T>, grid_for((long long)N * P), 256, smem, (cudaStream_t)s, 
                                   (const T*)x, w, b, out_nchw, N, P, C, K)));
  return check_launch("head_fwd");
}

extern "C" int pcm_head_bwd(const float* dout_nchw, const void* x, const float* w, void* dx, float* dw, float* db,
                            int N, int P, int C, int K, int dtype, pcm_stream_t s) {
  PCM_REQUIRE(C % 8 == 0 && K >= 1 && K <= 8, "head_bwd: K must be in 1..8");
  if (N == 0) return PCM_OK;
  const size_t smem = (size_t)(2 * K * C + K) * sizeof(float);
  long long blocks = ((long long)N * P + 255) / 256;
  if (blocks > 592) blocks = 592;
  PCM_DISPATCH_DTYPE(dtype, T, (pcm::launch(head_bwd_kernel<T>, (int)blocks, 256, smem, (cudaStream_t)s, 
                                   dout_nchw, (const T*)x, w, (T*)dx, dw, db, N, P, C, K)));
  return check_launch("head_bwd");
}

template <typename T, int CV, int KK>
static void head_mse_launch(const void* x, const float* w, const float* b, const float* target, const float* gscale, float* pred,
                            float* loss, void* dx, float* dw, float* db, int N, int P, pcm_stream_t s) {
  long long blocks = ((long long)N * P + 255) / 256;
  if (blocks > 296) blocks = 296;
  pcm::launch(head_mse_kernel<T, CV, KK>, (int)blocks, 256, 0, (cudaStream_t)s, (const T*)x, w, b, target, gscale, pred, loss,
              (T*)dx, dw, db, N, P);
}

extern "C" int pcm_head_mse_supported(int C, int K) { return ((C == 16 || C == 32) && K == 2) ? 1 : 0; }

static int head_mse_dispatch(const void* x, const float* w, const float* b, const float* target, const float* gscale, float* pred,
                             float* loss, void* dx, float* dw, float* db, int N, int P, int C, int K, int dtype, pcm_stream_t s) {
  PCM_REQUIRE(pcm_head_mse_supported(C, K), "head_mse: C must be 16 or 32 and K 2 (got C=%d K=%d)", C, K);
  if (N == 0) return PCM_OK;
  PCM_DISPATCH_DTYPE(dtype, T, {
    if (C == 16) head_mse_launch<T, 2, 2>(x, w, b, target, gscale, pred, loss, dx, dw, db, N, P, s);
    else head_mse_launch<T, 4, 2>(x, w, b, target, gscale, pred, loss, dx, dw, db, N, P, s);
  });
  return check_launch("head_mse");
}

extern "C" int pcm_head_mse_fwd(const void* x, const float* w, const float* b, const float* target, float* pred_nchw,
                                float* loss, int N, int P, int C, int K, int dtype, pcm_stream_t s) {
  return head_mse_dispatch(x, w, b, target, nullptr, pred_nchw, loss, nullptr, nullptr, nullptr, N, P, C, K, dtype, s);
}

extern "C" int pcm_head_mse_bwd(const void* x, const float* w, const float* b, const float* target, const float* gscale,
                                void* dx, float* dw, float* db, int N, int P, int C, int K, int dtype, pcm_stream_t s) {
  PCM_REQUIRE(dx != nullptr && dw != nullptr && db != nullptr && gscale != nullptr, "head_mse_bwd: null output");
  return head_mse_dispatch(x, w, b, target, gscale, nullptr, nullptr, dx, dw, db, N, P, C, K, dtype, s);
}

extern "C" int pcm_mse_fwd(const float* a, const float* b, float* loss, long long n, pcm_stream_t s) {
  if (n == 0) return PCM_OK;
  long long blocks = (n + 1023) / 1024;
  if (blocks > 296) blocks = 296;
  pcm::launch(mse_fwd_kernel, (int)blocks, 256, 0, (cudaStream_t)s, a, b, loss, n);
  return check_launch("mse_fwd");
}

extern "C" int pcm_mse_bwd(const float* a, const float* b, const float* gscale, float* da, long long n,
                           pcm_stream_t s) {
  if (n == 0) return PCM_OK;
  pcm::launch(mse_bwd_kernel, grid_for(n), 256, 0, (cudaStream_t)s, a, b, gscale, da, n);
  return check_launch("mse_bwd");
}

extern "C" int pcm_adam_step(float* p, const float* g, float* m, float* v, float* state, long long n, float lr,
                             float b1, float b2, float eps, float wd, float grad_scale, pcm_stream_t s) {
  pcm::launch(adam_tick_kernel, 1, 1, 0, (cudaStream_t)s, state, b1, b2);
  if (n > 0)
    pcm::launch(adam_apply_kernel, grid_for(n), 256, 0, (cudaStream_t)s, p, g, m, v, state, n, lr, b1, b2, eps, wd, grad_scale);
  return check_launch("adam_step");
}

// The same two kernels separately: under data parallelism the optimizer runs bucket by bucket as each gradient bucket's
// all-reduce completes (trainer.py) — ONE tick per step, then one apply per parameter range.
extern "C" int pcm_adam_tick(float* state, float b1, float b2, pcm_stream_t s) {
  pcm::launch(adam_tick_kernel, 1, 1, 0, (cudaStream_t)s, state, b1, b2);
  return check_launch("adam_tick");
}

extern "C" int pcm_adam_apply(float* p, const float* g, float* m, float* v, const float* state, long long n, float lr,
                              float b1, float b2, float eps, float wd, float grad_scale, pcm_stream_t s) {
  if (n > 0)
    pcm::launch(adam_apply_kernel, grid_for(n), 256, 0, (cudaStream_t)s, p, g, m, v, state, n, lr, b1, b2, eps, wd, grad_scale);
  return check_launch("adam_apply");
}
