// BatchNorm2d in training mode (+ReLU, +residual) on NHWC rows, channel dropout masks and small elementwise helpers
// for the SimpleCNN / ResidualBlock path (reference src/models.py:44-123).  All HBM-bound: every thread owns a fixed
// 8-channel vector column (so per-channel scale/shift live in registers) and strides over pixel rows.
#include "common.cuh"

namespace pcm {

constexpr int kBnThreads = 256;
constexpr int kBnGridCap = 148 * 8;

__device__ __forceinline__ float hash_uniform(unsigned long long seed, unsigned long long idx) {
  return dropout_uniform(seed, idx);      // common.cuh: one definition for every mask-drawing kernel
}

// sums[c][0] += sum_r x(r,c); sums[c][1] += sum_r x(r,c)^2          (caller zeroes sums)
// The per-thread partials (a few rows each) are fp32; they are combined across threads and blocks in DOUBLE: the
// order of those atomics varies from run to run, and in fp32 that moved mean / variance in the last bit — enough to
// flip the ReLU that follows for pre-activations within ~1e-6 of zero and change every gradient upstream by ~1e-3
// (seen as a 1-in-15 flake of the fp32 parity test).  In double the order-dependence is ~1e-16, and E[x^2] - mean^2
// loses nothing to cancellation.
template <typename T>
__global__ void __launch_bounds__(kBnThreads)
bn_stats_kernel(const T* __restrict__ x, double* __restrict__ sums, long long R, int C) {
  PCM_PDL_ENTRY();
  extern __shared__ double shd[];                // cv <= 32: [warps][C][2] per-warp slots; else [C][2] accumulators
  const int cv = C / 8, rpb = kBnThreads / cv;
  const int cb = threadIdx.x % cv, rl = threadIdx.x / cv;
  constexpr int kWarps = kBnThreads / 32;
  if (cv > 32) {
    for (int i = threadIdx.x; i < 2 * C; i += kBnThreads) shd[i] = 0.0;
    __syncthreads();
  }
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  if (rl < rpb) {
    for (long long r = (long long)blockIdx.x * rpb + rl; r < R; r += (long long)gridDim.x * rpb) {
      float v[8];
      load8(x + r * C + cb * 8, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) { s[j] += v[j]; q[j] = fmaf(v[j], v[j], q[j]); }
    }
  }
  // lanes of a warp that own the same channel block (cv < 32: lane stride cv) are combined by a fixed shuffle tree
  // first, so that only cv lanes per warp touch the shared accumulators (shared-memory double atomics are CAS loops)
  for (int off = cv; off < 32; off <<= 1) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j] += __shfl_xor_sync(0xffffffffu, s[j], off);
      q[j] += __shfl_xor_sync(0xffffffffu, q[j], off);
    }
  }
  if (cv <= 32) {
    // every warp holds all cv channel blocks: lanes < cv park the warp's totals in the warp's own slot, then 2*C
    // threads add the slots in a fixed order (no shared-memory atomics at all: they are CAS loops for double)
    if ((int)(threadIdx.x & 31) < cv) {
      double* slot = shd + (size_t)(threadIdx.x >> 5) * 2 * C;
#pragma unroll
      for (int j = 0; j < 8; ++j) { slot[(cb * 8 + j) * 2] = (double)s[j]; slot[(cb * 8 + j) * 2 + 1] = (double)q[j]; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += kBnThreads) {
      double t = 0.0;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) t += shd[(size_t)w * 2 * C + i];
      atomicAdd(sums + i, t);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&shd[(cb * 8 + j) * 2], (double)s[j]);
      atomicAdd(&shd[(cb * 8 + j) * 2 + 1], (double)q[j]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += kBnThreads) atomicAdd(sums + i, shd[i]);
  }
}

// (mean, biased variance) of channel c from the double sums
__device__ __forceinline__ void bn_mean_var(const double* sums, int c, double invR, float& mean, float& var) {
  const double m = sums[2 * c] * invR;
  mean = (float)m;
  var = (float)fmax(sums[2 * c + 1] * invR - m * m, 0.0);
}

__device__ __forceinline__ void bn_coeffs(const double* sums, const float* gamma, const float* beta, int c0, double invR,
                                          float eps, float mean[8], float rstd[8], float a[8], float b[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float m, var;
    bn_mean_var(sums, c0 + j, invR, m, var);
    mean[j] = m;
    rstd[j] = rsqrtf(var + eps);
    a[j] = gamma[c0 + j] * rstd[j];
    b[j] = beta[c0 + j] - m * a[j];
  }
}

// y = [relu]( (x-mean)*rstd*gamma + beta [+ res] )
template <typename T>
__global__ void __launch_bounds__(kBnThreads)
bn_apply_fwd_kernel(const T* __restrict__ x, const double* __restrict__ sums, const float* __restrict__ gamma,
                    const float* __restrict__ beta, const T* __restrict__ res, T* __restrict__ y, long long R, double invRd, int C,
                    float eps, int relu) {
  PCM_PDL_ENTRY();
  const int cv = C / 8, rpb = kBnThreads / cv;
  const int cb = threadIdx.x % cv, rl = threadIdx.x / cv;
  if (rl >= rpb) return;
  float mean[8], rstd[8], a[8], b[8];
  bn_coeffs(sums, gamma, beta, cb * 8, invRd, eps, mean, rstd, a, b);
  for (long long r = (long long)blockIdx.x * rpb + rl; r < R; r += (long long)gridDim.x * rpb) {
    float v[8];
    load8(x + r * C + cb * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], a[j], b[j]);
    if (res != nullptr) {
      float t[8];
      load8(res + r * C + cb * 8, t);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += t[j];
    }
    if (relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    store8(y + r * C + cb * 8, v);
  }
}

// dsum[c][0] += sum dz ; dsum[c][1] += sum dz*xhat, with dz = dy * (y > 0 when y != null)
template <typename T>
__global__ void __launch_bounds__(kBnThreads)
bn_bwd_reduce_kernel(const T* __restrict__ dy, const T* __restrict__ y, const T* __restrict__ x,
                     const double* __restrict__ sums, float* __restrict__ dsum, long long R, double invRd, int C, float eps) {
  PCM_PDL_ENTRY();
  extern __shared__ float sh[];                  // cv <= 32: [warps][C][2] per-warp slots; else [C][2] accumulators
  const int cv = C / 8, rpb = kBnThreads / cv;
  const int cb = threadIdx.x % cv, rl = threadIdx.x / cv;
  constexpr int kWarps = kBnThreads / 32;
  if (cv > 32) {
    for (int i = threadIdx.x; i < 2 * C; i += kBnThreads) sh[i] = 0.f;
    __syncthreads();
  }
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  if (rl < rpb) {
    float mean[8], rstd[8];
    const float invR = 1.f / (float)R;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float m, var;
      bn_mean_var(sums, cb * 8 + j, invRd, m, var);
      mean[j] = m; rstd[j] = rsqrtf(var + eps);
    }
    for (long long r = (long long)blockIdx.x * rpb + rl; r < R; r += (long long)gridDim.x * rpb) {
      float g[8], v[8];
      load8(dy + r * C + cb * 8, g);
      load8(x + r * C + cb * 8, v);
      if (y != nullptr) {
        float o[8];
        load8(y + r * C + cb * 8, o);
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = o[j] > 0.f ? g[j] : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) { s[j] += g[j]; q[j] = fmaf(g[j], (v[j] - mean[j]) * rstd[j], q[j]); }
    }
  }
  for (int off = cv; off < 32; off <<= 1) {          // same-channel lanes of the warp first (see bn_stats_kernel)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j] += __shfl_xor_sync(0xffffffffu, s[j], off);
      q[j] += __shfl_xor_sync(0xffffffffu, q[j], off);
    }
  }
  if (cv <= 32) {
    if ((int)(threadIdx.x & 31) < cv) {
      float* slot = sh + (size_t)(threadIdx.x >> 5) * 2 * C;
#pragma unroll
      for (int j = 0; j < 8; ++j) { slot[(cb * 8 + j) * 2] = s[j]; slot[(cb * 8 + j) * 2 + 1] = q[j]; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += kBnThreads) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) t += sh[(size_t)w * 2 * C + i];
      atomicAdd(dsum + i, t);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&sh[(cb * 8 + j) * 2], s[j]);
      atomicAdd(&sh[(cb * 8 + j) * 2 + 1], q[j]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += kBnThreads) atomicAdd(dsum + i, sh[i]);
  }
}

// dx = gamma*rstd*(dz - mean(dz) - xhat*mean(dz*xhat)); dres = dz (nullable); block 0 adds dgamma/dbeta
template <typename T>
__global__ void __launch_bounds__(kBnThreads)
bn_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ y, const T* __restrict__ x,
                    const double* __restrict__ sums, const float* __restrict__ gamma, const float* __restrict__ dsum,
                    T* __restrict__ dx, T* __restrict__ dres, float* __restrict__ dgamma, float* __restrict__ dbeta,
                    long long R, double invRd, int C, float eps) {
  PCM_PDL_ENTRY();
  const int cv = C / 8, rpb = kBnThreads / cv;
  const int cb = threadIdx.x % cv, rl = threadIdx.x / cv;
  if (blockIdx.x == 0) {
    for (int c = threadIdx.x; c < C; c += kBnThreads) {
      atomicAdd(dbeta + c, dsum[2 * c]);
      atomicAdd(dgamma + c, dsum[2 * c + 1]);
    }
  }
  if (rl >= rpb) return;
  const float invR = 1.f / (float)R;
  float mean[8], rstd[8], a[8], m1[8], m2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cb * 8 + j;
    float m, var;
    bn_mean_var(sums, c, invRd, m, var);
    mean[j] = m; rstd[j] = rsqrtf(var + eps);
    a[j] = gamma[c] * rstd[j];
    m1[j] = dsum[2 * c] * invR; m2[j] = dsum[2 * c + 1] * invR;
  }
  for (long long r = (long long)blockIdx.x * rpb + rl; r < R; r += (long long)gridDim.x * rpb) {
    float g[8], v[8];
    load8(dy + r * C + cb * 8, g);
    load8(x + r * C + cb * 8, v);
    if (y != nullptr) {
      float o[8];
      load8(y + r * C + cb * 8, o);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] = o[j] > 0.f ? g[j] : 0.f;
    }
    if (dres != nullptr) store8(dres + r * C + cb * 8, g);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = a[j] * (g[j] - m1[j] - (v[j] - mean[j]) * rstd[j] * m2[j]);
    store8(dx + r * C + cb * 8, v);
  }
}

__global__ void bn_update_running_kernel(const double* __restrict__ sums, float* __restrict__ rm, float* __restrict__ rv,
                                         long long* __restrict__ nbt, long long R, double invRd, int C, float momentum) {
  PCM_PDL_ENTRY();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && nbt != nullptr) *nbt += 1;
  if (c >= C) return;
  float m, var;
  bn_mean_var(sums, c, invRd, m, var);
  const float unb = R > 1 ? var * (float)R / (float)(R - 1) : var;
  rm[c] = (1.f - momentum) * rm[c] + momentum * m;
  rv[c] = (1.f - momentum) * rv[c] + momentum * unb;
}

// ---- small elementwise helpers -----------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) add_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ o,
                                                   long long n8) {
  PCM_PDL_ENTRY();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float u[8], v[8];
    load8(a + i * 8, u); load8(b + i * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) u[j] += v[j];
    store8(o + i * 8, u);
  }
}

// y[i] = x[i] + b[i % Rn]  (b fp32; positional embedding / broadcast bias)
template <typename T>
__global__ void __launch_bounds__(256) add_bcast_kernel(const T* __restrict__ x, const float* __restrict__ b,
                                                         T* __restrict__ y, long long n8, long long R8) {
  PCM_PDL_ENTRY();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float u[8], v[8];
    load8(x + i * 8, u);
    load8(b + (i % R8) * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) u[j] += v[j];
    store8(y + i * 8, u);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) relu_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ y, T* __restrict__ dx,
                                                        long long n8) {
  PCM_PDL_ENTRY();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float g[8], o[8];
    load8(dy + i * 8, g); load8(y + i * 8, o);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = o[j] > 0.f ? g[j] : 0.f;
    store8(dx + i * 8, g);
  }
}

// dx = y > 0 ? dy * scale : 0 — backward of ReLU followed by dropout from the saved OUTPUT alone: y > 0 exactly where the
// pre-activation was positive and the element was kept, and the kept gradient is scaled by 1 / (1 - p)
template <typename T>
__global__ void __launch_bounds__(256) relu_bwd_scaled_kernel(const T* __restrict__ dy, const T* __restrict__ y,
                                                               T* __restrict__ dx, long long n8, float scale) {
  PCM_PDL_ENTRY();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float g[8], o[8];
    load8(dy + i * 8, g); load8(y + i * 8, o);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = o[j] > 0.f ? g[j] * scale : 0.f;
    store8(dx + i * 8, g);
  }
}

// y = x * keep(seed, i) / (1 - p): counter-based mask, so applying the same call to dy is the backward
template <typename T>
__global__ void __launch_bounds__(256) dropout_kernel(const T* __restrict__ x, T* __restrict__ y, long long n8, float p,
                                                       unsigned long long seed,
                                                       const unsigned long long* __restrict__ epoch) {
  PCM_PDL_ENTRY();
  seed = mix_epoch(seed, epoch);
  const float sc = 1.f / (1.f - p);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    float u[8];
    load8(x + i * 8, u);
#pragma unroll
    for (int j = 0; j < 8; ++j) u[j] = hash_uniform(seed, (unsigned long long)(i * 8 + j)) >= p ? u[j] * sc : 0.f;
    store8(y + i * 8, u);
  }
}

__global__ void dropout_mask_kernel(float* __restrict__ mask, long long n, float p, unsigned long long seed,
                                    const unsigned long long* __restrict__ epoch) {
  PCM_PDL_ENTRY();
  seed = mix_epoch(seed, epoch);
  const float sc = 1.f / (1.f - p);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    mask[i] = hash_uniform(seed, (unsigned long long)i) >= p ? sc : 0.f;
}

// out[r] += sum_b x[b][r]   (gradient of a parameter broadcast over the batch, e.g. pos_embedding)
template <typename T>
__global__ void __launch_bounds__(128) batch_sum_kernel(const T* __restrict__ x, float* __restrict__ out, int B,
                                                         long long R8) {
  PCM_PDL_ENTRY();
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= R8) return;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int b = 0; b < B; ++b) {
    float v[8];
    load8(x + ((long long)b * R8 + i) * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += v[j];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) out[i * 8 + j] += acc[j];
}

static inline int bn_grid(long long R, int C) {
  const int rpb = kBnThreads / (C / 8);
  long long g = (R + (long long)rpb * 4 - 1) / ((long long)rpb * 4);
  if (g > kBnGridCap) g = kBnGridCap;
  if (g < 1) g = 1;
  return (int)g;
}
static inline int ew_grid(long long n8) {
  long long g = (n8 + 255) / 256;
  if (g > kBnGridCap) g = kBnGridCap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace pcm

using namespace pcm;

#define BN_CHECK_C(fn, C)                                                                                         \
  PCM_REQUIRE((C) % 8 == 0 && (C) >= 8 && (C) <= 2048 && kBnThreads % ((C) / 8) == 0,                            \
              fn ": C must be 8 * (a divisor of 256) (got %d)", (int)(C))

extern "C" int pcm_bn_stats(const void* x, double* sums, long long R, int C, int dtype, pcm_stream_t s) {
  BN_CHECK_C("bn_stats", C);
  if (R == 0) return PCM_OK;
  PCM_DISPATCH_DTYPE(dtype, T, (pcm::launch(bn_stats_kernel<T>, bn_grid(R, C), kBnThreads, (C <= 256 ? kBnThreads / 32 : 1) * 2 * C * sizeof(double), (cudaStream_t)s,
                                   static_cast<const T*>(x), sums, R, C)));
  return check_launch("bn_stats");
}

extern "C" int pcm_bn_apply_fwd(const void* x, const double* sums, const float* gamma, const float* beta, const void* res,
                                void* y, long long R, int C, float eps, int relu, int dtype, pcm_stream_t s) {
  BN_CHECK_C("bn_apply_fwd", C);
  if (R == 0) return PCM_OK;
  PCM_DISPATCH_DTYPE(dtype, T, (pcm::launch(bn_apply_fwd_kernel<T>, bn_grid(R, C), kBnThreads, 0, (cudaStream_t)s, 
                                   static_cast<const T*>(x), sums, gamma, beta, static_cast<const T*>(res),
                                   static_cast<T*>(y), R, 1.0 / (double)R, C, eps, relu)));
  return check_launch("bn_apply_fwd");
}

extern "C" int pcm_bn_update_running(const double* sums, float* running_mean, float* running_var,
                                     long long* num_batches_tracked, long long R, int C, float momentum, pcm_stream_t s) {
  pcm::launch(bn_update_running_kernel, ceil_div(C, 128), 128, 0, (cudaStream_t)s, sums, running_mean, running_var,
                                                                          num_batches_tracked, R, 1.0 / (double)R, C, momentum);
  return check_launch("bn_update_running");
}

extern "C" int pcm_bn_bwd_reduce(const void* dy, const void* y, const void* x, const double* sums, float* dsum, long long R,
                                 int C, float eps, int dtype, pcm_stream_t s) {
  BN_CHECK_C("bn_bwd_reduce", C);
  if (R == 0) return PCM_OK;
  PCM_DISPATCH_DTYPE(dtype, T, (pcm::launch(bn_bwd_reduce_kernel<T>, bn_grid(R, C), kBnThreads, (C <= 256 ? kBnThreads / 32 : 1) * 2 * C * sizeof(float), (cudaStream_t)s, 
                                   static_cast<const T*>(dy), static_cast<const T*>(y), static_cast<const T*>(x), sums, dsum,
                                   R, 1.0 / (double)R, C, eps)));
  return check_launch("bn_bwd_reduce");
}

extern "C" int pcm_bn_bwd_apply(const void* dy, const void* y, const void* x, const double* sums, const float* gamma,
                                const float* dsum, void* dx, void* dres, float* dgamma, float* dbeta, long long R, int C,
                                float eps, int dtype, pcm_stream_t s) {
  BN_CHECK_C("bn_bwd_apply", C);
  if (R == 0) return PCM_OK;
  PCM_DISPATCH_DTYPE(dtype, T, (pcm::launch(bn_bwd_apply_kernel<T>, bn_grid(R, C), kBnThreads, 0, (cudaStream_t)s, 
                                   static_cast<const T*>(dy), static_cast<const T*>(y), static_cast<const T*>(x), sums, gamma,
                                   dsum, static_cast<T*>(dx), static_cast<T*>(dres), dgamma, dbeta, R, 1.0 / (double)R, C, eps)));
  return check_launch("bn_bwd_apply");
}

extern "C" int pcm_add(const void* a, const void* b, void* out, long long n, int dtype, pcm_stream_t s) {
  PCM_REQUIRE(n % 8 == 0, "add: n must be a multiple of 8");
  if (n == 0) return PCM_OK;
  PCM_DISPATCH_DTYPE(dtype, T, (pcm::launch(add_kernel<T>, ew_grid(n / 8), 256, 0, (cudaStream_t)s, 
                                   static_cast<const T*>(a), static_cast<const T*>(b), static_cast<T*>(out), n / 8)));
  return check_launch("add");
}

extern "C" int pcm_add_bcast(const void* x, const float* b, void* y, long long n, long long period, int dtype,
                             pcm_stream_t s) {
  PCM_REQUIRE(n % 8 == 0 && period % 8 == 0 && period > 0, "add_bcast: n and period must be multiples of 8");
  if (n == 0) return PCM_OK;
  PCM_DISPATCH_DTYPE(dtype, T, (pcm::launch(add_bcast_kernel<T>, ew_grid(n / 8), 256, 0, (cudaStream_t)s, 
                                   static_cast<const T*>(x), b, static_cast<T*>(y), n / 8, period / 8)));
  return check_launch("add_bcast");
}

extern "C" int pcm_relu_bwd(const void* dy, const void* y, void* dx, long long n, int dtype, pcm_stream_t s) {
  PCM_REQUIRE(n % 8 == 0, "relu_bwd: n must be a multiple of 8");
  if (n == 0) return PCM_OK;
  PCM_DISPATCH_DTYPE(dtype, T, (pcm::launch(relu_bwd_kernel<T>, ew_grid(n / 8), 256, 0, (cudaStream_t)s, 
                                   static_cast<const T*>(dy), static_cast<const T*>(y), static_cast<T*>(dx), n / 8)));
  return check_launch("relu_bwd");
}

extern "C" int pcm_relu_bwd_scaled(const void* dy, const void* y, void* dx, long long n, float scale, int dtype,
                                   pcm_stream_t s) {
  PCM_REQUIRE(n % 8 == 0, "relu_bwd_scaled: n must be a multiple of 8");
  if (n == 0) return PCM_OK;
  PCM_DISPATCH_DTYPE(dtype, T, (pcm::launch(relu_bwd_scaled_kernel<T>, ew_grid(n / 8), 256, 0, (cudaStream_t)s,
                                   static_cast<const T*>(dy), static_cast<const T*>(y), static_cast<T*>(dx), n / 8, scale)));
  return check_launch("relu_bwd_scaled");
}

extern "C" int pcm_dropout(const void* x, void* y, long long n, float p, long long seed, int dtype, pcm_stream_t s) {
  PCM_REQUIRE(n % 8 == 0 && p >= 0.f && p < 1.f, "dropout: n must be a multiple of 8 and 0 <= p < 1");
  if (n == 0) return PCM_OK;
  const unsigned long long* epoch = dropout_epoch_cell();
  PCM_REQUIRE(epoch != nullptr, "dropout: could not allocate the epoch cell");
  PCM_DISPATCH_DTYPE(dtype, T, (pcm::launch(dropout_kernel<T>, ew_grid(n / 8), 256, 0, (cudaStream_t)s,
                                   static_cast<const T*>(x), static_cast<T*>(y), n / 8, p, (unsigned long long)seed, epoch)));
  return check_launch("dropout");
}

extern "C" int pcm_dropout_mask(float* mask, long long n, float p, long long seed, pcm_stream_t s) {
  PCM_REQUIRE(p >= 0.f && p < 1.f, "dropout_mask: 0 <= p < 1");
  if (n == 0) return PCM_OK;
  const unsigned long long* epoch = dropout_epoch_cell();
  PCM_REQUIRE(epoch != nullptr, "dropout_mask: could not allocate the epoch cell");
  pcm::launch(dropout_mask_kernel, ew_grid(n), 256, 0, (cudaStream_t)s, mask, n, p, (unsigned long long)seed, epoch);
  return check_launch("dropout_mask");
}

extern "C" int pcm_batch_sum(const void* x, float* out, int B, long long R, int dtype, pcm_stream_t s) {
  PCM_REQUIRE(R % 8 == 0, "batch_sum: row length must be a multiple of 8");
  if (B == 0 || R == 0) return PCM_OK;
  PCM_DISPATCH_DTYPE(dtype, T, (pcm::launch(batch_sum_kernel<T>, ceil_div(R / 8, 128), 128, 0, (cudaStream_t)s, 
                                   static_cast<const T*>(x), out, B, R / 8)));
  return check_launch("batch_sum");
}
