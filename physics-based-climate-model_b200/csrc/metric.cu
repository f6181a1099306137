// cos(lat)-area-weighted metric triplet (src/utils_final.py:282-302, main_final.py:616-631):
// monthly RMSE, time-mean RMSE and time-std MAE per variable, in one pass over pred/truth.
// Pass 1 (HBM-bound, reads 2*T*V*Y*X*4 bytes once): per pixel time sums of p, p^2, t, t^2, (p-t)^2
// accumulated in fp64 (tas ~ 273 K: fp32 sums of squares would cancel catastrophically in the
// variance) plus the three counts of non-NaN terms.  Pass 2: per-pixel terms -> latitude-weighted warp-shuffle
// reduction.
// NaN handling = xarray's (skipna=True reductions and DataArray.weighted().mean(), src/utils_final.py:296): a NaN
// prediction/target drops out of that pixel's time mean / std; (p-t)^2 drops out where either is NaN; the weighted
// means divide by the weights of the terms that remain.  Without NaNs this is the plain definition.
#include "common.cuh"

namespace pcm {

constexpr int kMS = 8;       // doubles per pixel: sum p, p^2, t, t^2, (p-t)^2, count p, count t, count (p-t)

// Validation epilogue (main_final.py:563-574 + src/utils_final.py:130-206): predictions and targets may arrive
// NORMALISED; the inverse transform of Normalizer.inverse_transform_output is applied on the fly in fp64, so the
// de-normalised tensors are never materialised.  tr[v] = (kind, a, b, c): x_phys = g(x*a + b) with
//   kind 0 zscore / minimax: identity   1 log1p: expm1   2 sqrt: square   3 pow: (.)^(1/c)
__device__ __forceinline__ double denorm(double x, int kind, double a, double b, double c) {
  const double u = fma(x, a, b);
  return kind == 0 ? u : kind == 1 ? expm1(u) : kind == 2 ? u * u : pow(u, 1.0 / c);
}

// grid: (ceil(V*Y*X / 32), time_chunks); block = 32 pixels x 8 time lanes.  A warp reads 32 consecutive pixels of one time
// step (128 bytes), the 8 warps of a block stride over the time steps of the chunk with kMU steps in flight per thread;
// the 8 time lanes of a pixel are combined through shared memory, so a block issues ONE fp64 atomic per (pixel, statistic)
// — the first version gave every thread its own pixel and 22 time chunks: 1.2 M fp64 atomics for 6 912 pixels, and
// 1.6 TB/s (profiles/r2_metric_kernel.md).
constexpr int kMU = 8;       // time steps in flight per thread

template <typename TI, bool DENORM>
__global__ void __launch_bounds__(256)
metric_partial_kernel(const TI* __restrict__ pred, const TI* __restrict__ truth, const float* __restrict__ tr,
                      double* __restrict__ partial, int T, int YX, int VYX) {
  PCM_PDL_ENTRY();
  __shared__ double red[kMS][8][33];
  const int lane = threadIdx.x & 31, tl = threadIdx.x >> 5;
  const int pix = blockIdx.x * 32 + lane;
  const bool valid = pix < VYX;
  int kind = 0;
  double a = 1.0, b = 0.0, c = 1.0;
  if (DENORM && valid) {
    const int v = pix / YX;
    kind = (int)__ldg(tr + 4 * v);
    a = (double)__ldg(tr + 4 * v + 1); b = (double)__ldg(tr + 4 * v + 2); c = (double)__ldg(tr + 4 * v + 3);
  }
  const int per = (T + gridDim.y - 1) / gridDim.y;
  const int t0 = blockIdx.y * per, t1 = min(T, t0 + per);
  double sp = 0, spp = 0, st = 0, stt = 0, sd = 0;
  int np = 0, nt = 0, nd = 0;
  if (valid) {
    for (int tb = t0 + tl; tb < t1; tb += 8 * kMU) {
      TI pv[kMU], qv[kMU];
#pragma unroll
      for (int u = 0; u < kMU; ++u) {
        const int t = tb + 8 * u;
        const long long off = (long long)min(t, t1 - 1) * VYX + pix;
        pv[u] = __ldg(pred + off);
        qv[u] = __ldg(truth + off);
      }
#pragma unroll
      for (int u = 0; u < kMU; ++u) {
        if (tb + 8 * u < t1) {
          double p = (double)pv[u], q = (double)qv[u];
          if (DENORM) { p = denorm(p, kind, a, b, c); q = denorm(q, kind, a, b, c); }
          const bool okp = p == p, okq = q == q;
          if (okp) { sp += p; spp = fma(p, p, spp); ++np; }
          if (okq) { st += q; stt = fma(q, q, stt); ++nt; }
          if (okp && okq) { const double d = p - q; sd = fma(d, d, sd); ++nd; }
        }
      }
    }
  }
  red[0][tl][lane] = sp; red[1][tl][lane] = spp; red[2][tl][lane] = st; red[3][tl][lane] = stt; red[4][tl][lane] = sd;
  red[5][tl][lane] = (double)np; red[6][tl][lane] = (double)nt; red[7][tl][lane] = (double)nd;
  __syncthreads();
  if (valid) {
    // thread (lane, tl) finishes statistic `tl` of pixel `lane`: fixed order over the time lanes
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc += red[tl][k][lane];
    atomicAdd(partial + (long long)pix * kMS + tl, acc);
  }
}

// one block per variable; out[v][0..2]
__global__ void __launch_bounds__(256)
metric_finalize_kernel(const double* __restrict__ partial, const double* __restrict__ w_lat, double* __restrict__ out,
                       int Y, int X) {
  PCM_PDL_ENTRY();
  __shared__ double red[32];
  const int v = blockIdx.x;
  double a0 = 0, w0 = 0, a1 = 0, a2 = 0, w1 = 0;
  for (int i = threadIdx.x; i < Y * X; i += blockDim.x) {
    const double w = w_lat[i / X];
    const double* q = partial + ((long long)v * Y * X + i) * kMS;
    const double np = q[5], nt = q[6], nd = q[7];
    a0 += w * q[4];                            // sum over time of (p-t)^2 at this pixel, w * (terms that exist)
    w0 += w * nd;
    if (np > 0 && nt > 0) {                    // time mean / std exist for both fields at this pixel
      const double mp = q[0] / np, mt = q[2] / nt;
      const double vp = fmax(q[1] / np - mp * mp, 0.0), vt = fmax(q[3] / nt - mt * mt, 0.0);
      a1 += w * (mp - mt) * (mp - mt);
      a2 += w * fabs(sqrt(vp) - sqrt(vt));
      w1 += w;
    }
  }
  a0 = block_sum(a0, red);
  w0 = block_sum(w0, red);
  a1 = block_sum(a1, red);
  a2 = block_sum(a2, red);
  w1 = block_sum(w1, red);
  if (threadIdx.x == 0) {
    out[v * 3 + 0] = sqrt(a0 / w0);
    out[v * 3 + 1] = sqrt(a1 / w1);
    out[v * 3 + 2] = a2 / w1;
  }
}

}  // namespace pcm

using namespace pcm;

template <typename TI, bool DENORM>
static int metric_partial_launch(const TI* pred, const TI* truth, const float* tr, double* partial, int T, int V, int Y,
                                 int X, int zero_first, pcm_stream_t s, const char* what) {
  const int VYX = V * Y * X;
  if (zero_first) {
    cudaError_t e = cudaMemsetAsync(partial, 0, (size_t)VYX * kMS * sizeof(double), (cudaStream_t)s);
    if (e != cudaSuccess) { set_error("%s memset: %s", what, cudaGetErrorString(e)); return PCM_ERR_CUDA; }
  }
  if (T == 0) return PCM_OK;
  const int gx = ceil_div(VYX, 32);
  int chunks = (4 * 148 + gx - 1) / gx;                    // ~4 blocks per SM
  if (chunks > (T + 63) / 64) chunks = (T + 63) / 64;      // at least one round of 8 time lanes x kMU steps per block
  if (chunks < 1) chunks = 1;
  dim3 grid(gx, chunks);
  pcm::launch(metric_partial_kernel<TI, DENORM>, grid, 256, 0, (cudaStream_t)s, pred, truth, tr, partial, T, Y * X, VYX);
  return check_launch(what);
}

extern "C" int pcm_metric_partial(const float* pred, const float* truth, double* partial, int T, int V, int Y, int X,
                                  int zero_first, pcm_stream_t s) {
  return metric_partial_launch<float, false>(pred, truth, nullptr, partial, T, V, Y, X, zero_first, s, "metric_partial");
}

extern "C" int pcm_metric_partial_f64(const double* pred, const double* truth, double* partial, int T, int V, int Y,
                                      int X, int zero_first, pcm_stream_t s) {
  return metric_partial_launch<double, false>(pred, truth, nullptr, partial, T, V, Y, X, zero_first, s,
                                              "metric_partial_f64");
}

extern "C" int pcm_metric_partial_denorm(const float* pred, const float* truth, const float* tr, double* partial, int T,
                                         int V, int Y, int X, int zero_first, pcm_stream_t s) {
  PCM_REQUIRE(tr != nullptr, "metric_partial_denorm: transform table is null");
  return metric_partial_launch<float, true>(pred, truth, tr, partial, T, V, Y, X, zero_first, s, "metric_partial_denorm");
}

extern "C" int pcm_metric_finalize(const double* partial, const double* w_lat, double* out, long long T_total, int V,
                                   int Y, int X, pcm_stream_t s) {
  (void)T_total;      // kept in the ABI: the per-pixel counts in `partial` carry the number of time steps
  pcm::launch(metric_finalize_kernel, V, 256, 0, (cudaStream_t)s, partial, w_lat, out, Y, X);
  return check_launch("metric_finalize");
}
