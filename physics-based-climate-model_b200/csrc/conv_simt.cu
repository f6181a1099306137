// SIMT convolution family (gather form), weight gradient and weight packing.
// These are the general-shape kernels (any K, stride, channel count multiple of 8); the
// tensor-core implicit-GEMM path for the 3x3/stride-1 layers lives in conv_tc.cu.
#include "common.cuh"

namespace pcm {

// ------------------------------------------------------------------------------------------------
// weight packing: out[t][o][i] = w[o*so + i*si + t*st] (zero padded)
// ------------------------------------------------------------------------------------------------
// Pixel-group form of a 3x3 / stride-1 kernel (group = g adjacent pixels of a row act as ONE pixel with g times the
// channels; see pcm_conv3x3_tc_grouped): out[kh*3 + s][pa*Op + o][pb*Ip + i] = base[kh*3 + dx][o][i] with
// dx = g*(s-1) + pb - pa + 1 when 0 <= dx <= 2, else 0 — output pixel pa of a group sees input pixel pb of the group
// s-1 groups to the right through tap dx.  base[t][o][i] = (o<O && i<I) ? w[o*so + i*si + t*st] : 0 as below.
__device__ __forceinline__ float pack_weight_value(const float* __restrict__ w, long long so, long long si, long long st,
                                                   int O, int I, int Op, int Ip, int g, unsigned idx) {
  const unsigned Ig = (unsigned)Ip * g, Og = (unsigned)Op * g;
  int i = (int)(idx % Ig), o = (int)((idx / Ig) % Og), t = (int)(idx / (Ig * Og));
  if (g > 1) {
    const int pb = i / Ip, pa = o / Op, kh = t / 3, s = t - kh * 3;
    i -= pb * Ip; o -= pa * Op;
    const int dx = g * (s - 1) + pb - pa + 1;
    if (dx < 0 || dx > 2) return 0.f;
    t = kh * 3 + dx;
  }
  return (o < O && i < I) ? __ldg(w + o * so + i * si + (long long)t * st) : 0.f;
}

template <typename T>
__global__ void pack_weight_kernel(const float* __restrict__ w, long long so, long long si, long long st, int O, int I,
                                   int taps, int Op, int Ip, int g, T* __restrict__ out) {
  PCM_PDL_ENTRY();
  const unsigned total = (unsigned)taps * Op * Ip * g * g;
  for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x)
    out[idx] = from_f<T>(pack_weight_value(w, so, si, st, O, I, Op, Ip, g, idx));
}

// All weight re-packs of a training step in ONE launch: blockIdx.y = job, the x-grid strides over that job's
// elements.  Job record (8 x int64, device memory): w ptr, out ptr, so, si, st, (O | I << 32), (taps | Op << 32),
// (Ip | dtype << 32 | group << 40).
// Work item = (job, 1024-element block of that job): blockIdx.x indexes the flattened work list the host built, so
// every job proceeds in parallel (a per-thread sweep over the jobs serialises ~2 us of memory latency per job).
constexpr int kBatchBlockElems = 1024;

__global__ void __launch_bounds__(256) pack_weights_batched_kernel(const long long* __restrict__ jobs,
                                                                    const int* __restrict__ work) {
  PCM_PDL_ENTRY();
  const int job = work[2 * blockIdx.x], blk = work[2 * blockIdx.x + 1];
  const long long* j = jobs + (long long)job * 8;
  const float* w = reinterpret_cast<const float*>(j[0]);
  void* out = reinterpret_cast<void*>(j[1]);
  const long long so = j[2], si = j[3], st = j[4];
  const int O = (int)(j[5] & 0xffffffffll), I = (int)(j[5] >> 32);
  const int taps = (int)(j[6] & 0xffffffffll), Op = (int)(j[6] >> 32);
  const int Ip = (int)(j[7] & 0xffffffffll), dtype = (int)((j[7] >> 32) & 0xff);
  int g = (int)((j[7] >> 40) & 0xff);                       // pixel-group factor (0 / 1: plain)
  if (g < 1) g = 1;
  const unsigned total = (unsigned)taps * Op * Ip * g * g;
#pragma unroll
  for (int r = 0; r < kBatchBlockElems / 256; ++r) {
    const unsigned idx = (unsigned)blk * kBatchBlockElems + r * 256 + threadIdx.x;
    if (idx >= total) break;
    const float v = pack_weight_value(w, so, si, st, O, I, Op, Ip, g, idx);
    if (dtype == PCM_BF16) reinterpret_cast<__nv_bfloat16*>(out)[idx] = __float2bfloat16_rn(v);
    else reinterpret_cast<float*>(out)[idx] = v;
  }
}

// Packed weight gradients [tap][Co][Cpad] (what the tensor-core weight-gradient kernel reduces into with vector
// atomics) -> the parameter's own layout: dst[co*sa + ci*sb + tap*st] += packed[(tap*Co + co)*Cpad + ci]; the packed
// buffer is zeroed on the way (ready for the next step).  blockIdx.y = job.  Job record (8 x int64): packed ptr,
// dst ptr, sa, sb, st, (Co | Ci_real << 32), (Cpad | taps << 32), unused.
// Work item = (job, co): the [taps][Cpad] slab of one output channel is staged through shared memory so that both
// the packed reads (rows of Cpad floats) and the parameter-layout read-modify-writes (taps*Ci contiguous floats when
// sb == taps, st == 1 — every conv / convT weight) are coalesced.
__global__ void __launch_bounds__(256) unpack_grads_batched_kernel(const long long* __restrict__ jobs,
                                                                    const int* __restrict__ work) {
  PCM_PDL_ENTRY();
  extern __shared__ float slab[];       // [taps][Cpad]
  const int job = work[2 * blockIdx.x], co = work[2 * blockIdx.x + 1];
  const long long* j = jobs + (long long)job * 8;
  float* packed = reinterpret_cast<float*>(j[0]);
  float* dst = reinterpret_cast<float*>(j[1]);
  const long long sa = j[2], sb = j[3], st = j[4];
  const int Co = (int)(j[5] & 0xffffffffll), Ci = (int)(j[5] >> 32);
  const int Cpad = (int)(j[6] & 0xffffffffll), taps = (int)(j[6] >> 32);
  for (int i = threadIdx.x; i < taps * Cpad; i += blockDim.x) {
    const int t = i / Cpad, ci = i - t * Cpad;
    float* p = packed + ((long long)t * Co + co) * Cpad + ci;
    slab[i] = *p;
    *p = 0.f;
  }
  __syncthreads();
  float* d = dst + (long long)co * sa;
  if (sb == taps && st == 1) {
    for (int i = threadIdx.x; i < Ci * taps; i += blockDim.x) {
      const int ci = i / taps, t = i - ci * taps;
      d[i] += slab[t * Cpad + ci];
    }
  } else {
    for (int i = threadIdx.x; i < Ci * taps; i += blockDim.x) {
      const int ci = i / taps, t = i - ci * taps;
      d[ci * sb + t * st] += slab[t * Cpad + ci];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// gather convolution: one thread = one destination pixel x 8 destination channels
// ------------------------------------------------------------------------------------------------
struct GatherGeom {
  long long src_ns, dst_ns;
  int src_ps, dst_ps;
  int Hs, Ws, Sc, Hd, Wd, Dc;
  int N, KH, KW, stride, pad, mode;
};

template <typename T, typename TO>
__global__ void __launch_bounds__(128)
conv_gather_kernel(const T* __restrict__ src, TO* __restrict__ dst, const T* __restrict__ wk,
                   const float* __restrict__ bias, GatherGeom g, int accumulate, int relu) {
  PCM_PDL_ENTRY();
  const long long pix = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long npix = (long long)g.N * g.Hd * g.Wd;
  if (pix >= npix) return;
  const int dc0 = blockIdx.y * 8;
  const int wd = (int)(pix % g.Wd);
  const int hd = (int)((pix / g.Wd) % g.Hd);
  const int n = (int)(pix / ((long long)g.Wd * g.Hd));

  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;

  const T* src_n = src + n * g.src_ns;
  for (int kh = 0; kh < g.KH; ++kh) {
    int hs;
    if (g.mode == 0) {
      hs = hd * g.stride - g.pad + kh;
    } else {
      const int t = hd + g.pad - kh;
      if (t < 0 || (t % g.stride) != 0) continue;
      hs = t / g.stride;
    }
    if (hs < 0 || hs >= g.Hs) continue;
    for (int kw = 0; kw < g.KW; ++kw) {
      int ws;
      if (g.mode == 0) {
        ws = wd * g.stride - g.pad + kw;
      } else {
        const int t = wd + g.pad - kw;
        if (t < 0 || (t % g.stride) != 0) continue;
        ws = t / g.stride;
      }
      if (ws < 0 || ws >= g.Ws) continue;
      const T* sp = src_n + ((long long)hs * g.Ws + ws) * g.src_ps;
      const T* wp = wk + ((long long)(kh * g.KW + kw) * g.Dc + dc0) * g.Sc;
      for (int sc = 0; sc < g.Sc; sc += 8) {
        float xv[8];
        load8(sp + sc, xv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float wv[8];
          load8(wp + (long long)j * g.Sc + sc, wv);
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[j] = fmaf(xv[k], wv[k], acc[j]);
        }
      }
    }
  }
  if (bias != nullptr) {
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += __ldg(bias + dc0 + j);
  }
  TO* dp = dst + n * g.dst_ns + ((long long)hd * g.Wd + wd) * g.dst_ps + dc0;
  if (accumulate) {
    float old[8];
    load8(dp, old);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += old[j];
  }
  if (relu) {
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = fmaxf(acc[j], 0.f);
  }
  store8(dp, acc);
}

// ------------------------------------------------------------------------------------------------
// weight gradient.  Thread = (8 A-channels x 8 B-channels) tile for one tap, looping over the
// A-grid pixels of its chunk; `lanes` threads share a tile and are reduced in shared memory.
// ------------------------------------------------------------------------------------------------
struct WgradGeom {
  long long a_ns, b_ns, sa, sb, st;
  int a_ps, b_ps;
  int Ha, Wa, Ca, Ca_real, Hb, Wb, Cb, Cb_real;
  int N, KH, KW, stride, pad;
  int tiles_a, tiles_b, tiles, lanes, tiles_per_block;
  long long chunk;   // A-pixels per block
};

template <typename T>
__global__ void __launch_bounds__(256)
conv_wgrad_kernel(const T* __restrict__ A, const T* __restrict__ B, float* __restrict__ dw, WgradGeom g) {
  PCM_PDL_ENTRY();
  extern __shared__ float red[];   // [tiles_per_block][64] when lanes > 1
  const int tap = blockIdx.z;
  const int kh = tap / g.KW, kw = tap % g.KW;
  const int tile_local = threadIdx.x % g.tiles_per_block;
  const int lane = threadIdx.x / g.tiles_per_block;
  const int tile = blockIdx.y * g.tiles_per_block + tile_local;
  const bool active = (tile < g.tiles) && (lane < g.lanes);
  const int ta = tile / g.tiles_b, tb = tile % g.tiles_b;

  float acc[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) acc[i] = 0.f;

  if (g.lanes > 1) {
    for (int i = threadIdx.x; i < g.tiles_per_block * 64; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
  }

  if (active) {
    const long long npix = (long long)g.N * g.Ha * g.Wa;
    const long long p0 = blockIdx.x * g.chunk;
    const long long p1 = min(npix, p0 + g.chunk);
    for (long long p = p0 + lane; p < p1; p += g.lanes) {
      const int wa = (int)(p % g.Wa);
      const int ha = (int)((p / g.Wa) % g.Ha);
      const int n = (int)(p / ((long long)g.Wa * g.Ha));
      const int hb = ha * g.stride - g.pad + kh;
      const int wb = wa * g.stride - g.pad + kw;
      if (hb < 0 || hb >= g.Hb || wb < 0 || wb >= g.Wb) continue;
      float av[8], bv[8];
      load8(A + n * g.a_ns + ((long long)ha * g.Wa + wa) * g.a_ps + ta * 8, av);
      load8(B + n * g.b_ns + ((long long)hb * g.Wb + wb) * g.b_ps + tb * 8, bv);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i * 8 + j] = fmaf(av[i], bv[j], acc[i * 8 + j]);
    }
  }
  if (g.lanes > 1) {
    if (active) {
#pragma unroll
      for (int i = 0; i < 64; ++i) atomicAdd(&red[tile_local * 64 + i], acc[i]);
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < g.tiles_per_block * 64; idx += blockDim.x) {
      const int tl = idx / 64, e = idx % 64;
      const int t2 = blockIdx.y * g.tiles_per_block + tl;
      if (t2 >= g.tiles) continue;
      const int ac = (t2 / g.tiles_b) * 8 + e / 8, bc = (t2 % g.tiles_b) * 8 + e % 8;
      if (ac < g.Ca_real && bc < g.Cb_real) atomicAdd(dw + ac * g.sa + bc * g.sb + tap * g.st, red[idx]);
    }
  } else if (active) {
#pragma unroll
    for (int i = 0; i < 64; ++i) {
      const int ac = ta * 8 + i / 8, bc = tb * 8 + i % 8;
      if (ac < g.Ca_real && bc < g.Cb_real) atomicAdd(dw + ac * g.sa + bc * g.sb + tap * g.st, acc[i]);
    }
  }
}

// per-channel sum over pixels
template <typename T>
__global__ void __launch_bounds__(256)
channel_sum_kernel(const T* __restrict__ x, long long ns, int ps, int N, int P, int C, int C_real,
                   float* __restrict__ out, int per_image) {
  PCM_PDL_ENTRY();
  extern __shared__ float sacc[];   // [C]
  for (int i = threadIdx.x; i < C; i += blockDim.x) sacc[i] = 0.f;
  __syncthreads();
  const int cv = C / 8;
  // per_image: blockIdx.y = image, the x-grid covers that image's pixels only
  const int n_fixed = per_image ? blockIdx.y : -1;
  const long long total = per_image ? (long long)P * cv : (long long)N * P * cv;
  if (per_image) out += (long long)n_fixed * C_real;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  int my_cb = -1;
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < total;
       v += (long long)gridDim.x * blockDim.x) {
    const int cb = (int)(v % cv);
    const long long pix = v / cv;
    const int n = per_image ? n_fixed : (int)(pix / P);
    const int p = (int)(pix % P);
    if (cb != my_cb) {
      if (my_cb >= 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { atomicAdd(&sacc[my_cb * 8 + j], acc[j]); acc[j] = 0.f; }
      }
      my_cb = cb;
    }
    float xv[8];
    load8(x + n * ns + (long long)p * ps + cb * 8, xv);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += xv[j];
  }
  // cv a power of two < 32 (and the thread stride a multiple of cv): every thread kept one channel block, equal to
  // lane % cv — combine the lanes that share it with shuffles so that only cv lanes per warp touch shared memory
  // (256 threads hammering 16 addresses was the whole cost of the thin-layer bias gradients)
  const bool uniform = cv < 32 && (cv & (cv - 1)) == 0 && ((long long)gridDim.x * blockDim.x) % cv == 0;
  if (uniform) {
    for (int off = cv; off < 32; off <<= 1) {
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], off);
    }
    const int lane_cb = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) % cv);
    if ((int)(threadIdx.x & 31) < cv) {
#pragma unroll
      for (int j = 0; j < 8; ++j) atomicAdd(&sacc[lane_cb * 8 + j], acc[j]);
    }
  } else if (my_cb >= 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(&sacc[my_cb * 8 + j], acc[j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C_real; i += blockDim.x) atomicAdd(out + i, sacc[i]);
}

}  // namespace pcm

using namespace pcm;

extern "C" int pcm_pack_weight(const float* w, long long so, long long si, long long st, int O, int I, int taps,
                               int Op, int Ip, void* out, int dtype, pcm_stream_t s) {
  PCM_REQUIRE(O <= Op && I <= Ip && taps > 0, "pack_weight: bad sizes");
  const long long total = (long long)taps * Op * Ip;
  const int blocks = (int)min((long long)1184, (total + 255) / 256);
  PCM_DISPATCH_DTYPE(dtype, T, (pcm::launch(pack_weight_kernel<T>, blocks, 256, 0, (cudaStream_t)s, 
                                   w, so, si, st, O, I, taps, Op, Ip, 1, (T*)out)));
  return check_launch("pack_weight");
}

extern "C" int pcm_pack_weight_grouped(const float* w, long long so, long long si, long long st, int O, int I, int Op,
                                       int Ip, int group, void* out, int dtype, pcm_stream_t s) {
  PCM_REQUIRE(O <= Op && I <= Ip && group >= 1 && group <= 8, "pack_weight_grouped: bad sizes");
  const long long total = 9ll * Op * Ip * group * group;
  PCM_REQUIRE(total < (1ll << 31), "pack_weight_grouped: too large");
  const int blocks = (int)min((long long)1184, (total + 255) / 256);
  PCM_DISPATCH_DTYPE(dtype, T, (pcm::launch(pack_weight_kernel<T>, blocks, 256, 0, (cudaStream_t)s,
                                   w, so, si, st, O, I, 9, Op, Ip, group, (T*)out)));
  return check_launch("pack_weight_grouped");
}

extern "C" int pcm_pack_weights_batched(const long long* jobs, const int* work, int nwork, pcm_stream_t s) {
  PCM_REQUIRE(jobs != nullptr && work != nullptr && nwork >= 0, "pack_weights_batched: bad arguments");
  if (nwork == 0) return PCM_OK;
  pcm::launch(pack_weights_batched_kernel, (unsigned)nwork, 256, 0, (cudaStream_t)s, jobs, work);
  return check_launch("pack_weights_batched");
}

extern "C" int pcm_unpack_grads_batched(const long long* jobs, const int* work, int nwork, pcm_stream_t s) {
  PCM_REQUIRE(jobs != nullptr && work != nullptr && nwork >= 0, "unpack_grads_batched: bad arguments");
  if (nwork == 0) return PCM_OK;
  // dynamic shared memory: the largest [taps][Cpad] slab the library produces (9 x 512 floats)
  pcm::launch(unpack_grads_batched_kernel, (unsigned)nwork, 256, 9 * 512 * sizeof(float), (cudaStream_t)s, jobs, work);
  return check_launch("unpack_grads_batched");
}

extern "C" int pcm_conv_gather(const void* src, long long src_ns, int src_ps, int Hs, int Ws, int Sc, void* dst,
                               long long dst_ns, int dst_ps, int Hd, int Wd, int Dc, const void* wk,
                               const float* bias, int N, int KH, int KW, int stride, int pad, int mode, int dst_f32,
                               int accumulate, int relu, int dtype, pcm_stream_t s) {
  PCM_REQUIRE(Sc % 8 == 0 && Dc % 8 == 0, "conv_gather: channels must be multiples of 8 (Sc=%d Dc=%d)", Sc, Dc);
  PCM_REQUIRE(src_ps % 8 == 0 && dst_ps % 8 == 0, "conv_gather: pixel strides must be multiples of 8");
  PCM_REQUIRE(mode == 0 || mode == 1, "conv_gather: bad mode");
  if (N == 0) return PCM_OK;
  GatherGeom g{src_ns, dst_ns, src_ps, dst_ps, Hs, Ws, Sc, Hd, Wd, Dc, N, KH, KW, stride, pad, mode};
  const long long npix = (long long)N * Hd * Wd;
  dim3 grid(ceil_div(npix, 128), Dc / 8);
  cudaStream_t st = (cudaStream_t)s;
  if (dst_f32) {
    PCM_DISPATCH_DTYPE(dtype, T, (pcm::launch(conv_gather_kernel<T, float>, grid, 128, 0, st, 
                                     (const T*)src, (float*)dst, (const T*)wk, bias, g, accumulate, relu)));
  } else {
    PCM_DISPATCH_DTYPE(dtype, T, (pcm::launch(conv_gather_kernel<T, T>, grid, 128, 0, st, 
                                     (const T*)src, (T*)dst, (const T*)wk, bias, g, accumulate, relu)));
  }
  return check_launch("conv_gather");
}

extern "C" int pcm_conv_wgrad(const void* A, long long a_ns, int a_ps, int Ha, int Wa, int Ca, int Ca_real,
                              const void* B, long long b_ns, int b_ps, int Hb, int Wb, int Cb, int Cb_real,
                              float* dw, long long sa, long long sb, long long st, int N, int KH, int KW, int stride,
                              int pad, int dtype, pcm_stream_t s) {
  PCM_REQUIRE(Ca % 8 == 0 && Cb % 8 == 0, "conv_wgrad: channels must be multiples of 8");
  if (N == 0) return PCM_OK;
  WgradGeom g;
  g.a_ns = a_ns; g.b_ns = b_ns; g.sa = sa; g.sb = sb; g.st = st; g.a_ps = a_ps; g.b_ps = b_ps;
  g.Ha = Ha; g.Wa = Wa; g.Ca = Ca; g.Ca_real = Ca_real; g.Hb = Hb; g.Wb = Wb; g.Cb = Cb; g.Cb_real = Cb_real;
  g.N = N; g.KH = KH; g.KW = KW; g.stride = stride; g.pad = pad;
  g.tiles_a = Ca / 8; g.tiles_b = Cb / 8; g.tiles = g.tiles_a * g.tiles_b;
  g.tiles_per_block = g.tiles < 256 ? g.tiles : 256;
  g.lanes = 256 / g.tiles_per_block;
  const int grid_y = ceil_div(g.tiles, g.tiles_per_block);
  const long long npix = (long long)N * Ha * Wa;
  // aim for ~4 waves of blocks over 148 SMs, but at least 64 pixels per lane
  long long want_chunks = (4 * 148 + grid_y * KH * KW - 1) / (grid_y * KH * KW);
  if (want_chunks < 1) want_chunks = 1;
  long long chunk = (npix + want_chunks - 1) / want_chunks;
  const long long min_chunk = 64LL * g.lanes;
  if (chunk < min_chunk) chunk = min_chunk;
  g.chunk = chunk;
  dim3 grid(ceil_div(npix, chunk), grid_y, KH * KW);
  const size_t smem = g.lanes > 1 ? (size_t)g.tiles_per_block * 64 * sizeof(float) : 0;
  if (smem > 48 * 1024) {
    PCM_DISPATCH_DTYPE(dtype, T, cudaFuncSetAttribute(conv_wgrad_kernel<T>,
                                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  PCM_DISPATCH_DTYPE(dtype, T, (pcm::launch(conv_wgrad_kernel<T>, grid, 256, smem, (cudaStream_t)s, 
                                   (const T*)A, (const T*)B, dw, g)));
  return check_launch("conv_wgrad");
}

namespace pcm {
// Column sums of a [rows = N*P][C] view (bias gradients: out[c] += sum over all pixels).  A CTA owns a block of up to 64
// channels (8 vectors of 8) and a chunk of rows: 32 row lanes x 8 channel vectors, every warp reads whole 128-byte row
// segments; each thread accumulates its 8 channels in registers over its rows, the 32 row lanes are then combined
// through shared memory and ONE global atomic per channel leaves the CTA.  (The general kernel above walks a flat
// vector index: with C/8 not a power of two every iteration flushed through shared-memory atomics, and its 592 CTAs
// each sent C global atomics — 25 us for 10 MB at C = 384.)
template <typename T>
__global__ void __launch_bounds__(256)
channel_sum_rows_kernel(const T* __restrict__ x, long long ns, int ps, int P, long long rows, int C, int C_real,
                        float* __restrict__ out, int rows_per_cta) {
  PCM_PDL_ENTRY();
  __shared__ float sm[32][65];
  const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;
  const int cb = blockIdx.x * 8 + tx;                     // channel vector
  const bool live = cb * 8 < C;
  const long long r0 = (long long)blockIdx.y * rows_per_cta;
  const long long r1 = min(rows, r0 + rows_per_cta);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (live) {
    long long r = r0 + ty;
    // two rows in flight per thread
    for (; r + 32 < r1; r += 64) {
      float a[8], b[8];
      const long long ra = r, rb = r + 32;
      load8(x + (ra / P) * ns + (ra % P) * (long long)ps + cb * 8, a);
      load8(x + (rb / P) * ns + (rb % P) * (long long)ps + cb * 8, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += a[j] + b[j];
    }
    for (; r < r1; r += 32) {
      float a[8];
      load8(x + (r / P) * ns + (r % P) * (long long)ps + cb * 8, a);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += a[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) sm[ty][tx * 8 + j] = acc[j];
  __syncthreads();
  if (threadIdx.x < 64) {
    const int c = blockIdx.x * 64 + threadIdx.x;
    float t = 0.f;
#pragma unroll 8
    for (int k = 0; k < 32; ++k) t += sm[k][threadIdx.x];
    if (c < C_real) atomicAdd(out + c, t);
  }
}
}  // namespace pcm

extern "C" int pcm_channel_sum(const void* x, long long ns, int ps, int N, int P, int C, int C_real, float* out,
                               int per_image, int dtype, pcm_stream_t s) {
  PCM_REQUIRE(C % 8 == 0, "channel_sum: C must be a multiple of 8");
  if (N == 0) return PCM_OK;
  if (!per_image) {
    const long long rows = (long long)N * P;
    const int cblocks = (C / 8 + 7) / 8;
    // ~4 CTAs per SM in total, at least 64 rows per CTA
    long long chunks = (148 * 4 + cblocks - 1) / cblocks;
    if (chunks > (rows + 63) / 64) chunks = (rows + 63) / 64;
    if (chunks < 1) chunks = 1;
    const int rows_per_cta = (int)((rows + chunks - 1) / chunks);
    dim3 grid(cblocks, (unsigned)((rows + rows_per_cta - 1) / rows_per_cta));
    PCM_DISPATCH_DTYPE(dtype, T, (pcm::launch(pcm::channel_sum_rows_kernel<T>, grid, 256, 0, (cudaStream_t)s,
                                     (const T*)x, ns, ps, P, rows, C, C_real, out, rows_per_cta)));
    return check_launch("channel_sum");
  }
  const long long total = per_image ? (long long)P * (C / 8) : (long long)N * P * (C / 8);
  int bx = (int)min((long long)592, (total + 255) / 256);
  if (per_image) bx = (int)min((long long)max(1, 592 / N), (total + 1023) / 1024);
  dim3 grid(bx, per_image ? N : 1);
  PCM_DISPATCH_DTYPE(dtype, T, (pcm::launch(channel_sum_kernel<T>, grid, 256, C * sizeof(float), (cudaStream_t)s, 
                                   (const T*)x, ns, ps, N, P, C, C_real, out, per_image)));
  return check_launch("channel_sum");
}
