// Per-image fused ConvBlock tails (reference src/unet.py:35-49 with SEBlock :6-17 and SpatialGate :19-29).
//
// At the emulator's grid sizes one image of one level fits in a single SM's shared memory (48x72x16 bf16 =
// 108 KB, 24x36x32 = 54 KB, ...), so everything that follows a 3x3 convolution of the block runs as ONE kernel
// with one CTA per image and the image resident in shared memory:
//
//   tail 1  (after conv1):  GroupNorm(8) statistics -> normalise -> SiLU                      1 read, 1 write
//   tail 2  (after conv2):  GroupNorm(8) -> SiLU -> SE squeeze/excite -> channel mean/max map
//                           -> 7x7 gate conv -> sigmoid -> out = a*se*gate                    1 read, 1 write
//                           (+ 13 bytes per pixel saved for the backward: mean / max maps, gate, tie count)
//   and the two backward tails (dout -> dy2, da1 -> dy1): the gate's pre-activation gradient comes from dout*out
//   and the saved maps, the activation a2 is recomputed once from the saved pre-normalisation tensor.
//
// The multi-kernel path in convblock.cu (grid-wide passes, 6 forward + 8 backward launches per block) remains for
// images that do not fit (config 5, fp32 at full size) — see pcm_convblock_fused_supported.
#include "tc_common.cuh"
#include "f32x2.cuh"

namespace pcm {

constexpr int kFT = 512;       // max threads per CTA (one CTA per image); small images launch 256 / 128 so that
                               // several CTAs share an SM and their latency-bound phases overlap
#define NT ((int)blockDim.x)
constexpr int kGroups = 8;     // nn.GroupNorm(8, c)
constexpr int kDwFloats = 200; // gate-weight gradient staging: two halves of the image x 98 sums (+ pad)

template <typename T> __device__ __forceinline__ float sigmoid_t(float z);
template <> __device__ __forceinline__ float sigmoid_t<float>(float z) { return sigmoidf_(z); }
template <> __device__ __forceinline__ float sigmoid_t<__nv_bfloat16>(float z) {
  // one SFU op (tanh.approx, rel. error ~2^-11 < bf16 resolution) instead of ex2 + rcp
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * z));
  return fmaf(0.5f, t, 0.5f);
}

// Per-channel sums over the image without shared-memory atomics (fp32 atomicAdd on shared memory is a CAS loop,
// and every warp of the CTA would spin on the same C addresses).  chan_put: this thread's 8 per-channel partials
// (channel block cb = tid % cv) go to part[slot][warp][C] after the lanes of the warp that hold the same cb are
// combined with xor-shuffles; chan_finish: dst[slot][c] = sum over warps.  Both are called by every thread.
__device__ __forceinline__ void chan_put(float (&v)[8], float* part, int slot, int cb, int cv, int C) {
  for (int off = cv; off < 32; off <<= 1) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] += __shfl_xor_sync(0xffffffffu, v[j], off);
  }
  if ((int)(threadIdx.x & 31) < cv) {
    float4* d = reinterpret_cast<float4*>(part + ((size_t)slot * (NT >> 5) + (threadIdx.x >> 5)) * C + cb * 8);
    d[0] = make_float4(v[0], v[1], v[2], v[3]);
    d[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
}
// the same for eight partials held as four fp32 pairs
__device__ __forceinline__ void chan_put(float2 (&v)[4], float* part, int slot, int cb, int cv, int C) {
  for (int off = cv; off < 32; off <<= 1) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[j].x += __shfl_xor_sync(0xffffffffu, v[j].x, off);
      v[j].y += __shfl_xor_sync(0xffffffffu, v[j].y, off);
    }
  }
  if ((int)(threadIdx.x & 31) < cv) {
    float4* d = reinterpret_cast<float4*>(part + ((size_t)slot * (NT >> 5) + (threadIdx.x >> 5)) * C + cb * 8);
    d[0] = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
    d[1] = make_float4(v[2].x, v[2].y, v[3].x, v[3].y);
  }
}
__device__ __forceinline__ void chan_finish(const float* part, int nslots, float* dst, int C) {
  __syncthreads();
  const int nw = NT >> 5;
  for (int i = threadIdx.x; i < nslots * C; i += NT) {
    const int slot = i / C, c = i - slot * C;
    const float* src = part + (size_t)slot * nw * C + c;
    float t = 0.f;
    for (int w = 0; w < nw; ++w) t += src[w * C];
    dst[i] = t;
  }
  __syncthreads();
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

// An 8-element vector as it sits in memory (16 bytes of bf16 / 32 bytes of fp32): the global-memory passes issue
// the loads of several rounds back to back into these and convert afterwards — with 16 warps per SM a loop that
// loads, waits and computes one round at a time is bound by memory latency, not bandwidth.
template <typename T> struct Raw8;
template <> struct Raw8<__nv_bfloat16> { uint4 u; };
template <> struct Raw8<float> { float4 a, b; };
template <bool NC> __device__ __forceinline__ Raw8<__nv_bfloat16> ld_raw(const __nv_bfloat16* p) {
  Raw8<__nv_bfloat16> r;
  r.u = NC ? __ldg(reinterpret_cast<const uint4*>(p)) : *reinterpret_cast<const uint4*>(p);
  return r;
}
template <bool NC> __device__ __forceinline__ Raw8<float> ld_raw(const float* p) {
  Raw8<float> r;
  const float4* q = reinterpret_cast<const float4*>(p);
  r.a = NC ? __ldg(q) : q[0];
  r.b = NC ? __ldg(q + 1) : q[1];
  return r;
}
__device__ __forceinline__ void unpack8(const Raw8<__nv_bfloat16>& r, float d[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r.u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); d[2 * i] = f.x; d[2 * i + 1] = f.y; }
}
__device__ __forceinline__ void unpack8(const Raw8<float>& r, float d[8]) {
  d[0] = r.a.x; d[1] = r.a.y; d[2] = r.a.z; d[3] = r.a.w; d[4] = r.b.x; d[5] = r.b.y; d[6] = r.b.z; d[7] = r.b.w;
}
template <typename T> constexpr int raw_batch() { return sizeof(T) == 2 ? 4 : 2; }   // vectors in flight per operand

// ---- the same vectors as four fp32 PAIRS (f32x2.cuh): the element-wise passes below run on pairs of channels
__device__ __forceinline__ void unpack8(const Raw8<__nv_bfloat16>& r, float2 (&d)[4]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r.u);
#pragma unroll
  for (int i = 0; i < 4; ++i) d[i] = __bfloat1622float2(h[i]);
}
__device__ __forceinline__ void unpack8(const Raw8<float>& r, float2 (&d)[4]) {
  d[0] = make_float2(r.a.x, r.a.y); d[1] = make_float2(r.a.z, r.a.w);
  d[2] = make_float2(r.b.x, r.b.y); d[3] = make_float2(r.b.z, r.b.w);
}
template <typename T>
__device__ __forceinline__ void load8_rw(const T* p, float2 (&d)[4]) {      // generic-address load (shared / own global writes)
  unpack8(ld_raw<false>(p), d);
}
__device__ __forceinline__ void store8(float* p, const float2 (&v)[4]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
  reinterpret_cast<float4*>(p)[1] = make_float4(v[2].x, v[2].y, v[3].x, v[3].y);
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float2 (&v)[4]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[i].x, v[i].y);
  *reinterpret_cast<uint4*>(p) = u;
}

// Gate maps (channel mean / max, dq) live in zero-padded planes: pixel (h, w) at (h + 3) * Wp + (w + 4), with
// Wp = roundup8(W) + 8, so that the 16 floats [w0, w0 + 16) a run of 8 pixels needs from one row are four aligned
// 16-byte shared-memory loads.
__host__ __device__ inline int plane_wp(int W) { return ((W + 7) / 8) * 8 + 8; }

struct TailSmem {
  size_t img, cm0, cm1, dq, gate, dm, scr, bar, part, fl, total;
};
// threads per CTA: enough vectors per thread to amortise the phase barriers, few enough that small images get
// several CTAs per SM (the kernels are register-limited to 512 threads per SM)
__host__ __device__ inline int tail_threads(int H, int W, int C) {
  const int nvec = H * W * (C / 8);
  return nvec >= 4096 ? 512 : nvec >= 1536 ? 256 : 128;
}
// bwd: 0 = forward tails, 1 = backward tails (needs dq / dm / cnt as well)
// scr (backward tails only): room for a second image-sized array — the scratch (dxhat) that GroupNorm pass 1 hands to
// pass 2 then stays in shared memory instead of making a round trip through global memory.  It OVERLAYS the gate planes
// (all dead once pass B is done) and only the excess over them is extra shared memory.
__host__ __device__ inline TailSmem tail_smem_layout(int H, int W, int C, int elt, int full, int bwd, int scr = 0) {
  TailSmem L;
  const size_t P = (size_t)H * W, Pp = (size_t)(H + 6) * plane_wp(W);
  size_t off = 0;
#define PCM_TAKE(bytes) ([&]() { size_t o = off; off += ((size_t)(bytes) + 15) & ~(size_t)15; return o; }())
  L.img = PCM_TAKE(P * C * elt);
  L.cm0 = full ? PCM_TAKE(Pp * 4) : 0;
  L.cm1 = full ? PCM_TAKE(Pp * 4) : 0;
  L.dq = (full && bwd) ? PCM_TAKE(Pp * 4) : 0;
  L.gate = full ? PCM_TAKE(P * 4) : 0;
  L.dm = (full && bwd) ? PCM_TAKE(P * 8) : 0;
  L.scr = 0;
  if (scr && bwd) {
    L.scr = full ? L.cm0 : off;
    const size_t end = L.scr + ((P * C * elt + 15) & ~(size_t)15);
    if (end > off) off = end;
  }
  L.bar = PCM_TAKE(64);                                  // mbarriers of the bulk copies that bring the image in (one per 32 KB piece)
  L.part = PCM_TAKE((size_t)2 * (tail_threads(H, W, C) / 32) * C * 4);    // chan_put: 2 slots x warps x C floats (launches never
                                                                          // use more threads than tail_threads)
  // floats: 5 channel arrays | a[C] b[C] | se[C] pool[C] dpool[C] dpre2[C] | hid[64] dpre1[64] | mu[8] rs[8] m1[8] m2[8]
  //         | wt[2][7][8] | wtf[7][8][2] | sw1[Cr*C] sw2[C*Cr] (Cr = C/8 at most: C*C/4 floats) | dw[2][98 + 2] (backward
  //         tail only, LAST: 384 forward images of 24x36x32 fit three CTAs per SM by 144 bytes)
  L.fl = PCM_TAKE((size_t)(11 * C + 128 + 32 + 112 + 112 + (full ? C * C / 4 : 0) + ((full && bwd) ? kDwFloats : 0)) * 4);
#undef PCM_TAKE
  L.total = off;
  return L;
}

struct TailPtrs {
  float *part;
  float *ch0, *ch1, *ch2, *ch3, *ch4, *ca, *cb_, *se, *pool, *dpool, *dpre2, *hid, *dpre1, *mu, *rs, *m1, *m2, *wt, *wtf, *dw, *sw1, *sw2;
};
__device__ __forceinline__ TailPtrs tail_ptrs(uint8_t* smem, const TailSmem& L, int C) {
  float* f = reinterpret_cast<float*>(smem + L.fl);
  TailPtrs p;
  p.part = reinterpret_cast<float*>(smem + L.part);
  p.ch0 = f; p.ch1 = f + C; p.ch2 = f + 2 * C; p.ch3 = f + 3 * C; p.ch4 = f + 4 * C;
  p.ca = f + 5 * C; p.cb_ = f + 6 * C;
  p.se = f + 7 * C; p.pool = f + 8 * C; p.dpool = f + 9 * C; p.dpre2 = f + 10 * C;
  float* g = f + 11 * C;
  p.hid = g; p.dpre1 = g + 64;
  p.mu = g + 128; p.rs = g + 136; p.m1 = g + 144; p.m2 = g + 152;
  p.wt = g + 160; p.wtf = g + 272;
  p.sw1 = g + 384; p.sw2 = p.sw1 + C * C / 8;
  p.dw = p.sw1 + C * C / 4;
  return p;
}

// 7x7 gate weights (2, 7, 7):
//   wt[k][dy][8]      (dx padded to 8) for the forward gate conv over the two scalar planes (stencil_run8);
//   wtf[dy][8][2]     the flipped kernel as PAIRS over the two maps, wtf[dy][dx] = (w[0][6-dy][6-dx], w[1][6-dy][6-dx]),
//                     for the backward tail: the transposed stencil then produces (dmean, dmax) with one packed FMA
//                     (f32x2.cuh) per tap from a broadcast dq, and the weight gradient accumulates (dw[0], dw[1]) the
//                     same way from an interleaved (mean, max) plane — half the FMA instructions of the scalar form.
__device__ __forceinline__ void load_gate_weights(const float* __restrict__ wsp, const TailPtrs& sp) {
  for (int i = threadIdx.x; i < 112; i += NT) {
    {
      const int k = i / 56, dy = (i % 56) / 8, dx = i % 8;
      sp.wt[i] = dx < 7 ? __ldg(wsp + k * 49 + dy * 7 + dx) : 0.f;
    }
    {
      const int k = i & 1, e = i >> 1, dy = e >> 3, dx = e & 7;
      sp.wtf[i] = dx < 7 ? __ldg(wsp + k * 49 + (6 - dy) * 7 + (6 - dx)) : 0.f;
    }
  }
}

// Bank swizzles of the gate planes.  A stencil work item reads 16 consecutive pixels of a plane row, and the lanes of a
// warp own adjacent 8-pixel runs: their 16-byte loads start 32 bytes apart in a scalar plane (lanes l and l + 4 of a
// quarter-warp hit the same banks: 2-way conflict) and 64 bytes apart in the interleaved pair plane (4-way) — the
// gate-weight gradient phase of the backward tail was bound by exactly these replays (17 k of its 96 k cycles per
// 48x72x16 image, against 5 k of issue).  Storing 16-byte granule g of a plane at g ^ ((g >> 3) & 1) (scalar) /
// g ^ ((g >> 3) & 3) (pairs) makes the eight loads of a quarter-warp land in eight different bank groups.  Indices are
// element indices relative to the plane (float / float2); the permutation stays inside aligned 8-element groups, so
// zero-filling the plane linearly is still right.
__device__ __forceinline__ int pswz(int i) { return i ^ ((i >> 3) & 4); }      // float index: byte-address bit 7 -> bit 4
__device__ __forceinline__ int pswz2(int p) { return p ^ ((p >> 3) & 6); }     // float2 index: bits 7, 8 -> bits 4, 5
// one row of 16 consecutive pixels of the interleaved plane (8 aligned 16-byte loads) as 16 pairs; i = element index of
// the first pixel (a multiple of 8)
__device__ __forceinline__ void load_pair_row16(const float2* plane, int i, float2 (&m)[16]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 v = *reinterpret_cast<const float4*>(plane + pswz2(i + 2 * j));
    m[2 * j] = make_float2(v.x, v.y);
    m[2 * j + 1] = make_float2(v.z, v.w);
  }
}
__device__ __forceinline__ void load_row16(const float* plane, int i, float (&m)[16]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 v = *reinterpret_cast<const float4*>(plane + pswz(i + 4 * j));
    m[4 * j] = v.x; m[4 * j + 1] = v.y; m[4 * j + 2] = v.z; m[4 * j + 3] = v.w;
  }
}
__device__ __forceinline__ void load_weight_row(const float* wrow, float2 (&w)[7]) {   // wt / wtf + dy * 16
  const float4* r = reinterpret_cast<const float4*>(wrow);
  const float4 a = r[0], b = r[1], c = r[2], d = r[3];
  w[0] = make_float2(a.x, a.y); w[1] = make_float2(a.z, a.w); w[2] = make_float2(b.x, b.y); w[3] = make_float2(b.z, b.w);
  w[4] = make_float2(c.x, c.y); w[5] = make_float2(c.z, c.w); w[6] = make_float2(d.x, d.y);
}

// Transposed gate conv: q[i] += (sum_taps wf[0] * dq(..), sum_taps wf[1] * dq(..)) from the scalar dq plane
// (i0 = h * Wp + w0, the element index of the window's first pixel): 4 + 4 vector loads feed 56 packed FMAs with a
// broadcast operand.
__device__ __forceinline__ void stencil_dq_run8(const float* plane, int i0, int Wp, const float* wtf, float2 (&q)[8]) {
#pragma unroll
  for (int dy = 0; dy < 7; ++dy) {
    float m[16];
    float2 w[7];
    load_row16(plane, i0 + dy * Wp, m);
    load_weight_row(wtf + dy * 16, w);
#pragma unroll
    for (int dx = 0; dx < 7; ++dx) {
#pragma unroll
      for (int i = 0; i < 8; ++i) q[i] = fma2(w[dx], bc2(m[i + dx + 1]), q[i]);
    }
  }
}

// q[i] += sum_{dy,dx} wt[dy][dx] * plane(h + dy - 3, w0 + i + dx - 3) for the 8 pixels (h, w0 .. w0 + 7);
// i0 = h * Wp + w0 (element index of the window's first pixel).  Per kernel row: 4 + 2 vector loads feed 56 FMAs.
__device__ __forceinline__ void stencil_run8(const float* plane, int i0, int Wp, const float* wt, float (&q)[8]) {
#pragma unroll
  for (int dy = 0; dy < 7; ++dy) {
    float m[16];
    load_row16(plane, i0 + dy * Wp, m);
    const float4 wa = reinterpret_cast<const float4*>(wt + dy * 8)[0], wb = reinterpret_cast<const float4*>(wt + dy * 8)[1];
    const float w[7] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z};
#pragma unroll
    for (int dx = 0; dx < 7; ++dx) {
#pragma unroll
      for (int i = 0; i < 8; ++i) q[i] = fmaf(w[dx], m[i + dx + 1], q[i]);
    }
  }
}

// (mean, 1/std) of one group from its raw sums — ONE definition, so that the forward kernel and the backward
// kernels (which start from the saved sums) derive bit-identical coefficients
__device__ __forceinline__ void group_mu_rs(float S, float Q, float cnt, float eps, float& mu, float& rs) {
  mu = S / cnt;
  const float var = fmaxf(fmaf(-mu, mu, Q / cnt), 0.f);
  rs = rsqrtf(var + eps);
}
// z = za*x + zb with za = gamma*rs, zb = beta - mu*za ; a = silu(z) rounded to the storage type.  The backward
// tail compares u = a*se against the channel maximum the forward kernel saved, so both use this one definition.
__device__ __forceinline__ void gn_coef(float gamma, float beta, float mu, float rs, float& za, float& zb) {
  za = gamma * rs;
  zb = fmaf(-mu, za, beta);
}
// Rounding to the storage type goes through the PACKING conversion (two values per F2FP, ALU pipe) instead of one
// F2F per value (SFU pipe, which the sigmoid already loads); same round-to-nearest-even result.
template <typename T> __device__ __forceinline__ void round8_to(float (&a)[8]);
template <> __device__ __forceinline__ void round8_to<float>(float (&)[8]) {}
template <> __device__ __forceinline__ void round8_to<__nv_bfloat16>(float (&a)[8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(__floats2bfloat162_rn(a[2 * i], a[2 * i + 1]));
    a[2 * i] = f.x; a[2 * i + 1] = f.y;
  }
}
// store 8 values and leave them in `a` as the storage type holds them
__device__ __forceinline__ void store8_rounded(float* p, float (&a)[8]) { store8(p, a); }
__device__ __forceinline__ void store8_rounded(__nv_bfloat16* p, float (&a)[8]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = __floats2bfloat162_rn(a[2 * i], a[2 * i + 1]);
    const float2 f = __bfloat1622float2(h[i]);
    a[2 * i] = f.x; a[2 * i + 1] = f.y;
  }
  *reinterpret_cast<uint4*>(p) = u;
}
// pack 8 values to bf16 (16 bytes) and leave them in `a` as bf16 holds them
__device__ __forceinline__ uint4 pack8_rounded(float* a) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = __floats2bfloat162_rn(a[2 * i], a[2 * i + 1]);
    const float2 f = __bfloat1622float2(h[i]);
    a[2 * i] = f.x; a[2 * i + 1] = f.y;
  }
  return u;
}
// a = silu(z), unrounded (callers round eight at a time)
template <typename T>
__device__ __forceinline__ float silu_raw(float x, float za, float zb) {
  const float z = fmaf(za, x, zb);
  return z * sigmoid_t<T>(z);
}

// ---- the same on fp32 pairs: per lane exactly the operations of sigmoid_t / silu_raw / the rounding helpers above
template <typename T> __device__ __forceinline__ float2 sigmoid2(float2 z);
template <> __device__ __forceinline__ float2 sigmoid2<float>(float2 z) { return make_float2(sigmoidf_(z.x), sigmoidf_(z.y)); }
template <> __device__ __forceinline__ float2 sigmoid2<__nv_bfloat16>(float2 z) {
  const float2 h = mul2(z, bc2(0.5f));
  float2 t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(h.x));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(h.y));
  return fma2(bc2(0.5f), t, bc2(0.5f));
}
// t <- silu(za*t + zb), unrounded
template <typename T>
__device__ __forceinline__ void silu8(float2 (&t)[4], const float2 (&za)[4], const float2 (&zb)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 z = fma2(za[i], t[i], zb[i]);
    t[i] = mul2(z, sigmoid2<T>(z));
  }
}
template <typename T> __device__ __forceinline__ void round8_to(float2 (&a)[4]);
template <> __device__ __forceinline__ void round8_to<float>(float2 (&)[4]) {}
template <> __device__ __forceinline__ void round8_to<__nv_bfloat16>(float2 (&a)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) a[i] = __bfloat1622float2(__floats2bfloat162_rn(a[i].x, a[i].y));
}
__device__ __forceinline__ void store8_rounded(float* p, float2 (&a)[4]) { store8(p, a); }
__device__ __forceinline__ void store8_rounded(__nv_bfloat16* p, float2 (&a)[4]) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = __floats2bfloat162_rn(a[i].x, a[i].y);
    a[i] = __bfloat1622float2(h[i]);
  }
  *reinterpret_cast<uint4*>(p) = u;
}
// the eight per-channel values of channel block cb of a shared-memory array as four pairs
__device__ __forceinline__ void load_chan8(const float* a, int cb, float2 (&d)[4]) {
  const float4 lo = reinterpret_cast<const float4*>(a + cb * 8)[0], hi = reinterpret_cast<const float4*>(a + cb * 8)[1];
  d[0] = make_float2(lo.x, lo.y); d[1] = make_float2(lo.z, lo.w); d[2] = make_float2(hi.x, hi.y); d[3] = make_float2(hi.z, hi.w);
}

// ---- the image comes in through the bulk-copy engine in 32 KB pieces, one mbarrier per piece (one thread issues;
// a pass that walks the image front to back waits piece by piece, so it starts while the rest is still in flight)
constexpr uint32_t kPieceShift = 15, kPiece = 1u << kPieceShift;              // <= 8 pieces: 256 KB > any image
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void image_copy_start(uint64_t* bars, void* dst, const void* src, uint32_t bytes) {
  const uint32_t npiece = (bytes + kPiece - 1) >> kPieceShift;
  for (uint32_t i = 0; i < npiece; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(bars + i)));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  for (uint32_t i = 0; i < npiece; ++i) {
    const uint32_t off = i << kPieceShift, n = min(kPiece, bytes - off), b = smem_addr(bars + i);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(n) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst) + off), "l"(reinterpret_cast<const uint8_t*>(src) + off), "r"(n), "r"(b) : "memory");
  }
}
__device__ __forceinline__ void image_piece_wait(uint64_t* bars, uint32_t piece) {
  const uint32_t b = smem_addr(bars + piece);
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(b) : "memory");
    if (ok) return;
    if (clock64() - t0 > 4000000000LL) __trap();       // ~2 s: never hang the GPU
  }
}
__device__ __forceinline__ void image_copy_wait(uint64_t* bars, uint32_t bytes) {
  const uint32_t npiece = (bytes + kPiece - 1) >> kPieceShift;
  for (uint32_t i = 0; i < npiece; ++i) image_piece_wait(bars, i);
}
// front-to-back walkers: vector v (16 or 32 bytes) may be read once pieces [0, piece(v)] have landed
struct PieceWalk {
  uint64_t* bars;
  int ready;                       // pieces [0, ready) are known to have landed
  __device__ __forceinline__ PieceWalk(uint64_t* b) : bars(b), ready(0) {}
  __device__ __forceinline__ void need(uint32_t byte_end) {          // data up to byte_end (exclusive) is about to be read
    const int want = (int)((byte_end + kPiece - 1) >> kPieceShift);
    while (ready < want) image_piece_wait(bars, (uint32_t)ready++);
  }
};

// GroupNorm statistics of the image in shared memory -> mu/rs per group (+ raw sums to `stats_out` when non-null)
template <typename T>
__device__ __forceinline__ void image_group_stats(const T* s_img, int nvec, int cv, int cg, int P, float eps,
                                                  const TailPtrs& sp, float* stats_out, uint64_t* bars) {
  const int cb = threadIdx.x & (cv - 1);
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  PieceWalk land(bars);
#pragma unroll 2
  for (int v = threadIdx.x; v < nvec; v += NT) {
    float x[8];
    land.need((uint32_t)(v + 1) * 8u * (uint32_t)sizeof(T));
    load8_rw(s_img + (size_t)v * 8, x);
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] += x[j]; q[j] = fmaf(x[j], x[j], q[j]); }
  }
  land.need((uint32_t)nvec * 8u * (uint32_t)sizeof(T));       // every thread has seen the whole image land
  chan_put(s, sp.part, 0, cb, cv, cv * 8);
  chan_put(q, sp.part, 1, cb, cv, cv * 8);
  chan_finish(sp.part, 2, sp.ch0, cv * 8);
  if (threadIdx.x < kGroups) {
    const int g = threadIdx.x;
    float S = 0.f, Q = 0.f;
    for (int k = 0; k < cg; ++k) { S += sp.ch0[g * cg + k]; Q += sp.ch1[g * cg + k]; }
    group_mu_rs(S, Q, (float)cg * (float)P, eps, sp.mu[g], sp.rs[g]);
    if (stats_out != nullptr) { stats_out[2 * g] = S; stats_out[2 * g + 1] = Q; }
  }
  __syncthreads();
}

__device__ __forceinline__ void group_mu_rs_from_stats(const float* stats_n, int cg, int P, float eps, const TailPtrs& sp) {
  if (threadIdx.x < kGroups) {
    const int g = threadIdx.x;
    group_mu_rs(stats_n[2 * g], stats_n[2 * g + 1], (float)cg * (float)P, eps, sp.mu[g], sp.rs[g]);
  }
}

// Walks the pixels this thread's vectors belong to: vector v = tid + r*NT is (pixel p = v / cv, channel block
// cb = v % cv); the cv lanes of a pixel are adjacent lanes of one warp.  (h, w) advance without divisions.
struct PixWalk {
  int p, h, w, dh, dw, pstep;
  __device__ __forceinline__ PixWalk(int cvs, int W) {
    pstep = NT >> cvs;
    p = (int)threadIdx.x >> cvs;
    h = p / W; w = p - h * W;
    dh = pstep / W; dw = pstep - dh * W;
  }
  __device__ __forceinline__ void next(int W) {
    p += pstep; w += dw; h += dh;
    if (w >= W) { w -= W; ++h; }
  }
};

// Second half of the full forward tail, shared by convblock_tail_fwd_kernel and the fused block kernel
// (convblock_fwd_tc_kernel): on entry s_img holds a = silu(GN(x)) of image n (rounded to T, linear [pixel][C]) and
// sp.ch2[c] the per-channel sums of a (published by a barrier).  Squeeze / excite, channel mean / max maps, the 7x7 gate
// and out = a*se*gate; saves pool / hid / se and (training) maps / ties.  Every thread of the CTA calls it.
template <typename T>
__device__ __forceinline__ void tail_finish_from_pool(T* s_img, uint8_t* smem, const TailSmem& L, const TailPtrs& sp, int n,
                                                      int H, int W, int C, int Cr, float* __restrict__ pool_g,
                                                      float* __restrict__ se_g, float* __restrict__ hid_g,
                                                      float* __restrict__ maps, uint8_t* __restrict__ ties,
                                                      T* __restrict__ on) {
  const int P = H * W, cv = C / 8, nvec = P * cv, Wp = plane_wp(W);
  const int cvs = __ffs(cv) - 1;
  const int cb = threadIdx.x & (cv - 1);
  // ---- squeeze / excite (SEBlock.forward, src/unet.py:16-17)
  const float invP = 1.f / (float)P;
  for (int c = threadIdx.x; c < C; c += NT) pool_g[(size_t)n * C + c] = sp.ch2[c];
  // the two 1x1 "fc" matrices were staged in shared memory at kernel start; one warp per hidden unit
  for (int j = threadIdx.x >> 5; j < Cr; j += NT / 32) {
    float a = 0.f;
    for (int c = threadIdx.x & 31; c < C; c += 32) a = fmaf(sp.sw1[j * C + c], sp.ch2[c], a);
    a = fmaxf(warp_sum(a) * invP, 0.f);
    if ((threadIdx.x & 31) == 0) { sp.hid[j] = a; hid_g[(size_t)n * Cr + j] = a; }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += NT) {
    float a = 0.f;
    for (int j = 0; j < Cr; ++j) a = fmaf(sp.sw2[c * Cr + j], sp.hid[j], a);
    a = sigmoidf_(a);
    sp.se[c] = a;
    se_g[(size_t)n * C + c] = a;
  }
  __syncthreads();

  // ---- channel mean / max of u = a*se (SpatialGate.forward, src/unet.py:26-27): each thread reduces its 8
  // channels, the cv lanes of a pixel finish with a shuffle butterfly.  When `maps` is given (training) the maps,
  // the gate and the number of channels that attain the maximum (amax splits its gradient between ties) are saved
  // for the backward tail: maps[n] = mean[P] | max[P] | gate[P], ties[n][P].
  float* cm0 = reinterpret_cast<float*>(smem + L.cm0);
  float* cm1 = reinterpret_cast<float*>(smem + L.cm1);
  float* s_gate = reinterpret_cast<float*>(smem + L.gate);
  float* mp = maps != nullptr ? maps + (size_t)n * 3 * P : nullptr;
  uint8_t* tp = maps != nullptr ? ties + (size_t)n * P : nullptr;
  float sc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) sc[j] = sp.se[cb * 8 + j];
  {
    const int nround = (nvec + NT - 1) / NT;
    const float invC = 1.f / (float)C;
    PixWalk pw(cvs, W);
#pragma unroll 2
    for (int r = 0; r < nround; ++r, pw.next(W)) {
      const bool valid = pw.p < P;
      float sum = 0.f, mx = -INFINITY;
      float u[8];
      if (valid) {
        load8_rw(s_img + ((size_t)pw.p * cv + cb) * 8, u);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          u[j] *= sc[j];
          sum += u[j];
          mx = fmaxf(mx, u[j]);
        }
      }
      for (int off = 1; off < cv; off <<= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, off);
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      }
      int cnt = 0;
      if (mp != nullptr) {
        if (valid) {
#pragma unroll
          for (int j = 0; j < 8; ++j) cnt += (u[j] == mx) ? 1 : 0;
        }
        for (int off = 1; off < cv; off <<= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
      }
      if (valid && cb == 0) {
        const int ip = (pw.h + 3) * Wp + pw.w + 4;
        const float mean = sum * invC;
        cm0[pswz(ip)] = mean;
        cm1[pswz(ip)] = mx;
        if (mp != nullptr) {
          mp[pw.p] = mean;
          mp[P + pw.p] = mx;
          tp[pw.p] = (uint8_t)min(cnt, 255);
        }
      }
    }
  }
  __syncthreads();
  // ---- gate = sigmoid(conv7x7([mean, max]))  (:28), 8 pixels of a row per work item
  {
    const int nrun = (W + 7) / 8;
    for (int item = threadIdx.x; item < H * nrun; item += NT) {
      const int h = item / nrun, w0 = (item - h * nrun) * 8;
      float q[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) q[i] = 0.f;
      stencil_run8(cm0, h * Wp + w0, Wp, sp.wt, q);
      stencil_run8(cm1, h * Wp + w0, Wp, sp.wt + 56, q);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (w0 + i < W) {
          const float gt = sigmoidf_(q[i]);
          s_gate[h * W + w0 + i] = gt;
          if (mp != nullptr) mp[2 * P + h * W + w0 + i] = gt;
        }
      }
    }
  }
  __syncthreads();
#pragma unroll 2
  for (int v = threadIdx.x; v < nvec; v += NT) {
    float t[8];
    load8_rw(s_img + (size_t)v * 8, t);
    const float gt = s_gate[v >> cvs];
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = t[j] * sc[j] * gt;
    store8(on + (size_t)v * 8, t);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// forward tails.  FULL = false: y = silu(GN(x)).  FULL = true: out = a*se*gate with a = silu(GN(x)).
// ---------------------------------------------------------------------------------------------------------------
// MAXT = 256: the instantiation for images that launch at most 256 threads — compiled for three CTAs per SM (<= 85
// registers instead of 90), which lets 384 images of 24x36x32 run as ONE wave (444 slots) instead of two.
template <typename T, bool FULL, int MAXT = kFT>
__global__ void __launch_bounds__(MAXT, MAXT == 256 ? 3 : 1)
convblock_tail_fwd_kernel(const T* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                          const float* __restrict__ w1, const float* __restrict__ w2, const float* __restrict__ wsp,
                          float* __restrict__ stats, float* __restrict__ pool_g, float* __restrict__ se_g,
                          float* __restrict__ hid_g, float* __restrict__ maps, uint8_t* __restrict__ ties,
                          T* __restrict__ out, int H, int W, int C, int Cr, float eps) {
  pdl_launch_dependents();
  extern __shared__ __align__(16) uint8_t smem[];
  const int n = blockIdx.x, P = H * W, cv = C / 8, cg = C / kGroups, nvec = P * cv, Wp = plane_wp(W);
  const int cvs = __ffs(cv) - 1;
  const TailSmem L = tail_smem_layout(H, W, C, (int)sizeof(T), FULL ? 1 : 0, 0);
  T* s_img = reinterpret_cast<T*>(smem + L.img);
  const TailPtrs sp = tail_ptrs(smem, L, C);
  const T* xn = x + (size_t)n * P * C;
  T* on = out + (size_t)n * P * C;
  const int cb = threadIdx.x & (cv - 1);

  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L.bar);
  if (FULL) {
    float4* z4 = reinterpret_cast<float4*>(smem + L.cm0);      // cm0 and cm1 are contiguous
    for (int i = threadIdx.x; i < 2 * (H + 6) * Wp / 4; i += NT) z4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  pdl_wait();                                                  // global memory from here on
  if (threadIdx.x == 0) image_copy_start(bar, s_img, xn, (uint32_t)((size_t)P * C * sizeof(T)));
  if (FULL) {
    load_gate_weights(wsp, sp);
    for (int i = threadIdx.x; i < C * Cr; i += NT) { sp.sw1[i] = __ldg(w1 + i); sp.sw2[i] = __ldg(w2 + i); }
  }
  __syncthreads();
  image_group_stats<T>(s_img, nvec, cv, cg, P, eps, sp, stats + (size_t)n * kGroups * 2, bar);

  float ga[8], be[8], acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cb * 8 + j, g = c / cg;
    gn_coef(__ldg(gamma + c), __ldg(beta + c), sp.mu[g], sp.rs[g], ga[j], be[j]);
    acc[j] = 0.f;
  }
#pragma unroll 2
  for (int v = threadIdx.x; v < nvec; v += NT) {
    float t[8];
    load8_rw(s_img + (size_t)v * 8, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = silu_raw<T>(t[j], ga[j], be[j]);
    if (FULL) store8_rounded(s_img + (size_t)v * 8, t);
    else store8_rounded(on + (size_t)v * 8, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += t[j];
  }
  if (!FULL) return;

  // ---- squeeze / excite pool sums, then the shared second half
  chan_put(acc, sp.part, 0, cb, cv, C);
  chan_finish(sp.part, 1, sp.ch2, C);
  tail_finish_from_pool<T>(s_img, smem, L, sp, n, H, W, C, Cr, pool_g, se_g, hid_g, maps, ties, on);
}

// ---------------------------------------------------------------------------------------------------------------
// Whole ConvBlock forward in ONE kernel, one CTA per image (reference src/unet.py:35-49):
//     x -> conv3x3 -> GN -> SiLU -> conv3x3 -> GN -> SiLU -> SE -> spatial gate -> out
// for the thin layers (C = 16 / 32 / 64 output channels) whose image — and both weight tensors — fit one SM.
//
//   * the input image arrives by TMA as a zero-ringed halo image [(H+2) x (W+2) pixels][Cin] in the canonical K-major
//     swizzled UMMA layout (out-of-bounds fill = the conv padding); both weight tensors sit next to it, resident;
//   * each 3x3 convolution is an implicit GEMM on the 5th-gen tensor cores over the resident image: an M = 128 tile is
//     128 consecutive halo rows, the nine taps are nine descriptors whose start address is advanced by kh*(W+2) + kw
//     rows (conv3x3_tc_halo_kernel's trick), and the WHOLE output image stays in TENSOR MEMORY as fp32
//     (tiles x C columns: 448 of 512 columns at 48 x 72 x 16) — it never goes through shared memory;
//   * GroupNorm needs whole-image statistics before it can normalise, so each conv output is read from TMEM twice:
//     once as its tiles complete (statistics of the bf16-rounded values + the copy of y the backward pass needs, to
//     global), once to apply normalise + SiLU.  After conv1 that second read writes a1 as the NEXT conv's operand, i.e.
//     straight into the halo image in shared memory (ring re-zeroed) — conv1 -> GN -> SiLU -> conv2 without the
//     activation ever leaving the SM.  After conv2 it writes a2 as the linear image the SE / gate half of the
//     tail (tail_finish_from_pool, shared with convblock_tail_fwd_kernel) works on.
//   * replaces 4 launches (conv, gn_silu_img_fwd, conv, convblock_tail_fwd) and their 8 tensor passes over HBM by one
//     launch and 5.4 (x in; y1, a1, y2, out + 13 B / pixel of gate maps out — all of which the backward needs).
//
// Thread roles: warps 0-15 read TMEM (warp w: lane quarter w & 3, tiles (w >> 2) + 4 i) and do all the tail work;
// warp 16 issues the TMA loads and the MMAs (one elected thread).  Results are bit-compatible with the 4-kernel path: the
// statistics are taken from the bf16-rounded conv outputs, the activations use the same silu_raw / rounding.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kBlkThreads = 544;

// -DPCM_BLK_PROFILE: thread 0 of CTA 0 accumulates clock64 deltas per phase over its images and prints them at exit
#ifdef PCM_BLK_PROFILE
#define BLK_PROF_DECL long long prof_t[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; long long prof_c = clock64(); const bool prof_on = blockIdx.x == 0 && threadIdx.x == 0
#define BLK_PROF(i) do { if (prof_on) { const long long c_ = clock64(); prof_t[i] += c_ - prof_c; prof_c = c_; } } while (0)
#define BLK_PROF_PRINT() do { if (prof_on) printf("convblock_fwd_tc[%d,%d,%d,%d] cycles: prologue %lld | xwait+conv1+stats %lld  coef+zero %lld  a1 %lld  conv2+stats %lld  coef %lld  a2+pool %lld  tail %lld\n", p.H, p.W, p.Cin, C, prof_t[0], prof_t[1], prof_t[2], prof_t[3], prof_t[4], prof_t[5], prof_t[6], prof_t[7]); } while (0)
#else
#define BLK_PROF_DECL
#define BLK_PROF(i)
#define BLK_PROF_PRINT()
#endif

// -DPCM_TAIL_PROFILE: thread 0 of CTA 0 of convblock_tail_bwd records clock64 at every phase boundary (with a CTA
// barrier, so a phase ends when its last warp does) and prints the cycles per phase (measurement builds only)
#ifdef PCM_TAIL_PROFILE
#define TPROF_DECL long long tp_t[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; long long tp_c = clock64(); const bool tp_on = blockIdx.x == 0 && threadIdx.x == 0
#define TPROF(i) do { __syncthreads(); if (tp_on) { const long long c_ = clock64(); tp_t[i] += c_ - tp_c; tp_c = c_; } } while (0)
#define TPROF_PRINT() do { if (tp_on) printf("convblock_tail_bwd[%d,%d,%d] x%d thr, cycles: prologue %lld | passA %lld | gate wgrad %lld | gate dgrad %lld | coef+xwait %lld | passB %lld | SE %lld | pass1 %lld | stats %lld | pass2 %lld\n", H, W, C, NT, tp_t[0], tp_t[1], tp_t[2], tp_t[3], tp_t[4], tp_t[5], tp_t[6], tp_t[7], tp_t[8], tp_t[9]); } while (0)
#else
#define TPROF_DECL
#define TPROF(i)
#define TPROF_PRINT()
#endif

struct BlockFwdParams {
  int N, H, W, Cin, C, Cr;           // Cin = padded input channels (16 / 32 / 64), C = output channels
  int Wp, M, tiles, P;               // Wp = W + 2, M = H * Wp halo-row extent of the output, tiles = ceil(M / 128)
  int rows_a;                        // halo rows addressable in the image region: tiles*128 + 2*Wp + 2
  int box_h, nbox;                   // the input arrives in nbox TMA boxes of box_h halo rows each
  uint32_t tmem_cols;
  uint32_t off_w1, off_w2, off_cm0, off_cm1, off_gate, off_part, off_fl, off_bar;   // shared-memory layout (region 0 = image)
  float eps;
};

// byte offset of 16-byte chunk `chunk` of halo row `row` in a 1024-byte aligned K-major tile with rows of rb bytes
// (rb = 32 / 64 / 128 with the swizzle of the same width: chunk bits ^= row bits, a function of the address)
__device__ __forceinline__ uint32_t swz(uint32_t row, uint32_t chunk, uint32_t rb) {
  const uint32_t off = row * rb + (chunk << 4);
  return off ^ (((off >> 7) & ((rb >> 4) - 1u)) << 4);
}

__device__ __forceinline__ void issue_conv_tiles(uint32_t tmem_base, uint32_t a_base, uint32_t b_base, int tiles, int Wp,
                                                 int Cin, int C, uint64_t* tile_done) {
  using namespace tc;
  const uint32_t rb = (uint32_t)Cin * 2u;
  const uint32_t idesc = make_idesc_bf16(128, C, 0, 0);
  const uint32_t ltype = layout_type_for_row_bytes((int)rb);
  const uint64_t adesc0 = make_smem_desc(a_base, 16, 8 * rb, ltype);
  const uint64_t bdesc0 = make_smem_desc(b_base, 16, 8 * rb, ltype);
  const int ksteps = Cin / 16;
  uint32_t a_off[9], b_off[9];
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    a_off[tap] = ((uint32_t)((tap / 3) * Wp + (tap % 3)) * rb) >> 4;
    b_off[tap] = ((uint32_t)tap * (uint32_t)C * rb) >> 4;
  }
  const uint32_t tile_step = (128u * rb) >> 4;
  for (int t = 0; t < tiles; ++t) {
    const uint32_t d = tmem_base + (uint32_t)(t * C);
    const uint64_t ab = adesc0 + (uint64_t)(t * tile_step);
#pragma unroll
    for (int tap = 0; tap < 9; ++tap)
      for (int k = 0; k < ksteps; ++k)
        umma_bf16(d, ab + (uint64_t)(a_off[tap] + 2 * k), bdesc0 + (uint64_t)(b_off[tap] + 2 * k), idesc, (tap | k) != 0);
    umma_commit(&tile_done[t]);
  }
}

// packed weights [9][C][Cin] bf16 (global) -> swizzled K-major tiles in shared memory
__device__ __forceinline__ void stage_weights(uint8_t* dst, const __nv_bfloat16* __restrict__ wk, int C, int Cin) {
  const int cpr = Cin / 8;                                   // 16-byte chunks per row
  const int total = 9 * C * cpr;
  const uint32_t rb = (uint32_t)Cin * 2u;
  for (int i = threadIdx.x; i < total; i += NT) {
    const int chunk = i % cpr, row = i / cpr;                // row = tap*C + co
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(wk) + i);
    *reinterpret_cast<uint4*>(dst + swz((uint32_t)row, (uint32_t)chunk, rb)) = v;
  }
}

template <int C>       // output channels (16 / 32 / 64): GroupNorm group of a channel and the TMEM column loops are compile-time
__global__ void __launch_bounds__(kBlkThreads, 1)
convblock_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __nv_bfloat16* __restrict__ wk1,
                        const __nv_bfloat16* __restrict__ wk2, const float* __restrict__ g1, const float* __restrict__ b1,
                        const float* __restrict__ g2, const float* __restrict__ b2, const float* __restrict__ sw1,
                        const float* __restrict__ sw2, const float* __restrict__ wsp, __nv_bfloat16* __restrict__ y1,
                        __nv_bfloat16* __restrict__ a1, __nv_bfloat16* __restrict__ y2, float* __restrict__ stats1,
                        float* __restrict__ stats2, float* __restrict__ pool_g, float* __restrict__ se_g,
                        float* __restrict__ hid_g, float* __restrict__ maps, uint8_t* __restrict__ ties,
                        __nv_bfloat16* __restrict__ out, unsigned int* __restrict__ err, const BlockFwdParams p) {
  using namespace tc;
  typedef __nv_bfloat16 T;
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sImg = smem;                                    // halo image of x, then of a1, then the linear image of a2
  uint8_t* sW1 = smem + p.off_w1;
  uint8_t* sW2 = smem + p.off_w2;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bar);
  uint64_t* tile_done = bars;                              // [32]
  uint64_t* xfull = bars + 32;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 33);
  TailSmem L;
  L.img = 0; L.cm0 = p.off_cm0; L.cm1 = p.off_cm1; L.dq = 0; L.gate = p.off_gate; L.dm = 0; L.scr = 0; L.bar = p.off_bar;
  L.part = p.off_part; L.fl = p.off_fl; L.total = 0;
  const TailPtrs sp = tail_ptrs(smem, L, p.C);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Cin = p.Cin, H = p.H, W = p.W, Wp = p.Wp, P = p.P;
  constexpr int cg = C / kGroups;
  const uint32_t rb1 = (uint32_t)Cin * 2u, rb2 = (uint32_t)C * 2u;
  BLK_PROF_DECL;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    for (int i = 0; i < p.tiles; ++i) mbar_init(&tile_done[i], 1);
    mbar_init(xfull, 1);
    mbar_fence_init();
  }
  if (warp == 16) tmem_alloc(tmem_slot, p.tmem_cols);
  {  // zero the gate planes once: their rings must read as zero, the interiors are rewritten for every image
    float4* z4 = reinterpret_cast<float4*>(smem + p.off_cm0);              // cm0 and cm1 are contiguous
    const int Wpl = plane_wp(W);
    for (int i = threadIdx.x; i < 2 * (H + 6) * Wpl / 4; i += NT) z4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();                                                  // global memory from here on
  const uint32_t tmem_base = *tmem_slot;

  // the weights of both convolutions, the gate stencil and the squeeze-excite matrices: staged ONCE per CTA, resident for
  // every image this (persistent) CTA processes
  stage_weights(sW1, wk1, C, Cin);
  stage_weights(sW2, wk2, C, C);
  load_gate_weights(wsp, sp);
  for (int i = threadIdx.x; i < C * p.Cr; i += NT) { sp.sw1[i] = __ldg(sw1 + i); sp.sw2[i] = __ldg(sw2 + i); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const int quarter = warp & 3, grp = warp >> 2;              // TMEM readers: warps 0..15
  BLK_PROF(0);

  uint32_t xphase = 0;
  for (int n = blockIdx.x; n < p.N; n += gridDim.x, xphase ^= 1u) {
  if (warp == 16) {
    if (elect_one()) {
      // the whole halo image of image n: nbox boxes {Cin, W+2, box_h, 1} at (0, -1, -1 + b*box_h, n).  The image region
      // is free: the previous image's tail (which read it as the linear a2 image) ended with a CTA barrier.  (Plain
      // 16-byte loads by all threads were measured slower than these boxes: 43k vs 35k cycles per image for load + conv1.)
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(xfull, (uint32_t)(p.nbox * p.box_h * Wp) * rb1);
      for (int b = 0; b < p.nbox; ++b)
        tma_load_4d(sImg + (size_t)b * p.box_h * Wp * rb1, &tmX, xfull, 0, -1, -1 + b * p.box_h, n);
    }
    __syncwarp();
  }

  // ---------------- conv1 on the tensor cores ----------------
  if (warp == 16) {
    if (elect_one()) {
      if (mbar_wait(xfull, xphase, err)) {
        tc_fence_after();
        issue_conv_tiles(tmem_base, smem_u32(sImg), smem_u32(sW1), p.tiles, Wp, Cin, C, tile_done);
      }
    }
    __syncwarp();
  }

  // Pass over this warp's tiles of a conv output in TMEM: statistics of the bf16-rounded values per GroupNorm group and
  // the copy of y for the backward pass.  phase = mbarrier parity of this convolution.
  auto stats_pass = [&](uint32_t phase, __nv_bfloat16* __restrict__ ydst, float* __restrict__ stats_out) {
    float gs[kGroups], gq[kGroups];
#pragma unroll
    for (int g = 0; g < kGroups; ++g) gs[g] = gq[g] = 0.f;
    bool ok = true;
    if (warp < 16) {
      for (int t = grp; t < p.tiles; t += 4) {
        const int m = t * 128 + quarter * 32 + lane;
        const int h = m / Wp, w = m - h * Wp;
        const bool valid = w < W && h < H;
        ok = mbar_wait(&tile_done[t], phase, err) && ok;
        ok = __all_sync(0xffffffffu, ok);
        if (!ok) break;
        tc_fence_after();
        const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(t * C);
#pragma unroll
        for (int c0 = 0; c0 < C; c0 += 16) {
          float v[16];
          tmem_ld16(t_addr + c0, v);
          if (valid) {
            const uint4 u0 = pack8_rounded(v), u1 = pack8_rounded(v + 8);          // v now holds the rounded values
            uint4* dp = reinterpret_cast<uint4*>(ydst + ((size_t)n * P + h * W + w) * C + c0);
            dp[0] = u0; dp[1] = u1;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int g = (c0 + j) / cg;                     // compile-time: both loops are unrolled
              gs[g] += v[j];
              gq[g] = fmaf(v[j], v[j], gq[g]);
            }
          }
        }
      }
    }
    // CTA reduction of the 16 group sums: lanes, then warps through sp.part
#pragma unroll
    for (int g = 0; g < kGroups; ++g) { gs[g] = warp_sum(gs[g]); gq[g] = warp_sum(gq[g]); }
    if (lane == 0) {
#pragma unroll
      for (int g = 0; g < kGroups; ++g) { sp.part[warp * 16 + 2 * g] = gs[g]; sp.part[warp * 16 + 2 * g + 1] = gq[g]; }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x < kGroups) {
      const int g = threadIdx.x;
      float S = 0.f, Q = 0.f;
      for (int wv = 0; wv < NT / 32; ++wv) { S += sp.part[wv * 16 + 2 * g]; Q += sp.part[wv * 16 + 2 * g + 1]; }
      group_mu_rs(S, Q, (float)cg * (float)P, p.eps, sp.mu[g], sp.rs[g]);
      stats_out[(size_t)n * kGroups * 2 + 2 * g] = S;
      stats_out[(size_t)n * kGroups * 2 + 2 * g + 1] = Q;
    }
    __syncthreads();
    return ok;
  };
  // per-channel GroupNorm + SiLU coefficients into shared memory (sp.ca = za, sp.cb_ = zb)
  auto publish_coef = [&](const float* __restrict__ gamma, const float* __restrict__ beta) {
    for (int c = threadIdx.x; c < C; c += NT) {
      const int g = c / cg;
      float za, zb;
      gn_coef(__ldg(gamma + c), __ldg(beta + c), sp.mu[g], sp.rs[g], za, zb);
      sp.ca[c] = za; sp.cb_[c] = zb;
    }
  };

  bool ok = stats_pass(0u, y1, stats1);
  BLK_PROF(1);
  publish_coef(g1, b1);
  // conv1 is complete (every tile barrier was waited for): the image region becomes a1's halo image — all zero first
  {
    const uint32_t bytes = (uint32_t)p.rows_a * rb2;
    for (uint32_t o = threadIdx.x * 16u; o < bytes; o += NT * 16u) *reinterpret_cast<uint4*>(sImg + o) = make_uint4(0, 0, 0, 0);
  }
  __syncthreads();
  BLK_PROF(2);
  // ---------------- a1 = silu(GN1(y1)): TMEM -> halo image (the next conv's operand) + global (saved for backward) ----
  if (warp < 16 && ok) {
    for (int t = grp; t < p.tiles; t += 4) {
      const int m = t * 128 + quarter * 32 + lane;
      const int h = m / Wp, w = m - h * Wp;
      const bool valid = w < W && h < H;
      const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(t * C);
      const uint32_t arow = (uint32_t)((h + 1) * Wp + w + 1);
#pragma unroll
      for (int c0 = 0; c0 < C; c0 += 16) {
        float v[16];
        tmem_ld16(t_addr + c0, v);
        if (valid) {
          pack8_rounded(v); pack8_rounded(v + 8);                   // the bf16 values y1 holds in memory
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = silu_raw<T>(v[j], sp.ca[c0 + j], sp.cb_[c0 + j]);
          const uint4 u0 = pack8_rounded(v), u1 = pack8_rounded(v + 8);
          *reinterpret_cast<uint4*>(sImg + swz(arow, (uint32_t)(c0 >> 3), rb2)) = u0;
          *reinterpret_cast<uint4*>(sImg + swz(arow, (uint32_t)(c0 >> 3) + 1u, rb2)) = u1;
          uint4* dp = reinterpret_cast<uint4*>(a1 + ((size_t)n * P + h * W + w) * C + c0);
          dp[0] = u0; dp[1] = u1;
        }
      }
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes of a1 -> the tensor core's async proxy
  tc_fence_before();
  __syncthreads();
  BLK_PROF(3);
  // ---------------- conv2 ----------------
  if (warp == 16) {
    if (elect_one()) {
      tc_fence_after();
      issue_conv_tiles(tmem_base, smem_u32(sImg), smem_u32(sW2), p.tiles, Wp, C, C, tile_done);
    }
    __syncwarp();
  }
  ok = stats_pass(1u, y2, stats2) && ok;
  BLK_PROF(4);
  publish_coef(g2, b2);
  __syncthreads();
  BLK_PROF(5);
  // ---------------- a2 = silu(GN2(y2)): TMEM -> linear image [pixel][C] (conv2 is complete: the region is free) + SE pool sums
  T* s_img = reinterpret_cast<T*>(sImg);
#pragma unroll
  for (int c0 = 0; c0 < C; c0 += 16) {
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.f;
    if (warp < 16 && ok) {
      for (int t = grp; t < p.tiles; t += 4) {
        const int m = t * 128 + quarter * 32 + lane;
        const int h = m / Wp, w = m - h * Wp;
        const bool valid = w < W && h < H;
        const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(t * C);
        float v[16];
        tmem_ld16(t_addr + c0, v);
        if (valid) {
          pack8_rounded(v); pack8_rounded(v + 8);                 // the bf16 values y2 holds in memory
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = silu_raw<T>(v[j], sp.ca[c0 + j], sp.cb_[c0 + j]);
          const uint4 u0 = pack8_rounded(v), u1 = pack8_rounded(v + 8);
          uint4* dp = reinterpret_cast<uint4*>(s_img + (size_t)(h * W + w) * C + c0);
          dp[0] = u0; dp[1] = u1;
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] += v[j];
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = warp_sum(acc[j]);
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < 16; ++j) sp.part[warp * C + c0 + j] = acc[j];      // one row of C partial sums per warp
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += NT) {
    float t = 0.f;
    for (int wv = 0; wv < NT / 32; ++wv) t += sp.part[wv * C + c];
    sp.ch2[c] = t;
  }
  tc_fence_before();
  __syncthreads();
  BLK_PROF(6);
  // ---------------- SE, channel maps, 7x7 gate, out = a2*se*gate ----------------
  tail_finish_from_pool<T>(s_img, smem, L, sp, n, H, W, C, p.Cr, pool_g, se_g, hid_g, maps, ties,
                           out + (size_t)n * P * C);
  __syncthreads();                  // the image region and the gate planes are free for the next image
  BLK_PROF(7);
  }   // images
  BLK_PROF_PRINT();
  tc_fence_before();
  __syncthreads();
  if (warp == 16) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// Per-thread channel coefficients of a GroupNorm + SiLU backward (fixed channel block cb), as pairs:
//   xhat = xa*x + xb ;  z = za*x + zb  (za, zb exactly as the forward kernel derives them: gn_coef)
struct GnCoef {
  float2 xa[4], xb[4], za[4], zb[4], gm[4];
};
__device__ __forceinline__ void gn_coef_load(GnCoef& k, const float* __restrict__ gamma, const float* __restrict__ beta,
                                             const TailPtrs& sp, int cb, int cg) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = cb * 8 + j, g = c / cg;
    const float gmj = __ldg(gamma + c), xaj = sp.rs[g], xbj = -sp.mu[g] * sp.rs[g];
    float zaj, zbj;
    gn_coef(gmj, __ldg(beta + c), sp.mu[g], sp.rs[g], zaj, zbj);
    if (j & 1) { k.gm[j >> 1].y = gmj; k.xa[j >> 1].y = xaj; k.xb[j >> 1].y = xbj; k.za[j >> 1].y = zaj; k.zb[j >> 1].y = zbj; }
    else { k.gm[j >> 1].x = gmj; k.xa[j >> 1].x = xaj; k.xb[j >> 1].x = xbj; k.za[j >> 1].x = zaj; k.zb[j >> 1].x = zbj; }
  }
}
// pass 2: dx = rs*(dxh - m1 - xh*m2) = xa*dxh + k2*x + k1 with k1 = -(rs*m1 + rs*m2*xb), k2 = -rs*m2*xa
__device__ __forceinline__ void gn_bwd_pass2_coef(const GnCoef& k, const TailPtrs& sp, int cb, int cg, float2 (&k1)[4],
                                                  float2 (&k2)[4]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int g0 = (cb * 8 + 2 * j) / cg, g1 = (cb * 8 + 2 * j + 1) / cg;
    k1[j].x = -k.xa[j].x * (sp.m1[g0] + sp.m2[g0] * k.xb[j].x);
    k1[j].y = -k.xa[j].y * (sp.m1[g1] + sp.m2[g1] * k.xb[j].y);
    k2[j].x = -k.xa[j].x * sp.m2[g0] * k.xa[j].x;
    k2[j].y = -k.xa[j].y * sp.m2[g1] * k.xa[j].y;
  }
}
// one vector of pass 1: d (gradient reaching a) -> d = gamma * dz (dxhat), r0 += dz*xhat, r1 += dz
template <typename T>
__device__ __forceinline__ void gn_silu_bwd_vec(float2 (&d)[4], const float2 (&t)[4], const GnCoef& k, float2 (&r0)[4],
                                                float2 (&r1)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 xh = fma2(k.xa[i], t[i], k.xb[i]);
    const float2 z = fma2(k.za[i], t[i], k.zb[i]);
    const float2 sg = sigmoid2<T>(z);
    const float2 om = fma2(sg, bc2(-1.f), bc2(1.f));                       // 1 - sg
    const float2 dz = mul2(mul2(d[i], sg), fma2(z, om, bc2(1.f)));         // d * sg * (1 + z*(1 - sg))
    r0[i] = fma2(dz, xh, r0[i]);
    r1[i] = add2(r1[i], dz);
    d[i] = mul2(dz, k.gm[i]);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// backward of tail 1: da (gradient w.r.t. a = silu(GN(x))) -> dx ; dgamma, dbeta accumulate
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kFT, 1)
gn_silu_img_bwd_kernel(const T* __restrict__ da, const T* __restrict__ x, const float* __restrict__ stats,
                       const float* __restrict__ gamma, const float* __restrict__ beta, T* __restrict__ dx,
                       float* __restrict__ dgamma, float* __restrict__ dbeta, int H, int W, int C, float eps, int scr) {
  pdl_launch_dependents();
  extern __shared__ __align__(16) uint8_t smem[];
  const int n = blockIdx.x, P = H * W, cv = C / 8, cg = C / kGroups, nvec = P * cv;
  const TailSmem L = tail_smem_layout(H, W, C, (int)sizeof(T), 0, 1, scr);
  T* s_img = reinterpret_cast<T*>(smem + L.img);
  const TailPtrs sp = tail_ptrs(smem, L, C);
  const T* xn = x + (size_t)n * P * C;
  const T* dan = da + (size_t)n * P * C;
  T* dxn = dx + (size_t)n * P * C;
  T* sxn = scr ? reinterpret_cast<T*>(smem + L.scr) : dxn;     // pass 1 -> pass 2 scratch: shared memory when it fits
  const int cb = threadIdx.x & (cv - 1);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L.bar);
  pdl_wait();                                                  // global memory from here on
  if (threadIdx.x == 0) image_copy_start(bar, s_img, xn, (uint32_t)((size_t)P * C * sizeof(T)));
  group_mu_rs_from_stats(stats + (size_t)n * kGroups * 2, cg, P, eps, sp);
  __syncthreads();
  GnCoef k;
  gn_coef_load(k, gamma, beta, sp, cb, cg);
  float2 r0[4], r1[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) r0[j] = r1[j] = make_float2(0.f, 0.f);
  // pass 1: dxhat (stored to dx as scratch) and the reductions; x is consumed as its pieces land
  constexpr int KB = raw_batch<T>();
  PieceWalk land(bar);
  for (int v0 = threadIdx.x; v0 < nvec; v0 += NT * KB) {
    Raw8<T> raw[KB];
#pragma unroll
    for (int kk = 0; kk < KB; ++kk) raw[kk] = ld_raw<true>(dan + (size_t)min(v0 + kk * NT, nvec - 1) * 8);
    land.need((uint32_t)min(v0 + (KB - 1) * NT + 1, nvec) * 8u * (uint32_t)sizeof(T));
#pragma unroll
    for (int kk = 0; kk < KB; ++kk) {
      const int v = v0 + kk * NT;
      if (v < nvec) {
        float2 t[4], d[4];
        load8_rw(s_img + (size_t)v * 8, t);
        unpack8(raw[kk], d);
        gn_silu_bwd_vec<T>(d, t, k, r0, r1);
        store8(sxn + (size_t)v * 8, d);
      }
    }
  }
  chan_put(r0, sp.part, 0, cb, cv, C);
  chan_put(r1, sp.part, 1, cb, cv, C);
  chan_finish(sp.part, 2, sp.ch0, C);
  if (threadIdx.x < kGroups) {
    // group means of dxhat = gamma*dz and of dxhat*xhat, from the per-channel sums of dz and dz*xhat
    const int g = threadIdx.x;
    float a = 0.f, b = 0.f;
    for (int kk = 0; kk < cg; ++kk) {
      const float gmc = __ldg(gamma + g * cg + kk);
      a = fmaf(gmc, sp.ch1[g * cg + kk], a);
      b = fmaf(gmc, sp.ch0[g * cg + kk], b);
    }
    const float cnt = (float)cg * (float)P;
    sp.m1[g] = a / cnt;
    sp.m2[g] = b / cnt;
  }
  for (int c = threadIdx.x; c < C; c += NT) {
    atomicAdd(dgamma + c, sp.ch0[c]);
    atomicAdd(dbeta + c, sp.ch1[c]);
  }
  __syncthreads();
  // dx = rs*(dxh - m1 - xh*m2) = rs*dxh - (rs*m1 + rs*m2*xb) - (rs*m2*xa)*x
  float2 k1[4], k2[4];
  gn_bwd_pass2_coef(k, sp, cb, cg, k1, k2);
  // pass 2: every thread re-reads exactly the vectors it wrote in pass 1
  for (int v0 = threadIdx.x; v0 < nvec; v0 += NT * KB) {
    Raw8<T> raw[KB];
#pragma unroll
    for (int kk = 0; kk < KB; ++kk) raw[kk] = ld_raw<false>(sxn + (size_t)min(v0 + kk * NT, nvec - 1) * 8);
#pragma unroll
    for (int kk = 0; kk < KB; ++kk) {
      const int v = v0 + kk * NT;
      if (v < nvec) {
        float2 t[4], d[4];
        load8_rw(s_img + (size_t)v * 8, t);
        unpack8(raw[kk], d);
#pragma unroll
        for (int j = 0; j < 4; ++j) d[j] = fma2(k.xa[j], d[j], fma2(k2[j], t[j], k1[j]));
        store8(dxn + (size_t)v * 8, d);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// backward of tail 2: dout -> dx (gradient w.r.t. the conv2 output), all parameter gradients of GN2 / SE / gate.
// The forward tail saved the channel maps, the gate and the tie counts, and the block output `out` = a*se*gate is
// alive anyway (the next layer keeps it), so the gradient reaching the gate's pre-activation needs no activation:
//     dq[p] = gate'(q) * sum_c dout*u = (1 - gate[p]) * sum_c dout[p,c]*out[p,c].
// That first pass streams dout and out from global memory while the bulk-copy engine brings x into shared memory.
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kFT, 1)
convblock_tail_bwd_kernel(const T* __restrict__ dout, const T* __restrict__ x, const T* __restrict__ out,
                          const float* __restrict__ stats, const float* __restrict__ gamma,
                          const float* __restrict__ beta, const float* __restrict__ w1, const float* __restrict__ w2,
                          const float* __restrict__ wsp, const float* __restrict__ pool_g,
                          const float* __restrict__ se_g, const float* __restrict__ hid_g,
                          const float* __restrict__ maps, const uint8_t* __restrict__ ties, T* __restrict__ dx,
                          float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dw1,
                          float* __restrict__ dw2, float* __restrict__ dwsp, float* __restrict__ dq_out,
                          const float* __restrict__ sdot, int H, int W, int C, int Cr, float eps, int scr) {
  // sdot != NULL: sdot[n][p] = sum_c dout[p,c]*out[p,c] was already formed by the kernel that produced dout
  // (pcm_maxpool2_bwd_skip_dot, which holds both operands in registers) — pass A then reads 4 bytes per pixel instead of
  // streaming dout and out (2 x 2C bytes per pixel: a quarter of this kernel's cycles at 48x72x16).
  // dq_out != NULL: the gate's pre-activation gradient dq [N][H*W] is also written to global memory and the gate-weight
  // gradient (98 sums over the image: 11 % of this kernel's instructions on 14 of its 16 warps, 13 % of its samples) is
  // left to pcm_gate_wgrad on the side stream — a parameter gradient nothing in the backward chain waits for.
  pdl_launch_dependents();
  extern __shared__ __align__(16) uint8_t smem[];
  const int n = blockIdx.x, P = H * W, cv = C / 8, cg = C / kGroups, nvec = P * cv, Wp = plane_wp(W);
  const int cvs = __ffs(cv) - 1;
  const int nround = (nvec + NT - 1) / NT;
  const TailSmem L = tail_smem_layout(H, W, C, (int)sizeof(T), 1, 1, scr);
  T* s_img = reinterpret_cast<T*>(smem + L.img);
  float2* cm = reinterpret_cast<float2*>(smem + L.cm0);          // (mean, max) pairs: the cm0 | cm1 regions as one plane
  float* s_dq = reinterpret_cast<float*>(smem + L.dq);
  float* s_gate = reinterpret_cast<float*>(smem + L.gate);
  float2* s_dm = reinterpret_cast<float2*>(smem + L.dm);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L.bar);
  const TailPtrs sp = tail_ptrs(smem, L, C);
  const T* xn = x + (size_t)n * P * C;
  const T* don = dout + (size_t)n * P * C;
  const T* outn = out + (size_t)n * P * C;
  const float* mp = maps + (size_t)n * 3 * P;
  const uint8_t* tp = ties + (size_t)n * P;
  T* dxn = dx + (size_t)n * P * C;
  const int cb = threadIdx.x & (cv - 1);
  const float invP = 1.f / (float)P;
  TPROF_DECL;

  {
    float4* z4 = reinterpret_cast<float4*>(cm);                                 // cm0 | cm1 | dq are contiguous
    for (int i = threadIdx.x; i < 3 * (H + 6) * Wp / 4; i += NT) z4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  pdl_wait();                                                                   // global memory from here on
  if (threadIdx.x == 0) image_copy_start(bar, s_img, xn, (uint32_t)((size_t)P * C * sizeof(T)));
  load_gate_weights(wsp, sp);
  for (int i = threadIdx.x; i < C * Cr; i += NT) { sp.sw1[i] = __ldg(w1 + i); sp.sw2[i] = __ldg(w2 + i); }
  for (int c = threadIdx.x; c < C; c += NT) {
    sp.se[c] = __ldg(se_g + (size_t)n * C + c);
    sp.pool[c] = __ldg(pool_g + (size_t)n * C + c);
  }
  for (int j = threadIdx.x; j < Cr; j += NT) sp.hid[j] = __ldg(hid_g + (size_t)n * Cr + j);
  group_mu_rs_from_stats(stats + (size_t)n * kGroups * 2, cg, P, eps, sp);
  __syncthreads();
  TPROF(0);

  // ---- pass A (global operands only): the saved maps go into the padded planes and dq = (1 - gate) * sum_c dout*out.
  // Nothing here depends on shared memory, so the loads of kBatch rounds are issued back to back (the loop is
  // bound by global-load latency otherwise: 16 warps per SM).
  {
    const float* sdn = sdot != nullptr ? sdot + (size_t)n * P : nullptr;
#pragma unroll 4
    for (int p = threadIdx.x; p < P; p += NT) {
      const int h = p / W, w = p - h * W, ip = (h + 3) * Wp + w + 4;
      const float gtp = __ldg(mp + 2 * P + p);
      cm[pswz2(ip)] = make_float2(__ldg(mp + p), __ldg(mp + P + p));
      s_gate[p] = gtp;
      if (sdn != nullptr) {
        const float dqv = __ldg(sdn + p) * (1.f - gtp);
        s_dq[pswz(ip)] = dqv;
        if (dq_out != nullptr) dq_out[(size_t)n * P + p] = dqv;
      }
    }
    constexpr int kBatch = 2 * raw_batch<T>();
    PixWalk pw(cvs, W);
    for (int r0 = sdn != nullptr ? nround : 0; r0 < nround; r0 += kBatch) {
      Raw8<T> dr[kBatch], orw[kBatch];
      float gt[kBatch];
      int ip[kBatch], pk[kBatch];
#pragma unroll
      for (int k = 0; k < kBatch; ++k) {
        const bool valid = pw.p < P;
        const size_t v = valid ? (size_t)pw.p * cv + cb : 0;
        dr[k] = ld_raw<true>(don + v * 8);
        orw[k] = ld_raw<true>(outn + v * 8);
        gt[k] = (valid && cb == 0) ? __ldg(mp + 2 * P + pw.p) : 1.f;
        ip[k] = (valid && cb == 0) ? pswz((pw.h + 3) * Wp + pw.w + 4) : -1;
        pk[k] = pw.p;
        pw.next(W);
      }
#pragma unroll
      for (int k = 0; k < kBatch; ++k) {
        float2 d[4], o[4];
        unpack8(dr[k], d);
        unpack8(orw[k], o);
        float2 a2 = mul2(d[0], o[0]);
#pragma unroll
        for (int j = 1; j < 4; ++j) a2 = fma2(d[j], o[j], a2);
        float acc = a2.x + a2.y;
        for (int off = 1; off < cv; off <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (ip[k] >= 0) {
          const float dqv = acc * (1.f - gt[k]);
          s_dq[ip[k]] = dqv;
          if (dq_out != nullptr) dq_out[(size_t)n * P + pk[k]] = dqv;
        }
      }
    }
  }
  __syncthreads();
  TPROF(1);
  // ---- dwsp[k][dy][dx] = sum_p dq[p] * cmap_k[p + (dy-3, dx-3)]: one warp per (kernel row dy, half of the 8-pixel
  // runs), lanes over the runs; the 7 dx sums of BOTH maps accumulate as pairs (broadcast dq x (mean, max)), combined
  // with shuffles (no atomics: every (half, k, dy, dx) has one owner; the two halves are added when dwsp is updated)
  if (dq_out == nullptr) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nrun = (W + 7) / 8;
    const int nitem = H * nrun, ihalf = (nitem + 1) >> 1;
    for (int combo = warp; combo < 14; combo += NT >> 5) {
      const int half = combo / 7, dy = combo - half * 7;
      const int i0 = half ? ihalf : 0, i1 = half ? nitem : ihalf;
      float2 a[7];
#pragma unroll
      for (int i = 0; i < 7; ++i) a[i] = make_float2(0.f, 0.f);
      for (int item = i0 + lane; item < i1; item += 32) {
        const int h = item / nrun, w0 = (item - h * nrun) * 8;
        const int id = (h + 3) * Wp + w0 + 4;
        const float4 d0 = *reinterpret_cast<const float4*>(s_dq + pswz(id)), d1 = *reinterpret_cast<const float4*>(s_dq + pswz(id + 4));
        const float dq8[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
        float2 m[16];
        load_pair_row16(cm, (h + dy) * Wp + w0, m);
#pragma unroll
        for (int dxx = 0; dxx < 7; ++dxx) {
#pragma unroll
          for (int i = 0; i < 8; ++i) a[dxx] = fma2(bc2(dq8[i]), m[i + dxx + 1], a[dxx]);
        }
      }
#pragma unroll
      for (int dxx = 0; dxx < 7; ++dxx) {
        const float t0 = warp_sum(a[dxx].x), t1 = warp_sum(a[dxx].y);
        if (lane == 0) {
          sp.dw[half * 98 + dy * 7 + dxx] = t0;
          sp.dw[half * 98 + 49 + dy * 7 + dxx] = t1;
        }
      }
    }
  }
  TPROF(2);
  // ---- gradient reaching (mean, max) through the transposed stencil (flipped weights)
  {
    const int nrun = (W + 7) / 8;
    const float invC = 1.f / (float)C;
    for (int item = threadIdx.x; item < H * nrun; item += NT) {
      const int h = item / nrun, w0 = (item - h * nrun) * 8;
      float2 q[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) q[i] = make_float2(0.f, 0.f);
      stencil_dq_run8(s_dq, h * Wp + w0, Wp, sp.wtf, q);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (w0 + i < W) {
          const int p = h * W + w0 + i;
          s_dm[p] = make_float2(q[i].x * invC, q[i].y / (float)max((int)__ldg(tp + p), 1));
        }
      }
    }
  }
  __syncthreads();
  if (dq_out == nullptr && threadIdx.x < 98) atomicAdd(dwsp + threadIdx.x, sp.dw[threadIdx.x] + sp.dw[98 + threadIdx.x]);
  TPROF(3);

  // per-thread channel coefficients (fixed channel block), as pairs: xhat = xa*x + xb ; z = za*x + zb
  GnCoef k;
  gn_coef_load(k, gamma, beta, sp, cb, cg);
  float2 sc[4];
  load_chan8(sp.se, cb, sc);
  image_copy_wait(bar, (uint32_t)((size_t)P * C * sizeof(T)));      // x is in shared memory from here on
  TPROF(4);

  // ---- pass B (first of the two activation evaluations): du = dout*gate + dmean + [u == max]*dmax/ties ; r = du*se (scratch in dx) ;
  // dse = sum_p du*a.  a and u are recomputed exactly as the forward kernel computed them (silu8 + rounding), so the
  // comparison against the saved maximum selects the same channels.
  {
    float2 acc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = make_float2(0.f, 0.f);
    constexpr int KB = raw_batch<T>();
    PixWalk pw(cvs, W);
    for (int r0 = 0; r0 < nround; r0 += KB) {
      Raw8<T> dr[KB];
      int pp[KB], ipk[KB];
#pragma unroll
      for (int kk = 0; kk < KB; ++kk) {
        pp[kk] = pw.p < P ? pw.p : -1;
        ipk[kk] = (pw.h + 3) * Wp + pw.w + 4;
        dr[kk] = ld_raw<true>(don + ((size_t)max(pp[kk], 0) * cv + cb) * 8);
        pw.next(W);
      }
#pragma unroll
      for (int kk = 0; kk < KB; ++kk) {
        if (pp[kk] >= 0) {
          const size_t v = (size_t)pp[kk] * cv + cb;
          float2 t[4], d[4];
          load8_rw(s_img + v * 8, t);
          unpack8(dr[kk], d);
          const float2 gt = bc2(s_gate[pp[kk]]);
          const float2 dm = s_dm[pp[kk]];
          const float mx = cm[pswz2(ipk[kk])].y;
          silu8<T>(t, k.za, k.zb);
          round8_to<T>(t);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 u = mul2(t[j], sc[j]);
            float2 du = fma2(d[j], gt, bc2(dm.x));
            if (u.x == mx) du.x += dm.y;
            if (u.y == mx) du.y += dm.y;
            acc[j] = fma2(du, t[j], acc[j]);
            d[j] = mul2(du, sc[j]);
          }
          store8(dxn + v * 8, d);
        }
      }
    }
    chan_put(acc, sp.part, 0, cb, cv, C);
  }
  chan_finish(sp.part, 1, sp.ch0, C);
  TPROF(5);
  // ---- SE backward (tiny): dpool, dw1, dw2
  for (int c = threadIdx.x; c < C; c += NT) {
    const float s = sp.se[c];
    sp.dpre2[c] = sp.ch0[c] * s * (1.f - s);
  }
  __syncthreads();
  for (int j = threadIdx.x >> 5; j < Cr; j += NT / 32) {
    float a = 0.f;
    for (int c = threadIdx.x & 31; c < C; c += 32) a = fmaf(sp.sw2[c * Cr + j], sp.dpre2[c], a);
    a = warp_sum(a);
    if ((threadIdx.x & 31) == 0) sp.dpre1[j] = sp.hid[j] > 0.f ? a : 0.f;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += NT) {
    float a = 0.f;
    for (int j = 0; j < Cr; ++j) a = fmaf(sp.sw1[j * C + c], sp.dpre1[j], a);
    sp.dpool[c] = a * invP;
  }
  // dw2[c][j] += dpre2[c]*hid[j] ; dw1[j][c] += dpre1[j]*pool[c]/P.  Every image of the launch adds into the same
  // 2*C*Cr values: as scalar atomics that was a quarter of this kernel at 6x9x128 (4096 per image); four consecutive
  // elements share c (dw2, Cr % 4 == 0) resp. j (dw1), so they go out as 16-byte vector reductions when the buffers allow.
  if ((Cr & 3) == 0 && ((reinterpret_cast<uintptr_t>(dw1) | reinterpret_cast<uintptr_t>(dw2)) & 15) == 0) {
    for (int i = threadIdx.x * 4; i < C * Cr; i += NT * 4) {
      {
        const int c = i / Cr, j = i - c * Cr;
        const float a = sp.dpre2[c];
        const float v0 = a * sp.hid[j], v1 = a * sp.hid[j + 1], v2 = a * sp.hid[j + 2], v3 = a * sp.hid[j + 3];
        if (a != 0.f && (v0 != 0.f || v1 != 0.f || v2 != 0.f || v3 != 0.f))
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dw2 + i), "f"(v0), "f"(v1), "f"(v2), "f"(v3) : "memory");
      }
      {
        const int j = i / C, c = i - j * C;
        const float a = sp.dpre1[j] * invP;
        if (a != 0.f)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dw1 + i), "f"(a * sp.pool[c]), "f"(a * sp.pool[c + 1]),
                       "f"(a * sp.pool[c + 2]), "f"(a * sp.pool[c + 3]) : "memory");
      }
    }
  } else {
    for (int i = threadIdx.x; i < C * Cr; i += NT) {
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
  }
  __syncthreads();
  TPROF(6);
  // ---- GroupNorm + SiLU backward, pass 1 (second activation evaluation): dxhat -> scratch, reductions.  With `scr` the
  // scratch is the shared-memory overlay of the gate planes (dead since pass B; the barriers of the SE backward separate
  // their last reader from the first scratch write) and pass 2 has no global load left.
  T* sxn = scr ? reinterpret_cast<T*>(smem + L.scr) : dxn;
  float2 dp[4], r0[4], r1[4];
  load_chan8(sp.dpool, cb, dp);
#pragma unroll
  for (int j = 0; j < 4; ++j) r0[j] = r1[j] = make_float2(0.f, 0.f);
  constexpr int KB = raw_batch<T>();
  for (int v0 = threadIdx.x; v0 < nvec; v0 += NT * KB) {
    Raw8<T> raw[KB];
#pragma unroll
    for (int kk = 0; kk < KB; ++kk) raw[kk] = ld_raw<false>(dxn + (size_t)min(v0 + kk * NT, nvec - 1) * 8);
#pragma unroll
    for (int kk = 0; kk < KB; ++kk) {
      const int v = v0 + kk * NT;
      if (v < nvec) {
        float2 t[4], d[4];
        load8_rw(s_img + (size_t)v * 8, t);
        unpack8(raw[kk], d);
#pragma unroll
        for (int j = 0; j < 4; ++j) d[j] = add2(d[j], dp[j]);
        gn_silu_bwd_vec<T>(d, t, k, r0, r1);
        store8(sxn + (size_t)v * 8, d);
      }
    }
  }
  chan_put(r0, sp.part, 0, cb, cv, C);
  chan_put(r1, sp.part, 1, cb, cv, C);
  chan_finish(sp.part, 2, sp.ch1, C);
  TPROF(7);
  if (threadIdx.x < kGroups) {
    const int g = threadIdx.x;
    float a = 0.f, b = 0.f;
    for (int kk = 0; kk < cg; ++kk) {
      const float gmc = __ldg(gamma + g * cg + kk);
      a = fmaf(gmc, sp.ch2[g * cg + kk], a);
      b = fmaf(gmc, sp.ch1[g * cg + kk], b);
    }
    const float cnt = (float)cg * (float)P;
    sp.m1[g] = a / cnt;
    sp.m2[g] = b / cnt;
  }
  for (int c = threadIdx.x; c < C; c += NT) {
    atomicAdd(dgamma + c, sp.ch1[c]);
    atomicAdd(dbeta + c, sp.ch2[c]);
  }
  __syncthreads();
  float2 k1[4], k2[4];
  gn_bwd_pass2_coef(k, sp, cb, cg, k1, k2);
  TPROF(8);
  for (int v0 = threadIdx.x; v0 < nvec; v0 += NT * KB) {
    Raw8<T> raw[KB];
#pragma unroll
    for (int kk = 0; kk < KB; ++kk) raw[kk] = ld_raw<false>(sxn + (size_t)min(v0 + kk * NT, nvec - 1) * 8);
#pragma unroll
    for (int kk = 0; kk < KB; ++kk) {
      const int v = v0 + kk * NT;
      if (v < nvec) {
        float2 t[4], d[4];
        load8_rw(s_img + (size_t)v * 8, t);
        unpack8(raw[kk], d);
#pragma unroll
        for (int j = 0; j < 4; ++j) d[j] = fma2(k.xa[j], d[j], fma2(k2[j], t[j], k1[j]));
        store8(dxn + (size_t)v * 8, d);
      }
    }
  }
  TPROF(9);
  TPROF_PRINT();
}

// Gate-weight gradient of SpatialGate's 7x7 conv (src/unet.py:24,28) for ALL images of a launch, off the critical path:
//   dwsp[k][dy][dx] += sum_n sum_p dq[n][p] * map_k[n][p + (dy-3, dx-3)]      (k = channel mean / max map)
// from the dq the backward tail wrote and the maps the forward tail saved.  One CTA walks images; the three planes of an
// image sit zero-padded in shared memory; warp w owns kernel row dy = w for both maps (the 8-pixel run of dq a lane
// loads serves 2 x 56 FMAs), accumulating in registers across its images; one shuffle reduction + 98 atomics per CTA.
constexpr int kGwThreads = 224;        // 7 warps = 7 kernel rows

__global__ void __launch_bounds__(kGwThreads)
gate_wgrad_kernel(const float* __restrict__ dq, const float* __restrict__ maps, float* __restrict__ dwsp, int N, int H, int W) {
  pdl_launch_dependents();
  extern __shared__ __align__(16) uint8_t smem[];
  const int P = H * W, Wp = plane_wp(W), Pp = (H + 6) * Wp;
  float* cm0 = reinterpret_cast<float*>(smem);
  float* cm1 = cm0 + Pp;
  float* s_dq = cm1 + Pp;
  for (int i = threadIdx.x; i < 3 * Pp / 4; i += kGwThreads) reinterpret_cast<float4*>(cm0)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  pdl_wait();
  const int dy = threadIdx.x >> 5, lane = threadIdx.x & 31, nrun = (W + 7) / 8;
  float a0[7], a1[7];
#pragma unroll
  for (int i = 0; i < 7; ++i) a0[i] = a1[i] = 0.f;
  for (int n = blockIdx.x; n < N; n += gridDim.x) {
    __syncthreads();                                           // zeros written / previous image consumed
    const float* mp = maps + (size_t)n * 3 * P;
    const float* dn = dq + (size_t)n * P;
#pragma unroll 4
    for (int p = threadIdx.x; p < P; p += kGwThreads) {
      const int h = p / W, w = p - h * W, ip = (h + 3) * Wp + w + 4;
      cm0[ip] = __ldg(mp + p);
      cm1[ip] = __ldg(mp + P + p);
      s_dq[ip] = __ldg(dn + p);
    }
    __syncthreads();
    for (int item = lane; item < H * nrun; item += 32) {
      const int h = item / nrun, w0 = (item - h * nrun) * 8;
      const float4* dr = reinterpret_cast<const float4*>(s_dq + (h + 3) * Wp + w0 + 4);
      const float4 d0 = dr[0], d1 = dr[1];
      const float dq8[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const float4* cr = reinterpret_cast<const float4*>((k ? cm1 : cm0) + (h + dy) * Wp + w0);
        const float4 c0 = cr[0], c1 = cr[1], c2 = cr[2], c3 = cr[3];
        const float m[16] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w, c2.x, c2.y, c2.z, c2.w, c3.x, c3.y, c3.z, c3.w};
#pragma unroll
        for (int dxx = 0; dxx < 7; ++dxx) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (k) a1[dxx] = fmaf(dq8[i], m[i + dxx + 1], a1[dxx]);
            else a0[dxx] = fmaf(dq8[i], m[i + dxx + 1], a0[dxx]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int dxx = 0; dxx < 7; ++dxx) {
    const float t0 = warp_sum(a0[dxx]), t1 = warp_sum(a1[dxx]);
    if (lane == 0) {
      atomicAdd(dwsp + dy * 7 + dxx, t0);
      atomicAdd(dwsp + 49 + dy * 7 + dxx, t1);
    }
  }
}

// Launch width of a per-image kernel.  tail_threads() sizes the CTA by the image; when that width needs more than one
// wave of CTAs over the SMs and half the width fits everything in fewer, the narrower CTA wins: measured on 384 images of
// 12x18x64, two waves of 256-thread CTAs (two per SM by registers) 56 us, one wave of 128-thread CTAs 42 us.  Cost model:
// waves x (c0 + vector rounds per thread), c0 = the per-image fixed cost in units of one round (fitted: ~6.75 rounds for
// the backward kernels, ~2.5 for the forward ones).  PCM_TAIL_WAVES=0 keeps tail_threads().
template <typename K>
static int tail_launch_threads(K kern, int N, int H, int W, int C, size_t smem, float c0) {
  static int sms = 0;
  const char* e = getenv("PCM_TAIL_WAVES");          // read per launch (host side, cheap): tests flip it inside one process
  const int on = e ? atoi(e) : 1;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  const int tmax = tail_threads(H, W, C);
  if (!on || tmax <= 128) return tmax;
  cudaFuncAttributes fa;
  if (cudaFuncGetAttributes(&fa, kern) != cudaSuccess || fa.numRegs <= 0) return tmax;
  const int nvec = H * W * (C / 8);
  int best = tmax;
  float best_cost = 0.f;
  for (int t = tmax; t >= tmax / 2 && t >= 128; t >>= 1) {
    long long per = (228 * 1024) / (long long)(smem + 1024);                 // CTAs per SM: shared memory (+1 KB reserved each),
    const long long by_regs = 65536 / ((long long)((fa.numRegs + 7) / 8 * 8) * t);   // registers, threads
    if (by_regs < per) per = by_regs;
    if (2048 / t < per) per = 2048 / t;
    if (per > 32) per = 32;
    if (per < 1) continue;
    const long long waves = (N + sms * per - 1) / (sms * per);
    const float cost = (float)waves * (c0 + (float)((nvec + t - 1) / t));
    if (t == tmax || cost < best_cost - 0.5f) { best = t; best_cost = cost; }
  }
  return best;
}

// Whether a backward tail launched with `threads` per CTA may take the larger shared-memory layout (`smem_scr`: scratch
// overlay, see tail_smem_layout) without losing resident CTAs per SM against `smem`: true when registers / threads, not
// shared memory, bound the residency either way (the level-1 images: one 512-thread CTA per SM by registers alone).
// PCM_TAIL_SCRATCH=0 keeps the scratch in global memory.
template <typename K>
static int tail_scratch_ok(K kern, int threads, size_t smem, size_t smem_scr) {
  const char* e = getenv("PCM_TAIL_SCRATCH");
  if (e != nullptr && atoi(e) == 0) return 0;
  if (smem_scr > 227 * 1024) return 0;
  cudaFuncAttributes fa;
  if (cudaFuncGetAttributes(&fa, kern) != cudaSuccess || fa.numRegs <= 0) return 0;
  long long other = 65536 / ((long long)((fa.numRegs + 7) / 8 * 8) * threads);
  if (2048 / threads < other) other = 2048 / threads;
  if (other > 32) other = 32;
  const long long a = (228 * 1024) / (long long)(smem + 1024), b = (228 * 1024) / (long long)(smem_scr + 1024);
  return (b < other ? b : other) >= (a < other ? a : other) ? 1 : 0;
}

static bool fused_shape_ok(int H, int W, int C, int Cr) {
  const int cv = C / 8;
  return C % 8 == 0 && C >= 8 && cv <= 32 && (cv & (cv - 1)) == 0 && Cr >= 1 && Cr <= 64 && Cr * 8 <= C && H >= 1 && W >= 1;
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

// ---- fused block forward: shared-memory layout + support check -----------------------------------------------------
static bool block_fwd_layout(int H, int W, int Cin, int C, int Cr, BlockFwdParams* p, size_t* smem_total) {
  if (!(C == 16 || C == 32 || C == 64) || !(Cin == 16 || Cin == 32 || Cin == 64)) return false;
  if (Cr < 1 || Cr > 64 || Cr * 8 > C || H < 1 || W < 1 || W + 2 > 256) return false;
  p->H = H; p->W = W; p->Cin = Cin; p->C = C; p->Cr = Cr;
  p->Wp = W + 2; p->P = H * W;
  p->M = H * p->Wp;
  p->tiles = (p->M + 127) / 128;
  if (p->tiles > 32 || p->tiles * C > 512) return false;                    // tile barriers / TMEM columns
  p->rows_a = p->tiles * 128 + 2 * p->Wp + 2;
  // TMA boxes of the input: whole halo rows, <= 32 KB each, and every box must start on a swizzle-atom boundary of the
  // image (8 rows x Cin*2 bytes; the swizzle is a function of the shared-memory address) — box_h*(W+2) a multiple of 8.
  // Rows past H+2 are out of bounds (zero filled); pick the box height that overshoots least.
  {
    const int max_h = (32 * 1024) / (p->Wp * Cin * 2);
    int best_h = 0, best_rows = 1 << 30;
    for (int bh = 1; bh <= max_h && bh <= 256; ++bh) {
      if ((bh * p->Wp) % 8 != 0) continue;
      const int rows = ((H + 2 + bh - 1) / bh) * bh;
      if (rows < best_rows || (rows == best_rows && bh > best_h)) { best_rows = rows; best_h = bh; }
    }
    if (best_h == 0) return false;
    p->box_h = best_h;
    p->nbox = best_rows / best_h;
  }
  const int box_h = p->box_h;
  uint32_t cols = 32;
  while (cols < (uint32_t)(p->tiles * C)) cols <<= 1;
  p->tmem_cols = cols;
  const size_t rb1 = (size_t)Cin * 2, rb2 = (size_t)C * 2;
  size_t rows = (size_t)p->rows_a;
  if ((size_t)p->nbox * box_h * p->Wp > rows) rows = (size_t)p->nbox * box_h * p->Wp;
  size_t img = rows * rb1;
  if ((size_t)p->rows_a * rb2 > img) img = (size_t)p->rows_a * rb2;
  if ((size_t)p->P * C * 2 > img) img = (size_t)p->P * C * 2;
  size_t off = (img + 1023) & ~(size_t)1023;
  auto take = [&](size_t bytes, size_t align) { off = (off + align - 1) & ~(align - 1); const size_t o = off; off += bytes; return (uint32_t)o; };
  p->off_w1 = take((size_t)9 * C * rb1, 1024);
  p->off_w2 = take((size_t)9 * C * rb2, 1024);
  const size_t Pp = (size_t)(H + 6) * plane_wp(W);
  p->off_cm0 = take(Pp * 4, 16);
  p->off_cm1 = take(Pp * 4, 16);
  if (p->off_cm1 != p->off_cm0 + Pp * 4) return false;                      // the kernel zeroes them as one range
  p->off_gate = take((size_t)p->P * 4, 16);
  p->off_part = take((size_t)2 * (kBlkThreads / 32) * C * 4 + 64, 16);       // chan_put slots; >= 17 warps x 16 floats
  p->off_fl = take((size_t)(11 * C + 128 + 32 + 112 + 112 + C * C / 4) * 4, 16);
  p->off_bar = take(34 * 8 + 16, 16);
  *smem_total = off + 1024;                                                  // + alignment slack of the dynamic base
  return *smem_total <= 227 * 1024;
}

extern "C" int pcm_convblock_fwd_tc_supported(int H, int W, int Cin, int C, int Cr) {
  BlockFwdParams p;
  size_t smem = 0;
  return block_fwd_layout(H, W, Cin, C, Cr, &p, &smem) ? 1 : 0;
}

extern "C" int pcm_convblock_fwd_tc(const void* x, const void* wk1, const void* wk2, const float* g1, const float* b1,
                                    const float* g2, const float* b2, const float* sw1, const float* sw2, const float* wsp,
                                    void* y1, void* a1, void* y2, float* stats1, float* stats2, float* pool, float* se,
                                    float* hid, float* maps, unsigned char* ties, void* out, int N, int H, int W, int Cin,
                                    int C, int Cr, float eps, pcm_stream_t s) {
  BlockFwdParams p;
  size_t smem = 0;
  PCM_REQUIRE(block_fwd_layout(H, W, Cin, C, Cr, &p, &smem), "convblock_fwd_tc: unsupported shape H=%d W=%d Cin=%d C=%d Cr=%d",
              H, W, Cin, C, Cr);
  PCM_REQUIRE(maps != nullptr && ties != nullptr, "convblock_fwd_tc: maps and ties are required (training forward)");
  PCM_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(wk1) | reinterpret_cast<uintptr_t>(wk2) |
                reinterpret_cast<uintptr_t>(y1) | reinterpret_cast<uintptr_t>(a1) | reinterpret_cast<uintptr_t>(y2) |
                reinterpret_cast<uintptr_t>(out)) & 15) == 0, "convblock_fwd_tc: pointers must be 16-byte aligned");
  if (N == 0) return PCM_OK;
  p.N = N; p.eps = eps;
  CUtensorMap tmX;
  {
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t strides[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
    uint32_t box[4] = {(uint32_t)Cin, (uint32_t)p.Wp, (uint32_t)p.box_h, 1};
    const int rc = make_tensor_map(&tmX, x, 4, dims, strides, box, Cin * 2);
    if (rc != PCM_OK) return rc;
  }
  unsigned int* err = tc_error_counter();
  PCM_REQUIRE(err != nullptr, "convblock_fwd_tc: could not allocate the error counter");
  auto kern = C == 16 ? convblock_fwd_tc_kernel<16> : C == 32 ? convblock_fwd_tc_kernel<32> : convblock_fwd_tc_kernel<64>;
  const int ki = C == 16 ? 0 : C == 32 ? 1 : 2;
  static size_t smem_set[3] = {0, 0, 0};
  if (smem > smem_set[ki]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("convblock_fwd_tc: smem attribute (%zu B): %s", smem, cudaGetErrorString(e)); return PCM_ERR_CUDA; }
    smem_set[ki] = smem;
  }
  typedef __nv_bfloat16 B16;
  static int num_sms = 0;
  if (num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (num_sms <= 0) num_sms = 148;
  }
  // persistent CTAs (one per SM): weights, gate stencil and SE matrices are staged once per CTA, not once per image
  const int grid = N < num_sms ? N : num_sms;
  pcm::launch(kern, grid, kBlkThreads, smem, (cudaStream_t)s, tmX, (const B16*)wk1, (const B16*)wk2, g1, b1, g2, b2, sw1, sw2, wsp,
              (B16*)y1, (B16*)a1, (B16*)y2, stats1, stats2, pool, se, hid, maps, ties, (B16*)out, err, p);
  return check_launch("convblock_fwd_tc");
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
      pcm::launch(convblock_tail_fwd_kernel<T, false>, N, tail_launch_threads(convblock_tail_fwd_kernel<T, false>, N, H, W, C, smem, 2.5f), smem, (cudaStream_t)s, 
          (const T*)x, gamma, beta, nullptr, nullptr, nullptr, stats, nullptr, nullptr, nullptr, nullptr, nullptr,
          (T*)y, H, W, C, 1, eps);
  });
  if (rc != PCM_OK) return rc;
  return check_launch("gn_silu_img_fwd");
}

extern "C" int pcm_convblock_tail_fwd(const void* x, const float* gamma, const float* beta, const float* w1,
                                      const float* w2, const float* wsp, float* stats, float* pool, float* se,
                                      float* hid, float* maps, unsigned char* ties, void* out, int N, int H, int W,
                                      int C, int Cr, float eps, int dtype, pcm_stream_t s) {
  PCM_REQUIRE(fused_shape_ok(H, W, C, Cr), "convblock_tail_fwd: unsupported shape H=%d W=%d C=%d Cr=%d", H, W, C, Cr);
  if (N == 0) return PCM_OK;
  const size_t smem = tail_smem_layout(H, W, C, dtype == PCM_BF16 ? 2 : 4, 1, 0).total;
  PCM_REQUIRE(smem <= 227 * 1024, "convblock_tail_fwd: image does not fit shared memory (%zu B)", smem);
  PCM_REQUIRE((maps == nullptr) == (ties == nullptr), "convblock_tail_fwd: maps and ties are saved together");
  int rc = PCM_OK;
  const bool narrow = tail_threads(H, W, C) <= 256;
  PCM_DISPATCH_DTYPE(dtype, T, {
    auto kern = narrow ? convblock_tail_fwd_kernel<T, true, 256> : convblock_tail_fwd_kernel<T, true, kFT>;
    rc = tail_set_smem(kern, smem, "convblock_tail_fwd");
    if (rc == PCM_OK)
      pcm::launch(kern, N, tail_launch_threads(kern, N, H, W, C, smem, 2.5f), smem, (cudaStream_t)s,
          (const T*)x, gamma, beta, w1, w2, wsp, stats, pool, se, hid, maps, ties, (T*)out, H, W, C, Cr, eps);
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
    const int threads = tail_launch_threads(gn_silu_img_bwd_kernel<T>, N, H, W, C, smem, 6.75f);
    const size_t smem_scr = tail_smem_layout(H, W, C, (int)sizeof(T), 0, 1, 1).total;
    const int scr = tail_scratch_ok(gn_silu_img_bwd_kernel<T>, threads, smem, smem_scr);
    const size_t bytes = scr ? smem_scr : smem;
    rc = tail_set_smem(gn_silu_img_bwd_kernel<T>, bytes, "gn_silu_img_bwd");
    if (rc == PCM_OK)
      pcm::launch(gn_silu_img_bwd_kernel<T>, N, threads, bytes, (cudaStream_t)s, (const T*)da, (const T*)x, stats, gamma, beta,
                  (T*)dx, dgamma, dbeta, H, W, C, eps, scr);
  });
  if (rc != PCM_OK) return rc;
  return check_launch("gn_silu_img_bwd");
}

extern "C" int pcm_gate_wgrad(const float* dq, const float* maps, float* dwsp, int N, int H, int W, pcm_stream_t s) {
  PCM_REQUIRE(dq != nullptr && maps != nullptr && dwsp != nullptr && H >= 1 && W >= 1, "gate_wgrad: bad arguments");
  if (N == 0) return PCM_OK;
  const size_t smem = (size_t)3 * (H + 6) * plane_wp(W) * sizeof(float);
  PCM_REQUIRE(smem <= 227 * 1024, "gate_wgrad: planes do not fit shared memory (%zu B)", smem);
  int rc = tail_set_smem(gate_wgrad_kernel, smem, "gate_wgrad");
  if (rc != PCM_OK) return rc;
  int grid = N < 148 * 4 ? N : 148 * 4;
  pcm::launch(gate_wgrad_kernel, grid, kGwThreads, smem, (cudaStream_t)s, dq, maps, dwsp, N, H, W);
  return check_launch("gate_wgrad");
}

extern "C" int pcm_convblock_tail_bwd_dq(const void* dout, const void* x, const void* out, const float* stats,
                                         const float* gamma, const float* beta, const float* w1, const float* w2,
                                         const float* wsp, const float* pool, const float* se, const float* hid,
                                         const float* maps, const unsigned char* ties, void* dx, float* dgamma,
                                         float* dbeta, float* dw1, float* dw2, float* dwsp, float* dq_out, int N, int H,
                                         int W, int C, int Cr, float eps, int dtype, pcm_stream_t s);
extern "C" int pcm_convblock_tail_bwd(const void* dout, const void* x, const void* out, const float* stats,
                                      const float* gamma, const float* beta, const float* w1, const float* w2,
                                      const float* wsp, const float* pool, const float* se, const float* hid,
                                      const float* maps, const unsigned char* ties, void* dx, float* dgamma,
                                      float* dbeta, float* dw1, float* dw2, float* dwsp, int N, int H, int W, int C,
                                      int Cr, float eps, int dtype, pcm_stream_t s) {
  return pcm_convblock_tail_bwd_dq(dout, x, out, stats, gamma, beta, w1, w2, wsp, pool, se, hid, maps, ties, dx, dgamma, dbeta,
                                   dw1, dw2, dwsp, nullptr, N, H, W, C, Cr, eps, dtype, s);
}

extern "C" int pcm_convblock_tail_bwd_sdot(const void* dout, const void* x, const void* out, const float* stats,
                                           const float* gamma, const float* beta, const float* w1, const float* w2,
                                           const float* wsp, const float* pool, const float* se, const float* hid,
                                           const float* maps, const unsigned char* ties, void* dx, float* dgamma,
                                           float* dbeta, float* dw1, float* dw2, float* dwsp, float* dq_out,
                                           const float* sdot, int N, int H, int W, int C, int Cr, float eps, int dtype,
                                           pcm_stream_t s);
extern "C" int pcm_convblock_tail_bwd_dq(const void* dout, const void* x, const void* out, const float* stats,
                                         const float* gamma, const float* beta, const float* w1, const float* w2,
                                         const float* wsp, const float* pool, const float* se, const float* hid,
                                         const float* maps, const unsigned char* ties, void* dx, float* dgamma,
                                         float* dbeta, float* dw1, float* dw2, float* dwsp, float* dq_out, int N, int H,
                                         int W, int C, int Cr, float eps, int dtype, pcm_stream_t s) {
  return pcm_convblock_tail_bwd_sdot(dout, x, out, stats, gamma, beta, w1, w2, wsp, pool, se, hid, maps, ties, dx, dgamma, dbeta,
                                     dw1, dw2, dwsp, dq_out, nullptr, N, H, W, C, Cr, eps, dtype, s);
}

extern "C" int pcm_convblock_tail_bwd_sdot(const void* dout, const void* x, const void* out, const float* stats,
                                           const float* gamma, const float* beta, const float* w1, const float* w2,
                                           const float* wsp, const float* pool, const float* se, const float* hid,
                                           const float* maps, const unsigned char* ties, void* dx, float* dgamma,
                                           float* dbeta, float* dw1, float* dw2, float* dwsp, float* dq_out,
                                           const float* sdot, int N, int H, int W, int C, int Cr, float eps, int dtype,
                                           pcm_stream_t s) {
  PCM_REQUIRE(fused_shape_ok(H, W, C, Cr), "convblock_tail_bwd: unsupported shape H=%d W=%d C=%d Cr=%d", H, W, C, Cr);
  if (N == 0) return PCM_OK;
  const size_t smem = tail_smem_layout(H, W, C, dtype == PCM_BF16 ? 2 : 4, 1, 1).total;
  PCM_REQUIRE(smem <= 227 * 1024, "convblock_tail_bwd: image does not fit shared memory (%zu B)", smem);
  PCM_REQUIRE((out != nullptr || sdot != nullptr) && maps != nullptr && ties != nullptr,
              "convblock_tail_bwd: the forward tail's saved out (or sdot) / maps / ties are required");
  PCM_REQUIRE(((uintptr_t)x & 15) == 0, "convblock_tail_bwd: x must be 16-byte aligned (bulk copy)");
  int rc = PCM_OK;
  PCM_DISPATCH_DTYPE(dtype, T, {
    const int threads = tail_launch_threads(convblock_tail_bwd_kernel<T>, N, H, W, C, smem, 6.75f);
    const size_t smem_scr = tail_smem_layout(H, W, C, (int)sizeof(T), 1, 1, 1).total;
    const int scr = tail_scratch_ok(convblock_tail_bwd_kernel<T>, threads, smem, smem_scr);
    const size_t bytes = scr ? smem_scr : smem;
    rc = tail_set_smem(convblock_tail_bwd_kernel<T>, bytes, "convblock_tail_bwd");
    if (rc == PCM_OK)
      pcm::launch(convblock_tail_bwd_kernel<T>, N, threads, bytes, (cudaStream_t)s,
          (const T*)dout, (const T*)x, (const T*)out, stats, gamma, beta, w1, w2, wsp, pool, se, hid, maps, ties,
          (T*)dx, dgamma, dbeta, dw1, dw2, dwsp, dq_out, sdot, H, W, C, Cr, eps, scr);
  });
  if (rc != PCM_OK) return rc;
  return check_launch("convblock_tail_bwd");
}
