// Persistent ConvLSTM recurrence (reference src/convlstm.py:27-35): ONE launch for all T steps forward, ONE for the
// whole back-propagation through time, instead of one gate-convolution launch (+ one cell launch) per step.
//
// The recurrence is sequential in t but independent per sample, and at the bottleneck grid (6 x 9) a sample is tiny,
// so the unit of work is a CLUSTER OF FOUR CTAs that owns two samples for all T steps:
//
//   * CTA r of the cluster owns hidden channels [16r, 16r + 16) of both samples: its slice of the recurrent weight
//     Wh (the 4 x 16 gate rows of those channels, 9 taps x 64 rows x 64 input channels bf16 = 72 KB) is loaded into
//     shared memory ONCE and stays resident for all steps (the per-step kernel re-fetched it every step);
//   * h_{t-1} of a sample lives in shared memory as a zero-ringed halo image [(H+2) x (W+2) pixels][64 ch] in the
//     canonical K-major 128-byte-swizzled UMMA layout; the nine taps of the gate convolution are nine descriptors whose
//     start address is advanced by kh*(W+2) + kw pixel rows (the trick of conv3x3_tc_halo_kernel), so one step is
//     36 tcgen05.mma (M128 x N64 x K16) per sample with NO global-memory operand traffic;
//   * the accumulator (Wh.h, fp32) stays in TMEM; the epilogue adds the hoisted Wx.x + b (`gx`, computed for all T by one
//     batched launch), applies sigmoid/tanh and the cell update with the cell state c held in REGISTERS across steps, and
//     writes its 16 channels of h_t straight into the halo images of all four CTAs of the cluster through distributed
//     shared memory (st.shared::cluster); one cluster barrier per step is the only synchronisation.  The halo images
//     are double buffered, so a CTA may run one step ahead of a peer that is still multiplying h_{t-1}.
//   * BPTT runs the same structure backwards: the cell backward (src/convlstm.py:14-18 differentiated) is computed in
//     registers from the saved activations, its 64 gate gradients per pixel go into the CTA's OWN halo image (K = this
//     CTA's slice of the 256 gate channels), the data-gradient convolution Wh^T * dgates is therefore split over K across
//     the four CTAs, and the four fp32 partial sums [pixels x 64 ch] are exchanged through distributed shared memory so
//     that each CTA ends the step with the complete dh_{t-1} of its own 16 channels.  dc stays in registers.
//
// Global traffic per step is what must exist anyway: gx in; activations, c and h out (saved for the backward /
// weight gradients); dgates out.  Supported: Ch = 64, bf16, grids whose flattened halo rows fit one M = 128 tile
// ((H-1)*(W+2) + W <= 128: the 6 x 9 bottleneck of the 48 x 72 emulator grid) — other shapes keep the per-step path.
#include "tc_common.cuh"

namespace pcm {

using namespace tc;

// -DPCM_SEQ_PROFILE: thread 0 of cluster 0 / CTA 0 accumulates clock64 deltas per phase and prints them at exit
#ifdef PCM_SEQ_PROFILE
#define SEQ_PROF_DECL long long prof_t[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; long long prof_c = clock64(); const bool prof_on = blockIdx.x == 0 && threadIdx.x == 0
#define SEQ_PROF(i) do { if (prof_on) { const long long c_ = clock64(); prof_t[i] += c_ - prof_c; prof_c = c_; } } while (0)
#define SEQ_PROF_PRINT(name) do { if (prof_on) printf("%s phases (cycles): %lld %lld %lld %lld %lld %lld %lld %lld\n", name, prof_t[0], prof_t[1], prof_t[2], prof_t[3], prof_t[4], prof_t[5], prof_t[6], prof_t[7]); if (prof_on) printf("   prologue: zero/init %lld weights %lld\n", prof_t[8], prof_t[9]); } while (0)
#else
#define SEQ_PROF_DECL
#define SEQ_PROF(i)
#define SEQ_PROF_PRINT(name)
#endif

constexpr int kSeqThreads = 192;      // warps 0,1: epilogue of sample 0 (TMEM lanes 0-63); 4,5: sample 1; 2: MMA issuer
constexpr int kSeqCh = 64;
constexpr int kSeqCluster = 4;
constexpr uint32_t kABuf = 160 * 128; // one sample's halo image: 8 x 11 = 88 pixel rows used, 128 + 24 rows addressable

struct SeqParams {
  int T, B, H, W, Wh, P;              // Wh = W + 2 (halo row pitch), P = H * W
  int ext_all;                        // backward: dh_ext has T steps (1) or only the last one (0)
};

__device__ __forceinline__ float tanh_apx(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x));
  return t;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address of this CTA -> the same offset in CTA `rank` of the cluster (shared::cluster window)
__device__ __forceinline__ uint32_t map_to_rank(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster16(uint32_t raddr, uint4 v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(raddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async;" ::: "memory"); }

__device__ __forceinline__ uint4 pack8_bf16(const float* v) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  return u;
}
__device__ __forceinline__ void unpack8_bf16(uint4 u, float* v) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 f = __bfloat1622float2(h[i]); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}

// byte offset of 16-byte chunk `chunk` (0..7) of pixel row `row` inside a 1024-byte aligned K-major tile of 128-byte rows
// with the 128-byte swizzle (chunk index XOR row-within-8)
__device__ __forceinline__ uint32_t sw128(uint32_t row, uint32_t chunk) { return row * 128u + ((chunk ^ (row & 7u)) << 4); }

// The 36 MMAs of one sample's gate (or data-gradient) convolution: D[128 x 64] = sum_tap A(row shift kh*Wh + kw) * B_tap^T
__device__ __forceinline__ void issue_conv(uint32_t d_tmem, uint32_t a_base, uint32_t b_base, int Wh) {
  const uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
  const uint64_t adesc0 = make_smem_desc(a_base, 16, 1024, 2);
  const uint64_t bdesc0 = make_smem_desc(b_base, 16, 1024, 2);
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const uint32_t a_off = ((uint32_t)((tap / 3) * Wh + (tap % 3)) * 128u) >> 4;
    const uint32_t b_off = ((uint32_t)tap * 8192u) >> 4;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      umma_bf16(d_tmem, adesc0 + (uint64_t)(a_off + 2 * k), bdesc0 + (uint64_t)(b_off + 2 * k), idesc, (tap | k) != 0);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// forward: h_t, c_t, activations for t = 0 .. T-1
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSeqThreads, 1)
convlstm_seq_fwd_kernel(const float* __restrict__ gx, const __nv_bfloat16* __restrict__ wh,
                        __nv_bfloat16* __restrict__ h_all, float* __restrict__ c_all, __nv_bfloat16* __restrict__ acts,
                        unsigned int* __restrict__ err, const SeqParams p) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sB = smem;                                   // 9 taps x [64 gate rows][64 k] bf16, swizzled
  uint8_t* sA = smem + 9 * 8192;                        // [2 buffers][2 samples] halo images
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + 4 * kABuf);
  uint64_t* mma_done = bars;                            // [2]: one per sample tile
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t r = cluster_ctarank();
  const int pair = blockIdx.x / kSeqCluster;
  const int Ch = kSeqCh, G = 4 * kSeqCh;
  SEQ_PROF_DECL;

  {  // zero both buffers of both halo images (the ring and the slack rows must read as zero)
    uint4* z = reinterpret_cast<uint4*>(sA);
    for (uint32_t i = threadIdx.x; i < 4 * kABuf / 16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
  }
  if (warp == 0 && lane == 0) {
    mbar_init(&mma_done[0], 1);
    mbar_init(&mma_done[1], 1);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 128);
  pdl_wait();                       // predecessor complete: global memory may be touched from here on
  SEQ_PROF(8);
  // resident weight slice: gate q, hidden channel 16r + j  ->  B row q*16 + j  (all 64 input channels, 9 taps);
  // eight 16-byte loads in flight per thread (the loop is latency-bound otherwise: 24 dependent round trips)
  for (int i0 = threadIdx.x; i0 < 9 * 64 * 8; i0 += 8 * kSeqThreads) {
    uint4 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = i0 + k * kSeqThreads;
      const int chunk = i & 7, n = (i >> 3) & 63, tap = i >> 9;
      const int grow = (n >> 4) * Ch + (int)r * 16 + (n & 15);
      if (i < 9 * 64 * 8) v[k] = __ldg(reinterpret_cast<const uint4*>(wh + ((size_t)tap * G + grow) * Ch) + chunk);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = i0 + k * kSeqThreads;
      const int chunk = i & 7, n = (i >> 3) & 63, tap = i >> 9;
      if (i < 9 * 64 * 8) *reinterpret_cast<uint4*>(sB + tap * 8192 + sw128(n, chunk)) = v[k];
    }
  }
  SEQ_PROF(9);
  fence_async_proxy();
  tc_fence_before();
  cluster_sync_all();               // every CTA of the cluster has zeroed its images before anyone writes into them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // epilogue role: (sample s, pixel row m of the flattened halo tile)
  const bool is_epi = (warp & 3) < 2 && warp != 2 && warp != 3;        // warps 0, 1, 4, 5
  const int s = warp >> 2, quarter = warp & 3;
  const int m = quarter * 32 + lane;
  const int hh = m / p.Wh, ww = m - hh * p.Wh;
  const int n = pair * 2 + s;
  const bool valid = is_epi && ww < p.W && hh < p.H && n < p.B;
  const long long pix = valid ? ((long long)n * p.P + hh * p.W + ww) : 0;
  const uint32_t arow = (uint32_t)((hh + 1) * p.Wh + ww + 1);           // this pixel's row in the halo image
  float cst[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) cst[j] = 0.f;
  const long long BP = (long long)p.B * p.P;

  bool ok = true;
  SEQ_PROF(0);                      // 0: first cluster barrier
  for (int t = 0; t < p.T; ++t) {      // every thread runs every step (cluster barriers); a timed-out wait only skips work
    if (warp == 2 && t > 0) {
      if (elect_one()) {
        fence_async_proxy();
        tc_fence_after();
        const uint32_t abuf = smem_u32(sA) + (uint32_t)(t & 1) * 2u * kABuf;
        issue_conv(tmem_base, abuf, smem_u32(sB), p.Wh);
        umma_commit(&mma_done[0]);
        issue_conv(tmem_base + 64, abuf + kABuf, smem_u32(sB), p.Wh);
        umma_commit(&mma_done[1]);
      }
      __syncwarp();
    }
    if (is_epi) {
      float gi[16], gf[16], go[16], gg[16];
      if (valid) {      // hoisted Wx.x + b of this step: issued before the wait, overlaps the MMAs
        const float* gp = gx + ((long long)t * BP + pix) * G + r * 16;
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(gp + j));
          const float4 b = __ldg(reinterpret_cast<const float4*>(gp + Ch + j));
          const float4 c = __ldg(reinterpret_cast<const float4*>(gp + 2 * Ch + j));
          const float4 d = __ldg(reinterpret_cast<const float4*>(gp + 3 * Ch + j));
          gi[j] = a.x; gi[j + 1] = a.y; gi[j + 2] = a.z; gi[j + 3] = a.w;
          gf[j] = b.x; gf[j + 1] = b.y; gf[j + 2] = b.z; gf[j + 3] = b.w;
          go[j] = c.x; go[j + 1] = c.y; go[j + 2] = c.z; go[j + 3] = c.w;
          gg[j] = d.x; gg[j + 1] = d.y; gg[j + 2] = d.z; gg[j + 3] = d.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) gi[j] = gf[j] = go[j] = gg[j] = 0.f;
      }
      SEQ_PROF(1);                  // 1: gx loads issued
      if (t > 0) {
        ok = mbar_wait(&mma_done[s], (uint32_t)((t - 1) & 1), err) && ok;
        ok = __all_sync(0xffffffffu, ok);
        SEQ_PROF(2);                // 2: waiting for the MMAs
        if (ok) {
          tc_fence_after();
          const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)s * 64u;
          float v[16];
          tmem_ld16(t_addr, v);
#pragma unroll
          for (int j = 0; j < 16; ++j) gi[j] += v[j];
          tmem_ld16(t_addr + 16, v);
#pragma unroll
          for (int j = 0; j < 16; ++j) gf[j] += v[j];
          tmem_ld16(t_addr + 32, v);
#pragma unroll
          for (int j = 0; j < 16; ++j) go[j] += v[j];
          tmem_ld16(t_addr + 48, v);
#pragma unroll
          for (int j = 0; j < 16; ++j) gg[j] += v[j];
          tc_fence_before();
        }
      }
      SEQ_PROF(3);                  // 3: TMEM loads
      if (ok && valid) {
        // i,f,o = sigmoid, g = tanh ; c' = f*c + i*g ; h' = o*tanh(c')  (src/convlstm.py:14-18); same one-instruction
        // tanh.approx forms and bf16 rounding of the saved activations as the per-step kernel (conv_tc.cu, EPI == 1)
        float hn[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float a_i = round_to<__nv_bfloat16>(fmaf(0.5f, tanh_apx(0.5f * gi[j]), 0.5f));
          const float a_f = round_to<__nv_bfloat16>(fmaf(0.5f, tanh_apx(0.5f * gf[j]), 0.5f));
          const float a_o = round_to<__nv_bfloat16>(fmaf(0.5f, tanh_apx(0.5f * go[j]), 0.5f));
          const float a_g = round_to<__nv_bfloat16>(tanh_apx(gg[j]));
          gi[j] = a_i; gf[j] = a_f; go[j] = a_o; gg[j] = a_g;
          cst[j] = fmaf(a_f, cst[j], a_i * a_g);
          hn[j] = a_o * tanh_apx(cst[j]);
        }
        const long long row = (long long)t * BP + pix;
        __nv_bfloat16* ap = acts + row * G + r * 16;
        store8(ap, gi); store8(ap + 8, gi + 8);
        store8(ap + Ch, gf); store8(ap + Ch + 8, gf + 8);
        store8(ap + 2 * Ch, go); store8(ap + 2 * Ch + 8, go + 8);
        store8(ap + 3 * Ch, gg); store8(ap + 3 * Ch + 8, gg + 8);
        float* cq = c_all + row * Ch + r * 16;
        store8(cq, cst); store8(cq + 8, cst + 8);
        const uint4 h0 = pack8_bf16(hn), h1 = pack8_bf16(hn + 8);
        uint4* hp = reinterpret_cast<uint4*>(h_all + row * Ch + r * 16);
        hp[0] = h0; hp[1] = h1;
        if (t + 1 < p.T) {
          // h_t (these 16 channels = 16-byte chunks 2r, 2r+1 of the pixel's row) into the NEXT step's halo image of
          // every CTA of the cluster
          const uint32_t base = smem_u32(sA) + (uint32_t)((t + 1) & 1) * 2u * kABuf + (uint32_t)s * kABuf;
          const uint32_t o0 = base + sw128(arow, 2 * r), o1 = base + sw128(arow, 2 * r + 1);
#pragma unroll
          for (uint32_t peer = 0; peer < kSeqCluster; ++peer) {
            st_cluster16(map_to_rank(o0, peer), h0);
            st_cluster16(map_to_rank(o1, peer), h1);
          }
        }
      }
    }
    SEQ_PROF(4);                    // 4: cell math, global stores, DSMEM stores
    if (t + 1 < p.T) {
      fence_async_proxy();          // the halo-image writes (generic proxy) will be read by the tensor core's async proxy
      tc_fence_before();
      cluster_sync_all();
      tc_fence_after();
    }
    SEQ_PROF(5);                    // 5: fence + cluster barrier
  }
  SEQ_PROF_PRINT("convlstm_seq_fwd");
  tc_fence_before();
  cluster_sync_all();               // no CTA leaves while a peer could still write into its shared memory
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// backward through time: dgates for t = T-1 .. 0
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSeqThreads, 1)
convlstm_seq_bwd_kernel(const __nv_bfloat16* __restrict__ dh_ext, const __nv_bfloat16* __restrict__ acts,
                        const float* __restrict__ c_all, const __nv_bfloat16* __restrict__ wht,
                        __nv_bfloat16* __restrict__ dgates, unsigned int* __restrict__ err, const SeqParams p) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sB = smem;                                   // 9 taps x [64 output channels][64 k = this CTA's gate channels]
  uint8_t* sA = smem + 9 * 8192;                        // [2 samples] halo images of this CTA's dgates slice
  float* red = reinterpret_cast<float*>(sA + 2 * kABuf);   // [2 parity][4 source CTAs][2 samples][64 rows][16] partial dh
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(red) + 2 * 4 * 2 * 64 * 64);
  uint64_t* mma_done = bars;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t r = cluster_ctarank();
  const int pair = blockIdx.x / kSeqCluster;
  const int Ch = kSeqCh, G = 4 * kSeqCh;
  SEQ_PROF_DECL;

  {
    uint4* z = reinterpret_cast<uint4*>(sA);
    for (uint32_t i = threadIdx.x; i < 2 * kABuf / 16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
  }
  if (warp == 0 && lane == 0) {
    mbar_init(&mma_done[0], 1);
    mbar_init(&mma_done[1], 1);
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 128);
  pdl_wait();                       // predecessor complete: global memory may be touched from here on
  SEQ_PROF(8);
  // resident weight slice of the data-gradient convolution: wht[tap][n = hidden channel][k = gate channel] (taps
  // flipped by the packer); this CTA's K = gate q, channel 16r + j  ->  k' = q*16 + j
  for (int i0 = threadIdx.x; i0 < 9 * 64 * 8; i0 += 8 * kSeqThreads) {
    uint4 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = i0 + k * kSeqThreads;
      const int chunk = i & 7, nn = (i >> 3) & 63, tap = i >> 9;
      const int q = chunk >> 1, half = chunk & 1;
      if (i < 9 * 64 * 8)
        v[k] = __ldg(reinterpret_cast<const uint4*>(wht + ((size_t)tap * Ch + nn) * G + q * Ch + (int)r * 16 + half * 8));
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = i0 + k * kSeqThreads;
      const int chunk = i & 7, nn = (i >> 3) & 63, tap = i >> 9;
      if (i < 9 * 64 * 8) *reinterpret_cast<uint4*>(sB + tap * 8192 + sw128(nn, chunk)) = v[k];
    }
  }
  SEQ_PROF(9);
  fence_async_proxy();
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const bool is_epi = (warp & 3) < 2 && warp != 2 && warp != 3;
  const int s = warp >> 2, quarter = warp & 3;
  const int m = quarter * 32 + lane;
  const int hh = m / p.Wh, ww = m - hh * p.Wh;
  const int n = pair * 2 + s;
  const bool valid = is_epi && ww < p.W && hh < p.H && n < p.B;
  const long long pix = valid ? ((long long)n * p.P + hh * p.W + ww) : 0;
  const uint32_t arow = (uint32_t)((hh + 1) * p.Wh + ww + 1);
  const long long BP = (long long)p.B * p.P;
  float dcs[16], dhn[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) dcs[j] = dhn[j] = 0.f;

  // operands of the cell backward of one step, as loaded: 4 x 16 activations (bf16), c_t, c_{t-1}, external dh_t.  The
  // loads of step t-1 are issued while the tensor core works on step t and the cluster exchanges partial sums.
  uint4 ra[8], re[2];
  float4 rc[4], rp[4];
  auto load_step = [&](int t) {
    if (!valid) return;
    const long long row = (long long)t * BP + pix;
    const __nv_bfloat16* ap = acts + row * G + r * 16;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      ra[2 * q] = __ldg(reinterpret_cast<const uint4*>(ap + q * Ch));
      ra[2 * q + 1] = __ldg(reinterpret_cast<const uint4*>(ap + q * Ch) + 1);
    }
    const float4* cq = reinterpret_cast<const float4*>(c_all + row * Ch + r * 16);
#pragma unroll
    for (int j = 0; j < 4; ++j) rc[j] = __ldg(cq + j);
    if (t > 0) {
      const float4* cr = reinterpret_cast<const float4*>(c_all + (row - BP) * Ch + r * 16);
#pragma unroll
      for (int j = 0; j < 4; ++j) rp[j] = __ldg(cr + j);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) rp[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (p.ext_all || t == p.T - 1) {
      const uint4* ep = reinterpret_cast<const uint4*>(dh_ext + ((p.ext_all ? row : pix) * Ch + r * 16));
      re[0] = __ldg(ep); re[1] = __ldg(ep + 1);
    } else {
      re[0] = re[1] = make_uint4(0, 0, 0, 0);
    }
  };
  load_step(p.T - 1);

  bool ok = true;
  int par = 0;
  SEQ_PROF(0);                      // 0: first cluster barrier, first operand loads issued
  for (int t = p.T - 1; t >= 0; --t) {
    if (valid) {
      // ---- cell backward at step t for this pixel's 16 channels (same arithmetic as lstm_cell_bwd_kernel)
      const long long row = (long long)t * BP + pix;
      float dh[16], cc[16], cp[16], gi[16], gf[16], go[16], gg[16];
      {
        float e[16];
        unpack8_bf16(re[0], e); unpack8_bf16(re[1], e + 8);
#pragma unroll
        for (int j = 0; j < 16; ++j) dh[j] = dhn[j] + e[j];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        cc[4 * j] = rc[j].x; cc[4 * j + 1] = rc[j].y; cc[4 * j + 2] = rc[j].z; cc[4 * j + 3] = rc[j].w;
        cp[4 * j] = rp[j].x; cp[4 * j + 1] = rp[j].y; cp[4 * j + 2] = rp[j].z; cp[4 * j + 3] = rp[j].w;
      }
      unpack8_bf16(ra[0], gi); unpack8_bf16(ra[1], gi + 8);
      unpack8_bf16(ra[2], gf); unpack8_bf16(ra[3], gf + 8);
      unpack8_bf16(ra[4], go); unpack8_bf16(ra[5], go + 8);
      unpack8_bf16(ra[6], gg); unpack8_bf16(ra[7], gg + 8);
      float di[16], df[16], dO[16], dg[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float tc = tanhf(cc[j]);
        const float dct = fmaf(dh[j] * go[j], 1.f - tc * tc, dcs[j]);
        dO[j] = dh[j] * tc * go[j] * (1.f - go[j]);
        di[j] = dct * gg[j] * gi[j] * (1.f - gi[j]);
        df[j] = dct * cp[j] * gf[j] * (1.f - gf[j]);
        dg[j] = dct * gi[j] * (1.f - gg[j] * gg[j]);
        dcs[j] = dct * gf[j];
      }
      uint4 u[8];
      u[0] = pack8_bf16(di); u[1] = pack8_bf16(di + 8); u[2] = pack8_bf16(df); u[3] = pack8_bf16(df + 8);
      u[4] = pack8_bf16(dO); u[5] = pack8_bf16(dO + 8); u[6] = pack8_bf16(dg); u[7] = pack8_bf16(dg + 8);
      __nv_bfloat16* dp = dgates + row * G + r * 16;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4* d4 = reinterpret_cast<uint4*>(dp + q * Ch);
        d4[0] = u[2 * q]; d4[1] = u[2 * q + 1];
      }
      if (t > 0) {
        uint8_t* a = sA + (size_t)s * kABuf;
#pragma unroll
        for (uint32_t c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(a + sw128(arow, c)) = u[c];
      }
    }
    SEQ_PROF(1);                    // 1: cell backward, dgates stores
    if (t == 0) break;
    // ---- dh_{t-1} = conv(dgates_t, flipped Wh^T): this CTA's K slice -> partial sums for all 64 hidden channels
    fence_async_proxy();
    tc_fence_before();
    __syncthreads();
    SEQ_PROF(2);                    // 2: fence + CTA barrier
    if (warp == 2) {
      if (elect_one()) {
        tc_fence_after();
        issue_conv(tmem_base, smem_u32(sA), smem_u32(sB), p.Wh);
        umma_commit(&mma_done[0]);
        issue_conv(tmem_base + 64, smem_u32(sA) + kABuf, smem_u32(sB), p.Wh);
        umma_commit(&mma_done[1]);
      }
      __syncwarp();
    }
    load_step(t - 1);               // next step's operands: in flight during the MMAs and the cluster exchange
    if (is_epi) {
      ok = mbar_wait(&mma_done[s], (uint32_t)((p.T - 1 - t) & 1), err) && ok;
      ok = __all_sync(0xffffffffu, ok);
      SEQ_PROF(3);                  // 3: waiting for the MMAs
      if (ok) {
        tc_fence_after();
        const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)s * 64u;
        // the 16-channel block of hidden channels [16 r', 16 r' + 16) goes to CTA r' (slot = this CTA's rank)
        const uint32_t slot = smem_u32(red) + (uint32_t)(((par * 4 + (int)r) * 2 + s) * 64 + m) * 64u;
#pragma unroll
        for (uint32_t peer = 0; peer < kSeqCluster; ++peer) {
          float v[16];
          tmem_ld16(t_addr + peer * 16, v);
          const uint32_t ra = map_to_rank(slot, peer);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            st_cluster16(ra + 16u * j, make_uint4(__float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]),
                                                  __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3])));
        }
        tc_fence_before();
      }
    }
    SEQ_PROF(4);                    // 4: TMEM loads + DSMEM stores of the partial sums
    cluster_sync_all();             // all four partial sums of this step have landed in every CTA
    SEQ_PROF(5);                    // 5: cluster barrier
    if (is_epi) {
#pragma unroll
      for (int j = 0; j < 16; ++j) dhn[j] = 0.f;
#pragma unroll
      for (int src = 0; src < kSeqCluster; ++src) {
        const float4* q4 = reinterpret_cast<const float4*>(red + (size_t)(((par * 4 + src) * 2 + s) * 64 + m) * 16);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 v = q4[j];
          dhn[4 * j] += v.x; dhn[4 * j + 1] += v.y; dhn[4 * j + 2] += v.z; dhn[4 * j + 3] += v.w;
        }
      }
    }
    par ^= 1;
    SEQ_PROF(6);                    // 6: summing the four partials
  }
  SEQ_PROF_PRINT("convlstm_seq_bwd");
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

static bool seq_shape_ok(int H, int W, int Ch) {
  // valid accumulator rows within TMEM lanes 0-63 (the two epilogue warps of a sample); the furthest row a tap of the
  // 128-row tile reads (127 + 2*(W+2) + 2) inside the 160-row image buffer
  return Ch == kSeqCh && H >= 1 && W >= 1 && (H - 1) * (W + 2) + W <= 64 && 2 * (W + 2) + 2 + 128 <= 160;
}

}  // namespace pcm

using namespace pcm;

extern "C" int pcm_convlstm_seq_supported(int H, int W, int Ch) { return seq_shape_ok(H, W, Ch) ? 1 : 0; }

extern "C" int pcm_convlstm_seq_fwd_tc(const float* gx, const void* wh, void* h_all, float* c_all, void* acts, int T, int B,
                                       int H, int W, int Ch, pcm_stream_t s) {
  PCM_REQUIRE(seq_shape_ok(H, W, Ch), "convlstm_seq_fwd_tc: unsupported shape H=%d W=%d Ch=%d", H, W, Ch);
  PCM_REQUIRE(((reinterpret_cast<uintptr_t>(gx) | reinterpret_cast<uintptr_t>(wh) | reinterpret_cast<uintptr_t>(h_all) |
                reinterpret_cast<uintptr_t>(c_all) | reinterpret_cast<uintptr_t>(acts)) & 15) == 0,
              "convlstm_seq_fwd_tc: pointers must be 16-byte aligned");
  if (T == 0 || B == 0) return PCM_OK;
  SeqParams p;
  p.T = T; p.B = B; p.H = H; p.W = W; p.Wh = W + 2; p.P = H * W; p.ext_all = 0;
  const size_t smem = 1024 + 9 * 8192 + 4 * kABuf + 64;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(convlstm_seq_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("convlstm_seq_fwd_tc: smem attribute (%zu B): %s", smem, cudaGetErrorString(e)); return PCM_ERR_CUDA; }
    attr_set = true;
  }
  unsigned int* err = tc_error_counter();
  PCM_REQUIRE(err != nullptr, "convlstm_seq_fwd_tc: could not allocate the error counter");
  const int clusters = (B + 1) / 2;
  pcm::launch_cluster(convlstm_seq_fwd_kernel, dim3(clusters * kSeqCluster), dim3(kSeqThreads), smem, (cudaStream_t)s, kSeqCluster,
                      gx, reinterpret_cast<const __nv_bfloat16*>(wh), reinterpret_cast<__nv_bfloat16*>(h_all), c_all,
                      reinterpret_cast<__nv_bfloat16*>(acts), err, p);
  return check_launch("convlstm_seq_fwd_tc");
}

extern "C" int pcm_convlstm_seq_bwd_tc(const void* dh_ext, int ext_all_steps, const void* acts, const float* c_all,
                                       const void* wht, void* dgates, int T, int B, int H, int W, int Ch, pcm_stream_t s) {
  PCM_REQUIRE(seq_shape_ok(H, W, Ch), "convlstm_seq_bwd_tc: unsupported shape H=%d W=%d Ch=%d", H, W, Ch);
  PCM_REQUIRE(((reinterpret_cast<uintptr_t>(dh_ext) | reinterpret_cast<uintptr_t>(acts) | reinterpret_cast<uintptr_t>(c_all) |
                reinterpret_cast<uintptr_t>(wht) | reinterpret_cast<uintptr_t>(dgates)) & 15) == 0,
              "convlstm_seq_bwd_tc: pointers must be 16-byte aligned");
  if (T == 0 || B == 0) return PCM_OK;
  SeqParams p;
  p.T = T; p.B = B; p.H = H; p.W = W; p.Wh = W + 2; p.P = H * W; p.ext_all = ext_all_steps ? 1 : 0;
  const size_t smem = 1024 + 9 * 8192 + 2 * kABuf + 2 * 4 * 2 * 64 * 64 + 64;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(convlstm_seq_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("convlstm_seq_bwd_tc: smem attribute (%zu B): %s", smem, cudaGetErrorString(e)); return PCM_ERR_CUDA; }
    attr_set = true;
  }
  unsigned int* err = tc_error_counter();
  PCM_REQUIRE(err != nullptr, "convlstm_seq_bwd_tc: could not allocate the error counter");
  const int clusters = (B + 1) / 2;
  pcm::launch_cluster(convlstm_seq_bwd_kernel, dim3(clusters * kSeqCluster), dim3(kSeqThreads), smem, (cudaStream_t)s, kSeqCluster,
                      reinterpret_cast<const __nv_bfloat16*>(dh_ext), reinterpret_cast<const __nv_bfloat16*>(acts), c_all,
                      reinterpret_cast<const __nv_bfloat16*>(wht), reinterpret_cast<__nv_bfloat16*>(dgates), err, p);
  return check_launch("convlstm_seq_bwd_tc");
}
