// 3x3 / stride-1 / pad-1 convolution as a tcgen05 implicit GEMM (sm_100a), NHWC bf16.
//
//   D[128 pixels x Cout] (fp32, TMEM) = sum over 9 taps, Cin chunks:  A_tap[128 x KC] * B_tap[Cout x KC]^T
//
// * A tiles are TMA boxes {KC channels, Wb, Hb, Nb} of the activation tensor, fetched at the tap's
//   (kh-1, kw-1) offset — the hardware zero-fills out-of-bounds coordinates, which IS the conv padding,
//   so halo tiles need no im2col buffer and no predication.  The box lands in shared memory as
//   consecutive pixel rows of KC*2 bytes with the 32/64/128-byte swizzle == the canonical K-major UMMA
//   operand layout.
// * B tiles are boxes {KC, Cout, 1} of the packed weight [9][Cout][Cin].
// * Persistent, warp-specialised CTA (1 per SM): warp 0 = TMA producer, warp 1 = MMA issuer (one elected
//   thread), warps 2-5 = epilogue.  smem ring (full/empty mbarriers) between producer and MMA; the fp32
//   accumulator is double buffered in TMEM (tmem_full/tmem_empty mbarriers) so the epilogue of tile i
//   overlaps the MMAs of tile i+1.
// * Epilogue: tcgen05.ld -> registers -> (+bias) (+= previous fp32 contents) -> bf16 / fp32 NHWC store.
//
// The same kernel is the data-gradient convolution (weights packed flipped + transposed) and the
// ConvLSTM gate convolution (fp32 output, accumulate = Wx.x + Wh.h split).
#include <stdlib.h>

#include "tc_common.cuh"

namespace pcm {

using namespace tc;

// Work item -> (channel group, tile column, tile row, image block) for item = first + k*step WITHOUT per-item
// divisions: a mixed-radix counter advanced by the (pre-decomposed) step.  The persistent loops of all three roles
// used three integer divisions per tile; in the thin-layer kernels (epilogue-bound, ~250 instructions per tile and
// warp) they were 30 % of the epilogue warps' stall samples.
struct TileWalk {
  int g, tw, th, tn, dg, dtw, dth, dtn, ng, nw, nh;
  __device__ __forceinline__ TileWalk(int first, int step, int ngroups, int tiles_w, int tiles_h)
      : ng(ngroups), nw(tiles_w), nh(tiles_h) {
    g = first % ng; int r = first / ng; tw = r % nw; r /= nw; th = r % nh; tn = r / nh;
    dg = step % ng; r = step / ng; dtw = r % nw; r /= nw; dth = r % nh; dtn = r / nh;
  }
  __device__ __forceinline__ void next() {
    g += dg; int c = g >= ng ? 1 : 0; g -= c ? ng : 0;
    tw += dtw + c; c = tw >= nw ? 1 : 0; tw -= c ? nw : 0;
    th += dth + c; c = th >= nh ? 1 : 0; th -= c ? nh : 0;
    tn += dtn + c;
  }
};

struct ConvTcParams {
  int N, H, W, Cin, Cout;
  int Wb, Hb, Nb, tiles_w, tiles_h, num_tiles;
  int KC, kchunks, stages;
  int ksz, relu;          // kernel size (1 or 3; pad = ksz/2), ReLU in the store epilogue
  float drop_p;           // > 0: nn.Dropout(p) after the (ReLU) epilogue — the mask of pcm_dropout on the dense destination
  unsigned long long drop_seed;          // (element index = offset from dst), so the fused and the two-kernel forms draw
  const unsigned long long* drop_epoch;  // the same mask; the library's epoch cell is mixed in on the device
  int mode;               // 0: conv taps (shifted boxes) ; 1: ConvTranspose2d(k2,s2) data gradient — 4 taps (kh,kw), tap q
                          //    gathers dy(2h+kh, 2w+kw) through the 5-D views tmA (kh=0) / tmA2 (kh=1) {C, kw, w, h, n}
                          // 2: 3x3 / stride-2 / pad-1 conv — the input is viewed as {2C (pw, c), W/2, ph, H/2, n}; 6 taps
                          //    (kh, dw): tap row 2h+kh-1 = (ph, h+oh), columns 2(w+dw)+pw with the unused (dw=-1, pw=0) weights zero
  int pad;                // conv taps: leading zero padding (ksz/2 for 'same', 0 for the stride-2 data-gradient taps)
  int ps_co;              // > 0: ConvTranspose2d(k2,s2) forward — GEMM column q*ps_co + co is stored at pixel
                          //    (2h + q/2, 2w + q%2), channel co of a (2H, 2W) destination (pixel shuffle in the epilogue)
  long long dst_ns;
  int dst_ps, dst_f32, accumulate;
  uint32_t tmem_cols, acc_stride, a_stage_bytes, b_stage_bytes, a_tx_bytes, b_tx_bytes;
  // fused ConvLSTM cell epilogue (EPI == 1): rows are pixels m = (n*H + h)*W + w of contiguous [B*P] tensors
  const float* gx;        // [M][4*Ch] fp32: Wx.x + bias for this step (computed for all T by one launch)
  const float* c_prev;    // [M][Ch] fp32
  float* c_out;           // [M][Ch] fp32
  __nv_bfloat16* acts;    // [M][4*Ch] bf16: activated gates i,f,o,g saved for backward
  int ngroups;            // EPI == 1: hidden channels are split into ngroups = Ch/16 groups; a work item is (pixel tile,
                          // group) and computes the 4 x 16 gate columns of its 16 hidden channels (4x more CTAs on the
                          // 27-tile recurrent step, 4x less epilogue work each); EPI == 0: 1
};

constexpr int kThreads = 192;

__device__ __forceinline__ float tanh_approx(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x));
  return t;
}

template <int KSTEPS, int EPI>   // KSTEPS = KC / 16 (UMMA K-steps per stage); EPI 0 = store, 1 = ConvLSTM cell
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                  const __grid_constant__ CUtensorMap tmB, void* __restrict__ dst, const float* __restrict__ bias, unsigned int* __restrict__ err,
                  const ConvTcParams p) {
  pdl_launch_dependents();          // the next kernel may start launching; it waits for us in its own pdl_wait()
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)p.stages * p.a_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)p.stages * p.b_stage_bytes);
  uint64_t* full = bars;                  // [stages]
  uint64_t* empty = bars + p.stages;      // [stages]
  uint64_t* tfull = empty + p.stages;     // [2]
  uint64_t* tempty = tfull + 2;           // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmA2);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < p.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 128); }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();                       // predecessor complete: global memory may be touched from here on
  const uint32_t tmem_base = *tmem_slot;

  const int ntaps = p.mode == 1 ? 4 : p.mode == 2 ? 6 : p.ksz * p.ksz, kpad = p.pad;
  const int kiters = ntaps * p.kchunks;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      bool ok = true;
      TileWalk tk(blockIdx.x, gridDim.x, p.ngroups, p.tiles_w, p.tiles_h);
      for (int item = blockIdx.x; item < p.num_tiles * p.ngroups && ok; item += gridDim.x, tk.next()) {
        const int cgrp = tk.g;
        const int w0 = tk.tw * p.Wb, h0 = tk.th * p.Hb, n0 = tk.tn * p.Nb;
        for (int tap = 0; tap < ntaps && ok; ++tap) {
          const int kh = tap / p.ksz, kw = tap % p.ksz;
          for (int kc = 0; kc < p.kchunks; ++kc) {
            ok = mbar_wait(&empty[stage], phase ^ 1, err);
            if (!ok) break;
            mbar_expect_tx(&full[stage], p.a_tx_bytes + p.b_tx_bytes);
            if (p.mode == 1)
              tma_load_5d(sA + (size_t)stage * p.a_stage_bytes, (tap >> 1) ? &tmA2 : &tmA, &full[stage], kc * p.KC, tap & 1, w0,
                          h0, n0);
            else if (p.mode == 2)     // tap = kh*2 + (dw+1): kh 0 -> (ph 1, h-1), 1 -> (ph 0, h), 2 -> (ph 1, h)
              tma_load_5d(sA + (size_t)stage * p.a_stage_bytes, &tmA, &full[stage], kc * p.KC, w0 + (tap & 1) - 1,
                          (tap >> 1) != 1 ? 1 : 0, h0 - ((tap >> 1) == 0 ? 1 : 0), n0);
            else
              tma_load_4d(sA + (size_t)stage * p.a_stage_bytes, &tmA, &full[stage], kc * p.KC, w0 + kw - kpad, h0 + kh - kpad, n0);
            if (EPI == 1) {
              // the 16 rows of each gate (i, f, o, g) that belong to this hidden-channel group: four 16-row boxes
              // land back to back as one 64-row B tile
              const int Ch = p.Cout >> 2;
#pragma unroll
              for (int gq = 0; gq < 4; ++gq)
                tma_load_3d(sB + (size_t)stage * p.b_stage_bytes + (size_t)gq * 16 * p.KC * 2, &tmB, &full[stage], kc * p.KC,
                            gq * Ch + cgrp * 16, tap);
            } else {
              tma_load_3d(sB + (size_t)stage * p.b_stage_bytes, &tmB, &full[stage], kc * p.KC, 0, tap);
            }
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      // The issuing thread is instruction-bound for small N: keep the loop to a handful of SASS
      // instructions per MMA (descriptors are start-address increments of two precomputed bases).
      const uint32_t idesc = make_idesc_bf16(128, EPI == 1 ? 64 : p.Cout, 0, 0);
      constexpr uint32_t row_bytes = KSTEPS * 32;
      const uint32_t ltype = layout_type_for_row_bytes(row_bytes);
      const uint64_t adesc0 = make_smem_desc(smem_u32(sA), 16, 8 * row_bytes, ltype);
      const uint64_t bdesc0 = make_smem_desc(smem_u32(sB), 16, 8 * row_bytes, ltype);
      const uint32_t a_step = p.a_stage_bytes >> 4, b_step = p.b_stage_bytes >> 4;
      const int nstages = p.stages, ntiles = p.num_tiles * p.ngroups, acc_stride = p.acc_stride;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      bool ok = true;
      for (int tile = blockIdx.x; tile < ntiles && ok; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        ok = mbar_wait(&tempty[acc], acc_phase ^ 1, err);
        if (!ok) break;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * acc_stride;
        for (int ki = 0; ki < kiters; ++ki) {
          ok = mbar_wait(&full[stage], phase, err);
          if (!ok) break;
          tc_fence_after();
          const uint64_t adesc = adesc0 + (uint64_t)(stage * a_step);
          const uint64_t bdesc = bdesc0 + (uint64_t)(stage * b_step);
#pragma unroll
          for (int k = 0; k < KSTEPS; ++k) {
            // advance 16 bf16 = 32 bytes along K inside the swizzle row: +2 in the (addr >> 4) field
            umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (ki | k) != 0);
          }
          umma_commit(&empty[stage]);
          if (++stage == nstages) { stage = 0; phase ^= 1; }
        }
        if (ok) umma_commit(&tfull[acc]);
      }
    }
  } else {
    // ===================== epilogue (warps 2..5 -> TMEM lane quarters 2,3,0,1) =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int mvalid = p.Wb * p.Hb * p.Nb;
    int it = 0;
    bool ok = true;
    const int wl = row % p.Wb, hl = (row / p.Wb) % p.Hb, nl = row / (p.Wb * p.Hb);
    TileWalk tk(blockIdx.x, gridDim.x, p.ngroups, p.tiles_w, p.tiles_h);
    for (int item = blockIdx.x; item < p.num_tiles * p.ngroups && ok; item += gridDim.x, ++it, tk.next()) {
      const int cgrp = tk.g;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int w = tk.tw * p.Wb + wl, h = tk.th * p.Hb + hl, n = tk.tn * p.Nb + nl;
      const bool valid = row < mvalid && w < p.W && h < p.H && n < p.N;
      const long long off = (long long)n * p.dst_ns + ((long long)h * p.W + w) * p.dst_ps;
      ok = mbar_wait(&tfull[acc], acc_phase, err);
      ok = __all_sync(0xffffffffu, ok);
      if (!ok) break;
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * p.acc_stride;
      if (EPI == 1) {
        // gates = Wh.h (TMEM) + gx ; i,f,o = sigmoid, g = tanh ; c' = f*c + i*g ; h' = o*tanh(c')
        // (src/convlstm.py:13-18).  This work item owns hidden channels [c0, c0 + 16); its accumulator holds their
        // i | f | o | g pre-activations in four 16-column blocks.  sigmoid / tanh use the one-instruction
        // tanh.approx (rel. error ~2^-11, below the bf16 resolution the gates are stored in).
        const int Ch = p.Cout >> 2;
        const int c0 = cgrp * 16;
        const long long m = ((long long)n * p.H + h) * p.W + w;
        float gi[16], gf[16], go[16], gg[16];
        tmem_ld16(t_addr, gi);
        tmem_ld16(t_addr + 16, gf);
        tmem_ld16(t_addr + 32, go);
        tmem_ld16(t_addr + 48, gg);
        if (valid) {
          const float* gxp = p.gx + m * p.Cout + c0;
          const float* cp = p.c_prev + m * Ch + c0;
          float* cq = p.c_out + m * Ch + c0;
          __nv_bfloat16* ap = p.acts + m * p.Cout + c0;
          __nv_bfloat16* hp = reinterpret_cast<__nv_bfloat16*>(dst) + m * Ch + c0;
          float cn[16], hn[16];
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 xi = *reinterpret_cast<const float4*>(gxp + j);
            const float4 xf = *reinterpret_cast<const float4*>(gxp + Ch + j);
            const float4 xo = *reinterpret_cast<const float4*>(gxp + 2 * Ch + j);
            const float4 xg = *reinterpret_cast<const float4*>(gxp + 3 * Ch + j);
            const float4 cc = *reinterpret_cast<const float4*>(cp + j);
            const float xi_[4] = {xi.x, xi.y, xi.z, xi.w}, xf_[4] = {xf.x, xf.y, xf.z, xf.w};
            const float xo_[4] = {xo.x, xo.y, xo.z, xo.w}, xg_[4] = {xg.x, xg.y, xg.z, xg.w};
            const float cc_[4] = {cc.x, cc.y, cc.z, cc.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float a_i = round_to<__nv_bfloat16>(fmaf(0.5f, tanh_approx(0.5f * (gi[j + k] + xi_[k])), 0.5f));
              const float a_f = round_to<__nv_bfloat16>(fmaf(0.5f, tanh_approx(0.5f * (gf[j + k] + xf_[k])), 0.5f));
              const float a_o = round_to<__nv_bfloat16>(fmaf(0.5f, tanh_approx(0.5f * (go[j + k] + xo_[k])), 0.5f));
              const float a_g = round_to<__nv_bfloat16>(tanh_approx(gg[j + k] + xg_[k]));
              gi[j + k] = a_i; gf[j + k] = a_f; go[j + k] = a_o; gg[j + k] = a_g;
              cn[j + k] = fmaf(a_f, cc_[k], a_i * a_g);
              hn[j + k] = a_o * tanh_approx(cn[j + k]);
            }
          }
          store8(ap, gi); store8(ap + 8, gi + 8);
          store8(ap + Ch, gf); store8(ap + Ch + 8, gf + 8);
          store8(ap + 2 * Ch, go); store8(ap + 2 * Ch + 8, go + 8);
          store8(ap + 3 * Ch, gg); store8(ap + 3 * Ch + 8, gg + 8);
          store8(cq, cn); store8(cq + 8, cn + 8);
          store8(hp, hn); store8(hp + 8, hn + 8);
        }
      } else
      for (int c0 = 0; c0 < p.Cout; c0 += 16) {
        float v[16];
        tmem_ld16(t_addr + c0, v);
        if (valid) {
          long long offc = off + c0;
          int cbias = c0;
          if (p.ps_co > 0) {
            const int q = c0 / p.ps_co;
            cbias = c0 - q * p.ps_co;
            offc = (long long)n * p.dst_ns + ((long long)(2 * h + (q >> 1)) * (2 * p.W) + 2 * w + (q & 1)) * p.dst_ps + cbias;
          }
          if (bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += __ldg(bias + cbias + j);
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          if (p.drop_p > 0.f) {
            const unsigned long long sd = mix_epoch(p.drop_seed, p.drop_epoch);
            const float sc = 1.f / (1.f - p.drop_p);
#pragma unroll
            for (int j = 0; j < 16; ++j)
              v[j] = dropout_uniform(sd, (unsigned long long)(offc + j)) >= p.drop_p ? v[j] * sc : 0.f;
          }
          if (p.dst_f32) {
            float* dp = reinterpret_cast<float*>(dst) + offc;
            if (p.accumulate) {
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                const float4 o = *reinterpret_cast<const float4*>(dp + j);
                v[j] += o.x; v[j + 1] += o.y; v[j + 2] += o.z; v[j + 3] += o.w;
              }
            }
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              *reinterpret_cast<float4*>(dp + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
            __nv_bfloat16* dp = reinterpret_cast<__nv_bfloat16*>(dst) + offc;
            store8(dp, v);
            store8(dp + 8, v + 8);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------
// Thin-layer variant (Cin = 16 or 32, i.e. 32/64-byte pixel rows).  TMA throughput is per box ROW
// (~4 cycles/row measured), so fetching nine shifted 32-byte-row boxes per tile starves the MMA.
// Here each tile fetches ONE halo box {Cin, Wb+2, Hb+2, Nb}; the nine taps are nine UMMA descriptors
// into the same shared-memory image whose start address is advanced by (kh*(Wb+2) + kw) pixel rows
// (the layout is dense, SBO = 8 rows, so a row shift is a pure start-address offset).  The GEMM M index
// is the flattened halo position; the two halo columns per row produce don't-care accumulator rows the
// epilogue skips.  All nine weight taps (<= 64 KB) are loaded once per CTA and stay resident.
// ------------------------------------------------------------------------------------------------
struct ConvHaloParams {
  int N, H, W, Cin, Cout;
  int Wb, Hb, Nb, Wh, Hh, tiles_w, tiles_h, num_tiles;
  int stages;
  int kchunks;          // Cin / KC: channel chunks of one swizzle row (KC = min(Cin, 64) channels) — one halo box each
  int cslice;           // output channels per CTA: blockIdx.y owns columns [y*cslice, (y+1)*cslice) and keeps the nine
                        // taps of exactly those rows of the weight resident (wide layers: the slice that fits smem)
  long long dst_ns;
  int dst_ps, dst_f32, accumulate;
  uint32_t tmem_cols, acc_stride, a_chunk_bytes, a_stage_bytes, a_tx_bytes, b_chunk_bytes, b_bytes;
  uint32_t kmask[3];    // per kernel column kw: the 16-channel K steps whose weights are not identically zero (all ones
                        // for a plain layer; the pixel-group form skips the all-zero blocks of its shifted taps)
};

template <int KSTEPS>   // KC / 16
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_tc_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       void* __restrict__ dst, const float* __restrict__ bias, unsigned int* __restrict__ err,
                       const ConvHaloParams p) {
  pdl_launch_dependents();          // the next kernel may start launching; it waits for us in its own pdl_wait()
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t b_region = (p.b_bytes + 1023u) & ~1023u;
  uint8_t* sB = smem;
  uint8_t* sA = smem + b_region;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + (size_t)p.stages * p.a_stage_bytes);
  uint64_t* full = bars;                  // [stages]
  uint64_t* empty = bars + p.stages;      // [stages]
  uint64_t* tfull = empty + p.stages;     // [2]
  uint64_t* tempty = tfull + 2;           // [2]
  uint64_t* bfull = tempty + 2;           // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bfull + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int co0 = blockIdx.y * p.cslice;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < p.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 128); }
    mbar_init(bfull, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();                       // predecessor complete: global memory may be touched from here on
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t rb = KSTEPS * 32;     // bytes of one pixel row of one channel chunk

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(bfull, p.b_bytes);
      for (int kc = 0; kc < p.kchunks; ++kc)
        tma_load_3d(sB + (size_t)kc * p.b_chunk_bytes, &tmB, bfull, kc * (int)(rb / 2), co0, 0);
      int stage = 0;
      uint32_t phase = 0;
      TileWalk tk(blockIdx.x, gridDim.x, 1, p.tiles_w, p.tiles_h);
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, tk.next()) {
        const int tw = tk.tw, th = tk.th, tn = tk.tn;
        if (!mbar_wait(&empty[stage], phase ^ 1, err)) break;
        mbar_expect_tx(&full[stage], p.a_tx_bytes);
        for (int kc = 0; kc < p.kchunks; ++kc)
          tma_load_4d(sA + (size_t)stage * p.a_stage_bytes + (size_t)kc * p.a_chunk_bytes, &tmA, &full[stage],
                      kc * (int)(rb / 2), tw * p.Wb - 1, th * p.Hb - 1, tn * p.Nb);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // Tight issue loop: per-tap descriptor increments live in registers (fully unrolled), so each
      // MMA costs a 64-bit add + the UTCHMMA (the single issuing thread is otherwise instruction-bound:
      // ncu showed ~400 cycles of address arithmetic per MMA in the first version of this loop).
      const uint32_t idesc = make_idesc_bf16(128, p.cslice, 0, 0);
      const uint32_t ltype = layout_type_for_row_bytes(rb);
      const uint64_t adesc0 = make_smem_desc(smem_u32(sA), 16, 8 * rb, ltype);
      const uint64_t bdesc0 = make_smem_desc(smem_u32(sB), 16, 8 * rb, ltype);
      uint32_t a_off[9], b_off[9];
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        a_off[tap] = ((uint32_t)((tap / 3) * p.Wh + (tap % 3)) * rb) >> 4;
        b_off[tap] = ((uint32_t)tap * p.cslice * rb) >> 4;
      }
      const uint32_t a_step = p.a_stage_bytes >> 4, a_chunk = p.a_chunk_bytes >> 4, b_chunk = p.b_chunk_bytes >> 4;
      const int nstages = p.stages, ntiles = p.num_tiles, acc_stride = p.acc_stride, kchunks = p.kchunks;
      const uint32_t km[3] = {p.kmask[0], p.kmask[1], p.kmask[2]};
      int stage = 0, it = 0;
      uint32_t phase = 0;
      bool ok = mbar_wait(bfull, 0, err);
      for (int tile = blockIdx.x; tile < ntiles && ok; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        ok = mbar_wait(&tempty[acc], acc_phase ^ 1, err);
        if (!ok) break;
        ok = mbar_wait(&full[stage], phase, err);
        if (!ok) break;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * acc_stride;
        const uint64_t abase = adesc0 + (uint64_t)(stage * a_step);
        uint32_t accum = 0;
        for (int kc = 0; kc < kchunks; ++kc) {
          const uint64_t ab = abase + (uint64_t)(kc * a_chunk), bb = bdesc0 + (uint64_t)(kc * b_chunk);
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
            for (int k = 0; k < KSTEPS; ++k) {
              if ((km[tap % 3] >> k) & 1u) {
                umma_bf16(d_tmem, ab + (uint64_t)(a_off[tap] + 2 * k), bb + (uint64_t)(b_off[tap] + 2 * k), idesc, accum);
                accum = 1;
              }
            }
          }
        }
        umma_commit(&empty[stage]);
        umma_commit(&tfull[acc]);
        if (++stage == nstages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int per_img = p.Hh * p.Wh;
    const int nl = row / per_img, rem = row % per_img;
    const int hl = rem / p.Wh, wl = rem % p.Wh;
    const bool row_ok = nl < p.Nb && hl < p.Hb && wl < p.Wb;
    int it = 0;
    bool ok = true;
    TileWalk tk(blockIdx.x, gridDim.x, 1, p.tiles_w, p.tiles_h);
    for (int tile = blockIdx.x; tile < p.num_tiles && ok; tile += gridDim.x, ++it, tk.next()) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int w = tk.tw * p.Wb + wl, h = tk.th * p.Hb + hl, n = tk.tn * p.Nb + nl;
      const bool valid = row_ok && w < p.W && h < p.H && n < p.N;
      const long long off = (long long)n * p.dst_ns + ((long long)h * p.W + w) * p.dst_ps + co0;
      ok = mbar_wait(&tfull[acc], acc_phase, err);
      ok = __all_sync(0xffffffffu, ok);
      if (!ok) break;
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * p.acc_stride;
      for (int c0 = 0; c0 < p.cslice; c0 += 16) {
        float v[16];
        tmem_ld16(t_addr + c0, v);
        if (valid) {
          if (bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += __ldg(bias + co0 + c0 + j);
          }
          if (p.dst_f32) {
            float* dp = reinterpret_cast<float*>(dst) + off + c0;
            if (p.accumulate) {
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                const float4 o = *reinterpret_cast<const float4*>(dp + j);
                v[j] += o.x; v[j + 1] += o.y; v[j + 2] += o.z; v[j + 3] += o.w;
              }
            }
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              *reinterpret_cast<float4*>(dp + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
            __nv_bfloat16* dp = reinterpret_cast<__nv_bfloat16*>(dst) + off + c0;
            store8(dp, v);
            store8(dp + 8, v + 8);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// halo tile chooser: last valid flattened row (Nb-1)*Hh*Wh + (Hb-1)*Wh + Wb-1 must be <= 127
void choose_tile_halo(int N, int H, int W, int* Wb, int* Hb, int* Nb) {
  double best = -1.0;
  for (int wb = 1; wb <= W && wb <= 126; ++wb) {
    if (W % wb != 0) continue;
    const int wh = wb + 2;
    for (int hb = 1; hb <= H; ++hb) {
      if ((hb - 1) * wh + wb - 1 > 127) break;
      int nb = 1;
      if (hb == H && wb == W) {
        const int per = (hb + 2) * wh;
        nb = 1 + (127 - ((hb - 1) * wh + wb - 1)) / per;
        if (nb > N) nb = N;
      }
      const long long tiles = (long long)(W / wb) * ((H + hb - 1) / hb) * ((N + nb - 1) / nb);
      const double eff = (double)N * H * W / ((double)tiles * 128.0);
      if (eff > best + 1e-9) { best = eff; *Wb = wb; *Hb = hb; *Nb = nb; }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tensor_map(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, int row_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return PCM_ERR_CUDA; }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUtensorMapSwizzle sw = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : row_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base),
                  reinterpret_cast<const cuuint64_t*>(dims), reinterpret_cast<const cuuint64_t*>(strides_bytes),
                  reinterpret_cast<const cuuint32_t*>(box), estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (CUresult %d)", (int)r); return PCM_ERR_CUDA; }
  return PCM_OK;
}

// choose the pixel box {Wb, Hb, Nb} (<= 128 rows) that wastes the fewest MMA rows
void choose_tile(int N, int H, int W, int* Wb, int* Hb, int* Nb) {
  double best = -1.0;
  for (int wb = 1; wb <= W && wb <= 128; ++wb) {
    if (W % wb != 0 && wb != 128) continue;
    int hmax = 128 / wb;
    if (hmax < 1) continue;
    for (int hb = 1; hb <= hmax && hb <= H; ++hb) {
      int nb = 1;
      if (hb == H && wb == W) nb = 128 / (wb * hb);
      if (nb > N) nb = N;
      if (nb < 1) nb = 1;
      const long long tiles = (long long)((W + wb - 1) / wb) * ((H + hb - 1) / hb) * ((N + nb - 1) / nb);
      const double eff = (double)N * H * W / ((double)tiles * 128.0);
      if (eff > best + 1e-9) { best = eff; *Wb = wb; *Hb = hb; *Nb = nb; }
    }
  }
}

static int g_num_sms = 0;

unsigned int* tc_error_counter() {
  static unsigned int* ptr = nullptr;
  if (ptr == nullptr) {
    if (cudaMalloc(&ptr, sizeof(unsigned int)) != cudaSuccess) { ptr = nullptr; return nullptr; }
    cudaMemset(ptr, 0, sizeof(unsigned int));
  }
  return ptr;
}

}  // namespace pcm

using namespace pcm;

extern "C" int pcm_tc_error_count(void) {
  unsigned int v = 0;
  unsigned int* p = pcm::tc_error_counter();
  if (p == nullptr || cudaMemcpy(&v, p, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  return (int)v;
}

static int conv3x3_tc_impl(const void* src, long long src_ns, int src_ps, int H, int W, int Cin, void* dst,
                           long long dst_ns, int dst_ps, int Cout, const void* wk, const float* bias, int N,
                           int dst_f32, int accumulate, const float* lstm_gx, const float* lstm_c_prev,
                           float* lstm_c_out, void* lstm_acts, pcm_stream_t s, int ksz = 3, int relu = 0, int mode = 0,
                           int ps_co = 0, int pad = -1, int group = 1, float drop_p = 0.f, long long drop_seed = 0) {
  // group > 1: pcm_conv3x3_tc_grouped — H, W, Cin, Cout are already those of the grouped image (W/g pixels of g*C channels)
  PCM_REQUIRE(Cin % 16 == 0 && Cin >= 16, "conv3x3_tc: Cin must be a multiple of 16 (got %d)", Cin);
  PCM_REQUIRE(Cin <= 64 || Cin % 64 == 0, "conv3x3_tc: Cin above 64 must be a multiple of 64 (got %d)", Cin);
  PCM_REQUIRE(Cin == 16 || Cin == 32 || Cin >= 64, "conv3x3_tc: unsupported Cin %d", Cin);
  PCM_REQUIRE(Cout % 16 == 0 && Cout >= 16 && Cout <= 256, "conv3x3_tc: Cout must be a multiple of 16 in [16,256] (got %d)", Cout);
  PCM_REQUIRE(src_ps % 8 == 0 && src_ns % 8 == 0 && dst_ps % 8 == 0, "conv3x3_tc: strides must be multiples of 8 elements");
  PCM_REQUIRE(!accumulate || dst_f32, "conv3x3_tc: accumulate needs an fp32 destination");
  PCM_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(wk) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(dst) & 15) == 0, "conv3x3_tc: pointers must be 16-byte aligned");
  if (N == 0) return PCM_OK;
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  static int halo_env = -1;
  if (halo_env < 0) {
    const char* e = getenv("PCM_TC_HALO");      // PCM_TC_HALO=0: force the nine-box kernel (A/B experiments)
    halo_env = e ? atoi(e) : 1;
  }
  unsigned int* err = tc_error_counter();
  PCM_REQUIRE(err != nullptr, "conv3x3_tc: could not allocate the error counter");
  // halo kernel: one halo box per (tile, channel chunk), nine shifted descriptors, weights resident.  Wide layers
  // keep only a slice of the output channels per CTA (blockIdx.y) so that the resident weights stay <= ~100 KB.
  int cslice = 0;
  PCM_REQUIRE(drop_p == 0.f || (drop_p > 0.f && drop_p < 1.f && ksz == 1 && ps_co == 0 && lstm_gx == nullptr),
              "conv_tc: fused dropout is wired for the 1x1 / linear form with 0 < p < 1");
  if (halo_env && ksz == 3 && mode == 0 && ps_co == 0 && !relu && lstm_gx == nullptr && Cin <= 128) {
    const size_t budget = (size_t)(Cin <= 32 ? 64 : Cin <= 64 ? 148 : 100) * 1024;
    for (int cs = Cout; cs >= 16; cs >>= 1) {
      if (Cout % cs != 0 || cs % 16 != 0) break;
      if ((size_t)9 * cs * Cin * 2 <= budget) { cslice = cs; break; }
    }
    if (Cin <= 64 && cslice != Cout) cslice = 0;        // thin layers: all-or-nothing (as measured)
    // measured: slices narrower than half the layer lose — every slice CTA re-reads the whole A tile from shared memory
    // for N = 32 columns only (128->128 in four slices: 24 us vs 17 us for the nine-box kernel; 128->64 in two: 7.4 vs 9.7)
    if (cslice > 0 && Cout / cslice > 2) cslice = 0;
  }
  if (cslice > 0) {
    ConvHaloParams h;
    h.N = N; h.H = H; h.W = W; h.Cin = Cin; h.Cout = Cout;
    h.cslice = cslice;
    const int KC = Cin < 64 ? Cin : 64;
    h.kchunks = Cin / KC;
    const int nslices = Cout / cslice;
    h.Wb = h.Hb = h.Nb = 1;
    choose_tile_halo(N, H, W, &h.Wb, &h.Hb, &h.Nb);
    h.Wh = h.Wb + 2; h.Hh = h.Hb + 2;
    h.tiles_w = W / h.Wb;
    h.tiles_h = (H + h.Hb - 1) / h.Hb;
    h.num_tiles = h.tiles_w * h.tiles_h * ((N + h.Nb - 1) / h.Nb);
    h.dst_ns = dst_ns; h.dst_ps = dst_ps; h.dst_f32 = dst_f32; h.accumulate = accumulate;
    h.kmask[0] = h.kmask[1] = h.kmask[2] = 0xffffffffu;
    if (group > 1) {
      // shifted taps of the pixel-group kernel: the group to the left contributes through its LAST pixel only, the group
      // to the right through its FIRST (pcm_pack_weight_grouped: dx = g*(s-1) + pb - pa + 1 in [0, 2])
      const int c16 = Cin / group / 16;                            // K steps per pixel of the group
      h.kmask[0] = ((1u << c16) - 1u) << ((group - 1) * c16);
      h.kmask[2] = (1u << c16) - 1u;
    }
    h.acc_stride = cslice < 32 ? 32 : cslice;
    uint32_t cols = 32;
    while (cols < 2 * h.acc_stride) cols <<= 1;
    h.tmem_cols = cols;
    const uint32_t row_bytes = KC * 2;
    const int box_rows = h.Nb * h.Hh * h.Wh;
    int need_rows = 128 + 2 * h.Wh + 2;
    if (box_rows > need_rows) need_rows = box_rows;
    h.a_chunk_bytes = ((uint32_t)need_rows * row_bytes + 1023u) & ~1023u;
    h.a_stage_bytes = h.a_chunk_bytes * h.kchunks;
    h.a_tx_bytes = (uint32_t)box_rows * row_bytes * h.kchunks;
    h.b_chunk_bytes = 9u * cslice * row_bytes;                      // multiple of 1024 for cslice >= 16, KC >= 16? checked below
    h.b_bytes = h.b_chunk_bytes * h.kchunks;
    PCM_REQUIRE(h.kchunks == 1 || h.b_chunk_bytes % 1024 == 0, "conv3x3_tc(halo): weight chunk not 1 KB aligned");
    const size_t b_region = ((size_t)h.b_bytes + 1023) & ~(size_t)1023;
    int stages = (int)(((Cin <= 32 ? 160 : 212) * 1024 - b_region) / h.a_stage_bytes);
    if (stages > 6) stages = 6;
    {
      static int cap = -1;                       // PCM_HALO_STAGES: cap the pipeline depth (A/B experiments)
      if (cap < 0) { const char* e = getenv("PCM_HALO_STAGES"); cap = e ? atoi(e) : 0; }
      if (cap > 0 && stages > cap) stages = cap;
    }
    if (stages < 2) stages = 2;
    h.stages = stages;
    const size_t smem = 1024 + b_region + (size_t)stages * h.a_stage_bytes + (2 * stages + 5) * sizeof(uint64_t) + 16;
    CUtensorMap tmA, tmB;
    {
      uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)N};
      uint64_t strides[3] = {(uint64_t)src_ps * 2, (uint64_t)W * src_ps * 2, (uint64_t)src_ns * 2};
      uint32_t box[4] = {(uint32_t)KC, (uint32_t)h.Wh, (uint32_t)h.Hh, (uint32_t)h.Nb};
      int rc = make_tensor_map(&tmA, src, 4, dims, strides, box, row_bytes);
      if (rc != PCM_OK) return rc;
    }
    {
      uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)Cout, 9};
      uint64_t strides[2] = {(uint64_t)Cin * 2, (uint64_t)Cout * Cin * 2};
      uint32_t box[3] = {(uint32_t)KC, (uint32_t)cslice, 9};
      int rc = make_tensor_map(&tmB, wk, 3, dims, strides, box, row_bytes);
      if (rc != PCM_OK) return rc;
    }
    auto kern = KC == 16 ? conv3x3_tc_halo_kernel<1> : KC == 32 ? conv3x3_tc_halo_kernel<2> : conv3x3_tc_halo_kernel<4>;
    const int hk = KC == 16 ? 0 : KC == 32 ? 1 : 2;
    static size_t smem_set_h[3] = {0, 0, 0};
    if (smem > smem_set_h[hk]) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) { set_error("conv3x3_tc(halo): smem attribute (%zu B): %s", smem, cudaGetErrorString(e)); return PCM_ERR_CUDA; }
      smem_set_h[hk] = smem;
    }
    int gx = g_num_sms / nslices;
    if (gx < 1) gx = 1;
    if (gx > h.num_tiles) gx = h.num_tiles;
    pcm::launch(kern, dim3(gx, nslices), kThreads, smem, (cudaStream_t)s, tmA, tmB, dst, bias, err, h);
    return check_launch("conv3x3_tc(halo)");
  }
  ConvTcParams p;
  p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
  choose_tile(N, H, W, &p.Wb, &p.Hb, &p.Nb);
  p.tiles_w = (W + p.Wb - 1) / p.Wb;
  p.tiles_h = (H + p.Hb - 1) / p.Hb;
  const int tiles_n = (N + p.Nb - 1) / p.Nb;
  p.num_tiles = p.tiles_w * p.tiles_h * tiles_n;
  p.KC = Cin < 64 ? Cin : 64;
  p.kchunks = Cin / p.KC;
  p.ksz = ksz; p.relu = relu; p.mode = mode; p.ps_co = ps_co;
  p.drop_p = drop_p; p.drop_seed = (unsigned long long)drop_seed; p.drop_epoch = nullptr;
  if (drop_p > 0.f) {
    p.drop_epoch = dropout_epoch_cell();
    PCM_REQUIRE(p.drop_epoch != nullptr, "conv_tc: could not allocate the dropout epoch cell");
  }
  p.pad = pad >= 0 ? pad : (ksz >> 1);
  const int ntaps = mode == 1 ? 4 : mode == 2 ? 6 : ksz * ksz;
  p.dst_ns = dst_ns; p.dst_ps = dst_ps; p.dst_f32 = dst_f32; p.accumulate = accumulate;
  const int ncols = lstm_gx != nullptr ? 64 : Cout;        // accumulator columns per work item (fused ConvLSTM: 4 x 16)
  p.acc_stride = ncols < 32 ? 32 : ncols;
  uint32_t cols = 32;
  while (cols < 2 * p.acc_stride) cols <<= 1;
  p.tmem_cols = cols;
  p.a_stage_bytes = 128u * p.KC * 2;
  p.b_stage_bytes = ((uint32_t)ncols * p.KC * 2 + 1023u) & ~1023u;
  p.a_tx_bytes = (uint32_t)(p.Wb * p.Hb * p.Nb) * p.KC * 2;
  p.b_tx_bytes = (uint32_t)ncols * p.KC * 2;
  const size_t per_stage = (size_t)p.a_stage_bytes + p.b_stage_bytes;
  int stages = (int)((200 * 1024) / per_stage);
  if (stages > 8) stages = 8;
  if (stages > ntaps * p.kchunks) stages = ntaps * p.kchunks;
  if (stages < 2) stages = 2;
  p.stages = stages;
  const size_t smem = 1024 + stages * per_stage + (2 * stages + 4) * sizeof(uint64_t) + 16;

  CUtensorMap tmA, tmA2, tmB;
  if (mode == 1) {
    // src is the (2H, 2W) gradient image; view {C, kw, w, h, n}: pixel (2h + kh, 2w + kw) with kh folded into the base
    const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(src);
    uint64_t dims[5] = {(uint64_t)Cin, 2, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t strides[4] = {(uint64_t)src_ps * 2, (uint64_t)2 * src_ps * 2, (uint64_t)4 * W * src_ps * 2, (uint64_t)src_ns * 2};
    uint32_t box[5] = {(uint32_t)p.KC, 1, (uint32_t)p.Wb, (uint32_t)p.Hb, (uint32_t)p.Nb};
    int rc = make_tensor_map(&tmA, base, 5, dims, strides, box, p.KC * 2);
    if (rc != PCM_OK) return rc;
    rc = make_tensor_map(&tmA2, base + (size_t)2 * W * src_ps, 5, dims, strides, box, p.KC * 2);
    if (rc != PCM_OK) return rc;
  } else if (mode == 2) {
    // src is the (2H, 2W) input with dense channels (Cin = 2 * pixel stride): view {(pw, c), w, ph, h, n}
    uint64_t dims[5] = {(uint64_t)Cin, (uint64_t)W, 2, (uint64_t)H, (uint64_t)N};
    uint64_t strides[4] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)2 * W * Cin * 2, (uint64_t)src_ns * 2};
    uint32_t box[5] = {(uint32_t)p.KC, (uint32_t)p.Wb, 1, (uint32_t)p.Hb, (uint32_t)p.Nb};
    int rc = make_tensor_map(&tmA, src, 5, dims, strides, box, p.KC * 2);
    if (rc != PCM_OK) return rc;
    tmA2 = tmA;
  } else {
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t strides[3] = {(uint64_t)src_ps * 2, (uint64_t)W * src_ps * 2, (uint64_t)src_ns * 2};
    uint32_t box[4] = {(uint32_t)p.KC, (uint32_t)p.Wb, (uint32_t)p.Hb, (uint32_t)p.Nb};
    int rc = make_tensor_map(&tmA, src, 4, dims, strides, box, p.KC * 2);
    if (rc != PCM_OK) return rc;
    tmA2 = tmA;
  }
  {
    uint64_t dims[3] = {(uint64_t)Cin, (uint64_t)Cout, (uint64_t)ntaps};
    uint64_t strides[2] = {(uint64_t)Cin * 2, (uint64_t)Cout * Cin * 2};
    uint32_t box[3] = {(uint32_t)p.KC, (uint32_t)(lstm_gx != nullptr ? 16 : Cout), 1};
    int rc = make_tensor_map(&tmB, wk, 3, dims, strides, box, p.KC * 2);
    if (rc != PCM_OK) return rc;
  }
  p.ngroups = lstm_gx != nullptr ? (Cout >> 2) / 16 : 1;
  p.gx = lstm_gx; p.c_prev = lstm_c_prev; p.c_out = lstm_c_out; p.acts = reinterpret_cast<__nv_bfloat16*>(lstm_acts);
  const int ksi = (p.KC == 16 ? 0 : p.KC == 32 ? 1 : 2) + (lstm_gx != nullptr ? 3 : 0);
  auto kern = ksi == 0 ? conv3x3_tc_kernel<1, 0> : ksi == 1 ? conv3x3_tc_kernel<2, 0> : ksi == 2 ? conv3x3_tc_kernel<4, 0>
            : ksi == 3 ? conv3x3_tc_kernel<1, 1> : ksi == 4 ? conv3x3_tc_kernel<2, 1> : conv3x3_tc_kernel<4, 1>;
  static size_t smem_set[6] = {0, 0, 0, 0, 0, 0};
  if (smem > smem_set[ksi]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("conv3x3_tc: smem attribute (%zu B): %s", smem, cudaGetErrorString(e)); return PCM_ERR_CUDA; }
    smem_set[ksi] = smem;
  }
  const int nitems = p.num_tiles * p.ngroups;
  const int grid = nitems < g_num_sms ? nitems : g_num_sms;
  pcm::launch(kern, grid, kThreads, smem, (cudaStream_t)s, tmA, tmA2, tmB, dst, bias, err, p);
  return check_launch("conv3x3_tc");
}

extern "C" int pcm_conv3x3_tc(const void* src, long long src_ns, int src_ps, int H, int W, int Cin, void* dst,
                              long long dst_ns, int dst_ps, int Cout, const void* wk, const float* bias, int N,
                              int dst_f32, int accumulate, pcm_stream_t s) {
  return conv3x3_tc_impl(src, src_ns, src_ps, H, W, Cin, dst, dst_ns, dst_ps, Cout, wk, bias, N, dst_f32, accumulate,
                         nullptr, nullptr, nullptr, nullptr, s);
}

extern "C" int pcm_conv3x3_tc_grouped(const void* src, long long src_ns, int src_ps, int H, int W, int Cin, void* dst,
                                      long long dst_ns, int dst_ps, int Cout, const void* wk, int N, int dst_f32,
                                      int group, pcm_stream_t s) {
  PCM_REQUIRE(group >= 1 && W % group == 0, "conv3x3_tc_grouped: W (%d) must be a multiple of the group (%d)", W, group);
  PCM_REQUIRE(src_ps == Cin && dst_ps == Cout, "conv3x3_tc_grouped: pixels must be dense (src_ps %d / Cin %d, dst_ps %d / Cout %d)",
              src_ps, Cin, dst_ps, Cout);
  PCM_REQUIRE(group * Cin <= 64 && group * Cout <= 64, "conv3x3_tc_grouped: group*Cin and group*Cout must be <= 64");
  return conv3x3_tc_impl(src, src_ns, src_ps * group, H, W / group, Cin * group, dst, dst_ns, dst_ps * group, Cout * group,
                         wk, nullptr, N, dst_f32, 0, nullptr, nullptr, nullptr, nullptr, s, 3, 0, 0, 0, -1, group);
}

extern "C" int pcm_convlstm_step_tc(const void* h_prev, const void* wh, const float* gx, const float* c_prev,
                                    void* h_out, float* c_out, void* acts, int B, int H, int W, int Ch,
                                    pcm_stream_t s) {
  PCM_REQUIRE(gx != nullptr && c_prev != nullptr && c_out != nullptr && acts != nullptr && h_out != nullptr,
              "convlstm_step_tc: null pointer");
  PCM_REQUIRE(Ch % 16 == 0 && 4 * Ch <= 256, "convlstm_step_tc: Ch must be a multiple of 16, <= 64 (got %d)", Ch);
  return conv3x3_tc_impl(h_prev, (long long)H * W * Ch, Ch, H, W, Ch, h_out, (long long)H * W * Ch, Ch, 4 * Ch, wh,
                         nullptr, B, 0, 0, gx, c_prev, c_out, acts, s);
}

extern "C" int pcm_conv1x1_tc(const void* src, long long src_ns, int src_ps, int H, int W, int Cin, void* dst,
                              long long dst_ns, int dst_ps, int Cout, const void* wk, const float* bias, int N,
                              int dst_f32, int accumulate, int relu, pcm_stream_t s) {
  return conv3x3_tc_impl(src, src_ns, src_ps, H, W, Cin, dst, dst_ns, dst_ps, Cout, wk, bias, N, dst_f32, accumulate,
                         nullptr, nullptr, nullptr, nullptr, s, 1, relu);
}

// nn.Linear -> ReLU -> nn.Dropout(p) (the inner half of the transformer FFN, src/cnn_transformer.py:25-31) in one launch:
// the dropout mask of pcm_dropout on the dense destination is applied in the store epilogue.
extern "C" int pcm_conv1x1_drop_tc(const void* src, long long src_ns, int src_ps, int H, int W, int Cin, void* dst,
                                   long long dst_ns, int dst_ps, int Cout, const void* wk, const float* bias, int N, int relu,
                                   float drop_p, long long seed, pcm_stream_t s) {
  PCM_REQUIRE(dst_ps == Cout && dst_ns == (long long)H * W * Cout, "conv1x1_drop_tc: the destination must be dense");
  return conv3x3_tc_impl(src, src_ns, src_ps, H, W, Cin, dst, dst_ns, dst_ps, Cout, wk, bias, N, 0, 0, nullptr, nullptr, nullptr,
                         nullptr, s, 1, relu, 0, 0, -1, 1, drop_p, seed);
}

// ConvTranspose2d(kernel 2, stride 2) forward: one GEMM [pixels x Cin] x [Cin x 4*Cout] whose epilogue scatters
// quadrant q = (kh, kw) of every input pixel to output pixel (2h + kh, 2w + kw).
extern "C" int pcm_convT2x2_tc(const void* src, long long src_ns, int src_ps, int H, int W, int Cin, void* dst,
                               long long dst_ns, int dst_ps, int Cout, const void* wk, const float* bias, int N,
                               int relu, pcm_stream_t s) {
  PCM_REQUIRE(Cout % 16 == 0 && 4 * Cout <= 256, "convT2x2_tc: Cout must be a multiple of 16, <= 64 (got %d)", Cout);
  return conv3x3_tc_impl(src, src_ns, src_ps, H, W, Cin, dst, dst_ns, dst_ps, 4 * Cout, wk, bias, N, 0, 0, nullptr, nullptr,
                         nullptr, nullptr, s, 1, relu, 0, Cout);
}

// ... and its data gradient: dx(n,h,w,ci) = sum_{kh,kw,co} dy(n, 2h+kh, 2w+kw, co) * wk[kh*2+kw][ci][co]
extern "C" int pcm_convT2x2_dgrad_tc(const void* dy, long long dy_ns, int dy_ps, int H, int W, int Cout, void* dx,
                                     long long dx_ns, int dx_ps, int Cin, const void* wk, int N, pcm_stream_t s) {
  return conv3x3_tc_impl(dy, dy_ns, dy_ps, H, W, Cout, dx, dx_ns, dx_ps, Cin, wk, nullptr, N, 0, 0, nullptr, nullptr, nullptr,
                         nullptr, s, 1, 0, 1, 0);
}

// 3x3 / stride-2 / pad-1 convolution (src/cnn_transformer.py:10,12) on the tensor cores.  src: (2H, 2W) image with Cs
// dense channels (pixel stride == Cs); the GEMM sees pixel pairs: K = 2*Cs per tap, 6 taps (kh, dw).
// wk: bf16 [6][Cout][2*Cs], tap kh*2 + (dw+1), channel pw*Cs + c = w[co][c][kh][2*(dw+1) + pw - 1] (0 where that kw < 0).
extern "C" int pcm_conv3x3s2_tc(const void* src, long long src_ns, int Cs, int H, int W, void* dst, long long dst_ns,
                                int dst_ps, int Cout, const void* wk, const float* bias, int N, int relu, pcm_stream_t s) {
  PCM_REQUIRE(Cs % 8 == 0, "conv3x3s2_tc: source channels must be a multiple of 8 (got %d)", Cs);
  return conv3x3_tc_impl(src, src_ns, Cs, H, W, 2 * Cs, dst, dst_ns, dst_ps, Cout, wk, bias, N, 0, 0, nullptr, nullptr, nullptr,
                         nullptr, s, 3, relu, 2, 0);
}

// ... and its data gradient: dx(2h+ph, 2w+pw, c) = sum_{oh,ow in {0,1}} sum_co dy(h+oh, w+ow, co) * wk[oh*2+ow][(ph*2+pw)*Cq + c][co]
// (a 2x2-tap convolution of dy without leading padding whose 4*Cq output columns are pixel-shuffled into the (2H, 2W)
// gradient image).  Cq: channel count (= pixel stride granularity) of dx, multiple of 16, 4*Cq <= 256.
extern "C" int pcm_conv3x3s2_dgrad_tc(const void* dy, long long dy_ns, int dy_ps, int H, int W, int Cout, void* dx,
                                      long long dx_ns, int dx_ps, int Cq, const void* wk, int N, pcm_stream_t s) {
  PCM_REQUIRE(Cq % 16 == 0 && 4 * Cq <= 256, "conv3x3s2_dgrad_tc: Cq must be a multiple of 16, <= 64 (got %d)", Cq);
  return conv3x3_tc_impl(dy, dy_ns, dy_ps, H, W, Cout, dx, dx_ns, dx_ps, 4 * Cq, wk, nullptr, N, 0, 0, nullptr, nullptr, nullptr,
                         nullptr, s, 2, 0, 0, Cq, 0);
}
