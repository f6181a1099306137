// Weight gradient of the 3x3 / stride-1 / pad-1 convolution as a tcgen05 GEMM over the PIXEL dimension:
//
//   dW[co][ci][tap] += sum_p dy[p][co] * x[p + shift(tap)][ci]        (K = pixels, M = co, N = ci per tap)
//
// Both operands are used exactly as they sit in HBM (NHWC): a TMA box {C-chunk, W+2, rows, images}
// lands in shared memory as consecutive pixel rows of C*2 bytes == the canonical MN-major UMMA operand
// layout (channels contiguous, K = pixel rows in 8-row groups).  The x box carries a one-pixel halo, the
// dy box is fetched (W+2) wide with its two extra columns out of bounds (= zero), so both share the row
// pitch W+2 and the nine taps are nine start-address offsets (kh*(W+2)+kw rows) into the SAME x tile.
// All taps of a group accumulate side by side in TMEM (tap-major columns); K is split over CTAs and
// the fp32 partials are reduced with atomics straight into the parameter-gradient tensor.
//
// Roles as in conv_tc.cu: warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue.
#include <stdlib.h>

#include "tc_common.cuh"

namespace pcm {

using namespace tc;

struct WgradTcParams {
  int N, H, W, Wp;
  int Co_real, Ci_real, Ci;     // Ci = padded N extent per tap (multiple of 16)
  int Cca, Ccb, na_chunks, nb_chunks;
  int Hb, Nb, tiles_h, num_ktiles, Kpad;
  int tpg, ngroups;             // taps per group, groups
  int ksz, ntaps;               // kernel size (1 or 3), ksz*ksz
  int wide;                     // 1: one MMA covers the ksz taps of a kernel row (N = ksz*Ci, LBO = one pixel row)
  int convt;                    // 1: ConvTranspose2d(k2,s2) weight gradient — 4 taps (kh,kw); tap q's B tile is its own
                                //    sub-tile, gathered from pixels (2h+kh, 2w+kw) through the 5-D views tmB / tmB2
  uint32_t b_sub_bytes;         // convt: bytes of one tap's B sub-tile (all chunks)
  int group, Cog, Cig;          // pixel-group form (group > 1): `group` adjacent pixels of a row are one K row with group*C
                                // channels; Cog / Cig = channels of ONE pixel (M index = pa*Cog + co, B row = pb*Cig + ci)
  int ncol, sub16;              // wide / group: N extent of the MMA of one kernel row; group: B start offset inside a row (>>4)
  uint32_t bar_off;             // barriers sit behind the pipeline stages or the epilogue's staging array, whichever is larger
  int noatomic;                 // PCM_WGRAD_NOATOMIC=1 (measurement only): skip the global reductions of the epilogue
  int stages;
  int bx0;                      // first column of the x box in its tensor map (-pad; 0 for an interior column strip)
  long long sa, sb, st;
  uint32_t a_chunk_bytes, b_chunk_bytes, a_stage_bytes, b_stage_bytes, tx_bytes, tmem_cols, lbo_a, lbo_b;
};

constexpr int kWgThreads = 192;

// Epilogue of the pixel-group form: accumulator lane = (pa, co) (output pixel pa of the group), column = (kh, j, ci) with
// input pixel p0 - 1 + j; tap dx = j - pa.  The g values of one (co, kh, dx, ci) sit in lanes of different pa, mostly in
// different warps: every lane stores its three tap blocks per kernel row into its pa's copy of a [g][9][Cog][Cig] fp32
// array in shared memory (the pipeline stages are idle by now), one barrier, and the 128 epilogue threads sum the g
// copies and issue ONE coalesced vector reduction per 4 outputs — as many atomics as the plain form, on consecutive
// addresses.  TMEM is read a whole kernel row at a time (one tcgen05.wait::ld per row, not per 16 columns).
template <int COG, int CIG>
__device__ __forceinline__ void wgrad_group_epilogue(uint8_t* smem, uint32_t tmem_base, float* __restrict__ dw,
                                                     const WgradTcParams& p, int q, int lane) {
  constexpr int PAW = 32 / COG;                   // group pixels held by one warp (2 for 16 channels, 1 for 32)
  constexpr int NB = PAW + 2;                     // input-pixel blocks of CIG columns this warp needs per kernel row
  const int g = p.group;
  const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16);
  const int m = q * 32 + lane, pa_l = m / COG, cog = m - pa_l * COG, pa_first = (q * 32) / COG;
  float* acc = reinterpret_cast<float*>(smem);    // [g][9][COG][CIG]
  if (pa_first < g) {
#pragma unroll 1
    for (int kh = 0; kh < 3; ++kh) {
      uint32_t r[NB * CIG];
#pragma unroll
      for (int c = 0; c < NB * CIG; c += 16) tc::tmem_ld16_nowait(t_addr + kh * p.ncol + pa_first * CIG + c, r + c);
      tc::tmem_wait_ld();
      const bool hi = PAW == 2 && pa_l != pa_first;          // this lane's blocks start one block further
#pragma unroll
      for (int dxx = 0; dxx < 3; ++dxx) {
        float4* d = reinterpret_cast<float4*>(acc + (((size_t)pa_l * 9 + kh * 3 + dxx) * COG + cog) * CIG);
#pragma unroll
        for (int k = 0; k < CIG / 4; ++k) {
          const int b0 = dxx * CIG + 4 * k, b1 = PAW == 2 ? b0 + CIG : b0;
          float4 o;
          o.x = __uint_as_float(hi ? r[b1] : r[b0]);
          o.y = __uint_as_float(hi ? r[b1 + 1] : r[b0 + 1]);
          o.z = __uint_as_float(hi ? r[b1 + 2] : r[b0 + 2]);
          o.w = __uint_as_float(hi ? r[b1 + 3] : r[b0 + 3]);
          d[k] = o;
        }
      }
    }
  }
  asm volatile("bar.sync 1, 128;" ::: "memory");
  const int et = threadIdx.x - 64;                // 0..127 over the epilogue warps
  constexpr int nv4 = 9 * COG * CIG / 4, per_row = CIG / 4;
  for (int i = et; i < nv4; i += 128) {
    const int row = i / per_row, c4 = i - row * per_row;                        // row = t*COG + co
    const int t = row / COG, co = row - t * COG;
    float4 v = reinterpret_cast<const float4*>(acc)[i];
    for (int pa = 1; pa < g; ++pa) {
      const float4 u = reinterpret_cast<const float4*>(acc)[pa * nv4 + i];
      v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
    }
    float* base = dw + (long long)co * p.sa + (long long)t * p.st;
    if (p.sb == 1) {
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(base + c4 * 4), "f"(v.x), "f"(v.y), "f"(v.z),
                   "f"(v.w) : "memory");
    } else {
      const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (c4 * 4 + k < p.Ci_real) atomicAdd(base + (long long)(c4 * 4 + k) * p.sb, e[k]);
    }
  }
}

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad3x3_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmB2, float* __restrict__ dw, unsigned int* __restrict__ err, const WgradTcParams p) {
  pdl_launch_dependents();          // the next kernel may start launching; it waits for us in its own pdl_wait()
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)p.stages * p.a_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.bar_off);
  uint64_t* full = bars;
  uint64_t* empty = bars + p.stages;
  uint64_t* done = empty + p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = blockIdx.x, nsplit = gridDim.x, mtile = blockIdx.y, group = blockIdx.z;
  const int tap0 = group * p.tpg;
  const int ntap = min(p.tpg, p.ntaps - tap0);

  // rows TMA never writes (K padding, shifted-view overrun) must read as zero: clear the ring once
  {
    const uint32_t total16 = (uint32_t)(((size_t)p.stages * (p.a_stage_bytes + p.b_stage_bytes)) >> 4);
    uint4* z = reinterpret_cast<uint4*>(smem);
    for (uint32_t i = threadIdx.x; i < total16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < p.stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(done, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();                       // predecessor complete: global memory may be touched from here on
  const uint32_t tmem_base = *tmem_slot;
  const int my_tiles = (p.num_ktiles - split + nsplit - 1) / nsplit;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kt = split; kt < p.num_ktiles; kt += nsplit) {
        const int th = kt % p.tiles_h, tn = kt / p.tiles_h;
        const int h0 = th * p.Hb, n0 = tn * p.Nb;
        if (!mbar_wait(&empty[stage], phase ^ 1, err)) break;
        mbar_expect_tx(&full[stage], p.tx_bytes);
        uint8_t* a = sA + (size_t)stage * p.a_stage_bytes;
        uint8_t* b = sB + (size_t)stage * p.b_stage_bytes;
        for (int c = 0; c < p.na_chunks; ++c)
          tma_load_4d(a + (size_t)c * p.a_chunk_bytes, &tmA, &full[stage], mtile * 128 + c * p.Cca, 0, h0, n0);
        if (p.convt == 1) {
          for (int q = 0; q < 4; ++q)
            for (int c = 0; c < p.nb_chunks; ++c)
              tma_load_5d(b + (size_t)q * p.b_sub_bytes + (size_t)c * p.b_chunk_bytes, (q >> 1) ? &tmB2 : &tmB, &full[stage],
                          c * p.Ccb, q & 1, 0, h0, n0);
        } else if (p.convt == 2) {
          // 3x3 / stride-2 conv: tap q = kh*2 + (dw+1) reads rows 2h+kh-1 = (ph, h+oh) and pixel pairs w+dw of the input,
          // viewed as {(pw, c), w, ph, h, n}
          for (int q = 0; q < p.ntaps; ++q) {
            const int kh = q >> 1;
            for (int c = 0; c < p.nb_chunks; ++c)
              tma_load_5d(b + (size_t)q * p.b_sub_bytes + (size_t)c * p.b_chunk_bytes, &tmB, &full[stage], c * p.Ccb, (q & 1) - 1,
                          kh != 1 ? 1 : 0, h0 - (kh == 0 ? 1 : 0), n0);
          }
        } else {
          for (int c = 0; c < p.nb_chunks; ++c)
            tma_load_4d(b + (size_t)c * p.b_chunk_bytes, &tmB, &full[stage], c * p.Ccb, p.bx0, h0 - (p.ksz >> 1), n0);
        }
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t rba = p.Cca * 2, rbb = p.Ccb * 2;
      // A (dy) and B (x) are MN-major.  `wide`: the ksz taps of one kernel row are ksz consecutive pixel shifts of
      // the x tile, i.e. ksz N-chunks one pixel row (LBO) apart -> ONE MMA with N = ksz*Ci per kernel row.  The A
      // operand (128 rows, re-read from shared memory by every MMA) is then read ksz times less often — for the
      // thin layers the kernel was bound by exactly that read (ncu: 68 cycles per N=16 MMA, tensor pipe 12 %).
      const int wide = p.wide;
      const uint32_t idesc = make_idesc_bf16(128, wide ? p.ncol : p.Ci, 1, 1);
      const uint64_t adesc0 = make_smem_desc(smem_u32(sA), p.lbo_a, 8 * rba, layout_type_for_row_bytes(rba));
      const uint64_t bdesc0 = make_smem_desc(smem_u32(sB), wide ? rbb : p.lbo_b, 8 * rbb, layout_type_for_row_bytes(rbb));
      const uint32_t a_step = p.a_stage_bytes >> 4, b_step = p.b_stage_bytes >> 4;
      const uint32_t a_k = rba, b_k = rbb;                                 // 16 pixel rows = 16*rb bytes -> (>>4) = rb
      const uint32_t b_tap_row = (uint32_t)p.Wp * rbb >> 4, b_px = rbb >> 4;
      const int ksteps = p.Kpad / 16, nstages = p.stages, ci = p.Ci;
      // per-issue offsets of this group: B start-address offset and TMEM column offset (at most 9)
      uint32_t b_off[9], d_off[9];
      int nissue = 0;
      if (p.group > 1) {
        // pixel-group form: output pixel pa of a group sees input pixels p0-1 .. p0+g of the row (p0 = first pixel of the
        // group) — g+2 CONSECUTIVE pixels of the pixel-linear tile, starting g-1 pixels into the group to the left: one MMA
        // per kernel row with N = (g+2)*Cig, its B start address advanced by that sub-row offset (the swizzle is a function
        // of the absolute address, so a 16-byte-granular start inside a row is exact like the row shifts are)
        for (int kh = 0; kh < 3; ++kh) {
          b_off[nissue] = (uint32_t)kh * b_tap_row + (uint32_t)p.sub16;
          d_off[nissue++] = (uint32_t)(kh * p.ncol);
        }
      } else if (wide) {
        for (int tl = 0; tl < ntap; tl += p.ksz) {
          b_off[nissue] = (uint32_t)((tap0 + tl) / p.ksz) * b_tap_row;
          d_off[nissue++] = (uint32_t)(tl * ci);
        }
      } else {
        for (int tl = 0; tl < ntap; ++tl) {
          const int tap = tap0 + tl;
          b_off[nissue] = p.convt ? (uint32_t)tap * (p.b_sub_bytes >> 4)
                                  : (uint32_t)(tap / p.ksz) * b_tap_row + (uint32_t)(tap % p.ksz) * b_px;
          d_off[nissue++] = (uint32_t)(tl * ci);
        }
      }
      int stage = 0;
      uint32_t phase = 0;
      bool ok = true;
      for (int t = 0; t < my_tiles && ok; ++t) {
        ok = mbar_wait(&full[stage], phase, err);
        if (!ok) break;
        tc_fence_after();
        const uint64_t ad = adesc0 + (uint64_t)(stage * a_step);
        const uint64_t bd = bdesc0 + (uint64_t)(stage * b_step);
        for (int i = 0; i < nissue; ++i) {
          const uint64_t bt = bd + (uint64_t)b_off[i];
          const uint32_t d = tmem_base + d_off[i];
          for (int k = 0; k < ksteps; ++k)
            umma_bf16(d, ad + (uint64_t)(k * a_k), bt + (uint64_t)(k * b_k), idesc, (t | k) != 0);
        }
        umma_commit(&empty[stage]);
        if (++stage == nstages) { stage = 0; phase ^= 1; }
      }
      umma_commit(done);
    }
  } else {
    // epilogue: TMEM lane = output channel, columns = (tap, ci)
    const int q = warp & 3;
    const int co = mtile * 128 + q * 32 + lane;
    bool ok = my_tiles > 0 ? mbar_wait(done, 0, err) : false;
    ok = __all_sync(0xffffffffu, ok) && !p.noatomic;
    if (ok && p.group > 1) {
      tc_fence_after();
      if (p.Cog == 16 && p.Cig == 16) wgrad_group_epilogue<16, 16>(smem, tmem_base, dw, p, q, lane);
      else if (p.Cog == 16) wgrad_group_epilogue<16, 32>(smem, tmem_base, dw, p, q, lane);
      else if (p.Cig == 16) wgrad_group_epilogue<32, 16>(smem, tmem_base, dw, p, q, lane);
      else wgrad_group_epilogue<32, 32>(smem, tmem_base, dw, p, q, lane);
    } else if (ok) {
      tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16);
      const int ncols = ntap * p.Ci;
      // 64 accumulator columns per tcgen05.wait::ld (four loads in flight): the epilogue is a serial tail behind the K loop
      // of every CTA, and one wait per 16 columns made it 5-10 us long on the wide layers
      for (int c00 = 0; c00 < ncols; c00 += 64) {
        uint32_t r[64];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (c00 + 16 * k < ncols) tmem_ld16_nowait(t_addr + c00 + 16 * k, r + 16 * k);
        tmem_wait_ld();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int c0 = c00 + 16 * k;
          if (c0 < ncols && co < p.Co_real) {
            const float* v = reinterpret_cast<const float*>(r + 16 * k);
            const int tap = tap0 + c0 / p.Ci;
            const int ci0 = c0 % p.Ci;
            float* base = dw + (long long)co * p.sa + (long long)tap * p.st;
            if (p.sb == 1) {
              // packed gradient layout [tap][co][ci] (ci contiguous, padded): 16-byte vector reductions — one L2
              // sector operation per 4 values instead of one per value (the scattered scalar atomics of the
              // reference layout were the whole cost of the wide layers: 49 splits x 147k values for enc4)
              float* qd = base + ci0;
#pragma unroll
              for (int j = 0; j < 16; j += 4)
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(qd + j), "f"(v[j]), "f"(v[j + 1]),
                             "f"(v[j + 2]), "f"(v[j + 3]) : "memory");
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (ci0 + j < p.Ci_real) atomicAdd(base + (long long)(ci0 + j) * p.sb, v[j]);
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

static int g_wg_sms = 0;

}  // namespace pcm

using namespace pcm;

static int wgrad_tc_impl(const void* dy, long long dy_ns, int dy_ps, int Co, int Co_real, const void* x,
                         long long x_ns, int x_ps, int Ci, int Ci_real, float* dw, long long sa, long long sb,
                         long long st, int N, int H, int W, int ksz, pcm_stream_t s, int convt = 0, int Wfull = 0,
                         int w0 = 0, int group = 1) {
  // group > 1 (pixel-group form, 3x3 only): the kernel sees images of W/group pixels with group*C channels; Ci_real keeps
  // its per-pixel meaning
  const int Cog = Co, Cig = Ci;
  if (group > 1) {
    PCM_REQUIRE(convt == 0 && ksz == 3 && Wfull <= 0 && W % group == 0, "wgrad3x3_tc: pixel groups need a 3x3 layer and W %% group == 0");
    PCM_REQUIRE(dy_ps == Co && x_ps == Ci && Co == Co_real, "wgrad3x3_tc: pixel groups need dense pixels");
    PCM_REQUIRE(group * Co <= 64 && group * Ci <= 64, "wgrad3x3_tc: group*C must be <= 64");
    Co *= group; Co_real = Co; Ci *= group; W /= group; dy_ps *= group; x_ps *= group;
  }
  // Wfull > 0: this call covers the column strip [w0, w0 + W) of images that are Wfull pixels wide (grids whose rows do
  // not fit one TMA box, e.g. 360 columns): dy is viewed as a W-wide tensor starting at column w0 (everything outside
  // the strip is out of bounds = zero, so it contributes nothing), x as the strip plus its real neighbour columns.
  const int pad = ksz >> 1, ntaps = convt == 1 ? 4 : convt == 2 ? 6 : ksz * ksz;
  if (Wfull <= 0) { Wfull = W; w0 = 0; }
  PCM_REQUIRE(Wfull == W || (convt == 0 && ksz == 3), "wgrad3x3_tc: column strips are implemented for the 3x3 stride-1 form");
  PCM_REQUIRE(Co % 16 == 0 && (Co <= 64 ? (Co == 16 || Co == 32 || Co == 64) : Co % 128 == 0),
              "wgrad3x3_tc: Co must be 16, 32, 64 or a multiple of 128 (got %d)", Co);
  PCM_REQUIRE(Ci % 16 == 0 && (Ci <= 64 ? (Ci == 16 || Ci == 32 || Ci == 64) : (Ci % 64 == 0 && Ci <= 256)),
              "wgrad3x3_tc: Ci must be 16, 32, 64, 128, 192 or 256 (got %d)", Ci);
  PCM_REQUIRE(W + 2 * pad <= 256 && H + 2 * pad <= 256, "wgrad3x3_tc: grid too large for one TMA box (W=%d H=%d)", W, H);
  PCM_REQUIRE(dy_ps % 8 == 0 && x_ps % 8 == 0 && dy_ns % 8 == 0 && x_ns % 8 == 0, "wgrad3x3_tc: strides must be multiples of 8");
  PCM_REQUIRE(sb != 1 || ((reinterpret_cast<uintptr_t>(dw) & 15) == 0 && sa % 4 == 0 && st % 4 == 0),
              "wgrad3x3_tc: the packed layout (sb == 1) needs a 16-byte aligned dw and sa, st multiples of 4");
  if (N == 0) return PCM_OK;
  if (g_wg_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_wg_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_wg_sms <= 0) g_wg_sms = 148;
  }
  WgradTcParams p;
  p.N = N; p.H = H; p.W = W; p.Wp = W + 2 * pad;
  p.bx0 = -pad;
  p.ksz = ksz; p.ntaps = ntaps; p.convt = convt;
  p.Co_real = Co_real; p.Ci_real = Ci_real; p.Ci = Ci;
  p.Cca = Co < 64 ? Co : 64;
  p.Ccb = Ci < 64 ? Ci : 64;
  const int m_extent = Co < 128 ? Co : 128;
  p.na_chunks = m_extent / p.Cca;
  p.nb_chunks = Ci / p.Ccb;
  const int mtiles = (Co + 127) / 128;
  p.sa = sa; p.sb = sb; p.st = st;
  // K tile: whole images (with their zero rows) when small, else a block of image rows; capped so that one
  // pipeline stage (A: m_extent channels, B: Ci channels, bf16) stays near 100 KB (two stages fit)
  const int rows_full = (H + 2 * pad) * p.Wp;
  int rows_cap = (int)((100 * 1024) / ((size_t)(m_extent + (convt ? ntaps : 1) * Ci) * 2)) - 2 * pad * p.Wp - 2 * pad;
  if (rows_cap > 240) rows_cap = 240;
  if (rows_cap < 16) rows_cap = 16;
  int a_box_h, b_box_h;
  if (rows_full <= 128 && rows_full <= rows_cap) {
    p.Nb = (rows_cap + 16) / rows_full;
    if (p.Nb > N) p.Nb = N;
    if (p.Nb < 1) p.Nb = 1;
    p.Hb = H;
    a_box_h = H + 2 * pad; b_box_h = H + 2 * pad;
    p.tiles_h = 1;
  } else {
    p.Nb = 1;
    int hb = rows_cap / p.Wp;
    if (hb < 1) hb = 1;
    if (hb > H) hb = H;
    p.tiles_h = (H + hb - 1) / hb;
    p.Hb = (H + p.tiles_h - 1) / p.tiles_h;
    a_box_h = p.Hb; b_box_h = p.Hb + 2 * pad;
  }
  const int tiles_n = (N + p.Nb - 1) / p.Nb;
  p.num_ktiles = tiles_n * p.tiles_h;
  const int a_rows = p.Nb * a_box_h * p.Wp;
  const int b_rows = p.Nb * b_box_h * p.Wp;
  p.Kpad = (a_rows + 15) / 16 * 16;
  const uint32_t rba = p.Cca * 2, rbb = p.Ccb * 2;
  // M blocks beyond the real channels alias the tile shifted by 8 rows per block: needs <= 56 rows of slack
  const int a_alloc = p.Kpad + (p.na_chunks * p.Cca >= 128 ? 0 : 64);
  int b_alloc = p.Kpad + 2 * pad * p.Wp + 2 * pad;
  if (b_alloc < b_rows) b_alloc = b_rows;
  p.a_chunk_bytes = ((uint32_t)a_alloc * rba + 1023u) & ~1023u;
  p.b_chunk_bytes = ((uint32_t)b_alloc * rbb + 1023u) & ~1023u;
  p.a_stage_bytes = p.a_chunk_bytes * p.na_chunks;
  p.b_sub_bytes = p.b_chunk_bytes * p.nb_chunks;
  p.b_stage_bytes = p.b_sub_bytes * (convt ? ntaps : 1);
  p.tx_bytes = (uint32_t)a_rows * rba * p.na_chunks + (uint32_t)b_rows * rbb * p.nb_chunks * (convt ? ntaps : 1);
  // M blocks beyond the real channels alias the tile shifted by 8 rows (results unused, reads stay in bounds)
  p.lbo_a = (p.na_chunks * p.Cca >= 128) ? p.a_chunk_bytes : 8 * rba;
  p.lbo_b = p.b_chunk_bytes;
  static int wide_env = -1;
  if (wide_env < 0) {
    const char* e = getenv("PCM_WGRAD_WIDE");     // PCM_WGRAD_WIDE=0: one MMA per tap (A/B experiments)
    wide_env = e ? atoi(e) : 1;
  }
  p.wide = (wide_env && !convt && ksz == 3 && p.nb_chunks == 1 && 3 * Ci <= 256) ? 1 : 0;
  p.group = group; p.Cog = Cog; p.Cig = Cig;
  {
    static int na = -1;
    if (na < 0) { const char* e = getenv("PCM_WGRAD_NOATOMIC"); na = e ? atoi(e) : 0; }
    p.noatomic = na;
  }
  p.ncol = ksz * Ci; p.sub16 = 0;
  if (group > 1) {
    p.wide = 1;
    p.ncol = (group + 2) * Cig; p.sub16 = (group - 1) * Cig * 2 / 16;
    PCM_REQUIRE(3 * p.ncol <= 512 && p.ncol <= 256, "wgrad3x3_tc: pixel-group accumulator does not fit tensor memory");
    p.tpg = 9; p.ngroups = 1;
  } else if (p.wide) {
    int rows = 512 / (3 * Ci);                                  // whole kernel rows per group (TMEM: 512 columns)
    if (rows > 3) rows = 3;
    p.tpg = 3 * rows;
    p.ngroups = (3 + rows - 1) / rows;
  } else {
    p.tpg = 512 / Ci;
    if (p.tpg > ntaps) p.tpg = ntaps;
    p.ngroups = (ntaps + p.tpg - 1) / p.tpg;
    p.tpg = (ntaps + p.ngroups - 1) / p.ngroups;                // balance the groups
  }
  uint32_t cols = 32;
  while (cols < (uint32_t)(group > 1 ? 3 * p.ncol : p.tpg * Ci)) cols <<= 1;
  p.tmem_cols = cols;
  const size_t per_stage = (size_t)p.a_stage_bytes + p.b_stage_bytes;
  int stages = (int)((200 * 1024) / per_stage);
  if (stages > 4) stages = 4;
  PCM_REQUIRE(stages >= 1, "wgrad3x3_tc: tile does not fit shared memory (%zu B per stage)", per_stage);
  if (stages > p.num_ktiles) stages = p.num_ktiles;
  p.stages = stages;
  size_t pipe_bytes = stages * per_stage;
  if (group > 1 && pipe_bytes < (size_t)group * 9 * Cog * Cig * 4) pipe_bytes = (size_t)group * 9 * Cog * Cig * 4;
  p.bar_off = (uint32_t)((pipe_bytes + 15) & ~(size_t)15);
  const size_t smem = 1024 + p.bar_off + (2 * stages + 1) * sizeof(uint64_t) + 16;

  CUtensorMap tmA, tmB, tmB2;
  {
    uint64_t dims[4] = {(uint64_t)Co, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t strides[3] = {(uint64_t)dy_ps * 2, (uint64_t)Wfull * dy_ps * 2, (uint64_t)dy_ns * 2};
    uint32_t box[4] = {(uint32_t)p.Cca, (uint32_t)p.Wp, (uint32_t)a_box_h, (uint32_t)p.Nb};
    int rc = make_tensor_map(&tmA, reinterpret_cast<const __nv_bfloat16*>(dy) + (size_t)w0 * dy_ps, 4, dims, strides, box, rba);
    if (rc != PCM_OK) return rc;
  }
  if (convt == 2) {
    // B is the (2H, 2W) conv input with dense channels Ci/2: view {(pw, c), w, ph, h, n}
    uint64_t dims[5] = {(uint64_t)Ci, (uint64_t)W, 2, (uint64_t)H, (uint64_t)N};
    uint64_t strides[4] = {(uint64_t)Ci * 2, (uint64_t)W * Ci * 2, (uint64_t)2 * W * Ci * 2, (uint64_t)x_ns * 2};
    uint32_t box[5] = {(uint32_t)p.Ccb, (uint32_t)p.Wp, 1, (uint32_t)b_box_h, (uint32_t)p.Nb};
    int rc = make_tensor_map(&tmB, x, 5, dims, strides, box, rbb);
    if (rc != PCM_OK) return rc;
    tmB2 = tmB;
  } else if (convt) {
    // B is the (2H, 2W) image; view {C, kw, w, h, n}: pixel (2h + kh, 2w + kw), kh folded into the base pointer
    const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(x);
    uint64_t dims[5] = {(uint64_t)Ci, 2, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t strides[4] = {(uint64_t)x_ps * 2, (uint64_t)2 * x_ps * 2, (uint64_t)4 * W * x_ps * 2, (uint64_t)x_ns * 2};
    uint32_t box[5] = {(uint32_t)p.Ccb, 1, (uint32_t)p.Wp, (uint32_t)b_box_h, (uint32_t)p.Nb};
    int rc = make_tensor_map(&tmB, base, 5, dims, strides, box, rbb);
    if (rc != PCM_OK) return rc;
    rc = make_tensor_map(&tmB2, base + (size_t)2 * W * x_ps, 5, dims, strides, box, rbb);
    if (rc != PCM_OK) return rc;
  } else {
    // x strip with its neighbour columns: an interior strip starts one REAL column to the left (box column 0), the
    // first strip starts at the image edge (box column -1 = out of bounds = the zero padding)
    const int xb = w0 > 0 ? w0 - 1 : 0;
    int xw = (w0 > 0 ? W + 2 : W + 1);
    if (xb + xw > Wfull) xw = Wfull - xb;
    if (w0 > 0) p.bx0 = 0;
    uint64_t dims[4] = {(uint64_t)Ci, (uint64_t)(Wfull == W ? W : xw), (uint64_t)H, (uint64_t)N};
    uint64_t strides[3] = {(uint64_t)x_ps * 2, (uint64_t)Wfull * x_ps * 2, (uint64_t)x_ns * 2};
    uint32_t box[4] = {(uint32_t)p.Ccb, (uint32_t)p.Wp, (uint32_t)b_box_h, (uint32_t)p.Nb};
    int rc = make_tensor_map(&tmB, reinterpret_cast<const __nv_bfloat16*>(x) + (size_t)xb * x_ps, 4, dims, strides, box, rbb);
    if (rc != PCM_OK) return rc;
    tmB2 = tmB;
  }
  static size_t smem_set = 0;
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad3x3_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("wgrad3x3_tc: smem attribute (%zu B): %s", smem, cudaGetErrorString(e)); return PCM_ERR_CUDA; }
    smem_set = smem;
  }
  unsigned int* err = tc_error_counter();
  PCM_REQUIRE(err != nullptr, "wgrad3x3_tc: could not allocate the error counter");
  int nsplit = g_wg_sms / (mtiles * p.ngroups);
  if (nsplit < 1) nsplit = 1;
  if (nsplit > p.num_ktiles) nsplit = p.num_ktiles;
  dim3 grid(nsplit, mtiles, p.ngroups);
  pcm::launch(wgrad3x3_tc_kernel, grid, kWgThreads, smem, (cudaStream_t)s, tmA, tmB, tmB2, dw, err, p);
  return check_launch("wgrad3x3_tc");
}

extern "C" int pcm_wgrad3x3_tc(const void* dy, long long dy_ns, int dy_ps, int Co, int Co_real, const void* x,
                               long long x_ns, int x_ps, int Ci, int Ci_real, float* dw, long long sa, long long sb,
                               long long st, int N, int H, int W, pcm_stream_t s) {
  if (W + 2 <= 256) return wgrad_tc_impl(dy, dy_ns, dy_ps, Co, Co_real, x, x_ns, x_ps, Ci, Ci_real, dw, sa, sb, st, N, H, W, 3, s);
  // rows wider than one TMA box (config 5: 360 columns): equal column strips, one launch each, all accumulating into dw
  const int nstrips = (W + 253) / 254;
  const int ws = (W + nstrips - 1) / nstrips;
  for (int w0 = 0; w0 < W; w0 += ws) {
    const int rc = wgrad_tc_impl(dy, dy_ns, dy_ps, Co, Co_real, x, x_ns, x_ps, Ci, Ci_real, dw, sa, sb, st, N, H,
                                 (w0 + ws <= W ? ws : W - w0), 3, s, 0, W, w0);
    if (rc != PCM_OK) return rc;
  }
  return PCM_OK;
}

extern "C" int pcm_wgrad3x3_tc_grouped(const void* dy, long long dy_ns, int Co, const void* x, long long x_ns, int Ci,
                                       int Ci_real, float* dw, long long sa, long long sb, long long st, int N, int H,
                                       int W, int group, pcm_stream_t s) {
  return wgrad_tc_impl(dy, dy_ns, Co, Co, Co, x, x_ns, Ci, Ci, Ci_real, dw, sa, sb, st, N, H, W, 3, s, 0, 0, 0, group);
}

extern "C" int pcm_wgrad1x1_tc(const void* dy, long long dy_ns, int dy_ps, int Co, int Co_real, const void* x,
                               long long x_ns, int x_ps, int Ci, int Ci_real, float* dw, long long sa, long long sb,
                               int N, int H, int W, pcm_stream_t s) {
  return wgrad_tc_impl(dy, dy_ns, dy_ps, Co, Co_real, x, x_ns, x_ps, Ci, Ci_real, dw, sa, sb, 0, N, H, W, 1, s);
}

// ConvTranspose2d(k2,s2) weight gradient: dw[ca*sa + cb*sb + q*st] += sum_{n,h,w} a(n,h,w,ca) * b(n, 2h+kh, 2w+kw, cb),
// q = kh*2 + kw; a is the (H, W) input of the transposed conv, b the (2H, 2W) output gradient.
extern "C" int pcm_convT2x2_wgrad_tc(const void* a, long long a_ns, int a_ps, int Ca, int Ca_real, const void* b,
                                     long long b_ns, int b_ps, int Cb, int Cb_real, float* dw, long long sa,
                                     long long sb, long long st, int N, int H, int W, pcm_stream_t s) {
  return wgrad_tc_impl(a, a_ns, a_ps, Ca, Ca_real, b, b_ns, b_ps, Cb, Cb_real, dw, sa, sb, st, N, H, W, 1, s, 1);
}

// Weight gradient of the 3x3 / stride-2 / pad-1 convolution in the pixel-pair form of pcm_conv3x3s2_tc:
// dw[co*sa + pc*sb + q*st] += sum_{n,h,w} dy(n,h,w,co) * x(n, 2h+kh-1, 2(w+dw)+pw, c), q = kh*2 + (dw+1), pc = pw*Cs + c.
// dy: (H, W) output gradient; x: the (2H, 2W) input with Cs dense channels.
extern "C" int pcm_wgrad3x3s2_tc(const void* dy, long long dy_ns, int dy_ps, int Co, const void* x, long long x_ns, int Cs,
                                 float* dw, long long sa, long long sb, long long st, int N, int H, int W, pcm_stream_t s) {
  return wgrad_tc_impl(dy, dy_ns, dy_ps, Co, Co, x, x_ns, Cs, 2 * Cs, 2 * Cs, dw, sa, sb, st, N, H, W, 1, s, 2);
}
