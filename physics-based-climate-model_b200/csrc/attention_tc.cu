// Multi-head self-attention core on the tensor cores (tcgen05, sm_100a) for the CNNTransformer geometry:
// L <= 224 tokens (12 x 18 = 216), head dim 32 (reference src/cnn_transformer.py:25-31, nn.MultiheadAttention).
//
// One CTA per (batch, head).  The whole problem of a head lives on chip:
//   S_t = Q_t K^T          two 128-row query tiles, N = Lk (keys padded to 16), K = 32      -> TMEM (fp32)
//   P   = softmax(scale S) one thread per query row (two warpgroups), bf16 -> shared memory in the canonical
//                          K-major 128B-swizzled layout (so it is directly the A operand of the next MMA)
//   O_t = P_t V            M = 128, N = 32, K = 256 (padded keys carry P = 0); V is used exactly as stored
//                          ([key][d] rows == MN-major B operand)                              -> TMEM
// Q, K, V tiles are TMA boxes of the packed in_proj output qkv [B*L][3E]; rows past the end of this batch element
// belong to the next one (finite values) or are zero-filled, and are masked (P = 0 / rows not stored).
// Dropout on the probabilities uses the same counter-based mask as the SIMT kernels (transformer.cu).
#include "tc_common.cuh"

namespace pcm {

using namespace tc;

constexpr int kAttThreads = 320;     // warp 0: TMA, warp 1: MMA issuer, warps 2-5 / 6-9: softmax + epilogue of tile 0 / 1
constexpr int kAttD = 32;

__device__ __forceinline__ float att_hash_uniform(unsigned long long seed, unsigned long long idx) {
  return dropout_uniform(seed, idx);      // common.cuh: one definition for every mask-drawing kernel
}

struct AttParams {
  int L, Lk, nh, E;
  float scale, drop_p;
  unsigned long long seed;
  const unsigned long long* epoch;      // library-owned dropout epoch cell (common.cuh)
};

// 16-byte store of 8 bf16 into a K-major 128B-swizzled tile: row r, 16-byte unit u (0..7) of 64-element chunk
__device__ __forceinline__ void st_sw128(uint8_t* chunk_base, int r, int u, const float v[8]) {
  uint4 w;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&w);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(chunk_base + r * 128 + ((u ^ (r & 7)) << 4)) = w;
}

__global__ void __launch_bounds__(kAttThreads, 1)
mha_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, __nv_bfloat16* __restrict__ out, float* __restrict__ lse,
                  unsigned int* __restrict__ err, const AttParams p) {
  pdl_launch_dependents();          // the next kernel may start launching; it waits for us in its own pdl_wait()
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                        // 2 x [128][64 B]
  uint8_t* sK = sQ + 2 * 8192;               // [256][64 B]
  uint8_t* sV = sK + 16384;                  // [256][64 B]
  uint8_t* sP = sV + 16384;                  // 2 x 4 chunks x [128][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * 65536);
  uint64_t* bar_load = bars;                 // [1]
  uint64_t* bar_S = bars + 1;                // [2]
  uint64_t* bar_P = bars + 3;                // [2]
  uint64_t* bar_O = bars + 5;                // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.x, b = bh / p.nh, h = bh - b * p.nh;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(bar_load, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&bar_S[i], 1); mbar_init(&bar_P[i], 128); mbar_init(&bar_O[i], 1); }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();                       // predecessor complete: global memory may be touched from here on
  const unsigned long long seed_eff = p.drop_p > 0.f ? mix_epoch(p.seed, p.epoch) : p.seed;
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t colO = 2 * p.Lk;            // S_t at columns t*Lk, O_t at 2*Lk + 32*t   (2*Lk + 64 <= 512)

  if (warp == 0) {
    if (elect_one()) {
      const int row0 = b * p.L;
      mbar_expect_tx(bar_load, (uint32_t)(2 * 128 + p.Lk + 256) * 64);
      tma_load_3d(sQ, &tmQ, bar_load, h * kAttD, row0, 0);
      tma_load_3d(sQ + 8192, &tmQ, bar_load, h * kAttD, row0 + 128, 0);
      tma_load_3d(sK, &tmK, bar_load, p.E + h * kAttD, row0, 0);
      tma_load_3d(sV, &tmV, bar_load, 2 * p.E + h * kAttD, row0, 0);
    }
  } else if (warp == 1) {
    if (elect_one()) {
      bool ok = mbar_wait(bar_load, 0, err);
      tc_fence_after();
      const uint32_t idesc_s = make_idesc_bf16(128, p.Lk, 0, 0);
      const uint32_t idesc_o = make_idesc_bf16(128, kAttD, 0, 1);          // B (= V) MN-major
      const uint64_t dK = make_smem_desc(smem_u32(sK), 16, 8 * 64, 4);
      const uint64_t dV = make_smem_desc(smem_u32(sV), 16384, 8 * 64, 4);
      for (int t = 0; t < 2 && ok; ++t) {
        const uint64_t dQ = make_smem_desc(smem_u32(sQ + t * 8192), 16, 8 * 64, 4);
#pragma unroll
        for (int k = 0; k < 2; ++k) umma_bf16(tmem_base + t * p.Lk, dQ + (uint64_t)(2 * k), dK + (uint64_t)(2 * k), idesc_s, k != 0);
        umma_commit(&bar_S[t]);
      }
      for (int t = 0; t < 2 && ok; ++t) {
        ok = mbar_wait(&bar_P[t], 0, err);
        if (!ok) break;
        tc_fence_after();
        const uint64_t dP = make_smem_desc(smem_u32(sP + t * 65536), 16, 8 * 128, 2);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_bf16(tmem_base + colO + 32 * t, dP + (uint64_t)(c * 1024 + 2 * ks), dV + (uint64_t)((c * 64 + ks * 16) * 4),
                      idesc_o, (c | ks) != 0);
        }
        umma_commit(&bar_O[t]);
      }
    }
  } else {
    // ===================== softmax + epilogue: warpgroup t owns query tile t, one row per thread =====================
    const int t = warp >= 6 ? 1 : 0;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const int i = t * 128 + r;                                   // query index inside this batch element
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    uint8_t* sPt = sP + t * 65536;
    bool ok = mbar_wait(&bar_S[t], 0, err);
    ok = __all_sync(0xffffffffu, ok);
    if (ok) {
      tc_fence_after();
      float mx = -INFINITY;
      for (int j0 = 0; j0 < p.Lk; j0 += 16) {
        float s[16];
        tmem_ld16(taddr + t * p.Lk + j0, s);
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (j0 + j < p.L) mx = fmaxf(mx, s[j]);
      }
      float sum = 0.f;
      const float keep_sc = 1.f / (1.f - p.drop_p);
      const bool row_ok = i < p.L;
      for (int j0 = 0; j0 < 256; j0 += 16) {
        float pv[16];
        if (j0 < p.Lk) {
          float s[16];
          tmem_ld16(taddr + t * p.Lk + j0, s);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float e = 0.f;
            if (j0 + j < p.L) {
              e = __expf((s[j] - mx) * p.scale);
              sum += e;
              if (p.drop_p > 0.f)
                e = att_hash_uniform(seed_eff, ((unsigned long long)bh * p.L + i) * p.L + j0 + j) >= p.drop_p ? e * keep_sc : 0.f;
            }
            pv[j] = row_ok ? e : 0.f;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) pv[j] = 0.f;
        }
        uint8_t* chunk = sPt + (j0 >> 6) * 16384;
        const int u = (j0 & 63) >> 3;
        st_sw128(chunk, r, u, pv);
        st_sw128(chunk, r, u + 1, pv + 8);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(&bar_P[t]);
      ok = mbar_wait(&bar_O[t], 0, err);
      ok = __all_sync(0xffffffffu, ok);
      if (ok) {
        tc_fence_after();
        float o[32];
        tmem_ld16(taddr + colO + 32 * t, o);
        tmem_ld16(taddr + colO + 32 * t + 16, o + 16);
        if (row_ok) {
          const float inv = 1.f / sum;
#pragma unroll
          for (int j = 0; j < 32; ++j) o[j] *= inv;
          __nv_bfloat16* op = out + ((long long)b * p.L + i) * p.E + h * kAttD;
          store8(op, o); store8(op + 8, o + 8); store8(op + 16, o + 16); store8(op + 24, o + 24);
          lse[(long long)bh * p.L + i] = mx * p.scale + __logf(sum);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Backward.  One CTA per (batch, head); iteration (kt, t) = (key tile of 128, query tile of 128):
//   S  = Q_t K_kt^T , dP = dO_t V_kt^T                                 (TMEM, N = 128 each)
//   P~ = dropout(exp(scale S - lse)) , dS' = scale * P o (dP~ - D)      two warpgroups, 64 keys each, -> smem (bf16)
//   dV_kt += P~^T dO_t , dK_kt += dS'^T Q_t , dQ_t += dS' K_kt          (TMEM accumulators)
// The [query][key] images of P~ and dS' in shared memory are K-major A operands (for dQ) and, read the other way,
// MN-major A operands (for dV, dK) — the same bytes; Q, dO, K as stored are MN-major B operands.
// D_i = dO_i . O_i and lse_i sit in shared memory.  TMEM: S 0..127 | dP 128..255 | dQ_0 256 | dQ_1 288 | dK 320 | dV 352.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kAttThreads, 1)
mha_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                  const __grid_constant__ CUtensorMap tmdO, const __nv_bfloat16* __restrict__ out,
                  const __nv_bfloat16* __restrict__ dout, const float* __restrict__ lse, __nv_bfloat16* __restrict__ dqkv,
                  unsigned int* __restrict__ err, const AttParams p) {
  pdl_launch_dependents();          // the next kernel may start launching; it waits for us in its own pdl_wait()
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                        // 2 x [128][64 B]
  uint8_t* sdO = sQ + 16384;                 // 2 x [128][64 B]
  uint8_t* sK = sdO + 16384;                 // [256][64 B]
  uint8_t* sV = sK + 16384;                  // [256][64 B]
  uint8_t* sP = sV + 16384;                  // 2 chunks x [128][128 B]
  uint8_t* sdS = sP + 32768;                 // 2 chunks x [128][128 B]
  float* sD = reinterpret_cast<float*>(sdS + 32768);      // [256]
  float* sLse = sD + 256;                                  // [256]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sLse + 256);
  uint64_t* bar_load = bars;                 // [1]
  uint64_t* bar_sdp = bars + 1;              // [4]  MMA -> warpgroups: S, dP of iteration it are in TMEM
  uint64_t* bar_pds = bars + 5;              // [4]  warpgroups -> MMA: P~, dS' of iteration it are in smem (256 arrivals)
  uint64_t* bar_acc = bars + 9;              // [4]  MMA -> warpgroups: the three products of iteration it are complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.x, b = bh / p.nh, h = bh - b * p.nh;
  const int nkt = p.Lk > 128 ? 2 : 1, nqt = p.L > 128 ? 2 : 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    tma_prefetch_desc(&tmdO);
    mbar_init(bar_load, 1);
    for (int i = 0; i < 4; ++i) { mbar_init(&bar_sdp[i], 1); mbar_init(&bar_pds[i], 256); mbar_init(&bar_acc[i], 1); }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();                       // predecessor complete: global memory may be touched from here on
  const unsigned long long seed_eff = p.drop_p > 0.f ? mix_epoch(p.seed, p.epoch) : p.seed;
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t cS = 0, cdP = 128, cdQ = 256, cdK = 320, cdV = 352;

  if (warp == 0) {
    if (elect_one()) {
      const int row0 = b * p.L;
      mbar_expect_tx(bar_load, (uint32_t)(4 * 128 + 2 * 256) * 64);
      tma_load_3d(sQ, &tmQ, bar_load, h * kAttD, row0, 0);
      tma_load_3d(sQ + 8192, &tmQ, bar_load, h * kAttD, row0 + 128, 0);
      tma_load_3d(sdO, &tmdO, bar_load, h * kAttD, row0, 0);
      tma_load_3d(sdO + 8192, &tmdO, bar_load, h * kAttD, row0 + 128, 0);
      tma_load_3d(sK, &tmKV, bar_load, p.E + h * kAttD, row0, 0);
      tma_load_3d(sV, &tmKV, bar_load, 2 * p.E + h * kAttD, row0, 0);
    }
  } else if (warp == 1) {
    if (elect_one()) {
      bool ok = mbar_wait(bar_load, 0, err);
      tc_fence_after();
      const uint32_t id_sdp = make_idesc_bf16(128, 128, 0, 0);    // A K-major (Q / dO), B K-major (K / V rows)
      const uint32_t id_kv = make_idesc_bf16(128, kAttD, 1, 1);    // A MN-major (P~^T / dS'^T), B MN-major (dO / Q)
      const uint32_t id_q = make_idesc_bf16(128, kAttD, 0, 1);     // A K-major (dS'), B MN-major (K rows)
      const uint64_t aP_mn = make_smem_desc(smem_u32(sP), 16384, 8 * 128, 2);
      const uint64_t adS_mn = make_smem_desc(smem_u32(sdS), 16384, 8 * 128, 2);
      const uint64_t adS_k = make_smem_desc(smem_u32(sdS), 16, 8 * 128, 2);
      int it = 0;
      for (int kt = 0; kt < nkt && ok; ++kt) {
        const uint64_t bK = make_smem_desc(smem_u32(sK + kt * 8192), 16, 8 * 64, 4);          // K-major B (rows = keys)
        const uint64_t bV = make_smem_desc(smem_u32(sV + kt * 8192), 16, 8 * 64, 4);
        const uint64_t bK_mn = make_smem_desc(smem_u32(sK + kt * 8192), 16384, 8 * 64, 4);    // MN-major B (K = key rows)
        for (int t = 0; t < nqt && ok; ++t, ++it) {
          const uint64_t aQ = make_smem_desc(smem_u32(sQ + t * 8192), 16, 8 * 64, 4);
          const uint64_t adO = make_smem_desc(smem_u32(sdO + t * 8192), 16, 8 * 64, 4);
          const uint64_t bQ_mn = make_smem_desc(smem_u32(sQ + t * 8192), 16384, 8 * 64, 4);
          const uint64_t bdO_mn = make_smem_desc(smem_u32(sdO + t * 8192), 16384, 8 * 64, 4);
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            umma_bf16(tmem_base + cS, aQ + (uint64_t)(2 * k), bK + (uint64_t)(2 * k), id_sdp, k != 0);
            umma_bf16(tmem_base + cdP, adO + (uint64_t)(2 * k), bV + (uint64_t)(2 * k), id_sdp, k != 0);
          }
          umma_commit(&bar_sdp[it]);
          ok = mbar_wait(&bar_pds[it], 0, err);
          if (!ok) break;
          tc_fence_after();
          // dV_kt += P~^T dO_t ; dK_kt += dS'^T Q_t : K = the 128 query rows of the tile, 16 per step
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            umma_bf16(tmem_base + cdV, aP_mn + (uint64_t)(ks * 128), bdO_mn + (uint64_t)(ks * 64), id_kv, (t | ks) != 0);
            umma_bf16(tmem_base + cdK, adS_mn + (uint64_t)(ks * 128), bQ_mn + (uint64_t)(ks * 64), id_kv, (t | ks) != 0);
          }
          // dQ_t += dS' K_kt : K = the 128 keys of the tile = 2 chunks x 4 steps
#pragma unroll
          for (int c = 0; c < 2; ++c) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_bf16(tmem_base + cdQ + 32 * t, adS_k + (uint64_t)(c * 1024 + 2 * ks), bK_mn + (uint64_t)((c * 64 + ks * 16) * 4),
                        id_q, (kt | c | ks) != 0);
          }
          umma_commit(&bar_acc[it]);
        }
      }
    }
  } else {
    // ===================== two warpgroups: 64 keys of the tile each, one query row per thread =====================
    const int wg = warp >= 6 ? 1 : 0;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    {   // D_i and lse_i of query wg*128 + r
      const int i = wg * 128 + r;
      float d = 0.f, l = 0.f;
      if (i < p.L) {
        const __nv_bfloat16* op = out + ((long long)b * p.L + i) * p.E + h * kAttD;
        const __nv_bfloat16* gp = dout + ((long long)b * p.L + i) * p.E + h * kAttD;
#pragma unroll
        for (int c = 0; c < kAttD; c += 8) {
          float o[8], g[8];
          load8(op + c, o); load8(gp + c, g);
#pragma unroll
          for (int k = 0; k < 8; ++k) d = fmaf(o[k], g[k], d);
        }
        l = lse[(long long)bh * p.L + i];
      }
      sD[i] = d; sLse[i] = l;
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float keep_sc = 1.f / (1.f - p.drop_p);
    bool ok = true;
    int it = 0;
    for (int kt = 0; kt < nkt && ok; ++kt) {
      for (int t = 0; t < nqt && ok; ++t, ++it) {
        const int i = t * 128 + r;
        const bool row_ok = i < p.L;
        const float Di = sD[i], li = sLse[i];
        ok = mbar_wait(&bar_sdp[it], 0, err);
        if (ok && it > 0) ok = mbar_wait(&bar_acc[it - 1], 0, err);     // sP / sdS of the previous iteration consumed
        ok = __all_sync(0xffffffffu, ok);
        if (!ok) break;
        tc_fence_after();
#pragma unroll 1
        for (int jj = 0; jj < 4; ++jj) {
          float s[16], dp[16], pt[16], ds[16];
          tmem_ld16(taddr + cS + wg * 64 + jj * 16, s);
          tmem_ld16(taddr + cdP + wg * 64 + jj * 16, dp);
          const int j0 = kt * 128 + wg * 64 + jj * 16;
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int j = j0 + e;
            float pv = 0.f, ks = 1.f;
            if (row_ok && j < p.L) {
              pv = __expf(s[e] * p.scale - li);
              if (p.drop_p > 0.f)
                ks = att_hash_uniform(seed_eff, ((unsigned long long)bh * p.L + i) * p.L + j) >= p.drop_p ? keep_sc : 0.f;
            }
            pt[e] = pv * ks;
            ds[e] = pv * (dp[e] * ks - Di) * p.scale;
          }
          st_sw128(sP + wg * 16384, r, jj * 2, pt);
          st_sw128(sP + wg * 16384, r, jj * 2 + 1, pt + 8);
          st_sw128(sdS + wg * 16384, r, jj * 2, ds);
          st_sw128(sdS + wg * 16384, r, jj * 2 + 1, ds + 8);
        }
        tc_fence_before();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(&bar_pds[it]);
      }
      if (!ok) break;
      // dK_kt (warpgroup 0) / dV_kt (warpgroup 1): TMEM lane = key row of the tile
      ok = mbar_wait(&bar_acc[it - 1], 0, err);
      ok = __all_sync(0xffffffffu, ok);
      if (!ok) break;
      tc_fence_after();
      {
        float v[32];
        tmem_ld16(taddr + (wg ? cdV : cdK), v);
        tmem_ld16(taddr + (wg ? cdV : cdK) + 16, v + 16);
        const int j = kt * 128 + r;
        if (j < p.L) {
          __nv_bfloat16* dp = dqkv + ((long long)b * p.L + j) * 3 * p.E + (wg ? 2 : 1) * p.E + h * kAttD;
          store8(dp, v); store8(dp + 8, v + 8); store8(dp + 16, v + 16); store8(dp + 24, v + 24);
        }
      }
      tc_fence_before();
    }
    if (ok && wg < nqt) {      // dQ_t: warpgroup t (all products complete: bar_acc of the last iteration was awaited above)
      tc_fence_after();
      float v[32];
      tmem_ld16(taddr + cdQ + 32 * wg, v);
      tmem_ld16(taddr + cdQ + 32 * wg + 16, v + 16);
      const int i = wg * 128 + r;
      if (i < p.L) {
        __nv_bfloat16* dp = dqkv + ((long long)b * p.L + i) * 3 * p.E + h * kAttD;
        store8(dp, v); store8(dp + 8, v + 8); store8(dp + 16, v + 16); store8(dp + 24, v + 24);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace pcm

using namespace pcm;

static int att_maps(const void* base, int B, int L, int E3, CUtensorMap* m128, CUtensorMap* mLk, int Lk, CUtensorMap* m256) {
  // qkv (or a gradient of the same shape) as {3E columns, B*L rows, 1}; boxes of 32 columns (64 B, 64B swizzle)
  uint64_t dims[3] = {(uint64_t)E3, (uint64_t)B * L, 1};
  uint64_t strides[2] = {(uint64_t)E3 * 2, (uint64_t)E3 * 2 * (uint64_t)B * L};
  int rc = PCM_OK;
  if (m128) { uint32_t box[3] = {32, 128, 1}; rc = make_tensor_map(m128, base, 3, dims, strides, box, 64); if (rc) return rc; }
  if (mLk) { uint32_t box[3] = {32, (uint32_t)Lk, 1}; rc = make_tensor_map(mLk, base, 3, dims, strides, box, 64); if (rc) return rc; }
  if (m256) { uint32_t box[3] = {32, 256, 1}; rc = make_tensor_map(m256, base, 3, dims, strides, box, 64); if (rc) return rc; }
  return PCM_OK;
}

extern "C" int pcm_mha_fwd_tc(const void* qkv, void* out, float* lse, int B, int L, int nh, float scale, float drop_p,
                              long long seed, pcm_stream_t s) {
  PCM_REQUIRE(L >= 1 && L <= 224 && nh >= 1 && drop_p >= 0.f && drop_p < 1.f, "mha_fwd_tc: needs 1 <= L <= 224 (got %d)", L);
  PCM_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
              "mha_fwd_tc: pointers must be 16-byte aligned");
  if (B == 0) return PCM_OK;
  AttParams p;
  p.L = L; p.Lk = (L + 15) / 16 * 16; p.nh = nh; p.E = nh * kAttD;
  p.scale = scale; p.drop_p = drop_p; p.seed = (unsigned long long)seed; p.epoch = dropout_epoch_cell();
  PCM_REQUIRE(p.epoch != nullptr, "mha_tc: could not allocate the dropout epoch cell");
  CUtensorMap tmQ, tmK, tmV;
  int rc = att_maps(qkv, B, L, 3 * p.E, &tmQ, &tmK, p.Lk, &tmV);
  if (rc != PCM_OK) return rc;
  const size_t smem = 1024 + 2 * 8192 + 16384 + 16384 + 2 * 65536 + 64;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(mha_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("mha_fwd_tc: smem attribute (%zu B): %s", smem, cudaGetErrorString(e)); return PCM_ERR_CUDA; }
    attr_set = true;
  }
  unsigned int* err = tc_error_counter();
  PCM_REQUIRE(err != nullptr, "mha_fwd_tc: could not allocate the error counter");
  pcm::launch(mha_fwd_tc_kernel, B * nh, kAttThreads, smem, (cudaStream_t)s, tmQ, tmK, tmV, reinterpret_cast<__nv_bfloat16*>(out), lse,
                                                                   err, p);
  return check_launch("mha_fwd_tc");
}

extern "C" int pcm_mha_bwd_tc(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int B, int L,
                              int nh, float scale, float drop_p, long long seed, pcm_stream_t s) {
  PCM_REQUIRE(L >= 1 && L <= 224 && nh >= 1 && drop_p >= 0.f && drop_p < 1.f, "mha_bwd_tc: needs 1 <= L <= 224 (got %d)", L);
  PCM_REQUIRE((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(dout) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(dqkv) & 15) == 0,
              "mha_bwd_tc: pointers must be 16-byte aligned");
  if (B == 0) return PCM_OK;
  AttParams p;
  p.L = L; p.Lk = (L + 15) / 16 * 16; p.nh = nh; p.E = nh * kAttD;
  p.scale = scale; p.drop_p = drop_p; p.seed = (unsigned long long)seed; p.epoch = dropout_epoch_cell();
  PCM_REQUIRE(p.epoch != nullptr, "mha_tc: could not allocate the dropout epoch cell");
  CUtensorMap tmQ, tmKV, tmdO;
  int rc = att_maps(qkv, B, L, 3 * p.E, &tmQ, nullptr, 0, &tmKV);
  if (rc != PCM_OK) return rc;
  rc = att_maps(dout, B, L, p.E, &tmdO, nullptr, 0, nullptr);
  if (rc != PCM_OK) return rc;
  const size_t smem = 1024 + 4 * 16384 + 2 * 32768 + 2 * 256 * 4 + 13 * 8 + 64;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(mha_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("mha_bwd_tc: smem attribute (%zu B): %s", smem, cudaGetErrorString(e)); return PCM_ERR_CUDA; }
    attr_set = true;
  }
  unsigned int* err = tc_error_counter();
  PCM_REQUIRE(err != nullptr, "mha_bwd_tc: could not allocate the error counter");
  pcm::launch(mha_bwd_tc_kernel, B * nh, kAttThreads, smem, (cudaStream_t)s, 
      tmQ, tmKV, tmdO, reinterpret_cast<const __nv_bfloat16*>(out), reinterpret_cast<const __nv_bfloat16*>(dout), lse,
      reinterpret_cast<__nv_bfloat16*>(dqkv), err, p);
  return check_launch("mha_bwd_tc");
}
