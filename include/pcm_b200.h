/* pcm_b200 — C ABI of the B200-native hot path of the Physics-Based-Climate-Model emulator.
 *
 * The reference (/root/reference) is pure Python on stock torch.nn modules and defines NO native
 * interface; these entry points are what its modules' forward/backward bodies
 * bind instead of the ATen calls they make today.  Each declaration cites the reference lines it
 * replaces.  Conventions:
 *   - every pointer is a DEVICE pointer unless named host_*; `stream` is a cudaStream_t;
 *   - activations are NHWC ("channels last") with the channel count padded to a multiple of 8,
 *     stored as `dtype` (PCM_F32 or PCM_BF16); statistics, gate maps, cell state, weight
 *     gradients and optimizer state are always fp32;
 *   - an activation view is (ptr, nstride, pstride): element (n, h, w, c) lives at
 *     ptr[n*nstride + (h*W + w)*pstride + c]  (lets kernels read/write slices of concat buffers
 *     and time-strided image sequences without copies);
 *   - functions return PCM_OK or an error code and never abort; pcm_last_error() gives the text;
 *   - nothing allocates or synchronises: every call is CUDA-Graph capturable.
 *   - "accumulates" means += into a buffer the caller zeroed (weight-gradient buffers).
 */
#ifndef PCM_B200_H_
#define PCM_B200_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCM_OK 0
#define PCM_ERR_INVALID 1
#define PCM_ERR_CUDA 2
#define PCM_F32 0
#define PCM_BF16 1

typedef void* pcm_stream_t; /* cudaStream_t */

#if defined(PCM_BUILDING)
#define PCM_API __attribute__((visibility("default")))
#else
#define PCM_API
#endif

PCM_API const char* pcm_last_error(void);
PCM_API int pcm_version(void);

/* ---- layout staging ------------------------------------------------------------------------
 * Module boundary tensors are NCHW fp32 (reference: every nn.Module in src/*.py).  T > 1 additionally
 * re-orders a (B, T) window batch: NCHW image b*T + t  <->  NHWC image t*B + b ("t-major", so that every
 * time step of the ConvLSTM is a contiguous block of B images). */
PCM_API int pcm_nchw_to_nhwc(const float* x, void* y, int N, int C, int H, int W, int Cp, int T, int dtype,
                             pcm_stream_t s);
PCM_API int pcm_nhwc_to_nchw(const void* x, float* y, int N, int C, int H, int W, int Cp, int T, int dtype,
                             pcm_stream_t s);
/* main_final.py:186-216 (seasonal channels): x5 (N,5,H,W) + month index (N) -> NHWC with
 * ch5 = sin(2 pi m/12), ch6 = cos(2 pi m/12), ch7.. = 0. */
PCM_API int pcm_season_embed_stage(const float* x5, const int* month, void* y, int N, int H, int W, int Cp, int T,
                                   int dtype, pcm_stream_t s);

/* Sliding windows from a device-resident series (SequenceDataset.__getitem__, main_final.py:97-154):
 * NHWC image n <- frame frames[n] of series [Ttot][C][H][W] (fp32 NCHW); frames[n] < 0 = the zero left-pad of windows
 * that start before the record.  With frames[t*B + b] = idx[b] - T + 1 + t the result is the t-major staged batch the
 * models' forward_staged takes — a training step then needs B window indices from the host, not B*T frames.
 * Fused while staging (both optional, NULL = off):
 *   norm [C][4] fp64 = (kind, a, b, lambda): Normalizer.normalize(.., "input") of src/utils_final.py:45-128,
 *     x_n = (g(x) - b) * a, g = identity (kind 0: zscore a = 1/(std + 1e-8), b = mean; minimax a = 1/range, b = min),
 *     log1p (1), sqrt (2), x^lambda (3); kind < 0 passes the channel through;
 *   month [Ttot] int32 (0..11): channels C, C+1 = sin / cos(2 pi month / 12) (main_final.py:186-216). */
PCM_API int pcm_window_stage(const float* series, const int* frames, void* y, int N, int C, int H, int W, int Cp,
                             const double* norm, const int* month, int dtype, pcm_stream_t s);

/* ---- weight packing: out[t][o][i] = (o<O && i<I) ? w[o*so + i*si + t*st] : 0, stored as dtype */
PCM_API int pcm_pack_weight(const float* w, long long so, long long si, long long st, int O, int I, int taps, int Op,
                    int Ip, void* out, int dtype, pcm_stream_t s);

/* Pixel-group form of a 3x3 / stride-1 kernel, for pcm_conv3x3_tc_grouped: `group` = g adjacent pixels of an image row
 * are treated as ONE pixel with g times the channels, so the packed kernel is [9][g*Op][g*Ip] with
 * out[kh*3 + s][pa*Op + o][pb*Ip + i] = base[kh*3 + dx][o][i], dx = g*(s-1) + pb - pa + 1 (zero unless 0 <= dx <= 2),
 * base[t][o][i] = (o<O && i<I) ? w[o*so + i*si + t*st] : 0 exactly as pcm_pack_weight (forward and flipped data-gradient
 * packings alike). */
PCM_API int pcm_pack_weight_grouped(const float* w, long long so, long long si, long long st, int O, int I, int Op,
                                    int Ip, int group, void* out, int dtype, pcm_stream_t s);

/* every re-pack of a training step in ONE launch.  jobs: device array of records of 8 x int64:
 * {w ptr, out ptr, so, si, st, O | I<<32, taps | Op<<32, Ip | dtype<<32 | group<<40} (same meaning as pcm_pack_weight /
 * pcm_pack_weight_grouped; group 0 or 1 = plain; a grouped job has taps*Op*Ip*group^2 outputs);
 * work: device array of nwork (job, block) int32 pairs — block b of job j covers elements [1024 b, 1024 b + 1024)
 * of that job's taps*Op*Ip outputs, so that all jobs proceed in parallel. */
PCM_API int pcm_pack_weights_batched(const long long* jobs, const int* work, int nwork, pcm_stream_t s);

/* packed weight gradients -> parameter layout, all layers in ONE launch (and the packed buffers are re-zeroed):
 * dst[co*sa + ci*sb + tap*st] += packed[(tap*Co + co)*Cpad + ci].  jobs: records of 8 x int64:
 * {packed ptr, dst ptr, sa, sb, st, Co | Ci_real<<32, Cpad | taps<<32, 0}; work: nwork (job, co) int32 pairs — one
 * CTA per output channel of each job (taps*Cpad <= 4608).
 * pcm_wgrad3x3_tc reduces into such a buffer with 16-byte vector atomics when called with sb == 1. */
PCM_API int pcm_unpack_grads_batched(const long long* jobs, const int* work, int nwork, pcm_stream_t s);

/* ---- convolution family (nn.Conv2d / nn.ConvTranspose2d call sites: src/convlstm.py:9,13;
 * src/unet.py:36,38,63; src/cnn_transformer.py:10,12,36,38; src/models.py:47,50,57,90,108).
 * "gather" form: dst(n,hd,wd,dc) = sum_taps sum_sc src(n,hs,ws,sc) * wk[tap][dc][sc] (+bias)(relu)
 *   mode 0 (conv forward / convT data-grad): hs = hd*stride - pad + kh
 *   mode 1 (conv data-grad / convT forward): hs = (hd + pad - kh)/stride when divisible
 * wk is packed [KH*KW][Dc][Sc] in `dtype`; dst is `dtype`, or fp32 when dst_f32. */
PCM_API int pcm_conv_gather(const void* src, long long src_ns, int src_ps, int Hs, int Ws, int Sc,
                    void* dst, long long dst_ns, int dst_ps, int Hd, int Wd, int Dc,
                    const void* wk, const float* bias, int N, int KH, int KW, int stride, int pad, int mode,
                    int dst_f32, int accumulate, int relu, int dtype, pcm_stream_t s);
/* Tensor-core path (tcgen05 implicit GEMM, TMA-staged halo tiles; bf16 in, fp32 accumulate) for the
 * 3x3 / stride 1 / pad 1 convolutions: dst(n,h,w,co) = sum_{kh,kw,ci} src(n,h+kh-1,w+kw-1,ci)*wk[kh*3+kw][co][ci]
 * (+bias) (+= dst when accumulate; fp32 dst only).  wk is bf16 [9][Cout][Cin]; Cin in {16,32,64k},
 * Cout multiple of 16, <= 256.  Data gradients use the same entry with flipped/transposed weights. */
PCM_API int pcm_conv3x3_tc(const void* src, long long src_ns, int src_ps, int H, int W, int Cin, void* dst,
                           long long dst_ns, int dst_ps, int Cout, const void* wk, const float* bias, int N,
                           int dst_f32, int accumulate, pcm_stream_t s);
/* The same convolution in PIXEL-GROUP form for the thin layers (Cin, Cout = 16 / 32): TMA moves a box one pixel row
 * (Cin*2 bytes) at a time at a fixed cost per row, which — not the tensor pipe, not HBM — bounds a 16-channel layer
 * (32-byte rows).  Here `group` = g adjacent pixels of a row (W % g == 0, dense pixels: src_ps == Cin, dst_ps == Cout)
 * are one GEMM row of g*Cin channels and produce g*Cout outputs, through a kernel packed by pcm_pack_weight_grouped
 * (bf16 [9][g*Cout][g*Cin], two thirds of it useful): g times fewer, g times longer TMA rows, the same number of UMMAs
 * (each g times wider), identical results up to the fp32 summation order.  g*Cin, g*Cout <= 64. */
PCM_API int pcm_conv3x3_tc_grouped(const void* src, long long src_ns, int src_ps, int H, int W, int Cin, void* dst,
                                   long long dst_ns, int dst_ps, int Cout, const void* wk, int N, int dst_f32,
                                   int group, pcm_stream_t s);
/* Fused ConvLSTM step (src/convlstm.py:11-19) for t >= 1: gates = conv3x3(h_prev, Wh) [tcgen05, fp32 in TMEM]
 * + gx (= Wx.x_t + bias, fp32 [B*P][4Ch], precomputed for all T by one pcm_conv3x3_tc launch); the epilogue
 * applies sigmoid/tanh, updates c (fp32) and h (bf16) and saves the activated gates for backward — the
 * gate pre-activations of the recurrent half never reach HBM.  wh: bf16 [9][4Ch][Ch]; Ch in {16,32,64}. */
PCM_API int pcm_convlstm_step_tc(const void* h_prev, const void* wh, const float* gx, const float* c_prev,
                                 void* h_out, float* c_out, void* acts, int B, int H, int W, int Ch, pcm_stream_t s);
/* Persistent ConvLSTM recurrence (src/convlstm.py:27-35): all T steps in ONE launch, forward, and the whole
 * back-propagation through time in ONE launch.  A cluster of four CTAs owns two samples for all steps; CTA r keeps the
 * Wh rows of hidden channels [16r, 16r+16) resident in shared memory, h_{t-1} lives in shared memory as a zero-ringed
 * halo image (nine taps = nine row-shifted UMMA descriptors), the accumulator in TMEM, c / dc in registers; h_t (forward)
 * and the K-split partial sums of dh_{t-1} (backward) are exchanged through distributed shared memory, one cluster
 * barrier per step.  Ch = 64, bf16, (H-1)*(W+2) + W <= 64 (pcm_convlstm_seq_supported).
 *   gx     fp32 [T][B][H*W][4*Ch]  Wx.x + bias for every step (one batched pcm_conv3x3_tc launch)
 *   wh     bf16 [9][4*Ch][Ch]      recurrent half of the gate weight, forward packing
 *   h_all  bf16 [T][B][H*W][Ch], c_all fp32 [T][B][H*W][Ch], acts bf16 [T][B][H*W][4*Ch]   (outputs; gate order i,f,o,g)
 *   dh_ext bf16: gradient w.r.t. h of every step [T][B][H*W][Ch] (ext_all_steps = 1) or of the last step only [B][H*W][Ch]
 *   wht    bf16 [9][Ch][4*Ch]      data-gradient packing of the recurrent half (taps flipped)
 *   dgates bf16 [T][B][H*W][4*Ch]  (output) gradient w.r.t. the gate pre-activations of every step */
PCM_API int pcm_convlstm_seq_supported(int H, int W, int Ch);
PCM_API int pcm_convlstm_seq_fwd_tc(const float* gx, const void* wh, void* h_all, float* c_all, void* acts, int T, int B,
                                    int H, int W, int Ch, pcm_stream_t s);
PCM_API int pcm_convlstm_seq_bwd_tc(const void* dh_ext, int ext_all_steps, const void* acts, const float* c_all,
                                    const void* wht, void* dgates, int T, int B, int H, int W, int Ch, pcm_stream_t s);

/* Tensor-core weight gradient of the same convolution (GEMM over the pixel dimension, MN-major operands
 * straight from NHWC, all taps from ONE halo tile): dw[co*sa + ci*sb + tap*st] += sum_p dy(p,co)*x(p+tap,ci).
 * dy: Co in {16,32,64,128k}; x: Ci in {16,32,64,128,192,256}; only co < Co_real, ci < Ci_real are written.
 * fp32, accumulates (atomics). */
PCM_API int pcm_wgrad3x3_tc(const void* dy, long long dy_ns, int dy_ps, int Co, int Co_real, const void* x,
                            long long x_ns, int x_ps, int Ci, int Ci_real, float* dw, long long sa, long long sb,
                            long long st, int N, int H, int W, pcm_stream_t s);
/* The same weight gradient in PIXEL-GROUP form for the thin layers (dense pixels, Co and Ci in {16, 32}, group*C <= 64,
 * W % group == 0): both tensors are fetched as images of W/group pixels with group*C channels, i.e. `group` times fewer
 * and longer TMA box rows (the per-row cost of TMA, not HBM or the tensor pipe, bounds these layers).  An output pixel pa
 * of a group needs the input pixels p0-1 .. p0+group of its row — group+2 consecutive pixels of the pixel-linear tile — so
 * one UMMA per kernel row with N = (group+2)*Ci covers all three taps for every pa; accumulator lane = (pa, co),
 * column = (kh, input pixel, ci); the epilogue adds each block to tap dx = input - output + 1. */
PCM_API int pcm_wgrad3x3_tc_grouped(const void* dy, long long dy_ns, int Co, const void* x, long long x_ns, int Ci,
                                    int Ci_real, float* dw, long long sa, long long sb, long long st, int N, int H,
                                    int W, int group, pcm_stream_t s);
/* 1x1 convolution / linear layer on the same tcgen05 pipeline (one tap): dst(n,h,w,co) = sum_ci src(n,h,w,ci)*wk[co][ci]
 * (+bias) (relu) (+= dst, fp32 only).  Call sites: the ResidualBlock skip conv (src/models.py:56), and — with the token
 * matrix [M][E] presented as an image (N=1, H*W=M) — every nn.Linear of the transformer encoder layer
 * (src/cnn_transformer.py:25-31: in_proj, out_proj, linear1(+ReLU), linear2).  wk: bf16 [Cout][Cin]. */
PCM_API int pcm_conv1x1_tc(const void* src, long long src_ns, int src_ps, int H, int W, int Cin, void* dst,
                           long long dst_ns, int dst_ps, int Cout, const void* wk, const float* bias, int N,
                           int dst_f32, int accumulate, int relu, pcm_stream_t s);
/* nn.Linear -> ReLU -> nn.Dropout(p) (inner half of the transformer FFN, src/cnn_transformer.py:25-31) in one launch: the
 * mask pcm_dropout(seed) would draw on the dense destination is applied in the store epilogue (0 < p < 1). */
PCM_API int pcm_conv1x1_drop_tc(const void* src, long long src_ns, int src_ps, int H, int W, int Cin, void* dst,
                                long long dst_ns, int dst_ps, int Cout, const void* wk, const float* bias, int N, int relu,
                                float drop_p, long long seed, pcm_stream_t s);
/* weight gradient of the above: dw[co*sa + ci*sb] += sum_p dy(p,co)*x(p,ci)  (fp32, accumulates) */
PCM_API int pcm_wgrad1x1_tc(const void* dy, long long dy_ns, int dy_ps, int Co, int Co_real, const void* x,
                            long long x_ns, int x_ps, int Ci, int Ci_real, float* dw, long long sa, long long sb,
                            int N, int H, int W, pcm_stream_t s);
/* nn.ConvTranspose2d(kernel_size=2, stride=2) (src/unet.py:63, src/cnn_transformer.py:36,38) on the tensor cores:
 *   forward   dst(n, 2h+kh, 2w+kw, co) = sum_ci src(n,h,w,ci) * wk[kh*2+kw][co][ci] + bias[co] (relu): ONE GEMM
 *             [pixels x Cin] x [Cin x 4*Cout] with a pixel-shuffle epilogue; H, W = input grid; Cout multiple of 16, <= 64
 *   dgrad     dx(n,h,w,ci) = sum_{kh,kw,co} dy(n, 2h+kh, 2w+kw, co) * wk[kh*2+kw][ci][co]: four stride-2 TMA views of dy
 *   wgrad     dw[ca*sa + cb*sb + q*st] += sum_{n,h,w} a(n,h,w,ca) * b(n, 2h+kh, 2w+kw, cb), q = kh*2 + kw
 * dy / b / dst may be channel slices of a wider (concat) buffer: pass its pixel and image strides. */
PCM_API int pcm_convT2x2_tc(const void* src, long long src_ns, int src_ps, int H, int W, int Cin, void* dst,
                            long long dst_ns, int dst_ps, int Cout, const void* wk, const float* bias, int N,
                            int relu, pcm_stream_t s);
PCM_API int pcm_convT2x2_dgrad_tc(const void* dy, long long dy_ns, int dy_ps, int H, int W, int Cout, void* dx,
                                  long long dx_ns, int dx_ps, int Cin, const void* wk, int N, pcm_stream_t s);
PCM_API int pcm_convT2x2_wgrad_tc(const void* a, long long a_ns, int a_ps, int Ca, int Ca_real, const void* b,
                                  long long b_ns, int b_ps, int Cb, int Cb_real, float* dw, long long sa,
                                  long long sb, long long st, int N, int H, int W, pcm_stream_t s);
/* nn.Conv2d(kernel 3, stride 2, padding 1) (src/cnn_transformer.py:10,12) on the tensor cores, in "pixel pair" form: the
 * (2H, 2W) input with Cs dense channels is viewed as {(pw, c), w, ph, h, n}; the GEMM has 6 taps q = kh*2 + (dw+1) of
 * K = 2*Cs (the (dw = -1, pw = 0) half of the weights is zero).  H, W = OUTPUT grid.
 *   forward  wk: bf16 [6][Cout][2*Cs], wk[q][co][pw*Cs + c] = w[co][c][kh][2*(dw+1) + pw - 1]
 *   dgrad    dx(2h+ph, 2w+pw, c) = sum_{oh,ow in {0,1}} sum_co dy(h+oh, w+ow, co) * wk[oh*2+ow][(ph*2+pw)*Cq + c][co]
 *   wgrad    dw[co*sa + (pw*Cs + c)*sb + q*st] += sum dy(n,h,w,co) * x(n, 2h+kh-1, 2(w+dw)+pw, c) */
PCM_API int pcm_conv3x3s2_tc(const void* src, long long src_ns, int Cs, int H, int W, void* dst, long long dst_ns,
                             int dst_ps, int Cout, const void* wk, const float* bias, int N, int relu, pcm_stream_t s);
PCM_API int pcm_conv3x3s2_dgrad_tc(const void* dy, long long dy_ns, int dy_ps, int H, int W, int Cout, void* dx,
                                   long long dx_ns, int dx_ps, int Cq, const void* wk, int N, pcm_stream_t s);
PCM_API int pcm_wgrad3x3s2_tc(const void* dy, long long dy_ns, int dy_ps, int Co, const void* x, long long x_ns, int Cs,
                              float* dw, long long sa, long long sb, long long st, int N, int H, int W, pcm_stream_t s);
/* number of bounded-wait timeouts recorded by the tensor-core kernels since load (0 when healthy; syncs) */
PCM_API int pcm_tc_error_count(void);
/* weight gradient: dw[ac*sa + bc*sb + tap*st] += sum_{n,ha,wa} A(n,ha,wa,ac) * B(n,hb,wb,bc),
 * hb = ha*stride - pad + kh; only ac < Ca_real, bc < Cb_real are written.  fp32, accumulates. */
PCM_API int pcm_conv_wgrad(const void* A, long long a_ns, int a_ps, int Ha, int Wa, int Ca, int Ca_real,
                   const void* B, long long b_ns, int b_ps, int Hb, int Wb, int Cb, int Cb_real,
                   float* dw, long long sa, long long sb, long long st, int N, int KH, int KW, int stride, int pad,
                   int dtype, pcm_stream_t s);
/* per-channel sum over pixels: out[c] += sum_{n,p} x(n,p,c) (bias gradients), or, when per_image,
 * out[n][c] += sum_p x(n,p,c) (SE squeeze); c < C_real */
PCM_API int pcm_channel_sum(const void* x, long long ns, int ps, int N, int P, int C, int C_real, float* out,
                            int per_image, int dtype, pcm_stream_t s);

/* ---- ConvBlock tail: GroupNorm(8)+SiLU, SE, SpatialGate (src/unet.py:6-29, 35-49) ------------ */
/* stats[n][g][2] += (sum, sum of squares) of x over group g of image n (caller zeroes stats) */
PCM_API int pcm_gn_stats(const void* x, float* stats, int N, int P, int C, int G, int dtype, pcm_stream_t s);
/* y = silu(gamma*(x-mu)*rstd+beta); pool[n][c] += sum_p y (nullable) */
PCM_API int pcm_gn_silu_fwd(const void* x, const float* stats, const float* gamma, const float* beta, void* y, float* pool,
                    int N, int P, int C, int G, float eps, int dtype, pcm_stream_t s);
/* SE excitation + channel statistics of u = a*se: se[n][c] = sigmoid(W2 relu(W1 pool/P));
 * cmap[n][p] = (mean_c u, max_c u).  w1 == NULL means "no excitation" (se = 1). */
PCM_API int pcm_se_chanstat_fwd(const void* a, const float* pool, const float* w1, const float* w2, float* se, float* hid,
                        float* cmap, int N, int P, int C, int Cr, int dtype, pcm_stream_t s);
/* out = x*scale[n][c] + add[n][c] (scale/add nullable) — stand-alone SEBlock (src/unet.py:16-17) */
PCM_API int pcm_scale_channels(const void* x, const float* scale, const float* add, void* out, int N, int P, int C,
                               int dtype, pcm_stream_t s);
/* gate[n][p] = sigmoid(conv7x7(cmap)); out = a*se*gate */
PCM_API int pcm_spatial_gate_fwd(const void* a, const float* se, const float* cmap, const float* wsp, float* gate, void* out,
                         int N, int H, int W, int C, int dtype, pcm_stream_t s);
/* backward of out = a*se*gate:  dq[n][p] = (sum_c dout*a*se) * gate*(1-gate) */
PCM_API int pcm_spatial_gate_bwd_dq(const void* dout, const void* a, const float* se, const float* gate, float* dq, int N,
                            int P, int C, int dtype, pcm_stream_t s);
/* dwsp[k][dy][dx] += sum dq[p]*cmap_k[p+off]  (98 values) */
PCM_API int pcm_spatial_gate_bwd_dw(const float* dq, const float* cmap, float* dwsp, int N, int H, int W, pcm_stream_t s);
/* da = (dout*gate + dmean/C + dmax*tie)*se ; dse[n][c] += sum_p du*a */
PCM_API int pcm_spatial_gate_bwd_da(const void* dout, const void* a, const float* se, const float* gate, const float* cmap,
                            const float* dq, const float* wsp, void* da, float* dse, int N, int H, int W, int C,
                            int dtype, pcm_stream_t s);
/* SE backward: dpool[n][c] (already divided by P) ; dw1, dw2 accumulate */
PCM_API int pcm_se_bwd(const float* dse, const float* se, const float* hid, const float* pool, const float* w1,
               const float* w2, float* dpool, float* dw1, float* dw2, int N, int P, int C, int Cr, pcm_stream_t s);
/* GroupNorm+SiLU backward, pass 1: with da_total = da + dpool[n][c] (dpool nullable):
 * gsum[n][g][2] += (sum dxhat, sum dxhat*xhat); dgamma[c] += sum dz*xhat; dbeta[c] += sum dz */
PCM_API int pcm_gn_silu_bwd_reduce(const void* da, const float* dpool, const void* x, const float* stats,
                           const float* gamma, const float* beta, float* gsum, float* dgamma, float* dbeta, int N,
                           int P, int C, int G, float eps, int dtype, pcm_stream_t s);
/* pass 2: dx = rstd*(dxhat - mean(dxhat) - xhat*mean(dxhat*xhat)) */
PCM_API int pcm_gn_silu_bwd_apply(const void* da, const float* dpool, const void* x, const float* stats,
                          const float* gamma, const float* beta, const float* gsum, void* dx, int N, int P, int C,
                          int G, float eps, int dtype, pcm_stream_t s);

/* ---- per-image fused ConvBlock tails (csrc/convblock_fused.cu): one CTA per image, the image resident in shared
 * memory.  GroupNorm has 8 groups (nn.GroupNorm(8, c), src/unet.py:37,39); stats[n][8][2] = (sum, sum of squares).
 * pcm_convblock_fused_supported: 1 when an H x W x C image of `dtype` (plus the gate maps) fits one SM. */
PCM_API int pcm_convblock_fused_supported(int H, int W, int C, int Cr, int dtype);
/* Whole ConvBlock forward (src/unet.py:35-49: conv3x3 -> GN(8) -> SiLU -> conv3x3 -> GN(8) -> SiLU -> SE -> spatial gate)
 * as ONE kernel, one CTA per image, for the thin layers (C = 16 / 32 / 64 output channels, Cin = 16 / 32 / 64 padded input
 * channels, bf16) whose halo image and weights fit one SM (pcm_convblock_fwd_tc_supported).  Both convolutions run on the
 * tensor cores over the shared-memory-resident image (nine row-shifted UMMA descriptors per tile), each conv output stays
 * in tensor memory as fp32 for the whole image, a1 = silu(GN(y1)) goes from TMEM straight into the next conv's operand
 * image.  x [N][H][W][Cin]; wk1 [9][C][Cin], wk2 [9][C][C] packed bf16 (pcm_pack_weight); outputs (all saved for the
 * backward pass, same contents as the 4-kernel path pcm_conv3x3_tc / pcm_gn_silu_img_fwd / pcm_conv3x3_tc /
 * pcm_convblock_tail_fwd): y1, a1, y2, out [N][H][W][C] bf16; stats1 / stats2 [N][8][2]; pool, se [N][C]; hid [N][Cr];
 * maps [N][3][H*W] fp32; ties [N][H*W]. */
PCM_API int pcm_convblock_fwd_tc_supported(int H, int W, int Cin, int C, int Cr);
PCM_API int pcm_convblock_fwd_tc(const void* x, const void* wk1, const void* wk2, const float* g1, const float* b1,
                                 const float* g2, const float* b2, const float* sw1, const float* sw2, const float* wsp,
                                 void* y1, void* a1, void* y2, float* stats1, float* stats2, float* pool, float* se,
                                 float* hid, float* maps, unsigned char* ties, void* out, int N, int H, int W, int Cin,
                                 int C, int Cr, float eps, pcm_stream_t s);

/* tail 1: y = silu(GroupNorm(x))  [statistics + normalise in one pass over shared memory] */
PCM_API int pcm_gn_silu_img_fwd(const void* x, const float* gamma, const float* beta, float* stats, void* y, int N,
                                int H, int W, int C, float eps, int dtype, pcm_stream_t s);
/* tail 2: out = a*se*gate, a = silu(GroupNorm(x)), se = sigmoid(W2 relu(W1 mean_p a)),
 * gate = sigmoid(conv7x7([mean_c a*se, max_c a*se])).  Saves stats, pool[n][C] (= sum_p a), se[n][C], hid[n][Cr]
 * and, when `maps` / `ties` are non-null (training; both or neither), maps[n] = mean[P] | max[P] | gate[P] (fp32)
 * and ties[n][P] = number of channels attaining the maximum (x.amax splits its gradient between them). */
PCM_API int pcm_convblock_tail_fwd(const void* x, const float* gamma, const float* beta, const float* w1,
                                   const float* w2, const float* wsp, float* stats, float* pool, float* se,
                                   float* hid, float* maps, unsigned char* ties, void* out, int N, int H, int W,
                                   int C, int Cr, float eps, int dtype, pcm_stream_t s);
/* backward of tail 1: da -> dx; dgamma, dbeta accumulate */
PCM_API int pcm_gn_silu_img_bwd(const void* da, const void* x, const float* stats, const float* gamma,
                                const float* beta, void* dx, float* dgamma, float* dbeta, int N, int H, int W, int C,
                                float eps, int dtype, pcm_stream_t s);
/* backward of tail 2: dout -> dx; dgamma, dbeta, dw1, dw2, dwsp accumulate.  `out`, `maps`, `ties` are what the
 * forward tail produced for the same x (the gate's pre-activation gradient is (1 - gate) * sum_c dout*out; a is
 * recomputed from x for the rest).  x must be 16-byte aligned. */
PCM_API int pcm_convblock_tail_bwd(const void* dout, const void* x, const void* out, const float* stats,
                                   const float* gamma, const float* beta, const float* w1, const float* w2,
                                   const float* wsp, const float* pool, const float* se, const float* hid,
                                   const float* maps, const unsigned char* ties, void* dx, float* dgamma,
                                   float* dbeta, float* dw1, float* dw2, float* dwsp, int N, int H, int W, int C,
                                   int Cr, float eps, int dtype, pcm_stream_t s);
/* The same with the gate-weight gradient taken off the critical path: dq_out [N][H*W] fp32 (non-NULL) receives the gate's
 * pre-activation gradient and dwsp is NOT touched; pcm_gate_wgrad(dq, maps, dwsp) then accumulates
 * dwsp[k][dy][dx] += sum_n sum_p dq[n][p] * map_k[n][p + (dy-3, dx-3)] for all images (maps as saved by the forward tail:
 * [N][3][H*W], mean | max | gate) — a parameter gradient, issued on a side stream by the caller. */
PCM_API int pcm_convblock_tail_bwd_dq(const void* dout, const void* x, const void* out, const float* stats,
                                      const float* gamma, const float* beta, const float* w1, const float* w2,
                                      const float* wsp, const float* pool, const float* se, const float* hid,
                                      const float* maps, const unsigned char* ties, void* dx, float* dgamma,
                                      float* dbeta, float* dw1, float* dw2, float* dwsp, float* dq_out, int N, int H,
                                      int W, int C, int Cr, float eps, int dtype, pcm_stream_t s);
PCM_API int pcm_gate_wgrad(const float* dq, const float* maps, float* dwsp, int N, int H, int W, pcm_stream_t s);
/* The same again with the per-pixel sum sdot[N][H*W] = sum_c dout*out supplied by the kernel that produced dout
 * (pcm_maxpool2_bwd_skip_dot: the backward of MaxPool2d + time-mean skip that follows the block in the encoder,
 * src/unet_convlstm_attention.py:21-24,91-93, holds dout and out in registers).  sdot non-NULL: dout / out are not streamed
 * for the gate gradient (out may be NULL); sdot NULL: identical to pcm_convblock_tail_bwd_dq.  dq_out may be NULL. */
PCM_API int pcm_convblock_tail_bwd_sdot(const void* dout, const void* x, const void* out, const float* stats,
                                        const float* gamma, const float* beta, const float* w1, const float* w2,
                                        const float* wsp, const float* pool, const float* se, const float* hid,
                                        const float* maps, const unsigned char* ties, void* dx, float* dgamma,
                                        float* dbeta, float* dw1, float* dw2, float* dwsp, float* dq_out,
                                        const float* sdot, int N, int H, int W, int C, int Cr, float eps, int dtype,
                                        pcm_stream_t s);

/* ---- BatchNorm2d, training mode (src/models.py:48,51,57,91,109; eps 1e-5, momentum 0.1) on NHWC rows
 * (R = N*H*W rows of C channels; C = 8 * a divisor of 256).  sums[c] = (sum x, sum x^2) in DOUBLE (2*C doubles),
 * zeroed by the caller: the cross-thread / cross-block accumulation order then cannot move mean or variance. */
PCM_API int pcm_bn_stats(const void* x, double* sums, long long R, int C, int dtype, pcm_stream_t s);
/* y = [relu]( gamma*(x-mean)*rstd + beta [+ res] )   (res nullable: the residual add of src/models.py:70-71) */
PCM_API int pcm_bn_apply_fwd(const void* x, const double* sums, const float* gamma, const float* beta, const void* res,
                             void* y, long long R, int C, float eps, int relu, int dtype, pcm_stream_t s);
/* running_mean/var <- (1-m)*old + m*(batch mean / unbiased batch var); num_batches_tracked (int64, nullable) += 1 */
PCM_API int pcm_bn_update_running(const double* sums, float* running_mean, float* running_var,
                                  long long* num_batches_tracked, long long R, int C, float momentum, pcm_stream_t s);
/* backward pass 1: with dz = dy * (y > 0) when y != NULL (ReLU mask), else dz = dy:
 * dsum[c] += (sum dz, sum dz*xhat)   (caller zeroes dsum) */
PCM_API int pcm_bn_bwd_reduce(const void* dy, const void* y, const void* x, const double* sums, float* dsum, long long R,
                              int C, float eps, int dtype, pcm_stream_t s);
/* pass 2: dx = gamma*rstd*(dz - mean(dz) - xhat*mean(dz*xhat)); dres = dz (nullable); dgamma/dbeta += dsum */
PCM_API int pcm_bn_bwd_apply(const void* dy, const void* y, const void* x, const double* sums, const float* gamma,
                             const float* dsum, void* dx, void* dres, float* dgamma, float* dbeta, long long R, int C,
                             float eps, int dtype, pcm_stream_t s);
/* elementwise helpers (n multiple of 8): out = a + b ; y[i] = x[i] + b[i % period] (pos_embedding add,
 * src/cnn_transformer.py:48) ; dx = dy * (y > 0) (nn.ReLU backward) */
PCM_API int pcm_add(const void* a, const void* b, void* out, long long n, int dtype, pcm_stream_t s);
PCM_API int pcm_add_bcast(const void* x, const float* b, void* y, long long n, long long period, int dtype,
                          pcm_stream_t s);
/* out[r] += sum_b x[b][r] (x: [B][R]; gradient of a parameter broadcast over the batch, e.g. pos_embedding) */
PCM_API int pcm_batch_sum(const void* x, float* out, int B, long long R, int dtype, pcm_stream_t s);
PCM_API int pcm_relu_bwd(const void* dy, const void* y, void* dx, long long n, int dtype, pcm_stream_t s);
/* nn.Dropout: y = x * keep(seed, i)/(1-p) with a counter-based mask (the same call on dy is the backward);
 * nn.Dropout2d (src/models.py:103): mask[n*C + c] in {0, 1/(1-p)}, applied with pcm_scale_channels */
/* Dropout epoch: every mask-drawing call (pcm_dropout, pcm_dropout_mask, pcm_mha_*) mixes ONE library-owned device
 * counter into its seed.  pcm_dropout_epoch_advance increments it with a one-thread kernel on `s`; captured in a CUDA
 * graph it makes every replay draw fresh masks (scalar seeds are frozen in the graph) while forward and backward of
 * one step still agree.  pcm_dropout_epoch reads it back (synchronous; tests). */
PCM_API int pcm_dropout_epoch_advance(pcm_stream_t s);
PCM_API long long pcm_dropout_epoch(void);
/* dx = y > 0 ? dy * scale : 0: backward of ReLU -> dropout from the saved output alone (scale = 1 / (1 - p)) */
PCM_API int pcm_relu_bwd_scaled(const void* dy, const void* y, void* dx, long long n, float scale, int dtype, pcm_stream_t s);
PCM_API int pcm_dropout(const void* x, void* y, long long n, float p, long long seed, int dtype, pcm_stream_t s);
PCM_API int pcm_dropout_mask(float* mask, long long n, float p, long long seed, pcm_stream_t s);

/* ---- transformer encoder layer (src/cnn_transformer.py:25-31; post-norm, batch_first) ---------------------
 * y = LayerNorm(a + dropout(b))*gamma + beta over the last dim E (eps inside sqrt); sum_out = a + dropout(b) (nullable)
 * and stat[m] = (mean, rstd) are saved for backward.  b nullable.  drop_p > 0 applies nn.Dropout's counter-based mask
 * (stream `seed`, element m*E + c — the mask pcm_dropout draws) to b on the fly: dropout -> residual add -> LayerNorm of
 * the post-norm encoder layer in one pass. */
PCM_API int pcm_add_layernorm_fwd(const void* a, const void* b, const float* gamma, const float* beta, void* sum_out,
                                  void* y, float* stat, int M, int E, float eps, float drop_p, long long seed, int dtype,
                                  pcm_stream_t s);
/* ds (gradient of a + dropout(b), i.e. of a); with drop_p > 0 also db = dropout(ds) (gradient of b, same mask); dgamma,
 * dbeta accumulate */
PCM_API int pcm_layernorm_bwd(const void* dy, const void* sum_in, const float* stat, const float* gamma, void* ds,
                              void* db, float* dgamma, float* dbeta, int M, int E, float drop_p, long long seed, int dtype,
                              pcm_stream_t s);
/* the same with the incoming gradient given as TWO addends (dy2 nullable): the gradients of the two consumers of a residual
 * fork are summed while loading instead of by a launch of their own */
PCM_API int pcm_layernorm_bwd2(const void* dy, const void* dy2, const void* sum_in, const float* stat, const float* gamma,
                               void* ds, void* db, float* dgamma, float* dbeta, int M, int E, float drop_p, long long seed,
                               int dtype, pcm_stream_t s);
/* multi-head self-attention core of nn.MultiheadAttention: qkv [B][L][3*nh*D] (q | k | v; head h = columns
 * h*D..), out [B][L][nh*D] = softmax(scale * q k^T) v per head, lse [B][nh][L] fp32 saved for backward.
 * drop_p: dropout on the attention probabilities (counter-based mask from `seed`).  D in {8, 16, 32, 64}. */
PCM_API int pcm_mha_fwd(const void* qkv, void* out, float* lse, int B, int L, int nh, int D, float scale, float drop_p,
                        long long seed, int dtype, pcm_stream_t s);
/* the same forward on the tensor cores (tcgen05; bf16, head dim 32, L <= 224): S = Q K^T and O = P V as UMMAs with
 * the scores in TMEM and the probabilities staged in shared memory as the next MMA's operand; same lse / dropout-mask
 * convention as pcm_mha_fwd, so either backward kernel pairs with it */
PCM_API int pcm_mha_fwd_tc(const void* qkv, void* out, float* lse, int B, int L, int nh, float scale, float drop_p,
                           long long seed, pcm_stream_t s);
/* ... and the backward on the tensor cores: per (key tile, query tile) S and dP as UMMAs, P~ / dS' through shared
 * memory (one image serves as K-major and as MN-major operand), dQ / dK / dV accumulated in TMEM */
PCM_API int pcm_mha_bwd_tc(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int B, int L,
                           int nh, float scale, float drop_p, long long seed, pcm_stream_t s);
PCM_API int pcm_mha_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv, int B, int L,
                        int nh, int D, float scale, float drop_p, long long seed, int dtype, pcm_stream_t s);

/* ---- pooling / skips (src/unet.py:54,57; src/unet_convlstm_attention.py:21,24,91-93) ----------- */
PCM_API int pcm_maxpool2_fwd(const void* x, void* y, int N, int H, int W, int C, int dtype, pcm_stream_t s);
/* dx(n,h,w,c) = [first max of its 2x2 window] * dy(n,h/2,w/2,c) + dskip(b(n), h, w, c)/T, with
 * b(n) = n % (N/T) for t-major image order, n / T otherwise
 * (dy nullable; dskip nullable, an activation view with channel offset applied by the caller) */
PCM_API int pcm_maxpool2_bwd_skip(const void* x, const void* dy, const void* dskip, long long dskip_ns, int dskip_ps,
                                  void* dx, int N, int H, int W, int C, int T, int t_major, int dtype, pcm_stream_t s);
/* the same, also writing sdot[n][h][w] = sum_c dx(n,h,w,c) * x(n,h,w,c) (fp32, dx as the storage type rounds it; sdot
 * nullable; C/8 a power of two <= 32) for pcm_convblock_tail_bwd_sdot: x is the output of the ConvBlock whose backward
 * consumes dx next (src/unet.py:29,47-49: out = u * gate, so d gate = sum_c dout * u) */
PCM_API int pcm_maxpool2_bwd_skip_dot(const void* x, const void* dy, const void* dskip, long long dskip_ns, int dskip_ps,
                                      void* dx, float* sdot, int N, int H, int W, int C, int T, int t_major, int dtype,
                                      pcm_stream_t s);
/* dst(b,p,c) = mean_t src(img(b,t), p, c), img = t*B + b (t-major) or b*T + t */
PCM_API int pcm_time_mean(const void* src, void* dst, long long dst_ns, int dst_ps, int B, int T, int P, int C,
                          int t_major, int dtype, pcm_stream_t s);

/* ---- ConvLSTM cell (src/convlstm.py:11-19) ----------------------------------------------------
 * gates: fp32 [M][4*Ch] pre-activations in order i,f,o,g (bias already added); c_prev nullable (=0).
 * acts: activated gates (dtype) saved for backward; c: fp32 [M][Ch]; h: dtype [M][Ch]. */
PCM_API int pcm_lstm_cell_fwd(const float* gates, const float* c_prev, void* acts, float* c, void* h, int M, int Ch,
                      int dtype, pcm_stream_t s);
/* dh = dh_a + dh_b (either nullable); dc_in nullable.  dgates: dtype [M][4*Ch]; dc_prev fp32 */
PCM_API int pcm_lstm_cell_bwd(const void* dh_a, const void* dh_b, const float* dc_in, const void* acts, const float* c_prev,
                      const float* c, void* dgates, float* dc_prev, int M, int Ch, int dtype, pcm_stream_t s);

/* ---- head (1x1 conv, src/unet_convlstm_attention.py:56,104) and loss (main_final.py:544,559) ---- */
PCM_API int pcm_head_fwd(const void* x, const float* w, const float* b, float* out_nchw, int N, int P, int C, int K,
                 int dtype, pcm_stream_t s);
PCM_API int pcm_head_bwd(const float* dout_nchw, const void* x, const float* w, void* dx, float* dw, float* db, int N, int P,
                 int C, int K, int dtype, pcm_stream_t s);
/* Head + loss fused for the training step (the prediction is needed only inside the loss): ONE pass forward —
 * loss[0] += mean over (n,k,p) of (head(x) - target)^2, pred_nchw optional (NULL: not written) — and ONE pass backward that
 * recomputes head(x) - target per pixel: dx = W^T dpred, dw += dpred x^T, db += dpred with dpred = 2 (pred - target) gscale[0] / n.
 * x NHWC [N][P][C] (dtype), target / pred NCHW fp32 [N][K][P].  C in {16, 32}, K = 2 (pcm_head_mse_supported). */
PCM_API int pcm_head_mse_supported(int C, int K);
PCM_API int pcm_head_mse_fwd(const void* x, const float* w, const float* b, const float* target, float* pred_nchw,
                             float* loss, int N, int P, int C, int K, int dtype, pcm_stream_t s);
PCM_API int pcm_head_mse_bwd(const void* x, const float* w, const float* b, const float* target, const float* gscale,
                             void* dx, float* dw, float* db, int N, int P, int C, int K, int dtype, pcm_stream_t s);
/* loss[0] += mean((a-b)^2) (caller zeroes) */
PCM_API int pcm_mse_fwd(const float* a, const float* b, float* loss, long long n, pcm_stream_t s);
/* da = 2*(a-b)/n * gscale[0] */
PCM_API int pcm_mse_bwd(const float* a, const float* b, const float* gscale, float* da, long long n, pcm_stream_t s);

/* ---- optimizer (torch.optim.Adam as configured at main_final.py:737-747) ------------------------
 * state[0] = step count (float), updated on device so a captured graph advances it. */
PCM_API int pcm_adam_step(float* p, const float* g, float* m, float* v, float* state, long long n, float lr, float b1,
                  float b2, float eps, float wd, float grad_scale, pcm_stream_t s);
/* pcm_adam_step = pcm_adam_tick (advance the device step counter / bias corrections once per step) + pcm_adam_apply
 * (update one parameter range).  Separately they let the data-parallel step run the optimizer bucket by bucket, as
 * each gradient bucket's all-reduce completes. */
PCM_API int pcm_adam_tick(float* state, float b1, float b2, pcm_stream_t s);
PCM_API int pcm_adam_apply(float* p, const float* g, float* m, float* v, const float* state, long long n, float lr,
                   float b1, float b2, float eps, float wd, float grad_scale, pcm_stream_t s);

/* ---- cos(lat)-weighted metric (src/utils_final.py:282-302, main_final.py:616-631) ----------------
 * pred/truth: fp32 [T][V][Y][X]; w_lat: fp64 [Y]; partial: fp64 workspace [V][Y][X][8] (zeroed by the call when
 * zero_first) = per-pixel time sums of p, p^2, t, t^2, (p-t)^2 and the counts of non-NaN p, t, (p-t);
 * out: fp64 [V][3] = monthly_rmse, time_mean_rmse, time_std_mae.
 * pcm_metric_partial accumulates the per-pixel time sums (shardable over T: partials add);
 * pcm_metric_finalize reduces them with the latitude weights.  NaNs are skipped the way xarray's
 * DataArray.weighted(w).mean() / .mean("time") / .std("time") skip them (src/utils_final.py:296, main_final.py:616-631).
 * With w_lat = cos(lat rounded to 2 dp) the same two calls give the Kaggle form of the triplet
 * (_climate_kaggle_metric.py:103-153: weights cos(lat)/sum over the unique lats, mean over lon). */
PCM_API int pcm_metric_partial(const float* pred, const float* truth, double* partial, int T, int V, int Y, int X,
                       int zero_first, pcm_stream_t s);
/* fp64 pred/truth: the DataFrame route of _climate_kaggle_metric.score carries float64 "Prediction" columns */
PCM_API int pcm_metric_partial_f64(const double* pred, const double* truth, double* partial, int T, int V, int Y, int X,
                           int zero_first, pcm_stream_t s);
/* the same accumulation on NORMALISED pred/truth with Normalizer.inverse_transform_output (src/utils_final.py:130-206)
 * fused in: tr: device fp32 [V][4] = (kind, a, b, c), x_phys = g(x*a + b); kind 0 identity (zscore: a = std, b = mean;
 * minimax: a = max - min, b = min), 1 expm1 (log1p), 2 square (sqrt), 3 (.)^(1/c) (pow) — SURVEY §8(f)3 */
PCM_API int pcm_metric_partial_denorm(const float* pred, const float* truth, const float* tr, double* partial, int T,
                                      int V, int Y, int X, int zero_first, pcm_stream_t s);
PCM_API int pcm_metric_finalize(const double* partial, const double* w_lat, double* out, long long T_total, int V, int Y,
                        int X, pcm_stream_t s);

/* ---- Kaggle submission I/O — HOST functions, host pointers, no device work --------------------
 * ID = "t%03d_%s_%.2f_%.2f" % (t_idx, var, lat, lon), rows ordered time > variable > lat > lon
 * (src/utils_final.py:409-449 convert_predictions_to_kaggle_format; var_names = V NUL-terminated strings back to back).
 * pcm_kaggle_format_ids: the IDs, '\n'-separated, into out[cap] (*written bytes; out == NULL: *written = size bound).
 * pcm_kaggle_write_csv: "<id_col>,Prediction" + one row per ID with the shortest float32 text that reads back exactly
 *   (what DataFrame.to_csv(index=False) prints; main_final.py:706-727); host_pred fp32 [T][V][Y][X].
 * pcm_kaggle_parse_ids: the inverse (_climate_kaggle_metric.py:82-96): n '\n'-separated IDs -> time / variable code /
 *   lat / lon; grammar of the reference's re.match pattern; variable names in first-appearance order, NUL-separated,
 *   in names_out; a malformed ID returns PCM_ERR_INVALID ("Invalid ID format: ...") and its index in *bad_row. */
PCM_API int pcm_kaggle_format_ids(char* out, long long cap, long long* written, int T, int V, int Y, int X,
                                  const double* lat, const double* lon, const char* var_names);
PCM_API int pcm_kaggle_write_csv(const char* path, const float* host_pred, int T, int V, int Y, int X,
                                 const double* lat, const double* lon, const char* var_names, const char* id_col);
PCM_API int pcm_kaggle_parse_ids(const char* buf, long long nbytes, long long n, long long* time, int* var_code,
                                 double* lat, double* lon, char* names_out, int names_cap, int* n_vars,
                                 long long* bad_row);

#ifdef __cplusplus
}
#endif
#endif /* PCM_B200_H_ */
