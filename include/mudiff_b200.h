/*
 * mudiff_b200.h - C ABI of libmudiff_b200.so (hand-written sm_100a CUDA kernels for the
 * MU-Diff reverse-sampling hot path).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless stated; inputs are borrowed, outputs are
 *     caller-allocated (the reference allocates with at::empty inside its op,
 *     utils/op/upfirdn2d_kernel.cu:244-245; here the Python wrapper allocates with torch);
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous and
 *     CUDA-graph-capture safe (no allocation, no host<->device copy, no sync);
 *   - return value: 0 = launched, >0 = cudaError_t of the launch, <0 = -EINVAL style
 *     argument reject (MUDIFF_EINVAL / MUDIFF_EUNSUPPORTED).  Nothing falls back to CPU;
 *   - dtype codes: MUDIFF_F32 = 0, MUDIFF_BF16 = 1, MUDIFF_F16 = 2;
 *   - activations are NHWC ("channels_last"): element (b,y,x,c) at ((b*H+y)*W+x)*ld + c.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the
 * MU-Diff repository root).
 */
#ifndef MUDIFF_B200_H
#define MUDIFF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MUDIFF_F32 0
#define MUDIFF_BF16 1
#define MUDIFF_F16 2

#define MUDIFF_EINVAL (-22)
#define MUDIFF_EUNSUPPORTED (-95)

#define MUDIFF_ACT_NONE 0
#define MUDIFF_ACT_SILU 1
#define MUDIFF_ACT_SIGMOID 2
#define MUDIFF_ACT_TANH 3
#define MUDIFF_ACT_LRELU 4   /* LeakyReLU(0.2): the discriminator's activation (backbones/discriminator.py:178) */

/* ABI version + build info (static string). */
int mudiff_abi_version(void);
const char* mudiff_build_info(void);
/* Number of kernels launched by this library since load (for bench.py `gpu_launches`). */
int64_t mudiff_launch_count(void);
/* sizeof(mudiff_conv_desc) as compiled, so bindings can verify their struct layout. */
int mudiff_conv_desc_size(void);

/* ---------------------------------------------------------------------------------
 * FIR resampling.  Replaces upfirdn2d_op.upfirdn2d(input[major,H,W,minor], kernel[kh,kw],
 * up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1)  (utils/op/upfirdn2d.cpp:20-31,
 * utils/op/upfirdn2d_kernel.cu:211-371).  Same semantics: zero-insert upsample, pad
 * (negative = crop), TRUE convolution with `kernel` (the kernel is flipped, :139),
 * decimate.  out is [major, out_h, out_w, minor], out_h = (in_h*up_y+pad_y0+pad_y1-kh)/down_y+1.
 * NCHW tensors use major=N*C, minor=1 (what utils/op/upfirdn2d.py:124 does); NHWC tensors
 * use major=N, minor=C (vectorised path when minor % 8 == 0).
 * `kernel` is float32 on the device regardless of dtype.  kh,kw <= 32.
 * ------------------------------------------------------------------------------- */
int mudiff_upfirdn2d(const void* in, void* out, const float* kernel, int dtype,
                     int64_t major, int in_h, int in_w, int minor, int kh, int kw,
                     int up_x, int up_y, int down_x, int down_y,
                     int pad_x0, int pad_x1, int pad_y0, int pad_y1, void* stream);

/* Both resampled branches of a resample ResnetBlockBigGANpp_Adagn (backbones/layerspp.py:293-305: h = act(AdaGN(x)),
 * h = up/downsample_2d(h), x = up/downsample_2d(x)) from ONE read of x:
 *   out_h = FIR(act(x * scale[b][c] + shift[b][c])),  out_x = FIR(x)
 * x / out_h / out_x channels-last [batch, H, W, channels] dense, dtype MUDIFF_F32 | MUDIFF_BF16; kernel = the 4x4 fp32 FIR
 * (gain folded in, as upfirdn2d gets it); table = float[batch][table_ld][2] from mudiff_gn_scale_shift;
 * act MUDIFF_ACT_NONE | MUDIFF_ACT_SILU.  Modes: (up 2, down 1, pad 2/1) = upsample_2d, (up 1, down 2, pad 1/1) =
 * downsample_2d (backbones/up_or_down_sampling.py:200-262); anything else MUDIFF_EUNSUPPORTED. */
int mudiff_upfirdn2d_gn(const void* x, void* out_h, void* out_x, const float* kernel, const float* table,
                        int table_ld, int act, int dtype, int batch, int in_h, int in_w, int channels,
                        int up, int down, int pad0, int pad1, void* stream);

/* fp32 parity path on the tensor cores: x (fp32 channels-last, pixel stride ld) -> bf16 with x == hi + mid + lo exactly;
 * layout 0: [pixels][3c] = (lo | mid | hi) (activation side), layout 1: [pixels][6c] = (lo | mid mid | hi hi hi) (the
 * K-major "weight" side when that operand is an activation too: attention scores / PV / V^T).  The host then runs the
 * tcgen05 conv on three channel-slice segments, small products first: [hi] x w_lo, [mid hi] x w_mid, [lo mid hi] x w_hi
 * (the six products above 2^-24) with fp32 accumulation in TMEM. */
int mudiff_split3_bf16(const float* x, int ld, void* out, int64_t pixels, int c, int layout, void* stream);

/* Replaces fused.fused_bias_act(input, bias, refer, act, grad, alpha, scale)
 * (utils/op/fused_bias_act.cpp:18-27, fused_bias_act_kernel.cu:20-51):
 *   x += bias[(i / step_b) % size_b] (if bias); act 1 = linear, 3 = leaky-relu(alpha);
 *   grad 0: y = act(x)*scale; grad 1: y = x * dact(ref) * scale; grad 2: 0 (second order).
 * bias/ref may be NULL. */
int mudiff_fused_bias_act(const void* x, const void* bias, const void* ref, void* out, int dtype,
                          int64_t n, int size_b, int64_t step_b, int act, int grad,
                          float alpha, float scale, void* stream);

/* Minibatch standard deviation feature of Discriminator_large.forward (backbones/discriminator.py:243-250):
 * x = channels-last [batch, hw, C] (pixel stride ld), group = min(batch, 4), n_sub = batch / group;
 * s[m] = mean over (c, pixel) of sqrt(var over g of x[g * n_sub + m] + 1e-8) (biased variance);
 * writes s[b % n_sub] into channel `out_c` of out (channels-last, pixel stride out_ld) for every pixel of sample b.
 * dtype = dtype of x and out.  batch must be a multiple of group. */
int mudiff_minibatch_stddev(const void* x, int ld, void* out, int out_ld, int out_c, int dtype,
                            int batch, int group, int channels, int hw, void* stream);

/* ---------------------------------------------------------------------------------
 * Posterior update, one fused kernel.  Replaces sample_posterior_combine
 * (engine/test.py:150-177):  mean = ((c1[t]*x01 + c2[t]*xt) + (c1[t]*x02 + c2[t]*xt))/2,
 * out = mean + (t!=0) * exp(0.5*logvar[t]) * noise.   All fp32; t is int64 [B];
 * c1/c2/logvar are fp32 device tables of length n_steps; per_sample = C*H*W;
 * x01/x02 may be channel-0 slices of wider tensors: sample stride given in elements.
 * ------------------------------------------------------------------------------- */
int mudiff_posterior_update(const float* x01, int64_t x01_bstride, const float* x02, int64_t x02_bstride,
                            const float* xt, const float* noise, const int64_t* t,
                            const float* coef1, const float* coef2, const float* logvar, int n_steps,
                            float* out, int batch, int64_t per_sample, void* stream);

/* ---------------------------------------------------------------------------------
 * GroupNorm (+ AdaGN scale/shift) (+ SiLU).  Replaces AdaptiveGroupNorm.forward
 * (backbones/layerspp.py:47-54), GroupNorm_Conv (:56-65), nn.GroupNorm call sites
 * (:103,:113; ncsnpp_generator_adagn_feat.py:265,436) and the following self.act.
 * Input is the channel-concatenation of up to two NHWC tensors (x0: C0 channels with
 * pixel stride ld0, x1: C1 with ld1; x1 may be NULL) - this is torch.cat([h, hs.pop()], 1)
 * of ncsnpp_generator_adagn_feat.py:383 without materialising it.
 * Statistics are PER CHANNEL: chstats double[B][st_ld][2] = (sum, sum of squares) over the H*W pixels; a
 * tensor's statistics are produced either by mudiff_gn_stats (stand-alone pass) or by the epilogue of
 * the convolution that wrote the tensor (mudiff_conv_desc.stats + mudiff_stats_finalize).  Both are
 * deterministic and batch-invariant (ordered partial sums, no floating-point atomics).  The apply
 * kernel derives the group mean / rstd of the (concatenated) input from the per-channel values, so any
 * grouping and any concat of tensors with known statistics needs no extra pass.
 * apply: y = act(gamma[b,c] * (x-mean)*rstd + beta[b,c]); gamma/beta fp32 with batch
 * stride gb_bstride (0 => shared affine [C]); NULL gamma/beta => 1/0.  G groups over C0+C1 channels.
 * ------------------------------------------------------------------------------- */
int mudiff_gn_stats(const void* x, int c, int ld, int dtype, int batch, int64_t hw,
                    double* chstats, int st_ld, int st_off, void* stream);
/* Single-pass GroupNorm (+ AdaGN gamma/beta, + SiLU) of a dense bf16 NHWC tensor: out = act(GN(x)*gamma + beta) AND the
 * per-channel (sum, sumsq) statistics of x (same layout as mudiff_gn_stats) with ONE read of x: a persistent
 * cooperative grid keeps every CTA's chunk of an image in shared memory between the statistics and the apply phase.
 * MUDIFF_EUNSUPPORTED when the image is too large for the stages (callers then use gn_stats + gn_apply). */
int mudiff_gn_fused(const void* x, void* out, int c, int batch, int64_t hw, int groups, const float* gamma,
                    const float* beta, int64_t gb_bstride, float eps, int act, double* chstats, int st_ld,
                    int st_off, void* stream);

/* Folded GroupNorm / AdaGN parameters table[b][c] = (gamma*rstd, beta - mean*gamma*rstd) (float pairs, [batch][c0+c1])
 * from per-channel (sum, sumsq) statistics of one or two concatenated sources; consumed by mudiff_conv_tc's
 * A-operand transform (a_xform). */
int mudiff_gn_scale_shift(const double* st0, int st0_ld, int c0, const double* st1, int st1_ld, int c1,
                          const float* gamma, const float* beta, int64_t gb_bstride, int batch, int64_t hw,
                          int groups, float eps, float* table, void* stream);
/* mudiff_gn_stats(x0 -> st0) followed by mudiff_gn_scale_shift([st0 | st1] -> table) as ONE launch (the block that completes an
 * image's statistics writes its table rows); identical values, one launch less per GroupNorm that a consumer applies itself. */
int mudiff_gn_stats_table(const void* x0, int c0, int ld0, int dtype, double* st0, int st0_ld,
                          int c1, const double* st1, int st1_ld, const float* gamma, const float* beta,
                          int64_t gb_bstride, int batch, int64_t hw, int groups, float eps, float* table, void* stream);

/* partial float[batch*rows_per_image][n][2] (written by mudiff_conv_tc: one row per pixel tile and TMEM lane quadrant,
 * rows_per_image = 4 * tiles per image) -> chstats[b][st_off + c][2]; fixed summation order (deterministic) */
int mudiff_stats_finalize(const float* partial, int tiles_per_image, int n, double* chstats, int st_ld,
                          int st_off, int batch, void* stream);
/* mudiff_gn_stats(x0) + mudiff_gn_apply([x0 | x1]) in ONE launch (AdaptiveGroupNorm + SiLU, backbones/layerspp.py:47-54,
 * 293, 314): every block re-reads its own 64 KB chunk from L2 after the image's statistics are complete, so the tensor crosses
 * HBM twice (read, write) instead of three times.  st0 receives x0's per-channel (sum, sumsq); x1 / st1 (optional) is the
 * channel-concat partner with known statistics.  bf16 only; bit-identical to the two-call sequence; MUDIFF_EUNSUPPORTED
 * when the shape does not qualify. */
int mudiff_gn_stats_apply(const void* x0, int c0, int ld0, double* st0, int st0_ld,
                          const void* x1, int c1, int ld1, const double* st1, int st1_ld, int dtype,
                          const float* gamma, const float* beta, int64_t gb_bstride,
                          void* out, int ld_out, int batch, int64_t hw, int groups, float eps, int act, void* stream);
int mudiff_gn_apply(const void* x0, int c0, int ld0, const double* st0, int st0_ld,
                    const void* x1, int c1, int ld1, const double* st1, int st1_ld, int dtype_in,
                    const float* gamma, const float* beta, int64_t gb_bstride,
                    void* out, int ld_out, int dtype_out, int batch, int64_t hw, int groups,
                    float eps, int act, void* stream);
int mudiff_zero(void* p, int64_t nbytes, void* stream);

/* ---------------------------------------------------------------------------------
 * Convolution / contraction descriptor shared by the tensor-core and SIMT kernels.
 * Replaces the nn.Conv2d / F.conv2d / NIN / einsum call sites of the generator
 * (backbones/layers.py:124,502-505; backbones/layerspp.py:118,122;
 *  backbones/up_or_down_sampling.py:183).
 *
 *  out[b,y,x,coff+n] = act( alpha * ( sum_seg sum_tap sum_c A_seg[b, y*s+dy, x*s+dx, c] * Wt[n, k(seg,tap,c)]
 *                                     + bias[n] + rowbias[b,n] ) + beta * residual[b,y,x,n] )
 *
 * A segments: up to 3 NHWC tensors with identical B,H,W (channel concat and the fused
 * 1x1 shortcut of ResnetBlockBigGANpp_Adagn, layerspp.py:318-324).  taps = 9 (3x3, pad 1)
 * or 1 (1x1).  Weights are packed [w_batch][N][Ktot] (K-major) with
 * k = seg_offset + tap*C_seg + c;  w_bstride = 0 => shared weights, else elements between
 * per-sample weight matrices (attention QK^T / PV).  a_bstride0 = 1 => segment tensors
 * are per-sample (normal); 0 => A shared across the batch (swapped GEMM, A = weights).
 * ------------------------------------------------------------------------------- */
typedef struct mudiff_conv_desc {
  const void* a[3];        /* NHWC tensors                                         */
  int32_t a_c[3];          /* channels used from each                              */
  int32_t a_ld[3];         /* pixel stride (elements) of each                      */
  int32_t a_taps[3];       /* 9 or 1                                               */
  int32_t nseg;
  int32_t a_batched;       /* 1: A indexed by b; 0: A shared over batch            */
  int32_t batch, h, w;     /* INPUT spatial size (output too when stride == 1)     */
  int32_t stride;          /* 1, or 2 (SIMT only; pad 0: conv_downsample_2d)       */
  int32_t pad;             /* 1 for 3x3 'same', 0 for 1x1 / strided                */
  const void* wt;          /* packed weights, dtype = a dtype                      */
  int64_t w_bstride;
  int32_t w_ld;            /* row stride of wt in elements; 0 => Ktot              */
  int32_t n;               /* output channels                                      */
  const float* bias;       /* [n] or NULL                                          */
  const float* rowbias;    /* [batch][rowbias_ld] or NULL (Dense_0(act(temb)))     */
  int32_t rowbias_ld;
  const void* residual;    /* NHWC [.., res_ld] dtype = out dtype, or NULL         */
  int32_t res_ld;
  float alpha, beta;
  int32_t act;             /* MUDIFF_ACT_*                                         */
  void* out;
  int32_t out_ld, out_coff;
  int32_t out_dtype;
  void* stats;             /* optional (mudiff_conv_tc, n <= 256): float partial    */
  int32_t stats_groups;    /*   [batch*tiles_per_image][n][2] per-tile per-channel  */
                           /*   (sum,sumsq) of the OUTPUT; stats_groups is unused   */
  int32_t flags;           /* debug/ablation: bit1 forbid halo staging, bit2 descriptor
                              base-offset variant, bit3 forbid stationary weights,
                              bit4 one pixel tile per unit.  bit15 (0x8000): decimate - keep only the
                              odd (y, x) outputs of the 'same' 3x3 conv and write them at ((y-1)/2,
                              (x-1)/2) of an [(h-1)/2, (w-1)/2] output: this IS the stride-2 VALID conv
                              of conv_downsample_2d (up_or_down_sampling.py:183) on the tensor cores */
  /* mudiff_conv_tc only: A-operand transform.  a_xform[i] != NULL: segment i is read as
   * act(x * scale + shift) with (scale, shift) float pairs a_xform[i][(b * a_xform_ld[i] + c) * 2 + {0,1}] -
   * the GroupNorm / AdaGN + SiLU that precedes the conv (layerspp.py:293-314) applied to the staged tile in
   * shared memory, so the normalised tensor is never stored.  Zero padding is applied AFTER the transform. */
  const float* a_xform[3];
  int32_t a_xform_ld[3];   /* row stride of the table in (scale, shift) pairs */
  int32_t a_xform_act;     /* MUDIFF_ACT_NONE or MUDIFF_ACT_SILU */
} mudiff_conv_desc;

/* tcgen05/TMEM/TMA implicit GEMM (bf16 in, fp32 accumulate).  Requires a_c[i] % 64 == 0,
 * n % 32 == 0, stride == 1.  Returns MUDIFF_EUNSUPPORTED otherwise. */
int mudiff_conv_tc(const mudiff_conv_desc* d, void* stream);

/* Fused stem of ConvFeatBlock / ConvBlock / ConvBlock_GAP (backbones/layerspp.py:394-501):
 * conv3x3(1 -> n) -> GroupNorm / AdaGN -> SiLU with the raw conv output never stored.  The GroupNorm statistics
 * follow from second moments of the 1-channel input: mudiff_stem_moments writes double[batch][54]
 * (9 patch sums + 45 patch products), mudiff_stem_conv_gn_act consumes them. */
int mudiff_stem_conv_tc(const float* x, const float* wt, const float* bias, const float* scale_shift, int act,
                        void* out, int out_ld, int out_coff, int out_dtype, int batch, int h, int w, int n, void* stream);
int mudiff_stem_moments(const float* x, int ld, int batch, int h, int w, double* moments, void* stream);
int mudiff_stem_conv_gn_act(const float* x, int ld, const float* wt, const float* bias, const double* moments,
                            const float* gamma, const float* beta, int64_t gb_bstride, int groups, float eps,
                            int act, float* scale_shift /* workspace float[batch][n][2] */,
                            void* out, int out_ld, int out_coff, int out_dtype,
                            int batch, int h, int w, int n, void* stream);
/* Planning only: out[0..9] = tile_h, tile_w, tiles_per_image, n_tile, tiles_per_unit, stationary_weights,
 * a_slots, b_slots, accumulator_stages, halo(seg0).  Callers size `stats` with out[2]. */
int mudiff_conv_tc_query(const mudiff_conv_desc* d, int32_t* out);

/* Fused attention of AttnBlockpp (backbones/layerspp.py:118-122: einsum -> softmax -> einsum) in one tcgen05 kernel:
 * out[b] = softmax(Q[b] K[b]^T * scale) V[b].  qk [B, L, 2C] bf16 (q | k per token), vt = V^T [B, C, L] bf16,
 * out [B, L, C] bf16.  C == 256, L % 128 == 0, else MUDIFF_EUNSUPPORTED (callers use the unfused kernels). */
int mudiff_attention_tc(const void* qk, const void* vt, void* out, int batch, int L, int C, float scale, void* stream);
/* -------------------------------------------------------------------------------
 * Volume prediction front / back end (engine/test_volume.py:135-191, 269-294).
 * mudiff_volume_window: exact robust [pmin, pmax] percentile window over the voxels != 0 of a fp32 volume
 * (np.percentile 'linear' semantics, fall back to min / max, degenerate -> all-zero output), kept on the device inside
 * `workspace` (mudiff_volume_workspace_bytes() bytes).  mudiff_volume_to_slices: out [n,1,sh,sw] = the normalised
 * ([-1, 1]) axial slices s0..s0+n-1 of vol [H, W, Z], bilinear-resized (align_corners = False) when (sh, sw) != (H, W).
 * mudiff_slices_to_volume: vol [H, W, Z] = zeros with slices s0.. = pred [n,1,H,W] (mapped (x+1)/2, clamped, if to01).
 * mudiff_volume_window_read: device float[3] <- lo, hi, status(int bits) of the window in `workspace`.
 * ------------------------------------------------------------------------------- */
int mudiff_volume_workspace_bytes(void);
int mudiff_volume_window(const float* vol, int64_t n, float pmin, float pmax, void* workspace, void* stream);
int mudiff_volume_window_read(const void* workspace, float* window, void* stream);
int mudiff_volume_to_slices(const float* vol, int h, int w, int z, int s0, int n, int sh, int sw,
                            const void* workspace, float* out, void* stream);
int mudiff_slices_to_volume(const float* pred, int h, int w, int z, int s0, int n, int to01, float* vol, void* stream);

/* Slice-test driver helpers (engine/test.py:265-400, dataset/dataset_brats.py:73-92):
 * mudiff_zscore_to_unit: out = clamp(in, -3, 3) / 3.  mudiff_minmax_keys: exact global min / max (order-preserving
 * uint32 keys, device uint32[2]; accumulate != 0 folds another tensor in).  mudiff_scale_to_u8: the reference's
 * np.clip((x - gmin) / (gmax - gmin) * 255, 0, 255).astype(uint8) with that window ([0, 1] if constant). */
int mudiff_zscore_to_unit(const float* in, float* out, int64_t n, void* stream);
int mudiff_minmax_keys(const float* x, int64_t n, int accumulate, unsigned int* keys, void* stream);
int mudiff_minmax_read(const unsigned int* keys, float* window, void* stream);
int mudiff_scale_to_u8(const float* x, int64_t n, const unsigned int* keys, unsigned char* out, void* stream);

/* Debug: out[0..7] = (timed_out, block, warp, lane, barrier smem address, parity, grid, 0) of the last mbarrier
 * wait that hit its 4e9-cycle bound inside mudiff_conv_tc (kept in mapped host memory). */
int mudiff_debug_last_timeout(int32_t* out);
/* Bound (clock64 cycles) of the mbarrier waits inside the tcgen05 kernels; 0 = unbounded (profilers, time-sliced GPUs).
 * Applies to the current device; synchronising (cudaMemcpyToSymbol). */
int mudiff_set_wait_timeout(long long cycles);
int mudiff_debug_dump(int32_t* out, int n);   /* header + the stuck CTA's barrier block / per-warp progress records */
int mudiff_debug_selftest(void);   /* 1 if the mapped-host debug channel works */
/* CUDA-core implicit GEMM (fp32 or bf16 storage, fp32 math): any shape, stride 1/2.
 * This is the fp32-parity path and the path for Cin=1 / Cout=1 / strided convs. */
int mudiff_conv_simt(const mudiff_conv_desc* d, int dtype, void* stream);

/* ---------------------------------------------------------------------------------
 * Attention pieces (backbones/layerspp.py:118-122).
 * Row softmax, in place allowed: y[r,:] = softmax(scale * x[r,:]) over `cols`.
 * ------------------------------------------------------------------------------- */
int mudiff_softmax_rows(const void* x, void* y, int dtype, int64_t rows, int cols, float scale, void* stream);

/* ---------------------------------------------------------------------------------
 * Small dense layers on embeddings (nn.Linear call sites: dense_layer.py:67-71,
 * ncsnpp_generator_adagn_feat.py:105-110,271-277; layerspp.py:42,277).
 * out[b,j] = act_out( sum_k act_in(in[b,k]) * W[j,k] + bias[j] ), all fp32.
 * Batched over many layers by concatenating W rows.
 * ------------------------------------------------------------------------------- */
int mudiff_linear(const float* in, int in_ld, const float* w, const float* bias, float* out, int out_ld,
                  int batch, int k, int j, int act_in, int act_out, void* stream);
/* layers.py:465-479 get_timestep_embedding(t int64 [B], dim) -> fp32 [B, dim] */
int mudiff_timestep_embedding(const int64_t* t, float* out, int batch, int dim, float max_positions, void* stream);
/* PixelNorm (ncsnpp_generator_adagn_feat.py:44-49): z / sqrt(mean(z^2, dim=1) + 1e-8) */
int mudiff_pixelnorm(const float* z, float* out, int batch, int dim, void* stream);

/* ---------------------------------------------------------------------------------
 * Elementwise glue on NHWC tensors (dtype f32/bf16), n = number of elements.
 *   gate_mul   : out = a * b                          (ncsnpp_generator_adagn_feat.py:778)
 *   gate_blend : out = g*a + (1-g)*b                  (:779), out written at channel offset
 *   add_scale  : out = (a + b) * scale                (:363 input-pyramid residual)
 *   copy_channels : dst[p, coff:coff+c] = src[p, 0:c] (torch.cat)
 *   gap        : out[b,c] = mean_p x[b,p,c]  fp32     (layerspp.py:473,491)
 *   cast       : dtype conversion / strided channel copy
 * ------------------------------------------------------------------------------- */
int mudiff_gate_mul(const void* a, int a_ld, const void* b, int b_ld, void* out, int out_ld,
                    int dtype, int64_t pixels, int c, void* stream);
int mudiff_gate_blend(const void* g, int g_ld, const void* a, int a_ld, const void* b, int b_ld,
                      void* out, int out_ld, int dtype, int64_t pixels, int c, void* stream);
int mudiff_add_scale(const void* a, const void* b, void* out, int dtype, int64_t n, float scale, void* stream);
int mudiff_copy_channels(const void* src, int src_ld, int src_dtype, void* dst, int dst_ld, int dst_dtype,
                         int64_t pixels, int c, void* stream);
int mudiff_gap(const void* x, int ld, int dtype, float* out, int batch, int64_t hw, int c, void* stream);
int mudiff_tanh(const void* x, void* out, int dtype_in, int dtype_out, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MUDIFF_B200_H */
