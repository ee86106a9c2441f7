/* mgf.h -- C ABI of libmgf_sm100a.so: B200 (sm_100a) kernels for the GANformer synthesis + latent-projection
 * hot path.  Plain pointers and sizes only; the caller owns every buffer (device memory unless stated),
 * sets the current device, and passes the CUDA stream to launch on.  No function allocates or frees
 * device memory, synchronises the host, or keeps a pointer after it returns, so every call can be
 * captured in a CUDA graph.
 *
 * Return value: 0 = ok, negative = bad argument (MGF_E_*), positive = cudaError_t.  The text of the last
 * error on the calling thread is returned by mgf_last_error().  There is no CPU fallback anywhere.
 *
 * Each entry point cites the reference interface (file:line under the MorphGANformer tree) it replaces.
 */
#ifndef MGF_H_
#define MGF_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define MGF_E_BADARG  (-1)
#define MGF_E_DTYPE   (-2)
#define MGF_E_SHAPE   (-3)
#define MGF_E_ALIGN   (-4)
#define MGF_E_UNSUP   (-5)
#define MGF_E_DRIVER  (-6)

/* dtype codes */
#define MGF_F32  0
#define MGF_BF16 1
#define MGF_F16  2
#define MGF_F64  3

const char* mgf_last_error(void);
int mgf_version(void);
/* number of kernel launches issued through this library since load (bench.py's gpu_launches claim) */
int64_t mgf_launch_count(void);
/* Element type of the 16-bit FORWARD tensors of the engine kernels below (activations, forward GEMM operands, cached LPIPS target
 * features): MGF_F16 (default: the parity-green mode) or MGF_BF16.  Gradient tensors are always bf16.  Process-wide setting, read at launch time. */
int mgf_set_forward_dtype(int dtype);
int mgf_get_forward_dtype(void);
/* fp16-forward mode stores saturate at +-65504; kernels that write forward activations (conv_tc epilogue, attention, skip FIR+add,
 * weight modulation, channel scaling) OR 1 into a per-device flag when a value was outside that range.  Reads the flag (synchronises
 * `stream`), optionally clears it.  A set flag means: results are clipped -- switch to MGF_BF16 forward storage for this checkpoint. */
int mgf_fp16_overflow_read(int* flag_host, int reset, void* stream);

/* ---- bias_act ---------------------------------------------------------------------------------------
 * Replaces bias_act_plugin.bias_act (torch_utils/ops/bias_act.cpp:24-82, kernel bias_act.cu:15-139).
 * y = clamp(act(x + b[(i / stepB) % sizeB]) * gain) for grad=0; first/second derivative forms for
 * grad=1/2 exactly as the reference kernel defines them (x = incoming gradient, xref/yref = saved
 * input/output, dy = saved gradient for grad=2).  Null pointer = absent tensor.  act = 1..9
 * (linear, relu, lrelu, tanh, sigmoid, elu, selu, softplus, swish), bias_act.py:15-25.  clamp<0 = off. */
int mgf_bias_act(const void* x, const void* b, const void* xref, const void* yref, const void* dy, void* y,
                 int dtype, int grad, int act, float alpha, float gain, float clamp,
                 int64_t sizeX, int64_t sizeB, int64_t stepB, void* stream);

/* ---- upfirdn2d --------------------------------------------------------------------------------------
 * Replaces upfirdn2d_plugin.upfirdn2d (torch_utils/ops/upfirdn2d.cpp:8-86, kernels upfirdn2d.cu:21-192):
 * zero-insert upsample (up), pad/crop (pad0 = padx0,pady0; the far-side padding is implied by outSize),
 * 2-D FIR with f (fp32, [fH,fW] with element strides fStride; flipped unless flip!=0), decimate (down),
 * times gain.  Sizes/strides are in elements, order {W,H,C,N} like upfirdn2d.h:6-32. */
int mgf_upfirdn2d(const void* x, const float* f, void* y, int dtype,
                  const int64_t inSize[4], const int64_t inStride[4],
                  const int32_t fSize[2], const int64_t fStride[2],
                  const int64_t outSize[4], const int64_t outStride[4],
                  const int32_t up[2], const int32_t down[2], const int32_t pad0[2],
                  int flip, float gain, void* stream);

/* ---- exact-fp32 direct convolution (the conv2d_gradfix surface) -----------------------------------
 * Replaces the ATen/cuDNN calls behind conv2d_gradfix.conv2d / conv_transpose2d
 * (torch_utils/ops/conv2d_gradfix.py:27-35) and their gradients (:96-157).  NCHW contiguous fp32,
 * weight [OC, IC/groups, KH, KW] contiguous; fp32 FFMA accumulation (no TF32), used for the 1e-4 parity path.
 *   fwd   : y[N,OC,HO,WO]   = conv(x[N,IC,H,W], w)
 *   dgrad : dx[N,IC,H,W]    = conv_transpose(dy[N,OC,HO,WO], w)      (also conv_transpose2d forward)
 *   wgrad : dw[OC,IC/g,KH,KW] = sum_n,oy,ox dy * x                   (dw must be zeroed by the caller; split-K atomics) */
typedef struct {
  int32_t N, IC, H, W, OC, HO, WO, KH, KW;
  int32_t stride_h, stride_w, pad_h, pad_w, dil_h, dil_w, groups;
} mgf_conv_shape;
int mgf_conv2d_fwd_f32(const float* x, const float* w, const float* bias, float* y, const mgf_conv_shape* s, void* stream);
int mgf_conv2d_dgrad_f32(const float* dy, const float* w, float* dx, const mgf_conv_shape* s, void* stream);
int mgf_conv2d_wgrad_f32(const float* dy, const float* x, float* dw, const mgf_conv_shape* s, void* stream);

/* ---- small elementwise helpers used by the Python host mirror -------------------------------------
 * fma: out = a*b + c with b broadcast per (n,c) and c broadcast per (h,w) or full (fma.py:7-17 as used by
 * networks.py:322).  bmode: 0 = b full, 1 = b is [N,C,1,1]; cmode: 0 = none, 1 = c full, 2 = c is [H,W] plane. */
int mgf_fma(const void* a, const void* b, const void* c, void* out, int dtype,
            int64_t N, int64_t C, int64_t HW, int bmode, int cmode, void* stream);

/* ---- tcgen05 implicit-GEMM convolution (bf16 in, fp32 accumulate in TMEM) --------------------------
 * Replaces, for the bf16 engine, the library calls behind modulated_conv2d (training/networks.py:252-328 ->
 * conv2d_resample.py:117-139 -> conv2d_gradfix.py:27-35 -> cuDNN grouped conv / conv_transpose) and the torchvision
 * VGG16 convolutions of LPIPS (lpips/pretrained_networks.py:97-135), forward and input-gradient.
 *
 * out[b, y*osy+ofy[ph], x*osx+ofx[ph], co] = epilogue( sum_{tap t, channel k}  A_{amap(t)}[b, y+dy(t), x+dx(t), k]
 *                                                     * W[g][wz(t)][ph*Cout + co][k] )
 * for (b, y, x) over the NB x GH x GW tile grid; reads outside an activation tensor are zero (TMA fill).
 * Activations are NHWC bf16 views (channel stride 1; W/H/N strides in elements, so strided phase views work);
 * W is a dense bf16 [G][T][NT = phases*Cout][K] tensor, G = NB for per-sample (modulated) weights or 1.
 * 16-bit element types: tensors flagged *_fwd follow mgf_set_forward_dtype, the others are bf16 gradients.
 * Epilogue, in order (each optional): reduce_out[b, n] += sum_pixels acc*X   (fp32 atomics; X like out),
 * acc *= scale_n[b, n], += noise[oy, ox] * *noise_strength, += bias[co], act (0 none, 1 leaky-ReLU(alpha), 2 ReLU) * gain,
 * += add (like out), *= (X > 0 ? 1 : ag_alpha) * ag_gain  (actgrad), store bf16.  bn = 0 picks the N tile. */
typedef struct { const void* ptr; int64_t C, W, H, N, sW, sH, sN; } mgf_tc_act;
typedef struct { int8_t amap, dy, dx, _pad; int32_t wz; } mgf_tc_tap;
typedef struct {
  mgf_tc_act a[4]; int32_t n_a;
  const void* w; int64_t w_G, w_T, w_NT, w_K;
  mgf_tc_tap taps[40]; int32_t ntaps;
  int32_t GW, GH, NB;
  int32_t phases, Cout;
  void* out; int64_t OH, OW, OC; int32_t osy, osx; int32_t ofy[4], ofx[4];
  const float* scale_n; float* reduce_out; const void* X;
  const float* noise; const float* noise_strength; const float* bias;
  int32_t act; float alpha, gain;
  const void* add;
  int32_t actgrad; float ag_alpha, ag_gain;
  int32_t bn; int32_t reduce_per_sample;
  /* which tensors are FORWARD-dtype tensors (see mgf_set_forward_dtype); 0 = gradient tensor (bf16) */
  int32_t ab_fwd, out_fwd, x_fwd, add_fwd;
  /* rows are pairs of 32-channel pixels viewed as one 64-channel super-pixel (the caller passes block-expanded weights): the
   * noise plane is then [GH, 2*GW] and the two 32-column halves of a row get their own noise value */
  int32_t superpix;
  /* elements between the noise planes of consecutive samples: 0 = one [OH, OW] plane shared by all samples (noise_mode 'const'),
   * OH*OW = per-sample planes [NB, OH, OW] (noise_mode 'random', reference networks.py:1015-1017) */
  int64_t noise_bstride;
  /* per-phase tap lists (transposed convolution, conv2d_resample.py:117-134 first stage): when phase_ntaps[0] > 0 the `taps` array is the
   * concatenation of `phases` lists (phase p owns phase_ntaps[p] entries), every phase reads the SAME [T][Cout][K] weights (w_NT == Cout)
   * and writes its own strided output view (osy/osx/ofy/ofx).  Forward-type launches only (no reduce_out / X). */
  int32_t phase_ntaps[4];
} mgf_conv_tc_desc;
int mgf_conv_tc(const mgf_conv_tc_desc* d, void* stream);
/* debugging / A-B measurement switches of mgf_conv_tc (process-wide; production value: 1).  Bit 0: halo (shared-memory tap reuse) kernel for the
 * 64- / 32-channel and 128 -> 64 channel 3x3 layers; bits 1..2: staging tiles per epilogue group in the halo kernel (0 = automatic); bit 3: halo kernel
 * for multi-phase launches too; bit 4: disable the CTA-pair (cta_group::2) kernel of the wide tiles; bit 5: two epilogue groups everywhere instead of
 * three where the tile tail has work; bit 6: enable the CTA-pair halo kernel (correct but measured slower, off by default); bits 8..12: stage-isolation
 * experiments of scripts/bench_halo.py (results are wrong while set). */
int mgf_conv_tc_set_halo(int mode);

/* ---- bf16 synthesis-engine helpers (engine_kernels.cu); activations NHWC bf16, coefficients fp32 ------------------
 * style_fwd : s[b,i] = ((wg[b,:] . A[i,:]) * again + abias[i]) * sgain  (FullyConnectedLayer affine, networks.py:138-150, :1022,
 *             :1056-1059) and, when Wsq is given, d[b,o] = rsqrt(sum_i s^2 Wsq[o,i] + 1e-8) (demodulation, :291).
 * style_bwd : dwg[b,:] += d(loss)/d(wg) from ds (direct style gradient) and R[b,o] = sum_pixels dy*y (demodulation path).
 * modulate_weights : out[b][t][n][k] = bf16(base[t][n][k] * rs[b][n % nmod] * cs[b][k])  -- the per-sample "fused_modconv"
 *             weights (:288-293) laid out as tcgen05 B operands.
 * small_gemm: out[b,m,n] (+)= sum_k A[b,m,k] Bm[n,k] + bias[n] (attention value/modulation fold, tiny).
 * torgb_*   : ToRGBLayer (:1054-1065) forward to the public fp32 NCHW image and its backward.
 * act_bwd   : leaky-ReLU backward of a fused conv epilogue + R[b,o] accumulation.
 * upfir2_*  : NHWC [1,3,3,1] FIR up-sampling x2 with residual add (resnet skip, :245-250, :1157-1160) and its adjoint;
 *             fk4 is a HOST array of the four 1-D taps. */
int mgf_style_fwd(const float* wg, int64_t wg_stride, const float* A, const float* abias, float again, float sgain,
                  const float* Wsq, float* s_out, float* d_out, int B, int Cin, int O, int wdim, void* stream);
int mgf_style_bwd(const float* ds, const float* R, const float* s, const float* d, const float* Wsq, const float* A,
                  float again, float sgain, float* dwg, int64_t dwg_stride, int B, int Cin, int O, int wdim, void* stream);
int mgf_modulate_weights(const float* base, const float* rs, int nmod, const float* cs, void* out, int out_fwd,
                         int B, int64_t T, int64_t NT, int64_t K, void* stream);
/* out[b,p,c] = x[b,p,c] * sc[b,c] on an NHWC 16-bit tensor (is_fwd: forward-dtype tensor, else bf16 gradient) */
int mgf_scale_channels(const void* x, const float* sc, void* out, int is_fwd, int B, int64_t HW, int C, void* stream);
int mgf_small_gemm(const float* A, int64_t sAb, int64_t sAm, const float* Bm, const float* bias, float* out,
                   int64_t sOb, int64_t sOm, int B, int M, int N, int K, int accumulate, void* stream);
int mgf_torgb_fwd(const void* y, const float* wrgb, const float* s, const float* bias, float* img, int B, int64_t HW, int C, void* stream);
int mgf_torgb_bwd(const float* dimg, const void* y, const float* wrgb, const float* s, void* dy, float* ds, float* R,
                  int B, int64_t HW, int C, void* stream);
int mgf_act_bwd(const void* dz, const void* z, void* dy, float* R, const float* noise, const float* nstr, const float* bias,
                float alpha, float gain, int mode, int B, int64_t HW, int C, int64_t noise_bstride, void* stream);
/* g [B,H+2,W+2,C] bf16 = 4-tap separable FIR (taps fk4, zero padding 2, * gain) of dy [B,H,W,C] bf16: first stage of the up-convolution's
 * input gradient (adjoint of upfirdn2d(pad 1, gain 4) after conv_transpose2d, conv2d_resample.py:117-134) */
/* same-resolution 4x4 separable FIR with zero padding on an NHWC 16-bit tensor (fir.cu): out[b,Y,X,c] = gain * sum_{t,u} fk4[t] fk4[u]
 * in[b,Y+off+t,X+off+u,c]; outputs outside the valid (Hv, Wv) window of the allocated (Ho, Wo) tensor are written as zeros.
 * off = -1: second stage of the up-convolution (conv_transpose2d(stride 2) -> upfirdn2d(pad 1, gain 4), conv2d_resample.py:117-134) with
 * the layer tail fused when noise / bias / act are given: (v + noise * nstr + bias) -> leaky-ReLU(alpha) (act == 1) -> * act_gain
 * (networks.py:1036-1040); in_fwd / out_fwd select forward-dtype tensors (else bf16 gradients).  off = -2: the adjoint (mgf_fir4_pad). */
int mgf_fir4(const void* in, void* out, const float* fk4, float gain, int off, int B, int Hi, int Wi, int Ho, int Wo, int Hv, int Wv,
             int C, int in_fwd, int out_fwd, const float* noise, const float* nstr, int64_t noise_bstride, const float* bias, int act,
             float alpha, float act_gain, void* stream);
int mgf_fir4_pad(const void* dy, void* g, const float* fk4, float gain, int B, int H, int W, int C, void* stream);
int mgf_upfir2_add(const void* v, const void* add, void* out, const float* fk4, float gain, int B, int h, int w, int C, void* stream);
/* 1x1 convolution with shared weights (the resnet-skip Conv2dLayer, reference training/networks.py:245-250 with kernel_size 1 as called from
 * SynthesisBlock.forward :1157-1160, and its input gradient): out[p, n] = sum_k x[p, k] w[n, k] over P pixels of NHWC 16-bit tensors;
 * w [N, K] is the layout of a [1, 1, N, K] mgf_conv_tc weight tensor.  is_fwd != 0: forward tensors (forward dtype, fp16 stores tracked by the
 * overflow flag); 0: bf16 gradients.  K in {32, 64, 128, 256}, N a multiple of 32 (mgf_pointwise_supported tells); else MGF_E_UNSUP. */
int mgf_pointwise(const void* x, const void* w, void* out, int64_t P, int K, int N, int is_fwd, void* stream);
int mgf_pointwise_supported(int K, int N);
int mgf_upfir2_bwd(const void* dout, void* dv, const float* fk4, float gain, int B, int h, int w, int C, void* stream);
/* A/B switch of the three FIR walkers above: bit 0 mgf_fir4 / mgf_fir4_pad, bit 1 mgf_upfir2_add, bit 2 mgf_upfir2_bwd run with 4 channels per thread
 * instead of 8 (half the registers per thread, twice the resident warps; identical results).  Default 2 (upfir2_add only: the one that gains).
 * Process-wide, not thread-safe: set once at start-up. */
int mgf_fir_set_mode(int bits);
int mgf_fir_get_mode(void);

/* ---- fused duplex attention layer (attention.cu): TransformerLayer.forward (networks.py:748-822, default GANformer config)
 * + noise + bias_act tail (:1036-1040) in one pass over X [B,HW,C] bf16; Kf [16,C], Sc [HW,16], maskbias [B,16], VM [B,16,C],
 * bm [C] are the host-folded constants described in attention.cu.  bwd writes dX, accumulates dVM [B,16,C] and R [B,C].
 * dmask (optional, [B,HW,16] fp32): attention dropout of training mode (networks.py:505-513) = keep-masks of the cell and column dropouts
 * times their 1/(1-p) scales; multiplies the probabilities after the softmax (probs, when requested, are the dropped ones, as in the reference). */
int mgf_attn_fwd(const void* X, const float* Kf, const float* Sc, const float* maskbias, const float* VM, const float* bm,
                 const float* noise, const float* nstr, const float* bias, float gain, float alpha,
                 void* out, float* probs, const float* dmask, const void* tabK, const void* tabV,
                 int B, int64_t HW, int C, int64_t noise_bstride, void* stream);
int mgf_attn_bwd(const void* X, const void* dz, const float* Kf, const float* Sc, const float* maskbias, const float* VM, const float* bm,
                 const float* noise, const float* nstr, const float* bias, float gain, float alpha,
                 void* dX, float* dVM, float* R, const float* dmask, const void* tabK, const void* tabV,
                 int B, int64_t HW, int C, int64_t noise_bstride, void* stream);
/* tabK / tabV (optional): the 16-bit coefficient tables of Kf / VM in the kernels' shared-memory layout, built once by mgf_attn_tables
 * (Kf: when the weights are folded; VM [B,16,C]: per step) instead of by every CTA of every launch.  mgf_attn_table_bytes(0 | 1, C) = bytes
 * of tabK / of tabV per sample.  The tables depend on the forward dtype in effect when they were built. */
int64_t mgf_attn_table_bytes(int which, int C);
int mgf_attn_tables(const float* Kf, const float* VM, void* tabK, void* tabV, int B, int C, void* stream);
/* A/B measurement switch (process-wide; default 1): 0 keeps the one-warp-per-pixel-tile backward kernel for every layer instead of the
 * split-channel kernel (several warps per tile) that mgf_attn_bwd picks for the low-resolution 512-channel layers. */
int mgf_attn_set_split(int enabled);

/* ---- mapping network z -> ws and its backward wrt z (mapping.cu): training/networks.py MappingNetwork.forward :894-942 with the
 * GANformer-default configuration (16 local + 1 global latents x 32, 4 resnet blocks, latent self-attention, positional maps).
 * params: packed fp32 weights, mgf_mapping_param_floats() floats (layout: mapping.cu; packed by morphganformer_b200.mapping_engine). */
int mgf_mapping_param_floats(void);
int mgf_mapping_fwd(const float* z, const float* params, const float* maskbias, float* ws, int B, int num_ws, void* stream);
int mgf_mapping_bwd(const float* z, const float* params, const float* maskbias, const float* dws, float* dz, int B, int num_ws, void* stream);

/* ---- projection-loss and optimizer kernels (lpips.cu): lpips/networks_basic.py:64-101, lpips/__init__.py:44-46,
 * MSELoss (1024_example_percept_MSE.py:143), Adam + latent noise (:117, :134-135, :153). */
int mgf_lpips_prep(const float* img, const float* target, void* col, float* mse, int B, int R, void* stream);
int mgf_lpips_prep_bwd(const void* dcol, const float* img, const float* target, float mcoef, float* dimg, int B, int R, void* stream);
/* LPIPS input stage fused with VGG conv1_1 (vgg_first.cu): ScalingLayer (networks_basic.py:94-101) + conv3x3(3->64) + bias + ReLU
 * (pretrained_networks.py:97-135, slice1 layers 0-1) straight from the fp32 NCHW image, and its backward to the image.
 * W [64][32] fp32, column (ky*3+kx)*3+c (27..31 unused); out [B,R,R,64] in the forward 16-bit type; gy [B,R,R,64] bf16. */
int mgf_vgg_conv1_fwd(const float* img, const float* target, float* mse, const float* W, const float* bias, void* out,
                      int B, int R, void* stream);
int mgf_vgg_conv1_bwd(const void* gy, const float* W, const float* img, const float* target, float mcoef, float* dimg,
                      int B, int R, void* stream);
int mgf_maxpool2_fwd(const void* x, void* y, int B, int H, int W, int C, void* stream);
int mgf_maxpool2_bwd(const void* x, const void* dy, const void* extra, void* dx, int B, int H, int W, int C, void* stream);
int mgf_lpips_head(int mode, const void* f, const void* n1, const float* lin, const float* coef, void* out, float* val,
                   int relu_mask, int B, int64_t HW, int C, void* stream);
/* fused forward of an LPIPS tap followed by a 2x2 max-pool: val[b] += head distance of x [B,H,W,C] vs n1 (as mgf_lpips_head mode 1) and
 * y [B,H/2,W/2,C] = maxpool2(x), one pass over x */
/* stats (optional, [B,H,W] float2): per-pixel (|f|, g . f) written by the forward kernel and consumed by the backward kernel below, which then
 * skips its own channel-reduction pass */
int mgf_lpips_tap_pool_fwd(const void* x, const void* n1, const float* lin, void* y, float* val, void* stats, int B, int H, int W, int C, void* stream);
/* fused backward of an LPIPS tap followed by a 2x2 max-pool: dx = (route(dy) + d(head)/dx) * (x > 0); x, n1 forward tensors, dy/dx bf16 */
int mgf_lpips_tap_pool_bwd(const void* x, const void* n1, const float* lin, const float* coef, const void* dy, void* dx,
                           const void* stats, int B, int H, int W, int C, void* stream);
int mgf_adam_noise_step(float* latent, const float* grad, float* m, float* v, const float* noise_all, int noise_rows, float* latent_n,
                        const float* sched, int* step_ptr, float beta1, float beta2, float eps, float weight_decay, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MGF_H_ */
