/* libwgs_b200 — C ABI of the B200-native WarpedGANSpace hot path.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the parameter name starts with `h_`;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); no call synchronises the
 *     host unless documented;
 *   - return value 0 = success, non-zero = failure, message via wgs_last_error();
 *   - tensors are dense row-major; activations are NHWC ("channels last") fp32 unless stated.
 *
 * Each entry point cites the reference interface it replaces (paths under chi0tzp/WarpedGANSpace).
 */
#ifndef WGS_B200_H
#define WGS_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library ---------------------------------------------------------------------------------- */
const char* wgs_last_error(void);
int  wgs_version(void);
unsigned long long wgs_launch_count(void);       /* kernels launched through this library so far */
void wgs_reset_launch_count(void);
int  wgs_device_info(int* sms, int* cc);         /* fails unless the current device is sm_100 */

/* ---- SupportSets RBF warp --------------------------------------------------------------------- *
 * Replaces SupportSets.forward, lib/support_sets.py:81-101 (and its autograd backward), with the
 * one-hot mask given as row indices:  out[b] = mag[b] * grad_f(z[b]) / ||grad_f(z[b])||,
 *   grad_f = -2 * sum_j alpha[k,j] * gamma_k * exp(-gamma_k * |z - s_kj|^2) * (z - s_kj),  k = idx[b].
 * support_sets [K, n_vec*d] (n_vec = 2 * num_support_dipoles), alphas [K, n_vec], loggamma [K] or NULL
 * (then gamma = fixed_gamma, the learn_gammas=False branch :93), idx [B] int64, z [B, d],
 * mag [B] or NULL (= 1, the bare module output), out [B, d].  d <= 1024 (128-bit row loads when d % 4 == 0,
 * guarded scalar loads otherwise: BigGAN-256 has dim_z = 119).                                     */
int wgs_rbf_warp_forward(const float* support_sets, const float* alphas, const float* loggamma,
                         float fixed_gamma, const long long* idx, const float* z, const float* mag,
                         float* out, int B, int K, int n_vec, int d, void* stream);

/* Backward of the above for upstream gradient dout [B, d].  d_support_sets [K, n_vec*d] and
 * d_loggamma [K] / d_alphas [K, n_vec] are ACCUMULATED into (atomics; zero them first; only the
 * rows named by idx are touched, as in the reference's one-hot matmul backward); dz [B, d] is
 * overwritten.  Any of the four outputs may be NULL.                                             */
int wgs_rbf_warp_backward(const float* support_sets, const float* alphas, const float* loggamma,
                          float fixed_gamma, const long long* idx, const float* z, const float* mag,
                          const float* dout, float* d_support_sets, float* d_loggamma, float* d_alphas,
                          float* dz, int B, int K, int n_vec, int d, void* stream);

/* Traversal chains, traverse_latent_space.py:369-438: for each chain c, `steps` sequential steps
 * shift = +-eps * S(path[c], code); code += shift, in both directions from start[c].
 * codes, shifts: [chains, 2*steps+1, d], most negative step first, centre frame = (start, 0).     */
int wgs_rbf_traverse(const float* support_sets, const float* alphas, const float* loggamma,
                     float fixed_gamma, const long long* path, const float* start, float eps, int steps,
                     float* codes, float* shifts, int chains, int K, int n_vec, int d, void* stream);

/* ---- op-level boundary: the reference's two native extensions ------------------------------------- *
 * wgs_fused_bias_act replaces `fused.fused_bias_act(input, bias, refer, act, grad, alpha, scale)`
 * (models/StyleGAN2/op/fused_bias_act.cpp:11-20; kernel op/fused_bias_act_kernel.cu:18-49):
 *   out[i] = f(x[i] + b[(i / step_b) % size_b]) * scale,  f selected by act (1 linear, 3 leaky ReLU with slope alpha)
 *   and grad (0 value; 1 first derivative taken at the sign of ref[i], the forward OUTPUT; 2 second derivative = 0).
 * b may be NULL (no bias), ref may be NULL when grad == 0.  n elements; step_b = product of the dims after the channel
 * dim, size_b = channels (op/fused_bias_act_kernel.cu:66-71).
 * wgs_upfirdn2d replaces `upfirdn2d_op.upfirdn2d(input, kernel, up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0,
 * pad_y1)` (op/upfirdn2d.cpp:12-22; kernel op/upfirdn2d_kernel.cu:52-272): in [major, in_h, in_w, minor] ->
 * out [major, out_h, out_w, minor], out_h = (in_h*up_y + pad_y0 + pad_y1 - kh) / down_y + 1 (same for w); zero-insert
 * up-sampling, padding (negative = crop), correlation with the FLIPPED kernel [kh, kw], decimation.  Any up / down /
 * kernel size (the reference silently reads an uninitialised tile size outside its six modes).   */
int wgs_fused_bias_act(const float* x, const float* b, const float* ref, float* out, long long n, long long step_b,
                       int size_b, int act, int grad, float alpha, float scale, void* stream);
int wgs_upfirdn2d(const float* in, const float* kernel, float* out, int major, int in_h, int in_w, int minor,
                  int kh, int kw, int up_x, int up_y, int down_x, int down_y, int pad_x0, int pad_x1, int pad_y0,
                  int pad_y1, void* stream);

/* ---- ProgGAN glue ----------------------------------------------------------------------------------- *
 * PixelNormLayer (models/ProgGAN/model.py:17-18) fused with the operand pack of the conv that follows it:
 *   a [R, C] fp32 (NHWC rows) -> split32( a / sqrt(mean_c a^2 + eps) ) [R, ceil(C/32), 64] and / or fp32 [R, C].
 * C % 4 == 0, C <= 512.                                                                              */
int wgs_pixelnorm_pack(const float* a, long long R, int C, float eps, void* out_split, float* out_f32, void* stream);
/* Backward of the above, fused with the LeakyReLU backward of the block that produced `a` (models/ProgGAN/model.py:48,62)
 * and the pack of the next data-gradient conv's operand:  g = r (dxn - xn mean_c(dxn xn)),  xn = a r,
 * r = rsqrt(mean_c a^2 + eps);  if lrelu_slope >= 0: g *= (a > 0 ? 1 : lrelu_slope).  Writes split32(g) and / or fp32 g. */
int wgs_pixelnorm_bwd_pack(const float* dxn, const float* a, long long R, int C, float eps, float lrelu_slope,
                           void* out_split, float* out_f32, void* stream);

/* ---- BigGAN glue: class-conditional BatchNorm (eval) + ReLU + operand pack ------------------------------ *
 * models/BigGAN/layers.py:303-322 (ccbn) and :395-405 (GBlock): with the generator frozen in eval mode, ccbn is the
 * per-sample affine map o[n,p,c] = A[n,c] * x[n,p,c] + B[n,c] (A = gain/sqrt(var+eps), B = bias - mean*A, formed by the host).
 * wgs_affine_act_pack: x [R, C] fp32 (R = groups * rows_per_group NHWC rows) -> [relu](A x + B) as split32 and / or fp32.
 * A, B [groups, C] or both NULL (plain [relu +] pack).  The same kernel serves per-channel eval BatchNorm (groups = 1). */
int wgs_affine_act_pack(const float* x, const float* A, const float* B, long long R, int C, long long rows_per_group,
                        int relu, void* out_split, float* out_f32, void* stream);
/* Backward: dpre = dz * [A x + B > 0]; dx = dpre * A written as split32 (operand of the next data-gradient conv) and / or
 * fp32; dA[n,c] += sum_p dpre * x, dB[n,c] += sum_p dpre (atomics; zero them first; NULL to skip).                      */
int wgs_affine_act_bwd(const float* dz, const float* x, const float* A, const float* B, int N, long long rows_per_group,
                       int C, int relu, void* dx_split, float* dx_f32, float* dA, float* dB, void* stream);

/* ---- output stage: GPU JPEG encode (nvJPEG, resolved with dlopen at first use) ------------------------- *
 * Replaces the host-side PIL encode of traverse_latent_space.py:466-483 / sample_gan.py:172-176 (quality 95, optimised
 * Huffman tables, progressive): pixels = uint8 [N, H, W, C] on the DEVICE (C = 3 interleaved RGB or 1 gray, the output of
 * wgs_image_to_u8); image i's bitstream is written to the HOST buffer h_out + i * h_capacity_per_image and its length to
 * h_sizes[i].  Synchronises the stream (the bitstream is returned to the host).  wgs_jpeg_available() = 1 when nvJPEG loads. */
int wgs_jpeg_available(void);
int wgs_jpeg_encode(const unsigned char* pixels, int N, int H, int W, int C, int quality, int progressive,
                    unsigned char* h_out, long long h_capacity_per_image, long long* h_sizes, void* stream);

/* ---- tensor-core convolution -------------------------------------------------------------------- *
 * "split32" operand format: every 32 fp32 channels become one 128-byte row of 64 bf16 —
 * [hi(32) | lo(32)], x = hi + lo — so a tensor of C channels (padded to a multiple of 32) is
 * [..., C/32, 64] bf16 and has the same byte size as its fp32 source.
 *
 * wgs_pack_split32: rows x C fp32 (row stride `ld` floats) -> [rows, ceil(C/32), 64] split32, each
 * element optionally multiplied by scale[(row / rows_per_group) * scale_ld + c] (per-sample per-channel
 * modulation; scale may be NULL) — replaces `weight = scale * W * style` of
 * models/StyleGAN2/model.py:190-191 by scaling the activations instead of the weights.            */
int wgs_pack_split32(const float* src, long long rows, int C, long long ld, const float* scale,
                     long long scale_ld, long long rows_per_group, void* dst, void* stream);

/* Stacked weight layout for narrow layers (see wgs_conv_desc.w_layout): src fp32 [taps*cout, C] rows (stride ld). */
int wgs_pack_weights_stacked(const float* src, int taps, int cout, int C, long long ld, void* dst, void* stream);

/* Grouped weight pack: every conv weight of a network in one launch, gathered from the torch [Co, Ci, kh, kw] layout.
 *   mode 0  forward taps          dst [kh*kw][Co][Ci]                 (tap = ky*kw + kx)
 *   mode 1  transposed taps       dst [kh*kw][Ci][Co]                 (data gradient of a stride-1 conv; taps flipped by the tap list)
 *   mode 2  im2col matrix         dst [1][Co][kh*kw*Ci], K = (ky, kx, c) (few-input-channel stem, wgs_im2col_split32)
 *   mode 3  phase-merged dgrad    dst [S][G*Ci][Co], block (s, g) = tap idx[s*G+g] transposed, idx < 0 = zero block
 *                                 (all stride^2 output phases of a strided data gradient stacked along N, wgs_conv_desc.group_size)
 *   mode 4  space-to-depth taps   dst [S*S][Co][4*Ci]: a stride-2 kh x kw conv as a stride-1 S x S conv over the 2x2
 *                                 space-to-depth input (wgs_s2d_pack_split32): k = (py*2+px)*Ci + c, tap (ty, tx) holds kernel
 *                                 element (2*ty + py - G, 2*tx + px - G) or zero; S = taps per axis, G = kernel offset
 * layout 0 = rows [T][rows][chunks][hi32 | lo32], 1 = stacked [T][chunks][hi | lo][rows][32] (rows <= 64).
 * h_problems is a HOST array (copied into the launch parameters).                                                     */
#define WGS_PACK_GROUP_MAX 24
typedef struct {
    const float* src;
    void* dst;
    int co, ci, kh, kw;
    int mode, layout, S, G;
    signed char idx[64];
    int ci_src, ci_off;   /* mode 3 only: pack input channels [ci_off, ci_off + ci) of a [Co, ci_src, kh, kw] weight (0 = all of ci) */
} wgs_pack_problem;
int wgs_pack_weights_group(const wgs_pack_problem* h_problems, int count, void* stream);
int wgs_pack_problem_size(void);
/* 2x2 space-to-depth + split32 pack: x fp32 NHWC [N,H,W,C] (H, W even) -> [N, H/2, W/2, ceil(4C/32), 64], channel
 * (py*2+px)*C + c = x[2Y+py, 2X+px, c] - the operand of a few-input-channel stride-2 conv run as a stride-1 conv
 * (ResNet stem 7x7/2 on 6 channels, lib/reconstructor.py:56-60 -> 4x4 taps on 24 channels).                        */
int wgs_s2d_pack_split32(const float* x, int N, int H, int W, int C, void* out, void* stream);
/* The same over the channel concatenation [x1 (C1 channels) ; x2 (C2 channels)] without materialising it: the Reconstructor's
 * torch.cat([x1, x2], dim=1) (lib/reconstructor.py:72) folded into the operand pack of its stem.                     */
int wgs_s2d_pack_split32_pair(const float* x1, const float* x2, int N, int H, int W, int C1, int C2, void* out, void* stream);


#define WGS_MAX_TAPS 64
typedef struct wgs_conv_desc {
    /* input activations, split32 [in_n][in_h][in_w][c_chunks][64] */
    const void* in;
    int in_n, in_h, in_w, c_chunks;
    /* weights, split32 [w_taps][w_cout][c_chunks][64] */
    const void* w;
    int w_taps, w_cout;
    /* virtual output grid per image; input pixel = grid coord * in_stride + tap offset (zero outside) */
    int out_n, grid_h, grid_w, in_stride;
    int num_taps;
    int tap_dy[WGS_MAX_TAPS], tap_dx[WGS_MAX_TAPS], tap_w[WGS_MAX_TAPS];
    /* output fp32: &out[n*out_sn + (oy*out_ystep+out_y0)*out_sy + (ox*out_xstep+out_x0)*out_sx + co] */
    float* out;
    long long out_sn, out_sy, out_sx;
    int out_y0, out_x0, out_ystep, out_xstep;
    int cout;
    const float* alpha;      /* [out_n, cout] per-sample per-channel scale (demodulation) or NULL */
    const float* beta;       /* [cout] bias or NULL */
    int act;                 /* 0 none, 1 relu, 2 leaky relu 0.2, 3 sqrt(2)*leaky relu 0.2 (FusedLeakyReLU), 4 tanh */
    int accumulate;          /* add to the existing output instead of overwriting */
    int force_bn;            /* 0 = choose the channel tile automatically */
    const float* noise;      /* per-pixel noise plane indexed by OUTPUT pixel [y*noise_ld + x], or NULL  */
    float noise_w;           /* NoiseInjection weight (models/StyleGAN2/model.py:231-241)                */
    int noise_ld;
    /* fused consumers of the activation (identity output mapping only; `out` may then be NULL):            */
    void* out_split;         /* split32 [out_n][grid_h][grid_w][cout/32][64] of act * split_scale[n,co]     */
    const float* split_scale;/* [out_n, cout] with row stride split_scale_ld, or NULL (next layer's style)  */
    long long split_scale_ld;
    int out_from_n;          /* fp32 `out` only for images n >= out_from_n (the half that is back-propagated) */
    const float* rgb_w;      /* [out_n][3][cout] modulated ToRGB weights (model.py:270-282) or NULL          */
    float* rgb_out;          /* [out_n][grid_h][grid_w][3] += act . rgb_w  (pre-initialised: bias + skip)    */
    /* phase-packed output (all output phases of a strided data-gradient / transposed conv in ONE launch): when
     * group_size > 0 the cout channels are cout/group_size groups of group_size channels; group g = gy*group_w + gx
     * is written to output pixel (oy*out_ystep+out_y0+gy, ox*out_xstep+out_x0+gx), channel co % group_size, and only
     * if that pixel lies inside out_h x out_w.  The weights hold one row block per group (zero where a phase has no
     * tap at a given input shift).  fp32 output only (no out_split / rgb_out / noise).                          */
    int group_size, group_w, out_h, out_w;
    /* weight layout: 0 = rows [w_taps][w_cout][c_chunks][hi32|lo32] (wgs_pack_split32);
     * 1 = stacked [w_taps][c_chunks][2][w_cout][32] (wgs_pack_weights_stacked; hi plane then lo plane, 64-byte rows),
     * w_cout <= 64 only: hi*hi and hi*lo then come out of ONE N = 2*BN MMA (the A operand is fetched twice per
     * K slice instead of three times; narrow-N MMAs are bound by that fetch).                                   */
    int w_layout;
    /* 1 = this launch may split its contraction over a thread-block cluster (tiny-M, deep-K layers: the Reconstructor's
     * layer 3 / 4 convs).  The split factor depends on the number of output tiles, i.e. on the batch size, so callers that
     * need results independent of how images are batched (the generator: G(z) inside a pair batch == G(z) alone, bit for
     * bit) leave it 0 or pass 2 = split by GEOMETRY only (feature maps up to 16 x 16; the factor is a function of the map
     * size, channel and tap counts, never of the batch).  WGS_CONV_SPLITK=1 / 0 in the environment forces splitting on / off. */
    int split_k;
    /* Train-mode BatchNorm statistics of the OUTPUT formed in the epilogue (the Reconstructor's convs, lib/reconstructor.py:54
     * -> torchvision BasicBlock): stat_sum[c] += sum_pixels (out - shift[c]), stat_sumsq[c] += sum_pixels (out - shift[c])^2
     * (atomics; zero them first; stat_shift may be NULL = 0).  Replaces a separate pass over the output (wgs_bn_stats).  Plain
     * fp32 output only (no fused / phase-packed / accumulate epilogue), cout % 16 == 0, output maps of >= 128 pixels.      */
    float* stat_sum;
    float* stat_sumsq;
    const float* stat_shift;
    /* > 0: the split32 output holds pixel_norm(act) = act * rsqrt(mean_c act^2 + pixnorm_eps) instead of act (ProgGAN's
     * PixelNormLayer, models/ProgGAN/model.py:17-18, fused with the conv that feeds it and the pack of the conv that follows):
     * cout <= 256 (one channel tile holds every channel of a pixel), output maps of >= 128 pixels.  The fp32 output stays the
     * un-normalised activation (the backward pass needs it).  With a strided output mapping (output-phase launches of an
     * up-sampling conv) out_split is addressed by output pixel and out_h / out_w give its dims.                           */
    float pixnorm_eps;
} wgs_conv_desc;

/* One implicit-GEMM convolution on tcgen05 tensor cores (see csrc/conv.cu).  Replaces the cuDNN calls
 * behind F.conv2d / F.conv_transpose2d at models/StyleGAN2/model.py:206,219,225,
 * models/ProgGAN/model.py:39,55, models/BigGAN/layers.py:105, models/SNGAN/sn_gen_resnet.py:28-29 and
 * torchvision resnet18 (lib/reconstructor.py:54), forward and data-gradient.                       */
int wgs_conv_split32(const wgs_conv_desc* desc, void* stream);
int wgs_conv_desc_size(void);

/* ---- output stage ---------------------------------------------------------------------------------- *
 * tensor2image (traverse_latent_space.py:26-41, sample_gan.py:10-25) on the device: n images of `count` fp32 values each
 * (any layout; the generators here produce NHWC, which is what PIL wants) -> uint8.  adaptive != 0: per-image
 * (x - min) / (max - min), else (x + 1) / 2; then uint8(255 * t) by truncation - the reference's fp32 operations in
 * the reference's order, so the pixels are bit-identical.  workspace: 2 * workspace_pairs floats of scratch
 * (per-block min/max), workspace_pairs >= n.                                                                    */
int wgs_image_to_u8(const float* images, int n, long long count, int adaptive, float* workspace, int workspace_pairs,
                    unsigned char* out, void* stream);

/* ---- StyleGAN2 glue (CUDA cores, HBM/latency-bound) --------------------------------------------- *
 * wgs_linear_small: out[b,o] = epi(wscale * sum_i f(x[b,i]) * W[o,i] + bscale * bias[o]); f = square when
 * in_square; epi 0 linear, 1 sqrt(2)*lrelu(0.2) (EqualLinear 'fused_lrelu', models/StyleGAN2/model.py:110-131),
 * 2 rsqrt(wscale*sum + eps) (demodulation :194-195 restated on squared styles). Row strides in floats.   */
int wgs_linear_small(const float* x, long long x_ld, const float* W, long long w_ld, const float* bias,
                     float* out, long long out_ld, int B, int I, int O, float wscale, float bscale,
                     int in_square, int epi, float eps, int accumulate, void* stream);
/* Grouped form: ONE launch over `count` independent small linears of the same batch size B (all 17 demodulation
 * vectors of a StyleGAN2 forward, their backward, the mapping-network backward with the fused-lrelu derivative folded
 * into the input).  out[b,o] (+)= mul[b,o] * epi(wscale * sum_i f(x[b,i], x2[b,i]) * W[o,i] + bscale * bias[o]);
 * in_mode 0: x, 1: x^2, 2: x * dlrelu(x2) (x2 = forward output of models/StyleGAN2/model.py:127-128),
 * 3: x * x2^3 (derivative of the rsqrt in :194-195).  x2, bias, mul may be NULL where unused.                  */
#define WGS_MAX_LINEAR_GROUP 32
typedef struct {
    const float* x;   long long x_ld;
    const float* x2;  long long x2_ld;
    const float* W;   long long w_ld;
    const float* bias;
    const float* mul; long long mul_ld;
    float* out;       long long out_ld;
    int I, O;
    float wscale, bscale, eps;
    int in_mode, epi, accumulate;   /* accumulate: 0 overwrite, 1 add to the output, 2 atomic add (problems of one launch sharing an output) */
} wgs_linear_problem;
int wgs_linear_group(const wgs_linear_problem* problems, int count, int B, void* stream);
int wgs_linear_problem_size(void);
/* A chain of L equal-width linears (d -> d) in one launch on a cluster of 8 CTAs (csrc/mlp.cu): the mapping network
 * (8 x EqualLinear + fused leaky-ReLU, model.py:110-131,291-295) forward, and its backward pass.
 *   out_l[b,o] = epi( wscale * sum_i f(in[b,i], aux_l[b,i]) W_l[o,i] + bscale * bias_l[o] ),  in = out_{l-1}, in_0 = x
 * in_mode 0: f = in; 2: f = in * lrelu'(aux) (aux = that layer's FORWARD output, [B, d] contiguous).  epi 0 linear, 1 sqrt(2)*lrelu_0.2.
 * W_l [d, d] row-major (row = output feature), bias_l [d] or NULL, out_l [B, d] contiguous or NULL (intermediate not kept).
 * h_layers is a HOST array.  d % 128 == 0, 2 * B * d * 4 bytes of shared memory (B <= 48 at d = 512).               */
#define WGS_MLP_MAX_LAYERS 16
typedef struct { const float* W; const float* bias; const float* aux; float* out; } wgs_mlp_layer;
int wgs_mlp_chain(const float* x, long long x_ld, const wgs_mlp_layer* h_layers, int L, int B, int d, float wscale,
                  float bscale, int in_mode, int epi, void* stream);
int wgs_mlp_layer_size(void);
/* PixelNorm over latent rows (model.py:9-15). */
int wgs_pixelnorm_rows(const float* x, float* out, int B, int d, void* stream);
/* Separable 4-tap FIR (upfirdn2d up=down=1: Blur, model.py:66-81; op/upfirdn2d_kernel.cu:52-137) fused with
 * demodulation scale alpha[n,c], NoiseInjection (:231-241) and FusedLeakyReLU (op/fused_bias_act_kernel.cu).
 * y [N,Hin,Win,C] -> out [N,Hout,Wout,C] (fp32, images n >= out_from_n only; may be NULL) and/or out_split
 * (split32 of the result times split_scale[n,c]); h_taps4 is a HOST pointer to the four 1-D taps.           */
int wgs_fir4_act(const float* y, float* out, int N, int Hin, int Win, int Hout, int Wout, int C, int pad0,
                 const float* h_taps4, const float* alpha, const float* beta, const float* noise,
                 float noise_w, int act, void* out_split, const float* split_scale, long long split_scale_ld,
                 int out_from_n, void* stream);
/* ToRGB accumulator init (bias + FIR-upsampled skip) and modulated ToRGB weights for the fused conv epilogue. */
int wgs_sg2_rgb_init(const float* bias, const float* prev, float* rgb, int N, int H, int W,
                     const float* h_taps4, void* stream);
int wgs_sg2_rgb_weights(const float* W, const float* s, long long s_ld, float* wm, int N, int C, float wscale,
                        void* stream);
/* ToRGB (model.py:270-282): 1x1 modulated conv (no demod) + bias + FIR-upsampled skip (Upsample :29-45).
 * a [N,H,W,C], s [N,C] (row stride s_ld), W [3,C], bias [3], prev [N,H/2,W/2,3] or NULL, rgb [N,H,W,3].  */
int wgs_sg2_torgb(const float* a, const float* s, long long s_ld, const float* W, const float* bias,
                  const float* prev, float* rgb, int N, int H, int Wd, int C, float wscale,
                  const float* h_taps4, void* stream);

/* Weight gradient of F.conv2d(x, w, stride, padding) on tensor cores (csrc/wgrad.cu):
 * xs split32 [n,h,w,ci_chunks,64], dys split32 [n,oh,ow,co_chunks,64] ->
 * dw fp32, ACCUMULATED (zero it first): layout 0 = [kh*kw][co_chunks*32][ci_chunks*32]; layout 1 = torch
 * [out_co][out_ci][kh][kw] (e.g. straight into the parameter's .grad).  Replaces the cuDNN wgrad reached from
 * loss.backward(), lib/trainer.py:250.                                                              */
int wgs_conv_wgrad_split32(const void* xs, int n, int h, int w, int ci_chunks, const void* dys, int oh,
                           int ow, int co_chunks, int kh, int kw, int stride, int pad, float* dw,
                           int layout, int out_co, int out_ci, void* stream);

/* ---- StyleGAN2 data-gradient glue (csrc/sg2_bwd.cu) — replaces autograd through
 * models/StyleGAN2/model.py:187-282 for the frozen generator (no weight gradients are formed).
 * P = pixels per image, tensors [N, P, C]; reductions are accumulated (zero the outputs first).     */
int wgs_sg2_act_bwd(const float* da, const float* a, const float* demod, const float* bias,
                    const float* noise, float noise_w, float* dpre, float* dd, void* g_split, int N,
                    long long P, int C, void* stream);
int wgs_sg2_mod_bwd(const float* dx, const float* a_prev, int a_bcast, const float* s, long long s_ld,
                    float* da_prev, int accumulate, float* ds, long long ds_ld, int N, long long P, int C,
                    void* stream);
int wgs_sg2_torgb_bwd(const float* drgb, const float* a, const float* s, long long s_ld, const float* W,
                      float wscale, float* da, int accumulate, float* ds, long long ds_ld, int N,
                      long long P, int C, void* stream);
int wgs_sg2_rgb_up_bwd(const float* drgb, float* dprev, int N, int H, int W, const float* h_taps4,
                       void* stream);
/* Fused layer boundary of the data-gradient chain (mod_bwd of the layer above + torgb_bwd + act_bwd of this layer):
 * every tensor is read once.  dx_up / drgb are optional (NULL) but not both.                                    */
int wgs_sg2_layer_bwd(const float* dx_up, const float* s_up, long long s_up_ld, float* ds_up, long long ds_up_ld,
                      const float* drgb, const float* s_rgb, long long s_rgb_ld, const float* Wrgb, float wscale,
                      float* ds_rgb, long long ds_rgb_ld, const float* a, const float* demod, const float* bias,
                      const float* noise, float noise_w, float* dd, float* dpre, void* g_split, int N,
                      long long P, int C, void* stream);

/* Fused Adam over a flat fp32 buffer, torch.optim.Adam defaults semantics (lib/trainer.py:153-156,253-254);
 * the gradient is multiplied by grad_scale first (1/world_size after the NCCL sum).                  */
int wgs_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2,
                  float eps, int step, float grad_scale, const int* step_dev, void* stream);
/* step_dev (device int) replaces `step` when non-NULL so that the launch can be replayed from a CUDA graph. */
int wgs_step_increment(int* step_dev, void* stream);

/* im2col into split32 for few-input-channel convs (ResNet stem 7x7/2 on 6 channels, lib/reconstructor.py:56-60):
 * x fp32 NHWC [N,H,W,C] -> out split32 [N,OH,OW,ceil(kh*kw*C/32),64], K index = (ky*kw+kx)*C + c.             */
int wgs_im2col_split32(const float* x, int N, int H, int W, int C, int kh, int kw, int stride, int pad,
                       int OH, int OW, void* out, void* stream);

/* ---- Reconstructor BatchNorm (train mode) fused with residual / ReLU / operand packing (csrc/bn.cu) -------------- *
 * Replaces cuDNN batch_norm fwd/bwd + ATen relu / add around every torchvision BasicBlock conv
 * (lib/reconstructor.py:54-69; train mode per lib/trainer.py:150).  Tensors are NHWC [R = N*H*W, C] fp32.         */
int wgs_bn_stats(const float* y, long long R, int C, float* sum, float* sumsq, void* stream);   /* accumulates sums of (y - y[0,c]) and (y - y[0,c])^2: shifted statistics, shift = first row */
int wgs_bn_finalize(const float* sum, const float* sumsq, const float* shift, long long R, int C, float eps,
                    float momentum, float* mean, float* rstd, float* running_mean, float* running_var, void* stream);
/* bn_finalize folded into bn_act_fwd: mean / rstd are derived inside the apply kernel from the shifted sums over the same y
 * (block 0 writes them out for the backward pass and updates the running statistics).  shift = the [C] vector the sums were
 * formed with (wgs_conv_desc.stat_shift, statistics from the conv epilogue) or NULL = row 0 of y (wgs_bn_stats).        */
int wgs_bn_fwd_fused(const float* y, const float* sum, const float* sumsq, const float* shift, long long R, int C, float eps,
                     float momentum, const float* gamma, const float* beta, const float* residual, int relu, float* z, void* zs,
                     float* mean, float* rstd, float* running_mean, float* running_var, void* stream);
int wgs_bn_act_fwd(const float* y, const float* mean, const float* rstd, const float* gamma, const float* beta,
                   const float* residual, int relu, float* z, void* zs, long long R, int C, void* stream);
int wgs_bn_act_bwd_reduce(const float* dz, const float* z, const float* y, const float* mean, const float* rstd,
                          const float* gamma, const float* beta, int relu, long long R, int C, float* sum_dz,
                          float* sum_dzx, void* stream);  /* accumulates; z may be NULL (no residual): the ReLU mask is then
                                                             re-derived from y, gamma, beta - the activation need not be kept */
int wgs_bn_act_bwd_apply(const float* dz, const float* z, const float* y, const float* mean, const float* rstd,
                         const float* gamma, const float* beta, const float* sum_dz, const float* sum_dzx, int relu,
                         long long R, int C, void* dys, float* dy, float* dres, void* stream);

/* 3x3 / stride 2 / pad 1 max-pool of the ResNet stem (torchvision resnet18.maxpool; lib/reconstructor.py:54), NHWC.
 * fwd: out fp32 [N,OH,OW,C], idx uint8 argmax tap, optional split32 pack; bwd: gather, dz [N,H,W,C] fully written.   */
int wgs_maxpool3s2_fwd(const float* z, int N, int H, int W, int C, float* out, void* idx, void* outs, void* stream);
int wgs_maxpool3s2_bwd(const float* dout, const void* idx, int N, int H, int W, int C, float* dz, void* stream);
/* The stem's train-mode BatchNorm + ReLU + max-pool fused (torchvision resnet18 bn1 / relu / maxpool, lib/reconstructor.py:54):
 * the normalised activation is never stored.  fwd: y [N,H,W,C] + the shifted sums of wgs_bn_stats -> pooled out fp32, arg-max
 * table, optional split32 pack, mean / rstd (+ running statistics).  bwd: the pooled gradient is gathered through the arg-max
 * table, the ReLU mask re-derived from y; reduce accumulates sum dzr / sum dzr*xhat, apply writes split32(dy).            */
int wgs_bn_pool_fwd(const float* y, const float* sum, const float* sumsq, const float* shift, int N, int H, int W, int C,
                    float eps, float momentum, const float* gamma, const float* beta, float* out, void* idx, void* outs,
                    float* mean, float* rstd, float* running_mean, float* running_var, void* stream);
int wgs_bn_pool_bwd_reduce(const float* dout, const void* idx, const float* y, const float* mean, const float* rstd,
                           const float* gamma, const float* beta, int N, int H, int W, int C, float* sum_dz,
                           float* sum_dzx, void* stream);
int wgs_bn_pool_bwd_apply(const float* dout, const void* idx, const float* y, const float* mean, const float* rstd,
                          const float* gamma, const float* beta, const float* sum_dz, const float* sum_dzx, int N,
                          int H, int W, int C, void* dys, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WGS_B200_H */
