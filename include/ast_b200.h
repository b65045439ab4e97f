/*
 * ast_b200.h -- C ABI of libast_b200.so: hand-written sm_100a kernels for the AdaIN
 * style-transfer hot path of rwickman/ArbitraryStyleTransfer.
 *
 * The reference has no FFI: its hot path is Python calling ATen.  Each
 * entry point below therefore cites the reference Python function (path relative to
 * /root/reference, file:line) whose device work it replaces.  The Python host side
 * (arbitrarystyletransfer_b200/*.py) mirrors the reference's call surface and binds these
 * symbols with ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; no torch / C++ types cross the boundary;
 *   - return 0 on success, >0 = cudaError_t, <0 = AST_E_* argument/shape error;
 *   - never throws, never allocates device memory, never synchronises the device;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = default stream);
 *   - device pointers are borrowed for the duration of the enqueued work; host arrays
 *     (style pointer lists, weights) are copied into kernel parameters before returning;
 *   - tensors are contiguous.  "NCHW fp32" is the reference's own layout; "NHWC16 padded"
 *     is the native inter-layer layout: bf16 [N][H+2][W+2][C] with a one-pixel halo that holds
 *     zeros (VGG, zero padding) or the reflection of the interior (decoder, ReflectionPad2d(1)).
 */
#ifndef AST_B200_H_
#define AST_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AST_ABI_VERSION 3

/* error codes (negative) */
#define AST_E_BADARG   (-1)  /* null pointer / non-positive size */
#define AST_E_SHAPE    (-2)  /* shape not supported by this kernel */
#define AST_E_ALIGN    (-3)  /* pointer not aligned as required */
#define AST_E_TOOMANY  (-4)  /* K > AST_MAX_STYLES */
#define AST_E_NODRIVER (-5)  /* CUDA driver entry point (TMA descriptor encode) unavailable */
#define AST_E_WORKSPACE (-6) /* workspace too small */

#define AST_MAX_STYLES 8

/* flags for the statistics / AdaIN family */
#define AST_F_CANONICAL 0x1u /* scale = sigma_s, shift = mu_s (paper).  Default 0 = the
                                reference's swapped unpack at models.py:44:
                                scale = mu_s, shift = sigma_s. */
#define AST_F_BIASED    0x2u /* divide M2 by HW instead of HW-1 (default: unbiased, as
                                torch.std / torch.var in model_util.py:5, models.py:59) */
#define AST_F_BF16      0x4u /* tensors are bf16 instead of fp32 (stats stay fp32) */

int ast_abi_version(void);
const char* ast_error_string(int code);
/* SM count / compute capability of the current device; returns 0 or a cudaError_t. */
int ast_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---------------------------------------------------------------------------------------
 * K1  fused AdaIN: per-(n,c) Welford mean/variance of content and K style maps, affine
 * re-normalisation, K-style mix and alpha blend in one pass over HBM.
 * Replaces AdaIN.forward (models.py:43-51) + channel_stats x2 (model_util.py:3-8) + the
 * alpha blend `t = alpha*t + (1-alpha)*content_map` (models.py:471).
 *   content : [N][C][HW]          styles[k] : [N][C][style_hw[k]]      out : [N][C][HW]
 *   style_w : K host floats (mix weights; K=1,w=1 is the reference)
 *   stats   : optional device [N*C][2+2K] fp32 = mu_c, sigma_c, (mu_s, sigma_s) x K
 *   eps     : added to the variance before sqrt (0 for the reference AdaIN)
 * out may alias content.  Algorithmic HBM bytes: (1 + K + 1) * N*C*HW * sizeof(elem).
 * ------------------------------------------------------------------------------------- */
int ast_adain_fwd(const void* content, const void* const* styles, const int64_t* style_hw,
                  const float* style_w, int K, void* out, float* stats,
                  int N, int C, int64_t HW, float alpha, float eps, unsigned flags,
                  void* stream);

/* channel_stats (model_util.py:3-8) / calc_mean_std (models.py:54-62): x [rows][HW] ->
 * mean[rows], std[rows] (std = sqrt(var + eps)). */
int ast_channel_stats_fwd(const void* x, float* mean, float* std_, int64_t rows, int64_t HW,
                          float eps, unsigned flags, void* stream);
/* backward of the above: gx = g_mean/HW + g_std * (x - mean) / ((HW-1|HW) * std).
 * g_mean / g_std may be NULL (treated as zero). */
int ast_channel_stats_bwd(const void* x, const float* mean, const float* std_,
                          const float* g_mean, const float* g_std, void* gx,
                          int64_t rows, int64_t HW, unsigned flags, void* stream);

/* mean_variance_norm (models.py:64-68): y = (x - mean) / sqrt(var_unbiased + eps).
 * stats: optional device [rows][2] fp32 (mean, std) saved for the backward.  y may alias x. */
int ast_mvn_fwd(const void* x, void* y, float* stats, int64_t rows, int64_t HW,
                float eps, unsigned flags, void* stream);
/* backward: gx = (gy - mean(gy))/std - y * sum(gy*y) / ((HW-1|HW) * std), with y recomputed
 * from x and the saved stats. */
int ast_mvn_bwd(const void* x, const void* gy, const float* stats, void* gx,
                int64_t rows, int64_t HW, unsigned flags, void* stream);

/* Backward of ast_adain_fwd (AdaIN.forward models.py:43-51 + the alpha blend :471, eps = 0) for a feature map that
 * requires grad.  gy = dL/d out.  stats = the [rows][2+2K] array ast_adain_fwd wrote; style_w = K host floats.
 *   g_content = alpha A / sigma_c * (gy - mean(gy) - z sum(gy z) / (HW-1)) + (1 - alpha) gy,  z = (c - mu_c) / sigma_c
 *   aux (optional, device [K][4][rows] fp32) = per style k: mean_k, std_k, dL/d mean_k, dL/d std_k -- the arguments
 *   of ast_channel_stats_bwd on style map k, which completes the style gradients.
 * flags: AST_F_BF16, AST_F_CANONICAL as in the forward call. */
int ast_adain_bwd(const void* content, const void* gy, const float* stats, const float* style_w, int K,
                  float alpha, void* g_content, float* aux, int64_t rows, int64_t HW, unsigned flags,
                  void* stream);

/* ---------------------------------------------------------------------------------------
 * K3  losses.
 * ------------------------------------------------------------------------------------- */
/* compute_content_loss = F.huber_loss(inp, tgt), delta 1, mean (losses.py:124-126).
 * Writes scale * mean(huber) to loss[0] (device fp32; overwritten).
 * ws: device workspace of >= ast_huber_ws_bytes(n) bytes. */
size_t ast_huber_ws_bytes(int64_t n);
int ast_huber_fwd(const float* inp, const float* tgt, float* loss, int64_t n, float scale,
                  void* ws, size_t ws_bytes, void* stream);
/* g_inp = g_loss[0] * scale / n * clamp(inp - tgt, -1, 1). */
int ast_huber_bwd(const float* inp, const float* tgt, const float* g_loss, float* g_inp,
                  int64_t n, float scale, void* stream);

/* Gram backward on the tensor cores: gx[b] = (gg[b] + gg[b]^T) X[b] / (C*HW) with bf16 copies of both operands
 * (MN-major UMMA operands straight from the NCHW tap), fp32 accumulation.  C % 8 == 0, HW % 8 == 0.
 * ws: >= ast_gram_bwd_tc_ws_bytes(B, C, HW) bytes. */
size_t ast_gram_bwd_tc_ws_bytes(int B, int C, int64_t HW);
int ast_gram_bwd_tc(const float* x, const float* gg, float* gx, int B, int C, int64_t HW, void* ws, size_t ws_bytes,
                    void* stream);

/* compute_hist_loss (losses.py:8-87): squared earth mover's distance between the soft 256-bin
 * histograms (sigmoid-difference bins, L = 1/256, W = L/2.5) of two fp32 image batches x (B, nx elements each) and
 * y (B, ny), averaged over the batch; norm_* = the reference's normaliser x.size(1) * x.size(2).  32.32 fixed-point
 * integer accumulation (bit-deterministic); no (B, 256, C*H*W) intermediate.  gd[B][256] = d loss / d cdf_x is saved
 * for the backward pass.  ws: >= ast_hist_ws_bytes(B) bytes, 8-byte aligned.
 * Backward: gx = sign * g_loss[0] * d loss / d x (sign +1: first argument, -1: second, with its own norm). */
size_t ast_hist_ws_bytes(int B);
int ast_hist_loss_fwd(const float* x, const float* y, int B, int64_t nx, int64_t ny, float norm_x, float norm_y,
                      float* loss, float* gd, void* ws, size_t ws_bytes, void* stream);
int ast_hist_loss_bwd(const float* x, const float* gd, const float* g_loss, float sign, float norm, float* gx, int B,
                      int64_t n, void* stream);

/* tv_loss (losses.py:90-103): loss[0] = sum of squared horizontal + vertical neighbour differences over
 * `planes` = N*C images of H x W (fp32 NCHW); ws as for ast_huber_fwd.  Backward: g_img = g_loss[0] * d loss/d img. */
int ast_tv_fwd(const float* img, float* loss, int64_t planes, int H, int W, void* ws, size_t ws_bytes, void* stream);
int ast_tv_bwd(const float* img, const float* g_loss, float* g_img, int64_t planes, int H, int W, void* stream);

/* gram_matrix (losses.py:105-109): G[b] = X[b] X[b]^T / (C*HW), X fp32 [B][C][HW] -> [B][C][C] */
int ast_gram_fwd(const float* x, float* g, int B, int C, int64_t HW, void* stream);
/* Same Gram matrix on the tensor cores in TF32 (inputs rounded to 10 mantissa bits, fp32
 * accumulate; ~1e-4 relative agreement).  Needs HW % 4 == 0 and C % 16 == 0. */
int ast_gram_fwd_tf32(const float* x, float* g, int B, int C, int64_t HW, void* stream);
/* gx[b] = (gG[b] + gG[b]^T) X[b] / (C*HW) */
int ast_gram_bwd(const float* x, const float* gg, float* gx, int B, int C, int64_t HW,
                 void* stream);

/* ---------------------------------------------------------------------------------------
 * K2  3x3 convolutions (stride 1) on the native layout.
 * Replaces, for VGG-19 (models.py:186-240): nn.Conv2d(k3, zero pad 1, bias) + nn.ReLU +
 * nn.MaxPool2d(2,2); for the classic decoder (models.py:598-628): nn.ReflectionPad2d(1) +
 * nn.Conv2d(k3, p0, bias) + nn.ReLU + nn.Upsample(x2, nearest).
 * ------------------------------------------------------------------------------------- */

/* epilogue placement */
#define AST_EPI_PLAIN 0 /* out[h][w]                          (Ho,Wo) = (H,W)      */
#define AST_EPI_POOL2 1 /* 2x2/2 max pool of relu(conv)       (Ho,Wo) = (H/2,W/2)  */
#define AST_EPI_UP2   2 /* nearest x2 upsample of relu(conv)  (Ho,Wo) = (2H,2W)    */
#define AST_EPI_UPFOLD 4 /* the conv AFTER an Upsample(x2, nearest) + ReflectionPad2d(1) (models.py:602-604, 616-618,
                            622-624), evaluated as four parity-specific 2x2 convs on the LOW-res map: `in` is the
                            low-res (H,W) tensor with an AST_HALO_CLAMP halo, `wpk` comes from
                            ast_pack_conv_weight_fold, (Ho,Wo) = (2H,2W).  tcgen05 path only. */
/* halo written by the epilogue into the padded output */
#define AST_HALO_KEEP    0 /* interior only (halo keeps the caller's zeros: next conv zero-pads) */
#define AST_HALO_REFLECT 1 /* also write the ReflectionPad2d(1) halo of the output grid */
#define AST_HALO_CLAMP   2 /* also write a replicate halo (-1 <- 0, X <- X-1): the reflection of the x2-upsampled
                              grid expressed on the low-res grid; feeds AST_EPI_UPFOLD.  tcgen05 path only. */
/* implementation selector */
#define AST_CONV_AUTO   0 /* tcgen05 implicit GEMM when Cin%64==0 && Cout%64==0, else direct */
#define AST_CONV_TC     1 /* force tcgen05 (AST_E_SHAPE if unsupported) */
#define AST_CONV_DIRECT 2 /* CUDA-core direct kernel (odd shapes; on-device cross-check) */
#define AST_CONV_TC_TAPBOX 3 /* one-CTA tcgen05 kernel variant that loads one 8x16-pixel A box per tap (round-1
                                A/B reference) */
/* AUTO / TC run the CTA-pair kernel (tcgen05.mma.cta_group::2, one {64 ch, 10 w, 18 h} A box per tile and channel
 * block for all nine taps, csrc/conv_pair.cuh) whenever the layer has at least two spatial tiles, else the one-CTA
 * kernel (one {64, 8, 18} box per kw).  impl = 64, 128 or 256 forces the N-block width; 1064, 1128 or 1256 does the
 * same for the TAPBOX variant and 2064, 2128 or 2256 for the pair kernel (tuning / tests).  Environment switches read
 * once per process (A/B measurements): AST_CONV_PAIR=0 (one-CTA kernels), AST_CONV_WIDEA=0 (pair
 * kernel with one A box per kw), AST_CONV_TMA_STORE=1 (TMA-store epilogue for plain tiles), AST_FIRST_NO_TMA_STORE=1
 * (conv1_1 with per-lane stores), AST_CONV_DEBUG=1 (per-role wait cycles on stderr; synchronises). */

typedef struct ast_conv_desc {
  int N, H, W;        /* conv input = conv output spatial size (stride 1, 3x3, pad 1)      */
  int Cin, Cout;
  int relu;           /* fuse ReLU                                                           */
  int epilogue;       /* AST_EPI_*                                                           */
  int halo;           /* AST_HALO_*                                                          */
  int impl;           /* AST_CONV_*                                                          */
  int tap_prerelu;    /* if tap != NULL: 1 = tap holds conv+bias before ReLU, 0 = after     */
} ast_conv_desc;

/* in  : bf16 [N][H+2][W+2][Cin] (halo already holds the padding of this conv)
 * wpk : bf16 packed weights [9][Cout][Cin] (ast_pack_conv_weight)
 * bias: fp32 [Cout] (may be NULL)
 * out : bf16 [N][Ho+2][Wo+2][Cout] or NULL
 * tap : optional fp32 NCHW [N][Cout][H][W] copy of the (pre- or post-ReLU) conv output at
 *       conv resolution, for callers that need reference-layout feature maps. */
int ast_conv3x3_fwd(const ast_conv_desc* d, const void* in, const void* wpk, const float* bias,
                    void* out, float* tap, void* stream);

/* OIHW fp32 [Cout][Cin][3][3] -> bf16 [9][Cout][Cin] (tap = kh*3+kw).
 * flip=1 additionally rotates the taps by 180 degrees and swaps the O/I roles, producing the
 * data-gradient weights: out is [9][Cin][Cout'] i.e. a conv from Cout channels to Cin.
 * cout_pad > Cout (flip must be 0) writes [9][cout_pad][Cin] with zero rows for co >= Cout: the
 * 16-row form ast_conv3x3_last's tensor-core path consumes. */
int ast_pack_conv_weight(const float* w_oihw, void* wpk, int Cout, int Cin, int flip, int cout_pad,
                         void* stream);

/* OIHW fp32 [Cout][Cin][3][3] -> bf16 [16][Cout][Cin], the weights of AST_EPI_UPFOLD: for output parity (py, px) and
 * low-res tap (a, b) in {0,1}^2, tile ((py*2+px)*2+a)*2+b holds the fp32 sum of the 3x3 taps (kh, kw) that land on
 * that low-res pixel -- rows: py=0: a=0 <- {kh 0}, a=1 <- {kh 1,2}; py=1: a=0 <- {kh 0,1}, a=1 <- {kh 2}; columns
 * alike -- rounded to bf16 once.  Replaces nn.Upsample + nn.ReflectionPad2d + nn.Conv2d, models.py:602-604 etc. */
int ast_pack_conv_weight_fold(const float* w_oihw, void* wpk, int Cout, int Cin, void* stream);

/* The first TWO VGG layers in one kernel (csrc/conv12_fused.cuh): Normalization (models.py:129-131) + conv_1 (3->64,
 * zero pad) + relu_1 + conv_2 (64->64, zero pad) + relu_2 + pool_2 (models.py:198-224, vgg19.features[0..4]) from the
 * reference's NCHW fp32 image; the 64-channel full-resolution map between the two convs never reaches HBM.  For passes
 * that tap neither relu_1 nor relu_2 (inference to relu4_1).
 *   img : fp32 [N][3][H][W], W % 4 == 0, H % 2 == 0; w1 : fp32 OIHW [64][3][3][3]; b1 fp32 [64]
 *   mean/std : 3 host floats each (NULL = no normalisation)
 *   wpk2 : ast_pack_conv_weight of conv_2, bf16 [9][64][64]; b2 fp32 [64]
 *   out : bf16 [N][H/2+2][W/2+2][64] (interior only). */
int ast_conv12_fused(const float* img, const float* w1, const float* b1, const float* mean, const float* std_,
                     const void* wpk2, const float* b2, void* out, int N, int H, int W, void* stream);

/* First VGG layer: Normalization (models.py:129-131) + conv_1 (3->Cout, zero pad) + ReLU from
 * the reference's NCHW fp32 image straight into the native layout.  Cout == 64 runs on the tensor
 * cores (im2col A tile built in shared memory, K = 27 padded to 32) unless impl = AST_CONV_DIRECT.
 *   img : fp32 [N][3][H][W]; w : fp32 OIHW [Cout][3][3][3]; bias fp32 [Cout]
 *   mean/std : 3 host floats each (NULL = no normalisation)
 *   out : bf16 [N][H+2][W+2][Cout] (interior only); tap as in ast_conv3x3_fwd. */
int ast_conv3x3_first(const float* img, const float* w, const float* bias, const float* mean,
                      const float* std_, void* out, float* tap, int tap_prerelu,
                      int N, int H, int W, int Cout, int impl, void* stream);

/* Last decoder layer: reflect-padded native input -> NCHW fp32 image, no ReLU.
 *   in : bf16 [N][H+2][W+2][Cin]; w : fp32 OIHW [Cout][Cin][3][3]; out fp32 [N][Cout][H][W]
 *   wpk16 : bf16 [9][16][Cin] from ast_pack_conv_weight(..., cout_pad = 16), or NULL.  With wpk16
 *           and Cin % 64 == 0 the layer runs on the tensor cores (impl AUTO / TC); otherwise the
 *           CUDA-core kernel reads `w` (Cin in {16, 32, 64}, Cout <= 4).
 *   clamp01 != 0 applies Hardtanh(0,1) (Decoder.last_act when exporting, models.py:304,315) */
int ast_conv3x3_last(const void* in, const float* w, const void* wpk16, const float* bias, float* out,
                     int N, int H, int W, int Cin, int Cout, int clamp01, int impl, void* stream);

/* layout converters between the reference layout and the native one.
 * nchw fp32 [N][C][H][W] <-> bf16 [N][H+2][W+2][C]; `halo` as AST_HALO_*. */
int ast_nchw_to_native(const float* nchw, void* native, int N, int C, int H, int W, int halo,
                       void* stream);
int ast_native_to_nchw(const void* native, float* nchw, int N, int C, int H, int W,
                       void* stream);

/* Input path: images cross PCIe as bytes.
 *   ast_u8hwc_to_nchw : uint8 [N][H][W][3] -> fp32 [N][3][H][W] = u8 / 255 (IEEE division), what the loader's
 *                       transforms.ToTensor() computes on the host (data_loader.py:114, 132).
 *   ast_nchw_to_u8hwc : fp32 [N][3][H][W] -> uint8 [N][H][W][3] = trunc(clamp(x, 0, 1) * 255): Hardtanh(0,1) of the
 *                       exporting decoder (models.py:315-316) followed by transforms.ToPILImage()'s
 *                       pic.mul(255).byte() (train.py:18).  NaN maps to 0.
 * The u8 pointer must be 4-byte aligned, the fp32 pointer 16-byte aligned. */
int ast_u8hwc_to_nchw(const void* u8_nhwc, float* nchw, int N, int H, int W, void* stream);
int ast_nchw_to_u8hwc(const float* nchw, void* u8_nhwc, int N, int H, int W, void* stream);

/* ---------------------------------------------------------------------------------------
 * K1n  AdaIN on the native layout (in-pipeline form of K1; same arithmetic).
 *   content : bf16 [N][H+2][W+2][C]; styles[k] : bf16 [N][Hs[k]+2][Ws[k]+2][C] (each style map has its own size)
 *   out     : bf16 [N][H+2][W+2][C], halo per `halo`
 *   ws      : >= ast_adain_native_ws_bytes(N, C, K) bytes
 * ------------------------------------------------------------------------------------- */
size_t ast_adain_native_ws_bytes(int N, int C, int K);
int ast_adain_native_fwd(const void* content, const void* const* styles, const float* style_w,
                         int K, void* out, int N, int C, int H, int W, const int* Hs, const int* Ws,
                         float alpha, float eps, unsigned flags, int halo,
                         void* ws, size_t ws_bytes, void* stream);


/* ---------------------------------------------------------------------------------------
 * Training (BASELINE config 2: decoder training step).  Gradients live in the native layout too.
 * Data gradient of a conv = ast_conv3x3_fwd on dZ with ast_pack_conv_weight(flip = 1) weights.
 * ------------------------------------------------------------------------------------- */

/* Generalised weight packer: bf16 [9][rows_pad][cols_pad], zero beyond the real rows / columns.
 * flip = 0: rows = Cout, cols = Cin.  flip = 1: rows = Cin, cols = Cout, taps rotated 180 degrees
 * (the data-gradient weights).  row_scale: optional device fp32 [rows] multiplied into each row
 * (1/std of Normalization, models.py:131, for the gradient w.r.t. the image). */
int ast_pack_conv_weight_ex(const float* w_oihw, void* wpk, int Cout, int Cin, int flip,
                            int rows_pad, int cols_pad, const float* row_scale, void* stream);

/* NCHW fp32 [N][C][H][W] -> channels [0,C) of a native tensor with Cdst channels and halo width
 * dst_halo (1 or 2); other channels and the halo are left untouched (caller zero-fills once). */
int ast_nchw_to_native_ex(const float* nchw, void* native, int N, int C, int H, int W, int Cdst,
                          int dst_halo, void* stream);
/* Zero the halo ring (width halo = 1 or 2) of a native tensor bf16 [N][H+2*halo][W+2*halo][C]: the zero padding of
 * nn.Conv2d(.., padding=1) in VGG-19 (models.py:192-228) around an interior the producing kernel rewrites fully. */
int ast_zero_halo(void* native, int N, int C, int H, int W, int halo, void* stream);
/* interior of a native tensor with halo width src_halo -> NCHW fp32 */
int ast_native_to_nchw_ex(const void* native, float* nchw, int N, int C, int H, int W, int src_halo,
                          void* stream);

/* nn.MaxPool2d(2,2) forward on the native layout (training keeps the un-pooled activation). */
int ast_maxpool2_native(const void* in, void* out, int N, int C, int H, int W, void* stream);

/* Backward through ReLU (+ MaxPool2d) of one VGG layer (torchvision vgg19 inside
 * models.py:186-240): dZ = [Y>0] * (route(G) + tap_post) + tap_pre, see csrc/train.cu.
 * Y, G, tap_* : native 1-halo (G at pooled size when pooled); dZ: native with halo dz_halo. */
int ast_vgg_bwd_prep(const void* Y, const void* G, const void* tap_post, const void* tap_pre,
                     void* dZ, int N, int C, int H, int W, int pooled, int dz_halo, void* stream);

/* Backward through ReflectionPad2d(1) (+ Upsample x2 nearest) (+ ReLU) between two decoder convs
 * (models.py:598-628): dXpad [N][Hi+4][Wi+4][C] -> dZ [N][Hc+4][Wc+4][C] (2-pixel zero halo). */
int ast_dec_bwd_fold(const void* dXpad, const void* Xi, void* dZ, int N, int C, int Hi, int Wi,
                     int up, int relu, void* stream);

/* native (halo width src_halo) -> channel-planar bf16 [nshift][C][ldq], q = (n*(H+2)+ph)*wp + pw
 * over the 1-halo padded grid with row pitch wp (multiple of 8, >= W+2; ldq = N*(H+2)*wp).
 * copy_halo = 0 writes a zero halo.  nshift = 1: planar[c][q] = x[c][q]; nshift = 3: the three
 * column-shifted copies planar[s][c][q] = x[c][q+s-1] that the kw taps of the wgrad GEMM read. */
int ast_native_to_planar(const void* native, void* planar, int N, int C, int H, int W, int src_halo,
                         int copy_halo, int wp, int nshift, void* stream);

/* Weight gradient on the tensor cores: dwpk [9][Cout][Cin] fp32 (overwritten) =
 * sum_q dzT[co][q] * x[ci][q + (kh-1)*wp + (kw-1)], dz_planar [Cout][ldq] (nshift 1, zero halo),
 * x_planar3 [3][Cin][ldq] (nshift 3, halo copied).  Cin % 16 == 0. */
int ast_conv3x3_wgrad(const void* dz_planar, const void* x_planar3, float* dwpk, int N, int H, int W,
                      int Cin, int Cout, int wp, void* stream);

/* K2wn: weight and bias gradient of a 3x3 conv straight from the native tensors (no planar copies): what autograd
 * derives for nn.Conv2d.weight / .bias of the decoder convs (models.py:598-628) in train.py:287-300.
 * dz: bf16 [N][H+2*dz_halo][W+2*dz_halo][cz], halo ZERO, channels >= Cout zero; x: bf16 [N][H+2][W+2][Cin] with the
 * halo the forward conv saw.  dwpk [9][Cout][Cin] fp32 (overwritten) = sum_p dz[p][co] * x[p + tap][ci];
 * db [Cout] fp32 (optional, overwritten) = sum_p dz[p][co].  cz % 8 == 0, Cin % 8 == 0, Cout <= cz <= 2048. */
int ast_conv3x3_wgrad_native(const void* dz, int cz, int dz_halo, const void* x, float* dwpk, float* db, int N,
                             int H, int W, int Cin, int Cout, void* stream);

/* dwpk -> OIHW fp32 gradient (overwrite or accumulate); b_grad (optional) = row sums of dz_planar. */
int ast_unpack_wgrad(const float* dwpk, float* w_grad, const void* dz_planar, float* b_grad, int Cout,
                     int Cin, int64_t ldq, int accumulate, void* stream);

/* ---------------------------------------------------------------------------------------
 * K4  MobileNet-style blocks (Encoder / Decoder / AutoEncoder, eval mode), plain NHWC 16-bit.
 * Replaces the layers of DepthWiseConv (mobilenetv2.py:95-165), SELayer (:63-81), conv_3x3_bn
 * (:38-43) and Decoder._ref_out/_img_out (models.py:300-316).
 *
 * Storage formats.  Forward ACTIVATIONS -- and the weights they meet in the tensor cores -- are IEEE fp16 (11-bit
 * significand, conversions saturate at +-65504): with bf16's 8 bits the roundings of ~30 blocks in series put 5 % on
 * the deepest encoder tap and 25 % on the stem's weight gradient, all of it from the forward roundings (DESIGN.md
 * section 5).  GRADIENTS are bf16 (range, not precision).  "act" / "grad" below name the role; entry points that
 * serve both roles take a `dtype` argument.  tcgen05.mma takes ONE 16-bit format per instruction, so the
 * weight-gradient GEMM (ast_pw_wgrad: bf16 x bf16) gets its activation operand through ast_cvt_f16_to_bf16.
 * ------------------------------------------------------------------------------------- */
#define AST_DT_BF16 0
#define AST_DT_F16  1
/* Process-wide format of the K4 forward activations: AST_DT_F16 (default: precision) or AST_DT_BF16 (fp32's exponent
 * range: e.g. training from the reference's fresh initialisation, whose closed SE gates leave decoder activations
 * around 1e-20, below fp16's 6e-8).  Set it before any K4 tensor exists; tensors written under one format must not be
 * read under the other.  Gradients are bf16 under both. */
int ast_set_act_format(int dtype);
int ast_get_act_format(void);

/* Pointwise conv as a tcgen05 GEMM: out[p][co] = act(sum_ci x[p][ci] * w[n?][co][ci] + bias[co]) (+ res).
 *   dtype : AST_DT_F16 (forward: activations and weights fp16) or AST_DT_BF16 (data gradients, attention rows);
 *           x, w, residual, out and out_act all have that format
 *   x   : [N*HW][ld_in]  (first Cin channels of each row are read)
 *   w   : [Cout][Cin], or [N][Cout][Cin] when per_sample_w (SE scaling folded in)
 *   act : 0 none, 1 Hardswish;  residual (optional): [N*HW][ld_res];  out: [N*HW][ld_out]
 * Cin, Cout, ld_* multiples of 8. */
/*   out_act (optional, training): `out` keeps the raw pre-activation and out_act [N*HW][ld_act] receives
 *   Hardswish of the rounded value (both are needed by the backward pass); excludes residual.
 *   res_up2_w > 0: the residual is the quarter-size tensor [N*HW/4][ld_res] read through a nearest x2
 *   upsample (DecoderBlock._upsample_3 + identity of _upsample_2); res_up2_w = width of the output grid. */
int ast_pw_conv(const void* x, int ld_in, const void* w, int per_sample_w, const float* bias, int act,
                const void* residual, int ld_res, void* out, int ld_out, int N, int64_t HW, int Cin,
                int Cout, void* out_act, int ld_act, int res_up2_w, int dtype, void* stream);

/* Depthwise k x k (3 or 5), stride 1 or 2, reflect padding (k-1)/2, + bias + optional Hardswish.
 *   x : act (fp16) [N][H][W][C]; w : fp32 [k*k][C]; out : act [N][Ho][Wo][C]
 *   pool (optional) : fp32 [N][C], receives the per-channel SUM of the output (SE squeeze)
 *   up2 : read x through a virtual nearest x2 upsample (conv input = 2H x 2W).
 *   act : 0 none, 1 store Hardswish(y), 2 (training) store raw y and pool Hardswish(y). */
int ast_dw_conv(const void* x, const float* w, const float* bias, void* out, float* pool, int N, int C,
                int H, int W, int k, int stride, int up2, int act, void* stream);

/* SE excitation: scale[n][c] = clamp(W2 relu(W1 (pool[n]*inv_hw) + b1) + b2, 0, 1). */
/* hid_out [N][S] / pre_out [N][C] (optional): post-ReLU hidden layer and pre-clamp output, for ast_se_bwd. */
int ast_se_fc(const float* pool, float inv_hw, const float* w1, const float* b1, const float* w2,
              const float* b2, float* scale, float* hid_out, float* pre_out, int N, int C, int S,
              void* stream);

/* out[n][co][ci] (fp16) = w[co][ci] * se[n][ci]  (se NULL: plain cast, N copies). */
int ast_scale_weights(const float* w, const float* se, void* out, int N, int Cout, int Cin, void* stream);

/* Stem: NCHW fp32 image -> conv3x3 (reflect pad, no bias) -> Hardswish -> NHWC act (fp16); Cout <= 32. */
/* out_raw (optional, training): the rounded pre-activation, NHWC act. */
int ast_stem_conv(const float* img, const float* w, void* out, void* out_raw, int N, int H, int W, int Cout,
                  void* stream);

/* Image head: NHWC act (fp16) -> ReflectionPad2d(1) -> conv3x3 + bias -> NCHW fp32 (+ Hardtanh(0,1)). */
int ast_head_conv(const void* x, const float* w, const float* bias, float* out, int N, int H, int W,
                  int Cin, int Cout, int clamp01, void* stream);

/* NHWC 16-bit (row stride ld, format `dtype`) -> NCHW fp32. */
int ast_nhwc_to_nchw(const void* x, int ld, float* out, int N, int C, int64_t HW, int dtype, void* stream);

/* NCHW fp32 -> NHWC 16-bit (row stride ld, format `dtype`). */
int ast_nchw_to_nhwc(const float* x, void* out, int ld, int N, int C, int64_t HW, int dtype, void* stream);

/* fp16 activation rows -> bf16 copy [rows][ld_out] (first C channels), the operand of ast_pw_wgrad. */
int ast_cvt_f16_to_bf16(const void* x, int64_t ld_x, void* out, int64_t ld_out, int64_t rows, int C, void* stream);

/* ---------------------------------------------------------------------------------------
 * K4t  Training mode of the MobileNet-style blocks (train_autoencoder.py:111-148): nn.BatchNorm2d with
 * batch statistics (mobilenetv2.py:108,128,137,149), Hardswish / SELayer / residual passes and all
 * backward kernels.  Tensors are NHWC [N*HW][ld]: activations (x, a, a_pre, z, res, out of forward passes) fp16,
 * gradients (du, dy, da, dx, dres) bf16; statistics fp32, cross-CTA sums fp64.
 * `stat` = float[4][C]: mean, invstd, scale = gamma*invstd, shift = beta - mean*scale.
 * ------------------------------------------------------------------------------------- */

/* sums[0][c] = sum x, sums[1][c] = sum x^2 over all N*HW rows. */
int ast_bn_stats(const void* x, int ld, double* sums, int N, int C, int64_t HW, void* stream);
/* sums -> stat; running_mean/var (nullable) updated like nn.BatchNorm2d (momentum, unbiased variance);
 * num_batches_tracked (nullable, device int64 scalar) += 1 as nn.BatchNorm2d.forward does in training mode. */
int ast_bn_finalize(const double* sums, double count, const float* gamma, const float* beta,
                    float* running_mean, float* running_var, float momentum, float eps, float* stat, int C,
                    long long* num_batches_tracked, void* stream);
/* out = act(x*sc[c] + sh[c]) [* se[n][c]] [+ res]; pool[n][c] (optional) = sum over pixels of act(..)
 * (before the se factor).  sc/sh null = identity; act 0 none, 1 Hardswish; out may be null (pool only). */
int ast_affine_act(const void* x, int ld_x, const float* sc, const float* sh, int act, const float* se,
                   const void* res, int ld_res, void* out, int ld_out, float* pool, int N, int C, int64_t HW,
                   void* stream);
/* Backward through u = Hardswish(z)*s, z = a*scale+shift: out[n][0..4][c] = sums over pixels of
 * du*h, du*h', h', du*h'*ahat, h'*ahat (h' = dHardswish/dz, ahat = (a-mean)*invstd; stat null: identity). */
int ast_dw_bwd_reduce(const void* du, const void* a, const float* stat, float* out, int N, int C, int64_t HW,
                      void* stream);
/* BatchNorm coefficients of the depthwise norm from those sums: dgamma, dbeta, coef[2][C] = (dbeta, dgamma)/count. */
int ast_se_bn_combine(const float* T, const float* s, const float* g, float* dgamma, float* dbeta, float* coef,
                      int N, int C, double count, void* stream);
/* da = ((du*s[n][c] + g[n][c]) * h'(z) - coef0 - ahat*coef1) * scale   (stat null: da = (du*s+g)*h'(a)). */
int ast_dw_bwd_apply(const void* du, const void* a, const float* s, const float* g, const float* stat,
                     const float* coef, void* da, int N, int C, int64_t HW, void* stream);
/* Generic BatchNorm backward: sums[0][c] = sum dy, sums[1][c] = sum dy*ahat;  finalize -> dgamma, dbeta,
 * coef;  apply: da = (dy - coef0 - ahat*coef1) * scale. */
int ast_bn_bwd_reduce(const void* dy, int ld_dy, const void* a, int ld_a, const float* stat, double* sums, int N,
                      int C, int64_t HW, void* stream);
int ast_bn_bwd_finalize(const double* sums, double count, float* dgamma, float* dbeta, float* coef, int C,
                        void* stream);
int ast_bn_bwd_apply(const void* dy, int ld_dy, const void* a, int ld_a, const float* stat, const float* coef,
                     void* da, int N, int C, int64_t HW, void* stream);
/* Depthwise conv data gradient (gather form, reflection and x2-upsample folded back):
 *   dy [N][Ho][Wo][C] -> dx [N][H][W][C];  w fp32 [k*k][C];  a_pre (optional): dx *= Hardswish'(a_pre*scale+shift). */
/*   dres (optional): residual-branch gradient [N][Ho][Wo][C] added to dx (summed over 2x2 when up2);
 *   requires stride 1 (Ho x Wo == conv-input size). */
int ast_dw_conv_dgrad(const void* dy, const float* w, const void* a_pre, const float* stat, const void* dres,
                      void* dx, int N, int C, int H, int W, int k, int stride, int up2, void* stream);
/* Depthwise conv weight gradient, accumulated (atomics) into dw fp32 (C,1,k,k); x = the conv's stored input. */
int ast_dw_conv_wgrad(const void* dy, const void* x, float* dw, int N, int C, int H, int W, int k, int stride,
                      int up2, void* stream);
/* SELayer backward (mobilenetv2.py:73-81): ds [N][ds_stride] = gradient of the scale; outputs dpre, dhid
 * (scratch, [N][C] / [N][S]), g[n][c] = gradient of the pooled mean / HW, and the four parameter gradients. */
int ast_se_bwd(const float* ds, int ds_stride, const float* pre, const float* hid, const float* pool,
               float inv_hw, const float* w1, const float* w2, float* dpre, float* dhid, float* g, float* dw1,
               float* db1, float* dw2, float* db2, int N, int C, int S, void* stream);
/* Pointwise conv weight gradient on tcgen05 (MN-major operands): out[i*si + j*sj] += sum_p a[p][i]*b[p][j]. */
int ast_pw_wgrad(const void* a, int ld_a, int Ca, const void* b, int ld_b, int Cb, int64_t P, float* out,
                 int64_t si, int64_t sj, void* stream);
/* Stem (conv_3x3_bn) weight gradient: dw (16,3,3,3) += sum dy*Hardswish'(z) (x) img[reflected tap]. */
int ast_stem_wgrad(const void* dy, const void* z, const float* img, float* dw, int N, int H, int W, int Cout,
                   void* stream);
/* Stem data gradient: NCHW fp32 gradient of the image (3 channels) from the stem output's gradient dy (bf16) and the
 * stored pre-activation z; reflection padding folded back.  Needed only when the Encoder's INPUT requires grad. */
int ast_stem_dgrad(const void* dy, const void* z, const float* w, float* dimg, int N, int H, int W, int Cout,
                   void* stream);
/* Hardtanh(0,1) backward of the exporting head (models.py:304, 315-316): dx = dy where 0 < y < 1 else 0 (fp32). */
int ast_hardtanh01_bwd(const float* dy, const float* y, float* dx, int64_t n, void* stream);
/* Image head (_ref_out + _img_out): weight / bias gradient (accumulated) and data gradient (NHWC bf16). */
int ast_head_wgrad(const float* dY, const void* x, float* dw, float* db, int N, int H, int W, int Cin, int Cout,
                   void* stream);
int ast_head_dgrad(const float* dY, const float* w, void* dx, int N, int H, int W, int Cin, int Cout,
                   void* stream);
/* fp32 [R][Cc] -> mode 0: bf16 [R][Cc]; 1: bf16 [Cc][R]; 2: fp32 [Cc][R]; 3: fp16 [R][Cc] (forward GEMM weights). */
int ast_prep_weight(const float* w, void* out, int R, int Cc, int mode, void* stream);

/* The five prepared forms of ONE DepthWiseConv block's weights (mobilenetv2.py:103-150) in one launch: pointwise
 * w1 [R1][C1] (nullable: expand_ratio == 1 blocks) and w2 [R2][C2] -> activation-format copy (w*_f, same layout, the
 * forward GEMM operand) + bf16 transposed copy (w*_t, the data-gradient GEMM operand); depthwise wd [Rd][Cd] ->
 * fp32 transposed [Cd][Rd]. */
int ast_prep_block_weights(const float* w1, int R1, int C1, void* w1_f, void* w1_t, const float* wd, int Rd, int Cd,
                           float* wd_t, const float* w2, int R2, int C2, void* w2_f, void* w2_t, void* stream);

/* ---------------------------------------------------------------------------------------
 * K6  AdaAttN (models.py:70-115).  The layer's contractions -- Q K^T (:97),
 * P V and P V^2 (:101-103) and the five products of their backward pass -- run on one batched tcgen05 GEMM that
 * reads either operand K-major or MN-major, so the NHWC activations are used as they lie in HBM; the passes
 * between them are row-matrix streaming kernels.  All bf16 tensors are [rows][channels] with the given row stride.
 * ------------------------------------------------------------------------------------- */

/* d[b][i][j] = sum_k A(b,i,k) * B(b,j,k), bf16 operands, fp32 accumulation, d fp32 (d_bf16 = 0) or bf16.
 * a_mn = 0: a is [M][K] (row stride ld_a); a_mn = 1: a is [K][M].  b_mn = 0: b is [N][K]; b_mn = 1: b is [K][N].
 * ld_a, ld_b, a_bs, b_bs (batch strides, elements) multiples of 8; a, b 16-byte aligned; any M, N, K. */
int ast_bgemm(const void* a, int a_mn, int ld_a, int64_t a_bs, const void* b, int b_mn, int ld_b, int64_t b_bs,
              void* d, int d_bf16, int64_t ld_d, int64_t d_bs, int M, int N, int K, int batch, void* stream);
/* p[r][:] = bf16(softmax(s[r][:])) (nn.Softmax(dim=-1), models.py:75, 99); lsum[r] = sum of the rounded weights. */
int ast_attn_softmax(const float* s, int64_t ld_s, void* p, int64_t ld_p, float* lsum, int64_t rows, int cols,
                     void* stream);
/* ds[r][j] = p[r][j] * (da[r][j] - sum_j' da[r][j'] p[r][j'] / lsum[r]). */
int ast_attn_softmax_bwd(const void* p, int64_t ld_p, const float* lsum, const float* da, int64_t ld_a, void* ds,
                         int64_t ld_o, int64_t rows, int cols, void* stream);
/* out[r] = [v | hi | lo] (3C bf16 per row), hi + lo = v^2 exactly. */
int ast_attn_vv3(const void* v, int64_t ld_v, void* out, int64_t rows, int C, void* stream);
/* out[r] = [v | v | hi | hi | lo] (5C): backward-pass partner of ast_attn_out_bwd's dmm. */
int ast_attn_vv5(const void* v, int64_t ld_v, void* out, int64_t rows, int C, void* stream);
/* mm = P [v | hi | lo] (fp32, 3C per row): mean = mm0/lsum, m2 = (mm1+mm2)/lsum;
 * out = sqrt(relu(m2 - mean^2)) * cn + mean  (models.py:103, 115), bf16. */
int ast_attn_out_fwd(const float* mm, const float* lsum, const void* cn, int64_t ld_cn, void* out, int64_t rows,
                     int C, void* stream);
/* From d(out): dmm[r] = [dMean_hi | dMean_lo | dM2_hi | dM2_lo | dM2_hi] (bf16, 5C per row: two-term splits of
 * d mean and d m2, already divided by lsum) and dcn = d(out) * std. */
int ast_attn_out_bwd(const float* mm, const float* lsum, const void* cn, int64_t ld_cn, const void* dout, void* dmm,
                     void* dcn, int64_t rows, int C, void* stream);
/* dv = (dvv0 + dvv1) + 2 v (dvv2 + dvv3)  (dvv = P^T dmm[:, :4C], fp32, 4C per row). */
int ast_attn_dv(const float* dvv, const void* v, int64_t ld_v, void* dv, int64_t rows, int C, void* stream);
/* Two-term bf16 split x = hi + lo of an fp32 row matrix / NCHW tensor into rows of 3C: pattern 0 [hi | lo | hi]
 * (A side), pattern 1 [hi | hi | lo] (B side): one K = 3C GEMM of an A-side by a B-side matrix is the product to
 * ~2^-17 -- used for the logits Q K^T and the 1x1 convolutions W_q, W_k in front of them (softmax exponentiates
 * the logits' absolute error; models.py:87-88, 97 have no 1/sqrt(d) scaling). */
int ast_split3_rows(const float* x, int64_t ld_x, void* out, int64_t rows, int C, int pattern, void* stream);
int ast_split3_nchw(const float* x, void* out, int N, int C, int64_t HW, int pattern, void* stream);
/* out = a*x + b*y (y nullable: a*x), fp32: the alpha blend of models.py:471 and its gradient. */
int ast_axpby(const float* x, const float* y, float a, float b, float* out, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AST_B200_H_ */
