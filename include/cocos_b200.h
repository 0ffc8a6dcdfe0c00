/* cocos_b200 — C-ABI of the B200-native CoCosNet hot path (libcocos_b200.so).
 *
 * The reference (microsoft/CoCosNet) has no native boundary: its hot path is
 * Python calling ATen.  This header is the boundary we put underneath the
 * reference's module API; each entry point cites the reference
 * expression it replaces.  Conventions:
 *   - every pointer is a DEVICE pointer owned by the caller (torch); the
 *     library never allocates or frees device memory and never synchronises;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*);
 *   - return 0 on success, negative on error; cocos_last_error() gives the
 *     message (thread local).  Nothing throws across the boundary;
 *   - fp16 operand buffers are produced by the cocos_pack_* entry points.
 */
#ifndef COCOS_B200_H
#define COCOS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define COCOS_ABI_VERSION 1

int cocos_abi_version(void);
const char* cocos_last_error(void);

/* fp32 [B, C, N] (channel-major, the layout of theta/phi after
 * correspondence.py:274/276 view/unfold and :277-280 normalisation) ->
 * fp16 [B, N, Kt], Kt = Kp * (split_mode ? 3 : 1), Kp >= C zero padded, Kp%2==0.
 * This is the `theta.permute(0, 2, 1)` of correspondence.py:281 plus the
 * operand rounding.  split_mode: 0 = plain fp16; 1 = query side [hi, lo, hi];
 * 2 = key side [hi, hi, lo] (one GEMM then sums hi*hi + lo*hi + hi*lo); 4 = plain bf16 instead of fp16.
 * rowscale_out (may be NULL): if given, every position n is first multiplied by
 * r[b,n] = 1 / max_c |src[b,c,n]| and r is written to rowscale_out [B,N] (used for
 * the upstream gradient dO in the backward so it always fits fp16). */
int cocos_pack_rows_f16(const float* src, void* dst, int B, int C, int N, int Kp, int split_mode, float* rowscale_out,
                        void* stream);

/* fp32 [B, Cv, Nk] (channel-major exemplar values: avg-pooled ref image
 * correspondence.py:313-315, unfolded patches :311, ref_seg :330-332) ->
 * fp16 (bf16 if `bf16` != 0) [B, Cvp, Nkp], zero padded; for K1: Cvp % 16 == 0, Nkp % 8 == 0. */
int cocos_pack_v_f16(const float* src, void* dst, int B, int Cv, int Nk, int Cvp, int Nkp, int bf16, void* stream);

/* K1, fused correlation + softmax + warp.  Replaces
 *   f = matmul(theta_permute, phi)          correspondence.py:291
 *   f_WTA = f / temperature                 correspondence.py:304
 *   f_div_C = softmax(f_WTA, -1)            correspondence.py:307
 *   y = matmul(f_div_C, ref)                correspondence.py:318 (and :334, :343-344,
 *                                           :362, :368, :370 with operands swapped)
 * q  : fp16 [B, Nq, Kd]   k : fp16 [B, Nk, Kd]   (Kd % 64 == 0)
 * vt : fp16 [B, Cvp, Nkp] (values, channel-major; cocos_pack_v_f16).  May be NULL when v32 is given and Cv <= 4.
 * v32: fp32 [B, Cv, Nk] the same values unpacked (may be NULL).  With Cv <= 4 (the avg-pooled RGB exemplar of
 *      correspondence.py:313-318) the product with the values runs on the CUDA cores inside the exp loop, in fp32.
 * out: fp32 [B, Cv, Nq]   = y.permute(0, 2, 1)  (correspondence.py:323)
 * lse: fp32 [B, Nq] natural-log row log-sum-exp of scale*f (may be NULL)
 * corr: fp32 [B, Nq, Nk] scaled logits f/temperature (may be NULL; the
 *       `return_corr=True` output of correspondence.py:305-306)
 * scale = 1 / temperature. */
int cocos_corr_warp_fwd(const void* q, const void* k, const void* vt, const float* v32, float* out, float* lse,
                        float* corr, int B, int Nq, int Nk, int Kd, int Cv, int Cvp, int Nkp, float scale,
                        void* stream);

/* Batched tcgen05 GEMM: C[b] (MxN fp32 row-major, ldc) = alpha * A[b] (MxK fp16,
 * K contiguous, lda) * B[b]^T (NxK fp16, ldb) (+ C[b] if accumulate); `bf16` != 0
 * treats both operands as bf16.  Strides in
 * elements.  The torch.matmul / 1x1-conv call sites of correspondence.py:291
 * (return_corr), architecture.py:116-125 and the K1 backward. */
int cocos_gemm_f16(const void* a, const void* b, float* c, int batch, int M, int N, int K, int lda, int ldb, int ldc,
                   long long stride_a, long long stride_b, long long stride_c, float alpha, int accumulate,
                   int bf16, void* stream);

/* K1 backward, stage A (what autograd does through correspondence.py:291-318):
 * recomputes S = q k^T and dP' = dO' V^T tile by tile and writes, as bf16,
 *   ds  [B, Nq, Nkp] = P * (dP' - D') * scale / r,   D'[i] = sum_c dO'[i,c] O[c,i]
 *   dst [B, Nk, Nqp] = ds^T,   pt [B, Nk, Nqp] = P^T (optional, may be NULL)
 * q,k as in cocos_corr_warp_fwd; do16 fp16 [B,Nq,Cvk] = row-scaled upstream gradient
 * dO' = r * dO with rscale = r [B,Nq] (cocos_pack_rows_f16 with rowscale_out), v16 fp16
 * [B,Nk,Cvk] = V position-major (Cvk % 64 == 0); out fp32 [B,Cv,Nq] and lse from the
 * forward.  The input gradients follow from three bf16 cocos_gemm_f16 calls:
 *   dq^T[Kd,Nq] = k_cm[Kd,Nk] x ds,  dk^T[Kd,Nk] = q_cm[Kd,Nq] x dst,  dv^T[Cv,Nk] = do_cm[Cv,Nq] x pt. */
int cocos_corr_warp_bwd_ds(const void* q, const void* k, const void* do16, const void* v16, const float* rscale,
                           const float* out, const float* lse, void* ds, void* dst, void* pt, int B, int Nq, int Nk,
                           int Kd, int Cv, int Cvk, int Nkp, int Nqp, float scale, void* stream);

/* Fused PONO + SPADE modulation + LeakyReLU + reflection pad, NCHW fp32:
 *   y[B,C,H+2p,W+2p] = reflect_pad_p( lrelu_slope( (x - mean_c)/sqrt(var_c + eps) * (1 + gamma) + beta ) )
 * with per-pixel statistics over channels, unbiased variance (normalization.py:63-68), the
 * modulation of normalization.py:149, architecture.py:94-95 (slope 0.2; pass 1.0 for the
 * shortcut's plain SPADE) and the ReflectionPad2d of architecture.py:73-74.  gb = [gamma ; beta]
 * as [B,2C,H,W].  mean/rstd [B,H,W] are saved for the backward, which returns dx [B,C,H,W] and
 * dgb [B,2C,H,W] from dy [B,C,H+2p,W+2p].  nhwc != 0: every tensor is channels-last ([B,H,W,C] in memory,
 * C % 4 == 0), one warp per pixel. */
int cocos_spade_mod_fwd(const float* x, const float* gb, float* y, float* mean, float* rstd, int B, int C, int H,
                        int W, int pad, float slope, float eps, int nhwc, void* stream);
int cocos_spade_mod_bwd(const float* dy, const float* x, const float* gb, const float* mean, const float* rstd,
                        float* dx, float* dgb, int B, int C, int H, int W, int pad, float slope, int nhwc,
                        void* stream);

/* K2: TMA-fed tcgen05 implicit-GEMM convolution, forward, KS in {1, 3}, stride 1 (the 3x3 convs of the SPADE blocks,
 * domain adaptor and residual blocks: architecture.py:31-33,73-74; normalization.py:112-120; correspondence.py:17-22).
 *   y[b,n,h,w] = bias[n] + sum_{r,s,c} x[b, h + r - off, w + s - off, c] * wt[n, (r*KS + s)*Cp + c]
 * x  : fp16 (bf16 if `bf16`) NHWC [B, Hin, Win, Cp], Cp % 64 == 0 (cocos_pack_rows_f16 of the NCHW tensor viewed as
 *      [B,C,Hin*Win]); reads outside [0,Hin)x[0,Win) are zero (TMA out-of-bounds fill).
 *      forward, producer already padded (reflection): Hin = H + KS - 1, off = 0; forward with zero padding: Hin = H,
 *      off = KS/2; backward-data (the same kernel): x = dy, wt = spatially flipped W^T, off = KS - 1 (- padding).
 * wt : fp16/bf16 [Cout, KS*KS*Cp].   bias: fp32 [Cout] or NULL.   y: fp32 NCHW [B, Cout, H, W]. */
int cocos_conv_fwd(const void* x, const void* wt, const float* bias, float* y, int B, int H, int W, int Hin, int Win,
                   int Cp, int Cout, int KS, int off, int bf16, void* stream);

/* fp32 [rows, Win] -> 16-bit (fp16, or bf16 if `bf16`) [nshift, rows, Wp], row pitch Wp >= Wout, Wp % 8 == 0 (TMA
 * needs 16-byte strides): dst[s][row][w] = src[row][w + s - off] for w < Wout, zero where that column does not exist.
 * Makes the NCHW 16-bit operand copies of cocos_conv_wgrad (dy: nshift 1, off 0; x: nshift KS column-shifted copies,
 * because a TMA box cannot start at an odd column). */
int cocos_cast_pitch(const float* src, void* dst, long long rows, int Win, int Wout, int Wp, int nshift, int off,
                     int bf16, void* stream);

/* K2w: convolution backward-weights (stride 1, KS in {1,3}) as a tcgen05 GEMM over the B*H*W pixels, split-K over the
 * grid.  Replaces the wgrad half of autograd's convolution_backward for the same nn.Conv2d call sites as
 * cocos_conv_fwd.
 *   ws[(r*KS + s), c, n] = sum_{b,h,w} dy[b, n, h, w] * x[b, c, h + r - off, w + s - off]
 * dy : 16-bit NCHW [B, Cout, H, Wp]  (Wp = W rounded up to 8);
 * x  : 16-bit [KS, B, Cin, Hin, Wp], the KS column-shifted copies cocos_cast_pitch(.., Win, W, Wp, KS, off, ..) makes;
 *      rows outside [0,Hin) are zero.  off = 0 when x carries its halo, KS/2 for zero padding.
 * W >= 64 only (narrower layers: error return; the host mirror sends them to the library wgrad); both operands bf16
 * (dy_bf16 = x_bf16 = 1): a bf16 x fp16 instruction descriptor is rejected by the hardware.
 * ws : fp32 [KS*KS, Cin, Cout], fully overwritten (zeroed + atomically accumulated when the pixel range is split). */
int cocos_conv_wgrad(const void* dy, const void* x, float* ws, int B, int H, int W, int Hin, int Win, int Cout, int Cin,
                     int KS, int off, int dy_bf16, int x_bf16, int flat, void* stream);

/* Operand copies for the flat mode of cocos_conv_wgrad (`flat` = 1; EXPERIMENTAL, not yet validated on hardware --
 * the host mirror only uses it with COCOS_WGRAD_NARROW=1): for maps narrower than 64 the pixels of one (image, channel)
 * plane are flattened to H*W and BOTH shifts of a tap are baked into KS*KS copies:
 *   dst[(r*KS + s), row, h*W + w] = src[row, h + r - off, w + s - off]   (zero outside [0,Hin)x[0,Win)),
 * 16-bit, pitch H*W rounded up to 8.  dy is then passed as [B, Cout, HWp] (cocos_cast_pitch of the flattened view);
 * H*W >= 64 is required. */
int cocos_cast_taps(const float* src, void* dst, long long rows, int Hin, int Win, int H, int W, int KS, int off,
                    int bf16, void* stream);

/* Fused operand prologue for `--PONO_C` (correspondence.py:273-281 / 283-289): x fp32 [B,C,h,w] (output of the theta
 * or phi 1x1 conv) -> unfold(match_kernel, zero pad) -> minus the mean over K = C*mk*mk -> / (L2 norm over K + eps)
 * -> fp16 [B, h*w, K], K laid out tap-major (k = tap*C + c; use the same call for both operands).  xt_workspace:
 * fp32 scratch of B*C*h*w elements.  match_kernel in {1, 3}; C % 4 == 0; K % 64 == 0. */
int cocos_normalize_pack(const float* x, float* xt_workspace, void* out, int B, int C, int h, int w, int match_kernel,
                         float eps, void* stream);

/* Fused InstanceNorm2d(affine=False, eps) + LeakyReLU(slope) over `planes` = B*C contiguous planes of HW fp32
 * elements (NCHW): the norm/activation pairs of the domain adaptor (generator.py:104-113,141-145) and of the
 * PatchGAN (discriminator.py:92-115; normalization.py:52-53).  mean/rstd [planes] are saved for the backward,
 * which returns dx from dy.  slope = 1 is the plain instance norm. */
int cocos_inst_act_fwd(const float* x, float* y, float* mean, float* rstd, int planes, int HW, float slope, float eps,
                       void* stream);
int cocos_inst_act_bwd(const float* dy, const float* x, const float* mean, const float* rstd, float* dx, int planes,
                       int HW, float slope, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* COCOS_B200_H */
