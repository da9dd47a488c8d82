/* libpidm -- C ABI of the B200-native physics-informed-diffusion hot path.
 *
 * The reference (jhbastek/PhysicsInformedDiffusionModels) has no FFI: its boundary is the Python class
 * surface used by main.py / sample.py.  This header is the operator interface those
 * classes are re-implemented on: plain pointers + sizes, no torch types.  Every entry point
 *   - takes raw DEVICE pointers borrowed from the caller (caller keeps them alive; nothing is allocated
 *     inside except through caller-provided workspaces),
 *   - launches on the `stream` argument (a cudaStream_t passed as void*), never synchronises,
 *   - returns 0 on success, non-zero on error with a message in pidm_last_error().
 * Activations are NHWC ([B, H*W, C]); `dtype` is PIDM_F32 (0) or PIDM_BF16 (1) and names the ACTIVATION /
 * packed-weight storage type; statistics, parameters, gradients of parameters and all reductions are fp32.
 * Each declaration cites the reference code it replaces (paths relative to the reference repository).
 */
#ifndef PIDM_H_
#define PIDM_H_

#ifdef __cplusplus
extern "C" {
#endif

#ifndef PIDM_F32
#define PIDM_F32 0
#define PIDM_BF16 1
#endif

const char* pidm_last_error(void);
int pidm_version(void);

/* ---- diffusion element-wise ops ------------------------------------------------------------------------- */
/* q_sample: x_t = sqrt(abar_t) x0 + sqrt(1-abar_t) eps.   src/denoising_utils.py:373-378 and inline :633-638 */
int pidm_qsample(const float* x0, const float* noise, const long long* t, const float* sqrt_ab,
                 const float* sqrt_1mab, float* xt, int B, int per_sample, void* stream);
/* ancestral step: out = coef1*x0_pred + coef2*x_t + sigma*z.   src/denoising_utils.py:441-455 */
int pidm_posterior_step(const float* x_t, const float* x0_pred, const float* z, float* out, float coef1, float coef2,
                        float sigma, long long n, void* stream);
/* out = a_b*x + b_b*y + c_b*z with per-sample coefficients [B]: the eta=0 DDIM jump of ddim_sample_x0
 * (src/denoising_utils.py:755-785) collapses to this form */
int pidm_axpby_per_sample(const float* a, const float* x, const float* b, const float* y, const float* c,
                          const float* z, float* out, int B, int per_sample, void* stream);
/* per-sample coefficients of the deterministic DDIM jump t -> t_next (src/denoising_utils.py:755-781, eta = 0):
 * x' = coef_x0 * x0_pred + coef_x * x  (coef_x0 = 0, coef_x = 1 where t == t_next); tables are the diff_dict entries */
int pidm_ddim_coefs(const long long* t, const long long* t_next, const float* posterior_mean_coef1,
                    const float* posterior_mean_coef2, const float* sqrt_recip_alphas, const float* noise_mean_coeff,
                    const float* alphas_prod, float* coef_x0, float* coef_x, int B, void* stream);
/* out = x * *alpha_dev  (chain-rule scaling of a precomputed gradient by the upstream scalar; out may alias x) */
int pidm_scale(const float* x, const float* alpha_dev, float* out, long long n, void* stream);

/* ---- toy study (main_toy.py, src/denoising_toy_utils.py:436-511): PIDM loss algebra on [B,D] points ---------- */
/* data term c_data*mean_b(w_b*mean_D(target-output)^2) with w_b = p2[t_b] (p2_loss_weight != NULL) or 1; Gaussian NLL of the
 * residual / inequality values with the log-likelihood clamped at -27.631 (:381); lambda*mean(opt).  ineq / opt may be
 * NULL.  sums7 = data, residual, inequality, optimisation terms, mean|r|, mean(ineq), mean(opt); gradients overwritten. */
int pidm_toy_pidm_loss(const float* target, const float* output, const float* residual, const float* ineq,
                       const float* opt, const long long* t, const float* p2_loss_weight,
                       const float* posterior_var_clipped, float c_data, float c_residual, float c_ineq, float lambda_opt,
                       float* sums7, float* grad_output, float* grad_residual, float* grad_ineq, float* grad_opt, int B,
                       int D, void* stream);

/* ---- Darcy residual (src/residuals_darcy.py:134-183 + src/grad_utils.py:64-146) ------------------------- */
/* x0hat [B,2,P,P] fp32 NCHW (p, K); f_s [P*P]; residual [B,P*P,3] = (eq_0, bc_x0, bc_x1).  P must be 64. */
int pidm_darcy_residual_fwd(const float* x0hat, const float* f_s, float* residual, int B, int pixels,
                            float domain_length, int reverse_d1, int pixels_at_boundary, void* stream);
/* vector-Jacobian product of the above: grad_x0hat [B,2,P,P] = J^T grad_residual */
int pidm_darcy_residual_bwd(const float* x0hat, const float* f_s, const float* grad_residual, float* grad_x0hat, int B,
                            int pixels, float domain_length, int reverse_d1, int pixels_at_boundary, void* stream);
/* one derivative field of u [planes,P,P]: mode 0..4 = d_d0, d_d1, d_d00, d_d11, d_d01 (StencilGradients.forward,
 * src/grad_utils.py:161-175; second-order, one-sided at the boundary) */
int pidm_fd_stencil(const float* u, float* out, int planes, int pixels, int mode, float d0, float d1, void* stream);
/* Fused PIDM loss (src/denoising_utils.py:669-692): sums3 = {c_data*mean_b(p2[t] mse_b), mean(c_res*0.5 r^2/var_t),
 * mean|r|}; optionally the gradient of (sums3[0]+sums3[1]) w.r.t. x0hat (residual operand) and model_out (data
 * operand; pass the same pointer twice in 'mean' mode, then only grad_x0hat is written).  The residual is never
 * materialised.  grad pointers may be NULL (loss only). */
int pidm_darcy_pidm_loss(const float* x0hat, const float* model_out, const float* target, const float* f_s,
                         const long long* t, const float* p2_loss_weight, const float* posterior_var_clipped,
                         float c_data, float c_residual, float* sums3, float* grad_x0hat, float* grad_model_out, int B,
                         int pixels, float domain_length, int reverse_d1, int pixels_at_boundary, void* stream);
/* CoCoGen step size (src/residuals_darcy.py:218-231): max_dr_dp[b] = largest entry (signed, as torch.max) of the Jacobian
 * d residual / d p of sample b -- evaluated analytically from the stencil coefficients and K, the reference materialises
 * the 12288 x 4096 Jacobian per sample with vmap(jacfwd). */
int pidm_darcy_jacobian_max(const float* x0hat, float* max_dr_dp, int B, int pixels, float domain_length, int reverse_d1,
                            int pixels_at_boundary, void* stream);

/* ---- layout ------------------------------------------------------------------------------------------- */
/* image_to_b_xy_c / b_xy_c_to_image (src/denoising_utils.py:36-55) fused with the dtype change + channel padding */
int pidm_nchw_to_nhwc(const float* src, void* dst, int B, int C, int HW, int Cpad, int dtype, void* stream);
int pidm_nhwc_to_nchw(const void* src, float* dst, int B, int C, int HW, int Cpad, int dtype, void* stream);
int pidm_add(const void* a, const void* b, void* out, long long n, int dtype, void* stream);
/* exact (erf) GELU on activations, n % 8 == 0: emb_conv of the residual-gradient guidance branch (src/unet_model.py:520-524) */
int pidm_gelu_fwd(const void* x, void* y, long long n, int dtype, void* stream);
int pidm_gelu_bwd(const void* x, const void* dy, void* dx, long long n, int dtype, void* stream);
/* torch.cat((x, skip), dim=1) on NHWC rows and its backward (src/unet_model.py:606,612) */
int pidm_concat_channels(const void* a, const void* b, void* out, long long rows, int Ca, int Cb, int dtype, void* stream);
int pidm_split_channels(const void* g, void* ga, void* gb, long long rows, int Ca, int Cb, int dtype, void* stream);

/* ---- convolutions as implicit GEMM (src/unet_model.py:163,197,227,253,275,279,453,517) ------------------- */
/* Packed weights: Wp[n][tap*Cin + c] in the activation dtype, built by pidm_pack_weights from the framework
 * layout through strides.  PackEntry (56 bytes, see pidm_pack_entry_size):
 *   { const float* src; void* dst; long long s_n, s_c; int N, C, Cpad, taps, flip, pad_; }
 *   src index = n*s_n + c*s_c + (flip ? taps-1-tap : tap) */
int pidm_pack_entry_size(void);
int pidm_pack_weights(const void* table_dev, int n_entries, int dtype, void* stream);
/* Forward AND dgrad operand of a layer from one read of its weights (layers with Cin % 32 == 0, Cout % 32 == 0,
 * <= 16 taps); one CTA per 32 x 32 channel block.  PackPairEntry (64 bytes, see pidm_pack_pair_entry_size):
 *   { const float* src; void* dst_f; void* dst_d (may be null); long long s_co, s_ci; int Cout, Cin, taps, flip;
 *     int tile0, pad_; }
 *   src index = co*s_co + ci*s_ci + tap;  dst_f[co][tap*Cin + ci];  dst_d[ci][(flip ? taps-1-tap : tap)*Cout + co]
 *   entry e owns blocks [tile0, tile0 + (Cout/32)*(Cin/32)); tile_map_dev[i] = index of the entry that owns block i;
 *   one launch packs blocks [tile_base, tile_base + n_tiles). */
int pidm_pack_pair_entry_size(void);
int pidm_pack_weights_pairs(const void* table_dev, const int* tile_map_dev, int tile_base, int n_tiles, int max_taps,
                            int dtype, void* stream);
/* y[b,oh,ow,n] = sum A(m,k) Wp[n,k] + bias[n] + residual;  transposed=0: A gathers x at (oh*s-p+r, ow*s-p+q);
 * transposed=1: at ((oh+p-r)/s, (ow+p-q)/s) when divisible (ConvTranspose forward / strided-conv dgrad).
 * CUDA-core fp32-accumulate kernel for every geometry (parity anchor + layers the tensor-core kernel skips). */
int pidm_conv2d_simt(const void* x, const void* w_packed, const float* bias, const void* residual, void* y, int B,
                     int H, int W, int Cin, int Ho, int Wo, int Cout, int KH, int KW, int stride, int pad,
                     int transposed, int dtype, void* stream);
/* dW (framework layout, index n*w_stride_n + c*w_stride_c + tap) += sum_m dy[m,n] A(m,tap,c); dbias[n] += sum_m dy */
int pidm_conv2d_wgrad_simt(const void* x, const void* dy, float* dw, float* dbias, int B, int H, int W, int Cin,
                           int Cin_real, int Ho, int Wo, int Cout, int KH, int KW, int stride, int pad, int transposed,
                           long long w_stride_n, long long w_stride_c, int dtype, void* stream);
/* debugging aid: device buffer (>= 4096 int64) receiving a clock64 timeline of CTA 0 of every tensor-core conv launch */
int pidm_debug_set_trace(void* buf);
/* tcgen05 + TMA implicit-GEMM convolution, bf16 operands, fp32 TMEM accumulation; same contract as pidm_conv2d_simt
 * (requires Cin % 32 == 0, Cout % 32 == 0): stride-1/2 regular convolution (input sampled through TMA elementStrides) and the
 * stride-2 transposed gather (ConvTranspose forward / dgrad of the stride-2 conv) as 4 output-parity classes.
 * gn_sums (optional, [B, gn_groups, 2]): per-(sample, group) sum and sum of squares of the fp32 output, accumulated in
 * the epilogue so that the following GroupNorm needs no statistics pass.  It is zeroed here (one memset node) unless
 * gn_sums_zeroed != 0, i.e. the caller hands in a slice of a buffer it has already cleared. */
int pidm_conv2d_tc_general(const void* x, const void* w_packed, const float* bias, const void* residual, void* y, int B,
                           int H, int W, int Cin, int Ho, int Wo, int Cout, int KH, int KW, int stride, int pad,
                           int transposed, float* gn_sums, int gn_groups, int gn_sums_zeroed, void* stream);
int pidm_conv2d_tc_general_supported(int B, int H, int W, int Cin, int Ho, int Wo, int Cout, int KH, int KW, int stride,
                                     int pad, int transposed);
/* wgrad on tcgen05: D[(tap,cA)][cB] = sum over grid pixels g of a[a_stride*g - pad + tap][cA] * b[g][cB], MN-major
 * (pixel-strided) TMA operands, split over pixel ranges, red.global.add into dw[cA*s_row + cB*s_col + tap] (fp32,
 * ACCUMULATED).  Regular conv: a = x, b = dy.  ConvTranspose: a = dy (a_stride 2), b = x.  Rows cA >= CA_real (channel
 * padding) are dropped. */
int pidm_conv2d_wgrad_tc(const void* a, const void* b, float* dw, int B, int HA, int WA, int CA, int CA_real, int GH,
                         int GW, int CB, int KH, int KW, int a_stride, int pad, long long s_row, long long s_col,
                         void* stream);
int pidm_conv2d_wgrad_tc_supported(int B, int GH, int GW, int CA, int CB, int KH, int KW, int a_stride);
/* out[c] += sum_m x[m][c]: bias gradients (column sums of an NHWC tensor) */
int pidm_colsum(const void* x, float* out, long long M, int C, int dtype, void* stream);

/* ---- normalisations ------------------------------------------------------------------------------------ */
/* Block.forward tail: GroupNorm(G) -> *(scale+1)+shift -> SiLU (src/unet_model.py:233-241).  scale_shift [B,2C] or NULL.
 * residual (optional, same shape as y) is added after the SiLU: the `h + x` of a ResnetBlock whose res_conv is the
 * identity (:262).  sums [B,G,2] (sum, sum of squares) is written here and consumed by the backward. */
int pidm_groupnorm_silu_fwd(const void* x, const float* gamma, const float* beta, const float* scale_shift,
                            const void* residual, void* y, float* sums, int stats_precomputed, int B, int HW, int C, int G,
                            float eps, int dtype, void* stream);
/* workspace: float[B*C*2]; dgamma/dbeta ACCUMULATE; d_scale_shift [B,2C] overwritten (may be NULL);
 * dbias_of_producer (may be NULL): column sums of dx ACCUMULATED = bias gradient of the convolution that produced x. */
int pidm_groupnorm_silu_bwd(const void* x, const void* dy, const float* sums, const float* gamma, const float* beta,
                            const float* scale_shift, void* dx, float* dgamma, float* dbeta, float* d_scale_shift,
                            float* dbias_of_producer, float* workspace, int B, int HW, int C, int G, float eps,
                            int dtype, void* stream);
/* channel LayerNorm, gain only, biased variance (src/unet_model.py:201-210); dgamma ACCUMULATES */
int pidm_layernorm_c_fwd(const void* x, const float* gamma, void* y, long long M, int C, float eps, int dtype, void* stream);
int pidm_layernorm_c_bwd(const void* x, const void* dy, const float* gamma, void* dx, float* dgamma,
                         const void* dx_residual /* optional: added to dx (skip-connection gradient) */, long long M, int C,
                         float eps, int dtype, void* stream);

/* ---- attention ------------------------------------------------------------------------------------------ */
/* SpatialLinearAttention core between to_qkv and to_out (src/unet_model.py:286-297), dim_head = 32.
 * qkv [B,N,3*heads*32]; out [B,N,heads*32]; ctx [B,heads,32,32], kmax/kzinv [B,heads,32] kept for backward. */
/* Linear attention fused with its to_qkv 1x1 projection (C = 32 input channels, 8 heads, bf16): q, k, v are recomputed
 * per head on the tensor cores from xn = PreNorm(x) instead of being materialised (reference unet_model.py:275-297 --
 * `qkv = self.to_qkv(x)` and everything after it).  w_qkv = packed to_qkv weights [768][32] bf16 (pidm_pack_weights).
 * Forward keeps ctx / kmax / kzinv for backward; backward returns dqkv [B,N,768] (gradient w.r.t. xn W^T) for the
 * ordinary dgrad / wgrad of to_qkv.  dctx is scratch. */
int pidm_linattn_fused_supported(int C, int heads, int N, int dtype);
int pidm_linattn_fused_workspace_floats(int B, int N);
int pidm_linattn_fused_fwd(const void* xn, const void* w_qkv, void* out, float* ctx, float* kmax, float* kzinv,
                           float* workspace, int B, int N, void* stream);
int pidm_linattn_fused_bwd(const void* xn, const void* w_qkv, const void* dout, const float* ctx, const float* kmax,
                           const float* kzinv, void* dqkv, float* dctx, int B, int N, void* stream);
int pidm_linattn_workspace_floats(int B, int N, int heads);
int pidm_linattn_fwd(const void* qkv, void* out, float* ctx, float* kmax, float* kzinv, float* workspace, int B, int N,
                     int heads, int dtype, void* stream);
int pidm_linattn_bwd(const void* qkv, const void* dout, const float* ctx, const float* kmax, const float* kzinv,
                     void* dqkv, float* dctx_scratch, int B, int N, int heads, int dtype, void* stream);
/* mid-block softmax attention over <= 64 tokens (src/unet_model.py:341-367) */
int pidm_attn_fwd(const void* qkv, void* out, int B, int n_tokens, int heads, int dtype, void* stream);
int pidm_attn_bwd(const void* qkv, const void* dout, void* dqkv, int B, int n_tokens, int heads, int dtype, void* stream);

/* ---- time conditioning (fp32) --------------------------------------------------------------------------- */
/* SinusoidalPosEmb + time_mlp (src/unet_model.py:147-159,464-469); also returns SiLU(temb) for the block MLPs */
int pidm_time_embed_fwd(const long long* t, const float* W1, const float* b1, const float* W2, const float* b2,
                        float* emb, float* h1, float* temb, float* silu_t, int B, int dim, int td, void* stream);
/* workspace float[2*B*td]; parts: bit 1 = activation gradients into the workspace, bit 0 = weight / bias gradients from
 * it (ACCUMULATED) -- the second half only feeds the optimizer and may be issued on another stream after the first. */
int pidm_time_embed_bwd(const float* d_silu_t, const float* emb, const float* h1, const float* temb, const float* W2,
                        float* dW1, float* db1, float* dW2, float* db2, float* workspace, int B, int dim, int td,
                        int parts, void* stream);
/* every ResnetBlock.mlp Linear in one launch (src/unet_model.py:246-249,258-262).  MlpEntry (56 bytes):
 *   { const float* W; const float* b; float* dW; float* db; float* out; const float* dout; int n, pad_; }
 *   out_e[b, j] = b_e[j] + W_e[j,:] . silu_t[b,:] ;  backward accumulates dW_e, db_e and overwrites d_silu_t. */
int pidm_mlp_entry_size(void);
int pidm_block_mlps_fwd(const void* table_dev, int n_entries, int max_rows, const float* silu_t, int B, int td,
                        void* stream);
/* parts: bit 0 = weight / bias gradients, bit 1 = input gradient d_silu_t (independent halves) */
int pidm_block_mlps_bwd(const void* table_dev, int n_entries, int max_rows, const float* silu_t, float* d_silu_t, int B,
                        int td, int parts, void* stream);

/* ---- output head: final 1x1 conv to NCHW fp32 (+ sigmoid on the last channel) src/unet_model.py:517,619-621 */
int pidm_head_fwd(const void* x, const float* w, const float* bias, float* y, int B, int HW, int C, int O,
                  int sigmoid_last, int dtype, void* stream);
int pidm_head_bwd(const void* x, const float* w, const float* y, const float* dy, void* dx, float* dw, float* db, int B,
                  int HW, int C, int O, int sigmoid_last, int dtype, void* stream);

/* ---- step glue on flat buffers (main.py:163-166,178-183; src/denoising_utils.py:163-205) ----------------- */
/* out[0] += sum x^2, deterministic (fixed summation order).  workspace: float[1185] zero-initialised once by the caller. */
int pidm_sumsq(const float* x, long long n, float* out, float* workspace, void* stream);
/* Adam (torch.optim.Adam semantics, bias corrections evaluated in double) + global-norm clip + EMA shadow, one pass.
 * step: 1-based host step count, ignored when step_counter_dev != NULL (device counter, incremented by the call).
 * ema_first_step: 0 = no EMA update; k >= 1 = update the shadow from the k-th step on (main.py:178 => ema_start+2). */
int pidm_adam_ema_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, float* ema_shadow, long long n,
                       float lr, double beta1, double beta2, float eps, int step, int* step_counter_dev,
                       const float* grad_norm_sq_dev, float grad_scale, float max_norm, float ema_mu,
                       int ema_first_step, int zero_grad, void* stream);

/* ---- mechanics residual, matrix-free (src/residuals_mechanics_K.py:166-274) ------------------------------ */
/* u [B,2,65,65] nodal displacements, rho [B,64,64], bcs [B,4,65,65] = (bc_x, bc_y, load_x, load_y), KE [8,8].
 * residual [B,8450] = K(rho) u - f with BC rows replaced by identity rows; compliance [B] = u^T K u. */
int pidm_mechanics_residual_fwd(const float* u, const float* rho, const float* bcs, const float* KE, float* residual,
                                float* compliance, int B, int nel, void* stream);
int pidm_mechanics_residual_bwd(const float* u, const float* rho, const float* bcs, const float* KE,
                                const float* grad_residual, const float* grad_compliance, float* grad_u, float* grad_rho,
                                float* workspace /* float[B*2*(nel+1)^2] */, int B, int nel, void* stream);
/* Fused PIDM loss of the mechanics branch + its gradients (src/denoising_utils.py:669-710), one launch:
 * u [B,2,n] displacements on the (nel+1)^2 node grid, rho [B,nel,nel], x0 [B,3,n] = (disp_x, disp_y, E) data target,
 * residual [B,2n], compliance [B], vf [B].  sums6 (zeroed here) = data, residual, inequality, optimisation loss terms,
 * mean|r|, mean_b(mean(rho_b) - vf_b).  grad_u / grad_rho / grad_residual / grad_compliance are overwritten. */
int pidm_mech_pidm_loss(const float* u, const float* rho, const float* x0, const float* residual, const float* compliance,
                        const float* vf, const long long* t, const float* p2_loss_weight,
                        const float* posterior_var_clipped, float c_data, float c_residual, float c_ineq,
                        float lambda_opt, float* sums6, float* grad_u, float* grad_rho, float* grad_residual,
                        float* grad_compliance, int B, int nel, void* stream);
/* bilinear resize, align_corners=False, antialias=False (resize_image, src/residuals_mechanics_K.py:10-21) */
int pidm_bilinear_resize_fwd(const float* x, float* y, int planes, int in, int out, void* stream);
int pidm_bilinear_resize_bwd(const float* dy, float* dx, int planes, int in, int out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PIDM_H_ */
