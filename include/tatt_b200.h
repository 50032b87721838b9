/* tatt_b200 -- C-ABI of the B200-native (sm_100a) TATT/TSRN hot path.
 *
 * Boundary contract: plain device pointers + sizes, the caller's CUDA stream as the
 * last argument (a cudaStream_t passed as void*), no torch types.  Every tensor is fp32, contiguous,
 * device-resident, caller-allocated and caller-owned; nothing is retained after the call returns
 * (all work is enqueued on `stream`).  Entry points return 0 on success and a non-zero code on
 * failure; `tatt_last_error()` then returns a thread-local message.  Nothing throws, nothing exits.
 * The library keeps no global mutable state and is re-entrant (DataParallel-style callers,
 * interfaces/base.py:390 in the reference, call replicas from several threads).
 *
 * The reference (mjq11302010044/TATT) is pure Python; it has no FFI.  Each group below names the
 * reference nn.Module / functional call (file:line under the reference tree) whose arithmetic it
 * replaces.  Feature maps are channels-last ("NHWC": [N][H][W][C]); token tensors are [N][L][64].
 */
#ifndef TATT_B200_H_
#define TATT_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

const char* tatt_last_error(void);
int tatt_version(void);
/* 1 when the library was compiled for sm_100a (always, in this tree) */
int tatt_arch(void);
/* Host-only census of a captured CUDA graph (cudaGraph_t as void*): counts[4] = {kernel, memcpy, memset, other} nodes.
 * bench.py reports counts[0] as the real number of kernel launches per replayed training step. */
int tatt_graph_node_counts(void* graph, int* counts);

/* ---- GEMM / linear layers: nn.Linear, 1x1 nn.Conv2d, GRU input/recurrent projections -------------
 * model/tsrn.py:170,1070 ; model/transformer_v2.py:453-458,785-790,177 ; model/stn_head.py:49-53
 * amode: 0 A[M][K] row-major, 1 A given as [K][M].  bmode: 0 B[K][N], 1 B given as [N][K].
 * flags: 1 accumulate into C, 2 ReLU epilogue, 4 split-K with atomics (C must be zero, or add 64 to
 * have a dense C zeroed here), 128 force the fp32 FFMA kernels instead of the tcgen05 bf16x3 path, 1024 bf16 mode
 * (operands rounded to bf16, one MMA per k-step, fp32 accumulate; also honoured by the conv entry points).
 * ws / ws_bytes: optional device scratch (16-byte aligned, >= 4*(|A|+|B|) bytes rounded up per row to 8
 * elements) for the pre-split bf16 operand planes of the v2 tcgen05 engine; NULL selects the in-loop split.  batch > 1 strides A/B/C/bias by sA/sB/sC/sBias elements. */
int tatt_gemm(int amode, int bmode, const float* A, long long lda, const float* B, long long ldb, float* C,
              long long ldc, const float* bias, int M, int N, int K, int batch, long long sA, long long sB,
              long long sC, long long sBias, int flags, long long loA, long long loB, void* ws, long long ws_bytes,
              void* stream);
/* Pre-split bf16 hi/lo operand planes (for operands reused by many GEMMs, e.g. W_hh over the RPE recurrence):
 * planes[rows][round8(cols)], or with transpose=1 planes[cols][round8(rows)].  Pass them to tatt_gemm with
 * flags 256 (A is the hi plane, lo plane at A + loA bf16 elements, lda/sA in plane elements) / 512 (same for B);
 * operand pairs (amode 0, bmode 1) or (amode 1, bmode 0).  colsum (optional, non-transposed only): out[c] =
 * sum_r src[r][c], i.e. the bias gradient for free while the gradient matrix is being split. */
int tatt_split_bf16(const float* src, long long ld, long long rows, int cols, int transpose, void* hi, void* lo,
                    float* colsum, void* stream);
/* out[c] (+)= sum_r X[r*ldx + c] */
int tatt_colsum(const float* X, long long ldx, float* out, long long P, int C, int zero_first, void* stream);

/* ---- per-pixel linear layers with the operand split fused into the loader (tc4_rows.cu) ----------
 * GruBlock.conv1 (1x1 conv over the channel concatenation, model/tsrn.py:902,1075), nn.GRU weight_ih
 * (tsrn.py:1071), attention in/out projections and FFN of the TP Interpreter
 * (model/transformer_v2.py:453-458,785-790): Y[M][N] (=|+=) act(sum_kb X_kb[M][64] W[:, 64kb:64kb+64]^T + bias)
 * over M pixel / token rows, fp32 in HBM, read once.  KB = 1..3 K blocks of 64 columns, each from its own
 * tensor (row strides ldx*; the concatenation never exists in memory); N in {64,128,192}, KB*N <= 192.
 * wtrans = 0: W[N][64 KB] (row stride ldw); 1: W given as [64 KB][N] (data gradient: Y = dY W).
 * flags: 1 accumulate into Y, 2 ReLU, 1024 bf16 mode (see tatt_gemm). */
int tatt_rows_gemm(const float* X0, long long ldx0, const float* X1, long long ldx1, const float* X2, long long ldx2,
                   const float* W, long long ldw, int wtrans, const float* bias, float* Y, long long ldy, long long M,
                   int N, int KB, int flags, void* stream);
/* Weight gradients of the same layers: D[64][64 NB] = A[M][64]^T B[M][64 NB], reduction over the M rows.
 * A's 64 columns are two 32-column segments a_off0 / a_off1 of its rows (a_off1 = a_off0 + 32 for a plain
 * matrix; the saved h_{t-1} of both GRU directions are two separate segments of the gate tensor);
 * B0..B2 are NB column blocks of 64.  out[b][i][j] = D[b rb + (T ? j : i)][b cb + (T ? i : j)] for b < nb,
 * i < ni, j < nj (T = transpose): e.g. dW[N][K] of a linear layer with A = dY, B = X (T = 0), or
 * dW_ih[192][64] with A = X, B = dGI (T = 1).  colsum_src 1 / 2: dbias[n] = column sums of A / of B (the bias
 * gradient), n < nbias.  ws: scratch of tatt_rows_wgrad_ws_bytes() bytes for per-CTA partial tiles. */
int tatt_rows_wgrad_ws_bytes(void);
int tatt_rows_wgrad(const float* A, long long lda, int a_off0, int a_off1, const float* B0, long long ldb0,
                    const float* B1, long long ldb1, const float* B2, long long ldb2, int NB, long long M,
                    int colsum_src, float* out, int nb, int ni, int nj, int transpose, int rb, int cb, float* dbias,
                    int nbias, void* ws, long long ws_bytes, int flags, void* stream);

/* ---- convolution (stride 1), implicit GEMM over NHWC: nn.Conv2d ------------------------------------
 * model/tsrn.py:597 (9x9 stem), 876,884 (SRB 3x3), 611 (block7), 1043 (upsample 64->256), 623 (9x9 out);
 * model/stn_head.py:15.  Weights are first packed to Wt[(ky,kx,ci)][co] (flip=0) or, for the
 * data-gradient, to the flipped/transposed Wt[(ky,kx,co)][ci] (flip=1).
 * flags of tatt_conv2d_igemm / tatt_conv2d_wgrad: 1 / 2 / 128 / 1024 as for tatt_gemm; 2048: the bf16 planes of X at the start of
 * `ws` are still valid from an earlier call on the same X (forward -> weight gradient, or the dY planes the weight-gradient
 * pass leaves behind the X planes -> data gradient with ws advanced past the X planes): skip the split pass.
 * 4096 (tatt_conv2d_wgrad, 3x3 / 64 input channels only): the dY planes behind the X planes were already written by the
 * caller (tatt_split_bf16 to ws + 4 * round8(|X|) bytes, lo plane round8(|dY|) elements later). */
int tatt_conv_weight_pack(const float* W, float* Wt, int Cout, int Cin, int KH, int KW, int CinP, int CoutP,
                          int flip, void* stream);
int tatt_conv_weight_unpack_grad(const float* dWt, float* dW, int Cout, int Cin, int KH, int KW, int CinP,
                                 int CoutP, void* stream);
int tatt_conv2d_igemm(const float* X, const float* Wt, const float* bias, float* Y, int nimg, int H, int W, int Cin,
                      int Cout, int KH, int KW, int padH, int padW, int flags, void* ws, long long ws_bytes,
                      void* stream);
int tatt_conv2d_wgrad(const float* X, const float* dY, float* dWt, int nimg, int H, int W, int Cin, int Cout,
                      int KH, int KW, int padH, int padW, int flags, void* ws, long long ws_bytes, void* stream);

/* 3x3 / 64 -> 64 convolution (tsrn.py:876,884,611) that also yields the BatchNorm statistics of its output from the same
 * pass: stats = TATT_CONV_STATS_ROWS rows of 128 floats, row c = {sum_ch[64], sumsq_ch[64]} over the pixels CTA c of the
 * persistent TMA kernel stored (zeroed here; no atomics); tatt_bn_finalize(nparts = TATT_CONV_STATS_ROWS) adds the rows in
 * double and produces mean / invstd (+ running statistics).  Served shapes only (tatt_conv3x3_stats_supported != 0:
 * W % 128 == 0, H % 2 == 0); flags 1024 / 2048 as for tatt_conv2d_igemm. */
#define TATT_CONV_STATS_ROWS 160
int tatt_conv3x3_stats_supported(int H, int W, int Cin, int Cout);
int tatt_conv3x3_stats(const float* X, const float* Wt, const float* bias, float* Y, int nimg, int H, int W, int flags,
                       void* ws, long long ws_bytes, void* stats, void* stream);

/* 9x9 convolution with <= 4 output channels (tsrn.py:623) as a K = KH*Cin, N = KW*CoP GEMM over the vertical
 * taps (tatt_conv2d_igemm / tatt_conv2d_wgrad with KW=1) plus these horizontal shift-sum / shift-expand passes */
int tatt_conv_kxexp_pack(const float* W, float* Wt, int Cout, int Cin, int KH, int KW, int CinP, int CoP,
                         void* stream);
int tatt_conv_kxexp_unpack_grad(const float* dWt, float* dW, int Cout, int Cin, int KH, int KW, int CinP, int CoP,
                                void* stream);
int tatt_conv_kxexp_reduce(const float* T, const float* bias, float* out, long long P, int W, int KW, int CoP,
                           int padW, void* stream);
int tatt_conv_kxexp_expand(const float* dOut, float* dT, long long P, int W, int KW, int CoP, int padW,
                           void* stream);

/* ---- BatchNorm (batch statistics over all rows of X[P][C]) with fused activation ----------------------
 * nn.BatchNorm2d/1d: model/tsrn.py:878,886,612 ; model/stn_head.py:19,51.  act: 0 none, 1 ReLU, 2 mish
 * (model/tsrn.py:1061-1064).  ws: scratch of >= 2*C doubles. */
int tatt_bn_stats(const float* X, long long P, int C, float eps, float momentum, float* mean, float* invstd,
                  float* running_mean, float* running_var, void* ws, void* stream);
/* parts: nparts rows of {sum[C], sumsq[C]} floats (per-CTA partial sums of a producing kernel), added in double */
int tatt_bn_finalize(const float* parts, int nparts, long long P, int C, float eps, float momentum, float* mean,
                     float* invstd, float* running_mean, float* running_var, void* stream);
int tatt_bn_eval_stats(const float* running_mean, const float* running_var, float eps, int C, float* mean,
                       float* invstd, void* stream);
int tatt_bn_apply_fwd(const float* X, float* Y, const float* mean, const float* invstd, const float* gamma,
                      const float* beta, int act, long long P, int C, void* stream);
/* the same normalisation + activation, written ONLY as bf16 hi / lo planes [P][C]: for a BatchNorm whose single consumer
 * is a convolution (tsrn.py:897: bn1 + mish -> conv2), the planes are that convolution's X operand (flag 2048) */
int tatt_bn_apply_planes(const float* X, void* y_hi, void* y_lo, const float* mean, const float* invstd, const float* gamma,
                         const float* beta, int act, long long P, int C, void* stream);
int tatt_bn_bwd(const float* X, const float* dY, const float* mean, const float* invstd, const float* gamma,
                const float* beta, int act, int training, long long P, int C, float* dX, float* dgamma,
                float* dbeta, void* ws, void* stream);
/* the same backward with dX written only as bf16 hi / lo planes [P][C] (operand format of the tcgen05 convolution kernels):
 * the BatchNorm behind a 3x3 convolution hands that convolution's backward passes their dY planes (flag 4096 of
 * tatt_conv2d_wgrad, flag 2048 of the data-gradient tatt_conv2d_igemm) -- no fp32 copy, no split pass */
int tatt_bn_bwd_planes(const float* X, const float* dY, const float* mean, const float* invstd, const float* gamma,
                       const float* beta, int act, int training, long long P, int C, void* dx_hi, void* dx_lo,
                       float* dgamma, float* dbeta, void* ws, void* stream);

/* ---- LayerNorm(64) with fused residual add: model/transformer_v2.py:460-461,792-794,166 -------------- */
int tatt_layernorm64_fwd(const float* X, const float* R, const float* gamma, const float* beta, float* Y, float* S,
                         float* mean, float* rstd, long long P, float eps, void* stream);
int tatt_layernorm64_bwd(const float* dY, const float* S, const float* mean, const float* rstd, const float* gamma,
                         float* dS, float* dgamma, float* dbeta, long long P, void* stream);

/* ---- BiGRU(hidden 32) scans on the NHWC map: GruBlock, model/tsrn.py:1067-1084 ------------------------
 * GI [rows][192], OUT [rows][64], GATES [rows][320] (saved for bwd), dGI/dGH [rows][192];
 * row(seq,t) = (seq / s_inner)*outer_stride + (seq % s_inner)*inner_stride + t*t_stride. */
int tatt_gru32_scan_fwd(const float* GI, const float* Whh, const float* bhh, float* OUT, float* GATES, int nseq,
                        int T, int s_inner, long long outer_stride, long long inner_stride, long long t_stride,
                        void* stream);
int tatt_gru32_scan_bwd(const float* dOUT, const float* GATES, const float* Whh, float* dGI, float* dGH, int nseq,
                        int T, int s_inner, long long outer_stride, long long inner_stride, long long t_stride,
                        void* stream);

/* ---- recurrent positional encoding (batch-axis BiGRU, quirk Q1): model/transformer_v2.py:177,215-221 -- */
/* Persistent recurrence (tc5_rpe.cu): ONE cooperative launch runs all N steps of both directions (weight-stationary
 * W_hh hi plane in shared memory, hidden state streamed as bf16 hi/lo planes by TMA, tcgen05 MMAs, gate math in the
 * epilogue, per-step grid barrier).  WPL: bf16 planes of W_hh [2][3Hd][Hd] (hi, lo plane w_lo elements later);
 * HPL: bf16 planes of HALL [2][N+1][Wd][Hd], step 0 zero, written by the kernel (lo plane h_lo elements later);
 * sync: tatt_rpe_sync_bytes(N) bytes of scratch.  tatt_rpe_persist_supported() == 0 when shape and device qualify
 * (Wd <= 128, Hd % 64 == 0, 2*Hd/16 CTAs co-resident); the per-step entry points below remain for other shapes. */
int tatt_rpe_persist_supported(int N, int Wd, int Hd, int C);
int tatt_rpe_sync_bytes(int N);
int tatt_rpe_fwd(const float* GI, const float* BHH, const void* WPL, long long w_lo, void* HPL, long long h_lo,
                 float* HALL, float* GATES, float* QPOS, void* sync, int N, int Wd, int Hd, int C, int Himg,
                 void* stream);
int tatt_rpe_gather(const float* emb, float* X, int H, int W, int C, void* stream);
int tatt_rpe_scatter(const float* dX, float* demb, int H, int W, int C, void* stream);
/* HPL / DGHPL: optional bf16 planes {hi, lo at +plane_lo elements} mirroring HALL / DGH (operands of the next
 * recurrent GEMM), or NULL */
int tatt_rpe_gate_fwd(const float* GI, const float* GH, float* HALL, float* GATES, float* QPOS, void* HPL,
                      long long plane_lo, int step, int N, int Wd, int Hd, int C, int Himg, void* stream);
int tatt_rpe_gate_bwd(const float* dQPOS, const float* HALL, const float* GATES, float* DH, float* DGISUM,
                      float* DGH, void* DGHPL, long long plane_lo, int step, int N, int Wd, int Hd, int C, int Himg,
                      void* stream);

/* ---- multi-head attention core (64 = 4 x 16, <= 32 keys): nn.MultiheadAttention as used at
 * model/transformer_v2.py:476-478 (encoder) and 820-823 (decoder cross-attention) ----------------------- */
int tatt_mha64_fwd(const float* Q, const float* K, const float* V, float* O, float* AW, int N, int Lq, int Lk,
                   float pdrop, const unsigned long long* rng, unsigned long long site, void* stream);
int tatt_mha64_bwd(const float* Q, const float* K, const float* V, const float* dO, float* dQ, float* dK, float* dV,
                   int N, int Lq, int Lk, float pdrop, const unsigned long long* rng, unsigned long long site,
                   void* stream);

/* Fused decoder layer (tc6_declayer.cu): ONE launch = TransformerDecoderLayer_TP.forward_post + the decoder's final
 * LayerNorm of that layer's output (model/transformer_v2.py:806-833, 380-390) for Lq %% 128 == 0 query tokens and
 * Lk <= 32 keys per sample; every contraction (Q / out / FFN projections, per-head QK^T and PV) on tcgen05.
 * in[18]  = tgt, query_pos [N*Lq][64]; projected K, V [N][Lk][64]; Wq, Wo, W1, W2 [64][64]; bq, bo, norm2.weight,
 *           norm2.bias, b1, b2, norm3.weight, norm3.bias, final norm weight, bias [64].
 * out[14] = out, inter [N*Lq][64]; head-averaged attention weights [N][Lq][Lk] or NULL; then, when train != 0, the
 *           tensors the backward pass consumes: qin, q, a, S2, t1, h1, h1d, S3 [N*Lq][64], stats2, stats3, statsF
 *           [2][N*Lq] (mean, rstd).  pdrop[4] / sites[4]: attention, dropout2, dropout (FFN hidden), dropout3 -- the masks
 *           equal those of tatt_dropout / tatt_mha64_* for the same rng state and sites. */
int tatt_tp_declayer_fwd(const float* const* in, float* const* out, int train, int N, int Lq, int Lk,
                         const float* pdrop, const unsigned long long* rng, const unsigned long long* sites,
                         void* stream);

/* ---- element-wise / layout ------------------------------------------------------------------------------
 * PReLU (tsrn.py:598,173), PixelShuffle(2)+mish (tsrn.py:1049-1053), tanh (tsrn.py:675), Dropout,
 * MaxPool2d (stn_head.py:34-44), residual adds. */
int tatt_axpby(const float* a, const float* b, float alpha, float beta, float* out, long long n, void* stream);
int tatt_add_bcast_rows(const float* a, const float* b, float* out, long long rows, long long period, int cols,
                        void* stream);
int tatt_prelu_fwd(const float* x, const float* w, float* y, long long n, void* stream);
int tatt_prelu_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw, long long n,
                   void* stream);
int tatt_pixshuf2_mish_fwd(const float* in, float* out, long long nimg, int H, int W, int C, void* stream);
/* PixelShuffle(2) + mish written as bf16 hi / lo planes [N*2H*2W][C]: the X operand of the convolution that follows
 * (tsrn.py:623 after 1049-1053) */
int tatt_pixshuf2_mish_planes(const float* in, void* out_hi, void* out_lo, long long nimg, int H, int W, int C,
                              void* stream);
int tatt_pixshuf2_mish_bwd(const float* in, const float* dout, float* din, long long nimg, int H, int W, int C,
                           void* stream);
int tatt_nchw_to_nhwc(const float* in, float* out, long long N, int C, int H, int W, int Cp, void* stream);
int tatt_nhwc_to_nchw(const float* in, float* out, long long N, int C, int H, int W, int Cp, int do_tanh,
                      void* stream);
int tatt_tanh_bwd_nchw_to_nhwc(const float* dout, const float* out, float* dpre, long long N, int C, int H, int W,
                               int Cp, void* stream);
/* rng: DEVICE pointer to {seed, counter} (graph-capture safe); site: per-call-site stream id */
int tatt_dropout(const float* x, float* y, long long n, float p, const unsigned long long* rng,
                 unsigned long long site, void* stream);
int tatt_rng_advance(unsigned long long* rng, void* stream);
int tatt_relu_bwd(const float* y, const float* dy, float* dx, long long n, void* stream);
int tatt_maxpool_fwd(const float* in, float* out, long long N, int H, int W, int C, int kh, int kw, void* stream);
int tatt_maxpool_bwd(const float* in, const float* dout, float* din, long long N, int H, int W, int C, int kh,
                     int kw, void* stream);

/* ---- TPS grid + bilinear grid_sample: model/tps_spatial_transformer.py:97-112,10-18 -------------------- */
int tatt_tps_sample_fwd(const float* X, const float* ctrl, const float* invK, const float* repr, float* OUT,
                        float* SRC, int N, int H, int W, void* stream);
int tatt_tps_sample_bwd(const float* X, const float* ctrl, const float* invK, const float* repr, const float* dOUT,
                        float* dctrl, int N, int H, int W, void* stream);

/* ---- gradient step on a flat buffer: clip_grad_norm_(0.25) + Adam, interfaces/super_resolution.py:1083-1085 */
int tatt_sqnorm(const float* x, long long n, float* out, int zero_first, void* stream);
/* the same sum with a fixed reduction order (per-block partials in ws[ws_floats], then one block): bit-identical on every
 * data-parallel replica, which all clip the same all-reduced gradient */
int tatt_sqnorm_det(const float* x, long long n, float* out, float* ws, int ws_floats, void* stream);
/* gradient packing in one launch: table = DEVICE array of nchunks x {source address, destination offset (floats),
 * count (floats, <= 16384)}; sq (optional) receives sum(x^2) over everything copied (zeroed first) */
int tatt_multi_copy(const unsigned long long* table, int nchunks, float* dst, float* sq, void* stream);
/* step_state: DEVICE {unused, step>=1} (advance it with tatt_rng_advance before the call) */
int tatt_adam_clip_step(float* p, const float* g, float* m, float* v, long long n, const float* sqnorm,
                        float max_norm, float lr, float beta1, float beta2, float eps,
                        const unsigned long long* step_state, float grad_scale, void* stream);

/* ---- plumbing ------------------------------------------------------------------------------------------- */
int tatt_memcpy_d2d(void* dst, const void* src, long long bytes, void* stream);
int tatt_memset0(void* dst, long long bytes, void* stream);

/* ---- image loss directly after the path: loss/image_loss.py:10-58 ----------------------
 * ImageLoss(gradient=True, loss_weight=[w0, w1]) on NCHW fp32 images with C >= 3 channels:
 * loss[n] = w0 * mean (out - tgt)^2 + w1 * mean_{RGB} |gradmag(out) - gradmag(tgt)|  (central differences, zero
 * padding, eps 1e-6).  G (optional, [N][3][H][W][2] floats) receives the per-pixel gradient-magnitude derivatives the
 * backward pass gathers; ws: >= 2*N doubles.  bwd: dout = gloss[n] * d loss[n] / d out. */
int tatt_image_loss_fwd(const float* out, const float* tgt, float* loss, float* G, int N, int C, int H, int W, float w0,
                        float w1, void* ws, void* stream);
int tatt_image_loss_bwd(const float* out, const float* tgt, const float* G, const float* gloss, float* dout, int N, int C,
                        int H, int W, float w0, float w1, void* stream);

/* ---- rest of the loss block -----------------------------------------------------------------
 * SemanticLoss.forward, loss/semantic_loss.py:20-37: loss[0] = mean |gt - pred| + mean t (log t - log(pred + 1e-20)),
 * t = gt + 1e-20 (nn.KLDivLoss default 'mean' reduction = over ALL n elements).  ws: >= 2 doubles.
 * bwd: dpred / dgt (either may be NULL) = gloss[0] * d loss / d pred|gt. */
int tatt_semantic_loss_fwd(const float* pred, const float* gt, long long n, float* loss, void* ws, void* stream);
int tatt_semantic_loss_bwd(const float* pred, const float* gt, const float* gloss, float* dpred, float* dgt, long long n,
                           void* stream);
/* TRI_SSIM.forward -> _tri_ssim, utils/ssim_psnr.py:28-37,99-128,231-256: three NCHW images, depthwise 11x11 Gaussian
 * (sigma 1.5) with zero padding, C1 = 0.01^2, C2 = 0.03^2.  per_sample 0: out[0] = mean of the ssim map (size_average);
 * 1: out[n] = per-sample mean.  G (optional, 5*N*C*H*W floats) receives the per-pixel derivatives the backward pass
 * blurs; ws: >= N doubles.  bwd: d1 / d2 / d3 (any may be NULL) = gout * d out / d x1|x2|x3, gout [1] or [N]. */
int tatt_tri_ssim_fwd(const float* x1, const float* x2, const float* x3, float* out, float* G, int N, int C, int H, int W,
                      int per_sample, void* ws, void* stream);
int tatt_tri_ssim_bwd(const float* x1, const float* x2, const float* x3, const float* G, const float* gout, float* d1,
                      float* d2, float* d3, int N, int C, int H, int W, int per_sample, void* stream);
/* TextSR.torch_rotate_img, interfaces/super_resolution.py:126-157: per-sample rotation by arcs[n] with the aspect term
 * H/W + offs[n]*2*off_range - off_range, as affine_grid + grid_sample (bilinear, zeros, align_corners=False), NCHW.
 * bwd: dimg = adjoint (scatter) of the same sampling applied to dout (dimg is zero-filled here). */
int tatt_rotate_img_fwd(const float* img, const float* arcs, const float* offs, float off_range, float* out, int N, int C,
                        int H, int W, void* stream);
int tatt_rotate_img_bwd(const float* dout, const float* arcs, const float* offs, float off_range, float* dimg, int N, int C,
                        int H, int W, void* stream);

/* ---- CRNN text-prior generator in front of the path: model/crnn/crnn.py:5-93 ----------------------
 * Its convolutions / BatchNorm+ReLU / linear layers / per-step recurrent GEMMs use tatt_conv2d_*, tatt_bn_*, tatt_gemm.
 * parse_crnn_data (interfaces/base.py:797-815): NCHW image (C >= 3) -> bicubic resize (torch upsample_bicubic2d,
 * align_corners=False, A = -0.75) -> 0.299 R + 0.587 G + 0.114 B, out [N][OH][OW]. */
int tatt_bicubic_gray(const float* img, float* out, int N, int C, int H, int W, int OH, int OW, void* stream);
/* nn.MaxPool2d(kernel, stride, padding) on NHWC maps, floor mode (crnn.py:56-66; the (2,2),(2,1),(0,1) windows overlap).
 * bwd: din = gradient routed to each window's first arg-max (torch's tie rule); gather form, no atomics. */
int tatt_maxpool2d_fwd(const float* in, float* out, long long N, int H, int W, int C, int kh, int kw, int sh, int sw, int ph,
                       int pw, void* stream);
int tatt_maxpool2d_bwd(const float* in, const float* dout, float* din, long long N, int H, int W, int C, int kh, int kw,
                       int sh, int sw, int ph, int pw, void* stream);
/* pad 0: out [N][OH][OW][C] = top-left crop of in [N][H][W][C]; pad 1: out [N][H][W][C] = in [N][OH][OW][C] zero-padded
 * (conv6 of crnn.py:68 is a 2x2 convolution without padding: same-size convolution + crop) */
int tatt_crop_nhwc(const float* in, float* out, long long N, int H, int W, int OH, int OW, int C, int pad, void* stream);
/* [A][B][C] -> [B][A][C] (crnn.py:85-86 permute(2, 0, 1) of the squeezed feature map; logits back to [T, N, C]) */
int tatt_permute_102(const float* in, float* out, int A, int B, int C, void* stream);
/* One time step of nn.LSTM(bidirectional=True) (crnn.py:9,19), both directions: step s handles time s (forward) and
 * T-1-s (reverse).  G [T*Nb][8H]: rows t*Nb+n, columns [dir][i f g o][H]; W_ih x + b_ih on entry, activated gates on exit.
 * GH [2][Nb][4H] = W_hh h_prev + b_hh (NULL at s == 0: BHH [2][4H] is used), CS [2][T][Nb][H] cell states,
 * OUT [T*Nb][2H] = [h_fwd | h_bwd].  bwd (steps T-1 .. 0): dG = pre-activation gate gradients of this step's rows,
 * DH [2][Nb][H] = recurrent hidden gradient (ignored at s == T-1), DC [2][Nb][H] cell gradient (in/out). */
int tatt_lstm_gate_fwd(float* G, const float* GH, const float* BHH, float* CS, float* OUT, int s, int T, int Nb, int H,
                       void* stream);
int tatt_lstm_gate_bwd(const float* G, const float* CS, const float* dOUT, const float* DH, float* DC, float* dG, int s,
                       int T, int Nb, int H, void* stream);
/* interfaces/super_resolution.py:796-799: probs [T*Nb][C] = softmax over the C <= 64 classes of logits [T][Nb][C];
 * prior (optional) [Nb][C][1][T] = the permuted text prior the SR model consumes.  bwd: dlogits from dprobs. */
int tatt_softmax_prior_fwd(const float* logits, float* probs, float* prior, int T, int Nb, int C, void* stream);
int tatt_softmax_bwd(const float* probs, const float* dprobs, float* dlogits, long long R, int C, void* stream);

/* ---- TextZoom collate, device half: dataset/dataset.py:1266-1319 (resizeNormalize after the PIL resize) --
 * img: uint8 [N][H][W][3] (RGB, HWC) -> out fp32 [N][3 + with_mask][H][W] = ToTensor() (x / 255) and, with_mask != 0, the
 * mean-threshold mask channel (1 where PIL luma L <= mean(L), else 0).  Bit-exact integer / IEEE arithmetic. */
int tatt_collate_u8(const void* img, float* out, int N, int H, int W, int with_mask, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TATT_B200_H_ */
