/*
 * npf_b200.h -- C ABI of libnpf_b200.so: the B200 (sm_100a) kernels behind the Neural-Process hot path.
 *
 * The reference (YannDubs/Neural-Process-Family) has no FFI: its operator boundary is the Python
 * sub-module factory protocol (tier 2).  Each entry point below replaces the ATen
 * op sequence of one reference function (cited as upstream file:line) and is what a binding of that
 * function would call (see INTEGRATION.md for the ctypes stubs).
 *
 * Conventions (all entry points):
 *   - C linkage, POD arguments only, no torch types.
 *   - every pointer is a CALLER-OWNED DEVICE pointer to contiguous fp32 unless stated otherwise;
 *     leading dimensions (ld*) are in elements.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).  No allocation, no
 *     synchronisation and no global mutable state inside => re-entrant per stream.  (Kernel function attributes --
 *     the opt-in to > 48 KB of dynamic shared memory -- are configured once per process on first use: the library
 *     follows the one-process-per-GPU model of the data-parallel design.)
 *   - returns 0 on success, a negative NPF_E* code otherwise; npf_last_error() gives the text
 *     (thread-local).  Kernel launch errors are reported via cudaGetLastError() after the launch.
 *   - "accumulate" outputs (weight / bias / theta gradients) are ADDED to: zero them first.
 *   - fp32 storage everywhere.  `precision` selects the arithmetic of GEMM-shaped inner products:
 *       NPF_PREC_FP32  : fp32 FFMA                                   (parity bar 1e-4 rel)
 *       NPF_PREC_BF16  : bf16 tensor-core operands, fp32 accumulate   (parity bar 1e-2 rel)
 *       NPF_PREC_BF16X3: 3-term split-bf16 tensor-core, fp32 accumulate (parity bar 1e-4 rel)
 */
#ifndef NPF_B200_H
#define NPF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NPF_ABI_VERSION 1

/* error codes */
#define NPF_OK 0
#define NPF_EINVAL (-1)   /* bad argument (shape, alignment, null pointer)        */
#define NPF_ECUDA (-2)    /* CUDA runtime / launch error                          */
#define NPF_ENOTSUP (-3)  /* valid request this build does not implement          */

/* precision of GEMM-shaped inner products */
#define NPF_PREC_FP32 0
#define NPF_PREC_BF16 1
#define NPF_PREC_BF16X3 2

/* flags */
#define NPF_RELU_OUT 1   /* apply relu to the output                               */
#define NPF_RELU_IN 2    /* apply relu to the (first / activation) input on load   */
#define NPF_ACCUM 4      /* add to the output instead of overwriting it            */
#define NPF_ADD_DY 8     /* npf_dwconv_bwd: dX += dY (gradient of a residual branch that reads the same X) */
#define NPF_MASK_X 16    /* npf_linear_bwd: multiply dX by the relu mask (X > 0) of the layer input X  */

typedef void* npf_stream_t;

#if defined(__GNUC__)
#define NPF_API __attribute__((visibility("default")))
#else
#define NPF_API
#endif

/* ---- multi-GPU: one-shot mean all-reduce of the flat gradient bucket over NVLink peer memory (one process per GPU) ----------
 * npf_p2p_alloc / _get_handle / _open / _close wrap cudaMalloc and the CUDA IPC handle calls so that the host language needs no
 * CUDA binding of its own: every rank allocates its bucket, output-less signal block (2 * world ints) and exchanges the 64-byte
 * handles through its own channel (torch.distributed here), then opens the peers' handles once.
 * npf_allreduce_mean_p2p: `in` / `sig` are HOST arrays of `world` device pointers (in[rank] / sig[rank] are this rank's own
 * allocations, the others peer mappings); out[n] = mean_r in_r[n]; `state` = 2 zero-initialised ints of local device memory.
 * The kernel carries its own inter-rank barriers (epoch flags in the signal blocks): the peers' buckets may be read as soon as it
 * starts and this rank's bucket may be overwritten as soon as it ends.  Same launch every step (CUDA-graph capturable). */
NPF_API int npf_p2p_alloc(void** ptr, size_t bytes);
NPF_API int npf_p2p_free(void* ptr);
NPF_API int npf_p2p_get_handle(void* ptr, unsigned char* handle64);
NPF_API int npf_p2p_open(const unsigned char* handle64, void** ptr);
NPF_API int npf_p2p_close(void* ptr);
NPF_API int npf_allreduce_mean_p2p(const float* const* in, int* const* sig, float* out, int* state, int rank, int world, long n,
                           npf_stream_t stream);
/* Two-shot variant: rank r reduces slice r only and writes its mean into EVERY rank's output buffer (`out` = HOST array of `world`
 * device pointers, out[rank] local); 2 (world - 1) / world bucket sizes cross NVLink per rank instead of (world - 1). */
NPF_API int npf_allreduce_mean_p2p2(const float* const* in, int* const* sig, float* const* out, int* state, int rank, int world, long n,
                            npf_stream_t stream);

/* Diagnostics only: device buffer (>= 4096 uint64, zeroed by the caller) into which CTA 0 of the kernels that support it
 * (npf_mlp_chain_bwd, npf_setconv_bwd tensor-core path) writes per-role (event, SM clock) records; NULL disables. */
NPF_API int npf_debug_set_trace(unsigned long long* device_buffer);
NPF_API int npf_abi_version(void);
NPF_API const char* npf_last_error(void);
/* number of kernels launched by this library in the calling process so far (monotonic; for bench.py) */
NPF_API unsigned long long npf_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Linear layers  (nn.Linear inside MLP.forward, npf/architectures/mlp.py:95-109; SetConv resizer
 * npf/architectures/setcnn.py:268; 1x1 "pointwise" convs of ResConvBlock, npf/architectures/cnn.py:214;
 * Q/K/V projections npf/architectures/attention.py:472-474)
 * ------------------------------------------------------------------------------------------------ */

/* Y[M,N] = act_out( act_in(X)[M,K] . W[N,K]^T  (+ u[M] (x) w2[N])  + b[N] )
 * b, u/w2 optional (NULL).  u (x) w2 is the rank-1 "density column" of SetConv's Linear(in+1 -> out). */
NPF_API int npf_linear_fwd(const float* X, int ldx, const float* W, int ldw, const float* b, float* Y, int ldy,
                   int M, int K, int N, int flags, const float* u, const float* w2, int ldw2, int precision,
                   npf_stream_t stream);

/* dX[M,K] = (dY[M,N] . W[N,K]) (.) (mask_src[M,K] > 0)      mask_src optional; NPF_ACCUM honoured.
 * Chains of Linear->ReLU pass the PRE-activation gradient between layers: mask_src is the layer's own
 * (post-relu) input. */
NPF_API int npf_linear_bwd_data(const float* dY, int lddy, const float* W, int ldw, float* dX, int lddx, int M, int K,
                        int N, const float* mask_src, int ldm, int flags, int precision, npf_stream_t stream);

/* Whole backward of one Linear in a single pass over dY and X (the layer's saved input):
 *     dX[M,K] = (dY[M,N] . W[N,K]) (.) (X > 0 if NPF_MASK_X)      (overwritten; NULL to skip)
 *     dW[N,K] += dY^T . act_in(X)      db[N] += colsum(dY)        (db optional)
 * Same results as npf_linear_bwd_weight followed by npf_linear_bwd_data(mask_src = X); for 128 -> 128 layers in the
 * tensor-core precisions, and for thin layers (K <= 8 or N <= 8 against a 128-wide side) in every precision, dY and X
 * are read from HBM once instead of twice. */
NPF_API int npf_linear_bwd(const float* dY, int lddy, const float* X, int ldx, const float* W, int ldw, float* dX, int lddx,
                   float* dW, int lddw, float* db, int M, int K, int N, int flags, int precision, npf_stream_t stream);

/* L consecutive square layers of one MLP (mlp.py:95-109, the hidden stack):  H_0 = act_in(X),
 *     Y_l = act_l( H_l . W_l^T + b_l ),  H_{l+1} = Y_l,   act_l = relu if bit l of relu_mask else identity,
 * every Y_l [M,width] stored (they are the activations saved for backward).  W / b / Y are HOST arrays of L device
 * pointers (b may be NULL, or hold NULL entries); W_l is [width,width] row-major contiguous, Y_l contiguous.
 * For width 128 in the tensor-core precisions and M <= 37 888 rows the row block stays on chip between layers
 * (one kernel: no intermediate activation is read back); otherwise identical to L npf_linear_fwd calls. */
NPF_API int npf_mlp_chain_fwd(const float* X, int ldx, const float* const* W, const float* const* b, float* const* Y, int L, int M,
                      int width, int relu_in, unsigned relu_mask, int precision, npf_stream_t stream);

/* Whole backward of the same L square layers (replaces L npf_linear_bwd calls; upstream mlp.py:95-109 through autograd):
 *     dZ_{L-1} = dY  (gradient w.r.t. the PRE-activation output of the last layer);  for l = L-1 .. 0:
 *     dW_l += dZ_l^T . X_l ,  db_l += colsum(dZ_l) ,  dZ_{l-1} = (dZ_l . W_l) (.) (X_l > 0)
 * with X_l the saved INPUT of layer l (X_0 = chain input, X_l = Y_{l-1} of npf_mlp_chain_fwd: post-ReLU, so every
 * intermediate layer is masked); dX (optional) = dZ_{-1}, masked with X_0 > 0 only if flags has NPF_MASK_X.
 * X / W / dW / db are HOST arrays of L device pointers (db may be NULL or hold NULL entries), all [.,width] contiguous.
 * The gradient stays on chip between layers: HBM sees dY and every X_l once and dX once.  Covered: width 128, tensor-core
 * precisions, M >= 64; returns NPF_ENOTSUP otherwise (no scratch memory is owned here: call npf_linear_bwd per layer). */
NPF_API int npf_mlp_chain_bwd(const float* dY, int lddy, const float* const* X, const float* const* W, float* dX, int lddx,
                      float* const* dW, float* const* db, int L, int M, int width, int flags, int precision, npf_stream_t stream);

/* dW[N,K] += dY[M,N]^T . act_in(X)[M,K] ;  db[N] += sum_m dY[m,:] ;  dw2[N*ldw2] += sum_m dY[m,:] u[m]
 * db, u/dw2 optional. */
NPF_API int npf_linear_bwd_weight(const float* dY, int lddy, const float* X, int ldx, float* dW, int lddw, float* db,
                          int M, int K, int N, int flags, const float* u, float* dw2, int ldw2, int precision,
                          npf_stream_t stream);

/* dZ = dH (.) (H > 0)   (stand-alone ReLU backward, n elements) */
NPF_API int npf_relu_bwd(const float* dH, const float* H, float* dZ, long n, npf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * SetConv with the exponential-quadratic RBF  (SetConv.forward npf/architectures/setcnn.py:234-268,
 * ExpRBF.forward :126-142).  x_dim == 1 (asserted upstream, :226).
 *   sigma = 1e-5 + softplus(theta);  a_qk = -((x_q - x_k)/sigma)^2
 *   feat[b,q,:] = sum_k softmax_k(a_qk) values[b,k,:]      dens[b,q] = sum_k exp(a_qk)
 * keys [B,K] with batch stride key_bs (0 => one key set shared by all tasks, e.g. the induced grid);
 * queries [B,Q] with batch stride qry_bs (0 => shared).  If keys_regular != 0 the keys are an
 * increasing, uniformly spaced grid and the kernel only visits the run-time sigma-window of keys whose
 * softmax weight is not below 2^-60 of the largest one (exact in fp32); otherwise all keys are visited.
 * mstat[B,Q,2] receives (max logit, sum_k exp(a - max)) per query (saved for backward).
 * ldf / ldd: row stride of feat (and dfeat) / element stride of dens (and ddens), in floats.  Cin and 1 for dense
 * buffers; the few-channel path (Cin <= 4, irregular keys) also takes ldf = ldd = Cin + 1 with dens = feat + Cin, i.e.
 * [feat | dens] interleaved as the [B*Q, Cin+1] input of SetConv's resizer Linear (no concatenation pass).
 * ------------------------------------------------------------------------------------------------ */
NPF_API int npf_setconv_fwd(const float* keys, long key_bs, const float* queries, long qry_bs, const float* values,
                    const float* theta, float* feat, float* dens, float* mstat, int B, int K, int Q, int Cin,
                    int keys_regular, int ldf, int ldd, npf_stream_t stream);

/* Given dfeat[B,Q,Cin], ddens[B,Q]:  dvalues[B,K,Cin] (overwritten; may be NULL) and dtheta[1] (+=). */
NPF_API int npf_setconv_bwd(const float* keys, long key_bs, const float* queries, long qry_bs, const float* values,
                    const float* theta, const float* feat, const float* dens, const float* mstat,
                    const float* dfeat, const float* ddens, float* dvalues, float* dtheta, int B, int K, int Q,
                    int Cin, int keys_regular, int ldf, int ldd, npf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Depthwise convolution, channel-last, zero padding k/2  (depthwise half of make_depth_sep_conv,
 * npf/utils/helpers.py:354-403, and conv2_depthwise of ResConvBlock.forward npf/architectures/cnn.py:204-215;
 * the CNN wrapper's channel permutes :363-370 disappear because we stay channel-last).
 *   Y[b,h,w,c] = sum_{i,j} Wt[c,i,j] * act(X)[b,h+i-kh/2,w+j-kw/2,c] + bias[c] (+ res[b,h,w,c])
 *   act(x) = relu(pre_scale[c]*x + pre_shift[c])  if NPF_RELU_IN (pre_scale/shift optional = folded
 *   BatchNorm affine of norm1/norm2), identity otherwise.   1-D signals: H = 1, kh = 1.
 * ------------------------------------------------------------------------------------------------ */
NPF_API int npf_dwconv_fwd(const float* X, const float* Wt, const float* bias, const float* res, float* Y, int B, int H,
                   int Wd, int C, int kh, int kw, int flags, const float* pre_scale, const float* pre_shift,
                   npf_stream_t stream);

/* dX (overwritten, or += with NPF_ACCUM; NULL to skip), dWt[C,kh,kw] (+=), dbias[C] (+=).
 * dX excludes the residual branch unless NPF_ADD_DY is set (then dX = conv^T(dY) * act' + dY, for blocks whose
 * residual is the conv input itself).  If pre_scale is
 * given, dX is the gradient w.r.t. X through the affine (i.e. multiplied by pre_scale[c]), and
 * dpre_scale[C]/dpre_shift[C] (+=, optional) receive the affine's gradients. */
NPF_API int npf_dwconv_bwd(const float* dY, const float* X, const float* Wt, float* dX, float* dWt, float* dbias, int B,
                   int H, int Wd, int C, int kh, int kw, int flags, const float* pre_scale,
                   const float* pre_shift, float* dpre_scale, float* dpre_shift, npf_stream_t stream);

/* The whole 1-D pre-activation residual block of the ConvCNP CNN in ONE kernel (upstream cnn.py:204-215 with
 * n_conv_layers = 1 and Normalization = Identity, helpers.py:354-403):
 *     O = depthwise_k(relu(X)) + bdw + X        Y = O . wpw^T + bpw
 * X, O, Y [B,L,128] channel-last contiguous; wdw [128,k] (the Conv1d weight [128,1,k]), wpw [128,128] (the 1x1 Conv1d
 * weight); O = NULL skips the write of the intermediate (a backward that recomputes it does not need it).
 * Raw rows reach shared memory by TMA bulk copies; O never makes a round trip through HBM.
 * Covered: C = 128, k = 11, NPF_PREC_BF16X3; NPF_ENOTSUP otherwise (run npf_dwconv_fwd + npf_linear_fwd). */
NPF_API int npf_resblock1d_fwd(const float* X, const float* wdw, const float* bdw, const float* wpw, const float* bpw, float* O,
                       float* Y, int B, int L, int C, int k, int precision, npf_stream_t stream);

/* Backward of npf_resblock1d_fwd in ONE kernel with O recomputed from X (so the forward may pass O = NULL):
 *     dX (overwritten) , dWdw [128,k] += , dbdw [128] += (optional) , dWpw [128,128] += , dbpw [128] += (optional)
 * HBM sees dY and X once and dX once; neither dO nor O exists in memory.  Same coverage as the forward. */
NPF_API int npf_resblock1d_bwd(const float* dY, const float* X, const float* wdw, const float* bdw, const float* wpw, float* dX,
                       float* dWdw, float* dbdw, float* dWpw, float* dbpw, int B, int L, int C, int k, int precision,
                       npf_stream_t stream);

/* per-channel batch statistics of a channel-last tensor X[M,C] (train-mode BatchNorm of the notebook CNN
 * configs, two-pass like ATen: first the mean, then the centred second moment):
 *   sum[C] += sum_m (x - center[c]) ;  sumsq[C] += sum_m (x - center[c])^2      center optional (NULL = 0)
 * these vectors are what a multi-GPU run would all-reduce for synchronised BatchNorm. */
NPF_API int npf_channel_stats(const float* X, const float* center, float* sum, float* sumsq, long M, int C,
                      npf_stream_t stream);
/* Y[m,c] (+)= a[c] * X[m,c] + b[c]   (BatchNorm backward through the batch statistics) */
NPF_API int npf_channel_affine(const float* X, const float* a, const float* b, float* Y, long M, int C, int accumulate,
                       npf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * On-grid context encoding  (GridConvCNP.cntxt_to_induced npf/neuralproc/gridconvnp.py:136-162 with the
 * abs-weight depthwise conv of npf/utils/helpers.py:316-331)
 *   sig = conv_|w|(Y * M), den = conv_|w|(M);   feat[b,h,w,:] = [sig / max(den, 1e-5) (y) ; den (y)]
 * img [B,H,W,y] fp32, mask [B,H,W,mc] uint8 (mc = 1 or y), Wt [y,k,k] (raw weights; abs applied here).
 * ------------------------------------------------------------------------------------------------ */
NPF_API int npf_gridconv_in_fwd(const float* img, const uint8_t* mask, int mc, const float* Wt, float* feat, int B, int H,
                        int Wd, int y, int k, npf_stream_t stream);
/* dWt[y,k,k] += gradient through sig, den and abs(), given dfeat[B,H,W,2y] */
NPF_API int npf_gridconv_in_bwd(const float* img, const uint8_t* mask, int mc, const float* Wt, const float* feat,
                        const float* dfeat, float* dWt, int B, int H, int Wd, int y, int k, npf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Sum-merge, pooling, normalisation glue
 * ------------------------------------------------------------------------------------------------ */
/* MergeFlatInputs sum-merge (npf/architectures/encoders.py:175-183) with the broadcasts of
 * CNP.trgt_dependent_representation (npf/neuralproc/np.py:103-110):
 *   out[z,b,t,:] = relu(x1[b,t,:] + x2[z,b,(x2_has_t ? t : 0),:])         x1 [B,T,C], x2 [Z,B,(T|1),C] */
NPF_API int npf_merge_relu_fwd(const float* x1, const float* x2, float* out, int Z, int B, int T, int C, int x2_has_t,
                       npf_stream_t stream);
/* dx1[B,T,C] (overwritten; NULL to skip) = sum_z dpre, dx2 (overwritten) = (sum_t) dpre, dpre = dout (.) (out>0) */
NPF_API int npf_merge_relu_bwd(const float* dout, const float* out, float* dx1, float* dx2, int Z, int B, int T, int C,
                       int x2_has_t, npf_stream_t stream);

/* mean over the middle axis: R[b,:] = mean_n X[b,n,:]  (CNP.encode_globally npf/neuralproc/np.py:95) */
NPF_API int npf_mean_pool_fwd(const float* X, float* R, int B, int N, int C, npf_stream_t stream);
/* dX[b,n,:] = dR[b,:] / N */
NPF_API int npf_mean_pool_bwd(const float* dR, float* dX, int B, int N, int C, npf_stream_t stream);

/* Y = LayerNorm(A + Bm) * gamma + beta over the last dim C, eps = 1e-5
 * (TransformerAttender.forward npf/architectures/attention.py:585-586).  rstat[M,2] = (mean, rstd). */
NPF_API int npf_add_layernorm_fwd(const float* A, const float* Bm, const float* gamma, const float* beta, float* Y,
                          float* rstat, long M, int C, npf_stream_t stream);
/* dS[M,C] = gradient w.r.t. the sum (A + Bm); dgamma[C], dbeta[C] (+=) */
NPF_API int npf_add_layernorm_bwd(const float* dY, const float* A, const float* Bm, const float* gamma,
                          const float* rstat, float* dS, float* dgamma, float* dbeta, long M, int C,
                          npf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Multi-head scaled-dot cross-attention of targets over context
 * (BaseAttender.forward / DotAttender.score npf/architectures/attention.py:129-156, 204-220; head split
 * of MultiheadAttender :507-527: head h owns channels [h*D, (h+1)*D) -- no permute is materialised).
 *   O[b,t,h,:] = sum_c softmax_c(Q[b,t,h,:].K[b,c,h,:] * scale) V[b,c,h,:]
 * Q [B,Tq,H*D], K [B,Tk,H*D], V [B,Tk,H*Dv], O [B,Tq,H*Dv], LSE [B,H,Tq] (log-sum-exp, saved).
 * The [Tq,Tk] logits are never written to memory.
 * ------------------------------------------------------------------------------------------------ */
NPF_API int npf_xattn_fwd(const float* Q, const float* K, const float* V, float* O, float* LSE, int B, int Tq, int Tk,
                  int H, int D, int Dv, float scale, int precision, npf_stream_t stream);
NPF_API int npf_xattn_bwd(const float* Q, const float* K, const float* V, const float* O, const float* LSE,
                  const float* dO, float* dQ, float* dK, float* dV, int B, int Tq, int Tk, int H, int D, int Dv,
                  float scale, int precision, npf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Predictive head and losses
 * ------------------------------------------------------------------------------------------------ */
/* NeuralProcessFamily.decode tail (npf/neuralproc/base.py:350-353, scale transform :116):
 *   loc = suff[:, :y];  scale = min_scale + (1 - min_scale) * softplus(suff[:, y:])        suff [M,2y] */
NPF_API int npf_gauss_head_fwd(const float* suff, float* loc, float* scale, long M, int y, float min_scale,
                       npf_stream_t stream);
NPF_API int npf_gauss_head_bwd(const float* suff, const float* dloc, const float* dscale, float* dsuff, long M, int y,
                       float min_scale, npf_stream_t stream);

/* sum_log_prob (npf/losses.py:18-24) of Independent(Normal(loc, scale)):
 *   slp[z,b] = sum_{t,j} -log(scale) - 0.5 log(2 pi) - 0.5 ((Y - loc)/scale)^2
 * loc/scale [Z,B,T*y] (T*y = n), Y [B,T*y] broadcast over z. */
NPF_API int npf_gauss_nll_fwd(const float* loc, const float* scale, const float* Y, float* slp, int Z, int B, long n,
                      npf_stream_t stream);
/* dloc, dscale given g[z,b] = dLoss/dslp[z,b] */
NPF_API int npf_gauss_nll_bwd(const float* loc, const float* scale, const float* Y, const float* g, float* dloc,
                      float* dscale, int Z, int B, long n, npf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Latent path (LatentNeuralProcessFamily.infer_latent_dist / latent_path npf/neuralproc/base.py:495-547,
 * scale transform :432; global latent of ConvLNP.add_global_latent npf/neuralproc/convnp.py:322-335)
 * ------------------------------------------------------------------------------------------------ */
/*   q_loc = suff[:, :zd];  q_scale = 0.1 + 0.9 sigmoid(suff[:, zd:]);  z[s,m,:] = q_loc + q_scale * eps[s,m,:]
 * suff [M,2zd], eps/z [S,M,zd].  q_loc/q_scale [M,zd] are also written (they parameterise q(z|C)). */
NPF_API int npf_latent_sample_fwd(const float* suff, const float* eps, float* q_loc, float* q_scale, float* z, int S,
                          long M, int zd, npf_stream_t stream);
/* dsuff[M,2zd] (overwritten) from dz[S,M,zd] plus optional direct grads dq_loc/dq_scale [M,zd] (e.g. KL) */
NPF_API int npf_latent_sample_bwd(const float* suff, const float* eps, const float* dz, const float* dq_loc,
                          const float* dq_scale, float* dsuff, int S, long M, int zd, npf_stream_t stream);
/* out[n,p,:] = [ zin[n,p,:C/2] ; mean_p zin[n,p,C/2:] ]       zin/out [N,P,C] */
NPF_API int npf_global_latent_fwd(const float* zin, float* out, int N, int P, int C, npf_stream_t stream);
NPF_API int npf_global_latent_bwd(const float* dout, float* dzin, int N, int P, int C, npf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Optimizer step on flat buffers -- the step AFTER the path (torch.optim.Adam as configured by utils/train.py:50, 237-254).
 * One pass over n contiguous fp32 parameters and their gradient / first / second moment buffers; `step` is the 1-based
 * step count (bias corrections computed on the host in double), grad_scale multiplies the gradient first (1 / world
 * size, or a clipping factor), weight_decay is the L2 (non-decoupled) form of torch.optim.Adam.
 * ------------------------------------------------------------------------------------------------ */
NPF_API int npf_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long n, int step, float lr,
                  float beta1, float beta2, float eps, float weight_decay, float grad_scale, npf_stream_t stream);

/* Gradient clipping by global norm without a host sync (skorch GradientNormClipping -> torch.nn.utils.clip_grad_norm_,
 * the callback of the ConvLNP / AttnLNP notebooks):  npf_sqnorm overwrites the device scalar sqnorm[0] with sum_i x_i^2
 * over the flat gradient bucket; npf_adam_step_clipped is npf_adam_step with the gradient additionally multiplied by
 * min(1, max_norm / (grad_scale * sqrt(sqnorm[0]) + 1e-6)), read on the device. */
NPF_API int npf_sqnorm(const float* x, long n, float* sqnorm, npf_stream_t stream);
NPF_API int npf_adam_step_clipped(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long n, int step, float lr,
                  float beta1, float beta2, float eps, float weight_decay, float grad_scale, const float* sqnorm,
                  float max_norm, npf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Context / target split on the device -- the step BEFORE the path (npf/utils/datasplit.py).
 * Index / byte work: results are bit-exact against oracle/datasplit_oracle.py (same Philox-4x32-10 draws).
 *
 * npf_random_subset   GetRandomIndcs.__call__ (datasplit.py:108-145, is_batch_share=False) for a given count n:
 *                     indcs[b, 0..n) (int32) = the first n entries of an independent uniformly random permutation of
 *                     0..N-1 per row b (partial Fisher-Yates; draw i of row b = word i&3 of
 *                     Philox4x32-10(counter {i>>2, b, 0, 0}, key {seed lo, seed hi}), j = i + mulhi(draw, N-i)).
 *                     N <= 12288.  The count itself (random.randint(a, b) upstream) stays a host decision.
 * npf_random_mask     RandomMasker.__call__ (datasplit.py:259-278): mask[b, p] (bytes, 0/1) = 1 at the same n positions
 *                     npf_random_subset would return for (B, P, n, seed); every other byte of the row is cleared.
 * npf_select_points   CntxtTrgtGetter.select (datasplit.py:246-255): Xo[b,i,:] = X[b, indcs[b,i], :] (xd features),
 *                     Yo[b,i,:] = Y[b, indcs[b,i], :] (yd values); indices must lie in [0, N).
 * npf_grid_select     GridCntxtTrgtGetter.select (datasplit.py:423-452): for each row the masked grid points in
 *                     row-major order (the order of mask.nonzero()): Xo[b,k,:] = grid coordinates normalised to
 *                     [-1, 1] * upscale (n_grid_dim = 2: (row, col) of an H x W grid; 1: H == 1), Yo[b,k,:] = img[b,p,:].
 *                     At most n points per row are written; counts[b] (int32, may be NULL) receives the number of
 *                     masked points so the caller can verify "same count in every row" without a sync in the hot loop.
 * ------------------------------------------------------------------------------------------------ */
NPF_API int npf_random_subset(int32_t* indcs, int B, int N, int n, unsigned long long seed, npf_stream_t stream);
NPF_API int npf_random_mask(uint8_t* mask, int B, int P, int n, unsigned long long seed, npf_stream_t stream);
NPF_API int npf_select_points(const float* X, const float* Y, const int32_t* indcs, float* Xo, float* Yo, int B, int N, int n,
                  int xd, int yd, npf_stream_t stream);
NPF_API int npf_grid_select(const uint8_t* mask, const float* img, float* Xo, float* Yo, int32_t* counts, int B, int H, int W,
                  int n_grid_dim, int yd, int n, float upscale, npf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Gaussian-process prior sampler -- the synthetic-task generator in front of the path (GPDataset._sample_targets,
 * utils/data/gaussian_process.py:201-231: sklearn GaussianProcessRegressor.sample_y of the un-fitted regressor, i.e.
 * y ~ N(0, k(X, X)); kernels of utils/ntbks_helpers.py:76-108).
 *   X [B, N] point positions (raw scale, e.g. [-2, 2]), eps [B, S, N] standard-normal draws, Y [B, S, N] samples:
 *   S samples share the positions of task b (upstream's n_same_samples).
 *   kernel: 0 = RBF(length_scale), 1 = Matern(length_scale, nu = 1.5), 2 = ExpSineSquared(length_scale, periodicity);
 *   noise_level > 0 adds WhiteKernel(noise_level) (on the diagonal).
 *   Method: diagonally pivoted Cholesky K = L L^T + E stopped when every residual variance is <= tol (these matrices are
 *   numerically rank-deficient), y_s = L eps_s with eps indexed by factorisation step.  Optional outputs (may be NULL):
 *   L [B, N, N] row-major (columns beyond the rank are zero; rows in the original point order), rank [B] (int32).
 *   N <= 232 (the factor lives in one CTA's shared memory).
 * ------------------------------------------------------------------------------------------------ */
NPF_API int npf_gp_sample(const float* X, const float* eps, float* Y, float* L, int32_t* rank, int B, int N, int S, int kernel,
                  float length_scale, float periodicity, float noise_level, float tol, npf_stream_t stream);
/* Same with PER-TASK hyper-parameters (GPDataset(is_vary_kernel_hyp=True), utils/data/gaussian_process.py:206-207, 233-242: every group
 * of n_same_samples functions is drawn from a kernel whose hyper-parameters were sampled uniformly in their bounds):
 * hyp [B, 3] device = (length_scale, periodicity, noise_level) of task b; the caller draws them. */
NPF_API int npf_gp_sample_hyp(const float* X, const float* eps, float* Y, float* L, int32_t* rank, const float* hyp, int B, int N, int S,
                      int kernel, float tol, npf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Input validation without a host sync  (NeuralProcessFamily._validate_inputs npf/neuralproc/base.py:241-247,
 * isin_range npf/utils/helpers.py:55-57): flag[0] |= 1 if any x outside [lo, hi]  (flag is device int32)
 * ------------------------------------------------------------------------------------------------ */
NPF_API int npf_range_check(const float* X, long n, float lo, float hi, int* flag, npf_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* NPF_B200_H */
