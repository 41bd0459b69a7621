/*
 * gail_carla_b200 - C ABI of the B200-native learning hot path of gustavokcouto/gail-carla.
 *
 * The reference has no FFI: its boundary is the Python class API in tools/storage.py, tools/model.py, algo/ppo.py,
 * algo/wdgail.py and common/running_mean_std.py.  This library is what the drop-in Python classes in gail_carla_b200/
 * bind (ctypes): one entry point per fused stage, plain device pointers and sizes, no torch types.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error (cudaError_t / CUresult values or negative argument
 *     errors); gc_last_error_string() describes the last failure on the calling thread; nothing throws.
 *   - all pointers are DEVICE pointers borrowed for the duration of the enqueue; fp32 unless stated; workspace is
 *     passed in by the caller (no hidden allocation); `stream` is a cudaStream_t; nothing synchronises.
 *   - image activations are NHWC with explicit pitches; dense contractions run as tcgen05.mma.kind::tf32 with fp32
 *     accumulation in TMEM (the reference's own cuDNN/cuBLAS path on Ampere+ is TF32 as well under torch 1.8).
 */
#ifndef GAIL_CARLA_B200_H_
#define GAIL_CARLA_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define GC_ABI_VERSION 2

const char* gc_last_error_string(void);
int gc_abi_version(void);
/* sha256 of the sources + flags this binary was built from (gail_carla_b200/build.py); the loader refuses a binary whose
 * digest differs from the source tree next to it. */
const char* gc_build_digest(void);

/* ------------------------------------------------------------------------------------------------------------
 * Rollout maths (HBM-bound)
 * ---------------------------------------------------------------------------------------------------------- */

/* RolloutStorage.compute_returns - tools/storage.py:37-50 (gail_coef=1, env_coef=0).
 * gail_rewards [T,N]; value_preds, masks, returns [T+1,N] time-major.  Writes returns[0:T] (row T untouched).
 * adv_raw (nullable) [T,N] receives returns-value_preds; stats (nullable) double[3] receives
 * {sum adv, sum adv^2, T*N} for algo/ppo.py:47-49. */
int gc_gae_returns(const float* gail_rewards, const float* value_preds, const float* masks, float* returns, float* adv_raw,
                   double* stats, int T, int N, float gamma, float gae_lambda, void* stream);

/* {sum, sumsq, n} of returns-value_preds over n elements - algo/ppo.py:47. */
int gc_adv_stats(const float* returns, const float* value_preds, double* stats, long n, void* stream);

/* adv_out = ((returns-value_preds) - mean) / (unbiased_std + 1e-5) - algo/ppo.py:48-49. */
int gc_adv_normalize(const float* returns, const float* value_preds, const double* stats, float* adv_out, long n, void* stream);

/* Policy head tail + PPO loss forward/backward - tools/model.py:45-53,80-85; algo/ppo.py:80-85,88-99,104-113.
 * head_out [B,4] = {value, mu0_raw, mu1_raw, pad}; actions [B,2]; the scalar columns are [B].
 * mode 0 PPO: d_head_out = d(value_coef*value_loss + action_weight*action_loss)/d head_out.  Advantages come from `adv`
 *             when non-null, else are formed in-kernel as ((returns-value_old)-mean)/(std+1e-5) from adv_stats.
 *             clipped_value != 0: 0.5*max((v-R)^2, (v_clipped-R)^2) (algo/ppo.py:104-111); 0: 0.5*(R-v)^2 (:112-113).
 * mode 1 BC : d_head_out = d(action_weight * -mean(logp))/d head_out.
 * mode 2    : forward only.
 * inv_norm > 0 replaces 1/B as the weight of one sample in the batch means (a rank that holds B rows of a global
 * minibatch of B_total rows passes 1/B_total and the per-rank gradients are summed); <= 0 selects 1/B.
 * out_value / out_logp (nullable) [B].  loss_acc (nullable) double[3] += {sum 0.5*max(.)^2 terms, sum -min(surr), sum -logp}. */
int gc_ppo_loss_fwd_bwd(const float* head_out, const float* actions, const float* old_logp, const float* value_old,
                        const float* returns, const float* adv, const double* adv_stats, float* d_head_out,
                        float* out_value, float* out_logp, double* loss_acc, int B, float logstd0, float logstd1,
                        int activation, float clip, float value_coef, float action_weight, int mode, int clipped_value,
                        float inv_norm, void* stream);

/* Policy.act - tools/model.py:25-36.  noise (nullable) [B,2] standard normal draws; null => deterministic. */
int gc_policy_act(const float* head_out, const float* noise, float* value, float* action, float* logp, int B, float logstd0,
                  float logstd1, int activation, void* stream);

/* RunningMeanStd.update for shape () - common/running_mean_std.py:10-31.  state double[3] = {mean,var,count};
 * scratch2 double[2]. */
int gc_welford_merge(double* state, const float* x, long n, double* scratch2, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Data movement / small elementwise stages
 * ---------------------------------------------------------------------------------------------------------- */

/* Minibatch image gather - tools/storage.py:65-66 + tools/model.py:157-161 (normalise).
 * src [rows,3,192,192] fp32 (CHW); idx (nullable) int64[B] row indices (null => rows 0..B-1).
 * out [B,96,96,16] space-to-depth NHWC: out[b,Y,X,dy*8+dx*4+c] = (src[idx[b],c,2Y+dy,2X+dx]-mean[c])/std[c], channel 3 = 1
 * (the ones column that turns conv1's bias gradient into a wgrad column; its conv1 weight is 0). */
int gc_gather_obs_s2d(const float* src, const long long* idx, float* out, int B, void* stream);
/* Same from a uint8 table [rows,3,192,192] (the expert data set as its PNGs store it - algo/wdgail.py:222-241 decodes each
 * image to uint8 and ToTensor() divides by 255): out = ((float)src/255 - mean)/std, bit-identical to decoding first. */
int gc_gather_obs_u8_s2d(const unsigned char* src, const long long* idx, float* out, int B, void* stream);

/* One critic minibatch in one pass (algo/wdgail.py:66-80,116,121), both sources uint8 tables: out rows [0,B) = expert images
 * src_e[idx_e[b]], rows [B,2B) = policy images src_p[idx_p[b]], rows [2B,3B) = alpha[b]*expert + (1-alpha[b])*policy, all as
 * gc_gather_obs_u8_s2d writes them (the mix equals gc_mixup applied to the first two row blocks, bit for bit). */
int gc_gather_pair_mix_u8_s2d(const unsigned char* src_e, const long long* idx_e, const unsigned char* src_p, const long long* idx_p,
                              const float* alpha, float* out, int B, void* stream);

/* out[b, 0:width] = src[idx[b], 0:width] (idx nullable), out row pitch ldo - tools/storage.py:66-76 scalar columns. */
int gc_gather_rows(const float* src, const long long* idx, float* out, int B, int width, long ldo, void* stream);

/* out = alpha[b]*xe + (1-alpha[b])*xp over per_sample floats per sample - algo/wdgail.py:66-80. */
int gc_mixup(const float* xe, const float* xp, const float* alpha, float* out, int B, long per_sample, void* stream);

/* ProcessMetrics (+ action passthrough) - tools/model.py:179-213, algo/wdgail.py:45-50.
 * metrics [B,4]; emb [10,8]; action (nullable) [B,2]; writes out[b,0:13] (+[13:15] action) and zero-fills up to `pad`
 * columns; out row pitch ldo.  When metrics2 / alpha are non-null the inputs are first mixed
 * alpha*metrics+(1-alpha)*metrics2 (and action likewise with action2) as algo/wdgail.py:72-80 does. */
int gc_metrics_features(const float* metrics, const float* metrics2, const float* action, const float* action2,
                        const float* alpha, const float* emb, float* out, long ldo, int pad, int B, void* stream);

/* d_emb[int(metrics[b,3]), j] += d_feat[b, 5+j] - backward of the embedding lookup (metrics mixed as above). */
int gc_metrics_features_bwd(const float* metrics, const float* metrics2, const float* alpha, const float* d_feat, long ldf,
                            float* d_emb, int B, void* stream);

/* y[b,j] = sum_k x[b,k]*w[j,k] + bias[j] for tiny N (<=4): head.2 / trunk.2 - tools/model.py:117-126, algo/wdgail.py:31. */
int gc_small_linear_fwd(const float* x, long ldx, const float* w, const float* bias, float* y, long ldy, int B, int N, int K,
                        void* stream);
/* dx[b,k] = lrelu'(x[b,k]) * sum_j dy[b,j]*w[j,k]  (x is the post-LeakyReLU input; slope<0 disables the mask);
 * dw[j,k] += sum_b dy[b,j]*x[b,k]; db[j] += sum_b dy[b,j]  over the first B_params rows (dw/db nullable). */
int gc_small_linear_bwd(const float* x, long ldx, const float* w, const float* dy, long lddy, float* dx, long lddx, float* dw,
                        float* db, int B, int B_params, int N, int K, float slope, void* stream);

/* Discriminator loss seeds - algo/wdgail.py:116-131.  d [3B] = {expert, policy, mixup} critic outputs.
 * dd[0:B] = -(1-tanh^2)/B, dd[B:2B] = +(1-tanh^2)/B, dd[2B:3B] = 1 (seed of dD/dx for the penalty).
 * acc double[4] += {sum d_e, sum d_p, sum tanh d_e, sum tanh d_p}.  inv_norm as for gc_ppo_loss_fwd_bwd (replaces 1/B). */
int gc_disc_loss_seed(const float* d, float* dd, double* acc, int B, float inv_norm, void* stream);

/* Gradient penalty - algo/wdgail.py:93-97.  g [B,per_sample] = dD/dx_n w.r.t. the *normalised* s2d input; the reference
 * differentiates w.r.t. the raw input, g_raw = g * s_c with s_c = 1/std_c and c = element index & 3 (s3 = 0 for the pad
 * channel).  acc[0] += sum_b (||g_raw,b||-1)^2;  u = d(lambda*mean_b(||g_raw,b||-1)^2)/dg, same layout as g; the mean is
 * over 1/inv_norm samples when inv_norm > 0, else over B. */
int gc_grad_penalty(const float* g, float* u, double* acc, int B, long per_sample, float lambda_, float s0, float s1, float s2,
                    float inv_norm, void* stream);

/* reward = -log(1 - sigmoid(d)) - algo/wdgail.py:185-186. */
int gc_reward_epilogue(const float* d, float* reward, long n, void* stream);

/* out[c] += sum_{r<rows} x[r*ld + c] for c < C  (bias gradients). */
int gc_colsum(const float* x, long ld, long rows, int C, float* out, void* stream);

/* out[m,n] = epi(sum_z part[z,m,n]); epilogue as gc_linear_fwd (0 store, 1 bias+LeakyReLU, 2 bias, 3 mask by sign of mask_src). */
int gc_splitk_reduce(const float* part, int splits, long M, int N, long ldp, const float* bias, const float* mask_src, long ldm,
                     float* out, long ldo, int epilogue, float slope, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Parameter layout preparation (reference layout <-> GEMM operand layout)
 * ---------------------------------------------------------------------------------------------------------- */

/* Conv2d weight [Cout,Cin,4,4] (tools/model.py:137-143) -> fprop operand [Cout][ky][kx][c] and dgrad operand
 * [cls=py*2+px][c][a][b'][n] (ky=py+2a, kx=px+2b').  layer1=1 selects the space-to-depth form of conv1:
 * fprop [32][ky2][px][dy][dx][c4] (c=3 zero) and dgrad [16 (dy,dx,c4)][a][b'][n]. */
int gc_prep_conv_weight(const float* w, float* w_fprop, float* w_dgrad, int Cout, int Cin, int layer1, void* stream);
/* inverse for gradients: dw[Cout,Cin,4,4] = sum_z part[z][n][ky][kx][c] (same fprop operand layout).
 * layer1 with dbias != NULL: the pad channel of the space-to-depth image holds 1.0 (gc_gather_obs_s2d), so the wgrad
 * column of that channel is sum_pixels dy[n] - the bias gradient - and is written to dbias[Cout] (no column-sum pass). */
int gc_unprep_conv_wgrad(const float* part, int splits, float* dw, float* dbias, int Cout, int Cin, int layer1, void* stream);
/* FC1 weight [out, 25600+tail] (NCHW-flatten columns c*100+p) -> [out, ld] with NHWC columns p*256+c, tail copied, pad 0. */
int gc_prep_fc1_weight(const float* w, float* w_gemm, int out, int tail, long ld, void* stream);
/* dw[out, 25600+tail] = sum_z part[z][out][ld] with the inverse column permutation. */
int gc_unprep_fc1_wgrad(const float* part, int splits, float* dw, int out, int tail, long ld, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Optimiser: clip_grad_norm_ + Adam - algo/ppo.py:115-119, algo/wdgail.py:140-145
 * ---------------------------------------------------------------------------------------------------------- */
/* Zero `rows` runs of `width_bytes` bytes that start `pitch_bytes` apart (a memset node on the stream, no kernel): loss
 * accumulators, the metric / action columns of the gradient-penalty rows of the feature matrix (algo/wdgail.py:85-91 takes
 * the gradient w.r.t. the image only). */
int gc_zero_block(void* ptr, long pitch_bytes, long rows, long width_bytes, void* stream);

/* sumsq double[1] = sum (grad_scale*g)^2 (the accumulator is cleared on the stream first).  grad_scale = 1/world_size after a NCCL SUM all-reduce of per-rank mean
 * gradients (the global-batch gradient of algo/ppo.py:115), 1 otherwise. */
int gc_grad_sumsq(const float* grad, long n, float grad_scale, double* sumsq, void* stream);
/* g' = grad_scale*g * min(1, max_norm/(sqrt(sumsq)+1e-6)); Adam step with g' (torch.optim.Adam, no amsgrad/weight decay).
 * bias_corr1 = 1-beta1^t, bias_corr2 = 1-beta2^t.  max_norm < 0 disables clipping.  zero_grad != 0 writes 0 over each
 * gradient element after reading it (optimizer.zero_grad() of the next minibatch, algo/ppo.py:115, without another pass).
 * dev_hyper (nullable) device float[3] = {lr, bias_corr1, bias_corr2} overrides the by-value arguments, so a captured CUDA
 * graph of the minibatch step can be replayed while the step count and the learning-rate schedule advance. */
int gc_clip_adam(float* param, float* grad, float* exp_avg, float* exp_avg_sq, long n, const double* sumsq, float max_norm,
                 float lr, float beta1, float beta2, float eps, float bias_corr1, float bias_corr2, float grad_scale,
                 int zero_grad, const float* dev_hyper, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Dense contractions on tcgen05 tensor cores (TF32 in, fp32 accumulate in TMEM, TMA-fed implicit GEMM)
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct gc_conv_geom {
  int B;                 /* samples */
  int H, W, Hp, Wp, Cin; /* input extent, row/col pitches (pixels), channels (NHWC) */
  int KH, KW, S;         /* taps and stride */
  int OH, OW, OHp, OWp, Cout;
  long in_batch_stride;  /* floats between consecutive samples of the input / output tensors */
  long out_batch_stride;
} gc_conv_geom;

/* epilogue codes */
#define GC_EPI_STORE 0
#define GC_EPI_BIAS_LRELU 1
#define GC_EPI_BIAS 2
#define GC_EPI_MASK 3

/* LeakyReLU' bit masks (optional accelerator, all `mask_bits` arguments nullable): one bit per fp32 element of an
 * activation tensor in the same linear order (word = element offset / 32; channel counts and batch strides must be
 * multiples of 32).  gc_conv_fprop writes them from its bias+LeakyReLU epilogue (bit = output > 0); the mask epilogues of
 * gc_conv_fprop / gc_conv_dgrad / gc_linear_dgrad read them instead of TMA-loading the fp32 `mask_src` tile
 * (semantically identical: mask = mask_src > 0 ? 1 : slope). */

/* nn.Conv2d forward (tools/model.py:137-143) on NHWC: y = epi(conv(x, w)); w in the fprop operand layout
 * [Cout][KH][KW][Cin].  GC_EPI_MASK multiplies by LeakyReLU'(sign of mask_src) instead of adding a bias (used for the
 * second-order chain of the gradient penalty, algo/wdgail.py:85-97). */
int gc_conv_fprop(const gc_conv_geom* g, const float* x, const float* w, const float* bias, const float* mask_src,
                  unsigned* mask_bits, float* y, int epilogue, float slope, void* stream);
/* data gradient: dx = LeakyReLU'(mask_src) * conv_transpose(dy, w); wd in the dgrad operand layout.
 * dbias_in (nullable, needs mask_bits): dbias_in[c] += sum over the pixels of the first `dbias_samples` samples (<= 0: all) of
 * the masked dx[.., c] - dx is the pre-activation gradient of the layer below, so this is that layer's bias gradient
 * (nn.Conv2d bias, tools/model.py:137-143), summed from the output tiles while they are staged for the store. */
int gc_conv_dgrad(const gc_conv_geom* g, const float* dy, const float* wd, const float* mask_src, const unsigned* mask_bits,
                  float* dx, float slope, float* dbias_in, int dbias_samples, void* stream);
/* weight gradient partials [splits][Cout][KH][KW*Cin]; gc_conv_wgrad_splits suggests `splits` for a geometry. */
int gc_conv_wgrad_splits(const gc_conv_geom* g);
int gc_conv_wgrad(const gc_conv_geom* g, const float* dy, const float* x, float* dw_partial, int splits, void* stream);

/* nn.Linear forward (tools/model.py:93-99,110-113; algo/wdgail.py:27-31): y[z] = epi(x[:, Kz] w[:, Kz]^T). */
int gc_linear_fwd(const float* x, long ldx, const float* w, long ldw, const float* bias, float* y, long ldy, int M, int N, int K,
                  int epilogue, float slope, int splits, void* stream);
/* dx[M,N] = LeakyReLU'(mask_src) * dy[M,K] w[K,N]  (w = forward weight [out=K, in=N]).
 * colsum (nullable, needs mask_bits): colsum[n % colsum_mod] += sum over the first `colsum_rows` rows (<= 0: all) of the
 * masked dx[.., n] - with dx the NHWC conv4 feature gradient [M, 100 pixels x 256 channels] this is conv4's bias gradient. */
int gc_linear_dgrad(const float* dy, long lddy, const float* w, long ldw, const float* mask_src, const unsigned* mask_bits,
                    long ldm, float* dx, long lddx, int M, int N, int K, float slope, float* colsum, int colsum_mod,
                    int colsum_rows, void* stream);
/* dw[z][M,N] = sum_{rows in split z} dy[row,M]^T x[row,N]. */
int gc_linear_wgrad(const float* dy, long lddy, const float* x, long ldx, float* dw, long lddw, int M, int N, int K, int splits,
                    void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GAIL_CARLA_B200_H_ */
