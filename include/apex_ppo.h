/* C-ABI of libapex_b200.so — learner-side kernels of the PPO path (float32, device pointers, cudaStream_t as void*).
 *
 * Each entry point replaces a piece of rl/algos/ppo.py that the reference runs as PyTorch-on-CPU code:
 *   apex_gae_scan        PPOBuffer.finish_path (ppo.py:73-89) for every env at once; lam = 1 reproduces it exactly
 *   apex_moments / apex_normalize   advantage normalisation (ppo.py:395-396)
 *   apex_mlp_forward / apex_mlp_backward   Gaussian_FF_Actor / FF_V trunks (rl/policies/actor.py:183-197, critic.py:65-74)
 *   apex_gaussian_sample torch.distributions.Normal(mu, sd).sample() + log_prob (actor.py:199-215)
 *   apex_prepare_obs     minibatch gather, observation normalisation, SymmetricEnv.mirror_clock_observation
 *                        (rl/envs/wrappers.py:59-67)
 *   apex_ppo_loss        the loss block of PPO.update_policy (ppo.py:276-317) forward + backward
 *   apex_grad_sumsq / apex_adam_step   clip_grad_norm_ + Adam.step (ppo.py:319-330)
 * Return value: 0 on success, negative cudaError_t otherwise (-1000 = bad argument).
 */
#ifndef APEX_PPO_H
#define APEX_PPO_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* y = W3 relu(W2 relu(W1 x + b1) + b2) + b3; W are torch Linear weights [out, in]; h1, h2 [rows, hid] are kept for backward */
int apex_mlp_forward(const float *x, int rows, int in_dim, int hid, int out_dim, const float *w1, const float *b1, const float *w2,
                     const float *b2, const float *w3, const float *b3, float *h1, float *h2, float *y, void *stream);
/* accumulates (+=) into gw*, gb*; dh2, dh1 [rows, hid] are scratch */
int apex_mlp_backward(const float *x, int rows, int in_dim, int hid, int out_dim, const float *w2, const float *w3, const float *h1,
                      const float *h2, const float *dy, float *dh2, float *dh1, float *gw1, float *gb1, float *gw2, float *gb2,
                      float *gw3, float *gb3, void *stream);
/* rows of obs (optionally gathered through idx) -> raw copy, normalised copy, mirrored + normalised copy (any may be NULL) */
int apex_prepare_obs(const float *obs, const int64_t *idx, int rows, int dim, const float *mean, const float *stdv,
                     const int *mir_src, const float *mir_sign, const int *clock_mask, float *raw, float *xn, float *xn_mir,
                     void *stream);
/* act = mu + sigma * anneal * N(0,1) (Philox keyed by seed, row0 + row, step), logp [rows] */
int apex_gaussian_sample(const float *mu, const float *sigma, float anneal, int rows, int adim, unsigned seed, unsigned step,
                         unsigned row0, float *act, float *logp, void *stream);
/* stats[6] (double, += ): sum surrogate, sum 0.5 (R-V)^2, sum ratio, sum KL (summed over action dims), sum mirror sq. err, count */
int apex_ppo_loss(int rows, int adim, const float *mu, const float *mu_mir, const int64_t *idx, const float *act_all,
                  const float *oldlogp_all, const float *adv_all, const float *ret_all, const float *oldmu_all, const float *value,
                  const float *sigma, float clip, float mirror_coeff, const int *amir_src, const float *amir_sign, float *dmu,
                  float *dmu_mir, float *dvalue, double *stats, void *stream);
/* apex_gaussian_sample with (seed, anneal as float bits) read from dyn[0..1] in device memory: PPO.sample_parallel captures a whole
 * rollout (T steps of observation copy, actor / critic inference, sampling, env step; then the return scan) as one CUDA graph and
 * replays it every iteration — only these two words change between replays */
int apex_gaussian_sample_dev(const float *mu, const float *sigma, const unsigned *dyn, int rows, int adim, unsigned step, unsigned row0,
                             float *act, float *logp, void *stream);
int apex_grad_sumsq(const float *g, int n, double *out, void *stream); /* out += sum g^2 */
int apex_adam_step(float *p, const float *g, float *m, float *v, int n, const double *sumsq, float gscale, float max_norm, float lr,
                   float beta1, float beta2, float eps, int step, void *stream);
/* rew, val, term_val, ret, adv [T, N]; done [T, N] (bit0 terminal, bit1 time-out); last_val [N]; T <= 512 */
int apex_gae_scan(int T, int N, const float *rew, const float *val, const int *done, const float *term_val, const float *last_val,
                  float gamma, float lam, float *ret, float *adv, void *stream);
int apex_moments(const float *x, long n, double *out3, void *stream);            /* out3 += (sum, sum sq, count) */
int apex_normalize(float *x, long n, const double *mom3, float eps, void *stream); /* (x - mean) / (std_unbiased + eps) */

/* out[0..dim) += column sums, out[dim..2dim) += column sums of squares of x [rows, dim]  (rl/envs/normalize.py:48) */
int apex_col_moments(const float *x, int rows, int dim, double *out, void *stream);

/* ---- TD3 (rl/algos/sync_td3.py:133-209, rl/utils/remote_replay.py:65-90) ---- */
/* apex_mlp_backward + dx [rows, in_dim] (may be NULL) and a switch for the weight gradients */
int apex_mlp_backward_dx(const float *x, int rows, int in_dim, int hid, int out_dim, const float *w1, const float *w2, const float *w3,
                         const float *h1, const float *h2, const float *dy, float *dh2, float *dh1, float *dx, int want_wgrads,
                         float *gw1, float *gb1, float *gw2, float *gb2, float *gw3, float *gb3, void *stream);
/* replay rows [state | next_state | action | reward | done] gathered by idx -> state, next_state, [state|action], reward, 1-done */
int apex_replay_gather(const float *storage, const int64_t *idx, int rows, int S, int A, float *state, float *next_state, float *sa,
                       float *reward, float *notdone, void *stream);
/* sa = [state | clamp(max_a tanh(pre) + clamp(noise, +-noise_clip), +-max_a)]; noise: explicit [rows, A] or Philox N(0, policy_noise) */
int apex_td3_action(const float *pre, const float *state, const float *noise, int rows, int S, int A, float max_a, float policy_noise,
                    float noise_clip, unsigned seed, unsigned step, float *sa, float *tanh_out, void *stream);
/* CUDA-graph friendly forms (TD3 at the reference's batch of 256 is launch-bound: TD3.train(use_graph=True) captures policy_freq
 * iterations once and replays them): counters that change between replays are read from device memory.
 * apex_adam_step_dev: Adam with the step count at *step_dev; apex_td3_action_dev: smoothing noise from Philox(seed, row, *step_dev);
 * apex_replay_sample: idx[i] uniform in [0, *size_dev) from Philox(seed, i, *ctr_dev) (remote_replay.py:78-79: sampling with
 * replacement); apex_counter_add: *counter_dev += inc. */
int apex_adam_step_dev(float *p, const float *g, float *m, float *v, int n, const double *sumsq, float gscale, float max_norm, float lr,
                       float beta1, float beta2, float eps, const int *step_dev, void *stream);
int apex_td3_action_dev(const float *pre, const float *state, int rows, int S, int A, float max_a, float policy_noise, float noise_clip,
                        unsigned seed, const int *step_dev, float *sa, float *tanh_out, void *stream);
int apex_replay_sample(int64_t *idx, int rows, const int *size_dev, unsigned seed, const int *ctr_dev, void *stream);
int apex_counter_add(int *counter_dev, int inc, void *stream);
/* twin-Q target and MSE gradients; stats[3] (double, +=): loss, sum Q1, sum Q2 */
int apex_td3_critic_loss(int rows, const float *q1, const float *q2, const float *q1t, const float *q2t, const float *reward,
                         const float *notdone, float discount, float *dq1, float *dq2, double *stats, void *stream);
int apex_td3_actor_grad(int rows, int S, int A, const float *dsa, const float *tanh_v, float max_a, float *dpre, void *stream);
int apex_polyak(float *target, const float *src, int n, float tau, void *stream); /* target = tau src + (1 - tau) target */

/* ARS (rl/algos/ars.py): act[e] = Linear_Actor(theta + sign[e] * noise[idx[dir[e]] : +P])(obs[e]); S <= 64, H, A <= 32 */
int apex_ars_policy(const float *obs, int n, int S, int H, int A, const float *theta, const float *noise, const int64_t *idx,
                    const int *dir, const float *sign, const float *obs_mean, const float *obs_std, float *act, void *stream);
/* theta += coef * sum_d weight[d] * noise[idx[d] : +P]   (ars.py:152-156) */
int apex_ars_update(float *theta, int P, const float *noise, const int64_t *idx, const float *weight, int ndir, float coef,
                    void *stream);

/* ---- tensor-core path (sm_100a tcgen05; opt-in reduced precision for BASELINE config "PPO CassieTraj-v0 ... bf16") ----
 * y [M, N] = act(x [M, K] W^T + b): bf16 operands converted while staging, float32 accumulation in tensor memory, float32
 * in / out.  N in {64, 128, 256}, K a multiple of 64, y 16-byte aligned; -1000 otherwise. */
int apex_tc_linear_forward(const float *x, int M, int K, const float *w, const float *bias, int N, int relu, float *y, void *stream);
/* test hook: 0 forces the single-stage tcgen05 kernel, 1 (default) the persistent warp-specialised one when W fits in shared memory */
void apex_set_tc_persistent(int on);
/* apex_mlp_forward with the hidden hid x hid layer on the tensor cores.  With `scratch` (caller-owned, zero-initialised once,
 * >= apex_mlp_bf16_scratch_bytes(rows, hid) bytes, 16-byte aligned; hid 128 or 256) layer 1 also writes h1 as bf16 in the tiled
 * shared-memory image and the hidden layer fetches its operands with cp.async.bulk (TMA); without it (NULL) the hidden layer
 * converts float32 h1 on the fly (apex_tc_linear_forward). */
long apex_mlp_bf16_scratch_bytes(int rows, int hid);
int apex_mlp_forward_bf16(const float *x, int rows, int in_dim, int hid, int out_dim, const float *w1, const float *b1,
                          const float *w2, const float *b2, const float *w3, const float *b3, float *h1, float *h2, float *y,
                          void *scratch, long scratch_bytes, void *stream);
/* the TMA hidden layer on its own: xt = tiled bf16 image of x [M, K] (rows padded to 128), wt_scratch = N * K * 2 bytes */
int apex_tc_linear_tiled(const void *xt, int M, int K, const float *w, void *wt_scratch, const float *bias, int N, int relu, float *y,
                         void *stream);
/* ---- float32-accurate tensor-core path (csrc/tc_gemm3.cu): tcgen05 kind::tf32 with every operand split as hi + lo (two tf32
 * terms) and three products per k step ("3xTF32"), used by apex_mlp_forward / apex_mlp_backward for the 256-wide layers
 * (rl/policies/actor.py:142-215: the reference's default 256 x 256 hidden layers) when rows >= apex_set_tc_min_rows (1024).
 * mode 3 (default) = split, 1 = plain TF32 (10-bit mantissa), 0 = SIMT float32 kernels everywhere. */
void apex_set_tc_mode(int mode);
int apex_get_tc_mode(void);
void apex_set_tc_min_rows(int rows);
/* C [M, 256] = epi(A [M, K] W^T), W(n, k) = w[n * swn + k * swk]; epi = + bias[n], ReLU, zero where mask[m, n] <= 0 (each
 * optional).  Any K <= 1024 (zero-padded to a multiple of 64 on chip); C and mask 16-byte aligned, ldc / ldmask multiples of 4;
 * passes 1 or 3. */
int apex_tc3_linear(const float *A, long lda, int M, int K, const float *w, long swn, long swk, const float *bias, int relu,
                    const float *mask, long ldmask, float *C, long ldc, int passes, void *stream);
/* C [256, nb] (+)= A [R, 256]^T B [R, nb], nb = 256 or <= 64 (weight gradient: reduction over the R rows, one CTA per SM, float
 * atomics into C) */
int apex_tc3_outer(const float *A, long lda, const float *B, long ldb, int nb, long R, float *C, long ldc, int accumulate, int passes,
                   void *stream);
/* test hook: 0 sends the narrow output layer (256 -> 10 / 1) through the GEMM kernels instead of the streaming kernels of
 * csrc/mlp_head.cu (forward; backward = dh2, gW3 and gb3 in one pass over h2) */
void apex_set_head_kernels(int on);
/* test hook: 0 routes every GEMM through the 64 x 64 tile kernel, 1 (default) uses the 128 x 128 one when M, N >= 128 */
void apex_set_gemm_large_tiles(int on);
/* tuning: minimum number of 128 x 128 tiles (x split-k) for the large-tile kernel to be chosen (default 148 = one per SM) */
void apex_set_gemm_min_ctas(int n);

#ifdef __cplusplus
}
#endif
#endif
