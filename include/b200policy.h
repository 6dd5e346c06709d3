/* b200policy.h -- C ABI of the shared per-agent policy of libb200env.so.
 *
 * The reference's callers evaluate ONE policy network on every agent row of the VecEnv: the
 * stable-baselines runner behind `PPO2/A2C(MlpPolicy, vec_env).learn()` and `.predict(obs)`
 * (run_multiagent_exp_single.py:37-49, play_optimize.py:79-98) calls `model.step(obs)` with the
 * [sum(P), 3H] observation matrix that OptVecEnv.step_wait returned (vectorize/optvecenv.py:78-88).
 * MlpPolicy is two tanh layers of 64 units and a linear head.  At BASELINE config 4 that matrix has
 * 2.08e8 rows; pulling it to the host costs 12.5 GB per step.  These entry points evaluate the
 * action head on the device instead:
 *
 *   action[row] = clip(mean(obs[row]) + noise_std * N(0, 1), low, high)
 *   mean(x)     = w3 . tanh(W2 tanh(W1 x + b1) + b2) + b3
 *
 * with bf16 operands and fp32 accumulation on the tcgen05 tensor cores (a policy is the caller's
 * model, not part of the reference's arithmetic: the parity bar of the env step does not apply;
 * tests/test_gpu_policy.py holds it to a bf16-rounded torch evaluation of the same network).
 *
 * b2p_act reads a dense observation matrix (what b2e_step wrote).  b2p_act_env reads the env's
 * adjusted-history rings directly -- the observation row of agent p of env e is
 * clip(nan_to_num([adj_w[e][newest..oldest][p] | adj_L[e][..] | adj_g[e][..][p]]), +-100) - 1
 * (envs/multioptlrs.py:96-101) -- so that b2e_step may be called with obs_out = NULL and the
 * 3H observation words per agent are never written to or read from HBM.
 *
 * Pointers are DEVICE pointers; `stream` is a cudaStream_t passed as void*; calls are asynchronous
 * on it.  Return 0 on success, non-zero with b2p_last_error() otherwise.
 */
#ifndef B200POLICY_H
#define B200POLICY_H

#include <stddef.h>
#include <stdint.h>

#include "b200env.h"

#ifdef __cplusplus
extern "C" {
#endif

#define B2P_HIDDEN 64            /* units of both hidden layers (stable-baselines MlpPolicy) */
#define B2P_MAX_OBS_DIM 15       /* 3H with H <= 5 (envs/multioptlrs.py:39, utils/utils_env.py:32-36) */

/* tanh_mode */
#define B2P_TANH_F32    0        /* tanh.approx.f32 on the fp32 accumulators (default)                      */
#define B2P_TANH_BF16X2 1        /* first layer: tanh.approx.bf16x2 on the packed operand of the second MMA */

typedef struct b2p_policy *b2p_handle;

int b2p_create(int device, int obs_dim, int tanh_mode, b2p_handle *out);
void b2p_destroy(b2p_handle h);
const char *b2p_last_error(b2p_handle h);

/* torch.nn.Linear layouts: w1 [64, obs_dim], b1 [64], w2 [64, 64], b2 [64], w3 [64], b3 [1]; float32.
 * The handle keeps its own copy. */
int b2p_set_weights(b2p_handle h, const float *w1, const float *b1, const float *w2,
                    const float *b2, const float *w3, const float *b3, void *stream);

/* Optional: a device counter added to `seed` by every later b2p_act / b2p_act_env launch (NULL = none).  The caller
 * advances it on the stream between launches; a CUDA graph that captured the launches then draws fresh noise on
 * every replay although the seed argument is baked into the graph. */
int b2p_set_seed_counter(b2p_handle h, const uint64_t *counter);

/* obs float32 [rows, obs_dim] dense -> actions_out float32 [rows].  noise_std = exp(log_std) of the
 * diagonal Gaussian (0 = deterministic, `model.predict(deterministic=True)`); the noise of a row is a
 * counter-based hash of (seed, row). */
int b2p_act(b2p_handle h, const float *obs, int64_t rows, float *actions_out, float noise_std,
            uint64_t seed, float low, float high, void *stream);

/* The same for every agent row of `env` (MultiOptLRs layout, large-problem pipeline), read from the env's
 * rings; actions_out float32 [E*P] in VecEnv row order, ready for the next b2e_step. */
int b2p_act_env(b2p_handle h, b2e_handle env, float *actions_out, float noise_std, uint64_t seed,
                float low, float high, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* B200POLICY_H */
