"""On-device rollout collection: the loop of `MAPPOTrainer.collect_rollouts` (MAPPO/trainer.py:154-318) with
every env-side piece on the GPU -- ONE fused kernel launch per rollout step writes the step's features straight
into the `[T, N, ...]` rollout buffers (no per-array `torch.from_numpy(...).to(device)`), the policy's int64
samples are decoded inside the kernel (no `.item()` / `inverse_transform` loop), and GAE is one kernel.

`policy(obs, vec, gstate, gvec) -> (actions int64 [N, R], log_probs f32 [N, R], values f32 [N])` stands for
`_get_actions_and_values` (MAPPO/trainer.py:133-152); `value(gstate, gvec) -> f32 [N]` for the critic call at
:291.  The networks themselves are out of scope (they are the caller of this path).
"""
from __future__ import annotations

import torch

from .batch import compute_gae


class RolloutBuffers:
    """The `mb_*` tensors of MAPPO/trainer.py:156-164 (same shapes and dtypes), allocated once and reused."""

    def __init__(self, batch, rollout_steps: int):
        T, N, R = int(rollout_steps), batch.n_envs, batch.n_robots
        H, W, dev = batch.rows_out, batch.cols_out, batch.device
        f32 = dict(dtype=torch.float32, device=dev)
        # the kernels want 16-byte aligned base pointers: pad each step's slab to a multiple of 4 floats
        def stepwise(per_step, shape):
            slab = torch.zeros((T + 1, (per_step + 3) // 4 * 4), **f32)
            return slab[:, :per_step].view(T + 1, *shape)

        self.obs = stepwise(N * R * 6 * H * W, (N, R, 6, H, W))  # row T holds the state after the last step
        self.vector_obs = stepwise(N * R * batch.vec_dim, (N, R, batch.vec_dim))
        self.global_states = stepwise(N * 4 * H * W, (N, 4, H, W))
        self.global_vector = stepwise(N * batch.gvec_dim, (N, batch.gvec_dim))
        self.actions = torch.zeros((T, N, R), dtype=torch.long, device=dev)
        self.log_probs = torch.zeros((T, N, R), **f32)
        self.rewards = torch.zeros((T, N), **f32)
        self.env_rewards = torch.zeros((T, N), dtype=torch.float64, device=dev)
        self.dones = torch.zeros((T, N), dtype=torch.uint8, device=dev)
        self.values = torch.zeros((T, N), **f32)
        self.T = T

    def feat(self, t):
        return self.obs[t], self.vector_obs[t], self.global_states[t], self.global_vector[t]


def collect_rollouts(vec_env, policy, value, rollout_steps: int, gamma: float = 0.99, gae_lambda: float = 0.95,
                     buffers: RolloutBuffers = None, fresh: bool = True, parts: int = 1):
    """-> (b_obs, b_vector_obs, b_global_states, b_global_vector, b_actions, b_log_probs, b_advantages,
    b_returns_critic, mb_rewards, buffers): the flattened training batch of MAPPO/trainer.py:303-318 plus the
    buffers (whose row T is the state the next rollout starts from).  `fresh=False` continues from row T of
    `buffers` instead of featurising the current state again.

    `parts > 1`: the env batch is split into that many contiguous ranges (VectorizedEnv.split), each with its own CUDA stream on which
    the policy is evaluated on the range's rows and the range's fused env launch is enqueued -- the policy's (compute-bound) kernels of
    one range overlap the (HBM-bound) env kernel of another.  `policy` must be row-wise (each env's outputs depend on that env's
    features only, as in MAPPO/trainer.py:133-152); with a deterministic row-wise policy the results equal parts = 1 (tested).
    Whether it pays depends on the policy: it needs env stepping and inference to take comparable time.  With the small CNN of
    examples/mappo_rollout_demo.py inference is 98 % of a step and halving its batch costs more than the overlap saves (32 768 envs:
    5.5 M env-steps/s with parts = 1, 4.4-4.9 M with parts = 2), so the default stays 1."""
    b = vec_env.batch
    T, N, R = int(rollout_steps), b.n_envs, b.n_robots
    buf = buffers if buffers is not None else RolloutBuffers(b, T)
    if fresh or buffers is None:
        b.featurize(out=buf.feat(0))
    else:
        for dst, src in zip(buf.feat(0), buf.feat(T)):
            dst.copy_(src)
    if parts > 1:
        sp = vec_env.split(parts, out=(buf.env_rewards[0], buf.rewards[0], buf.dones[0]), feat_out=buf.feat(1))
        with torch.no_grad():
            sp.fork()
            for t in range(T):
                step_out, step_feat = (buf.env_rewards[t], buf.rewards[t], buf.dones[t]), buf.feat(t + 1)
                for k, (lo, hi) in enumerate(sp.ranges):
                    with torch.cuda.stream(sp.streams[k]):
                        actions, log_probs, values = policy(*(x[lo:hi] for x in buf.feat(t)))
                        buf.actions[t, lo:hi].copy_(actions)
                        buf.log_probs[t, lo:hi].copy_(log_probs)
                        buf.values[t, lo:hi].copy_(values)
                        sp.step_part(k, buf.actions[t], out=step_out, feat_out=step_feat)
            sp.join()
            last_value = value(buf.global_states[T], buf.global_vector[T]).reshape(N)
        advantages, returns = compute_gae(buf.rewards, buf.values, buf.dones, last_value, gamma, gae_lambda)
        return _flatten(b, buf, T, N, R, advantages, returns)
    with torch.no_grad():
        for t in range(T):
            actions, log_probs, values = policy(*buf.feat(t))
            buf.actions[t].copy_(actions)
            buf.log_probs[t].copy_(log_probs)
            buf.values[t].copy_(values)
            # step + shaped reward + auto-reset + features of the next state, written in place (:208-286)
            b.step_featurize(buf.actions[t], auto_reset=True, out=(buf.env_rewards[t], buf.rewards[t], buf.dones[t]),
                             feat_out=buf.feat(t + 1))
        last_value = value(buf.global_states[T], buf.global_vector[T]).reshape(N)
    advantages, returns = compute_gae(buf.rewards, buf.values, buf.dones, last_value, gamma, gae_lambda)
    return _flatten(b, buf, T, N, R, advantages, returns)


def _flatten(b, buf, T, N, R, advantages, returns):
    """The flattened training batch of MAPPO/trainer.py:303-318."""
    H, W = b.rows_out, b.cols_out
    return (buf.obs[:T].reshape(-1, 6, H, W), buf.vector_obs[:T].reshape(-1, b.vec_dim),
            buf.global_states[:T].reshape(T * N, 4, H, W), buf.global_vector[:T].reshape(T * N, b.gvec_dim),
            buf.actions.reshape(-1), buf.log_probs.reshape(-1),
            advantages.reshape(T * N, 1).repeat(1, R).reshape(-1), returns.reshape(-1), buf.rewards, buf)
