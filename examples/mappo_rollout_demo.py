#!/usr/bin/env python
"""What `MAPPOTrainer.collect_rollouts` (MAPPO/trainer.py:154-318) looks like on top of the B200 path.

    python examples/mappo_rollout_demo.py [--envs 4096] [--steps 64] [--iters 5]

A small random-weight actor / critic (same input signature as MAPPO/networks.py: 6-channel local observation + vector
features per agent, 4-channel global state + global vector for the critic) drives the batched environment; every env-side
piece of the trainer's loop -- step, shaped reward, auto-reset, the four feature tensors written straight into the rollout
buffers, int64 action decoding, GAE -- runs on the GPU through libmdx.  The returned batch has the trainer's own flattened
layout (trainer.py:303-318), so the PPO update of the reference can consume it unchanged.
"""
import argparse
import os
import sys
import time

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import marl_delivery_b200 as mdb  # noqa: E402
from marl_delivery_b200.rollout import collect_rollouts  # noqa: E402


class Actor(nn.Module):
    def __init__(self, vec_dim, n_actions=15):
        super().__init__()
        self.cnn = nn.Sequential(nn.Conv2d(6, 16, 3, padding=1), nn.ReLU(), nn.AdaptiveAvgPool2d(1), nn.Flatten())
        self.head = nn.Sequential(nn.Linear(16 + vec_dim, 64), nn.ReLU(), nn.Linear(64, n_actions))

    def forward(self, obs, vec):
        return self.head(torch.cat([self.cnn(obs), vec], dim=1))


class Critic(nn.Module):
    def __init__(self, gvec_dim):
        super().__init__()
        self.cnn = nn.Sequential(nn.Conv2d(4, 16, 3, padding=1), nn.ReLU(), nn.AdaptiveAvgPool2d(1), nn.Flatten())
        self.head = nn.Sequential(nn.Linear(16 + gvec_dim, 64), nn.ReLU(), nn.Linear(64, 1))

    def forward(self, gstate, gvec):
        return self.head(torch.cat([self.cnn(gstate), gvec], dim=1)).squeeze(-1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--parts", type=int, default=1, help="env ranges, each with its own stream for its policy evaluation + env launch")
    args = ap.parse_args()
    N, R = args.envs, 5
    vec_env = mdb.VectorizedEnv(mdb.Environment, N, map_file="map1.txt", n_robots=R, n_packages=50, max_time_steps=500, seed=42,
                                k_others=R - 1, k_packages=5)  # the featuriser sizes of MAPPO/trainer.py:255-256
    b = vec_env.batch
    b.reset()
    actor, critic = Actor(b.vec_dim).cuda(), Critic(b.gvec_dim).cuda()

    def policy(obs, vec, gstate, gvec):  # -> int64 actions [N, R], log-probs [N, R], values [N]
        logits = actor(obs.flatten(0, 1), vec.flatten(0, 1)).view(obs.shape[0], R, -1)  # (a range of the envs when --parts > 1)
        dist = torch.distributions.Categorical(logits=logits)
        a = dist.sample()
        return a, dist.log_prob(a), critic(gstate, gvec)

    buffers = None
    for it in range(args.iters):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = collect_rollouts(vec_env, policy, critic, args.steps, gamma=0.99, gae_lambda=0.95, buffers=buffers, fresh=buffers is None,
                               parts=args.parts)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        b_obs, b_vec, b_gs, b_gv, b_act, b_lp, b_adv, b_ret, rewards, buffers = out
        stats = dict(zip(mdb.STAT_NAMES, b.episode_stats().tolist()))
        print(f"iter {it}: {N * args.steps / dt / 1e6:6.2f} M env-steps/s incl. the networks | batch obs {tuple(b_obs.shape)} "
              f"adv {tuple(b_adv.shape)} | mean shaped reward {rewards.mean().item():+.4f} | episodes done {stats['episodes_done']}")


if __name__ == "__main__":
    main()
