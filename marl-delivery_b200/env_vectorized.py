"""Drop-in for `MAPPO/env_vectorized.py` (== IDQ/, qmix/) and `QMIX/env_vectorized.py` (adds
`indices=`): `VectorizedEnv(env_cls, num_envs, **env_kwargs)` with `.envs`, `.num_envs`, `reset`,
`step`.  All environments live in ONE DeliveryBatch on the GPU; the list-of-dicts returns are the
reference-compatible slow path, `step_tensor` / `featurize` / `rollout_step` are the fast path.
"""
from __future__ import annotations

import numpy as np
import torch

from .batch import MAPPO_SHAPING, DeliveryBatch, encode_action
from .env import EnvView
from .maps import load_map

_ENV_KW = ("map_file", "max_time_steps", "n_robots", "n_packages", "move_cost", "delivery_reward", "delay_reward", "seed")


class VectorizedEnv:
    def __init__(self, env_cls, num_envs, k_others=100, k_packages=100, max_robots_in_state=100,
                 max_packages_in_state=100, shaping=MAPPO_SHAPING, device=None, ring_episodes=None, first_env_index=0,
                 bfs_shaping=False, stale_tracker_compat=False,
                 **env_kwargs):
        """`env_cls` is accepted for signature compatibility (the reference passes its Environment
        class); the environments are always slots of one CUDA batch.  Seeds follow
        MAPPO/env_vectorized.py:4-10: env idx uses seed + idx (all envs share the default seed 2025 when
        no seed is given, exactly like the reference).  `map_file` may be a list, then env idx uses
        map_files[idx % len] (4-map notebook).  `first_env_index` offsets idx for multi-GPU sharding so
        that results do not depend on the number of ranks.  `stale_tracker_compat=True` reproduces what MAPPO/trainer.py
        really feeds the converters and compute_shaped_rewards: its per-env package dict is never cleared on auto-reset
        (MAPPO/trainer.py:229-234), so stale entries shadow new packages; the default is the env truth (= a
        tracker cleared on reset, what the QMIX / IDQ trainers keep).  Call reset() before the first step, like the trainer."""
        unknown = set(env_kwargs) - set(_ENV_KW)
        if unknown:
            raise TypeError(f"unexpected Environment arguments: {sorted(unknown)}")
        if "map_file" not in env_kwargs:
            raise TypeError("Environment.__init__() missing 1 required positional argument: 'map_file'")
        map_file = env_kwargs["map_file"]
        files = list(map_file) if isinstance(map_file, (list, tuple)) and not _is_grid(map_file) else [map_file]
        grids = [load_map(f) for f in files]
        gidx = np.arange(first_env_index, first_env_index + num_envs)
        base_seed = env_kwargs.get("seed", None)
        seeds = np.full(num_envs, 2025, np.int64) if base_seed is None else (int(base_seed) + gidx).astype(np.int64)
        self.batch = DeliveryBatch(
            grids, num_envs, env_kwargs.get("n_robots", 5), env_kwargs.get("n_packages", 20),
            env_kwargs.get("max_time_steps", 100), env_kwargs.get("move_cost", -0.01),
            env_kwargs.get("delivery_reward", 10.), env_kwargs.get("delay_reward", 1.), seeds=seeds,
            env_map=(gidx % len(grids)).astype(np.int32) if len(grids) > 1 else None, k_others=k_others,
            k_packages=k_packages, max_robots_in_state=max_robots_in_state,
            max_packages_in_state=max_packages_in_state, shaping=shaping, device=device,
            bfs_shaping=bfs_shaping, stale_tracker=stale_tracker_compat)
        self.num_envs = num_envs
        self.envs = _EnvList(self.batch, [files[int(g) % len(files)] for g in gidx] if len(files) > 1 else files * 1)

    # ------------------------------------------------------------------ reference API (lists of dicts)
    def reset(self, indices=None):
        idx = list(range(self.num_envs)) if indices is None else [int(i) for i in indices]
        self.batch.reset(None if indices is None else idx)
        snaps = self.batch.get_states(None if indices is None else idx)  # one round trip for all listed envs
        for i, snap in zip(idx, snaps):
            self.envs[i]._seed(snap)
            self.envs[i].done = False
        return [self.envs[i].get_state() for i in idx]

    def step(self, actions, indices=None):
        idx = list(range(self.num_envs)) if indices is None else [int(i) for i in indices]
        idx = idx[:len(actions)]  # zip() semantics of the reference
        if not idx:
            raise ValueError("not enough values to unpack (expected 4, got 0)")
        R = self.batch.n_robots
        codes = np.zeros((len(idx), R), np.uint8)
        for k, acts in enumerate(actions[:len(idx)]):
            if len(acts) != R:
                raise ValueError("The number of actions must match the number of robots.")
            codes[k] = [encode_action(m, p) for (m, p) in acts]
        reward, _, done, rows = self.batch.step_host(codes, None if indices is None and len(idx) == self.num_envs else idx)
        rewards = reward.tolist()
        dones = done.tolist()
        snaps = self.batch._rows_to_states(rows)
        states, infos = [], []
        for k, i in enumerate(idx):
            e = self.envs[i]
            e._seed(snaps[k])
            info = {}
            if dones[k]:
                info["total_reward"] = e.total_reward
                info["total_time_steps"] = e.t
            states.append(e.get_state())
            infos.append(info)
        return states, rewards, dones, infos

    def render(self, indices=None):
        raise NotImplementedError("rendering is not part of the B200 hot path")

    # ------------------------------------------------------------------ tensor fast path
    def step_tensor(self, actions, indices=None, auto_reset=False):
        """actions: uint8 [n, R] codes (a % 5 -> 'DLRSU', a // 5 -> package op).  -> reward f64, shaped f32,
        done u8 tensors on the GPU."""
        return self.batch.step(actions, ids=indices, auto_reset=auto_reset)

    def featurize(self, indices=None, out=None):
        return self.batch.featurize(ids=indices, out=out)

    def rollout_step(self, actions, auto_reset=True, out=None, feat_out=None):
        """One fused launch per MAPPO rollout iteration (MAPPO/trainer.py:172-286 without the nets)."""
        return self.batch.step_featurize(actions, auto_reset=auto_reset, out=out, feat_out=feat_out)

    def rollout_graph(self, n_steps, auto_reset=True):
        """`n_steps` env steps captured in ONE CUDA graph (the hot loops of MAPPO/trainer.py:172 / QMIX/trainer.py:435 call
        vec_env.step once per step; for a few thousand envs a step is a ~7 us kernel and the host is the bottleneck); see
        RolloutGraph."""
        return RolloutGraph(self, n_steps, auto_reset=auto_reset)

    def host_pipeline(self, feat_out=None, auto_reset=True):
        """Double-buffered host <-> device stepping for callers whose actions live in HOST memory; see HostPipeline."""
        return HostPipeline(self, feat_out=feat_out, auto_reset=auto_reset)


class RolloutGraph:
    """`n_steps` consecutive `step_tensor(..., auto_reset)` calls over all envs as one CUDA graph.

        g = ve.rollout_graph(64)
        reward, shaped, done = g.run(actions)     # actions: uint8 [64, N, R] codes (copied into the graph's input buffer)

    `run` replays the graph on the current stream and returns the graph's OUTPUT BUFFERS ([n_steps, N] each, overwritten by
    the next run).  Results are identical to calling step_tensor step by step (tests/test_gpu_r2.py).  The launches are enqueued by
    one C call (include/mdx.h: mdx_step_n), so the same speed is also available without a graph through
    `ve.batch.step_n(actions)`."""

    def __init__(self, ve, n_steps, auto_reset=True):
        b = ve.batch
        self.batch, self.n_steps, self.auto_reset = b, int(n_steps), auto_reset
        dev, N, R = b.device, b.n_envs, b.n_robots
        self.actions = torch.zeros((self.n_steps, N, R), dtype=torch.uint8, device=dev)
        self.out = (torch.empty((self.n_steps, N), dtype=torch.float64, device=dev),
                    torch.empty((self.n_steps, N), dtype=torch.float32, device=dev),
                    torch.empty((self.n_steps, N), dtype=torch.uint8, device=dev))
        # one real launch before the capture (loads the kernel), on a snapshot that is put back afterwards
        blob = b.state_dict()
        b.step_n(self.actions[:1], auto_reset=auto_reset)
        torch.cuda.synchronize(dev)
        b.load_state_dict(blob)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            b.step_n(self.actions, auto_reset=auto_reset, out=self.out)
        b.load_state_dict(blob)  # (capturing does not run anything, but keep the version counter / views honest)

    def run(self, actions=None):
        if actions is not None:
            self.actions.copy_(actions, non_blocking=True)
        self.graph.replay()
        self.batch._version += 1
        return self.out


class HostPipeline:
    """`rollout_step` for actions that come from pinned HOST memory and rewards/dones that are wanted back on the host,
    with the copies taken off the critical path: two slots of device/host staging, one copy stream per direction.

        slot = pipe.submit(h_actions)      # H2D copy of step n's actions, fused launch, D2H of (shaped, done): all async
        shaped, done = pipe.result(slot)   # numpy views of the pinned result buffers, valid until the slot is reused

    While step n runs on the GPU, the copy of step n+1's actions and the read-back of step n-1's results are in flight, so
    a step costs one kernel.  Feature tensors stay in HBM (`feat_out`, where the policy / critic read them on the launch
    stream).  Results are the same as calling `rollout_step` step by step (tests/test_gpu_next.py)."""

    def __init__(self, ve, feat_out=None, auto_reset=True):
        b = ve.batch
        self.batch, self.auto_reset = b, auto_reset
        n, R, dev = b.n_envs, b.n_robots, b.device
        self.n = n
        self.feat = feat_out if feat_out is not None else b.alloc_features(n)
        self.main = torch.cuda.current_stream(dev)
        self.s_in, self.s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        self.d_act = [torch.empty((n, R), dtype=torch.uint8, device=dev) for _ in range(2)]
        # [shaped f32 | done u8] share one allocation per slot: one D2H copy per step
        self.d_sd = [torch.empty(5 * n, dtype=torch.uint8, device=dev) for _ in range(2)]
        self.d_rew = torch.empty(n, dtype=torch.float64, device=dev)
        self.h_sd = [torch.empty(5 * n, dtype=torch.uint8).pin_memory() for _ in range(2)]
        self.outs = [(self.d_rew, d[:4 * n].view(torch.float32), d[4 * n:]) for d in self.d_sd]  # (reward, shaped, done) per slot
        self.ev_in = [torch.cuda.Event() for _ in range(2)]
        self.ev_k = [torch.cuda.Event() for _ in range(2)]
        self.ev_out = [torch.cuda.Event() for _ in range(2)]
        self.used = [False, False]
        self.count = 0

    def submit(self, h_actions):
        """h_actions: uint8 [n, R] action codes in (preferably pinned) host memory.  Returns the slot to pass to result()."""
        s = self.count & 1
        self.count += 1
        n = self.n
        if self.used[s]:
            self.ev_out[s].synchronize()          # the host buffers of this slot have been filled (and read by the caller)
            self.s_in.wait_event(self.ev_k[s])    # the kernel that read d_act[s] two steps ago is done
        with torch.cuda.stream(self.s_in):
            self.d_act[s].copy_(h_actions, non_blocking=True)
            self.ev_in[s].record(self.s_in)
        self.main.wait_event(self.ev_in[s])
        if self.used[s]:
            self.main.wait_event(self.ev_out[s])  # d_sd[s] has been read back
        with torch.cuda.stream(self.main):
            self.batch.step_featurize(self.d_act[s], auto_reset=self.auto_reset, out=self.outs[s], feat_out=self.feat)
            self.ev_k[s].record(self.main)
        self.s_out.wait_event(self.ev_k[s])
        with torch.cuda.stream(self.s_out):
            self.h_sd[s].copy_(self.d_sd[s], non_blocking=True)
            self.ev_out[s].record(self.s_out)
        self.used[s] = True
        return s

    def result(self, slot):
        """Blocks until step `slot`'s results are on the host -> (shaped float32 [n], done uint8 [n]) numpy views."""
        self.ev_out[slot].synchronize()
        h = self.h_sd[slot].numpy()
        return h[:4 * self.n].view(np.float32), h[4 * self.n:]

    def drain(self):
        """Make the launch stream wait for every outstanding copy (call before timing / reusing the buffers elsewhere)."""
        for s in range(2):
            if self.used[s]:
                self.main.wait_event(self.ev_out[s])


class _EnvList:
    """`.envs` of the reference is a list of Environment objects; here the per-env views are created on
    first access (a 65k-env batch should not build 65k Python objects up front)."""

    def __init__(self, batch, files):
        self._b, self._files, self._cache = batch, files, {}

    def __len__(self):
        return self._b.n_envs

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[k] for k in range(*i.indices(len(self)))]
        i = int(i)
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError("list index out of range")
        v = self._cache.get(i)
        if v is None:
            v = self._cache[i] = EnvView(self._b, i, self._files[i % len(self._files)])
        return v

    def __iter__(self):
        return (self[i] for i in range(len(self)))


def _is_grid(x):
    return isinstance(x, np.ndarray) or (len(x) > 0 and isinstance(x[0], (list, tuple, np.ndarray)))
