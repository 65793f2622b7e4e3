"""DeliveryBatch: N environments on one B200 behind the C ABI of include/mdx.h.

This is the tensor-level fast path the drop-in wrappers (env.py, env_vectorized.py, helper.py) sit
on.  PyTorch is used for what it is good at here -- owning the output tensors in HBM, streams,
torch.distributed -- and nothing else: every transition, reward and feature is computed by the
hand-written kernels in csrc/.  No CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import MDX_F_ACTIONS_I64, MDX_F_AUTO_RESET

# compute_shaped_rewards constants (MAPPO/helper.py:271-279, QMIX/helper.py:270-278)
MAPPO_SHAPING = (5.0, 200.0, 20.0, 0.02, 0.0, 0.0, -0.05, -0.05, -0.01)
QMIX_SHAPING = (0.5, 2.0, 0.2, 0.02, -0.1, -0.1, -0.05, -0.02, -0.02)

STATUS_NAMES = ("None", "waiting", "in_transit", "delivered")
MOVE_CODES = {"D": 0, "L": 1, "R": 2, "S": 3, "U": 4}  # sorted LabelEncoder order, MAPPO/trainer.py:84-89
STAT_NAMES = ("episodes_done", "robot_moves", "on_time", "late", "pickups", "env_steps", "sum_episode_len",
              "starved_envs")


def encode_action(move, pkg_act) -> int:
    """(move_str, pkg_str) -> action code of include/mdx.h.  Mirrors the reference's comparisons:
    moves other than S/L/R/U/D leave the robot in place (env.py:333-334), and only the STRINGS '1' and
    '2' are package ops (env.py:264,277; ints are ignored)."""
    op = 1 if (isinstance(pkg_act, str) and pkg_act == "1") else 2 if (isinstance(pkg_act, str) and pkg_act == "2") else 0
    mv = MOVE_CODES.get(move) if isinstance(move, str) else None
    if mv is None:
        return 15 + op
    return mv + 5 * op


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _ptr(t: Optional[torch.Tensor]):
    """Device address as a plain int (ctypes converts it to void*; no c_void_p object per argument)."""
    return None if t is None else t.data_ptr()


class DeliveryBatch:
    """n_envs environments resident on one GPU."""

    def __init__(self, grids: Sequence[np.ndarray], n_envs: int, n_robots: int = 5, n_packages: int = 20,
                 max_time_steps: int = 100, move_cost: float = -0.01, delivery_reward: float = 10.0,
                 delay_reward: float = 1.0, seeds=None, base_seed: int = 2025, env_map=None,
                 k_others: int = 100, k_packages: int = 100, max_robots_in_state: int = 100,
                 max_packages_in_state: int = 100, shaping=MAPPO_SHAPING, device=None, ring_episodes=None,
                 sampler_threads=None, blank: bool = False, bfs_shaping: bool = False, lax_robots: bool = False,
                 stale_tracker: bool = False, specialize=None):
        """`specialize`: compile the fused step + featurise kernel at run time (NVRTC) for this instance's shape, so that its loops unroll
        and its index arithmetic folds (MDX_CFG_JIT_SHAPE; ~1 s per new shape, cached on disk; bit-identical results; +3 to +12 % on B200).
        None = automatically for batches of >= 4 096 envs, where a step is bandwidth- rather than launch-bound.
        `ring_episodes` / `sampler_threads` are accepted and ignored (v1 pre-sampled reset blocks on the host into a ring;
        since v2 every env draws its next reset block on the device, from its own MT19937 stream, when it resets)."""
        L = _lib.lib()
        if not torch.cuda.is_available():
            raise _lib.MdxError("no CUDA device: marl-delivery_b200 runs on the GPU only (no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise _lib.MdxError("marl-delivery_b200 needs a CUDA device")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        grids = [np.ascontiguousarray(np.asarray(g, dtype=np.uint8)) for g in grids]
        for g in grids:
            if g.ndim != 2:
                raise ValueError("a map must be a 2-D grid of 0/1")
        self.grids = grids
        cfg = _lib.Config()
        cfg.n_envs, cfg.n_robots, cfg.n_packages, cfg.max_time_steps = int(n_envs), int(n_robots), int(n_packages), int(max_time_steps)
        cfg.move_cost, cfg.delivery_reward, cfg.delay_reward = float(move_cost), float(delivery_reward), float(delay_reward)
        cfg.shaping = _lib.Shaping(*[float(x) for x in shaping])
        cfg.k_others, cfg.k_packages = int(k_others), int(k_packages)
        cfg.max_robots_in_state, cfg.max_packages_in_state = int(max_robots_in_state), int(max_packages_in_state)
        cfg.n_maps = len(grids)
        rows = (C.c_int32 * len(grids))(*[g.shape[0] for g in grids])
        cols = (C.c_int32 * len(grids))(*[g.shape[1] for g in grids])
        cells = (C.POINTER(C.c_uint8) * len(grids))(*[g.ctypes.data_as(C.POINTER(C.c_uint8)) for g in grids])
        cfg.map_rows, cfg.map_cols, cfg.map_cells = rows, cols, cells
        keep = [rows, cols, cells]
        if env_map is not None:
            em = np.ascontiguousarray(np.asarray(env_map, dtype=np.int32))
            if em.shape != (n_envs,):
                raise ValueError("env_map must have one entry per env")
            cfg.env_map = em.ctypes.data_as(C.POINTER(C.c_int32))
            keep.append(em)
            self.env_map = em.copy()
        else:
            self.env_map = np.zeros(n_envs, np.int32)
        if seeds is not None:
            sd = np.ascontiguousarray(np.asarray(seeds, dtype=np.int64))
            if sd.shape != (n_envs,):
                raise ValueError("seeds must have one entry per env")
            cfg.env_seed = sd.ctypes.data_as(C.POINTER(C.c_int64))
            keep.append(sd)
        cfg.base_seed = int(base_seed)
        cfg.device = int(self.device.index)
        cfg.ring_episodes, cfg.sampler_threads = 0, 0
        cfg.flags = ((_lib.MDX_CFG_BLANK if blank else 0) | (_lib.MDX_CFG_BFS_SHAPING if bfs_shaping else 0)
                     | (_lib.MDX_CFG_LAX_ROBOTS if lax_robots else 0) | (_lib.MDX_CFG_STALE_TRACKER if stale_tracker else 0)
                     | (_lib.MDX_CFG_JIT_SHAPE if (specialize if specialize is not None else (n_envs >= 4096 and not stale_tracker and not bfs_shaping)) else 0))
        self.stale_tracker = bool(stale_tracker)
        handle = C.c_void_p()
        _lib.check(L.mdx_create(C.byref(cfg), C.byref(handle)))
        self._h = handle
        self._L = L
        d = _lib.Dims()
        _lib.check(L.mdx_get_dims(self._h, C.byref(d)))
        self.n_envs, self.n_robots, self.n_packages, self.max_time_steps = d.n_envs, d.n_robots, d.n_packages, d.max_time_steps
        self.rows_out, self.cols_out, self.vec_dim, self.gvec_dim = d.rows_out, d.cols_out, d.vec_dim, d.gvec_dim
        self.bytes_per_env_step = int(d.bytes_per_env_step)
        self.shaping = tuple(float(x) for x in shaping)
        self._dev_index = int(self.device.index)
        self._bufs = {}
        self._ok_out = {}
        self._ok_ids = {}  # validated device id tensors, keyed by (identity, version)  # id(tuple) -> tuple of caller-owned output tensors that passed _check_out
        self._version = 0  # bumped by every call that changes env state (invalidates EnvView snapshots)
        self._grid_lists = [g.tolist() for g in self.grids]
        # packed state rows (include/mdx.h: mdx_get_states): one structured numpy record per env
        R, P = self.n_robots, self.n_packages
        Pp8, Pp16 = max(8, (P + 7) // 8 * 8), max(16, (P + 15) // 16 * 16)
        off_sched = 4 * ((6 + R + 1) // 2 * 2)
        self._row_bytes = int(L.mdx_state_row_bytes(self._h))
        self._row_dtype = np.dtype({"names": ["t", "episode", "map_id", "total_reward", "robots", "sched", "status"],
                                    "formats": ["<i4", "<i4", "<i4", "<f8", ("<u4", (R,)), ("<u2", (4, Pp8)), ("u1", (Pp16,))],
                                    "offsets": [0, 4, 8, 16, 24, off_sched, off_sched + 8 * Pp8], "itemsize": self._row_bytes})
        self._host = {}  # n -> reusable host arrays of the packed-row entry points

    def jit_status(self):
        """-> (state, message): 1 = the fused kernel compiled at run time for this shape (specialize=True) is in use, 0 = not
        requested, -1 = requested but unavailable (the precompiled run-time-shape kernels run; message says why)."""
        buf = C.create_string_buffer(512)
        st = int(self._L.mdx_jit_status(self._h, buf, 512))
        return st, buf.value.decode("utf-8", "replace")

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.mdx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        """The caller's current stream on this device (every kernel is enqueued there)."""
        if _raw_stream is not None:  # the handle without building a torch.cuda.Stream object (a few microseconds per launch)
            return _raw_stream(self._dev_index) or None
        return torch.cuda.current_stream(self.device).cuda_stream or None

    def _buf(self, name, shape, dtype):
        key = (name, tuple(shape), dtype)
        t = self._bufs.get(key)
        if t is None:
            t = torch.empty(shape, dtype=dtype, device=self.device)
            self._bufs[key] = t
        return t

    def _check_out(self, tensors, specs):
        """Caller-owned output tensors are written by the kernels without further checks, so a wrong device, dtype,
        size or a non-contiguous view must be refused here (None = output not wanted).  A tuple that passed once is
        remembered by identity: steady-state loops that reuse their buffers pay nothing."""
        key = id(tensors)
        hit = self._ok_out.get(key)
        if hit is not None and hit[0] is tensors:  # (only tuples are remembered: their members cannot be swapped)
            return hit[1]
        if len(tensors) != len(specs):
            raise ValueError(f"expected {len(specs)} output tensors, got {len(tensors)}")
        for t, (name, numel, dtype) in zip(tensors, specs):
            if t is None:
                continue
            on_device = isinstance(t, torch.Tensor) and t.is_cuda and t.device == self.device
            zero_copy = isinstance(t, torch.Tensor) and t.device.type == "cpu" and t.is_pinned()  # written by the kernel over PCIe
            if not (on_device or zero_copy):
                raise ValueError(f"output '{name}' must be a CUDA tensor on {self.device} (or a pinned host tensor: zero-copy)")
            if t.dtype != dtype or t.numel() != numel or not t.is_contiguous():
                raise ValueError(f"output '{name}' must be a contiguous {dtype} tensor of {numel} elements "
                                 f"(got {t.dtype}, {t.numel()} elements, contiguous={t.is_contiguous()})")
        ptrs = tuple(None if t is None else t.data_ptr() for t in tensors)
        if isinstance(tensors, tuple):
            if len(self._ok_out) > 64:
                self._ok_out.clear()
            self._ok_out[key] = (tensors, ptrs)  # the tuple keeps its tensors (and so their addresses) alive
        return ptrs

    def _rsd_specs(self, n, second):
        return (("reward", n, torch.float64), second, ("done", n, torch.uint8))

    def _feat_specs(self, n):
        R, HW = self.n_robots, self.rows_out * self.cols_out
        return (("obs", n * R * 6 * HW, torch.float32), ("vec", n * R * self.vec_dim, torch.float32),
                ("gstate", n * 4 * HW, torch.float32), ("gvec", n * self.gvec_dim, torch.float32))

    def _alt_specs(self, n):
        R, HW = self.n_robots, self.rows_out * self.cols_out
        return (("obs6", n * R * 6 * HW, torch.float32), ("gstate7", n * 7 * HW, torch.float32))

    def _ids(self, ids, unique=True):
        """-> (device int32 tensor or None, host list or None, n).  Ids must be in range and pairwise distinct (two lane
        groups stepping one env in the same launch would race; the reference would step it twice in sequence).  Host
        lists are checked here; a device tensor is checked on the device ONCE per (tensor, version) -- one small
        synchronous read the first time an ids tensor is used, nothing afterwards unless it is modified in place."""
        if ids is None:
            return None, None, self.n_envs
        if isinstance(ids, torch.Tensor):
            d = ids.to(device=self.device, dtype=torch.int32).contiguous()
            key = (id(ids), ids._version, ids.data_ptr(), int(ids.numel()), unique)
            if self._ok_ids.get(key) is not ids:
                n = int(d.numel())
                if n and (int(d.min()) < 0 or int(d.max()) >= self.n_envs):
                    raise IndexError("env index out of range")
                if unique and n > 1 and int(torch.unique(d).numel()) != n:
                    raise ValueError("an env index is listed twice")
                if len(self._ok_ids) > 64:
                    self._ok_ids.clear()
                self._ok_ids[key] = ids
            return d, None, int(d.numel())
        h = [int(i) for i in ids]
        for i in h:
            if not 0 <= i < self.n_envs:
                raise IndexError(f"env index {i} out of range")
        if unique and len(set(h)) != len(h):
            raise ValueError("an env index is listed twice")
        return torch.tensor(h, dtype=torch.int32, device=self.device), h, len(h)

    # ------------------------------------------------------------------ reset / prefetch
    def reset(self, ids=None):
        """Environment.reset for the listed envs (all when None): next host-drawn reset block."""
        if ids is None:
            _lib.check(self._L.mdx_reset(self._h, None, 0, self._stream()))
            self._version += 1
            return
        h = np.ascontiguousarray(np.asarray([int(i) for i in ids], dtype=np.int32))
        _lib.check(self._L.mdx_reset(self._h, h.ctypes.data_as(C.POINTER(C.c_int32)), int(h.size), self._stream()))
        self._version += 1

    def prefetch(self, episodes_ahead: int = 0):
        """No-op kept for v1 callers: reset blocks are no longer pre-sampled on the host -- an env that finishes under
        `auto_reset` draws its next block on the device right inside the step kernel, so it can never starve."""
        _lib.check(self._L.mdx_prefetch(self._h, int(episodes_ahead), self._stream()))

    # ------------------------------------------------------------------ step
    def _check_actions(self, actions, n):
        """-> (device tensor, flag bits).  uint8 action codes are used as they are; an int64 tensor is taken
        to be the policy's raw samples (mb_actions of MAPPO/trainer.py:193) and decoded inside the kernel."""
        if (actions.__class__ is torch.Tensor and actions.dtype is torch.uint8 and actions.is_cuda and actions.is_contiguous()
                and actions.dim() == 2 and actions.shape[0] == n and actions.shape[1] == self.n_robots
                and actions.get_device() == self._dev_index):
            return actions, 0  # the steady-state case: device-resident action codes of the right shape
        if (actions.__class__ is torch.Tensor and actions.device.type == "cpu" and actions.dtype in (torch.uint8, torch.int64) and actions.is_pinned()
                and actions.is_contiguous() and tuple(actions.shape) == (n, self.n_robots)):
            # ZERO-COPY: pinned host memory is device-addressable (unified addressing); the kernel reads the actions over PCIe itself,
            # no copy engine, no staging tensor (the caller must not rewrite the buffer before the launch has run)
            return actions, (MDX_F_ACTIONS_I64 if actions.dtype is torch.int64 else 0)
        if not isinstance(actions, torch.Tensor):
            actions = torch.as_tensor(np.asarray(actions, dtype=np.uint8))
        flags = 0
        if actions.dtype == torch.int64:
            flags = MDX_F_ACTIONS_I64
        elif actions.dtype != torch.uint8:
            actions = actions.to(torch.uint8)
        actions = actions.to(self.device).contiguous()
        if tuple(actions.shape) != (n, self.n_robots):
            raise ValueError("The number of actions must match the number of robots.")  # env.py:182-183
        return actions, flags

    def step(self, actions, ids=None, auto_reset: bool = False, want_shaped: bool = True, out=None):
        """-> (reward f64 [n], shaped f32 [n] or None, done u8 [n]); rows follow `ids`.

        With `out=None` the returned tensors are buffers OWNED BY THIS INSTANCE and shared between step(), step_idq() and
        step_featurize(): they are valid until the next call of any of them.  Pass `out=(reward, shaped, done)` (a tuple
        of caller-owned CUDA tensors; validated once, then remembered by identity) to keep results across steps."""
        if ids is None:
            d_ids, n = None, self.n_envs
        else:
            d_ids, _, n = self._ids(ids)
        actions, aflags = self._check_actions(actions, n)
        if out is None:
            reward = self._buf("reward", (n,), torch.float64)
            shaped = self._buf("shaped", (n,), torch.float32) if want_shaped else None
            done = self._buf("done", (n,), torch.uint8)
            p_r, p_s, p_d = reward.data_ptr(), (None if shaped is None else shaped.data_ptr()), done.data_ptr()
        else:
            p_r, p_s, p_d = self._check_out(out, self._rsd_specs(n, ("shaped", n, torch.float32)))
            reward, shaped, done = out
        rc = self._L.mdx_step(self._h, actions.data_ptr(), _ptr(d_ids), n, p_r, p_s, p_d,
                              (MDX_F_AUTO_RESET if auto_reset else 0) | aflags, self._stream())
        if rc:
            _lib.check(rc)
        self._version += 1
        return reward, shaped, done

    def stepper(self, out, feat_out=None, auto_reset: bool = True):
        """A pre-bound `step` / `step_featurize` for hot loops over ALL envs with fixed caller-owned outputs: everything that does
        not change from step to step (output addresses, flags, the C function, the validation of `out` / `feat_out`) is resolved
        once; the returned callable takes the step's uint8 action tensor [n_envs, R] on this device and costs one ctypes call
        (launch-bound batches: ~2 us less per step than step(); results are identical)."""
        n, R = self.n_envs, self.n_robots
        p_r, p_s, p_d = self._check_out(out, self._rsd_specs(n, ("shaped", n, torch.float32)))
        flags = MDX_F_AUTO_RESET if auto_reset else 0
        h, idx, raw = self._h, self._dev_index, _raw_stream
        u8 = torch.uint8
        if feat_out is None:
            fn = self._L.mdx_step

            def call(actions):
                if actions.dtype is not u8 or actions.shape[0] != n or actions.get_device() != idx or not actions.is_contiguous():
                    raise ValueError("actions must be a contiguous uint8 tensor [n_envs, n_robots] on the batch's device")
                rc = fn(h, actions.data_ptr(), None, n, p_r, p_s, p_d, flags, raw(idx) or None)
                if rc:
                    _lib.check(rc)
                self._version += 1
        else:
            fp = self._check_out(feat_out, self._feat_specs(n))
            fn = self._L.mdx_step_featurize

            def call(actions):
                if actions.dtype is not u8 or actions.shape[0] != n or actions.get_device() != idx or not actions.is_contiguous():
                    raise ValueError("actions must be a contiguous uint8 tensor [n_envs, n_robots] on the batch's device")
                rc = fn(h, actions.data_ptr(), None, n, p_r, p_s, p_d, fp[0], fp[1], fp[2], fp[3], flags, raw(idx) or None)
                if rc:
                    _lib.check(rc)
                self._version += 1
        call.out, call.feat_out = out, feat_out  # keep the tensors (and so their addresses) alive as long as the callable
        return call

    def step_n(self, actions, auto_reset: bool = True, out=None):
        """`n_steps` consecutive step() calls over all envs in ONE C call (include/mdx.h: mdx_step_n): actions uint8
        [n_steps, n_envs, R] on the GPU -> (reward f64, shaped f32, done u8), each [n_steps, n_envs].  For launch-bound
        batches the per-step cost is the kernel, not the binding."""
        if not (isinstance(actions, torch.Tensor) and actions.is_cuda and actions.dtype == torch.uint8 and actions.is_contiguous()
                and actions.dim() == 3 and tuple(actions.shape[1:]) == (self.n_envs, self.n_robots)):
            raise ValueError("actions must be a contiguous uint8 CUDA tensor [n_steps, n_envs, n_robots]")
        S, n = int(actions.shape[0]), self.n_envs
        if out is None:
            out = (torch.empty((S, n), dtype=torch.float64, device=self.device), torch.empty((S, n), dtype=torch.float32, device=self.device),
                   torch.empty((S, n), dtype=torch.uint8, device=self.device))
        p_r, p_s, p_d = self._check_out(out, (("reward", S * n, torch.float64), ("shaped", S * n, torch.float32), ("done", S * n, torch.uint8)))
        rc = self._L.mdx_step_n(self._h, actions.data_ptr(), S, p_r, p_s, p_d, MDX_F_AUTO_RESET if auto_reset else 0, self._stream())
        if rc:
            _lib.check(rc)
        self._version += 1
        return out

    def alloc_features(self, n=None, obs=True, vec=True, gstate=True, gvec=True):
        n = self.n_envs if n is None else n
        R, H, W = self.n_robots, self.rows_out, self.cols_out
        mk = lambda *s: torch.empty(s, dtype=torch.float32, device=self.device)  # noqa: E731
        return (mk(n, R, 6, H, W) if obs else None, mk(n, R, self.vec_dim) if vec else None,
                mk(n, 4, H, W) if gstate else None, mk(n, self.gvec_dim) if gvec else None)

    def _feat_out(self, n, out):
        if out is not None:
            self._check_out(out, self._feat_specs(n))
            return out
        key = ("feat", n)
        if key not in self._bufs:
            self._bufs[key] = self.alloc_features(n)
        return self._bufs[key]

    def featurize(self, ids=None, out=None):
        """-> (obs [n,R,6,H,W], vec [n,R,D], gstate [n,4,H,W], gvec [n,Dg]) float32 on the GPU."""
        d_ids, _, n = self._ids(ids, unique=False)  # read-only: repeated ids just produce repeated rows
        obs, vec, gstate, gvec = self._feat_out(n, out)
        _lib.check(self._L.mdx_featurize(self._h, _ptr(d_ids), n, _ptr(obs), _ptr(vec), _ptr(gstate), _ptr(gvec),
                                         self._stream()))
        return obs, vec, gstate, gvec

    def step_featurize(self, actions, ids=None, auto_reset: bool = True, out=None, feat_out=None):
        """One launch per MAPPO rollout iteration (MAPPO/trainer.py:172-286 minus the networks).  Like step(): with
        `out=None` / `feat_out=None` the results live in instance-owned buffers that the next call overwrites."""
        if ids is None:
            d_ids, n = None, self.n_envs
        else:
            d_ids, _, n = self._ids(ids)
        actions, aflags = self._check_actions(actions, n)
        if out is None:
            reward = self._buf("reward", (n,), torch.float64)
            shaped = self._buf("shaped", (n,), torch.float32)
            done = self._buf("done", (n,), torch.uint8)
            p_r, p_s, p_d = reward.data_ptr(), shaped.data_ptr(), done.data_ptr()
        else:
            p_r, p_s, p_d = self._check_out(out, self._rsd_specs(n, ("shaped", n, torch.float32)))
            reward, shaped, done = out
        if feat_out is None:
            feat_out = self._feat_out(n, None)
            fp = tuple(None if t is None else t.data_ptr() for t in feat_out)
        else:
            fp = self._check_out(feat_out, self._feat_specs(n))
        obs, vec, gstate, gvec = feat_out
        rc = self._L.mdx_step_featurize(self._h, actions.data_ptr(), _ptr(d_ids), n, p_r, p_s, p_d, fp[0], fp[1], fp[2], fp[3],
                                        (MDX_F_AUTO_RESET if auto_reset else 0) | aflags, self._stream())
        if rc:
            _lib.check(rc)
        self._version += 1
        return reward, shaped, done, obs, vec, gstate, gvec

    def shaped_reward(self, cur_robots, cur_t, actions, global_reward):
        """compute_shaped_rewards with the stored state as 'previous' (MAPPO/helper.py:257-369)."""
        n = self.n_envs
        cur_robots = cur_robots.to(self.device, torch.int64).contiguous()
        packed = (cur_robots[..., 0] | (cur_robots[..., 1] << 16)).to(torch.int32).contiguous()  # cell | carry << 16
        cur_t = cur_t.to(self.device, torch.int32).contiguous()
        actions, aflags = self._check_actions(actions, n)
        if aflags:
            raise ValueError("shaped_reward takes uint8 action codes")
        gr = global_reward.to(self.device, torch.float64).contiguous()
        out = torch.empty(n, dtype=torch.float32, device=self.device)
        _lib.check(self._L.mdx_shaped_reward(self._h, _ptr(packed), _ptr(cur_t), _ptr(actions), _ptr(gr), _ptr(out),
                                             self._stream()))
        return out

    def shaped_reward_host(self, cur_robots: np.ndarray, cur_t: np.ndarray, codes: np.ndarray, global_reward: np.ndarray) -> np.ndarray:
        """shaped_reward() for HOST callers: cur_robots uint32 [n, R] = cell | carry << 16, cur_t int32 [n], codes uint8 [n, R],
        global_reward float64 [n] -> float32 [n]; one launch + one synchronisation inside one C call."""
        n, R = self.n_envs, self.n_robots
        cur_robots = np.ascontiguousarray(cur_robots, dtype=np.uint32).reshape(n, R)
        cur_t = np.ascontiguousarray(cur_t, dtype=np.int32).reshape(n)
        codes = np.ascontiguousarray(codes, dtype=np.uint8).reshape(n, R)
        gr = np.ascontiguousarray(global_reward, dtype=np.float64).reshape(n)
        out = np.empty(n, np.float32)
        _lib.check(self._L.mdx_shaped_reward_host(self._h, cur_robots.ctypes.data, cur_t.ctypes.data, codes.ctypes.data, gr.ctypes.data,
                                                  out.ctypes.data, self._stream()))
        return out

    # ------------------------------------------------------------------ callers either side of the path
    def greedy_actions(self, ids=None, out=None) -> torch.Tensor:
        """GreedyAgents.get_actions (greedyagent.py:119-174) for the listed envs -> uint8 [n, R] action codes on
        the GPU, ready for step().  Call once per env step; the agent re-initialises itself at t = 0."""
        d_ids, _, n = self._ids(ids)
        if out is None:
            out = torch.empty((n, self.n_robots), dtype=torch.uint8, device=self.device)
        else:
            self._check_out((out,), (("actions", n * self.n_robots, torch.uint8),))
        _lib.check(self._L.mdx_greedy_actions(self._h, _ptr(d_ids), n, _ptr(out), self._stream()))
        return out

    def featurize_alt(self, ids=None, obs6=True, gstate7=True, out=None):
        """IDQ / legacy-qmix convert_state (IDQ/networks.py:112-217) for every robot and the legacy 7-channel
        global tensor (qmix/networks.py:350-468) -> (obs6 [n,R,6,H,W], gstate7 [n,7,H,W]) float32 on the GPU."""
        d_ids, _, n = self._ids(ids, unique=False)
        R, H, W = self.n_robots, self.rows_out, self.cols_out
        if out is None:
            out = (self._buf("obs6", (n, R, 6, H, W), torch.float32) if obs6 else None,
                   self._buf("gstate7", (n, 7, H, W), torch.float32) if gstate7 else None)
        else:
            self._check_out(out, self._alt_specs(n))
        _lib.check(self._L.mdx_featurize_alt(self._h, _ptr(d_ids), n, _ptr(out[0]), _ptr(out[1]), self._stream()))
        return out

    def step_idq(self, actions, ids=None, auto_reset: bool = False, ops_as_int: bool = False, out=None):
        """step + IDQ's per-agent reward_shaping (IDQ/networks.py:228-349) -> (reward f64 [n], agent rewards
        f64 [n, R], done u8 [n]).  ops_as_int=False is what IDQ/trainer.py does (string ops: stay penalty only)."""
        d_ids, _, n = self._ids(ids)
        actions, aflags = self._check_actions(actions, n)
        if out is None:
            out = (self._buf("reward", (n,), torch.float64), self._buf("agent_rewards", (n, self.n_robots), torch.float64),
                   self._buf("done", (n,), torch.uint8))
        else:
            self._check_out(out, self._rsd_specs(n, ("agent_rewards", n * self.n_robots, torch.float64)))
        reward, agent, done = out
        _lib.check(self._L.mdx_step_idq(self._h, _ptr(actions), _ptr(d_ids), n, _ptr(reward), _ptr(agent), _ptr(done),
                                        1 if ops_as_int else 0, (MDX_F_AUTO_RESET if auto_reset else 0) | aflags, self._stream()))
        self._version += 1
        return reward, agent, done

    def step_idq_featurize(self, actions, ids=None, auto_reset: bool = False, ops_as_int: bool = False, out=None, feat_out=None,
                           env_range=None, stream=None):
        """One IDQ trainer iteration in ONE launch (IDQ/trainer.py:276-300): step_idq, then featurize_alt of the state after
        the step (and after any auto-reset) -> (reward, agent rewards, done, obs6, gstate7).

        `env_range=(first_env, n)`: the contiguous env range only -- `actions`, `out` and `feat_out` are that range's rows (e.g.
        slices of full-batch tensors) -- launched on `stream` (a torch.cuda.Stream; default: the current one), so that several
        ranges of one instance can be in flight at once (see VectorizedEnv.split for the MAPPO path)."""
        if env_range is not None:
            if ids is not None:
                raise ValueError("give ids or env_range, not both")
            lo, n = int(env_range[0]), int(env_range[1])
            if lo < 0 or n < 1 or lo + n > self.n_envs:
                raise ValueError("env_range outside [0, n_envs)")
            d_ids = None
        else:
            d_ids, _, n = self._ids(ids)
        actions, aflags = self._check_actions(actions, n)
        if out is None:
            out = (self._buf("reward", (n,), torch.float64), self._buf("agent_rewards", (n, self.n_robots), torch.float64),
                   self._buf("done", (n,), torch.uint8))
        else:
            self._check_out(out, self._rsd_specs(n, ("agent_rewards", n * self.n_robots, torch.float64)))
        reward, agent, done = out
        R, H, W = self.n_robots, self.rows_out, self.cols_out
        if feat_out is None:
            feat_out = (self._buf("obs6", (n, R, 6, H, W), torch.float32), self._buf("gstate7", (n, 7, H, W), torch.float32))
        else:
            self._check_out(feat_out, self._alt_specs(n))
        flags = (MDX_F_AUTO_RESET if auto_reset else 0) | aflags
        st = self._stream() if stream is None else stream.cuda_stream
        if env_range is not None:
            _lib.check(self._L.mdx_step_idq_featurize_range(self._h, _ptr(actions), lo, n, _ptr(reward), _ptr(agent), _ptr(done),
                                                            _ptr(feat_out[0]), _ptr(feat_out[1]), 1 if ops_as_int else 0, flags, st))
        else:
            _lib.check(self._L.mdx_step_idq_featurize(self._h, _ptr(actions), _ptr(d_ids), n, _ptr(reward), _ptr(agent), _ptr(done),
                                                      _ptr(feat_out[0]), _ptr(feat_out[1]), 1 if ops_as_int else 0, flags, st))
        self._version += 1
        return reward, agent, done, feat_out[0], feat_out[1]

    def bfs_table(self, map_id: int = 0) -> np.ndarray:
        """All-pairs shortest-path lengths of one map, uint16 [H*W, H*W] in output-plane cell ids (0xFFFF =
        unreachable): run_bfs (greedyagent.py:6-29) from every free cell."""
        hw = self.rows_out * self.cols_out
        out = np.empty((hw, hw), np.uint16)
        _lib.check(self._L.mdx_bfs_table(self._h, int(map_id), out.ctypes.data_as(C.POINTER(C.c_uint16))))
        return out

    def state_dict(self) -> torch.Tensor:
        """Complete env state (device arrays, episode ring, counters, greedy bookkeeping, every env's RNG stream)
        as one uint8 host tensor; `load_state_dict` makes an instance of the same shape continue bit-identically."""
        nbytes = int(self._L.mdx_checkpoint_bytes(self._h))
        buf = torch.empty(nbytes, dtype=torch.uint8)
        _lib.check(self._L.mdx_checkpoint_save(self._h, C.c_void_p(buf.data_ptr()), nbytes, self._stream()))
        return buf

    def load_state_dict(self, blob: torch.Tensor):
        blob = blob.detach().to("cpu", torch.uint8).contiguous()
        _lib.check(self._L.mdx_checkpoint_load(self._h, C.c_void_p(blob.data_ptr()), int(blob.numel()), self._stream()))
        self._version += 1

    def reward_shaping_idq(self, cur_robots, cur_t, actions, ops_as_int: bool = False):
        """IDQ reward_shaping with the stored state as 'previous' (IDQ/networks.py:228-349) -> f64 [n, R]."""
        n = self.n_envs
        cur_robots = cur_robots.to(self.device, torch.int64).contiguous()
        packed = (cur_robots[..., 0] | (cur_robots[..., 1] << 16)).to(torch.int32).contiguous()
        cur_t = cur_t.to(self.device, torch.int32).contiguous()
        actions, aflags = self._check_actions(actions, n)
        if aflags:
            raise ValueError("reward_shaping_idq takes uint8 action codes")
        out = torch.empty((n, self.n_robots), dtype=torch.float64, device=self.device)
        _lib.check(self._L.mdx_reward_shaping_idq(self._h, _ptr(packed), _ptr(cur_t), _ptr(actions), 1 if ops_as_int else 0,
                                                  _ptr(out), self._stream()))
        return out

    # ------------------------------------------------------------------ introspection
    def _rows_to_states(self, rows):
        """structured rows -> list of the per-env dicts EnvView reads (0-based rows / cols, int32 arrays).  All arithmetic is
        done once over the whole batch of rows; the per-env dicts hold views."""
        P, W = self.n_packages, self.cols_out
        n = len(rows)
        if n == 0:
            return []
        rob = rows["robots"]
        cell = (rob & 0xffff).astype(np.int32)
        rr, rc, rk = cell // W, cell % W, (rob >> 16).astype(np.int32)
        sched = rows["sched"][:, :, :P]          # uint16 views of the rows (cells < 16 384, times / deadlines < 65 536)
        pos = sched[:, :2]
        prow, pcol = pos // W, pos % W
        status = rows["status"][:, :P]
        t, ep, mp, tot = rows["t"].tolist(), rows["episode"].tolist(), rows["map_id"].tolist(), rows["total_reward"].tolist()
        return [{"t": t[i], "episode": ep[i], "map_id": mp[i], "total_reward": tot[i],
                 "robot_row": rr[i], "robot_col": rc[i], "robot_carry": rk[i],
                 "pkg_start_row": prow[i, 0], "pkg_start_col": pcol[i, 0], "pkg_target_row": prow[i, 1], "pkg_target_col": pcol[i, 1],
                 "pkg_start_time": sched[i, 2], "pkg_deadline": sched[i, 3], "pkg_status": status[i]} for i in range(n)]

    def get_rows(self, ids=None) -> np.ndarray:
        """Packed snapshot of the listed envs (all when None) as a structured numpy array (fields t, episode, map_id,
        total_reward, robots = cell | carry << 16, sched [4, Pp8] = start cell / target cell / start_time / deadline, status):
        ONE gather launch, ONE D2H copy, ONE synchronisation for any number of envs."""
        if ids is None:
            h, n = None, self.n_envs
        else:
            h = np.ascontiguousarray(np.asarray(ids, dtype=np.int32).reshape(-1))
            n = int(h.size)
        rows = np.empty(n, dtype=self._row_dtype)
        if n:
            _lib.check(self._L.mdx_get_states(self._h, None if h is None else h.ctypes.data, n, rows.ctypes.data, self._stream()))
        return rows

    def get_states(self, ids=None) -> list:
        return self._rows_to_states(self.get_rows(ids))

    def get_state(self, i: int) -> dict:
        if not 0 <= int(i) < self.n_envs:
            raise ValueError(f"env id {i} out of range")
        return self._rows_to_states(self.get_rows([int(i)]))[0]

    def featurize_host(self, ids=None, obs=True, vec=True, gstate=True, gvec=True):
        """featurize() for HOST callers -> numpy arrays (None for the ones not asked for): one launch + one synchronisation
        inside one C call (include/mdx.h: mdx_featurize_host); small results are written by the kernel straight into pinned
        host memory."""
        if ids is None:
            h, n = None, self.n_envs
        else:
            h = np.ascontiguousarray(np.asarray(ids, dtype=np.int32).reshape(-1))
            n = int(h.size)
        R, H, W = self.n_robots, self.rows_out, self.cols_out
        o = np.empty((n, R, 6, H, W), np.float32) if obs else None
        v = np.empty((n, R, self.vec_dim), np.float32) if vec else None
        g = np.empty((n, 4, H, W), np.float32) if gstate else None
        gv = np.empty((n, self.gvec_dim), np.float32) if gvec else None
        vp = lambda a: None if a is None else a.ctypes.data  # noqa: E731
        _lib.check(self._L.mdx_featurize_host(self._h, vp(h), n, vp(o), vp(v), vp(g), vp(gv), self._stream()))
        return o, v, g, gv

    def set_rows(self, ids, rows: np.ndarray):
        """Packed state rows (the structured dtype of get_rows) -> the listed envs; validated in C like set_state."""
        h = np.ascontiguousarray(np.asarray(ids, dtype=np.int32).reshape(-1))
        rows = np.ascontiguousarray(rows)
        if rows.dtype != self._row_dtype or rows.shape != (h.size,):
            raise ValueError("rows must be a structured array of get_rows()'s dtype, one row per id")
        _lib.check(self._L.mdx_set_states(self._h, h.ctypes.data, int(h.size), rows.ctypes.data, self._stream()))
        self._version += 1

    def time_steps_host(self, sample: int = 256) -> np.ndarray:
        """Current time step of (a strided sample of) the envs as a host array -- one small synchronous read."""
        ids = np.unique(np.linspace(0, self.n_envs - 1, min(sample, self.n_envs)).astype(np.int32))
        return self.get_rows(ids)["t"].copy()

    def step_host(self, codes: np.ndarray, ids=None, auto_reset: bool = False, want_rows: bool = True, want_shaped: bool = False):
        """Environment.step for HOST callers: uint8 action codes [n, R] in, (reward f64 [n], shaped f32 [n] or None, done
        bool [n], structured state rows or None) out -- one H2D copy, the step kernel, a packed snapshot, one copy back and
        one synchronisation, all inside one C call (include/mdx.h: mdx_step_host)."""
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        if ids is None:
            h, n = None, self.n_envs
        else:
            h = np.ascontiguousarray(np.asarray(ids, dtype=np.int32).reshape(-1))
            n = int(h.size)
        if codes.shape != (n, self.n_robots):
            raise ValueError("The number of actions must match the number of robots.")  # env.py:182-183
        bufs = self._host.get(n)
        if bufs is None:
            bufs = self._host[n] = (np.empty(n, np.float64), np.empty(n, np.float32), np.empty(n, np.uint8), np.empty(n, dtype=self._row_dtype))
        reward, shaped, done, rows = bufs[0], (bufs[1] if want_shaped else None), bufs[2], (bufs[3] if want_rows else None)
        vp = lambda a: None if a is None else a.ctypes.data  # noqa: E731
        rc = self._L.mdx_step_host(self._h, codes.ctypes.data, vp(h), n, reward.ctypes.data, vp(shaped), done.ctypes.data, vp(rows),
                                   MDX_F_AUTO_RESET if auto_reset else 0, self._stream())
        if rc:
            _lib.check(rc)
        reward, done = reward.copy(), done.astype(bool)
        shaped = None if shaped is None else shaped.copy()
        rows = None if rows is None else rows.copy()
        self._version += 1
        return reward, shaped, done, rows

    # ------------------------------------------------------------------ the MAPPO trainer's tracker (stale_tracker=True)
    def tracker_clear(self, ids=None):
        """Empty the package dict of the listed envs (all when None): what QMIX/trainer.py:523-526 does on reset."""
        if ids is None:
            _lib.check(self._L.mdx_tracker_clear(self._h, None, 0, self._stream()))
        else:
            h = np.ascontiguousarray(np.asarray(ids, dtype=np.int32).reshape(-1))
            _lib.check(self._L.mdx_tracker_clear(self._h, h.ctypes.data, int(h.size), self._stream()))

    def get_tracker(self, i: int) -> list:
        """The dict of env i in dict (insertion) order: [(package id, start cell, target cell, start_time, deadline, status)]."""
        P = self.n_packages
        Pp8, Pp16 = max(8, (P + 7) // 8 * 8), max(16, (P + 15) // 16 * 16)
        n = C.c_int32(0)
        rows = np.zeros((5, Pp8), np.uint16)
        st = np.zeros(Pp16, np.uint8)
        _lib.check(self._L.mdx_tracker_get(self._h, int(i), C.byref(n), rows.ctypes.data, st.ctypes.data, self._stream()))
        return [(int(rows[0, k]), int(rows[1, k]), int(rows[2, k]), int(rows[3, k]), int(rows[4, k]), int(st[k])) for k in range(n.value)]

    def set_tracker(self, i: int, entries):
        """entries: [(package id, start cell, target cell, start_time, deadline, status 1|2)] in dict order."""
        P = self.n_packages
        Pp8, Pp16 = max(8, (P + 7) // 8 * 8), max(16, (P + 15) // 16 * 16)
        if len(entries) > P:
            raise ValueError("a tracker holds at most n_packages entries")
        rows = np.zeros((5, Pp8), np.uint16)
        st = np.zeros(Pp16, np.uint8)
        for k, e in enumerate(entries):
            rows[:, k] = e[:5]
            st[k] = e[5]
        _lib.check(self._L.mdx_tracker_set(self._h, int(i), len(entries), rows.ctypes.data, st.ctypes.data, self._stream()))

    def set_state(self, i: int, t, total_reward, robots, packages):
        """robots: [(row, col, carry)] 0-based; packages: [(start_row, start_col, target_row, target_col,
        start_time, deadline, status_code)] in id order, exactly n_packages entries."""
        R, P = self.n_robots, self.n_packages
        if len(robots) != R or len(packages) != P:
            raise ValueError("robots / packages must match n_robots / n_packages")
        rb = np.ascontiguousarray(np.asarray(robots, dtype=np.int32).reshape(R, 3).T)
        pk = np.ascontiguousarray(np.asarray(packages, dtype=np.int32).reshape(P, 7).T) if P else np.zeros((7, 1), np.int32)
        v = _lib.StateView()
        v.t, v.total_reward = int(t), float(total_reward)
        p32 = C.POINTER(C.c_int32)
        v.robot_row, v.robot_col, v.robot_carry = (rb[k].ctypes.data_as(p32) for k in range(3))
        (v.pkg_start_row, v.pkg_start_col, v.pkg_target_row, v.pkg_target_col, v.pkg_start_time, v.pkg_deadline,
         v.pkg_status) = (pk[k].ctypes.data_as(p32) for k in range(7))
        _lib.check(self._L.mdx_set_state(self._h, int(i), C.byref(v), self._stream()))
        self._version += 1

    def episode_stats(self, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """int64[8] (STAT_NAMES) summed over this instance's envs; all-reduce it across ranks.  `out`: a caller-owned
        int64[8] CUDA tensor to write into (steady-state loops: no allocation)."""
        if out is None:
            out = torch.empty(8, dtype=torch.int64, device=self.device)
        else:
            self._check_out((out,), (("stats", 8, torch.int64),))
        _lib.check(self._L.mdx_episode_stats(self._h, _ptr(out), self._stream()))
        return out


def compute_gae(rewards: torch.Tensor, values: torch.Tensor, dones: torch.Tensor, last_value: torch.Tensor,
                gamma: float = 0.99, gae_lambda: float = 0.95):
    """GAE + returns of MAPPO/trainer.py:288-301 on a device-resident [T, N] rollout (float32, the reference's order
    of operations) -> (advantages, returns).  One CUDA launch instead of the reference's T small torch ops."""
    if rewards.device.type != "cuda":
        raise _lib.MdxError("compute_gae runs on the GPU only (no CPU fallback)")
    T, N = rewards.shape
    rewards = rewards.to(torch.float32).contiguous()
    values = values.to(device=rewards.device, dtype=torch.float32).contiguous()
    dones = dones.to(device=rewards.device, dtype=torch.uint8).contiguous()
    last_value = last_value.to(device=rewards.device, dtype=torch.float32).reshape(N).contiguous()
    if tuple(values.shape) != (T, N) or tuple(dones.shape) != (T, N):
        raise ValueError("rewards, values and dones must share the shape [T, N]")
    adv = torch.empty_like(rewards)
    ret = torch.empty_like(rewards)
    with torch.cuda.device(rewards.device):
        st = C.c_void_p(torch.cuda.current_stream(rewards.device).cuda_stream)
        _lib.check(_lib.lib().mdx_gae(_ptr(rewards), _ptr(values), _ptr(dones), _ptr(last_value), int(T), int(N), float(gamma),
                                      float(gae_lambda), _ptr(adv), _ptr(ret), st))
    return adv, ret
