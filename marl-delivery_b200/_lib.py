"""ctypes binding of libmdx.so (include/mdx.h).  There is NO CPU fallback: if the CUDA library is
missing, or there is no CUDA device when an environment is created, this raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MDX_LIB_PATH") or os.path.join(_HERE, "csrc", "libmdx.so")  # override: A/B builds of the same library

MDX_OK, MDX_E_INVALID, MDX_E_CUDA, MDX_E_STARVED, MDX_E_NOMEM = 0, -1, -2, -3, -4
MDX_F_AUTO_RESET = 1
MDX_F_ACTIONS_I64 = 2
MDX_CFG_BLANK = 1
MDX_CFG_BFS_SHAPING = 2
MDX_CFG_LAX_ROBOTS = 4
MDX_CFG_STALE_TRACKER = 8
MDX_CFG_JIT_SHAPE = 16
MDX_MAX_ROBOTS = 32


class MdxError(RuntimeError):
    pass


class Shaping(C.Structure):
    _fields_ = [(n, C.c_float) for n in (
        "pickup", "delivery_on_time", "delivery_late", "moved_closer", "wasted_pickup", "wasted_drop",
        "collision_or_stuck", "idle_with_tasks", "moved_away")]


class Config(C.Structure):
    _fields_ = [
        ("n_envs", C.c_int32), ("n_robots", C.c_int32), ("n_packages", C.c_int32), ("max_time_steps", C.c_int32),
        ("move_cost", C.c_double), ("delivery_reward", C.c_double), ("delay_reward", C.c_double),
        ("shaping", Shaping),
        ("k_others", C.c_int32), ("k_packages", C.c_int32),
        ("max_robots_in_state", C.c_int32), ("max_packages_in_state", C.c_int32),
        ("n_maps", C.c_int32),
        ("map_rows", C.POINTER(C.c_int32)), ("map_cols", C.POINTER(C.c_int32)),
        ("map_cells", C.POINTER(C.POINTER(C.c_uint8))),
        ("env_map", C.POINTER(C.c_int32)), ("env_seed", C.POINTER(C.c_int64)),
        ("base_seed", C.c_int64),
        ("device", C.c_int32), ("ring_episodes", C.c_int32), ("sampler_threads", C.c_int32), ("flags", C.c_int32),
    ]


class Dims(C.Structure):
    _fields_ = [
        ("n_envs", C.c_int32), ("n_robots", C.c_int32), ("n_packages", C.c_int32), ("max_time_steps", C.c_int32),
        ("rows_out", C.c_int32), ("cols_out", C.c_int32), ("vec_dim", C.c_int32), ("gvec_dim", C.c_int32),
        ("bytes_per_env_step", C.c_int64), ("ring_episodes", C.c_int32), ("reserved", C.c_int32),
    ]


_I32P = C.POINTER(C.c_int32)


class StateView(C.Structure):
    _fields_ = [
        ("t", C.c_int32), ("episode", C.c_int32), ("map_id", C.c_int32), ("reserved", C.c_int32),
        ("total_reward", C.c_double),
        ("robot_row", _I32P), ("robot_col", _I32P), ("robot_carry", _I32P),
        ("pkg_start_row", _I32P), ("pkg_start_col", _I32P), ("pkg_target_row", _I32P), ("pkg_target_col", _I32P),
        ("pkg_start_time", _I32P), ("pkg_deadline", _I32P), ("pkg_status", _I32P),
    ]


# every symbol include/mdx.h declares: name -> (restype, argtypes)
_VP = C.c_void_p
SYMBOLS = {
    "mdx_last_error": (C.c_char_p, []),
    "mdx_version": (C.c_int32, []),
    "mdx_create": (C.c_int32, [C.POINTER(Config), C.POINTER(_VP)]),
    "mdx_destroy": (C.c_int32, [_VP]),
    "mdx_get_dims": (C.c_int32, [_VP, C.POINTER(Dims)]),
    "mdx_jit_status": (C.c_int32, [_VP, C.c_char_p, C.c_int32]),
    "mdx_reset": (C.c_int32, [_VP, _I32P, C.c_int32, _VP]),
    "mdx_prefetch": (C.c_int32, [_VP, C.c_int32, _VP]),
    "mdx_step": (C.c_int32, [_VP, _VP, _VP, C.c_int32, _VP, _VP, _VP, C.c_uint32, _VP]),
    "mdx_featurize": (C.c_int32, [_VP, _VP, C.c_int32, _VP, _VP, _VP, _VP, _VP]),
    "mdx_step_featurize": (C.c_int32, [_VP, _VP, _VP, C.c_int32, _VP, _VP, _VP, _VP, _VP, _VP, _VP, C.c_uint32, _VP]),
    "mdx_shaped_reward": (C.c_int32, [_VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "mdx_get_state": (C.c_int32, [_VP, C.c_int32, C.POINTER(StateView), _VP]),
    "mdx_set_state": (C.c_int32, [_VP, C.c_int32, C.POINTER(StateView), _VP]),
    "mdx_episode_stats": (C.c_int32, [_VP, _VP, _VP]),
    "mdx_shaped_reward_host": (C.c_int32, [_VP, _VP, _VP, _VP, _VP, _VP, _VP]),
    "mdx_featurize_host": (C.c_int32, [_VP, _VP, C.c_int32, _VP, _VP, _VP, _VP, _VP]),
    "mdx_tracker_clear": (C.c_int32, [_VP, _VP, C.c_int32, _VP]),
    "mdx_tracker_get": (C.c_int32, [_VP, C.c_int32, _VP, _VP, _VP, _VP]),
    "mdx_tracker_set": (C.c_int32, [_VP, C.c_int32, C.c_int32, _VP, _VP, _VP]),
    "mdx_step_range": (C.c_int32, [_VP, _VP, C.c_int32, C.c_int32, _VP, _VP, _VP, C.c_uint32, _VP]),
    "mdx_step_featurize_range": (C.c_int32, [_VP, _VP, C.c_int32, C.c_int32, _VP, _VP, _VP, _VP, _VP, _VP, _VP, C.c_uint32, _VP]),
    "mdx_step_n": (C.c_int32, [_VP, _VP, C.c_int32, _VP, _VP, _VP, C.c_uint32, _VP]),
    "mdx_state_row_bytes": (C.c_int64, [_VP]),
    "mdx_get_states": (C.c_int32, [_VP, _VP, C.c_int32, _VP, _VP]),
    "mdx_set_states": (C.c_int32, [_VP, _VP, C.c_int32, _VP, _VP]),
    "mdx_step_host": (C.c_int32, [_VP, _VP, _VP, C.c_int32, _VP, _VP, _VP, _VP, C.c_uint32, _VP]),
    "mdx_greedy_actions": (C.c_int32, [_VP, _VP, C.c_int32, _VP, _VP]),
    "mdx_bfs_table": (C.c_int32, [_VP, C.c_int32, C.POINTER(C.c_uint16)]),
    "mdx_featurize_alt": (C.c_int32, [_VP, _VP, C.c_int32, _VP, _VP, _VP]),
    "mdx_step_idq": (C.c_int32, [_VP, _VP, _VP, C.c_int32, _VP, _VP, _VP, C.c_int32, C.c_uint32, _VP]),
    "mdx_step_idq_featurize": (C.c_int32, [_VP, _VP, _VP, C.c_int32, _VP, _VP, _VP, _VP, _VP, C.c_int32, C.c_uint32, _VP]),
    "mdx_step_idq_featurize_range": (C.c_int32, [_VP, _VP, C.c_int32, C.c_int32, _VP, _VP, _VP, _VP, _VP, C.c_int32, C.c_uint32, _VP]),
    "mdx_reward_shaping_idq": (C.c_int32, [_VP, _VP, _VP, _VP, C.c_int32, _VP, _VP]),
    "mdx_gae": (C.c_int32, [_VP, _VP, _VP, _VP, C.c_int32, C.c_int32, C.c_double, C.c_double, _VP, _VP, _VP]),
    "mdx_checkpoint_bytes": (C.c_int64, [_VP]),
    "mdx_checkpoint_save": (C.c_int32, [_VP, _VP, C.c_int64, _VP]),
    "mdx_checkpoint_load": (C.c_int32, [_VP, _VP, C.c_int64, _VP]),
    "mdx_sampler_create": (C.c_int32, [C.c_uint32, C.POINTER(C.c_uint8), C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                       C.c_int32, C.POINTER(_VP)]),
    "mdx_sampler_next": (C.c_int32, [_VP, _I32P, _I32P]),
    "mdx_sampler_destroy": (C.c_int32, [_VP]),
    "mdx_rng_randint": (C.c_int32, [C.c_uint32, C.c_int64, C.c_int64, C.c_int32, C.POINTER(C.c_int64)]),
    "mdx_lut_quotient": (C.c_int32, [C.c_int32, C.c_int32, C.POINTER(C.c_float)]),
    "mdx_lut_rank": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_uint16), _I32P]),
    "mdx_np_sum_f32": (C.c_float, [C.POINTER(C.c_float), C.c_int32]),
}

_lib = None


def lib():
    """Load libmdx.so (built in-tree by __graft_entry__.build() / csrc/build.sh)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MdxError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  marl-delivery_b200 has no CPU fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int):
    if rc == MDX_OK:
        return
    msg = lib().mdx_last_error().decode("utf-8", "replace")
    if rc == MDX_E_INVALID:
        raise ValueError(msg)
    raise MdxError(f"libmdx error {rc}: {msg}")
