"""Drop-in for the reference's `greedyagent.GreedyAgents` (greedyagent.py:42-174): same class, same
`init_agents(state)` / `get_actions(state)` calls, same `(move_str, pkg_str)` tuples -- computed by
`mdx_greedy_actions` on the device-resident environment (per-map all-pairs BFS table instead of one
breadth-first search per call, greedyagent.py:6-41).

`state` must be a state dict returned by this package's Environment / VectorizedEnv (it carries a
back-reference to its environment slot); the agent's bookkeeping lives on the device next to the env
and restarts whenever the env is at t = 0, which is what evaluation.py:9-66 does by building a new
GreedyAgents per episode.  Batched use: `DeliveryBatch.greedy_actions()` returns the action codes of
all envs as one uint8 tensor that `step()` takes directly.
"""
from __future__ import annotations

_MOVES = "DLRSU"


def _view_of(state):
    v = getattr(state, "_view", None)
    if v is None:
        raise TypeError("GreedyAgents needs a state dict returned by marl_delivery_b200's Environment / VectorizedEnv "
                        "(the policy runs on the device-resident environment; there is no CPU fallback)")
    return v


class GreedyAgents:
    def __init__(self):
        self.n_robots = 0
        self.state = None
        self.is_init = False

    def init_agents(self, state):
        v = _view_of(state)
        self.state = state
        self.n_robots = len(state["robots"])
        self.map = state["map"]
        if v.t != 0:
            raise ValueError("init_agents expects the state of a freshly reset environment (t = 0)")

    def get_actions(self, state, deterministic=True):
        v = _view_of(state)
        self.is_init = True
        self.state = state
        codes = v._b.greedy_actions(ids=v._ids).cpu().tolist()[0]
        return [(_MOVES[a % 5], str(a // 5)) for a in codes]
