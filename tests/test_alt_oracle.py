"""Pins the f3 oracle (IDQ / legacy-qmix featurisers, IDQ reward_shaping) and the QMIX-sized MAPPO
featurisers to values recorded from the unmodified reference (tests/golden/kats_alt.npz)."""
import json
import os

import numpy as np

from golden_util import GOLDEN
from oracle import alt_oracle as alt
from oracle import oracle_env as orc


def replay_alt_case(z, m, visit):
    """Drive an OracleEnv along the recorded actions (auto-reset on done) and call
    visit(k, env, snapshot, shaping) for the state the features of row k were taken from; `shaping` is
    the oracle's (string-op, int-op) IDQ reward_shaping of the transition that led there, evaluated before
    any auto-reset (None for the initial state)."""
    e = orc.OracleEnv(z[f"{m['name']}_grid"], m["T"], m["R"], m["P"], seed=m["seed"])
    e.reset()
    visit(0, e, e.snapshot(), None)
    for s in range(m["n_steps"]):
        a = z[f"{m['name']}_actions"][s]
        before = e.snapshot()
        r, d = e.step(a)
        after = e.snapshot()
        assert d == bool(z[f"{m['name']}_done"][s])
        shaping = (alt.idq_reward_shaping(e, before, after, a, False), alt.idq_reward_shaping(e, before, after, a, True))
        if d:
            e.reset()
        visit(s + 1, e, e.snapshot(), shaping)


def test_alt_oracle_matches_reference():
    z = np.load(os.path.join(GOLDEN, "kats_alt.npz"))
    for m in json.loads(str(z["meta"])):
        n = m["name"]
        R = m["R"]

        def visit(k, e, snap, step):
            for i in range(R):
                assert np.array_equal(alt.idq_observation(e, snap, i).view(np.uint32), z[f"{n}_obs6"][k, i].view(np.uint32)), (n, k, i)
                got = orc.vector_features(e, snap, i, R - 1, 5)
                assert np.array_equal(got.view(np.uint32), z[f"{n}_qvec"][k, i].view(np.uint32)), (n, k, i)
            assert np.array_equal(alt.legacy_global_tensor(e, snap).view(np.uint32), z[f"{n}_g7"][k].view(np.uint32)), (n, k)
            assert np.array_equal(orc.global_state(e, snap, 10, 20)[1].view(np.uint32), z[f"{n}_qgvec"][k].view(np.uint32)), (n, k)
            if step is not None:
                assert step[0] == list(z[f"{n}_shape_str"][k - 1]), (n, k)
                assert step[1] == list(z[f"{n}_shape_int"][k - 1]), (n, k)

        replay_alt_case(z, m, visit)
