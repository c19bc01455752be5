"""CPU: the environment as look-up tables (navgraph_env.DeviceNavTables) against the environment itself —
observe() and step() of every discretised state of a real R2R navigation graph (tests/golden/nav_graphs.npz)."""
import math

import numpy as np
import pytest

from speaker_follower_b200.navgraph_env import DeviceNavTables, FakeR2RBatch, WorldState


@pytest.mark.parametrize("graph", ["8194nk5LbLH", "pLe4wQe7qrG"])
def test_tables_reproduce_observe_and_step_of_every_state(graph):
    env = FakeR2RBatch(n_instr=4, batch_size=4, seed=3, graph=graph, with_features=False)
    env.reset()
    nav = DeviceNavTables(env, "cpu", with_teacher=True)
    nvp = len(env.adj)
    assert nav.S == nvp * 12
    item = env.batch[0]
    for s in range(0, nav.S, 7):                                  # a stride keeps the test fast; every viewpoint is visited
        ws = nav.world_state(s)
        assert nav.state_ids([ws]) == [s]
        ob = env._observe_one(ws, item, include_teacher=False)
        adj = ob["adj_loc_list"]
        assert int(nav.vp[s]) == ob["vp_index"] and int(nav.view[s]) == ob["viewIndex"] and int(nav.nvalid[s]) == len(adj)
        for a, d in enumerate(adj):
            nxt = env._step_one(ws, a, ob)
            assert nav.state_ids([nxt]) == [int(nav.next[s, a])], (s, a)
            if a > 0:
                assert int(nav.cv[s, a]) == d["absViewIndex"]
                want = np.float32([math.sin(d["rel_heading"]), math.cos(d["rel_heading"]),
                                   math.sin(d["rel_elevation"]), math.cos(d["rel_elevation"])])
                assert np.array_equal(nav.trig[s, a].numpy(), want)
        assert (nav.cv[s, len(adj):] == -1).all() and int(nav.cv[s, 0]) == -1
        # the teacher table: next hop on a shortest path to every goal, 0 at the goal (env.py:742-761)
        for goal in range(0, nvp, 11):
            ob_t = env._observe_one(ws, dict(item, goal=goal), include_teacher=True)
            assert int(nav.teach[s, goal]) == ob_t["teacher"], (s, goal)


def test_observe_states_equals_env_observe_without_teacher():
    env = FakeR2RBatch(n_instr=3, batch_size=3, seed=5, graph="8194nk5LbLH", with_features=False)
    env.reset()
    nav = DeviceNavTables(env, "cpu", with_teacher=False)
    states = [[5, 17, 5], [40], [3, 4]]
    got = nav.observe_states(env, states)
    want = env.observe([[nav.world_state(s) for s in row] for row in states], beamed=True, include_teacher=False)
    for g_row, w_row in zip(got, want):
        for g, w in zip(g_row, w_row):
            assert g.keys() == w.keys()
            for k in ("instr_id", "viewpoint", "viewIndex", "heading", "elevation", "vp_index", "instr_length"):
                assert g[k] == w[k], k
            assert np.array_equal(g["instr_encoding"], w["instr_encoding"])
            assert g["adj_loc_list"] == w["adj_loc_list"]
