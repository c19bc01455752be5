"""The search oracle (oracle/search_oracle.py) checked against the step oracle and the reference's own invariants on
REAL R2R navigation graphs (tests/golden/nav_graphs.npz), CPU only: beam(1) == greedy rollout (follower.py:147-180),
per-step scores add up (rational_speaker.py:87-89), state-factored candidates end in distinct world states."""
import numpy as np
import torch

from fake_env import FakeR2RBatch
from oracle import r2r_oracle as O
from oracle import search_oracle as S
from speaker_follower_b200 import synth


def greedy_oracle(env, we, wd, episode_len, max_length):
    """follower.py:430-539 (argmax feedback) with the step oracle, on the same env protocol."""
    ws = env.reset(sort=True)
    obs = env.observe(ws)
    seq, mask, lengths = S.batch_instructions([o["instr_encoding"] for o in obs], max_length)
    ctx, h, c = O.encoder_lstm(seq[:, :max(lengths)], lengths, we)
    n = len(obs)
    u = torch.zeros(n, synth.FEAT)
    ended = np.zeros(n, bool)
    traj = [{"instr_id": o["instr_id"], "trajectory": [(o["viewpoint"], o["heading"], o["elevation"])], "actions": [], "score": 0.0}
            for o in obs]
    for t in range(episode_len):
        U, valid, _ = S.action_variable(obs)
        h, c, alpha, logit, _ = O.attn_decoder_step(u, U, S.feature_variable(obs), h, c, ctx, mask, wd)
        logit = logit.masked_fill(valid == 0, -float("inf"))
        a = logit.max(1)[1]
        lp = torch.log_softmax(logit, 1)
        u = U[torch.arange(n), a]
        ws = env.step(ws, a.tolist(), obs)
        obs = env.observe(ws)
        for i in range(n):
            if not ended[i]:
                traj[i]["actions"].append(int(a[i]))
                traj[i]["score"] += float(lp[i, a[i]])
                traj[i]["trajectory"].append((obs[i]["viewpoint"], obs[i]["heading"], obs[i]["elevation"]))
                if int(a[i]) == 0:
                    ended[i] = True
        if ended.all():
            break
    return traj


def test_oracle_beam1_is_greedy_on_a_real_graph():
    we, wd = synth.follower_encoder_weights(), synth.follower_decoder_weights()
    torch.set_num_threads(4)
    with torch.no_grad():
        g = greedy_oracle(FakeR2RBatch(n_instr=4, batch_size=4, seed=11, graph="8194nk5LbLH", beam_size=3), we, wd, 5, 20)
        b1, _ = S.follower_beam_search(FakeR2RBatch(n_instr=4, batch_size=4, seed=11, graph="8194nk5LbLH", beam_size=3),
                                       we, wd, 1, episode_len=5, max_length=20)
        b3, _ = S.follower_beam_search(FakeR2RBatch(n_instr=4, batch_size=4, seed=11, graph="8194nk5LbLH", beam_size=3),
                                       we, wd, 3, episode_len=5, max_length=20)
    for gg, one, three in zip(g, b1, b3):
        assert one[0]["instr_id"] == gg["instr_id"] and one[0]["actions"] == gg["actions"]
        assert one[0]["trajectory"] == gg["trajectory"]
        assert abs(one[0]["score"] - gg["score"]) < 1e-5
        sc = [x["score"] for x in three]
        assert sc == sorted(sc, reverse=True) and sc[0] >= one[0]["score"] - 1e-6
        for x in three:
            assert abs(sum(x["scores"]) - x["score"]) < 1e-5


def test_oracle_state_factored_search_on_a_real_graph():
    we, wd = synth.follower_encoder_weights(), synth.follower_decoder_weights()
    env = FakeR2RBatch(n_instr=3, batch_size=3, seed=12, graph="GdvgFV5R1Z5", beam_size=2)
    torch.set_num_threads(4)
    with torch.no_grad():
        trajs, completed, traversed = S.follower_state_factored_search(env, we, wd, 3, 1, episode_len=5, max_length=20)
    for cands, states, walk in zip(trajs, completed, traversed):
        assert 1 <= len(cands) <= 3
        keys = [tuple(s.world_state[0:4]) for s in states]
        assert len(set(keys)) == len(keys)
        sc = [c["score"] for c in cands]
        assert sc == sorted(sc, reverse=True)
        for c in cands:
            assert abs(sum(c["scores"]) - c["score"]) < 1e-4
            assert c["actions"][-1] == 0 or len(c["actions"]) == 5
        for a, b in zip(walk[:-1], walk[1:]):
            va, vb = a.world_state.viewpointId, b.world_state.viewpointId
            assert va == vb or vb in env.adj[va]
