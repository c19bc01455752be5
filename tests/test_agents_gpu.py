"""Agent-level parity on a fake environment (tests/fake_env.py): the CUDA agents (speaker_follower_b200.follower /
.speaker) against the CPU oracle fed with the observations the rollout actually saw, plus the invariants the
reference states in comments (follower.py:147-180): beam_search(1) == greedy rollout, teacher rollout ==
_score_obs_actions_and_instructions."""
import numpy as np
import pytest
import torch

from fake_env import FakeR2RBatch
from oracle import r2r_oracle as O
from speaker_follower_b200 import follower as Fo, model as M, ops, speaker as Sp, synth

pytestmark = pytest.mark.gpu


def make_follower(env, store=False, seed_shift=0):
    glove = synth.follower_encoder_weights()["embedding.weight"].numpy()
    enc = M.EncoderLSTM(synth.VOCAB, synth.WORD, synth.HID, 0, 0.5, glove=glove).cuda().eval()
    dec = M.AttnDecoderLSTM(synth.FEAT, synth.HID, 0.5).cuda().eval()
    we, wd = synth.follower_encoder_weights(), synth.follower_decoder_weights()
    enc.load_state_dict(we); dec.load_state_dict(wd)
    if store:
        dec.feature_store = ops.FeatureStore(torch.from_numpy(env.table).cuda(), torch.from_numpy(env.loc).cuda())
    return Fo.Seq2SeqAgent(env, "", enc, dec, episode_len=6, max_instruction_length=20), we, wd


def oracle_check(agent, traj, we, wd, feedback):
    """Re-run the recorded observation sequence through the CPU oracle and compare actions / scores."""
    B = len(traj)
    enc = [t["instr_encoding"] for t in traj]
    seq, mask, lengths = Fo.batch_instructions_from_encoded(enc, agent.max_instruction_length, reverse=True)
    T = max(len(t["actions"]) for t in traj)
    steps = []
    for s in range(T):
        obs = [t["observations"][min(s, len(t["observations"]) - 2)] for t in traj]
        A = max(len(o["adj_loc_list"]) for o in obs)
        U = torch.zeros(B, A, synth.FEAT); valid = torch.zeros(B, A)
        for i, o in enumerate(obs):
            n = len(o["adj_loc_list"]); U[i, :n] = torch.from_numpy(o["action_embedding"]); valid[i, :n] = 1
        vis = torch.from_numpy(np.stack([o["feature"][0] for o in obs]))
        tgt = torch.tensor([(t["actions"][s] if s < len(t["actions"]) else -1) for t in traj])
        steps.append({"visual": vis, "all_u_t": U, "is_valid": valid, "target": tgt if feedback == "teacher" else None})
    res, loss, score = O.follower_rollout(seq, mask, lengths, steps, we, wd, feedback=feedback)
    for i, t in enumerate(traj):
        for s, a in enumerate(t["actions"]):
            assert int(res[s]["a_t"][i]) == int(a), (i, s)
            assert abs(float(res[s]["scores"][i]) - t["scores"][s]) < 1e-4
    return res, loss, score


@pytest.mark.parametrize("store", [False, True])
def test_greedy_rollout_matches_oracle(store):
    env = FakeR2RBatch(n_viewpoints=20, n_instr=8, batch_size=8, seed=3)
    agent, we, wd = make_follower(env, store=store)
    agent.feedback = "argmax"
    with torch.no_grad():
        traj = agent.rollout()
    assert len(traj) == 8 and all(len(t["trajectory"]) == len(t["actions"]) + 1 for t in traj)
    # rows that ended keep stepping in the reference (follower.py:509-513) but stop recording
    oracle_check(agent, traj, we, wd, "argmax")


def test_beam1_equals_greedy_and_beams_are_sorted():
    env = FakeR2RBatch(n_viewpoints=20, n_instr=8, batch_size=8, seed=4, beam_size=4)
    agent, we, wd = make_follower(env)
    agent.feedback = "argmax"
    with torch.no_grad():
        greedy = agent._rollout_with_loss()
        beams1, _, _ = agent.beam_search(1, load_next_minibatch=False)
        beams4, _, _ = agent.beam_search(4, load_next_minibatch=False)
    for g, b in zip(greedy, beams1):
        assert g["instr_id"] == b[0]["instr_id"]
        assert g["trajectory"] == b[0]["trajectory"]
        assert abs(g["score"] - b[0]["score"]) < 1e-4
    for g, bs in zip(greedy, beams4):
        sc = [b["score"] for b in bs]
        assert sc == sorted(sc, reverse=True) and 1 <= len(bs) <= 4
        assert abs(sum(bs[0]["scores"]) - bs[0]["score"]) < 1e-4          # rational_speaker.py:87-89 invariant
    # teacher-forced rescoring of a returned beam reproduces its score (a8)
    obs_paths = [bs[0]["observations"] for bs in beams4]
    act_paths = [bs[0]["actions"] for bs in beams4]
    encs = [bs[0]["instr_encoding"] for bs in beams4]
    with torch.no_grad():
        scored, _ = agent._score_obs_actions_and_instructions(obs_paths, act_paths, encs)
    for s, bs in zip(scored, beams4):
        assert abs(s["score"] - bs[0]["score"]) < 2e-4 and [int(a) for a in s["actions"]] == [int(a) for a in bs[0]["actions"]]


def test_teacher_rollout_loss_and_sample_feedback():
    env = FakeR2RBatch(n_viewpoints=16, n_instr=8, batch_size=8, seed=5)
    agent, we, wd = make_follower(env)
    agent.feedback = "teacher"
    with torch.no_grad():
        traj = agent.rollout()
    res, loss, score = oracle_check(agent, traj, we, wd, "teacher")
    torch.manual_seed(0)
    agent.feedback = "sample"
    with torch.no_grad():
        traj = agent.rollout()
    for t in traj:                      # sampled actions are always valid actions
        for o, a in zip(t["observations"], t["actions"]):
            assert 0 <= a < len(o["adj_loc_list"])


def test_speaker_teacher_scoring_matches_oracle():
    env = FakeR2RBatch(n_viewpoints=16, n_instr=6, batch_size=6, seed=6)
    we, wd = synth.speaker_encoder_weights(), synth.speaker_decoder_weights()
    enc = M.SpeakerEncoderLSTM(synth.FEAT, synth.FEAT, synth.HID, 0.5).cuda().eval()
    dec = M.SpeakerDecoderLSTM(synth.VOCAB, synth.WORD, synth.HID, 0.5, glove=wd["embedding.weight"].numpy()).cuda().eval()
    enc.load_state_dict(we); dec.load_state_dict(wd)
    spk = Sp.Seq2SeqSpeaker(env, "", enc, dec, instruction_len=12, max_episode_len=6)
    path_obs, path_actions, encoded = env.gold_obs_actions_and_instructions(6)
    with torch.no_grad():
        outs, loss = spk._score_obs_actions_and_instructions(path_obs, path_actions, encoded, "teacher")
    # oracle on the same batch
    _, feats, acts, mask, _, _, _ = spk._batch_observations_and_actions(path_obs, path_actions, encoded)
    instr, _, _ = Fo.batch_instructions_from_encoded(encoded, 12)
    sc, l, words, wsc = O.speaker_score_teacher([a.cpu() for a in acts], [f.cpu() for f in feats], mask.cpu().bool(), instr, we, wd)
    for i, o in enumerate(outs):
        assert abs(o["score"] - float(sc[i])) < 2e-3 * max(1.0, abs(float(sc[i])))
        n = len(o["word_indices"])
        assert o["word_indices"] == [int(x) for x in words[i, :n]]
    assert abs(float(loss) - float(l)) < 1e-3 * max(1.0, abs(float(l)))
    spk.feedback = "argmax"
    with torch.no_grad():
        res = spk.test()
    assert len(res) == 6


def test_speaker_scoring_from_feature_store_equals_dense_batching():
    """speaker.py:68-121 without host slabs: with `encoder.feature_store` set the batch carries (viewpoint row, view index)
    pairs and the action embeddings are assembled on the device; ragged path lengths (padded steps must see zero input).
    Scores and words equal the dense-batching path (same kernels, gathered vs. dense source) and the CPU oracle."""
    env = FakeR2RBatch(n_viewpoints=24, n_instr=10, batch_size=10, seed=16)
    we, wd = synth.speaker_encoder_weights(), synth.speaker_decoder_weights()
    enc = M.SpeakerEncoderLSTM(synth.FEAT, synth.FEAT, synth.HID, 0.5).cuda().eval()
    dec = M.SpeakerDecoderLSTM(synth.VOCAB, synth.WORD, synth.HID, 0.5, glove=wd["embedding.weight"].numpy()).cuda().eval()
    enc.load_state_dict(we); dec.load_state_dict(wd)
    spk = Sp.Seq2SeqSpeaker(env, "", enc, dec, instruction_len=12, max_episode_len=6)
    path_obs, path_actions, encoded = env.gold_obs_actions_and_instructions(6)
    assert len(set(len(a) for a in path_actions)) > 1, "the batch must be ragged for this test to mean something"
    with torch.no_grad():
        dense, loss_d = spk._score_obs_actions_and_instructions(path_obs, path_actions, encoded, "teacher")
        _, feats_d, acts_d, _, _, _, _ = spk._batch_observations_and_actions(path_obs, path_actions, encoded)
        enc.feature_store = ops.FeatureStore(torch.from_numpy(env.table).cuda(), torch.from_numpy(env.loc).cuda())
        _, feats_i, acts_i, _, _, _, _ = spk._batch_observations_and_actions(path_obs, path_actions, encoded)
        assert isinstance(feats_i[0], tuple)
        for a_d, a_i in zip(acts_d, acts_i):                      # device-assembled action embeddings == env.py:60-75 on the host
            assert torch.equal(a_d, a_i)
        spk._step_masks = None
        idx, loss_i = spk._score_obs_actions_and_instructions(path_obs, path_actions, encoded, "teacher")
    for d, i in zip(dense, idx):
        assert d["word_indices"] == i["word_indices"]
        assert abs(d["score"] - i["score"]) < 1e-4 * max(1.0, abs(d["score"]))
    assert abs(float(loss_d) - float(loss_i)) < 1e-5 * max(1.0, abs(float(loss_d)))


def test_state_factored_search_invariants():
    """follower.py:720-980 on the fake env: distinct end states per instance, sorted by score, per-step scores add up
    (rational_speaker.py:87-89), teacher-forced rescoring of every candidate reproduces its score (the invariant
    rational_follower.py relies on), and the traversal list is a connected physical walk from the start."""
    env = FakeR2RBatch(n_viewpoints=20, n_instr=8, batch_size=8, seed=9, beam_size=3)
    agent, we, wd = make_follower(env)
    with torch.no_grad():
        trajs, completed, traversed = agent.state_factored_search(completion_size=3, successor_size=1,
                                                                  load_next_minibatch=True)
    assert len(trajs) == 8 and len(completed) == 8 and len(traversed) == 8
    for cands, states, walk in zip(trajs, completed, traversed):
        assert 1 <= len(cands) <= 3
        sc = [float(c["score"]) for c in cands]
        assert sc == sorted(sc, reverse=True)
        keys = [tuple(s.world_state[0:4]) for s in states]
        assert len(set(keys)) == len(keys)                                   # one candidate per end world state
        for c in cands:
            assert abs(sum(float(x) for x in c["scores"]) - float(c["score"])) < 1e-4
            assert len(c["trajectory"]) == len(c["actions"]) + 1
            assert c["actions"][-1] == 0 or len(c["actions"]) == agent.episode_len
        assert walk[0].prev_inference_state is None                          # starts at the root
        for a, b in zip(walk[:-1], walk[1:]):                                # a connected walk on the navigation graph
            va, vb = a.world_state.viewpointId, b.world_state.viewpointId
            assert va == vb or vb in env.adj[va], (va, vb)
        assert walk[-1].world_state.viewpointId == states[-1].world_state.viewpointId
    flat = [c for cands in trajs for c in cands]
    with torch.no_grad():
        scored, _ = agent._score_obs_actions_and_instructions([c["observations"] for c in flat], [c["actions"] for c in flat],
                                                              [c["instr_encoding"] for c in flat])
    for s, c in zip(scored, flat):
        assert abs(s["score"] - float(c["score"])) < 3e-4, (s["score"], c["score"])
    # successor_size > 1 expands more states per iteration and still satisfies the same contract
    env2 = FakeR2RBatch(n_viewpoints=20, n_instr=8, batch_size=8, seed=9, beam_size=3)
    agent2, _, _ = make_follower(env2)
    with torch.no_grad():
        trajs2, _, _ = agent2.state_factored_search(completion_size=2, successor_size=3, load_next_minibatch=True)
    assert all(1 <= len(c) <= 2 for c in trajs2)


def test_speaker_beam_search():
    """speaker.py:211-318: beam 1 == greedy argmax decode; wider beams are sorted, scores add up, best beam >= greedy."""
    env = FakeR2RBatch(n_viewpoints=16, n_instr=6, batch_size=6, seed=6)
    we, wd = synth.speaker_encoder_weights(), synth.speaker_decoder_weights()
    enc = M.SpeakerEncoderLSTM(synth.FEAT, synth.FEAT, synth.HID, 0.5).cuda().eval()
    dec = M.SpeakerDecoderLSTM(synth.VOCAB, synth.WORD, synth.HID, 0.5, glove=wd["embedding.weight"].numpy()).cuda().eval()
    enc.load_state_dict(we); dec.load_state_dict(wd)
    spk = Sp.Seq2SeqSpeaker(env, "", enc, dec, instruction_len=12, max_episode_len=6)
    path_obs, path_actions, encoded = env.gold_obs_actions_and_instructions(6)
    with torch.no_grad():
        greedy, _ = spk._score_obs_actions_and_instructions(path_obs, path_actions, encoded, "argmax")
        b1 = spk.beam_search(1, path_obs, path_actions)
        b4 = spk.beam_search(4, path_obs, path_actions)
    for g, one, four in zip(greedy, b1, b4):
        assert len(one) == 1 and one[0]["instr_id"] == g["instr_id"]
        assert one[0]["word_indices"] == g["word_indices"]
        assert abs(float(one[0]["score"]) - g["score"]) < 1e-3
        sc = [float(x["score"]) for x in four]
        assert 1 <= len(four) <= 4 and sc == sorted(sc, reverse=True)
        assert sc[0] >= float(one[0]["score"]) - 1e-4
        for x in four:
            assert abs(sum(float(s) for s in x["scores"]) - float(x["score"])) < 1e-3
            assert x["word_indices"][-1] == 2 or len(x["word_indices"]) == 12      # <EOS> or instruction_len


# ------------------------------------------------------------------ search loops against the search oracle (real graphs)
from oracle import search_oracle as SO   # noqa: E402


def _same_candidates(got, want, tol=2e-4, what=""):
    """Candidate lists of one instance: same length, same action sequences and end states in the same order, scores
    within tol — unless the oracle's own neighbouring scores are closer than 10 tol (an order flip would be legal)."""
    sw = [c["score"] for c in want]
    if any(abs(a - b) < 10 * tol for a, b in zip(sw[:-1], sw[1:])):
        return False
    assert len(got) == len(want), (what, len(got), len(want))
    for g, w in zip(got, want):
        assert [int(a) for a in g["actions"]] == [int(a) for a in w["actions"]], what
        assert [p[0] for p in g["trajectory"]] == [p[0] for p in w["trajectory"]], what
        assert abs(float(g["score"]) - float(w["score"])) < tol, (what, g["score"], w["score"])
        assert np.allclose([float(x) for x in g["scores"]], [float(x) for x in w["scores"]], atol=tol), what
    return True


@pytest.mark.parametrize("graph,beam", [("8194nk5LbLH", 3), ("pLe4wQe7qrG", 4)])
def test_beam_search_matches_search_oracle(graph, beam):
    """follower.py:541-718 — expansion, pruning, completion and ordering of the product's beam search equal the
    plain-Python oracle's on a real R2R navigation graph (twin environments, same seed)."""
    env = FakeR2RBatch(n_instr=6, batch_size=6, seed=21, graph=graph, beam_size=beam)
    twin = FakeR2RBatch(n_instr=6, batch_size=6, seed=21, graph=graph, beam_size=beam)
    agent, we, wd = make_follower(env)
    with torch.no_grad():
        got, _, _ = agent.beam_search(beam)
        want, _ = SO.follower_beam_search(twin, we, wd, beam, episode_len=agent.episode_len,
                                          max_length=agent.max_instruction_length)
    checked = sum(_same_candidates(g, w, what="instance %d" % i) for i, (g, w) in enumerate(zip(got, want)))
    assert checked >= 4


@pytest.mark.parametrize("graph,completion,successor,store", [("8194nk5LbLH", 3, 1, False), ("GdvgFV5R1Z5", 4, 2, False),
                                                             ("pLe4wQe7qrG", 5, 1, False), ("pLe4wQe7qrG", 5, 1, True),
                                                             ("8194nk5LbLH", 6, 1, True), ("GdvgFV5R1Z5", 4, 2, True)])
def test_state_factored_search_matches_search_oracle(graph, completion, successor, store):
    """follower.py:720-980 — world-state dedupe, strict-improvement replacement, heapq.nlargest expansion order,
    completion bookkeeping and the traversal walk equal the oracle's.  store=True: the index mode (observations WITHOUT
    slabs and action-embedding rows, everything gathered from the device feature store) with the lazy-heap selection."""
    env = FakeR2RBatch(n_instr=5, batch_size=5, seed=22, graph=graph, beam_size=max(successor, 2), with_features=not store)
    twin = FakeR2RBatch(n_instr=5, batch_size=5, seed=22, graph=graph, beam_size=max(successor, 2))
    agent, we, wd = make_follower(env, store=store)
    with torch.no_grad():
        got, _, walk_g = agent.state_factored_search(completion, successor)
        want, _, walk_w = SO.follower_state_factored_search(twin, we, wd, completion, successor, episode_len=agent.episode_len,
                                                            max_length=agent.max_instruction_length)
    checked = 0
    for i, (g, w) in enumerate(zip(got, want)):
        if _same_candidates(g, w, tol=3e-4, what="instance %d" % i):
            checked += 1
            assert [s.world_state.viewpointId for s in walk_g[i]] == [s.world_state.viewpointId for s in walk_w[i]], i
    assert checked >= 3


@pytest.mark.parametrize("graph,completion", [("8194nk5LbLH", 4), ("pLe4wQe7qrG", 6), ("GdvgFV5R1Z5", 8)])
def test_device_state_factored_search_matches_host_search_and_oracle(graph, completion):
    """the state-factored search with its state on the device (sfb_sf_search_update: cache / holding /
    completed tables, node pool, selection; successor_size = 1) against the host implementation of follower.py:720-980 AND
    against the plain-Python search oracle, on a real R2R navigation graph: same candidates (end state, actions, scores)
    and the same traversal walk."""
    from speaker_follower_b200.navgraph_env import DeviceNavTables
    mk = lambda feats: FakeR2RBatch(n_instr=6, batch_size=6, seed=23, graph=graph, beam_size=2, with_features=feats)
    env_d, env_h, twin = mk(False), mk(False), mk(True)
    agent_d, we, wd = make_follower(env_d, store=True)
    agent_h, _, _ = make_follower(env_h, store=True)
    nav = DeviceNavTables(env_d, "cuda", with_teacher=False)
    with torch.no_grad():
        got, _, walk_g = agent_d.device_state_factored_search(nav, completion, cuda_graph=(completion == 6))   # one case through the graph
        host, _, walk_h = agent_h.state_factored_search(completion, 1)
        want, _, walk_w = SO.follower_state_factored_search(twin, we, wd, completion, 1, episode_len=agent_d.episode_len,
                                                            max_length=agent_d.max_instruction_length)
    assert agent_d.last_search_iterations > completion
    checked = 0
    for i, (g, h, w) in enumerate(zip(got, host, want)):
        if _same_candidates(g, h, tol=3e-4, what="device vs host, instance %d" % i) and \
                _same_candidates(g, w, tol=3e-4, what="device vs oracle, instance %d" % i):
            checked += 1
            vps = lambda walk: [s_.world_state.viewpointId for s_ in walk]
            assert vps(walk_g[i]) == vps(walk_h[i]) == vps(walk_w[i]), i
            for cg, ch in zip(g, h):
                assert cg["trajectory"] == ch["trajectory"] and len(cg["attentions"]) == len(ch["attentions"])
    assert checked >= 4


def test_device_state_factored_search_long_run_equals_host_bit_for_bit():
    """The pragmatic-inference configuration at a size where near-ties would flip a rounding-level comparison (24 instances,
    completion 30, episode_len 10, > 100 iterations): both searches run the same kernels on the same batch (host side padded
    to one row per instance, device side without per-episode projections), so every candidate, score and the traversal
    walk must be EQUAL — what is compared is the search logic of sfb_sf_search_update against follower.py:886-924."""
    from speaker_follower_b200.navgraph_env import DeviceNavTables
    mk = lambda: FakeR2RBatch(n_viewpoints=120, n_instr=24, batch_size=24, seed=31, max_len=30, beam_size=30, with_features=False)
    env_d, env_h = mk(), mk()
    agent_d, _, _ = make_follower(env_d, store=True)
    agent_h, _, _ = make_follower(env_h, store=True)
    agent_d.episode_len = agent_h.episode_len = 10
    nav = DeviceNavTables(env_d, "cuda", with_teacher=False)
    with torch.no_grad():
        got, _, walk_g = agent_d.device_state_factored_search(nav, 30, use_ctx_proj=False)
        want, _, walk_w = agent_h.state_factored_search(30, 1, _pad_batch=True)
    assert agent_d.last_search_iterations > 100
    assert [len(g) for g in got] == [len(w) for w in want] and sum(len(g) for g in got) > 24 * 25   # (a few instances run out of states)
    for i, (g, w) in enumerate(zip(got, want)):
        for cg, cw in zip(g, w):
            assert [int(a) for a in cg["actions"]] == [int(a) for a in cw["actions"]], i
            assert cg["trajectory"] == cw["trajectory"], i
            assert abs(float(cg["score"]) - float(cw["score"])) < 1e-6, i
        assert [x.world_state for x in walk_g[i]] == [x.world_state for x in walk_w[i]], i


def test_speaker_beam_search_matches_search_oracle():
    """speaker.py:211-318 against the oracle on gold paths of a real graph."""
    env = FakeR2RBatch(n_instr=6, batch_size=6, seed=23, graph="8194nk5LbLH")
    we, wd = synth.speaker_encoder_weights(), synth.speaker_decoder_weights()
    # random-init word logits are nearly uniform (beam scores would tie within fp32 noise): spread them out
    wd = dict(wd)
    wd["decoder2action.weight"] = wd["decoder2action.weight"] * 40.0
    enc = M.SpeakerEncoderLSTM(synth.FEAT, synth.FEAT, synth.HID, 0.5).cuda().eval()
    dec = M.SpeakerDecoderLSTM(synth.VOCAB, synth.WORD, synth.HID, 0.5, glove=wd["embedding.weight"].numpy()).cuda().eval()
    enc.load_state_dict(we); dec.load_state_dict(wd)
    spk = Sp.Seq2SeqSpeaker(env, "", enc, dec, instruction_len=10, max_episode_len=6)
    path_obs, path_actions, _ = env.gold_obs_actions_and_instructions(6)
    with torch.no_grad():
        got = spk.beam_search(3, path_obs, path_actions)
        want = SO.speaker_beam_search(path_obs, path_actions, we, wd, 3, instruction_len=10)
    checked = 0
    for g, w in zip(got, want):
        sw = [c["score"] for c in w]
        if any(abs(a - b) < 2e-3 for a, b in zip(sw[:-1], sw[1:])):
            continue
        checked += 1
        assert len(g) == len(w)
        for a, b in zip(g, w):
            assert [int(x) for x in a["word_indices"]] == [int(x) for x in b["word_indices"]]
            assert abs(float(a["score"]) - float(b["score"])) < 1e-3
    assert checked >= 3


# ------------------------------------------------------------------ pragmatic inference / data augmentation (C4 / C5)
def make_speaker(env, instruction_len=10):
    we, wd = synth.speaker_encoder_weights(), synth.speaker_decoder_weights()
    enc = M.SpeakerEncoderLSTM(synth.FEAT, synth.FEAT, synth.HID, 0.5).cuda().eval()
    dec = M.SpeakerDecoderLSTM(synth.VOCAB, synth.WORD, synth.HID, 0.5, glove=wd["embedding.weight"].numpy()).cuda().eval()
    enc.load_state_dict(we); dec.load_state_dict(wd)
    return Sp.Seq2SeqSpeaker(env, "", enc, dec, instruction_len=instruction_len, max_episode_len=6), we, wd


def test_rational_follower_pipeline_and_combine():
    """rational_follower.py:11-150 on a real graph: state-factored candidates, speaker rescoring of every candidate,
    product combine == oracle combine on the same records, weight 0 picks the follower's own best candidate."""
    from speaker_follower_b200 import pragmatic as PR
    env = FakeR2RBatch(n_instr=6, batch_size=3, seed=31, graph="pLe4wQe7qrG", beam_size=4)
    agent, _, _ = make_follower(env)
    spk, _, _ = make_speaker(env)
    by_w, cands, records = PR.run_rational_follower(env, agent, spk, beam_size=4)
    assert len(cands) == 6 and records.shape[1] == 4 and records.shape[0] == sum(len(c) for c in cands.values())
    groups = [int(g) for g in records[:, 0]]
    for w in (0.0, 0.95):
        best = O.rational_combine(records[:, 3], records[:, 2], groups, w)
        order = list(cands.keys())
        for k, iid in enumerate(order):
            chosen = by_w[w]["results"][iid]
            assert chosen is cands[iid][int(records[best[k], 1])]
    for iid, c in cands.items():                       # weight 0 = follower score alone
        assert by_w[0.0]["results"][iid] is max(c, key=lambda x: x["follower_score"])
        assert all("speaker_score" in x and "observations" not in x for x in c)


def test_speaker_generation_records_in_trajectory_order(tmp_path):
    """data_augmentation_from_speaker.py:66 (literal speaker): one record per trajectory, in env order, JSON written."""
    from speaker_follower_b200 import pragmatic as PR
    env = FakeR2RBatch(n_instr=7, batch_size=3, seed=32, graph="8194nk5LbLH")
    spk, _, _ = make_speaker(env, instruction_len=8)
    out = PR.generate_speaker_instructions(env, spk, path=str(tmp_path / "aug.json"))
    assert [r["instr_id"] for r in out] == [it["instr_id"] for it in env.data]
    assert all(1 <= len(r["word_indices"]) <= 8 for r in out)
    import json
    assert json.load(open(tmp_path / "aug.json")) == out


@pytest.mark.parametrize("graph_capture", [False, True])
def test_device_rollout_matches_host_loop(graph_capture):
    """SURVEY f-2: the table-driven environment (DeviceNavTables + sfb_nav_step) reproduces R2RBatch.step / observe, and
    the fully device-resident rollout (optionally one CUDA graph per episode) takes the same actions with the same scores
    as the host-driven rollout — for the staged host loop and, through it, for the oracle (test_greedy_rollout...)."""
    from speaker_follower_b200.navgraph_env import DeviceNavTables
    env = FakeR2RBatch(n_instr=8, batch_size=8, seed=41, graph="pLe4wQe7qrG")
    agent, we, wd = make_follower(env, store=True)
    agent.feedback = "argmax"
    with torch.no_grad():
        traj = agent.rollout()                                  # staged host loop (vp_index observations + device store)
    oracle_check(agent, traj, we, wd, "argmax")
    nav = DeviceNavTables(env, "cuda")
    batch = env.batch                                           # the minibatch the rollout ran on, sorted by length
    ws0 = [WorldStateOf(it) for it in batch]
    res = agent.device_rollout(nav, nav.state_ids(ws0), [it["goal"] for it in batch], [it["instr_encoding"] for it in batch],
                               cuda_graph=graph_capture)
    acts, scores = res["actions"].cpu(), res["scores"].cpu()
    for i, t in enumerate(traj):
        n = len(t["actions"])
        assert acts[i, :n].tolist() == [int(a) for a in t["actions"]], i
        assert (acts[i, n:] == -1).all()
        assert np.allclose(scores[i, :n].numpy(), np.array(t["scores"], dtype=np.float32), atol=1e-5)
        # the table-driven environment ended where the simulator stand-in did
        assert int(res["final_state"][i]) // nav.HEADINGS == t["trajectory"][-1][0]


def WorldStateOf(item):
    from speaker_follower_b200.navgraph_env import WorldState
    return WorldState("fake", item["start"], item["heading"], 0.0)
