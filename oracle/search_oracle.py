"""CPU oracle for the search loops of the follower and the speaker — TEST INFRASTRUCTURE ONLY.

Plain-Python restatements (lists, dicts, namedtuples, ``heapq``; the model arithmetic comes from
``oracle/r2r_oracle.py`` on CPU tensors) of

    Seq2SeqAgent.beam_search              tasks/R2R/follower.py:541-718
    Seq2SeqAgent.state_factored_search    tasks/R2R/follower.py:720-980
    Seq2SeqSpeaker.beam_search            tasks/R2R/speaker.py:211-318

and of the host batching they rely on (``batch_instructions_from_encoded`` follower.py:75-105,
``_action_variable`` follower.py:300-320, ``_batch_observations_and_actions`` speaker.py:68-121,
``backchain_inference_states`` / ``least_common_viewpoint_path`` follower.py:31-73).  The reference's agent
code cannot run unchanged on a modern torch (``x.data[0]`` on 0-dim tensors, unconditional ``.cuda()``,
SURVEY.md §8c), which is why this is a restatement rather than an import; every block cites the lines it
follows.  Nothing under ``speaker_follower_b200/`` imports this file: the product's search loops are compared
against it in ``tests/`` on twin environments (same seed).

``env`` is any object with the R2RBatch methods the reference uses: ``reset(sort, beamed,
load_next_minibatch)``, ``observe(world_states, beamed)``, ``step(world_states, actions, last_obs, beamed)``
and observations that follow SURVEY.md A.3.
"""
from __future__ import annotations

import heapq
import itertools
from collections import namedtuple

import numpy as np
import torch

from . import r2r_oracle as O

PAD, UNK, EOS, BOS = 0, 1, 2, 3                      # utils.py:19-24

InferenceState = namedtuple(
    "InferenceState", "prev_inference_state, world_state, observation, flat_index, last_action, "
                      "last_action_embedding, action_count, score, h_t, c_t, last_alpha")      # follower.py:18
SpeakerState = namedtuple("SpeakerState", "prev_inference_state, flat_index, last_word, word_count, score, last_alpha")

Cons = namedtuple("Cons", "first, rest")


# ------------------------------------------------------------------ host batching
def batch_instructions(encoded, max_length=80, reverse=True):
    """follower.py:75-105 (sort=False): reversed tokens + <EOS>, truncated, <PAD>-padded; mask = pad positions."""
    n = len(encoded)
    seq = np.full((n, max_length), PAD, dtype=np.int64)
    lengths = []
    for i, inst in enumerate(encoded):
        inst = np.asarray(inst, dtype=np.int64)
        if reverse:
            inst = inst[::-1]
        inst = np.concatenate((inst, [EOS]))[:max_length]
        seq[i, :len(inst)] = inst
        lengths.append(len(inst))
    seq = torch.from_numpy(seq)
    mask = (seq == PAD)[:, :max(lengths)]
    return seq, mask, lengths


def action_variable(obs):
    """follower.py:300-320: zero-padded [N, max_a, E] action embeddings + validity."""
    max_a = max(len(ob["adj_loc_list"]) for ob in obs)
    E = obs[0]["action_embedding"].shape[-1]
    valid = np.zeros((len(obs), max_a), np.float32)
    emb = np.zeros((len(obs), max_a, E), np.float32)
    for i, ob in enumerate(obs):
        n = len(ob["adj_loc_list"])
        valid[i, :n] = 1.0
        emb[i, :n] = ob["action_embedding"]
    return torch.from_numpy(emb), torch.from_numpy(valid), valid


def feature_variable(obs):
    """follower.py:291-298 / env.py:330-332: np.stack of the single mean-pooled slab per observation."""
    return torch.from_numpy(np.stack([ob["feature"][0] for ob in obs]))


def backchain(last):
    """follower.py:31-50."""
    states, observations, actions, scores, attentions = [], [], [], [], []
    st, last_score = last, None
    while st is not None:
        states.append(st.world_state)
        observations.append(st.observation)
        actions.append(st.last_action)
        attentions.append(st.last_alpha)
        if last_score is not None:
            scores.append(last_score - st.score)
        last_score = st.score
        st = st.prev_inference_state
    scores.append(last_score)
    return (list(reversed(states)), list(reversed(observations)), list(reversed(actions))[1:],
            list(reversed(scores))[1:], list(reversed(attentions))[1:])


def least_common_viewpoint_path(a, b):
    """follower.py:52-73."""
    to_b = {}
    stack = Cons(b, None)
    while b is not None:
        to_b[b.world_state.viewpointId] = stack
        b = b.prev_inference_state
        stack = Cons(b, stack)
    path = [a]
    while a is not None:
        vp = a.world_state.viewpointId
        if vp in to_b:
            rest, c = [], to_b[vp]
            while c is not None:
                rest.append(c.first)
                c = c.rest
            return path + rest[1:]
        a = a.prev_inference_state
        path.append(a)
    raise AssertionError("no common ancestor found")


def _traj(st):
    states, observations, actions, scores, attentions = backchain(st)
    return {"instr_id": observations[0]["instr_id"], "instr_encoding": observations[0]["instr_encoding"],
            "trajectory": [(ob["viewpoint"], ob["heading"], ob["elevation"]) for ob in observations],   # follower.py:283
            "observations": observations, "actions": [int(a) for a in actions], "score": float(st.score),
            "scores": [float(s) for s in scores], "attentions": attentions}


def _encode(env, world_states, obs, enc_w, max_length, reverse):
    seq, mask, lengths = batch_instructions([o[0]["instr_encoding"] for o in obs], max_length, reverse)
    ctx, h_t, c_t = O.encoder_lstm(seq[:, :max(lengths)], lengths, enc_w)
    return ctx, h_t, c_t, mask


# ------------------------------------------------------------------ follower beam search (follower.py:541-718)
def follower_beam_search(env, enc_w, dec_w, beam_size, episode_len=10, max_length=80, reverse=True,
                         load_next_minibatch=True):
    assert env.beam_size >= beam_size
    world_states = env.reset(sort=True, beamed=True, load_next_minibatch=load_next_minibatch)
    obs = env.observe(world_states, beamed=True)
    n = len(world_states)
    ctx, h_t, c_t, seq_mask = _encode(env, world_states, obs, enc_w, max_length, reverse)
    E = dec_w["lstm.weight_ih"].shape[1] - dec_w["visual_attention_layer.linear_in_v.weight"].shape[1]
    u_begin = torch.zeros(E)
    completed = [[] for _ in range(n)]
    beams = [[InferenceState(None, ws[0], o[0], i, -1, u_begin, 0, 0.0, None, None, None)]
             for i, (ws, o) in enumerate(zip(world_states, obs))]
    for t in range(episode_len):
        flat_indices = [s.flat_index for beam in beams for s in beam]
        beam_indices = [bi for bi, beam in enumerate(beams) for _ in beam]
        u_prev = torch.stack([s.last_action_embedding for beam in beams for s in beam], 0)
        flat_obs = [o for os_ in obs for o in os_]
        f_t = feature_variable(flat_obs)
        all_u_t, is_valid, is_valid_np = action_variable(flat_obs)
        h_t, c_t, alpha, logit, _ = O.attn_decoder_step(u_prev, all_u_t, f_t, h_t[flat_indices], c_t[flat_indices],
                                                        ctx[beam_indices], seq_mask[beam_indices], dec_w)
        logit = logit.masked_fill(is_valid == 0, -float("inf"))                                   # 600
        log_probs = torch.log_softmax(logit, dim=1)
        _, action_indices = logit.topk(min(beam_size, logit.shape[1]), dim=1)                     # 608
        action_scores = log_probs.gather(1, action_indices)                                       # 609
        start, all_succ = 0, []
        for beam, beam_ws, beam_obs in zip(beams, world_states, obs):
            succ = []
            for j, (st, ws, ob) in enumerate(zip(beam, beam_ws, beam_obs)):
                fi = start + j
                for sc, ai in zip(action_scores[fi].tolist(), action_indices[fi].tolist()):
                    if is_valid_np[fi, ai] == 0:                                                  # 626
                        continue
                    succ.append(InferenceState(st, ws, ob, fi, ai, all_u_t[fi, ai], st.action_count + 1,
                                               float(st.score + sc), None, None, alpha[fi]))
            start += len(beam)
            all_succ.append(sorted(succ, key=lambda s: s.score, reverse=True)[:beam_size])        # 640
        new_ws = env.step([[s.world_state for s in ss] for ss in all_succ], [[s.last_action for s in ss] for ss in all_succ],
                          [[s.observation for s in ss] for ss in all_succ], beamed=True)
        new_obs = env.observe(new_ws, beamed=True)
        all_succ = [[s._replace(world_state=w, observation=o) for s, w, o in zip(ss, ws_, os_)]
                    for ss, ws_, os_ in zip(all_succ, new_ws, new_obs)]
        beams = []
        for bi, ss in enumerate(all_succ):
            nb = []
            for s in ss:
                if s.last_action == 0 or t == episode_len - 1:                                    # 670
                    completed[bi].append(s)
                else:
                    nb.append(s)
            if len(completed[bi]) >= beam_size:                                                   # 674-675
                nb = []
            beams.append(nb)
        world_states = [[s.world_state for s in beam] for beam in beams]
        obs = [[s.observation for s in beam] for beam in beams]
        if not any(beams):
            break
    trajs = []
    for done in completed:
        assert done
        trajs.append([_traj(s) for s in sorted(done, key=lambda s: s.score, reverse=True)[:beam_size]])
    return trajs, completed


# ------------------------------------------------------------------ state-factored search (follower.py:720-980)
def follower_state_factored_search(env, enc_w, dec_w, completion_size, successor_size, episode_len=10, max_length=80,
                                   reverse=True, first_n_ws_key=4, load_next_minibatch=True):
    assert env.beam_size >= successor_size
    world_states = env.reset(sort=True, beamed=True, load_next_minibatch=load_next_minibatch)
    initial_obs = env.observe(world_states, beamed=True)
    n = len(world_states)
    ctx, h_t, c_t, seq_mask = _encode(env, world_states, initial_obs, enc_w, max_length, reverse)
    E = dec_w["lstm.weight_ih"].shape[1] - dec_w["visual_attention_layer.linear_in_v.weight"].shape[1]
    u_begin = torch.zeros(E)
    completed = [dict() for _ in range(n)]
    holding = [dict() for _ in range(n)]
    cache = [{tuple(ws[0][0:first_n_ws_key]): (InferenceState(None, ws[0], o[0], None, -1, u_begin, 0, 0.0, h_t[i], c_t[i], None), True)}
             for i, (ws, o) in enumerate(zip(world_states, initial_obs))]                          # 741-752
    beams = [[st for _, (st, _e) in sorted(c.items())] for c in cache]
    last_expanded = [beam[0] for beam in beams]
    traversed = [[beam[0]] for beam in beams]

    def update_traversed(groups):                                                                  # 768-783
        for i, group in enumerate(groups):
            cur = last_expanded[i]
            assert cur.world_state.viewpointId == traversed[i][-1].world_state.viewpointId
            for st in group:
                walk = least_common_viewpoint_path(cur, st)
                assert walk[0].world_state.viewpointId == cur.world_state.viewpointId
                assert walk[-1].world_state.viewpointId == st.world_state.viewpointId
                traversed[i].extend(walk[1:])
                cur = st
            last_expanded[i] = cur

    while any(len(c) < completion_size for c in completed):                                        # 786
        beam_indices = [bi for bi, beam in enumerate(beams) for _ in beam]
        flat = [st for beam in beams for st in beam]
        flat_obs = [st.observation for st in flat]
        u_prev = torch.stack([st.last_action_embedding for st in flat], 0)
        hh = torch.stack([st.h_t for st in flat], 0)
        cc = torch.stack([st.c_t for st in flat], 0)
        f_t = feature_variable(flat_obs)
        all_u_t, is_valid, is_valid_np = action_variable(flat_obs)
        h_new, c_new, alpha, logit, _ = O.attn_decoder_step(u_prev, all_u_t, f_t, hh, cc, ctx[beam_indices],
                                                            seq_mask[beam_indices], dec_w)
        logit = logit.masked_fill(is_valid == 0, -float("inf"))                                   # 815
        log_probs = torch.log_softmax(logit, dim=1)
        start, all_succ = 0, []
        for beam in beams:
            succ = []
            for j, st in enumerate(beam):
                fi = start + j
                for ai, sc in enumerate(log_probs[fi].tolist()):                                  # every valid action: 846-861
                    if is_valid_np[fi, ai] == 0:
                        continue
                    succ.append(InferenceState(st, st.world_state, flat_obs[fi], None, ai, all_u_t[fi, ai],
                                               st.action_count + 1, float(np.float32(st.score) + np.float32(sc)),
                                               h_new[fi], c_new[fi], alpha[fi]))
            start += len(beam)
            all_succ.append(sorted(succ, key=lambda s: s.score, reverse=True))
        new_ws = env.step([[s.world_state for s in ss] for ss in all_succ], [[s.last_action for s in ss] for ss in all_succ],
                          [[s.observation for s in ss] for ss in all_succ], beamed=True)
        all_succ = [[s._replace(world_state=w) for s, w in zip(ss, ws_)] for ss, ws_ in zip(all_succ, new_ws)]
        new_beams = []
        for i, succ in enumerate(all_succ):
            if len(completed[i]) >= completion_size:                                              # 893-895
                new_beams.append([])
                continue
            for s in succ:
                k = tuple(s.world_state[0:first_n_ws_key])
                if s.last_action == 0 or s.action_count == episode_len:                           # 898-903
                    if k not in holding[i] or holding[i][k][0].score < s.score:
                        holding[i][k] = (s, False)
                elif k not in cache[i] or cache[i][k][0].score < s.score:
                    cache[i][k] = (s, False)
            pool = itertools.chain(((k, st, False) for k, (st, e) in cache[i].items() if not e),
                                   ((k, st, True) for k, (st, e) in holding[i].items() if not e))
            beam = []
            for k, st, fin in heapq.nlargest(successor_size, pool, key=lambda p: p[1].score):     # 906-912
                if fin:
                    holding[i][k] = (st, True)
                    if k not in completed[i] or completed[i][k].score < st.score:
                        completed[i][k] = st
                else:
                    cache[i][k] = (st, True)
                    beam.append(st)
            new_beams.append([] if len(completed[i]) >= completion_size else beam)
        beams = new_beams
        if not any(beams):
            break
        new_obs = env.observe([[st.world_state for st in beam] for beam in beams], beamed=True)
        beams = [[st._replace(observation=o) for st, o in zip(beam, os_)] for beam, os_ in zip(beams, new_obs)]
        update_traversed(beams)
    completed_list = [sorted(c.values(), key=lambda s: s.score, reverse=True)[:completion_size] for c in completed]
    final_obs = env.observe([[st.world_state for st in cl] for cl in completed_list], beamed=True)
    completed_list = [[st._replace(observation=o) for st, o in zip(cl, os_)] for cl, os_ in zip(completed_list, final_obs)]
    update_traversed(completed_list)
    trajs = [[_traj(st) for st in cl] for cl in completed_list]
    return trajs, completed_list, traversed


# ------------------------------------------------------------------ speaker beam search (speaker.py:211-318)
def batch_observations_and_actions(path_obs, path_actions):
    """speaker.py:68-121: T arrays [N,...] zero-initialised, per-item copy, mask = 1 on padded path steps."""
    lengths = np.array([len(a) for a in path_actions])
    T, N = int(lengths.max()), len(path_obs)
    mask = np.ones((N, T), np.uint8)
    E = path_obs[0][0]["action_embedding"].shape[-1]
    acts = [np.zeros((N, E), np.float32) for _ in range(T)]
    feats = [np.zeros((N,) + path_obs[0][0]["feature"][0].shape, np.float32) for _ in range(T)]
    for i, (obs, actions) in enumerate(zip(path_obs, path_actions)):
        assert len(obs) == len(actions) + 1
        mask[i, :len(actions)] = 0
        for t, (ob, a) in enumerate(zip(obs[:-1], actions)):
            feats[t][i] = ob["feature"][0]
            acts[t][i] = ob["action_embedding"][a]
    return ([torch.from_numpy(f) for f in feats], [torch.from_numpy(a) for a in acts], torch.from_numpy(mask).bool(),
            [obs[0] for obs in path_obs])


def speaker_beam_search(path_obs, path_actions, enc_w, dec_w, beam_size, instruction_len=80):
    feats, acts, path_mask, start_obs = batch_observations_and_actions(path_obs, path_actions)
    n = len(start_obs)
    ctx, h_t, c_t = O.speaker_encoder(acts, feats, enc_w)
    completed = [[] for _ in range(n)]
    beams = [[SpeakerState(None, i, BOS, 0, 0.0, None)] for i in range(n)]
    for t in range(instruction_len):
        flat_indices = [s.flat_index for beam in beams for s in beam]
        beam_indices = [bi for bi, beam in enumerate(beams) for _ in beam]
        w_t = torch.tensor([s.last_word for beam in beams for s in beam], dtype=torch.long)
        h_t, c_t, alpha, logit = O.speaker_decoder_step(w_t, h_t[flat_indices], c_t[flat_indices], ctx[beam_indices],
                                                        path_mask[beam_indices], dec_w)
        log_probs = torch.log_softmax(logit, dim=1)
        _, word_indices = logit.topk(min(beam_size, logit.shape[1]), dim=1)                       # 254
        word_scores = log_probs.gather(1, word_indices)
        start, new_beams = 0, []
        for bi, beam in enumerate(beams):
            succ = []
            for j, st in enumerate(beam):
                fi = start + j
                for sc, wi in zip(word_scores[fi].tolist(), word_indices[fi].tolist()):
                    succ.append(SpeakerState(st, fi, wi, st.word_count + 1, float(np.float32(st.score) + np.float32(sc)), alpha[fi]))
            start += len(beam)
            succ = sorted(succ, key=lambda s: s.score, reverse=True)[:beam_size]
            nb = []
            for s in succ:
                if s.last_word == EOS or t == instruction_len - 1:                                # 284
                    completed[bi].append(s)
                else:
                    nb.append(s)
            if len(completed[bi]) >= beam_size:
                nb = []
            new_beams.append(nb)
        beams = new_beams
        if not any(beams):
            break
    outputs = []
    for i in range(n):
        outs = []
        for st in sorted(completed[i], key=lambda s: s.score, reverse=True)[:beam_size]:
            words, scores, s, last = [], [], st, None
            while s is not None:                                                                  # speaker.py:20-34
                words.append(s.last_word)
                if last is not None:
                    scores.append(last - s.score)
                last = s.score
                s = s.prev_inference_state
            scores.append(last)
            outs.append({"instr_id": start_obs[i]["instr_id"], "word_indices": list(reversed(words))[1:],
                         "score": float(st.score), "scores": list(reversed(scores))[1:]})
        outputs.append(outs)
    return outputs
