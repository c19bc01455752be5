"""Follower agent on the sm_100a kernels: the control loops of tasks/R2R/follower.py:Seq2SeqAgent (261-1035)
with the same class / method names, arguments and result dictionaries, modernised for today's torch (bool masks,
`.item()`), and with every per-step tensor operation on the device:

    _feature_variables (291-298)      -> indices into a device-resident FeatureStore when the env provides
                                         `vp_index` (else the reference's host stack + copy)
    decoder(...) (473)                -> sfb_follower_step_fwd  (one call)
    mask / CE / argmax|sample / u_t_prev gather / score (476-505)  -> sfb_follower_step_tail (one call)

Only `a_t` (B int32) crosses back to the host each step, because the simulator needs it (follower.py:509-513).
Written from the behaviour described in SURVEY.md A.2; no reference code is reused.
"""
from __future__ import annotations

import json
import math
import random
from collections import namedtuple
from typing import List, Optional

import numpy as np
import torch

from . import ops

vocab_pad_idx, vocab_unk_idx, vocab_eos_idx, vocab_bos_idx = 0, 1, 2, 3      # utils.py:19-24

InferenceState = namedtuple("InferenceState", "prev_inference_state, world_state, observation, flat_index, last_action, "
                            "last_action_embedding, action_count, score, h_t, c_t, last_alpha")   # follower.py:19


def _device(module) -> torch.device:
    return next(module.parameters()).device


def path_element_from_observation(ob):
    return (ob["viewpoint"], ob["heading"], ob["elevation"])


def backchain_inference_states(last):
    """follower.py:32-50: walk parent pointers back to the start; actions/scores/attentions exclude the start."""
    chain = []
    s = last
    while s is not None:
        chain.append(s)
        s = s.prev_inference_state
    chain.reverse()
    states = [s.world_state for s in chain]
    observations = [s.observation for s in chain]
    actions = [s.last_action for s in chain][1:]
    scores = [b.score - a.score for a, b in zip(chain[:-1], chain[1:])]
    attentions = [s.last_alpha for s in chain][1:]
    return states, observations, actions, scores, attentions


def least_common_viewpoint_path(inf_state_a, inf_state_b):
    """follower.py:52-73: the physical walk from A to B through the search tree — up from A to its first ancestor X
    whose viewpoint also occurs among B's ancestors, then down from that ancestor of B (the OLDEST one with that
    viewpoint) to B.  Returns the list of inference states, X counted once."""
    down = {}                       # viewpointId -> [ancestor, ..., B]; older ancestors overwrite newer ones
    chain, x = [], inf_state_b
    while x is not None:
        chain.insert(0, x)
        down[x.world_state.viewpointId] = list(chain)
        x = x.prev_inference_state
    up, x = [], inf_state_a
    while x is not None:
        up.append(x)
        tail = down.get(x.world_state.viewpointId)
        if tail is not None:
            return up + tail[1:]
        x = x.prev_inference_state
    raise AssertionError("no common ancestor found")


def batch_instructions_from_encoded(encoded_instructions, max_length, reverse=False, sort=False, device=None):
    """follower.py:75-105: [reversed] tokens + <EOS>, truncated to max_length, padded with <PAD>; mask = PAD
    positions cut to the longest sequence.  Returns (seq int64 [N,max_length], mask bool [N,max(len)], lengths
    [, perm_idx when sort])."""
    n = len(encoded_instructions)
    seq = np.full((n, max_length), vocab_pad_idx, dtype=np.int64)
    lengths = []
    for i, inst in enumerate(encoded_instructions):
        inst = list(inst)
        if inst:
            assert inst[-1] != vocab_eos_idx
        if reverse:
            inst = inst[::-1]
        inst = (inst + [vocab_eos_idx])[:max_length]
        seq[i, :len(inst)] = inst
        lengths.append(len(inst))
    seq_t = torch.from_numpy(seq)
    perm = None
    if sort:
        order = np.argsort(-np.asarray(lengths), kind="stable")
        perm = [int(i) for i in order]
        seq_t = seq_t[order]
        lengths = [lengths[i] for i in perm]
    mask = (seq_t == vocab_pad_idx)[:, :max(lengths)]
    if device is not None:
        seq_t, mask = seq_t.to(device), mask.to(device)
    out = (seq_t, mask, lengths)
    return out + (perm,) if sort else out


class BaseAgent(object):
    """follower.py:107-195."""

    def __init__(self, env, results_path):
        self.env = env
        self.results_path = results_path
        random.seed(1)
        self.results = {}
        self.losses = []

    def write_results(self):
        results = {k: {"instr_id": v["instr_id"], "trajectory": v["trajectory"]} for k, v in self.results.items()}
        with open(self.results_path, "w") as f:
            json.dump(results, f)

    def rollout(self):
        raise NotImplementedError

    def test(self):
        self.env.reset_epoch()
        self.losses = []
        self.results = {}
        looped = False
        while True:
            for result in self.rollout():
                if result["instr_id"] in self.results:
                    looped = True
                else:
                    self.results[result["instr_id"]] = result
            if looped:
                break
        return self.results


class Seq2SeqAgent(BaseAgent):
    """follower.py:261-1035."""
    feedback_options = ["teacher", "argmax", "sample"]

    def __init__(self, env, results_path, encoder, decoder, episode_len=10, beam_size=1, reverse_instruction=True,
                 max_instruction_length=80):
        super().__init__(env, results_path)
        self.encoder, self.decoder = encoder, decoder
        self.episode_len = episode_len
        self.losses = []
        self.beam_size = beam_size
        self.reverse_instruction = reverse_instruction
        self.max_instruction_length = max_instruction_length
        self.feedback = "argmax"
        self.loss = 0
        self._sample_gen: Optional[torch.Generator] = None

    # ---------------------------------------------------------------- batching shims (follower.py:291-332)
    def _feature_variables(self, obs, beamed=False):
        """Device feature store + (viewpoint row, viewIndex) when available, else the dense [N,36,F] host stack."""
        flat = [o for beam in obs for o in beam] if beamed else list(obs)
        dev = _device(self.decoder)
        store = getattr(self.decoder, "feature_store", None)
        if store is not None and all("vp_index" in ob for ob in flat):
            vp = torch.tensor([ob["vp_index"] for ob in flat], dtype=torch.int32, device=dev)
            view = torch.tensor([ob["viewIndex"] for ob in flat], dtype=torch.int32, device=dev)
            return [(vp, view)]
        feats = np.stack([ob["feature"][0] for ob in flat])
        return [torch.from_numpy(feats).to(dev)]

    def _action_variable(self, obs):
        dev = _device(self.decoder)
        max_a = max(len(ob["adj_loc_list"]) for ob in obs)
        dim = obs[0]["action_embedding"].shape[-1]
        is_valid = np.zeros((len(obs), max_a), np.float32)
        emb = np.zeros((len(obs), max_a, dim), np.float32)
        for i, ob in enumerate(obs):
            n = len(ob["adj_loc_list"])
            is_valid[i, :n] = 1.0
            emb[i, :n] = ob["action_embedding"]
        return torch.from_numpy(emb).to(dev), torch.from_numpy(is_valid).to(dev), is_valid

    def _action_indices(self, obs):
        """The action candidates of `obs` as indices (env.py:60-75 without the rows): per candidate the view index into the
        observation's own slab (-1: stop / padding) and sin/cos of the relative heading / elevation.  Returns
        (cand_view int32 [N,A], cand_trig [N,A,4], is_valid [N,A]) on the device plus is_valid as numpy."""
        dev = _device(self.decoder)
        max_a = max(len(ob["adj_loc_list"]) for ob in obs)
        is_valid = np.zeros((len(obs), max_a), np.float32)
        cv = np.full((len(obs), max_a), -1, np.int32)
        tr = np.zeros((len(obs), max_a, 4), np.float32)
        for i, ob in enumerate(obs):
            adj = ob["adj_loc_list"]
            is_valid[i, :len(adj)] = 1.0
            for a in range(1, len(adj)):
                d = adj[a]
                cv[i, a] = d["absViewIndex"]
                rh, re = d["rel_heading"], d["rel_elevation"]
                tr[i, a] = (math.sin(rh), math.cos(rh), math.sin(re), math.cos(re))
        return torch.from_numpy(cv).to(dev), torch.from_numpy(tr).to(dev), torch.from_numpy(is_valid).to(dev), is_valid, cv, tr

    def _embed_actions(self, vp, aview, trig):
        """Action-embedding rows (env.py:60-75) assembled on the device: vp / aview int tensors [n], trig [n,4]."""
        store = self.decoder.feature_store
        has = (aview >= 0).unsqueeze(-1).float()
        rows = store.feat_table[vp.long(), aview.clamp(min=0).long()]
        loc = store.loc_table.shape[2]
        return torch.cat((rows * has, trig.repeat_interleave(loc // 4, dim=1) * has), dim=1).contiguous()

    def _teacher_action(self, obs, ended):
        a = [(-1 if ended[i] else int(ob["teacher"])) for i, ob in enumerate(obs)]
        return torch.tensor(a, dtype=torch.int32, device=_device(self.decoder))

    def _proc_batch(self, obs, beamed=False):
        flat = [o for beam in obs for o in beam] if beamed else obs
        enc = [ob["instr_encoding"] for ob in flat]
        return batch_instructions_from_encoded(enc, self.max_instruction_length, reverse=self.reverse_instruction,
                                               device=_device(self.encoder))

    def _sample_uniform(self, n, dev):
        if self._sample_gen is None:
            self._sample_gen = torch.Generator(device=dev)
            self._sample_gen.manual_seed(torch.initial_seed() & 0x7FFFFFFF)
        return torch.rand(n, device=dev, generator=self._sample_gen)

    def _step(self, u_t_prev, obs, h_t, c_t, ctx, seq_mask, target, feedback, carry=None):
        """One decode step + tail on the device.  Returns (h, c, alpha, masked logit, a_t[int32], u_next, score, ce).
        ``carry``: a dict that lives as long as (h_t, c_t) are fed back unchanged (a rollout) — it holds the state
        this step prepares for the next one (visual query + packed gate operand; packed path), so the next call starts
        directly with the fused gather + LSTM launch."""
        f_t = self._feature_variables(obs)[0]
        all_u_t, is_valid, _ = self._action_variable(obs)
        su = self._sample_uniform(len(obs), h_t.device) if feedback == "sample" else None
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.decoder.parameters()):
            # training: the decoder call is autograd-aware; the CE terms must be differentiable w.r.t. the logits
            from . import _functional as Fn
            h_t, c_t, alpha, logit, alpha_v = self.decoder(u_t_prev, all_u_t, f_t, h_t, c_t, ctx, seq_mask)
            ce = Fn.tail_loss_terms(logit, is_valid, target) if target is not None else None
            a_t, u_next, score, _ = ops.follower_tail(logit.detach().clone(), is_valid, all_u_t, feedback, target=target, sample_u=su)
            return h_t, c_t, alpha, logit, a_t, u_next, score, ce
        if getattr(self.decoder, "supports_fused_step", False):
            tail = {"is_valid": is_valid, "feedback": feedback, "target": target, "sample_u": su}
            c_in = c_out = None
            if carry is not None:
                c_in = carry.get("state")          # written by the previous step of this rollout
                c_out = carry.get("spare")
                if c_out is None:
                    c_out = self.decoder.new_carry(len(obs), h_t.device)
                carry["spare"] = c_in
            if carry is not None and "ctx_proj" not in carry:      # once per rollout: ctx is constant over its steps
                carry["ctx_proj"] = self.decoder.project_ctx(ctx)
            h_t, c_t, alpha, logit, alpha_v = self.decoder.decode_step(
                u_t_prev, all_u_t, f_t, h_t, c_t, ctx, seq_mask, tail=tail, carry_in=c_in, carry_out=c_out,
                ctx_proj=carry.get("ctx_proj") if carry is not None else None)
            if carry is not None:
                carry["state"] = c_out
            a_t, u_next, score, ce = tail["out"]
            return h_t, c_t, alpha, logit, a_t, u_next, score, ce
        h_t, c_t, alpha, logit, alpha_v = self.decoder(u_t_prev, all_u_t, f_t, h_t, c_t, ctx, seq_mask)
        a_t, u_next, score, ce = ops.follower_tail(logit, is_valid, all_u_t, feedback, target=target, sample_u=su)
        return h_t, c_t, alpha, logit, a_t, u_next, score, ce

    # ---------------------------------------------------------------- rollouts
    def rollout(self):
        if self.beam_size == 1:
            return self._rollout_with_loss()
        beams, _, _ = self.beam_search(self.beam_size)
        return [beam[0] for beam in beams]

    # ---------------------------------------------------------------- device-resident greedy rollout
    def _staged_rollout(self, world_states, obs):
        """follower.py:430-539 for inference (no autograd) when the slabs live in the decoder's device feature store:
        per step the host ships ONE pinned staging buffer (viewpoint row, view index, candidate view indices, 4 trig
        values per candidate, validity, teacher target: ~20 KB at B=100 instead of the 38 MB the reference copies,
        follower.py:291-320), the step runs as two launches on the carried state, and the only thing read back is a_t,
        written by the last kernel straight into page-locked host memory.  Scores stay on the device until the end."""
        dev = _device(self.decoder)
        B = len(obs)
        feedback = self.feedback
        seq, seq_mask, seq_lengths = self._proc_batch(obs)
        ctx, h_t, c_t = self.encoder(seq, seq_lengths)
        ctx_proj = self.decoder.project_ctx(ctx)
        E = self.decoder.embedding_size
        A_cap = 16
        ni = 2 * B + B * A_cap + B                     # int32: vp, view, cand_view, target
        nf = 5 * B * A_cap + B                         # f32: trig [B,A,4], valid [B,A], sample_u [B]
        host = torch.empty(ni + nf, dtype=torch.int32).pin_memory()
        hi, hf = host[:ni].numpy(), host[ni:].view(torch.float32).numpy()
        stage = torch.empty(ni + nf, dtype=torch.int32, device=dev)
        di, df = stage[:ni], stage[ni:].view(torch.float32)
        a_host = torch.full((B,), -1, dtype=torch.int32).pin_memory()
        a_np = a_host.numpy()
        carry = [self.decoder.new_carry(B, dev), self.decoder.new_carry(B, dev)]
        hb = [h_t.contiguous(), torch.empty_like(h_t)]
        cb = [c_t.contiguous(), torch.empty_like(c_t)]
        ub = [self.decoder.u_begin.expand(B, -1).contiguous(), torch.empty(B, E, device=dev)]
        L = ctx.shape[1]
        ws = ops.follower_workspace({k: v for k, v in self.decoder.named_parameters()}, B, L, A_cap, dev)
        alpha = torch.empty(B, L, device=dev)
        alpha_v = torch.empty(B, self.decoder.feature_store.feat_table.shape[1], device=dev)
        step_scores, step_ce, n_keeps = [], [], []
        traj = [{"instr_id": ob["instr_id"], "trajectory": [path_element_from_observation(ob)], "actions": [],
                 "scores": [], "observations": [ob], "instr_encoding": ob["instr_encoding"]} for ob in obs]
        ended = np.zeros(B, dtype=bool)
        end_step = np.full(B, -1)
        stream = torch.cuda.current_stream(dev)
        for t in range(self.episode_len):
            A = max(len(ob["adj_loc_list"]) for ob in obs)
            if A > A_cap:
                raise RuntimeError("more than %d action candidates" % A_cap)
            cv = hi[2 * B:2 * B + B * A].reshape(B, A)
            tg = hi[2 * B + B * A_cap:]
            trig = hf[:4 * B * A].reshape(B, A, 4)
            valid = hf[4 * B * A_cap:4 * B * A_cap + B * A].reshape(B, A)
            cv.fill(-1); trig.fill(0.0); valid.fill(0.0)
            for i, ob in enumerate(obs):
                hi[i] = ob["vp_index"]; hi[B + i] = ob["viewIndex"]
                tg[i] = -1 if ended[i] else int(ob.get("teacher", 0))
                adj = ob["adj_loc_list"]
                valid[i, :len(adj)] = 1.0
                for a in range(1, len(adj)):                       # env.py:60-75: row 0 (stop) stays zero
                    d = adj[a]
                    cv[i, a] = d["absViewIndex"]
                    rh, re = d["rel_heading"], d["rel_elevation"]
                    trig[i, a, 0] = np.sin(rh); trig[i, a, 1] = np.cos(rh); trig[i, a, 2] = np.sin(re); trig[i, a, 3] = np.cos(re)
            if feedback == "sample":
                hf[5 * B * A_cap:] = self._sample_uniform(B, dev).cpu().numpy()
            a_np.fill(-1)
            stage.copy_(host, non_blocking=True)                   # the one H2D of the step
            p = t % 2
            d_cv = di[2 * B:2 * B + B * A].view(B, A)
            d_trig = df[:4 * B * A].view(B, A, 4)
            d_valid = df[4 * B * A_cap:4 * B * A_cap + B * A].view(B, A)
            logit = torch.empty(B, A, device=dev)
            score = torch.empty(B, device=dev)
            ce = torch.empty(B, device=dev)
            tail = {"is_valid": d_valid, "feedback": feedback, "target": di[2 * B + B * A_cap:],
                    "sample_u": df[5 * B * A_cap:] if feedback == "sample" else None, "out": (a_host, ub[p ^ 1], score, ce)}
            self.decoder.decode_step(ub[p], None, (di[:B], di[B:2 * B]), hb[p], cb[p], ctx, seq_mask, tail=tail,
                                     carry_in=None if t == 0 else carry[p], carry_out=carry[p ^ 1], ctx_proj=ctx_proj,
                                     cand_view=d_cv, cand_trig=d_trig, out=(hb[p ^ 1], cb[p ^ 1], alpha, logit, alpha_v),
                                     workspace=ws)
            step_scores.append(score); step_ce.append(ce)
            n_keeps.append(int((~ended).sum()))
            stream.synchronize()                                   # a_t is in host memory: the simulator can move
            env_action = a_np.tolist()
            world_states = self.env.step(world_states, env_action, obs)
            obs = self.env.observe(world_states)
            for i, ob in enumerate(obs):                           # follower.py:518-530
                if not ended[i]:
                    traj[i]["trajectory"].append(path_element_from_observation(ob))
                    traj[i]["actions"].append(env_action[i])
                    traj[i]["observations"].append(ob)
                    end_step[i] = t
                if env_action[i] == 0:
                    ended[i] = True
            if ended.all():
                break
        sc = torch.stack(step_scores, 1).cpu().numpy()             # [B, T]: one read-back for the whole rollout
        cum = np.cumsum(sc.astype(np.float32), axis=1, dtype=np.float32)
        for i in range(B):
            e = int(end_step[i])
            traj[i]["scores"] = [float(x) for x in sc[i, :e + 1]]
            traj[i]["score"] = float(cum[i, e]) if e >= 0 else 0.0
        ces = torch.stack(step_ce, 1)
        loss = torch.zeros((), device=dev)
        for t, n in enumerate(n_keeps):
            loss = loss + (ces[:, t].sum() / n if n > 0 else ces[:, t].sum() * float("nan"))
        self.loss = loss
        self.losses.append(float(loss))
        return traj

    def device_rollout(self, nav, start_states, goals, instr_encodings, feedback="argmax", cuda_graph=False):
        """The whole rollout (follower.py:430-539, inference) on the device: the environment is a DeviceNavTables
        (SURVEY.md §8 f-2), so a step is [sfb_nav_step -> fused gather + LSTM -> fused text side + tail] with NO host
        round trip; actions and scores come back once, at the end.  ``start_states`` / ``goals``: int lists (state ids,
        goal viewpoints); ``instr_encodings``: token lists, already sorted by length like env.reset(sort=True) does.
        cuda_graph=True captures the episode's steps into one CUDA graph (returned for replay with new start states
        written into the returned buffers).  Returns dict(actions [B,T] int32 (-1 after the end), scores [B,T], loss)."""
        dev = _device(self.decoder)
        B, T, A = len(start_states), self.episode_len, nav.A
        seq, seq_mask, seq_lengths = batch_instructions_from_encoded(instr_encodings, self.max_instruction_length,
                                                                     reverse=self.reverse_instruction, device=dev)
        sd = {k: v for k, v in self.decoder.named_parameters()}
        E = self.decoder.embedding_size
        buf = {"state": torch.tensor(start_states, dtype=torch.int32, device=dev), "state0": None,
               "ended": torch.zeros(B, dtype=torch.int32, device=dev),
               "goal": torch.tensor(goals, dtype=torch.int32, device=dev),
               "vp_idx": torch.empty(B, dtype=torch.int32, device=dev), "view_idx": torch.empty(B, dtype=torch.int32, device=dev),
               "cand_view": torch.empty(B, A, dtype=torch.int32, device=dev), "cand_trig": torch.empty(B, A, 4, device=dev),
               "is_valid": torch.empty(B, A, device=dev), "target": torch.empty(B, dtype=torch.int32, device=dev),
               "a_t": torch.zeros(B, dtype=torch.int32, device=dev), "actions": torch.full((T, B), -1, dtype=torch.int32, device=dev),
               "scores": torch.zeros(T, B, device=dev), "ce": torch.zeros(T, B, device=dev)}
        buf["state0"] = buf["state"].clone()
        with torch.no_grad():
            ctx, h0, c0 = self.encoder(seq, seq_lengths)
            ctx_proj = self.decoder.project_ctx(ctx)
            L = ctx.shape[1]
            ws = ops.follower_workspace(sd, B, L, A, dev)
            carry = [self.decoder.new_carry(B, dev), self.decoder.new_carry(B, dev)]
            hb, cb = [h0.contiguous(), torch.empty_like(h0)], [c0.contiguous(), torch.empty_like(c0)]
            ub = [self.decoder.u_begin.expand(B, -1).contiguous(), torch.empty(B, E, device=dev)]
            h_init, c_init = hb[0].clone(), cb[0].clone()
            alpha = torch.empty(B, L, device=dev)
            alpha_v = torch.empty(B, self.decoder.feature_store.feat_table.shape[1], device=dev)
            logit = torch.empty(B, A, device=dev)

            def episode():
                buf["state"].copy_(buf["state0"]); buf["ended"].zero_()
                hb[0].copy_(h_init); cb[0].copy_(c_init); ub[0].zero_()
                for t in range(T):
                    ops.nav_step(nav, buf["state"], buf["ended"], buf["goal"], buf["a_t"] if t > 0 else None,
                                 buf["actions"][t - 1] if t > 0 else None, buf)
                    p = t % 2
                    tail = {"is_valid": buf["is_valid"], "feedback": feedback, "target": buf["target"],
                            "out": (buf["a_t"], ub[p ^ 1], buf["scores"][t], buf["ce"][t])}
                    self.decoder.decode_step(ub[p], None, (buf["vp_idx"], buf["view_idx"]), hb[p], cb[p], ctx, seq_mask, tail=tail,
                                             carry_in=None if t == 0 else carry[p], carry_out=carry[p ^ 1], ctx_proj=ctx_proj,
                                             cand_view=buf["cand_view"], cand_trig=buf["cand_trig"],
                                             out=(hb[p ^ 1], cb[p ^ 1], alpha, logit, alpha_v), workspace=ws, idx_dependent=True)
                ops.nav_step(nav, buf["state"], buf["ended"], buf["goal"], buf["a_t"], buf["actions"][T - 1], buf)

            graph = None
            if cuda_graph:
                side = torch.cuda.Stream(device=dev)
                with torch.cuda.stream(side):
                    episode()
                torch.cuda.current_stream(dev).wait_stream(side)
                torch.cuda.synchronize(dev)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    episode()
                graph.replay()
            else:
                episode()
        live = buf["actions"] >= 0
        return {"actions": buf["actions"].t(), "scores": torch.where(live, buf["scores"], torch.zeros_like(buf["scores"])).t(),
                "ce": buf["ce"].t(), "graph": graph, "buffers": buf, "final_state": buf["state"]}

    def _can_stage(self, obs):
        return (not torch.is_grad_enabled() and getattr(self.decoder, "supports_fused_step", False)
                and getattr(self.decoder, "feature_store", None) is not None and not self.decoder.training
                and all("vp_index" in ob for ob in obs))

    def _rollout_with_loss(self):
        """follower.py:430-539."""
        world_states = self.env.reset(sort=True)
        obs = self.env.observe(world_states)
        if self._can_stage(obs):
            return self._staged_rollout(world_states, obs)
        batch_size = len(obs)
        seq, seq_mask, seq_lengths = self._proc_batch(obs)
        dev = _device(self.decoder)
        self.loss = torch.zeros((), device=dev)
        feedback = self.feedback
        ctx, h_t, c_t = self.encoder(seq, seq_lengths)
        traj = [{"instr_id": ob["instr_id"], "trajectory": [path_element_from_observation(ob)], "actions": [],
                 "scores": [], "observations": [ob], "instr_encoding": ob["instr_encoding"]} for ob in obs]
        u_t_prev = self.decoder.u_begin.expand(batch_size, -1)
        ended = np.zeros(batch_size, dtype=bool)
        sequence_scores = torch.zeros(batch_size, device=dev)
        carry = {}   # visual query of the next step, produced by the current one (packed path)
        for t in range(self.episode_len):
            target = self._teacher_action(obs, ended)
            h_t, c_t, alpha, logit, a_t, u_t_prev, score, ce = self._step(u_t_prev, obs, h_t, c_t, ctx, seq_mask, target, feedback, carry)
            n_keep = int((target >= 0).sum())
            self.loss = self.loss + (ce.sum() / n_keep if n_keep > 0 else ce.sum() * float("nan"))   # CE(ignore_index=-1), 278,481
            sequence_scores = sequence_scores + score
            env_action = a_t.tolist()                                     # the one device->host read of the step
            scores_host = score.tolist()
            world_states = self.env.step(world_states, env_action, obs)
            obs = self.env.observe(world_states)
            seq_host = sequence_scores.tolist()
            for i, ob in enumerate(obs):                                  # follower.py:518-530
                if not ended[i]:
                    traj[i]["trajectory"].append(path_element_from_observation(ob))
                    traj[i]["score"] = seq_host[i]
                    traj[i]["scores"].append(scores_host[i])
                    traj[i]["actions"].append(env_action[i])
                    traj[i]["observations"].append(ob)
                if env_action[i] == 0:
                    ended[i] = True
            if ended.all():
                break
        self.losses.append(float(self.loss.detach()) if torch.is_tensor(self.loss) else float(self.loss))
        return traj

    def _score_obs_actions_and_instructions(self, path_obs, path_actions, encoded_instructions):
        """follower.py:342-428: teacher-forced scoring of given (observations, actions, instruction) triples."""
        batch_size = len(path_obs)
        assert len(path_actions) == batch_size and len(encoded_instructions) == batch_size
        for o, a in zip(path_obs, path_actions):
            assert len(o) == len(a) + 1
        dev = _device(self.decoder)
        seq, seq_mask, seq_lengths, perm = batch_instructions_from_encoded(
            encoded_instructions, self.max_instruction_length, reverse=self.reverse_instruction, sort=True, device=dev)
        loss = torch.zeros((), device=dev)
        ctx, h_t, c_t = self.encoder(seq, seq_lengths)
        u_t_prev = self.decoder.u_begin.expand(batch_size, -1)
        ended = np.zeros(batch_size, dtype=bool)
        sequence_scores = torch.zeros(batch_size, device=dev)
        traj = [{"instr_id": o[0]["instr_id"], "trajectory": [path_element_from_observation(o[0])], "actions": [],
                 "scores": [], "observations": [o[0]], "instr_encoding": o[0]["instr_encoding"]} for o in path_obs]
        obs = None
        carry = {}
        for t in range(self.episode_len):
            nxt_obs, nxt_tgt = [], []
            for pi, src in enumerate(perm):
                if t < len(path_actions[src]):
                    nxt_tgt.append(int(path_actions[src][t]))
                    nxt_obs.append(path_obs[src][t])
                else:
                    nxt_tgt.append(-1)
                    nxt_obs.append(obs[pi])
            obs = nxt_obs
            target = torch.tensor(nxt_tgt, dtype=torch.int32, device=dev)
            h_t, c_t, alpha, logit, a_t, u_t_prev, score, ce = self._step(u_t_prev, obs, h_t, c_t, ctx, seq_mask, target, "teacher", carry)
            n_keep = int((target >= 0).sum())
            loss = loss + (ce.sum() / n_keep if n_keep > 0 else 0.0)
            # reference: action_scores = -CE(logit, target, ignore_index=-1) -> 0 for finished rows (405)
            step_scores = torch.where(target >= 0, score, torch.zeros_like(score))
            sequence_scores = sequence_scores + step_scores
            a_host, s_host, seq_host = a_t.tolist(), step_scores.tolist(), sequence_scores.tolist()
            for pi, src in enumerate(perm):
                if not ended[pi]:
                    traj[src]["trajectory"].append(path_element_from_observation(obs[pi]))
                    traj[src]["score"] = seq_host[pi]
                    traj[src]["scores"].append(s_host[pi])
                    traj[src]["actions"].append(a_host[pi])
                if a_host[pi] == 0:
                    ended[pi] = True
            if ended.all():
                break
        return traj, loss

    # ---------------------------------------------------------------- beam search (follower.py:541-718)
    def beam_search(self, beam_size, load_next_minibatch=True, mask_undo=False):
        assert self.env.beam_size >= beam_size
        world_states = self.env.reset(sort=True, beamed=True, load_next_minibatch=load_next_minibatch)
        obs = self.env.observe(world_states, beamed=True)
        batch_size = len(world_states)
        seq, seq_mask, seq_lengths = self._proc_batch(obs, beamed=True)
        ctx, h_t, c_t = self.encoder(seq, seq_lengths)
        dev = _device(self.decoder)
        completed = [[] for _ in range(batch_size)]
        beams = [[InferenceState(None, ws[0], o[0], i, -1, self.decoder.u_begin.view(-1), 0, 0.0, None, None, None)]
                 for i, (ws, o) in enumerate(zip(world_states, obs))]
        for t in range(self.episode_len):
            flat_states = [s for beam in beams for s in beam]
            beam_of = [bi for bi, beam in enumerate(beams) for _ in beam]
            flat_obs = [s.observation for s in flat_states]
            flat_idx = torch.tensor([s.flat_index for s in flat_states], dtype=torch.long, device=dev)
            beam_idx = torch.tensor(beam_of, dtype=torch.long, device=dev)
            u_t_prev = torch.stack([s.last_action_embedding for s in flat_states], 0).contiguous()
            f_t = self._feature_variables(flat_obs)[0]
            all_u_t, is_valid, is_valid_np = self._action_variable(flat_obs)
            h_t, c_t, alpha, logit, alpha_v = self.decoder(u_t_prev, all_u_t, f_t, h_t[flat_idx].contiguous(),
                                                           c_t[flat_idx].contiguous(), ctx[beam_idx].contiguous(),
                                                           seq_mask[beam_idx].contiguous())
            logit = logit.masked_fill(is_valid == 0, -float("inf"))                       # 600
            log_probs = torch.log_softmax(logit, dim=1)
            k = min(beam_size, logit.shape[1])
            _, action_indices = logit.topk(k, dim=1)                                      # ranks on masked LOGITS (608)
            action_scores = log_probs.gather(1, action_indices)
            idx_h, sc_h = action_indices.tolist(), action_scores.tolist()
            all_successors, start = [], 0
            for bi, beam in enumerate(beams):
                succ = []
                for j, st in enumerate(beam):
                    fi = start + j
                    for sc, ai in zip(sc_h[fi], idx_h[fi]):
                        if is_valid_np[fi, ai] == 0:
                            continue
                        succ.append(InferenceState(st, st.world_state, st.observation, fi, ai, all_u_t[fi, ai],
                                                   st.action_count + 1, float(st.score + sc), None, None, alpha[fi]))
                start += len(beam)
                succ.sort(key=lambda s: s.score, reverse=True)
                all_successors.append(succ[:beam_size])
            new_ws = self.env.step([[s.world_state for s in ss] for ss in all_successors],
                                   [[s.last_action for s in ss] for ss in all_successors],
                                   [[s.observation for s in ss] for ss in all_successors], beamed=True)
            new_obs = self.env.observe(new_ws, beamed=True)
            all_successors = [[s._replace(world_state=w, observation=o) for s, w, o in zip(ss, ws_, os_)]
                              for ss, ws_, os_ in zip(all_successors, new_ws, new_obs)]
            beams = []
            for bi, ss in enumerate(all_successors):
                nb = []
                for s in ss:
                    (completed[bi] if (s.last_action == 0 or t == self.episode_len - 1) else nb).append(s)
                beams.append([] if len(completed[bi]) >= beam_size else nb)
            if not any(beams):
                break
        trajs = []
        for done in completed:
            assert done
            out = []
            for s in sorted(done, key=lambda s: s.score, reverse=True)[:beam_size]:
                states, observations, actions, scores, attentions = backchain_inference_states(s)
                out.append({"instr_id": observations[0]["instr_id"], "instr_encoding": observations[0]["instr_encoding"],
                            "trajectory": [path_element_from_observation(o) for o in observations],
                            "observations": observations, "actions": actions, "score": s.score, "scores": scores,
                            "attentions": attentions})
            trajs.append(out)
        return trajs, completed, None

    def state_factored_search(self, completion_size, successor_size, load_next_minibatch=True, mask_undo=False,
                              first_n_ws_key=4, _pad_batch=False):
        """follower.py:720-980.  Search over WORLD STATES instead of action sequences: per instance a table keeps the
        best-scoring inference state reaching each world-state key (scan, viewpoint, heading, elevation); every
        iteration the `successor_size` best not-yet-expanded entries (open or finished) are expanded; finished ones move
        to `completed` until `completion_size` distinct end states exist.  Returns (trajs, completed_list,
        traversed_lists) exactly like the reference; traversed_lists = physical states visited in expansion order."""
        import heapq
        assert self.env.beam_size >= successor_size
        world_states = self.env.reset(sort=True, beamed=True, load_next_minibatch=load_next_minibatch)
        initial_obs = self.env.observe(world_states, beamed=True)
        n_inst = len(world_states)
        seq, seq_mask, seq_lengths = self._proc_batch(initial_obs, beamed=True)
        ctx, h_t, c_t = self.encoder(seq, seq_lengths)
        dev = _device(self.decoder)

        def ws_key(ws):
            return tuple(ws[0:first_n_ws_key])

        # index mode: observations carry (viewpoint row, view index) and the decoder owns a device feature store — no slab
        # and no action-embedding row ever exists on the host; a state's last action embedding is kept as
        # (viewpoint row, view index of the chosen direction, 4 angle terms) and assembled on the device when needed
        store = getattr(self.decoder, "feature_store", None)
        index_mode = store is not None and all("vp_index" in o[0] for o in initial_obs) and \
            getattr(self.decoder, "supports_fused_step", False)
        NO_ACTION = (0, -1, (0.0, 0.0, 0.0, 0.0))

        completed = [dict() for _ in range(n_inst)]      # key -> finished inference state that has been expanded
        holding = [dict() for _ in range(n_inst)]        # key -> [finished inference state, expanded?]
        cache, beams = [], []                            # key -> [open inference state, expanded?]
        for i, (ws, o) in enumerate(zip(world_states, initial_obs)):
            root = InferenceState(None, ws[0], o[0], None, -1, NO_ACTION if index_mode else self.decoder.u_begin.view(-1), 0,
                                  np.float32(0.0), h_t[i], c_t[i], None)
            cache.append({ws_key(ws[0]): [root, True]})
            beams.append([root])
        # Selection of the best not-yet-expanded entry (follower.py:903-908, heapq.nlargest over cache + holding) as a lazy
        # max-heap per instance: an entry is pushed when it is inserted / replaced; stale entries (replaced since, or
        # expanded) are skipped when popped.  Ties resolve like nlargest over [cache entries..., holding entries...] in
        # dict order: open before finished, then first insertion of the key first.
        heaps = [[] for _ in range(n_inst)]
        first_seen = [dict() for _ in range(n_inst)]     # (finished?, key) -> order of first insertion
        use_heap = successor_size == 1
        last_expanded = [beam[0] for beam in beams]
        traversed_lists = [[beam[0]] for beam in beams]

        def extend_traversed(groups):                    # follower.py:764-779
            for i, group in enumerate(groups):
                cur = last_expanded[i]
                for st in group:
                    walk = least_common_viewpoint_path(cur, st)
                    assert walk[0].world_state.viewpointId == cur.world_state.viewpointId
                    assert walk[-1].world_state.viewpointId == st.world_state.viewpointId
                    traversed_lists[i].extend(walk[1:])
                    cur = st
                last_expanded[i] = cur

        while any(len(c) < completion_size for c in completed):
            if _pad_batch:
                # test hook (successor_size 1): every instance contributes a row every iteration (its root when it has no
                # state to expand; the row's results are dropped), so that the decode step sees the same batch as the
                # device search and the two can be compared bit for bit
                assert successor_size == 1
                real = [bool(beam) for beam in beams]
                beams = [beam if beam else [traversed_lists[i][0]] for i, beam in enumerate(beams)]
            flat = [st for beam in beams for st in beam]
            owner = [bi for bi, beam in enumerate(beams) for _ in beam]
            flat_obs = [st.observation for st in flat]
            own_t = torch.tensor(owner, dtype=torch.long, device=dev)
            h_in = torch.stack([st.h_t for st in flat], 0).contiguous()
            c_in = torch.stack([st.c_t for st in flat], 0).contiguous()
            f_t = self._feature_variables(flat_obs)[0]
            if index_mode:
                lae = [st.last_action_embedding for st in flat]
                u_prev = self._embed_actions(torch.tensor([x[0] for x in lae], dtype=torch.int32, device=dev),
                                             torch.tensor([x[1] for x in lae], dtype=torch.int32, device=dev),
                                             torch.tensor([x[2] for x in lae], dtype=torch.float32, device=dev))
                cand_view, cand_trig, is_valid, is_valid_np, cv_np, tr_np = self._action_indices(flat_obs)
                h_new, c_new, alpha, logit, _ = self.decoder.decode_step(u_prev, None, f_t, h_in, c_in, ctx[own_t].contiguous(),
                                                                        seq_mask[own_t].contiguous(), cand_view=cand_view,
                                                                        cand_trig=cand_trig)
            else:
                u_prev = torch.stack([st.last_action_embedding for st in flat], 0).contiguous()
                all_u_t, is_valid, is_valid_np = self._action_variable(flat_obs)
                h_new, c_new, alpha, logit, _ = self.decoder(u_prev, all_u_t, f_t, h_in, c_in, ctx[own_t].contiguous(),
                                                             seq_mask[own_t].contiguous())
            logit = logit.masked_fill(is_valid == 0, -float("inf"))                       # 808
            lp_host = torch.log_softmax(logit, dim=1).cpu().numpy()                       # the one D2H of the iteration

            # every valid action of every beam state is a successor (follower.py:832-857), best first
            all_succ, fi = [], 0
            for bi_, beam in enumerate(beams):
                succ = []
                for st in beam:
                    for ai in range(lp_host.shape[1]):
                        if is_valid_np[fi, ai] == 0 or (_pad_batch and not real[bi_]):
                            continue
                        emb = (flat_obs[fi]["vp_index"], int(cv_np[fi, ai]), tuple(float(v) for v in tr_np[fi, ai])) if index_mode \
                            else all_u_t[fi, ai]
                        succ.append(InferenceState(st, st.world_state, flat_obs[fi], None, ai, emb,
                                                   st.action_count + 1, np.float32(st.score + lp_host[fi, ai]),
                                                   h_new[fi], c_new[fi], alpha[fi]))
                    fi += 1
                succ.sort(key=lambda x: x.score, reverse=True)
                all_succ.append(succ)
            new_ws = self.env.step([[x.world_state for x in ss] for ss in all_succ],
                                   [[x.last_action for x in ss] for ss in all_succ],
                                   [[x.observation for x in ss] for ss in all_succ], beamed=True)
            all_succ = [[x._replace(world_state=w) for x, w in zip(ss, ws_)] for ss, ws_ in zip(all_succ, new_ws)]

            new_beams = []
            for i, succ in enumerate(all_succ):
                if len(completed[i]) >= completion_size:                                  # 889-891
                    new_beams.append([])
                    continue
                for x in succ:                                                            # keep the best per world state
                    k = ws_key(x.world_state)
                    fin = x.last_action == 0 or x.action_count == self.episode_len
                    table = holding[i] if fin else cache[i]
                    if k not in table or table[k][0].score < x.score:
                        table[k] = [x, False]
                        if use_heap:
                            order = first_seen[i].setdefault((fin, k), len(first_seen[i]))
                            heapq.heappush(heaps[i], (-float(x.score), fin, order, id(x), k, x))
                if use_heap:
                    chosen = []
                    while heaps[i]:
                        _, fin, _, _, k, x = heaps[i][0]
                        entry = (holding[i] if fin else cache[i]).get(k)
                        if entry is None or entry[0] is not x or entry[1]:
                            heapq.heappop(heaps[i])                                       # stale: replaced or already expanded
                            continue
                        heapq.heappop(heaps[i])
                        chosen.append((k, x, fin))
                        break
                else:
                    pool = [(k, e[0], False) for k, e in cache[i].items() if not e[1]] + \
                           [(k, e[0], True) for k, e in holding[i].items() if not e[1]]
                    chosen = heapq.nlargest(successor_size, pool, key=lambda t: t[1].score)
                beam = []
                for k, st, finished in chosen:
                    if finished:
                        holding[i][k][1] = True
                        if k not in completed[i] or completed[i][k].score < st.score:
                            completed[i][k] = st
                    else:
                        cache[i][k][1] = True
                        beam.append(st)
                new_beams.append([] if len(completed[i]) >= completion_size else beam)
            beams = new_beams
            if not any(beams):
                break
            new_obs = self.env.observe([[st.world_state for st in beam] for beam in beams], beamed=True)
            beams = [[st._replace(observation=o) for st, o in zip(beam, os_)] for beam, os_ in zip(beams, new_obs)]
            extend_traversed(beams)

        completed_list = [sorted(c.values(), key=lambda x: x.score, reverse=True)[:completion_size] for c in completed]
        final_obs = self.env.observe([[st.world_state for st in cl] for cl in completed_list], beamed=True)
        completed_list = [[st._replace(observation=o) for st, o in zip(cl, os_)] for cl, os_ in zip(completed_list, final_obs)]
        extend_traversed(completed_list)
        trajs = []
        for cl in completed_list:
            assert cl
            out = []
            for st in cl:
                states, observations, actions, scores, attentions = backchain_inference_states(st)
                out.append({"instr_id": observations[0]["instr_id"], "instr_encoding": observations[0]["instr_encoding"],
                            "trajectory": [path_element_from_observation(o) for o in observations],
                            "observations": observations, "actions": actions, "score": st.score, "scores": scores,
                            "attentions": attentions})
            trajs.append(out)
        return trajs, completed_list, traversed_lists

    def device_state_factored_search(self, nav, completion_size, load_next_minibatch=True, max_iter=1024, check_every=8,
                                     use_ctx_proj=True, cuda_graph=False):
        """state_factored_search(completion_size, successor_size=1) — the configuration of pragmatic inference
        (rational_follower.py:44-69) — with the search state ON THE DEVICE: `nav` is the environment as
        look-up tables (navgraph_env.DeviceNavTables), the per-instance cache / holding / completed tables, the inference-state
        pool and the selection of the next state to expand live in device arrays (ops.SfSearchState, sfb_sf_search_update),
        every expansion's (h, c, alpha) stays in a device pool.  The host only launches iterations and looks at one flag
        every `check_every` iterations; when the search has ended it downloads the node arrays once and rebuilds the
        reference's result structures (trajectories, completed lists, traversal walk) for the nodes that matter.
        Returns what state_factored_search returns.  `cuda_graph`: replay the iteration from one captured graph (the iteration
        takes its index from the device); measured at 64 instances x 192 iterations it does not pay — capture costs more than
        the ≈25 small launches per iteration it saves, and the rebuild of the result structures dominates the call."""
        assert getattr(self.decoder, "supports_fused_step", False) and getattr(self.decoder, "feature_store", None) is not None
        world_states = self.env.reset(sort=True, beamed=True, load_next_minibatch=load_next_minibatch)
        initial_obs = self.env.observe(world_states, beamed=True)
        B = len(world_states)
        seq, seq_mask, seq_lengths = self._proc_batch(initial_obs, beamed=True)
        ctx, h_t, c_t = self.encoder(seq, seq_lengths)
        dev = _device(self.decoder)
        H, L, A = h_t.shape[1], ctx.shape[1], nav.A
        st = ops.SfSearchState(B, nav.S, torch.tensor(nav.state_ids([ws[0] for ws in world_states]), dtype=torch.int32),
                               max_iter, max_iter * A, dev)
        pool_h = torch.empty(max_iter + 2, B, H, device=dev); pool_c = torch.empty(max_iter + 2, B, H, device=dev)
        pool_alpha = torch.zeros(max_iter + 2, B, L, device=dev)
        pool_h[0].copy_(h_t); pool_c[0].copy_(c_t)
        rows = torch.arange(B, device=dev)
        a_ids = torch.arange(A, device=dev).unsqueeze(0)
        h1 = torch.empty(B, H, device=dev); c1 = torch.empty(B, H, device=dev); alpha1 = torch.empty(B, L, device=dev)
        logit = torch.empty(B, A, device=dev)
        alpha_v = torch.empty(B, self.decoder.feature_store.feat_table.shape[1], device=dev)
        ctx, seq_mask = ctx.contiguous(), seq_mask.contiguous()
        cproj = self.decoder.project_ctx(ctx) if (use_ctx_proj and hasattr(self.decoder, "project_ctx")) else None

        def iteration(sst):
            """One search iteration, device-side only and with static shapes (so that it can be replayed from a CUDA graph):
            gather the selected states' inputs from the node pool + tables, decode, file the outputs under slot iter + 1,
            update the search state."""
            n = sst.beam_node.long().clamp(min=0)                     # instances without a state run a dummy row (ignored)
            s_ = sst.node_state[rows, n].long()
            par = sst.node_parent[rows, n].long().clamp(min=0)
            act = sst.node_action[rows, n].long()
            ps = sst.node_state[rows, par].long()
            slot = sst.node_slot[rows, n].long()
            h0, c0 = pool_h[slot, rows].contiguous(), pool_c[slot, rows].contiguous()
            u_prev = self._embed_actions(nav.vp[ps], torch.where(act > 0, nav.cv[ps, act.clamp(min=0)], torch.full_like(nav.vp[ps], -1)),
                                         nav.trig[ps, act.clamp(min=0)])
            is_valid = (a_ids < nav.nvalid[s_].unsqueeze(1))
            self.decoder.decode_step(u_prev, None, (nav.vp[s_].contiguous(), nav.view[s_].contiguous()), h0, c0, ctx, seq_mask,
                                     cand_view=nav.cv[s_].contiguous(), cand_trig=nav.trig[s_].contiguous(), ctx_proj=cproj,
                                     out=(h1, c1, alpha1, logit, alpha_v))
            dst = (sst.flags[3:4] + 1).long()                         # slot iter + 1 (the device's own iteration count)
            pool_h.index_copy_(0, dst, h1.unsqueeze(0)); pool_c.index_copy_(0, dst, c1.unsqueeze(0))
            pool_alpha.index_copy_(0, dst, alpha1.unsqueeze(0))
            lp = torch.log_softmax(logit.masked_fill(~is_valid, -float("inf")), dim=1).contiguous()   # follower.py:808-810
            ops.sf_search_update(sst, nav, -1, self.episode_len, completion_size, lp)

        start = torch.tensor(nav.state_ids([ws[0] for ws in world_states]), dtype=torch.int32)
        replay = lambda: iteration(st)
        if cuda_graph:
            iteration(ops.SfSearchState(B, nav.S, start, 2, 2 * A, dev))   # warm-up on a scratch state (allocations, kernel attributes)
            torch.cuda.synchronize()
            gph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gph):
                iteration(st)
            replay = gph.replay
        for t in range(max_iter):
            replay()
            if (t + 1) % check_every == 0 and int(st.flags[0]) != 0:  # the one host look at the device every few iterations
                break
        flags = st.flags.tolist()
        if flags[2]:
            raise RuntimeError("device_state_factored_search: node pool exhausted (max_iter=%d)" % max_iter)
        if not flags[0]:
            raise RuntimeError("device_state_factored_search: not finished after %d iterations" % max_iter)
        T = flags[3]
        # ---- one download, then the reference's result structures for the nodes that matter
        parent, state, action = st.node_parent.cpu().numpy(), st.node_state.cpu().numpy(), st.node_action.cpu().numpy()
        count, slot_h, score = st.node_count.cpu().numpy(), st.node_slot.cpu().numpy(), st.node_score.cpu().numpy()
        trav = st.trav[:, :T].cpu().numpy()
        d_score, d_node = st.d_score.cpu().numpy(), st.d_node.cpu().numpy()
        expanded = [sorted({0} | {int(m) for m in trav[i] if m >= 0}) for i in range(B)]
        if hasattr(nav, "observe_states"):               # every distinct state observed once (≈10 k expansions share ≈2 k states)
            exp_obs = nav.observe_states(self.env, [[state[i, m] for m in expanded[i]] for i in range(B)])
        else:
            exp_obs = self.env.observe([[nav.world_state(state[i, m]) for m in expanded[i]] for i in range(B)], beamed=True)
        own_obs = [dict(zip(expanded[i], exp_obs[i])) for i in range(B)]
        for i in range(B):
            own_obs[i][0] = initial_obs[i][0]            # the root keeps the observation env.reset produced
        built = [dict() for _ in range(B)]

        def node(i, m):
            """The InferenceState (follower.py:27-30) of node m of instance i, ancestors first."""
            chain, k = [], m
            while k >= 0 and k not in built[i]:
                chain.append(k)
                k = int(parent[i, k])
            for k in reversed(chain):
                p_ = int(parent[i, k])
                prev = built[i][p_] if p_ >= 0 else None
                ob = own_obs[i].get(k, prev.observation if prev is not None else None)   # a never-expanded state keeps its parent's
                if p_ < 0:
                    ws_k = world_states[i][0]
                elif int(action[i, k]) == 0:
                    ws_k = prev.world_state                  # the stop action leaves the world state untouched (env.py:628-641)
                else:
                    ws_k = nav.world_state(state[i, k])
                built[i][k] = InferenceState(prev, ws_k, ob, None,
                                             int(action[i, k]), None, int(count[i, k]), np.float32(score[i, k]), None, None,
                                             pool_alpha[int(slot_h[i, k]), i] if p_ >= 0 else None)
            return built[i][m]

        roots = [node(i, 0) for i in range(B)]
        last_expanded = list(roots)
        traversed_lists = [[r] for r in roots]

        def extend_traversed(groups):                    # follower.py:764-779
            for i, group in enumerate(groups):
                cur = last_expanded[i]
                for x in group:
                    walk = least_common_viewpoint_path(cur, x)
                    traversed_lists[i].extend(walk[1:])
                    cur = x
                last_expanded[i] = cur
        for t in range(T):
            extend_traversed([[node(i, int(trav[i, t]))] if trav[i, t] >= 0 else [] for i in range(B)])
        completed_list = []
        for i in range(B):
            done = [node(i, int(d_node[i, s_])) for s_ in np.nonzero(d_score[i] > -np.inf)[0]]
            completed_list.append(sorted(done, key=lambda x: x.score, reverse=True)[:completion_size])
        if hasattr(nav, "observe_states"):
            final_obs = nav.observe_states(self.env, [[nav.state_ids([x.world_state])[0] for x in cl] for cl in completed_list])
            for i, cl in enumerate(completed_list):      # a completed state reached by the stop action at the root keeps the
                for j, x in enumerate(cl):               # root's own world-state floats (heading as env.reset produced it)
                    if x.world_state is world_states[i][0]:
                        final_obs[i][j] = dict(final_obs[i][j], heading=x.world_state.heading, elevation=x.world_state.elevation)
        else:
            final_obs = self.env.observe([[x.world_state for x in cl] for cl in completed_list], beamed=True)
        completed_list = [[x._replace(observation=o) for x, o in zip(cl, os_)] for cl, os_ in zip(completed_list, final_obs)]
        extend_traversed(completed_list)
        trajs = []
        for cl in completed_list:
            assert cl
            out = []
            for x in cl:
                states, observations, actions, scores, attentions = backchain_inference_states(x)
                out.append({"instr_id": observations[0]["instr_id"], "instr_encoding": observations[0]["instr_encoding"],
                            "trajectory": [path_element_from_observation(o) for o in observations],
                            "observations": observations, "actions": actions, "score": x.score, "scores": scores,
                            "attentions": attentions})
            trajs.append(out)
        self.last_search_iterations = T
        return trajs, completed_list, traversed_lists

    # ---------------------------------------------------------------- driver methods (follower.py:982-1035)
    def set_beam_size(self, beam_size):
        if self.env.beam_size < beam_size:
            self.env.set_beam_size(beam_size)
        self.beam_size = beam_size

    def test(self, use_dropout=False, feedback="argmax", allow_cheat=False, beam_size=1):
        if not allow_cheat:
            assert feedback in ["argmax", "sample"]
        self.feedback = feedback
        (self.encoder.train if use_dropout else self.encoder.eval)()
        (self.decoder.train if use_dropout else self.decoder.eval)()
        self.set_beam_size(beam_size)
        with torch.no_grad():
            return super().test()

    def train(self, encoder_optimizer, decoder_optimizer, n_iters, feedback="teacher"):
        assert all(f in self.feedback_options for f in feedback.split("+"))
        self.feedback = feedback
        self.encoder.train()
        self.decoder.train()
        self.losses = []
        for _ in range(1, n_iters + 1):
            encoder_optimizer.zero_grad()
            decoder_optimizer.zero_grad()
            self._rollout_with_loss()
            self.loss.backward()          # gradients: torch autograd over the device-side restatement
            encoder_optimizer.step()
            decoder_optimizer.step()

    def _encoder_and_decoder_paths(self, base_path):
        return base_path + "_enc", base_path + "_dec"

    def save(self, path):
        e, d = self._encoder_and_decoder_paths(path)
        torch.save(self.encoder.state_dict(), e)
        torch.save(self.decoder.state_dict(), d)

    def load(self, path, **kwargs):
        e, d = self._encoder_and_decoder_paths(path)
        self.encoder.load_state_dict(torch.load(e, **kwargs))
        self.decoder.load_state_dict(torch.load(d, **kwargs))
