"""Speaker agent on the sm_100a kernels: the control loops of tasks/R2R/speaker.py:Seq2SeqSpeaker (34-410) with the
same class / method names and result dictionaries.  Path encoder = T x (visual attention + LSTMCell) on the device
(sfb_speaker_encoder_step_fwd), word decoder = sfb_speaker_decoder_step_fwd per word.  Written from the behaviour
described in SURVEY.md A.2 (BOS index 3, instructions NOT reversed, PAD ignored in scores and loss, padded path
steps carry zero features/actions and are masked only in the decoder's attention)."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from collections import namedtuple

from .follower import batch_instructions_from_encoded, vocab_bos_idx, vocab_eos_idx, vocab_pad_idx, _device

InferenceState = namedtuple("InferenceState", "prev_inference_state, flat_index, last_word, word_count, score, last_alpha")   # speaker.py:16


def backchain_inference_states(last):
    """speaker.py:18-32: words / per-word scores / attentions along the parent chain, BOS excluded."""
    chain, x = [], last
    while x is not None:
        chain.append(x)
        x = x.prev_inference_state
    chain.reverse()
    return ([x.last_word for x in chain][1:], [b.score - a.score for a, b in zip(chain[:-1], chain[1:])],
            [x.last_alpha for x in chain][1:])


class Seq2SeqSpeaker(object):
    feedback_options = ["teacher", "argmax", "sample"]

    def __init__(self, env, results_path, encoder, decoder, instruction_len, max_episode_len=10):
        self.env = env
        self.feature_size = getattr(env, "feature_size", None)
        self.results_path = results_path
        self.results = {}
        self.encoder, self.decoder = encoder, decoder
        self.instruction_len = instruction_len
        self.losses = []
        self.max_episode_len = max_episode_len
        self.feedback = "teacher"
        self.loss = 0

    # ---------------------------------------------------------------- batching (speaker.py:68-121)
    def _batch_observations_and_actions(self, path_obs, path_actions, encoded_instructions):
        dev = _device(self.decoder)
        seq_lengths = np.array([len(a) for a in path_actions])
        T = int(seq_lengths.max())
        N = len(path_obs)
        assert N == len(path_actions)
        mask = np.ones((N, T), np.uint8)
        store = getattr(self.encoder, "feature_store", None)
        if store is not None and all("vp_index" in ob for obs in path_obs for ob in obs[:-1]):
            return self._batch_by_index(store, path_obs, path_actions, encoded_instructions, mask, seq_lengths, T, N, dev)
        E = path_obs[0][0]["action_embedding"].shape[-1]
        fshape = path_obs[0][0]["feature"][0].shape
        acts = np.zeros((T, N, E), np.float32)
        feats = np.zeros((T, N) + fshape, np.float32)
        for i, (obs, actions) in enumerate(zip(path_obs, path_actions)):
            assert len(obs) == len(actions) + 1
            mask[i, :len(actions)] = 0
            for t, (ob, a) in enumerate(zip(obs[:-1], actions)):
                assert a >= 0
                feats[t, i] = ob["feature"][0]
                acts[t, i] = ob["action_embedding"][a]
        acts_t = torch.from_numpy(acts).to(dev)
        feats_t = torch.from_numpy(feats).to(dev)
        return ([o[0] for o in path_obs], [feats_t[t] for t in range(T)], [acts_t[t] for t in range(T)],
                torch.from_numpy(mask).to(dev), list(seq_lengths), encoded_instructions, list(range(N)))

    def _batch_by_index(self, store, path_obs, path_actions, encoded_instructions, mask, seq_lengths, T, N, dev):
        """The same batch (speaker.py:68-121) without materialising T x N slabs on the host: the encoder gathers each
        step's 36-view slab from the device-resident feature store by (viewpoint row, view index), and the action
        embeddings (env.py:60-75: slab row of the chosen direction + sin/cos of the relative angles) are assembled on the
        device from (view index, 4 angles).  Padded steps must see zero features and zero actions like the reference's
        zero-initialised arrays (speaker.py:87-95): their rows carry a zero input mask (`step_masks`)."""
        vp = np.zeros((T, N), np.int32); view = np.zeros((T, N), np.int32)
        aview = np.full((T, N), -1, np.int32); trig = np.zeros((T, N, 4), np.float32)
        valid = np.zeros((T, N), np.float32)
        for i, (obs, actions) in enumerate(zip(path_obs, path_actions)):
            assert len(obs) == len(actions) + 1
            n = len(actions)
            mask[i, :n] = 0
            valid[:n, i] = 1.0
            for t in range(n):
                ob, a = obs[t], actions[t]
                assert a >= 0
                vp[t, i] = ob["vp_index"]; view[t, i] = ob["viewIndex"]
                d = ob["adj_loc_list"][a]
                if d["absViewIndex"] >= 0:
                    aview[t, i] = d["absViewIndex"]
                    rh, re = d["rel_heading"], d["rel_elevation"]
                    trig[t, i] = (np.sin(rh), np.cos(rh), np.sin(re), np.cos(re))
        vp_t, view_t = torch.from_numpy(vp).to(dev), torch.from_numpy(view).to(dev)
        aview_t, trig_t, valid_t = torch.from_numpy(aview).to(dev), torch.from_numpy(trig).to(dev), torch.from_numpy(valid).to(dev)
        img = store.img_dim
        loc = store.loc_table.shape[2]
        rows = store.feat_table[vp_t.long(), aview_t.clamp(min=0).long()]                         # [T, N, img]
        has = (aview_t >= 0).unsqueeze(-1).float()
        acts_t = torch.cat((rows * has, (trig_t.repeat_interleave(loc // 4, dim=2)) * has), dim=2).contiguous()   # [T, N, img + loc]
        feats = [(vp_t[t].contiguous(), view_t[t].contiguous()) for t in range(T)]
        E = acts_t.shape[2]
        F_ = img + loc
        self._step_masks = [None if bool(valid[t].all()) else valid_t[t].unsqueeze(1).expand(N, E + F_).contiguous() for t in range(T)]
        return ([o[0] for o in path_obs], feats, [acts_t[t] for t in range(T)],
                torch.from_numpy(mask).to(dev), list(seq_lengths), encoded_instructions, list(range(N)))

    def _encode(self, acts, feats):
        """SpeakerEncoderLSTM over the batch from _batch_observations_and_actions (dense slabs or store indices)."""
        masks = getattr(self, "_step_masks", None)
        self._step_masks = None
        if feats and isinstance(feats[0], tuple):
            return self.encoder(acts, feats, step_masks=masks)
        return self.encoder(acts, feats)

    # ---------------------------------------------------------------- scoring / decoding (speaker.py:123-202)
    def _score_obs_actions_and_instructions(self, path_obs, path_actions, encoded_instructions, feedback):
        assert len(path_obs) == len(path_actions) == len(encoded_instructions)
        start_obs, feats, acts, path_mask, path_lengths, encoded_instructions, perm = \
            self._batch_observations_and_actions(path_obs, path_actions, encoded_instructions)
        dev = _device(self.decoder)
        instr_seq, _, _ = batch_instructions_from_encoded(encoded_instructions, self.instruction_len, device=dev)
        N = len(start_obs)
        ctx, h_t, c_t = self._encode(acts, feats)
        w_t = torch.full((N,), vocab_bos_idx, dtype=torch.long, device=dev)
        ended = np.zeros(N, dtype=bool)
        outputs = [{"instr_id": start_obs[i]["instr_id"], "word_indices": [], "scores": []} for i in range(N)]
        loss = torch.zeros((), device=dev)
        sequence_scores = torch.zeros(N, device=dev)
        rows = torch.arange(N, device=dev)
        for t in range(self.instruction_len):
            h_t, c_t, alpha, logit = self.decoder(w_t.view(-1, 1), h_t, c_t, ctx, path_mask)
            target = instr_seq[:, t].contiguous()
            if feedback == "teacher":
                w_t = target
            elif feedback == "argmax":
                w_t = logit.max(1)[1]
            elif feedback == "sample":
                w_t = torch.multinomial(F.softmax(logit, dim=1), 1).squeeze(1)
            else:
                raise ValueError("Invalid feedback option")
            log_probs = F.log_softmax(logit, dim=1)
            word_scores = log_probs[rows, w_t] * (w_t != vocab_pad_idx)           # -nll_loss(..., ignore_index=PAD) (180)
            sequence_scores = sequence_scores + word_scores
            keep = target != vocab_pad_idx
            if bool(keep.any()):
                loss = loss + (-(log_probs[rows, target]) * keep).sum() / keep.sum()   # mean NLL over non-PAD targets (182)
            w_host, ws_host, seq_host = w_t.tolist(), word_scores.tolist(), sequence_scores.tolist()
            for i in range(N):
                if not ended[i]:
                    outputs[i]["word_indices"].append(int(w_host[i]))
                    outputs[i]["score"] = float(seq_host[i])
                    outputs[i]["scores"].append(ws_host[i])
                if w_host[i] == vocab_eos_idx:
                    ended[i] = True
            if ended.all():
                break
        tok = getattr(self.env, "tokenizer", None)
        for item in outputs:
            item["words"] = tok.decode_sentence(item["word_indices"], break_on_eos=True, join=False) if tok else None
        return outputs, loss

    def rollout(self, load_next_minibatch=True):
        path_obs, path_actions, encoded = self.env.gold_obs_actions_and_instructions(
            self.max_episode_len, load_next_minibatch=load_next_minibatch)
        outputs, loss = self._score_obs_actions_and_instructions(path_obs, path_actions, encoded, self.feedback)
        self.loss = loss
        self.losses.append(float(loss.detach()))
        return outputs

    def beam_search(self, beam_size, path_obs, path_actions):
        """speaker.py:211-318: word-level beam search.  Ranking uses top-k of the logits, scores are the gathered
        log-probabilities; a hypothesis completes on <EOS> or at instruction_len; an instance stops expanding once it
        holds beam_size completions.  Returns, per input path, up to beam_size dicts sorted by score."""
        assert len(path_obs) == len(path_actions)
        start_obs, feats, acts, path_mask, _, _, perm = self._batch_observations_and_actions(path_obs, path_actions, None)
        N = len(start_obs)
        dev = _device(self.decoder)
        ctx, h_t, c_t = self._encode(acts, feats)
        completed = [[] for _ in range(N)]
        beams = [[InferenceState(None, i, vocab_bos_idx, 0, np.float32(0.0), None)] for i in range(N)]
        for t in range(self.instruction_len):
            flat = [st for beam in beams for st in beam]
            owner = torch.tensor([bi for bi, beam in enumerate(beams) for _ in beam], dtype=torch.long, device=dev)
            src = torch.tensor([st.flat_index for st in flat], dtype=torch.long, device=dev)
            w_t = torch.tensor([st.last_word for st in flat], dtype=torch.long, device=dev)
            h_t, c_t, alpha, logit = self.decoder(w_t.view(-1, 1), h_t[src].contiguous(), c_t[src].contiguous(),
                                                  ctx[owner].contiguous(), path_mask[owner].contiguous())
            log_probs = F.log_softmax(logit, dim=1)
            k = min(beam_size, logit.shape[1])
            _, word_indices = logit.topk(k, dim=1)
            word_scores = log_probs.gather(1, word_indices)
            wi_h, ws_h = word_indices.tolist(), word_scores.cpu().numpy()
            new_beams, fi = [], 0
            for bi, beam in enumerate(beams):
                succ = []
                for st in beam:
                    for j in range(k):
                        succ.append(InferenceState(st, fi, wi_h[fi][j], st.word_count + 1,
                                                   np.float32(st.score + ws_h[fi, j]), alpha[fi]))
                    fi += 1
                succ.sort(key=lambda x: x.score, reverse=True)
                nb = []
                for x in succ[:beam_size]:
                    (completed[bi] if (x.last_word == vocab_eos_idx or t == self.instruction_len - 1) else nb).append(x)
                new_beams.append([] if len(completed[bi]) >= beam_size else nb)
            beams = new_beams
            if not any(beams):
                break
        tok = getattr(self.env, "tokenizer", None)
        outputs = [[] for _ in range(N)]
        for pi, si in enumerate(perm):
            for st in sorted(completed[pi], key=lambda x: x.score, reverse=True)[:beam_size]:
                words, scores, attentions = backchain_inference_states(st)
                outputs[si].append({"instr_id": start_obs[pi]["instr_id"], "word_indices": words, "score": st.score,
                                    "scores": scores,
                                    "words": tok.decode_sentence(words, break_on_eos=True, join=False) if tok else None,
                                    "attentions": attentions})
        return outputs

    # ---------------------------------------------------------------- drivers (speaker.py:320-410)
    def test(self, use_dropout=False, feedback="argmax", allow_cheat=False, beam_size=1):
        if not allow_cheat:
            assert feedback in ["argmax", "sample"]
        self.feedback = feedback
        (self.encoder.train if use_dropout else self.encoder.eval)()
        (self.decoder.train if use_dropout else self.decoder.eval)()
        self.env.reset_epoch()
        self.losses = []
        self.results = {}
        looped = False
        with torch.no_grad():
            while True:
                for result in self.rollout():
                    if result["instr_id"] in self.results:
                        looped = True
                    else:
                        self.results[result["instr_id"]] = result
                if looped:
                    break
        return self.results

    def train(self, encoder_optimizer, decoder_optimizer, n_iters, feedback="teacher"):
        assert feedback in self.feedback_options
        self.feedback = feedback
        self.encoder.train()
        self.decoder.train()
        self.losses = []
        for _ in range(1, n_iters + 1):
            encoder_optimizer.zero_grad()
            decoder_optimizer.zero_grad()
            self.rollout()
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
