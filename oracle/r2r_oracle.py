"""CPU oracle for the speaker/follower recurrent hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package ``speaker_follower_b200``; only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may use it, and
only as the checker / the CPU arm being timed.

What this is: a functional (weights-dict in, tensors out) fp32 restatement, on
the CPU, of the arithmetic of ``tasks/R2R/model.py`` and of the per-step tails
of ``tasks/R2R/follower.py`` / ``tasks/R2R/speaker.py`` in
ronghanghu/speaker_follower.  Every function cites the reference lines it
follows.  The arithmetic itself lives in PyTorch (the reference pins
pytorch 0.3.x, ``requirements.txt:72``); this file uses today's torch CPU
kernels for the same operator definitions, so the contract against it is
1e-4 absolute on fp32 logits/states and exact equality on argmax indices.

Parity pin: the reference publishes no golden vectors for this path
(SURVEY.md §8c).  The oracle is instead pinned against the reference's own
``model.py`` executed in the build container: ``tests/golden/make_golden.py``
imports it read-only from /root/reference, runs its modules on seeded inputs
and commits the outputs under ``tests/golden/``; ``tests/test_oracle_golden.py``
checks this file against those vectors on every CPU test run.

Weights are passed as ``dict[str, Tensor]`` keyed by the reference modules'
``state_dict`` names (e.g. ``lstm.weight_ih``), so a reference checkpoint loads
without renaming.  Dropout is deterministic here: callers pass the already
scaled keep-masks (``mask / (1 - p)``) or ``None`` for eval mode.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

Tensor = torch.Tensor
Weights = Dict[str, Tensor]

NEG_INF = -float("inf")


def _cpu(*ts):
    for t in ts:
        if t is not None and isinstance(t, torch.Tensor):
            assert t.device.type == "cpu", "the oracle is a CPU restatement"


def _drop(x: Tensor, keep_scaled: Optional[Tensor]) -> Tensor:
    """nn.Dropout with an injected mask (model.py:370,392,394,414,433,473)."""
    return x if keep_scaled is None else x * keep_scaled


# --------------------------------------------------------------------------
# sub-modules
# --------------------------------------------------------------------------

def visual_soft_dot_attention(h: Tensor, visual_context: Tensor, w: Weights,
                              prefix: str = "visual_attention_layer.") -> Tuple[Tensor, Tensor]:
    """VisualSoftDotAttention.forward — model.py:310-326 (ctor 303-308).

    target = linear_in_h(h); context = linear_in_v(V); attn = softmax_i(context_i . target);
    weighted = sum_i attn_i V_i.  The ``mask`` argument is ignored by the reference.
    """
    _cpu(h, visual_context)
    target = torch.addmm(w[prefix + "linear_in_h.bias"], h, w[prefix + "linear_in_h.weight"].t())
    context = torch.matmul(visual_context, w[prefix + "linear_in_v.weight"].t()) + w[prefix + "linear_in_v.bias"]
    attn = torch.softmax(torch.bmm(context, target.unsqueeze(2)).squeeze(2), dim=1)
    weighted = torch.bmm(attn.unsqueeze(1), visual_context).squeeze(1)
    return weighted, attn


def soft_dot_attention(h: Tensor, context: Tensor, mask: Optional[Tensor], w: Weights,
                       prefix: str) -> Tuple[Tensor, Tensor]:
    """SoftDotAttention.forward — model.py:122-143.

    target = linear_in(h) (no bias); attn = softmax(ctx.target with masked = -inf);
    h_tilde = tanh(linear_out([weighted_ctx ; h])) (no bias, that concat order).
    """
    _cpu(h, context, mask)
    target = h @ w[prefix + "linear_in.weight"].t()
    attn = torch.bmm(context, target.unsqueeze(2)).squeeze(2)
    if mask is not None:
        attn = attn.masked_fill(mask.bool(), NEG_INF)
    attn = torch.softmax(attn, dim=1)
    weighted = torch.bmm(attn.unsqueeze(1), context).squeeze(1)
    h_tilde = torch.tanh(torch.cat((weighted, h), 1) @ w[prefix + "linear_out.weight"].t())
    return h_tilde, attn


def eltwise_prod_scoring(h: Tensor, all_u_t: Tensor, w: Weights,
                         prefix: str = "decoder2action.") -> Tensor:
    """EltwiseProdScoring.forward — model.py:342-352 (``mask`` ignored there)."""
    _cpu(h, all_u_t)
    target = torch.addmm(w[prefix + "linear_in_h.bias"], h, w[prefix + "linear_in_h.weight"].t()).unsqueeze(1)
    context = torch.matmul(all_u_t, w[prefix + "linear_in_a.weight"].t()) + w[prefix + "linear_in_a.bias"]
    eltprod = target * context
    return (torch.matmul(eltprod, w[prefix + "linear_out.weight"].t()) + w[prefix + "linear_out.bias"]).squeeze(2)


def lstm_cell(x: Tensor, h0: Tensor, c0: Tensor, w: Weights, prefix: str = "lstm.") -> Tuple[Tensor, Tensor]:
    """nn.LSTMCell as used at model.py:371,393 / 417,434 / 483,515.

    gates = W_ih x + b_ih + W_hh h + b_hh, chunked (i, f, g, o); two bias vectors.
    """
    _cpu(x, h0, c0)
    gates = (x @ w[prefix + "weight_ih"].t() + w[prefix + "bias_ih"]
             + h0 @ w[prefix + "weight_hh"].t() + w[prefix + "bias_hh"])
    i, f, g, o = gates.chunk(4, dim=1)
    c1 = torch.sigmoid(f) * c0 + torch.sigmoid(i) * torch.tanh(g)
    h1 = torch.sigmoid(o) * torch.tanh(c1)
    return h1, c1


# --------------------------------------------------------------------------
# follower
# --------------------------------------------------------------------------

def attn_decoder_step(u_t_prev: Tensor, all_u_t: Tensor, visual_context: Tensor, h_0: Tensor, c_0: Tensor,
                      ctx: Tensor, ctx_mask: Optional[Tensor], w: Weights,
                      drop_x: Optional[Tensor] = None, drop_h: Optional[Tensor] = None):
    """AttnDecoderLSTM.forward — model.py:377-397: one follower decode step.

    Returns (h_1, c_1, alpha[B,L], logit[B,A], alpha_v[B,36]); h_1/c_1 are returned un-dropped.
    """
    feature, alpha_v = visual_soft_dot_attention(h_0, visual_context, w)
    x = _drop(torch.cat((u_t_prev, feature), 1), drop_x)
    h_1, c_1 = lstm_cell(x, h_0, c_0, w)
    h_tilde, alpha = soft_dot_attention(_drop(h_1, drop_h), ctx, ctx_mask, w, "text_attention_layer.")
    logit = eltwise_prod_scoring(h_tilde, all_u_t, w)
    return h_1, c_1, alpha, logit, alpha_v


def lstm_sequence(x: Tensor, lengths: Sequence[int], w: Weights, suffix: str = "_l0",
                  reverse: bool = False) -> Tuple[Tensor, Tensor, Tensor]:
    """One direction of a 1-layer batch_first nn.LSTM over a length-sorted padded batch
    (what pack_padded_sequence + nn.LSTM + pad_packed_sequence compute, model.py:89-90,101).

    Returns (outputs[B,max_len,H] zero beyond each length, h_T[B,H], c_T[B,H]) where h_T/c_T are
    the states after each row's own last valid token.
    """
    B, T, _ = x.shape
    W_ih, W_hh = w["lstm.weight_ih" + suffix], w["lstm.weight_hh" + suffix]
    b = w["lstm.bias_ih" + suffix] + w["lstm.bias_hh" + suffix]
    H = W_hh.shape[1]
    h = x.new_zeros(B, H)
    c = x.new_zeros(B, H)
    max_len = int(max(lengths))
    out = x.new_zeros(B, max_len, H)
    lens = torch.as_tensor(list(lengths))
    steps = range(max_len - 1, -1, -1) if reverse else range(max_len)
    for t in steps:
        active = (lens > t).unsqueeze(1)
        gates = x[:, t] @ W_ih.t() + h @ W_hh.t() + b
        i, f, g, o = gates.chunk(4, dim=1)
        c_new = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h_new = torch.sigmoid(o) * torch.tanh(c_new)
        h = torch.where(active, h_new, h)
        c = torch.where(active, c_new, c)
        out[:, t] = torch.where(active, h_new, torch.zeros_like(h_new))
    return out, h, c


def encoder_lstm(inputs: Tensor, lengths: Sequence[int], w: Weights, bidirectional: bool = False,
                 use_glove: bool = True, drop_embed: Optional[Tensor] = None,
                 drop_ctx: Optional[Tensor] = None):
    """EncoderLSTM.forward — model.py:81-104.  ``lengths`` must be sorted descending (model.py:89).

    Embedding is not dropped when GloVe is used (86-87); ctx is dropped (102); c_t is returned raw (97).
    """
    _cpu(inputs)
    embeds = w["embedding.weight"][inputs]
    if not use_glove:
        embeds = _drop(embeds, drop_embed)
    out_f, h_f, c_f = lstm_sequence(embeds, lengths, w, "_l0")
    if bidirectional:
        out_b, h_b, c_b = lstm_sequence(embeds, lengths, w, "_l0_reverse", reverse=True)
        ctx = torch.cat((out_f, out_b), 2)
        # enc_h_t[-1] is the reverse direction, enc_h_t[-2] the forward one (model.py:93-94)
        h_t = torch.cat((h_b, h_f), 1)
        c_t = torch.cat((c_b, c_f), 1)
    else:
        ctx, h_t, c_t = out_f, h_f, c_f
    decoder_init = torch.tanh(h_t @ w["encoder2decoder.weight"].t() + w["encoder2decoder.bias"])
    return _drop(ctx, drop_ctx), decoder_init, c_t


def follower_step_tail(logit: Tensor, is_valid: Tensor, target: Optional[Tensor], feedback: str,
                       all_u_t: Tensor, sample_u: Optional[Tensor] = None):
    """Per-step tail of Seq2SeqAgent._rollout_with_loss — follower.py:476-505.

    logit[is_valid==0] = -inf (477); loss term = CrossEntropyLoss(ignore_index=-1)(logit, target)
    (278,481; mean over non-ignored rows); a_t = clamp(target,0) | argmax | sample (484-499);
    u_t_prev = all_u_t[arange, a_t] (502); action_scores = log_softmax(logit)[a_t] (504).
    ``sample_u``: uniform(0,1) numbers per row for feedback='sample' (inverse-CDF draw over
    softmax(logit)*valid, 491-497) so that a device implementation can reproduce the draw.
    Returns (masked_logit, loss_term or None, a_t, u_next, action_scores).
    """
    _cpu(logit, is_valid, all_u_t)
    logit = logit.masked_fill(is_valid == 0, NEG_INF)
    logp = torch.log_softmax(logit, dim=1)
    loss = None
    if target is not None:
        keep = target >= 0
        n = int(keep.sum())
        picked = -logp[torch.arange(logit.shape[0]), target.clamp(min=0)]
        loss = (picked * keep).sum() / n if n > 0 else picked.sum() * float("nan")
    if feedback == "teacher":
        a_t = target.clamp(min=0)
    elif feedback == "argmax":
        a_t = logit.max(1)[1]
    elif feedback == "sample":
        probs = torch.softmax(logit, dim=1) * (is_valid != 0)
        cdf = torch.cumsum(probs / probs.sum(1, keepdim=True), dim=1)
        a_t = (sample_u.unsqueeze(1) > cdf).sum(1).clamp(max=logit.shape[1] - 1)
    else:
        raise ValueError("Invalid feedback option")
    rows = torch.arange(logit.shape[0])
    u_next = all_u_t[rows, a_t]
    action_scores = logp[rows, a_t]
    return logit, loss, a_t, u_next, action_scores


def follower_rollout(seq: Tensor, seq_mask: Tensor, lengths: Sequence[int], steps: Sequence[dict],
                     enc_w: Weights, dec_w: Weights, feedback: str = "argmax"):
    """Model-side arithmetic of Seq2SeqAgent._rollout_with_loss — follower.py:430-539 — on a
    pre-recorded sequence of observations.

    ``steps[t]`` = {"visual": [B,36,F], "all_u_t": [B,A_t,E], "is_valid": [B,A_t], "target": [B] or None}.
    (The simulator is off this path; the observations a rollout would have produced are given.)
    Returns per-step dicts with logit, a_t, scores plus the accumulated loss and sequence scores.
    """
    ctx, h_t, c_t = encoder_lstm(seq, lengths, enc_w)
    B = seq.shape[0]
    u_prev = torch.zeros(B, steps[0]["all_u_t"].shape[2])          # decoder.u_begin (model.py:368)
    seq_scores = torch.zeros(B)
    loss = torch.zeros(())
    out = []
    for st in steps:
        h_t, c_t, alpha, logit, alpha_v = attn_decoder_step(
            u_prev, st["all_u_t"], st["visual"], h_t, c_t, ctx, seq_mask, dec_w)
        logit, l, a_t, u_prev, sc = follower_step_tail(
            logit, st["is_valid"], st.get("target"), feedback, st["all_u_t"], st.get("sample_u"))
        if l is not None:
            loss = loss + l
        seq_scores = seq_scores + sc
        out.append({"logit": logit, "a_t": a_t, "scores": sc, "h": h_t, "c": c_t,
                    "alpha": alpha, "alpha_v": alpha_v})
    return out, loss, seq_scores


# --------------------------------------------------------------------------
# speaker
# --------------------------------------------------------------------------

def speaker_encoder_step(h_0, c_0, action_embedding, world_state_embedding, w: Weights,
                         drop_x: Optional[Tensor] = None):
    """SpeakerEncoderLSTM._forward_one_step — model.py:429-435."""
    feature, _ = visual_soft_dot_attention(h_0, world_state_embedding, w)
    x = _drop(torch.cat((action_embedding, feature), 1), drop_x)
    return lstm_cell(x, h_0, c_0, w)


def speaker_encoder(batched_action_embeddings: List[Tensor], world_state_embeddings: List[Tensor], w: Weights,
                    drop_x: Optional[List[Tensor]] = None, drop_ctx: Optional[Tensor] = None):
    """SpeakerEncoderLSTM.forward — model.py:437-457.  Runs all T steps for every row (no
    per-row stop); ctx = drop(stack(h_t)); decoder_init = tanh(encoder2decoder(h_T))."""
    assert len(batched_action_embeddings) == len(world_state_embeddings)
    B = world_state_embeddings[0].shape[0]
    H = w["lstm.weight_hh"].shape[1]
    h = torch.zeros(B, H)
    c = torch.zeros(B, H)
    hs = []
    for t, (a, v) in enumerate(zip(batched_action_embeddings, world_state_embeddings)):
        h, c = speaker_encoder_step(h, c, a, v, w, None if drop_x is None else drop_x[t])
        hs.append(h)
    decoder_init = torch.tanh(h @ w["encoder2decoder.weight"].t() + w["encoder2decoder.bias"])
    ctx = _drop(torch.stack(hs, dim=1), drop_ctx)
    return ctx, decoder_init, c


def speaker_decoder_step(previous_word: Tensor, h_0: Tensor, c_0: Tensor, ctx: Tensor,
                         ctx_mask: Optional[Tensor], w: Weights, use_glove: bool = True,
                         drop_e: Optional[Tensor] = None, drop_h: Optional[Tensor] = None):
    """SpeakerDecoderLSTM.forward, default (no input-att-feed) branch — model.py:497-503,515-519."""
    _cpu(previous_word, h_0, c_0, ctx)
    e = w["embedding.weight"][previous_word.reshape(-1)]
    if not use_glove:
        e = _drop(e, drop_e)
    h_1, c_1 = lstm_cell(e, h_0, c_0, w)
    h_tilde, alpha = soft_dot_attention(_drop(h_1, drop_h), ctx, ctx_mask, w, "attention_layer.")
    logit = h_tilde @ w["decoder2action.weight"].t() + w["decoder2action.bias"]
    return h_1, c_1, alpha, logit


def speaker_score_teacher(action_embs: List[Tensor], feats: List[Tensor], path_mask: Tensor, instr_seq: Tensor,
                          enc_w: Weights, dec_w: Weights, bos_idx: int = 3, pad_idx: int = 0,
                          feedback: str = "teacher"):
    """Model-side arithmetic of Seq2SeqSpeaker._score_obs_actions_and_instructions — speaker.py:123-202
    (teacher / argmax feedback).  BOS index 3 (utils.py:24); PAD ignored in word scores (180) and in
    the mean NLL loss (182).  Runs all ``instr_seq.shape[1]`` steps (no early exit) and returns
    (sequence_scores[B], loss, word_indices[B,T], word_scores[B,T]).
    """
    ctx, h_t, c_t = speaker_encoder(action_embs, feats, enc_w)
    B, T = instr_seq.shape
    w_t = torch.full((B,), bos_idx, dtype=torch.long)
    seq_scores = torch.zeros(B)
    loss = torch.zeros(())
    words, wscores = [], []
    rows = torch.arange(B)
    for t in range(T):
        h_t, c_t, alpha, logit = speaker_decoder_step(w_t, h_t, c_t, ctx, path_mask, dec_w)
        target = instr_seq[:, t]
        w_t = target if feedback == "teacher" else logit.max(1)[1]
        logp = torch.log_softmax(logit, dim=1)
        ws = logp[rows, w_t] * (w_t != pad_idx)
        seq_scores = seq_scores + ws
        keep = target != pad_idx
        if int(keep.sum()) > 0:
            loss = loss + (-(logp[rows, target]) * keep).sum() / keep.sum()
        words.append(w_t)
        wscores.append(ws)
    return seq_scores, loss, torch.stack(words, 1), torch.stack(wscores, 1)


# --------------------------------------------------------------------------
# pragmatic-inference combine
# --------------------------------------------------------------------------

def rational_combine(speaker_scores, follower_scores, groups, weight: float):
    """run_rational_follower candidate combine — rational_follower.py:118-150: global population
    std (np.std, ddof 0) of all speaker / follower scores, then per instruction argmax of
    weight*spk/std_spk + (1-weight)*fol/std_fol.  ``groups[i]`` = instruction id of candidate i.
    Returns {instruction id: index of the best candidate}."""
    import numpy as np
    s = np.asarray(speaker_scores, dtype=np.float64)
    f = np.asarray(follower_scores, dtype=np.float64)
    ss, fs = np.std(s), np.std(f)
    comb = weight * s / ss + (1.0 - weight) * f / fs
    best = {}
    for i, g in enumerate(groups):
        if g not in best or comb[i] > comb[best[g]]:
            best[g] = i
    return best
