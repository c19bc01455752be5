"""Deterministic synthetic weights and inputs for tests, smoke and bench.

Everything is drawn from ``numpy.random.Generator(PCG64(seed))`` so that the golden fixtures made in
the build container (tests/golden/make_golden.py) and the tensors regenerated on the GPU box are the
same bits.  Weight distributions follow the reference modules' default initialisation
(U(-1/sqrt(fan), 1/sqrt(fan)) for nn.Linear / nn.LSTMCell — model.py:61-65,306-307,338-340,371), the
feature statistics follow ResNet pool5 (non-negative, about half zeros) and the orientation part is the
reference's sin/cos table (env.py:78-101).
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import numpy as np
import torch

V_NUM = 36
IMG_DIM = 2048
LOC_DIM = 128
FEAT = IMG_DIM + LOC_DIM          # 2176, train.py:38
HID = 512                         # train.py:33
DOT = 256                         # model.py:303,335
WORD = 300                        # train.py:30
VOCAB = 991                       # tasks/R2R/data/train_vocab.txt


def _rng(seed: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64(seed))


def _uniform(g, shape, bound) -> torch.Tensor:
    return torch.from_numpy(g.uniform(-bound, bound, size=shape).astype(np.float32))


def _linear(g, out_f, in_f, bias=True, prefix="") -> Dict[str, torch.Tensor]:
    b = 1.0 / math.sqrt(in_f)
    d = {prefix + "weight": _uniform(g, (out_f, in_f), b)}
    if bias:
        d[prefix + "bias"] = _uniform(g, (out_f,), b)
    return d


def _lstm_cell(g, in_f, hid, prefix="lstm.", suffix="") -> Dict[str, torch.Tensor]:
    b = 1.0 / math.sqrt(hid)
    return {
        prefix + "weight_ih" + suffix: _uniform(g, (4 * hid, in_f), b),
        prefix + "weight_hh" + suffix: _uniform(g, (4 * hid, hid), b),
        prefix + "bias_ih" + suffix: _uniform(g, (4 * hid,), b),
        prefix + "bias_hh" + suffix: _uniform(g, (4 * hid,), b),
    }


def follower_decoder_weights(seed=11, emb=FEAT, hid=HID, feat=FEAT, dot=DOT) -> Dict[str, torch.Tensor]:
    """state_dict of AttnDecoderLSTM (model.py:361-375), reference key names."""
    g = _rng(seed)
    w = {}
    w.update(_lstm_cell(g, emb + feat, hid))
    w.update(_linear(g, dot, hid, True, "visual_attention_layer.linear_in_h."))
    w.update(_linear(g, dot, feat, True, "visual_attention_layer.linear_in_v."))
    w.update(_linear(g, hid, hid, False, "text_attention_layer.linear_in."))
    w.update(_linear(g, hid, 2 * hid, False, "text_attention_layer.linear_out."))
    w.update(_linear(g, dot, hid, True, "decoder2action.linear_in_h."))
    w.update(_linear(g, dot, emb, True, "decoder2action.linear_in_a."))
    w.update(_linear(g, 1, dot, True, "decoder2action.linear_out."))
    return w


def follower_encoder_weights(seed=12, vocab=VOCAB, word=WORD, hid=HID, bidirectional=False) -> Dict[str, torch.Tensor]:
    """state_dict of EncoderLSTM (model.py:47-65).  Embedding ~ N(0,1)*0.4 (GloVe-like scale), PAD row zero."""
    g = _rng(seed)
    h = hid // 2 if bidirectional else hid
    w = {"embedding.weight": torch.from_numpy((g.standard_normal((vocab, word)) * 0.4).astype(np.float32))}
    w["embedding.weight"][0] = 0
    w.update(_lstm_cell(g, word, h, "lstm.", "_l0"))
    if bidirectional:
        w.update(_lstm_cell(g, word, h, "lstm.", "_l0_reverse"))
    w.update(_linear(g, hid, hid, True, "encoder2decoder."))
    return w


def speaker_encoder_weights(seed=13, emb=FEAT, feat=FEAT, hid=HID, dot=DOT) -> Dict[str, torch.Tensor]:
    """state_dict of SpeakerEncoderLSTM (model.py:406-419)."""
    g = _rng(seed)
    w = {}
    w.update(_linear(g, dot, hid, True, "visual_attention_layer.linear_in_h."))
    w.update(_linear(g, dot, feat, True, "visual_attention_layer.linear_in_v."))
    w.update(_lstm_cell(g, emb + feat, hid))
    w.update(_linear(g, hid, hid, True, "encoder2decoder."))
    return w


def speaker_decoder_weights(seed=14, vocab=VOCAB, word=WORD, hid=HID) -> Dict[str, torch.Tensor]:
    """state_dict of SpeakerDecoderLSTM, default branch (model.py:461-485)."""
    g = _rng(seed)
    w = {"embedding.weight": torch.from_numpy((g.standard_normal((vocab, word)) * 0.4).astype(np.float32))}
    w.update(_lstm_cell(g, word, hid))
    w.update(_linear(g, hid, hid, False, "attention_layer.linear_in."))
    w.update(_linear(g, hid, 2 * hid, False, "attention_layer.linear_out."))
    w.update(_linear(g, vocab, hid, True, "decoder2action."))
    return w


def loc_embedding_table() -> torch.Tensor:
    """[36 (agent viewIndex), 36 (absViewIndex), 128] — env.py:78-101 (build_viewpoint_loc_embedding)."""
    inc = math.pi / 6.0
    t = np.zeros((V_NUM, V_NUM, LOC_DIM), np.float32)
    for view in range(V_NUM):
        for a in range(V_NUM):
            rel = (a - view) % 12 + (a // 12) * 12
            rh = (rel % 12) * inc
            re = (rel // 12 - 1) * inc
            t[view, a, 0:32] = np.sin(rh)
            t[view, a, 32:64] = np.cos(rh)
            t[view, a, 64:96] = np.sin(re)
            t[view, a, 96:] = np.cos(re)
    return torch.from_numpy(t)


def feature_table(n_viewpoints: int, seed=21, img_dim=IMG_DIM) -> torch.Tensor:
    """[n_viewpoints, 36, img_dim] f32, relu(N(0,1))*1.1 — pool5-like."""
    g = _rng(seed)
    x = g.standard_normal((n_viewpoints, V_NUM, img_dim), dtype=np.float32)
    np.maximum(x, 0, out=x)
    x *= 1.1
    return torch.from_numpy(x)


def action_embeddings(table: torch.Tensor, vp: np.ndarray, n_actions: np.ndarray, a_max: int, g) -> Tuple[torch.Tensor, torch.Tensor]:
    """all_u_t [B,a_max,img+128] and is_valid [B,a_max] built like env.py:60-75 + follower.py:300-320:
    row 0 = stop = zeros; row a>=1 = [table[vp, absViewIndex_a], sin/cos(rel_heading), sin/cos(rel_elev)]."""
    B = len(vp)
    img = table.shape[2]
    U = torch.zeros(B, a_max, img + LOC_DIM)
    valid = torch.zeros(B, a_max)
    for b in range(B):
        n = int(n_actions[b])
        valid[b, :n] = 1.0
        for a in range(1, n):
            view = int(g.integers(0, V_NUM))
            rh = float(g.uniform(-math.pi, math.pi))
            re = float(g.uniform(-math.pi / 6, math.pi / 6))
            U[b, a, :img] = table[vp[b], view]
            U[b, a, img:img + 32] = math.sin(rh)
            U[b, a, img + 32:img + 64] = math.cos(rh)
            U[b, a, img + 64:img + 96] = math.sin(re)
            U[b, a, img + 96:] = math.cos(re)
    return U, valid


# empirical #actions (incl. stop) distribution of the R2R nav graphs: mean 5.06, p95 9, max 14
_ACTION_COUNTS = np.arange(2, 15)
_ACTION_P = np.array([6, 14, 22, 20, 14, 9, 6, 4, 2.2, 1.4, 0.8, 0.4, 0.2])
_ACTION_P = _ACTION_P / _ACTION_P.sum()


def follower_step_inputs(B: int, L: int, A: int, seed=31, n_viewpoints=64, table=None, loc=None,
                         fixed_len=False, img_dim=IMG_DIM, hid=HID) -> Dict[str, torch.Tensor]:
    """One decode step's inputs for AttnDecoderLSTM.forward (model.py:377): dict with
    u_t_prev [B,E], all_u_t [B,A,E], is_valid [B,A], visual_context [B,36,F], h_0, c_0 [B,H],
    ctx [B,L,H], ctx_mask [B,L] (bool, True = pad), vp_idx, view_idx [B] (int32) and lengths."""
    g = _rng(seed)
    if table is None:
        table = feature_table(n_viewpoints, seed + 1000, img_dim)
    if loc is None:
        loc = loc_embedding_table()
    n_vp = table.shape[0]
    vp = g.integers(0, n_vp, size=B)
    view = g.integers(0, V_NUM, size=B)
    visual = torch.cat((table[vp], loc[view]), dim=2).contiguous()
    n_act = np.minimum(g.choice(_ACTION_COUNTS, size=B, p=_ACTION_P), A)
    n_act[0] = A
    U, valid = action_embeddings(table, vp, n_act, A, g)
    prev = g.integers(0, n_act)
    u_prev = U[torch.arange(B), torch.from_numpy(prev)].clone()
    h0 = torch.from_numpy((np.tanh(g.standard_normal((B, hid)) * 0.5)).astype(np.float32))
    c0 = torch.from_numpy((g.standard_normal((B, hid)) * 0.5).astype(np.float32))
    ctx = torch.from_numpy(np.tanh(g.standard_normal((B, L, hid)) * 0.6).astype(np.float32))
    if fixed_len:
        lengths = np.full(B, L)
    else:
        lengths = np.sort(g.integers(min(10, L), L + 1, size=B))[::-1].copy()
        lengths[0] = L
    mask = torch.arange(L).unsqueeze(0) >= torch.from_numpy(lengths.copy()).unsqueeze(1)
    ctx = ctx * (~mask).unsqueeze(2)           # pad_packed_sequence leaves zeros beyond each length
    return {
        "u_t_prev": u_prev, "all_u_t": U, "is_valid": valid, "visual_context": visual,
        "h_0": h0, "c_0": c0, "ctx": ctx, "ctx_mask": mask,
        "vp_idx": torch.from_numpy(vp.astype(np.int32)), "view_idx": torch.from_numpy(view.astype(np.int32)),
        "lengths": torch.from_numpy(lengths.astype(np.int64)),
    }


def instruction_batch(B: int, L: int, seed=41, vocab=VOCAB) -> Tuple[torch.Tensor, torch.Tensor, List[int]]:
    """seq [B,L] int64 (tokens >= 4, EOS=2 at the end, PAD=0 after), mask [B,max_len] bool, lengths desc
    — the layout batch_instructions_from_encoded produces (follower.py:75-105)."""
    g = _rng(seed)
    lengths = np.sort(g.integers(min(8, L), L + 1, size=B))[::-1].copy()
    lengths[0] = L
    seq = np.zeros((B, L), np.int64)
    for b in range(B):
        n = int(lengths[b])
        seq[b, :n - 1] = g.integers(4, vocab, size=n - 1)
        seq[b, n - 1] = 2
    seq = torch.from_numpy(seq)
    mask = (seq == 0)[:, :int(lengths.max())]
    return seq, mask, [int(x) for x in lengths]
