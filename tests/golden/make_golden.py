"""Generate the golden vectors under tests/golden/ by running the REFERENCE's own model.py.

Run in the build container only (needs /root/reference, read-only):

    python tests/golden/make_golden.py

The reference (ronghanghu/speaker_follower, tasks/R2R/model.py) is imported unmodified with a stub
``MatterSim`` module (the simulator is off this path and cannot be built here) and
bool masks (torch >= 1.2 rejects the reference's uint8 masks at model.py:135).  Two kinds of cases:

* ``*_small``: reference modules instantiated with small dimensions; weights, inputs and outputs are
  all stored, so the fixture is self-contained.
* ``*_full``: the real dimensions (E=F=2176, H=512, vocab 991); weights and inputs are regenerated
  from ``speaker_follower_b200.synth`` seeds (PCG64, stable), only the reference OUTPUTS are stored.

Dropout placement is pinned by replacing the ``drop`` submodule of a reference *instance* with a
module that multiplies by pre-drawn masks in call order (no reference source is touched).
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/tasks/R2R")
sys.modules["MatterSim"] = types.ModuleType("MatterSim")

import model as ref                       # noqa: E402  the reference
from speaker_follower_b200 import synth   # noqa: E402

torch.set_num_threads(8)


class MaskSeq(nn.Module):
    """Stands in for nn.Dropout on a reference instance: multiplies by given scaled masks in call order."""

    def __init__(self, masks):
        super().__init__()
        self.masks = list(masks)
        self.i = 0

    def forward(self, x):
        m = self.masks[self.i]
        self.i += 1
        return x * m


def np_(d):
    return {k: (v.detach().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in d.items()}


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **np_(arrs))
    print(name, {k: tuple(np.asarray(v).shape) for k, v in np_(arrs).items()}, os.path.getsize(path) // 1024, "KB")


def keep_mask(g, shape, p=0.5):
    return torch.from_numpy((g.random(shape) >= p).astype(np.float32) / (1.0 - p))


def small_inputs(g, B, L, A, E, F, H, V=36):
    t = lambda *s: torch.from_numpy(g.standard_normal(s).astype(np.float32))
    lengths = np.sort(g.integers(2, L + 1, size=B))[::-1].copy()
    lengths[0] = L
    mask = torch.arange(L).unsqueeze(0) >= torch.from_numpy(lengths.copy()).unsqueeze(1)
    return dict(u_t_prev=t(B, E), all_u_t=t(B, A, E), visual_context=torch.relu(t(B, V, F)),
                h_0=torch.tanh(t(B, H)), c_0=t(B, H) * 0.5, ctx=torch.tanh(t(B, L, H)), ctx_mask=mask)


def follower_step_small():
    g = np.random.Generator(np.random.PCG64(101))
    E, F, H, B, L, A = 48, 40, 32, 5, 7, 4
    torch.manual_seed(1)
    dec = ref.AttnDecoderLSTM(E, H, 0.5, feature_size=F).eval()
    x = small_inputs(g, B, L, A, E, F, H)
    with torch.no_grad():
        h1, c1, alpha, logit, alpha_v = dec(x["u_t_prev"], x["all_u_t"], x["visual_context"], x["h_0"], x["c_0"],
                                            x["ctx"], x["ctx_mask"])
    w = {"w." + k: v for k, v in dec.state_dict().items()}
    save("follower_step_small", **w, **{"in." + k: v for k, v in x.items()},
         **{"out.h_1": h1, "out.c_1": c1, "out.alpha": alpha, "out.logit": logit, "out.alpha_v": alpha_v})

    # train mode with pinned dropout masks + gradients of a scalar loss (pins backward too)
    dx, dh = keep_mask(g, (B, E + F)), keep_mask(g, (B, H))
    dec.drop = MaskSeq([dx, dh])
    dec.train()
    leaves = {k: v.clone().requires_grad_(True) for k, v in x.items() if v.dtype == torch.float32}
    h1, c1, alpha, logit, alpha_v = dec(leaves["u_t_prev"], leaves["all_u_t"], leaves["visual_context"],
                                        leaves["h_0"], leaves["c_0"], leaves["ctx"], x["ctx_mask"])
    r = {k: torch.from_numpy(g.standard_normal(tuple(v.shape)).astype(np.float32))
         for k, v in dict(h_1=h1, c_1=c1, logit=logit).items()}
    loss = (h1 * r["h_1"]).sum() + (c1 * r["c_1"]).sum() + (logit * r["logit"]).sum()
    dec.zero_grad()
    loss.backward()
    grads = {"gin." + k: v.grad for k, v in leaves.items()}
    grads.update({"gw." + k: p.grad for k, p in dec.named_parameters()})
    save("follower_step_small_train", **w, **{"in." + k: v for k, v in x.items()},
         **{"drop.x": dx, "drop.h": dh}, **{"cot." + k: v for k, v in r.items()},
         **{"out.h_1": h1, "out.c_1": c1, "out.alpha": alpha, "out.logit": logit, "out.alpha_v": alpha_v}, **grads)


def follower_step_full(name, B, L, A, seed):
    w = synth.follower_decoder_weights()
    dec = ref.AttnDecoderLSTM(synth.FEAT, synth.HID, 0.5).eval()
    dec.load_state_dict(w)
    x = synth.follower_step_inputs(B, L, A, seed=seed)
    with torch.no_grad():
        h1, c1, alpha, logit, alpha_v = dec(x["u_t_prev"], x["all_u_t"], x["visual_context"], x["h_0"], x["c_0"],
                                            x["ctx"], x["ctx_mask"])
    save(name, B=B, L=L, A=A, seed=seed, **{"out.h_1": h1, "out.c_1": c1, "out.alpha": alpha,
                                             "out.logit": logit, "out.alpha_v": alpha_v})


def encoder_cases():
    g = np.random.Generator(np.random.PCG64(102))
    for bidir in (False, True):
        V, Wd, H, B, L = 30, 12, 16, 4, 6
        torch.manual_seed(2)
        enc = ref.EncoderLSTM(V, Wd, H // 2 if bidir else H, 0, 0.5, bidirectional=bidir,
                              glove=g.standard_normal((V, Wd)).astype(np.float32)).eval()
        seq, mask, lengths = synth.instruction_batch(B, L, seed=5, vocab=V)
        with torch.no_grad():
            ctx, h, c = enc(seq, lengths)
        save("encoder_small_bi" if bidir else "encoder_small",
             **{"w." + k: v for k, v in enc.state_dict().items()}, seq=seq, lengths=np.array(lengths),
             **{"out.ctx": ctx, "out.h": h, "out.c": c})
    # full dims
    w = synth.follower_encoder_weights()
    enc = ref.EncoderLSTM(synth.VOCAB, synth.WORD, synth.HID, 0, 0.5, glove=w["embedding.weight"].numpy()).eval()
    enc.load_state_dict(w)
    B, L = 8, 20
    seq, mask, lengths = synth.instruction_batch(B, L, seed=41)
    with torch.no_grad():
        ctx, h, c = enc(seq, lengths)
    save("encoder_full", B=B, L=L, seed=41, **{"out.ctx": ctx, "out.h": h, "out.c": c})


def speaker_cases():
    g = np.random.Generator(np.random.PCG64(103))
    t = lambda *s: torch.from_numpy(g.standard_normal(s).astype(np.float32))
    E, F, H, N, T = 24, 20, 16, 3, 4
    torch.manual_seed(3)
    enc = ref.SpeakerEncoderLSTM(E, F, H, 0.5).eval()
    acts = [t(N, E) for _ in range(T)]
    feats = [torch.relu(t(N, 36, F)) for _ in range(T)]
    with torch.no_grad():
        ctx, h, c = enc(acts, feats)
    save("speaker_encoder_small", **{"w." + k: v for k, v in enc.state_dict().items()},
         acts=torch.stack(acts), feats=torch.stack(feats), **{"out.ctx": ctx, "out.h": h, "out.c": c})

    V, Wd = 30, 12
    dec = ref.SpeakerDecoderLSTM(V, Wd, H, 0.5, glove=g.standard_normal((V, Wd)).astype(np.float32)).eval()
    prev = torch.from_numpy(g.integers(0, V, size=(N, 1)))
    mask = torch.zeros(N, T, dtype=torch.bool)
    mask[1, 3:] = True
    mask[2, 2:] = True
    with torch.no_grad():
        h1, c1, alpha, logit = dec(prev, h, c, ctx, mask)
    save("speaker_decoder_small", **{"w." + k: v for k, v in dec.state_dict().items()},
         prev=prev, h_0=h, c_0=c, ctx=ctx, mask=mask,
         **{"out.h_1": h1, "out.c_1": c1, "out.alpha": alpha, "out.logit": logit})

    # full dims: encoder over T=3 path steps of N=4, then 5 teacher-forced decoder steps
    we, wd = synth.speaker_encoder_weights(), synth.speaker_decoder_weights()
    enc = ref.SpeakerEncoderLSTM(synth.FEAT, synth.FEAT, synth.HID, 0.5).eval()
    enc.load_state_dict(we)
    dec = ref.SpeakerDecoderLSTM(synth.VOCAB, synth.WORD, synth.HID, 0.5, glove=wd["embedding.weight"].numpy()).eval()
    dec.load_state_dict(wd)
    N, T, S = 4, 3, 5
    acts, feats = [], []
    for k in range(T):
        x = synth.follower_step_inputs(N, 8, 6, seed=200 + k)
        acts.append(x["u_t_prev"])
        feats.append(x["visual_context"])
    mask = torch.zeros(N, T, dtype=torch.bool)
    mask[2, 2:] = True
    mask[3, 1:] = True
    words = synth.instruction_batch(N, S, seed=77)[0]
    with torch.no_grad():
        ctx, h, c = enc(acts, feats)
        outs = {"out.ctx": ctx, "out.h": h, "out.c": c}
        w_t = torch.full((N, 1), 3, dtype=torch.long)
        for s in range(S):
            h, c, alpha, logit = dec(w_t, h, c, ctx, mask)
            outs["out.logit%d" % s] = logit
            outs["out.h%d" % s] = h
            w_t = words[:, s:s + 1]
    save("speaker_full", N=N, T=T, S=S, mask=mask, words=words, **outs)


def follower_rollout_full():
    """Encoder + 4 greedy decode steps at full dims (C1-like, B=8, L=20) with the per-step tail done the
    reference's way (follower.py:476-505) in plain torch here."""
    we, wd = synth.follower_encoder_weights(), synth.follower_decoder_weights()
    enc = ref.EncoderLSTM(synth.VOCAB, synth.WORD, synth.HID, 0, 0.5, glove=we["embedding.weight"].numpy()).eval()
    enc.load_state_dict(we)
    dec = ref.AttnDecoderLSTM(synth.FEAT, synth.HID, 0.5).eval()
    dec.load_state_dict(wd)
    B, L, A, S = 8, 20, 6, 4
    seq, mask, lengths = synth.instruction_batch(B, L, seed=41)
    outs = {}
    with torch.no_grad():
        ctx, h, c = enc(seq, lengths)
        u_prev = dec.u_begin.expand(B, -1)
        score = torch.zeros(B)
        for s in range(S):
            x = synth.follower_step_inputs(B, L, A, seed=300 + s)
            h, c, alpha, logit, alpha_v = dec(u_prev, x["all_u_t"], x["visual_context"], h, c, ctx, mask)
            logit[x["is_valid"] == 0] = -float("inf")
            _, a_t = logit.max(1)
            u_prev = x["all_u_t"][np.arange(B), a_t, :]
            sc = -torch.nn.functional.cross_entropy(logit, a_t, reduction="none")
            score = score + sc
            outs["out.logit%d" % s] = logit
            outs["out.a%d" % s] = a_t
            outs["out.score%d" % s] = sc
        outs["out.h"] = h
        outs["out.c"] = c
        outs["out.seq_score"] = score
    save("follower_rollout_full", B=B, L=L, A=A, S=S, **outs)


if __name__ == "__main__":
    follower_step_small()
    follower_step_full("follower_step_c1", 8, 20, 6, 31)
    follower_step_full("follower_step_c2", 100, 80, 8, 32)
    follower_step_full("follower_step_b3", 3, 13, 14, 33)
    encoder_cases()
    speaker_cases()
    follower_rollout_full()
