"""Pins the CPU oracle (oracle/r2r_oracle.py) to vectors produced by the reference's own model.py
(tests/golden/make_golden.py).  Tolerance 1e-5 abs: same operator definitions on the same torch CPU
kernels, only the composition is restated."""
import numpy as np
import torch

from conftest import load_golden, split_golden
from oracle import r2r_oracle as O
from speaker_follower_b200 import synth

TOL = 1e-5


def close(a, b, tol=TOL):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    fin = torch.isfinite(b)
    assert torch.equal(torch.isfinite(a), fin)
    err = (a[fin] - b[fin]).abs().max().item() if fin.any() else 0.0
    assert err <= tol, err


def test_follower_step_small():
    w, x, out, _ = split_golden(load_golden("follower_step_small"))
    h1, c1, alpha, logit, alpha_v = O.attn_decoder_step(
        x["u_t_prev"], x["all_u_t"], x["visual_context"], x["h_0"], x["c_0"], x["ctx"], x["ctx_mask"], w)
    for k, v in dict(h_1=h1, c_1=c1, alpha=alpha, logit=logit, alpha_v=alpha_v).items():
        close(v, out[k])


def test_follower_step_small_train_and_grads():
    w, x, out, rest = split_golden(load_golden("follower_step_small_train"))
    w = {k: v.clone().requires_grad_(True) for k, v in w.items()}
    xin = {k: (v.clone().requires_grad_(True) if v.dtype == torch.float32 else v) for k, v in x.items()}
    h1, c1, alpha, logit, alpha_v = O.attn_decoder_step(
        xin["u_t_prev"], xin["all_u_t"], xin["visual_context"], xin["h_0"], xin["c_0"], xin["ctx"],
        xin["ctx_mask"], w, drop_x=rest["drop.x"], drop_h=rest["drop.h"])
    for k, v in dict(h_1=h1, c_1=c1, alpha=alpha, logit=logit, alpha_v=alpha_v).items():
        close(v.detach(), out[k])
    loss = (h1 * rest["cot.h_1"]).sum() + (c1 * rest["cot.c_1"]).sum() + (logit * rest["cot.logit"]).sum()
    loss.backward()
    for k, v in xin.items():
        if v.dtype == torch.float32:
            close(v.grad, rest["gin." + k], 2e-5)
    for k, v in w.items():
        close(v.grad, rest["gw." + k], 5e-5)


def _full_step(name):
    z = load_golden(name)
    B, L, A, seed = int(z["B"]), int(z["L"]), int(z["A"]), int(z["seed"])
    w = synth.follower_decoder_weights()
    x = synth.follower_step_inputs(B, L, A, seed=seed)
    h1, c1, alpha, logit, alpha_v = O.attn_decoder_step(
        x["u_t_prev"], x["all_u_t"], x["visual_context"], x["h_0"], x["c_0"], x["ctx"], x["ctx_mask"], w)
    _, _, out, _ = split_golden(z)
    for k, v in dict(h_1=h1, c_1=c1, alpha=alpha, logit=logit, alpha_v=alpha_v).items():
        close(v, out[k], 2e-5)
    assert torch.equal(logit.max(1)[1], out["logit"].max(1)[1])


def test_follower_step_c1():
    _full_step("follower_step_c1")


def test_follower_step_c2():
    _full_step("follower_step_c2")


def test_follower_step_b3_ragged_actions():
    _full_step("follower_step_b3")


def test_encoder_small_uni_and_bi():
    for name, bi in (("encoder_small", False), ("encoder_small_bi", True)):
        w, _, out, rest = split_golden(load_golden(name))
        ctx, h, c = O.encoder_lstm(rest["seq"], rest["lengths"].tolist(), w, bidirectional=bi)
        close(ctx, out["ctx"]); close(h, out["h"]); close(c, out["c"])


def test_encoder_full():
    z = load_golden("encoder_full")
    w = synth.follower_encoder_weights()
    seq, mask, lengths = synth.instruction_batch(int(z["B"]), int(z["L"]), seed=int(z["seed"]))
    ctx, h, c = O.encoder_lstm(seq, lengths, w)
    _, _, out, _ = split_golden(z)
    close(ctx, out["ctx"], 2e-5); close(h, out["h"], 2e-5); close(c, out["c"], 2e-5)


def test_speaker_small():
    w, _, out, rest = split_golden(load_golden("speaker_encoder_small"))
    ctx, h, c = O.speaker_encoder(list(rest["acts"]), list(rest["feats"]), w)
    close(ctx, out["ctx"]); close(h, out["h"]); close(c, out["c"])
    w, _, out, rest = split_golden(load_golden("speaker_decoder_small"))
    h1, c1, alpha, logit = O.speaker_decoder_step(rest["prev"], rest["h_0"], rest["c_0"], rest["ctx"], rest["mask"], w)
    close(h1, out["h_1"]); close(c1, out["c_1"]); close(alpha, out["alpha"]); close(logit, out["logit"])


def test_speaker_full():
    z = load_golden("speaker_full")
    _, _, out, rest = split_golden(z)
    N, T, S = int(z["N"]), int(z["T"]), int(z["S"])
    we, wd = synth.speaker_encoder_weights(), synth.speaker_decoder_weights()
    acts, feats = [], []
    for k in range(T):
        x = synth.follower_step_inputs(N, 8, 6, seed=200 + k)
        acts.append(x["u_t_prev"]); feats.append(x["visual_context"])
    ctx, h, c = O.speaker_encoder(acts, feats, we)
    close(ctx, out["ctx"], 2e-5); close(h, out["h"], 2e-5); close(c, out["c"], 2e-5)
    w_t = torch.full((N,), 3, dtype=torch.long)
    for s in range(S):
        h, c, alpha, logit = O.speaker_decoder_step(w_t, h, c, ctx, rest["mask"], wd)
        close(logit, out["logit%d" % s], 5e-5); close(h, out["h%d" % s], 2e-5)
        w_t = rest["words"][:, s]


def test_follower_rollout_full():
    z = load_golden("follower_rollout_full")
    _, _, out, _ = split_golden(z)
    B, L, A, S = int(z["B"]), int(z["L"]), int(z["A"]), int(z["S"])
    we, wd = synth.follower_encoder_weights(), synth.follower_decoder_weights()
    seq, mask, lengths = synth.instruction_batch(B, L, seed=41)
    steps = []
    for s in range(S):
        x = synth.follower_step_inputs(B, L, A, seed=300 + s)
        steps.append({"visual": x["visual_context"], "all_u_t": x["all_u_t"], "is_valid": x["is_valid"]})
    res, loss, score = O.follower_rollout(seq, mask, lengths, steps, we, wd, feedback="argmax")
    for s in range(S):
        close(res[s]["logit"], out["logit%d" % s], 5e-5)
        assert torch.equal(res[s]["a_t"], out["a%d" % s])
        close(res[s]["scores"], out["score%d" % s], 5e-5)
    close(score, out["seq_score"], 1e-4)
    close(res[-1]["h"], out["h"], 2e-5)


def test_rational_combine():
    best = O.rational_combine([-1.0, -3.0, -2.0, -0.5], [-2.0, -0.1, -1.0, -4.0], ["a", "a", "b", "b"], 0.95)
    assert best == {"a": 0, "b": 3}
    best = O.rational_combine([-1.0, -3.0, -2.0, -0.5], [-2.0, -0.1, -1.0, -4.0], ["a", "a", "b", "b"], 0.0)
    assert best == {"a": 1, "b": 2}
