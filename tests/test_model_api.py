"""The drop-in module layer (speaker_follower_b200.model) mirrors the reference's tasks/R2R/model.py:
state_dict key names / shapes are pinned on the CPU against the golden files written from the reference's own
modules; forwards are checked on the GPU."""
import numpy as np
import pytest
import torch

from conftest import load_golden, split_golden
from speaker_follower_b200 import model as M, synth


def _keys(name):
    w, _, _, _ = split_golden(load_golden(name))
    return {k: tuple(v.shape) for k, v in w.items()}


def test_state_dict_keys_match_reference():
    ref = _keys("follower_step_small")           # E=48, F=40, H=32
    dec = M.AttnDecoderLSTM(48, 32, 0.5, feature_size=40)
    assert {k: tuple(v.shape) for k, v in dec.state_dict().items()} == ref
    ref = _keys("encoder_small")
    enc = M.EncoderLSTM(30, 12, 16, 0, 0.5, glove=np.zeros((30, 12), np.float32))
    assert {k: tuple(v.shape) for k, v in enc.state_dict().items()} == ref
    ref = _keys("encoder_small_bi")
    enc = M.EncoderLSTM(30, 12, 8, 0, 0.5, bidirectional=True, glove=np.zeros((30, 12), np.float32))
    assert {k: tuple(v.shape) for k, v in enc.state_dict().items()} == ref
    ref = _keys("speaker_encoder_small")
    se = M.SpeakerEncoderLSTM(24, 20, 16, 0.5)
    assert {k: tuple(v.shape) for k, v in se.state_dict().items()} == ref
    ref = _keys("speaker_decoder_small")
    sd = M.SpeakerDecoderLSTM(30, 12, 16, 0.5, glove=np.zeros((30, 12), np.float32))
    assert {k: tuple(v.shape) for k, v in sd.state_dict().items()} == ref


def test_full_size_parameter_counts():
    """P_dec = 12,129,537 parameters; frozen GloVe embeddings are excluded from the optimiser
    exactly like the reference (train.py:64-65 filters on requires_grad)."""
    dec = M.AttnDecoderLSTM(synth.FEAT, synth.HID, 0.5)
    assert sum(p.numel() for p in dec.parameters()) == 12129537
    assert dec.u_begin.shape == (synth.FEAT,) and float(dec.u_begin.abs().sum()) == 0.0
    enc = M.EncoderLSTM(synth.VOCAB, synth.WORD, synth.HID, 0, 0.5, glove=np.zeros((synth.VOCAB, synth.WORD), np.float32))
    assert not enc.embedding.weight.requires_grad
    assert sum(p.numel() for p in enc.parameters() if p.requires_grad) == 4 * 512 * (300 + 512 + 2) + 512 * 513


def test_modules_refuse_cpu():
    """No CPU fallback: the forward always runs on the CUDA kernels (also under autograd, where only the gradients come
    from torch), so CPU tensors are refused with or without grad mode; the stand-alone attention sub-modules are
    forward-only."""
    from speaker_follower_b200._lib import SfbError
    dec = M.AttnDecoderLSTM(48, 32, 0.5, feature_size=40).eval()
    x = torch.zeros(2, 48), torch.zeros(2, 3, 48), torch.zeros(2, 36, 40), torch.zeros(2, 32), torch.zeros(2, 32), torch.zeros(2, 5, 32)
    with pytest.raises(SfbError):
        dec(*x)                                   # autograd recording: forward still goes to the (absent) GPU
    with torch.no_grad(), pytest.raises(SfbError):
        dec(*x)                                   # CPU tensors: no fallback
    with pytest.raises(NotImplementedError):
        dec.visual_attention_layer(x[3], x[2])    # sub-module on its own under autograd: not differentiable here


@pytest.mark.gpu
def test_modules_forward_golden():
    w, x, out, rest = split_golden(load_golden("follower_step_small_train"))
    dec = M.AttnDecoderLSTM(48, 32, 0.5, feature_size=40).cuda().eval()
    dec.load_state_dict(w)
    xc = {k: v.cuda() for k, v in x.items()}
    w0, x0, out0, _ = split_golden(load_golden("follower_step_small"))
    with torch.no_grad():
        res = dec(dec.u_begin.expand(5, -1) * 0 + xc["u_t_prev"], xc["all_u_t"], xc["visual_context"], xc["h_0"], xc["c_0"],
                  xc["ctx"], xc["ctx_mask"])
    for k, v in zip(("h_1", "c_1", "alpha", "logit", "alpha_v"), res):
        assert (v.cpu() - out0[k]).abs().max() < 1e-4, k
    # sub-modules are callable on their own like the reference's
    with torch.no_grad():
        f, av = dec.visual_attention_layer(xc["h_0"], xc["visual_context"])
        assert (av.cpu() - out0["alpha_v"]).abs().max() < 1e-5
        ht, al = dec.text_attention_layer(res[0], xc["ctx"], xc["ctx_mask"])
        assert (al.cpu() - out0["alpha"]).abs().max() < 1e-5

    for name, bi in (("encoder_small", False), ("encoder_small_bi", True)):
        w, _, out, rest = split_golden(load_golden(name))
        enc = M.EncoderLSTM(30, 12, 8 if bi else 16, 0, 0.5, bidirectional=bi, glove=np.zeros((30, 12), np.float32)).cuda().eval()
        enc.load_state_dict(w)
        with torch.no_grad():
            ctx, h, c = enc(rest["seq"].cuda(), rest["lengths"].tolist())
        assert (ctx.cpu() - out["ctx"]).abs().max() < 1e-4 and (h.cpu() - out["h"]).abs().max() < 1e-4

    w, _, out, rest = split_golden(load_golden("speaker_encoder_small"))
    se = M.SpeakerEncoderLSTM(24, 20, 16, 0.5).cuda().eval()
    se.load_state_dict(w)
    with torch.no_grad():
        ctx, h, c = se([a.cuda() for a in rest["acts"]], [f.cuda() for f in rest["feats"]])
    assert (ctx.cpu() - out["ctx"]).abs().max() < 1e-4 and (h.cpu() - out["h"]).abs().max() < 1e-4
    assert (c.cpu() - out["c"]).abs().max() < 1e-4

    w, _, out, rest = split_golden(load_golden("speaker_decoder_small"))
    sd = M.SpeakerDecoderLSTM(30, 12, 16, 0.5, glove=np.zeros((30, 12), np.float32)).cuda().eval()
    sd.load_state_dict(w)
    with torch.no_grad():
        h1, c1, alpha, logit = sd(rest["prev"].cuda(), rest["h_0"].cuda(), rest["c_0"].cuda(), rest["ctx"].cuda(), rest["mask"].cuda())
    assert (logit.cpu() - out["logit"]).abs().max() < 1e-4 and (alpha.cpu() - out["alpha"]).abs().max() < 1e-4


@pytest.mark.gpu
def test_train_mode_dropout_statistics():
    """train(): dropout masks are drawn and applied where the reference applies nn.Dropout (model.py:392,394);
    eval(): deterministic."""
    torch.manual_seed(0)
    dec = M.AttnDecoderLSTM(synth.FEAT, synth.HID, 0.5).cuda()
    dec.load_state_dict(synth.follower_decoder_weights())
    x = {k: v.cuda() for k, v in synth.follower_step_inputs(16, 20, 6, seed=3).items()}
    args = (x["u_t_prev"], x["all_u_t"], x["visual_context"], x["h_0"], x["c_0"], x["ctx"], x["ctx_mask"])
    with torch.no_grad():
        dec.eval()
        a, b = dec(*args), dec(*args)
        assert torch.equal(a[3], b[3])
        dec.train()
        c, d = dec(*args), dec(*args)
        assert not torch.equal(c[3], d[3])
        assert torch.isfinite(c[3]).all()


def test_feature_store_from_tsv(tmp_path):
    """The reference's precomputed-feature TSV (env.py:350-375) -> one contiguous table + index map;
    CPU-only check of the loader (device='cpu'), two files concatenated along the feature axis like env.py:372-375."""
    import base64
    import numpy as np
    from speaker_follower_b200 import ops
    rng = np.random.default_rng(0)
    ids = [("scanA", "vp%02d" % i) for i in range(5)] + [("scanB", "vp00")]
    feats = {k: rng.standard_normal((36, 2048)).astype(np.float32) for k in ids}
    feats2 = {k: rng.standard_normal((36, 2048)).astype(np.float32) for k in ids}

    def write(path, table, order):
        with open(path, "wt") as f:
            for k in order:
                f.write("\t".join([k[0], k[1], "640", "480", "60", base64.b64encode(table[k].tobytes()).decode()]) + "\n")

    p1, p2 = str(tmp_path / "a.tsv"), str(tmp_path / "b.tsv")
    write(p1, feats, ids)
    write(p2, feats2, ids[::-1])                                  # other file lists the viewpoints in another order
    loc = torch.zeros(36, 36, 128)
    store = ops.FeatureStore.from_tsv(p1, loc, device="cpu", chunk_rows=4)
    assert store.feat_table.shape == (6, 36, 2048) and store.img_dim == 2048
    for k in ids:
        assert torch.equal(store.feat_table[store.index[k[0] + "_" + k[1]]], torch.from_numpy(feats[k]))
    both = ops.FeatureStore.from_tsv([p1, p2], loc, device="cpu", chunk_rows=4)
    assert both.feat_table.shape == (6, 36, 4096)
    for k in ids:
        row = both.feat_table[both.index[k[0] + "_" + k[1]]]
        assert torch.equal(row[:, :2048], torch.from_numpy(feats[k])) and torch.equal(row[:, 2048:], torch.from_numpy(feats2[k]))
    assert both.rows(["scanB_vp00", "scanA_vp03"]).tolist() == [5, 3]
