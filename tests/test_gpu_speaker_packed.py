"""GPU parity of the packed-weight tcgen05 paths of the two speaker steps (sfb_vis_lstm_pack_weights /
sfb_speaker_encoder_step_packed_fwd, sfb_speaker_decoder_pack_weights / sfb_speaker_decoder_step_packed_fwd) against the
CPU oracle and the in-place path.  Tolerance 1e-4 absolute (BASELINE.json north_star), argmax word identical."""
import pytest
import torch

from oracle import r2r_oracle as O
from speaker_follower_b200 import ops, synth
from test_gpu_parity import close, cu

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,T", [(1, 3), (37, 10), (256, 6), (300, 4), (2560, 7)])
def test_speaker_decoder_packed_vs_oracle(N, T):
    """C3 shape (N=256 paths, T=6 path steps, vocabulary 991), ragged neighbours, and the C4 rescoring size (N=2560)."""
    g = torch.Generator().manual_seed(9 + N)
    H = synth.HID
    wd = synth.speaker_decoder_weights()
    wc = cu(wd)
    blob = ops.PackedSpeakerDecoder().get(wc)
    assert blob is not None
    ctx = torch.tanh(torch.randn(N, T, H, generator=g)); h0 = torch.tanh(torch.randn(N, H, generator=g))
    c0 = torch.randn(N, H, generator=g) * 0.5
    mask = torch.arange(T).unsqueeze(0) >= torch.randint(1, T + 1, (N, 1), generator=g)
    prev = torch.randint(0, synth.VOCAB, (N,), generator=g)
    ref = O.speaker_decoder_step(prev, h0, c0, ctx, mask, wd)
    res = ops.speaker_decoder_step(wc, prev.cuda(), h0.cuda(), c0.cuda(), ctx.cuda(), mask.cuda(), packed=blob)
    for a, b, k in zip(res, ref, ("h1", "c1", "alpha", "logit")):
        close(a, b, what=k)
    top = ref[3].topk(2, 1)[0]
    safe = top[:, 0] - top[:, 1] > 1e-4
    assert torch.equal(res[3].cpu().max(1)[1][safe], ref[3].max(1)[1][safe])
    inplace = ops.speaker_decoder_step(wc, prev.cuda(), h0.cuda(), c0.cuda(), ctx.cuda(), mask.cuda())
    for a, b, k in zip(res, inplace, ("h1", "c1", "alpha", "logit")):
        close(a, b.cpu(), 5e-5, "packed-vs-inplace:" + k)
    # train mode: dropout on the embedding (no-GloVe configuration) and on h_1
    drop_e = (torch.rand(N, synth.WORD, generator=g) > 0.5).float() * 2.0
    drop_h = (torch.rand(N, H, generator=g) > 0.5).float() * 2.0
    ref = O.speaker_decoder_step(prev, h0, c0, ctx, mask, wd, use_glove=False, drop_e=drop_e, drop_h=drop_h)
    res = ops.speaker_decoder_step(wc, prev.cuda(), h0.cuda(), c0.cuda(), ctx.cuda(), mask.cuda(), drop_e.cuda(), drop_h.cuda(),
                                   packed=blob)
    for a, b, k in zip(res, ref, ("h1", "c1", "alpha", "logit")):
        close(a, b, what="train:" + k)


@pytest.mark.parametrize("N", [1, 64, 256, 2560])
def test_speaker_encoder_step_packed_vs_oracle(N):
    we = synth.speaker_encoder_weights()
    wc = cu(we)
    blob = ops.PackedVisLstm().get(wc)
    assert blob is not None
    x = synth.follower_step_inputs(N, 8, 6, seed=400 + N)
    a, V, h0, c0 = x["all_u_t"][:, 1].contiguous(), x["visual_context"], x["h_0"], x["c_0"]
    ref = O.speaker_encoder_step(h0, c0, a, V, we)
    res = ops.speaker_encoder_step(wc, a.cuda(), V.cuda(), h0.cuda(), c0.cuda(), packed=blob)
    close(res[0], ref[0], what="h1"); close(res[1], ref[1], what="c1")
    g = torch.Generator().manual_seed(1)
    dx = (torch.rand(N, 2 * synth.FEAT, generator=g) > 0.5).float() * 2.0
    ref = O.speaker_encoder_step(h0, c0, a, V, we, dx)
    res = ops.speaker_encoder_step(wc, a.cuda(), V.cuda(), h0.cuda(), c0.cuda(), dx.cuda(), packed=blob)
    close(res[0], ref[0], what="train:h1"); close(res[1], ref[1], what="train:c1")
    # several path steps chained (SpeakerEncoderLSTM.forward, model.py:437-457)
    h, c, hr, cr = h0.cuda(), c0.cuda(), h0, c0
    for s in range(3):
        xs = synth.follower_step_inputs(N, 8, 6, seed=500 + s)
        a, V = xs["all_u_t"][:, 1].contiguous(), xs["visual_context"]
        h, c = ops.speaker_encoder_step(wc, a.cuda(), V.cuda(), h, c, packed=blob)
        hr, cr = O.speaker_encoder_step(hr, cr, a, V, we)
    close(h, hr, what="chain:h"); close(c, cr, what="chain:c")
