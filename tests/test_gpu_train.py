"""Training support: forward on the sm_100a kernels, gradients by torch autograd
over the device-side restatement in speaker_follower_b200/_functional.py with the same dropout masks.
Pinned against the gradients the REFERENCE's own modules produce (tests/golden/follower_step_small_train.npz, written
by tests/golden/make_golden.py from tasks/R2R/model.py) and against autograd through the CPU oracle at full size."""
import pytest
import torch

from conftest import load_golden, split_golden
from fake_env import FakeR2RBatch
from oracle import r2r_oracle as O
from speaker_follower_b200 import _functional as Fn, follower as Fo, model as M, ops, synth
from test_gpu_parity import NAMES, close, cu
from test_gpu_parity import close as _close

pytestmark = pytest.mark.gpu
IN = ("u_t_prev", "all_u_t", "visual_context", "h_0", "c_0", "ctx")


def test_reference_gradients_small_train_golden():
    """Same inputs, same dropout masks, same cotangents as the reference run -> same input and weight gradients."""
    w, x, out, rest = split_golden(load_golden("follower_step_small_train"))
    names = list(w.keys())
    wc = {k: v.cuda().requires_grad_(True) for k, v in w.items()}
    xin = [x[k].cuda().requires_grad_(True) for k in IN]
    mask, dx, dh = x["ctx_mask"].cuda(), rest["drop.x"].cuda(), rest["drop.h"].cuda()

    def run_cuda():
        with torch.no_grad():
            return ops.follower_step({k: v.detach() for k, v in wc.items()}, *[t.detach() for t in xin], mask, dx, dh)
    res = Fn.FollowerStepFn.apply(run_cuda, names, len(xin), mask, dx, dh, *xin, *[wc[k] for k in names])
    for k, v in zip(NAMES, res):
        close(v, out[k], what="fwd:" + k)                       # the forward really came from the CUDA kernels
    loss = (res[0] * rest["cot.h_1"].cuda()).sum() + (res[1] * rest["cot.c_1"].cuda()).sum() + \
           (res[3] * rest["cot.logit"].cuda()).sum()
    loss.backward()
    for k, t in zip(IN, xin):
        close(t.grad, rest["gin." + k], 5e-5, "gin." + k)
    for k in names:
        close(wc[k].grad, rest["gw." + k], 1e-4, "gw." + k)


def test_kernel_backward_reproduces_reference_gradients_small_train_golden():
    """The hand-written backward (sfb_follower_step_bwd, backward.cu — no torch op computes a gradient) against the
    gradients the REFERENCE's own modules produced for the same inputs, dropout masks and cotangents."""
    import ctypes as C
    w, x, out, rest = split_golden(load_golden("follower_step_small_train"))
    names = list(w.keys())
    wc = {k: v.cuda().requires_grad_(True) for k, v in w.items()}
    xin = [x[k].cuda().requires_grad_(True) for k in IN]
    mask, dx, dh = x["ctx_mask"].cuda(), rest["drop.x"].cuda(), rest["drop.h"].cuda()
    B, A = xin[1].shape[0], xin[1].shape[1]

    def run_cuda():
        with torch.no_grad():
            wd = {k: v.detach() for k, v in wc.items()}
            d = ops.follower_dims(wd, xin[2].shape[1])
            need = ops._lib.load().sfb_follower_step_workspace_bytes(C.byref(d), B, xin[5].shape[1], A)
            fwd_ws = torch.zeros(need, dtype=torch.uint8, device="cuda")
            return ops.follower_step(wd, *[t.detach() for t in xin], mask, dx, dh, workspace=fwd_ws), fwd_ws
    res = Fn.FollowerStepKernelFn.apply(run_cuda, names, len(xin), mask, dx, dh, *xin, *[wc[k] for k in names])
    for k, v in zip(NAMES, res):
        close(v, out[k], what="fwd:" + k)
    loss = (res[0] * rest["cot.h_1"].cuda()).sum() + (res[1] * rest["cot.c_1"].cuda()).sum() + \
           (res[3] * rest["cot.logit"].cuda()).sum()
    loss.backward()
    for k, t in zip(IN, xin):
        if k in ("h_0", "c_0", "ctx"):                     # u_t_prev / candidates / slab are constants of the step
            close(t.grad, rest["gin." + k], 5e-5, "gin." + k)
    for k in names:
        if k == "visual_attention_layer.linear_in_v.bias":  # cancels in the softmax: the reference's gradient is ~0 too
            assert rest["gw." + k].abs().max() < 1e-6
            continue
        close(wc[k].grad, rest["gw." + k], 1e-4, "gw." + k)


def test_module_gradients_match_oracle_autograd_full_size():
    """nn.Module level (eval mode: no dropout), reference dimensions, packed forward: grads == autograd through the oracle."""
    B, L, A = 8, 20, 6
    w = synth.follower_decoder_weights()
    x = synth.follower_step_inputs(B, L, A, seed=77)
    dec = M.AttnDecoderLSTM(synth.FEAT, synth.HID, 0.5).cuda().eval()
    dec.load_state_dict(w)
    xc = cu(x)
    h0, c0, ctx = xc["h_0"].clone().requires_grad_(True), xc["c_0"].clone().requires_grad_(True), xc["ctx"].clone().requires_grad_(True)
    g = torch.Generator().manual_seed(1)
    cot = [torch.randn(B, synth.HID, generator=g), torch.randn(B, synth.HID, generator=g), torch.randn(B, A, generator=g)]
    h1, c1, alpha, logit, av = dec(xc["u_t_prev"], xc["all_u_t"], xc["visual_context"], h0, c0, ctx, xc["ctx_mask"])
    ((h1 * cot[0].cuda()).sum() + (c1 * cot[1].cuda()).sum() + (logit * cot[2].cuda()).sum()).backward()
    # oracle autograd on the CPU
    wr = {k: v.clone().requires_grad_(True) for k, v in w.items()}
    hr, cr, ctr = x["h_0"].clone().requires_grad_(True), x["c_0"].clone().requires_grad_(True), x["ctx"].clone().requires_grad_(True)
    r = O.attn_decoder_step(x["u_t_prev"], x["all_u_t"], x["visual_context"], hr, cr, ctr, x["ctx_mask"], wr)
    ((r[0] * cot[0]).sum() + (r[1] * cot[1]).sum() + (r[3] * cot[2]).sum()).backward()
    close(h1, r[0], what="h1"); close(logit, r[3], what="logit")
    close(h0.grad, hr.grad, 1e-4, "d h_0"); close(c0.grad, cr.grad, 1e-4, "d c_0"); close(ctx.grad, ctr.grad, 1e-4, "d ctx")
    for k, p in dec.named_parameters():
        scale = max(1.0, float(wr[k].grad.abs().max()))
        close(p.grad / scale, wr[k].grad / scale, 1e-4, "d " + k)


def test_agent_train_iterations_reduce_the_loss():
    """Seq2SeqAgent.train (follower.py:1001-1020) runs end to end: encoder + decoder receive gradients, Adam steps
    change the weights (the packed copies follow), and the teacher-forced loss on the same batch goes down."""
    env = FakeR2RBatch(n_viewpoints=20, n_instr=8, batch_size=8, seed=11)
    glove = synth.follower_encoder_weights()["embedding.weight"].numpy()
    enc = M.EncoderLSTM(synth.VOCAB, synth.WORD, synth.HID, 0, 0.0, glove=glove).cuda()
    dec = M.AttnDecoderLSTM(synth.FEAT, synth.HID, 0.0).cuda()
    enc.load_state_dict(synth.follower_encoder_weights()); dec.load_state_dict(synth.follower_decoder_weights())
    agent = Fo.Seq2SeqAgent(env, "", enc, dec, episode_len=6, max_instruction_length=20)
    eo = torch.optim.Adam([p for p in enc.parameters() if p.requires_grad], lr=1e-3)
    do = torch.optim.Adam(dec.parameters(), lr=1e-3)
    w0 = dec.lstm.weight_ih.detach().clone()
    agent.train(eo, do, 1, feedback="teacher")
    # (linear_in_v.bias cancels in the softmax over the views: its exact gradient is zero)
    assert all(p.grad is not None and torch.isfinite(p.grad).all() and (p.grad.abs().sum() > 0 or k.endswith("linear_in_v.bias"))
               for k, p in dec.named_parameters())
    assert all(p.grad is not None and p.grad.abs().sum() > 0 for p in enc.parameters() if p.requires_grad)
    first = agent.losses[0]
    agent.train(eo, do, 5, feedback="teacher")
    assert not torch.equal(dec.lstm.weight_ih.detach(), w0)
    assert agent.losses[-1] < first, (first, agent.losses)
    # inference after training uses the re-packed weights: packed step == in-place step on the updated parameters
    dec.eval(); enc.eval()
    agent.feedback = "argmax"
    with torch.no_grad():
        traj = agent.rollout()
    assert len(traj) == 8


def test_speaker_train_iteration_and_gradients():
    """Seq2SeqSpeaker.train (speaker.py:376-395): speaker encoder/decoder gradients == autograd through the CPU oracle
    (teacher-forced loss on the same batch), and Adam iterations reduce the loss."""
    from speaker_follower_b200 import speaker as Sp
    env = FakeR2RBatch(n_viewpoints=16, n_instr=6, batch_size=6, seed=6)
    we, wd = synth.speaker_encoder_weights(), synth.speaker_decoder_weights()
    enc = M.SpeakerEncoderLSTM(synth.FEAT, synth.FEAT, synth.HID, 0.0).cuda()
    dec = M.SpeakerDecoderLSTM(synth.VOCAB, synth.WORD, synth.HID, 0.0, glove=wd["embedding.weight"].numpy()).cuda()
    enc.load_state_dict(we); dec.load_state_dict(wd)
    spk = Sp.Seq2SeqSpeaker(env, "", enc, dec, instruction_len=12, max_episode_len=6)
    path_obs, path_actions, encoded = env.gold_obs_actions_and_instructions(6)
    outs, loss = spk._score_obs_actions_and_instructions(path_obs, path_actions, encoded, "teacher")
    loss.backward()
    # the same loss through the CPU oracle with autograd
    _, feats, acts, mask, _, _, _ = spk._batch_observations_and_actions(path_obs, path_actions, encoded)
    instr, _, _ = Fo.batch_instructions_from_encoded(encoded, 12)
    wer = {k: v.clone().requires_grad_(True) for k, v in we.items()}
    wdr = {k: v.clone().requires_grad_(k != "embedding.weight") for k, v in wd.items()}
    sc, l, words, wsc = O.speaker_score_teacher([a.cpu() for a in acts], [f.cpu() for f in feats], mask.cpu().bool(), instr, wer, wdr)
    l.backward()
    assert abs(float(loss.detach()) - float(l.detach())) < 1e-3 * max(1.0, abs(float(l.detach())))
    for mod, ref_w in ((enc, wer), (dec, wdr)):
        for k, p in mod.named_parameters():
            if not p.requires_grad:
                continue
            ref = ref_w[k].grad
            scale = max(1.0, float(ref.abs().max()))
            close(p.grad / scale, ref / scale, 2e-4, "d " + k)
    eo, do = torch.optim.Adam(enc.parameters(), lr=1e-3), torch.optim.Adam([p for p in dec.parameters() if p.requires_grad], lr=1e-3)
    env.reset_epoch()
    spk.train(eo, do, 1)
    first = spk.losses[0]
    for _ in range(5):
        env.reset_epoch()
        spk.train(eo, do, 1)
    assert spk.losses[-1] < first


def test_encoder_kernel_backward_matches_oracle_autograd():
    """EncoderLSTM under autograd (model.py:81-104): taped forward + hand-written BPTT (sfb_encoder_lstm_bwd) against torch
    autograd through the CPU oracle — ragged lengths, all three outputs carrying gradient, no cuDNN involved."""
    B, L = 12, 20
    we = synth.follower_encoder_weights()
    seq, mask, lengths = synth.instruction_batch(B, L, seed=5)
    enc = M.EncoderLSTM(synth.VOCAB, synth.WORD, synth.HID, 0, 0.0, glove=we["embedding.weight"].numpy()).cuda().train()
    enc.load_state_dict(we)
    g = torch.Generator().manual_seed(3)
    maxlen = max(lengths)
    cots = [torch.randn(B, maxlen, synth.HID, generator=g), torch.randn(B, synth.HID, generator=g), torch.randn(B, synth.HID, generator=g)]
    ctx, dec, c = enc(seq.cuda(), lengths)
    assert ctx.grad_fn is not None and "EncoderLstmKernelFn" in type(ctx.grad_fn).__name__
    ((ctx * cots[0].cuda()).sum() + (dec * cots[1].cuda()).sum() + (c * cots[2].cuda()).sum()).backward()
    wr = {k: v.clone().requires_grad_(k != "embedding.weight") for k, v in we.items()}
    rc, rd, rct = O.encoder_lstm(seq[:, :maxlen], lengths, wr)
    ((rc * cots[0]).sum() + (rd * cots[1]).sum() + (rct * cots[2]).sum()).backward()
    close(ctx, rc, what="ctx"); close(dec, rd, what="decoder_init"); close(c, rct, what="c_t")
    for k, p in enc.named_parameters():
        if not p.requires_grad:
            continue
        scale = max(1.0, float(wr[k].grad.abs().max()))
        close(p.grad / scale, wr[k].grad / scale, 1e-4, "d " + k)


def test_speaker_step_kernel_backward_with_dropout_masks():
    """sfb_speaker_encoder_step_bwd / sfb_speaker_decoder_step_bwd with dropout masks (train_speaker.py trains with p=0.5):
    gradients of a random linear functional of the outputs against torch autograd over the fp32 restatement of the same
    step (speaker_follower_b200/_functional.py: model.py:429-435, 487-519) with the SAME masks."""
    torch.manual_seed(3)
    close = lambda a, b, tol, what: _close(a, torch.as_tensor(b).detach().cpu(), tol, what)
    N, T = 24, 5
    we = cu(synth.speaker_encoder_weights()); wd = cu(synth.speaker_decoder_weights())
    x = cu(synth.follower_step_inputs(N, 8, 6, seed=9))
    a, V, h0, c0 = x["all_u_t"][:, 1].contiguous(), x["visual_context"], x["h_0"], x["c_0"]
    keep = lambda *s: (torch.rand(*s, device="cuda") > 0.5).float() * 2.0
    # ---- encoder step
    dx = keep(N, 2 * synth.FEAT)
    lib = ops._lib.load()
    need = lib.sfb_follower_step_workspace_bytes(ops.C.byref(ops.follower_dims(we, V.shape[1])), N, 1, 1)
    for packed in (None, ops.PackedVisLstm().get(we)):
        ws = torch.zeros(need, dtype=torch.uint8, device="cuda")
        h1, c1 = ops.speaker_encoder_step(we, a, V, h0, c0, dx, packed=packed, workspace=ws)
        gh, gc = torch.randn_like(h1), torch.randn_like(c1)
        grads = {k: torch.zeros_like(v) for k, v in we.items() if not k.startswith("encoder2decoder.")}
        d_h0, d_c0 = ops.speaker_encoder_step_bwd(we, a, V, h0, c0, dx, c1, ws, gh, gc, grads, accumulate=False)
        wr = {k: v.clone().requires_grad_(True) for k, v in we.items()}
        h0r, c0r = h0.clone().requires_grad_(True), c0.clone().requires_grad_(True)
        h1r, c1r = Fn.speaker_encoder_step(wr, a, V, h0r, c0r, dx)
        ((h1r * gh).sum() + (c1r * gc).sum()).backward()
        close(h1, h1r.detach(), 1e-4, "enc h1"); close(d_h0, h0r.grad, 2e-4, "enc d_h0"); close(d_c0, c0r.grad, 2e-4, "enc d_c0")
        for k, g in grads.items():
            ref = wr[k].grad if wr[k].grad is not None else torch.zeros_like(g)
            scale = max(1.0, float(ref.abs().max()))
            close(g / scale, ref / scale, 2e-4, "enc d " + k)
    # ---- decoder step
    ctx = torch.tanh(torch.randn(N, T, synth.HID, device="cuda"))
    mask = torch.arange(T, device="cuda").unsqueeze(0) >= torch.randint(2, T + 1, (N, 1), device="cuda")
    prev = torch.randint(4, synth.VOCAB, (N,), device="cuda")
    dh = keep(N, synth.HID)
    need = lib.sfb_speaker_decoder_step_workspace_bytes(synth.HID, synth.WORD, N, T)
    for packed in (None, ops.PackedSpeakerDecoder().get(wd)):
        ws = torch.zeros(need, dtype=torch.uint8, device="cuda")
        h1, c1, alpha, logit = ops.speaker_decoder_step(wd, prev, h0, c0, ctx, mask, None, dh, packed=packed, workspace=ws)
        gh, gc, gl = torch.randn_like(h1), torch.randn_like(c1), torch.randn_like(logit) * 0.1
        grads = {k: torch.zeros_like(v) for k, v in wd.items() if k != "embedding.weight"}
        d_h0, d_c0, d_ctx = ops.speaker_decoder_step_bwd(wd, prev, h0, c0, ctx, mask, None, dh, c1, alpha, ws, gh, gc, gl, grads,
                                                          accumulate=False)
        wr = {k: v.clone().requires_grad_(k != "embedding.weight") for k, v in wd.items()}
        h0r, c0r, ctxr = h0.clone().requires_grad_(True), c0.clone().requires_grad_(True), ctx.clone().requires_grad_(True)
        h1r, c1r, alphar, logitr = Fn.speaker_decoder_step(wr, prev, h0r, c0r, ctxr, mask, None, dh)
        ((h1r * gh).sum() + (c1r * gc).sum() + (logitr * gl).sum()).backward()
        close(logit, logitr.detach(), 1e-4, "dec logit")
        close(d_h0, h0r.grad, 2e-4, "dec d_h0"); close(d_c0, c0r.grad, 2e-4, "dec d_c0"); close(d_ctx, ctxr.grad, 2e-4, "dec d_ctx")
        for k, g in grads.items():
            ref = wr[k].grad
            scale = max(1.0, float(ref.abs().max()))
            close(g / scale, ref / scale, 2e-4, "dec d " + k)
