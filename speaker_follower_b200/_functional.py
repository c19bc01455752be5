"""Device-side torch restatement of the follower decoder step, used ONLY to obtain gradients.

Forward passes always run on the sm_100a kernels (ops.py).  Hand-written backward kernels are not part of this round
(DESIGN.md §10); until they are, a training step differentiates by RE-COMPUTING the step with torch ops on the same
device from the saved inputs and dropout masks (library kernels, fp32) and calling torch.autograd on that graph.
The arithmetic follows tasks/R2R/model.py line by line (cited below); it is checked against the gradients the
reference's own modules produce (tests/golden/follower_step_small_train.npz, tests/test_gpu_train.py).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def follower_step(w: Dict[str, Tensor], u_t_prev: Tensor, all_u_t: Tensor, visual_context: Tensor, h_0: Tensor,
                  c_0: Tensor, ctx: Tensor, ctx_mask: Optional[Tensor], drop_x: Optional[Tensor],
                  drop_h: Optional[Tensor]):
    """AttnDecoderLSTM.forward — model.py:377-397 -> (h_1, c_1, alpha, logit, alpha_v)."""
    # VisualSoftDotAttention.forward — model.py:310-326 (mask ignored there too)
    t = F.linear(h_0, w["visual_attention_layer.linear_in_h.weight"], w["visual_attention_layer.linear_in_h.bias"])
    k = F.linear(visual_context, w["visual_attention_layer.linear_in_v.weight"], w["visual_attention_layer.linear_in_v.bias"])
    alpha_v = torch.softmax(torch.bmm(k, t.unsqueeze(2)).squeeze(2), dim=1)
    feature = torch.bmm(alpha_v.unsqueeze(1), visual_context).squeeze(1)
    # model.py:391-394
    x = torch.cat((u_t_prev, feature), 1)
    if drop_x is not None:
        x = x * drop_x
    gates = F.linear(x, w["lstm.weight_ih"], w["lstm.bias_ih"]) + F.linear(h_0, w["lstm.weight_hh"], w["lstm.bias_hh"])
    i, f, g, o = gates.chunk(4, 1)
    c_1 = torch.sigmoid(f) * c_0 + torch.sigmoid(i) * torch.tanh(g)
    h_1 = torch.sigmoid(o) * torch.tanh(c_1)
    h_d = h_1 * drop_h if drop_h is not None else h_1
    # SoftDotAttention.forward — model.py:122-143
    target = F.linear(h_d, w["text_attention_layer.linear_in.weight"]).unsqueeze(2)
    attn = torch.bmm(ctx, target).squeeze(2)
    if ctx_mask is not None:
        attn = attn.masked_fill(ctx_mask.bool(), -float("inf"))
    alpha = torch.softmax(attn, dim=1)
    weighted = torch.bmm(alpha.unsqueeze(1), ctx).squeeze(1)
    h_tilde = torch.tanh(F.linear(torch.cat((weighted, h_d), 1), w["text_attention_layer.linear_out.weight"]))
    # EltwiseProdScoring.forward — model.py:342-352
    th = F.linear(h_tilde, w["decoder2action.linear_in_h.weight"], w["decoder2action.linear_in_h.bias"]).unsqueeze(1)
    ta = F.linear(all_u_t, w["decoder2action.linear_in_a.weight"], w["decoder2action.linear_in_a.bias"])
    logit = F.linear(th * ta, w["decoder2action.linear_out.weight"], w["decoder2action.linear_out.bias"]).squeeze(2)
    return h_1, c_1, alpha, logit, alpha_v


class FollowerStepKernelFn(torch.autograd.Function):
    """One decode step, forward AND backward on the sm_100a library: the forward call runs in its own workspace, which
    keeps the step's intermediates (attention output, activated gates, dropped h_1, h~) for sfb_follower_step_bwd
    (backward.cu); parameter gradients are returned to autograd, which accumulates them over the steps of a rollout.
    u_t_prev, the action candidates and the visual context are constants of the step (follower.py:502)."""

    @staticmethod
    def forward(ctx_, run_cuda, names, n_in, ctx_mask, drop_x, drop_h, *tensors):
        inputs, params = tensors[:n_in], tensors[n_in:]
        outs, fwd_ws = run_cuda()
        ctx_.names, ctx_.n_in = names, n_in
        ctx_.ctx_mask, ctx_.drop_x, ctx_.drop_h, ctx_.fwd_ws = ctx_mask, drop_x, drop_h, fwd_ws
        ctx_.save_for_backward(*inputs, *params, outs[1], outs[2], outs[4])
        ctx_.mark_non_differentiable(outs[2], outs[4])
        return tuple(outs)

    @staticmethod
    def backward(ctx_, g_h1, g_c1, g_alpha, g_logit, g_alpha_v):
        from . import ops
        saved = ctx_.saved_tensors
        n_in, names = ctx_.n_in, ctx_.names
        u, U, V, h0, c0, c = saved[:n_in]
        params = saved[n_in:n_in + len(names)]
        c1, alpha, alpha_v = saved[n_in + len(names):]
        w = {k: p.detach() for k, p in zip(names, params)}
        grads = {k: torch.zeros_like(p) for k, p in w.items()}   # linear_in_v.bias stays zero: it cancels in the softmax
        with torch.no_grad():
            d_h0, d_c0, d_ctx = ops.follower_step_bwd(w, u.detach(), U.detach(), V.detach(), h0.detach(), c0.detach(), c.detach(),
                                                       ctx_.ctx_mask, ctx_.drop_x, ctx_.drop_h, c1, alpha, alpha_v, ctx_.fwd_ws,
                                                       g_h1, g_c1, g_logit, grads, accumulate=False)
        need = ctx_.needs_input_grad[6:]
        gin = [None, None, None, d_h0, d_c0, d_ctx]
        gin = [g if nd else None for g, nd in zip(gin, need[:n_in])]
        gpar = [(grads.get(k) if nd else None) for k, nd in zip(names, need[n_in:])]
        return (None, None, None, None, None, None, *gin, *gpar)


class EncoderLstmKernelFn(torch.autograd.Function):
    """EncoderLSTM.forward (model.py:81-104, unidirectional) forward AND backward on the library: the forward keeps a
    tape of every time step, sfb_encoder_lstm_bwd back-propagates through time over it (backward.cu + skinny GEMMs)."""

    @staticmethod
    def forward(ctx_, seq, lengths, drop_embed, names, *params):
        from . import ops
        w = {k: p.detach() for k, p in zip(names, params)}
        with torch.no_grad():
            ctx, dec, c_t, saved = ops.encoder_lstm_train(w, seq, lengths, drop_embed)
        ctx_.names, ctx_.saved, ctx_.drop_embed = names, saved, drop_embed
        ctx_.save_for_backward(dec, *params)
        return ctx, dec, c_t

    @staticmethod
    def backward(ctx_, g_ctx, g_dec, g_c):
        from . import ops
        dec, *params = ctx_.saved_tensors
        w = {k: p.detach() for k, p in zip(ctx_.names, params)}
        grads = {k: torch.zeros_like(p) for k, p in w.items() if k != "embedding.weight"}
        with torch.no_grad():
            ops.encoder_lstm_bwd(w, ctx_.saved, dec, g_ctx, g_dec, g_c, grads, ctx_.drop_embed)
        need = ctx_.needs_input_grad[4:]
        return (None, None, None, None, *[(grads.get(k) if nd else None) for k, nd in zip(ctx_.names, need)])


class SpeakerEncStepKernelFn(torch.autograd.Function):
    """SpeakerEncoderLSTM._forward_one_step (model.py:429-435) forward AND backward on the library.  The forward runs in its
    own workspace (attention output, attention weights and activated gates stay there for sfb_speaker_encoder_step_bwd);
    the action embedding and the image features are data."""

    @staticmethod
    def forward(ctx_, run_cuda, names, drop_x, a, v, h0, c0, *params):
        (h1, c1), fwd_ws = run_cuda()
        ctx_.names, ctx_.drop_x, ctx_.fwd_ws = names, drop_x, fwd_ws
        ctx_.save_for_backward(a, v, h0, c0, c1, *params)
        return h1, c1

    @staticmethod
    def backward(ctx_, g_h1, g_c1):
        from . import ops
        a, v, h0, c0, c1, *params = ctx_.saved_tensors
        w = {k: p.detach() for k, p in zip(ctx_.names, params)}
        grads = {k: torch.zeros_like(p) for k, p in w.items() if not k.startswith("encoder2decoder.")}
        with torch.no_grad():
            d_h0, d_c0 = ops.speaker_encoder_step_bwd(w, a.detach(), v.detach(), h0.detach(), c0.detach(), ctx_.drop_x, c1, ctx_.fwd_ws,
                                                      g_h1, g_c1, grads, accumulate=False)
        need = ctx_.needs_input_grad
        gpar = [(grads.get(k) if nd else None) for k, nd in zip(ctx_.names, need[7:])]
        return (None, None, None, None, None, d_h0 if need[5] else None, d_c0 if need[6] else None, *gpar)


class SpeakerDecStepKernelFn(torch.autograd.Function):
    """SpeakerDecoderLSTM.forward (model.py:487-519, default branch) forward AND backward on the library.  The word
    embedding is frozen GloVe in the reference configuration and receives no gradient here."""

    @staticmethod
    def forward(ctx_, run_cuda, names, prev_word, ctx_mask, drop_e, drop_h, h0, c0, ctx, *params):
        (h1, c1, alpha, logit), fwd_ws = run_cuda()
        ctx_.names, ctx_.prev_word, ctx_.ctx_mask, ctx_.drop_e, ctx_.drop_h, ctx_.fwd_ws = names, prev_word, ctx_mask, drop_e, drop_h, fwd_ws
        ctx_.save_for_backward(h0, c0, ctx, c1, alpha, *params)
        ctx_.mark_non_differentiable(alpha)
        return h1, c1, alpha, logit

    @staticmethod
    def backward(ctx_, g_h1, g_c1, g_alpha, g_logit):
        from . import ops
        h0, c0, ctx, c1, alpha, *params = ctx_.saved_tensors
        w = {k: p.detach() for k, p in zip(ctx_.names, params)}
        grads = {k: torch.zeros_like(p) for k, p in w.items() if k != "embedding.weight"}
        with torch.no_grad():
            d_h0, d_c0, d_ctx = ops.speaker_decoder_step_bwd(w, ctx_.prev_word, h0.detach(), c0.detach(), ctx.detach(), ctx_.ctx_mask,
                                                             ctx_.drop_e, ctx_.drop_h, c1, alpha, ctx_.fwd_ws, g_h1, g_c1, g_logit,
                                                             grads, accumulate=False)
        need = ctx_.needs_input_grad
        gin = [d_h0 if need[6] else None, d_c0 if need[7] else None, d_ctx if need[8] else None]
        gpar = [(grads.get(k) if nd else None) for k, nd in zip(ctx_.names, need[9:])]
        return (None, None, None, None, None, None, *gin, *gpar)


class FollowerStepFn(torch.autograd.Function):
    """Forward on the CUDA kernels, backward by torch autograd over the restatement above (same inputs, same masks).
    Kept as the independent check of FollowerStepKernelFn in tests/test_gpu_train.py."""

    @staticmethod
    def forward(ctx_, run_cuda, names, n_in, ctx_mask, drop_x, drop_h, *tensors):
        inputs, params = tensors[:n_in], tensors[n_in:]
        outs = run_cuda()
        ctx_.names, ctx_.n_in = names, n_in
        ctx_.ctx_mask, ctx_.drop_x, ctx_.drop_h = ctx_mask, drop_x, drop_h
        ctx_.save_for_backward(*inputs, *params)
        ctx_.mark_non_differentiable(outs[2], outs[4])     # the reference never differentiates through alpha / alpha_v
        return tuple(outs)

    @staticmethod
    def backward(ctx_, g_h1, g_c1, g_alpha, g_logit, g_alpha_v):
        saved = ctx_.saved_tensors
        n_in = ctx_.n_in
        with torch.enable_grad():
            leaves = [t.detach().requires_grad_(t.is_floating_point()) for t in saved]
            w = dict(zip(ctx_.names, leaves[n_in:]))
            u, U, V, h0, c0, c = leaves[:n_in]
            h1, c1, _, logit, _ = follower_step(w, u, U, V, h0, c0, c, ctx_.ctx_mask, ctx_.drop_x, ctx_.drop_h)
            outs, gouts = [], []
            for o, g in ((h1, g_h1), (c1, g_c1), (logit, g_logit)):
                if g is not None:
                    outs.append(o); gouts.append(g)
            need = [i for i, (t, nd) in enumerate(zip(leaves, ctx_.needs_input_grad[6:])) if nd and t.requires_grad]
            grads = torch.autograd.grad(outs, [leaves[i] for i in need], gouts, allow_unused=True) if need and outs else []
        full = [None] * len(saved)
        for i, g in zip(need, grads):
            full[i] = g
        return (None, None, None, None, None, None, *full)


def _visual_attention(w, h, V):
    """VisualSoftDotAttention.forward — model.py:310-326."""
    t = F.linear(h, w["visual_attention_layer.linear_in_h.weight"], w["visual_attention_layer.linear_in_h.bias"])
    k = F.linear(V, w["visual_attention_layer.linear_in_v.weight"], w["visual_attention_layer.linear_in_v.bias"])
    a = torch.softmax(torch.bmm(k, t.unsqueeze(2)).squeeze(2), dim=1)
    return torch.bmm(a.unsqueeze(1), V).squeeze(1)


def _lstm_cell(w, x, h0, c0):
    gates = F.linear(x, w["lstm.weight_ih"], w["lstm.bias_ih"]) + F.linear(h0, w["lstm.weight_hh"], w["lstm.bias_hh"])
    i, f, g, o = gates.chunk(4, 1)
    c1 = torch.sigmoid(f) * c0 + torch.sigmoid(i) * torch.tanh(g)
    return torch.sigmoid(o) * torch.tanh(c1), c1


def speaker_encoder_step(w, action_embedding, V, h0, c0, drop_x):
    """SpeakerEncoderLSTM._forward_one_step — model.py:429-435 -> (h_1, c_1)."""
    x = torch.cat((action_embedding, _visual_attention(w, h0, V)), 1)
    if drop_x is not None:
        x = x * drop_x
    return _lstm_cell(w, x, h0, c0)


def speaker_decoder_step(w, prev_word, h0, c0, ctx, ctx_mask, drop_e, drop_h):
    """SpeakerDecoderLSTM.forward (default branch) — model.py:497-503,515-519 -> (h_1, c_1, alpha, logit)."""
    e = F.embedding(prev_word.view(-1).long(), w["embedding.weight"])
    if drop_e is not None:
        e = e * drop_e
    h1, c1 = _lstm_cell(w, e, h0, c0)
    hd = h1 * drop_h if drop_h is not None else h1
    target = F.linear(hd, w["attention_layer.linear_in.weight"]).unsqueeze(2)
    attn = torch.bmm(ctx, target).squeeze(2)
    if ctx_mask is not None:
        attn = attn.masked_fill(ctx_mask.bool(), -float("inf"))
    alpha = torch.softmax(attn, dim=1)
    weighted = torch.bmm(alpha.unsqueeze(1), ctx).squeeze(1)
    h_tilde = torch.tanh(F.linear(torch.cat((weighted, hd), 1), w["attention_layer.linear_out.weight"]))
    return h1, c1, alpha, F.linear(h_tilde, w["decoder2action.weight"], w["decoder2action.bias"])


class RecomputeFn(torch.autograd.Function):
    """Generic form of FollowerStepFn: forward = `run_cuda()` (sm_100a kernels, no graph), backward = torch autograd over
    `restate(*leaves)` evaluated on the saved tensors.  `nondiff` = indices of outputs that carry no gradient."""

    @staticmethod
    def forward(ctx_, run_cuda, restate, nondiff, *tensors):
        outs = run_cuda()
        ctx_.restate, ctx_.nondiff = restate, set(nondiff)
        ctx_.save_for_backward(*tensors)
        ctx_.mark_non_differentiable(*[outs[i] for i in nondiff])
        return tuple(outs)

    @staticmethod
    def backward(ctx_, *gouts):
        saved = ctx_.saved_tensors
        with torch.enable_grad():
            leaves = [t.detach().requires_grad_(t.is_floating_point()) for t in saved]
            res = ctx_.restate(*leaves)
            outs, gs = [], []
            for i, (o, g) in enumerate(zip(res, gouts)):
                if g is not None and i not in ctx_.nondiff:
                    outs.append(o); gs.append(g)
            need = [i for i, (t, nd) in enumerate(zip(leaves, ctx_.needs_input_grad[3:])) if nd and t.requires_grad]
            grads = torch.autograd.grad(outs, [leaves[i] for i in need], gs, allow_unused=True) if need and outs else []
        full = [None] * len(saved)
        for i, g in zip(need, grads):
            full[i] = g
        return (None, None, None, *full)


def tail_loss_terms(logit: Tensor, is_valid: Tensor, target: Tensor) -> Tensor:
    """Per-row cross-entropy terms of follower.py:477,481 (0 where target < 0), differentiable w.r.t. logit."""
    masked = logit.masked_fill(is_valid == 0, -float("inf"))
    logp = torch.log_softmax(masked, dim=1)
    keep = target >= 0
    picked = -logp.gather(1, target.clamp(min=0).long().unsqueeze(1)).squeeze(1)
    return torch.where(keep, picked, torch.zeros_like(picked))
