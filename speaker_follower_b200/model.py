"""Drop-in module layer: the classes of the reference's tasks/R2R/model.py, same names, constructor and
forward signatures and state_dict keys, with every forward running on the sm_100a library through the C ABI.

    EncoderLSTM          model.py:43-104      AttnDecoderLSTM      model.py:355-397
    SoftDotAttention     model.py:107-143     SpeakerEncoderLSTM   model.py:405-457
    VisualSoftDotAttention model.py:300-326   SpeakerDecoderLSTM   model.py:460-519
    EltwiseProdScoring   model.py:329-352

Parameters live in ordinary nn.Linear / nn.LSTMCell / nn.LSTM / nn.Embedding sub-modules so that
``state_dict()`` / ``load_state_dict()`` interoperate with the reference's snapshots (follower.py:1022-1035);
those sub-modules are parameter containers only — their own forward is never called.  The library reads the
parameter storage in place, or from packed tcgen05 operand copies that are refreshed whenever a parameter's
storage or version changes (``invalidate_packed()`` forces it, e.g. after writes through ``.data``).

Under ``torch.no_grad()`` the forward runs on the sm_100a kernels alone.  While autograd is recording, the
decoder / encoder modules stay differentiable (``_functional.py``: forward on the kernels, backward as documented
there); the bare attention sub-modules are forward-only on their own.  Dropout in ``train()`` mode is applied with
masks drawn here (torch RNG) and passed to the kernels, exactly where the reference applies nn.Dropout.  Tensors
must be CUDA: there is no CPU path in this package.
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn

from . import _functional as Fn
from . import ops


def _no_autograd(*tensors):
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors):
        raise NotImplementedError(
            "speaker_follower_b200: this attention sub-module is forward-only when called on its own; call it under "
            "torch.no_grad(), or go through AttnDecoderLSTM / SpeakerDecoderLSTM, which are differentiable")


def _sd(module: nn.Module):
    """name -> parameter tensor (no copies), reference state_dict key names."""
    return {k: v for k, v in module.named_parameters()}


class _PackedWeights:
    """Mixin of the modules that keep packed tcgen05 operand copies of their weights (``self._packer``)."""

    def invalidate_packed(self) -> None:
        """Re-pack on the next call.  Needed only after parameter writes that bypass autograd's version counter
        (``p.data.copy_()``, EMA swaps, old-style optimizers); ordinary in-place updates are detected."""
        self._packer.invalidate()

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self._packer.invalidate()
        return out


class _Drop:
    """nn.Dropout semantics with an explicit, kernel-visible mask (scaled keep mask or None)."""

    def __init__(self, p: float):
        self.p = p

    def mask(self, training: bool, shape, device) -> Optional[torch.Tensor]:
        if not training or self.p <= 0.0:
            return None
        keep = (torch.rand(shape, device=device) >= self.p).to(torch.float32)
        return keep / (1.0 - self.p)


class EncoderLSTM(nn.Module):
    """model.py:43-104."""

    def __init__(self, vocab_size, embedding_size, hidden_size, padding_idx, dropout_ratio, bidirectional=False,
                 num_layers=1, glove=None):
        super().__init__()
        assert num_layers == 1, "the reference only ever uses one layer (train.py:197)"
        self.embedding_size = embedding_size
        self.hidden_size = hidden_size
        self.drop = nn.Dropout(p=dropout_ratio)
        self._drop = _Drop(dropout_ratio)
        self.num_directions = 2 if bidirectional else 1
        self.num_layers = num_layers
        self.embedding = nn.Embedding(vocab_size, embedding_size, padding_idx)
        self.use_glove = glove is not None
        if self.use_glove:
            print("Using GloVe embedding")
            self.embedding.weight.data[...] = torch.from_numpy(glove)
            self.embedding.weight.requires_grad = False
        self.lstm = nn.LSTM(embedding_size, hidden_size, self.num_layers, batch_first=True,
                            dropout=0.0, bidirectional=bidirectional)
        self.encoder2decoder = nn.Linear(hidden_size * self.num_directions, hidden_size * self.num_directions)

    def forward(self, inputs, lengths):
        """inputs [B, seq_len] int64 (length-sorted, model.py:89), lengths list -> (ctx, decoder_init, c_t)."""
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            return self._forward_autograd(inputs, lengths)
        B = inputs.size(0)
        maxlen = int(max(int(x) for x in lengths))
        drop_e = None
        if not self.use_glove:
            drop_e = self._drop.mask(self.training, (B * maxlen, self.embedding_size), inputs.device)
        ctx, decoder_init, c_t = ops.encoder_lstm(_sd(self), inputs, lengths, self.num_directions == 2, drop_e)
        m = self._drop.mask(self.training, ctx.shape, ctx.device)
        if m is not None:
            ctx = ctx * m                                     # model.py:102
        return ctx, decoder_init, c_t


    def _forward_autograd(self, inputs, lengths):
        """Training.  Unidirectional (the reference default, train.py:194) with frozen GloVe: forward on the sm_100a kernels
        with a tape, backward by the hand-written BPTT (sfb_encoder_lstm_bwd) through _functional.EncoderLstmKernelFn.
        Otherwise (bidirectional, or a trainable embedding table): torch's own LSTM, model.py:81-104 verbatim."""
        if self.num_directions == 1 and not self.embedding.weight.requires_grad:
            B = inputs.size(0)
            maxlen = int(max(int(x) for x in lengths))
            drop_e = None
            if not self.use_glove:
                drop_e = self._drop.mask(self.training, (B * maxlen, self.embedding_size), inputs.device)
            sd = _sd(self)
            names = list(sd.keys())
            ctx, decoder_init, c_t = Fn.EncoderLstmKernelFn.apply(inputs, [int(x) for x in lengths], drop_e, names, *[sd[k] for k in names])
            m = self._drop.mask(self.training, ctx.shape, ctx.device)
            return (ctx * m if m is not None else ctx), decoder_init, c_t
        embeds = self.embedding(inputs)
        if not self.use_glove:
            embeds = self.drop(embeds)
        packed = nn.utils.rnn.pack_padded_sequence(embeds, [int(x) for x in lengths], batch_first=True)
        enc_h, (enc_h_t, enc_c_t) = self.lstm(packed)
        if self.num_directions == 2:
            h_t = torch.cat((enc_h_t[-1], enc_h_t[-2]), 1)
            c_t = torch.cat((enc_c_t[-1], enc_c_t[-2]), 1)
        else:
            h_t, c_t = enc_h_t[-1], enc_c_t[-1]
        decoder_init = torch.tanh(self.encoder2decoder(h_t))
        ctx, _ = nn.utils.rnn.pad_packed_sequence(enc_h, batch_first=True)
        return self.drop(ctx), decoder_init, c_t


class SoftDotAttention(nn.Module):
    """model.py:107-143."""

    def __init__(self, dim):
        super().__init__()
        self.linear_in = nn.Linear(dim, dim, bias=False)
        self.linear_out = nn.Linear(dim * 2, dim, bias=False)

    def forward(self, h, context, mask=None):
        _no_autograd(h, context, *self.parameters())
        w = {"a." + k: v for k, v in self.named_parameters()}
        return ops.soft_dot_attention(w, "a.", h.contiguous(), context.contiguous(), mask)


class VisualSoftDotAttention(nn.Module):
    """model.py:300-326 (``mask`` is ignored there too)."""

    def __init__(self, h_dim, v_dim, dot_dim=256):
        super().__init__()
        self.linear_in_h = nn.Linear(h_dim, dot_dim, bias=True)
        self.linear_in_v = nn.Linear(v_dim, dot_dim, bias=True)

    def forward(self, h, visual_context, mask=None):
        _no_autograd(h, visual_context, *self.parameters())
        w = {"visual_attention_layer." + k: v for k, v in self.named_parameters()}
        return ops.visual_attention(w, h.contiguous(), visual_context.contiguous())


class EltwiseProdScoring(nn.Module):
    """model.py:329-352.  Inside AttnDecoderLSTM the scoring runs fused with the step; called on its own it is the
    re-associated three-launch form (``mask`` is ignored by the reference too, model.py:342)."""

    def __init__(self, h_dim, a_dim, dot_dim=256):
        super().__init__()
        self.linear_in_h = nn.Linear(h_dim, dot_dim, bias=True)
        self.linear_in_a = nn.Linear(a_dim, dot_dim, bias=True)
        self.linear_out = nn.Linear(dot_dim, 1, bias=True)

    def forward(self, h, all_u_t, mask=None):
        _no_autograd(h, all_u_t, *self.parameters())
        w = {"s." + k: v for k, v in self.named_parameters()}
        return ops.eltwise_prod_scoring(w, h.contiguous(), all_u_t.contiguous(), prefix="s.")


class AttnDecoderLSTM(_PackedWeights, nn.Module):
    """model.py:355-397: one follower decode step per call."""

    def __init__(self, embedding_size, hidden_size, dropout_ratio, feature_size=2048 + 128, image_attention_layers=None):
        super().__init__()
        self.embedding_size = embedding_size
        self.feature_size = feature_size
        self.hidden_size = hidden_size
        self.register_buffer("_u_begin", torch.zeros(embedding_size), persistent=False)
        self.drop = nn.Dropout(p=dropout_ratio)
        self._drop = _Drop(dropout_ratio)
        self.lstm = nn.LSTMCell(embedding_size + feature_size, hidden_size)
        self.visual_attention_layer = VisualSoftDotAttention(hidden_size, feature_size)
        self.text_attention_layer = SoftDotAttention(hidden_size)
        self.decoder2action = EltwiseProdScoring(hidden_size, embedding_size)
        self.feature_store: Optional[ops.FeatureStore] = None   # set to gather slabs on the device (K13)
        self._packer = ops.PackedFollower()   # bf16 hi/lo tcgen05 operand copies, refreshed when a weight changes

    @property
    def u_begin(self):
        """zeros[E] on the module's device (model.py:368; read at follower.py:356,462,563,744)."""
        return self._u_begin

    def forward(self, u_t_prev, all_u_t, visual_context, h_0, c_0, ctx, ctx_mask=None):
        """-> (h_1, c_1, alpha, logit, alpha_v).  ``visual_context`` is either the dense [B,36,F] tensor of the
        reference or a ``(vp_idx, view_idx)`` pair of int tensors when ``self.feature_store`` is set."""
        return self.decode_step(u_t_prev, all_u_t, visual_context, h_0, c_0, ctx, ctx_mask)

    def project_ctx(self, ctx):
        """Per-episode (ctx W_in, ctx W_out_c^T) for decode_step(ctx_proj=...); None when the packed path is off."""
        sd = _sd(self)
        packed = self._packer.get(sd)
        return ops.follower_project_ctx(sd, packed, ctx.contiguous()) if packed is not None else None

    def decode_step(self, u_t_prev, all_u_t, visual_context, h_0, c_0, ctx, ctx_mask=None, tail=None, carry_in=None,
                    carry_out=None, ctx_proj=None, cand_view=None, cand_trig=None, out=None, workspace=None,
                    idx_dependent=False):
        """forward() plus the fast-path extras of the packed C ABI (include/sf_b200.h): ``tail`` fuses the rollout
        tail (follower.py:476-505) behind the logits, ``carry_in``/``carry_out`` (``new_carry()`` buffers) hand the next step's visual query and packed gate
        operand across steps,
        ``ctx_proj`` = project_ctx(ctx) takes the text-side projections off the step's dependency chain."""
        B = h_0.shape[0]
        dev = h_0.device
        drop_x = self._drop.mask(self.training, (B, self.embedding_size + self.feature_size), dev)
        drop_h = self._drop.mask(self.training, (B, self.hidden_size), dev)
        u_t_prev = u_t_prev.contiguous()        # u_begin.expand(B, -1) is a stride-0 view (follower.py:462)
        sd = _sd(self)
        if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (u_t_prev, all_u_t, h_0, c_0, ctx, *sd.values())):
            return self._decode_step_autograd(sd, u_t_prev, all_u_t, visual_context, h_0, c_0, ctx, ctx_mask, drop_x, drop_h)
        packed = self._packer.get(sd)           # None for dimensions the packed path does not cover
        extra = {}
        if packed is not None:
            extra = dict(packed=packed, tail=tail, carry_in=carry_in, carry_out=carry_out, ctx_proj=ctx_proj)
        elif tail is not None or carry_in is not None or carry_out is not None or ctx_proj is not None:
            raise NotImplementedError("fused tail / carried query need the packed path (H % 128 == 0)")
        if isinstance(visual_context, (tuple, list)):
            vp, view = visual_context
            if all_u_t is None:   # action candidates gathered on the device too (env.py:60-75): view index + 4 trig values
                extra.update(cand_view=cand_view, cand_trig=cand_trig, out=out, workspace=workspace, idx_dependent=idx_dependent)
            return ops.follower_step(sd, u_t_prev, None if all_u_t is None else all_u_t.contiguous(), None, h_0.contiguous(),
                                     c_0.contiguous(), ctx.contiguous(), ctx_mask, drop_x, drop_h,
                                     store=self.feature_store, vp_idx=vp, view_idx=view, **extra)
        return ops.follower_step(sd, u_t_prev, all_u_t.contiguous(), visual_context.contiguous(),
                                 h_0.contiguous(), c_0.contiguous(), ctx.contiguous(), ctx_mask, drop_x, drop_h, **extra)

    def new_carry(self, B, device=None):
        """A state buffer for decode_step(carry_in=, carry_out=); allocate two and ping-pong them over a rollout."""
        return ops.follower_carry(_sd(self), B, device)

    def _decode_step_autograd(self, sd, u_t_prev, all_u_t, visual_context, h_0, c_0, ctx, ctx_mask, drop_x, drop_h):
        """Training: forward on the CUDA kernels in a per-step workspace, backward on the hand-written kernels of
        backward.cu (sfb_follower_step_bwd) through _functional.FollowerStepKernelFn."""
        if isinstance(visual_context, (tuple, list)):
            vp, view = visual_context
            visual_context = self.feature_store.dense(vp, view)
        names = list(sd.keys())
        inputs = [u_t_prev, all_u_t.contiguous(), visual_context.contiguous(), h_0.contiguous(), c_0.contiguous(), ctx.contiguous()]
        params = [sd[k] for k in names]

        def run_cuda():
            with torch.no_grad():
                d = [t.detach() for t in inputs]
                w = {k: v.detach() for k, v in sd.items()}
                B, A = d[1].shape[0], d[1].shape[1]
                # the step's own workspace: its intermediates are what sfb_follower_step_bwd reads
                need = ops._lib.load().sfb_follower_step_workspace_bytes(ops.C.byref(ops.follower_dims(w, d[2].shape[1])), B, d[5].shape[1], A)
                fwd_ws = torch.zeros(need, dtype=torch.uint8, device=d[3].device)
                packed = self._packer.get(w)
                # per-episode projections of ctx (constant over the steps of a rollout): computed once per (ctx, weights)
                cproj = None
                if packed is not None:
                    key = (d[5].data_ptr(), d[5]._version, tuple(d[5].shape), self._packer.key)
                    if getattr(self, "_cproj_key", None) != key:
                        self._cproj = ops.follower_project_ctx(w, packed, d[5])
                        self._cproj_key = key
                    cproj = self._cproj
                outs = ops.follower_step(w, d[0], d[1], d[2], d[3], d[4], d[5], ctx_mask, drop_x, drop_h,
                                         packed=packed, workspace=fwd_ws, ctx_proj=cproj)
                return outs, fwd_ws
        return Fn.FollowerStepKernelFn.apply(run_cuda, names, len(inputs), ctx_mask, drop_x, drop_h, *inputs, *params)

    @property
    def supports_fused_step(self) -> bool:
        return self._packer.get(_sd(self)) is not None


class SpeakerEncoderLSTM(_PackedWeights, nn.Module):
    """model.py:405-457."""

    def __init__(self, action_embedding_size, world_embedding_size, hidden_size, dropout_ratio, bidirectional=False):
        super().__init__()
        assert not bidirectional, "Bidirectional is not implemented yet"
        self.action_embedding_size = action_embedding_size
        self.word_embedding_size = world_embedding_size
        self.hidden_size = hidden_size
        self.drop = nn.Dropout(p=dropout_ratio)
        self._drop = _Drop(dropout_ratio)
        self.visual_attention_layer = VisualSoftDotAttention(hidden_size, world_embedding_size)
        self.lstm = nn.LSTMCell(action_embedding_size + world_embedding_size, hidden_size)
        self.encoder2decoder = nn.Linear(hidden_size, hidden_size)
        self._packer = ops.PackedVisLstm()      # tcgen05 operand copies, refreshed when a weight changes
        self._ws = None                         # step workspace reused across the path steps of a call
        self.feature_store = None               # ops.FeatureStore: world_state_embeddings may then be (vp_idx, view_idx) pairs

    def forward(self, batched_action_embeddings: List[torch.Tensor], world_state_embeddings: List[torch.Tensor], step_masks=None):
        """model.py:437-457.  Besides the reference's dense [N,36,F] tensors, a path step's world state may be given as a
        (viewpoint row, view index) pair of int32 tensors when `feature_store` is set: the kernel gathers the slab from the
        device table.  `step_masks` (optional, per step [N, E+F] or None): multiplied into the LSTM input like a dropout mask
        — an all-zero row makes that row's step see zero features and a zero action, which is how the reference's
        zero-padded path steps behave (speaker.py:87-95)."""
        assert isinstance(batched_action_embeddings, list)
        assert isinstance(world_state_embeddings, list)
        assert len(batched_action_embeddings) == len(world_state_embeddings)
        w = _sd(self)
        grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        names = list(w.keys())
        packed = self._packer.get({k: v.detach() for k, v in w.items()})
        B = batched_action_embeddings[0].shape[0]
        dev = batched_action_embeddings[0].device
        h = torch.zeros(B, self.hidden_size, device=dev)
        c = torch.zeros(B, self.hidden_size, device=dev)
        hs = []
        for t, (a, v) in enumerate(zip(batched_action_embeddings, world_state_embeddings)):
            dx = self._drop.mask(self.training, (B, self.action_embedding_size + self.word_embedding_size), dev)
            if step_masks is not None and step_masks[t] is not None:
                dx = step_masks[t] if dx is None else dx * step_masks[t]
            a = a.contiguous()
            if isinstance(v, (tuple, list)):     # gathered from the device-resident store (inference paths)
                assert self.feature_store is not None and not grad, "index-addressed world states need feature_store (eval)"
                h, c = ops.speaker_encoder_step(w, a, None, h, c, dx, store=self.feature_store, vp_idx=v[0], view_idx=v[1], packed=packed)
                hs.append(h)
                continue
            v = v.contiguous()
            if grad:   # forward AND backward on the CUDA kernels (sfb_speaker_encoder_step_bwd); a per-step workspace is the tape
                def run_cuda(a=a, v=v, h=h, c=c, dx=dx):
                    with torch.no_grad():
                        wd = {k: t.detach() for k, t in w.items()}
                        need = ops._lib.load().sfb_follower_step_workspace_bytes(ops.C.byref(ops.follower_dims(wd, v.shape[1])), B, 1, 1)
                        fwd_ws = torch.zeros(need, dtype=torch.uint8, device=dev)
                        return ops.speaker_encoder_step(wd, a, v, h.detach(), c.detach(), dx, packed=packed, workspace=fwd_ws), fwd_ws
                h, c = Fn.SpeakerEncStepKernelFn.apply(run_cuda, names, dx, a, v, h, c, *[w[k] for k in names])
            else:
                h, c = ops.speaker_encoder_step(w, a, v, h, c, dx, packed=packed)
            hs.append(h)
        # decoder_init = tanh(encoder2decoder(h_T)) (model.py:453): one [B,H]x[H,H] product per path, host plumbing
        decoder_init = torch.tanh(torch.addmm(self.encoder2decoder.bias, h, self.encoder2decoder.weight.t()))
        ctx = torch.stack(hs, dim=1)
        m = self._drop.mask(self.training, ctx.shape, dev)
        if m is not None:
            ctx = ctx * m
        return ctx, decoder_init, c


class SpeakerDecoderLSTM(_PackedWeights, nn.Module):
    """model.py:460-519, default branch (use_input_att_feed=False is what train_speaker.py:188-193 builds)."""

    def __init__(self, vocab_size, vocab_embedding_size, hidden_size, dropout_ratio, glove=None, use_input_att_feed=False):
        super().__init__()
        if use_input_att_feed:
            raise NotImplementedError("use_input_att_feed=True is never constructed by the reference's CLIs")
        self.vocab_size = vocab_size
        self.vocab_embedding_size = vocab_embedding_size
        self.hidden_size = hidden_size
        self.embedding = nn.Embedding(vocab_size, vocab_embedding_size)
        self.use_glove = glove is not None
        if self.use_glove:
            print("Using GloVe embedding")
            self.embedding.weight.data[...] = torch.from_numpy(glove)
            self.embedding.weight.requires_grad = False
        self.drop = nn.Dropout(p=dropout_ratio)
        self._drop = _Drop(dropout_ratio)
        self.use_input_att_feed = use_input_att_feed
        self.lstm = nn.LSTMCell(vocab_embedding_size, hidden_size)
        self.attention_layer = SoftDotAttention(hidden_size)
        self.decoder2action = nn.Linear(hidden_size, vocab_size)
        self._packer = ops.PackedSpeakerDecoder()   # tcgen05 operand copies, refreshed when a weight changes

    def forward(self, previous_word, h_0, c_0, ctx, ctx_mask=None):
        B = h_0.shape[0]
        dev = h_0.device
        drop_e = None if self.use_glove else self._drop.mask(self.training, (B, self.vocab_embedding_size), dev)
        drop_h = self._drop.mask(self.training, (B, self.hidden_size), dev)
        w = _sd(self)
        h_0, c_0, ctx = h_0.contiguous(), c_0.contiguous(), ctx.contiguous()
        packed = self._packer.get({k: v.detach() for k, v in w.items()})
        if torch.is_grad_enabled() and any(t.requires_grad for t in (h_0, c_0, ctx, *w.values())):
            names = list(w.keys())
            if w["embedding.weight"].requires_grad:
                # a trainable embedding (no GloVe) needs a scatter-add the library does not have: differentiate through the
                # torch restatement, forward still on the kernels
                def run_cuda():
                    with torch.no_grad():
                        return ops.speaker_decoder_step({k: t.detach() for k, t in w.items()}, previous_word, h_0.detach(),
                                                        c_0.detach(), ctx.detach(), ctx_mask, drop_e, drop_h, packed=packed)

                def restate(h_, c_, x_, *ps):
                    return Fn.speaker_decoder_step(dict(zip(names, ps)), previous_word, h_, c_, x_, ctx_mask, drop_e, drop_h)
                return Fn.RecomputeFn.apply(run_cuda, restate, (2,), h_0, c_0, ctx, *[w[k] for k in names])

            # forward AND backward on the CUDA kernels (sfb_speaker_decoder_step_bwd); a per-step workspace is the tape
            def run_cuda():
                with torch.no_grad():
                    T = ctx.shape[1]
                    need = ops._lib.load().sfb_speaker_decoder_step_workspace_bytes(self.hidden_size, self.vocab_embedding_size, B, T)
                    fwd_ws = torch.zeros(need, dtype=torch.uint8, device=dev)
                    return ops.speaker_decoder_step({k: t.detach() for k, t in w.items()}, previous_word, h_0.detach(), c_0.detach(),
                                                    ctx.detach(), ctx_mask, drop_e, drop_h, packed=packed, workspace=fwd_ws), fwd_ws
            return Fn.SpeakerDecStepKernelFn.apply(run_cuda, names, previous_word, ctx_mask, drop_e, drop_h, h_0, c_0, ctx,
                                                   *[w[k] for k in names])
        return ops.speaker_decoder_step(w, previous_word, h_0, c_0, ctx, ctx_mask, drop_e, drop_h, packed=packed)
