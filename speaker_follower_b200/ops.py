"""Tensor-level wrappers over the C ABI (include/sf_b200.h): torch owns device memory and streams, the
library owns every kernel.  All functions require CUDA fp32 contiguous tensors and raise otherwise —
there is no CPU path in the product (the CPU restatement lives in oracle/ and is test-only).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch

from . import _lib
from ._lib import (Dims, EncoderWeights, ScoringWeights, SoftDotWeights, SpeakerDecoderWeights, VisLstmWeights,
                   VisualSource, check)

Tensor = torch.Tensor


def _p(t: Optional[Tensor], dtype=torch.float32, name="tensor", pinned_ok: bool = False) -> Optional[int]:
    if t is None:
        return None
    if pinned_ok and not t.is_cuda and t.is_pinned():
        # page-locked host memory is device-accessible at the same address (UVA): small step inputs / results can be
        # read / written by the kernels directly over PCIe instead of through a separate memcpy node
        if t.dtype != dtype or not t.is_contiguous():
            raise _lib.SfbError("%s must be contiguous %s" % (name, dtype))
        return t.data_ptr()
    if not t.is_cuda:
        raise _lib.SfbError("%s must be a CUDA tensor (sf_b200 has no CPU path)" % name)
    if t.dtype != dtype:
        raise _lib.SfbError("%s must have dtype %s, got %s" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise _lib.SfbError("%s must be contiguous" % name)
    return t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _first_cuda_tensor(args, kwargs):
    for a in list(args) + list(kwargs.values()):
        if isinstance(a, torch.Tensor) and a.is_cuda:
            return a
        if isinstance(a, dict):
            for v in a.values():
                if isinstance(v, torch.Tensor) and v.is_cuda:
                    return v
    return None


def _on_tensor_device(fn):
    """Run `fn` with the CUDA device of its first device tensor current: the library launches on the current device's
    stream and keeps per-device state (SM count, shared-memory attributes), so a process that drives several GPUs
    (speaker on cuda:0, follower on cuda:1) must not launch one device's tensors from another device's context."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        t = _first_cuda_tensor(args, kwargs)
        if t is None or t.device.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(t.device):
            return fn(*args, **kwargs)
    return wrapper


def _mask_u8(mask: Optional[Tensor]) -> Optional[Tensor]:
    if mask is None:
        return None
    if mask.dtype == torch.bool:
        return mask.contiguous().view(torch.uint8)
    if mask.dtype != torch.uint8:
        mask = mask.to(torch.uint8)
    return mask.contiguous()


def _i32(t: Optional[Tensor]) -> Optional[Tensor]:
    if t is None:
        return None
    return t.to(torch.int32).contiguous()


class FeatureStore:
    """Device-resident feature table + orientation table: the B200-native replacement of the per-step host
    np.stack + H2D copy of the 36-view slab (follower.py:291-298, env.py:330-332)."""

    def __init__(self, feat_table: Tensor, loc_table: Tensor):
        assert feat_table.dim() == 3 and loc_table.dim() == 3
        self.feat_table = feat_table.contiguous()
        self.loc_table = loc_table.contiguous()
        self.img_dim = feat_table.shape[2]

    @classmethod
    def from_tsv(cls, tsv_paths, loc_table: Tensor, device="cuda", num_views: int = 36, dim: int = 2048,
                 chunk_rows: int = 512):
        """Build the device table straight from the reference's precomputed feature files (env.py:350-375:
        tab-separated scanId, viewpointId, image_w, image_h, vfov, base64(float32[36*2048])); several files are
        concatenated along the feature axis like MeanPooledImageFeatures does (372-375).  Returns the store; its
        ``index`` maps "scanId_viewpointId" (env.py:377-378) to the table row, which is what an env puts into
        ``ob['vp_index']`` so that a step ships two integers instead of a 36 x 2176 slab.  Rows are decoded into a
        pinned staging buffer and uploaded in chunks: the 3 GB Python dict of the reference never exists."""
        import base64
        import csv
        import numpy as np
        if isinstance(tsv_paths, str):
            tsv_paths = [tsv_paths]
        tsv_paths = sorted(tsv_paths)
        fields = ["scanId", "viewpointId", "image_w", "image_h", "vfov", "features"]
        csv.field_size_limit(1 << 30)
        index, tables = {}, []
        for fi, path in enumerate(tsv_paths):
            rows, ids = [], []
            stage = torch.empty(chunk_rows, num_views, dim, dtype=torch.float32).pin_memory() if torch.cuda.is_available() \
                else torch.empty(chunk_rows, num_views, dim, dtype=torch.float32)
            parts, n_in = [], 0
            with open(path, "rt") as f:
                for item in csv.DictReader(f, delimiter="\t", fieldnames=fields):
                    buf = base64.b64decode(item["features"])
                    feat = np.frombuffer(buf, dtype=np.float32)
                    if feat.size != num_views * dim:
                        raise ValueError("%s: feature blob of %s_%s has %d floats, expected %d" % (
                            path, item["scanId"], item["viewpointId"], feat.size, num_views * dim))
                    stage[n_in].copy_(torch.from_numpy(feat.reshape(num_views, dim).copy()))
                    ids.append(item["scanId"] + "_" + item["viewpointId"])
                    n_in += 1
                    if n_in == chunk_rows:
                        parts.append(stage[:n_in].to(device, non_blocking=False).clone())
                        n_in = 0
            if n_in:
                parts.append(stage[:n_in].to(device).clone())
            table = torch.cat(parts, 0) if parts else torch.empty(0, num_views, dim, device=device)
            if fi == 0:
                index = {k: i for i, k in enumerate(ids)}
                if len(index) != len(ids):
                    raise ValueError("%s: duplicate scanId_viewpointId" % path)
            else:   # same viewpoints, possibly another order: align to the first file
                if set(ids) != set(index):
                    raise ValueError("%s: viewpoint set differs from %s" % (path, tsv_paths[0]))
                order = torch.empty(len(ids), dtype=torch.long)
                for i, k in enumerate(ids):
                    order[index[k]] = i
                table = table[order.to(table.device)]
            tables.append(table)
        store = cls(torch.cat(tables, 2) if len(tables) > 1 else tables[0], loc_table.to(device))
        store.index = index
        return store

    def rows(self, long_ids) -> Tensor:
        """int32 table rows of "scanId_viewpointId" ids (host list -> device tensor), for vp_idx."""
        return torch.tensor([self.index[k] for k in long_ids], dtype=torch.int32, device=self.feat_table.device)

    def dense(self, vp_idx: Tensor, view_idx: Tensor) -> Tensor:
        """Materialise [B,36,F] (what the reference builds on the host) — for tests only."""
        return torch.cat((self.feat_table[vp_idx.long()], self.loc_table[view_idx.long()]), dim=2).contiguous()


def _visual_source(visual, store: Optional[FeatureStore], vp_idx, view_idx, keep, idx_dependent: bool = False):
    vs = VisualSource()
    vs.idx_dependent = 1 if idx_dependent else 0
    if visual is not None:
        vs.visual = _p(visual, name="visual_context")
        keep.append(visual)
    else:
        vp, vw = _i32(vp_idx), _i32(view_idx)
        keep.extend([vp, vw])
        vs.visual = None
        vs.feat_table = _p(store.feat_table, name="feat_table")
        vs.loc_table = _p(store.loc_table, name="loc_table")
        vs.vp_idx = _p(vp, torch.int32, "vp_idx")
        vs.view_idx = _p(vw, torch.int32, "view_idx")
        vs.img_dim = store.img_dim
    return vs


def _vis_lstm_weights(w: Dict[str, Tensor]) -> VisLstmWeights:
    s = VisLstmWeights()
    s.lstm_w_ih = _p(w["lstm.weight_ih"]); s.lstm_w_hh = _p(w["lstm.weight_hh"])
    s.lstm_b_ih = _p(w["lstm.bias_ih"]); s.lstm_b_hh = _p(w["lstm.bias_hh"])
    s.va_w_h = _p(w["visual_attention_layer.linear_in_h.weight"])
    s.va_b_h = _p(w["visual_attention_layer.linear_in_h.bias"])
    s.va_w_v = _p(w["visual_attention_layer.linear_in_v.weight"])
    s.va_b_v = _p(w["visual_attention_layer.linear_in_v.bias"])
    return s


def _softdot_weights(w: Dict[str, Tensor], prefix: str) -> SoftDotWeights:
    s = SoftDotWeights()
    s.w_in = _p(w[prefix + "linear_in.weight"]); s.w_out = _p(w[prefix + "linear_out.weight"])
    return s


def _scoring_weights(w: Dict[str, Tensor], prefix="decoder2action.") -> ScoringWeights:
    s = ScoringWeights()
    s.w_h = _p(w[prefix + "linear_in_h.weight"]); s.b_h = _p(w[prefix + "linear_in_h.bias"])
    s.w_a = _p(w[prefix + "linear_in_a.weight"]); s.b_a = _p(w[prefix + "linear_in_a.bias"])
    s.w_out = _p(w[prefix + "linear_out.weight"]); s.b_out = _p(w[prefix + "linear_out.bias"])
    return s


def follower_dims(w: Dict[str, Tensor], V: int = 36) -> Dims:
    H = w["lstm.weight_hh"].shape[1]
    F = w["visual_attention_layer.linear_in_v.weight"].shape[1]
    E = w["lstm.weight_ih"].shape[1] - F
    D = w["visual_attention_layer.linear_in_h.weight"].shape[0]
    return Dims(E, F, H, D, V)


_WS_CACHE: "Dict[tuple, Tensor]" = {}
_WS_CACHE_MAX = 24


def _workspace(nbytes: int, device, sig: tuple = ()) -> Tensor:
    """Workspace for one (entry point, shape) signature on one device, zero-filled ONCE when it is allocated and then
    reused by every later call with the same signature (no allocation, no memset per step).  The library's barrier
    words sit at shape-dependent offsets and return to zero after every launch, so a buffer must not be shared
    between different layouts — hence the signature in the key.  One stream at a time per device; callers that
    overlap streams pass their own `workspace=`."""
    dev = torch.device(device)
    idx = dev.index if dev.index is not None else (torch.cuda.current_device() if dev.type == "cuda" else -1)
    key = (dev.type, idx, int(nbytes)) + tuple(sig)
    ws = _WS_CACHE.get(key)
    if ws is None:
        if len(_WS_CACHE) >= _WS_CACHE_MAX:
            _WS_CACHE.pop(next(iter(_WS_CACHE)))
        ws = torch.zeros(max(int(nbytes), 256), dtype=torch.uint8, device=dev)
        _WS_CACHE[key] = ws
    return ws


_FOLLOWER_KEYS = ("lstm.weight_ih", "lstm.weight_hh", "lstm.bias_ih", "lstm.bias_hh",
                  "visual_attention_layer.linear_in_h.weight", "visual_attention_layer.linear_in_h.bias",
                  "visual_attention_layer.linear_in_v.weight", "text_attention_layer.linear_in.weight",
                  "text_attention_layer.linear_out.weight", "decoder2action.linear_in_h.weight",
                  "decoder2action.linear_in_h.bias", "decoder2action.linear_in_a.weight",
                  "decoder2action.linear_in_a.bias", "decoder2action.linear_out.weight",
                  "decoder2action.linear_out.bias")


class PackedFollower:
    """Device blob of sfb_follower_pack_weights, re-packed when any weight tensor's storage or version changes
    (shadow copies keyed on param._version)."""

    def __init__(self):
        self.blob: Optional[Tensor] = None
        self.key = None

    def invalidate(self) -> None:
        """Force a re-pack on the next get().  The cache key is (storage pointer, tensor version) per weight: in-place
        updates through ``.data`` (``p.data.copy_``, EMA swaps, old-style optimizers) do NOT bump the version, so code
        that writes weights that way must call this (the nn.Module layer does it from load_state_dict)."""
        self.key = None

    @staticmethod
    def _key(w):
        return tuple((w[k].data_ptr(), w[k]._version) for k in _FOLLOWER_KEYS)

    def get(self, w: Dict[str, Tensor], V: int = 36) -> Optional[Tensor]:
        key = self._key(w)
        if self.blob is not None and key == self.key:
            return self.blob
        lib = _lib.load()
        d = follower_dims(w, V)
        n = lib.sfb_follower_packed_bytes(C.byref(d))
        if n == 0:
            return None   # dimensions the packed path does not cover -> caller uses the in-place path
        dev = w["lstm.weight_ih"].device
        if self.blob is None or self.blob.numel() < n or self.blob.device != dev:
            self.blob = torch.empty(n, dtype=torch.uint8, device=dev)
        wl, wt, ws = _vis_lstm_weights(w), _softdot_weights(w, "text_attention_layer."), _scoring_weights(w)
        with torch.cuda.device(dev):
            check(lib.sfb_follower_pack_weights(C.byref(d), C.byref(wl), C.byref(wt), C.byref(ws), self.blob.data_ptr(),
                                                self.blob.numel(), _stream()))
        self.key = key
        return self.blob


class _PackedCache:
    """Blob re-packed when any of `keys`' storage or version changes (same policy as PackedFollower)."""
    keys: tuple = ()

    def __init__(self):
        self.blob: Optional[Tensor] = None
        self.key = None

    def _pack(self, w, blob_or_none):   # -> (nbytes, pack_fn)
        raise NotImplementedError

    def invalidate(self) -> None:
        """See PackedFollower.invalidate()."""
        self.key = None

    def get(self, w: Dict[str, Tensor]) -> Optional[Tensor]:
        key = tuple((w[k].data_ptr(), w[k]._version) for k in self.keys)
        if self.blob is not None and key == self.key:
            return self.blob
        n, pack = self._pack(w)
        if n == 0:
            return None
        dev = w[self.keys[0]].device
        if self.blob is None or self.blob.numel() < n or self.blob.device != dev:
            self.blob = torch.empty(n, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            pack(self.blob)
        self.key = key
        return self.blob


class PackedVisLstm(_PackedCache):
    """sfb_vis_lstm_pack_weights: visual attention + LSTM cell of SpeakerEncoderLSTM (model.py:415-418)."""
    keys = ("lstm.weight_ih", "lstm.weight_hh", "visual_attention_layer.linear_in_h.weight",
            "visual_attention_layer.linear_in_h.bias", "visual_attention_layer.linear_in_v.weight")

    def _pack(self, w):
        lib = _lib.load()
        d = follower_dims(w)
        wl = _vis_lstm_weights(w)
        n = lib.sfb_vis_lstm_packed_bytes(C.byref(d))
        return n, lambda blob: check(lib.sfb_vis_lstm_pack_weights(C.byref(d), C.byref(wl), blob.data_ptr(), blob.numel(), _stream()))


def _spk_dec_weights(w: Dict[str, Tensor]) -> SpeakerDecoderWeights:
    sw = SpeakerDecoderWeights()
    sw.embedding = _p(w["embedding.weight"])
    sw.lstm_w_ih = _p(w["lstm.weight_ih"]); sw.lstm_w_hh = _p(w["lstm.weight_hh"])
    sw.lstm_b_ih = _p(w["lstm.bias_ih"]); sw.lstm_b_hh = _p(w["lstm.bias_hh"])
    sw.attn = _softdot_weights(w, "attention_layer.")
    sw.w_voc = _p(w["decoder2action.weight"]); sw.b_voc = _p(w["decoder2action.bias"])
    return sw


class PackedSpeakerDecoder(_PackedCache):
    """sfb_speaker_decoder_pack_weights (model.py:469-485)."""
    keys = ("lstm.weight_ih", "lstm.weight_hh", "attention_layer.linear_in.weight", "attention_layer.linear_out.weight",
            "decoder2action.weight", "embedding.weight")   # the blob holds the per-token table Emb W_ih^T

    def _pack(self, w):
        lib = _lib.load()
        vocab, Ew = w["embedding.weight"].shape
        H = w["lstm.weight_hh"].shape[1]
        sw = _spk_dec_weights(w)
        n = lib.sfb_speaker_decoder_packed_bytes(H, Ew, vocab)
        return n, lambda blob: check(lib.sfb_speaker_decoder_pack_weights(C.byref(sw), H, Ew, vocab, blob.data_ptr(),
                                                                          blob.numel(), _stream()))


@_on_tensor_device
def follower_step(w: Dict[str, Tensor], u_prev: Tensor, all_u_t: Tensor, visual: Optional[Tensor], h0: Tensor,
                  c0: Tensor, ctx: Tensor, ctx_mask: Optional[Tensor], drop_x: Optional[Tensor] = None,
                  drop_h: Optional[Tensor] = None, store: Optional[FeatureStore] = None,
                  vp_idx: Optional[Tensor] = None, view_idx: Optional[Tensor] = None,
                  workspace: Optional[Tensor] = None, out: Optional[tuple] = None,
                  packed: Optional[Tensor] = None, carry_in: Optional[Tensor] = None, carry_out: Optional[Tensor] = None,
                  tail: Optional[dict] = None, cand_view: Optional[Tensor] = None, cand_trig: Optional[Tensor] = None,
                  ctx_proj: Optional[tuple] = None, idx_dependent: bool = False):
    """AttnDecoderLSTM.forward (model.py:377-397) -> (h1, c1, alpha, logit, alpha_v).
    `packed`: blob from PackedFollower.get(w) -> the packed-weight tcgen05 path (sfb_follower_step_packed_fwd).
    Packed path only: `carry_in` / `carry_out` (follower_carry() buffers) hand the state one step prepares for the
    next across the call boundary — the next visual query and the gate GEMM's packed [u_prev | . | h_0] operand blocks
    (see include/sf_b200.h); a step given the previous step's carry_out starts directly with the fused gather + LSTM
    launch; `carry_out` is complete only when `tail` is given (its u_next is the next step's u_prev);
    `tail` = dict(is_valid, feedback, target=None, sample_u=None, out=(a_t, u_next, score, ce)) fuses the rollout
    tail (follower.py:476-505) behind the logits; the outputs are left in tail["out"];
    `all_u_t=None` with `cand_view` [B,A] int32 / `cand_trig` [B,A,4] (+ store, vp_idx): action candidates gathered
    on the device from the feature table (env.py:60-75) instead of being shipped as a dense [B,A,E] tensor;
    `ctx_proj` = (ctx_k, ctx_o) from follower_project_ctx(): per-episode projections of ctx that take the text-side
    projections off the step's dependency chain; `idx_dependent`: the index tensors (vp_idx, view_idx, cand_view, ...)
    are produced by the kernel enqueued just before this call (nav_step) — no prefetch ahead of the dependency wait."""
    lib = _lib.load()
    L = ctx.shape[1]
    V = visual.shape[1] if visual is not None else store.feat_table.shape[1]
    d = follower_dims(w, V)
    if all_u_t is not None:
        B, A, E = all_u_t.shape
    else:
        if packed is None or cand_view is None or cand_trig is None or store is None or vp_idx is None:
            raise _lib.SfbError("all_u_t=None needs packed=, store=, vp_idx=, cand_view= and cand_trig=")
        (B, A), E = cand_view.shape, d.E
    dev = h0.device
    keep = []
    vs = _visual_source(visual, store, vp_idx, view_idx, keep, idx_dependent)
    mask = _mask_u8(ctx_mask)
    need = lib.sfb_follower_step_workspace_bytes(C.byref(d), B, L, A)
    if workspace is None or workspace.numel() < need:
        workspace = _workspace(need, dev, ("follower_step", B, L, A, packed is not None))
    if out is None:
        h1 = torch.empty(B, d.H, device=dev); c1 = torch.empty(B, d.H, device=dev)
        alpha = torch.empty(B, L, device=dev); logit = torch.empty(B, A, device=dev)
        alpha_v = torch.empty(B, V, device=dev)
    else:
        h1, c1, alpha, logit, alpha_v = out
    if packed is None and (carry_in is not None or carry_out is not None or tail is not None):
        raise _lib.SfbError("carry_in / carry_out / tail need the packed path (packed=PackedFollower.get(w))")
    if packed is not None:
        nc = lib.sfb_follower_carry_bytes(C.byref(d), B)
        for nm, cbuf in (("carry_in", carry_in), ("carry_out", carry_out)):
            if cbuf is not None and (not cbuf.is_cuda or cbuf.dtype != torch.float32 or not cbuf.is_contiguous() or cbuf.numel() * 4 < nc):
                raise _lib.SfbError("%s must be a follower_carry() buffer for this batch size" % nm)
    if packed is not None:
        wl = _vis_lstm_weights(w)
        tl = None
        if tail is not None:
            tgt = _i32(tail.get("target"))
            if tail.get("out") is None:
                tail["out"] = (torch.empty(B, dtype=torch.int32, device=dev), torch.empty(B, E, device=dev),
                               torch.empty(B, device=dev), torch.empty(B, device=dev) if tgt is not None else None)
            a_t, u_next, score, ce = tail["out"]
            tl = _lib.StepTail(_p(tail["is_valid"], name="is_valid"), _p(tgt, torch.int32, "target"),
                               {"teacher": 0, "argmax": 1, "sample": 2}[tail["feedback"]],
                               _p(tail.get("sample_u"), name="sample_u"), _p(a_t, torch.int32, "a_t", pinned_ok=True),
                               _p(u_next), _p(score),
                               _p(ce))
            keep.append(tgt)
        act = None
        if all_u_t is None:
            act = _lib.ActionSource(None, _p(store.feat_table), _p(_i32(vp_idx), torch.int32, "vp_idx"),
                                    _p(cand_view, torch.int32, "cand_view"), _p(cand_trig, name="cand_trig"),
                                    store.feat_table.shape[2])
        check(lib.sfb_follower_step_packed_fwd(
            C.byref(d), C.byref(wl), packed.data_ptr(), packed.numel(), B, L, A,
            _p(u_prev, name="u_t_prev"), _p(all_u_t, name="all_u_t"), C.byref(vs), _p(h0, name="h_0"),
            _p(c0, name="c_0"), _p(ctx, name="ctx"), _p(mask, torch.uint8, "ctx_mask"), _p(drop_x, name="drop_x"),
            _p(drop_h, name="drop_h"), _p(h1), _p(c1), _p(alpha), _p(logit), _p(alpha_v),
            carry_in.data_ptr() if carry_in is not None else None,
            carry_out.data_ptr() if carry_out is not None else None, C.byref(tl) if tl is not None else None,
            C.byref(act) if act is not None else None,
            _p(ctx_proj[0], name="ctx_k") if ctx_proj is not None else None,
            _p(ctx_proj[1], name="ctx_o") if ctx_proj is not None else None,
            workspace.data_ptr(), workspace.numel(), _stream()))
        return h1, c1, alpha, logit, alpha_v
    wl, wt, ws = _vis_lstm_weights(w), _softdot_weights(w, "text_attention_layer."), _scoring_weights(w)
    check(lib.sfb_follower_step_fwd(
        C.byref(d), C.byref(wl), C.byref(wt), C.byref(ws), B, L, A,
        _p(u_prev, name="u_t_prev"), _p(all_u_t, name="all_u_t"), C.byref(vs), _p(h0, name="h_0"), _p(c0, name="c_0"),
        _p(ctx, name="ctx"), _p(mask, torch.uint8, "ctx_mask"), _p(drop_x, name="drop_x"), _p(drop_h, name="drop_h"),
        _p(h1), _p(c1), _p(alpha), _p(logit), _p(alpha_v), workspace.data_ptr(), workspace.numel(), _stream()))
    return h1, c1, alpha, logit, alpha_v


def follower_workspace(w: Dict[str, Tensor], B: int, L: int, A: int, device=None) -> Tensor:
    """A zero-filled workspace for follower_step(workspace=...): one buffer serves every candidate count <= A of a
    rollout (the layout does not depend on A), so an agent allocates it once instead of once per distinct A."""
    d = follower_dims(w)
    n = _lib.load().sfb_follower_step_workspace_bytes(C.byref(d), B, L, A)
    dev = device if device is not None else w["lstm.weight_ih"].device
    return _workspace(n, dev, ("follower_rollout", B, L))


def follower_carry(w: Dict[str, Tensor], B: int, device=None) -> Tensor:
    """Opaque per-step state buffer for follower_step(carry_in=, carry_out=) (sfb_follower_carry_bytes); float32 so that
    carry_query() can view the visual query it starts with."""
    d = follower_dims(w)
    n = _lib.load().sfb_follower_carry_bytes(C.byref(d), B)
    if n == 0:
        raise _lib.SfbError("these dimensions have no packed path")
    dev = device if device is not None else w["lstm.weight_ih"].device
    return torch.zeros((n + 3) // 4, dtype=torch.float32, device=dev)


def carry_query(carry: Tensor, B: int, F: int) -> Tensor:
    """The [B, F] visual query M_q h + b_q stored at the head of a carry buffer."""
    return carry[:B * F].view(B, F)


_GRAD_FIELDS = {"lstm_w_ih": "lstm.weight_ih", "lstm_w_hh": "lstm.weight_hh", "lstm_b_ih": "lstm.bias_ih", "lstm_b_hh": "lstm.bias_hh",
                "va_w_h": "visual_attention_layer.linear_in_h.weight", "va_b_h": "visual_attention_layer.linear_in_h.bias",
                "va_w_v": "visual_attention_layer.linear_in_v.weight", "w_in": "text_attention_layer.linear_in.weight",
                "w_out": "text_attention_layer.linear_out.weight", "sc_w_h": "decoder2action.linear_in_h.weight",
                "sc_b_h": "decoder2action.linear_in_h.bias", "sc_w_a": "decoder2action.linear_in_a.weight",
                "sc_b_a": "decoder2action.linear_in_a.bias", "sc_w_out": "decoder2action.linear_out.weight",
                "sc_b_out": "decoder2action.linear_out.bias"}


@_on_tensor_device
def follower_step_bwd(w: Dict[str, Tensor], u_prev: Tensor, all_u_t: Tensor, visual: Optional[Tensor], h0: Tensor, c0: Tensor,
                      ctx: Tensor, ctx_mask: Optional[Tensor], drop_x: Optional[Tensor], drop_h: Optional[Tensor],
                      c1: Tensor, alpha: Tensor, alpha_v: Tensor, fwd_workspace: Tensor,
                      g_h1: Optional[Tensor], g_c1: Optional[Tensor], g_logit: Optional[Tensor],
                      grads: Dict[str, Tensor], accumulate: bool = True, store: Optional[FeatureStore] = None,
                      vp_idx=None, view_idx=None, cand_view=None, cand_trig=None):
    """sfb_follower_step_bwd: hand-written backward of one decode step.  `fwd_workspace` is the workspace tensor the
    forward call of this step ran in; `grads` maps reference state_dict names to gradient buffers (accumulated into).
    Returns (d_h0, d_c0, d_ctx)."""
    lib = _lib.load()
    B, L, H = ctx.shape
    V = visual.shape[1] if visual is not None else store.feat_table.shape[1]
    d = follower_dims(w, V)
    A = all_u_t.shape[1] if all_u_t is not None else cand_view.shape[1]
    keep = []
    vs = _visual_source(visual, store, vp_idx, view_idx, keep)
    if all_u_t is not None:
        act = _lib.ActionSource(_p(all_u_t, name="all_u_t"), None, None, None, None, 0)
    else:
        act = _lib.ActionSource(None, _p(store.feat_table), _p(_i32(vp_idx), torch.int32), _p(cand_view, torch.int32),
                                _p(cand_trig), store.feat_table.shape[2])
    wl, wt, wsc = _vis_lstm_weights(w), _softdot_weights(w, "text_attention_layer."), _scoring_weights(w)
    gs = _lib.FollowerGrads()
    for f, k in _GRAD_FIELDS.items():
        setattr(gs, f, _p(grads.get(k), name="grad " + k))
    dev = h0.device
    need = lib.sfb_follower_step_bwd_workspace_bytes(C.byref(d), B, L, A)
    ws = _workspace(need, dev, ("follower_step_bwd", B))
    d_h0 = torch.empty(B, H, device=dev); d_c0 = torch.empty(B, H, device=dev); d_ctx = torch.empty(B, L, H, device=dev)
    mask = _mask_u8(ctx_mask)
    gc = lambda t: None if t is None else t.contiguous()
    g_h1, g_c1, g_logit = gc(g_h1), gc(g_c1), gc(g_logit)
    check(lib.sfb_follower_step_bwd(
        C.byref(d), C.byref(wl), C.byref(wt), C.byref(wsc), B, L, A, _p(u_prev, name="u_t_prev"), C.byref(act), C.byref(vs),
        _p(h0, name="h_0"), _p(c0, name="c_0"), _p(ctx, name="ctx"), _p(mask, torch.uint8, "ctx_mask"), _p(drop_x), _p(drop_h),
        _p(c1, name="c_1"), _p(alpha, name="alpha"), _p(alpha_v, name="alpha_v"), fwd_workspace.data_ptr(),
        _p(g_h1), _p(g_c1), _p(g_logit), _p(d_h0), _p(d_c0), _p(d_ctx), C.byref(gs), 1 if accumulate else 0,
        ws.data_ptr(), ws.numel(), _stream()))
    return d_h0, d_c0, d_ctx


@_on_tensor_device
def follower_gather_lstm(w: Dict[str, Tensor], packed: Tensor, carry_in: Tensor, c0: Tensor, visual: Optional[Tensor] = None,
                         store: Optional[FeatureStore] = None, vp_idx=None, view_idx=None, out: Optional[tuple] = None,
                         workspace: Optional[Tensor] = None):
    """sfb_follower_gather_lstm_fwd: attention gather + gate GEMM + LSTM cell from carried state, one launch.
    -> (h1, c1, alpha_v)."""
    lib = _lib.load()
    B, H = c0.shape
    V = visual.shape[1] if visual is not None else store.feat_table.shape[1]
    d = follower_dims(w, V)
    keep = []
    vs = _visual_source(visual, store, vp_idx, view_idx, keep)
    wl = _vis_lstm_weights(w)
    need = lib.sfb_follower_step_workspace_bytes(C.byref(d), B, 1, 1)
    if workspace is None or workspace.numel() < need:
        workspace = _workspace(need, c0.device, ("gather_lstm", B))
    h1, c1, av = out if out is not None else (torch.empty_like(c0), torch.empty_like(c0), torch.empty(B, V, device=c0.device))
    check(lib.sfb_follower_gather_lstm_fwd(C.byref(d), C.byref(wl), packed.data_ptr(), packed.numel(), B, carry_in.data_ptr(),
                                           C.byref(vs), _p(c0, name="c_0"), _p(h1), _p(c1), None, _p(av),
                                           workspace.data_ptr(), workspace.numel(), _stream()))
    return h1, c1, av


def ctx_rows(lengths, L: int, device) -> Tensor:
    """Flat (b*L + l) indices of the un-padded positions of a padded [B, L] batch (int32, on `device`)."""
    idx = [b * L + l for b, n in enumerate(lengths) for l in range(min(int(n), L))]
    return torch.tensor(idx, dtype=torch.int32, device=device)


@_on_tensor_device
def follower_project_ctx(w: Dict[str, Tensor], packed: Tensor, ctx: Tensor, out: Optional[tuple] = None,
                         rows: Optional[Tensor] = None, workspace: Optional[Tensor] = None):
    """Per-episode (ctx_k, ctx_o) = (ctx W_in, ctx W_out_c^T) for follower_step(ctx_proj=...) — include/sf_b200.h.
    `rows` = ctx_rows(lengths, L, device): project only the un-padded positions (padded ones stay zero)."""
    lib = _lib.load()
    d = follower_dims(w)
    B, L, H = ctx.shape
    ctx_k, ctx_o = out if out is not None else (torch.zeros_like(ctx), torch.zeros_like(ctx))
    need = lib.sfb_follower_project_ctx_workspace_bytes(C.byref(d), B, L)
    ws = workspace if workspace is not None and workspace.numel() >= need else _workspace(need, ctx.device, ("project_ctx", B, L))
    check(lib.sfb_follower_project_ctx(C.byref(d), packed.data_ptr(), packed.numel(), B, L, _p(ctx, name="ctx"),
                                       _p(rows, torch.int32, "rows"), 0 if rows is None else rows.numel(),
                                       _p(ctx_k), _p(ctx_o), ws.data_ptr(), ws.numel(), _stream()))
    return ctx_k, ctx_o


@_on_tensor_device
def follower_tail(logit: Tensor, is_valid: Tensor, all_u_t: Tensor, feedback: str, target: Optional[Tensor] = None,
                  sample_u: Optional[Tensor] = None, out: Optional[tuple] = None):
    """follower.py:476-505 -> (a_t int32 [B], u_next [B,E], action_score [B], ce [B]); masks `logit` in place."""
    lib = _lib.load()
    B, A, E = all_u_t.shape
    dev = logit.device
    fb = {"teacher": 0, "argmax": 1, "sample": 2}[feedback]
    if out is None:
        a_t = torch.empty(B, dtype=torch.int32, device=dev)
        u_next = torch.empty(B, E, device=dev)
        score = torch.empty(B, device=dev)
        ce = torch.empty(B, device=dev) if target is not None else None
    else:
        a_t, u_next, score, ce = out
    tgt = _i32(target)
    check(lib.sfb_follower_step_tail(B, A, E, _p(logit, name="logit"), _p(is_valid, name="is_valid"),
                                     _p(tgt, torch.int32, "target"), fb, _p(sample_u, name="sample_u"),
                                     _p(all_u_t, name="all_u_t"), _p(a_t, torch.int32), _p(u_next), _p(score), _p(ce),
                                     _stream()))
    return a_t, u_next, score, ce


@_on_tensor_device
def visual_attention(w: Dict[str, Tensor], h: Tensor, visual: Optional[Tensor], store=None, vp_idx=None,
                     view_idx=None):
    """VisualSoftDotAttention.forward (model.py:310-326) -> (feature [B,F], alpha_v [B,V])."""
    lib = _lib.load()
    B = h.shape[0]
    V = visual.shape[1] if visual is not None else store.feat_table.shape[1]
    H = h.shape[1]
    F = w["visual_attention_layer.linear_in_v.weight"].shape[1]
    D = w["visual_attention_layer.linear_in_h.weight"].shape[0]
    d = Dims(F, F, H, D, V)
    keep = []
    vs = _visual_source(visual, store, vp_idx, view_idx, keep)
    s = VisLstmWeights()
    s.va_w_h = _p(w["visual_attention_layer.linear_in_h.weight"]); s.va_b_h = _p(w["visual_attention_layer.linear_in_h.bias"])
    s.va_w_v = _p(w["visual_attention_layer.linear_in_v.weight"]); s.va_b_v = _p(w["visual_attention_layer.linear_in_v.bias"])
    need = lib.sfb_follower_step_workspace_bytes(C.byref(d), B, 1, 1)
    ws = _workspace(need, h.device, ("visual_attention", B))
    feat = torch.empty(B, F, device=h.device); alpha_v = torch.empty(B, V, device=h.device)
    check(lib.sfb_visual_attention_fwd(C.byref(d), C.byref(s), B, _p(h, name="h"), C.byref(vs), _p(feat), _p(alpha_v),
                                       ws.data_ptr(), ws.numel(), _stream()))
    return feat, alpha_v


@_on_tensor_device
def visual_attention_core(q: Tensor, visual: Optional[Tensor], store=None, vp_idx=None, view_idx=None,
                          out: Optional[tuple] = None, workspace: Optional[Tensor] = None):
    """The attention-gather kernel alone: q [B,F] -> (feature [B,F], alpha_v [B,V]); one launch."""
    lib = _lib.load()
    B, F = q.shape
    V = visual.shape[1] if visual is not None else store.feat_table.shape[1]
    d = Dims(F, F, 4, 4, V)
    keep = []
    vs = _visual_source(visual, store, vp_idx, view_idx, keep)
    if out is None:
        feat = torch.empty(B, F, device=q.device); alpha_v = torch.empty(B, V, device=q.device)
    else:
        feat, alpha_v = out
    need = lib.sfb_follower_step_workspace_bytes(C.byref(d), B, 1, 1)
    if workspace is None or workspace.numel() < need:
        workspace = _workspace(need, q.device, ("visual_attention_core", B))
    check(lib.sfb_visual_attention_core_fwd(C.byref(d), B, _p(q, name="q"), C.byref(vs), _p(feat), _p(alpha_v),
                                            workspace.data_ptr(), workspace.numel(), _stream()))
    return feat, alpha_v


@_on_tensor_device
def soft_dot_attention(w: Dict[str, Tensor], prefix: str, h: Tensor, ctx: Tensor, mask: Optional[Tensor]):
    """SoftDotAttention.forward (model.py:122-143) -> (h_tilde [B,H], alpha [B,L])."""
    lib = _lib.load()
    B, L, H = ctx.shape
    d = Dims(4, 4, H, 4, 1)
    sw = _softdot_weights(w, prefix)
    m = _mask_u8(mask)
    need = lib.sfb_follower_step_workspace_bytes(C.byref(d), B, L, 1)
    ws = _workspace(need, h.device, ("soft_dot_attention", B, L))
    ht = torch.empty(B, H, device=h.device); alpha = torch.empty(B, L, device=h.device)
    check(lib.sfb_soft_dot_attention_fwd(C.byref(d), C.byref(sw), B, L, _p(h, name="h"), _p(ctx, name="ctx"),
                                         _p(m, torch.uint8, "mask"), _p(ht), _p(alpha), ws.data_ptr(), ws.numel(),
                                         _stream()))
    return ht, alpha


@_on_tensor_device
def eltwise_prod_scoring(w: Dict[str, Tensor], h_tilde: Tensor, all_u_t: Tensor, prefix: str = "decoder2action."):
    """EltwiseProdScoring.forward (model.py:342-352) -> logit [B,A]."""
    lib = _lib.load()
    B, A, E = all_u_t.shape
    H = h_tilde.shape[1]
    D = w[prefix + "linear_in_h.weight"].shape[0]
    d = Dims(E, E, H, D, 1)
    sw = _scoring_weights(w, prefix)
    need = lib.sfb_follower_step_workspace_bytes(C.byref(d), B, 1, A)
    ws = _workspace(need, h_tilde.device, ("eltwise_prod_scoring", B, A))
    logit = torch.empty(B, A, device=h_tilde.device)
    check(lib.sfb_eltwise_prod_scoring_fwd(C.byref(d), C.byref(sw), B, A, _p(h_tilde, name="h"), _p(all_u_t, name="all_u_t"),
                                           _p(logit), ws.data_ptr(), ws.numel(), _stream()))
    return logit


@_on_tensor_device
def encoder_lstm(w: Dict[str, Tensor], seq: Tensor, lengths, bidirectional: bool = False,
                 drop_embed: Optional[Tensor] = None):
    """EncoderLSTM.forward (model.py:81-104), without the final ctx dropout -> (ctx, decoder_init, c_t)."""
    lib = _lib.load()
    dev = seq.device
    ndir = 2 if bidirectional else 1
    Hd = w["lstm.weight_hh_l0"].shape[1]
    Ew = w["embedding.weight"].shape[1]
    H = ndir * Hd
    B = seq.shape[0]
    lens = torch.as_tensor([int(x) for x in lengths], dtype=torch.int32)
    maxlen = int(lens.max())
    seq32 = seq[:, :maxlen].to(torch.int32).contiguous()
    lens_d = lens.to(dev)
    ew = EncoderWeights()
    ew.embedding = _p(w["embedding.weight"])
    for i, suf in enumerate(["_l0", "_l0_reverse"][:ndir]):
        ew.w_ih[i] = _p(w["lstm.weight_ih" + suf]); ew.w_hh[i] = _p(w["lstm.weight_hh" + suf])
        ew.b_ih[i] = _p(w["lstm.bias_ih" + suf]); ew.b_hh[i] = _p(w["lstm.bias_hh" + suf])
    ew.e2d_w = _p(w["encoder2decoder.weight"]); ew.e2d_b = _p(w["encoder2decoder.bias"])
    need = lib.sfb_encoder_lstm_workspace_bytes(ndir, Hd, Ew, B, maxlen)
    ws = _workspace(need, dev, ("encoder_lstm", ndir, Hd, Ew, B, maxlen))
    ctx = torch.empty(B, maxlen, H, device=dev); dec = torch.empty(B, H, device=dev); c_t = torch.empty(B, H, device=dev)
    vocab = int(w["embedding.weight"].shape[0])
    check(lib.sfb_encoder_lstm_fwd_vocab(C.byref(ew), vocab, ndir, Hd, Ew, B, maxlen, _p(seq32, torch.int32, "seq"),
                                         _p(lens_d, torch.int32, "lengths"), _p(drop_embed, name="drop_embed"),
                                         _p(ctx), _p(dec), _p(c_t), ws.data_ptr(), ws.numel(), _stream()))
    return ctx, dec, c_t


def _encoder_weights(w: Dict[str, Tensor], ndir: int) -> EncoderWeights:
    ew = EncoderWeights()
    ew.embedding = _p(w["embedding.weight"])
    for i, suf in enumerate(["_l0", "_l0_reverse"][:ndir]):
        ew.w_ih[i] = _p(w["lstm.weight_ih" + suf]); ew.w_hh[i] = _p(w["lstm.weight_hh" + suf])
        ew.b_ih[i] = _p(w["lstm.bias_ih" + suf]); ew.b_hh[i] = _p(w["lstm.bias_hh" + suf])
    ew.e2d_w = _p(w["encoder2decoder.weight"]); ew.e2d_b = _p(w["encoder2decoder.bias"])
    return ew


@_on_tensor_device
def encoder_lstm_train(w: Dict[str, Tensor], seq: Tensor, lengths, drop_embed: Optional[Tensor] = None):
    """sfb_encoder_lstm_train_fwd (unidirectional): forward that keeps the tape sfb_encoder_lstm_bwd needs.
    -> (ctx, decoder_init, c_t, saved) with saved = (tape, seq32, lens_d, maxlen)."""
    lib = _lib.load()
    dev = seq.device
    Hd = w["lstm.weight_hh_l0"].shape[1]
    Ew = w["embedding.weight"].shape[1]
    B = seq.shape[0]
    lens = torch.as_tensor([int(x) for x in lengths], dtype=torch.int32)
    maxlen = int(lens.max())
    seq32 = seq[:, :maxlen].to(torch.int32).contiguous()
    lens_d = lens.to(dev)
    ew = _encoder_weights(w, 1)
    need = lib.sfb_encoder_lstm_workspace_bytes(1, Hd, Ew, B, maxlen)
    ws = _workspace(need, dev, ("encoder_lstm", 1, Hd, Ew, B, maxlen))
    tape = torch.empty(lib.sfb_encoder_lstm_tape_bytes(Hd, B, maxlen), dtype=torch.uint8, device=dev)
    ctx = torch.empty(B, maxlen, Hd, device=dev); dec = torch.empty(B, Hd, device=dev); c_t = torch.empty(B, Hd, device=dev)
    check(lib.sfb_encoder_lstm_train_fwd(C.byref(ew), Hd, Ew, B, maxlen, _p(seq32, torch.int32, "seq"), _p(lens_d, torch.int32, "lengths"),
                                         _p(drop_embed, name="drop_embed"), _p(ctx), _p(dec), _p(c_t), tape.data_ptr(), tape.numel(),
                                         ws.data_ptr(), ws.numel(), _stream()))
    return ctx, dec, c_t, (tape, seq32, lens_d, maxlen)


@_on_tensor_device
def encoder_lstm_bwd(w: Dict[str, Tensor], saved, decoder_init: Tensor, g_ctx, g_dec, g_c, grads: Dict[str, Tensor],
                     drop_embed: Optional[Tensor] = None, accumulate: bool = False) -> None:
    """sfb_encoder_lstm_bwd: hand-written BPTT over the tape of encoder_lstm_train; fills `grads` (state_dict names)."""
    lib = _lib.load()
    tape, seq32, lens_d, maxlen = saved
    Hd = w["lstm.weight_hh_l0"].shape[1]
    Ew = w["embedding.weight"].shape[1]
    B = seq32.shape[0]
    ew = _encoder_weights(w, 1)
    gs = _lib.EncoderGrads(_p(grads.get("lstm.weight_ih_l0")), _p(grads.get("lstm.weight_hh_l0")), _p(grads.get("lstm.bias_ih_l0")),
                           _p(grads.get("lstm.bias_hh_l0")), _p(grads.get("encoder2decoder.weight")), _p(grads.get("encoder2decoder.bias")))
    need = lib.sfb_encoder_lstm_bwd_workspace_bytes(Hd, Ew, B, maxlen)
    ws = _workspace(need, seq32.device, ("encoder_lstm_bwd", Hd, Ew, B, maxlen))
    gc = lambda t: None if t is None else t.contiguous()
    g_ctx, g_dec, g_c = gc(g_ctx), gc(g_dec), gc(g_c)
    check(lib.sfb_encoder_lstm_bwd(C.byref(ew), Hd, Ew, B, maxlen, _p(seq32, torch.int32), _p(lens_d, torch.int32), _p(drop_embed),
                                   tape.data_ptr(), _p(decoder_init), _p(g_ctx), _p(g_dec), _p(g_c), C.byref(gs),
                                   1 if accumulate else 0, ws.data_ptr(), ws.numel(), _stream()))


@_on_tensor_device
def speaker_encoder_step(w: Dict[str, Tensor], action_embedding: Tensor, visual: Optional[Tensor], h0: Tensor,
                         c0: Tensor, drop_x: Optional[Tensor] = None, store=None, vp_idx=None, view_idx=None,
                         packed: Optional[Tensor] = None, workspace: Optional[Tensor] = None):
    """SpeakerEncoderLSTM._forward_one_step (model.py:429-435) -> (h1, c1).  `packed` = PackedVisLstm().get(w)."""
    lib = _lib.load()
    B = h0.shape[0]
    V = visual.shape[1] if visual is not None else store.feat_table.shape[1]
    d = follower_dims(w, V)
    keep = []
    vs = _visual_source(visual, store, vp_idx, view_idx, keep)
    wl = _vis_lstm_weights(w)
    need = lib.sfb_follower_step_workspace_bytes(C.byref(d), B, 1, 1)
    ws = workspace if workspace is not None and workspace.numel() >= need else _workspace(need, h0.device, ("spk_enc_step", B))
    h1 = torch.empty_like(h0); c1 = torch.empty_like(c0)
    if packed is not None:
        check(lib.sfb_speaker_encoder_step_packed_fwd(
            C.byref(d), C.byref(wl), packed.data_ptr(), packed.numel(), B, _p(action_embedding, name="action_embedding"),
            C.byref(vs), _p(h0, name="h_0"), _p(c0, name="c_0"), _p(drop_x, name="drop_x"), _p(h1), _p(c1),
            ws.data_ptr(), ws.numel(), _stream()))
        return h1, c1
    check(lib.sfb_speaker_encoder_step_fwd(C.byref(d), C.byref(wl), B, _p(action_embedding, name="action_embedding"),
                                           C.byref(vs), _p(h0, name="h_0"), _p(c0, name="c_0"), _p(drop_x, name="drop_x"),
                                           _p(h1), _p(c1), ws.data_ptr(), ws.numel(), _stream()))
    return h1, c1


@_on_tensor_device
def speaker_decoder_step(w: Dict[str, Tensor], prev_word: Tensor, h0: Tensor, c0: Tensor, ctx: Tensor,
                         ctx_mask: Optional[Tensor], drop_e: Optional[Tensor] = None, drop_h: Optional[Tensor] = None,
                         packed: Optional[Tensor] = None, workspace: Optional[Tensor] = None):
    """SpeakerDecoderLSTM.forward default branch (model.py:497-503,515-519) -> (h1, c1, alpha, logit).
    `packed` = PackedSpeakerDecoder().get(w) -> every projection on tcgen05 from packed weights."""
    lib = _lib.load()
    B, T, H = ctx.shape
    vocab, Ew = w["embedding.weight"].shape
    dev = h0.device
    sw = _spk_dec_weights(w)
    pw = _i32(prev_word.reshape(-1))
    m = _mask_u8(ctx_mask)
    need = lib.sfb_speaker_decoder_step_workspace_bytes(H, Ew, B, T)
    ws = workspace if workspace is not None and workspace.numel() >= need else _workspace(need, dev, ("spk_dec_step", B, T))
    h1 = torch.empty(B, H, device=dev); c1 = torch.empty(B, H, device=dev)
    alpha = torch.empty(B, T, device=dev); logit = torch.empty(B, vocab, device=dev)
    if packed is not None:
        check(lib.sfb_speaker_decoder_step_packed_fwd(
            C.byref(sw), packed.data_ptr(), packed.numel(), H, Ew, vocab, B, T, _p(pw, torch.int32, "previous_word"),
            _p(h0, name="h_0"), _p(c0, name="c_0"), _p(ctx, name="ctx"), _p(m, torch.uint8, "ctx_mask"),
            _p(drop_e, name="drop_e"), _p(drop_h, name="drop_h"), _p(h1), _p(c1), _p(alpha), _p(logit),
            ws.data_ptr(), ws.numel(), _stream()))
        return h1, c1, alpha, logit
    check(lib.sfb_speaker_decoder_step_fwd(C.byref(sw), H, Ew, vocab, B, T, _p(pw, torch.int32, "previous_word"),
                                           _p(h0, name="h_0"), _p(c0, name="c_0"), _p(ctx, name="ctx"),
                                           _p(m, torch.uint8, "ctx_mask"), _p(drop_e, name="drop_e"),
                                           _p(drop_h, name="drop_h"), _p(h1), _p(c1), _p(alpha), _p(logit),
                                           ws.data_ptr(), ws.numel(), _stream()))
    return h1, c1, alpha, logit


_SPK_DEC_GRAD_FIELDS = {"lstm_w_ih": "lstm.weight_ih", "lstm_w_hh": "lstm.weight_hh", "lstm_b_ih": "lstm.bias_ih",
                        "lstm_b_hh": "lstm.bias_hh", "w_in": "attention_layer.linear_in.weight",
                        "w_out": "attention_layer.linear_out.weight", "w_voc": "decoder2action.weight", "b_voc": "decoder2action.bias"}


@_on_tensor_device
def speaker_encoder_step_bwd(w: Dict[str, Tensor], action_embedding: Tensor, visual: Tensor, h0: Tensor, c0: Tensor,
                             drop_x: Optional[Tensor], c1: Tensor, fwd_workspace: Tensor, g_h1: Optional[Tensor],
                             g_c1: Optional[Tensor], grads: Dict[str, Tensor], accumulate: bool = True):
    """sfb_speaker_encoder_step_bwd: hand-written backward of SpeakerEncoderLSTM._forward_one_step.  `fwd_workspace` is the
    workspace the forward call of this step ran in.  Returns (d_h0, d_c0)."""
    lib = _lib.load()
    B, H = h0.shape
    d = follower_dims(w, visual.shape[1])
    keep = []
    vs = _visual_source(visual, None, None, None, keep)
    wl = _vis_lstm_weights(w)
    gs = _lib.FollowerGrads()
    for f, k in _GRAD_FIELDS.items():
        if f.startswith("lstm_") or f.startswith("va_"):
            setattr(gs, f, _p(grads.get(k), name="grad " + k))
    need = lib.sfb_speaker_encoder_step_bwd_workspace_bytes(C.byref(d), B)
    ws = _workspace(need, h0.device, ("spk_enc_step_bwd", B))
    d_h0 = torch.empty_like(h0); d_c0 = torch.empty_like(c0)
    gc = lambda t: None if t is None else t.contiguous()
    g_h1, g_c1 = gc(g_h1), gc(g_c1)
    check(lib.sfb_speaker_encoder_step_bwd(
        C.byref(d), C.byref(wl), B, _p(action_embedding, name="action_embedding"), C.byref(vs), _p(h0, name="h_0"), _p(c0, name="c_0"),
        _p(drop_x, name="drop_x"), _p(c1, name="c_1"), fwd_workspace.data_ptr(), _p(g_h1), _p(g_c1), _p(d_h0), _p(d_c0),
        C.byref(gs), 1 if accumulate else 0, ws.data_ptr(), ws.numel(), _stream()))
    return d_h0, d_c0


@_on_tensor_device
def speaker_decoder_step_bwd(w: Dict[str, Tensor], prev_word: Tensor, h0: Tensor, c0: Tensor, ctx: Tensor, ctx_mask: Optional[Tensor],
                             drop_e: Optional[Tensor], drop_h: Optional[Tensor], c1: Tensor, alpha: Tensor, fwd_workspace: Tensor,
                             g_h1: Optional[Tensor], g_c1: Optional[Tensor], g_logit: Optional[Tensor], grads: Dict[str, Tensor],
                             accumulate: bool = True):
    """sfb_speaker_decoder_step_bwd: hand-written backward of SpeakerDecoderLSTM.forward.  Returns (d_h0, d_c0, d_ctx)."""
    lib = _lib.load()
    B, T, H = ctx.shape
    vocab, Ew = w["embedding.weight"].shape
    sw = _spk_dec_weights(w)
    gs = _lib.SpeakerDecoderGrads()
    for f, k in _SPK_DEC_GRAD_FIELDS.items():
        setattr(gs, f, _p(grads.get(k), name="grad " + k))
    pw = _i32(prev_word.reshape(-1))
    m = _mask_u8(ctx_mask)
    need = lib.sfb_speaker_decoder_step_bwd_workspace_bytes(H, Ew, vocab, B)
    ws = _workspace(need, h0.device, ("spk_dec_step_bwd", B))
    d_h0 = torch.empty_like(h0); d_c0 = torch.empty_like(c0); d_ctx = torch.empty_like(ctx)
    gc = lambda t: None if t is None else t.contiguous()
    g_h1, g_c1, g_logit = gc(g_h1), gc(g_c1), gc(g_logit)
    check(lib.sfb_speaker_decoder_step_bwd(
        C.byref(sw), H, Ew, vocab, B, T, _p(pw, torch.int32, "previous_word"), _p(h0, name="h_0"), _p(c0, name="c_0"),
        _p(ctx, name="ctx"), _p(m, torch.uint8, "ctx_mask"), _p(drop_e, name="drop_e"), _p(drop_h, name="drop_h"),
        _p(c1, name="c_1"), _p(alpha, name="alpha"), fwd_workspace.data_ptr(), _p(g_h1), _p(g_c1), _p(g_logit),
        _p(d_h0), _p(d_c0), _p(d_ctx), C.byref(gs), 1 if accumulate else 0, ws.data_ptr(), ws.numel(), _stream()))
    return d_h0, d_c0, d_ctx


def nav_step(nav, state: Tensor, ended: Tensor, goal: Optional[Tensor], a_prev: Optional[Tensor], actions_log: Optional[Tensor],
             out: dict, with_target: bool = True) -> None:
    """sfb_nav_step: advance the table-driven environment `nav` (navgraph_env.DeviceNavTables) by the actions `a_prev`
    and write the next decode step's inputs into `out` (vp_idx, view_idx, cand_view, cand_trig, is_valid, target)."""
    lib = _lib.load()
    t = _lib.NavTables(_p(nav.vp, torch.int32), _p(nav.view, torch.int32), _p(nav.nvalid, torch.int32), _p(nav.cv, torch.int32),
                       _p(nav.trig), _p(nav.next, torch.int32), _p(nav.teach, torch.int32) if nav.teach is not None else None,
                       nav.S, nav.A, nav.G)
    with torch.cuda.device(state.device):
        check(lib.sfb_nav_step(C.byref(t), state.numel(), _p(state, torch.int32), _p(ended, torch.int32),
                               _p(goal, torch.int32), _p(a_prev, torch.int32), _p(actions_log, torch.int32),
                               _p(out["vp_idx"], torch.int32), _p(out["view_idx"], torch.int32), _p(out["cand_view"], torch.int32),
                               _p(out["cand_trig"]), _p(out["is_valid"]),
                               _p(out["target"], torch.int32) if with_target else None, _stream()))


class SfSearchState:
    """Device arrays of the state-factored search bookkeeping (include/sf_b200.h: sfb_sf_search_state) for B instances
    over a table-driven environment with S states; node 0 of every instance is the root (start state, score 0)."""

    def __init__(self, B: int, S: int, start_states: Tensor, max_iter: int, max_nodes: int, device):
        i32 = lambda *shape, fill=0: torch.full(shape, fill, dtype=torch.int32, device=device)
        f32 = lambda *shape: torch.full(shape, -float("inf"), dtype=torch.float32, device=device)
        u8 = lambda *shape: torch.zeros(shape, dtype=torch.uint8, device=device)
        self.B, self.S, self.max_iter, self.max_nodes = B, S, max_iter, max_nodes
        self.beam_node = i32(B)
        self.c_score, self.c_node, self.c_exp = f32(B, S), i32(B, S), u8(B, S)
        self.h_score, self.h_node, self.h_exp = f32(B, S), i32(B, S), u8(B, S)
        self.d_score, self.d_node, self.n_done = f32(B, S), i32(B, S), i32(B)
        self.n_nodes = i32(B, fill=1)
        self.node_parent, self.node_state, self.node_action = i32(B, max_nodes, fill=-1), i32(B, max_nodes), i32(B, max_nodes, fill=-1)
        self.node_count, self.node_slot = i32(B, max_nodes), i32(B, max_nodes)
        self.node_score = torch.zeros(B, max_nodes, dtype=torch.float32, device=device)
        self.trav = i32(B, max_iter, fill=-1)
        self.flags = i32(4)
        rows = torch.arange(B, device=device)
        st = start_states.to(device=device, dtype=torch.int32)
        self.node_state[:, 0] = st
        self.c_score[rows, st.long()] = 0.0            # cache[key(start)] = [root, expanded] (follower.py:752-756)
        self.c_exp[rows, st.long()] = 1

    def struct(self):
        s = _lib.SfSearchState()
        for name in ("beam_node", "c_node", "h_node", "d_node", "n_done", "n_nodes", "node_parent", "node_state", "node_action",
                     "node_count", "node_slot", "trav", "flags"):
            setattr(s, name, _p(getattr(self, name), torch.int32))
        for name in ("c_score", "h_score", "d_score", "node_score"):
            setattr(s, name, _p(getattr(self, name)))
        for name in ("c_exp", "h_exp"):
            setattr(s, name, _p(getattr(self, name), torch.uint8))
        s.max_nodes, s.max_iter = self.max_nodes, self.max_iter
        return s


def nav_tables_struct(nav):
    return _lib.NavTables(_p(nav.vp, torch.int32), _p(nav.view, torch.int32), _p(nav.nvalid, torch.int32), _p(nav.cv, torch.int32),
                          _p(nav.trig), _p(nav.next, torch.int32), _p(nav.teach, torch.int32) if nav.teach is not None else None,
                          nav.S, nav.A, nav.G)


def sf_search_update(state: SfSearchState, nav, iteration: int, episode_len: int, completion_size: int, lp: Tensor) -> None:
    """sfb_sf_search_update: one iteration of the state-factored search bookkeeping (follower.py:886-924) on the device."""
    lib = _lib.load()
    with torch.cuda.device(lp.device):
        check(lib.sfb_sf_search_update(C.byref(state.struct()), C.byref(nav_tables_struct(nav)), state.B, iteration, episode_len,
                                       completion_size, _p(lp, name="lp"), _stream()))


def set_option(name: str, value: int) -> None:
    """Testing hook: 'disable_tc' (1 = exact-fp32 FFMA gates GEMM), 'tc_debug'."""
    check(_lib.load().sfb_set_option(name.encode(), int(value)))


def last_launch_count() -> int:
    return int(_lib.load().sfb_last_launch_count())
