"""Import the staged reference ``model.py`` (oracle/_ref, see make_ref.py) — TEST / BENCH INFRASTRUCTURE ONLY.

``load()`` returns the reference's ``model`` module or ``None`` when ``oracle/_ref`` is absent.  The simulator module
``MatterSim`` (env.py:5) is replaced by an empty stub: nothing on the timed path touches it.  The reference feeds
uint8 masks to ``masked_fill_`` (model.py:135), which torch >= 1.2 rejects, so callers pass ``torch.bool`` masks; no
reference source is modified.
"""
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DST = os.path.join(HERE, "_ref")
_MOD = None


def load():
    global _MOD
    if _MOD is not None:
        return _MOD
    if not os.path.exists(os.path.join(REF_DST, "model.py")):
        return None
    sys.modules.setdefault("MatterSim", types.ModuleType("MatterSim"))
    saved = {k: sys.modules.get(k) for k in ("model", "env", "utils", "paths")}
    sys.path.insert(0, REF_DST)
    try:
        for k in saved:
            sys.modules.pop(k, None)
        _MOD = importlib.import_module("model")
    finally:
        sys.path.remove(REF_DST)
        for k, v in saved.items():          # do not leave generic names ("utils", "env") shadowing anything else
            if k != "model":
                mod = sys.modules.pop(k, None)
                if mod is not None:
                    sys.modules["_sf_ref_" + k] = mod
            if v is not None:
                sys.modules[k] = v
    return _MOD


def follower_decoder(weights, device="cpu"):
    """The reference's AttnDecoderLSTM (model.py:361-397) carrying `weights` (state_dict names), eval mode."""
    m = load()
    if m is None:
        return None
    H = weights["lstm.weight_hh"].shape[1]
    F = weights["visual_attention_layer.linear_in_v.weight"].shape[1]
    E = weights["lstm.weight_ih"].shape[1] - F
    dec = m.AttnDecoderLSTM(E, H, 0.5, feature_size=F)
    missing = dec.load_state_dict({k: v for k, v in weights.items()}, strict=False)
    assert not missing.unexpected_keys, missing
    return dec.to(device).eval()


def speaker_modules(enc_weights, dec_weights, device="cpu"):
    """The reference's SpeakerEncoderLSTM / SpeakerDecoderLSTM (model.py:400-519) carrying the given weights, eval mode."""
    m = load()
    if m is None:
        return None
    import numpy as np
    H = enc_weights["lstm.weight_hh"].shape[1]
    F = enc_weights["visual_attention_layer.linear_in_v.weight"].shape[1]
    E = enc_weights["lstm.weight_ih"].shape[1] - F
    enc = m.SpeakerEncoderLSTM(E, F, H, 0.5)
    vocab, word = dec_weights["embedding.weight"].shape
    dec = m.SpeakerDecoderLSTM(vocab, word, H, 0.5, glove=np.asarray(dec_weights["embedding.weight"].cpu().numpy()))
    for mod, w in ((enc, enc_weights), (dec, dec_weights)):
        missing = mod.load_state_dict({k: v for k, v in w.items()}, strict=False)
        assert not missing.unexpected_keys, missing
    return enc.to(device).eval(), dec.to(device).eval()
