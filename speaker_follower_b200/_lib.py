"""ctypes binding of libsf_b200.so (the C ABI in include/sf_b200.h).

The product path has no CPU fallback: if the shared library is missing or cannot be loaded this module
raises, and every op raises when handed a non-CUDA tensor.
"""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libsf_b200.so")

c_float_p = C.c_void_p     # device pointers are passed as integers
c_int_p = C.c_void_p
c_u8_p = C.c_void_p


class SfbError(RuntimeError):
    pass


class Dims(C.Structure):
    _fields_ = [("E", C.c_int32), ("F", C.c_int32), ("H", C.c_int32), ("D", C.c_int32), ("V", C.c_int32)]


class VisLstmWeights(C.Structure):
    _fields_ = [(n, c_float_p) for n in
                ("lstm_w_ih", "lstm_w_hh", "lstm_b_ih", "lstm_b_hh", "va_w_h", "va_b_h", "va_w_v", "va_b_v")]


class SoftDotWeights(C.Structure):
    _fields_ = [("w_in", c_float_p), ("w_out", c_float_p)]


class ScoringWeights(C.Structure):
    _fields_ = [(n, c_float_p) for n in ("w_h", "b_h", "w_a", "b_a", "w_out", "b_out")]


class VisualSource(C.Structure):
    _fields_ = [("visual", c_float_p), ("feat_table", c_float_p), ("loc_table", c_float_p),
                ("vp_idx", c_int_p), ("view_idx", c_int_p), ("img_dim", C.c_int32), ("idx_dependent", C.c_int32)]


class StepTail(C.Structure):
    _fields_ = [("is_valid", c_float_p), ("target", c_int_p), ("feedback", C.c_int32), ("sample_u", c_float_p),
                ("a_t", c_int_p), ("u_next", c_float_p), ("action_score", c_float_p), ("ce", c_float_p)]


class ActionSource(C.Structure):
    _fields_ = [("all_u_t", c_float_p), ("feat_table", c_float_p), ("vp_idx", c_int_p), ("cand_view", c_int_p),
                ("cand_trig", c_float_p), ("img_dim", C.c_int32)]


class SpeakerDecoderWeights(C.Structure):
    _fields_ = [("embedding", c_float_p), ("lstm_w_ih", c_float_p), ("lstm_w_hh", c_float_p),
                ("lstm_b_ih", c_float_p), ("lstm_b_hh", c_float_p), ("attn", SoftDotWeights),
                ("w_voc", c_float_p), ("b_voc", c_float_p)]


class EncoderWeights(C.Structure):
    _fields_ = [("embedding", c_float_p), ("w_ih", c_float_p * 2), ("w_hh", c_float_p * 2),
                ("b_ih", c_float_p * 2), ("b_hh", c_float_p * 2), ("e2d_w", c_float_p), ("e2d_b", c_float_p)]


class NavTables(C.Structure):
    _fields_ = [("vp", c_int_p), ("view", c_int_p), ("nvalid", c_int_p), ("cv", c_int_p), ("trig", c_float_p),
                ("next", c_int_p), ("teach", c_int_p), ("S", C.c_int32), ("A", C.c_int32), ("G", C.c_int32)]


class EncoderGrads(C.Structure):
    _fields_ = [(n, c_float_p) for n in ("w_ih", "w_hh", "b_ih", "b_hh", "e2d_w", "e2d_b")]


class FollowerGrads(C.Structure):
    _fields_ = [(n, c_float_p) for n in ("lstm_w_ih", "lstm_w_hh", "lstm_b_ih", "lstm_b_hh", "va_w_h", "va_b_h", "va_w_v",
                                         "w_in", "w_out", "sc_w_h", "sc_b_h", "sc_w_a", "sc_b_a", "sc_w_out", "sc_b_out")]


class SfSearchState(C.Structure):
    _fields_ = [("beam_node", c_int_p), ("c_score", c_float_p), ("c_node", c_int_p), ("c_exp", c_u8_p),
                ("h_score", c_float_p), ("h_node", c_int_p), ("h_exp", c_u8_p),
                ("d_score", c_float_p), ("d_node", c_int_p), ("n_done", c_int_p), ("n_nodes", c_int_p),
                ("node_parent", c_int_p), ("node_state", c_int_p), ("node_action", c_int_p), ("node_count", c_int_p),
                ("node_slot", c_int_p), ("node_score", c_float_p), ("trav", c_int_p), ("flags", c_int_p),
                ("max_nodes", C.c_int32), ("max_iter", C.c_int32)]


class SpeakerDecoderGrads(C.Structure):
    _fields_ = [(n, c_float_p) for n in ("lstm_w_ih", "lstm_w_hh", "lstm_b_ih", "lstm_b_hh", "w_in", "w_out", "w_voc", "b_voc")]


# name -> (restype, argtypes); every symbol include/sf_b200.h declares
SIGNATURES = {
    "sfb_abi_version": (C.c_int32, []),
    "sfb_last_error": (C.c_char_p, []),
    "sfb_last_launch_count": (C.c_int32, []),
    "sfb_set_option": (C.c_int32, [C.c_char_p, C.c_int32]),
    "sfb_debug_read_trace": (C.c_int32, [C.c_void_p, C.c_int32]),
    "sfb_debug_read_cta_trace": (C.c_int32, [C.c_void_p, C.c_int32]),
    "sfb_debug_read_timestamps": (C.c_int32, [C.c_void_p, C.c_int32]),
    "sfb_debug_max_active_clusters": (C.c_int32, [C.c_int32, C.c_int32]),
    "sfb_device_info": (C.c_int32, [C.POINTER(C.c_int32)] * 3),
    "sfb_follower_step_workspace_bytes": (C.c_size_t, [C.POINTER(Dims), C.c_int32, C.c_int32, C.c_int32]),
    "sfb_speaker_decoder_step_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "sfb_visual_attention_fwd": (C.c_int32, [C.POINTER(Dims), C.POINTER(VisLstmWeights), C.c_int32, c_float_p,
                                             C.POINTER(VisualSource), c_float_p, c_float_p,
                                             C.c_void_p, C.c_size_t, C.c_void_p]),
    "sfb_visual_attention_core_fwd": (C.c_int32, [C.POINTER(Dims), C.c_int32, c_float_p, C.POINTER(VisualSource),
                                                  c_float_p, c_float_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "sfb_soft_dot_attention_fwd": (C.c_int32, [C.POINTER(Dims), C.POINTER(SoftDotWeights), C.c_int32, C.c_int32,
                                               c_float_p, c_float_p, c_u8_p, c_float_p, c_float_p,
                                               C.c_void_p, C.c_size_t, C.c_void_p]),
    "sfb_follower_step_fwd": (C.c_int32, [C.POINTER(Dims), C.POINTER(VisLstmWeights), C.POINTER(SoftDotWeights),
                                          C.POINTER(ScoringWeights), C.c_int32, C.c_int32, C.c_int32,
                                          c_float_p, c_float_p, C.POINTER(VisualSource), c_float_p, c_float_p,
                                          c_float_p, c_u8_p, c_float_p, c_float_p,
                                          c_float_p, c_float_p, c_float_p, c_float_p, c_float_p,
                                          C.c_void_p, C.c_size_t, C.c_void_p]),
    "sfb_follower_packed_bytes": (C.c_size_t, [C.POINTER(Dims)]),
    "sfb_follower_pack_weights": (C.c_int32, [C.POINTER(Dims), C.POINTER(VisLstmWeights), C.POINTER(SoftDotWeights),
                                              C.POINTER(ScoringWeights), C.c_void_p, C.c_size_t, C.c_void_p]),
    "sfb_follower_step_packed_fwd": (C.c_int32, [C.POINTER(Dims), C.POINTER(VisLstmWeights), C.c_void_p, C.c_size_t,
                                                 C.c_int32, C.c_int32, C.c_int32,
                                                 c_float_p, c_float_p, C.POINTER(VisualSource), c_float_p, c_float_p,
                                                 c_float_p, c_u8_p, c_float_p, c_float_p,
                                                 c_float_p, c_float_p, c_float_p, c_float_p, c_float_p,
                                                 C.c_void_p, C.c_void_p, C.POINTER(StepTail), C.POINTER(ActionSource),
                                                 c_float_p, c_float_p,
                                                 C.c_void_p, C.c_size_t, C.c_void_p]),
    "sfb_follower_carry_bytes": (C.c_size_t, [C.POINTER(Dims), C.c_int32]),
    "sfb_follower_gather_lstm_fwd": (C.c_int32, [C.POINTER(Dims), C.POINTER(VisLstmWeights), C.c_void_p, C.c_size_t, C.c_int32,
                                                 C.c_void_p, C.POINTER(VisualSource), c_float_p, c_float_p, c_float_p, c_float_p,
                                                 c_float_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "sfb_encoder_lstm_tape_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "sfb_encoder_lstm_train_fwd": (C.c_int32, [C.POINTER(EncoderWeights), C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                               c_int_p, c_int_p, c_float_p, c_float_p, c_float_p, c_float_p,
                                               C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]),
    "sfb_encoder_lstm_bwd_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "sfb_encoder_lstm_bwd": (C.c_int32, [C.POINTER(EncoderWeights), C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                         c_int_p, c_int_p, c_float_p, C.c_void_p, c_float_p, c_float_p, c_float_p, c_float_p,
                                         C.POINTER(EncoderGrads), C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p]),
    "sfb_follower_step_bwd_workspace_bytes": (C.c_size_t, [C.POINTER(Dims), C.c_int32, C.c_int32, C.c_int32]),
    "sfb_follower_step_bwd": (C.c_int32, [C.POINTER(Dims), C.POINTER(VisLstmWeights), C.POINTER(SoftDotWeights),
                                          C.POINTER(ScoringWeights), C.c_int32, C.c_int32, C.c_int32,
                                          c_float_p, C.POINTER(ActionSource), C.POINTER(VisualSource), c_float_p, c_float_p,
                                          c_float_p, c_u8_p, c_float_p, c_float_p,
                                          c_float_p, c_float_p, c_float_p, C.c_void_p,
                                          c_float_p, c_float_p, c_float_p,
                                          c_float_p, c_float_p, c_float_p,
                                          C.POINTER(FollowerGrads), C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p]),
    "sfb_speaker_encoder_step_bwd_workspace_bytes": (C.c_size_t, [C.POINTER(Dims), C.c_int32]),
    "sfb_speaker_encoder_step_bwd": (C.c_int32, [C.POINTER(Dims), C.POINTER(VisLstmWeights), C.c_int32, c_float_p,
                                                 C.POINTER(VisualSource), c_float_p, c_float_p, c_float_p, c_float_p, C.c_void_p,
                                                 c_float_p, c_float_p, c_float_p, c_float_p, C.POINTER(FollowerGrads), C.c_int32,
                                                 C.c_void_p, C.c_size_t, C.c_void_p]),
    "sfb_speaker_decoder_step_bwd_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "sfb_speaker_decoder_step_bwd": (C.c_int32, [C.POINTER(SpeakerDecoderWeights), C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                                 C.c_int32, c_int_p, c_float_p, c_float_p, c_float_p, c_u8_p, c_float_p, c_float_p,
                                                 c_float_p, c_float_p, C.c_void_p, c_float_p, c_float_p, c_float_p,
                                                 c_float_p, c_float_p, c_float_p, C.POINTER(SpeakerDecoderGrads), C.c_int32,
                                                 C.c_void_p, C.c_size_t, C.c_void_p]),
    "sfb_sf_search_update": (C.c_int32, [C.POINTER(SfSearchState), C.POINTER(NavTables), C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                         c_float_p, C.c_void_p]),
    "sfb_nav_step": (C.c_int32, [C.POINTER(NavTables), C.c_int32, c_int_p, c_int_p, c_int_p, c_int_p, c_int_p,
                                 c_int_p, c_int_p, c_int_p, c_float_p, c_float_p, c_int_p, C.c_void_p]),
    "sfb_eltwise_prod_scoring_fwd": (C.c_int32, [C.POINTER(Dims), C.POINTER(ScoringWeights), C.c_int32, C.c_int32,
                                                 c_float_p, c_float_p, c_float_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "sfb_follower_project_ctx_workspace_bytes": (C.c_size_t, [C.POINTER(Dims), C.c_int32, C.c_int32]),
    "sfb_follower_project_ctx": (C.c_int32, [C.POINTER(Dims), C.c_void_p, C.c_size_t, C.c_int32, C.c_int32, c_float_p,
                                             c_int_p, C.c_int32, c_float_p, c_float_p, C.c_void_p, C.c_size_t,
                                             C.c_void_p]),
    "sfb_follower_step_tail": (C.c_int32, [C.c_int32, C.c_int32, C.c_int32, c_float_p, c_float_p, c_int_p, C.c_int32,
                                           c_float_p, c_float_p, c_int_p, c_float_p, c_float_p, c_float_p,
                                           C.c_void_p]),
    "sfb_encoder_lstm_workspace_bytes": (C.c_size_t, [C.c_int32] * 5),
    "sfb_encoder_lstm_fwd": (C.c_int32, [C.POINTER(EncoderWeights), C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                         C.c_int32, c_int_p, c_int_p, c_float_p, c_float_p, c_float_p, c_float_p,
                                         C.c_void_p, C.c_size_t, C.c_void_p]),
    "sfb_encoder_lstm_fwd_vocab": (C.c_int32, [C.POINTER(EncoderWeights), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                               C.c_int32, c_int_p, c_int_p, c_float_p, c_float_p, c_float_p, c_float_p,
                                               C.c_void_p, C.c_size_t, C.c_void_p]),
    "sfb_speaker_encoder_step_fwd": (C.c_int32, [C.POINTER(Dims), C.POINTER(VisLstmWeights), C.c_int32, c_float_p,
                                                 C.POINTER(VisualSource), c_float_p, c_float_p, c_float_p,
                                                 c_float_p, c_float_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "sfb_speaker_decoder_step_fwd": (C.c_int32, [C.POINTER(SpeakerDecoderWeights), C.c_int32, C.c_int32, C.c_int32,
                                                 C.c_int32, C.c_int32, c_int_p, c_float_p, c_float_p, c_float_p,
                                                 c_u8_p, c_float_p, c_float_p, c_float_p, c_float_p, c_float_p,
                                                 c_float_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "sfb_vis_lstm_packed_bytes": (C.c_size_t, [C.POINTER(Dims)]),
    "sfb_vis_lstm_pack_weights": (C.c_int32, [C.POINTER(Dims), C.POINTER(VisLstmWeights), C.c_void_p, C.c_size_t, C.c_void_p]),
    "sfb_speaker_encoder_step_packed_fwd": (C.c_int32, [C.POINTER(Dims), C.POINTER(VisLstmWeights), C.c_void_p, C.c_size_t,
                                                        C.c_int32, c_float_p, C.POINTER(VisualSource), c_float_p, c_float_p,
                                                        c_float_p, c_float_p, c_float_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "sfb_speaker_decoder_packed_bytes": (C.c_size_t, [C.c_int32, C.c_int32, C.c_int32]),
    "sfb_speaker_decoder_pack_weights": (C.c_int32, [C.POINTER(SpeakerDecoderWeights), C.c_int32, C.c_int32, C.c_int32,
                                                     C.c_void_p, C.c_size_t, C.c_void_p]),
    "sfb_speaker_decoder_step_packed_fwd": (C.c_int32, [C.POINTER(SpeakerDecoderWeights), C.c_void_p, C.c_size_t, C.c_int32,
                                                        C.c_int32, C.c_int32, C.c_int32, C.c_int32, c_int_p, c_float_p,
                                                        c_float_p, c_float_p, c_u8_p, c_float_p, c_float_p, c_float_p,
                                                        c_float_p, c_float_p, c_float_p, C.c_void_p, C.c_size_t, C.c_void_p]),
}

_lib = None


def load() -> C.CDLL:
    """Load the shared library (once) and attach the signatures.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SfbError(
            "libsf_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` or "
            "`python speaker_follower_b200/build.py`. There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if lib.sfb_abi_version() != 1:
        raise SfbError("libsf_b200.so ABI version mismatch")
    # bring-up: SFB_OPTIONS="name=value,..." applies sfb_set_option() at load time (e.g. disable_merged=1)
    for kv in filter(None, os.environ.get("SFB_OPTIONS", "").split(",")):
        k, _, v = kv.partition("=")
        if lib.sfb_set_option(k.strip().encode(), int(v)) != 0:
            raise SfbError("SFB_OPTIONS: unknown option %r" % k)
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != 0:
        msg = load().sfb_last_error()
        raise SfbError("sf_b200 error %d: %s" % (status, msg.decode() if msg else "?"))
