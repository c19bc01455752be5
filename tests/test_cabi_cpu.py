"""CPU-side checks of the boundary: the C-ABI library loads here (no GPU) and exports every symbol that
include/sf_b200.h declares; argument errors are reported before any launch; the product path refuses to run
without CUDA (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest
import torch

from speaker_follower_b200 import _lib, build, ops, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "sf_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(sfb_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 12
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_abi_version_and_workspace_queries(lib):
    assert lib.sfb_abi_version() == 1
    d = _lib.Dims(2176, 2176, 512, 256, 36)
    n = lib.sfb_follower_step_workspace_bytes(C.byref(d), 100, 80, 8)
    assert n > 100 * 2048 * 4 and n % 256 == 0
    assert lib.sfb_follower_step_workspace_bytes(None, 100, 80, 8) == 0
    assert lib.sfb_speaker_decoder_step_workspace_bytes(512, 300, 256, 6) > 0
    assert lib.sfb_encoder_lstm_workspace_bytes(1, 512, 300, 100, 80) > 100 * 80 * 2048 * 4


def test_argument_errors_before_launch(lib):
    d = _lib.Dims(2176, 2176, 512, 256, 36)
    st = lib.sfb_follower_step_fwd(C.byref(d), None, None, None, 1, 1, 1, *([None] * 2), None, *([None] * 11),
                                   None, 0, None)
    assert st == -1 and b"NULL" in lib.sfb_last_error()
    bad = _lib.Dims(2175, 2176, 512, 256, 36)
    st = lib.sfb_follower_step_fwd(C.byref(bad), None, None, None, 1, 1, 1, *([None] * 2), None, *([None] * 11),
                                   None, 0, None)
    assert st == -1 and b"multiples of 4" in lib.sfb_last_error()
    st = lib.sfb_follower_step_tail(4, 3, 2176, None, None, None, 1, None, None, None, None, None, None, None)
    assert st == -1


def test_no_cpu_fallback():
    w = synth.follower_decoder_weights(emb=48, hid=32, feat=40)
    x = synth.follower_step_inputs(2, 5, 3, seed=1, img_dim=40 - 128 if False else 2048)
    with pytest.raises(_lib.SfbError):
        ops.follower_step(w, x["u_t_prev"], x["all_u_t"], x["visual_context"], x["h_0"], x["c_0"], x["ctx"],
                          x["ctx_mask"])


def test_packed_path_queries_and_argument_errors(lib):
    """The packed fast paths report which dimensions they cover and validate arguments before any launch."""
    d = _lib.Dims(2176, 2176, 512, 256, 36)
    n = lib.sfb_follower_packed_bytes(C.byref(d))
    assert n > 48_000_000 and n % 256 == 0                      # >= the 48.5 MB of fp32 weights, as bf16 hi/lo tiles
    assert lib.sfb_vis_lstm_packed_bytes(C.byref(d)) > 39_000_000
    assert lib.sfb_speaker_decoder_packed_bytes(512, 300, 991) > 0
    small = _lib.Dims(48, 40, 32, 16, 36)                        # H % 128 != 0 -> in-place path only
    assert lib.sfb_follower_packed_bytes(C.byref(small)) == 0
    assert lib.sfb_vis_lstm_packed_bytes(C.byref(small)) == 0
    assert lib.sfb_speaker_decoder_packed_bytes(96, 300, 991) == 0
    assert lib.sfb_speaker_decoder_packed_bytes(512, 301, 991) == 0
    assert lib.sfb_follower_project_ctx_workspace_bytes(C.byref(d), 100, 80) > 100 * 80 * 512 * 4
    assert lib.sfb_follower_project_ctx_workspace_bytes(None, 100, 80) == 0
    st = lib.sfb_follower_pack_weights(C.byref(d), None, None, None, None, 0, None)
    assert st == -1 and b"NULL" in lib.sfb_last_error()
    st = lib.sfb_follower_pack_weights(C.byref(small), None, None, None, None, 0, None)
    assert st == -1 and b"packed path needs" in lib.sfb_last_error()
    st = lib.sfb_follower_step_packed_fwd(C.byref(d), None, None, 0, 1, 1, 1, None, None, None, *([None] * 11), None, None,
                                          None, None, None, None, None, 0, None)
    assert st == -1
    st = lib.sfb_follower_project_ctx(C.byref(d), None, 0, 1, 1, None, None, 0, None, None, None, 0, None)
    assert st == -1
    st = lib.sfb_speaker_decoder_pack_weights(None, 512, 300, 991, None, 0, None)
    assert st == -1


def test_ctx_rows_and_workspace_cache():
    rows = ops.ctx_rows([3, 1, 0], 4, "cpu")
    assert rows.dtype == torch.int32 and rows.tolist() == [0, 1, 2, 4]
    a = ops._workspace(1000, "cpu", ("t", 1))
    assert a.numel() >= 1000 and int(a.sum()) == 0
    assert ops._workspace(1000, "cpu", ("t", 1)) is a            # same signature -> same buffer, no re-zeroing
    assert ops._workspace(1000, "cpu", ("t", 2)) is not a        # different layout -> its own buffer
