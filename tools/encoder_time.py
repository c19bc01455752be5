"""Time the per-rollout pieces around the decode steps at the benchmark shape: EncoderLSTM (B=100, L=80) and the ctx
projection.  GPU box only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from speaker_follower_b200 import ops, synth
torch.cuda.set_device(0)
B, L = 100, 80
we = {k: v.cuda() for k, v in synth.follower_encoder_weights().items()}
seq, mask, lengths = synth.instruction_batch(B, L, seed=47)
seq = seq.cuda()
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / n
for opt in (0, 1):
    ops.set_option("disable_persist", opt)
    print("encoder_lstm B=%d L=%d (disable_persist=%d): %.1f us per call, %d launches" % (B, L, opt, t(lambda: ops.encoder_lstm(we, seq, lengths)), ops.last_launch_count()))
for opt in (0, 1):
    ops.set_option("disable_tc", opt)
    print("encoder_lstm B=%d L=%d (disable_tc=%d): %.1f us per call, %d launches" % (B, L, opt, t(lambda: ops.encoder_lstm(we, seq, lengths)), ops.last_launch_count()))
ops.set_option("disable_tc", 0)
ops.set_option("disable_persist", 0)
import ctypes as C
from speaker_follower_b200 import _lib
ops.set_option("trace", 1)
ops.encoder_lstm(we, seq, lengths); torch.cuda.synchronize()
buf = (C.c_int64 * (64 * 16))()
n = _lib.load().sfb_debug_read_trace(buf, 64)
t0 = buf[0]
for k in range(min(n, 8)):
    e, wt, x = buf[16 * k], buf[16 * k + 1], buf[16 * k + 2]
    print("launch %d: entry %.1f wait_done %.1f exit %.1f us" % (k, (e - t0) / 1e3, (wt - t0) / 1e3, (x - t0) / 1e3))
    ph = [buf[16 * k + j] for j in range(4, 16)]
    if any(ph):
        base = min(v for v in ph if v)
        print("      marks (us since the first): " + "  ".join("%d:%.2f" % (j + 4, (v - base) / 1e3) for j, v in enumerate(ph) if v))
print("launch %d: entry %.1f" % (n - 1, (buf[16 * (n - 1)] - t0) / 1e3))
