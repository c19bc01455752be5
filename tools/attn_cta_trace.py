"""Per-CTA timeline of the attention-gather kernel (B=100, gather mode): launch skew, first data, stream end, exit.
GPU box only.  Usage: python tools/attn_cta_trace.py [attn_cl=4]"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from speaker_follower_b200 import ops, synth, _lib
torch.cuda.set_device(0)
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(1)
NVP, B, F = 4096, 100, 2176
table = torch.empty(NVP, 36, 2048, device=dev).normal_(generator=g).clamp_(min=0)
store = ops.FeatureStore(table, synth.loc_embedding_table().to(dev))
q = torch.randn(B, F, device=dev) * 0.05
feat = torch.empty(B, F, device=dev); av = torch.empty(B, 36, device=dev)
ws = torch.zeros(1 << 26, dtype=torch.uint8, device=dev)
vps = [torch.randint(0, NVP, (B,), device=dev, dtype=torch.int32, generator=g) for _ in range(64)]
view = torch.randint(0, 36, (B,), device=dev, dtype=torch.int32, generator=g)
def run(i): ops.visual_attention_core(q, None, store=store, vp_idx=vps[i % 64], view_idx=view, out=(feat, av), workspace=ws)
def b2b(n=200):
    for i in range(5): run(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n): run(i + 5)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / n
def timeline(tag, cl=4):
    for i in range(3): run(i)
    torch.cuda.synchronize()
    ops.set_option("cta_trace", 1)
    run(7); torch.cuda.synchronize()
    n = B * cl
    buf = (C.c_int64 * (n * 8))()
    _lib.load().sfb_debug_read_cta_trace(buf, n)
    ops.set_option("cta_trace", 0)
    t = np.array(list(buf), dtype=np.int64).reshape(n, 8)
    t0 = t[:, 0].min()
    ent, first, done, ex = [(t[:, k] - t0) / 1e3 for k in (0, 1, 2, 3)]
    print("%-28s b2b %.2f us | entry max %.2f | first-row med %.2f | stream-done med %.2f max %.2f | exit med %.2f max %.2f" % (
        tag, b2b(), ent.max(), np.median(first), np.median(done), done.max(), np.median(ex), ex.max()))
ops.set_option("attn_cl", 4)
for st in (4, 5, 6):
    for nh in (0, 1):
        ops.set_option("attn_stages", st); ops.set_option("attn_nohint", nh)
        timeline("cl=4 stages=%d nohint=%d" % (st, nh))
ops.set_option("attn_stages", 0); ops.set_option("attn_nohint", 0)
ops.set_option("attn_cl", 2)
for st in (6, 8):
    ops.set_option("attn_stages", st)
    timeline("cl=2 stages=%d" % st, cl=2)
