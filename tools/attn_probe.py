"""Probe the attention-gather kernel: dense vs gather source, batch size, against torch streaming kernels timed the
same way (CUDA events around single launches, fresh data each launch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from speaker_follower_b200 import ops, synth
torch.cuda.set_device(0)
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(1)
NVP = 10567
table = torch.empty(NVP, 36, 2048, device=dev)
for i in range(0, NVP, 1024):
    table[i:i + 1024].normal_(generator=g).clamp_(min=0)
store = ops.FeatureStore(table, synth.loc_embedding_table().to(dev))
F = 2176

def timeit(fn, n=100, warm=5):
    for i in range(warm): fn(i)
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for i in range(n):
        evs[i][0].record(); fn(i + warm); evs[i][1].record()
    torch.cuda.synchronize()
    t = np.array([a.elapsed_time(b) for a, b in evs]) * 1e3
    return float(np.median(t)), float(t.min())

def timeit_batch(fn, n=100, warm=5):
    for i in range(warm): fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n): fn(i + warm)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / n

for B in (100, 400):
    q = torch.randn(B, F, device=dev) * 0.05
    feat = torch.empty(B, F, device=dev); av = torch.empty(B, 36, device=dev)
    ws = torch.zeros(1 << 26, dtype=torch.uint8, device=dev)
    vps = [torch.randint(0, NVP, (B,), device=dev, dtype=torch.int32, generator=g) for _ in range(120)]
    view = torch.randint(0, 36, (B,), device=dev, dtype=torch.int32, generator=g)
    dense = [store.dense(vps[i], view) for i in range(12)]          # 12 x 31 MB (B=100) rotating > L2
    mb = B * 36 * F * 4 / 1e6
    med, mn = timeit(lambda i: ops.visual_attention_core(q, None, store=store, vp_idx=vps[i % 120], view_idx=view, out=(feat, av), workspace=ws))
    print("B=%d gather : median %.2f us (min %.2f)  -> %.0f GB/s" % (B, med, mn, mb / med * 1e3))
    med, mn = timeit(lambda i: ops.visual_attention_core(q, dense[i % 12], out=(feat, av), workspace=ws))
    print("B=%d dense  : median %.2f us (min %.2f)  -> %.0f GB/s" % (B, med, mn, mb / med * 1e3))
    bt = timeit_batch(lambda i: ops.visual_attention_core(q, None, store=store, vp_idx=vps[i % 120], view_idx=view, out=(feat, av), workspace=ws))
    print("B=%d gather back-to-back (PDL chained): %.2f us/launch -> %.0f GB/s" % (B, bt, mb / bt * 1e3))
    ops.set_option("disable_pdl", 1)
    bt = timeit_batch(lambda i: ops.visual_attention_core(q, None, store=store, vp_idx=vps[i % 120], view_idx=view, out=(feat, av), workspace=ws))
    print("B=%d gather back-to-back (no PDL): %.2f us/launch -> %.0f GB/s" % (B, bt, mb / bt * 1e3))
    ops.set_option("disable_pdl", 0)
    # torch streaming kernels over the same bytes
    outs = torch.empty(B, 36, device=dev)
    med, mn = timeit(lambda i: torch.sum(dense[i % 12], dim=2, out=outs))
    print("B=%d torch.sum(dim=2) over the dense slab: median %.2f us (min %.2f) -> %.0f GB/s" % (B, med, mn, mb / med * 1e3))
    dst = torch.empty_like(dense[0])
    med, mn = timeit(lambda i: dst.copy_(dense[i % 12]))
    print("B=%d torch copy_ (read+write): median %.2f us -> %.0f GB/s (r+w)" % (B, med, 2 * mb / med * 1e3))
    del dense
# empty-ish kernel launch floor
x = torch.zeros(32, device=dev)
med, mn = timeit(lambda i: x.add_(1.0))
print("tiny torch kernel between events: median %.2f us (min %.2f)" % (med, mn))
