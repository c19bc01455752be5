"""Probe the attention-gather kernel's geometry: cluster size x ring budget x rows per barrier, back-to-back launches
with fresh slabs (the bench's roofline measurement), plus the cluster-occupancy query.  GPU box only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from speaker_follower_b200 import ops, synth, _lib
torch.cuda.set_device(0)
dev = torch.device("cuda")
lib = _lib.load()
for cl in (2, 4, 8, 16):
    print("max active clusters: cluster=%d smem=200KB -> %d ; smem=100KB -> %d ; smem=60KB -> %d" % (
        cl, lib.sfb_debug_max_active_clusters(cl, 200 * 1024), lib.sfb_debug_max_active_clusters(cl, 100 * 1024),
        lib.sfb_debug_max_active_clusters(cl, 60 * 1024)))
g = torch.Generator(device=dev).manual_seed(1)
NVP = 10567
table = torch.empty(NVP, 36, 2048, device=dev)
for i in range(0, NVP, 1024):
    table[i:i + 1024].normal_(generator=g).clamp_(min=0)
store = ops.FeatureStore(table, synth.loc_embedding_table().to(dev))
F, B = 2176, int(os.environ.get("B", 100))
q = torch.randn(B, F, device=dev) * 0.05
ws = torch.zeros(1 << 26, dtype=torch.uint8, device=dev)
N = 200
vps = [torch.randint(0, NVP, (B,), device=dev, dtype=torch.int32, generator=g) for _ in range(N)]
view = torch.randint(0, 36, (B,), device=dev, dtype=torch.int32, generator=g)
mb = B * 36 * F * 4 / 1e6

def run(tag):
    feat = torch.empty(B, F, device=dev); av = torch.empty(B, 36, device=dev)
    f = lambda i: ops.visual_attention_core(q, None, store=store, vp_idx=vps[i % N], view_idx=view, out=(feat, av), workspace=ws)
    for i in range(5): f(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(N): f(i)
    b.record(); torch.cuda.synchronize()
    t = a.elapsed_time(b) * 1e3 / N
    f(0); torch.cuda.synchronize()
    print("%-34s %6.2f us/launch  %5.0f GB/s" % (tag, t, mb / t * 1e3), flush=True)
    return feat.clone(), av.clone()

ref = run("default (plan)")
for cl, ring, rb in [(4, 52, 0), (2, 100, 0), (2, 100, 2), (1, 100, 1), (1, 100, 2), (1, 100, 4), (1, 180, 1), (1, 180, 2), (1, 180, 4), (1, 140, 2), (2, 180, 2)]:
    ops.set_option("attn_cl", cl); ops.set_option("attn_ring_kb", ring); ops.set_option("attn_rb", rb)
    try:
        out = run("cl=%d ring=%dKB rb=%d" % (cl, ring, rb))
        print("    max|dfeat| %.2e  max|dalpha| %.2e" % ((out[0] - ref[0]).abs().max().item(), (out[1] - ref[1]).abs().max().item()))
    except Exception as e:
        print("cl=%d ring=%d rb=%d FAILED: %s" % (cl, ring, rb, e))
