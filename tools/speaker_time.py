"""Time the speaker steps at the C3 shape (N=256 paths, T=6, vocabulary 991): packed tcgen05 path vs in-place path.
GPU box only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from speaker_follower_b200 import ops, synth
torch.cuda.set_device(0)
N, T, H = 256, 6, synth.HID
g = torch.Generator().manual_seed(1)
wd = {k: v.cuda() for k, v in synth.speaker_decoder_weights().items()}
we = {k: v.cuda() for k, v in synth.speaker_encoder_weights().items()}
ctx = torch.tanh(torch.randn(N, T, H, generator=g)).cuda(); h0 = torch.tanh(torch.randn(N, H, generator=g)).cuda()
c0 = (torch.randn(N, H, generator=g) * 0.5).cuda()
mask = (torch.arange(T).unsqueeze(0) >= torch.randint(1, T + 1, (N, 1), generator=g)).cuda()
prev = torch.randint(0, synth.VOCAB, (N,), generator=g).cuda()
x = {k: v.cuda() for k, v in synth.follower_step_inputs(N, 8, 6, seed=3).items()}
a, V = x["all_u_t"][:, 1].contiguous(), x["visual_context"]
wsd = torch.zeros(1 << 26, dtype=torch.uint8, device="cuda"); wse = torch.zeros(1 << 27, dtype=torch.uint8, device="cuda")
pd, pe = ops.PackedSpeakerDecoder().get(wd), ops.PackedVisLstm().get(we)
def graph_time(fn, n=50):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(8): fn()
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): gr.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (n * 8)
for name, blob in (("in-place", None), ("packed  ", pd)):
    us = graph_time(lambda: ops.speaker_decoder_step(wd, prev, h0, c0, ctx, mask, packed=blob, workspace=wsd))
    print("speaker decoder step N=%d T=%d %s: %.1f us (%d launches)" % (N, T, name, us, ops.last_launch_count()))
for name, blob in (("in-place", None), ("packed  ", pe)):
    us = graph_time(lambda: ops.speaker_encoder_step(we, a, V, h0, c0, packed=blob, workspace=wse))
    print("speaker encoder step N=%d        %s: %.1f us (%d launches)" % (N, name, us, ops.last_launch_count()))
