"""Bring-up: one packed follower step at a given shape (under compute-sanitizer if wanted).  GPU box only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from speaker_follower_b200 import ops, synth
B, L, A = (int(v) for v in sys.argv[1:4])
w = {k: v.cuda() for k, v in synth.follower_decoder_weights().items()}
x = {k: v.cuda() for k, v in synth.follower_step_inputs(B, L, A, seed=1).items()}
blob = ops.PackedFollower().get(w)
torch.cuda.synchronize(); print("packed", flush=True)
for i in range(2):
    res = ops.follower_step(w, x["u_t_prev"], x["all_u_t"], x["visual_context"], x["h_0"], x["c_0"], x["ctx"], x["ctx_mask"], packed=blob)
    torch.cuda.synchronize(); print("step", i, "ok", float(res[3].abs().sum()), flush=True)
