"""Bring-up check of the tcgen05 LSTM-gate GEMM: compare against the exact-fp32 FFMA path for each
descriptor-encoding variant (sfb_set_option tc_debug).  Run on the GPU box."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from speaker_follower_b200 import ops, synth

torch.cuda.set_device(0)
w = {k: v.cuda() for k, v in synth.follower_decoder_weights().items()}
for (B, L, A) in ((100, 80, 8), (8, 20, 6), (256, 12, 5)):
    x = {k: v.cuda() for k, v in synth.follower_step_inputs(B, L, A, seed=77).items()}
    def run():
        return ops.follower_step(w, x["u_t_prev"], x["all_u_t"], x["visual_context"], x["h_0"], x["c_0"], x["ctx"], x["ctx_mask"])
    ops.set_option("disable_tc", 1)
    ref = run(); torch.cuda.synchronize()
    ops.set_option("disable_tc", 0)
    for dbg in (0, 1, 2, 3):
        ops.set_option("tc_debug", dbg)
        try:
            res = run(); torch.cuda.synchronize()
            print("B=%d dbg=%d  max|dh1|=%.3e max|dc1|=%.3e max|dlogit|=%.3e" % (
                B, dbg, (res[0] - ref[0]).abs().max().item(), (res[1] - ref[1]).abs().max().item(),
                (res[3] - ref[3]).abs().max().item()), flush=True)
        except Exception as e:
            print("B=%d dbg=%d FAILED: %r" % (B, dbg, e), flush=True)
            raise
    ops.set_option("tc_debug", 0)
