"""Phase timeline (SM clocks) of CTA 0 of the tcgen05 gates GEMM inside a follower step.  GPU box only."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from speaker_follower_b200 import ops, synth, _lib
torch.cuda.set_device(0)
w = {k: v.cuda() for k, v in synth.follower_decoder_weights().items()}
x = {k: v.cuda() for k, v in synth.follower_step_inputs(100, 80, 8, seed=77).items()}
def run():
    return ops.follower_step(w, x["u_t_prev"], x["all_u_t"], x["visual_context"], x["h_0"], x["c_0"], x["ctx"], x["ctx_mask"])
for _ in range(3): run()
ops.set_option("tc_debug", 4)
run(); torch.cuda.synchronize()
buf = (C.c_int64 * 256)()
_lib.load().sfb_debug_read_timestamps(buf, 256)
t = list(buf)
base = t[0]
print("setup->pdl_wait", t[1] - t[0], "| producer loop end", t[2] - t[0], "| tmem->L2", t[3] - t[2], "| semaphore", t[4] - t[3], "| reduce+lstm", t[5] - t[4], "| exit", t[6] - t[5])
for it in range(12):
    r = t[8 + it * 8: 8 + it * 8 + 3]
    if r[0] == 0: break
    print("it %2d start %7d | wait_empty %5d convert+arrive %5d" % (it, r[0] - base, r[1] - r[0], r[2] - r[1]))
