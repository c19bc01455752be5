"""In-graph timeline of one follower decode step: every kernel's {entry, after-PDL-wait, exit} (block 0) in ns.
GPU box only.  Usage: python tools/step_trace.py [disable_pdl=1 ...]"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from speaker_follower_b200 import ops, synth, _lib
torch.cuda.set_device(0)
PACKED = 0
for kv in sys.argv[1:]:
    k, v = kv.split("=")
    if k == "packed":
        PACKED = int(v)
    else:
        ops.set_option(k, int(v)); print("option", k, v)
dev = torch.device("cuda")
B, L, A = 100, 80, 8
w = {k: v.cuda() for k, v in synth.follower_decoder_weights().items()}
xs = [{k: v.cuda() for k, v in synth.follower_step_inputs(B, L, A, seed=70 + i).items()} for i in range(2)]
ws = torch.zeros(1 << 26, dtype=torch.uint8, device=dev)
hb = [xs[0]["h_0"].clone(), torch.empty(B, 512, device=dev)]
cb = [xs[0]["c_0"].clone(), torch.empty(B, 512, device=dev)]
ub = [xs[0]["u_t_prev"].clone(), torch.empty(B, 2176, device=dev)]
alpha = torch.empty(B, L, device=dev); logit = torch.empty(B, A, device=dev); av = torch.empty(B, 36, device=dev)
a_t = torch.empty(B, dtype=torch.int32, device=dev); score = torch.empty(B, device=dev)
blob = ops.PackedFollower().get(w) if PACKED else None
qb = [ops.follower_carry(w, B), ops.follower_carry(w, B)] if PACKED else [None, None]
cp = [ops.follower_project_ctx(w, blob, x["ctx"]) for x in xs] if PACKED == 3 else None
def step(i):
    x, s = xs[i % 2], i % 2
    if PACKED == 3:   # + per-episode ctx projections (text side off the chain, helper stream)
        ops.follower_step(w, ub[s], x["all_u_t"], x["visual_context"], hb[s], cb[s], x["ctx"], x["ctx_mask"], workspace=ws,
                          out=(hb[s ^ 1], cb[s ^ 1], alpha, logit, av), packed=blob, carry_in=qb[s], carry_out=qb[s ^ 1],
                          ctx_proj=cp[i % 2],
                          tail={"is_valid": x["is_valid"], "feedback": "argmax", "out": (a_t, ub[s ^ 1], score, None)})
        return
    if PACKED == 2:   # carried query + fused tail
        ops.follower_step(w, ub[s], x["all_u_t"], x["visual_context"], hb[s], cb[s], x["ctx"], x["ctx_mask"], workspace=ws,
                          out=(hb[s ^ 1], cb[s ^ 1], alpha, logit, av), packed=blob, carry_in=qb[s], carry_out=qb[s ^ 1],
                          tail={"is_valid": x["is_valid"], "feedback": "argmax", "out": (a_t, ub[s ^ 1], score, None)})
        return
    ops.follower_step(w, ub[s], x["all_u_t"], x["visual_context"], hb[s], cb[s], x["ctx"], x["ctx_mask"], workspace=ws,
                      out=(hb[s ^ 1], cb[s ^ 1], alpha, logit, av), packed=blob)
    ops.follower_tail(logit, x["is_valid"], x["all_u_t"], "argmax", out=(a_t, ub[s ^ 1], score, None))
side = torch.cuda.Stream()
with torch.cuda.stream(side):
    for i in range(4): step(i)
torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
ops.set_option("trace", 1)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for i in range(1): step(i)
for _ in range(6): g.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): g.replay()
e1.record(); torch.cuda.synchronize()
print("graph of 1 step: %.1f us per step" % (e0.elapsed_time(e1) * 1e3 / 50))
buf = (C.c_int64 * (64 * 16))()
n = _lib.load().sfb_debug_read_trace(buf, 64)
names = ["gemm t_v", "gemm q", "attn visual", "gemm gates(TC)+lstm", "gemm t", "attn text", "gemm h~", "gemm t'", "gemm g", "scoring", "tail"]
if PACKED:
    names = ["pk q", "attn visual(+pack)", "pk gates+lstm", "pk [t|hh]", "attn text", "pk h~", "pk g", "scoring", "tail"]
if PACKED == 3:
    names = ["fused gather+gates+lstm", "fused text attn+proj+scoring+tail"]
if PACKED == 2:
    names = ["attn visual(+pack)", "pk gates+lstm", "pk [t|hh|q']", "attn text", "pk h~", "pk g", "scoring+tail"]
t0 = buf[0]
prev_exit = t0
for k in range(n):
    e, wt, x, x2 = buf[16 * k], buf[16 * k + 1], buf[16 * k + 2], buf[16 * k + 3]
    print("%2d %-22s entry %8.2f  wait_done %8.2f  exit %8.2f  (dur %6.2f us, after wait %6.2f us)%s" % (
        k, names[k % len(names)], (e - t0) / 1e3, (wt - t0) / 1e3, (x - t0) / 1e3, (x - e) / 1e3, (x - wt) / 1e3,
        ("  merge_exit %.2f" % ((x2 - t0) / 1e3)) if x2 else ""))
    ph = [buf[16 * k + j] for j in range(4, 16)]
    if any(ph):
        print("      phases(us since wait): " + "  ".join("%d:%.2f" % (j + 4, (v - wt) / 1e3) for j, v in enumerate(ph) if v))
# per-CTA timeline of the LAST attention launch of the step (the text attention)
import numpy as np
ops.set_option("trace", 0)
ops.set_option("cta_trace", 1)
g2 = torch.cuda.CUDAGraph()          # the trace pointer is a kernel argument: capture again with it set
with torch.cuda.graph(g2):
    step(0)
for _ in range(3): g2.replay()
torch.cuda.synchronize()
nb = 4096
cbuf = (C.c_int64 * (nb * 8))()
_lib.load().sfb_debug_read_cta_trace(cbuf, nb)
t = np.array(list(cbuf), dtype=np.int64).reshape(nb, 8)
t = t[t[:, 0] > 0]
if len(t):
    base = np.median(t[:, 0])
    ent, first, done, ex, qr = [(t[:, k] - base) / 1e3 for k in (0, 1, 2, 3, 5)]
    ok = t[:, 1] > 0
    print("text attention CTAs (us rel. to median entry): %d | entry min %.2f max %.2f | q ready med %.2f | first rows med %.2f max %.2f | stream done med %.2f max %.2f | exit med %.2f max %.2f"
          % (len(t), ent.min(), ent.max(), np.median(qr), np.median(first[ok]), first[ok].max(), np.median(done), done.max(), np.median(ex), ex.max()))
