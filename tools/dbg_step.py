"""Bring-up: packed follower steps at a given shape through every fast-path option (under compute-sanitizer if wanted).
GPU box only.  Usage: python tools/dbg_step.py B L A"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from speaker_follower_b200 import ops, synth
B, L, A = (int(v) for v in sys.argv[1:4])
w = {k: v.cuda() for k, v in synth.follower_decoder_weights().items()}
table, loc = synth.feature_table(64, 1031), synth.loc_embedding_table()
x = {k: v.cuda() for k, v in synth.follower_step_inputs(B, L, A, seed=1, table=table, loc=loc).items()}
store = ops.FeatureStore(table.cuda(), loc.cuda())
blob = ops.PackedFollower().get(w)
lengths = (~x["ctx_mask"]).sum(1).tolist()
cproj = ops.follower_project_ctx(w, blob, x["ctx"], rows=ops.ctx_rows(lengths, L, "cuda"))
torch.cuda.synchronize(); print("packed + projected", flush=True)
q = [ops.follower_carry(wc, B) for _ in range(2)]
cv = torch.randint(-1, 36, (B, A), device="cuda").int(); cv[:, 0] = -1
trig = torch.rand(B, A, 4, device="cuda")
h, c, u = x["h_0"], x["c_0"], x["u_t_prev"]
for i in range(3):
    tail = {"is_valid": x["is_valid"], "feedback": "argmax"}
    h, c, alpha, logit, av = ops.follower_step(w, u, None, None, h, c, x["ctx"], x["ctx_mask"], store=store, vp_idx=x["vp_idx"],
                                               view_idx=x["view_idx"], packed=blob, ctx_proj=cproj, carry_in=q[i % 2] if i else None,
                                               carry_out=q[(i + 1) % 2], tail=tail, cand_view=cv, cand_trig=trig)
    u = tail["out"][1]
    torch.cuda.synchronize(); print("step", i, "ok", float(h.abs().sum()), flush=True)
res = ops.follower_step(w, x["u_t_prev"], x["all_u_t"], x["visual_context"], x["h_0"], x["c_0"], x["ctx"], x["ctx_mask"], packed=blob)
torch.cuda.synchronize(); print("plain packed step ok", float(res[3].abs().sum()))
res = ops.follower_step(w, x["u_t_prev"], x["all_u_t"], x["visual_context"], x["h_0"], x["c_0"], x["ctx"], x["ctx_mask"])
torch.cuda.synchronize(); print("in-place step ok", float(res[3].abs().sum()))
