"""Device vs host state-factored search at the C4 bench configuration (64 instructions, completion 40, 160-viewpoint graph):
same number of completions per instance, same candidates.  GPU box only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from speaker_follower_b200.navgraph_env import DeviceNavTables, FakeR2RBatch

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
mk = lambda: FakeR2RBatch(n_viewpoints=160, n_instr=64, batch_size=64, seed=77, max_len=40, beam_size=40, with_features=False)
env_d, env_h = mk(), mk()
fd, _ = bench._agents(env_d, dev, instruction_len=30)
fh, _ = bench._agents(env_h, dev, instruction_len=30)
nav = DeviceNavTables(env_d, dev, with_teacher=False)
with torch.no_grad():
    # same kernels on the same batch in both runs (no per-episode projections on the device side, every instance a row on
    # the host side): any difference is a difference of the search logic, not of rounding
    got, _, walk_g = fd.device_state_factored_search(nav, 40, use_ctx_proj=False)
    want, _, walk_w = fh.state_factored_search(40, 1, _pad_batch=True)
print("iterations on the device:", fd.last_search_iterations)
print("completions device:", sum(len(g) for g in got), "host:", sum(len(w) for w in want))
bad = 0
for i, (g, w) in enumerate(zip(got, want)):
    if len(g) != len(w):
        print("instance", i, "counts", len(g), len(w)); bad += 1; continue
    for k, (cg, cw) in enumerate(zip(g, w)):
        if [int(a) for a in cg["actions"]] != [int(a) for a in cw["actions"]] or abs(float(cg["score"]) - float(cw["score"])) > 3e-4:
            gap = abs(float(w[k]["score"]) - float(w[min(k + 1, len(w) - 1)]["score"]))
            print("instance", i, "candidate", k, "differs: scores", float(cg["score"]), float(cw["score"]), "next gap", gap); bad += 1
            break
    if [s.world_state.viewpointId for s in walk_g[i]] != [s.world_state.viewpointId for s in walk_w[i]]:
        print("instance", i, "walk differs", len(walk_g[i]), len(walk_w[i])); bad += 1
print("mismatching instances:", bad)
import time
for graph in (True, False):
    for rep in range(2):
        env_d.reset_epoch()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        with torch.no_grad():
            fd.device_state_factored_search(nav, 40, cuda_graph=graph)
        torch.cuda.synchronize()
        print("device search, cuda_graph=%s: %.1f ms (%d iterations)" % (graph, (time.perf_counter() - t0) * 1e3, fd.last_search_iterations))
