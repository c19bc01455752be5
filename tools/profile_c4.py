"""cProfile of one C4 pass (state-factored search + speaker rescoring) on one GPU: where the HOST time goes.
GPU box only.  Usage: python tools/profile_c4.py [n_instructions]"""
import cProfile, io, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from speaker_follower_b200 import pragmatic as PR
from speaker_follower_b200.navgraph_env import FakeR2RBatch

n_inst = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)


def make_env():
    return FakeR2RBatch(n_viewpoints=160, n_instr=n_inst, batch_size=64, seed=77, max_len=40, beam_size=40, with_features=False)


def one_pass():
    env = make_env()
    follower, speaker = bench._agents(env, dev, instruction_len=30)
    return PR.run_rational_follower(env, follower, speaker, beam_size=40)


one_pass()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
one_pass()
torch.cuda.synchronize()
pr.disable()
print("phases:", PR.LAST_PHASES)
for key in ("cumulative", "tottime"):
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats(key).print_stats(28)
    print(s.getvalue()[:6000])
