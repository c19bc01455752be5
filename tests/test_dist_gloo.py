"""world_size-2 gloo tests (CPU) of the N>1 plumbing: instance sharding, max-over-ranks timing, and the sharded
pragmatic-inference combine reproducing the single-process result (rational_follower.py:118-150)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import r2r_oracle as O
from speaker_follower_b200 import dist as D
from speaker_follower_b200 import pragmatic as PR


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = np.random.Generator(np.random.PCG64(5))
        n_instr, n_cand = 7, 5
        spk = g.standard_normal((n_instr, n_cand)) * 3 - 20
        fol = g.standard_normal((n_instr, n_cand)) * 2 - 5
        mine = D.shard_indices(n_instr, rank, world)
        recs = torch.tensor([[i, c, fol[i, c], spk[i, c]] for i in mine for c in range(n_cand)], dtype=torch.float64)
        allr = D.gather_records(recs)
        s_std = D.global_std(recs[:, 3])
        f_std = D.global_std(recs[:, 2])
        rate, ms = D.aggregate_rate(100, 10.0 * (rank + 1))
        # C5: generated instructions of this rank's trajectories, gathered into single-process order
        traj = [(i, {"path_id": i, "instr_id": "%d_0" % i, "words": ["w%d" % ((i * 7 + k) % 11) for k in range(3 + i % 4)],
                     "score": float(spk[i % n_instr, 0])}) for i in D.shard_indices(11, rank, world)]
        shards = D.gather_json_shards(traj)
        # the PRODUCT combine (speaker_follower_b200/pragmatic.py) on the gathered records + reduced statistics
        choice = PR.rational_combine(allr.numpy(), (0.0, 0.95), stds=(f_std, s_std))
        q.put((rank, allr.numpy(), s_std, f_std, rate, ms, spk, fol, shards, choice))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_combine_equals_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    outs.sort(key=lambda o: o[0])
    _, allr0, s_std, f_std, rate, ms, spk, fol, shards0, choice0 = outs[0]
    assert choice0 == outs[1][9]                                   # every rank forms the same decision
    # JSON shard gather (data_augmentation_from_speaker.py:52-82 sharded): both ranks hold all 11 records in index order
    assert shards0 == outs[1][8] and [r["path_id"] for r in shards0] == list(range(11))
    assert all(r["words"] == ["w%d" % ((r["path_id"] * 7 + k) % 11) for k in range(3 + r["path_id"] % 4)] for r in shards0)
    assert np.array_equal(allr0, outs[1][1])                       # every rank sees the same, ordered records
    assert allr0.shape == (35, 4)
    assert np.array_equal(allr0[:, 0], np.repeat(np.arange(7), 5))
    assert abs(s_std - np.std(spk)) < 1e-12 and abs(f_std - np.std(fol)) < 1e-12     # global population std
    assert ms == 20.0 and abs(rate - 2 * 100 / 0.020) < 1e-9                         # max over ranks, whole-job units
    # weighted argmax with the sharded statistics == single-process oracle combine
    groups = [int(i) for i in allr0[:, 0]]
    best = O.rational_combine(allr0[:, 3], allr0[:, 2], groups, 0.95)
    comb = 0.95 * spk / np.std(spk) + 0.05 * fol / np.std(fol)
    for i in range(7):
        assert int(allr0[best[i], 1]) == int(np.argmax(comb[i]))
        assert choice0[0.95][i] == int(np.argmax(comb[i]))         # product combine == oracle combine == numpy
        assert choice0[0.0][i] == int(np.argmax(fol[i]))


def test_shard_indices_cover_and_balance():
    for n in (1, 7, 64, 2560):
        for w in (1, 2, 4, 8):
            parts = [D.shard_indices(n, r, w) for r in range(w)]
            assert sorted(sum(parts, [])) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


class _Env(object):
    def __init__(self, n, bs):
        self.data = [{"instr_id": "%d_0" % i} for i in range(n)]
        self.batch_size = bs
        self.ix = 0


def _shard_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        env = _Env(22, 4)
        mine = PR.shard_env(env, whole_batches=True)
        env2 = _Env(22, 4)
        strided = PR.shard_env(env2)
        q.put((rank, mine, [d["instr_id"] for d in env.data], env.batch_size, strided))
    finally:
        dist.destroy_process_group()


def test_two_rank_whole_minibatch_sharding_keeps_single_process_batches():
    """C5 sharding (pragmatic.shard_env(whole_batches=True)): every minibatch a rank runs has exactly the members it has in
    the single-process run (consecutive chunks of batch_size), the ranks' shares are disjoint and cover the data."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=120) for _ in procs], key=lambda o: o[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    single = [list(range(i, min(22, i + 4))) for i in range(0, 22, 4)]
    for rank, mine, ids, bs, strided in outs:
        assert ids == ["%d_0" % i for i in mine] and bs == 4
        local = [mine[i:i + 4] for i in range(0, len(mine), 4)]
        assert local == single[rank::2]                      # whole single-process minibatches, dealt out round-robin
        assert strided == list(range(rank, 22, 2))
    assert sorted(outs[0][1] + outs[1][1]) == list(range(22))
