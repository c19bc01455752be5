"""Multi-GPU plumbing for the instance-sharded path: one process per GPU, torch.distributed
(NCCL on the GPU box, gloo in CPU tests).  The decode path itself has no collective — instances are independent
given replicated weights and feature table — so this is only: which instances a rank owns, the max-over-ranks
step time, and the small end-of-pass exchanges of the pragmatic-inference combine (rational_follower.py:118-150)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_indices(n: int, rank: int, world_size: int) -> List[int]:
    """Instances owned by `rank`: strided, so sorted-by-length inputs stay balanced."""
    return list(range(rank, n, world_size))


def aggregate_rate(units_per_rank: int, elapsed_ms: float, device=None) -> Tuple[float, float]:
    """(whole-job units/s, max-over-ranks ms): value = units all ranks processed / slowest rank's time."""
    rank, ws = world()
    t = torch.tensor([float(elapsed_ms)], device=device)
    if ws > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    return ws * units_per_rank / (ms * 1e-3), ms


def global_std(values: torch.Tensor) -> float:
    """np.std (population, ddof=0) over the values of ALL ranks via one all-reduce of (n, sum, sum of squares) in
    fp64 — the normaliser of rational_follower.py:125-126 when candidates are sharded."""
    v = values.detach().to(torch.float64).flatten()
    acc = torch.stack([torch.tensor(float(v.numel()), dtype=torch.float64, device=v.device), v.sum(), (v * v).sum()])
    _, ws = world()
    if ws > 1:
        if acc.is_cuda:
            dist.all_reduce(acc)
        else:
            dist.all_reduce(acc)
    n, s, ss = acc.tolist()
    mean = s / n
    return float(max(ss / n - mean * mean, 0.0) ** 0.5)


def gather_records(local: torch.Tensor) -> torch.Tensor:
    """All-gather variable-length per-candidate records [(instr_idx, cand_idx, follower_score, speaker_score), ...]
    (float64 [n_local, 4]) from every rank, returned sorted by (instr_idx, cand_idx) so the result equals the
    single-process order."""
    rank, ws = world()
    local = local.to(torch.float64).reshape(-1, 4)
    if ws == 1:
        allr = local
    else:
        n = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
        sizes = [torch.zeros_like(n) for _ in range(ws)]
        dist.all_gather(sizes, n)
        mx = int(max(int(s.item()) for s in sizes))
        pad = torch.zeros(mx, 4, dtype=torch.float64, device=local.device)
        pad[: local.shape[0]] = local
        bufs = [torch.zeros_like(pad) for _ in range(ws)]
        dist.all_gather(bufs, pad)
        allr = torch.cat([b[: int(s.item())] for b, s in zip(bufs, sizes)], 0)
    key = allr[:, 0] * 1e6 + allr[:, 1]
    return allr[torch.argsort(key)]


def gather_json_shards(items: Sequence[Tuple[int, dict]]) -> List[dict]:
    """C5 (data_augmentation_from_speaker.py:52-82 sharded by trajectory): every rank holds the generated records of
    its own instances as (global instance index, JSON-serialisable dict).  One all-gather of the byte lengths and one of
    the padded uint8 buffers; every rank returns the records of ALL ranks in global-index order, i.e. exactly the list a
    single process would have written (the JSON file of the single-GPU run is reproduced byte for byte by dumping it)."""
    import json
    rank, ws = world()
    payload = json.dumps([[int(i), rec] for i, rec in items]).encode("utf-8")
    if ws == 1:
        merged = json.loads(payload.decode("utf-8"))
    else:
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        n = torch.tensor([len(payload)], dtype=torch.int64, device=dev)
        sizes = [torch.zeros_like(n) for _ in range(ws)]
        dist.all_gather(sizes, n)
        mx = int(max(int(s.item()) for s in sizes))
        buf = torch.zeros(mx, dtype=torch.uint8, device=dev)
        buf[: len(payload)] = torch.frombuffer(bytearray(payload), dtype=torch.uint8).to(dev)
        bufs = [torch.zeros_like(buf) for _ in range(ws)]
        dist.all_gather(bufs, buf)
        merged = []
        for b, s in zip(bufs, sizes):
            merged.extend(json.loads(bytes(b[: int(s.item())].cpu().tolist()).decode("utf-8")))
    merged.sort(key=lambda t: t[0])
    return [rec for _, rec in merged]
