"""Pragmatic inference and speaker-driven data augmentation on the CUDA agents, sharded by instance.

    run_rational_follower        tasks/R2R/rational_follower.py:11-150   (config C4)
    generate_speaker_instructions tasks/R2R/data_augmentation_from_speaker.py:52-82 (literal speaker, config C5)

Both passes are embarrassingly parallel over instructions / trajectories: every rank owns a strided
shard of the environment's instances, runs the search + rescoring (or the greedy speaker decode) for it on its own GPU,
and the only exchange happens once at the end:
  * C4: an all-gather of (instruction, candidate, follower score, speaker score) records and an all-reduce of
    (n, sum, sum of squares) for the two GLOBAL population standard deviations rational_follower.py:125-126 divides by,
    after which every rank forms the same weighted argmax;
  * C5: a length + padded-byte all-gather of the generated records, re-ordered by global instance index, so that the JSON
    a sharded run writes equals the single-process file byte for byte.
"""
from __future__ import annotations

import json
from collections import Counter
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import dist as D


LAST_PHASES: Dict[str, float] = {}   # wall-clock split of the last run_rational_follower call on this rank (reporting only)


def shard_env(env, whole_batches: bool = False) -> List[int]:
    """Keep only this rank's share of ``env.data`` (instances are independent); returns the GLOBAL indices kept, in local
    order.  Default: strided by instance.  ``whole_batches``: the consecutive minibatches a single process would form
    (env.batch_size items each) are dealt out round-robin instead, so that every minibatch has exactly the members it has
    in the single-process run — needed where a result depends on its minibatch: SpeakerEncoderLSTM runs every row for the
    longest path of its batch (model.py:437-457), so generated instructions are a function of the batch composition in
    the reference as well.  The batch size is clipped so that the last minibatch does not wrap around a short shard."""
    rank, ws = D.world()
    n = len(env.data)
    if whole_batches and hasattr(env, "batch_size"):
        bs = max(1, int(env.batch_size))
        chunks = [list(range(i, min(n, i + bs))) for i in range(0, n, bs)]
        mine = [i for c in chunks[rank::ws] for i in c]
    else:
        mine = D.shard_indices(n, rank, ws)
    env.data = [env.data[i] for i in mine]
    env.ix = 0
    if hasattr(env, "batch_size"):
        env.batch_size = max(1, min(env.batch_size, len(env.data)))
    return mine


def rational_combine(records, weights: Sequence[float] = (0.0, 0.95), stds: Optional[tuple] = None) -> Dict[float, Dict[int, int]]:
    """rational_follower.py:118-150.  ``records``: float64 [N, 4] rows (instruction index, candidate index, follower
    score, speaker score) of ALL candidates of the split; the two normalisers are the population standard deviations
    (np.std, ddof 0) over all of them — pass ``stds=(follower_std, speaker_std)`` when they were reduced across ranks.
    Returns {speaker weight: {instruction index: index of the winning candidate}} (first maximum wins, like max())."""
    r = np.asarray(records, dtype=np.float64).reshape(-1, 4)
    fol, spk = r[:, 2], r[:, 3]
    f_std, s_std = stds if stds is not None else (np.std(fol), np.std(spk))
    out = {}
    for w in weights:
        comb = spk * (w / s_std) + fol * ((1.0 - w) / f_std)
        best: Dict[int, int] = {}
        top: Dict[int, float] = {}
        for i in range(r.shape[0]):
            g = int(r[i, 0])
            if g not in best or comb[i] > top[g]:
                best[g], top[g] = int(r[i, 1]), float(comb[i])
        out[float(w)] = best
    return out


def run_rational_follower(env, follower, speaker, beam_size: int, state_factored_search: bool = True,
                          state_first_n_ws_key: int = 4, physical_traversal: bool = False,
                          weights: Sequence[float] = (0.0, 0.95), global_index: Optional[List[int]] = None, nav=None):
    """The candidate generation + rescoring + combine of run_rational_follower for the instances of ``env`` (already
    sharded with shard_env() under torch.distributed).  Returns (results_by_weight, candidate_lists, records) where
    results_by_weight[w][instr_id] is the chosen candidate dict (only for this rank's instructions), candidate_lists
    maps instr_id -> candidates with 'follower_score' / 'speaker_score', and records is the all-rank float64 [N,4] table."""
    from .follower import least_common_viewpoint_path, path_element_from_observation
    follower.env = env
    env.reset_epoch()
    follower.encoder.eval(); follower.decoder.eval()
    speaker.encoder.eval(); speaker.decoder.eval()
    follower.set_beam_size(beam_size)
    follower.feedback = "argmax"
    candidate_lists: Dict[str, list] = {}
    order: List[str] = []
    looped = False
    import time as _time
    phases = {"search_s": 0.0, "rescoring_s": 0.0, "exchange_s": 0.0}   # wall clock with a device sync at each boundary
    LAST_PHASES.clear(); LAST_PHASES.update(phases)

    def _tick():
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        return _time.perf_counter()
    with torch.no_grad():
        while not looped:
            t0 = _tick()
            if state_factored_search and nav is not None and state_first_n_ws_key == 4:
                # the environment as device tables: the whole search state lives on the device
                beam_candidates, inf_states, traversed = follower.device_state_factored_search(nav, beam_size, load_next_minibatch=True)
            elif state_factored_search:
                beam_candidates, inf_states, traversed = follower.state_factored_search(
                    beam_size, 1, load_next_minibatch=True, first_n_ws_key=state_first_n_ws_key)
            else:
                beam_candidates, inf_states, traversed = follower.beam_search(beam_size, load_next_minibatch=True)
            t1 = _tick()
            flat = [c for cands in beam_candidates for c in cands]
            scored, _ = speaker._score_obs_actions_and_instructions(
                [c["observations"] for c in flat], [c["actions"] for c in flat], [c["instr_encoding"] for c in flat], "teacher")
            assert len(scored) == len(flat)
            t2 = _tick()
            LAST_PHASES["search_s"] += t1 - t0
            LAST_PHASES["rescoring_s"] += t2 - t1
            start = 0
            for ii, cands in enumerate(beam_candidates):
                for i, c in enumerate(cands):
                    s = scored[start + i]
                    assert c["instr_id"] == s["instr_id"]
                    c["follower_score"] = float(c["score"])
                    c["speaker_score"] = float(s["score"])
                    del c["observations"]
                    if physical_traversal:                                            # rational_follower.py:82-99
                        last = traversed[ii][-1]
                        walk = least_common_viewpoint_path(last, inf_states[ii][i])
                        phys = [path_element_from_observation(st.observation) for st in traversed[ii] + walk[1:]]
                        assert phys[-1][0] == c["trajectory"][-1][0]
                        c["trajectory"] = phys
                start += len(cands)
                iid = cands[0]["instr_id"]
                if iid in candidate_lists:
                    looped = True
                else:
                    candidate_lists[iid] = cands
                    order.append(iid)
            # the reference keeps drawing minibatches until an instruction repeats (rational_follower.py:44-69), i.e. it runs
            # one more search + rescoring pass whose results it throws away whenever the data divides into whole batches;
            # every instruction of the split has its candidates once each id has been seen
            if hasattr(env, "data") and len(candidate_lists) >= len({it["instr_id"] for it in env.data}):
                looped = True
    # ---- the one exchange of the pass: candidate records + global standard deviations
    t3 = _tick()
    gi = global_index if global_index is not None else list(range(len(order)))
    local = torch.tensor([[gi[k], j, c["follower_score"], c["speaker_score"]]
                          for k, iid in enumerate(order) for j, c in enumerate(candidate_lists[iid])], dtype=torch.float64)
    dev = next(follower.decoder.parameters()).device
    _, ws = D.world()
    if ws > 1 and torch.distributed.get_backend() == "nccl":
        local = local.to(dev)
    records = D.gather_records(local)
    f_std, s_std = D.global_std(local[:, 2]), D.global_std(local[:, 3])
    best = rational_combine(records.cpu().numpy(), weights, stds=(f_std, s_std))
    LAST_PHASES["exchange_s"] = _tick() - t3
    results_by_weight = {}
    for w, choice in best.items():
        res, counts = {}, Counter()
        for k, iid in enumerate(order):
            j = choice[gi[k]]
            res[iid] = candidate_lists[iid][j]
            counts[j] += 1
        results_by_weight[w] = {"results": res, "index_counts": counts}
    return results_by_weight, candidate_lists, records.cpu().numpy()


def generate_speaker_instructions(env, speaker, global_index: Optional[List[int]] = None, path: Optional[str] = None):
    """C5: literal-speaker generation (``speaker.test(use_dropout=False, feedback='argmax')``,
    data_augmentation_from_speaker.py:66) for this rank's trajectories, then ONE gather of the JSON records in global
    trajectory order.  Every rank returns the full list; rank 0 writes ``path`` (the file a single process would write)."""
    speaker.env = env
    with torch.no_grad():
        res = speaker.test(use_dropout=False, feedback="argmax")
    ids = [it["instr_id"] for it in env.data]
    gi = global_index if global_index is not None else list(range(len(ids)))
    items = []
    for k, iid in enumerate(ids):
        r = res[iid]
        items.append((gi[k], {"instr_id": iid, "word_indices": [int(w) for w in r["word_indices"]],
                              "score": round(float(r["score"]), 4)}))
    merged = D.gather_json_shards(items)
    rank, _ = D.world()
    if path is not None and rank == 0:
        with open(path, "w") as f:
            json.dump(merged, f, sort_keys=True, indent=1)
    return merged
