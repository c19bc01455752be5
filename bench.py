#!/usr/bin/env python
"""bench.py — follower decode-steps/sec (BASELINE.json metric) on N B200s, one process per GPU.

A "step" is one follower decode step over the whole batch: AttnDecoderLSTM.forward (model.py:377-397)
fed from the device-resident feature table (replaces follower.py:291-298) plus the rollout tail
(follower.py:476-505: mask, log-softmax, argmax, next-u gather).  Workload = BASELINE.json configs[1]
shape: batch 100, instruction length 80, 8 action candidates, 36 x 2176 features, fp32.

  python bench.py [--gpus N --steps K --warmup W]            this repo's CUDA path
  python bench.py --impl reference [...]                     the reference's CPU arithmetic (oracle port)

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

B, L, A = 100, 80, 8
N_VIEWPOINTS = 10567          # R2R viewpoints -> 3.1 GB table, >> L2
ATTN_NCU_TRAFFIC = 74742528   # dram__bytes_read.sum + dram__bytes_write.sum of one vis_lstm_fused_kernel launch (ncu --set full, cold)
STEP_NCU_TRAFFIC = 110616064  # same for one step_kernel launch (profiles/r02_ncu_full_summary.txt: 103.64 MB read + 6.98 MB written)
POOL = 10                     # per-step input sets = the steps of one episode (episode_len = 10, train.py:29)
N_CTX = 4                     # rotating episodes: instruction contexts (16 MB each + 32 MB of per-episode projections)


def algorithmic_bytes(E, F, H, n_params):
    """bytes one decode step must move (every operand once)."""
    return 4 * (B * 36 * F + B * L * H + B * A * E + n_params + B * (E + 4 * H) + B * (36 + L + A))


def bench_config(world):
    """One config dict for both arms (the driver compares them key by key)."""
    return {"workload": "follower decode step B=%d L=%d A=%d V=36 F=2176 (AttnDecoderLSTM.forward + rollout tail), per GPU" % (B, L, A),
            "batch": B, "instr_len": L, "actions": A,
            "parallelism": "replicas x%d (instance-sharded, no data-path collective)" % world,
            "cache": "inputs larger than L2: every step gathers fresh slabs + action candidates from a 3.1 GB device table "
                     "(%d step sets = 295 MB of distinct slabs), %d rotating episode contexts; ctx is constant within a "
                     "10-step episode as in the reference rollout; weights (48.5 MB) are step-invariant" % (POOL, N_CTX),
            "launch": "one CUDA graph per 10-step episode: per-episode ctx projection (inside the timed region) + 10 decode steps"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons through NVML during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # pragma: no cover
            self.nv, self.err = None, repr(e)

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def result(self):
        self.stop_flag = True
        self.join(timeout=1.0)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------- reference arms
_CPU_STATE = {}
MIN_REF_SECONDS = 2.0         # the reference arm decides the headline ratio: never time less than this


def _ref_setup():
    """The reference's own AttnDecoderLSTM (oracle/_ref/model.py, staged unmodified by oracle/make_ref.py) when it
    travelled with the snapshot, else the oracle port of the same arithmetic."""
    if not _CPU_STATE:
        from oracle import r2r_oracle as O
        from oracle import make_ref, ref_loader
        from speaker_follower_b200 import synth
        make_ref.make()
        w = synth.follower_decoder_weights()
        xs = [synth.follower_step_inputs(B, L, A, seed=900 + i, n_viewpoints=128) for i in range(2)]
        dec = ref_loader.follower_decoder(w)
        _CPU_STATE.update(O=O, w=w, xs=xs, dec=dec, kind="reference" if dec is not None else "port")
    return _CPU_STATE


def _ref_step(st, dec, w, u, x, h, c):
    """One decode step + rollout tail with the reference arithmetic on whatever device the tensors live on."""
    O = st["O"]
    if dec is not None:   # model.py:377-397, unmodified; bool mask (torch >= 1.2 rejects the uint8 of model.py:135)
        h, c, alpha, logit, alpha_v = dec(u, x["all_u_t"], x["visual_context"], h, c, x["ctx"], x["ctx_mask"])
    else:
        h, c, alpha, logit, alpha_v = O.attn_decoder_step(u, x["all_u_t"], x["visual_context"], h, c, x["ctx"], x["ctx_mask"], w)
    if logit.is_cuda:     # the reference's own tail ops (follower.py:476-505, argmax feedback) on the tensors' device
        logit = logit.masked_fill(x["is_valid"] == 0, -float("inf"))
        a_t = logit.max(1)[1]
        u = x["all_u_t"][torch.arange(logit.shape[0], device=logit.device), a_t].detach()
        sc = torch.log_softmax(logit, 1).gather(1, a_t.unsqueeze(1))
        return h, c, u, a_t
    _, _, a_t, u, sc = O.follower_step_tail(logit, x["is_valid"], None, "argmax", x["all_u_t"])   # follower.py:476-505
    return h, c, u, a_t


def cpu_steps(n_steps, warmup, threads, budget_s=None, min_s=0.0):
    """Times the reference AttnDecoderLSTM.forward + tail on the host cores (torch CPU).  Runs at least `min_s`
    seconds and `n_steps` steps, at most `budget_s` seconds; returns (steps/s, seconds, steps done)."""
    st = _ref_setup()
    xs, w, dec = st["xs"], st["w"], st["dec"]
    torch.set_num_threads(threads)
    xs = [dict(x, ctx_mask=x["ctx_mask"].bool()) for x in xs]
    h, c, u = xs[0]["h_0"], xs[0]["c_0"], xs[0]["u_t_prev"]
    t0, done, i = None, 0, 0
    with torch.no_grad():
        while True:
            if i == warmup:
                t0 = time.perf_counter()
            h, c, u, _ = _ref_step(st, dec, w, u, xs[i % 2], h, c)
            i += 1
            if i > warmup:
                done += 1
                el = time.perf_counter() - t0
                if budget_s is not None and el > budget_s:
                    break
                if done >= n_steps and el >= min_s:
                    break
    dt = time.perf_counter() - t0
    return done / dt, dt, done


def cpu_best_threads():
    """The reference is plain torch-CPU code: give it the thread count it runs fastest with on this host (intra-op
    parallelism of MKL GEMMs stops paying long before all hardware threads are used)."""
    n = os.cpu_count() or 1
    cands = sorted({t for t in (4, 8, 16, 32, 64, n) if t <= n})
    best, best_sps = cands[0], 0.0
    for t in cands:
        sps, _, _ = cpu_steps(6, 2, t)
        if sps > best_sps:
            best, best_sps = t, sps
    return best


def gpu_reference_steps(dev, seconds=2.0):
    """BASELINE.md §2 'reference-GPU row' — the 10x denominator of north_star: the reference's own modules moved to
    the GPU with stock torch kernels, (a) inputs already on the device, (b) including the reference's per-step host
    batching: np.stack of the B observation slabs + H2D of the [B,36,F] slab and of the dense [B,A,E] action tensor
    (follower.py:291-320, env.py:330-332).  Returns a dict for the bench line."""
    st = _ref_setup()
    w = {k: v.to(dev) for k, v in st["w"].items()}
    if st["dec"] is None:
        return {"unavailable": "oracle/_ref (the staged reference sources) is absent on this box"}
    from oracle import ref_loader
    dec = ref_loader.follower_decoder(st["w"], device=dev)
    xs_host = st["xs"]
    xs = [{k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in x.items()} for x in xs_host]
    xs = [dict(x, ctx_mask=x["ctx_mask"].bool()) for x in xs]
    # what the reference holds on the host after env.observe: one float32 [36,F] array per observation + the action rows
    slabs = [[np.ascontiguousarray(x["visual_context"][b].numpy()) for b in range(B)] for x in xs_host]
    acts = [np.ascontiguousarray(x["all_u_t"].numpy()) for x in xs_host]

    def run(with_host_batching):
        h, c, u = xs[0]["h_0"], xs[0]["c_0"], xs[0]["u_t_prev"]
        done, t0 = 0, None
        with torch.no_grad():
            i = 0
            while True:
                if i == 5:
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                x = xs[i % 2]
                if with_host_batching:
                    vis = torch.from_numpy(np.stack(slabs[i % 2])).to(dev)           # follower.py:291-298, env.py:330-332
                    allu = torch.from_numpy(acts[i % 2]).to(dev)                     # follower.py:300-320
                    x = dict(x, visual_context=vis, all_u_t=allu)
                h, c, u, a_t = _ref_step(st, dec, w, u, x, h, c)
                a_host = a_t.cpu()                                                   # follower.py:509-513: the env needs a_t
                i += 1
                if i > 5:
                    done += 1
                    if time.perf_counter() - t0 > seconds:
                        break
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return done / dt, done, dt

    dev_sps, dev_n, dev_dt = run(False)
    host_sps, host_n, host_dt = run(True)
    return {"value": host_sps, "unit": "steps/s", "value_device_inputs": dev_sps, "kind": st["kind"],
            "sample": "reference AttnDecoderLSTM.forward + rollout tail on this GPU with stock torch %s kernels, a_t read back "
                      "every step: %d steps (%.1f s) with the reference's per-step host np.stack + H2D of the slab and action "
                      "tensors (`value`), %d steps (%.1f s) with inputs already on the device (`value_device_inputs`)"
                      % (torch.__version__, host_n, host_dt, dev_n, dev_dt)}


def run_reference(args, rank, world):
    if rank != 0:
        return
    st = _ref_setup()
    threads = cpu_best_threads()
    # bounded sample: every step is the full B=100 workload; at least MIN_REF_SECONDS, at most ~60 s of CPU work
    sps, dt, done = cpu_steps(args.steps, max(args.warmup, 3) if args.warmup < 20 else 5, threads, budget_s=60.0,
                              min_s=MIN_REF_SECONDS)
    what = ("the reference's own tasks/R2R/model.py (oracle/_ref, staged unmodified)" if st["kind"] == "reference"
            else "torch-CPU oracle port of tasks/R2R/model.py (oracle/_ref absent)")
    line = {
        "impl": "reference", "metric": "follower decode-steps/sec", "value": sps, "unit": "steps/s",
        "n_gpus": args.gpus, "steps": done, "warmup": args.warmup, "ms_per_step": 1e3 / sps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args.gpus),
        "cpu_baseline": {"value": sps, "unit": "steps/s", "cores": threads, "kind": st["kind"],
                         "sample": "%d decode steps (%.1f s, never less than %.0f s) of the same workload, %s; "
                                   "thread count picked as the fastest of a short sweep up to %d hardware threads"
                                   % (done, dt, MIN_REF_SECONDS, what, os.cpu_count() or 1)},
        "e2e": {"value": sps, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------- GPU arm
_T0 = time.perf_counter()


def log(msg):
    print("[bench %.1fs] %s" % (time.perf_counter() - _T0, msg), file=sys.stderr, flush=True)


def run_gpu(args, rank, local_rank, world):
    from speaker_follower_b200 import ops, synth
    import __graft_entry__ as ge
    ge.build()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    for kv in filter(None, os.environ.get("SFB_OPTIONS", "").split(",")):   # e.g. SFB_OPTIONS=disable_pdl=1 (A/B runs)
        k, v = kv.split("=")
        ops.set_option(k, int(v))
        log("option %s=%s" % (k, v))

    E = F = synth.FEAT
    H = synth.HID
    w_cpu = synth.follower_decoder_weights()
    n_params = sum(v.numel() for v in w_cpu.values())
    w = {k: v.to(dev) for k, v in w_cpu.items()}
    g = torch.Generator(device=dev).manual_seed(1234 + rank)

    # device-resident feature store (3.1 GB) — synthetic pool5-like features
    table = torch.empty(N_VIEWPOINTS, 36, synth.IMG_DIM, device=dev)
    for i in range(0, N_VIEWPOINTS, 1024):
        table[i:i + 1024].normal_(generator=g).clamp_(min=0).mul_(1.1)
    store = ops.FeatureStore(table, synth.loc_embedding_table().to(dev))

    log("feature table ready")
    # rotating input sets
    vp = [torch.randint(0, N_VIEWPOINTS, (B,), device=dev, dtype=torch.int32, generator=g) for _ in range(POOL)]
    view = [torch.randint(0, 36, (B,), device=dev, dtype=torch.int32, generator=g) for _ in range(POOL)]
    ctx = [torch.tanh(torch.randn(B, L, H, device=dev, generator=g) * 0.6) for _ in range(N_CTX)]
    lens = torch.sort(torch.randint(10, L + 1, (B,), generator=torch.Generator().manual_seed(5)), descending=True)[0]
    lens[0] = L
    mask = (torch.arange(L).unsqueeze(0) >= lens.unsqueeze(1)).to(dev)
    # action candidates (env.py:60-75): candidate a of row b is view cview[b,a] of the current viewpoint's slab plus
    # the sin/cos of its relative heading/elevation; row 0 = stop = zeros.  The packed path gathers them on the device
    # from (cview, ctrig); the dense [B,A,E] tensor U is what the reference ships per step (in-place path / e2e bytes).
    U, valid, cview, ctrig = [], [], [], []
    for j in range(POOL):
        n_act = torch.randint(2, A + 1, (B,), generator=torch.Generator().manual_seed(70 + j))
        n_act[0] = A
        v = (torch.arange(A).unsqueeze(0) < n_act.unsqueeze(1)).float().to(dev)
        cv = torch.randint(0, 36, (B, A), device=dev, generator=g).int()
        cv = torch.where(v > 0, cv, torch.full_like(cv, -1))
        cv[:, 0] = -1
        ang = torch.rand(B, A, 2, device=dev, generator=g) * 6.28 - 3.14
        trig = torch.stack([torch.sin(ang[..., 0]), torch.cos(ang[..., 0]), torch.sin(ang[..., 1]), torch.cos(ang[..., 1])], 2).contiguous()
        cview.append(cv.contiguous()); ctrig.append(trig); valid.append(v.contiguous())
        if os.environ.get("SFB_INPLACE"):
            rows = table[vp[j].long().unsqueeze(1), cv.clamp(min=0).long()]   # [B,A,2048]
            u = torch.cat([rows, trig.repeat_interleave(32, dim=2)], 2) * (cv >= 0).unsqueeze(2)
            U.append(u.contiguous())
            del rows, u
        else:
            U.append(None)

    d = ops.follower_dims(w)
    # weights re-laid out once per weight version (sfb_follower_pack_weights); SFB_INPLACE=1 benches the in-place path
    packer = ops.PackedFollower()
    blob = None if os.environ.get("SFB_INPLACE") else packer.get(w)
    hbuf = [torch.tanh(torch.randn(B, H, device=dev, generator=g) * 0.5), torch.empty(B, H, device=dev)]
    cbuf = [torch.randn(B, H, device=dev, generator=g) * 0.5, torch.empty(B, H, device=dev)]
    ubuf = [torch.zeros(B, E, device=dev), torch.empty(B, E, device=dev)]
    alpha = torch.empty(B, L, device=dev); logit = torch.empty(B, A, device=dev); alpha_v = torch.empty(B, 36, device=dev)
    a_t = torch.empty(B, dtype=torch.int32, device=dev); score = torch.empty(B, device=dev)
    ws = torch.zeros(1 << 26, dtype=torch.uint8, device=dev)   # zero-filled once (semaphores)
    launches_per_step = [0]

    # state one step prepares for the next (visual query + packed gate-operand blocks), ping-pong
    qbuf = [ops.follower_carry(w, B), ops.follower_carry(w, B)] if blob is not None else [None, None]
    # per-episode projections of ctx (ctx is constant over the 10 decode steps of a rollout): recomputed INSIDE the
    # timed region at the start of every episode (project()), never cached across episodes
    use_proj = blob is not None and not os.environ.get("SFB_NO_CTXPROJ")
    cproj = [(torch.empty_like(c), torch.empty_like(c)) if use_proj else None for c in ctx]

    crows = ops.ctx_rows(lens.tolist(), L, dev) if use_proj else None   # only the un-padded positions are projected
    pws = torch.zeros(1 << 27, dtype=torch.uint8, device=dev) if use_proj else None
    if use_proj:
        for ck, co in cproj:
            ck.zero_(); co.zero_()

    def project(e):
        if use_proj:
            ops.follower_project_ctx(w, blob, ctx[e], out=cproj[e], rows=crows, workspace=pws)
            return ops.last_launch_count()
        return 0

    def step(i, first=False, e=0):
        j, s = i % POOL, i % 2
        if blob is None:   # in-place weights: 11 launches + separate tail
            ops.follower_step(w, ubuf[s], U[j], None, hbuf[s], cbuf[s], ctx[e], mask, store=store, vp_idx=vp[j],
                              view_idx=view[j], workspace=ws, out=(hbuf[s ^ 1], cbuf[s ^ 1], alpha, logit, alpha_v))
            n = ops.last_launch_count()
            ops.follower_tail(logit, valid[j], U[j], "argmax", out=(a_t, ubuf[s ^ 1], score, None))
            launches_per_step[0] = n + ops.last_launch_count()
            return
        # packed weights; carry_out of this step is the carry_in of the next one; rollout tail fused into the last kernel
        ops.follower_step(w, ubuf[s], None, None, hbuf[s], cbuf[s], ctx[e], mask, store=store, vp_idx=vp[j],
                          view_idx=view[j], workspace=ws, out=(hbuf[s ^ 1], cbuf[s ^ 1], alpha, logit, alpha_v),
                          packed=blob, carry_in=None if first else qbuf[s], carry_out=qbuf[s ^ 1],
                          cand_view=cview[j], cand_trig=ctrig[j], ctx_proj=cproj[e],
                          tail={"is_valid": valid[j], "feedback": "argmax", "out": (a_t, ubuf[s ^ 1], score, None)})
        launches_per_step[0] = ops.last_launch_count()

    log("inputs ready")
    # warm-up outside graphs (also configures kernel attributes), then capture one graph per episode: the per-episode
    # ctx projection followed by the 10 decode steps that share it
    side = torch.cuda.Stream(device=dev)
    proj_launches = 0
    with torch.cuda.stream(side):
        for e in range(N_CTX):
            proj_launches = project(e)
        for i in range(POOL):
            step(i, first=(i == 0))
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    if args.profile_steps:
        project(0)
        for i in range(args.profile_steps):
            step(i)
        torch.cuda.synchronize()
        log("profile steps done (%d launches per step + %d per episode)" % (launches_per_step[0], proj_launches))
        return
    episodes = []
    for e in range(N_CTX):
        gph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gph):
            project(e)
            for i in range(POOL):
                step(i, first=(i == 0), e=e)   # a new episode starts from fresh (h_0, u_begin): no carried state
        episodes.append(gph)
    singles = []
    for i in range(POOL):
        gph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gph):
            if i == 0:
                project(0)
            step(i, first=(i == 0), e=0)
        singles.append(gph)

    # the step kernel alone (the dominant kernel: one launch = one decode step): steps 1..POOL-1 of an episode chained in
    # one graph, no per-episode projection, fresh slabs / candidates per launch
    chain = None
    if blob is not None:
        chain = torch.cuda.CUDAGraph()
        with torch.cuda.graph(chain):
            for i in range(1, POOL):
                step(i, first=False, e=0)
    log("graphs captured")

    def run_steps(k):
        for n in range(k // POOL):
            episodes[n % N_CTX].replay()
        for i in range(k % POOL):
            singles[i].replay()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    run_steps(max(args.warmup, 3))
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_steps(args.steps)
    e1.record()
    barrier()
    clocks = sampler.result()
    from speaker_follower_b200 import dist as sfdist
    value, ms = sfdist.aggregate_rate(args.steps, e0.elapsed_time(e1), device=dev)   # whole job / slowest rank

    log("timed region done: %.3f ms/step" % (ms / args.steps))
    step_kernel_us = None
    if chain is not None and launches_per_step[0] == 1:
        for _ in range(3):
            chain.replay()
        torch.cuda.synchronize()
        ec0, ec1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_chain = 40
        ec0.record()
        for _ in range(n_chain):
            chain.replay()
        ec1.record(); torch.cuda.synchronize()
        step_kernel_us = ec0.elapsed_time(ec1) * 1e3 / (n_chain * (POOL - 1))
        log("step kernel alone: %.2f us per launch" % step_kernel_us)
    # ---- e2e: same step through the public ops API with HOST buffers (pinned), H2D + D2H inside the timed region.
    # Host inputs per step (what the agent holds on the host after env.observe): viewpoint / view indices, the
    # action-candidate embeddings + validity (follower.py:300-320); result read back: a_t + logits (follower.py:510).
    e2e_steps = min(args.steps, 2000)
    if blob is None:
        # in-place path: the reference's per-step host inputs (follower.py:291-320 minus the slab): indices, the dense
        # candidate embeddings and validity
        h_in = [[t.cpu().pin_memory() for t in (vp[j], view[j], U[j], valid[j])] for j in range(POOL)]
        d_in = [torch.empty_like(t) for t in (vp[0], view[0], U[0], valid[0])]
        d_vp, d_view, d_U, d_valid = d_in
        d_out = [a_t, logit]
        h_out = [torch.empty(B, dtype=torch.int32).pin_memory(), torch.empty(B, A).pin_memory()]
    else:
        # packed path: everything the agent holds on the host after env.observe fits ONE pinned staging buffer:
        # viewpoint / view indices, candidate view indices + 4 trig values, validity (20 KB); one H2D, one D2H
        n_i, n_c = B, B * A
        host_i = [torch.cat([vp[j].cpu().view(-1), view[j].cpu().view(-1), cview[j].cpu().view(-1)]).int() for j in range(POOL)]
        host_f = [torch.cat([ctrig[j].cpu().view(-1), valid[j].cpu().view(-1)]).float() for j in range(POOL)]
        h_in = [[torch.cat([hi.view(torch.uint8), hf.view(torch.uint8)]).pin_memory()] for hi, hf in zip(host_i, host_f)]
        stage = torch.empty_like(h_in[0][0], device=dev)
        d_in = [stage]
        ib = (2 * n_i + n_c) * 4
        di = stage[:ib].view(torch.int32)
        df = stage[ib:].view(torch.float32)
        d_vp, d_view, d_cview = di[:n_i], di[n_i:2 * n_i], di[2 * n_i:].view(B, A)
        d_ctrig, d_valid = df[:n_c * 4].view(B, A, 4), df[n_c * 4:].view(B, A)
        # result: a_t (what the agent needs on the host to step the simulator, follower.py:509-513) is written by the
        # last kernel straight into page-locked host memory; the host polls it instead of paying a stream-sync wake-up
        h_at = torch.full((B,), -1, dtype=torch.int32).pin_memory()
        np_at = h_at.numpy()
        a_t2, logit2 = h_at, logit
        d_out, h_out = [], [h_at]
    graphs2, graphs2_first = [], []          # [parity]; *_first = first step of an episode (ctx projection inside)
    for with_proj in (False, True):
        for s in range(2):
            gph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gph):
                if with_proj:
                    project(0)                         # new episode: per-episode ctx projections (2 launches)
                if blob is None:
                    ops.follower_step(w, ubuf[s], d_U, None, hbuf[s], cbuf[s], ctx[0], mask, store=store, vp_idx=d_vp,
                                      view_idx=d_view, workspace=ws, out=(hbuf[s ^ 1], cbuf[s ^ 1], alpha, logit, alpha_v))
                    ops.follower_tail(logit, d_valid, d_U, "argmax", out=(a_t, ubuf[s ^ 1], score, None))
                else:
                    ops.follower_step(w, ubuf[s], None, None, hbuf[s], cbuf[s], ctx[0], mask, store=store, vp_idx=d_vp,
                                      view_idx=d_view, workspace=ws, out=(hbuf[s ^ 1], cbuf[s ^ 1], alpha, logit2, alpha_v),
                                      packed=blob, carry_in=None if with_proj else qbuf[s], carry_out=qbuf[s ^ 1], cand_view=d_cview, cand_trig=d_ctrig,
                                      ctx_proj=cproj[0],
                                      tail={"is_valid": d_valid, "feedback": "argmax", "out": (a_t2, ubuf[s ^ 1], score, None)})
            (graphs2_first if with_proj else graphs2).append(gph)

    def e2e_step(i):
        j = i % POOL
        if blob is not None:
            np_at.fill(-1)
        for dst, src in zip(d_in, h_in[j]):            # (a copy node inside the step's graph was measured slower: 80.4 vs 78.1 us)
            dst.copy_(src, non_blocking=True)
        (graphs2_first if j == 0 else graphs2)[i % 2].replay()
        if blob is None:
            for dst, src in zip(h_out, d_out):
                dst.copy_(src, non_blocking=True)
            torch.cuda.current_stream().synchronize()  # the agent needs a_t on the host to step the simulator
        else:
            spins = 0
            while np_at.min() < 0:                     # all B actions have landed in host memory
                spins += 1
                if spins > 2_000_000:                  # ~seconds: something is wrong -> surface the CUDA error, do not hang
                    torch.cuda.synchronize()
                    if np_at.min() < 0:
                        raise RuntimeError("e2e: a_t never reached the host buffer")
                    break

    for i in range(4):
        e2e_step(i)
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for i in range(e2e_steps):
        e2e_step(i)
    e1.record()
    barrier()
    e2e_ms = torch.tensor([max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)], device=dev)
    if dist is not None:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = world * e2e_steps / (float(e2e_ms.item()) * 1e-3)
    h2d = sum(t.numel() * t.element_size() for t in h_in[0])
    d2h = sum(t.numel() * t.element_size() for t in h_out)

    log("e2e done")
    # ---- agent level: Seq2SeqAgent on a navigation-graph environment, (a) the whole rollout on the device (table-driven
    # environment, one CUDA graph per 10-step episode, nav kernels included), (b) the host loop with the Python
    # environment and one pinned staging buffer per step
    agent_stats = None
    if blob is not None and not os.environ.get("SFB_NO_AGENT"):
        try:
            from speaker_follower_b200 import follower as Fo, model as M
            from speaker_follower_b200.navgraph_env import DeviceNavTables, FakeR2RBatch, WorldState
            nenv = FakeR2RBatch(n_viewpoints=256, n_instr=B, batch_size=B, seed=9, max_len=L, with_features=False)
            glove = synth.follower_encoder_weights()["embedding.weight"].numpy()
            enc = M.EncoderLSTM(synth.VOCAB, synth.WORD, synth.HID, 0, 0.5, glove=glove).to(dev).eval()
            dec = M.AttnDecoderLSTM(synth.FEAT, synth.HID, 0.5).to(dev).eval()
            enc.load_state_dict(synth.follower_encoder_weights()); dec.load_state_dict(w_cpu)
            dec.feature_store = store
            agent = Fo.Seq2SeqAgent(nenv, "", enc, dec, episode_len=POOL, max_instruction_length=L)
            nav = DeviceNavTables(nenv, dev)
            nenv.reset(sort=True)
            batch = nenv.batch
            starts = nav.state_ids([WorldState("fake", it["start"], it["heading"], 0.0) for it in batch])
            res = agent.device_rollout(nav, starts, [it["goal"] for it in batch], [it["instr_encoding"] for it in batch], cuda_graph=True)
            gph = res["graph"]
            for _ in range(5):
                gph.replay()
            torch.cuda.synchronize()
            n_ep = 100
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ea.record()
            for _ in range(n_ep):
                gph.replay()
            eb.record(); torch.cuda.synchronize()
            dev_sps = n_ep * POOL / (ea.elapsed_time(eb) * 1e-3)
            agent.feedback = "argmax"
            with torch.no_grad():
                agent.rollout()
                torch.cuda.synchronize()
                t0 = time.perf_counter(); n_steps = 0
                for _ in range(5):
                    tr = agent.rollout()
                    n_steps += max(len(t["actions"]) for t in tr)
                torch.cuda.synchronize()
                host_sps = n_steps / (time.perf_counter() - t0)
            agent_stats = {"device_rollout_steps_per_s": dev_sps, "host_loop_steps_per_s": host_sps,
                           "how": "Seq2SeqAgent at B=%d on a 256-viewpoint navigation graph: device_rollout = table-driven environment "
                                  "(sfb_nav_step) + decode steps, one CUDA graph per %d-step episode, encoder excluded; host loop = "
                                  "agent.rollout() with the Python environment (index-only observations), encoder included" % (B, POOL)}
            log("agent level: %.0f steps/s on the device, %.0f steps/s host loop" % (dev_sps, host_sps))
        except Exception as e:  # the agent numbers are an extra: never lose the headline line over them
            agent_stats = {"error": repr(e)}
    # ---- roofline of the dominant kernel (vis_lstm_fused_kernel: attention gather + gate GEMM + LSTM cell), timed alone
    # with CUDA events on its launch stream; every launch reads a different random set of slabs from the 3.1 GB table
    # (inputs >> L2) and the step-invariant LSTM weights (35.7 + 4.2 MB fp32), which may be L2-resident as in a rollout.
    feat = torch.empty(B, F, device=dev)
    n_attn = 400
    vps = [torch.randint(0, N_VIEWPOINTS, (B,), device=dev, dtype=torch.int32, generator=g) for _ in range(n_attn)]
    hh1, cc1 = torch.empty(B, H, device=dev), torch.empty(B, H, device=dev)

    def ka(i):
        if blob is not None:
            ops.follower_gather_lstm(w, blob, qbuf[0], cbuf[0], store=store, vp_idx=vps[i], view_idx=view[0],
                                     out=(hh1, cc1, alpha_v), workspace=ws)
        else:
            ops.visual_attention_core(torch.zeros(B, F, device=dev), None, store=store, vp_idx=vps[i], view_idx=view[0],
                                      out=(feat, alpha_v), workspace=ws)
    for i in range(5):
        ka(i)
    torch.cuda.synchronize()
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ea.record()
    for i in range(n_attn):
        ka(i)
    eb.record()
    torch.cuda.synchronize()
    attn_ms = ea.elapsed_time(eb) / n_attn
    lstm_w_bytes = 4 * (4 * H * (E + F) + 4 * H * H + 8 * H)
    attn_bytes = 4 * B * 36 * F + (lstm_w_bytes if blob is not None else 0) + (4 * B * (E + 4 * H) if blob is not None else 0)
    hbm_peak, peak_src = peaks()
    attn_gbs = attn_bytes / (attn_ms * 1e-3) / 1e9
    step_bytes = algorithmic_bytes(E, F, H, n_params)
    step_gbs = step_bytes / (ms / args.steps * 1e-3) / 1e9

    log("attention kernel timed: %.2f us" % (attn_ms * 1e3))
    if rank == 0:
        threads = cpu_best_threads()
        cpu_sps, cpu_dt, cpu_n = cpu_steps(4000, 3, threads, budget_s=15.0)
        gpu_base = gpu_reference_steps(dev)
        log("reference-GPU baseline done")
        line = {
            "metric": "follower decode-steps/sec", "value": value, "unit": "steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(world),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps},
            "agent": agent_stats,
            "gpu_launches": launches_per_step[0] * args.steps + proj_launches * (args.steps // POOL + (1 if args.steps % POOL else 0)),
            "roofline": ({"kernel": "step_kernel (the whole decode step in one launch)", "bound": "hbm",
                          "achieved": step_bytes / (step_kernel_us * 1e-6) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                          "frac": step_bytes / (step_kernel_us * 1e-6) / 1e9 / hbm_peak, "traffic": STEP_NCU_TRAFFIC,
                          "peak_source": peak_src, "bytes_per_launch": step_bytes, "us_per_launch": step_kernel_us,
                          "how": "360 launches (40 replays of a graph of 9 chained steps, no per-episode projection) between two "
                                 "CUDA events, fresh slabs and candidates per launch; algorithmic bytes = SURVEY 8d's 104.9 MB per "
                                 "step; traffic = dram read+write of one launch from profiles/r02_ncu_full_summary.txt (cold caches)"}
                         if step_kernel_us else None),
            "roofline_first_half": {"kernel": "vis_lstm_fused_kernel (36-view attention gather + gate GEMM + LSTM cell) launched alone",
                         "bound": "hbm",
                         "achieved": attn_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": attn_gbs / hbm_peak,
                         "traffic": ATTN_NCU_TRAFFIC, "peak_source": peak_src, "bytes_per_launch": attn_bytes,
                         "us_per_launch": attn_ms * 1e3,
                         "how": "%d back-to-back launches between two CUDA events, fresh slabs per launch; algorithmic bytes = slabs "
                                "4*B*36*F + fp32 LSTM weights 4*(4H*(E+F)+4H*H+8H) + states 4*B*(E+4H) (DESIGN.md section 6); traffic = "
                                "dram read+write per launch (ncu --set full, cold caches)" % n_attn},
            "roofline_step": {"bound": "hbm", "achieved": step_gbs, "peak": hbm_peak, "unit": "GB/s",
                              "frac": step_gbs / hbm_peak, "bytes_per_step": step_bytes},
            "cpu_baseline": {"value": cpu_sps, "unit": "steps/s", "cores": threads, "kind": _CPU_STATE["kind"],
                             "sample": "%d decode steps (%.1f s) of the same workload, %s, fastest thread count of a short sweep"
                                       % (cpu_n, cpu_dt, "the reference's own tasks/R2R/model.py (oracle/_ref)"
                                          if _CPU_STATE["kind"] == "reference" else "torch-CPU oracle port of tasks/R2R/model.py")},
            "gpu_baseline": gpu_base if "unavailable" in gpu_base else dict(gpu_base, vs={
                "value_over_gpu_baseline": value / world / gpu_base["value"],
                "e2e_over_gpu_baseline": e2e_value / world / gpu_base["value"],
                "value_over_gpu_baseline_device_inputs": value / world / gpu_base["value_device_inputs"],
                "e2e_over_gpu_baseline_device_inputs": e2e_value / world / gpu_base["value_device_inputs"]}),
        }
        if line["roofline"] is None:   # the step was not one launch on this shape / with these options
            line["roofline"] = line["roofline_first_half"]
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------- C4 / C5 (BASELINE.json configs[3], [4])
def _agents(env, dev, beam=1, instruction_len=80):
    from speaker_follower_b200 import follower as Fo, model as M, ops, speaker as Sp, synth
    glove = synth.follower_encoder_weights()["embedding.weight"].numpy()
    enc = M.EncoderLSTM(synth.VOCAB, synth.WORD, synth.HID, 0, 0.5, glove=glove).to(dev).eval()
    dec = M.AttnDecoderLSTM(synth.FEAT, synth.HID, 0.5).to(dev).eval()
    enc.load_state_dict(synth.follower_encoder_weights()); dec.load_state_dict(synth.follower_decoder_weights())
    dec.feature_store = ops.FeatureStore(torch.from_numpy(env.table).to(dev), torch.from_numpy(env.loc).to(dev))
    follower = Fo.Seq2SeqAgent(env, "", enc, dec, episode_len=10, max_instruction_length=80)
    swd = synth.speaker_decoder_weights()
    senc = M.SpeakerEncoderLSTM(synth.FEAT, synth.FEAT, synth.HID, 0.5).to(dev).eval()
    sdec = M.SpeakerDecoderLSTM(synth.VOCAB, synth.WORD, synth.HID, 0.5, glove=swd["embedding.weight"].numpy()).to(dev).eval()
    senc.load_state_dict(synth.speaker_encoder_weights()); sdec.load_state_dict(swd)
    senc.feature_store = dec.feature_store   # rescoring batches carry (viewpoint, view) indices instead of T x N slabs
    speaker = Sp.Seq2SeqSpeaker(env, "", senc, sdec, instruction_len=instruction_len, max_episode_len=10)
    return follower, speaker


def run_pragmatic(args, rank, local_rank, world):
    """--config c4: state-factored search (completion 40, successor 1) over 64 instructions on a navigation-graph
    environment, speaker rescoring of every candidate (up to 2 560), sharded by instruction over the ranks; one
    all-gather of the score records + one all-reduce of the statistics; every rank forms the weighted argmax.
    --config c5: greedy speaker generation over synthetic trajectories, sharded, JSON records gathered in order."""
    import __graft_entry__ as ge
    ge.build()
    from speaker_follower_b200 import dist as sfdist, pragmatic as PR
    from speaker_follower_b200.navgraph_env import FakeR2RBatch
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    c4 = args.config == "c4"
    n_inst = 64 if c4 else 256 * max(world, 1) * 2
    beam = 40 if c4 else 1

    def make_env():
        # index-only observations: slabs and action-embedding rows are gathered from the device feature store, the
        # environment ships (viewpoint row, view index, candidate view indices + angles) only
        return FakeR2RBatch(n_viewpoints=160, n_instr=n_inst, batch_size=64 if c4 else 256, seed=77, max_len=40,
                            beam_size=max(beam, 1), with_features=False)

    def one_pass(shard):
        env = make_env()
        # C4: instructions strided over the ranks (search + rescoring are per instruction); C5: whole minibatches dealt out
        # (a generated instruction depends on its minibatch's longest path, in the reference too)
        gi = PR.shard_env(env, whole_batches=not c4) if shard else list(range(n_inst))
        follower, speaker = _agents(env, dev, instruction_len=30)
        nav = None
        if c4 and not os.environ.get("SFB_HOST_SEARCH"):
            from speaker_follower_b200.navgraph_env import DeviceNavTables
            nav = DeviceNavTables(env, dev, with_teacher=False)   # environment set-up, like the feature-store upload
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if c4:
            by_w, cands, records = PR.run_rational_follower(env, follower, speaker, beam_size=beam, global_index=gi, nav=nav)
            out = {w: {int(k): int(v) for k, v in ch.items()} for w, ch in PR.rational_combine(records).items()}
            n_cand = int(records.shape[0])
        else:
            merged = PR.generate_speaker_instructions(env, speaker, global_index=gi)
            out, n_cand = merged, len(merged)
        torch.cuda.synchronize()
        return out, n_cand, time.perf_counter() - t0

    ref_out = None
    if world > 1 and rank == 0:      # the single-process answer (untimed), to compare the sharded one against
        saved = (sfdist.world,)
        sfdist.world = lambda: (0, 1)
        try:
            ref_out, _, _ = one_pass(False)
        finally:
            sfdist.world = saved[0]
    if dist is not None:
        dist.barrier()
    one_pass(True)                   # warm-up (kernel attributes, allocator)
    if dist is not None:
        dist.barrier()
    out, n_cand, dt = one_pass(True)
    t = torch.tensor([dt], device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    if rank == 0:
        same = None if ref_out is None else (ref_out == out)
        line = {"metric": "pragmatic-inference instructions/sec" if c4 else "speaker-generated trajectories/sec",
                "value": n_inst / dt, "unit": "instr/s" if c4 else "traj/s", "n_gpus": world, "steps": 1, "warmup": 1,
                "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": ("C4: state-factored search (search state on the device), completion 40, successor 1, 64 instructions, %d candidates "
                                        "rescored by the speaker" % n_cand) if c4 else
                                       ("C5: greedy speaker generation over %d synthetic trajectories, batch 256 per GPU" % n_inst),
                           "env": "navigation-graph stand-in (160 viewpoints), random-init weights",
                           "parallelism": "instances strided over %d ranks; collectives: all_gather(score records) + "
                                          "all_reduce(n, sum, sum^2)" % world if c4 else
                                          "minibatches of 256 dealt out over %d ranks; collectives: all_gather(JSON bytes)" % world},
                "identical_to_single_process": same}
        if c4:   # where the time of rank 0's pass went (wall clock, device synchronised at the boundaries)
            line["phases_rank0"] = {k: round(v, 4) for k, v in PR.LAST_PHASES.items()}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------- C2 (BASELINE.json configs[1]: training step)
def run_train(args, rank, local_rank, world):
    """--config c2: one follower training iteration = encoder forward, 10 teacher-forced decode steps with dropout 0.5,
    summed cross-entropy, backward through everything, Adam step (follower.py:1001-1020) at B=100, L=80.  Forward AND
    backward run on this library's kernels (hand-written backward: backward.cu); torch contributes the loss terms, the
    gradient accumulation of autograd and the optimizer."""
    import __graft_entry__ as ge
    ge.build()
    from speaker_follower_b200 import model as M, synth
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    T = POOL
    glove = synth.follower_encoder_weights()["embedding.weight"].numpy()
    enc = M.EncoderLSTM(synth.VOCAB, synth.WORD, synth.HID, 0, 0.5, glove=glove).to(dev).train()
    dec = M.AttnDecoderLSTM(synth.FEAT, synth.HID, 0.5).to(dev).train()
    enc.load_state_dict(synth.follower_encoder_weights()); dec.load_state_dict(synth.follower_decoder_weights())
    eo = torch.optim.Adam([p for p in enc.parameters() if p.requires_grad], lr=1e-4)
    do = torch.optim.Adam(dec.parameters(), lr=1e-4)
    seq, mask, lengths = synth.instruction_batch(B, L, seed=41)
    seq, mask = seq.to(dev), mask.to(dev)
    g = torch.Generator(device=dev).manual_seed(7 + rank)
    steps = []
    for t in range(T):
        x = synth.follower_step_inputs(B, L, A, seed=300 + t, n_viewpoints=256)
        steps.append({"U": x["all_u_t"].to(dev), "V": x["visual_context"].to(dev), "valid": x["is_valid"].to(dev),
                      "target": torch.randint(0, 2, (B,), device=dev, generator=g)})

    def iteration():
        eo.zero_grad(set_to_none=True); do.zero_grad(set_to_none=True)
        ctx, h, c = enc(seq, lengths)
        u = dec.u_begin.expand(B, -1)
        loss = torch.zeros((), device=dev)
        for st in steps:
            h, c, alpha, logit, av = dec(u, st["U"], st["V"], h, c, ctx, mask)
            lg = logit.masked_fill(st["valid"] == 0, -float("inf"))
            loss = loss + torch.nn.functional.cross_entropy(lg, st["target"])
            u = st["U"][torch.arange(B, device=dev), st["target"]].detach()
        loss.backward()
        eo.step(); do.step()
        return loss

    for _ in range(max(3, min(args.warmup, 5))):
        iteration()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    n = max(5, min(args.steps, 30))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        loss = iteration()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    t = torch.tensor([ms], device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    if rank == 0:
        # the reference's CPU path for the same iteration (oracle port of model.py under torch autograd), bounded sample
        cpu_ms = None
        try:
            from oracle import r2r_oracle as O
            we = {k: v.clone().requires_grad_(k != "embedding.weight") for k, v in synth.follower_encoder_weights().items()}
            wd = {k: v.clone().requires_grad_(True) for k, v in synth.follower_decoder_weights().items()}
            cs = [{k: v.cpu() for k, v in st.items()} for st in steps[:T]]
            t0 = time.perf_counter(); reps = 0
            while time.perf_counter() - t0 < 10.0 and reps < 3:
                ctxc, hc, cc = O.encoder_lstm(seq.cpu()[:, :max(lengths)], lengths, we)
                uc = torch.zeros(B, synth.FEAT)
                lc = torch.zeros(())
                for st in cs:
                    hc, cc, _, lg, _ = O.attn_decoder_step(uc, st["U"], st["V"], hc, cc, ctxc, mask.cpu(), wd)
                    lc = lc + torch.nn.functional.cross_entropy(lg.masked_fill(st["valid"] == 0, -float("inf")), st["target"])
                    uc = st["U"][torch.arange(B), st["target"]]
                lc.backward()
                reps += 1
            cpu_ms = (time.perf_counter() - t0) * 1e3 / max(reps, 1)
        except Exception as e:  # pragma: no cover
            cpu_ms = None
        line = {"metric": "follower training iterations/sec", "value": world * 1e3 / ms, "unit": "iter/s", "n_gpus": world, "steps": n,
                "warmup": max(3, min(args.warmup, 5)), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "C2: follower training iteration, B=%d, L=%d, %d teacher-forced decode steps, dropout 0.5, "
                                       "encoder + decoder forward and hand-written backward, Adam step" % (B, L, T),
                           "batch": B, "instr_len": L, "actions": A},
                "final_loss": float(loss),
                "cpu_baseline": {"value": (1e3 / cpu_ms) if cpu_ms else None, "unit": "iter/s", "ms_per_iter": cpu_ms, "cores": torch.get_num_threads(),
                                 "kind": "port", "sample": "the same iteration (fwd + bwd, no optimizer) through the torch-CPU oracle port under autograd"}}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def run_speaker(args, rank, local_rank, world):
    """--config c3: one speaker scoring pass (speaker.py:123-202) at the C3 shape — N=256 paths of T=6 steps through
    SpeakerEncoderLSTM, then 80 teacher-forced words through SpeakerDecoderLSTM — through the module API of
    speaker_follower_b200.model, replayed from one CUDA graph.  Launch-latency bound at this size: the
    line reports the time per word step and the algorithmic bytes of a decoder step beside it, not a roofline claim."""
    import __graft_entry__ as ge
    ge.build()
    from speaker_follower_b200 import model as M, synth
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    N, T, W = 256, 6, 80
    we, wd = synth.speaker_encoder_weights(), synth.speaker_decoder_weights()
    enc = M.SpeakerEncoderLSTM(synth.FEAT, synth.FEAT, synth.HID, 0.5).to(dev).eval()
    dec = M.SpeakerDecoderLSTM(synth.VOCAB, synth.WORD, synth.HID, 0.5, glove=wd["embedding.weight"].numpy()).to(dev).eval()
    enc.load_state_dict(we, strict=False); dec.load_state_dict(wd, strict=False)
    g = torch.Generator().manual_seed(5 + rank)
    x = synth.follower_step_inputs(N, 8, 6, seed=3)
    feats = [(x["visual_context"] * (0.9 + 0.02 * t)).to(dev) for t in range(T)]          # T x [N, 36, 2176]
    acts = [x["all_u_t"][:, 1 + (t % 4)].contiguous().to(dev) for t in range(T)]          # T x [N, 2176]
    mask = (torch.arange(T).unsqueeze(0) >= torch.randint(4, T + 1, (N, 1), generator=g)).to(dev)
    words = torch.randint(4, synth.VOCAB, (N, W), generator=g).to(dev)

    def scoring_pass():
        with torch.no_grad():
            ctx, h, c = enc(acts, feats)
            score = torch.zeros(N, device=dev)
            for w in range(W):
                h, c, alpha, logit = dec(words[:, w:w + 1], h, c, ctx, mask)
                score = score + torch.log_softmax(logit, 1).gather(1, words[:, w:w + 1]).squeeze(1)
        return score

    for _ in range(3):
        ref_score = scoring_pass()
    torch.cuda.synchronize()
    launch = "eager"
    replay = scoring_pass
    try:
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            out = scoring_pass()
        gr.replay(); torch.cuda.synchronize()
        assert torch.allclose(out, ref_score, atol=1e-3, rtol=1e-4)
        replay, launch = gr.replay, "one CUDA graph per scoring pass"
    except Exception as e:  # pragma: no cover
        launch = "eager (graph capture failed: %s)" % str(e)[:80]
    if dist is not None:
        dist.barrier()
    n = max(5, min(args.steps, 50))
    for _ in range(3):
        replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        replay()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    t = torch.tensor([ms], device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    if rank == 0:
        cpu = {"value": None, "unit": "word-steps/s", "cores": torch.get_num_threads(), "kind": "reference", "sample": "unavailable"}
        try:
            from oracle import make_ref, ref_loader
            make_ref.make()
            mods = ref_loader.speaker_modules(we, wd)
            if mods is not None:
                renc, rdec = mods
                fc, ac, mc, wc = [f.cpu() for f in feats], [a.cpu() for a in acts], mask.cpu().bool(), words.cpu()
                torch.cuda.disabled = True      # the reference's own switch (utils.py:195-204, --no_cuda): keep its zeros on the host
                t0 = time.perf_counter()
                with torch.no_grad():
                    ctx, h, c = renc(ac, fc)
                    nw = 0
                    while nw < W and time.perf_counter() - t0 < 20.0:
                        h, c, alpha, logit = rdec(wc[:, nw:nw + 1], h, c, ctx, mc)
                        nw += 1
                dt = time.perf_counter() - t0
                torch.cuda.disabled = False
                cpu.update(value=nw / dt, sample="encoder (T=%d) + %d of %d word steps of the same pass (%.1f s), the reference's own "
                                                "model.py (oracle/_ref) on the host cores" % (T, nw, W, dt))
        except Exception as e:  # pragma: no cover
            cpu["sample"] = "failed: %s" % str(e)[:120]
        P_dec = 2961887
        dec_bytes = 4 * (N * T * synth.HID + P_dec + N * synth.WORD + 4 * N * synth.HID + N * synth.VOCAB)
        line = {"metric": "speaker scoring word-steps/sec", "value": world * W * 1e3 / ms, "unit": "word-steps/s", "n_gpus": world, "steps": n,
                "warmup": 3, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": "C3: speaker scoring pass, N=%d paths, T=%d path steps (SpeakerEncoderLSTM) + %d teacher-forced "
                                       "words (SpeakerDecoderLSTM + log-softmax gather), per GPU" % (N, T, W),
                           "launch": launch, "parallelism": "replicas x%d" % world},
                "us_per_word_step_incl_encoder": ms * 1e3 / W,
                "decoder_step_algorithmic_bytes": dec_bytes,
                "note": "latency-bound at this size: %.1f MB per decoder step would take %.1f us at the measured HBM peak" % (
                    dec_bytes / 1e6, dec_bytes / (peaks()[0] * 1e3)),
                "cpu_baseline": cpu}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="decode", choices=["decode", "c2", "c3", "c4", "c5"],
                    help="decode (default): the headline decode-steps/sec; c2: training step; c3: speaker scoring pass; "
                         "c4: pragmatic inference; c5: data augmentation")
    ap.add_argument("--profile-steps", type=int, default=0,
                    help="run this many eager (no CUDA graph) steps after warm-up and exit: the command ncu wraps")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.config in ("c4", "c5"):
        run_pragmatic(args, rank, local_rank, world)
    elif args.config == "c2":
        run_train(args, rank, local_rank, world)
    elif args.config == "c3":
        run_speaker(args, rank, local_rank, world)
    else:
        run_gpu(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
