"""A small deterministic stand-in for tasks/R2R/env.py:R2RBatch over a plain navigation graph (no renderer).

The Matterport3D simulator is off the timed path and cannot be built here, so agent-level
tests and the C4 / C5 bench configurations run on navigation graphs (random, or the real R2R scans of
tests/golden/nav_graphs.npz) that honour the reference's observation-dict contract (env.py:775-794): keys `instr_id`,
`scan`, `viewpoint`, `viewIndex`, `heading`, `elevation`, `feature` (list of one [36,2176] float32 array: image
part from the feature table, orientation part `loc[viewIndex]`), `adj_loc_list` (entry 0 = stay, others sorted by
|rel_heading|), `action_embedding` ([len(adj), 2176], row 0 zeros, env.py:60-75), `teacher` (index into
adj_loc_list of the next node on the shortest path, 0 at the goal), `instr_encoding`, `instr_length`.
Extra (not in the reference): `vp_index` = row of the viewpoint in the feature table, so that agents may gather
slabs on the device instead of copying them from the host.
"""
from __future__ import annotations

import math
from collections import deque, namedtuple

import numpy as np

from speaker_follower_b200 import synth

WorldState = namedtuple("WorldState", ["scanId", "viewpointId", "heading", "elevation"])   # env.py:227


class FakeImageFeatures:
    """Plays the role of env.py:MeanPooledImageFeatures for `env.image_features_list`."""
    feature_dim = synth.IMG_DIM

    def __init__(self, table, loc):
        self.table, self.loc = table, loc

    def get_name(self):
        return "fake_mean_pooled"


class FakeR2RBatch:
    def __init__(self, n_viewpoints=24, n_instr=16, batch_size=8, seed=0, max_len=20, vocab=synth.VOCAB, beam_size=1,
                 img_dim=synth.IMG_DIM, graph=None, with_features=True):
        """graph: None = random ring + chords; a scan id from tests/golden/nav_graphs.npz (written from the reference's
        connectivity/*.json by tests/golden/make_nav_graphs.py) = that REAL R2R navigation graph, edge headings and
        elevations derived from the viewpoint positions (heading 0 = +y, clockwise, like the simulator)."""
        g = np.random.Generator(np.random.PCG64(seed))
        self.g = g
        self.batch_size, self.beam_size = batch_size, beam_size
        # with_features=False: observations carry indices only (vp_index, viewIndex, adj_loc_list) — the configuration
        # in which the slabs live in a device feature store and nothing of size 36 x 2176 is touched on the host
        self.with_features = with_features
        real = None
        if graph is not None:
            import os
            z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "nav_graphs.npz"))
            real = (z[graph + "/pos"], z[graph + "/adj"])
            n_viewpoints = real[0].shape[0]
        self.table = synth.feature_table(n_viewpoints, seed + 5, img_dim).numpy()          # [V,36,img]
        self.loc = synth.loc_embedding_table().numpy()                                      # [36,36,128]
        self.image_features_list = [FakeImageFeatures(self.table, self.loc)]
        self.splits = ["fake"]
        self.print_progress = False
        # random connected graph: ring + chords; every edge gets a direction
        self.adj = {v: {} for v in range(n_viewpoints)}

        def connect(a, b):
            if a == b or b in self.adj[a]:
                return
            head = float(g.uniform(0, 2 * math.pi))
            self.adj[a][b] = (head, float(g.uniform(-0.4, 0.4)))
            self.adj[b][a] = ((head + math.pi) % (2 * math.pi), -self.adj[a][b][1])

        if real is None:
            for v in range(n_viewpoints):
                connect(v, (v + 1) % n_viewpoints)
            for _ in range(n_viewpoints):
                connect(int(g.integers(n_viewpoints)), int(g.integers(n_viewpoints)))
        else:
            pos, adj = real
            for a in range(n_viewpoints):
                for b in range(n_viewpoints):
                    if adj[a, b] and a != b:
                        d = pos[b] - pos[a]
                        head = math.atan2(d[0], d[1]) % (2 * math.pi)
                        self.adj[a][b] = (float(head), float(math.atan2(d[2], math.hypot(d[0], d[1]))))
        self.dist = {v: self._bfs(v) for v in range(n_viewpoints)}
        # instructions: start, goal, random tokens
        self.data = []
        for i in range(n_instr):
            start = int(g.integers(n_viewpoints))
            reach = sorted(self.dist[start])                       # real graphs may have several components
            goal = int(reach[int(g.integers(len(reach)))])
            n_tok = int(g.integers(5, max_len))
            self.data.append({"instr_id": "%d_0" % i, "path_id": i, "scan": "fake", "start": start, "goal": goal,
                              "heading": float(g.integers(12)) * (math.pi / 6),   # step * increment, the one expression every
                              # discretised heading is formed with (MatterSim.cpp:339-367 snaps the same way): a state
                              # reached by stepping has the SAME float as the state a reset produces
                              "instr_encoding": g.integers(4, vocab, size=n_tok).astype(np.int64)})
        self.ix = 0
        self.batch = None

    def _bfs(self, src):
        d = {src: 0}
        q = deque([src])
        while q:
            u = q.popleft()
            for w in self.adj[u]:
                if w not in d:
                    d[w] = d[u] + 1
                    q.append(w)
        return d

    # ---- minibatching (env.py:707-740)
    def reset_epoch(self):
        self.ix = 0

    def set_beam_size(self, n):
        self.beam_size = n

    def _next_minibatch(self, sort):
        batch = [self.data[(self.ix + k) % len(self.data)] for k in range(self.batch_size)]
        self.ix = (self.ix + self.batch_size) % len(self.data)
        if sort:
            batch = sorted(batch, key=lambda it: len(it["instr_encoding"]), reverse=True)   # env.py:733-734
        self.batch = batch

    def reset(self, sort=False, beamed=False, load_next_minibatch=True):
        if load_next_minibatch or self.batch is None:
            self._next_minibatch(sort)
        ws = [WorldState("fake", it["start"], it["heading"], 0.0) for it in self.batch]
        return [[w] for w in ws] if beamed else ws

    # ---- observations (env.py:763-804)
    @staticmethod
    def _view_index(heading, elevation):
        return (int(round(elevation / (math.pi / 6))) + 1) * 12 + int(round(heading / (math.pi / 6))) % 12

    def _observe_one(self, ws, item, include_teacher=True):
        v = ws.viewpointId
        view = self._view_index(ws.heading, ws.elevation)
        feat = np.concatenate((self.table[v], self.loc[view]), axis=1).astype(np.float32) if self.with_features else None
        adj = [{"absViewIndex": -1, "nextViewpointId": v, "rel_heading": 0.0, "rel_elevation": 0.0, "distance": 0.0}]
        others = []
        for w, (head, elev) in self.adj[v].items():
            rel = (head - ws.heading + math.pi) % (2 * math.pi) - math.pi
            others.append({"absViewIndex": 12 + int(round(head / (math.pi / 6))) % 12, "nextViewpointId": w,
                           "rel_heading": rel, "rel_elevation": elev - ws.elevation, "distance": 1.0, "abs_heading": head})
        adj += sorted(others, key=lambda a: abs(a["rel_heading"]))                                # env.py:218-222
        emb = np.zeros((len(adj), self.table.shape[2] + self.loc.shape[2]), np.float32) if self.with_features else None   # env.py:60-75
        img = self.table.shape[2]
        for a, d in enumerate(adj):
            if a == 0 or emb is None:
                continue
            emb[a, :img] = self.table[v, d["absViewIndex"]]
            emb[a, img:img + 32] = math.sin(d["rel_heading"]); emb[a, img + 32:img + 64] = math.cos(d["rel_heading"])
            emb[a, img + 64:img + 96] = math.sin(d["rel_elevation"]); emb[a, img + 96:] = math.cos(d["rel_elevation"])
        ob = {"instr_id": item["instr_id"], "scan": "fake", "viewpoint": v, "viewIndex": view, "heading": ws.heading,
              "elevation": ws.elevation, "feature": [feat], "step": 0, "adj_loc_list": adj, "action_embedding": emb,
              "navigableLocations": adj, "instructions": "", "instr_encoding": item["instr_encoding"],
              "instr_length": len(item["instr_encoding"]), "vp_index": v}
        if include_teacher:
            goal = item["goal"]
            if v == goal:
                ob["teacher"] = 0
            else:
                best = min(range(1, len(adj)), key=lambda a: (self.dist[goal].get(adj[a]["nextViewpointId"], 1e9), a))
                ob["teacher"] = best
        return ob

    def observe(self, world_states, beamed=False, include_teacher=True):
        if beamed:
            return [[self._observe_one(w, it, include_teacher) for w in beam] for beam, it in zip(world_states, self.batch)]
        return [self._observe_one(w, it, include_teacher) for w, it in zip(world_states, self.batch)]

    # ---- transitions (env.py:628-641)
    def _step_one(self, ws, action, ob):
        action = int(action)
        if action <= 0:
            return ws
        d = ob["adj_loc_list"][action]
        return WorldState("fake", d["nextViewpointId"], round(d["abs_heading"] / (math.pi / 6)) % 12 * (math.pi / 6), 0.0)

    def step(self, world_states, actions, last_obs, beamed=False):
        if beamed:
            return [[self._step_one(w, a, o) for w, a, o in zip(ws, acts, obs)]
                    for ws, acts, obs in zip(world_states, actions, last_obs)]
        return [self._step_one(w, a, o) for w, a, o in zip(world_states, actions, last_obs)]

    # ---- speaker side: gold (shortest-path) rollouts (env.py:850)
    def gold_obs_actions_and_instructions(self, max_steps, load_next_minibatch=True):
        ws = self.reset(sort=False, load_next_minibatch=load_next_minibatch)
        path_obs = [[] for _ in ws]
        path_actions = [[] for _ in ws]
        done = [False] * len(ws)
        for _ in range(max_steps):
            obs = self.observe(ws)
            acts = []
            for i, ob in enumerate(obs):
                if done[i]:
                    acts.append(0)
                    continue
                path_obs[i].append(ob)
                path_actions[i].append(ob["teacher"])
                acts.append(ob["teacher"])
                if ob["teacher"] == 0:
                    done[i] = True
            ws = self.step(ws, acts, obs)
            if all(done):
                break
        final = self.observe(ws)
        for i in range(len(ws)):
            path_obs[i].append(final[i])          # len(obs) == len(actions) + 1 (speaker.py:98)
        return path_obs, path_actions, [it["instr_encoding"] for it in self.batch]


class DeviceNavTables:
    """the environment as device-resident look-up tables.  ``observe`` is a pure function of the
    discretised world state (viewpoint, heading bin; env.py:763-804 + MatterSim's 30-degree discretisation) and ``step``
    a pure function of (state, action) (env.py:628-641), so both are tabulated once per environment: per state the
    feature-table row + view index of the slab, the candidate list (view index, sin/cos of relative heading /
    elevation, count), the successor of every candidate and the teacher action per goal.  A rollout then needs no host
    round trip at all: ``sfb_nav_step`` advances the states with the actions the decode step left on the device and
    writes the next step's inputs."""
    HEADINGS = 12

    def __init__(self, env: FakeR2RBatch, device, a_cap: int = 16, with_teacher: bool = True):
        import torch
        nvp = len(env.adj)
        S = nvp * self.HEADINGS
        dummy = {"instr_id": "", "instr_encoding": np.zeros(1, np.int64), "goal": 0}
        vp = np.zeros(S, np.int32); view = np.zeros(S, np.int32); nvalid = np.zeros(S, np.int32)
        cv = np.full((S, a_cap), -1, np.int32); trig = np.zeros((S, a_cap, 4), np.float32)
        nxt = np.zeros((S, a_cap), np.int32)
        teach = np.zeros((S, nvp), np.int32) if with_teacher else None
        keep = env.with_features
        env.with_features = False
        try:
            for v in range(nvp):
                for hb in range(self.HEADINGS):
                    s = v * self.HEADINGS + hb
                    ob = env._observe_one(WorldState("fake", v, hb * (math.pi / 6), 0.0), dummy, include_teacher=False)
                    adj = ob["adj_loc_list"]
                    if len(adj) > a_cap:
                        raise ValueError("state with %d candidates > a_cap" % len(adj))
                    vp[s], view[s], nvalid[s] = v, ob["viewIndex"], len(adj)
                    nxt[s, :] = s
                    for a, d in enumerate(adj):
                        if a == 0:
                            continue
                        cv[s, a] = d["absViewIndex"]
                        trig[s, a] = (math.sin(d["rel_heading"]), math.cos(d["rel_heading"]),
                                      math.sin(d["rel_elevation"]), math.cos(d["rel_elevation"]))
                        nh = round(d["abs_heading"] / (math.pi / 6)) % 12
                        nxt[s, a] = d["nextViewpointId"] * self.HEADINGS + nh
                    if teach is not None:
                        for goal in range(nvp):
                            if v != goal:
                                teach[s, goal] = min(range(1, len(adj)),
                                                     key=lambda a: (env.dist[goal].get(adj[a]["nextViewpointId"], 1e9), a)) if len(adj) > 1 else 0
        finally:
            env.with_features = keep
        t = lambda x: torch.from_numpy(x).to(device).contiguous()
        self.vp, self.view, self.nvalid, self.cv, self.trig, self.next = t(vp), t(view), t(nvalid), t(cv), t(trig), t(nxt)
        self.teach = t(teach) if teach is not None else None
        self.S, self.A, self.G = S, a_cap, nvp
        self.device = device

    def world_state(self, s: int):
        """The WorldState (env.py:227) of state id `s`."""
        return WorldState("fake", int(s) // self.HEADINGS, (int(s) % self.HEADINGS) * (math.pi / 6), 0.0)

    def observe_states(self, env, per_instance_states):
        """Observations (env.py:763-804, without the teacher action) of lists of state ids, one list per instance of
        env.batch: an observation is a function of the state apart from the instruction fields of its instance, so every
        distinct state is observed once and stamped per instance."""
        uniq = sorted({int(s) for states in per_instance_states for s in states})
        tmpl = {s: env._observe_one(self.world_state(s), env.batch[0], include_teacher=False) for s in uniq}
        out = []
        for item, states in zip(env.batch, per_instance_states):
            stamp = {"instr_id": item["instr_id"], "instr_encoding": item["instr_encoding"], "instr_length": len(item["instr_encoding"])}
            out.append([dict(tmpl[int(s)], **stamp) for s in states])
        return out

    def state_ids(self, world_states):
        return [int(ws.viewpointId) * self.HEADINGS + int(round(ws.heading / (math.pi / 6))) % 12 for ws in world_states]
