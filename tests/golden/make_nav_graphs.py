"""Write tests/golden/nav_graphs.npz from the reference's connectivity files (build container only).

    python tests/golden/make_nav_graphs.py

For a handful of small Matterport scans it keeps what ``utils.load_nav_graphs`` (tasks/R2R/utils.py:26-52) keeps:
the included viewpoints, their positions (pose[3], pose[7], pose[11]) and the unobstructed & included adjacency.
``tests/fake_env.py`` turns these into navigation graphs with headings derived from the positions, so that the
agent / search tests walk REAL R2R graphs instead of random ones.  /root/reference is not read at test time.
"""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SCANS = ["gZ6f7yhEvPG", "YmJkqBEsHnH", "GdvgFV5R1Z5", "8194nk5LbLH", "pLe4wQe7qrG"]

out = {}
for scan in SCANS:
    data = json.load(open("/root/reference/connectivity/%s_connectivity.json" % scan))
    keep = [i for i, it in enumerate(data) if it["included"]]
    ids = [data[i]["image_id"] for i in keep]
    pos = np.array([[data[i]["pose"][3], data[i]["pose"][7], data[i]["pose"][11]] for i in keep], np.float64)
    adj = np.zeros((len(keep), len(keep)), np.bool_)
    for a, i in enumerate(keep):
        for b, j in enumerate(keep):
            if data[i]["unobstructed"][j]:
                assert data[j]["unobstructed"][i], "graph should be undirected (utils.py:47)"
                adj[a, b] = True
    out[scan + "/ids"] = np.array(ids)
    out[scan + "/pos"] = pos
    out[scan + "/adj"] = adj
    print(scan, len(keep), "viewpoints", int(adj.sum()) // 2, "edges")
np.savez_compressed(os.path.join(HERE, "nav_graphs.npz"), **out)
