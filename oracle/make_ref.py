"""Recipe for ``oracle/_ref/``: the reference's own Python sources for the hot path, staged UNMODIFIED so that
``bench.py --impl reference`` (and its ``gpu_baseline`` leg) can time the reference itself on the GPU box.

TEST / BENCH INFRASTRUCTURE ONLY.  ``oracle/_ref/`` is build output: git-ignored (never in history), not
gpurun-ignored (it travels to the GPU box like the built ``.so``).  It is produced in the build container, where
``/root/reference`` is mounted read-only, by ``__graft_entry__.build()`` calling :func:`make`; on the GPU box the
staged files are used as they are.  Nothing here edits a reference file: the modules are byte-for-byte copies of

    tasks/R2R/model.py  tasks/R2R/env.py  tasks/R2R/utils.py  tasks/R2R/paths.py

(``model.py`` imports ``utils`` and ``env``; ``env`` imports ``paths`` and the compiled ``MatterSim`` simulator, which is
off the timed path and cannot be built here — SURVEY.md §8c — so :mod:`oracle.ref_loader` registers an empty stub
module under that name before importing).  A MANIFEST with the sha256 of every staged file is written beside them.
"""
import hashlib
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/tasks/R2R"
REF_DST = os.path.join(HERE, "_ref")
FILES = ("model.py", "env.py", "utils.py", "paths.py")


def make(force: bool = False) -> bool:
    """Stage the reference sources; returns True when oracle/_ref is usable afterwards."""
    have = all(os.path.exists(os.path.join(REF_DST, f)) for f in FILES)
    if not os.path.isdir(REF_SRC):
        return have                       # GPU box: use what travelled with the snapshot
    if have and not force:
        same = all(open(os.path.join(REF_SRC, f), "rb").read() == open(os.path.join(REF_DST, f), "rb").read() for f in FILES)
        if same:
            return True
    os.makedirs(REF_DST, exist_ok=True)
    lines = []
    for f in FILES:
        shutil.copyfile(os.path.join(REF_SRC, f), os.path.join(REF_DST, f))
        lines.append("%s  %s" % (hashlib.sha256(open(os.path.join(REF_DST, f), "rb").read()).hexdigest(), f))
    with open(os.path.join(REF_DST, "MANIFEST"), "w") as fh:
        fh.write("# byte-for-byte copies from %s (ronghanghu/speaker_follower), staged by oracle/make_ref.py\n" % REF_SRC)
        fh.write("\n".join(lines) + "\n")
    return True


# ---------------------------------------------------------------------------------------------------------------
# Drop-in acceptance tree (tests/test_dropin_gpu.py): the reference's command-line entry points, byte for byte, plus
# the small fixtures they read, laid out as the reference expects them relative to the working directory
# (SURVEY.md A.4).  Data files are SUBSETS of the reference's own R2R_sub_*.json (first N paths) under the names
# train.py asks for; only the connectivity graphs of the scans those paths visit are staged.
DROPIN_DST = os.path.join(REF_DST, "dropin")
DROPIN_SOURCES = ("train.py", "train_speaker.py", "eval.py", "eval_speaker.py", "vocab.py", "env.py", "utils.py", "paths.py", "bleu.py")
DROPIN_SPLITS = (("sub_train", "train", 40), ("sub_val_seen", "val_seen", 36), ("sub_val_unseen", "val_unseen", 36))


def make_dropin(force: bool = False) -> bool:
    """Stage the acceptance tree; returns True when oracle/_ref/dropin is usable afterwards."""
    import json
    marker = os.path.join(DROPIN_DST, "MANIFEST")
    if not os.path.isdir(REF_SRC):
        return os.path.exists(marker)
    if os.path.exists(marker) and not force:
        return True
    ref_root = os.path.dirname(os.path.dirname(REF_SRC))
    r2r = os.path.join(DROPIN_DST, "tasks", "R2R")
    os.makedirs(os.path.join(r2r, "data"), exist_ok=True)
    os.makedirs(os.path.join(DROPIN_DST, "connectivity"), exist_ok=True)
    lines = []
    for f in DROPIN_SOURCES:
        src = os.path.join(REF_SRC, f)
        if not os.path.exists(src):
            continue
        shutil.copyfile(src, os.path.join(r2r, f))
        lines.append("%s  tasks/R2R/%s" % (hashlib.sha256(open(src, "rb").read()).hexdigest(), f))
    for f in ("train_vocab.txt", "train_glove.npy", "trainval_vocab.txt", "sub_train_vocab.txt"):
        shutil.copyfile(os.path.join(REF_SRC, "data", f), os.path.join(r2r, "data", f))
    scans = set()
    for src_split, dst_split, n in DROPIN_SPLITS:
        with open(os.path.join(REF_SRC, "data", "R2R_%s.json" % src_split)) as fh:
            items = json.load(fh)[:n]
        scans.update(it["scan"] for it in items)
        with open(os.path.join(r2r, "data", "R2R_%s.json" % dst_split), "w") as fh:
            json.dump(items, fh)
        lines.append("subset(first %d paths of R2R_%s.json)  tasks/R2R/data/R2R_%s.json" % (n, src_split, dst_split))
    for scan in sorted(scans):
        shutil.copyfile(os.path.join(ref_root, "connectivity", "%s_connectivity.json" % scan),
                        os.path.join(DROPIN_DST, "connectivity", "%s_connectivity.json" % scan))
    with open(marker, "w") as fh:
        fh.write("# staged by oracle/make_ref.py:make_dropin from %s; sources are byte-for-byte copies\n" % ref_root)
        fh.write("\n".join(lines) + "\n")
    return True


if __name__ == "__main__":
    print("oracle/_ref ready:", make(force=True))
    print("oracle/_ref/dropin ready:", make_dropin(force=True))
