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


if __name__ == "__main__":
    print("oracle/_ref ready:", make(force=True))
