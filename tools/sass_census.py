"""SASS mnemonic census of libsf_b200.so per kernel -> profiles/<tag>_sass_census.txt (runs here, no GPU needed).
Usage: python tools/sass_census.py > profiles/r02_sass_census.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "speaker_follower_b200", "libsf_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = {}
cur = None
counts = collections.OrderedDict()
COLS = ["UTCHMMA", "LDTM", "UTCBAR", "UBLKCP", "HMMA", "SYNCS", "UCGABAR", "LDGSTS", "SHFL", "BAR", "FFMA", "MUFU", "ATOMG", "RED", "LDS", "STS", "LDG", "STG"]
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        op = m.group(1)
        counts[cur]["total"] += 1
        for c in COLS:
            if op == c or op.startswith(c + ".") or op.startswith(c + "_") or (c in ("UTCHMMA", "UBLKCP", "UTCBAR", "LDTM", "UCGABAR", "SYNCS", "LDGSTS") and op.startswith(c)):
                counts[cur][c] += 1
dem = subprocess.run(["cu++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines()
print("# SASS census of speaker_follower_b200/libsf_b200.so (cuobjdump -sass, sm_100a), instruction counts per kernel.")
print("# UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (TMA engine), SYNCS = mbarrier ops,")
print("# UCGABAR = barrier.cluster, HMMA = legacy mma.sync (only in the stateless in-place path), LDGSTS = cp.async.")
print("%-66s %7s " % ("kernel", "total") + " ".join("%7s" % c for c in COLS))
for (k, c), d in zip(counts.items(), dem):
    d = re.sub(r"^(void )?sfb::", "", d.split("(")[0].replace("void sfb::", ""))
    print("%-66s %7d " % (d[:66], c["total"]) + " ".join("%7d" % c[x] for x in COLS))
