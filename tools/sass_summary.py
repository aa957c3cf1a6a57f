"""Per-object SASS evidence (B200_PROFILING.md "What proves a Blackwell-native kernel"): counts of the tcgen05 / TMEM / TMA
/ packed-FMA mnemonics in every object of idrlnet_b200/csrc/build.   python tools/sass_summary.py > profiles/r02/sass_summary.txt"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "UTMASTG", "UTCATOMSWS", "SYNCS", "FFMA2", "FFMA", "HMMA",
        "MUFU.TANH", "MUFU.EX2", "MUFU.RCP", "LDS.128", "STS.128", "LDG", "STG", "RED", "ATOM"]
print("SASS mnemonic counts per object (cuobjdump -sass, sm_100a).  UTCHMMA = tcgen05.mma kind::tf32/f16, LDTM/STTM = tcgen05.ld/st,")
print("UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (TMA engine, 1-D), FFMA2 = fma.rn.f32x2, HMMA = mma.sync.m16n8k8 TF32 (the per-warp dW contraction of the packed narrow program, 3xTF32).\n")
print(f"{'object':28s}" + "".join(f"{k:>10s}" for k in KEYS))
for obj in sorted(glob.glob(os.path.join(ROOT, "idrlnet_b200", "csrc", "build", "*.o"))):
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    c = collections.Counter()
    for line in sass.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        for k in KEYS:
            if op == k or op.startswith(k + ".") or (k in ("FFMA",) and op.split(".")[0] == k):
                c[k] += 1
    print(f"{os.path.basename(obj):28s}" + "".join(f"{c[k]:10d}" for k in KEYS))
