"""Opcode census of the built objects: python tools/sass_census.py > profiles/r2_sass_census.md
Counts selected SASS opcode families per object of lair_b200/build (cuobjdump -sass)."""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "lair_b200", "build")
FAM = ["UTCHMMA", "LDTM", "UTMALDG", "UTCBAR", "UTCATOMSWS", "DMMA", "SYNCS", "LDGSTS", "UCGABAR", "STAS", "CREDUX", "VOTE", "SHFL", "FFMA2", "FADD2", "DFMA", "LDS", "STS"]
print("# SASS opcode census of liblair_b200.so's objects (round 2 final tree, nvcc 12.9 -gencode arch=compute_100a,code=sm_100a)\n")
print("Counts of selected opcode families per object (`cuobjdump -sass lair_b200/build/<obj>.o`, written by `tools/sass_census.py`): tensor / TMEM / TMA / "
      "async-copy / barrier / remote-store / reduction / shuffle / shared-memory instructions.\n")
print("| object | total | " + " | ".join(FAM) + " |")
print("|---|---|" + "---|" * len(FAM))
for f in sorted(os.listdir(BUILD)):
    if not f.endswith(".o"):
        continue
    out = subprocess.run(["cuobjdump", "-sass", os.path.join(BUILD, f)], capture_output=True, text=True).stdout
    ops = re.findall(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", out, flags=re.M)
    cnt = {k: 0 for k in FAM}
    for o in ops:
        for k in FAM:
            if o.startswith(k):
                cnt[k] += 1
                break
    print(f"| {f} | {len(ops)} | " + " | ".join(str(cnt[k]) for k in FAM) + " |")
