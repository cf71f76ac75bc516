#!/usr/bin/env python
"""SASS opcode census of libholo_b200.so (no GPU needed): per kernel, the counts of the mnemonics that prove a
Blackwell-native kernel (B200_PROFILING.md): UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG/UBLKCP =
TMA, HMMA = legacy mma.sync.  Writes profiles/<tag>_sass_census.json and prints a table."""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "holo_diffusion_b200", "libholo_b200.so")
PAT = {"UTCHMMA": r"\bUTCHMMA", "UTC*MMA(other)": r"\bUTC(?!HMMA)[A-Z]*MMA", "LDTM": r"\bLDTM", "STTM": r"\bSTTM",
       "UTMALDG": r"\bUTMALDG", "UTMASTG": r"\bUTMASTG", "UBLKCP": r"\bUBLKCP", "UTCBAR": r"\bUTCBAR", "HMMA": r"\bHMMA",
       "SYNCS": r"\bSYNCS", "ACQBULK/griddep": r"\bACQBULK|\bPMTRIG", "LDGSTS": r"\bLDGSTS"}


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    census, cur, n_ins = collections.OrderedDict(), None, collections.Counter()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", cur.replace("(anonymous namespace)::", "").replace("void ", "", 1))
            census[cur] = collections.Counter()
            continue
        if cur and re.search(r"/\*[0-9a-f]{4,}\*/", line):
            n_ins[cur] += 1
            for k, p in PAT.items():
                if re.search(p, line):
                    census[cur][k] += 1
    rows = {k: dict(v, instructions=n_ins[k]) for k, v in census.items()}
    path = os.path.join(ROOT, "profiles", f"{tag}_sass_census.json")
    json.dump({"library": "holo_diffusion_b200/libholo_b200.so", "how": "cuobjdump -sass | tools/sass_census.py", "kernels": rows},
              open(path, "w"), indent=1)
    keys = list(PAT)
    print(f"{'kernel':58s} " + " ".join(f"{k[:8]:>8s}" for k in keys) + "    instr")
    for k, v in rows.items():
        if any(v.get(x) for x in keys[:7]):
            print(f"{k[:58]:58s} " + " ".join(f"{v.get(x, 0):8d}" for x in keys) + f" {v['instructions']:8d}")
    print("written", path)


if __name__ == "__main__":
    main()
