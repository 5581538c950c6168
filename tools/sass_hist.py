"""Per-kernel SASS opcode histogram of the shipped library (cuobjdump -sass; no GPU needed).

  python tools/sass_hist.py [profiles/sass_rNN.txt]

What to read in it: UBLKCP + SYNCS (TMA bulk copy + mbarrier) in the table-staging kernels; ATOMS.POPC.INC / REDUX and
UBLKRED (TMA bulk reduction of the histogram, cp.reduce.async.bulk) in the enumeration, whose multimem.red / multimem.st show
up as REDG / STG .STRONG.SYS (the multicast is in the address mapping); UTCHMMA / LDTM / STTM / UTMALDG (tcgen05 + TMEM + TMA
tensor loads) in k_attention_tc only: the hot path itself is integer / shared-memory lookup work that does not use tensor
cores (north_star)."""
import collections
import importlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
lib = importlib.import_module("texasholdempocker-rl_b200.build").LIB
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
out, name, hist = [], None, None


def flush():
    if name:
        n = sum(hist.values())
        top = ", ".join(f"{k} {v}" for k, v in hist.most_common(18))
        flags = [k for k in hist if k.startswith(("UBLKCP", "UBLKRED", "SYNCS", "STTM", "ATOMS", "REDUX", "HMMA", "UTC", "LDTM", "UTMA", "RED", "ATOMG", "LDGSTS"))]
        out.append(f"{name}\n  {n} instructions; notable: {', '.join(sorted(flags)) or '-'}\n  {top}\n")


for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        flush()
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        hist = collections.Counter()
        continue
    m = re.match(r"^\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and hist is not None:
        op = m.group(2)
        hist[".".join(op.split(".")[:3]) if op.startswith(("ATOMS", "SYNCS", "HMMA", "RED", "ATOMG")) else op.split(".")[0]] += 1
flush()
text = "SASS opcode histogram per kernel, " + os.path.relpath(lib, ROOT) + " (sm_100a)\n\n" + "\n".join(out)
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(text)
print(text[:3000])
