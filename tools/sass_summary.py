"""Per-kernel SASS opcode census of tatt_b200/lib/libtatt_b200.so (cuobjdump -sass): the mnemonics that prove which
hardware path a kernel uses -- UTCHMMA (tcgen05.mma), LDTM/STTM (tcgen05.ld/st), UTCBAR (tcgen05.commit), UTMALDG /
UTMASTG (TMA bulk tensor load / store), SYNCS (mbarrier), HMMA (mma.sync), FFMA (CUDA-core fp32), LDGSTS (cp.async),
RED/ATOM (atomics).  Usage: python tools/sass_summary.py > profiles/<round>_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tatt_b200", "lib", "libtatt_b200.so")
OPS = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UTMASTG", "SYNCS", "HMMA", "FFMA", "LDGSTS", "RED", "ATOM", "MUFU"]

out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
counts = collections.OrderedDict()
name = None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = m.group(1)
        counts[name] = collections.Counter()
        continue
    if name is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m:
        op = m.group(1)
        for key in OPS:
            if op.startswith(key):
                counts[name][key] += 1
        counts[name]["_total"] += 1


def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n
    except Exception:
        return n


print("# SASS opcode census of %s -- architectures in the binary: %s" % (os.path.relpath(LIB, ROOT), ", ".join(arch)))
print("# %-78s %7s " % ("kernel", "instrs") + " ".join("%7s" % k for k in OPS))
tot = collections.Counter()
for n, c in sorted(counts.items(), key=lambda kv: -kv[1]["UTCHMMA"] * 100000 - kv[1]["_total"]):
    d = re.sub(r"\(.*", "", demangle(n).replace("(anonymous namespace)::", "").replace("void ", ""))
    print("%-80s %7d " % (d[:80], c["_total"]) + " ".join("%7d" % c[k] for k in OPS))
    tot.update(c)
print("%-80s %7d " % ("TOTAL (%d kernels)" % len(counts), tot["_total"]) + " ".join("%7d" % tot[k] for k in OPS))
