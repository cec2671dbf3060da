"""SASS evidence for the Blackwell-native claim (B200_PROFILING.md: the PTX names never appear in SASS): per kernel of
libwunet_b200.so the counts of UTC*MMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG / UBLKCP (TMA),
HMMA (legacy mma.sync, expected 0), plus registers. Runs on the build box (cuobjdump needs no GPU).

    python tools/sass_digest.py > profiles/rNN_sass_digest.txt
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from wave_u_net_for_speech_enhancement_b200 import _lib  # noqa: E402

so = _lib.SO_PATH
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
res = subprocess.run(["cuobjdump", "-res-usage", so], capture_output=True, text=True).stdout
regs = dict(re.findall(r"Function (\S+):\s*\n\s*REG:(\d+)", res))
pats = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "SYNCS", "HMMA", "SHFL", "FFMA"]
print("# %s" % _lib.load().wunet_version().decode())
print("# cuobjdump -sass: instruction counts per kernel (static). UTCHMMA = tcgen05.mma kind::f16, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor")
print("# load / store, UBLKCP = bulk copy, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops; HMMA (legacy mma.sync) must be 0.")
print("%-72s %5s " % ("kernel", "regs") + " ".join("%7s" % p for p in pats))
for m in re.finditer(r"Function : (\S+)\n(.*?)(?=\n\s*Function : |\Z)", sass, re.S):
    name, body = m.group(1), m.group(2)
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    short = re.sub(r"\(.*", "", dem.replace("(anonymous namespace)::", "").replace("wunet::", ""))
    counts = []
    for p in pats:
        if p == "HMMA":
            counts.append(len(re.findall(r"\bHMMA\b", body)))
        else:
            counts.append(len(re.findall(r"\b%s" % p, body)))
    print("%-72s %5s " % (short[:72], regs.get(name, "?")) + " ".join("%7d" % c for c in counts))
