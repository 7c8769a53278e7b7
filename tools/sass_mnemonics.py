"""Static SASS evidence per kernel of egotap_b200/libegotap_b200.so (cuobjdump -sass): counts of the sm_100a mnemonics that
prove the tcgen05 / TMA / TMEM path.  Usage: python tools/sass_mnemonics.py [substring ...] > profiles/<name>.csv"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "egotap_b200", "libegotap_b200.so")
want = sys.argv[1:]
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cols = ["UTCHMMA", "UTMALDG", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "HMMA", "MUFU.EX2"]
rows, cur, cnt = [], None, None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        if cur:
            rows.append((cur, cnt))
        cur, cnt = m.group(1), collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        cnt["total"] += 1
        for c in cols:
            if op == c or op.startswith(c + "."):
                cnt[c] += 1
if cur:
    rows.append((cur, cnt))
names = subprocess.run(["c++filt"] + [r[0] for r in rows], capture_output=True, text=True).stdout.splitlines()
print("kernel," + ",".join(cols) + ",total_instructions")
for (_, c), d in zip(rows, names):
    short = re.sub(r"\(.*", "", d).replace("void ", "").replace("eb::", "")
    if not want or any(w in short for w in want):
        print('"%s",' % short + ",".join(str(c[k]) for k in cols) + ",%d" % c["total"])
