"""Print the per-kernel table of a bench.py --dump file."""
import collections
import json
import sys

d = json.load(open(sys.argv[1]))
agg = collections.OrderedDict()
for r in d["kernel_launches"]:
    k = (r["name"],) + ((r["M"], r["N"], r["K"], r["groups"], r["variant"]) if "flops" in r else ())
    a = agg.setdefault(k, [0, 0.0, 0.0])
    a[0] += 1; a[1] += r["ms"]; a[2] += r.get("flops", 0.0)
tot = sum(a[1] for a in agg.values())
L = d["line"]
print("step %.3f ms  (%s)  sum of kernels %.3f ms  value %.0f %s  e2e %.0f  clocks %s" % (
    L["ms_per_step"], L["config"]["precision"], tot, L["value"], L["unit"], L["e2e"]["value"], L["clocks"]))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    tf = " %6.0f TF/s(alg)" % (a[2] / a[1] / 1e9) if a[2] else ""
    print("%-60s n=%3d %8.3f ms %5.1f%%%s" % (" ".join(str(x) for x in k), a[0], a[1], 100 * a[1] / tot, tf))
