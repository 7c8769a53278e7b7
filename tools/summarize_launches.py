"""ncu launch list (--csv, long format: one row per launch and metric) -> per-kernel-family table + DRAM-traffic JSON.

    python tools/summarize_launches.py RAW.csv OUT.csv [TRAFFIC.json PRESET BATCH PRECISION] [--title "..."]

The table is what profiles/r0N_launches_step_*.csv hold; the traffic JSON is what bench.py reads for roofline.traffic."""
import collections
import csv
import json
import re
import sys


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("eb::", "").replace("(int)", "").replace("(bool)", "")
    return name


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    title = ""
    if "--title" in sys.argv:
        title = sys.argv[sys.argv.index("--title") + 1]
        args = [a for a in args if a != title]
    raw, out = args[0], args[1]
    rows = [r for r in csv.reader(l for l in open(raw) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    ix = {h: i for i, h in enumerate(hdr)}
    launches = collections.OrderedDict()
    for r in rows:
        d = launches.setdefault(r[ix["ID"]], dict(name=short(r[ix["Kernel Name"]])))
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        m = r[ix["Metric Name"]]
        if m == "gpu__time_duration.sum":
            v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "msecond": 1.0, "ms": 1.0, "nsecond": 1e-6, "second": 1e3}[unit]
        if m.startswith("dram__bytes"):
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        d[m] = v
    fam = collections.OrderedDict()
    for d in launches.values():
        f = fam.setdefault(d["name"], dict(n=0, ms=0.0, rd=0.0, wr=0.0, tp=0.0))
        f["n"] += 1
        f["ms"] += d.get("gpu__time_duration.sum", 0.0)
        f["rd"] += d.get("dram__bytes_read.sum", 0.0)
        f["wr"] += d.get("dram__bytes_write.sum", 0.0)
        f["tp"] += d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0) * d.get("gpu__time_duration.sum", 0.0)
    tot = sum(f["ms"] for f in fam.values())
    with open(out, "w") as fo:
        if title:
            fo.write("# %s\n" % title)
        fo.write("# per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes\n")
        fo.write("kernel,launches,total_ms,share,dram_read_GB,dram_write_GB,tensor_pipe_active_pct_time_weighted\n")
        for k, f in sorted(fam.items(), key=lambda kv: -kv[1]["ms"]):
            fo.write("%s,%d,%.3f,%.3f,%.3f,%.3f,%.1f\n" % (k.replace(",", ";"), f["n"], f["ms"], f["ms"] / tot, f["rd"] / 1e9, f["wr"] / 1e9,
                                                      f["tp"] / f["ms"] if f["ms"] else 0.0))
        fo.write("TOTAL,%d,%.3f,1.0,%.3f,%.3f,\n" % (sum(f["n"] for f in fam.values()), tot, sum(f["rd"] for f in fam.values()) / 1e9,
                                                   sum(f["wr"] for f in fam.values()) / 1e9))
    if len(args) >= 6:
        tj, preset, batch, precision = args[2], args[3], int(args[4]), args[5]
        g = {k: f for k, f in fam.items() if k.startswith("gemm_tc_kernel")}
        json.dump(dict(note="DRAM bytes of ONE forward step from ncu dram__bytes_read.sum + dram__bytes_write.sum, summed per kernel "
                            "family; source %s" % raw, batch=batch, preset=preset, precision=precision,
                       gemm_tc_kernel_bytes_per_step=sum(f["rd"] + f["wr"] for f in g.values()),
                       gemm_tc_kernel_launches=sum(f["n"] for f in g.values()),
                       per_kernel_bytes={k: f["rd"] + f["wr"] for k, f in fam.items()}), open(tj, "w"), indent=1)


if __name__ == "__main__":
    main()
