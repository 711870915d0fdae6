"""Turn ncu outputs (read here, no GPU needed) into the committed summaries under profiles/.

  python scripts/ncu_summarize.py launches <launches.csv> <n_launches_per_step> <out.md>
  python scripts/ncu_summarize.py traffic  <metrics.csv> <out.json>
  python scripts/ncu_summarize.py rep      <file.ncu-rep> <out.md>
"""
import collections
import csv
import json
import re
import subprocess
import sys


def read_csv(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    return list(csv.DictReader(lines))


def launches(path, per_step, out):
    rows = [r for r in read_csv(path) if r.get("Metric Name") == "gpu__time_duration.sum"]
    step = rows[-per_step:]
    tot = sum(float(r["Metric Value"]) for r in step)
    agg = collections.OrderedDict()
    for r in step:
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").strip()
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"])
    with open(out, "w") as f:
        f.write("# ncu launch list — one step (last %d launches of %s)\n\n" % (per_step, path))
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` (serialised, cold cache: compare SHARES).\n\n")
        f.write("step total: %.3f ms over %d launches\n\n| kernel | launches | time (us) | share |\n|---|---:|---:|---:|\n" % (tot / 1e6, len(step)))
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.1f | %.1f %% |\n" % (k, n, t / 1e3, 100 * t / tot))
        conv = sum(t for k, (n, t) in agg.items() if "conv_gemm" in k or "conv3x3" in k)
        f.write("\ntensor-core kernel share of the step: **%.1f %%**\n" % (100 * conv / tot))
    print(open(out).read())


def traffic(path, out):
    rows = read_csv(path)
    per = collections.defaultdict(dict)
    for r in rows:
        per[r["ID"]][r["Metric Name"]] = (float(r["Metric Value"]), r["Metric Unit"])
    unit = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot_b, tot_t, n = 0.0, 0.0, 0
    for k, m in per.items():
        if "dram__bytes_read.sum" not in m:
            continue
        b = sum(m[x][0] * unit[m[x][1]] for x in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        tot_b += b
        tot_t += m.get("gpu__time_duration.sum", (0.0, "ns"))[0]
        n += 1
    res = {"kernel": "conv_gemm_kernel (all tensor-core launches of one step)", "launches": n,
           "dram_bytes_total": tot_b, "dram_bytes_per_launch": tot_b / max(n, 1),
           "sum_duration_ms": tot_t / 1e6, "source": path}
    json.dump(res, open(out, "w"), indent=1)
    print(res)


def rep(path, out, index=None):
    """index: which kernel of a multi-kernel report (default: the last one)"""
    sel = [] if index is None else ["--launch-skip", str(index), "--launch-count", "1"]
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"] + sel, capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "sm__inst_executed_pipe_uniform.sum"]
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"] + sel, capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    shdr, sdata = srows[1], srows[2:]
    ix = {h: i for i, h in enumerate(shdr)}
    # a report with imported sources repeats the header per view: keep the SASS rows of the first view only
    keep = []
    for r in sdata:
        if len(r) <= ix["# Samples"] or not r[ix["# Samples"]].strip().isdigit():
            if keep:
                break
            continue
        keep.append(r)
    sdata = keep
    stalls = [h for h in shdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {s: sum(int(r[ix[s]] or 0) for r in sdata if len(r) > ix[s]) for s in stalls}
    mn = {"UTCHMMA": 0, "UTMALDG": 0, "UTMASTG": 0, "LDTM": 0, "SYNCS": 0, "ELECT": 0, "BRA.U.ANY": 0}
    for r in sdata:
        for k in mn:
            if k in r[ix["Source"]]:
                mn[k] += 1
    with open(out, "w") as f:
        f.write("# ncu --set full: %s\n\n| metric | value |\n|---|---|\n" % path)
        for i, h in enumerate(hdr):
            if any(h == w or h.startswith(w + ".") and h == w for w in want) or h in want:
                f.write("| %s | %s %s |\n" % (h, vals[i], units[i]))
        f.write("\nSASS evidence (static instruction counts): %s\n" % mn)
        f.write("\nWarp-stall samples by reason: %s\n\nTop stall sites:\n\n" %
                sorted(agg.items(), key=lambda kv: -kv[1])[:8])
        top = sorted(sdata, key=lambda r: -int(r[ix["# Samples"]] or 0))[:14]
        for r in top:
            st = sorted([(int(r[ix[s]] or 0), s[6:]) for s in stalls], reverse=True)[:2]
            f.write("* %s samples — `%s` %s\n" % (r[ix["# Samples"]], r[ix["Source"]].strip()[:90], st))
    print(open(out).read())


if __name__ == "__main__":
    {"launches": lambda: launches(sys.argv[2], int(sys.argv[3]), sys.argv[4]),
     "traffic": lambda: traffic(sys.argv[2], sys.argv[3]),
     "rep": lambda: rep(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else None)}[sys.argv[1]]()
