"""Per-launch table from the raw page of an `ncu --set full` capture of a whole step (scripts/gpu_evidence.sh):
    python scripts/ncu_step_table.py <raw.csv> <out.md> [title]
One row per launch: kernel (template arguments kept), grid, duration, tensor-pipe activity, DRAM throughput (% of peak and
achieved GB/s from dram bytes / duration), L2 throughput, registers; then the totals per kernel instantiation."""
import collections
import csv
import re
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6,
        "nsecond": 1e-9, "msecond": 1e-3, "second": 1.0}


def main():
    path, out = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else path
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def col(r, name, scale=False):
        if name not in ix or ix[name] >= len(r) or r[ix[name]] in ("", "n/a"):
            return float("nan")
        v = float(r[ix[name]].replace(",", ""))
        return v * UNIT.get(units[ix[name]], 1.0) if scale else v

    lines, agg = [], collections.OrderedDict()
    tot_t = 0.0
    for n, r in enumerate(data):
        if len(r) < len(hdr) // 2:
            continue
        name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "").replace("p2l::", "").replace("(int)", "").replace("(bool)", "")
        t = col(r, "gpu__time_duration.sum", True)
        rd, wr = col(r, "dram__bytes_read.sum", True), col(r, "dram__bytes_write.sum", True)
        tp = col(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
        if tp != tp:
            tp = col(r, "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active")
        dr = col(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
        l2 = col(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed")
        regs = col(r, "launch__registers_per_thread")
        grid = r[ix["Grid Size"]] if "Grid Size" in ix else str(col(r, "launch__grid_size"))
        lines.append("| %d | `%s` | %s | %.1f | %.1f | %.1f | %.0f | %.1f | %.0f | %.0f |" %
                     (n, name, grid, t * 1e6, tp, dr, (rd + wr) / t / 1e9, (rd + wr) / 1e6, l2, regs))
        a = agg.setdefault(name, [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += t
        a[2] += rd + wr
        a[3] += tp * t
        tot_t += t
    with open(out, "w") as f:
        f.write("# %s\n\n`ncu --set full --clock-control none` over every matching launch of one step (serialised, caches flushed "
                "between launches: durations are upper bounds, compare shares and the utilisation columns).\n\n" % title)
        f.write("## Per kernel instantiation\n\n| kernel | launches | time (us) | share | tensor pipe active (time-weighted %) | DRAM GB/s (avg) | DRAM MB / launch |\n|---|---:|---:|---:|---:|---:|---:|\n")
        for k, (n, t, b, tpw) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.1f | %.1f %% | %.1f | %.0f | %.1f |\n" % (k, n, t * 1e6, 100 * t / tot_t, tpw / t, b / t / 1e9, b / n / 1e6))
        f.write("| total | %d | %.1f | | | | |\n" % (sum(a[0] for a in agg.values()), tot_t * 1e6))
        f.write("\n## Per launch\n\n| # | kernel | grid | us | tensor pipe % | DRAM % of peak | DRAM GB/s | DRAM MB | L2 % | regs |\n|---:|---|---|---:|---:|---:|---:|---:|---:|---:|\n")
        f.write("\n".join(lines) + "\n")
    print(open(out).read()[:3000])


main()
