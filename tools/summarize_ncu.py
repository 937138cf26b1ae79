"""Turn ncu outputs into the markdown summaries committed under profiles/.

  python tools/summarize_ncu.py launches <launch list .csv>          (ncu --metrics gpu__time_duration.sum --csv)
  python tools/summarize_ncu.py full <report .ncu-rep>               (ncu --set full; reads it with `ncu -i --page raw`)

Prints markdown to stdout.  Per-launch times under ncu are cold-cache and serialised: compare SHARES with the live
CUDA-event figures of bench.py, not absolute values."""
import collections
import csv
import io
import subprocess
import sys


def launches(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if r]
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    k, v, u = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= v:
            continue
        try:
            t = float(r[v].replace(",", ""))
        except ValueError:
            continue
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[u], 1e-3)
        d = tot.setdefault(r[k], [0.0, 0])
        d[0] += t * scale
        d[1] += 1
    total = sum(d[0] for d in tot.values())
    n = sum(d[1] for d in tot.values())
    print(f"Total captured device time: {total / 1e3:.2f} ms over {n} launches.\n")
    print("| share | total us | launches | kernel |\n|---:|---:|---:|---|")
    ours = 0.0
    for name, (t, c) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
        if "umhs::" in name:
            ours += t
        if t / total >= 0.0015:
            print(f"| {100 * t / total:.2f}% | {t:.1f} | {c} | `{name[:86]}` |")
    print(f"\nShare of device time in this repo's kernels (`umhs::*`): {100 * ours / total:.1f}%.")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = csv.reader(io.StringIO(out))
    hdr = next(rd)
    next(rd)
    ix = {h: i for i, h in enumerate(hdr)}
    cols = [("gpu__time_duration.sum", "us", 1.0), ("dram__bytes_read.sum", "DRAM rd MB", 1.0),
            ("dram__bytes_write.sum", "DRAM wr MB", 1.0),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %", 1.0),
            ("lts__t_sector_hit_rate.pct", "L2 hit %", 1.0), ("l1tex__t_sector_hit_rate.pct", "L1 hit %", 1.0),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %", 1.0),
            ("smsp__issue_active.avg.per_cycle_active", "IPC / scheduler", 1.0),
            ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes / instr", 1.0),
            ("launch__registers_per_thread", "regs", 1.0), ("launch__grid_size", "grid", 1.0),
            ("launch__block_size", "block", 1.0)]
    best = {}
    for r in rd:
        name = r[ix["Kernel Name"]]
        t = float(r[ix["gpu__time_duration.sum"]])
        if name not in best or t > best[name][0]:
            best[name] = (t, r)
    print("| kernel | " + " | ".join(c[1] for c in cols) + " |")
    print("|---|" + "---:|" * len(cols))
    for name, (t, r) in sorted(best.items(), key=lambda kv: -kv[1][0]):
        vals = []
        for key, _, _ in cols:
            v = r[ix[key]] if key in ix else ""
            try:
                f = float(v)
                vals.append(f"{f:.0f}" if f >= 100 or f == int(f) else f"{f:.2f}")
            except ValueError:
                vals.append(v or "-")
        print(f"| `{name[:60]}` | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
