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
    unit_row = next(rd)
    ix = {h: i for i, h in enumerate(hdr)}
    cols = [("gpu__time_duration.sum", "us", 1.0), ("dram__bytes_read.sum", "DRAM rd MB", 1.0),
            ("dram__bytes_write.sum", "DRAM wr MB", 1.0),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %", 1.0),
            ("lts__t_bytes.sum", "L2 MB", 1.0), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %", 1.0),
            ("lts__t_sector_hit_rate.pct", "L2 hit %", 1.0), ("l1tex__t_sector_hit_rate.pct", "L1 hit %", 1.0),
            ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %", 1.0),
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
    # machine-readable per-launch traffic for bench.py's roofline entries (profiles/ncu_traffic.json)
    if len(sys.argv) > 3:
        import json
        import re

        entry = {"field_mlp_tc_bwd_saved": "umhs_field_mlp_tc_bwd_saved", "field_mlp_tc_bwd_kernel": "umhs_field_mlp_tc_bwd",
                 "field_mlp_tc_fwd_kernel<0": "umhs_field_mlp_tc_fwd", "field_mlp_tc_fwd_kernel<false": "umhs_field_mlp_tc_fwd",
                 "field_mlp_tc_fwd_kernel<1": "umhs_density_tc_fwd", "field_mlp_tc_fwd_kernel<true": "umhs_density_tc_fwd", "hash_encode_fwd": "umhs_hash_encode_fwd",
                 "hash_encode_bwd": "umhs_hash_encode_bwd", "composite_fwd": "umhs_composite_fwd",
                 "composite_bwd": "umhs_composite_bwd", "adam_kernel": "umhs_adam_step_dev", "train_loss": "umhs_train_loss",
                 "march_kernel": "umhs_march_count"}

        def num(r, key):
            try:
                return float(r[ix[key]]) if key in ix else None
            except ValueError:
                return None

        units = {h: u for h, u in zip(hdr, unit_row)}
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        out = {}
        for name, (t, r) in best.items():
            for pat, api in entry.items():
                if pat in name and api not in out:
                    rd, wr, l2 = num(r, "dram__bytes_read.sum"), num(r, "dram__bytes_write.sum"), num(r, "lts__t_bytes.sum")
                    out[api] = {"dram_bytes": (rd or 0) * scale.get(units.get("dram__bytes_read.sum", "byte"), 1.0)
                                + (wr or 0) * scale.get(units.get("dram__bytes_write.sum", "byte"), 1.0),
                                "lts_bytes": None if l2 is None else l2 * scale.get(units.get("lts__t_bytes.sum", "byte"), 1.0),
                                "us_under_ncu": t}
        out["umhs_adam_step"] = out.get("umhs_adam_step_dev")
        json.dump(out, open(sys.argv[3], "w"), indent=1)
    print("| kernel | " + " | ".join(c[1] for c in cols) + " |")
    print("|---|" + "---:|" * len(cols))
    for name, (t, r) in sorted(best.items(), key=lambda kv: -kv[1][0]):
        vals = []
        for key, label, _ in cols:
            v = r[ix[key]] if key in ix else ""
            try:
                f = float(v)
                if label.endswith(" MB") and key in ix:  # ncu picks a unit per column: bring byte counts to MB
                    f *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(unit_row[ix[key]], 1.0)
                vals.append(f"{f:.0f}" if f >= 100 or f == int(f) else f"{f:.2f}")
            except ValueError:
                vals.append(v or "-")
        print(f"| `{name[:60]}` | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
