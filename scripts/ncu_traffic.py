"""Summarise `ncu --set full` captures (exported with `ncu -i X.ncu-rep --page raw --csv`) into
profiles/ncu_traffic.json (the `roofline.traffic` source of bench.py) and a markdown table.

    python scripts/ncu_traffic.py KEY=raw.csv [KEY=raw.csv ...] [--md profiles/r02_ncu_full.md]

KEY is the bench.py lookup key of the kernel and shape, e.g. 'gemm_nt_syrk:n=16384,m=262144'."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed_pipe_fp64_op_dmma.sum"]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3,
        "second": 1.0, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}


def read_raw(path):
    rows = list(csv.reader(open(path, newline="")))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units = rows[hdr], rows[hdr + 1]
    out = []
    for r in rows[hdr + 2:]:
        if len(r) != len(names):
            continue
        d = {"Kernel Name": r[names.index("Kernel Name")]}
        for i, (nm, u) in enumerate(zip(names, units)):
            try:
                d[nm] = float(r[i].replace(",", "")) * UNIT.get(u, 1.0)
            except ValueError:
                pass
        out.append(d)
    return out


def main():
    args = [a for a in sys.argv[1:] if "=" in a]
    md = sys.argv[sys.argv.index("--md") + 1] if "--md" in sys.argv else None
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    tab = json.load(open(path)) if os.path.exists(path) else {}
    lines = ["| key | kernel | time | DRAM read | DRAM write | FP64 pipe % of peak (active) | L2 hit % | shared bank conflicts | regs |",
             "|---|---|---|---|---|---|---|---|---|"]
    for a in args:
        key, f = a.rsplit("=", 1)
        for d in read_raw(f):
            rd, wr = d.get("dram__bytes_read.sum", 0.0), d.get("dram__bytes_write.sum", 0.0)
            tab[key] = {"dram_bytes": rd + wr, "source": f"profiles/{os.path.basename(f)} (ncu --set full: dram__bytes_read.sum "
                        f"{rd / 1e9:.3f} GB + dram__bytes_write.sum {wr / 1e9:.3f} GB, gpu__time_duration {d.get('gpu__time_duration.sum', 0) * 1e3:.3f} ms)"}
            fp = d.get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", d.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", float("nan")))
            lines.append(f"| {key} | `{d['Kernel Name'][:48]}` | {d.get('gpu__time_duration.sum', 0) * 1e3:.3f} ms | {rd / 1e9:.3f} GB | {wr / 1e9:.3f} GB | "
                         f"{fp:.1f} | {d.get('lts__t_sector_hit_rate.pct', float('nan')):.1f} | "
                         f"{d.get('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', float('nan')):.0f} | {d.get('launch__registers_per_thread', float('nan')):.0f} |")
            break                                   # first launch of the capture
    json.dump(tab, open(path, "w"), indent=1)
    if md:
        open(md, "a").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
