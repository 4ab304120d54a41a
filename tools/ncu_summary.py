"""Condense `ncu --page raw --csv` exports into one small table (profiles/*.md): duration, DRAM traffic / throughput,
tensor-pipe activity, issue activity, occupancy, top stall reasons per captured launch."""
import csv, sys, os, re

KEYS = [
    ("gpu__time_duration.sum", "dur_us"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct"),
    ("sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active", "hmma_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_pct"),
    ("launch__registers_per_thread", "regs"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
]
STALL = "smsp__average_warps_issue_stalled_"

def norm(val, unit):
    v = float(val.replace(",", "")) if val not in ("", "n/a") else float("nan")
    u = unit.lower()
    scale = {"kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "byte": 1.0, "ns": 1e-3, "us": 1.0, "ms": 1e3, "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3}
    return v * scale.get(u, 1.0)

rows_out = []
for path in sys.argv[1:]:
    with open(path) as f:
        rows = list(csv.reader(f))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "")
        d = {"kernel": name[:44], "grid": r[idx["Grid Size"]].replace(" ", "") if "Grid Size" in idx else ""}
        for k, short in KEYS:
            if k in idx:
                d[short] = norm(r[idx[k]], units[idx[k]])
        stalls = sorted(((float(r[i].replace(",", "")) if r[i] not in ("", "n/a") else 0.0, h[len(STALL):].replace("_per_issue_active.ratio", ""))
                         for h, i in idx.items() if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and "selected" not in h),
                        reverse=True)[:3]
        d["stalls"] = ", ".join(f"{n} {v:.1f}" for v, n in stalls)
        rows_out.append(d)
print("| kernel | grid | dur us | DRAM rd MB | DRAM wr MB | DRAM % | tensor % | issue % | warps % | regs | top stalls (warps per issue) |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for d in rows_out:
    g = lambda k, s=1.0, f="{:.1f}": f.format(d[k] * s) if k in d and d[k] == d[k] else "-"
    print(f"| {d['kernel']} | {d['grid']} | {g('dur_us')} | {g('dram_rd', 1e-6)} | {g('dram_wr', 1e-6)} | {g('dram_pct')} | {g('tensor_pct')} | {g('issue_pct')} | "
          f"{g('occ_pct')} | {g('regs', 1.0, '{:.0f}')} | {d['stalls']} |")
