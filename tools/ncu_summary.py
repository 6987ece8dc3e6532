"""Summarise an .ncu-rep (ncu --set full) as text: one line per profiled launch with duration, DRAM bytes, L2 sectors,
tensor-pipe activity and issue rate.   python tools/ncu_summary.py report.ncu-rep [--json out.json --kind gemm|gather]"""
import csv
import json
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
COLS = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "dram_rd"),
        ("dram__bytes_write.sum", "dram_wr"), ("lts__t_sectors.sum", "l2_sectors"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts_pct"),
        ("TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor_pipe_pct"),
        ("sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed", "bf16_ops_pct"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
        ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs")]
idx = [(hdr.index(c), n) for c, n in COLS if c in hdr]
units = {n: rows[1][i] for i, n in idx}


def num(v):
    try:
        return float(str(v).replace(",", ""))
    except ValueError:
        return 0.0   # "no data"


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


recs = []
for r in rows[2:]:
    d = {n: r[i] for i, n in idx}
    for k in ("dram_rd", "dram_wr"):
        if k in d:
            d[k] = to_bytes(d[k], units[k])
    recs.append(d)
print(f"# {rep}: ncu --set full --clock-control none (cold caches, serialised launches: durations are NOT the in-step times)")
print(f"{'kernel':42s} {'us':>8s} {'DRAM rd MB':>10s} {'DRAM wr MB':>10s} {'L2 MB':>9s} {'L2 %':>6s} {'tensor %':>8s} {'bf16 ops %':>10s} {'issue %':>7s} {'warps %':>7s}")
for d in recs:
    l2 = num(d.get("l2_sectors", "0")) * 32 / 1e6
    print(f"{d['kernel'][:42]:42s} {num(d['us']):8.2f} {d.get('dram_rd', 0) / 1e6:10.2f} {d.get('dram_wr', 0) / 1e6:10.2f} {l2:9.1f} "
          f"{num(d.get('lts_pct', 0)):6.1f} {num(d.get('tensor_pipe_pct', 0)):8.1f} {num(d.get('bf16_ops_pct', 0)):10.2f} "
          f"{num(d.get('issue_pct', 0)):7.1f} {num(d.get('occupancy_pct', 0)):7.1f}")
if "--json" in sys.argv:
    path = sys.argv[sys.argv.index("--json") + 1]
    kind = sys.argv[sys.argv.index("--kind") + 1]
    per = [d.get("dram_rd", 0) + d.get("dram_wr", 0) for d in recs]
    if kind == "gemm":
        blob = {"kernel": recs[0]["kernel"].split("(")[0], "source": rep, "dram_bytes_per_launch": per,
                "dram_bytes_per_step": sum(per), "launch_us_under_ncu": [float(d["us"]) for d in recs]}
    else:
        blob = {"kernel": recs[0]["kernel"].split("(")[0], "source": rep, "dram_bytes_per_launch": per[0],
                "dram_bytes_read": recs[0].get("dram_rd", 0), "dram_bytes_write": recs[0].get("dram_wr", 0),
                "launch_us_under_ncu": float(recs[0]["us"])}
    json.dump(blob, open(path, "w"), indent=1)
