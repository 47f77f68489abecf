"""Summarise `ncu --set full` reports (gpurun_out/*.ncu-rep) into one markdown table + JSON per report.

    python scripts/ncu_summarize.py gpurun_out/r01b_ncu_attn_decode_bf16.ncu-rep [...] > profiles/r01b_ncu_summary.md

Reads the raw page through `ncu -i <rep> --page raw --csv` (works without a GPU) and keeps the metrics the roofline
arguments in DESIGN.md use: duration, DRAM bytes read/written (-> "traffic"), DRAM throughput %, tensor-pipe activity %,
SM throughput %, achieved occupancy, registers, shared memory, grid/block."""
import csv
import io
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "dur_us",
    "dram__bytes_read.sum": "dram_rd_MB",
    "dram__bytes_write.sum": "dram_wr_MB",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct2",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct",
    "sm__inst_executed_pipe_tensor.sum": "tensor_inst",
    "lts__t_bytes.sum": "l2_MB",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occ_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__shared_mem_per_block_dynamic": "dsmem_B",
    "launch__occupancy_limit_registers": "occ_lim_regs",
    "launch__occupancy_limit_shared_mem": "occ_lim_smem",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio": "stall_long_sb",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_sb",
}
SCALE = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "ns": 1e-3,
         "us": 1.0, "ms": 1e3, "second": 1e6}


def rows_of(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    lines = [l for l in out.splitlines() if l.startswith('"')]
    rd = list(csv.reader(io.StringIO("\n".join(lines))))
    if len(rd) < 3:
        return []
    header, units = rd[0], rd[1]
    res = []
    for r in rd[2:]:
        d = {"kernel": r[header.index("Kernel Name")][:70], "id": r[0]}
        for i, (h, u) in enumerate(zip(header, units)):
            if h in WANT and i < len(r) and r[i] != "":
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                key = WANT[h]
                if key.endswith("_MB") or key == "dur_us":
                    v *= SCALE.get(u, 1.0)
                d[key] = v
        res.append(d)
    return res


def main():
    cols = ["dur_us", "dram_rd_MB", "dram_wr_MB", "dram_pct", "tensor_pct", "sm_pct", "l2_MB", "l2_hit_pct", "occ_pct", "regs",
            "grid", "block", "dsmem_B"]
    allr = {}
    for rep in sys.argv[1:]:
        rs = rows_of(rep)
        allr[rep] = rs
        print(f"\n### {rep}\n")
        print("| # | kernel | " + " | ".join(cols) + " | DRAM GB/s |")
        print("|---|---|" + "---|" * (len(cols) + 1))
        for d in rs:
            if "dram_pct" not in d and "dram_pct2" in d:
                d["dram_pct"] = d["dram_pct2"]
            gbs = (d.get("dram_rd_MB", 0) + d.get("dram_wr_MB", 0)) / max(d.get("dur_us", 1e9), 1e-9) * 1e3
            cells = []
            for c in cols:
                v = d.get(c)
                cells.append("" if v is None else (f"{v:.0f}" if abs(v) >= 100 else f"{v:.2f}"))
            print(f"| {d['id']} | `{d['kernel']}` | " + " | ".join(cells) + f" | {gbs:.0f} |")
    json.dump(allr, open("/tmp/ncu_summary.json", "w"), indent=1)


if __name__ == "__main__":
    main()
