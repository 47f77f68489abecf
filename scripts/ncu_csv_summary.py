"""Markdown summary + profiles/ncu_traffic.json from the raw-page CSVs that scripts/gpu_round.sh writes on the GPU box
(`ncu -i X.ncu-rep --page raw --csv`).  Usage: python scripts/ncu_csv_summary.py TAG   (reads gpurun_out/TAG_ncu_*.raw.csv or
profiles/TAG_ncu/*.raw.csv.gz, writes profiles/TAG_ncu/summary.md and profiles/ncu_traffic.json)."""
import csv
import glob
import gzip
import io
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = [("gpu__time_duration.sum", "dur us", 1), ("dram__bytes_read.sum", "DRAM rd MB", 1), ("dram__bytes_write.sum", "DRAM wr MB", 1),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %", 0),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %", 0),
        ("sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active", "HMMA pipe %", 0),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %", 0), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %", 0),
        ("lts__t_sector_hit_rate.pct", "L2 hit %", 0), ("launch__registers_per_thread", "regs", 0), ("launch__grid_size", "grid", 0),
        ("launch__block_size", "block", 0), ("launch__cluster_size", "cluster", 0), ("smsp__inst_executed.sum", "warp inst (M)", 2)]
UNIT = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6, "ns": 1e-3,
        "us": 1.0, "ms": 1e3}
CATS = (("attn_decode", "attn_decode"), ("gemm_bf16_tcgen05", "gemm_bf16_tcgen05"), ("vq_gather", "vq_gather"), ("vq_argmin", "vq_argmin"),
        ("attn_prefill", "attn_prefill_f32"), ("layer_norm", "layer_norm"), ("instance_norm", "instance_norm"))


def read(path):
    f = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
    rows = list(csv.reader(io.StringIO("".join(l for l in f if l.startswith('"')))))
    return (rows[0], rows[1], rows[2:]) if len(rows) > 2 else (None, None, [])


def main():
    tag = sys.argv[1]
    paths = sorted(glob.glob(os.path.join(ROOT, "gpurun_out", f"{tag}_ncu_*.raw.csv")) +
                   glob.glob(os.path.join(ROOT, "profiles", f"{tag}_ncu", "*.raw.csv.gz")))
    out = [f"# ncu --set full captures, run {tag} (B200, full workload B=256 x T=300 unless noted; --clock-control none)\n",
           "Per launch; cold-cache and serialised by the profiler: compare SHARES and utilisation, not absolute times.\n"]
    traffic = {}
    seen = set()
    for path in paths:
        name = os.path.basename(path).split(".raw.csv")[0]
        if name in seen:
            continue
        seen.add(name)
        hdr, units, rows = read(path)
        if not rows:
            continue
        out.append(f"\n## {name}\n")
        out.append("| # | kernel | " + " | ".join(c[1] for c in COLS) + " | DRAM GB/s |")
        out.append("|---|---|" + "---|" * (len(COLS) + 1))
        for r in rows:
            kn = r[hdr.index("Kernel Name")].replace("void ", "").replace("<unnamed>::", "")[:64]
            vals = {}
            for key, label, kind in COLS:
                if key not in hdr or r[hdr.index(key)] == "":
                    vals[label] = None
                    continue
                v = float(r[hdr.index(key)].replace(",", ""))
                u = units[hdr.index(key)]
                if kind == 1:
                    v *= UNIT.get(u, 1.0)
                if kind == 2:
                    v *= 1e-6
                vals[label] = v
            tot_mb = (vals["DRAM rd MB"] or 0) + (vals["DRAM wr MB"] or 0)
            gbs = tot_mb / max(vals["dur us"] or 1e9, 1e-9) * 1e3
            cells = ["" if vals[c[1]] is None else (f"{vals[c[1]]:.0f}" if abs(vals[c[1]]) >= 100 else f"{vals[c[1]]:.2f}") for c in COLS]
            out.append(f"| {r[0]} | `{kn}` | " + " | ".join(cells) + f" | {gbs:.0f} |")
            for pat, cat in CATS:
                if pat in kn:
                    traffic.setdefault(cat, []).append(tot_mb * 1e6)
    os.makedirs(os.path.join(ROOT, "profiles", f"{tag}_ncu"), exist_ok=True)
    open(os.path.join(ROOT, "profiles", f"{tag}_ncu", "summary.md"), "w").write("\n".join(out) + "\n")
    tj = {k: sum(v) / len(v) for k, v in traffic.items()}
    tj["_note"] = (f"dram__bytes_read.sum + dram__bytes_write.sum per launch (bytes), averaged over the launches captured in run {tag} "
                   "(profiles/%s_ncu/summary.md); attn_decode: 256 clips x 12 heads, self (~230 keys) and cross (300 keys) launches" % tag)
    json.dump(tj, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
    print("\n".join(out))


if __name__ == "__main__":
    main()
