"""Summarise an `ncu --set full` report (read here with `ncu -i ... --page raw --csv`) as a markdown table and, for the
decode capture, the per-launch DRAM traffic JSON bench.py reads (profiles/r02_int4_traffic.json).
  python scripts/ncu_summary.py gpurun_out/r02_int4_decode.ncu-rep profiles/r02_int4_decode_ncu.md [profiles/r02_int4_traffic.json]"""
import csv
import io
import json
import re
import subprocess
import sys

rep, out_md = sys.argv[1], sys.argv[2]
out_json = sys.argv[3] if len(sys.argv) > 3 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]


def col(pat):
    for i, h in enumerate(hdr):
        if re.fullmatch(pat, h):
            return i
    return None


COLS = [
    ("kernel", r"Kernel Name"), ("grid", r"launch__grid_size"), ("regs", r"launch__registers_per_thread"),
    ("us", r"gpu__time_duration\.sum"), ("dram rd MB", r"dram__bytes_read\.sum"), ("dram wr MB", r"dram__bytes_write\.sum"),
    ("dram % peak", r"gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed"),
    ("tensor pipe % (realtime)", r".*sm__pipe_tensor_cycles_active_realtime\.avg\.pct_of_peak_sustained_elapsed"),
    ("hmma % active", r"sm__pipe_tensor_subpipe_hmma_cycles_active\.avg\.pct_of_peak_sustained_active"),
    ("issue active %", r"smsp__issue_active\.avg\.pct_of_peak_sustained_active"),
    ("warps active %", r"sm__warps_active\.avg\.pct_of_peak_sustained_active"),
    ("smem ld bank conflicts", r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld\.sum"),
    ("smem ld wavefronts", r"l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld\.sum"),
    ("inst executed", r"smsp__inst_executed\.sum"),
]
idx = [(n, col(p)) for n, p in COLS]
lines = ["| # | " + " | ".join(n for n, _ in idx) + " |", "|" + "---|" * (len(idx) + 1)]
recs = []
for k, r in enumerate(data):
    vals = []
    rec = {}
    for n, c in idx:
        v = r[c] if c is not None else ""
        if n == "kernel":
            v = re.sub(r"\(.*", "", v).replace("void ", "")
        rec[n] = v
        try:
            f = float(v.replace(",", ""))
            v = f"{f:.2f}" if abs(f) < 1000 and not f.is_integer() else f"{f:.0f}"
        except ValueError:
            pass
        vals.append(v)
    recs.append(rec)
    lines.append(f"| {k} | " + " | ".join(vals) + " |")
with open(out_md, "w") as f:
    f.write(f"Source: `{rep}` (`ncu --set full --clock-control none`, one launch per row, serialised and cold: durations are not\n"
            "bench values).  Launch order: see scripts/gpu_ncu_shapes_r2.py.\n\n" + "\n".join(lines) + "\n")
print("\n".join(lines))
if out_json:
    def mb(r):
        return (float(r["dram rd MB"]) + float(r["dram wr MB"])) * 1e6
    half = len(recs) // 2
    js = {"source": rep, "note": "dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over the four launches of a fused layer",
          "bs32": {"dram_bytes_per_launch": sum(mb(r) for r in recs[:half]) / half, "launches": half},
          "bs1": {"dram_bytes_per_launch": sum(mb(r) for r in recs[half:]) / (len(recs) - half), "launches": len(recs) - half}}
    json.dump(js, open(out_json, "w"), indent=1)
    print(js)
