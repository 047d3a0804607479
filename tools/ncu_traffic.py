"""Extract the DRAM traffic of one kernel from an `ncu --set full` report (run here, no GPU needed) and record it in
profiles/northstar_traffic.json, the file bench.py's `roofline.traffic` is read from.

    python tools/ncu_traffic.py gpurun_out/northstar_r02.ncu-rep conv3x3 32 256 64 profiles/r02_northstar_ncu_metrics.csv
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, pattern, batch, res, ch, out_csv = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), sys.argv[6]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
name_i = hdr.index("Kernel Name")
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.sum",
        "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
sel = [r for r in data if pattern in r[name_i]]
assert sel, "no launch of %r in %s" % (pattern, rep)


def val(r, key):
    i = hdr.index(key)
    return float(r[i].replace(",", "")) * SCALE.get(units[i], 1)


tr = [val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum") for r in sel]
with open(os.path.join(ROOT, out_csv), "w") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit"] + ["launch%d" % i for i in range(len(sel))])
    for k in hdr:
        if k in KEEP or k.startswith("dram__") or k.startswith("sm__pipe_tensor") or k.startswith("smsp__warp_issue_stalled"):
            i = hdr.index(k)
            w.writerow([k, units[i]] + [r[i] for r in sel])
path = os.path.join(ROOT, "profiles", "northstar_traffic.json")
try:
    rec = json.load(open(path))
except Exception:
    rec = {"captures": []}
rec["captures"] = [c for c in rec["captures"] if (c["batch"], c["res"], c["ch"]) != (batch, res, ch)]
rec["captures"].append({"batch": batch, "res": res, "ch": ch, "dram_bytes": round(sum(tr) / len(tr)),
                        "launches": len(tr), "kernel": sel[0][name_i][:80],
                        "source": "%s (ncu --set full --clock-control none; summary %s)" % (os.path.basename(rep), out_csv)})
json.dump(rec, open(path, "w"), indent=1)
print(json.dumps(rec["captures"][-1]))
