"""Turn an `ncu --set full` report into the JSON summaries kept under profiles/.
   python tools/ncu_summarize.py gpurun_out/gen5.ncu-rep profiles/r1_ncu_gen5_summary.json [profiles/r1_traffic.json]
The second output (optional) holds per-launch DRAM bytes and executed warp-instructions, which bench.py
copies into roofline.traffic / roofline_issue."""
import csv
import io
import json
import re
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "lts__t_bytes.sum",
]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    traffic_out = sys.argv[3] if len(sys.argv) > 3 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    kernels, dram, inst = [], {}, {}
    for r in rows[2:]:
        name = re.sub(r"^void ", "", r[col["Kernel Name"]]).split("(")[0].split("<")[0]
        k = {"kernel": name}
        for m in METRICS:
            if m in col:
                k[m] = f"{r[col[m]]} {units[col[m]]}".strip()
        kernels.append(k)
        b = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            b += float(r[col[m]]) * SCALE[units[col[m]]]
        dram[name] = int(round(b))
        inst[name] = int(float(r[col["smsp__inst_executed.sum"]]))
    json.dump({"report": rep, "kernels": kernels}, open(out, "w"), indent=1)
    if traffic_out:
        try:
            old = json.load(open(traffic_out))
        except Exception:  # noqa: BLE001
            old = {"workload": {"P": 500000, "m": 2.0, "W": 1280, "H": 1024}}
        old["source"] = f"{out} (ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum, smsp__inst_executed.sum; per launch)"
        old["dram_bytes_per_launch"] = dram
        old["warp_instructions_per_launch"] = inst
        json.dump(old, open(traffic_out, "w"), indent=1)
    for k in kernels:
        print(k["kernel"], k.get("gpu__time_duration.sum"), inst[k["kernel"]], dram[k["kernel"]])


if __name__ == "__main__":
    main()
