"""profiles/r2_solver_counters.json from an `ncu --set full` capture of the MPC launches of one config-2 step (development
tool; run where ncu is installed):   python scripts/ncu_counters.py gpurun_out/<tag>_backtest.ncu-rep <decisions in the step>"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, decisions = sys.argv[1], int(sys.argv[2])
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, body = rows[0], rows[1], rows[2:]
    col = {n: i for i, n in enumerate(head)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    nbytes = lambda r, n: float(r[col[n]].replace(",", "")) * scale[units[col[n]]]
    stall = {n: i for n, i in col.items() if re.fullmatch(r"smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio", n)}
    kernels = []
    for r in body:
        f = lambda n: float(r[col[n]].replace(",", ""))
        name = re.sub(r"^void ", "", r[col["Kernel Name"]]).split("(")[0]
        st = {re.sub(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio", r"\1", n): round(float(r[i].replace(",", "")), 2)
              for n, i in stall.items()}
        st = dict(sorted(((k, v) for k, v in st.items() if v >= 0.1), key=lambda kv: -kv[1])[:8])
        kernels.append({
            "kernel": name,
            "duration_ms_under_ncu": round(f("gpu__time_duration.sum"), 3),
            "smsp__inst_executed.sum": int(f("smsp__inst_executed.sum")),
            "dram_bytes": int(nbytes(r, "dram__bytes_read.sum") + nbytes(r, "dram__bytes_write.sum")),
            "issue_active_pct": round(f("sm__issue_active.avg.pct_of_peak_sustained_elapsed"), 2),
            "fp64_pipe_pct": round(f("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"), 2),
            "l1tex_throughput_pct": round(f("l1tex__throughput.avg.pct_of_peak_sustained_active"), 2),
            "registers_per_thread": int(f("launch__registers_per_thread")),
            "dynamic_smem_per_block_kbytes": round(f("launch__shared_mem_per_block_dynamic"), 3),
            "stall_cycles_per_issued_instruction": st,
        })
    inst = sum(k["smsp__inst_executed.sum"] for k in kernels)
    dram = sum(k["dram_bytes"] for k in kernels)
    out = {
        "kernels": kernels,
        "kernel": "MPC stage = backtest_lane_kernel<5,2,4,true> (dense start) + backtest_active_kernel<5,8,true,2,1> (reduced solves) "
                  "+ second-chance / straggler launches (no work)",
        "source": f"profiles/r2_backtest_active_pipeline.txt (ncu --set full of the five MPC launches of one step: dense start, reduced solves, "
                  f"and the second-chance / straggler launches that find no work; python bench.py --paths 1184 --steps 1 --warmup 3 "
                  f"--no-cpu-baseline --no-other-configs; 1184 backtests x 246 decisions = {decisions} decisions per step)",
        "decisions_in_capture": decisions,
        "warp_instructions_per_decision": round(inst / decisions, 1),
        "dram_bytes_per_decision": round(dram / decisions, 1),
    }
    with open(os.path.join(ROOT, "profiles", "r2_solver_counters.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps({k: out[k] for k in ("warp_instructions_per_decision", "dram_bytes_per_decision")}))
    for k in kernels:
        print(k["kernel"], k["duration_ms_under_ncu"], k["issue_active_pct"], k["stall_cycles_per_issued_instruction"])


if __name__ == "__main__":
    main()
