#!/usr/bin/env python
"""Compact per-launch summary of an `ncu --set full` report (run where ncu is installed; no GPU needed):
  python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_<name>_ncu.json
Keeps the metrics the roofline discussion needs: duration, DRAM bytes, L2->SM bytes, tensor-pipe activity, issue slots."""
import csv
import io
import json
import subprocess
import sys

KEEP = {
    "Kernel Name": "kernel", "Grid Size": "grid", "Block Size": "block",
    "gpu__time_duration.sum": "duration_us",
    "sm__cycles_elapsed.max": "sm_cycles",
    "dram__bytes_read.sum": "dram_read_MB", "dram__bytes_write.sum": "dram_write_MB",
    "l1tex__m_xbar2l1tex_read_bytes.sum": "l2_to_sm_MB",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "launch__registers_per_thread": "regs",
    "launch__shared_mem_per_block_dynamic": "dyn_smem_KB",
    "launch__waves_per_multiprocessor": "waves",
}


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    idx = {hdr.index(k): v for k, v in KEEP.items() if k in hdr}
    res = []
    for r in rows[2:]:
        d = {}
        for i, name in idx.items():
            v = r[i]
            try:
                v = float(v.replace(",", ""))
            except ValueError:
                v = v[:90]
            d[name] = v
        res.append(d)
    json.dump(dict(report=path, units="as printed by ncu (duration us, bytes MB, smem KB)", launches=res), sys.stdout, indent=1)


if __name__ == "__main__":
    main(sys.argv[1])
