"""Summarise ncu reports (--set full) into one compact JSON: per launch, the metrics the roofline discussion uses.

    python tools/ncu_summary.py gpurun_out/ncu_eb.ncu-rep gpurun_out/ncu_gemm.ncu-rep ... > profiles/r01_ncu_full_summary_v2.json
"""
import csv
import io
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "us",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
    "sm__inst_executed_pipe_tensor.sum": "tensor_inst",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__inst_executed.sum": "warp_inst",
    "sm__cycles_elapsed.max": "cycles",
}


def main():
    out = []
    for path in sys.argv[1:]:
        txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units = rows[0], rows[1]
        ik = hdr.index("Kernel Name")
        for r in rows[2:]:
            if len(r) < len(hdr):
                continue
            name = r[ik]
            short = name.split("(")[0].replace("void ", "").replace("unnamed>::", "")
            e = {"kernel": short + (name[name.index("<"):name.index(">") + 1] if "<" in name.split("(")[0] else ""), "src": path.split("/")[-1]}
            for m, k in WANT.items():
                if m in hdr:
                    v = r[hdr.index(m)]
                    try:
                        v = float(v.replace(",", ""))
                    except ValueError:
                        pass
                    e[k] = v
                    u = units[hdr.index(m)]
                    if k in ("dram_read", "dram_write", "us"):
                        e[k + "_unit"] = u
            out.append(e)
    json.dump(out, sys.stdout, indent=0)


if __name__ == "__main__":
    main()
