"""ncu report(s) -> compact JSON summary for profiles/ (run where the .ncu-rep files are; needs the `ncu` CLI, no GPU).
Usage: python tests/debug_tools/summarize_ncu.py out.json rep1.ncu-rep [rep2.ncu-rep ...]"""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "time_us",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "lts__t_bytes.sum": "l2_bytes",
    "l1tex__m_xbar2l1tex_read_bytes.sum": "l2_to_sm_read_bytes",
    "smsp__inst_executed.sum": "warp_instructions",
    "sm__cycles_active.avg": "sm_cycles_active_avg",
    "sm__cycles_elapsed.max": "sm_cycles_elapsed_max",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__shared_mem_per_block_dynamic": "smem_dynamic",
    "launch__occupancy_limit_shared_mem": "occupancy_limit_smem_blocks",
}


def unit_scale(u):
    return {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3,
            "msecond": 1e3}.get(u, 1.0)


def summarize(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")][:200]}
        for i, h in enumerate(hdr):
            base = h.split(".TriageCompute.")[-1] if ".TriageCompute." in h else h
            if base in KEYS and r[i] not in ("", "n/a"):
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                if KEYS[base] in ("time_us", "dram_read", "dram_write", "l2_bytes", "l2_to_sm_read_bytes"):
                    v *= unit_scale(units[i])
                d[KEYS[base]] = v
        out.append(d)
    return out


def main():
    out_path, reps = sys.argv[1], sys.argv[2:]
    res = {"tool": "ncu --set full --clock-control none (B200); times are cold-cache single launches under the profiler, never bench values",
           "captures": {p.split("/")[-1]: summarize(p) for p in reps}}
    json.dump(res, open(out_path, "w"), indent=1)
    for k, v in res["captures"].items():
        for d in v:
            print(k, {kk: d[kk] for kk in ("kernel", "time_us", "dram_read", "dram_write", "warp_instructions") if kk in d})


if __name__ == "__main__":
    main()
