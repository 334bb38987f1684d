"""Condense an .ncu-rep (ncu --set full) into a small JSON: one record per captured launch with the metrics the profiles/
notes quote. Usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep profiles/r02_x.json"""
import csv
import json
import subprocess
import sys

METRICS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "lsu_wavefronts_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "shared_wavefronts",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "shared_bank_conflicts",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__shared_mem_per_block_dynamic": "dyn_smem",
    "launch__occupancy_limit_shared_mem": "occ_limit_smem_blocks",
    "launch__occupancy_limit_registers": "occ_limit_reg_blocks",
    "smsp__inst_executed.sum": "warp_instructions",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum": "global_ld_requests",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum": "global_ld_sectors",
    "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum": "global_st_requests",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum": "global_st_sectors",
}
SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3,
         "nsecond": 1e-9, "second": 1.0}


def main(rep, out):
    txt = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True, stderr=subprocess.DEVNULL)
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    recs = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for m, name in METRICS.items():
            if m in hdr:
                i = hdr.index(m)
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                u = units[i]
                if name in ("duration", "dram_read", "dram_write", "dyn_smem") and u in SCALE:
                    v *= SCALE[u]
                d[name] = v
        if "dram_read" in d and "dram_write" in d:
            d["dram_bytes"] = d["dram_read"] + d["dram_write"]
            if d.get("duration"):
                d["dram_gbs"] = d["dram_bytes"] / d["duration"] / 1e9
        recs.append(d)
    with open(out, "w") as f:
        json.dump({"source": rep, "units": "seconds, bytes, percent", "launches": recs}, f, indent=1)
    for d in recs:
        print("%-48s %9.1f us  dram %7.1f MB  %6.0f GB/s  lsu %5.1f%%  fp64 %5.1f%%  occ %5.1f%%" % (
            d["kernel"][:48], d.get("duration", 0) * 1e6, d.get("dram_bytes", 0) / 1e6, d.get("dram_gbs", 0), d.get("lsu_wavefronts_pct", 0),
            d.get("fp64_pipe_pct", 0), d.get("warps_active_pct", 0)))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
