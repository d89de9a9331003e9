"""Summarise an ncu report (ncu -i X.ncu-rep --page raw --csv) into the JSON kept under profiles/:
    python tools/ncu_summary.py gpurun_out/X.ncu-rep 'csr_stream_kernel' > profiles/rN_ncu_<workload>_summary.json
One entry per distinct kernel name matching the regex (first profiled launch of each)."""
import csv
import io
import json
import re
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1tex_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "lts_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "regs",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "sm__inst_executed.avg.per_cycle_elapsed": "ipc",
    "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed": "l1tex2xbar_req_pct",
    "lts__t_requests_srcunit_tex.sum": "l2_requests_from_sm",
    "l1tex__t_output_wavefronts_pipe_tex_mem_texture.sum": "tex_wavefronts",
    "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum": "global_load_wavefronts",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "shared_wavefronts",
    "sm__cycles_elapsed.sum": "sm_cycles_elapsed_sum",
    "launch__grid_size": "grid",
    "launch__occupancy_limit_registers": "occ_limit_regs",
    "launch__occupancy_limit_shared_mem": "occ_limit_smem",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard_per_issue",
}


def main():
    rep, pat = sys.argv[1], re.compile(sys.argv[2])
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    name_i = hdr.index("Kernel Name")
    out = {}
    for r in rows[2:]:
        name = r[name_i]
        if not pat.search(name) or name in out:
            continue
        e = {}
        for i, h in enumerate(hdr):
            if h in WANT and r[i] not in ("", "n/a"):
                e[WANT[h]] = (r[i] + " " + units[i]).strip()
        out[name] = e
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
