"""Condense `ncu --set full` reports into the few numbers the profiles/ summaries quote.

    python tools/ncu_summary.py gpurun_out/prof_*.ncu-rep

Reads each report with `ncu -i <rep> --page raw --csv` and prints, per captured launch: duration,
DRAM bytes, achieved occupancy / issue utilisation, the tensor (DMMA) / fp64 / fma pipe
utilisation, shared-memory wavefronts and the top warp-stall reasons per issued instruction.
"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__cluster_size", "cluster"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "dmma_pipe_pct"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma_pipe_pct"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_pct"),
]


def main():
    for rep in sys.argv[1:]:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        if len(rows) < 3:
            print(f"== {rep}: no launches captured")
            continue
        hdr, units = rows[0], rows[1]
        unit_of = dict(zip(hdr, units))
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            print(f"== {rep.split('/')[-1]} :: {d.get('Kernel Name', '?')[:90]}")
            for key, name in WANT:
                if key in d and d[key] not in ("", "n/a"):
                    print(f"   {name:20s} {d[key]} {unit_of.get(key, '')}")
            stalls = []
            for h in hdr:
                if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                    try:
                        stalls.append((float(d[h].replace(",", "")), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                    except ValueError:
                        pass
            stalls.sort(reverse=True)
            print("   stalls/issue        " + ", ".join(f"{n} {v:.2f}" for v, n in stalls[:6]))


if __name__ == "__main__":
    main()
