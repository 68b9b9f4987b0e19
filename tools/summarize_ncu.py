"""Summarise an .ncu-rep (read here, no GPU) into a small CSV for profiles/.
python tools/summarize_ncu.py gpurun_out/prof_gemm_r1.ncu-rep profiles/gemm_r1.csv"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "id"] + KEYS)
        w.writerow(["(unit)", ""] + [units[ix[k]] if k in ix else "" for k in KEYS])
        for r in rows[2:]:
            name = r[ix["Kernel Name"]].split("(")[0]
            w.writerow([name, r[ix["ID"]]] + [r[ix[k]] if k in ix else "" for k in KEYS])
    print("wrote", out, len(rows) - 2, "launches")


if __name__ == "__main__":
    main()
