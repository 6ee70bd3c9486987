"""Key metrics per kernel launch of an .ncu-rep (read with `ncu -i ... --page raw --csv`) as a small CSV on stdout."""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sectors_srcunit_tex.sum", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed.sum", "smsp__warps_issue_stalled_long_scoreboard_per_warp_active.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active"]
idx = [hdr.index(w) for w in want if w in hdr]
w = csv.writer(sys.stdout)
w.writerow([hdr[i] for i in idx])
w.writerow([units[i] for i in idx])
for r in data:
    w.writerow([r[i][:60] for i in idx])
