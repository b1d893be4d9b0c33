"""Summarise an .ncu-rep (raw page) into the handful of metrics the roofline discussion uses."""
import csv, subprocess, sys
WANT = ["gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
 "sm__throughput.avg.pct_of_peak_sustained_elapsed","launch__registers_per_thread","launch__grid_size","launch__block_size",
 "launch__occupancy_limit_registers","launch__occupancy_limit_shared_mem","sm__warps_active.avg.pct_of_peak_sustained_active",
 "smsp__inst_executed.sum","smsp__issue_active.avg.pct_of_peak_sustained_active","sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
 "sm__inst_executed_pipe_xu.sum","lts__t_bytes.sum","l1tex__t_bytes.sum","lts__t_sector_hit_rate.pct","sm__cycles_elapsed.max","launch__waves_per_multiprocessor"]
def main(rep):
    raw = subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, units = rows[0], rows[1]
    for r in rows[2:]:
        print("kernel:", r[h.index("Kernel Name")][:90])
        for w in WANT:
            if w in h:
                i = h.index(w); print(f"  {w:72s} {r[i]:>18s} {units[i]}")
        stalls = sorted(((float(r[i].replace(',','')), h[i]) for i in range(len(h)) if h[i].startswith("smsp__average_warps_issue_stalled_") and r[i]), reverse=True)[:7]
        for v, k in stalls: print(f"  warps stalled per issue: {k.split('smsp__average_warps_issue_stalled_')[1].replace('_per_issue_active.ratio',''):30s} {v:8.3f}")
if __name__ == "__main__": main(sys.argv[1])
