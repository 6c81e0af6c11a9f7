"""Summarise an `ncu --page raw --csv` export: one row per captured launch with the metrics the README tables quote.
  python profiles/ncu_summary.py gpurun_out/r02_loop.raw.csv [more.csv ...] > profiles/r02_..._summary.csv"""
import csv, sys
COLS = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
        ("sm__cycles_elapsed.avg.per_second", "sm_clock"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_active_pct"),
        ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"), ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1_pct"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma_pipe_pct"), ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu_pipe_pct"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_pipe_pct"), ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu_pipe_pct"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_wavefront_pct"),
        ("smsp__average_warp_latency_issue_stalled_long_scoreboard.pct", "stall_long_sb"), ("smsp__average_warp_latency_issue_stalled_short_scoreboard.pct", "stall_short_sb"),
        ("smsp__average_warp_latency_issue_stalled_barrier.pct", "stall_barrier"), ("smsp__average_warp_latency_issue_stalled_math_pipe_throttle.pct", "stall_math"),
        ("smsp__average_warp_latency_issue_stalled_lg_throttle.pct", "stall_lg"), ("smsp__average_warp_latency_issue_stalled_mio_throttle.pct", "stall_mio")]
w = csv.writer(sys.stdout)
first = True
for path in sys.argv[1:]:
    rows = [r for r in csv.reader(open(path)) if len(r) > 20]
    hdr, units = rows[0], rows[1]
    idx = [(hdr.index(c), n) for c, n in COLS if c in hdr]
    if first:  # same two header rows as the raw export: metric names, then units
        w.writerow([hdr[i] for i, n in idx])
        w.writerow([units[i] for i, n in idx])
        first = False
    for r in rows[2:]:
        w.writerow([(r[i][:90] if n == "kernel" else r[i]) for i, n in idx])
