"""ncu -i X.ncu-rep --page raw --csv | python profiles/scripts/ncu_summary.py  ->  one row per captured launch with the
columns the roofline discussion uses (profiles/README.md)."""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
cols = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "dram_rd"),
        ("dram__bytes_write.sum", "dram_wr"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "tensor_inst%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
        ("launch__shared_mem_per_block_dynamic", "dsmem"), ("lts__t_sector_hit_rate.pct", "l2hit%"),
        ("smsp__inst_executed.sum", "warp_inst")]
cols = [(c, n) for c, n in cols if c in ix]
w = csv.writer(sys.stdout)
w.writerow([n + ("" if not units[ix[c]] or n in ("kernel",) else " [" + units[ix[c]] + "]") for c, n in cols])
for r in rows[2:]:
    out = []
    for c, n in cols:
        v = r[ix[c]]
        if n == "kernel":
            v = v.split("(")[0][:70]
        out.append(v)
    w.writerow(out)
