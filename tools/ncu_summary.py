"""Condense `ncu --page raw --csv` (+ optionally `--page source --csv`) exports into a readable per-kernel summary.
usage: python tools/ncu_summary.py RAW.csv [SOURCE.csv] > profiles/xyz_summary.txt"""
import csv
import sys

csv.field_size_limit(10 ** 9)
KEYS = [
    ("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"), ("launch__occupancy_limit_registers", "occupancy limit (regs), blocks"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 throughput %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__inst_executed.sum", "warp instructions"),
]


def main():
    rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
    h, u = rows[0], rows[1]
    stall = [x for x in h if "issue_stalled" in x and "per_issue_active" in x and "not_issued" not in x]
    for n, r in enumerate(rows[2:]):
        d = dict(zip(h, r))
        print("=== kernel %d: %s" % (n, d.get("Kernel Name", "?")[:160]))
        for k, label in KEYS:
            if k in d and d[k] != "":
                print("  %-34s %s %s" % (label, d[k], u[h.index(k)]))
        st = sorted(((float(d[k] or 0), k) for k in stall), reverse=True)[:5]
        print("  stalls per issued instruction: " + ", ".join(
            "%s %.2f" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v) for v, k in st))
    if len(sys.argv) > 2:
        rows = list(csv.reader(open(sys.argv[2], errors="ignore")))
        kern, data, seen = None, [], set()

        def flush():
            if not data or kern in seen:
                return
            seen.add(kern)
            tot = sum(x[0] for x in data) or 1
            print("=== hottest instructions (warp stall samples): %s  [%d samples]" % (kern[:120], tot))
            for smp, src, ex in sorted(data, reverse=True)[:14]:
                print("  %5.1f %%  exec %9s  %s" % (100.0 * smp / tot, ex, src[:100]))
        for r in rows:
            if not r:
                continue
            if r[0] == "Kernel Name":
                flush(); kern, data = r[1], []
            elif r[0] != "Address" and len(r) >= 6:
                try:
                    data.append((int(r[4] or 0), r[1].strip(), r[5]))
                except ValueError:
                    pass
        flush()


main()
