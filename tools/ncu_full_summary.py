"""Key metrics of every kernel in an `ncu --page raw --csv` dump: python tools/ncu_full_summary.py raw.csv"""
import csv
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.max", "SM cycles"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/CTA"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe (tcgen05) cycles active %"),
    ("sm__inst_executed_pipe_tc.sum", "tcgen05 instructions"), ("sm__inst_executed_pipe_tmem.sum", "TMEM ld/st instructions"),
    ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor-memory cycles active %"),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "tensor-core smem wavefronts"),
    ("sm__inst_executed_pipe_xu.sum", "MUFU instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("== %s" % r[idx["Kernel Name"]][:110])
        for k, label in KEYS:
            if k in idx and r[idx[k]] not in ("", "n/a"):
                print("   %-40s %16s %s" % (label, r[idx[k]], units[idx[k]]))


if __name__ == "__main__":
    main()
