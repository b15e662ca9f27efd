"""Print a compact per-kernel summary (time, occupancy limits, issue rate, stall reasons, DRAM bytes)
from an .ncu-rep captured with --set full.  usage: python tools/ncu_brief.py report.ncu-rep"""
import csv, subprocess, sys, io

def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    want = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
            'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor',
            'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
            'smsp__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
            'launch__shared_mem_per_block_dynamic', 'lts__t_bytes.sum',
            'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
            'dram__throughput.avg.pct_of_peak_sustained_elapsed']
    for r in rows[2:]:
        print(r[hdr.index('Kernel Name')][:90], r[hdr.index('Grid Size')])
        for w in want:
            if w in hdr:
                print(f"   {w:60s} {r[hdr.index(w)]:>16s} {units[hdr.index(w)]}")
        st = []
        for i, h in enumerate(hdr):
            if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio'):
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v > 0.15:
                    st.append((v, h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]))
        print("   stalls/issue: " + ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)))

if __name__ == "__main__":
    main(sys.argv[1])
