"""Summarise one `ncu --set full` report: python tools/summarize_ncu_full.py REPORT.ncu-rep OUT_SUMMARY.csv [TRAFFIC.json]

Writes the per-launch columns the DESIGN.md kernel table quotes (duration, registers, occupancy, issue utilisation,
instructions, DRAM bytes, pipe utilisation, bank conflicts, the four dominant stall ratios) and, optionally, the
per-kernel DRAM traffic file bench.py reads for `roofline.traffic` (first launch of each kernel)."""
import csv
import json
import subprocess
import sys

KEEP = ['ID', 'Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_static', 'launch__shared_mem_per_block_dynamic',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio']
SCALE = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    traffic_path = sys.argv[3] if len(sys.argv) > 3 else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = [hdr.index(k) for k in KEEP if k in hdr]

    def nbytes(r, k):
        return float(r[hdr.index(k)]) * SCALE.get(units[hdr.index(k)], 1.0)

    traffic = {}
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in ix])
        w.writerow([units[i] for i in ix])
        for r in rows[2:]:
            w.writerow([r[i][:60] for i in ix])
            name = r[hdr.index('Kernel Name')].split('(')[0].replace('void ', '').split('<')[0]
            traffic.setdefault(name, {
                'dram_bytes': int(nbytes(r, 'dram__bytes_read.sum') + nbytes(r, 'dram__bytes_write.sum')),
                'duration_us': float(r[hdr.index('gpu__time_duration.sum')]) *
                               {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(units[hdr.index('gpu__time_duration.sum')], 1.0)})
    if traffic_path:
        json.dump({'source': f'{out} (ncu --set full --clock-control none, tools/dev_profile.py ours 2000000 1920 1280 '
                             'sh 1; first launch of each kernel)',
                   'workload': '2000000 Gaussians, 1920x1280, waymo_ring[1], SH degree 3', 'kernels': traffic},
                  open(traffic_path, 'w'), indent=1)
    for k, v in traffic.items():
        print(f"{k:34s} {v['duration_us']:9.1f} us  {v['dram_bytes'] / 1e6:9.1f} MB")


if __name__ == "__main__":
    main()
