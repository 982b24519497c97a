"""Turn an .ncu-rep into the small JSON summaries kept under profiles/ (run in the build container)."""
import csv
import io
import json
import subprocess
import sys

KEEP = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'sm__cycles_active.avg',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active']


def main(rep, out_summary, out_traffic):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out, traffic = [], {}
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        item = {k: (d[k] + ' ' + units[hdr.index(k)]).strip() for k in KEEP if k in d}
        item['stalls_per_issue'] = {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''): round(float(d[h]), 2)
                                    for h in hdr if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('per_issue_active.ratio') and float(d[h]) > 0.15}
        out.append(item)
        name = d['Kernel Name'].split('(')[0]
        rd, wr = float(d['dram__bytes_read.sum']), float(d['dram__bytes_write.sum'])
        mult = {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1.0}
        rdb = rd * mult.get(units[hdr.index('dram__bytes_read.sum')], 1.0)
        wrb = wr * mult.get(units[hdr.index('dram__bytes_write.sum')], 1.0)
        traffic[name] = {"dram_read_bytes": rdb, "dram_write_bytes": wrb, "bytes_per_launch": rdb + wrb}
    json.dump(out, open(out_summary, 'w'), indent=1)
    json.dump(traffic, open(out_traffic, 'w'), indent=1)
    print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    main(*sys.argv[1:4])
