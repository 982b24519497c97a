"""Turn an .ncu-rep into the small JSON summaries kept under profiles/ (run in the build container).

usage: python tools/summarize_ncu.py <rep> <summary.json> <traffic.json> [git-sha]
The capture covers ONE frame (tools/collect_profiles.sh: -s <warm-up launches> -c <launches per frame>); a kernel that
is launched several times per frame (field_kernel / xf_kernel run as launch pairs over ray-group ranges) is
aggregated: durations, DRAM bytes, L2 bytes and instruction counts are summed, rates are duration-weighted means.
"""
import csv
import io
import json
import subprocess
import sys

SUM = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum', 'lts__t_bytes.sum']
MEAN = ['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active']
ONCE = ['launch__registers_per_thread', 'launch__grid_size', 'launch__block_size']
MULT = {'Tbyte': 1e12, 'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'second': 1e3, 'msecond': 1.0, 'usecond': 1e-3,
        'nsecond': 1e-6, 'ms': 1.0, 'us': 1e-3, 'ns': 1e-6, 's': 1e3}


def main(rep, out_summary, out_traffic, sha="unknown"):
    raw = (open(rep).read() if rep.endswith(".csv") else
           subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    unit = lambda k: units[hdr.index(k)]
    val = lambda d, k: float(d[k].replace(",", "")) * MULT.get(unit(k), 1.0)
    agg = {}
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d['Kernel Name'].split('(')[0]
        a = agg.setdefault(name, {"launches": 0, "sum": {k: 0.0 for k in SUM if k in d}, "mean": {k: 0.0 for k in MEAN if k in d},
                                  "stall": {}, "once": {k: d[k] for k in ONCE if k in d}})
        a["launches"] += 1
        dur = val(d, 'gpu__time_duration.sum')
        for k in a["sum"]:
            a["sum"][k] += val(d, k)
        for k in a["mean"]:
            a["mean"][k] += dur * float(d[k].replace(",", ""))
        for h in hdr:
            if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('per_issue_active.ratio'):
                key = h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')
                a["stall"][key] = a["stall"].get(key, 0.0) + dur * float(d[h])
    out, traffic = [], {"_source": f"ncu --set full, one frame of bench.py cfg3, build {sha}", "_l2_peak_gbs": 20000.0}
    for name, a in agg.items():
        dur = a["sum"]['gpu__time_duration.sum']
        item = {"kernel": name, "launches_per_frame": a["launches"], "gpu_time_ms_sum": round(dur, 4)}
        item.update({k: round(v, 1) for k, v in a["sum"].items() if k != 'gpu__time_duration.sum'})
        item.update({k: round(v / max(dur, 1e-12), 2) for k, v in a["mean"].items()})
        item.update(a["once"])
        item["stalls_per_issue"] = {k: round(v / max(dur, 1e-12), 2) for k, v in a["stall"].items() if v / max(dur, 1e-12) > 0.15}
        out.append(item)
        rd, wr = a["sum"].get('dram__bytes_read.sum', 0.0), a["sum"].get('dram__bytes_write.sum', 0.0)
        traffic[name] = {"dram_read_bytes": rd, "dram_write_bytes": wr, "bytes_per_launch": rd + wr,
                         "launches_per_frame": a["launches"], "l2_bytes": a["sum"].get('lts__t_bytes.sum', 0.0),
                         "note": "summed over the kernel's launches of one frame"}
    json.dump(out, open(out_summary, 'w'), indent=1)
    json.dump(traffic, open(out_traffic, 'w'), indent=1)
    print(json.dumps(traffic, indent=1))


if __name__ == "__main__":
    main(*sys.argv[1:5])
