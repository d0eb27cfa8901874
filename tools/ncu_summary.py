"""Summarise ncu artefacts from gpurun_out/ into small, committed text files under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches_r1.csv profiles/r1_launches.md
  python tools/ncu_summary.py kernel   gpurun_out/prof_conv_r1.ncu-rep profiles/r1_conv_kernel.md
  python tools/ncu_summary.py traffic  gpurun_out/prof_conv_r2.ncu-rep conv_gate_B1_T862 EpiGate profiles/r2_conv_kernel_full.md
        (records dram bytes per launch of the kernels matching <substring> in profiles/traffic.json, which
         bench.py reads for roofline.traffic)
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.sum",
    "sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.sum.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_op_utchmma_src_fp16_dst_fp32_sparsity_off.max.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "sm__inst_executed.sum",
]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        if row["Metric Unit"] in ("us", "usecond"):
            v *= 1000.0
        elif row["Metric Unit"] in ("ms", "msecond"):
            v *= 1e6
        a = agg.setdefault(row["Kernel Name"], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)\n\n")
        f.write("source: %s   total %.1f us over %d launches\n\n" % (src, tot / 1000, sum(a[0] for a in agg.values())))
        f.write("| launches | avg us | share | kernel |\n|---:|---:|---:|---|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| %d | %.2f | %.1f%% | `%s` |\n" % (a[0], a[1] / a[0] / 1000, 100 * a[1] / tot, k[:140]))


def kernel(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write("# ncu --set full summary of %s\n\n" % src)
        for r in rows[2:]:
            f.write("## %s  (id %s)\n\n| metric | value | unit |\n|---|---:|---|\n" % (r[hdr.index("Kernel Name")][:120], r[0]))
            for i, h in enumerate(hdr):
                if h in KEYS:
                    f.write("| %s | %s | %s |\n" % (h, r[i], units[i]))
            f.write("\n")


def traffic(src, key, match, cite):
    import json
    import os
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    vals = []
    for r in rows[2:]:
        if match not in r[hdr.index("Kernel Name")]:
            continue
        tot = 0.0
        for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(name)
            tot += float(r[i].replace(",", "")) * scale[units[i]]
        vals.append(tot)
    assert vals, "no kernel matching %r in %s" % (match, src)
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
    table = json.load(open(path)) if os.path.exists(path) else {}
    table[key] = {"dram_bytes_per_launch": sum(vals) / len(vals), "launches_captured": len(vals),
                  "source": "%s (ncu --set full, kernels matching '%s', dram__bytes_read.sum + dram__bytes_write.sum)" % (cite, match)}
    json.dump(table, open(path, "w"), indent=2)
    print(key, table[key])


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(*sys.argv[2:6])
    else:
        {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2], sys.argv[3])
