#!/usr/bin/env python
"""profiles/r2_kernel_traffic.json (what bench.py quotes as `traffic`) and the per-launch summaries profiles/r2_step_{snp,all}_ncu.csv
from the metric logs of tools/ncu_step_metrics.sh:
    python tools/kernel_traffic.py gpurun_out/r2f_snp_metrics.csv gpurun_out/r2f_all_metrics.csv SITES ALIGNED_BASES INDEL_SITES
SITES / ALIGNED_BASES: candidate sites and aligned bases of the 20 Mb snps probe, INDEL_SITES: key positions of the 20 Mb all probe
(bench.py prints them in `config`)."""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "nsecond": 1e-6, "ns": 1e-6, "usecond": 1e-3, "us": 1e-3, "msecond": 1.0, "ms": 1.0, "second": 1e3}
SHORT = {"gpu__time_duration.sum": "duration_ms", "dram__bytes_read.sum": "dram_read_B", "dram__bytes_write.sum": "dram_write_B",
         "smsp__inst_executed.sum": "warp_inst", "sm__inst_issued.avg.pct_of_peak_sustained_active": "issue_pct",
         "sm__warps_active.avg.pct_of_peak_sustained_active": "occupancy_pct", "launch__registers_per_thread": "regs",
         "launch__grid_size": "grid", "launch__block_size": "block", "smsp__thread_inst_executed_per_inst_executed.ratio": "lanes_per_inst",
         "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct"}


def kernel_name(s):
    s = re.sub(r"\(.*", "", s).replace("void ", "").replace("nc::", "")
    return re.sub(r"<.*", "", s)


def launches(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    col = {n: i for i, n in enumerate(rows[hdr])}
    out = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) < len(col):
            continue
        d = out.setdefault(r[col["ID"]], {"kernel": kernel_name(r[col["Kernel Name"]])})
        m = r[col["Metric Name"]]
        if m in SHORT:
            try:
                d[SHORT[m]] = float(r[col["Metric Value"]].replace(",", "")) * SCALE.get(r[col["Metric Unit"]], 1.0)
            except ValueError:
                pass
    ls = list(out.values())
    # exactly one step: a step begins with K0a (cigar_scan_kernel)
    starts = [i for i, d in enumerate(ls) if d["kernel"] == "cigar_scan_kernel"]
    if len(starts) >= 2:
        return ls[starts[0]:starts[1]]
    return ls[starts[0]:] if starts else ls


def write_summary(ls, path):
    keys = ["kernel"] + list(SHORT.values())
    with open(path, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(keys)
        for d in ls:
            w.writerow([d.get(k, "") if k == "kernel" else ("%.6g" % d[k] if k in d else "") for k in keys])


def per_kernel(ls):
    agg = collections.OrderedDict()
    for d in ls:
        a = agg.setdefault(d["kernel"], {"dram_bytes": 0.0, "duration_ms": 0.0, "launches": 0})
        a["dram_bytes"] += d.get("dram_read_B", 0.0) + d.get("dram_write_B", 0.0)
        a["duration_ms"] += d.get("duration_ms", 0.0)
        a["launches"] += 1
    return agg


def main():
    snp, al = launches(sys.argv[1]), launches(sys.argv[2])
    sites, bases, isites = int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    write_summary(snp, os.path.join(ROOT, "profiles", "r2_step_snp_ncu.csv"))
    write_summary(al, os.path.join(ROOT, "profiles", "r2_step_all_ncu.csv"))
    ks, ka = per_kernel(snp), per_kernel(al)
    tot = lambda agg, names: sum(agg[n]["dram_bytes"] for n in names if n in agg)
    out = {"source": "profiles/r2_step_snp_ncu.csv and r2_step_all_ncu.csv (tools/ncu_step_metrics.sh: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,"
                     "gpu__time_duration.sum,... --clock-control none over ONE step of bench.py on 20 Mb probes, NC_BENCH_LEN / NC_BENCH_ALL_LEN = 2e7): "
                     "(dram__bytes_read.sum + dram__bytes_write.sum) summed over the step's launches of a kernel / units of that step, scaled by bench.py to its own unit counts",
           "sites": sites, "aligned_bases": bases, "indel_sites_approx": isites,
           "cnn_dram_bytes_per_site": tot(ks, ["tc_trunk_a_kernel", "tc_trunk_b_kernel", "tc_fc_kernel"]) / sites,
           "tensor_dram_bytes_per_site": tot(ks, ["tensor_kernel"]) / sites,
           "scan_dram_bytes_per_base": tot(ks, ["cigar_scan_kernel", "seq_codes_kernel", "row_fill_kernel", "scan_kernel", "site_list_kernel", "nmat_len_kernel",
                                                "nmat_fill_kernel", "prefix_max_kernel"]) / bases,
           "indel_build_dram_bytes_per_site": tot(ka, ["indel_site_reads_kernel", "indel_align2_kernel", "indel_align_kernel", "indel_msa_kernel", "indel_allele_kernel"]) / isites,
           "indel_cnn_dram_bytes_per_site": tot(ka, ["tci_trunk_a_kernel", "tci_trunk_b_kernel", "tci_fc_kernel"]) / isites,
           "per_kernel": {"snps": ks, "all": ka}}
    # bench.py looks the dominant kernel up by its plain name
    out["per_kernel"].update({k: v for k, v in ks.items()})
    json.dump(out, open(os.path.join(ROOT, "profiles", "r2_kernel_traffic.json"), "w"), indent=1)
    for name, agg in (("snps", ks), ("all", ka)):
        t = sum(a["duration_ms"] for a in agg.values())
        print("== %s step (20 Mb probe): %.3f ms under ncu" % (name, t))
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["duration_ms"])[:14]:
            print("   %-28s %8.3f ms %5.1f%%  %9.1f MB DRAM  x%d" % (k, a["duration_ms"], 100 * a["duration_ms"] / t, a["dram_bytes"] / 1e6, a["launches"]))


if __name__ == "__main__":
    main()
