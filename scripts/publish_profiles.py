"""Copy the evidence of one GPU session (gpurun_out/<session>/, written by scripts/gpu_evidence_session.sh) into profiles/
under a round tag, and derive the two tables bench.py / DESIGN.md cite:
  <tag>_launch_shares.txt   per-kernel share of the ncu launch list of the bench command
  ncu_traffic.json          dram read+write bytes per launch of every library kernel (ncu --set full capture)
Usage: python scripts/publish_profiles.py gpurun_out/s16 r01v"""
import collections
import csv
import gzip
import json
import os
import shutil
import sys

# bench.py's stage names (sfb_profile_name) in launch order of one lego_1m forward + backward
STAGES = [("preprocess_kernel", "preprocess"), ("radix_hist_all_kernel", "depth_sort.hist"),
          ("onesweep_pass_kernel<16, 2", "depth_sort.scatter"), ("instance_block_sums_kernel", "instance_block_sums"),
          ("scan_exclusive_kernel", "instance_block_scan"), ("duplicate_kernel", "duplicate"),
          ("radix_hist_all_kernel", "tile_sort.hist"), ("onesweep_pass_kernel<16, 0", "tile_sort.scatter"), ("onesweep_pass_kernel<16, 1", "tile_sort.scatter"),
          ("tile_ranges_kernel", "tile_ranges"), ("render_forward_kernel", "render_forward"),
          ("render_backward", "render_backward"), ("geom_backward_kernel", "geom_backward")]


def launch_shares(path):
    tot = collections.defaultdict(float)
    cnt = collections.Counter()
    for r in csv.DictReader(l for l in open(path) if l.startswith('"')):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = r["Kernel Name"].split("(")[0].replace("void ", "")
        tot[k] += float(r["Metric Value"]) / 1e3
        cnt[k] += 1
    s = sum(tot.values())
    lines = ["ncu --metrics gpu__time_duration.sum --clock-control none -c 400, python bench.py --steps 2 --warmup 1 "
             "(cold-cache, serialised; shares only)"]
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        lines.append(f"{100 * v / s:6.2f}%  {v:9.1f} us  x{cnt[k]:3d}  {k[:100]}")
    return "\n".join(lines) + "\n"


def traffic(raw_csv, src):
    rows = list(csv.reader(open(raw_csv)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    kn, rd, wr = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")

    def to_bytes(v, u):
        return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]

    per = collections.defaultdict(list)
    seen_hist = 0
    for r in data:
        name = r[kn]
        b = to_bytes(r[rd], units[rd]) + to_bytes(r[wr], units[wr])
        if "radix_hist_all_kernel" in name:      # first = depth sort, second = tile sort
            per["depth_sort.hist" if seen_hist == 0 else "tile_sort.hist"].append(b)
            seen_hist += 1
            continue
        for pat, stage in STAGES:
            if pat in name and "radix_hist" not in pat:
                per[stage].append(b)
                break
    out = {k: int(sum(v) / len(v)) for k, v in per.items()}
    out["_source"] = src
    return out


def kernel_stats(raw_csv, src):
    """Per stage (average over its launches): issue-slot utilisation, DRAM / SM throughput (% of peak), registers,
    warp instructions — the numbers that say what bounds a kernel whose HBM fraction is low by construction."""
    rows = list(csv.reader(open(raw_csv)))
    hdr, data = rows[0], rows[2:]
    kn = hdr.index("Kernel Name")
    cols = {"issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "dram_pct_of_peak": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "sm_pct_of_peak": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "registers": "launch__registers_per_thread", "warp_instructions": "smsp__inst_executed.sum",
            "duration_us_under_ncu": "gpu__time_duration.sum"}
    idx = {k: hdr.index(v) for k, v in cols.items() if v in hdr}
    per = collections.defaultdict(list)
    seen_hist = 0
    for r in data:
        name = r[kn]
        vals = {k: float(r[i].replace(",", "")) for k, i in idx.items()}
        if "radix_hist_all_kernel" in name:
            per["depth_sort.hist" if seen_hist == 0 else "tile_sort.hist"].append(vals)
            seen_hist += 1
            continue
        for pat, stage in STAGES:
            if pat in name and "radix_hist" not in pat:
                per[stage].append(vals)
                break
    out = {st: {k: round(sum(v[k] for v in lst) / len(lst), 2) for k in lst[0]} for st, lst in per.items()}
    out["_source"] = src
    return out


def main(sess, tag):
    os.makedirs("profiles", exist_ok=True)
    for f, dst in [("bench_n1.json", f"{tag}_bench_n1.json"), ("launches.csv", f"{tag}_launches.csv"),
                   ("ncu_full_summary.txt", f"{tag}_ncu_full_summary.txt"),
                   ("quick_perf.jsonl", f"{tag}_quick_perf_all_configs.jsonl")]:
        if os.path.exists(os.path.join(sess, f)):
            shutil.copy(os.path.join(sess, f), os.path.join("profiles", dst))
    raw = os.path.join(sess, "ncu_raw.csv")
    if os.path.exists(raw):
        with open(raw, "rb") as fi, gzip.open(f"profiles/{tag}_ncu_raw.csv.gz", "wb") as fo:
            shutil.copyfileobj(fi, fo)
        t = traffic(raw, f"profiles/{tag}_ncu_full_summary.txt (ncu --set full --clock-control none, lego_1m, one "
                         "forward+backward, average per launch of the kernels that run several times per step; "
                         "dram__bytes_read.sum + dram__bytes_write.sum)")
        json.dump(t, open("profiles/ncu_traffic.json", "w"), indent=1)
        json.dump(kernel_stats(raw, f"profiles/{tag}_ncu_full_summary.txt (same capture as ncu_traffic.json)"),
                  open("profiles/ncu_kernel_stats.json", "w"), indent=1)
    if os.path.exists(os.path.join(sess, "launches.csv")):
        open(f"profiles/{tag}_launch_shares.txt", "w").write(launch_shares(os.path.join(sess, "launches.csv")))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
