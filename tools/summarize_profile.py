"""Turn gpurun_out/launches.csv (ncu --metrics gpu__time_duration.sum) and *.ncu-rep into small text summaries
under profiles/ (run in the build container: ncu can read reports without a GPU)."""
import collections, csv, re, subprocess, sys, os

def launches(path, out):
    """Launch list in ncu's long CSV format (one row per launch and metric).  gpu__time_duration.sum is required;
    dram__bytes_read.sum / dram__bytes_write.sum, when collected in the same pass, give the DRAM traffic per kernel
    (also written to <out>.json -> bench.py's roofline.traffic)."""
    import json
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[hi]
    c_id, c_name, c_grid = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Grid Size")
    c_m, c_u, c_v = hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "usecond": 1e3, "nsecond": 1.0,
             "msecond": 1e6}
    per = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= c_v:
            continue
        d = per.setdefault(r[c_id], {"name": r[c_name], "grid": r[c_grid]})
        d[r[c_m]] = float(r[c_v].replace(",", "")) * scale.get(r[c_u], 1.0)
    data = list(per.values())
    names = [d["name"] for d in data]
    t = [d.get("gpu__time_duration.sum", 0.0) for d in data]
    dram = [d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0) for d in data]
    have_dram = any("dram__bytes_read.sum" in d for d in data)
    def short(n):
        m = re.search(r"(\w+_kernel)", n)
        return "torch/other" if "at::" in n or not m else m.group(1)
    sn = [short(n) for n in names]
    rend = [i for i, n in enumerate(sn) if n.startswith("render_")]
    note = ""
    if len(rend) >= 2:
        a, b = rend[-2] + 1, rend[-1] + 1
        idx = list(range(a, b))
    else:
        # the capture ended inside the second step: take the FIRST step (a warm-up step: cold caches, one-time weight
        # packing by torch kernels interleaved -- those are left out), from the first own kernel to the render kernel
        a = next(i for i, n in enumerate(sn) if n != "torch/other")
        b = rend[0] + 1
        idx = [i for i in range(a, b) if sn[i] != "torch/other"]
        note = " [FIRST (warm-up) step of a truncated capture: use for DRAM bytes, not for time shares]"
    agg, cnt, byt = collections.OrderedDict(), collections.Counter(), collections.Counter()
    for i in idx:
        agg[sn[i]] = agg.get(sn[i], 0) + t[i]; cnt[sn[i]] += 1; byt[sn[i]] += dram[i]
    tot = sum(agg.values())
    with open(out, "w") as f:
        f.write(f"# one bench step (eager launches under ncu, cold-cache serialized): {len(idx)} launches, {tot/1e6:.3f} ms{note}\n")
        f.write("# compare SHARES, not absolutes (B200_PROFILING.md)" + ("; dram = dram__bytes_read.sum + dram__bytes_write.sum per launch (average)" if have_dram else "") + "\n")
        for k, v in sorted(agg.items(), key=lambda x: -x[1]):
            extra = f"  dram {byt[k] / cnt[k] / 1e6:8.2f} MB/launch" if have_dram else ""
            f.write(f"{k:30s} n={cnt[k]:4d} {v/1e6:8.3f} ms {100*v/tot:5.1f}%{extra}\n")
        f.write("\n# conv_tc / halo launches in step order: grid, us\n")
        f.write(" ".join(f"{data[i]['grid'].replace(' ','')}:{t[i]/1e3:.1f}" for i in idx if "conv_tc" in names[i]) + "\n")
    if have_dram:
        json.dump({"what": "ncu dram__bytes_read.sum + dram__bytes_write.sum, average bytes per launch over one bench step",
                   "source": os.path.basename(out),
                   "bytes_per_launch": {k: byt[k] / cnt[k] for k in agg}, "launches": dict(cnt)},
                  open(out.replace(".txt", "_traffic.json"), "w"), indent=1)

KEYS = ["Kernel Name", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum"]

# every counter that talks about the tcgen05 pipe / tensor memory (profiles/README.md says which one to trust)
TENSOR_SUBSTR = ["pipe_tensor", "mem_tensor_reads", "mem_tensor_writes", "inst_executed_pipe_tensor", "inst_executed_pipe_tmem",
                 "data_pipe_tc_wavefronts", "inst_executed_pipe_uniform"]


def rep(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        for r in rows[2:]:
            for i, h in enumerate(hdr):
                if any(h == k or h.endswith(k) for k in KEYS) or any(t in h for t in TENSOR_SUBSTR):
                    f.write(f"{h} [{units[i]}] = {r[i]}\n")
            st = [(float(r[i]), h) for i, h in enumerate(hdr) if "pcsamp_warps_issue_stalled" in h and not h.endswith("_not_issued") and r[i]]
            tot = sum(v for v, _ in st) or 1
            f.write("stall reasons: " + ", ".join(f"{h.split('stalled_')[1]} {100*v/tot:.1f}%" for v, h in sorted(st, reverse=True)[:6]) + "\n---\n")

if __name__ == "__main__":
    tag = sys.argv[1]
    if os.path.exists("gpurun_out/launches.csv"):
        launches("gpurun_out/launches.csv", f"profiles/{tag}_launches.txt")
    for p in sys.argv[2:]:
        rep(p, f"profiles/{tag}_{os.path.basename(p).replace('.ncu-rep','')}.txt")
