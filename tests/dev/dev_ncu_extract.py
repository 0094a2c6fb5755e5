"""Developer aid (not a test): turn the ncu captures under gpurun_out/ into the committed summaries under profiles/.

  python tests/dev/dev_ncu_extract.py launches gpurun_out/launches_r1.csv profiles/launches_r1_summary.txt
  python tests/dev/dev_ncu_extract.py reps profiles/ncu_summary_r1.md gpurun_out/prof_a.ncu-rep [gpurun_out/prof_b.ncu-rep ...]
The second form also writes profiles/ncu_decode_traffic.json (DRAM bytes per launch of the decode kernels, read by
bench.py for roofline.traffic)."""
import csv, io, json, os, subprocess, sys
from collections import OrderedDict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__cluster_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]


def launches(src, dst):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] in ("ns", "nsecond") else v * (1e3 if r[ui] in ("ms", "msecond") else 1.0)
        name = r[ki].split("(")[0][:64]
        n, t = agg.get(name, (0, 0.0))
        agg[name] = (n + 1, t + v)
    tot = sum(t for _, t in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list of the device-timed region of `bench.py --steps 1 --warmup 3 --profile-region --no-e2e --no-cpu-baseline`\n"
                "# ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off ; cold-cache + serialised: compare SHARES\n")
        f.write(f"{'kernel':64s} {'launches':>8s} {'total_us':>11s} {'share':>7s}\n")
        for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{name:64s} {n:8d} {t:11.1f} {100 * t / tot:6.2f}%\n")
    print(open(dst).read())


def reps(dst, files):
    out = [f"# ncu summaries ({os.path.basename(dst)})\n",
           "Captured on a B200 with `ncu --set full --clock-control none --import-source on -k regex:<kernel>` (the .ncu-rep files stay in",
           "gpurun_out/, which is scratch); extracted with `ncu -i <rep> --page raw --csv`.  Durations under ncu are NOT bench numbers.\n"]
    traffic = {}
    for fpath in files:
        txt = subprocess.run(["ncu", "-i", fpath, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        hdr, units = rows[0], rows[1]
        out.append(f"## {os.path.basename(fpath)}")
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            out.append(f"- {name[:110]}")
            vals = {}
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    vals[k] = (r[i], units[i])
                    out.append(f"    {k} = {r[i]} {units[i]}")
            def gb(k):
                v, u = vals.get(k, ("0", "byte"))
                v = float(v.replace(",", ""))
                return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}.get(u, 1.0)
            key = None
            if "whisper_decode" in name:
                key = "whisper_decode_cluster" if "cluster" in name else "whisper_decode"
            elif "llama_decode_kernel" in name:
                key = "llama_decode"
            elif "conv1d_tc" in name:
                key = "conv1d_tc"
            if key and key not in traffic:
                traffic[key] = {"kernel": name.split("(")[0], "dram_bytes_per_launch": gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum"),
                                "duration_under_ncu": vals.get("gpu__time_duration.sum"), "source": os.path.basename(fpath)}
        out.append("")
    open(dst, "w").write("\n".join(out) + "\n")
    if traffic:
        json.dump(traffic, open(os.path.join(os.path.dirname(dst), "ncu_traffic_r2.json"), "w"), indent=1)
    print("\n".join(out))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        reps(sys.argv[2], sys.argv[3:])
