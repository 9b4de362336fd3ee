"""Turn the raw captures in gpurun_out/ into the committed summaries under profiles/ (run here, CPU only; needs `ncu`)."""
import csv, io, json, os, subprocess, sys
from collections import defaultdict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

def raw_metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return dict(zip(rows[0], rows[2] if len(rows) > 2 else rows[1])), dict(zip(rows[0], rows[1]))

def source_stalls(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
    tot = defaultdict(int)
    ops = defaultdict(int)
    for r in rows[2:]:
        if len(r) != len(hdr): continue
        for c in cols: tot[c] += int(r[hdr.index(c)] or 0)
        op = r[hdr.index("Source")].replace("@P0", "").replace("@!P0", "").split()
        op = [o for o in op if not o.startswith("@")]
        if op: ops[op[0].split(".")[0]] += int(r[hdr.index("Instructions Executed")] or 0)
    return tot, ops

def mega(rep, out_name, cmd):
    v, units = raw_metrics(rep)
    keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "dram__bytes_read.sum.per_second", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "launch__shared_mem_per_block_dynamic"]
    tot, ops = source_stalls(rep)
    s = sum(tot.values())
    def to_bytes(k):
        x, u = float(v[k]), units[k].lower()
        return x * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1}.get(u, 1)
    traffic = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
    with open(os.path.join(P, out_name), "w") as f:
        f.write("ncu --set full --clock-control none, kernel k_mega_decode<64,4> (one launch = one decoded token), Qwen3-8B gs64, pos 3\n")
        f.write("command: %s\n(durations under ncu are cold-cache/serialised; the bench number is the CUDA-event one)\n\n" % cmd)
        for k in keys:
            if k in v: f.write("%-70s %s %s\n" % (k, v[k], units[k]))
        f.write("\ntraffic (dram read+write per launch) = %.4f GB ; algorithmic bytes per token at pos 3 = 8.0420 GB\n" % (traffic / 1e9))
        f.write("\nwarp stall sampling (all samples):\n")
        for c, n in sorted(tot.items(), key=lambda x: -x[1])[:9]:
            f.write("  %-24s %5.1f%%\n" % (c, 100.0 * n / s))
        f.write("\nwarp instructions executed by opcode (top 12):\n")
        for o, n in sorted(ops.items(), key=lambda x: -x[1])[:12]:
            f.write("  %-10s %8.1f M\n" % (o, n / 1e6))
    return traffic

def launches(csv_name, out_txt, out_csv, cmd):
    rows = [r for r in csv.reader(open(os.path.join(G, csv_name))) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r is hdr or len(r) != len(hdr) or r[iv] in ("Metric Value", ""): continue
        name = r[ik].split("(")[0].replace("void ", "").replace("q3::", "").replace("(int)", "")
        agg[name][0] += 1
        agg[name][1] += float(r[iv].replace(",", "")) / 1e3  # ns -> us
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(P, out_txt), "w") as f, open(os.path.join(P, out_csv), "w") as fc:
        f.write(cmd + "\n(cold-cache, serialised; includes load-time k_build_stream / k_transpose_f32, the persistent decode launches\n"
                " (timed region + e2e), the prefill pass bench.py reports, and the graph-path GEMVs it times for its comparison table)\n\n")
        f.write("%-60s %8s %12s %7s %10s\n" % ("kernel", "launches", "total_us", "share", "avg_us"))
        fc.write("kernel,launches,total_us,share,avg_us\n")
        for k, (n, us) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write("%-60s %8d %12.1f %6.1f%% %10.1f\n" % (k, n, us, 100 * us / tot, us / n))
            fc.write("%s,%d,%.1f,%.4f,%.1f\n" % (k.replace(",", ";"), n, us, us / tot, us / n))

if __name__ == "__main__":
    t = mega(os.path.join(G, "mega_8b_v2.ncu-rep"), "r01_ncu_mega_8b.txt",
             "ncu --set full --clock-control none --import-source on -k regex:k_mega -s 3 -c 1 python scripts/ncu_target.py qwen3-8b 1 6")
    json.dump({"mega": t, "workload": "qwen3-8b/gs64/tp1",
               "source": "profiles/r01_ncu_mega_8b.txt (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, one launch of k_mega_decode on Qwen3-8B gs64 at pos 3)"},
              open(os.path.join(P, "roofline_traffic.json"), "w"), indent=1)
    launches("launches2.csv", "r01_launches_bench.txt", "r01_launches_bench.csv",
             "ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv python bench.py --steps 1 --warmup 1 --tokens-per-step 16 --no-cpu-baseline")
    for src, dst in (("bench8_final2.json", "r01_bench_8b.json"), ("bench4_final2.json", "r01_bench_4b.json"), ("bench06_final2.json", "r01_bench_06b.json"),
                     ("phase_final_pos64.txt", "r01_mega_phase_profile_8b_pos64.txt"), ("phase_final_pos1500.txt", "r01_mega_phase_profile_8b_pos1500.txt")):
        open(os.path.join(P, dst), "w").write(open(os.path.join(G, src)).read())
    print("ok")
