"""Turn the round-2 captures in gpurun_out/ (scripts/evidence_run.sh) into the committed summaries under profiles/ (CPU only)."""
import csv, gzip, io, json, os, re, subprocess, sys
from collections import defaultdict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
LIB = os.path.join(ROOT, "qwen3_rs_b200", "lib", "libqwen3cuda.so")


def to_bytes(x, u):
    return float(x) * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1}.get(u.lower(), 1)


def source_stalls(path):
    rows = list(csv.reader(io.TextIOWrapper(gzip.open(path))))
    hdr = rows[1] if "Source" in rows[1] else rows[0]
    body = rows[rows.index(hdr) + 1:]
    cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
    tot, ops = defaultdict(int), defaultdict(int)
    for r in body:
        if len(r) != len(hdr):
            continue
        for c in cols:
            tot[c] += int(r[hdr.index(c)] or 0)
        op = [o for o in r[hdr.index("Source")].split() if not o.startswith("@")]
        if op:
            ops[op[0].split(".")[0]] += int(r[hdr.index("Instructions Executed")] or 0)
    return tot, ops


def mega():
    rows = list(csv.reader(open(os.path.join(G, "r2_ncu_mega_raw.csv"))))
    hdr, units, v = rows[0], dict(zip(rows[0], rows[1])), dict(zip(rows[0], rows[2]))
    keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "dram__bytes_read.sum.per_second", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "launch__shared_mem_per_block_dynamic"]
    tot, ops = source_stalls(os.path.join(G, "r2_ncu_mega_source.csv.gz"))
    s = sum(tot.values())
    traffic = to_bytes(v["dram__bytes_read.sum"], units["dram__bytes_read.sum"]) + to_bytes(v["dram__bytes_write.sum"], units["dram__bytes_write.sum"])
    alg = 8.041103360e9 + 294912.0 * 516
    with open(os.path.join(P, "r02_ncu_mega_8b.txt"), "w") as f:
        f.write("ncu --set full --clock-control none, kernel k_mega_decode<64,4> (one launch = one decoded token), Qwen3-8B gs64, position 515\n")
        f.write("command: ncu --set full --clock-control none --import-source on -k regex:k_mega -s 3 -c 1 python scripts/ncu_target.py qwen3-8b 1 6 512\n"
                "(durations under ncu are cold-cache/serialised; the bench number is the CUDA-event one; pages exported on the box with\n"
                " `ncu -i ... --page raw --csv` / `--page source --csv`, scripts/evidence_run.sh)\n\n")
        for k in keys:
            if k in v:
                f.write("%-70s %s %s\n" % (k, v[k], units[k]))
        f.write("\ntraffic (dram read+write per launch) = %.4f GB ; algorithmic bytes per token at pos 515 = %.4f GB ; ratio %.4f\n" % (traffic / 1e9, alg / 1e9, traffic / alg))
        f.write("\nwarp stall sampling (all samples):\n")
        for c, n in sorted(tot.items(), key=lambda x: -x[1])[:10]:
            f.write("  %-24s %5.1f%%\n" % (c, 100.0 * n / s))
        f.write("\nwarp instructions executed by opcode (top 14):\n")
        for o, n in sorted(ops.items(), key=lambda x: -x[1])[:14]:
            f.write("  %-10s %8.1f M\n" % (o, n / 1e6))
    json.dump({"mega": traffic, "workload": "qwen3-8b/gs64/tp1", "position": 515, "algorithmic_bytes_at_that_position": alg,
               "source": "profiles/r02_ncu_mega_8b.txt (ncu --set full on the round-2 kernel, dram__bytes_read.sum + dram__bytes_write.sum, one launch of "
                         "k_mega_decode on Qwen3-8B gs64 at position 515; the bench's own launches run at other positions: +294 912 B per position)"},
              open(os.path.join(P, "roofline_traffic.json"), "w"), indent=1)


def prefill():
    rows = list(csv.reader(open(os.path.join(G, "r2_ncu_prefill_raw.csv"))))
    hdr, units = rows[0], dict(zip(rows[0], rows[1]))
    keys = ["gpu__time_duration.sum", "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "smsp__inst_executed.sum"]
    with open(os.path.join(P, "r02_ncu_prefill_4b.txt"), "w") as f:
        f.write("ncu --set full --clock-control none, Qwen3-4B gs64, T = 2048 prefill: k_gemm_q8<64,EPI,0> (tcgen05.mma.kind::i8 + TMA + TMEM, persistent,\n"
                "fast drain; EPI 3 = gate/up + SwiGLU, 2 = o_proj / down + residual, 1 = QKV) and k_pf_attention_h<4> (mma.sync m16n8k16, FP16 hi / lo split)\n"
                "command: ncu --set full --clock-control none --import-source on -k regex:'k_gemm_q8|k_pf_attention_h' -s 8 -c 5 python scripts/ncu_prefill_target.py qwen3-4b 2048\n"
                "(durations under ncu are cold-cache / serialised; the CUDA-event numbers are in r02_gemm_q8_ceilings.txt and the bench line)\n\n")
        for r in rows[2:]:
            v = dict(zip(hdr, r))
            f.write(v["Kernel Name"] + "\n")
            for k in keys:
                if k in v:
                    f.write("  %-78s %s %s\n" % (k, v[k], units[k]))
            st = {k: float(x.replace(",", "")) for k, x in v.items() if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and x}
            f.write("  warps stalled per issue slot: " + ", ".join("%s %.2f" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), x)
                                                                  for k, x in sorted(st.items(), key=lambda kv: -kv[1])[:6]) + "\n\n")
        f.write("Reading (GEMM): the int8 tensor pipe is busy 15-18 % of the time (round 1: 8.6 %, start of round 2: 12 %); the issue slots are 67-71 % busy and the\n"
                "leading stall is 'not selected' -- the epilogue warps (I2FP + FMUL2 + FFMA2 per accumulator element and quantisation group) now fill the\n"
                "schedulers, i.e. the kernel runs at ~70 % of its own instruction-issue ceiling.  The TMEM read path is NOT the limit (an earlier note in this\n"
                "file said so): scripts/micro/tmem_ld_bw.cu measures 387 B/clk/SM for tcgen05.ld.32x32b.x16 with two loads in flight per warp, i.e. 170 clk for\n"
                "the 64 KB of one 128 x 128 group accumulator, against ~130 clk of MMA and ~370 issue slots per scheduler for scaling it (r02_tmem_ld_bw.txt).\n"
                "What held the kernel at 12 % was found with the timing-experiment modes and the in-kernel trace (r02_gemm_q8_ceilings.txt): scale rows shipped as\n"
                "eight 512-byte bulk copies per stage from a divergent single-lane producer, the MMA issuer starved in warp 1, uncoalesced output stores.\n"
                "Reading (attention): 56 % of the legacy tensor pipe with 8 warps per SM (218 registers, 96 KB of shared memory per CTA); stall 'wait' (fixed-latency\n"
                "dependencies of back-to-back MMAs / ldmatrix) leads.  3.0x faster than the 3xTF32 kernel it replaces (1014 -> 370 us per 4B layer at T = 2048).\n")


def launches():
    rows = [r for r in csv.reader(open(os.path.join(G, "r2_launches.csv"), errors="ignore")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r is hdr or len(r) != len(hdr) or r[iv] in ("Metric Value", ""):
            continue
        name = r[ik].split("(")[0].replace("void ", "").replace("q3::", "").replace("(int)", "")
        agg[name][0] += 1
        agg[name][1] += float(r[iv].replace(",", "")) / 1e3  # ns -> us
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(P, "r02_launches_bench.txt"), "w") as f, open(os.path.join(P, "r02_launches_bench.csv"), "w") as fc:
        f.write("ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv python bench.py --steps 1 --warmup 3 --tokens-per-step 16 --no-cpu-baseline --no-extras\n"
                "(cold-cache, serialised; includes load-time k_build_stream / k_transpose_f32, the persistent decode launches (timed region + e2e),\n"
                " the prefill pass bench.py reports, and the graph-path GEMVs it times for its comparison table)\n\n")
        f.write("%-60s %8s %12s %7s %10s\n" % ("kernel", "launches", "total_us", "share", "avg_us"))
        fc.write("kernel,launches,total_us,share,avg_us\n")
        for k, (n, us) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write("%-60s %8d %12.1f %6.1f%% %10.1f\n" % (k, n, us, 100 * us / tot, us / n))
            fc.write("%s,%d,%.1f,%.4f,%.1f\n" % (k.replace(",", ";"), n, us, us / tot, us / n))


def sass():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    fns = re.split(r"(?=\tFunction : )", txt)
    want = {"r02_sass_k_mega_decode.txt": "k_mega_decodeILi64ELi4E", "r02_sass_k_gemm_q8.txt": "k_gemm_q8ILi64ELi3ELi0E", "r02_sass_k_pf_attention_h.txt": "k_pf_attention_hILi4E"}
    for out, key in want.items():
        body = next(f for f in fns if key in f.split("\n")[0])
        lines = [l for l in body.split("\n") if not re.match(r"^\s+/\* 0x[0-9a-f]{16} \*/\s*$", l)]
        lines = [re.sub(r"\s+/\* 0x[0-9a-f]{16} \*/\s*$", "", l) for l in lines]
        ops = defaultdict(int)
        for l in lines:
            m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", l)
            if m:
                ops[m.group(1)] += 1
        with open(os.path.join(P, out), "w") as f:
            f.write("cuobjdump -sass qwen3_rs_b200/lib/libqwen3cuda.so, function %s (sm_100a; encodings stripped)\n" % body.split("\n")[0].strip())
            f.write("mnemonic counts: " + ", ".join("%s %d" % kv for kv in sorted(ops.items(), key=lambda x: -x[1])[:40]) + "\n")
            f.write("Blackwell-specific: " + ", ".join("%s %d" % (k, ops.get(k, 0)) for k in ("UTCIMMA", "UTCHMMA", "LDTM", "UTMALDG", "UBLKCP", "UTCBAR", "SYNCS", "IDP", "HMMA", "LDSM", "FFMA2", "FMUL2")) + "\n\n")
            f.write("\n".join(lines))


def copies():
    for src, dst in (("r2_phase_pos900_final.txt", "r02_mega_phase_profile_8b_pos900.txt"), ("r2_phase_pos64_final.txt", "r02_mega_phase_profile_8b_pos64.txt"),
                     ("r2_sanitizer_memcheck.log", "r02_sanitizer_memcheck.log"), ("r2_sanitizer_synccheck.log", "r02_sanitizer_synccheck.log"),
                     ("r2_flips_06b.txt", "r02_flip_attribution_06b.txt")):
        if os.path.exists(os.path.join(G, src)):
            open(os.path.join(P, dst), "w").write(open(os.path.join(G, src)).read())
    race = open(os.path.join(G, "r2_sanitizer_racecheck.log")).read()
    head = ("compute-sanitizer --tool racecheck --racecheck-report all python scripts/sanitize_target.py micro   (first 60 KB of the log + its summary)\n"
            "Every hazard reported is of ONE class: a consumer warp's 128-bit shared-memory read of a ring stage (lds128, q3_mega.cuh) against the producer\n"
            "thread's cp.async.bulk / TMA refill of the same stage (bulk_g2s / tma_load_2d).  These are ordered by the stage's `empty` mbarrier (consumer:\n"
            "__syncwarp + mbarrier.arrive after its last read; producer: mbarrier.try_wait before issuing the copy) -- the standard TMA pipeline hand-off,\n"
            "which racecheck does not model for asynchronous-proxy writes.  memcheck and synccheck are clean (r02_sanitizer_memcheck.log, _synccheck.log).\n\n")
    kinds = defaultdict(int)
    for m in re.finditer(r"Read Thread \([^)]*\) at (.+?)\+0x[0-9a-f]+ in \S+\n=========     Write Thread \([^)]*\) at (.+?)\+0x", race):
        kinds[(m.group(1).split("(")[0], m.group(2).split("(")[0])] += 1
    head += "hazard classes in this excerpt (reader function, writer function): " + "; ".join("%s / %s x %d" % (a, b, n) for (a, b), n in kinds.items()) + "\n\n"
    open(os.path.join(P, "r02_sanitizer_racecheck.log"), "w").write(head + race[:40000] + "\n...\n" + race[-400:])


if __name__ == "__main__":
    which = sys.argv[1:] or ["mega", "prefill", "launches", "sass", "copies"]
    for w in which:
        globals()[w]()
    print("ok", which)
