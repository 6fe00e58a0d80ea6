"""Turn the ncu outputs under gpurun_out/ into the committed summaries under profiles/."""
import csv, json, re, subprocess, sys, collections, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"

def launch_list(src=None, dst=None, title="one train step (forward + CE + backward), cfg2 crossattention, B=4096, fp32-strict"):
    with open(os.path.join(G, src or f"launches_{tag}.csv")) as f:
        lines = [l for l in f if l.startswith('"')]
    r = csv.reader(lines); hdr = next(r)
    ki, vi, ui, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Grid Size")
    rows = list(r)
    marks = [i for i, row in enumerate(rows) if "rng_advance" in row[ki]]
    s, e = marks[0] + 1, marks[1] + 1
    out, agg, tot = [], collections.OrderedDict(), 0.0
    for row in rows[s:e]:
        v = float(row[vi].replace(",", "")); v = v / 1000 if row[ui] == "ns" else v
        name = re.sub(r"\(.*", "", row[ki]).replace("fb200::", "").replace("void ", "")
        out.append(f"{v:9.1f} us  grid {row[gi]:>14s}  {name}")
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
    with open(os.path.join(P, dst or f"{tag}_launch_list_cfg2_B4096.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none; {title}\n")
        f.write(f"# cold-cache, serialised durations: compare SHARES, not absolutes.  step total {tot:.1f} us over {e - s} launches\n")
        f.write("# --- per kernel ---\n")
        for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{t:9.1f} us  {100 * t / tot:5.1f}%  x{c:3d}  {n}\n")
        f.write("# --- launch order ---\n" + "\n".join(out) + "\n")
    return tot, agg

def full():
    rep = os.path.join(G, f"prof_{tag}_tc.ncu-rep")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines())); hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    want = ["Kernel Name", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sectors_srcunit_tex_op_read.sum"]
    recs = []
    for r in rows[2:]:
        recs.append({w: (r[idx[w]] + " " + units[idx[w]]).strip() for w in want if w in idx})
    json.dump(recs, open(os.path.join(P, f"{tag}_tc_gemm_ncu_full_summary.json"), "w"), indent=1)
    def num(x): return float(x.split()[0].replace(",", ""))
    def to_bytes(x):
        v, u = x.split()[0], x.split()[1] if len(x.split()) > 1 else "byte"
        return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    first = [r for r in recs if "tc_gemm_kernel" in r["Kernel Name"]][0]
    tr = to_bytes(first["dram__bytes_read.sum"]) + to_bytes(first["dram__bytes_write.sum"])
    json.dump({"kernel": first["Kernel Name"][:80], "grid": first["Grid Size"], "dram_bytes_per_launch": tr,
               "note": "dram__bytes_read.sum + dram__bytes_write.sum of one launch (ncu --set full, cold L2)"},
              open(os.path.join(P, f"{tag}_dominant_kernel_traffic.json"), "w"), indent=1)
    return recs

def attention():
    rep = os.path.join(G, f"prof_{tag}_attn.ncu-rep")
    if not os.path.exists(rep):
        return []
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines())); hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
            "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
    recs = [{w: (r[idx[w]] + " " + units[idx[w]]).strip() for w in want if w in idx} for r in rows[2:]]
    json.dump({"command": "ncu --set full --clock-control none --import-source on -k regex:attn_ -c 3 python tools/attn_once.py",
               "shape": "Sq=197 image tokens, Skv=85 metadata tokens, B=32, D=512, H=8 (hd=64), fp32", "kernels": recs},
              open(os.path.join(P, f"{tag}_attention_ncu_summary.json"), "w"), indent=1)
    return recs


def rep_summary(rep, out_name, command, shape, extra_metrics=()):
    """ncu --set full report -> JSON of the per-kernel counters that explain it (tensor / FMA pipe, issue slots, stalls, DRAM)."""
    if not os.path.exists(rep):
        print("missing", rep); return []
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines())); hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
            "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"] + list(extra_metrics)
    recs = [{w: (r[idx[w]] + " " + units[idx[w]]).strip() for w in want if w in idx} for r in rows[2:]]
    json.dump({"command": command, "shape": shape, "kernels": recs}, open(os.path.join(P, out_name), "w"), indent=1)
    return recs


def r02c():
    a = rep_summary(os.path.join(G, "r02c_attn_tc.ncu-rep"), "r02c_attention_tc_ncu_summary.json",
                    "ncu --set full --clock-control none --import-source on -k regex:attn_tc -s 3 -c 3 python tools/attn_once.py",
                    "Sq=197 image tokens, Skv=85 metadata tokens, B=32, D=512, H=8 (hd=64), fp32 via 3xTF32 mma.sync")
    t = rep_summary(os.path.join(G, "r02c_tabt.ncu-rep"), "r02c_tabt_ncu_summary.json",
                    "ncu --set full --clock-control none --import-source on -k regex:tabt_ -s 3 -c 3 python tools/tabt_once.py 4096",
                    "TabTransformer encoder, B=4096 samples x 82 tokens x d=32, 4 heads, ff=128, 2 layers, train mode (Philox dropout), fp32")
    for r in a + t:
        print({k: v for k, v in r.items() if k in ("Kernel Name", "gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
                                                   "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")})
    for src, dst in (("r02c_bench_head.json", "r02c_bench_head.json"), ("r02c_tabt_bench.log", "r02c_tabt_bench.txt"),
                     ("r02c_attn_bench.log", "r02c_attn_bench.txt"), ("r02c_gputest.log", "r02c_gputest.txt")):
        if os.path.exists(os.path.join(G, src)):
            open(os.path.join(P, dst), "w").write(open(os.path.join(G, src)).read())


def r02d():
    """fourth session of round 2: cluster split-K GEMM (medium batches), grb_bwd with two CTAs per SM, bf16 plain epilogue"""
    for src, dst, title in (("launches_r02d_b256.csv", "r02d_launch_list_cfg2_B256.txt", "one train step, cfg2 crossattention, B=256, fp32-strict (cluster split-K GEMMs)"),
                            ("launches_r02d_cfg5.csv", "r02d_launch_list_cfg5_bf16.txt", "one train step, cfg5 RG-ATT, B=4096, bf16")):
        if os.path.exists(os.path.join(G, src)):
            tot, agg = launch_list(src, dst, title); print(dst, "total", tot)
    rep_summary(os.path.join(G, "r02d_tc_csplit.ncu-rep"), "r02d_tc_gemm_csplit_ncu_summary.json",
                "ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 30 -c 6 python bench.py --batch 256 --steps 2 --warmup 3 --no-graph ...",
                "cfg2 crossattention B=256, fp32-strict 3xTF32, 128x64 tiles, cluster split-K (1,1,4)",
                extra_metrics=("launch__cluster_dim_z", "launch__cluster_size", "sm__ctas_launched.sum"))
    g = rep_summary(os.path.join(G, "r02d_grb.ncu-rep"), "r02d_grb_ncu_summary.json",
                    "ncu --set full --clock-control none -k regex:grb_ -s 4 -c 4 python bench.py --workload cfg5 --steps 2 --warmup 3 --no-graph ...",
                    "cfg5 RG-ATT B=4096 bf16: grb_fwd / grb_bwd rows 512 wide, fp32 element-wise")
    for r in g:
        print({k: v for k, v in r.items() if k in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread")})
    for src in ("r02d_bench_head.json", "r02d_bench_cfg3a.json", "r02d_bench_cfg4b.json", "r02d_bench_cfg5.json", "r02d_gputest.log", "r02d_csplit_ab.txt"):
        if os.path.exists(os.path.join(G, src)):
            open(os.path.join(P, src.replace(".log", ".txt")), "w").write(open(os.path.join(G, src)).read())


def r02e():
    """fifth session of round 2: grid-wide GEMM timeline, tensor memory released early (bf16 / cluster kernels), bf16 with two CTAs
    per SM, finer split of the deferred FFMA weight gradient"""
    global tag
    for src, dst, title in (("launches_r02e.csv", "r02e_launch_list_cfg2_B4096.txt", "one train step (forward + CE + backward), cfg2 crossattention, B=4096, fp32-strict"),
                            ("launches_r02e_cfg5.csv", "r02e_launch_list_cfg5_bf16.txt", "one train step, cfg5 RG-ATT, B=4096, bf16 (two GEMM CTAs per SM)")):
        if os.path.exists(os.path.join(G, src)):
            tot, agg = launch_list(src, dst, title); print(dst, "total", tot)
    if os.path.exists(os.path.join(G, "prof_r02e_tc.ncu-rep")):
        for r in full(): print({k: v for k, v in r.items() if k in ("Kernel Name", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed")})
    for src in ("r02e_bench_head.json", "r02e_bench_cfg3a.json", "r02e_bench_cfg4b.json", "r02e_bench_cfg5.json", "r02e_bench_2gpu.json", "r02e_gputest.log", "r02e_smoke.log",
                "r02e_timeline_cfg2_B4096.txt", "r02e_timeline_cfg5_B4096.txt", "r02e_timeline_cfg2_B256.txt", "r02e_handover_gemm.txt", "r02e_handover_micro.txt",
                "r02e_occ2.txt", "r02e_iso.txt", "r02e_early.txt", "r02e_smalln.txt", "r02e_kloop_dbg.txt"):
        if os.path.exists(os.path.join(G, src)):
            open(os.path.join(P, src.replace(".log", ".txt")), "w").write(open(os.path.join(G, src)).read())


if __name__ == "__main__":
    if tag == "r02e":
        r02e(); sys.exit(0)
    if tag == "r02c":
        r02c(); sys.exit(0)
    if tag == "r02d":
        r02d(); sys.exit(0)
    tot, agg = launch_list(); print("launch list total", tot)
    for r in attention(): print({k: v for k, v in r.items() if k in ("Kernel Name", "gpu__time_duration.sum")})
    for r in full(): print({k: v for k, v in r.items() if k in ("Kernel Name", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed")})
