"""Turn one gpurun round's raw captures (gpurun_out/<tag>_*) into the tracked summaries under profiles/.

    python tools/make_profile_summary.py <tag> <out_prefix>      e.g.  r01a r01

Writes profiles/<out_prefix>_launches.csv (the ncu launch list of `bench.py --steps 3 --warmup 3 --skip-e2e
--skip-cpu-baseline`, product kernels only), profiles/<out_prefix>_step_share.txt (one step's kernels with their
share of the step), profiles/<out_prefix>_ncu_step.txt / _ncu_imit.txt (key `ncu --set full` metrics per kernel)
and refreshes profiles/roofline_traffic.json (dram bytes per launch of the dominant kernels; bench.py reads it).
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, out = sys.argv[1], sys.argv[2]
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

WANT = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
    "local_load_bytes" , "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
]


def to_bytes(val, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(val.replace(",", "")) * mult.get(unit, 1)


def ncu_table(rep):
    res = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(res.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    return hdr, units, data, idx


def write_ncu_summary(rep, path, cmd):
    hdr, units, data, idx = ncu_table(rep)
    traffic = {}
    with open(path, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on, command: {cmd}\n")
        f.write(f"# source report: gpurun_out/{os.path.basename(rep)} (scratch, not tracked)\n")
        for k, r in enumerate(data):
            name = r[idx["Kernel Name"]]
            f.write(f"\n[{k}] {name[:150]}\n")
            for w in WANT:
                if w in idx:
                    f.write(f"    {w:72s} {r[idx[w]]:>16s} {units[idx[w]]}\n")
            rd = to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]])
            wr = to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]])
            f.write(f"    {'dram traffic (read+write) per launch':72s} {rd + wr:16.0f} byte\n")
            traffic.setdefault(name.split("(")[0], []).append(rd + wr)
            traffic.setdefault("__order__", []).append((name.split("(")[0], rd + wr))
    return traffic


# ---- launch list --------------------------------------------------------------------------------------------
lp = os.path.join(G, f"{tag}_launches.csv")
if os.path.exists(lp):
    rows = list(csv.reader(open(lp)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    gi, bi = hdr.index("Grid Size"), hdr.index("Block Size")
    seq = [(r[ki], r[gi], r[bi], float(r[vi])) for r in rows[start + 1:] if len(r) > vi]
    with open(os.path.join(P, f"{out}_launches.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["launch", "kernel", "grid", "block", "gpu__time_duration.sum [ns]"])
        for i, (k, g, b, v) in enumerate(seq):
            if "simstep::" in k:
                w.writerow([i, k.split("(")[0], g, b, int(v)])
    starts = [i for i, s in enumerate(seq) if "prep_input_kernel" in s[0]]   # a step begins with the input preparation
    if len(starts) >= 2:
        a, b = starts[-2], starts[-1]
        tot = sum(s[3] for s in seq[a:b])
        with open(os.path.join(P, f"{out}_step_share.txt"), "w") as f:
            f.write("# one step of bench.py (40000 env-steps, fp16 operands) under\n"
                    "# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised launches:\n"
                    "# compare SHARES with bench.py's kernels_ms_per_step, not absolutes)\n")
            for k, g, bsz, v in seq[a:b]:
                f.write(f"{k.split('(')[0][-70:]:72s} grid {g:14s} {v / 1e3:9.1f} us {100 * v / tot:5.1f}%\n")
            f.write(f"{'total':72s} {'':19s} {tot / 1e3:9.1f} us\n")

# ---- full captures ------------------------------------------------------------------------------------------
tr = {}
sp = os.path.join(G, f"{tag}_prof_step.ncu-rep")
if os.path.exists(sp):
    t = write_ncu_summary(sp, os.path.join(P, f"{out}_ncu_step.txt"),
                          "python bench.py --steps 3 --warmup 3 --skip-e2e --skip-cpu-baseline")
    # ensemble layer launches (epilogue mode 0 = hidden, 1 = final) in capture order; a step is 5 consecutive ones
    # (the capture window may start mid-step: any 5 consecutive launches cover each layer once)
    gem = [b for (k, b) in t.get("__order__", []) if "gemm_tcgen05_kernel" in k and
           any(m in k for m in (", 0>", ", 1>", ", 0,", ", 1,"))]
    tr["ensemble_gemm_dram_bytes_per_step"] = sum(gem[:5]) if len(gem) >= 5 else None
    chain = [b for (k, b) in t.get("__order__", []) if "ensemble_chain_kernel" in k]
    if chain:   # the column-fused forward: one launch per step
        tr["ensemble_gemm_dram_bytes_per_step"] = chain[0]
    for k, v in t.items():
        if k == "__order__":
            continue
        if "post_step" in k:
            tr["post_step_dram_bytes_per_launch"] = v[0]
ip = os.path.join(G, f"{tag}_prof_imit.ncu-rep")
if os.path.exists(ip):
    t = write_ncu_summary(ip, os.path.join(P, f"{out}_ncu_imit.txt"), "python tools/bench_imitation.py --iters 3 --warmup 1")
    for k, v in t.items():
        if k != "__order__" and "imitation_reward" in k:
            tr["imitation_dram_bytes_per_launch"] = v[0]
    hdr, units, data, idx = ncu_table(ip)
    key = "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"
    if key in idx and data:
        tr["imitation_fp32_pipe_frac"] = float(data[0][idx[key]].replace(",", "")) / 100.0
if tr:
    old_path = os.path.join(P, "roofline_traffic.json")
    if os.path.exists(old_path):   # keep the keys this round did not re-measure
        with open(old_path) as f:
            prev = json.load(f)
        for k, v in prev.items():
            tr.setdefault(k, v)
    tr["source"] = f"profiles/{out}_ncu_step.txt, profiles/{out}_ncu_imit.txt (ncu --set full, per launch)"
    with open(os.path.join(P, "roofline_traffic.json"), "w") as f:
        json.dump(tr, f, indent=1)
print(json.dumps(tr, indent=1))
