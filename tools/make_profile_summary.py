"""Condense gpurun_out/ ncu artefacts into small, committed summaries under profiles/ (the .ncu-rep files stay scratch)."""
import collections
import csv
import re
import subprocess
import sys


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[start]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    seq = []
    for r in rows[start + 1:]:
        name = re.sub(r"\(.*", "", r[ki])
        v = float(r[vi].replace(",", ""))
        v = {"us": v / 1e3, "ns": v / 1e6, "ms": v}.get(r[ui], v * 1e3)
        agg.setdefault(name, []).append(v)
        seq.append((name, v))
    tot = sum(sum(v) for v in agg.values())
    with open(dst, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none : per-kernel totals (ms), then every launch in order\n")
        f.write("kernel,launches,mean_ms,total_ms,share_pct\n")
        for k, v in sorted(agg.items(), key=lambda x: -sum(x[1])):
            f.write("%s,%d,%.4f,%.3f,%.2f\n" % (k, len(v), sum(v) / len(v), sum(v), 100 * sum(v) / tot))
        f.write("\n#launch,kernel,ms\n")
        for i, (k, v) in enumerate(seq):
            f.write("%d,%s,%.4f\n" % (i, k, v))


KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed",
        "sm__sass_thread_inst_executed_op_dfma_pred_on.sum.peak_sustained",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpc__cycles_elapsed.avg.per_second", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores"]


def full(rep, dst, options):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    shdr, sdata = srows[1], srows[2:]
    ci, si = shdr.index("Instructions Executed"), shdr.index("Source")
    ops = collections.Counter()
    tot = 0
    for r in sdata:
        if len(r) <= ci or not r[ci].isdigit():
            continue
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[si].strip())
        ops[m.group(2) if m else r[si].split()[0]] += int(r[ci])
        tot += int(r[ci])
    with open(dst, "w") as f:
        f.write("# %s : ncu --set full --clock-control none, one launch over %d options\n" % (d["Kernel Name"][0] if "Kernel Name" in d else rep, options))
        f.write("metric,value,unit\n")
        for k in KEYS + sorted(k for k in hdr if "smsp__average_warps_issue_stalled" in k):
            if k in d:
                f.write("%s,%s,%s\n" % (k, d[k][0], d[k][1]))
        f.write("\n# dynamic SASS instruction mix (warp instructions), from the source page\nopcode,executed,share_pct,per_option\n")
        for op, n in ops.most_common(24):
            f.write("%s,%d,%.2f,%.1f\n" % (op, n, 100.0 * n / tot, n / options))
        f.write("TOTAL,%d,100,%.1f\n" % (tot, tot / options))


def kernel_json(rep, dst, name, options, mean_iters):
    """Merge the counters bench.py quotes for kernel `name` into profiles/kernel_profile.json, stamped with the hash of the
    CUDA sources the capture was taken from (bench.py prints "stale": true once they differ)."""
    import json
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from fastamericanoptionpricing_b200 import _lib
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--print-units", "base"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    d = dict(zip(rows[0], rows[2]))
    units = dict(zip(rows[0], rows[1]))

    def num(k):
        return float(d[k].replace(",", ""))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    shdr, sdata = srows[1], srows[2:]
    ci, si, ti = shdr.index("Instructions Executed"), shdr.index("Source"), shdr.index("Predicated-On Thread Instructions Executed")
    fp64 = other = 0
    flops = 0
    for r in sdata:
        if len(r) <= ci or not r[ci].isdigit():
            continue
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[si].strip())
        op = m.group(2) if m else ""
        if op in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"):
            fp64 += int(r[ci])
        else:
            other += int(r[ci])
        if op in ("DFMA", "DMUL", "DADD"):
            flops += int(r[ti]) * (2 if op == "DFMA" else 1)
    assert units["gpu__time_duration.sum"] in ("ns", "nsecond"), units["gpu__time_duration.sum"]
    entry = {"kernel": d.get("Kernel Name", name), "options": options, "mean_iterations": mean_iters,
             "duration_ms": num("gpu__time_duration.sum") / 1e6,
             "registers": int(num("launch__registers_per_thread")),
             "fp64_pipe_active_pct": num("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
             "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
             "flops_executed_per_option": flops / options,
             "fp64_warp_inst_per_option": fp64 / options, "other_warp_inst_per_option": other / options,
             "dram_bytes_per_option": (num("dram__bytes_read.sum") + num("dram__bytes_write.sum")) / options,
             "local_loads": num("sass__inst_executed_local_loads") if "sass__inst_executed_local_loads" in d else None,
             "capture": os.path.basename(rep)}
    try:
        prof = json.load(open(dst))
    except Exception:
        prof = {"kernels": {}}
    h = _lib.kernel_source_hash()
    if prof.get("source_hash") != h:        # a new build: counters of the old one do not carry over
        prof = {"kernels": {}}
    prof["source_hash"] = h
    prof["kernels"][name] = entry
    json.dump(prof, open(dst, "w"), indent=1, sort_keys=True)
    print(json.dumps(entry))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    elif sys.argv[1] == "json":        # json <rep> <dst.json> <kernel key> <options> <mean sweeps>
        kernel_json(sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5]), float(sys.argv[6]))
    else:
        full(sys.argv[2], sys.argv[3], int(sys.argv[4]))
