#!/usr/bin/env python
"""Summarise an .ncu-rep into the CSVs committed under profiles/ (metrics + stall/opcode mix).
usage: tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/prefix "note" """
import csv, subprocess, sys, collections, io
rep, prefix, note = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ['Kernel Name', 'gpu__time_duration.sum', 'sm__cycles_elapsed.avg', 'sm__cycles_active.avg',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'launch__shared_mem_per_block_static']
with open(prefix + "_ncu_summary.csv", "w") as f:
    f.write("# " + note + "\nmetric,unit,value\n")
    for i, h in enumerate(hdr):
        if h in keep:
            f.write('"%s",%s,"%s"\n' % (h, units[i], vals[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
body = [r for r in rows[2:] if len(r) >= len(h)]
ix = {k: i for i, k in enumerate(h)}
stalls = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
tot = sum(int(r[ix["# Samples"]]) for r in body)
inst = sum(int(r[ix["Instructions Executed"]]) for r in body)
st = {k: sum(int(r[ix[k]]) for r in body) for k in stalls}
ops, ops_s = collections.Counter(), collections.Counter()
for r in body:
    t = r[ix["Source"]].split()
    op = t[1] if t[0].startswith("@") else t[0]
    ops[op] += int(r[ix["Instructions Executed"]])
    ops_s[op] += int(r[ix["# Samples"]])
ex = [i for i, r in enumerate(body) if "MUFU.EX2" in r[ix["Source"]]]
with open(prefix + "_ncu_stalls.csv", "w") as f:
    f.write("# " + note + "\n# warp-state samples over the whole kernel (source page)\nstall,samples,pct\n")
    for k, v in sorted(st.items(), key=lambda kv: -kv[1]):
        f.write("%s,%d,%.2f\n" % (k[6:], v, 100.0 * v / tot))
    f.write("# executed warp instructions by opcode (top 25) and the share of samples on them\nopcode,executed,pct_executed,pct_samples\n")
    for op, v in ops.most_common(25):
        f.write("%s,%d,%.2f,%.2f\n" % (op, v, 100.0 * v / inst, 100.0 * ops_s[op] / tot))
    f.write("# instructions per MUFU.EX2 over the whole kernel: %.3f\n" % (inst / ops["MUFU.EX2"]))
print(open(prefix + "_ncu_summary.csv").read())
print(open(prefix + "_ncu_stalls.csv").read())
