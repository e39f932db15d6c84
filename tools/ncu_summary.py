"""Summarises an .ncu-rep (read here, no GPU needed): python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-regex-for-stalls]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor']
seen = collections.OrderedDict()
for d in data:
    name = d[idx['Kernel Name']].split('(')[0]
    seen.setdefault(name, []).append(d)
for name, ds in seen.items():
    d = ds[len(ds) // 2]
    print('----', name, f'({len(ds)} captures)')
    for w in want:
        if w in idx:
            print(f"  {w:70s} {d[idx[w]]:>16s} {units[idx[w]]}")
if len(sys.argv) > 2:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + sys.argv[2], "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]; data = [r for r in rows[2:] if len(r) == len(hdr) and r[0] != 'Address']
    idx = {h: i for i, h in enumerate(hdr)}
    f = lambda x: float(x) if x not in ('', None) else 0.0
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    seenaddr = set(); uniq = []
    for r in data:
        if r[0] in seenaddr: continue
        seenaddr.add(r[0]); uniq.append(r)
    tot = {s: sum(f(r[idx[s]]) for r in uniq) for s in stalls}
    T = sum(tot.values()) or 1
    print("stall reasons:", ", ".join(f"{s[6:]} {v / T * 100:.1f}%" for s, v in sorted(tot.items(), key=lambda kv: -kv[1])[:7]))
    print("total instr executed:", sum(f(r[idx['Instructions Executed']]) for r in uniq))
    for r in sorted(uniq, key=lambda r: -f(r[idx['# Samples']]))[:18]:
        top = max(stalls, key=lambda s: f(r[idx[s]]))
        print(r[idx['# Samples']].rjust(6), top[6:].ljust(14), r[idx['Instructions Executed']].rjust(9), r[idx['Source']][:90])
