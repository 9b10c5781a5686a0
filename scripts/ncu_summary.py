"""Summarise an .ncu-rep: per-kernel key metrics and top source lines by executed instructions.
usage: python scripts/ncu_summary.py report.ncu-rep [top_n]"""
import csv, collections, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 35
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'launch__grid_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active']
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')]
    print("=====", name[:60])
    for w in want:
        if w in hdr:
            print("  %-72s %s %s" % (w, r[hdr.index(w)], units[hdr.index(w)]))
    for i, h in enumerate(hdr):
        if 'warp_issue_stalled' in h and h.endswith('per_warp_active.pct'):
            try:
                v = float(r[i].replace(',', ''))
            except ValueError:
                continue
            if v > 4:
                print("  STALL %-66s %.1f" % (h.replace('smsp__warp_issue_stalled_', ''), v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
cur = None; kern = None; hdr2 = None; agg = collections.OrderedDict()
for r in csv.reader(io.StringIO(src)):
    if len(r) == 2 and r[0] == 'File Path': cur = r[1]; continue
    if len(r) == 2 and r[0] == 'Function Name': kern = r[1][:40]; continue
    if len(r) == 2: continue
    if r and r[0] == 'Line No': hdr2 = r; continue
    if hdr2 is None: continue
    try:
        ln = int(r[0]); ie = hdr2.index('Instructions Executed'); te = hdr2.index('Thread Instructions Executed')
        ss = hdr2.index('# Samples')
        inst = float(r[ie] or 0); tinst = float(r[te] or 0); smp = float(r[ss] or 0)
    except (ValueError, IndexError):
        continue
    a = agg.setdefault((kern, cur.split('/')[-1], ln, r[1].strip()[:80]), [0, 0, 0])
    a[0] += inst; a[1] += tinst; a[2] += smp
for k in sorted(set(x[0] for x in agg)):
    sub = {kk: v for kk, v in agg.items() if kk[0] == k}
    tot = sum(v[0] for v in sub.values()); tots = sum(v[2] for v in sub.values())
    if tot < 1e6: continue
    print("\n=== top lines for", k, " total warp-inst %.3g" % tot)
    for (kk, f, ln, s), (i, t, sm) in sorted(sub.items(), key=lambda kv: -kv[1][0])[:topn]:
        print("%5.2f%% inst %5.2f%% smp lanes=%4.1f %s:%d  %s" % (100 * i / tot, 100 * sm / max(tots, 1), t / max(i, 1), f, ln, s))
