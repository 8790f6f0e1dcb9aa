import csv, sys, subprocess
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
f = {'Gbyte': 1, 'Mbyte': 1e-3, 'Kbyte': 1e-6, 'byte': 1e-9}
keys = [h for h in hdr if h.startswith('smsp__pcsamp_warps_issue_stalled') and 'not_issued' not in h]
def g(r, k):
    return r[idx[k]] if k in idx else ''
print("| kernel | time ms | DRAM rd GB | DRAM wr GB | DRAM % | L2 hit % | SM thr % | regs | occ % | warp-instr M | top stalls |")
print("|---|---|---|---|---|---|---|---|---|---|---|")
for r in rows[2:]:
    name = r[idx['Kernel Name']].split('(')[0].replace('void ', '')
    rd = float(g(r, 'dram__bytes_read.sum')) * f[units[idx['dram__bytes_read.sum']]]
    wr = float(g(r, 'dram__bytes_write.sum')) * f[units[idx['dram__bytes_write.sum']]]
    tu = units[idx['gpu__time_duration.sum']]
    t = float(g(r, 'gpu__time_duration.sum')) * {'ms': 1, 'us': 1e-3, 'ns': 1e-6, 's': 1e3}[tu]
    st = sorted(((float(r[idx[k]] or 0), k.replace('smsp__pcsamp_warps_issue_stalled_', '')) for k in keys), reverse=True)[:4]
    tot = sum(float(r[idx[k]] or 0) for k in keys) or 1
    print(f"| {name} | {t:.3f} | {rd:.3f} | {wr:.3f} | {float(g(r,'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')):.1f} | {float(g(r,'lts__t_sector_hit_rate.pct')):.1f} | {float(g(r,'sm__throughput.avg.pct_of_peak_sustained_elapsed')):.1f} | {g(r,'launch__registers_per_thread')} | {float(g(r,'sm__warps_active.avg.pct_of_peak_sustained_active')):.0f} | {float(g(r,'smsp__inst_executed.sum'))/1e6:.0f} | " + ", ".join(f"{k} {100*v/tot:.0f}%" for v, k in st) + " |")
