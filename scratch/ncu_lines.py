import csv, sys, subprocess
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
cur = None
data = []
ie_col = smp_col = None
for r in csv.reader(raw.splitlines()):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Line No":
        ie_col = r.index("Instructions Executed"); smp_col = r.index("# Samples"); continue
    if r[0] in ("Function Name", "Kernel Name", "File Name") or ie_col is None: continue
    if r[0] == "": continue
    try:
        data.append((float(r[ie_col] or 0), int(r[smp_col] or 0), cur, r[0], r[1].strip()[:100]))
    except Exception:
        pass
tot = sum(d[0] for d in data); ts = sum(d[1] for d in data)
print(f"kernel {kern}: {tot/1e6:.1f} M warp-instr, {ts} samples")
for ie, smp, f, ln, src in sorted(data, reverse=True)[:top]:
    print(f"{100*ie/tot:5.1f}% instr {100*smp/max(ts,1):5.1f}% smp  {f}:{ln}  {src}")
