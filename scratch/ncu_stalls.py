import csv, sys, subprocess
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
cur = None; hdr = None; data = []
for r in csv.reader(raw.splitlines()):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or r[0] in ("Function Name", "Kernel Name", "File Name", ""): continue
    try:
        smp = int(r[hdr.index("# Samples")] or 0)
    except Exception:
        continue
    reasons = {}
    for i, h in enumerate(hdr):
        if h.startswith("stall_") and "Not Issued" not in h:
            try: reasons[h[6:]] = int(r[i] or 0)
            except Exception: pass
    data.append((smp, cur, r[0], r[1].strip()[:90], reasons))
ts = sum(d[0] for d in data)
agg = {}
for d in data:
    for k, v in d[4].items(): agg[k] = agg.get(k, 0) + v
print(f"kernel {kern}: {ts} samples; by reason:", ", ".join(f"{k} {100*v/ts:.0f}%" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
for smp, f, ln, src, rs in sorted(data, key=lambda d: -d[0])[:top]:
    rr = ", ".join(f"{k} {v}" for k, v in sorted(rs.items(), key=lambda x: -x[1])[:3] if v)
    print(f"{100*smp/ts:5.1f}%  {f}:{ln}  {src}   [{rr}]")
