"""Group an ncu source-page CSV of one kernel launch into basic-block-like runs with equal execution counts."""
import csv, sys, subprocess, io
rep, skip = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "0")
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ia = hdr.index('Source'); ie = hdr.index('Instructions Executed'); it = hdr.index('Avg. Threads Executed'); iss = hdr.index('Warp Stall Sampling (All Samples)')
data = []
for r in rows[2:]:
    if len(r) <= max(ie, it, iss) or not r[ie].isdigit():
        if data: break
        continue
    data.append(r)
dump = open('/tmp/src_dump.txt', 'w')
tot = sum(int(r[ie]) for r in data); print(rows[0][1][:60], "total warp instr", tot, "n sass", len(data))
groups = []
for k, r in enumerate(data):
    e = int(r[ie]); th = float(r[it]); st = int(r[iss]); dump.write(f"{k:4d} {e:10d} {th:5.1f} {st:6d}  {r[ia].strip()[:90]}\n")
    if groups and abs(groups[-1]['e'] - e) < 0.02 * max(e, 1) and groups[-1]['end'] == k - 1:
        g = groups[-1]; g['end'] = k; g['n'] += 1; g['tot'] += e; g['st'] += st; g['th'] += th
    else:
        groups.append(dict(start=k, end=k, e=e, n=1, tot=e, st=st, th=th, first=r[ia].strip()))
tst = sum(g['st'] for g in groups)
for g in groups:
    if g['tot'] > 0.004 * tot:
        print(f"{g['start']:4d}-{g['end']:4d} n={g['n']:3d} exec={g['e']:9d} lanes={g['th']/g['n']:5.1f} share_instr={g['tot']/tot:.3f} share_stall={g['st']/tst:.3f}  {g['first'][:50]}")
