"""Top stall sites of a kernel from `ncu --page source --print-source sass --csv` output."""
import csv, subprocess, sys
rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if 'Source' in r and any('Sampl' in c for c in r)]
h = rows[hi[0]]; ci = {c: i for i, c in enumerate(h)}
end = hi[1] if len(hi) > 1 else len(rows)
body = [r for r in rows[hi[0] + 1:end] if len(r) == len(h)]
tot = sum(float(r[ci['# Samples']] or 0) for r in body)
stalls = [c for c in h if c.startswith('stall_') and 'Not Issued' not in c]
print("kernel 1 of %d, samples %d" % (len(hi), tot))
agg = {s: sum(float(r[ci[s]] or 0) for r in body) for s in stalls}
print("stall mix:", ", ".join("%s %.0f%%" % (k[6:], 100 * v / max(1, sum(agg.values()))) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:6]))
for n, r in sorted(((float(r[ci['# Samples']] or 0), r) for r in body), key=lambda t: -t[0])[:top]:
    best = max(stalls, key=lambda s: float(r[ci[s]] or 0))
    print("%6.0f %5.1f%%  %-14s %s" % (n, 100 * n / max(1, tot), best[6:], r[ci['Source']][:100]))
