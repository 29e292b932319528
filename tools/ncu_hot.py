""" development tool: top stall locations of one kernel launch from an .ncu-rep (needs -lineinfo + --import-source on)
    usage: python tools/ncu_hot.py rep.ncu-rep <kernel regex> [launch-skip] [top-n] """
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat, "--launch-skip", skip,
                      "--launch-count", "1", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
samp = col["# Samples"]
data = [r for r in rows[hi + 1:] if len(r) > samp and r[samp].strip().isdigit()]
tot = sum(int(r[samp]) for r in data)
print(rows[0][1][:100], " total samples", tot)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
order = sorted(range(len(data)), key=lambda i: -int(data[i][samp]))[:topn]
for i in sorted(order):
    r = data[i]
    st = sorted(((int(r[col[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:2]
    print(f"{int(r[samp]):6d} {100.0 * int(r[samp]) / tot:5.1f}%  [{i:5d}] {r[col['Source']][:90]:90s} {st}")
