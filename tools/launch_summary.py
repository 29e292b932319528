""" development tool: condense `ncu --metrics gpu__time_duration.sum --csv` launch lists into per-kernel totals/shares.
    usage: python tools/launch_summary.py launches.csv "<header comment>" > profiles/rNN_launch_summary.csv """
import collections, csv, re, sys
rows = list(csv.reader(l for l in open(sys.argv[1], errors="replace") if l.startswith('"')))
hdr = rows[0]
c_name, c_val, c_unit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= c_val or r[hdr.index("Metric Name")] != "gpu__time_duration.sum":
        continue
    v = float(r[c_val].replace(",", ""))
    u = r[c_unit]
    us = v / 1000.0 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1000.0)
    name = re.sub(r"\(.*$", "", r[c_name]).strip()
    name = re.sub(r"^void ", "", name)[:160]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
n = sum(a[0] for a in agg.values())
for line in sys.argv[2:]:
    print("# " + line)
print("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes")
print(f"# total captured kernel time: {tot / 1000.0:.2f} ms over {n} launches")
print("kernel,launches,total_us,share")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"\"{k}\",{c},{t:.1f},{t / tot:.4f}")
