"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table (profiles/)."""
import collections, csv, re, sys
src, dst, title = sys.argv[1], sys.argv[2], sys.argv[3]
rows = [r for r in csv.reader(open(src)) if len(r) > 5]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ki])
    v = float(r[vi].replace(",", ""))
    v = v / 1e6 if r[ui] in ("ns", "nsecond") else (v / 1e3 if r[ui] in ("us", "usecond") else (v * 1e3 if r[ui] in ("s", "second") else v))
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
ours = {k: v for k, v in agg.items() if "at::" not in k and "at_cuda" not in k}
lines = ["# " + title, "", "| kernel | launches | total ms | share |", "|---|---|---|---|"]
for k, (n, ms) in sorted(ours.items(), key=lambda x: -x[1][1]):
    lines.append("| `%s` | %d | %.2f | %.2f %% |" % (k[:90], n, ms, 100 * ms / tot))
other = sum(v[1] for k, v in agg.items() if k not in ours)
lines.append("| PyTorch helper kernels (scene set-up: depth2normal, copies) | %d | %.2f | %.2f %% |" % (
    sum(v[0] for k, v in agg.items() if k not in ours), other, 100 * other / tot))
lines += ["", "captured %d launches, %.1f ms total; per-launch times are cold-cache and serialised by the profiler: compare shares, not absolutes." % (len(rows) - 1, tot)]
open(dst, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
