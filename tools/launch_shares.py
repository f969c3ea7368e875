"""Kernel launch shares of one pass of the path from an ncu launch list (--metrics gpu__time_duration.sum --csv):
    python tools/launch_shares.py profiles/r02_launches_split_batch2.csv > profiles/r02_launch_shares.txt
The CSV holds the warm-up pass and the measured pass of tools/ncu_path_once.py: the second half of this library's launches is used
(for the whole model: everything from the last stem conv on)."""
import csv, re, sys, collections

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]
ik, iv, iu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
ours = [(r[ik], float(r[iv].replace(",", "")) / (1000.0 if r[iu] in ("ns", "nsecond") else 1.0)) for r in rows[1:]
        if "<unnamed>" in r[ik] or "ss_" in r[ik]]
# pass boundary: the whole model starts a pass at its (single) stem conv; a multi-pass list of the path (bench.py under ncu) at the
# gwc volume kernel -- a pass from the middle of the run is used.  A two-pass list of tools/ncu_path_once.py: the second half.
stems = [i for i, (k, _) in enumerate(ours) if "stem_conv_kernel" in k]
gwcs = [i for i, (k, _) in enumerate(ours) if "gwc_volume" in k]
if stems:
    ours = ours[stems[-1]:]
elif len(gwcs) > 2:
    m = len(gwcs) // 2                         # a pass from the middle of the run (bench.py: the timed K steps)
    per, lead = gwcs[m + 1] - gwcs[m], gwcs[0]  # launches per pass; launches of a pass before its gwc kernel
    ours = ours[gwcs[m] - lead:gwcs[m] - lead + per]
else:
    ours = ours[len(ours) // 2:]
tot = sum(t for _, t in ours)
agg = collections.OrderedDict()
for k, t in ours:
    name = re.sub(r"\(.*$", "", k.replace("void ", "").replace("<unnamed>::", "")).strip()
    n, s = agg.get(name, (0, 0.0))
    agg[name] = (n + 1, s + t)
print("# kernel launch shares of ONE pass of the default-precision (split) hot path at batch 2 (1024x1024), from the ncu launch list")
print("# %s (gpu__time_duration.sum, --clock-control none; cold-cache serialised: compare SHARES)" % sys.argv[1])
print("# %d launches of this library, %.1f us total" % (len(ours), tot))
print("# launches   us   share   kernel")
for name, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%4d %9.1f  %.3f  %s" % (n, s, s / tot, name))
