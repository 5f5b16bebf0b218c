"""Per-launch table from an `ncu --csv --metrics ...` log: id, kernel, duration ns, other metrics."""
import collections, csv, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]
ki, mi, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
d = collections.OrderedDict()
for r in rows[1:]:
    d.setdefault((r[ii], r[ki].replace("void ", "").replace("mcmcb::", "")[:34].replace(" ", "")), {})[r[mi]] = r[vi].replace(",", "")
for (i, k), v in d.items():
    t = v.pop("gpu__time_duration.sum", "0")
    print(i, k, t, " ".join("%s=%s" % (m.split(".")[0].split("__")[-1][:14], x) for m, x in v.items()))
