"""Summarise the SASS source page of an ncu report: per kernel, the instruction mix, the stall reasons and the most-sampled instructions.
    python scripts/ncu_sass_summary.py report.ncu-rep [kernel substring]"""
import collections, csv, io, re, subprocess, sys
rep = sys.argv[1]
flt = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
kern = hdr = None
data = collections.OrderedDict()
idx = 0
for r in rows:
    if r and r[0] == "Kernel Name":
        idx += 1
        kern = "%d:%s" % (idx, r[1]); hdr = None; continue
    if r and r[0] == "Address":
        hdr = r; continue
    if kern and hdr and len(r) == len(hdr):
        data.setdefault(kern, []).append(dict(zip(hdr, r)))
for k, v in data.items():
    if flt not in k: continue
    tot = sum(int(x["Instructions Executed"]) for x in v); samp = sum(int(x["# Samples"]) for x in v)
    print("==", k, "static", len(v), "executed", tot, "samples", samp)
    st = collections.Counter()
    for x in v:
        for c in hdr:
            if c.startswith("stall_") and "Not Issued" not in c: st[c] += int(x[c])
    t = sum(st.values()) or 1
    print("  stalls:", ", ".join("%s %.1f%%" % (c[6:], 100 * n / t) for c, n in st.most_common(9)))
    byop = collections.Counter(); 
    for x in v:
        m = re.match(r"\s*(@!?U?P\w+\s+)?([A-Z0-9_]+)", x["Source"]); byop[m.group(2)] += int(x["Instructions Executed"])
    print("  mix:", ", ".join("%s %.1f%%" % (o, 100 * c / tot) for o, c in byop.most_common(18)))
    top = sorted(enumerate(v), key=lambda ix: -int(ix[1]["# Samples"]))[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]
    for i, x in top:
        print("   %5d %6s %-72s %s" % (i, x["# Samples"], x["Source"].strip()[:72], {c[6:]: x[c] for c in hdr if c.startswith("stall_") and "Not" not in c and int(x[c]) > max(5, int(x["# Samples"]) // 5)}))
