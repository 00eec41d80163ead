#!/usr/bin/env python3
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per source line and per kernel phase.
usage: ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > cs.csv ; python tools/ncu_lines.py cs.csv [top]"""
import csv, sys, re, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hdr = None; cur = None; curfile = ""
per = collections.OrderedDict()
for r in rows:
    if len(r) == 2 and r[0] == "File Path": curfile = r[1].split("/")[-1]; continue
    if len(r) > 5 and r[0] == "Line No": hdr = r; iI = hdr.index("Instructions Executed"); iS = hdr.index("# Samples"); iW = hdr.index("L1 Wavefronts Shared"); iX = hdr.index("L1 Wavefronts Shared Excessive"); iST = [(h2, hdr.index(h2)) for h2 in hdr if h2.startswith("stall_") and "Not Issued" not in h2]; continue
    if hdr is None or len(r) < len(hdr): continue
    if r[0] != "":
        cur = (curfile, int(r[0]), r[1].strip()); per.setdefault(cur, [0, 0, 0, collections.Counter(), 0, collections.Counter()]); continue
    if cur is None: continue
    try: ins = int(r[iI]); smp = int(r[iS]); wv = int(r[iW]); xs = int(r[iX])
    except ValueError: continue
    a = per[cur]; a[0] += ins; a[1] += smp; a[2] += wv; a[4] += xs
    for nm, ix in iST:
        try: a[5][nm] += int(r[ix])
        except ValueError: pass
    op = r[3].split()[0] if not r[3].strip().startswith("@") else r[3].split()[1]
    a[3][op.split(".")[0]] += ins
totI = sum(v[0] for v in per.values()); totS = sum(v[1] for v in per.values()); totW = sum(v[2] for v in per.values())
print("total warp instructions %d, samples %d, shared wavefronts %d" % (totI, totS, totW))
# phases by marker comments in the kernel source ("// ---- Pn")
phase = collections.OrderedDict(); name = "prologue"
byfile = [k for k in per if k[0].startswith("hfx_assemble")]
marks = {}
import os
srcp = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "hyperfox_b200", "csrc", "hfx_assemble.cuh")
if len(sys.argv) > 3: srcp = sys.argv[3]
for ln, text in enumerate(open(srcp), 1):
    m = re.match(r"\s*// -+ (P\w+)", text)
    if m: marks[ln] = m.group(1)
    if "once per CTA" in text: marks[ln] = "prologue"
import bisect
mk = sorted(marks)
def ph(line):
    i = bisect.bisect_right(mk, line) - 1
    return marks[mk[i]] if i >= 0 else "prologue/helpers"
agg = collections.OrderedDict()
for k, v in per.items():
    p = ph(k[1]) if k[0].startswith("hfx_assemble") and k[1] >= 403 else "helpers(<403)"
    a = agg.setdefault(p, [0, 0, 0, collections.Counter()]); a[0] += v[0]; a[1] += v[1]; a[2] += v[2]; a[3].update(v[5])
print("%-18s %8s %8s %8s  top stall reasons (share of the phase's samples)" % ("phase", "inst%", "samples%", "smemwf%"))
for p, a in agg.items():
    ts_ = max(sum(a[3].values()), 1)
    print("%-18s %8.1f %8.1f %8.1f  %s" % (p, 100.0 * a[0] / totI, 100.0 * a[1] / max(totS, 1), 100.0 * a[2] / max(totW, 1), ", ".join("%s %.0f%%" % (k.replace("stall_", ""), 100.0 * c / ts_) for k, c in a[3].most_common(5))))
print("\ntop lines by samples:")
for k, v in sorted(per.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%5d %5.1f%% smp %5.1f%% ins %5.1f%% wf | %s | %s" % (k[1], 100.0 * v[1] / max(totS, 1), 100.0 * v[0] / totI, 100.0 * v[2] / max(totW, 1), k[2][:90], dict(v[3].most_common(4))))

totX = sum(v[4] for v in per.values())
print("\ntop lines by shared-memory wavefronts (excessive = beyond the conflict-free minimum; total excessive %.1f%% of all):" % (100.0 * totX / max(totW, 1)))
for k, v in sorted(per.items(), key=lambda kv: -kv[1][2])[:top]:
    print("%5d %5.1f%% wf (%5.1f%% of them excessive) %5.1f%% ins | %s" % (k[1], 100.0 * v[2] / max(totW, 1), 100.0 * v[4] / max(v[2], 1), 100.0 * v[0] / totI, k[2][:100]))
