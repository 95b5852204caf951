#!/usr/bin/env python
"""Summarise the source page of an ncu report: opcode histogram weighted by executed count, and stall samples per
region (contiguous SASS lines with the same execution count).  usage: ncu_src_summary.py report.ncu-rep [min_samples]"""
import collections, csv, io, re, subprocess, sys
rep = sys.argv[1]
minsmp = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[1]
isrc, ie, ismp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for i, r in enumerate(rows[2:]):
    if len(r) <= ie: continue
    src = r[isrc].strip()
    op = re.sub(r"^@!?U?P\d+\s+", "", src).split()[0].split(".")[0]
    data.append((i, int(r[ie] or 0), int(r[ismp] or 0), op, src, [int(r[k] or 0) for k in stall]))
print(rows[0][1][:120])
tot = sum(d[1] for d in data); ts = sum(d[2] for d in data)
print("instructions %.4g  samples %d" % (tot, ts))
ops = collections.Counter(); osm = collections.Counter()
for d in data: ops[d[3]] += d[1]; osm[d[3]] += d[2]
for op, n in ops.most_common(16): print("  %-10s %14d %5.1f%%  samples %d" % (op, n, 100.0 * n / tot, osm[op]))
runs = []; cur = None
for d in data:
    if cur and cur["cnt"] > 0 and abs(d[1] - cur["cnt"]) <= 0.03 * cur["cnt"]:
        cur["n"] += 1; cur["tot"] += d[1]; cur["smp"] += d[2]; cur["end"] = d[0]; cur["ops"][d[3]] += 1
        for k in range(len(stall)): cur["st"][k] += d[5][k]
    else:
        if cur: runs.append(cur)
        cur = dict(start=d[0], end=d[0], cnt=d[1], n=1, tot=d[1], smp=d[2], ops=collections.Counter({d[3]: 1}), st=list(d[5]))
runs.append(cur)
print("regions with >= %d samples:" % minsmp)
for r in runs:
    if r["smp"] >= minsmp:
        top = sorted(range(len(stall)), key=lambda k: -r["st"][k])[:4]
        print("  lines %5d-%5d n=%4d count=%11d samples=%8d (%4.1f%%) %s | %s" % (
            r["start"], r["end"], r["n"], r["cnt"], r["smp"], 100.0 * r["smp"] / ts,
            " ".join("%s:%d" % (k, v) for k, v in r["ops"].most_common(4)),
            " ".join("%s:%d" % (hdr[stall[k]][6:], r["st"][k]) for k in top)))
