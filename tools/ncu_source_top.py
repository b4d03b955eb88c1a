#!/usr/bin/env python3
"""Aggregate the ncu source-page CSV (gz): stall-reason totals, instruction-class mix, hottest SASS lines.
usage: python tools/ncu_source_top.py profiles/ncu_source_<k>.csv.gz [top_n]"""
import collections
import csv
import gzip
import sys

path = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 15
rows = list(csv.reader(gzip.open(path, "rt")))
hdr = next(i for i, r in enumerate(rows[:20]) if "Address" in r and "Source" in r)
names = rows[hdr]
ix = {n: i for i, n in enumerate(names)}
stall_cols = [n for n in names if n.startswith("stall_")]
tot = collections.Counter()
cls = collections.Counter()
cls_samples = collections.Counter()
lines = []
launch = 0
for r in rows[hdr + 1:]:
    if len(r) != len(names):
        continue
    if r[ix["Address"]] == names[0]:
        continue
    try:
        samples = int(r[ix["# Samples"]] or 0)
        execd = int(r[ix["Instructions Executed"]] or 0)
    except ValueError:
        continue
    src = r[ix["Source"]].strip()
    op = src.split()[0] if src else "?"
    if op.startswith("@"):
        op = src.split()[1] if len(src.split()) > 1 else op
    key = op.split(".")[0]
    cls[key] += execd
    cls_samples[key] += samples
    for sc in stall_cols:
        try:
            tot[sc] += int(r[ix[sc]] or 0)
        except ValueError:
            pass
    lines.append((samples, execd, src))
S = sum(tot.values()) or 1
print("stall samples by reason:")
for k, v in tot.most_common(10):
    print("   %-28s %9d  %5.1f%%" % (k, v, 100.0 * v / S))
E = sum(cls.values()) or 1
print("instruction mix (warp-level executed) and share of samples:")
SS = sum(cls_samples.values()) or 1
for k, v in cls.most_common(14):
    print("   %-14s %12d  %5.1f%%   samples %5.1f%%" % (k, v, 100.0 * v / E, 100.0 * cls_samples[k] / SS))
print("hottest SASS lines:")
for s, e, src in sorted(lines, reverse=True)[:topn]:
    print("   %7d samples  %10d exec   %s" % (s, e, src[:110]))
