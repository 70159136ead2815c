#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections, csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    name = row["Kernel Name"]
    m = re.search(r"conv3x3_tc_kernel<(\d+), (\d+), (\d+), (\d+)", name)
    if "fused_pair" in name:
        key = "tc_fused_pair<conv3+conv4>"
    else:
        key = f"tc<PT={m.group(1)},N={m.group(2)},kind={m.group(4)}>" if m else re.sub(r"\(.*", "", name)[-48:]
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    v = v / 1000 if unit in ("ns", "nsecond") else v * 1000 if unit in ("ms", "msecond") else v
    agg.setdefault(key, []).append(v)
# a pass launches every kernel once: shares are taken over the per-kernel averages (the capture window may cut a pass)
tot = sum(sum(v) / len(v) for v in agg.values())
for k, v in agg.items():
    print(f"{k:50s} n={len(v):3d} avg={sum(v)/len(v):9.1f} us  share of one pass={100*(sum(v)/len(v))/tot:5.1f}%")
print(f"{'one pass (64 frames), serialised under ncu':50s}       sum={tot:9.1f} us")
