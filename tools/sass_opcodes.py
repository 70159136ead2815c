#!/usr/bin/env python3
"""Per-kernel counts of the SASS mnemonics that prove the Blackwell paths (tcgen05 MMA / TMEM / bulk copies / SFU):
    python tools/sass_opcodes.py [lib.so] > profiles/rNN_sass_opcodes.txt      (cuobjdump -sass, no GPU needed)"""
import collections, os, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                         "fs_uae_image_enhancer_project_b200", "libfsuae_enhancer.so")
OPS = ["UTCHMMA", "UTCBAR", "LDTM", "UBLKCP", "LDGSTS", "SYNCS", "MUFU", "RED", "MEMBAR", "CCTL"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
fn, counts, arch = None, collections.OrderedDict(), set()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        counts[fn] = collections.Counter()
        continue
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        arch.add(m.group(1))
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
    if m and fn:
        counts[fn]["_all"] += 1
        for op in OPS:
            if m.group(1).startswith(op):
                counts[fn][op] += 1
print(f"# {os.path.basename(lib)}: arch = {', '.join(sorted(arch))}; instructions per kernel (cuobjdump -sass)")
print(f"# {'kernel':110s} {'total':>7s} " + " ".join(f"{o:>7s}" for o in OPS))
tot = collections.Counter()
for fn, c in counts.items():
    short = re.sub(r"fsuae::\(anonymous namespace\)::", "", fn)
    short = re.sub(r"\(anonymous namespace\)::", "", short)
    print(f"{short[:112]:112s} {c['_all']:7d} " + " ".join(f"{c[o]:7d}" for o in OPS))
    tot.update(c)
print(f"{'ALL KERNELS':112s} {tot['_all']:7d} " + " ".join(f"{tot[o]:7d}" for o in OPS))
