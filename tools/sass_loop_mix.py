#!/usr/bin/env python
"""Instruction mix of the largest loop of every kernel in a cuobjdump -sass listing (static issue-slot count of a probe's solve loop).
usage: cuobjdump -sass xpbd-fem_b200/build/xf_probe_coop.cu.o | python tools/sass_loop_mix.py"""
import collections
import re
import sys

txt = sys.stdin.read()
for f in re.split(r'\n\s*Function : ', txt)[1:]:
    name = f.split('\n')[0].strip()
    ins = []
    for line in f.split('\n'):
        m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);', line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2)))
    loops = []
    for addr, t in ins:
        m = re.search(r'\bBRA\S*\s+.*?0x([0-9a-f]+)', t)
        if m and int(m.group(1), 16) < addr:
            loops.append((int(m.group(1), 16), addr))
    print("%s: %d instructions" % (name, len(ins)))
    if not loops:
        continue
    a, b = max(loops, key=lambda x: x[1] - x[0])
    cnt = collections.Counter()
    for ad, t in ins:
        if a <= ad <= b:
            t = re.sub(r'^@!?U?P\d+\s+', '', t)
            cnt[t.split()[0].split('.')[0]] += 1
    print("  largest loop 0x%x-0x%x: %d instructions: %s" % (a, b, sum(cnt.values()), dict(cnt.most_common())))
