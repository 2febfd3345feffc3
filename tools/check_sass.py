#!/usr/bin/env python
"""Build-time gate for the one hardware assumption of the barrier-free schedules: a vertex record moves
as ONE 256-bit strong access.  Disassembles the given objects / library with cuobjdump and fails (exit 1) if any kernel whose
name matches --kernels contains a strong global load/store of a vertex record that is NOT 256 bits wide, or contains no 256-bit
strong access at all (the toolchain split the access, or dropped the .relaxed qualifier).

  python tools/check_sass.py xpbd-fem_b200/build/xf_dataflow.cu.o xpbd-fem_b200/build/xf_part.cu.o [--excerpt profiles/r2_sass_records.txt]
"""
import argparse
import re
import subprocess
import sys

ap = argparse.ArgumentParser()
ap.add_argument("objects", nargs="+")
ap.add_argument("--kernels", default=r"k_substeps_(dataflow|chain|cluster|tiles)|k_part_dataflow")
ap.add_argument("--excerpt", default=None, help="write the matching SASS lines (per kernel, de-duplicated opcodes) to this file")
a = ap.parse_args()

pat = re.compile(a.kernels)
bad, seen, excerpt = [], 0, []
for obj in a.objects:
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True)
    if out.returncode != 0:
        print("check_sass: cuobjdump failed on %s: %s" % (obj, out.stderr.strip()), file=sys.stderr)
        sys.exit(1)
    fn, ops = None, {}
    def flush():
        global seen
        if fn is None or not pat.search(fn):
            return
        seen += 1
        strong = {k: v for k, v in ops.items() if "STRONG" in k}
        wide = [k for k in strong if ".256." in k]
        narrow = [k for k in strong if ".128" in k and not k.startswith("ATOM")]  # a split record would show up as 2 x 128
        has_ld = any(k.startswith("LDG") for k in wide)
        has_st = any(k.startswith("STG") for k in wide)
        excerpt.append("%s  (%s)" % (fn, obj))
        for k in sorted(ops):
            excerpt.append("    %5d x %s" % (ops[k], k))
        if not (has_ld and has_st):
            bad.append("%s: no 256-bit strong load+store pair (found %s)" % (fn, sorted(strong)))
        if narrow:
            bad.append("%s: strong vector access narrower than 256 bits: %s" % (fn, narrow))
    for line in out.stdout.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            flush()
            fn, ops = m.group(1), {}
            continue
        m = re.search(r"\b(LDG|STG|ATOMG|RED)(\.[A-Z0-9_.]+)?", line)
        if m and fn:
            op = m.group(0)
            ops[op] = ops.get(op, 0) + 1
    flush()

if a.excerpt:
    with open(a.excerpt, "w") as f:
        f.write("# global memory opcodes per barrier-free kernel (cuobjdump -sass, counts of static instructions)\n")
        f.write("\n".join(excerpt) + "\n")
if seen == 0:
    print("check_sass: no kernel matched %r" % a.kernels, file=sys.stderr)
    sys.exit(1)
if bad:
    print("check_sass: FAILED\n  " + "\n  ".join(bad), file=sys.stderr)
    sys.exit(1)
print("check_sass: %d barrier-free kernels move their vertex records with LDG/STG.E.ENL2.256.STRONG only" % seen)
