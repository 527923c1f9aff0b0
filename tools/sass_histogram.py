#!/usr/bin/env python
"""SASS opcode histogram of one kernel of the built library (cuobjdump -sass on the per-file object):
    python tools/sass_histogram.py dmc_bwrf8u_h2 'bwrf8u_h2_kernelILi5ELi4ELi0ELi1E' > profiles/r02_sass_bwrf8u_h2_r5.md"""
import collections, os, re, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
obj = os.path.join(root, "depthmapcompression_b200", "_obj", sys.argv[1] + ".o")
pat = sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
cur, body = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); body[cur] = []; continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
    if m and cur:
        body[cur].append(m.group(1).strip())
names = [n for n in body if pat in n]
if not names:
    raise SystemExit("no function matches %r; have: %s" % (pat, ", ".join(list(body)[:20])))
for n in names:
    ins = body[n]
    ops = collections.Counter()
    for i in ins:
        t = i.split()
        op = t[1] if t[0].startswith("@") else t[0]
        ops[op.split(".")[0]] += 1
    demangled = subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
    print("# SASS opcode histogram: `%s`\n" % demangled)
    print("`cuobjdump -sass depthmapcompression_b200/_obj/%s.o` (nvcc 12.9, sm_100a), %d instructions (%.1f KB)\n" % (sys.argv[1], len(ins), len(ins) * 16 / 1024.0))
    print("| opcode | count | share |\n|---|---|---|")
    for op, c in ops.most_common():
        print("| %s | %d | %.1f %% |" % (op, c, 100.0 * c / len(ins)))
    print()
