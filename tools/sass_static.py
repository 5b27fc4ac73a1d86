#!/usr/bin/env python3
"""Development aid: static SASS opcode histogram of one kernel in an object file (no GPU needed).

    tools/sass_static.py <object.o> <mangled-kernel-substring> [top]"""
import collections, re, subprocess, sys
obj, sub = sys.argv[1:3]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
ops, active, n = collections.Counter(), False, 0
for l in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", l)
    if m:
        active = sub in m.group(1)
        continue
    if not active:
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(.*?);", l)
    if m:
        s = m.group(1).strip()
        t = s.split()
        op = t[1] if s.startswith("@") else t[0]
        ops[op if op.startswith("IMAD.MOV") else op.split(".")[0]] += 1
        n += 1
print("static instructions:", n)
print("  ".join("%s %d" % kv for kv in ops.most_common(top)))
