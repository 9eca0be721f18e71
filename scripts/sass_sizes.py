"""SASS size and opcode histogram of the kernels of a built library (cuobjdump -sass).
usage: sass_sizes.py <lib.so> [kernel substring for the opcode histogram]"""
import collections, re, subprocess, sys
lib = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else "k_renderILb0ELi1ELi1E"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
n = collections.Counter(); ops = collections.Counter(); name = None
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        name = m.group(1); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m:
        n[name] += 1
        if want in name:
            ops[m.group(2).split(".")[0]] += 1
for k, v in sorted(n.items(), key=lambda x: x[1]):
    print("%7.1f KB %6d instr  %s" % (v * 16 / 1024, v, k))
print("opcode histogram of", want)
tot = sum(ops.values())
for k, v in ops.most_common(25):
    print("  %-10s %5d  %4.1f%%" % (k, v, 100.0 * v / max(tot, 1)))
