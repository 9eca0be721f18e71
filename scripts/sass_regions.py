"""Static SASS size of a kernel by source region: every instruction is attributed to its OUTERMOST source line in the kernel
(nvdisasm -g -gi prints the inline chain, innermost first).  usage: sass_regions.py <cubin> <mangled kernel> [bin size in lines]"""
import collections, re, subprocess, sys
cubin, kernel = sys.argv[1], sys.argv[2]
binsz = int(sys.argv[3]) if len(sys.argv) > 3 else 10
txt = subprocess.run(["nvdisasm", "-g", "-gi", cubin], capture_output=True, text=True).stdout.splitlines()
inside = False
outer = None; sub = "kernel"
count = collections.Counter(); subs = collections.Counter()
for ln in txt:
    if ln.startswith(".text."):
        inside = ln.strip().rstrip(":") == ".text." + kernel
        sub = "kernel"
        continue
    if not inside:
        continue
    m = re.match(r"\$?(\S+):$", ln.strip())
    if m and not ln.strip().startswith(".L_"):
        sub = m.group(1)[-60:]
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        outer = (m.group(1).split("/")[-1], int(m.group(2)))   # the last marker before an instruction is the outermost frame
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+\S", ln):
        if sub == "kernel":
            f, l = outer if outer else ("?", 0)
            count[(f, (l // binsz) * binsz)] += 1
        else:
            subs[sub] += 1
tot = sum(count.values()) + sum(subs.values())
print("total %d instr = %.1f KB" % (tot, tot * 16 / 1024))
for (f, l), v in sorted(count.items()):
    print("  %-16s %5d-%-5d %5d instr %5.2f KB" % (f, l, l + binsz - 1, v, v * 16 / 1024))
for k, v in subs.most_common():
    print("  sub %-62s %5d instr %5.2f KB" % (k, v, v * 16 / 1024))
