"""Join an ncu SASS-level source export with nvdisasm line info: per source line, instructions executed and
stall samples.   usage: ncu_lines.py <src.csv from `ncu --page source --csv`> <nvdisasm -g dump> <mangled kernel> [top]"""
import csv, re, sys, collections
src_csv, dis, kernel = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# address -> (file,line) from nvdisasm
addr2line = {}
cur = None
inside = False
for ln in open(dis, errors="replace"):
    if ln.startswith(".text."):
        inside = ln.strip().rstrip(":") == ".text." + kernel
        cur = None
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", ln)
    if m:
        addr2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ia, ii, isamp, isrc = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
base = None
inst = collections.Counter(); samp = collections.Counter(); stalls = collections.defaultdict(collections.Counter)
fp64 = collections.Counter()
for r in rows[2:]:
    try:
        a = int(r[ia], 16) if r[ia].startswith("0x") else int(r[ia])
        n, s = int(r[ii]), int(r[isamp])
    except ValueError:
        continue
    if base is None:
        base = a
    key = addr2line.get(a - base)
    inst[key] += n; samp[key] += s
    op = r[isrc].split()[1] if r[isrc].startswith("@") else r[isrc].split()[0]
    if op.split(".")[0] in ("DADD", "DMUL", "DFMA", "DSETP"):
        fp64[key] += n
    for c in stall_cols:
        try:
            v = int(r[c])
        except ValueError:
            v = 0
        if v:
            stalls[key][hdr[c]] += v
ti, ts = sum(inst.values()), sum(samp.values())
print("total warp-inst %d, samples %d, fp64 warp-inst %d" % (ti, ts, sum(fp64.values())))
for key, s in samp.most_common(top):
    st = ", ".join("%s %.0f%%" % (k.replace("stall_", ""), 100.0 * v / max(s, 1)) for k, v in stalls[key].most_common(3))
    print("%-22s inst %5.2f%%  fp64 %5.2f%%  samples %5.2f%%   %s" % (str(key), 100.0 * inst[key] / ti, 100.0 * fp64[key] / ti, 100.0 * s / ts, st))
