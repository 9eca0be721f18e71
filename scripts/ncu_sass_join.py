"""Per source line / per source function shares of a kernel's executed instructions and stall samples, from an ncu report's SASS
page joined with the line table of the cubin the report was taken from (works when `--print-source cuda` cannot correlate because
the source file has changed since).
    ncu -i x.ncu-rep --page source --csv --print-source sass [-k regex:k_render] > sass.csv
    cuobjdump -xelf all lib.so ; nvdisasm -g -gi trt_render.sm_100a.cubin > render.txt
    python scripts/ncu_sass_join.py sass.csv render.txt <mangled kernel> <source file the cubin was built from> [top]
Every instruction is attributed to its INNERMOST source line (the first marker of nvdisasm's inline chain) and to the function that
line belongs to (nearest preceding line of the source that looks like a function head)."""
import collections, csv, re, sys

sass_csv, disasm, kernel, source = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40

# ---- line table: offset -> (file, line) innermost, and outermost line in the kernel's own file
table = {}
inside = False
chain = []
fresh = False
for ln in open(disasm):
    if ln.startswith(".text."):
        inside = ln.strip().rstrip(":") == ".text." + kernel
        chain = []
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        if not fresh:
            chain = []
            fresh = True
        chain.append((m.group(1).split("/")[-1], int(m.group(2))))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", ln)
    if m:
        fresh = False
        table[int(m.group(1), 16)] = (chain[0] if chain else ("?", 0), chain[-1] if chain else ("?", 0), m.group(2).strip(), list(chain))

# ---- function heads of the source file
src = open(source, errors="replace").read().splitlines()
heads = []
for i, text in enumerate(src, 1):
    if re.match(r"^(static |template|__device__|__global__|TRT_HD|inline)", text) and "(" in text and not text.rstrip().endswith(";"):
        m = re.search(r"\b(k_[A-Za-z0-9_]+)\s*\(", text) if "__global__" in text else re.search(r"([A-Za-z_][A-Za-z0-9_<>]*)\s*\(", text)
        if m:
            heads.append((i, m.group(1) + (" (kernel body)" if "__global__" in text else "")))


def other_text(k):
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "terminalraytracer_b200", "csrc", k[0])
    try:
        return open(path, errors="replace").read().splitlines()[k[1] - 1].strip()[:100]
    except (OSError, IndexError):
        return ""


_heads_other = {}


def func_in(name, line):
    import os
    if name not in _heads_other:
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "terminalraytracer_b200", "csrc", name)
        hs = []
        for i, text in enumerate(open(path, errors="replace").read().splitlines(), 1):
            if re.match(r"^(static |template|__device__|__global__|TRT_HD|inline)", text) and "(" in text and not text.rstrip().endswith(";"):
                m = re.search(r"([A-Za-z_][A-Za-z0-9_<>+*\-]*)\s*\(", text.replace("operator", "operator"))
                if m:
                    hs.append((i, m.group(1)))
        _heads_other[name] = hs
    out = "?"
    for i, n in _heads_other[name]:
        if i <= line:
            out = n
        else:
            break
    return out


def func_of(line):
    name = "?"
    for i, n in heads:
        if i <= line:
            name = n
        else:
            break
    return name


rows = list(csv.reader(open(sass_csv)))
hdr = None
data = []
use = False
for r in rows:
    if r and r[0] == "Kernel Name":
        use = kernel in r[1] or re.sub(r"[^A-Za-z0-9_]", "", kernel) in re.sub(r"[^A-Za-z0-9_]", "", r[1]) or True
        continue
    if r and r[0] == "Address":
        hdr = r
        ia, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr and r and r[0].startswith("0x"):
        data.append((int(r[0], 16), r[1].strip(), int(r[ia]), int(r[isamp])))
base = data[0][0]
ti = sum(d[2] for d in data)
ts = sum(d[3] for d in data)
mism = 0
by_line = collections.Counter(); s_line = collections.Counter()
by_func = collections.Counter(); s_func = collections.Counter()
by_outer = collections.Counter(); s_outer = collections.Counter()
for addr, text, inst, samp in data:
    off = addr - base
    if off not in table:
        mism += 1
        continue
    inner, outer, dis, chain = table[off]
    if dis.split()[0].lstrip("@!P0123456789U ") .split(".")[0] != text.split()[0].lstrip("@!P0123456789U ").split(".")[0] and not text.startswith("@"):
        mism += 1
    by_line[inner] += inst; s_line[inner] += samp
    # the innermost frame that lies in one of our own files names the function
    f = inner[0]
    for fr in chain:
        if fr[0] == "trt_render.cu":
            f = func_of(fr[1]); break
        if fr[0] in ("trt_device.cuh", "trt_cert.h"):
            f = fr[0] + ":" + func_in(fr[0], fr[1]); break
    by_func[f] += inst; s_func[f] += samp
    by_outer[(outer[1] // 10) * 10] += inst; s_outer[(outer[1] // 10) * 10] += samp
print("warp-instructions %d  samples %d  (instructions without a line / opcode mismatch: %d of %d)" % (ti, ts, mism, len(data)))
print("---- by function of the innermost line")
for f, v in by_func.most_common(25):
    print("  %-34s inst %5.2f%%  samples %5.2f%%" % (f, 100.0 * v / ti, 100.0 * s_func[f] / ts))
print("---- by innermost line")
for k, v in by_line.most_common(top):
    text = src[k[1] - 1].strip()[:100] if k[0] == "trt_render.cu" and 0 < k[1] <= len(src) else other_text(k)
    print("  %-22s %5d inst %5.2f%% samp %5.2f%%  %s" % (k[0][:22], k[1], 100.0 * v / ti, 100.0 * s_line[k] / ts, text))
print("---- by region of the kernel body (outermost line, bins of 10)")
for k, v in sorted(by_outer.items()):
    if v * 200 > ti:
        print("  line %5d+  inst %5.2f%%  samples %5.2f%%" % (k, 100.0 * v / ti, 100.0 * s_outer[k] / ts))
