"""Per CUDA-source-line instruction and stall-sample shares of k_render from `ncu --page source --csv --print-source cuda`.
usage: ncu_cuda_lines.py <csv> [top] [kernel substring]"""
import csv, collections, sys
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
want = sys.argv[3] if len(sys.argv) > 3 else "k_render"
rows = list(csv.reader(open(path)))
cur = None; hdr = None; data = []; kern = ""
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': kern = r[1]; continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr and want in kern:
        try:
            ln = int(r[0]); inst = int(r[hdr.index('Instructions Executed')]); s = int(r[hdr.index('# Samples')])
        except ValueError:
            continue
        data.append((cur, ln, r[1], inst, s))
ti = sum(d[3] for d in data); ts = sum(d[4] for d in data)
print("warp-inst", ti, "samples", ts)
byfile = collections.Counter(); sfile = collections.Counter()
for d in data:
    byfile[d[0]] += d[3]; sfile[d[0]] += d[4]
for f in byfile: print("  %-18s inst %5.2f%% samples %5.2f%%" % (f, 100 * byfile[f] / ti, 100 * sfile[f] / ts))
for d in sorted(data, key=lambda d: -d[4])[:top]:
    print("%-16s %4d inst %5.2f%% samp %5.2f%%  %s" % (d[0], d[1], 100 * d[3] / ti, 100 * d[4] / ts, d[2][:110]))
