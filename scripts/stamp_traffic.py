"""Writes profiles/k1_k2_traffic_stamp.json from an `ncu --set full` capture of the bench workload (raw-page CSV):
DRAM bytes per launch and executed-flop counters of k_render / k_encode, stamped with the commit and the hash of the kernel
sources they were taken at.  bench.py reports these as roofline.traffic / roofline.executed and says whether the kernels
have changed since.    usage: stamp_traffic.py <raw.csv> <profile file name the numbers are from> [<flop counters csv>]
The optional third file is the long-format `ncu --metrics sm__sass_thread_inst_executed_op_* --csv` log of the same
command (`--set full` does not collect the per-opcode counters); its k_render row supplies the executed flops."""
import csv, hashlib, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]


def val(d, k):
    try:
        return float(d[k].replace(",", ""))
    except (KeyError, ValueError):
        return 0.0


def scaled(d, units, k):
    v = val(d, k)
    u = units.get(k, "")
    return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(u, 1.0)


out = {"profile": sys.argv[2]}
units = dict(zip(hdr, rows[1]))
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    name = d.get("Kernel Name", "")
    dram = scaled(d, units, "dram__bytes_read.sum") + scaled(d, units, "dram__bytes_write.sum")
    if "k_render" in name and "k_render_dram_bytes_per_launch" not in out:
        out["k_render_dram_bytes_per_launch"] = dram
        g = lambda op: val(d, f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum") or val(d, f"sm__sass_thread_inst_executed_op_{op}_pred_on.sum")
        out["k_render_executed"] = {
            "fp64_flop": g("dadd") + g("dmul") + 2 * g("dfma"), "fp32_flop": g("fadd") + g("fmul") + 2 * g("ffma"),
            "warp_instructions": val(d, "smsp__inst_executed.sum"), "issue_active_pct": val(d, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "kernel_ms_under_ncu": val(d, "gpu__time_duration.sum")}
    elif "k_tile_certs" in name and "k_tile_certs_dram_bytes_per_launch" not in out:
        out["k_tile_certs_dram_bytes_per_launch"] = dram
    elif "k_encode" in name and "k_encode_dram_bytes_per_launch" not in out:
        out["k_encode_dram_bytes_per_launch"] = dram
if len(sys.argv) > 3 and "k_render_executed" in out:
    m = {}
    for r in csv.reader(l for l in open(sys.argv[3]) if l.startswith('"')):
        if len(r) > 14 and "k_render" in r[4]:
            try:
                m[r[12]] = float(r[14].replace(",", ""))
            except ValueError:
                pass
    g = lambda op: m.get(f"sm__sass_thread_inst_executed_op_{op}_pred_on.sum", 0.0)
    out["k_render_executed"]["fp64_flop"] = g("dadd") + g("dmul") + 2 * g("dfma")
    out["k_render_executed"]["fp32_flop"] = g("fadd") + g("fmul") + 2 * g("ffma")
    out["k_render_executed"]["flop_counters_from"] = "profiles/" + os.path.basename(sys.argv[3])
h = hashlib.sha256()
csrc = os.path.join(ROOT, "terminalraytracer_b200", "csrc")
for name in sorted(os.listdir(csrc)):
    if name.endswith((".cu", ".cuh", ".h")):
        h.update(open(os.path.join(csrc, name), "rb").read())
out["kernel_sources_sha256"] = h.hexdigest()
out["commit"] = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
json.dump(out, open(os.path.join(ROOT, "profiles", "k1_k2_traffic_stamp.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
