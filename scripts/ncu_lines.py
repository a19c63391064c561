"""Per-source-line warp-stall samples of one kernel of an ncu report (needs -lineinfo + --import-source on).
Usage: python scripts/ncu_lines.py <report.ncu-rep> <kernel index (0-based)> [top N]"""
import csv, subprocess, sys
rep, kid = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-id", ":::%d" % (kid + 1)],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None
hdr = None
lines = {}
stall_cols = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        fn = r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        si = hdr.index("Warp Stall Sampling (All Samples)")
        ie = hdr.index("Instructions Executed")
        stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or len(r) <= si:
        continue
    if r[0] != "":            # a CUDA source line (aggregate over its SASS)
        try:
            n = int(r[si])
        except ValueError:
            continue
        key = (cur_file, int(r[0]))
        st = {h: int(r[i]) for i, h in stall_cols if r[i].isdigit() and int(r[i]) > 0}
        lines[key] = (n, int(r[ie]) if r[ie].isdigit() else 0, r[1].strip()[:100], st)
tot = sum(v[0] for v in lines.values())
print(fn[:120])
print("total samples", tot)
for key, v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    st = sorted(v[3].items(), key=lambda kv: -kv[1])[:3]
    print("%5.1f%% %7d inst %-14s:%-5d %-100s %s" % (100.0 * v[0] / max(tot, 1), v[1], key[0], key[1], v[2], st))
