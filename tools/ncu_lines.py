"""Warp-stall samples per CUDA source line of one kernel: python tools/ncu_lines.py REPORT KERNEL_REGEX [N]"""
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + pat, "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
agg, cur_file, cur = {}, None, None
first_kernel_done = False
for row in csv.reader(raw.splitlines()):
    if not row:
        continue
    if row[0] == "Kernel Name":
        if first_kernel_done:
            break
        first_kernel_done = True
        continue
    if row[0] == "File Name":
        cur_file = row[1].split("/")[-1]
        continue
    if row[0] == "Line No":
        hdr = row
        si = hdr.index("# Samples")
        stall_ix = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if row[0] != "":  # a CUDA source line
        try:
            cur = (cur_file, int(row[0]), row[1].strip()[:90])
        except ValueError:
            pass
        continue
    if cur is None or len(row) <= si:
        continue
    try:
        n = int(row[si] or 0)
    except ValueError:
        continue
    a = agg.setdefault(cur, [0, {}])
    a[0] += n
    for i, name in stall_ix:
        try:
            v = int(row[i] or 0)
        except ValueError:
            v = 0
        if v:
            a[1][name] = a[1].get(name, 0) + v
tot = sum(v[0] for v in agg.values())
print("total samples", tot)
for (f, ln, src), (n, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    rs = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print("%5d %5.1f%% %s:%d  %s   [%s]" % (n, 100.0 * n / max(tot, 1), f, ln, src, " ".join("%s=%d" % kv for kv in rs)))
