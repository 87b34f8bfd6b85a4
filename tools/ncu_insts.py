"""Executed warp instructions per CUDA source line of one kernel: python tools/ncu_insts.py REPORT KERNEL_REGEX [N]"""
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + pat, "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
agg, cur_file, cur, ii = {}, None, None, None
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
        ii = row.index("Instructions Executed")
        continue
    if row[0] != "":
        try:
            cur = (cur_file, int(row[0]), row[1].strip()[:90])
        except ValueError:
            pass
        continue
    if cur is None or ii is None or len(row) <= ii:
        continue
    try:
        n = int(row[ii] or 0)
    except ValueError:
        continue
    agg[cur] = agg.get(cur, 0) + n
tot = sum(agg.values())
print("total warp instructions", tot)
for (f, ln, src), n in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
    print("%11d %5.1f%%  %s:%d  %s" % (n, 100.0 * n / max(tot, 1), f, ln, src))
