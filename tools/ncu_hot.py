"""Top stall sites of one kernel from an ncu report: python tools/ncu_hot.py REPORT.ncu-rep KERNEL_REGEX [N]
Prints the N SASS instructions with the most warp-stall samples, with their dominant stall reasons."""
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + pat, "--print-source", "sass"],
                     capture_output=True, text=True).stdout
lines = raw.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(lines[start:]))
# several launches may follow each other: keep the first block
block = []
for r in rows:
    if r["Address"] == "Address":
        break
    block.append(r)
stalls = [k for k in block[0] if k.startswith("stall_") and "Not Issued" not in k]
tot = sum(int(r["# Samples"] or 0) for r in block)
reason_tot = {k: sum(int(r[k] or 0) for r in block) for k in stalls}
print("total samples", tot, {k: v for k, v in sorted(reason_tot.items(), key=lambda kv: -kv[1])[:8]})
for r in sorted(block, key=lambda r: -int(r["# Samples"] or 0))[:top]:
    rs = sorted(((int(r[k] or 0), k) for k in stalls), reverse=True)[:3]
    print("%6s %5.1f%%  %-70s %s" % (r["# Samples"], 100.0 * int(r["# Samples"] or 0) / max(tot, 1), r["Source"][:70],
                                   " ".join("%s=%d" % (k[6:], v) for v, k in rs if v)))
