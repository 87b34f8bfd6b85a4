"""Per-kernel counts of the SASS instructions that show what the kernels use -> profiles/r2_sass_evidence.md
(cuobjdump -sass of ocrfdet_b200/_build/*.o; run after `python -m ocrfdet_b200.build`)."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "ocrfdet_b200", "_build")
PATS = ["UBLKCP", "UTMA", "LDGSTS", "SYNCS", "MATCH", "REDUX", "RED.E", "ATOMG", "ATOM.E", "ACQBULK", "PREEXIT", "MUFU.EX2",
        "DFMA", "HMMA", "UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "FFMA2", "FMUL2", "FADD2", "UCGABAR", "SHFL", "VOTE", "LDS.128", "STS.128", "LDG.E.128", "STG.E.128", "WARPSYNC"]
INFIX = ("RED.E", "ATOM.E", "LDS.128", "STS.128", "LDG.E.128", "STG.E.128", "MUFU.EX2")
out = ["# SASS evidence (cuobjdump -sass of the sm_100a objects; count of instructions per kernel)", "",
       "UBLKCP = 1-D bulk copy issued to the TMA unit (cp.async.bulk), SYNCS = mbarrier operations, LDGSTS = cp.async,",
       "ACQBULK / PREEXIT = griddepcontrol.wait / launch_dependents (programmatic dependent launch), MATCH = match_any ranking,",
       "RED/ATOM = global reductions, DFMA = fp64 (the cross-tile gradient sums and the preprocess backward),",
       "HMMA = mma.sync on the tensor cores (TF32, 3-term split: the feature-gradient product of the many-channel backward),",
       "UTCHMMA = tcgen05.mma (kind::tf32), UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05.alloc / dealloc, LDTM / STTM =",
       "tcgen05.ld / tcgen05.st (tensor memory): the 80-channel blend kernels of render_tc_fwd.cu / render_tc_bwd.cu,",
       "FFMA2 / FMUL2 / FADD2 = packed FP32 (the C = 3 blend kernels), UCGABAR = cluster barrier (visible_sort.cu).",
       "No tensor-core instruction appears on the C = 3 path: none of its stages is a dense contraction (DESIGN.md 2.10).", ""]
for o in sorted(f for f in os.listdir(BUILD) if f.endswith(".o")):
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(BUILD, o)], capture_output=True, text=True).stdout
    cur, cnt = None, collections.OrderedDict()
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            cur = name.replace("void ", "").replace("ocrf::", "")
            cnt[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(2)
            cnt[cur]["TOTAL"] += 1
            for p in PATS:
                if op.startswith(p) or (p in INFIX and p in op):
                    cnt[cur][p] += 1
    out.append("## " + o.replace(".o", ".cu"))
    for k, c in cnt.items():
        items = ", ".join("%s %d" % (p, c[p]) for p in PATS if c[p])
        out.append("- `%s` (%d instructions): %s" % (k, c["TOTAL"], items or "-"))
    out.append("")
open(os.path.join(ROOT, "profiles", "r2_sass_evidence.md"), "w").write("\n".join(out))
print("wrote profiles/r2_sass_evidence.md")
