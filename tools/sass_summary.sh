#!/bin/bash
# SASS evidence of the shipped library: per-kernel instruction counts of the mnemonics that matter on this path
# (bulk async copy / mbarrier, 128-bit loads and stores, atomics and reductions, match / redux, cluster barrier),
# and a check that no tensor-core instruction is present (integer / memory path).
#   tools/sass_summary.sh > profiles/sass_r02.txt
SO=${1:-throttlecrab_b200/libgcra_b200.so}
cuobjdump -sass "$SO" > /tmp/gcra_sass.txt
echo "# cuobjdump -sass $SO  ($(date -u +%Y-%m-%dT%H:%MZ), $(nvcc --version | tail -2 | head -1))"
echo "# arch: $(grep -m1 -o 'sm_[0-9a-z]*' /tmp/gcra_sass.txt)"
echo
python - <<'PY'
import re, collections
fn, per = None, collections.OrderedDict()
pat = {"UBLKCP (1-D bulk async copy, TMA)": r"\bUBLKCP", "SYNCS (mbarrier)": r"\bSYNCS\.", "LDG.E.128": r"\bLDG\.E\.128", "STG.E.128": r"\bSTG\.E\.128",
       "ATOMG.*CAS": r"\bATOMG\.[A-Z.0-9]*CAS", "ATOMG (returning)": r"\bATOMG\.", "REDG (posted)": r"\bREDG?\.E", "ATOMS (shared)": r"\bATOMS\.",
       "MATCH.ANY": r"\bMATCH\.ANY", "REDUX": r"\bREDUX", "UCGABAR (cluster barrier)": r"\bUCGABAR", "CCTL": r"\bCCTL", "MEMBAR.SYS / fence": r"\bMEMBAR\.[A-Z.]*SYS",
       "tensor core (HMMA/IMMA/UTC*MMA)": r"\b(HMMA|IMMA|UTC[A-Z]*MMA|QGMMA|HGMMA)"}
for line in open("/tmp/gcra_sass.txt"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        per[fn] = collections.Counter()
        continue
    if fn and re.search(r"/\*[0-9a-f]{4}\*/", line):
        per[fn]["instructions"] += 1
        for k, p in pat.items():
            if re.search(p, line):
                per[fn][k] += 1
import subprocess
def demangle(n):
    try:
        return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().split("(")[0].replace("void gcra::", "")
    except Exception:
        return n
tot = collections.Counter()
for fn, c in per.items():
    tot.update(c)
    items = ", ".join("%s %d" % (k, v) for k, v in c.items() if k != "instructions")
    print("%-42s %5d instr  %s" % (demangle(fn)[:42], c["instructions"], items))
print()
print("TOTAL " + ", ".join("%s %d" % (k, v) for k, v in tot.items()))
print("tensor-core instructions: %d (expected 0: integer / memory path)" % tot.get("tensor core (HMMA/IMMA/UTC*MMA)", 0))
PY
