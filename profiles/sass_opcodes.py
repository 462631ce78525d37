"""Per-kernel SASS opcode counts of the built library (CPU only: cuobjdump -sass).
usage: python profiles/sass_opcodes.py > profiles/r2_04_sass_opcodes.md"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dgvit-depth-goal-guided-vision-transformer-_b200", "libdgvit.so")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
COLS = ["UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "LDGMC", "MUFU.TANH", "MUFU.EX2", "FADD2", "LDGSTS", "HMMA"]
rows, cur, i = [], None, -1
for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        i += 1
        cur = collections.Counter()
        nm = re.sub(r"\(.*", "", names[i]).replace("void ", "").replace("dgvit::", "")
        rows.append((nm, cur))
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur is not None:
        op = m.group(1)
        for c in COLS:
            if op == c or op.startswith(c + "."):
                cur[c] += 1
print("# SASS opcode summary of libdgvit.so (round 2, final code)\n")
print("`cuobjdump -sass libdgvit.so`, counted per kernel (sm_100a; `profiles/sass_opcodes.py`). `UTCHMMA` = `tcgen05.mma kind::f16`, `LDTM` = `tcgen05.ld`, "
      "`UTMALDG` / `UTMASTG` = TMA tile\nload / store (`cp.async.bulk.tensor`), `UTCBAR` = `tcgen05.commit`, `SYNCS` = mbarrier ops, `LDGMC` = "
      "`multimem.ld_reduce` (NVLS), `FADD2` = packed fp32 pairs,\n`LDGSTS` = `cp.async`. There is no `HMMA` (legacy `mma.sync`) in any kernel of the library.\n")
print("| kernel | " + " | ".join(COLS) + " |")
print("|---|" + "---:|" * len(COLS))
tot = collections.Counter()
for nm, c in sorted(rows, key=lambda r: (-r[1]["UTCHMMA"], -sum(r[1].values()), r[0])):
    tot.update(c)
    if sum(c.values()) == 0:
        continue
    print(f"| `{nm[:90]}` | " + " | ".join(str(c[k]) for k in COLS) + " |")
print(f"| **all {len(rows)} kernels** | " + " | ".join(str(tot[k]) for k in COLS) + " |")
