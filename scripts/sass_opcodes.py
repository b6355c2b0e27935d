"""Per-kernel counts of the SASS opcodes that prove the Blackwell paths.   python scripts/sass_opcodes.py > profiles/r2/sass_opcodes.txt"""
import collections
import re
import subprocess

LIB = "rnnwavefunctions_b200/librnnwf_b200.so"
KEYS = ["UTCHMMA", "LDTM", "STTM", "UBLKCP", "UTCBAR", "SYNCS", "DMMA", "LDGSTS", "MUFU", "FFMA2", "DFMA", "FFMA", "ELECT"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.splitlines()
print(f"# cuobjdump -sass {LIB} : per-kernel counts of the opcodes that prove the Blackwell paths")
print("# UTCHMMA = tcgen05.mma kind::f16/tf32, LDTM/STTM = tcgen05.ld/st, UBLKCP = cp.async.bulk (TMA), UTCBAR = tcgen05.commit, SYNCS = mbarrier, "
      "DMMA = mma.sync f64, LDGSTS = cp.async")
total = collections.Counter()
for name, body in zip(names, re.split(r"Function : \S+", out)[1:]):
    c = collections.Counter()
    for m in re.finditer(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", body, flags=re.M):
        op = m.group(1)
        for k in KEYS:
            if op == k or (k in ("UTCHMMA", "UBLKCP", "SYNCS", "UTCBAR", "MUFU", "ELECT", "LDGSTS") and op.startswith(k)):
                c[k] += 1
                break
    total.update(c)
    if c:
        print(name[:150])
        print("    " + ", ".join(f"{k} {c[k]}" for k in KEYS if c[k]))
print("# library total: " + ", ".join(f"{k} {total[k]}" for k in KEYS if total[k]))
