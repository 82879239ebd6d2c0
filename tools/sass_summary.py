"""Per-kernel count of the Blackwell tensor-core / TMA / TMEM SASS mnemonics in libscp_b200.so (cuobjdump -sass, no GPU
needed): UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA load / store, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit,
SYNCS = mbarrier operations.  Writes a markdown table (profiles/r02_sass_tensor_ops.md)."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "scp_b200/libscp_b200.so"
out = sys.argv[2] if len(sys.argv) > 2 else "profiles/r02_sass_tensor_ops.md"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
ops = ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "SYNCS", "REDUX", "MATCH")
cur, arch, table = None, None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch = m.group(1)
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        table.setdefault(cur, collections.Counter())["_arch_" + str(arch)] += 1
        continue
    if cur:
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            table[cur]["_n"] += 1
            for o in ops:
                if m.group(1).startswith(o):
                    table[cur][o] += 1
with open(out, "w") as f:
    f.write(f"# SASS mnemonics per kernel of `{lib}` (`cuobjdump -sass`; all cubins are {arch})\n\n")
    f.write("Only kernels that use the tensor cores, TMA, tensor memory or warp-wide reduction / match instructions are listed.\n\n")
    f.write("| kernel | SASS instructions | " + " | ".join(ops) + " |\n|---|---|" + "---|" * len(ops) + "\n")
    for k, c in table.items():
        if any(c[o] for o in ops):
            f.write(f"| `{k}` | {c['_n']} | " + " | ".join(str(c[o]) if c[o] else "" for o in ops) + " |\n")
    f.write(f"\nTotals: " + ", ".join(f"{o} {sum(c[o] for c in table.values())}" for o in ops) + f"; {len(table)} kernels in the library.\n")
print(open(out).read())
