"""
SASS evidence of the Blackwell-native paths: per kernel of libplayslot_b200.so, the count of tcgen05 / TMEM / TMA instructions
(mnemonics per /opt/skills/guides/B200_PROFILING.md).  Runs in the build container (cuobjdump needs no GPU).

    python tools/sass_summary.py > profiles/r02_sass_summary.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "playslot_b200", "libplayslot_b200.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "STTM", "UTCCP", "SYNCS", "HMMA", "FFMA"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    fn = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.split("\n")
    names = iter(fn)
    counts, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = next(names)
            cur = re.sub(r"\(.*", "", cur.replace("(anonymous namespace)::", "")).replace("void ", "").strip()
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1).split(".")[0]
            counts[cur][op] += 1
            counts[cur]["_total"] += 1
    print("# SASS summary of playslot_b200/libplayslot_b200.so (cuobjdump -sass, sm_100a)\n")
    print("Instruction counts per kernel; UTCHMMA = tcgen05.mma (kind::f16), UTCBAR = tcgen05.commit, UTMALDG = TMA load, LDTM/STTM = "
          "tcgen05.ld/st (TMEM), SYNCS = mbarrier ops.\n")
    print("| kernel | total | " + " | ".join(MNEMONICS) + " |")
    print("|---|---|" + "---|" * len(MNEMONICS))
    tot = collections.Counter()
    for k, c in counts.items():
        if not any(c[m] for m in MNEMONICS[:-1]):
            continue
        print(f"| `{k}` | {c['_total']} | " + " | ".join(str(c[m]) for m in MNEMONICS) + " |")
        tot.update({m: c[m] for m in MNEMONICS})
    print("| **all kernels above** | | " + " | ".join(str(tot[m]) for m in MNEMONICS) + " |")
    print(f"\n{len(counts)} kernels in the library; kernels without tensor-core / TMA instructions are omitted from the table.")


if __name__ == "__main__":
    main()
