"""SASS opcode digest of libctl_b200.so: counts of the Blackwell-native mnemonics (tcgen05.mma -> UTC*MMA, tcgen05.ld ->
LDTM, TMA -> UTMALDG / UTMASTG / UBLKCP, tcgen05.commit -> UTCBAR, mbarrier -> SYNCS, griddepcontrol -> ACQBULK/...)
per kernel family, from `cuobjdump -sass`.  usage: python tools/sass_digest.py [out.txt]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cooperative_training_and_latent_space_data_augmentation_b200", "libctl_b200.so")
WATCH = ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF", "SYNCS", "HMMA",
         "HGMMA", "LDGSTS", "RED.E", "ATOMG", "ACQBULK", "PREEXIT", "ELECT", "SHFL", "BAR.SYNC")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
            name = name.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("void ", "")
            name = re.sub(r"\(.*", "", name)
            fam = re.sub(r"<.*", "", name)
            cur = per.setdefault(fam, {"kernels": 0, "ops": collections.Counter(), "instr": 0})
            cur["kernels"] += 1
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        cur["instr"] += 1
        for w in WATCH:
            if op.startswith(w):
                cur["ops"][w] += 1
    total = collections.Counter()
    lines = ["SASS digest of %s (cuobjdump -sass; sm_100a)" % os.path.relpath(LIB, ROOT),
             "%-42s %8s %9s  %s" % ("kernel family", "kernels", "instr", "watched opcodes")]
    for fam, d in per.items():
        total.update(d["ops"])
        lines.append("%-42s %8d %9d  %s" % (fam[:42], d["kernels"], d["instr"],
                                           " ".join("%s=%d" % kv for kv in sorted(d["ops"].items()))))
    lines.append("TOTAL " + " ".join("%s=%d" % kv for kv in sorted(total.items())))
    text = "\n".join(lines)
    print(text)
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(text + "\n")


if __name__ == "__main__":
    main()
