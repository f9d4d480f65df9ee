"""Per-kernel counts of the Blackwell-specific SASS mnemonics in libsdvg.so (cuobjdump -sass; no GPU needed):
UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load, LDTM = tcgen05.ld (TMEM -> registers), UTCBAR = tcgen05.commit,
UBLKCP = cp.async.bulk (DSMEM copies), LDGSTS = cp.async, HMMA = mma.sync (legacy tensor path, attention only).

    python tools/sass_counts.py > profiles/r2_sass_counts.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sd-video-gen_b200", "libsdvg.so")
MNEMONICS = ("UTCHMMA", "UTMALDG", "LDTM", "UTCBAR", "UBLKCP", "LDGSTS", "HMMA", "SYNCS", "UCGABAR")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1).split(".")[0]
            counts[cur]["_total"] += 1
            if op in MNEMONICS:
                counts[cur][op] += 1
    names = list(counts)
    out = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    for n, d in zip(names, out):
        demangle[n] = re.sub(r"\(.*", "", d)
    arch = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout.strip().splitlines()
    print("# libsdvg.so:", ", ".join(a.split()[-1] for a in arch if "sm_" in a))
    print("# kernel | instructions | " + " | ".join(MNEMONICS))
    for n in names:
        c = counts[n]
        if not any(c[m] for m in MNEMONICS[:7]):
            continue
        print(f"{demangle[n]} | {c['_total']} | " + " | ".join(str(c[m]) for m in MNEMONICS))


if __name__ == "__main__":
    sys.exit(main())
