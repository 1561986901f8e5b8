#!/usr/bin/env python
"""Static evidence of what the kernels of libnerf_b200.so are made of: per kernel, the SASS mnemonics that mark the
Blackwell paths (UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UBLKCP/UTMALDG/UTMASTG = TMA-engine bulk copies,
SYNCS = mbarrier, HMMA = legacy mma.sync, which must not appear) and the resource usage ptxas recorded.

    python tools/sass_evidence.py > profiles/<round>_sass_evidence.txt          (cuobjdump; no GPU needed)
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "nerf-and-dietnerf_b200", "libnerf_b200.so")
MARKS = [("UTC*MMA", r"\bUTC[A-Z]*MMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UBLKCP", r"\bUBLKCP"),
         ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("SYNCS", r"\bSYNCS"), ("HMMA", r"\bHMMA"),
         ("MUFU", r"\bMUFU"), ("SHFL", r"\bSHFL"), ("LDG.128", r"\bLDG\.E\.128|\bLDG\.E\.[A-Z.]*128"),
         ("STG.128", r"\bSTG\.E\.128|\bSTG\.E\.[A-Z.]*128")]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, order, cur = collections.defaultdict(collections.Counter), [], None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            order.append(cur)
            continue
        if cur and "/*" in line:
            counts[cur]["instructions"] += 1
            for tag, pat in MARKS:
                if re.search(pat, line):
                    counts[cur][tag] += 1
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True, check=True).stdout
    usage, fn = {}, None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            fn = m.group(1)
            continue
        if fn and "REG:" in line:
            usage[fn] = " ".join(re.findall(r"(?:REG|STACK|SHARED|LOCAL):\d+", line))
            fn = None
    names = demangle(order)
    print(f"# cuobjdump -sass / --dump-resource-usage of {os.path.relpath(LIB, ROOT)} (sm_100a); counts are static instruction counts")
    print("# UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UBLKCP = cp.async.bulk (TMA engine), SYNCS = mbarrier ops, "
          "HMMA = legacy mma.sync (expected: none)")
    total_hmma = 0
    for fn in order:
        c = counts[fn]
        total_hmma += c["HMMA"]
        short = re.sub(r"\(.*", "", names[fn])[:70]
        marks = " ".join(f"{tag}={c[tag]}" for tag, _ in MARKS if c[tag])
        print(f"{short:70s} inst={c['instructions']:6d} {usage.get(fn, ''):38s} {marks}")
    print(f"# kernels: {len(order)}; HMMA instructions in the whole library: {total_hmma}")


if __name__ == "__main__":
    sys.exit(main())
