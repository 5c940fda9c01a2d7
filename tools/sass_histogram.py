"""Opcode histogram per kernel of libdeer_b200.so (cuobjdump -sass): the Blackwell-specific instructions that show which
kernels run on tcgen05 / TMEM / TMA / DSMEM (mnemonics: /opt/skills/guides/B200_PROFILING.md).

    python tools/sass_histogram.py > profiles/r2_sass_histogram.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "uncertainty-aware-multimodal-emotion-recognition_b200", "csrc", "libdeer_b200.so")
KEY = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP",
       "SYNCS", "UCGABAR", "HMMA", "LDGSTS", "MUFU", "FFMA2", "FADD2", "FMUL2", "ATOM", "RED", "STAS"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, check=True).stdout.decode()
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE).stdout.decode().strip()
            name = re.sub(r"\(.*", "", name).replace("void ", "")
            cur = per.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            cur[m.group(2)] += 1
            cur["_total"] += 1
    tot = collections.Counter()
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(per)} kernels; columns = static SASS instruction counts")
    print("# " + " ".join(KEY))
    for name, c in per.items():
        hits = [(k, c[k]) for k in KEY if c[k]]
        for k, v in hits:
            tot[k] += v
        print(f"{name[:110]:110s} total={c['_total']:6d}  " + "  ".join(f"{k}={v}" for k, v in hits))
    print("\n# whole library: " + "  ".join(f"{k}={tot[k]}" for k in KEY if tot[k]))


if __name__ == "__main__":
    main()
