"""SASS opcode histogram per kernel of libnjf_b200.so (profiles/sass_r02_opcodes.txt).
usage: python tools/sass_histogram.py [lib] > profiles/sass_rNN_opcodes.txt   (needs cuobjdump and c++filt on PATH)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEY = ["UTCHMMA", "UTCQMMA", "UTCMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "HFMA2", "FFMA", "FFMA2", "LDG", "STG",
       "LDS", "STS", "F2FP", "MUFU", "BAR", "ATOMS", "ATOMG", "RED", "REDG"]


def main(lib):
    sha = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    fn, hist = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            fn = fn.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0]
            hist[fn] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and fn:
            hist[fn][m.group(1)] += 1
    print(f"SASS opcode histogram per kernel of {os.path.basename(lib)} (sm_100a), build {sha}: cuobjdump -sass, first mnemonic component.")
    print("Blackwell evidence: UTCHMMA = tcgen05.mma (kind::f16 in the render kernels, kind::tf32 in tt_gemm_tc / tt_wgrad_tc), "
          "LDTM/STTM = tcgen05.ld/st (TMEM), UBLKCP = cp.async.bulk (TMA engine), UTCBAR = tcgen05.commit, SYNCS = mbarrier.\n")
    for fn in sorted(hist):
        h = hist[fn]
        print(f"{fn}  [{sum(h.values())} instructions]")
        print("   " + "  ".join(f"{k}:{h[k]}" for k in sorted(KEY, key=lambda k: -h[k]) if h[k]))
        print("   top: " + "  ".join(f"{k}:{v}" for k, v in h.most_common(14)))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "neural-jacobian-field_b200", "lib", "libnjf_b200.so"))
