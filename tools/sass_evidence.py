"""Count the Blackwell-native SASS mnemonics of every kernel in the in-tree library (CPU-only: cuobjdump -sass).
    python tools/sass_evidence.py > profiles/sass_evidence_r1.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "learning-at-home_b200", "_C", "liblah_cuda.so")
PAT = re.compile(r"\b(UTC[A-Z0-9]*MMA[.\w]*|UTCCP[.\w]*|UTCBAR[.\w]*|UTCATOMSWS[.\w]*|LDTM[.\w]*|STTM[.\w]*|UTMALDG[.\w]*|UTMASTG[.\w]*|"
                 r"SYNCS[.\w]*|UCGABAR_\w+|MEMBAR[.\w]*|LDG\.E[.\w]*SYS|STG\.E[.\w]*SYS|ATOMG[.\w]*|ATOMS[.\w]*|REDG[.\w]*|CCTL[.\w]*|"
                 r"F2FP[.\w]*E4M3[.\w]*|REDUX[.\w]*)")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    print("# SASS evidence (cuobjdump -sass learning-at-home_b200/_C/liblah_cuda.so), counts of Blackwell-native mnemonics per kernel")
    print("# UTC*MMA = tcgen05.mma (UTCQMMA/UTCOMMA...: 8-bit / block-scaled kinds), UTCCP = tcgen05.cp (scale factors smem->TMEM),")
    print("# LDTM = tcgen05.ld, UTMALDG = TMA load, SYNCS = mbarrier, UTCBAR = tcgen05.commit, .2CTA = cta_group::2,")
    print("# LDG/STG .STRONG.SYS + MEMBAR.SYS = NVLink flag protocol, F2FP...E4M3 = fp8 conversion\n")
    kernel, counts = None, None
    out = []
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if kernel:
                out.append((kernel, counts))
            kernel, counts = m.group(1), collections.Counter()
            continue
        if kernel:
            for tok in PAT.findall(line):
                counts[tok] += 1
    if kernel:
        out.append((kernel, counts))
    for kernel, counts in out:
        print(kernel)
        print("    " + (", ".join(f"{k} x{v}" for k, v in sorted(counts.items())) or "(no tcgen05/TMA/sys-scope instructions)"))


if __name__ == "__main__":
    main()
