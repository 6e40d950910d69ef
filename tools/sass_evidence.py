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
                 r"F2FP[.\w]*E4M3[.\w]*|REDUX[.\w]*|LDGMC[.\w]*|UBLKCP[.\w]*)")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    print("# SASS evidence (cuobjdump -sass learning-at-home_b200/_C/liblah_cuda.so), counts of Blackwell-native mnemonics per kernel")
    print("# UTC*MMA = tcgen05.mma (UTCQMMA/UTCOMMA...: 8-bit / block-scaled kinds), UTCCP = tcgen05.cp (scale factors smem->TMEM),")
    print("# LDTM = tcgen05.ld, UTMALDG = TMA load, SYNCS = mbarrier, UTCBAR = tcgen05.commit, .2CTA = cta_group::2,")
    print("# LDG/STG .STRONG.SYS + MEMBAR.SYS = NVLink flag protocol, F2FP...E4M3 = fp8 conversion, UTMASTG = TMA store,")
    print("# LDGMC = multimem.ld_reduce (NVLS in-switch reduction; multimem.st compiles to STG...STRONG.SYS on the multicast address)\n")
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


if __name__ == "__main__" and "--listings" not in sys.argv:
    main()


def dump_listings(outdir, patterns=("wgrad_adam_kernel", "swapab_kernelILb0", "gemm2_kernelILi0ELb0ELb0ELb0", "attention_bwd_kernel",
                                    "attention_fwd_v2", "nvls_allreduce", "gemm_fp8_kernelILb0")):
    """full SASS listing of the hot kernels (one file each) -> profiles/sass/"""
    os.makedirs(outdir, exist_ok=True)
    names = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    funcs = re.findall(r"Function : (\S+)", names)
    for pat in patterns:
        for fn in [f for f in funcs if pat in f][:1]:
            txt = subprocess.run(["cuobjdump", "-sass", "-fun", fn, SO], capture_output=True, text=True).stdout
            short = re.sub(r"[^A-Za-z0-9_]", "", pat)
            with open(os.path.join(outdir, short + ".sass"), "w") as f:
                f.write(txt)
            print("wrote", short + ".sass", len(txt.splitlines()), "lines")


if __name__ == "__main__" and "--listings" in sys.argv:
    dump_listings(os.path.join(ROOT, "profiles", "sass"))
