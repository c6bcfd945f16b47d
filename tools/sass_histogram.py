"""Developer tool (no GPU): static SASS opcode histograms of selected kernels of the in-tree objects -> profiles/*_sass_histograms.txt.

    python tools/sass_histogram.py > profiles/r2zz_sass_histograms.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "medical_image_analysis_b200", "build")
# object, regex on the demangled kernel name
PICK = [
    ("scan_fwd_bf16.o", r"ss_fwd_cw_kernel<__nv_bfloat16, true, false, [14]>"),
    ("scan_bwd_bf16.o", r"ss_bwd_cw_kernel<__nv_bfloat16, true, false, 1, 12>"),
    ("scan_bwd_bf16.o", r"ss_bwd_cw_kernel<__nv_bfloat16, true, false, 4, 8>"),
    ("scan_bwd_bf16.o", r"ss_bwd_rowsn_kernel<__nv_bfloat16, __nv_bfloat16, true, false>"),
    ("scan_fwd_bf16.o", r"ss_fwd_rowsn_kernel<__nv_bfloat16, true, false, 16, false>"),
    ("gemm_tcgen05.o", r"gemm_tn_kernel<128, 6, false, false>"),
]


def main():
    print("# SASS opcode histograms (static instruction counts per kernel, cuobjdump -sass of the in-tree objects, final build of round 2)")
    print("# evidence mnemonics: UTMALDG / UTMASTG = tensor-map TMA box load / store, SYNCS = mbarrier, LDGSTS = cp.async, FFMA2 / FMUL2 / FADD2 =")
    print("#                     packed f32x2, LDS.128 / STS.128 = the 16-byte tile accesses, UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UBLKCP = 1-D TMA bulk copy")
    for obj, pat in PICK:
        sass = subprocess.run(["cuobjdump", "-sass", os.path.join(BUILD, obj)], capture_output=True, text=True, check=True).stdout
        name, hist, keep = None, None, False
        for line in sass.splitlines():
            m = re.match(r"\s*Function : (\S+)", line)
            if m:
                if keep:
                    emit(name, hist)
                dem = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().replace("(bool)1", "true").replace(
                    "(bool)0", "false").replace("(int)", "")
                name, hist, keep = dem, collections.Counter(), re.search(pat, dem) is not None
                continue
            m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_]+)?)", line)
            if keep and m:
                op = m.group(1)
                base = op.split(".")[0]
                hist[op if base in ("LDS", "STS", "MUFU", "UTMALDG", "UTMASTG", "SYNCS", "LDGSTS", "LDTM", "UTCHMMA") else base] += 1
        if keep:
            emit(name, hist)


def emit(name, hist):
    print(f"\n== {name[:170]}  ({sum(hist.values())} instructions)")
    print("; ".join(f"{v} {k}" for k, v in hist.most_common()))


if __name__ == "__main__":
    main()
