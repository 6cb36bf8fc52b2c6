#!/usr/bin/env python
"""SASS opcode histogram of the shipped sm_100a kernels (no GPU needed: cuobjdump on the objects libb200pa.so is linked
from).  One block per kernel: instruction count and the opcodes that prove how it moves data and computes -
UBLKCP (TMA bulk copy), SYNCS (mbarrier), LDGSTS (cp.async), DFMA / DMUL / DADD (FP64 pipe), DMMA (FP64 tensor core:
absent), LDS / STS, LDG / STG, LDCU / LDC (constant-bank operands), BAR, SHFL.  ATOMG = 1 in the kernels that end in a reduction is
the last-block ticket of the deterministic two-level sum (csrc/reduce.cuh); no kernel has an atomic or a RED on data.
usage: python tools/sass_hist.py > profiles/r2_sass_histogram.txt"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "cardiac-ablation-ecm2_b200", "build")
KEY = ["UBLKCP", "SYNCS", "LDGSTS", "DFMA", "DMUL", "DADD", "DMMA", "LDS", "STS", "LDG", "STG", "LDCU", "LDC", "R2UR", "BAR", "SHFL", "ATOM", "ATOMG", "RED",
       "MUFU", "IMAD", "UTCHMMA", "HMMA"]
WANT = re.compile(r"pa_apply_kernel|k_segment_sum|k_diag_sf|k_pcg_|k_px_|k_dot|k_diffusion_setup|k_mass_setup|pa_element_kernel")


def main():
    print("# SASS opcode histogram, sm_100a, from cardiac-ablation-ecm2_b200/build/*.o (cuobjdump -sass); columns:")
    print("# kernel | instructions | " + " ".join(KEY))
    for obj in sorted(glob.glob(os.path.join(OBJ, "*.o"))):
        out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
        name, hist = None, None
        blocks = []
        for line in out.splitlines():
            m = re.match(r"\s+Function : (\S+)", line)
            if m:
                name, hist = m.group(1), collections.Counter()
                blocks.append((name, hist))
                continue
            m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
            if m and hist is not None:
                hist[m.group(1)] += 1
        shown = False
        for name, hist in blocks:
            dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
            if not WANT.search(dem):
                continue
            if not shown:
                print(f"\n## {os.path.basename(obj)}")
                shown = True
            short = dem.replace("b200pa::", "").replace("(int)", "").replace("(bool)", "").replace("void ", "")
            short = re.sub(r"\(.*", "", short)
            print(f"{short[:70]:70s} | {sum(hist.values()):6d} | " + " ".join(f"{k}={hist[k]}" for k in KEY if hist[k]))


if __name__ == "__main__":
    main()
