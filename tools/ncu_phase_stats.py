#!/usr/bin/env python
"""Per-phase view of one captured pa_apply_kernel launch from the ncu SASS source page (read on the CPU box):
the instruction stream is cut at the BAR.SYNCs (= the kernel's phases: prologue, stage-in + A, B, C1, C2[, stage-out])
and for every segment the warp-stall samples, executed instructions, FP64 / R2UR+MOV / LDCU shares and shared-memory
wavefronts (measured / conflict-free) of LDS, STS and LDGSTS are summed; then the 12 most-sampled instructions.
usage: tools/ncu_phase_stats.py gpurun_out/prof.ncu-rep > profiles/<name>.txt   (capture with --import-source on)"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    # a report may hold several launches: keep the first kernel's section (rows up to the next "Kernel Name" line)
    nxt = [i for i, r in enumerate(rows) if i > 0 and r and r[0] == "Kernel Name"]
    if nxt:
        rows = rows[:nxt[0]]
    rows = [r for r in rows if r]
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    R = rows[2:]
    src = lambda r: r[ix["Source"]].strip()
    op = lambda r: (src(r).split()[1] if src(r).startswith("@") else src(r).split()[0]).split(".")[0]
    num = lambda r, k: int(r[ix[k]])
    bars = [n for n, r in enumerate(R) if "BAR.SYNC" in src(r)]
    bounds = [0] + bars + [len(R)]
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tots = sum(num(r, "# Samples") for r in R)
    print(f"# {rows[0][1] if len(rows[0]) > 1 else ''}")
    print(f"# {rep}: {sum(num(r, 'L1 Wavefronts Shared') for r in R)} shared-memory wavefronts from SM instructions, {tots} stall samples")
    for k in range(len(bounds) - 1):
        seg = R[bounds[k]:bounds[k + 1]]
        ex = lambda pred: sum(num(r, "Instructions Executed") for r in seg if pred(r))
        wf = lambda pred: (sum(num(r, "L1 Wavefronts Shared") for r in seg if pred(r)) / 1e6,
                           sum(num(r, "L1 Wavefronts Shared Ideal") for r in seg if pred(r)) / 1e6)
        samples = sum(num(r, "# Samples") for r in seg)
        agg = sorted(((sum(num(r, s) for r in seg), s[6:]) for s in stalls), reverse=True)[:4]
        lds, sts, gs = wf(lambda r: op(r) == "LDS"), wf(lambda r: op(r) == "STS"), wf(lambda r: op(r) == "LDGSTS")
        print(f"segment {k} [{bounds[k]:4d},{bounds[k + 1]:4d}) samples {samples:6d} ({samples / max(tots, 1):.2f})  inst {ex(lambda r: True) / 1e6:6.1f}M"
              f"  fp64 {ex(lambda r: op(r) in ('DFMA', 'DMUL', 'DADD')) / 1e6:5.1f}M  r2ur+mov {ex(lambda r: op(r) in ('R2UR', 'MOV')) / 1e6:5.1f}M"
              f"  ldcu {ex(lambda r: op(r) == 'LDCU') / 1e6:4.1f}M  wavefronts M (measured/ideal): LDS {lds[0]:.1f}/{lds[1]:.1f}"
              f" STS {sts[0]:.1f}/{sts[1]:.1f} LDGSTS {gs[0]:.1f}/{gs[1]:.1f} | " + " ".join(f"{n}={v}" for v, n in agg))
    top = sorted(((num(r, "# Samples"), n, src(r)[:70]) for n, r in enumerate(R)), reverse=True)[:12]
    print("# most-sampled instructions (samples, index, SASS)")
    for t in top:
        print(f"  {t[0]:6d} {t[1]:5d}  {t[2]}")


if __name__ == "__main__":
    main()
