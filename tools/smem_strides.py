#!/usr/bin/env python
"""Search the shared-memory strides (SQ, ES, SXS) of pa_apply_kernel that minimise bank-conflict
wavefronts, per order.  Model: 8-byte accesses are served per half-warp (16 lanes x 8 B = 128 B = all 32
banks once); lanes reading the same address are one request; the cost of one warp instruction is the sum
over its two half-warps of the maximum number of distinct addresses falling on one 8-byte bank."""
import itertools
import sys

CFG = {2: 28, 3: 8, 4: 5, 5: 3, 6: 2, 7: 2}  # D -> NEB (pa_apply_kernel.cuh)


def cost(addr_of_lane, nlanes):
    """wavefronts of one warp-wide instruction family: lanes 0..nlanes-1 grouped in warps of 32"""
    tot = 0
    for w0 in range(0, nlanes, 32):
        for h0 in (w0, w0 + 16):
            banks = {}
            for l in range(h0, min(h0 + 16, nlanes)):
                a = addr_of_lane(l)
                if a is None:
                    continue
                banks.setdefault(a % 16, set()).add(a)
            if banks:
                tot += max(len(v) for v in banks.values())
    return tot


def total(D, Q, NEB, SQ, ES, SXS):
    Q2, D2 = Q * Q, D * D
    NT = ((NEB * Q2 + 31) // 32) * 32
    c = 0
    nA = NEB * D * Q
    # phase A: reads of the x slab (one per (dy,dx)), writes of 3 fields x Q values
    for k in range(D2):
        c += cost(lambda l: (l // Q) * SXS + k, nA)
    rowpat = lambda l, x, f: ((l // Q) // D) * ES + ((l // Q) % D) * SQ + (l % Q) * Q + x + f * D * SQ
    for f in range(3):
        for qx in range(Q):
            c += 2 * cost(lambda l: rowpat(l, qx, f), nA)      # A write + C1 read
    for f in range(2):
        for dx in range(D):
            c += cost(lambda l: rowpat(l, dx, f), nA)           # C1 write
    # phase B: column reads + writes
    nB = NEB * Q2
    for f in range(3):
        for dz in range(D):
            c += 2 * cost(lambda l: (l // Q2) * ES + (l % Q2) + (f * D + dz) * SQ, nB)
    # phase C2: reads (2 fields x Q), writes D
    nC = NEB * D * D
    for f in range(2):
        for qy in range(Q):
            c += cost(lambda l: ((l // D) // D) * ES + ((l // D) % D) * SQ + (l % D) + qy * Q + f * D * SQ, nC)
    for dy in range(D):
        c += cost(lambda l: (l // D) * SXS + dy * D + (l % D), nC)
    # stage in / out
    nio = NEB * D * D2
    for r in range((nio + NT - 1) // NT):
        c += 2 * cost(lambda l: ((l + r * NT) // D2) * SXS + (l + r * NT) % D2 if l + r * NT < nio else None, NT)
    return c


def main():
    for D, NEB in CFG.items():
        Q = D + 1
        base = (Q * Q | 1, None, D * D | 1)
        cur_sq = 19 if D == 3 else (Q * Q | 1)
        cur_es = 185 if D == 3 else (3 * D * cur_sq + (1 if (3 * D * cur_sq) % 2 == 0 else 0))
        cur = total(D, Q, NEB, cur_sq, cur_es, D * D | 1)
        best = None
        for SQ in range(Q * Q, Q * Q + 17):
            for pad in range(0, 17):
                ES = 3 * D * SQ + pad
                for SXS in range(D * D, D * D + 9):
                    t = total(D, Q, NEB, SQ, ES, SXS)
                    key = (t, ES * NEB + 2 * NEB * D * SXS)
                    if best is None or key < best[0]:
                        best = (key, SQ, ES, SXS)
        print(f"D={D} Q={Q} NEB={NEB}: current (SQ={cur_sq}, ES={cur_es}, SXS={D*D|1}) cost {cur}; best SQ={best[1]} ES={best[2]} "
              f"SXS={best[3]} cost {best[0][0]}", flush=True)


if __name__ == "__main__":
    main()
