#!/usr/bin/env python
"""Bank-conflict model of pa_apply_kernel's shared-memory phases and the search that chose its per-order strides
and lane -> task maps (ApplyCfg in cardiac-ablation-ecm2_b200/csrc/pa_apply_kernel.cuh).

Model: 8-byte accesses are served per half-warp (16 lanes x 8 B = 128 B = all 32 banks once); lanes reading the same
address are one request; the cost of one warp-wide instruction is the sum over its half-warps of the largest number
of distinct addresses that fall on one 8-byte bank.  Checked against ncu (`l1tex__data_pipe_lsu_wavefronts_mem_shared`,
source page of profiles' r1e / r1h captures): 1486 modelled vs 1471 measured wavefronts per batch at p = 4.

    python tools/smem_strides.py report          # wavefronts / ideal per phase for the configuration in the kernel
    python tools/smem_strides.py search D NEB    # search (RQ, SQ, ES, SXS, BS, maps) for one order (minutes)
    python tools/smem_strides.py alt             # the "whole slabs per half-warp" lane maps for the row phases at p = 3..5
"""
import sys


def cost(addrs):
    """wavefronts of one warp-wide 8-byte access; addrs[lane] = double index or None (inactive lane)"""
    tot = 0
    for h0 in range(0, len(addrs), 16):
        banks = {}
        for a in addrs[h0:h0 + 16]:
            if a is not None:
                banks.setdefault(a % 16, set()).add(a)
        if banks:
            tot += max(len(v) for v in banks.values())
    return tot


def ideal(addrs):
    return sum(1 for h0 in range(0, len(addrs), 16) if any(a is not None for a in addrs[h0:h0 + 16]))


def phases(D, Q, NEB, SXS, RQ, SQ, ES, BS, mapA="qy", mapC="dx", QES=None, QMS=None):
    """{phase: [wavefronts, conflict-free wavefronts]} per batch of NEB elements (diffusion + mass, stored q-data)"""
    Q2, D2, D3, Q3 = Q * Q, D * D, D ** 3, Q ** 3
    NT = ((NEB * Q2 + 31) // 32) * 32
    FS = D * SQ
    QES = 6 * Q3 if QES is None else QES
    QMS = Q3 if QMS is None else QMS
    res = {}

    def add(name, fn, n):
        L = [fn(l) if l < n else None for l in range(((n + 31) // 32) * 32)]
        c = i = 0
        for s in range(0, len(L), NT):          # tasks beyond NT: the kernel loops
            c += cost(L[s:s + NT])
            i += ideal(L[s:s + NT])
        r = res.setdefault(name, [0, 0])
        r[0] += c
        r[1] += i

    nsl = NEB * D
    tA = (lambda l: (l // Q, l % Q)) if mapA == "qy" else (lambda l: (l % nsl, l // nsl))     # (slab, qy)
    tC = (lambda l: (l // D, l % D)) if mapC == "dx" else (lambda l: (l % nsl, l // nsl))     # (slab, dx)
    sE = lambda slab, f, qy, x: (slab // D) * ES + f * FS + (slab % D) * SQ + qy * RQ + x
    nA, nB, nC, nio = NEB * D * Q, NEB * Q2, NEB * D * D, NEB * D3
    for k in range(D2):
        add("A.x", lambda l: tA(l)[0] * SXS + k, nA)
    for dy in range(D):
        for _ in range(2):
            add("A.bg", lambda l: 10 ** 6 + tA(l)[1] * BS + dy, nA)
    for f in range(3):
        for qx in range(Q):
            add("A.w", lambda l: sE(tA(l)[0], f, tA(l)[1], qx), nA)
            add("C1.r", lambda l: sE(tA(l)[0], f, tA(l)[1], qx), nA)
    for f in range(2):
        for dx in range(D):
            add("C1.w", lambda l: sE(tA(l)[0], f, tA(l)[1], dx), nA)
    for f in range(3):
        for dz in range(D):
            fn = lambda l: (l // Q2) * ES + f * FS + dz * SQ + ((l % Q2) // Q) * RQ + (l % Q2) % Q
            add("B.r", fn, nB)
            add("B.w", fn, nB)
    for k in range(6):
        for qz in range(Q):
            add("B.qd", lambda l: (l // Q2) * QES + k * Q3 + qz * Q2 + l % Q2, nB)
    for qz in range(Q):
        add("B.qm", lambda l: (l // Q2) * QMS + qz * Q2 + l % Q2, nB)
    for f in range(2):
        for qy in range(Q):
            add("C2.r", lambda l: sE(tC(l)[0], f, qy, tC(l)[1]), nC)
    for dy in range(D):
        add("C2.w", lambda l: tC(l)[0] * SXS + dy * D + tC(l)[1], nC)   # staged output (FUSE_OUT stores to global instead)
    add("gather", lambda l: (l // D2) * SXS + l % D2, nio)
    return res


# the configuration compiled into the kernel: D -> (NEB, SXS, RQ, SQ, ES, BS, mapA, mapC, QES, QMS)
KERNEL = {
    2: (28, 6, 3, 11, 73, 2, "qy", "dx", None, None),
    3: (16, 9, 4, 19, 185, 3, "qy", "dx", None, None),
    4: (5, 17, 5, 28, 345, 4, "slab", "slab", None, None),
    5: (3, 25, 6, 37, 564, 5, "qy", "dx", 6 * 216 + 4, 216 + 12),
    6: (2, 38, 7, 49, 886, 6, "qy", "slab", None, None),
    7: (1, 55, 10, 87, 1841, 7, "qy", "dx", None, None),
}


def report():
    for D, (NEB, SXS, RQ, SQ, ES, BS, mA, mC, QES, QMS) in KERNEL.items():
        r = phases(D, D + 1, NEB, SXS, RQ, SQ, ES, BS, mA, mC, QES, QMS)
        tot, idl = sum(v[0] for v in r.values()), sum(v[1] for v in r.values())
        bad = {k: tuple(v) for k, v in r.items() if v[0] != v[1]}
        print(f"p={D - 1} NEB={NEB}: {tot} wavefronts per batch, {idl} conflict-free ({tot / idl:.2f}x); phases with conflicts: {bad}")


def search(D, NEB):
    Q = D + 1
    Q2, D2, D3 = Q * Q, D * D, D ** 3
    NT = ((NEB * Q2 + 31) // 32) * 32
    nsl = NEB * D
    nA, nB, nC, nio = NEB * D * Q, NEB * Q2, NEB * D * D, NEB * D3
    mapsA = {"qy": lambda l: (l // Q, l % Q), "slab": lambda l: (l % nsl, l // nsl)}
    mapsC = {"dx": lambda l: (l // D, l % D), "slab": lambda l: (l % nsl, l // nsl)}

    def ev(fn, n):
        L = [fn(l) if l < n else None for l in range(((n + NT - 1) // NT) * NT)]
        return sum(cost(L[s:s + NT]) for s in range(0, len(L), NT))

    sx, bs = {}, {}
    for mA, tA in mapsA.items():
        for BS in range(D, D + 4):
            c = 2 * sum(ev(lambda l: tA(l)[1] * BS + dy, nA) for dy in range(D))
            if mA not in bs or c < bs[mA][0]:
                bs[mA] = (c, BS)
        for mC, tC in mapsC.items():
            for SXS in range(D2, D2 + 17):
                c = sum(ev(lambda l: tA(l)[0] * SXS + k, nA) for k in range(D2))
                c += sum(ev(lambda l: tC(l)[0] * SXS + dy * D + tC(l)[1], nC) for dy in range(D))
                c += ev(lambda l: (l // D2) * SXS + l % D2, nio)
                if (mA, mC) not in sx or c < sx[(mA, mC)][0]:
                    sx[(mA, mC)] = (c, SXS)
    out = []
    for RQ in range(Q, Q + 4):
        for SQ in range((Q - 1) * RQ + Q, (Q - 1) * RQ + Q + 16):
            FS = D * SQ
            for pad in range(16):
                ES = 3 * FS + pad
                sE = lambda slab, f, qy, x: (slab // D) * ES + f * FS + (slab % D) * SQ + qy * RQ + x
                cB = 2 * sum(ev(lambda l: (l // Q2) * ES + f * FS + dz * SQ + ((l % Q2) // Q) * RQ + (l % Q2) % Q, nB)
                             for f in range(3) for dz in range(D))
                for mA, tA in mapsA.items():
                    cA = 2 * sum(ev(lambda l: sE(tA(l)[0], f, tA(l)[1], qx), nA) for f in range(3) for qx in range(Q))
                    cA += sum(ev(lambda l: sE(tA(l)[0], f, tA(l)[1], dx), nA) for f in range(2) for dx in range(D))
                    for mC, tC in mapsC.items():
                        cC = sum(ev(lambda l: sE(tC(l)[0], f, qy, tC(l)[1]), nC) for f in range(2) for qy in range(Q))
                        tot = cA + cB + cC + sx[(mA, mC)][0] + bs[mA][0]
                        out.append((tot, NEB * ES, dict(RQ=RQ, SQ=SQ, ES=ES, SXS=sx[(mA, mC)][1], BS=bs[mA][1], mapA=mA, mapC=mC)))
    out.sort(key=lambda t: (t[0], t[1]))
    for t in out[:5]:
        print(t)
    print("best without changing the lane maps:", [t for t in out if t[2]["mapA"] == "qy" and t[2]["mapC"] == "dx"][0])


def alt():
    """Row phases with WHOLE slabs per half-warp (floor(16/Q) slabs x Q rows for A / C1, floor(16/D) slabs x D columns for
    C2, the other lanes idle) instead of tasks packed into consecutive lanes: within a half-warp the addresses are then
    s*SQ + qy*RQ (+ const), which an odd SQ and a suitable element stride ES make conflict-free - but the idle lanes cost
    half-warps.  Prints the row-phase wavefronts per batch next to the shipped configuration's: the gain is 6-7 % of a
    batch's wavefronts at p = 4 and nothing elsewhere, which is why the kernel keeps the packed maps."""
    for D, ES_list in ((4, (345, 346, 349)), (5, (564, 565, 569)), (6, (886, 887, 889))):
        NEB, SXS, RQ, SQ, ES0, BS, mA, mC, QES, QMS = KERNEL[D]
        Q, nsl = D + 1, NEB * D
        Q2 = Q * Q
        NT = ((NEB * Q2 + 31) // 32) * 32
        FS = D * SQ
        GA, GC = max(1, 16 // Q), max(1, 16 // D)
        base = phases(D, Q, NEB, SXS, RQ, SQ, ES0, BS, mA, mC, QES, QMS)
        rows = ("A.w", "C1.r", "C1.w", "C2.r", "B.r", "B.w")
        print(f"p={D - 1} NEB={NEB}: shipped " + " ".join(f"{k}={base[k][0]}" for k in rows) + f" | all phases {sum(v[0] for v in base.values())}")
        for ES in ES_list:
            res = {}

            def add(name, fn, nl):
                L = [fn(l) for l in range(((nl + 31) // 32) * 32)]
                r = res.setdefault(name, 0)
                res[name] = r + sum(cost(L[s0:s0 + NT]) for s0 in range(0, len(L), NT))

            def task(l, G, W):
                h, j = divmod(l, 16)
                slab = G * h + j // W
                return (slab, j % W) if j < G * W and slab < nsl else None

            sE = lambda t, f, a, b: (t // D) * ES + f * FS + (t % D) * SQ + a * RQ + b
            nlA, nlC = ((nsl + GA - 1) // GA) * 16, ((nsl + GC - 1) // GC) * 16
            for f in range(3):
                for qx in range(Q):
                    fn = lambda l: (lambda t: None if t is None else sE(t[0], f, t[1], qx))(task(l, GA, Q))
                    add("A.w", fn, nlA)
                    add("C1.r", fn, nlA)
            for f in range(2):
                for dx in range(D):
                    add("C1.w", lambda l: (lambda t: None if t is None else sE(t[0], f, t[1], dx))(task(l, GA, Q)), nlA)
                for qy in range(Q):
                    add("C2.r", lambda l: (lambda t: None if t is None else sE(t[0], f, qy, t[1]))(task(l, GC, D)), nlC)
            for f in range(3):
                for dz in range(D):
                    fn = lambda l: (l // Q2) * ES + f * FS + dz * SQ + ((l % Q2) // Q) * RQ + (l % Q2) % Q if l < NEB * Q2 else None
                    add("B.r", fn, NEB * Q2)
                    add("B.w", fn, NEB * Q2)
            delta = sum(res[k] - base[k][0] for k in rows)
            print(f"      slabs per half-warp, ES={ES}: " + " ".join(f"{k}={res[k]}" for k in rows) + f" | change {delta:+d} wavefronts per batch")


if __name__ == "__main__":
    if len(sys.argv) >= 4 and sys.argv[1] == "search":
        search(int(sys.argv[2]), int(sys.argv[3]))
    elif len(sys.argv) >= 2 and sys.argv[1] == "alt":
        alt()
    else:
        report()
