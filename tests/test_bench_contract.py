"""CPU: bench.py's reference arm (the one leg that may execute oracle/_ref) prints the contract's JSON line,
and the algorithmic-bytes table of SURVEY.md §8(d) is what bench.py uses for the roofline."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT


def test_algorithmic_bytes_match_survey():
    sys.path.insert(0, ROOT)
    import bench
    # diff+mass / diff-only bytes per dof for p = 1..6
    expect = {1: (1596, 1380), 2: (495, 431), 3: (298, 261), 4: (225, 198), 5: (188, 166), 6: (165, 147)}
    for p, (both, diff) in expect.items():
        assert abs(bench.algorithmic_bytes_per_dof(p, 7)[0] - both) <= 1.0   # the survey rounds to whole bytes
        assert abs(bench.algorithmic_bytes_per_dof(p, 6)[0] - diff) <= 1.0
        tot, elem = bench.algorithmic_bytes_per_dof(p, 7)
        assert elem == tot - 12


def test_reference_arm_json_line():
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ref_driver")):
        pytest.skip("oracle/_ref/ref_driver not built")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--elems", "12", "--steps", "2",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GDOF/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("GDOF/s of FP64 PA diffusion+mass apply")
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "GDOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["dtype"] == "f64"


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True,
                         text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_shared_memory_layout_model_of_the_kernel():
    """tools/smem_strides.py holds the bank-conflict model the per-order layouts of pa_apply_kernel were chosen with;
    the configuration compiled into the kernel must stay conflict-free in every phase over the work array at p=2
    and p=3 and must agree with ApplyCfg in the source"""
    import re
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import smem_strides as ss
    for D in (3, 4):
        NEB, SXS, RQ, SQ, ES, BS, mA, mC, QES, QMS = ss.KERNEL[D]
        r = ss.phases(D, D + 1, NEB, SXS, RQ, SQ, ES, BS, mA, mC, QES, QMS)
        for ph in ("A.w", "B.r", "B.w", "C1.r", "C1.w", "C2.r"):
            assert r[ph][0] == r[ph][1], (D, ph, r[ph])
    src = open(os.path.join(ROOT, "cardiac-ablation-ecm2_b200", "csrc", "pa_apply_kernel.cuh")).read()
    new = src[src.index("#else", src.index("B200PA_TUNE_LAYOUT0")):]
    for name, col in (("SXS", 1), ("SQ", 3), ("ES", 4)):
        m = re.search(r"static constexpr int %s = ([^;]+);" % name, new)
        vals = [int(v) for v in re.findall(r"\? (\d+)", m.group(1))] + [int(re.findall(r": (\d+)$", m.group(1).strip())[0])]
        assert vals == [ss.KERNEL[D][col] for D in (2, 3, 4, 5, 6, 7)], (name, vals)


def test_grouped_lane_maps_gain_little(capsys):
    """tools/smem_strides.py alt: giving every half-warp whole slabs in the row phases removes at most ~7 % of a batch's
    shared-memory wavefronts at p = 4 and nothing at p = 3, 5 - the evidence DESIGN.md 4.1 cites for keeping the packed maps"""
    import re
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import smem_strides as ss
    ss.alt()
    out = capsys.readouterr().out
    best = {}
    cur = None
    for line in out.splitlines():
        m = re.match(r"p=(\d) NEB=\d+: shipped .* all phases (\d+)", line)
        if m:
            cur = int(m.group(1))
            best[cur] = [int(m.group(2)), 0]
        m = re.search(r"change ([+-]\d+) wavefronts", line)
        if m:
            best[cur][1] = min(best[cur][1], int(m.group(1)))
    assert set(best) == {3, 4, 5}
    assert best[3][1] == 0 and best[5][1] == 0                      # only worse there
    assert 0.05 <= -best[4][1] / best[4][0] <= 0.08                 # 102 of 1378 wavefronts at p = 4
