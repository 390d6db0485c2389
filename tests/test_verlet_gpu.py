"""GPU: Verlet (skin) list. Bit-exact list parity against the UNMODIFIED reference compiled from /root/reference
(oracle/_ref/ref_lj_verlet: numberNeighbours, neighbourList [k*N+i], groupIndex), LJ forces over the list within the fp32
tolerance model, rebuild policy (drift threshold, forced rebuild, list growth) and a trajectory against the cell-list engine."""
import json
import os
import subprocess

import numpy as np
import pytest
import torch

from uammd_b200 import synthetic as syn
from uammd_b200.md import Box, LJ, LJMD, PairForces, VerletList, VerletNVE

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "ref_lj_verlet")


def _run_ref(tmp_path, pos, L, rc=2.5):
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/ref_lj_verlet not built (needs the reference tree at build time)")
    pf = tmp_path / "pos.bin"
    pos.tofile(pf)
    out = str(tmp_path / "ref")
    r = subprocess.run([REF, "forces", str(pos.shape[0]), str(L[0]), str(L[1]), str(L[2]), str(rc), "1", "1", str(pf), out],
                       check=True, capture_output=True, text=True, timeout=600)
    info = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")][-1]
    N = pos.shape[0]
    return {"nn": np.fromfile(out + ".nn.bin", np.int32), "index": np.fromfile(out + ".index.bin", np.int32),
            "list": np.fromfile(out + ".list.bin", np.int32).reshape(info["maxk"], N),
            "force": np.fromfile(out + ".force.bin", np.float32).reshape(N, 4), "info": info}


@pytest.mark.parametrize("N,kind", [(20000, "uniform"), (1_000_000, "fcc")])
def test_list_bit_exact_vs_reference(orc, cuda, tmp_path, N, kind):
    Lb = syn.lj_box_length(N)
    pos = syn.uniform_cloud(N, Lb, seed=2024) if kind == "uniform" else syn.fcc_lattice(N, Lb)
    if kind == "fcc":
        pos[:, :3] += np.random.default_rng(5).normal(0, 0.05, (N, 3)).astype(np.float32)
    L = (Lb,) * 3
    ref = _run_ref(tmp_path, pos, L)
    pot = LJ(); pot.setPotParameters(0, 0, cutOff=2.5)
    nl = VerletList()
    pf = PairForces(pot, Box(L), nl=nl)
    dpos = torch.from_numpy(pos).to(cuda)
    force = torch.zeros(N, 4, device=cuda)
    pf.sum(dpos, force=force)
    torch.cuda.synchronize()
    d = nl.getVerletList()
    assert d["particleStride"] == ref["info"]["stride"] == N
    nn = d["numberNeighbours"].cpu().numpy()
    assert np.array_equal(d["groupIndex"].cpu().numpy(), ref["index"])
    assert np.array_equal(nn, ref["nn"])
    maxk = int(nn.max())
    mine = d["neighbourList"].cpu().numpy()[:maxk]
    valid = np.arange(maxk)[:, None] < nn[None, :]      # entries beyond numberNeighbours[i] are undefined in both
    assert np.array_equal(mine[valid], ref["list"][valid])
    # forces over the list: both implementations against the fp64 truth (same tolerance model as the cell-list path)
    g = orc.make_grid_f(L, orc.neighbour_celldim(L, 2.5))
    ocl = orc.celllist_build(g, pos)
    f64, _, _, sc = orc.lj_f64(g, ocl, pot.table(), 1, N)
    tol = sc.force_tol(L, 2.5)
    err_new = (np.abs(force.cpu().numpy()[:, :3] - f64).max(axis=1) / tol).max()
    err_ref = (np.abs(ref["force"][:, :3] - f64).max(axis=1) / tol).max()
    print(f"[verlet N={N} {kind}] maxk={maxk} mean neighbours {nn.mean():.1f}; force error new {err_new:.3f} ref {err_ref:.3f} x fp32 tolerance")
    assert err_new < 1.0 and err_ref < 1.0


def test_rebuild_policy_and_growth(cuda):
    N = 4 * 12 ** 3
    Lb = syn.lj_box_length(N)
    pos = torch.from_numpy(syn.fcc_lattice(N, Lb)).to(cuda)
    box = Box(Lb)
    nl = VerletList()
    assert nl.update(pos, box, 2.5) is True            # first call always builds
    assert nl.update(pos, box, 2.5) is False           # nothing moved
    assert nl.getNumberOfStepsSinceLastUpdate() == 1
    thr = (1.08 * 2.5 - 2.5) / 2
    p2 = pos.clone(); p2[7, 0] += 0.9 * thr
    assert nl.update(p2, box, 2.5) is False            # below the drift threshold
    p2[7, 0] += 0.2 * thr
    assert nl.update(p2, box, 2.5) is True             # above it
    assert nl.update(p2, box, 2.6) is True             # cut-off changed
    assert nl.update(p2, Box(Lb * 1.01), 2.6) is True  # box changed
    nl.forceNextUpdate = True
    assert nl.update(p2, Box(Lb * 1.01), 2.6) is True  # pos-write / reorder signal
    # a drift across the periodic boundary is measured through the minimum image
    p3 = p2.clone(); p3[:, 1] += Lb * 1.01
    assert nl.update(p3, Box(Lb * 1.01), 2.6) is False
    # growth: rho = 0.8, rc*1.08 = 2.7 -> ~66 neighbours + self > 64: the list must have grown from 32 in steps of 32
    d = nl.getVerletList()
    assert d["maxNeighboursPerParticle"] % 32 == 0 and d["maxNeighboursPerParticle"] > int(d["numberNeighbours"].max())


def test_md_trajectory_matches_cell_list_engine(cuda):
    # VerletNVE + PairForces<LJ, VerletList> vs the fused cell-list engine: same physics, different summation order
    N, steps, dt = 4 * 16 ** 3, 60, 0.005
    Lb = syn.lj_box_length(N)
    pos0 = syn.fcc_lattice(N, Lb); vel0 = syn.maxwell_velocities(N, 1.0)
    pot = LJ(); pot.setPotParameters(0, 0, cutOff=2.5)
    p = torch.from_numpy(pos0).to(cuda); v = torch.from_numpy(vel0).to(cuda)
    nl = VerletList()
    integ = VerletNVE(p, v, dt)
    integ.addInteractor(PairForces(pot, Box(Lb), nl=nl))
    for _ in range(steps):
        integ.forwardTime()
    p2 = torch.from_numpy(pos0).to(cuda); v2 = torch.from_numpy(vel0).to(cuda); f2 = torch.zeros(N, 4, device=cuda)
    LJMD(Box(Lb), pot, dt).run(p2, v2, f2, steps)
    torch.cuda.synchronize()
    rebuilds = nl.getVerletList()["rebuilds"]
    assert 2 <= rebuilds < steps, rebuilds             # the skin is actually used
    assert (p - p2).abs().max().item() < 2e-3          # chaotic divergence of fp32 round-off over 60 steps stays tiny


def test_dense_cloud_takes_the_unstaged_path(orc, cuda, tmp_path):
    """rho = 2.5: ~1400 candidates per home cell (beyond the 640 staged per warp) and ~210 neighbours per particle (beyond the
    96 assembled in shared memory): the global-memory walk and the long-list stores must give the same bits."""
    N = 20000
    Lb = float(np.float32((N / 2.5) ** (1.0 / 3.0)))
    pos = syn.uniform_cloud(N, Lb, seed=77)
    L = (Lb,) * 3
    ref = _run_ref(tmp_path, pos, L)
    pot = LJ(); pot.setPotParameters(0, 0, cutOff=2.5)
    nl = VerletList()
    nl.update(torch.from_numpy(pos).to(cuda), Box(L), 2.5)
    d = nl.getVerletList()
    nn = d["numberNeighbours"].cpu().numpy()
    assert np.array_equal(nn, ref["nn"]) and nn.max() > 150
    maxk = int(nn.max())
    valid = np.arange(maxk)[:, None] < nn[None, :]
    assert np.array_equal(d["neighbourList"].cpu().numpy()[:maxk][valid], ref["list"][valid])
