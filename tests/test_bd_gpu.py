"""GPU: BASELINE config 0 - BD::EulerMaruyama, ideal particles, N = 1e5, fp64 (README.md:82-106).
BIT-EXACT against the unmodified reference compiled from /root/reference (oracle/_ref/ref_bd): same Xorshift128plus
initial positions, same Saru seed drawn at construction, same Saru(i, step, seed) noise stream, 100 steps.
Against the C oracle (host libm) the agreement is to the last ulps of the float noise."""
import json
import os
import subprocess

import numpy as np
import pytest
import torch

from uammd_b200 import bd
from uammd_b200.synthetic import Xorshift128plus, gaussian_forces

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BD = os.path.join(ROOT, "oracle", "_ref", "ref_bd")
N, STEPS, T, ETA, A, DT, SYSSEED = 100_000, 100, 1.0, 1.0, 1.0, 0.1, 1234


def _readme_positions():
    rng = Xorshift128plus(SYSSEED)
    pos = np.zeros((N, 4))
    for i in range(N):
        pos[i, :3] = rng.uniform3(-0.5, 0.5)
    return pos, rng


def _run_ref(tmp_path, extra=()):
    if not os.path.exists(REF_BD):
        pytest.skip("oracle/_ref/ref_bd not built (needs the reference tree at build time)")
    out = str(tmp_path / "bd")
    r = subprocess.run([REF_BD, str(N), str(STEPS), str(T), str(ETA), str(A), str(DT), str(SYSSEED), out, *extra],
                       check=True, capture_output=True, text=True, timeout=600)
    info = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")][-1]
    return (np.fromfile(out + ".pos0.bin", np.float64).reshape(N, 4), np.fromfile(out + ".pos.bin", np.float64).reshape(N, 4),
            info)


def test_readme_example_bit_exact(cuda, tmp_path):
    pos0_ref, pos_ref, info = _run_ref(tmp_path)
    pos0, rng = _readme_positions()
    assert np.array_equal(pos0, pos0_ref), "Xorshift128plus restatement differs from the reference's initial positions"
    sys = bd.System(); sys._rng = rng
    p = torch.from_numpy(pos0).to(cuda)
    integ = bd.EulerMaruyama(p, bd.Parameters(temperature=T, viscosity=ETA, hydrodynamicRadius=A, dt=DT), sys=sys)
    assert integ.seed == info["seed"], "Saru seed drawn at construction differs"
    for _ in range(STEPS):
        integ.forwardTime()
    torch.cuda.synchronize()
    got = p.cpu().numpy()
    assert np.array_equal(got.view(np.uint64), pos_ref.view(np.uint64)), \
        f"positions after {STEPS} steps are not bit-identical: max |d| = {np.abs(got - pos_ref).max()}"


def test_with_forces_bit_exact(cuda, tmp_path):
    pos0, _ = _readme_positions()
    force = np.zeros((N, 4)); force[:, :3] = gaussian_forces(N, seed=3)
    pf, ff = tmp_path / "p.bin", tmp_path / "f.bin"
    pos0.tofile(pf); force.tofile(ff)
    _, pos_ref, info = _run_ref(tmp_path, (str(pf), str(ff)))
    sys = bd.System(SYSSEED)
    p = torch.from_numpy(pos0).to(cuda)
    dforce = torch.from_numpy(force).to(cuda)
    integ = bd.EulerMaruyama(p, bd.Parameters(temperature=T, viscosity=ETA, hydrodynamicRadius=A, dt=DT), sys=sys)
    assert integ.seed == info["seed"]
    integ.addInteractor(lambda f: f.add_(dforce))
    for _ in range(STEPS):
        integ.forwardTime()
    torch.cuda.synchronize()
    assert np.array_equal(p.cpu().numpy().view(np.uint64), pos_ref.view(np.uint64))


def test_against_oracle(orc, cuda):
    pos0, rng = _readme_positions()
    force = np.zeros((N, 4)); force[:, :3] = gaussian_forces(N, seed=3)
    sys = bd.System(); sys._rng = rng
    K = [[0.0, 0.1, 0.0], [0.0, 0.0, 0.0], [0.02, 0.0, 0.0]]
    p = torch.from_numpy(pos0).to(cuda)
    dforce = torch.from_numpy(force).to(cuda)
    integ = bd.EulerMaruyama(p, bd.Parameters(temperature=T, viscosity=ETA, hydrodynamicRadius=A, dt=DT, K=K), sys=sys)
    integ.addInteractor(lambda f: f.add_(dforce))
    ref = pos0.copy()
    for s in range(1, 11):
        integ.forwardTime()
        orc.bd_euler_maruyama_f64(ref, force, integ.selfMobility, DT, T, s, integ.seed, K9=np.array(K).ravel())
    torch.cuda.synchronize()
    B = np.sqrt(2 * T * integ.selfMobility * DT)
    assert np.abs(p.cpu().numpy() - ref).max() < 1e-5 * B * 10  # float libm ulps of the Box-Muller transform


def test_fp32_and_group(cuda):
    # group indirection + fp32: only the selected particles move, and they move like the full update moves them
    pos0, _ = _readme_positions()
    p_all = torch.from_numpy(pos0.astype(np.float32)).to(cuda)
    p_grp = p_all.clone()
    idx = torch.arange(0, N, 3, device=cuda, dtype=torch.int32)
    par = bd.Parameters(temperature=T, viscosity=ETA, hydrodynamicRadius=A, dt=DT)
    a = bd.EulerMaruyama(p_all, par, sys=bd.System(7))
    b = bd.EulerMaruyama(p_grp, par, sys=bd.System(7), groupIndex=idx)
    a.forwardTime(); b.forwardTime()
    torch.cuda.synchronize()
    sel = idx.long()
    assert torch.equal(p_all[sel], p_grp[sel])
    mask = torch.ones(N, dtype=torch.bool, device=cuda); mask[sel] = False
    assert torch.equal(p_grp[mask], torch.from_numpy(pos0.astype(np.float32)).to(cuda)[mask])
