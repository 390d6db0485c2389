"""GPU: the row list behind ub200_verletlist (uammd_b200/csrc/lj_vlist.cu), the list the built-in LJ traversal walks.
Oracle = the reference-layout list of the same handle (itself bit-exact against the compiled reference, test_verlet_gpu.py):
the two must hold the same PAIRS; the forces must be those of the cell-list engine, also after the particles have drifted
inside the skin and across the periodic boundary."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from uammd_b200 import synthetic as syn
from uammd_b200.md import Box, LJ, PairForces, VerletList, _ptr, f3, i3

pytestmark = pytest.mark.gpu


def _liquid(N, Lb, seed=3, jitter=0.25):
    pos = syn.fcc_lattice(N, Lb)
    rng = np.random.default_rng(seed)
    pos[:, :3] += rng.uniform(-jitter, jitter, (N, 3)).astype(np.float32)
    return pos


def _row_pairs(rows, L, periodic):
    """(i, j) in group indices from the row list; checks on the way that the stored image is the minimum image."""
    lst = rows["list"].cpu().numpy().view(np.uint32)
    cnt = rows["count"].cpu().numpy()
    idx = rows["index"].cpu().numpy().astype(np.int64)
    pos = rows["pos"].cpu().numpy().astype(np.float64)
    bits = rows["indexBits"]
    N, stride = lst.shape
    valid = np.arange(stride)[None, :] < cnt[:, None]
    slot_i = np.repeat(np.arange(N), cnt)
    e = lst[valid]
    slot_j = (e & np.uint32((1 << bits) - 1)).astype(np.int64)
    code = (e >> np.uint32(bits)).astype(np.int64)
    shift = np.stack([code % 3 - 1, (code // 3) % 3 - 1, code // 9 - 1], axis=1) * np.asarray(L, np.float64)[None, :]
    d = pos[slot_j, :3] + shift - pos[slot_i, :3]
    per = np.asarray(periodic, bool)
    mic = d - np.where(per, np.asarray(L) * np.round(d / np.asarray(L)), 0.0)
    assert np.abs(d - mic).max() < 1e-3, "a stored image is not the minimum image"
    assert (slot_i != slot_j).all(), "self must not be listed"
    return idx[slot_i], idx[slot_j], np.sqrt((d * d).sum(axis=1))


def _reference_pairs(d):
    nl = d["neighbourList"].cpu().numpy()
    nn = d["numberNeighbours"].cpu().numpy()
    gi = d["groupIndex"].cpu().numpy().astype(np.int64)
    N = nn.shape[0]
    maxk = int(nn.max())
    valid = np.arange(maxk)[:, None] < nn[None, :]
    si = np.broadcast_to(np.arange(N)[None, :], (maxk, N))[valid]
    sj = nl[:maxk][valid].astype(np.int64)
    keep = si != sj
    return gi[si[keep]], gi[sj[keep]]


def _check_same_pairs(nl, N, L, periodic, rlist):
    rows = nl.getRowList()
    ri, rj, rd = _row_pairs(rows, L, periodic)
    fi, fj = _reference_pairs(nl.getVerletList())
    a = set((ri * N + rj).tolist())
    b = set((fi * N + fj).tolist())
    # the only pairs allowed to differ sit within rounding of the list radius (the row list keeps them on purpose)
    far = {int(k): float(r) for k, r in zip(ri * N + rj, rd)}
    extra = a - b
    assert not (b - a), f"{len(b - a)} pairs of the reference list are missing"
    assert all(abs(far[k] - rlist) < 2e-5 * rlist for k in extra), f"{len(extra)} unexpected pairs"
    assert len(a) == len(ri), "duplicate entries"
    return rows


@pytest.mark.parametrize("case", ["cubic", "slab", "thin"])
def test_row_list_holds_the_pairs_of_the_reference_list(cuda, case):
    if case == "cubic":
        N = 4 * 14 ** 3
        Lb = syn.lj_box_length(N)
        L, periodic, pos = (Lb, Lb, Lb), (1, 1, 1), _liquid(N, Lb)
    elif case == "slab":   # non periodic z, uneven box
        N = 20000
        L, periodic = (31.0, 27.5, 36.0), (1, 1, 0)
        pos = syn.uniform_cloud(N, 1.0, seed=5)
        pos[:, :3] = pos[:, :3] * np.asarray(L, np.float32)[None, :] * 0.999
    else:                  # exactly five half cells across y: every row piece of the stencil is a different image
        N = 9000
        L, periodic = (40.0, 7.0, 33.0), (1, 1, 1)
        pos = syn.uniform_cloud(N, 1.0, seed=6)
        pos[:, :3] = pos[:, :3] * np.asarray(L, np.float32)[None, :]
    nl = VerletList()
    box = Box(L)
    box.setPeriodicity(*periodic)
    assert nl.update(torch.from_numpy(pos).to(cuda), box, 2.5)
    _check_same_pairs(nl, N, L, periodic, 2.5 * 1.08)


def test_dense_cloud_grows_the_rows_and_takes_the_unstaged_fill(cuda):
    N = 20000
    Lb = float(np.float32((N / 2.5) ** (1.0 / 3.0)))
    pos = syn.uniform_cloud(N, Lb, seed=77)
    nl = VerletList()
    nl.update(torch.from_numpy(pos).to(cuda), Box(Lb), 2.5)
    rows = _check_same_pairs(nl, N, (Lb,) * 3, (1, 1, 1), 2.5 * 1.08)
    assert int(rows["count"].max()) > 150 and rows["stride"] >= int(rows["count"].max())


def _engine_forces(pot, box, pos, N, ev=False):
    f = torch.zeros(N, 4, device=pos.device)
    e = torch.zeros(N, device=pos.device) if ev else None
    v = torch.zeros(N, device=pos.device) if ev else None
    PairForces(pot, box).sum(pos, f, energy=e, virial=v)
    return f, e, v


@pytest.mark.parametrize("ntypes", [1, 2])
def test_forces_energy_virial_match_the_cell_list_engine(cuda, ntypes):
    N = 4 * 16 ** 3
    Lb = syn.lj_box_length(N)
    posh = _liquid(N, Lb, seed=9)
    pot = LJ()
    pot.setPotParameters(0, 0, cutOff=2.5)
    if ntypes > 1:
        posh[::3, 3] = 1.0
        pot.setPotParameters(0, 1, sigma=0.9, epsilon=1.3, cutOff=2.2)
        pot.setPotParameters(1, 1, sigma=1.1, epsilon=0.7, cutOff=2.5, shift=True)
    pos = torch.from_numpy(posh).to(cuda)
    box = Box(Lb)
    nl = VerletList()
    pf = PairForces(pot, box, nl=nl)
    f = torch.zeros(N, 4, device=cuda)
    pf.sum(pos, f)
    f0, _, _ = _engine_forces(pot, box, pos, N)
    scale = f0[:, :3].abs().max().item()
    assert (f - f0)[:, :3].abs().max().item() < 2e-5 * scale
    f = torch.zeros(N, 4, device=cuda); e = torch.zeros(N, device=cuda); v = torch.zeros(N, device=cuda)
    pf.sum(pos, f, energy=e, virial=v)
    f1, e1, v1 = _engine_forces(pot, box, pos, N, ev=True)
    assert (f - f1)[:, :3].abs().max().item() < 2e-5 * scale
    assert (e - e1).abs().max().item() < 2e-5 * e1.abs().max().item()
    assert (v - v1).abs().max().item() < 2e-5 * v1.abs().max().item()


def test_list_stays_valid_while_particles_drift_inside_the_skin_and_across_the_boundary(cuda):
    N = 4 * 16 ** 3
    Lb = syn.lj_box_length(N)
    posh = _liquid(N, Lb, seed=11)
    pot = LJ(); pot.setPotParameters(0, 0, cutOff=2.5)
    box = Box(Lb)
    nl = VerletList()
    pf = PairForces(pot, box, nl=nl)
    pos = torch.from_numpy(posh).to(cuda)
    f = torch.zeros(N, 4, device=cuda)
    pf.sum(pos, f)
    assert nl.rebuilds() == 1
    thr = (1.08 * 2.5 - 2.5) / 2
    rng = np.random.default_rng(4)
    u = rng.normal(size=(N, 3)); u /= np.linalg.norm(u, axis=1)[:, None]
    moved = posh.copy()
    moved[:, :3] += (u * rng.uniform(0.0, 0.98 * thr, (N, 1))).astype(np.float32)
    moved[::7, 0] += np.float32(Lb)          # a caller that wraps (or unwraps) coordinates between rebuilds
    moved[::11, 2] -= np.float32(2 * Lb)
    pos2 = torch.from_numpy(moved).to(cuda)
    f.zero_()
    pf.sum(pos2, f)
    assert nl.rebuilds() == 1, "the drift stayed under the threshold: no rebuild expected"
    f0, _, _ = _engine_forces(pot, box, pos2, N)
    assert (f - f0)[:, :3].abs().max().item() < 2e-5 * f0[:, :3].abs().max().item()


def test_group_indirection_and_reference_layout_switch(cuda):
    """A ParticleGroup (index list) as the reference's PairForces passes it, and UB200_VERLET_FAST=0 (reference-layout list
    only) giving the same forces."""
    from uammd_b200 import _lib
    lib = _lib.lib()
    Ntot = 4 * 14 ** 3
    Lb = syn.lj_box_length(Ntot)
    posh = _liquid(Ntot, Lb, seed=13)
    rng = np.random.default_rng(8)
    group = np.sort(rng.choice(Ntot, Ntot // 2, replace=False)).astype(np.int32)
    pos = torch.from_numpy(posh).to(cuda)
    gidx = torch.from_numpy(group).to(cuda)
    pot = LJ(); pot.setPotParameters(0, 0, cutOff=2.5)
    tab = pot.table()
    out = []
    for fast in ("1", "0"):
        os.environ["UB200_VERLET_FAST"] = fast
        try:
            nl = VerletList()
            nl.update(pos, Box(Lb), 2.5, groupIndex=gidx)
            f = torch.zeros(Ntot, 4, device=cuda)
            rc = lib.ub200_lj_sum_verlet_f32(nl._h, tab.ctypes.data_as(C.POINTER(C.c_float)), 1, _ptr(f), None, None, _ptr(gidx), None)
            assert rc == 0
            torch.cuda.synchronize()
            if fast == "1":
                nl.getRowList()
            else:
                with pytest.raises(Exception):
                    nl.getRowList()
            out.append(f)
        finally:
            os.environ.pop("UB200_VERLET_FAST", None)
    sub = torch.from_numpy(posh[group]).to(cuda)
    f0, _, _ = _engine_forces(pot, Box(Lb), sub, group.shape[0])
    scale = f0[:, :3].abs().max().item()
    for f in out:
        assert (f[gidx.long(), :3] - f0[:, :3]).abs().max().item() < 2e-5 * scale
        mask = torch.ones(Ntot, dtype=torch.bool, device=cuda); mask[gidx.long()] = False
        assert f[mask].abs().max().item() == 0.0
