"""GPU: drop-in integration. UAMMD programs (reference headers, ParticleData, VerletNVE, BDHI::EulerMaruyama) compiled
in the build container with only the hot-path module swapped for the glue classes of include/uammd_b200/uammd_b200.cuh
(examples/dropin_lj.cu, examples/dropin_fcm.cu -> oracle/_ref/dropin_*). Skipped when the binaries were not built."""
import json
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(name, *args):
    exe = os.path.join(ROOT, "oracle", "_ref", name)
    if not os.path.exists(exe):
        pytest.skip(f"oracle/_ref/{name} not built (needs the reference tree at build time)")
    out = subprocess.run([exe, *map(str, args)], check=True, capture_output=True, text=True, timeout=900).stdout
    return json.loads([l for l in out.splitlines() if l.startswith("{")][-1])


def test_pairforces_dropin_matches_reference():
    r = _run("dropin_lj", 200000, 63.0)
    print(r)
    assert r["celllist_mismatches"] == 0                 # CellListData bit identical to the reference's
    assert r["generic_vs_ref"] == 0.0                    # reference kernel + user functor on our list: same bits
    assert r["fast_vs_ref"] < 1e-4                       # specialised traversal: fp32 rounding only (units of max |F|)
    # the reference's LJ functor through the generic-Transverser column traversal (b200::ColumnList): force, energy, virial
    assert r["column_generic_vs_ref"] < 1e-4 and r["column_energy_vs_ref"] < 1e-4 and r["column_virial_vs_ref"] < 1e-4
    assert r["nve20_max_dpos"] < 1e-4                    # 20 VerletNVE steps next to the reference


def test_nbody_fallback_dropin_matches_reference():
    """Box <= 3 cut-offs: the reference's PairForces takes NBody::transverse (PairForces.cu:49-53); ours ub200_lj_nbody_f32."""
    r = _run("dropin_nbody", 300, 7.0)
    print(r)
    assert r["nbody"] == 1
    assert r["force_vs_ref"] < 1e-5 and r["energy_vs_ref"] < 1e-5 and r["virial_vs_ref"] < 1e-5


@pytest.mark.xfail(strict=False, reason="first execution pending (round 2 GPU budget spent)")
@pytest.mark.parametrize("N,shift", [(4000, 1), (50000, 0)])
def test_double_precision_pairforces_dropin_matches_reference(N, shift):
    """A -DDOUBLE_PRECISION UAMMD program: the stock PairForces<LJ> (cell list and transverser in double) next to an
    Interactor over ub200_lj_sum_f64. Same pair terms in double, another summation order: flat 1e-12 of the largest value.
    (The program is its own process: the newest kernel cannot disturb the CUDA context of this session.)"""
    r = _run("dropin_lj64", N, shift)
    print(r)
    assert r["fmax"] > 1.0
    assert r["force_vs_ref"] < 1e-12 and r["energy_vs_ref"] < 1e-12 and r["virial_vs_ref"] < 1e-12


def test_langevin_verlet_dropin_matches_reference():
    """benchmark.cu's configuration: VerletNVT::GronbechJensen + PairForces<LJ, VerletList> with both modules swapped."""
    r = _run("dropin_nvt", 32768, 38.0, 20)
    print(r)
    assert r["ideal_mismatch_words"] == 0                # b200::VerletNVTGronbechJensen alone: the reference's bits
    assert r["basic_mismatch_words"] == 0                # b200::VerletNVTBasic alone: the bits of VerletNVT::Basic
    assert r["lj_max_dpos"] < 1e-4 and r["lj_max_dvel"] < 1e-2   # with the LJ forces: fp32 summation order only


def test_poisson_dropin_matches_reference():
    """b200::Poisson next to the unmodified reference Poisson (double precision): the reference test's analytic known answer
    for both, then forces / energies / field / potential of a neutral cloud of 2000 charges against each other."""
    r = _run("dropin_poisson", 2000)
    print(r)
    assert r["kat_reference"] < 1e-3 and r["kat_ours"] < 1e-3 and r["kat_field_ours"] < 1e-3   # test_poisson.cu:206,217
    # same algorithm, same tables; the FFTs differ (hand written vs cuFFT) and so does the order of the atomic spreading
    assert r["force_vs_ref"] < 1e-9 and r["energy_vs_ref"] < 1e-9
    assert r["field_vs_ref"] < 1e-9 and r["potential_vs_ref"] < 1e-9


def test_fcm_dropin_matches_reference():
    r = _run("dropin_fcm", 20000, 64)
    print(r)
    assert r["max_dpos_vs_reference"] < 1e-12
    assert abs(r["a"] - 1.0) < 1e-12


def test_verlet_bd_pse_dropin_matches_reference():
    r = _run("dropin_more", 50000, 40.0)
    print(r)
    assert r["verlet_generic_vs_ref"] == 0.0             # reference traversal + functor over OUR Verlet list: same bits
    assert r["verlet_fast_vs_ref"] < 1e-4                # specialised LJ traversal over the list (units of max |F|)
    assert r["bd_max_dpos"] == 0.0                       # b200::BDEulerMaruyama: bit-identical positions after 50 steps
    assert r["bd_force_max_dpos"] == 0.0                 # ... also with an interactor
    assert r["pse_displacement"] > 1e-3                  # the particles did move
    assert r["pse_max_dpos"] < 2e-5 * max(1.0, r["pse_displacement"])  # BDHI::EulerMaruyama<b200::PSE>, fp32


def test_dpd_dropin_matches_reference_and_thermostats():
    """b200::DPDPotential lets the reference's PairForces drive its own DPD transverser; b200::PairForcesDPD is the fast
    path. Same Saru seed and step: forces agree to fp32 summation order. Then 2000 VerletNVE steps of the DPD fluid at
    rho = 3 from a random cloud: the kinetic temperature must settle at kT = 1 (fluctuation-dissipation: a wrong sign or
    scale of the dissipative / random force would pass force parity against our own restatement but not this)."""
    r = _run("dropin_dpd", 81000, 2000)
    print(r)
    assert r["fast_vs_ref"] < 2e-5
    assert r["column_generic_vs_ref"] < 2e-5                # general transverser (getInfo) through b200::ColumnList
    assert abs(r["kT_measured"] - r["kT_target"]) < 0.03


def test_fcm_impl_and_ibm_glue_pass_the_reference_tests():
    """test/BDHI/FCM/fcm_test.cu:85-144 and test/misc/ibm/test_ibm_regular.cu:113-136,156-214,240-274 re-hosted on
    b200::FCM_impl<Gaussian, GaussianTorque> and b200::IBM<Peskin::threePoint> (double precision build)."""
    r = _run("dropin_fcm_impl")
    print(r)
    assert r["worst_vs_hasimoto"] < r["tolerance"]          # the reference's own assertion (1e-8)
    assert r["max_vs_reference_fcm_impl"] < 1e-10           # next to the reference's FCM_impl on the same inputs
    assert r["ibm_spread_vs_manual"] < 1e-10 and r["ibm_gather_vs_manual"] < 1e-10
    assert r["ibm_spread_vs_reference"] < 1e-12 and r["ibm_adjointness"] < 1e-10
