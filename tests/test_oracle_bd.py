"""CPU: host logic of BASELINE config 0 against golden vectors generated from the reference's own host code
(tests/golden/gen_golden_bd.cu): the Xorshift128plus restatement (initial positions + integrator seed of the README
example) and the oracle's BD::EulerMaruyama noise."""
import os

import numpy as np

from uammd_b200.synthetic import Xorshift128plus


def _golden(golden_dir):
    return open(os.path.join(golden_dir, "xorshift_bd.bin"), "rb").read()


def test_xorshift128plus_matches_reference(golden_dir):
    buf = _golden(golden_dir)
    off = 0
    for seed in (1234, 0xdeadbeef, None):
        r = Xorshift128plus(seed)
        u = np.frombuffer(buf, np.uint32, 16, off); off += 64
        assert [r.next32() for _ in range(16)] == list(u)
        d = np.frombuffer(buf, np.float64, 48, off); off += 48 * 8
        assert np.array_equal(np.array([r.uniform(-0.5, 0.5) for _ in range(48)]), d)


def test_readme_seed_and_oracle_noise(orc, golden_dir):
    buf = _golden(golden_dir)
    off = 3 * (64 + 48 * 8)
    s3 = np.frombuffer(buf, np.uint32, 3, off); off += 12
    g = np.frombuffer(buf, np.float32, 256, off).reshape(64, 4)
    r = Xorshift128plus(1234)
    for _ in range(100000):
        r.uniform3(-0.5, 0.5)
    assert [r.next32(), r.next32(), r.next32()] == list(s3)
    # one oracle step from the origin with F = 0: the displacement IS the noise (x, y from the first pair, z from the second)
    pos = np.zeros((64, 4))
    M = 1.0 / (6.0 * np.pi)
    orc.bd_euler_maruyama_f64(pos, None, M, 0.1, 1.0, 5, int(s3[2]))
    assert np.array_equal(pos[:, 0].astype(np.float32), g[:, 0])
    assert np.array_equal(pos[:, 1].astype(np.float32), g[:, 1])
    assert np.array_equal(pos[:, 2].astype(np.float32), g[:, 2])
    assert np.all(pos[:, 3] == 0)


def test_oracle_drift_term(orc):
    # T = 0: x += dt (K x + M F) exactly (fma-contracted like the reference kernel)
    rng = np.random.default_rng(1)
    pos = np.zeros((100, 4)); pos[:, :3] = rng.normal(size=(100, 3))
    F = np.zeros((100, 4)); F[:, :3] = rng.normal(size=(100, 3))
    K = rng.normal(size=(3, 3))
    want = pos.copy(); want[:, :3] += 0.01 * (pos[:, :3] @ K.T + 0.3 * F[:, :3])
    orc.bd_euler_maruyama_f64(pos, F, 0.3, 0.01, 0.0, 1, 42, K9=K.ravel())
    assert np.allclose(pos, want, rtol=0, atol=1e-15)
