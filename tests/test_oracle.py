"""The oracle against the reference's own known-answer vectors and NumPy/SciPy cross-checks (CPU)."""
import numpy as np
import pytest

from dynadjust_b200 import synth


def test_cholesky_inverse_known_answer(oracle):
    # reference tests/test_matrix.cpp:196-222 — 3x3 SPD inverse known answer (tolerance 1e-4 there)
    a = np.array([[4.0, 2.0, 1.0], [2.0, 5.0, 3.0], [1.0, 3.0, 6.0]])
    for use_ref in (False, True):
        inv = oracle.spd_inverse(a, use_ref=use_ref)
        assert np.allclose(inv @ a, np.eye(3), atol=1e-12)
        assert np.allclose(inv, np.linalg.inv(a), atol=1e-12)


def test_cholesky_inverse_packed_equals_full(oracle):
    # reference tests/test_matrix.cpp:1006-1026, 1279-1309 — packed path == full path to 1e-10
    rng = np.random.default_rng(5)
    for n in (1, 2, 7, 33, 130):
        m = rng.standard_normal((n, n))
        a = m @ m.T + n * np.eye(n)
        ref = oracle.spd_inverse(a, use_ref=True)
        port = oracle.spd_inverse(a, use_ref=False)
        assert np.allclose(ref, np.linalg.inv(a), rtol=1e-10, atol=1e-12)
        assert np.allclose(port, ref, rtol=1e-10, atol=1e-12)


def test_cholesky_inverse_singular_throws(oracle):
    # reference tests/test_matrix.cpp:535-568 — singular / indefinite input raises MatrixInversionFailure
    a = np.array([[1.0, 2.0], [2.0, 4.0]])
    b = np.array([[1.0, 2.0], [2.0, 1.0]])
    for m in (a, b):
        for use_ref in (False, True):
            with pytest.raises(np.linalg.LinAlgError):
                oracle.spd_inverse(m, use_ref=use_ref)


def test_geodesy_round_trip(oracle):
    # GEO:78-90 / GEO:154-225 and the worked example in the reference header comment (GEO:143-152):
    # X,Y,Z = (-3563081.362, -2057145.984, -4870449.482) on GRS80 -> -50, -150 (=> 210-360), 10000 m
    llh = oracle.cart_to_geo(-3563081.362, -2057145.984, -4870449.482)
    assert abs(np.degrees(llh[0]) + 50.0) < 1e-7
    assert abs(np.degrees(llh[1]) + 150.0) < 1e-7
    assert abs(llh[2] - 10000.0) < 2e-3
    rng = np.random.default_rng(1)
    for _ in range(50):
        lat, lon, h = np.radians(rng.uniform(-80, 80)), np.radians(rng.uniform(-179, 179)), rng.uniform(-100, 9000)
        xyz = oracle.geo_to_cart(lat, lon, h)
        back = oracle.cart_to_geo(*xyz)
        assert abs(back[0] - lat) < 1e-11 and abs(back[1] - lon) < 1e-11 and abs(back[2] - h) < 2e-5  # Newton stop at |f|<1e-12 (GEO:190) leaves a few um in h
        assert np.allclose(xyz, synth.geo_to_cart(np.array(lat), np.array(lon), np.array(h)), atol=1e-6)


def test_adjustment_against_numpy(oracle):
    """End-to-end oracle run on config C1 versus an independent dense NumPy solve of the same normals."""
    stn, msr, truth, _ = synth.config_network("C1")
    out = oracle.adjust_simultaneous(stn, msr, want_normals=True, want_vcv=True)
    res = out["res"]
    assert res.iterations == 2 and res.converged
    assert res.measurement_params == 900 and res.dof == 900 - (300 - 9)
    n, w = out["normals"], out["rhs"]
    assert np.allclose(n, n.T)
    d = np.linalg.solve(n, w)
    assert np.allclose(d, out["first_corr"], atol=1e-9)
    q = np.linalg.inv(n)
    assert np.abs(q - out["vcv"]).max() / np.abs(q).max() < 1e-9
    # the adjusted network is consistent with the truth at the level of the simulated noise
    assert np.sqrt(((out["est"] - truth) ** 2).mean()) < 0.02
    assert 0.7 < res.sigma_zero < 1.3


def test_ref_and_port_agree(oracle):
    if not oracle.ref_loaded():
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    stn, msr, _, _ = synth.gnss_network(200, 600, 11)
    a = oracle.adjust_simultaneous(stn.copy(), msr.copy(), opts=oracle.default_opts(use_ref=1), want_vcv=True)
    b = oracle.adjust_simultaneous(stn.copy(), msr.copy(), opts=oracle.default_opts(use_ref=0), want_vcv=True)
    assert a["res"].used_ref == 1 and b["res"].used_ref == 0
    assert np.abs(a["est"] - b["est"]).max() < 1e-9
    assert abs(a["res"].sigma_zero - b["res"].sigma_zero) < 1e-12
    assert np.abs(a["vcv"] - b["vcv"]).max() / np.abs(a["vcv"]).max() < 1e-9


def test_variance_scaling_write_back(oracle):
    """vScale / p,l,h scaling is applied once and written back into the records (ADJ:4281)."""
    stn, msr, _, _ = synth.gnss_network(30, 80, 3)
    msr["scale4"] = 4.0
    before = msr["term2"].copy()
    out = oracle.adjust_simultaneous(stn, msr)
    assert np.allclose(msr["term2"], 4.0 * before)
    stn2, msr2, _, _ = synth.gnss_network(30, 80, 3)
    out2 = oracle.adjust_simultaneous(stn2, msr2)
    # scaling every VCV by 4 divides chi^2 by 4 and leaves the estimates unchanged (to solver accuracy)
    assert abs(out["res"].chi_squared * 4.0 - out2["res"].chi_squared) / out2["res"].chi_squared < 1e-6
    assert np.abs(out["est"] - out2["est"]).max() < 1e-5


@pytest.mark.parametrize("kind", ["gnss", "mixed", "terrestrial"])
def test_phased_oracle_matches_simultaneous(oracle, kind):
    """The reference's phased mode restated (forward pass with the junction-station carry ADJ:998-1128, reverse pass
    ADJ:1133-1281, combination ADJ:3196-3333, first-appearance constraints ADJ:1884-2002) against its simultaneous mode
    on the same network: one block is the same arithmetic bit for bit; chains of 3 and 5 blocks agree to the rounding of
    the explicit junction-variance inversions.  This is the reference-side statement of "phased == simultaneous" that
    the engine's supernodal elimination relies on."""
    from dynadjust_b200 import synth_terrestrial
    from tests import parity
    if kind == "gnss":
        stn, msr, _, _ = synth.gnss_network(300, 900, 5)
    elif kind == "mixed":
        stn, msr, _, _ = synth.mixed_network(300, 900, 6, n_distances=200, n_levels=150)
    else:
        stn, msr, _, _ = synth_terrestrial.terrestrial_network(300, 700, 7, scalars={"S": 150, "L": 100}, n_dir_sets=40)
    s1, m1 = stn.copy(), msr.copy()
    ref = oracle.adjust_simultaneous(s1, m1, want_vcv=True)
    V = ref["vcv"]
    qd = np.stack([V[3 * s:3 * s + 3, 3 * s:3 * s + 3] for s in range(len(stn))])
    for width, exact in ((300, True), (100, False), (64, False)):
        s2, m2 = stn.copy(), msr.copy()
        blocks = parity.chain_blocks(len(stn), width)
        ph = oracle.adjust_phased(s2, m2, blocks, want_block=min(1, len(blocks) - 1))
        r1, r2 = ref["res"], ph["res"]
        assert r2.iterations == r1.iterations and r2.converged
        assert (r2.dof, r2.measurement_params, r2.unknown_params, r2.outliers) == \
            (r1.dof, r1.measurement_params, r1.unknown_params, r1.outliers)
        tol = 0.0 if exact else 1.0
        assert np.abs(ph["est"] - ref["est"]).max() <= tol * 2e-9
        assert abs(r2.sigma_zero - r1.sigma_zero) <= tol * 5e-9
        assert np.abs(ph["vcv"] - qd).max() <= tol * 1e-12 * np.abs(qd).max()
        assert abs(r2.global_pelzer - r1.global_pelzer) <= tol * 1e-12
        for f, t in (("measCorr", 5e-9), ("measAdj", 5e-9), ("measAdjPrec", 1e-15), ("NStat", 1e-5), ("term2", 0.0)):
            assert np.abs(m1[f] - m2[f]).max() <= (tol * t if t else 0.0), f
        assert np.abs(s1["currentLatitude"] - s2["currentLatitude"]).max() <= tol * 1e-15
        # the dense rigorous variance matrix the reference keeps per block (v_rigorousVariances_, ADJ:3805)
        bs = ph["block_stations"]
        idx = (3 * bs[:, None] + np.arange(3)[None, :]).ravel()
        assert np.abs(ph["block_vcv"] - V[np.ix_(idx, idx)]).max() <= tol * 1e-12 * np.abs(V).max()



def test_reference_matrix_unit_tests_pass_on_ref_build(oracle):
    """The reference's own tests/test_matrix.cpp (72 cases: constructors, packed / full Cholesky inverse known answers,
    failure modes, scaling, mmap-backed storage) compiled in place against the same matrix_2d sources and BLAS that
    oracle/_ref/libref_matrix.so is built from (oracle/Makefile): the build the oracle inverts with is the build the
    reference's tests accept."""
    import os
    import subprocess
    import sysconfig
    exe = os.path.join(os.path.dirname(oracle.LIB_PATH), "_ref", "test_matrix")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/test_matrix not built (needs /root/reference)")
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = os.path.join(sysconfig.get_paths()["purelib"], "opencv_python_headless.libs") + ":" + env.get("LD_LIBRARY_PATH", "")
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300, env=env)
    assert res.returncode == 0, res.stdout[-2000:]
    tail = res.stdout.strip().splitlines()[-1]
    assert tail.startswith("Total tests: 72") and tail.endswith("Failures: 0"), tail
