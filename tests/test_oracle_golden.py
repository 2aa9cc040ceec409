"""Pins the CPU oracle against the reference's own fixtures.

tests/golden/<case>/ holds verbatim copies of the reference's tests/<case>/{entrada.ini,movedor.ini,
chunk.xyz,ref.xyz}.  ref.xyz is the last Li.xyz frame written by the reference binary; the oracle must
reproduce every coordinate bit-for-bit (the 17-digit list-directed output round-trips doubles exactly).
"""
import os
import numpy as np
import pytest
from oracle import oracle as O


def test_rng_known_answers():
    # first ran()/gasdev() values for seed -104012 (SURVEY.md §8c bring-up checkpoints)
    r, g = O.rng_kat(-104012, 3, 4)
    assert list(r) == [0.662347872220825, 0.06550750168065857, 0.7375988305196832]
    assert list(g) == [0.1918187457932713, -0.5133655584400709, 0.4613120242960044, 0.6784472874214147]


CHECK = {
    # case: (final N, nupd_vlist, choques, ran calls) — SURVEY.md §8c
    "brown": (1505, 1004, 21470, 2422993),
    "gcmc": (594, 1623, 15148, 10953828),
    "ermak": (1204, 451, 49905, 75662156),
}


@pytest.mark.parametrize("case", ["brown", "gcmc", "ermak"])
def test_reference_fixture_bit_exact(case, golden_dir):
    d = O.read_case(os.path.join(golden_dir, case))
    o = O.Oracle(**d)
    o.step(d["nst"])
    zmax, z, pos, _ = o.frame()
    rzmax, rz, rpos = O.read_xyz_frame(os.path.join(golden_dir, case, "ref.xyz"))
    assert len(z) == len(rz) == CHECK[case][0]
    assert zmax == rzmax
    assert np.array_equal(z, rz)
    assert np.array_equal(pos, rpos)          # bit-for-bit
    s = o.scalars()
    assert (s.nupd, s.choques, s.ran_calls) == CHECK[case][1:]


def test_ermak_constants():
    # values at h=0.01, gamma=1, T=300
    o = O.from_case(os.path.join(os.path.dirname(__file__), "golden", "ermak"))
    s = o.scalars()
    assert s.cc0 == 0.9900498337491681
    assert s.cc1 == 0.009950166250831893
    assert s.cc2 == 0.004983374916810668
    assert s.sdr == 8.134432475981699e-4
    assert s.sdv == 0.14071718691490614
    assert s.crv1 == 0.8649405763640867
    assert s.crv2 == 0.501874286409417
    assert s.skt == 15.793445971222525


def test_fast_pos_inic_same_stream():
    # the cell-accelerated generator consumes the same RNG stream and yields the same configuration
    a, ca = O.pos_inic(-104012, 100.0, 100.0, 200.0, fast=False)
    b, cb = O.pos_inic(-104012, 100.0, 100.0, 200.0, fast=True)
    assert ca == cb and np.array_equal(a, b) and len(a) == 1204


def test_chunk_fixture_is_pos_inic_rule():
    # the shipped chunk.xyz is the pos_inic rule at 100x100x50 up to its text precision
    ch = O.read_chunk(os.path.join(os.path.dirname(__file__), "golden", "brown", "chunk.xyz"))
    assert ch.shape == (301, 3)


def test_logf_restatement_matches_libm(tmp_path):
    """DML_RNG_REFERENCE draws gasdev on the device, whose single-precision logarithm must be the host libm's logf bit for bit
    (src/dana.F90:1379-1404: rsq is real(sp)).  din_mol_li_b200/csrc/dml_device.cuh::ref_logf restates glibc's algorithm; the same
    restatement in C (tests/ref_logf_check.c, same table and polynomial) is compared here with logf over EVERY float in (0, 1]."""
    import subprocess
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_logf_check.c")
    exe = os.path.join(tmp_path, "ref_logf_check")
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-o", exe, src, "-lm"])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "0 differ" in r.stdout, r.stdout
    # the constants of the device function are the ones of the C file
    dev = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "din_mol_li_b200", "csrc", "dml_device.cuh")).read()
    for tok in ("0x1.661ec79f8f3bep+0", "-0x1.57bf7808caadep-2", "0x1.767dcf5534862p-1", "0x1.4043057b6ee09p-2", "-0x1.00ea348b88334p-2",
                "0x1.5575b0be00b6ap-2", "-0x1.ffffef20a4123p-2", "0x1.62e42fefa39efp-1"):
        assert tok in dev and tok in open(src).read()
