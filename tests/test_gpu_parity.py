"""GPU parity tests proper: the CUDA path (through the C ABI of libdml.so) against the CPU oracle on the same
inputs.  Bit-exact for cells, neighbour rows, integrator/overlap/reservoir state; forces bit-exact in strict
order and within 1e-12 relative in row order (BASELINE.json north_star)."""
import os
import sys
import numpy as np
import pytest
from oracle import oracle as O
from din_mol_li_b200 import dml
import parity as P

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def case(name, **over):
    d = O.read_case(os.path.join(GOLD, name))
    d.update(over)
    return d, O.Oracle(**d)


@pytest.mark.parametrize("name", ["ermak", "brown", "gcmc"])
def test_cells_and_rows_bit_exact_t0(name):
    d, o = case(name)
    ctx = P.ctx_from_oracle(o)
    ctx.test_update()
    a = P.oracle_slot_arrays(o)
    n = len(a["z"])
    # cell triples and in-cell chain order (Cells.F90:267-302)
    ocell, ochain = o.cells()
    gcell, gchain = ctx.cells(n)
    sb = a["slot_b"]
    assert np.array_equal(gcell[a["alive"]], ocell[sb[a["alive"]]])
    assert np.array_equal(gchain[a["alive"]], ochain[sb[a["alive"]]])
    # rows: same entries in the same order (Neighbor.F90:465-548)
    nn, rows, _ = o.rows(width=64)
    gnn, grows = ctx.neighbors(n, width=64)
    assert np.array_equal(gnn, nn[:n])
    assert P.rows_as_lists(gnn, grows, 0) == P.rows_as_lists(nn, rows, 1)
    assert ctx.counters().list_entries == int(nn.sum())


@pytest.mark.parametrize("strict", [1, 0])
def test_fuerza_matches_oracle(strict):
    d, o = case("ermak")
    o.step(150)                       # deposits exist, some pairs are inside the cut-off after ermak_a
    o.call(O.ERMAK_A)
    ctx = P.ctx_from_oracle(o, strict=strict)
    P.push_rows(o, ctx)
    o.call(O.FUERZA)
    ctx.fuerza()
    dd, a = P.compare_state(o, ctx, fields=("pos",), force=True, rtol=0.0 if strict else 1e-12, what="fuerza")
    ref = a["alive"] & ((a["flags"] & 1) > 0)
    assert (np.abs(a["force"][ref]).sum(axis=1) > 0).sum() > 5      # the comparison is not vacuous


@pytest.mark.parametrize("name,nsteps", [("ermak", 60), ("brown", 40), ("gcmc", 400)])
def test_lockstep_replay_bit_exact(name, nsteps):
    d, o = case(name)
    ls = P.Lockstep(o, strict=1, chunk_xyz=d.get("chunk_xyz"))
    for i in range(nsteps):
        ls.step(check=True, tag="%s step %d" % (name, i + 1))


@pytest.mark.parametrize("name,nsteps,over", [
    ("ermak", 400, dict(prob=0.5)), ("brown", 60, dict(prob=0.5)), ("brown", 60, dict(dif_sei=50.0)),
    ("brown", 120, dict(prob=0.3, dif_sei=20.0)), ("gcmc", 200, dict(prob=0.5, dif_sei=50.0))])
@pytest.mark.parametrize("coop", [1, 0])
def test_lockstep_replay_reference_branches(name, nsteps, over, coop, monkeypatch):
    """Branches of the reference that none of its three fixtures takes (all have prob = 1.0 and dif_sei = dif_sc): a deposition
    attempt that fails with probability 1 - prob (src/dana.F90:898-911 in overlap_moveback, 1236-1240 in atom_pbc: the atom goes back
    to old_cg, skip = .false., and draws again at its next visit) and the slower diffusion below z = 80 (src/dana.F90:817-821).
    Lock-step against the oracle with its random numbers injected (per-slot queues for the repeated draws), both launch forms."""
    if not coop:
        monkeypatch.setenv("DML_NO_COOP", "1")
    d, o = case(name, nwr=10 ** 9, **over)               # salida() never runs: try / depo accumulate over the whole test
    if name == "ermak":
        o.step(150)                                       # deposits exist: contacts with metal happen
    ls = P.Lockstep(o, strict=1, chunk_xyz=d.get("chunk_xyz"))
    t0, d0 = o.scalars().try_, o.scalars().depo
    for i in range(nsteps):
        ls.step(check=True, tag="%s %s step %d" % (name, over, i + 1))
    sc = o.scalars()
    if "prob" in over:
        assert sc.try_ - t0 > (sc.depo - d0) >= 0 and sc.try_ - t0 >= 3, "no failed deposition attempt happened: the test is vacuous"


@pytest.mark.parametrize("name,nsteps", [("ermak", 300), ("brown", 120), ("gcmc", 300)])
@pytest.mark.parametrize("coop", [1, 0])
def test_fused_step_replay_bit_exact(name, nsteps, coop, monkeypatch):
    """dml_step itself (not the call-site API) against the oracle: the fused loop body the bench runs — test_update launches that
    carry overlap_moveback's first / last pass and the tail of the step (msd, promotion, calc_rho, maxz), the deferred cell sort of
    a Brownian step's second rebuild, rows built on demand — must leave every array bit-identical to the oracle's after every step
    when it consumes the oracle's random numbers.  coop=0: the one-launch-per-phase forms."""
    if not coop:
        monkeypatch.setenv("DML_NO_COOP", "1")
    else:
        monkeypatch.setenv("DML_COOP_MAX_N", "100")          # multi-launch overlap_moveback between one-launch test_updates,
        monkeypatch.setenv("DML_COOP_TU_MAX_N", "4194304")   # i.e. the combination boxes above 65 536 slots (bench.py's) take
    d, o = case(name, nwr=10 ** 9)
    ls = P.FusedLockstep(o, strict=1, chunk_xyz=d.get("chunk_xyz"))
    for i in range(nsteps):
        ls.step(check=True, tag="%s fused step %d" % (name, i + 1))
    assert o.scalars().nupd > 3


@pytest.mark.parametrize("name,nsteps", [("ermak", 60), ("brown", 40), ("gcmc", 150)])
def test_lockstep_replay_multi_launch_path(name, nsteps, monkeypatch):
    """Same lock-step comparison with the persistent cooperative kernels switched off: the one-launch-per-phase path that
    boxes above 65 536 slots take (bench.py's 100 k and 1 M workloads) must be bit-exact too."""
    monkeypatch.setenv("DML_NO_COOP", "1")
    d, o = case(name)
    ls = P.Lockstep(o, strict=1, chunk_xyz=d.get("chunk_xyz"))
    for i in range(nsteps):
        ls.step(check=True, tag="%s (multi-launch) step %d" % (name, i + 1))


def test_brown_fixture_on_device():
    """The whole tests/brown case on the device in replay mode must end on the reference's ref.xyz."""
    d, o = case("brown")
    ls = P.Lockstep(o, strict=1, chunk_xyz=d["chunk_xyz"])
    for i in range(d["nst"]):
        ls.step(check=(i % 50 == 49), tag="brown step %d" % (i + 1))
    rzmax, rz, rpos = O.read_xyz_frame(os.path.join(GOLD, "brown", "ref.xyz"))
    a = P.oracle_slot_arrays(o)
    g = ls.ctx.download(len(a["z"]))
    st = o.state()
    s = st["slot_hs"] - 1
    assert np.array_equal(g["pos"][s], rpos) and np.array_equal(g["z"][s], rz)
    assert ls.ctx.scalars().zmax == rzmax


def test_ermak_fixture_on_device():
    """tests/ermak (10000 steps, forces + piston) replayed on the device: bit-identical to ref.xyz."""
    d, o = case("ermak")
    ls = P.Lockstep(o, strict=1)
    zmax_pre = None
    for i in range(d["nst"]):
        if i == d["nst"] - 1:
            pass
        ls.step(check=(i % 1000 == 999), tag="ermak step %d" % (i + 1))
    # ref.xyz holds the frame written by salida() BEFORE the last maxz: compare the oracle's
    # frame with ref.xyz (CPU test) and the device with the oracle's final state here.
    P.compare_state(o, ls.ctx, what="ermak final")
    c = ls.ctx.counters()
    assert c.choques == o.scalars().choques == 49905


def test_gcmc_fixture_on_device():
    """tests/gcmc (5000 steps: insertions, deletions, index reuse, limbo, incremental rows) replayed on the device."""
    d, o = case("gcmc")
    ls = P.Lockstep(o, strict=1)
    for i in range(d["nst"]):
        ls.step(check=(i % 250 == 249), tag="gcmc step %d" % (i + 1))
    rzmax, rz, rpos = O.read_xyz_frame(os.path.join(GOLD, "gcmc", "ref.xyz"))
    a = P.oracle_slot_arrays(o)
    g = ls.ctx.download(len(a["z"]))
    st = o.state()
    s = st["slot_hs"] - 1
    assert np.array_equal(g["pos"][s], rpos) and np.array_equal(g["z"][s], rz)
    c = ls.ctx.counters()
    assert c.gcmc_created > 0 and c.gcmc_destroyed > 0
    assert c.nat_sys == 594


def test_dana_b200_driver_writes_reference_layout(tmp_path):
    """tools/dana_host.cpp (C++ stand-in for dana's main program) runs tests/ermak through the C ABI with Philox noise and
    writes Li.xyz in the reference's list-directed layout (same column structure as ref.xyz, atoms conserved)."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "din_mol_li_b200", "dana_b200")
    for f in ("entrada.ini", "movedor.ini"):
        shutil.copy(os.path.join(GOLD, "ermak", f), tmp_path)
    r = subprocess.run([exe, str(tmp_path), "--steps", "300"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = open(os.path.join(tmp_path, "Li.xyz")).read().split("\n")
    ref = open(os.path.join(GOLD, "ermak", "ref.xyz")).read().split("\n")
    assert lines[0] == ref[0]                                   # "        1204"
    assert len(lines[1]) == len(ref[1]) and lines[1].startswith(" info:")
    last = lines[-1207:-1]                                      # last frame
    assert last[0] == ref[0]
    assert all(len(a) == 93 for a in last[2:]) and all(len(a) == 93 for a in ref[2:1206])
    # same initial configuration as the reference run: frame 0 of Li.xyz equals pos_inic of the oracle (same RNG, same rule)
    d = O.read_case(os.path.join(GOLD, "ermak"))
    o = O.Oracle(**d)
    st = o.state()
    frame0 = lines[2:2 + 1204]
    p0 = np.array([[float(x) for x in ln.split()[1:4]] for ln in frame0])
    assert np.array_equal(p0, st["pos"])
    assert "vecinos actualizados" in r.stdout


@pytest.mark.parametrize("name", ["ermak", "brown", "gcmc"])
def test_reference_test_cases_on_the_gpu_backed_binary(name, tmp_path):
    """The reference's own regression test (tests/test.sh: run dana in the case directory, compare the last frame of Li.xyz textually
    with ref.xyz) on dana_b200 --rng reference: the device draws the reference's ran / gasdev stream in the reference's order
    (DML_RNG_REFERENCE), the host writes the frame in gfortran's list-directed layout.  No oracle involved: input files in, text out."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "din_mol_li_b200", "dana_b200")
    for f in ("entrada.ini", "movedor.ini", "chunk.xyz"):
        if os.path.exists(os.path.join(GOLD, name, f)):
            shutil.copy(os.path.join(GOLD, name, f), tmp_path)
    r = subprocess.run([exe, str(tmp_path), "--rng", "reference"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    ref = open(os.path.join(GOLD, name, "ref.xyz")).read().split("\n")
    out = open(os.path.join(tmp_path, "Li.xyz")).read().split("\n")
    assert out[-len(ref):] == ref, "last frame of Li.xyz differs from the reference's ref.xyz"


def _deposited(z):
    return int((z >= 2).sum())


@pytest.mark.parametrize("name,nsteps,nseeds", [("ermak", 2000, 8), ("gcmc", 1500, 8)])
def test_long_run_observables_statistical(name, nsteps, nseeds):
    """Production mode (Philox noise) against the oracle's own RNG: long-run observables agree within statistical error
    (north_star criterion 4).  Observables: deposited atoms (CG+F), particles in the system, mean height of the ions, the Li
    density profile, the pair-distance histogram and (gcmc) the accepted insertions + deletions.
    A run may end in the reference's own abort `supero z0` (src/dana.F90:907): maxz never updates box(3), so ions pushed above it
    sit in halo cells that no stencil visits (SURVEY.md Q-list, Cells.F90:248), pass through each other and blow up when they
    come back; such a draw is replaced by the next seed on either side (at most two per side)."""
    from oracle import observables as OB
    dep_o, dep_g, n_o, n_g, zm_o, zm_g = [], [], [], [], [], []
    prof_o, prof_g, gr_o, gr_g = [], [], [], []
    acc_o, acc_g = [], []
    NZ, NR, RMAX = 6, 8, 12.0
    k, aborted = -1, 0
    while len(dep_g) < nseeds:
        k += 1
        d, o = case(name, idum=-104012 - 17 * k)
        uid0, nsys0 = int(o.state()["uid"].max()), o.scalars().nat_sys
        ztop = o.scalars().zmax * 1.1
        box = list(o.scalars().box)
        ctx = P.ctx_from_oracle(o, rng_mode=dml.RNG_PHILOX, strict=0, seed=9000 + k)
        try:
            ctx.step(nsteps)
            o.step(nsteps)
        except (dml.DmlError, RuntimeError) as e:
            assert "supero z0" in str(e), e
            aborted += 1
            assert aborted <= 4
            ctx.close()
            continue
        c = ctx.counters()
        g = ctx.download(c.n_slots)
        alive = g["z"] > 0
        dep_g.append(_deposited(g["z"][alive])); n_g.append(int(alive.sum())); zm_g.append(g["pos"][alive & (g["z"] == 1), 2].mean())
        # Li density profile and pair-distance histogram of everything, both computed on the device (dml_density_profile, dml_gr)
        prof_g.append(ctx.density_profile(0.0, ztop, NZ, (1,)).astype(float))
        h, nsel = ctx.gr(RMAX, NR, (1, 2, 3))
        gr_g.append(h / float(nsel))
        st = o.state()
        dep_o.append(_deposited(st["z"])); n_o.append(len(st["z"])); zm_o.append(st["pos"][st["z"] == 1, 2].mean())
        prof_o.append(OB.density_profile(st["pos"][:, 2], st["z"], 0.0, ztop, NZ, (1,)).astype(float))
        h, nsel = OB.gr(st["pos"], st["z"], box, (1, 1, 0), RMAX, NR, (1, 2, 3))
        gr_o.append(h / float(nsel))
        # accepted insertions + deletions (GCMC acceptance ratio x nadj x nsteps): creation ranks only grow through insertions and
        # atoms only leave sys through deletions, so both follow from the oracle's lists
        created_o = int(st["uid"].max()) - uid0
        acc_o.append(created_o + (created_o - (len(st["z"]) - nsys0)))
        acc_g.append(int(c.gcmc_created + c.gcmc_destroyed))
        ctx.close()

    def agree(a, b, what, floor):
        a, b = np.array(a, float), np.array(b, float)
        se = np.sqrt((a.var(ddof=1) + b.var(ddof=1)) / len(a)) + floor
        assert abs(a.mean() - b.mean()) <= 4.0 * se, "%s: device %s vs oracle %s (se %.3g)" % (what, a, b, se)

    agree(dep_g, dep_o, name + " deposited atoms", 1.0)
    agree(n_g, n_o, name + " particles", 1.0)
    agree(zm_g, zm_o, name + " mean ion height", 0.05)
    assert np.mean(dep_g) > 5          # the comparison is not vacuous
    # north_star criterion 4 also names the Li density profile and g(r): bin by bin, same rule
    pg, po, gg, go = np.array(prof_g), np.array(prof_o), np.array(gr_g), np.array(gr_o)
    for b in range(NZ):
        agree(pg[:, b], po[:, b], "%s Li density profile, z bin %d" % (name, b), 2.0)
    for b in range(NR):
        agree(gg[:, b], go[:, b], "%s pair-distance histogram per particle, r bin %d" % (name, b), 0.01)
    assert pg.sum() > 100 and gg.sum() > 0.5
    if name == "gcmc":
        agree(acc_g, acc_o, "gcmc accepted insertions + deletions", 4.0)
        assert np.mean(acc_g) > 50


def test_production_force_kernel_with_gather_skip_matches_reference_order():
    """The production pair-force kernel (sub-warp lanes, gather skipping by build-distance bound, doubled ref-ref weight)
    against the reference-order kernel on the same device state, sampled across several rebuild intervals of a Philox
    run of tests/ermak: forces and energies within 1e-12 relative (north_star tolerance)."""
    d, o = case("ermak")
    ctx = P.ctx_from_oracle(o, rng_mode=dml.RNG_PHILOX, strict=0, seed=31337)
    checked = nonzero = 0
    for it in range(40):
        ctx.step(13)
        ctx.ermak_a()                       # positions moved since the last test_update, like inside a step
        ctx.set_strict_order(0)
        ctx.fuerza()
        n = ctx.counters().n_slots
        a = ctx.download(n)
        ctx.set_strict_order(1)
        ctx.fuerza()
        b = ctx.download(n)
        ctx.set_strict_order(0)
        ref = (b["flags"] & 1) > 0
        for k in ("force", "epot"):
            x, y = a[k][ref], b[k][ref]
            scale = np.maximum(np.abs(y), np.abs(y).max() * 1e-3 + 1e-300)
            assert (np.abs(x - y) <= 1e-12 * scale).all(), "iteration %d: %s differs (max rel %g)" % (it, k, (np.abs(x - y) / scale).max())
        checked += int(ref.sum())
        nonzero += int((np.abs(b["force"][ref]).sum(axis=1) > 0).sum())
        ctx.ermak_b()
        ctx.test_update(); ctx.overlap_moveback(); ctx.test_update(); ctx.promote(); ctx.calc_rho(); ctx.maxz()
    assert checked > 30000 and nonzero > 50


def test_fused_ermak_b_is_bit_identical(monkeypatch):
    """With DML_FUSE_ERMAK_B dml_step applies ermak_b inside the production pair-force kernel instead of launching k_ermak_b
    after it.  Same arithmetic on the same forces: every array must be bit-identical after a Philox run of tests/ermak."""
    d, o = case("ermak")
    out = []
    for fuse in (False, True):
        if fuse:
            monkeypatch.setenv("DML_FUSE_ERMAK_B", "1")
        ctx = P.ctx_from_oracle(o, rng_mode=dml.RNG_PHILOX, strict=0, seed=4711)
        ctx.step(300)
        n = ctx.counters().n_slots
        out.append(ctx.download(n))
        ctx.close()
    for k in ("pos", "vel", "acel", "force", "epot", "z", "flags"):
        assert np.array_equal(out[0][k], out[1][k]), k
    assert np.abs(out[0]["vel"]).sum() > 0


@pytest.mark.parametrize("env", ["DML_NO_FLAT_B=1", "DML_NO_DQ=1", "DML_NO_GRAPH=1", "DML_FORCE_MINB=6", "DML_FORCE_MINB=8", "DML_ROWS_LEGACY=1", "DML_NO_COOP=1", "DML_NO_COOP=1,DML_OV_LANES=4", "DML_NO_COOP=1,DML_OV_UNSTAGED=1",
                                 "DML_NO_COOP=1,DML_OV_LANES=2"])
def test_kernel_launch_variants_bit_identical(env, monkeypatch):
    """Regression test, device against device (the oracle comparisons are the lock-step tests): other register budgets of the
    production pair-force kernel, the one-thread ordered row walk instead of the staged one, and the multi-launch forms of
    test_update / overlap_moveback give the bit-identical Philox trajectory of tests/ermak (deposit grows, so cut-off hits, CG
    partners and long rows occur)."""
    d, o = case("ermak")
    out = []
    for variant in ("", env):
        for kv in [x for x in variant.split(",") if x]:
            k, v = kv.split("=")
            monkeypatch.setenv(k, v)
        ctx = P.ctx_from_oracle(o, rng_mode=dml.RNG_PHILOX, strict=0, seed=4711)
        ctx.step(400)
        n = ctx.counters().n_slots
        out.append(ctx.download(n))
        ctx.close()
    for k in ("pos", "vel", "acel", "force", "epot", "z", "flags"):
        assert np.array_equal(out[0][k], out[1][k]), k
    assert np.abs(out[0]["force"]).sum() > 0


def _slab_oracle():
    """Two atomic layers of bcc metal (CG, a = 3.51 A) under ions placed by the reference's pos_inic rule: every ion within a list
    radius of the electrode has a row of 60-90 entries, far beyond the 24 that fit a slot's own row storage."""
    a = 3.51
    nx = 24
    xi = yi = nx * a
    g = np.stack(np.meshgrid(np.arange(nx), np.arange(nx), np.arange(1), indexing="ij"), -1).reshape(-1, 3).astype(float)
    cg = np.concatenate([g * a + [0.9, 0.9, 0.5], (g + 0.5) * a + [0.9, 0.9, 0.5]])
    ions, _ = O.pos_inic(-104012, xi, yi, 100.0)
    ions = ions + [0.0, 0.0, 5.7]
    pos = np.concatenate([cg, ions])
    z = np.concatenate([np.full(len(cg), 2, np.int32), np.ones(len(ions), np.int32)])
    return O.Oracle(idum=-31, xi=xi, yi=yi, z0=60.0, zmax=110.0, h=1e-2, nb_dcut=10.0, integrador=1, reservoir=1, init_xyz=pos, init_z=z)


def test_long_rows_next_to_dense_metal():
    """Rows longer than a slot's own storage are built by a whole warp (rows_long_warp): same entries in the same order as the
    reference's serial walk, at t = 0 and after every rebuild of a replayed run over the electrode (forces from CG partners,
    depositions on contact, hard-sphere move-backs)."""
    o = _slab_oracle()
    ctx = P.ctx_from_oracle(o)
    ctx.test_update()
    a = P.oracle_slot_arrays(o)
    n = len(a["z"])
    nn, rows, _ = o.rows(width=256)
    gnn, grows = ctx.neighbors(n, width=256)
    assert nn.max() > 40 and (nn > 24).sum() > 15                   # the long-row path is exercised
    assert np.array_equal(gnn, nn[:n])
    assert P.rows_as_lists(gnn, grows, 0) == P.rows_as_lists(nn[:n], rows, 1)
    ctx.close()
    ls = P.Lockstep(o, strict=1)
    nupd0 = o.scalars().nupd
    for i in range(120):
        ls.step(check=True, tag="slab step %d" % (i + 1))
        if i % 20 == 19:
            nn, rows, _ = o.rows(width=256)
            n = o.scalars().hs_amax
            gnn, grows = ls.ctx.neighbors(n, width=256)
            assert P.rows_as_lists(gnn, grows, 0) == P.rows_as_lists(nn[:n], rows, 1), "rows differ after step %d" % (i + 1)
    assert o.scalars().nupd > nupd0 + 2                             # several device-side rebuilds happened


@pytest.mark.parametrize("name,nsteps", [("ermak", 400), ("brown", 150)])
def test_step_with_folded_overlap_passes_bit_identical(name, nsteps, monkeypatch):
    """Inside dml_step the first and last pass of overlap_moveback over the slots (k_ov_init, k_ov_apply) ride on the two
    test_update launches around it (boxes above the cooperative-overlap size, i.e. bench.py's).  Forced here on the small fixtures:
    the Philox trajectory must be bit-identical with and without the folding."""
    d, o = case(name)
    monkeypatch.setenv("DML_COOP_MAX_N", "100")              # multi-launch overlap_moveback ...
    monkeypatch.setenv("DML_COOP_TU_MAX_N", "4194304")       # ... between one-launch test_updates
    out = []
    for nofuse in (False, True):
        if nofuse:
            monkeypatch.setenv("DML_NO_TU_FUSE", "1")
        ctx = P.ctx_from_oracle(o, rng_mode=dml.RNG_PHILOX, strict=0, seed=2024)
        if name == "brown":
            ch = P.ChunkTemplate(d["chunk_xyz"], o.scalars().zmax, o.params.dist + 3.2)
            ctx.set_chunk_template(ch.pos, ch.pos_old, ch.dist, P.RHOMEDIA)
        ctx.step(nsteps)
        c = ctx.counters()
        out.append((ctx.download(c.n_slots), c.choques, c.choques2, c.overlap_passes, c.nupd_vlist))
        ctx.close()
    for k in ("pos", "vel", "acel", "pos_old", "z", "flags"):
        assert np.array_equal(out[0][0][k], out[1][0][k]), k
    assert out[0][1:] == out[1][1:] and out[0][1] > 0


@pytest.mark.parametrize("env", ["DML_NO_BI_FUSE=1", "DML_NO_GRAPH=1", "DML_NO_COOP=1", "DML_COOP_MAX_N=100,DML_NO_BI_FUSE=1"])
def test_brownian_step_variants_bit_identical(env, monkeypatch):
    """Inside dml_step the Brownian integrator (cbrownian_hs + atom_pbc) rides on the first pass of the test_update behind it
    (k_test_update_coop<true>).  The Philox trajectory of tests/brown (chunk reservoir, depositions, move-backs) must be bit-identical
    with the integrator as a launch of its own, without the step graph, and with every phase as its own launch."""
    d, o = case("brown")
    out = []
    for variant in ("", env):
        for kv in [x for x in variant.split(",") if x]:
            k, v = kv.split("=")
            monkeypatch.setenv(k, v)
        ctx = P.ctx_from_oracle(o, rng_mode=dml.RNG_PHILOX, strict=0, seed=77)
        ch = P.ChunkTemplate(d["chunk_xyz"], o.scalars().zmax, o.params.dist + 3.2)
        ctx.set_chunk_template(ch.pos, ch.pos_old, ch.dist, P.RHOMEDIA)
        ctx.step(150)
        c = ctx.counters()
        out.append((ctx.download(c.n_slots), c.choques, c.depo, c.try_, c.nupd_vlist, c.max_vel, c.msd_t))
        ctx.close()
    for k in ("pos", "vel", "pos_old", "old_cg", "z", "flags"):
        assert np.array_equal(out[0][0][k], out[1][0][k]), k
    assert out[0][1:6] == out[1][1:6] and out[0][1] > 0
    assert abs(out[0][6] - out[1][6]) <= 1e-12 * abs(out[0][6])      # msd_t: a sum of per-block partials added atomically (order not fixed)


@pytest.mark.parametrize("integrador,nsteps", [(1, 300), (0, 120)])
def test_box_without_cell_lists_uses_verlet_rows(integrador, nsteps):
    """A box with fewer than 4 cells on every axis is not tessellated (Cells.F90:231): the reference builds its rows with the
    O(N^2) loop ngroup_verlet (Neighbor.F90:358-424: candidates in ascending hs%b index, inclusive <=).  Same rows at t = 0 and a
    bit-identical replayed trajectory over several rebuilds."""
    o = O.Oracle(idum=-77, xi=40.0, yi=40.0, z0=25.0, zmax=50.0, h=1e-2, nb_dcut=10.0, integrador=integrador, reservoir=1)
    sc = o.scalars()
    assert sc.tessellated == 0 and list(sc.ncells) == [3, 3, 3]
    ctx = P.ctx_from_oracle(o)
    ctx.test_update()
    assert ctx.counters().tessellated == 0
    P.compare_rows(o, ctx, what="verlet rows t=0")
    ctx.close()
    ls = P.Lockstep(o, strict=1)
    nupd0 = o.scalars().nupd
    for i in range(nsteps):
        ls.step(check=True, tag="small box step %d" % (i + 1))
        if i % 40 == 39:
            P.compare_rows(o, ls.ctx, what="verlet rows after step %d" % (i + 1))
    assert o.scalars().nupd > nupd0 + 2
    # the production path (Philox, dml_step with its folded passes) runs on such a box too
    ctx = P.ctx_from_oracle(o, rng_mode=dml.RNG_PHILOX, strict=0, seed=3)
    ctx.step(50)
    assert ctx.counters().nat_sys == o.scalars().nat_sys
    ctx.close()


def test_gcmc_step_equals_call_site_sequence():
    """dml_step on the grand-canonical case (whose first test_update skips the forced cell sort that only gcmc_run needs) against
    the same loop body issued call site by call site (src/dana.F90:173-265), Philox noise: bit-identical state and counters."""
    d, o = case("gcmc")
    out = []
    for fused in (True, False):
        ctx = P.ctx_from_oracle(o, rng_mode=dml.RNG_PHILOX, strict=0, seed=77)
        if fused:
            ctx.step(300)
        else:
            for i in range(300):
                ctx.cbrownian_hs(); ctx.test_update(); ctx.overlap_moveback(); ctx.test_update(); ctx.msd_book(); ctx.promote()
                ctx.gcmc_run(); ctx.calc_rho()
        c = ctx.counters()
        out.append((ctx.download(c.n_slots), (c.nupd_vlist, c.choques, c.gcmc_created, c.gcmc_destroyed, c.nat_sys, c.nat_ref, c.list_entries, c.nat_gcmc)))   # msd_t (log only) is an atomic sum: last-ulp noise
        ctx.close()
    assert out[0][1] == out[1][1] and out[0][1][2] > 10 and out[0][1][3] > 10
    for k in ("pos", "vel", "pos_old", "z", "flags", "uid", "slot_b"):
        assert np.array_equal(out[0][0][k], out[1][0][k]), k


def test_slab_decomposition_two_gpus():
    """z-slab decomposition over NCCL (tests/slab_check.py under torchrun, 2 ranks): identical pair sets and forces within
    1e-12 of the single-GPU result, before and after a move + halo refresh.  Needs two GPUs on the box."""
    import subprocess
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29617", os.path.join(root, "tests", "slab_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("-> OK") == 2


def test_slab_step_two_gpus():
    """dml_slab_step (global rebuild decision, migration, ghost re-selection, global piston, cross-face overlap rule) against
    dml_step on one GPU: tests/slab_step_check.py under torchrun, 2 ranks, 200 steps.  Needs two GPUs on the box."""
    import subprocess
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SLAB_STEPS="200")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29619", os.path.join(root, "tests", "slab_step_check.py")], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("-> OK") == 2


# ------------------------------------------------------------------------------------------------------------------------
# Parity at the sizes bench.py runs (BASELINE configs 2-4), production launch configuration: no DML_* overrides, row order
# (strict_order = 0), n > 65 536 so overlap_moveback takes k_ov_detect / k_ov_link / k_ov_resolve, 400-1000 A boxes (wider
# fp32 prefilter band and build-distance quantisation), the z-layer tables and the persisting-L2 window live.
# ------------------------------------------------------------------------------------------------------------------------
def _bench():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    import bench
    return bench


def _oracle_of(w, mnb=10000, **over):
    kw = dict(idum=w["seed"], prob=1.0, h=w["h"], nst=10 ** 9, nwr=10 ** 9, xi=w["xi"], yi=w["yi"], dist=w["dist"], z0=w["z0"], zmax=w["zmax"],
              dif_sc=250.0, dif_sei=250.0, nb_dcut=w["nb_dcut"], integrador=w["integrador"], reservoir=w["reservoir"], chunk_xyz=w["chunk"],
              init_xyz=w["pos"], init_z=w.get("z"), mnb=mnb, act=w.get("act", 0.0), nadj=w.get("nadj", 0), ov_guard_pass=64)
    kw.update(over)
    return O.Oracle(**kw)


@pytest.mark.parametrize("which,nsteps", [("brown", 8), ("gcmc", 6)])
def test_lockstep_at_bench_size_100k(which, nsteps):
    """BASELINE configs 2 and 3 exactly as bench.py builds them (100 k particles): rows at t = 0 and a lock-step replay against the
    oracle, every state array bit-for-bit after every call site, rows compared after every gcmc_run."""
    B = _bench()
    w = B.workload_brown(100000, -104012) if which == "brown" else B.workload_gcmc(100000, -104012)
    O.set_threads(len(os.sched_getaffinity(0)))
    o = _oracle_of(w)
    assert o.scalars().nat_sys > 99000
    ls = P.Lockstep(o, strict=0, chunk_xyz=w["chunk"], capacity=int(o.scalars().nat_sys * 1.3) + 8192, fresh=True)
    ch0 = o.scalars().choques
    for i in range(nsteps):
        ls.step(check=True, tag="%s 100k step %d" % (which, i + 1))
    assert o.scalars().choques - ch0 > 100                  # hard-sphere move-backs happened: the overlap resolver was exercised
    if which == "brown":
        assert o.scalars().nat_sys > w["pos"].shape[0]      # the chunk reservoir added a block
    else:
        c = ls.ctx.counters()
        assert c.gcmc_created + c.gcmc_destroyed > 10
    P.compare_rows(o, ls.ctx, what=which + " 100k rows at the end")


@pytest.mark.parametrize("which,nsteps", [("brown", 6), ("gcmc", 4)])
def test_fused_step_at_bench_size_100k(which, nsteps):
    """The same 100 k workloads through dml_step (what bench.py times), production launch configuration, against the oracle's step."""
    B = _bench()
    w = B.workload_brown(100000, -104012) if which == "brown" else B.workload_gcmc(100000, -104012)
    O.set_threads(len(os.sched_getaffinity(0)))
    o = _oracle_of(w)
    ls = P.FusedLockstep(o, strict=0, chunk_xyz=w["chunk"], capacity=int(o.scalars().nat_sys * 1.3) + 8192)
    for i in range(nsteps):
        ls.step(check=True, tag="%s 100k fused step %d" % (which, i + 1))
    P.compare_rows(o, ls.ctx, what=which + " 100k rows at the end")


def test_fused_step_at_1m():
    """BASELINE config 4's box on one GPU through dml_step: Ermak + piston, 1 M particles, production pair force (1e-12)."""
    B = _bench()
    w = B.workload_ermak(1000000, -104012)
    O.set_threads(len(os.sched_getaffinity(0)))
    o = _oracle_of(w, mnb=256)
    ls = P.FusedLockstep(o, strict=0, capacity=o.scalars().nat_sys + 65536)
    for i in range(3):
        ls.step(check=True, tag="1M fused step %d" % (i + 1))


@pytest.mark.parametrize("slab", [False, True])
def test_rows_and_forces_at_1m(slab):
    """BASELINE config 4's box on one GPU (1 M particles, Ermak + piston; slab=True adds the two-layer CG electrode: rows of 60-90
    entries built by whole warps): neighbour rows bit-exact at t = 0, then replayed steps with the production pair-force kernel
    within 1e-12 of the oracle's forces and energies and every other array bit-for-bit."""
    B = _bench()
    w = B.workload_ermak(1000000, -104012, slab=slab)
    O.set_threads(len(os.sched_getaffinity(0)))
    o = _oracle_of(w, mnb=256)
    n = o.scalars().nat_sys
    assert n > 990000
    ls = P.Lockstep(o, strict=0, capacity=n + 65536, fresh=True)          # compares the rows at t = 0
    ne = ls.ctx.counters().list_entries
    assert ne > 4 * n
    nsteps = 4 if not slab else 2
    for i in range(nsteps):
        ls.step(check=True, tag="1M%s step %d" % (" + CG slab" if slab else "", i + 1))
    st = o.state()
    ref = (st["flags"] & 1) > 0
    assert (np.abs(st["force"][ref]).sum(axis=1) > 0).sum() > 20          # the force comparison is not vacuous
