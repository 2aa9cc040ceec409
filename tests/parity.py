"""Test helpers: move state between the CPU oracle (checker) and the CUDA library (thing under test)."""
import numpy as np
from oracle import oracle as O
from din_mol_li_b200 import dml

RHOMEDIA = 5.775329e-4          # dana.F90:445


def oracle_slot_arrays(o):
    """Oracle state re-indexed by hs slot (slot = hs index - 1), the layout libdml uses."""
    st = o.state()
    sc = o.scalars()
    amax = sc.hs_amax
    _, _, slot_uid = o.rows(width=8) if amax else (None, None, np.zeros(0, np.int64))
    out = dict(pos=np.zeros((amax, 3)), vel=np.zeros((amax, 3)), acel=np.zeros((amax, 3)), force=np.zeros((amax, 3)),
               epot=np.zeros(amax), pos_old=np.zeros((amax, 3)), old_cg=np.zeros((amax, 3)), z=np.zeros(amax, np.int32),
               flags=np.zeros(amax, np.int32), uid=np.full(amax, -1, np.int32), slot_b=np.zeros(amax, np.int32))
    s = st["slot_hs"] - 1
    for k in ("pos", "vel", "acel", "force", "epot", "pos_old", "old_cg", "z", "flags"):
        out[k][s] = st[k]
    out["uid"][s] = st["uid"]
    out["slot_b"][s] = st["slot_b"] - 1
    out["flags"][slot_uid[:amax] == -2] = dml.F_LIMBO
    out["alive"] = out["z"] > 0
    return out


def ctx_from_oracle(o, rng_mode=dml.RNG_REPLAY, strict=1, capacity=None, seed=1):
    p = o.params
    sc = o.scalars()
    a = oracle_slot_arrays(o)
    n = a["pos"].shape[0]
    cap = capacity or max(2 * n + 1024, 4096)
    cfg = dml.make_config(box=list(sc.box), h=p.h, nb_dcut=p.nb_dcut, z0=sc.z0, zmax=sc.zmax, z1=sc.z1,
                          integrador=p.integrador, reservoir=p.reservoir, capacity=cap, prob=p.prob, dif_sc=p.dif_sc,
                          dif_sei=p.dif_sei, act=p.act, nadj=p.nadj, rng_mode=rng_mode, seed=seed, strict_order=strict)
    ctx = dml.Ctx(cfg)
    ctx.upload(a["pos"], a["z"], a["flags"], vel=a["vel"], acel=a["acel"], pos_old=a["pos_old"], old_cg=a["old_cg"],
               uid=a["uid"], slot_b=a["slot_b"])
    ctx.set_scalars(list(sc.box), sc.z0, sc.z1, sc.zmax, sc.rho, sc.rho0, sc.t, sc.step)
    return ctx


def push_rows(o, ctx):
    """Give the device exactly the oracle's neighbour rows (hs slots -> 0-based)."""
    nn, rows, _ = o.rows(width=64)
    r = rows.astype(np.int32) - 1
    ctx.set_neighbors(nn, r)


def rows_as_lists(nn, rows, base):
    return [list(rows[i, :nn[i]] - base) for i in range(len(nn))]


def compare_state(o, ctx, fields=("pos", "vel", "acel", "pos_old", "old_cg", "z", "flags"), force=False, rtol=0.0, what=""):
    a = oracle_slot_arrays(o)
    n = a["pos"].shape[0]
    d = ctx.download(n)
    alive = a["alive"]
    assert np.array_equal(d["z"] > 0, alive), what + ": set of occupied slots differs"
    for k in fields:
        x, y = d[k][alive], a[k][alive]
        if k == "flags":
            x, y = x & 7, y & 7
        if k == "old_cg":     # never-set old_cg (1e8) of CG atoms is not uploaded state we track
            m = (a["flags"][alive] & 1) > 0
            x, y = x[m], y[m]
        if not np.array_equal(x, y):
            bad = np.flatnonzero((x != y).reshape(len(x), -1).any(axis=1))
            raise AssertionError("%s: field %s differs for %d atoms (first idx %s: gpu=%s oracle=%s)" %
                                 (what, k, len(bad), bad[:3], x[bad[:3]], y[bad[:3]]))
    if force:
        ref = alive & ((a["flags"] & 1) > 0)
        for k in ("force", "epot"):
            x, y = d[k][ref], a[k][ref]
            if rtol == 0.0:
                assert np.array_equal(x, y), "%s: %s differs (max abs %g)" % (what, k, np.abs(x - y).max())
            else:
                scale = np.maximum(np.abs(y), np.abs(y).max() * 1e-3 + 1e-300)
                assert (np.abs(x - y) <= rtol * scale).all(), "%s: %s beyond rtol (max rel %g)" % (what, k, (np.abs(x - y) / scale).max())
    return d, a


def compare_rows(o, ctx, what="", width=64):
    """Neighbour rows of every ref atom: same entries in the same order (vectorised: usable at 1 M particles)."""
    a = oracle_slot_arrays(o)
    n = len(a["z"])
    nn, rows, _ = o.rows(width=width)
    gnn, grows = ctx.neighbors(n, width=rows.shape[1])
    ref = a["alive"] & ((a["flags"] & 1) > 0)
    assert np.array_equal(gnn[ref], nn[:n][ref]), what + ": row lengths differ for %d ref atoms" % int((gnn[ref] != nn[:n][ref]).sum())
    w = min(rows.shape[1], grows.shape[1])
    assert nn[:n][ref].max(initial=0) <= w
    live = (np.arange(w)[None, :] < nn[:n, None]) & ref[:, None]
    bad = np.flatnonzero(((rows[:n, :w] - 1 != grows[:, :w]) & live).any(axis=1))
    if len(bad):
        i = int(bad[0])
        raise AssertionError("%s: rows of %d slots differ; slot %d: oracle %s device %s" % (what, len(bad), i, list(rows[i, :nn[i]] - 1), list(grows[i, :gnn[i]])))
    return int(nn[:n][ref].sum())


def replay_from_trace(o, slot_of_uid, amax, ng):
    """Split the oracle's RNG trace of one call into per-slot injection arrays.  The deposition uniforms of overlap_moveback come
    back as per-slot queues (qstart[amax+1], values): an atom whose deposition fails (prob < 1) draws again at its next visit."""
    kind, uid, val = o.trace()
    gauss = np.zeros((amax, 6))
    upbc = np.zeros(amax)
    g = kind == O.TR_GAUSS_INTEG
    if g.any():
        gu, gv = uid[g].reshape(-1, ng), val[g].reshape(-1, ng)
        assert (gu == gu[:, :1]).all()
        gauss[slot_of_uid[gu[:, 0]], :ng] = gv
    m = kind == O.TR_UNIF_PBC
    upbc[slot_of_uid[uid[m]]] = val[m]
    m = kind == O.TR_UNIF_OVERLAP
    sl = slot_of_uid[uid[m]]
    order = np.argsort(sl, kind="stable")                 # draws of one slot stay in the order they were made
    qstart = np.zeros(amax + 1, np.int32)
    np.cumsum(np.bincount(sl, minlength=amax), out=qstart[1:])
    uovl = (qstart, np.ascontiguousarray(val[m][order]))
    gm = kind == O.TR_UNIF_GCMC
    gg = kind == O.TR_GAUSS_GCMC
    return gauss, upbc, uovl, val[gm], val[gg]


def uid_to_slot(o):
    st = o.state()
    m = np.full(int(st["uid"].max()) + 2, -1, np.int64)
    m[st["uid"]] = st["slot_hs"] - 1
    return m


class ChunkTemplate:
    """Host copy of dana's `chunk` group (dana.F90:552-587, 762-763)."""

    def __init__(self, chunk_xyz, zmax, dist_eff):
        self.pos = np.array(chunk_xyz, dtype=np.float64).copy()
        self.pos_old = self.pos.copy()
        self.pos[:, 2] = self.pos[:, 2] + zmax
        self.pos_old[:, 2] = self.pos[:, 2] + zmax       # sic (dana.F90:573-574)
        self.dist = dist_eff

    def shift(self):
        self.pos[:, 2] = self.pos[:, 2] + self.dist
        self.pos_old[:, 2] = self.pos_old[:, 2] + self.dist


class Lockstep:
    """Runs dana's loop body (dana.F90:173-265) on the oracle and on the device side by side, the device fed with the
    oracle's random numbers (trace-replay mode), comparing every state array bit-for-bit after each call site."""

    def __init__(self, o, strict=1, chunk_xyz=None, capacity=None, fresh=False):
        """fresh=True: the oracle has not moved since its last list build (e.g. just constructed), so the device builds its own rows
        from the same positions (the production build path) instead of being handed the oracle's."""
        self.o = o
        self.rtol = 0.0 if strict else 1e-12                 # forces: bit-exact in reference order, 1e-12 in row order (north_star)
        self.ctx = ctx_from_oracle(o, rng_mode=dml.RNG_REPLAY, strict=strict, capacity=capacity)
        if fresh:
            self.ctx.test_update()
            compare_rows(o, self.ctx, what="rows at start")
        else:
            push_rows(o, self.ctx)
        self.chunk = None
        if o.params.reservoir == 2:
            self.chunk = ChunkTemplate(chunk_xyz, o.scalars().zmax, o.params.dist + 3.2)
        self.nupd0 = o.scalars().nupd - self.ctx.counters().nupd_vlist
        o.trace_enable(True)

    def step(self, check=True, tag=""):
        o, ctx, chunk = self.o, self.ctx, self.chunk
        p = o.params
        amax = o.scalars().hs_amax
        u2s = uid_to_slot(o)
        o.trace_clear()
        if p.integrador:
            o.call(O.ERMAK_A)
            gauss, upbc, _, _, _ = replay_from_trace(o, u2s, amax, 6)
            ctx.set_replay_integrator(gauss, upbc, None)
            ctx.ermak_a()
            if check:
                compare_state(o, ctx, what=tag + " ermak_a")
            o.call(O.FUERZA)
            ctx.fuerza()
            if check:
                compare_state(o, ctx, force=True, rtol=self.rtol, what=tag + " fuerza")
            o.call(O.ERMAK_B)
            ctx.ermak_b()
        else:
            o.call(O.CBROWNIAN)
            gauss, upbc, _, _, _ = replay_from_trace(o, u2s, amax, 3)
            ctx.set_replay_integrator(gauss, upbc, None)
            ctx.cbrownian_hs()
        if check:
            compare_state(o, ctx, what=tag + " integrator")
        o.call(O.TEST_UPDATE)
        ctx.test_update()
        if check:
            compare_state(o, ctx, what=tag + " test_update 1")
        o.trace_clear()
        o.call(O.OVERLAP)
        _, _, uovl, _, _ = replay_from_trace(o, u2s, amax, 3)
        ctx.set_replay_overlap(*uovl)
        ctx.overlap_moveback()
        if check:
            compare_state(o, ctx, what=tag + " overlap")
        o.call(O.TEST_UPDATE)
        ctx.test_update()
        o.call(O.MSD)
        ctx.msd_book()
        o.call(O.PROMOTE)
        ctx.promote()
        if p.reservoir == 3:
            o.trace_clear()
            o.call(O.GCMC)
            _, _, _, gu, gg = replay_from_trace(o, u2s, amax, 3)
            ctx.set_replay_gcmc(gu, gg)
            ctx.gcmc_run()
            if check:
                compare_state(o, ctx, fields=("pos", "vel", "acel", "pos_old", "z", "flags", "uid", "slot_b"), what=tag + " gcmc_run")
                compare_rows(o, ctx, what=tag + " gcmc_run")
        o.call(O.CALC_RHO)
        rho = ctx.calc_rho()
        assert rho == o.scalars().rho, tag + " rho"
        if p.reservoir == 2:
            o.call(O.BLOQUES)
            if ctx.bloques(chunk.pos, chunk.pos_old, chunk.dist, RHOMEDIA):
                chunk.shift()
        if p.reservoir == 1:
            o.call(O.MAXZ)
            zmax = ctx.maxz()
            assert zmax == o.scalars().zmax, tag + " zmax"
        o.call(O.STEP_END)
        if check:
            compare_state(o, ctx, what=tag + " end of step")
        so, cg = o.scalars(), ctx.counters()
        a = (so.nupd - self.nupd0, so.nat_sys, so.nat_ref)
        b = (cg.nupd_vlist, cg.nat_sys, cg.nat_ref)
        assert a == b, tag + " counters oracle %s vs device %s" % (a, b)


class FusedLockstep:
    """dml_step(1) — the fused loop body the bench runs (test_update carrying overlap_moveback's first / last pass and the tail of
    the step, deferred cell sorts, lazily built rows) — against the oracle's own step, the device fed with the random numbers the
    oracle consumed in that step (trace-replay), every state array compared bit-for-bit at the end of each step."""

    def __init__(self, o, strict=1, chunk_xyz=None, capacity=None):
        self.o = o
        self.rtol = 0.0 if strict else 1e-12
        self.ctx = ctx_from_oracle(o, rng_mode=dml.RNG_REPLAY, strict=strict, capacity=capacity)
        self.ctx.test_update()                              # the oracle has not moved since its last list build: same rows
        compare_rows(o, self.ctx, what="rows at start")
        if o.params.reservoir == 2:
            ch = ChunkTemplate(chunk_xyz, o.scalars().zmax, o.params.dist + 3.2)
            self.ctx.set_chunk_template(ch.pos, ch.pos_old, ch.dist, RHOMEDIA)
        self.nupd0 = o.scalars().nupd - self.ctx.counters().nupd_vlist
        o.trace_enable(True)

    def step(self, check=True, tag=""):
        o, ctx = self.o, self.ctx
        p = o.params
        amax = o.scalars().hs_amax
        u2s = uid_to_slot(o)
        o.trace_clear()
        o.step(1)
        gauss, upbc, uovl, gu, gg = replay_from_trace(o, u2s, amax, 6 if p.integrador else 3)
        ctx.set_replay_integrator(gauss, upbc, None)
        ctx.set_replay_overlap(*uovl)
        if p.reservoir == 3:
            ctx.set_replay_gcmc(gu, gg)
        ctx.step(1)
        if check:
            fields = ("pos", "vel", "acel", "pos_old", "z", "flags", "uid", "slot_b") if p.reservoir == 3 else ("pos", "vel", "acel", "pos_old", "old_cg", "z", "flags")
            compare_state(o, ctx, fields=fields, force=bool(p.integrador), rtol=self.rtol, what=tag)
            so, sg, cg = o.scalars(), ctx.scalars(), ctx.counters()
            assert (so.rho, so.zmax, so.z0) == (sg.rho, sg.zmax, sg.z0), tag + " scalars oracle %s vs device %s" % ((so.rho, so.zmax, so.z0), (sg.rho, sg.zmax, sg.z0))
            a = (so.nupd - self.nupd0, so.nat_sys, so.nat_ref, so.choques)
            b = (cg.nupd_vlist, cg.nat_sys, cg.nat_ref, cg.choques)
            assert a == b, tag + " counters oracle %s vs device %s" % (a, b)
