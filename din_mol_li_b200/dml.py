"""ctypes binding of libdml.so (include/dml.h) — the Python face of the C ABI.

This is plumbing only: every method is one C-ABI call.  There is no CPU fallback; constructing a
`Ctx` without a CUDA device raises.
"""
import ctypes as C
import os
import numpy as np

from . import build as _build

F_REF, F_GCMC, F_SKIP, F_LIMBO = 1, 2, 4, 8
RNG_PHILOX, RNG_REPLAY, RNG_REFERENCE = 0, 1, 2
CLS_FORCE, CLS_LIST, CLS_INTEG, CLS_OVERLAP, CLS_ALL, CLS_BIN, CLS_OTHER, CLS_GCMC = range(8)


class Config(C.Structure):
    _fields_ = [
        ("device", C.c_int32), ("capacity", C.c_int32), ("box", C.c_double * 3), ("pbc", C.c_int32 * 3),
        ("rcut", C.c_double), ("nb_dcut", C.c_double), ("eps", C.c_double * 9), ("r0", C.c_double * 9),
        ("mass", C.c_double * 3), ("h", C.c_double), ("gama", C.c_double), ("Tsist", C.c_double), ("kB_ui", C.c_double),
        ("kB_ui_gcmc", C.c_double), ("dif_sc", C.c_double), ("dif_sei", C.c_double), ("z_sei", C.c_double),
        ("prob", C.c_double), ("z0", C.c_double), ("z1", C.c_double), ("zmax", C.c_double), ("tau", C.c_double),
        ("act", C.c_double), ("nadj", C.c_int32), ("integrador", C.c_int32), ("reservoir", C.c_int32),
        ("rng_mode", C.c_int32), ("seed", C.c_uint64), ("strict_order", C.c_int32),
    ]


class Counters(C.Structure):
    _fields_ = [
        ("nupd_vlist", C.c_int64), ("try_", C.c_int64), ("depo", C.c_int64), ("choques", C.c_int64),
        ("choques2", C.c_int64), ("choques3", C.c_int64), ("list_entries", C.c_int64), ("overlap_passes", C.c_int64),
        ("gcmc_created", C.c_int64), ("gcmc_destroyed", C.c_int64), ("row_overflow", C.c_int64),
        ("max_vel", C.c_double), ("msd_t", C.c_double), ("msd_max", C.c_double),
        ("n_slots", C.c_int32), ("nat_sys", C.c_int32), ("nat_ref", C.c_int32), ("nat_gcmc", C.c_int32),
        ("ncells", C.c_int32 * 3), ("cell", C.c_double * 3), ("tessellated", C.c_int32), ("listed", C.c_int32),
        ("rows_asym", C.c_int32),
    ]


class Scalars(C.Structure):
    _fields_ = [("box", C.c_double * 3), ("z0", C.c_double), ("z1", C.c_double), ("zmax", C.c_double),
                ("rho", C.c_double), ("rho0", C.c_double), ("t", C.c_double), ("step", C.c_int64)]


# names every exported symbol of include/dml.h (tests check the library exports each of them)
SYMBOLS = [
    "dml_create", "dml_destroy", "dml_last_error", "dml_version", "dml_upload", "dml_download", "dml_set_scalars",
    "dml_get_scalars", "dml_get_counters", "dml_reset_try_depo", "dml_test_update", "dml_fuerza", "dml_ermak_a",
    "dml_ermak_b", "dml_cbrownian_hs", "dml_overlap_moveback", "dml_msd_book", "dml_promote", "dml_gcmc_run",
    "dml_calc_rho", "dml_maxz", "dml_bloques", "dml_set_chunk_template", "dml_step", "dml_get_cells",
    "dml_get_neighbors", "dml_set_neighbors", "dml_set_replay_integrator", "dml_set_replay_overlap", "dml_set_replay_gcmc", "dml_profile",
    "dml_profile_get", "dml_profile_kernel", "dml_n_slots", "dml_set_strict_order", "dml_comm_unique_id", "dml_comm_init", "dml_slab_plan", "dml_slab_setup",
    "dml_slab_halo_exchange", "dml_slab_step", "dml_slab_info", "dml_launch_count", "dml_stream",
    "dml_salida_sums", "dml_density_profile", "dml_gr", "dml_membership_changes",
    "dml_set_rng_state", "dml_get_rng_state", "dml_ensemble_step", "dml_set_ensemble_member", "dml_upload_positions", "dml_download_frame",
]

class HostRng(C.Structure):
    """dana's RNG state (include/dml_host.h)."""
    _fields_ = [("idum", C.c_int32), ("ix", C.c_int32), ("iy", C.c_int32), ("stored", C.c_int32), ("g", C.c_double), ("calls", C.c_uint64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = _build.LIB
        if not os.path.exists(path):
            _build.build()
        L = C.CDLL(path)
        vp, i32, dbl = C.c_void_p, C.c_int32, C.c_double
        L.dml_create.argtypes = [C.POINTER(vp), C.POINTER(Config)]
        L.dml_destroy.argtypes = [vp]
        L.dml_destroy.restype = None
        L.dml_last_error.argtypes = [vp]
        L.dml_last_error.restype = C.c_char_p
        L.dml_version.restype = C.c_char_p
        L.dml_upload.argtypes = [vp, i32] + [vp] * 9
        L.dml_download.argtypes = [vp, i32] + [vp] * 11
        L.dml_set_scalars.argtypes = [vp, C.POINTER(Scalars)]
        L.dml_get_scalars.argtypes = [vp, C.POINTER(Scalars)]
        L.dml_get_counters.argtypes = [vp, C.POINTER(Counters)]
        for f in ("dml_reset_try_depo", "dml_test_update", "dml_fuerza", "dml_ermak_a", "dml_ermak_b", "dml_cbrownian_hs",
                  "dml_overlap_moveback", "dml_msd_book", "dml_promote", "dml_gcmc_run"):
            getattr(L, f).argtypes = [vp]
        L.dml_calc_rho.argtypes = [vp, C.POINTER(dbl)]
        L.dml_maxz.argtypes = [vp, C.POINTER(dbl)]
        L.dml_bloques.argtypes = [vp, i32, vp, vp, dbl, dbl, C.POINTER(i32)]
        L.dml_set_chunk_template.argtypes = [vp, i32, vp, vp, dbl, dbl]
        L.dml_step.argtypes = [vp, i32]
        L.dml_get_cells.argtypes = [vp, i32, vp, vp]
        L.dml_get_neighbors.argtypes = [vp, i32, i32, vp, vp]
        L.dml_set_neighbors.argtypes = [vp, i32, i32, vp, vp]
        L.dml_set_replay_integrator.argtypes = [vp, i32, vp, vp, vp]
        L.dml_set_replay_gcmc.argtypes = [vp, i32, vp, i32, vp]
        L.dml_set_replay_overlap.argtypes = [vp, i32, vp, i32, vp]
        L.dml_profile.argtypes = [vp, i32]
        L.dml_profile_get.argtypes = [vp, i32, C.POINTER(dbl), C.POINTER(C.c_int64), i32]
        L.dml_profile_kernel.argtypes = [vp, i32, C.POINTER(C.c_char_p), C.POINTER(dbl), C.POINTER(C.c_int64)]
        L.dml_n_slots.argtypes = [vp]
        L.dml_set_strict_order.argtypes = [vp, i32]
        L.dml_comm_unique_id.argtypes = [vp]
        L.dml_comm_init.argtypes = [vp, vp, i32, i32]
        L.dml_slab_plan.argtypes = [i32, vp, i32, dbl, dbl, vp]
        L.dml_slab_setup.argtypes = [vp, dbl, dbl]
        L.dml_slab_halo_exchange.argtypes = [vp]
        L.dml_slab_step.argtypes = [vp, i32]
        L.dml_slab_info.argtypes = [vp] + [C.POINTER(i32)] * 4
        L.dml_salida_sums.argtypes = [vp, C.POINTER(dbl), C.POINTER(dbl), C.POINTER(dbl), C.POINTER(i32)]
        L.dml_density_profile.argtypes = [vp, dbl, dbl, i32, i32, vp]
        L.dml_gr.argtypes = [vp, dbl, i32, i32, vp, C.POINTER(i32)]
        L.dml_membership_changes.argtypes = [vp, i32, vp, vp, vp, vp, C.POINTER(i32)]
        L.dml_ensemble_step.argtypes = [C.POINTER(vp), i32, i32]
        L.dml_set_ensemble_member.argtypes = [vp, i32]
        L.dml_upload_positions.argtypes = [vp, i32, vp, vp]
        L.dml_download_frame.argtypes = [vp, i32, vp, vp]
        L.dml_set_rng_state.argtypes = [vp, C.POINTER(HostRng)]
        L.dml_get_rng_state.argtypes = [vp, C.POINTER(HostRng)]
        L.dml_launch_count.argtypes = [vp]
        L.dml_launch_count.restype = C.c_int64
        L.dml_stream.argtypes = [vp]
        L.dml_stream.restype = vp
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a


def _i32(a):
    if a is None:
        return None
    return np.ascontiguousarray(a, dtype=np.int32)


def comm_unique_id():
    buf = (C.c_char * 128)()
    rc = lib().dml_comm_unique_id(C.cast(buf, C.c_void_p))
    if rc != 0:
        raise DmlError("ncclGetUniqueId failed (%d)" % rc)
    return bytes(buf)


def slab_plan(z, nranks, lo, hi):
    z = np.ascontiguousarray(z, dtype=np.float64)
    cuts = np.empty(nranks + 1)
    lib().dml_slab_plan(z.shape[0], _p(z), nranks, lo, hi, _p(cuts))
    return cuts


def host_pos_inic(idum, xi, yi, alto):
    """pos_inic (src/dana.F90:330-396) on the host side of libdml: returns (xyz[n,3], rng state)."""
    L = lib()
    L.dmlh_rng_init.argtypes = [C.POINTER(HostRng), C.c_int32]
    L.dmlh_pos_inic.argtypes = [C.POINTER(HostRng), C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_int32]
    r = HostRng()
    L.dmlh_rng_init(C.byref(r), idum)
    cap = int(xi * yi * alto * 6.1e-4) + 16
    xyz = np.empty((cap, 3))
    n = L.dmlh_pos_inic(C.byref(r), xi, yi, alto, _p(xyz), cap)
    if n < 0:
        raise DmlError("pos_inic failed (%d)" % n)
    return xyz[:n].copy(), r


# dana's pair tables (src/dana.F90:87-100), stored [(k-1)*3+(m-1)]
def dana_tables():
    eps = np.zeros((3, 3))
    r0 = np.zeros((3, 3))
    r0[1, 0] = 3.5
    r0[0, 1] = 3.5
    eps[0, 0] = 2313.6
    r0[0, 0] = 3.2
    eps[2, 2] = 121.0
    r0[2, 2] = 3.61
    eps[0, 2] = 529.1
    eps[2, 0] = eps[0, 2]
    r0[0, 2] = 1.564
    r0[2, 0] = r0[0, 2]
    return eps.ravel(), r0.ravel()


def kB_ui_dana():
    return 8.617330350e-5 * (96.485 * 100.0)          # src/dana.F90:22-24


def kB_ui_module():
    axps_mxs, uma_kg, qe_si = 1.0e2, 1.6605402e-27, 1.60219e-19   # src/Constants.F90:100-120,165-167
    joule_ev = 1.0 / qe_si
    ui_ev = axps_mxs * axps_mxs * uma_kg * joule_ev
    return 8.617385e-05 * (1.0 / ui_ev)


def make_config(box, h, nb_dcut, z0, zmax, integrador, reservoir, capacity, prob=1.0, dif_sc=250.0, dif_sei=250.0,
                z1=0.0, act=0.0, nadj=0, rng_mode=RNG_PHILOX, seed=12345, strict_order=0, device=0):
    c = Config()
    c.device, c.capacity = device, capacity
    for k in range(3):
        c.box[k] = box[k]
    c.pbc[0], c.pbc[1], c.pbc[2] = 1, 1, 0
    c.rcut, c.nb_dcut = 3.2, nb_dcut
    eps, r0 = dana_tables()
    for i in range(9):
        c.eps[i], c.r0[i] = eps[i], r0[i]
    for i in range(3):
        c.mass[i] = 6.94
    c.h, c.gama, c.Tsist, c.kB_ui, c.kB_ui_gcmc = h, 1.0, 300.0, kB_ui_dana(), kB_ui_module()
    c.dif_sc, c.dif_sei, c.z_sei = dif_sc, dif_sei, 80.0
    c.prob, c.z0, c.z1, c.zmax, c.tau = prob, z0, z1, zmax, 0.1
    c.act, c.nadj = act, nadj
    c.integrador, c.reservoir, c.rng_mode, c.seed, c.strict_order = integrador, reservoir, rng_mode, seed, strict_order
    return c


class DmlError(RuntimeError):
    pass


def ensemble_step(ctxs, nsteps=1):
    """dml_ensemble_step: nsteps of every replica, enqueued by this one host thread on the replicas' own streams."""
    arr = (C.c_void_p * len(ctxs))(*[c.h for c in ctxs])
    rc = lib().dml_ensemble_step(arr, len(ctxs), nsteps)
    if rc != 0:
        msgs = [lib().dml_last_error(c.h).decode() for c in ctxs]
        raise DmlError("dml_ensemble_step failed: " + "; ".join(m for m in msgs if m))


class Ctx:
    """One device context (one GPU, one stream)."""

    def __init__(self, cfg):
        self.h = C.c_void_p()
        self.cfg = cfg
        rc = lib().dml_create(C.byref(self.h), C.byref(cfg))
        if rc != 0:
            msg = lib().dml_last_error(self.h).decode() if self.h else "no CUDA device (libdml has no CPU fallback)"
            if self.h:
                lib().dml_destroy(self.h)
                self.h = None
            raise DmlError("dml_create failed (%d): %s" % (rc, msg))

    def close(self):
        if getattr(self, "h", None):
            lib().dml_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != 0:
            raise DmlError(lib().dml_last_error(self.h).decode())

    # --- state ---
    def upload(self, pos, z, flags, vel=None, acel=None, pos_old=None, old_cg=None, uid=None, slot_b=None):
        pos = _f64(pos)
        n = pos.shape[0]
        keep = [pos, _f64(vel), _f64(acel), _f64(pos_old), _f64(old_cg), _i32(z), _i32(flags), _i32(uid), _i32(slot_b)]
        self._chk(lib().dml_upload(self.h, n, *[_p(a) for a in keep]))
        self.n = n

    def download(self, n=None):
        n = self.counters().n_slots if n is None else n
        d = dict(pos=np.empty((n, 3)), vel=np.empty((n, 3)), acel=np.empty((n, 3)), force=np.empty((n, 3)), epot=np.empty(n),
                 pos_old=np.empty((n, 3)), old_cg=np.empty((n, 3)), z=np.empty(n, np.int32), flags=np.empty(n, np.int32),
                 uid=np.empty(n, np.int32), slot_b=np.empty(n, np.int32))
        self._chk(lib().dml_download(self.h, n, *[_p(d[k]) for k in ("pos", "vel", "acel", "force", "epot", "pos_old", "old_cg",
                                                                    "z", "flags", "uid", "slot_b")]))
        return d

    def set_scalars(self, box, z0, z1, zmax, rho, rho0, t=0.0, step=0):
        s = Scalars()
        for k in range(3):
            s.box[k] = box[k]
        s.z0, s.z1, s.zmax, s.rho, s.rho0, s.t, s.step = z0, z1, zmax, rho, rho0, t, step
        self._chk(lib().dml_set_scalars(self.h, C.byref(s)))

    def scalars(self):
        s = Scalars()
        self._chk(lib().dml_get_scalars(self.h, C.byref(s)))
        return s

    def counters(self):
        c = Counters()
        self._chk(lib().dml_get_counters(self.h, C.byref(c)))
        return c

    def reset_try_depo(self):
        self._chk(lib().dml_reset_try_depo(self.h))

    # --- call sites of dana's loop body (same names as the reference procedures) ---
    def test_update(self):
        self._chk(lib().dml_test_update(self.h))

    def fuerza(self):
        self._chk(lib().dml_fuerza(self.h))

    def ermak_a(self):
        self._chk(lib().dml_ermak_a(self.h))

    def ermak_b(self):
        self._chk(lib().dml_ermak_b(self.h))

    def cbrownian_hs(self):
        self._chk(lib().dml_cbrownian_hs(self.h))

    def overlap_moveback(self):
        self._chk(lib().dml_overlap_moveback(self.h))

    def msd_book(self):
        self._chk(lib().dml_msd_book(self.h))

    def promote(self):
        self._chk(lib().dml_promote(self.h))

    def gcmc_run(self):
        self._chk(lib().dml_gcmc_run(self.h))

    def calc_rho(self):
        r = C.c_double()
        self._chk(lib().dml_calc_rho(self.h, C.byref(r)))
        return r.value

    def maxz(self):
        z = C.c_double()
        self._chk(lib().dml_maxz(self.h, C.byref(z)))
        return z.value

    def bloques(self, chunk_pos, chunk_pos_old, dist, rhomedia):
        cp, co = _f64(chunk_pos), _f64(chunk_pos_old)
        fired = C.c_int32()
        self._chk(lib().dml_bloques(self.h, cp.shape[0], _p(cp), _p(co), dist, rhomedia, C.byref(fired)))
        return bool(fired.value)

    def set_chunk_template(self, chunk_pos, chunk_pos_old, dist, rhomedia):
        cp, co = _f64(chunk_pos), _f64(chunk_pos_old)
        self._chk(lib().dml_set_chunk_template(self.h, cp.shape[0], _p(cp), _p(co), dist, rhomedia))

    def step(self, n=1):
        self._chk(lib().dml_step(self.h, n))

    def upload_positions(self, pos, pos_old=None):
        pos, po = _f64(pos), _f64(pos_old)
        self._keep = (pos, po)                                # the copy is stream-ordered: keep the arrays alive
        self._chk(lib().dml_upload_positions(self.h, pos.shape[0], _p(pos), _p(po)))

    def download_frame(self, n=None):
        n = self.n_slots() if n is None else n
        pos, z = np.empty((n, 3)), np.empty(n, np.int32)
        self._chk(lib().dml_download_frame(self.h, n, _p(pos), _p(z)))
        return pos, z

    def set_rng_state(self, r):
        self._chk(lib().dml_set_rng_state(self.h, C.byref(r)))

    def rng_state(self):
        r = HostRng()
        self._chk(lib().dml_get_rng_state(self.h, C.byref(r)))
        return r

    def set_ensemble_member(self, on=True):
        self._chk(lib().dml_set_ensemble_member(self.h, 1 if on else 0))

    # --- output reductions and observables (salida/kion sums, rho(z), g(r)) ---
    def salida_sums(self):
        """energia (sum of epot over sys, dana.F90:1160), energia over hs%ref only, temp (kion, dana.F90:1342-1376), j."""
        e, er, t, j = C.c_double(), C.c_double(), C.c_double(), C.c_int32()
        self._chk(lib().dml_salida_sums(self.h, C.byref(e), C.byref(er), C.byref(t), C.byref(j)))
        return e.value, er.value, t.value, j.value

    def density_profile(self, zlo, zhi, nbins, types=(1,)):
        """counts[b] of particles of the given elements (1 Li, 2 CG, 3 F) per z bin."""
        out = np.zeros(nbins, np.int64)
        self._chk(lib().dml_density_profile(self.h, float(zlo), float(zhi), int(nbins), sum(1 << t for t in types), _p(out)))
        return out

    def gr(self, rmax, nbins, types=(1,)):
        """Pair-distance histogram (unordered pairs, vdistance minimum image) of the given elements; returns (counts, n_selected)."""
        out = np.zeros(nbins, np.int64)
        ns = C.c_int32()
        self._chk(lib().dml_gr(self.h, float(rmax), int(nbins), sum(1 << t for t in types), _p(out), C.byref(ns)))
        return out, ns.value

    def membership_changes(self, max_changes=65536):
        """(slot, kind, uid_now, z_now) arrays of the slots whose occupant / membership changed since the previous call (or upload);
        kind bits: 1 new occupant, 2 previous occupant gone, 4 element changed, 8 left hs%ref, 16 left gcmc."""
        sl, kd, ud, zn = (np.zeros(max_changes, np.int32) for _ in range(4))
        n = C.c_int32()
        self._chk(lib().dml_membership_changes(self.h, max_changes, _p(sl), _p(kd), _p(ud), _p(zn), C.byref(n)))
        m = min(n.value, max_changes)
        return sl[:m], kd[:m], ud[:m], zn[:m], n.value

    # --- inspection / parity ---
    def cells(self, n=None):
        n = self.counters().n_slots if n is None else n
        cell = np.empty((n, 3), np.int32)
        chain = np.empty(n, np.int32)
        self._chk(lib().dml_get_cells(self.h, n, _p(cell), _p(chain)))
        return cell, chain

    def neighbors(self, n=None, width=64):
        n = self.counters().n_slots if n is None else n
        while True:
            nn = np.zeros(n, np.int32)
            rows = np.full((n, width), -1, np.int32)
            rc = lib().dml_get_neighbors(self.h, n, width, _p(nn), _p(rows))
            if rc == 1:
                width *= 4
                continue
            self._chk(rc)
            return nn, rows

    def set_neighbors(self, nn, rows):
        nn, rows = _i32(nn), _i32(rows)
        self._chk(lib().dml_set_neighbors(self.h, nn.shape[0], rows.shape[1], _p(nn), _p(rows)))

    def set_replay_integrator(self, gauss, unif_pbc=None, unif_ovl=None):
        g, u, o = _f64(gauss), _f64(unif_pbc), _f64(unif_ovl)
        n = [a.shape[0] for a in (g, u, o) if a is not None][0]
        self._chk(lib().dml_set_replay_integrator(self.h, n, _p(g), _p(u), _p(o)))

    def set_replay_overlap(self, qstart, vals):
        """Per-slot queues of the deposition uniforms overlap_moveback draws (k-th draw of slot s = vals[qstart[s] + k])."""
        q, v = _i32(qstart), _f64(vals)
        self._chk(lib().dml_set_replay_overlap(self.h, q.shape[0] - 1, _p(q), v.shape[0], _p(v)))

    def set_replay_gcmc(self, unif, gauss):
        u, g = _f64(unif), _f64(gauss)
        self._chk(lib().dml_set_replay_gcmc(self.h, u.shape[0], _p(u), g.shape[0], _p(g)))

    def profile(self, on=True):
        self._chk(lib().dml_profile(self.h, 1 if on else 0))

    def profile_only(self, kernel_name):
        """CUDA events around the launches of ONE kernel (by its dml_profile_kernel name): the rest of the pipeline runs undisturbed."""
        names = list(self.profile_kernels().keys())
        self._chk(lib().dml_profile(self.h, 2 + names.index(kernel_name)))
        self.profile_get(CLS_ALL, reset=True)

    def profile_get(self, cls, reset=False):
        ms = C.c_double()
        nl = C.c_int64()
        self._chk(lib().dml_profile_get(self.h, cls, C.byref(ms), C.byref(nl), 1 if reset else 0))
        return ms.value, nl.value

    def profile_kernels(self):
        """{kernel name: (total ms, launches)} accumulated while profiling was on."""
        out = {}
        kid = 0
        while True:
            name, ms, nl = C.c_char_p(), C.c_double(), C.c_int64()
            if lib().dml_profile_kernel(self.h, kid, C.byref(name), C.byref(ms), C.byref(nl)) != 0:
                return out
            out[name.value.decode()] = (ms.value, nl.value)
            kid += 1

    def set_strict_order(self, on):
        self._chk(lib().dml_set_strict_order(self.h, 1 if on else 0))

    def n_slots(self):
        return lib().dml_n_slots(self.h)

    # --- slab decomposition (one ctx per rank) ---
    def comm_init(self, id128, rank, nranks):
        buf = (C.c_char * 128).from_buffer_copy(bytes(id128))
        self._keep_id = buf
        self._chk(lib().dml_comm_init(self.h, C.cast(buf, C.c_void_p), rank, nranks))

    def slab_setup(self, zlo, zhi):
        self._chk(lib().dml_slab_setup(self.h, zlo, zhi))

    def slab_halo_exchange(self):
        self._chk(lib().dml_slab_halo_exchange(self.h))

    def slab_step(self, n=1):
        """dana's loop body on the decomposed box (every rank calls it with the same n)."""
        self._chk(lib().dml_slab_step(self.h, n))

    def slab_info(self):
        v = [C.c_int32() for _ in range(4)]
        lib().dml_slab_info(self.h, *[C.byref(x) for x in v])
        return tuple(x.value for x in v)

    def launch_count(self):
        return lib().dml_launch_count(self.h)
