"""ctypes binding of the CPU oracle (oracle/dana_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the cpu_baseline /
--impl reference legs of bench.py.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "build", "liboracle.so")

ERMAK_A, FUERZA, ERMAK_B, CBROWNIAN, TEST_UPDATE, OVERLAP, PROMOTE, GCMC, CALC_RHO, BLOQUES, SALIDA, MAXZ, MSD, STEP_END = range(1, 15)
TR_GAUSS_INTEG, TR_UNIF_PBC, TR_UNIF_OVERLAP, TR_UNIF_GCMC, TR_GAUSS_GCMC = range(5)


class Params(C.Structure):
    _fields_ = [
        ("idum", C.c_int32), ("prob", C.c_double), ("h", C.c_double), ("nst", C.c_int32), ("nwr", C.c_int32),
        ("xi", C.c_double), ("yi", C.c_double), ("dist", C.c_double), ("z0", C.c_double), ("zmax", C.c_double),
        ("dif_sc", C.c_double), ("dif_sei", C.c_double), ("nb_dcut", C.c_double),
        ("integrador", C.c_int32), ("reservoir", C.c_int32), ("act", C.c_double), ("nadj", C.c_int32),
        ("nchunk", C.c_int32), ("chunk_xyz", C.POINTER(C.c_double)), ("mnb", C.c_int32), ("fast_init", C.c_int32),
        ("n_init", C.c_int32), ("init_xyz", C.POINTER(C.c_double)), ("init_z", C.POINTER(C.c_int32)),
        ("ov_guard_pass", C.c_int32),
    ]


class Scalars(C.Structure):
    _fields_ = [
        ("box", C.c_double * 3), ("z0", C.c_double), ("z1", C.c_double), ("zmax", C.c_double), ("rho", C.c_double),
        ("rho0", C.c_double), ("t", C.c_double), ("h", C.c_double), ("cell", C.c_double * 3), ("ncells", C.c_int32 * 3),
        ("tessellated", C.c_int32), ("listed", C.c_int32),
        ("nat_sys", C.c_int32), ("nat_ref", C.c_int32), ("nat_b", C.c_int32), ("nat_hs", C.c_int32), ("nat_gcmc", C.c_int32),
        ("hs_amax", C.c_int32), ("b_amax", C.c_int32),
        ("nupd", C.c_int64), ("choques", C.c_int64), ("choques2", C.c_int64), ("choques3", C.c_int64),
        ("try_", C.c_int64), ("depo", C.c_int64),
        ("max_vel", C.c_double), ("msd_t", C.c_double), ("msd_max", C.c_double),
        ("ran_calls", C.c_uint64), ("step", C.c_int32),
        ("cc0", C.c_double), ("cc1", C.c_double), ("cc2", C.c_double), ("sdr", C.c_double), ("sdv", C.c_double),
        ("crv1", C.c_double), ("crv2", C.c_double), ("skt", C.c_double),
    ]


def build(force=False):
    """Compile oracle/build/liboracle.so with the committed Makefile (g++ only)."""
    src = [os.path.join(_HERE, f) for f in ("dana_oracle.cpp", "dana_oracle.h")]
    if (not force) and os.path.exists(_LIB) and all(os.path.getmtime(_LIB) >= os.path.getmtime(s) for s in src):
        return _LIB
    subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(Params)]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_last_error.restype = C.c_char_p
        L.orc_last_error.argtypes = [C.c_void_p]
        L.orc_step.argtypes = [C.c_void_p, C.c_int]
        L.orc_call.argtypes = [C.c_void_p, C.c_int]
        L.orc_get_scalars.argtypes = [C.c_void_p, C.POINTER(Scalars)]
        L.orc_get_state.argtypes = [C.c_void_p] + [C.c_void_p] * 12
        L.orc_get_rows.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_get_cells.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_get_frame.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        L.orc_trace_enable.argtypes = [C.c_void_p, C.c_int]
        L.orc_trace_clear.argtypes = [C.c_void_p]
        L.orc_trace_size.restype = C.c_int64
        L.orc_trace_size.argtypes = [C.c_void_p]
        L.orc_trace_get.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_threads.restype = C.c_int
        L.orc_set_threads.argtypes = [C.c_int]
        L.orc_rng_kat.argtypes = [C.c_int32, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.orc_pos_inic.argtypes = [C.c_int32, C.c_double, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def read_case(case_dir):
    """Parse entrada.ini / movedor.ini / chunk.xyz of a reference test case into a dict."""
    def vals(path):
        out = []
        for line in open(path):
            line = line.split("!")[0].strip()
            if line:
                out.append(line)
        return out
    e = vals(os.path.join(case_dir, "entrada.ini"))
    m = vals(os.path.join(case_dir, "movedor.ini"))
    d = dict(idum=int(e[0]), prob=float(e[1]), h=float(e[2]), nst=int(e[3]), nwr=int(e[4]), xi=float(e[5]),
             yi=float(e[6]), dist=float(e[7]), z0=float(e[8]), zmax=float(e[9]), dif_sc=float(e[10]),
             dif_sei=float(e[11]), nb_dcut=float(e[12]))
    d["integrador"] = 1 if m[0].lower().startswith(".t") else 0
    d["reservoir"] = {"piston": 1, "chunks": 2, "gcmc": 3}[m[1].split()[0]]
    if d["reservoir"] == 3:
        a = m[2].split()
        d["act"], d["nadj"] = float(a[0]), int(a[1])
    if d["reservoir"] == 2:
        d["chunk_xyz"] = read_chunk(os.path.join(case_dir, "chunk.xyz"))
    return d


def read_chunk(path):
    lines = open(path).read().split("\n")
    n = int(lines[0].split()[0])
    xyz = np.array([[float(x) for x in lines[2 + i].split()[1:4]] for i in range(n)], dtype=np.float64)
    return xyz


def read_xyz_frame(path):
    """Last-frame fixture (ref.xyz): returns (zmax, z[int], pos[n,3]) with exact double parsing."""
    lines = open(path).read().split("\n")
    n = int(lines[0].split()[0])
    zmax = float(lines[1].split()[1])
    z = np.empty(n, dtype=np.int32)
    pos = np.empty((n, 3), dtype=np.float64)
    for i in range(n):
        f = lines[2 + i].split()
        pos[i] = [float(f[1]), float(f[2]), float(f[3])]
        z[i] = int(f[4])
    return zmax, z, pos


class Oracle:
    """One instance of the reference program state (everything dana.F90 does before its loop has run)."""

    def __init__(self, idum=-104012, prob=1.0, h=1e-2, nst=1, nwr=1, xi=100.0, yi=100.0, dist=50.0, z0=100.0,
                 zmax=200.0, dif_sc=250.0, dif_sei=250.0, nb_dcut=10.0, integrador=1, reservoir=1, act=0.0, nadj=0,
                 chunk_xyz=None, mnb=10000, fast_init=0, init_xyz=None, init_z=None, ov_guard_pass=0):
        L = lib()
        p = Params()
        p.idum, p.prob, p.h, p.nst, p.nwr = idum, prob, h, nst, nwr
        p.xi, p.yi, p.dist, p.z0, p.zmax = xi, yi, dist, z0, zmax
        p.dif_sc, p.dif_sei, p.nb_dcut = dif_sc, dif_sei, nb_dcut
        p.integrador, p.reservoir, p.act, p.nadj = integrador, reservoir, act, nadj
        p.mnb, p.fast_init, p.ov_guard_pass = mnb, fast_init, ov_guard_pass
        self._keep = []
        if chunk_xyz is not None:
            ch = np.ascontiguousarray(chunk_xyz, dtype=np.float64)
            self._keep.append(ch)
            p.nchunk = ch.shape[0]
            p.chunk_xyz = ch.ctypes.data_as(C.POINTER(C.c_double))
        if init_xyz is not None:
            ix = np.ascontiguousarray(init_xyz, dtype=np.float64)
            self._keep.append(ix)
            p.n_init = ix.shape[0]
            p.init_xyz = ix.ctypes.data_as(C.POINTER(C.c_double))
            if init_z is not None:
                iz = np.ascontiguousarray(init_z, dtype=np.int32)
                self._keep.append(iz)
                p.init_z = iz.ctypes.data_as(C.POINTER(C.c_int32))
        self.params = p
        self.h = L.orc_create(C.byref(p))
        err = L.orc_last_error(self.h)
        if err:
            raise RuntimeError(err.decode())

    def close(self):
        if self.h:
            lib().orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError(lib().orc_last_error(self.h).decode())

    def step(self, n=1):
        self._chk(lib().orc_step(self.h, n))

    def call(self, op):
        self._chk(lib().orc_call(self.h, op))

    def scalars(self):
        s = Scalars()
        lib().orc_get_scalars(self.h, C.byref(s))
        return s

    def state(self):
        n = self.scalars().nat_sys
        d = dict(uid=np.empty(n, np.int64), z=np.empty(n, np.int32), pos=np.empty((n, 3)), vel=np.empty((n, 3)),
                 acel=np.empty((n, 3)), force=np.empty((n, 3)), epot=np.empty(n), pos_old=np.empty((n, 3)),
                 old_cg=np.empty((n, 3)), flags=np.empty(n, np.int32), slot_hs=np.empty(n, np.int32),
                 slot_b=np.empty(n, np.int32))
        lib().orc_get_state(self.h, *[_p(d[k]) for k in ("uid", "z", "pos", "vel", "acel", "force", "epot", "pos_old",
                                                        "old_cg", "flags", "slot_hs", "slot_b")])
        return d

    def rows(self, width=512):
        """Neighbour rows as {uid: [uid, ...]} in row order plus the raw tables."""
        amax = self.scalars().hs_amax
        while True:
            nn = np.zeros(amax, np.int32)
            rows = np.zeros((amax, width), np.int32)
            slot_uid = np.zeros(amax, np.int64)
            rc = lib().orc_get_rows(self.h, width, _p(nn), _p(rows), _p(slot_uid))
            if rc >= 0:
                break
            width *= 4
        return nn, rows, slot_uid

    def cells(self):
        amax = self.scalars().b_amax
        cell = np.zeros((amax, 3), np.int32)
        chain = np.zeros(amax, np.int32)
        lib().orc_get_cells(self.h, _p(cell), _p(chain))
        return cell, chain

    def frame(self):
        n = C.c_int32()
        lib().orc_get_frame(self.h, C.byref(n), None, None, None, None)
        zmax = C.c_double()
        z = np.empty(n.value, np.int32)
        pos = np.empty((n.value, 3))
        scal = np.empty(6)
        lib().orc_get_frame(self.h, C.byref(n), C.byref(zmax), _p(z), _p(pos), _p(scal))
        return zmax.value, z, pos, scal

    def trace_enable(self, on=True):
        lib().orc_trace_enable(self.h, 1 if on else 0)

    def trace_clear(self):
        lib().orc_trace_clear(self.h)

    def trace(self):
        n = lib().orc_trace_size(self.h)
        kind = np.empty(n, np.int32)
        uid = np.empty(n, np.int64)
        val = np.empty(n, np.float64)
        if n:
            lib().orc_trace_get(self.h, _p(kind), _p(uid), _p(val))
        return kind, uid, val


def from_case(case_dir, **over):
    d = read_case(case_dir)
    d.update(over)
    return Oracle(**d)


def threads():
    return lib().orc_threads()


def set_threads(n):
    lib().orc_set_threads(int(n))


def rng_kat(idum, n_ran, n_gas):
    r = np.empty(n_ran)
    g = np.empty(n_gas)
    lib().orc_rng_kat(idum, n_ran, _p(r), n_gas, _p(g))
    return r, g


def pos_inic(idum, xi, yi, alto, fast=True):
    cap = int(xi * yi * alto * 6.1e-4) + 16
    xyz = np.empty((cap, 3))
    calls = C.c_uint64()
    n = lib().orc_pos_inic(idum, xi, yi, alto, 1 if fast else 0, _p(xyz), cap, C.byref(calls))
    if n < 0:
        raise RuntimeError("pos_inic failed")
    return xyz[:n].copy(), calls.value
