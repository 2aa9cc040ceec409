"""oracle/observables.py — numpy restatement of the output reductions and observables (TEST INFRASTRUCTURE ONLY: imported by
tests/ and nothing else; the product computes these on the device, din_mol_li_b200/csrc/dml_observe.cuh).

  energia / kion     src/dana.F90:1155-1163 (sum of epot over sys%alist), 1342-1376 (vdac = sum m*v.v over sym/='CG', temp =
                     vdac/(j*3*kB_ui)); serial sums in list order like the reference.
  density_profile    the reference only has the scalar calc_rho (src/dana.F90:521-549); the profile is its per-bin version.
  gr                 pair-distance histogram with vdistance's minimum image (src/Groups.F90:995-1016: vd - box*idnint(vd/box)
                     on the periodic axes, |.|^2 = (x^2+y^2)+z^2).  The reference computes no g(r): parity unpinned by the
                     reference, pinned by this brute-force O(N^2) definition.
"""
import numpy as np


def idnint(x):
    """Fortran idnint: nearest integer, halves away from zero."""
    r = np.rint(x)
    tie = np.abs(x - np.trunc(x)) == 0.5
    return np.where(tie, np.trunc(x) + np.sign(x), r)


def energia(epot_sys_order):
    e = 0.0
    for v in epot_sys_order:
        e = e + float(v)
    return e


def kion(vel, z, mass, kB_ui):
    vdac, j = 0.0, 0
    for i in range(len(z)):
        if z[i] == 2:
            continue
        j += 1
        vd = (vel[i, 0] * vel[i, 0] + vel[i, 1] * vel[i, 1]) + vel[i, 2] * vel[i, 2]
        vdac = vdac + vd * mass[z[i] - 1]
    return vdac / (j * 3.0 * kB_ui), j


def density_profile(zpos, z, zlo, zhi, nbins, types=(1,)):
    dz = (zhi - zlo) / float(nbins)
    sel = np.isin(z, types)
    q = (zpos[sel] - zlo) / dz
    q = q[(q >= 0.0) & (q < float(nbins))]
    return np.bincount(q.astype(np.int64), minlength=nbins)[:nbins].astype(np.int64)


def gr(pos, z, box, pbc, rmax, nbins, types=(1,), block=512):
    sel = np.isin(z, types)
    p = pos[sel]
    n = len(p)
    dr_bin = rmax / float(nbins)
    out = np.zeros(nbins, np.int64)
    one_box = 1.0 / np.asarray(box, float)
    for a0 in range(0, n, block):
        a1 = min(n, a0 + block)
        vd = p[a0:a1, None, :] - p[None, :, :]
        for k in range(3):
            if pbc[k]:
                vd[:, :, k] = vd[:, :, k] - box[k] * idnint(vd[:, :, k] * one_box[k])
        d2 = (vd[:, :, 0] * vd[:, :, 0] + vd[:, :, 1] * vd[:, :, 1]) + vd[:, :, 2] * vd[:, :, 2]
        ia = np.arange(a0, a1)[:, None]
        ib = np.arange(n)[None, :]
        m = (ib > ia) & (d2 < rmax * rmax)
        b = (np.sqrt(d2[m]) / dr_bin).astype(np.int64)
        b = b[b < nbins]
        out += np.bincount(b, minlength=nbins)[:nbins]
    return out, n
