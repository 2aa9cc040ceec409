"""Run under torchrun with >= 2 ranks (one GPU each): dml_slab_step (the whole loop body of dana on a z-slab decomposed
box: global rebuild decision, migration, ghost re-selection, global piston, cross-face overlap rule) against dml_step on
one GPU, same configuration, same Philox key (the noise is keyed by creation rank and step, so it does not depend on the
decomposition).  Checks:
  1. partition: every particle is owned by exactly one rank after the run, none lost, none duplicated;
  2. the number of list rebuilds (global decision) equals the single-GPU run's on every rank;
  3. positions / elements of (nearly) all particles equal the single-GPU trajectory (summation order differs at 1e-16, and
     what a rank cannot see beyond its halo makes the overlap resolution an approximation: SURVEY.md §8e), and the
     observables agree: deposited atoms, zmax, rho;
  4. particles migrated between slabs during the run (the test is not vacuous).
Exit code 0 = all ranks passed."""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from din_mol_li_b200 import dml  # noqa: E402

NSTEPS = int(os.environ.get("SLAB_STEPS", "60"))


def make_cfg(box, cap, dev):
    return dml.make_config(box=box, h=1e-2, nb_dcut=10.0, z0=box[2] / 2, zmax=box[2], integrador=1, reservoir=1, capacity=cap,
                           rng_mode=dml.RNG_PHILOX, seed=99, strict_order=0, device=dev)


def check(rank, world, lr, nsteps=NSTEPS):
    """The comparison itself, on an initialised process group (bench.py's slab section calls it too and puts the verdict in its JSON).
    Returns a dict with ok (all ranks) and the figures printed below."""
    NSTEPS = nsteps
    box = [200.0, 200.0, 400.0]
    pos0, _ = dml.host_pos_inic(-104012, box[0], box[1], box[2])
    n = len(pos0)
    z0 = np.ones(n, np.int32)
    z0[pos0[:, 2] < 6.0] = 2                                   # a layer of metal so that deposition happens
    fl0 = np.where(z0 == 2, 0, dml.F_REF).astype(np.int32)
    gen = dml.Ctx(make_cfg(box, 2 * n + 4096, lr))
    gen.upload(pos0, z0, fl0, old_cg=np.full((n, 3), 1e8))
    gen.test_update(); gen.fuerza(); rho0 = gen.calc_rho(); gen.set_scalars(box, box[2] / 2, 0.0, box[2], rho0, rho0)
    gen.step(25)
    st = gen.download(n)
    sc = gen.scalars()
    step0 = sc.step
    gen.close()

    def fresh(idx, cap):
        c = dml.Ctx(make_cfg(box, cap, lr))
        c.upload(st["pos"][idx], st["z"][idx], st["flags"][idx] & 7, vel=st["vel"][idx], acel=st["acel"][idx], pos_old=st["pos"][idx],
                 old_cg=st["old_cg"][idx], uid=st["uid"][idx])
        c.set_scalars(box, sc.z0, sc.z1, sc.zmax, sc.rho, sc.rho0, t=sc.t, step=step0)
        return c

    alive = np.flatnonzero(st["z"] > 0)
    ntot = len(alive)
    # ---- single-GPU reference ----
    full = fresh(alive, 2 * n + 4096)
    full.test_update(); full.fuerza()
    full.step(NSTEPS)
    f = full.download(full.counters().n_slots)
    fc, fs = full.counters(), full.scalars()
    fa = f["z"] > 0
    ref = {int(u): i for i, u in enumerate(f["uid"]) if fa[i]}
    full.close()

    # ---- slab decomposition ----
    cuts = dml.slab_plan(st["pos"][alive, 2], world, -1.0e9, 1.0e9)
    zlo, zhi = cuts[rank], cuts[rank + 1]
    own = alive[(st["pos"][alive, 2] >= zlo) & (st["pos"][alive, 2] < zhi)]
    own_uid0 = set(int(u) for u in st["uid"][own])
    slab = fresh(own, 2 * n + 4096)
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(dml.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    slab.comm_init(bytes(idt.cpu().numpy().tobytes()), rank, world)
    slab.slab_setup(zlo, zhi)
    slab.test_update(); slab.fuerza()        # like dana before its loop (dana.F90:140-142); every rank rebuilds (not listed yet)
    slab.slab_step(NSTEPS)
    n_owned, n_ghost, _, _ = slab.slab_info()
    s = slab.download(n_owned)
    c, ss = slab.counters(), slab.scalars()
    sa = s["z"] > 0
    mine = {int(u): i for i, u in enumerate(s["uid"]) if sa[i]}

    # 1. partition
    allu = [None] * world
    dist.all_gather_object(allu, sorted(mine.keys()))
    flat = [u for lst in allu for u in lst]
    part_ok = len(flat) == len(set(flat)) == len(ref) and set(flat) == set(ref.keys())
    inside = all(zlo - 15.0 <= s["pos"][i, 2] < zhi + 15.0 for i in mine.values())      # owned particles sit in (or just left) the slab
    # 2. rebuild count
    nupd_ok = c.nupd_vlist == fc.nupd_vlist
    # 3. trajectory + observables
    idx_s = np.array([mine[u] for u in mine], dtype=np.int64)
    idx_f = np.array([ref[u] for u in mine], dtype=np.int64)
    d = np.abs(s["pos"][idx_s] - f["pos"][idx_f]).max(axis=1) if len(idx_s) else np.zeros(0)
    same_z = s["z"][idx_s] == f["z"][idx_f]
    close = (d <= 1e-6) & same_z
    tot = torch.tensor([float(close.sum()), float(len(close)), float(c.depo), float(c.try_), float(c.choques),
                        float(len(set(mine.keys()) - own_uid0)), float((s["z"][idx_s] >= 2).sum())], dtype=torch.float64, device="cuda")
    dist.all_reduce(tot)
    nclose, nall, depo, tr, ch, migrated, ndep = [float(x) for x in tot.cpu()]
    frac = nclose / max(nall, 1.0)
    dep_ref = float((f["z"][fa] >= 2).sum())
    obs_ok = abs(ndep - dep_ref) <= max(2.0, 0.02 * dep_ref) and abs(ss.zmax - fs.zmax) <= 1e-3 * abs(fs.zmax)
    ok = part_ok and inside and nupd_ok and frac >= 0.995 and obs_ok and migrated > 0
    print("rank %d: owned %d ghosts %d  partition %s  nupd %d/%d  close %.4f (max dev %.3g)  deposited %d/%d  depo %d/%d try %d/%d choques %d/%d  "
          "zmax %.10g/%.10g  rho %.6g/%.6g  migrated %d -> %s" % (
              rank, len(mine), n_ghost, part_ok and inside, c.nupd_vlist, fc.nupd_vlist, frac, float(d.max()) if len(d) else 0.0, ndep, dep_ref,
              depo, fc.depo, tr, fc.try_, ch, fc.choques, ss.zmax, fs.zmax, ss.rho, fs.rho, migrated, "OK" if ok else "FAIL"), flush=True)
    t = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(t)
    slab.close()
    dist.barrier()
    return {"ok": t.item() == 0, "ranks": world, "steps": NSTEPS, "particles": int(ntot), "partition_exact": bool(part_ok and inside),
            "rebuilds_slab_vs_single": [int(c.nupd_vlist), int(fc.nupd_vlist)], "fraction_of_particles_within_1e-6_of_single_gpu": frac,
            "deposited_slab_vs_single": [int(ndep), int(dep_ref)], "zmax_slab_vs_single": [ss.zmax, fs.zmax], "migrated": int(migrated)}


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    r = check(rank, world, lr)
    dist.destroy_process_group()
    sys.exit(0 if r["ok"] else 1)


if __name__ == "__main__":
    main()
