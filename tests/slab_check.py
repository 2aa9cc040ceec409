"""Run under torchrun with >= 2 ranks (one GPU each): z-slab decomposition with NCCL halo exchange against the
single-GPU result on the same configuration.  Checks (per rank, on its owned particles):
  1. neighbour pair sets identical to the single-GPU rows (as sets of creation ranks),
  2. pair forces / energies within 1e-12 relative,
  3. the same after one integrator move + halo refresh (dml_slab_halo_exchange).
Exit code 0 = all ranks passed."""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from din_mol_li_b200 import dml  # noqa: E402


def make_cfg(box, cap, dev):
    return dml.make_config(box=box, h=1e-2, nb_dcut=10.0, z0=box[2] / 2, zmax=box[2], integrador=1, reservoir=1, capacity=cap,
                           rng_mode=dml.RNG_PHILOX, seed=99, strict_order=0, device=dev)


def rows_as_uid_sets(ctx, n, uid):
    nn, rows = ctx.neighbors(n, width=128)
    return {int(uid[i]): frozenset(int(uid[j]) for j in rows[i, :nn[i]]) for i in range(n) if nn[i] > 0}


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    box = [200.0, 200.0, 400.0]
    # the same state on every rank: pos_inic rule, a few Ermak steps so that some pairs sit inside the cut-off,
    # and the lowest particles turned into metal (CG) so that non-ref neighbours are exercised
    pos0, _ = dml.host_pos_inic(-104012, box[0], box[1], box[2])
    n = len(pos0)
    z0 = np.ones(n, np.int32)
    z0[pos0[:, 2] < 6.0] = 2
    fl0 = np.where(z0 == 2, 0, dml.F_REF).astype(np.int32)
    gen = dml.Ctx(make_cfg(box, 2 * n + 4096, lr))
    gen.upload(pos0, z0, fl0, old_cg=np.full((n, 3), 1e8))
    gen.test_update(); gen.fuerza(); rho0 = gen.calc_rho(); gen.set_scalars(box, box[2] / 2, 0.0, box[2], rho0, rho0)
    gen.step(25)
    st = gen.download(n)
    sc = gen.scalars()
    gen.close()

    def fresh(idx, cap):
        c = dml.Ctx(make_cfg(box, cap, lr))
        c.upload(st["pos"][idx], st["z"][idx], st["flags"][idx] & 3, vel=st["vel"][idx], acel=st["acel"][idx], pos_old=st["pos"][idx],
                 old_cg=st["old_cg"][idx], uid=st["uid"][idx])
        c.set_scalars(box, sc.z0, sc.z1, sc.zmax, sc.rho, sc.rho0)
        return c

    alive = np.flatnonzero(st["z"] > 0)
    # ---- single-GPU reference ----
    full = fresh(alive, 2 * n + 4096)
    full.test_update(); full.fuerza()
    nf = len(alive)
    f1 = full.download(nf)
    ref_rows = rows_as_uid_sets(full, nf, f1["uid"])
    full.ermak_a(); full.fuerza()
    f2 = full.download(nf)
    by_uid = {int(u): i for i, u in enumerate(f1["uid"])}

    # ---- slab decomposition ----
    cuts = dml.slab_plan(st["pos"][alive, 2], world, -1.0, box[2] * 1.5)
    zlo, zhi = cuts[rank], cuts[rank + 1]
    own = alive[(st["pos"][alive, 2] >= zlo) & (st["pos"][alive, 2] < zhi)]
    slab = fresh(own, 2 * n + 4096)
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(dml.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    slab.comm_init(bytes(idt.cpu().numpy().tobytes()), rank, world)
    slab.slab_setup(zlo, zhi)
    n_owned, n_ghost, ns_lo, ns_hi = slab.slab_info()
    assert n_owned == len(own) and (world == 1 or n_ghost > 0)
    slab.test_update(); slab.fuerza()
    ntot = n_owned + n_ghost
    s1 = slab.download(ntot)
    rows = rows_as_uid_sets(slab, ntot, s1["uid"])
    bad_rows = sum(1 for i in range(n_owned) if (s1["flags"][i] & 1) and rows.get(int(s1["uid"][i]), frozenset()) != ref_rows.get(int(s1["uid"][i]), frozenset()))

    def cmp_force(s, f, what):
        idx = np.array([by_uid[int(u)] for u in s["uid"][:n_owned]])
        refm = (s["flags"][:n_owned] & 1) > 0
        worst = 0.0
        for k in ("force", "epot"):
            x, y = s[k][:n_owned][refm], f[k][idx][refm]
            scale = np.maximum(np.abs(y), np.abs(y).max() * 1e-3 + 1e-300)
            worst = max(worst, float((np.abs(x - y) / scale).max()))
        nz = int((np.abs(f["force"][idx][refm]).sum(axis=1) > 0).sum())
        return worst, nz

    w1, nz1 = cmp_force(s1, f1, "t0")
    slab.ermak_a(); slab.slab_halo_exchange(); slab.fuerza()
    s2 = slab.download(ntot)
    assert np.array_equal(s2["pos"][:n_owned], f2["pos"][[by_uid[int(u)] for u in s2["uid"][:n_owned]]]), "owned positions differ after ermak_a"
    g_idx = [by_uid[int(u)] for u in s2["uid"][n_owned:ntot]]
    ghosts_ok = np.array_equal(s2["pos"][n_owned:ntot], f2["pos"][g_idx])
    w2, nz2 = cmp_force(s2, f2, "after move")
    ok = bad_rows == 0 and w1 <= 1e-12 and w2 <= 1e-12 and ghosts_ok and (nz1 + nz2) > 0
    print("rank %d: owned %d ghosts %d (send lo/hi %d/%d) bad_rows %d  force rel err %.2e / %.2e  nonzero %d/%d  ghosts_refreshed %s -> %s" % (
        rank, n_owned, n_ghost, ns_lo, ns_hi, bad_rows, w1, w2, nz1, nz2, ghosts_ok, "OK" if ok else "FAIL"), flush=True)
    t = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(t)
    slab.close(); full.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 0 else 1)


if __name__ == "__main__":
    main()
