#!/usr/bin/env python
"""tools/sanitize_run.py — a short tour of every device path for compute-sanitizer (memcheck / racecheck / initcheck):

    compute-sanitizer --tool memcheck python tools/sanitize_run.py

Replayed steps of the three reference fixtures (cooperative and multi-launch forms), the dense-metal long-row builder, the folded
overlap passes of dml_step, Philox steps, the observables and the membership report.  Results are checked by tests/; this script
only has to touch the code."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from oracle import oracle as O
from din_mol_li_b200 import dml
import parity as P
import test_gpu_parity as T

GOLD = os.path.join(ROOT, "tests", "golden")
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
for env in ({}, {"DML_NO_COOP": "1"}, {"DML_COOP_MAX_N": "100", "DML_COOP_TU_MAX_N": "4194304"}):
    os.environ.update(env)
    for name in ("ermak", "brown", "gcmc"):
        d = O.read_case(os.path.join(GOLD, name))
        o = O.Oracle(**d)
        ls = P.Lockstep(o, strict=1, chunk_xyz=d.get("chunk_xyz"))
        for i in range(steps):
            ls.step(check=False)
        ls.ctx.membership_changes()
        ls.ctx.salida_sums(); ls.ctx.density_profile(0.0, 200.0, 64, (1, 2, 3)); ls.ctx.gr(9.0, 32, (1,))
        ls.ctx.close()
        ctx = P.ctx_from_oracle(O.Oracle(**d), rng_mode=dml.RNG_PHILOX, strict=0, seed=5)
        if name == "brown":
            ch = P.ChunkTemplate(d["chunk_xyz"], 100.0 if False else O.Oracle(**d).scalars().zmax, d["dist"] + 3.2)
            ctx.set_chunk_template(ch.pos, ch.pos_old, ch.dist, P.RHOMEDIA)
        ctx.step(steps)
        ctx.close()
    for k in env:
        os.environ.pop(k)
o = T._slab_oracle()
ls = P.Lockstep(o, strict=1)
ls.ctx.test_update()
for i in range(steps):
    ls.step(check=False)
ls.ctx.close()
# round 2 paths: several members stepped from one host thread (dml_ensemble_step, one-block-per-SM cooperative grids), the frame
# path of the host-buffer interface, the CUDA-graph replay of dml_step (any Philox ctx.step above), the block-staged row build
# (bulk copies + mbarrier) and the staged overlap replay (every DML_NO_COOP=1 run above)
d = O.read_case(os.path.join(GOLD, "ermak"))
members = []
for k in range(3):
    c = P.ctx_from_oracle(O.Oracle(**d), rng_mode=dml.RNG_PHILOX, strict=0, seed=11 + k)
    c.set_ensemble_member(True)
    members.append(c)
dml.ensemble_step(members, steps)
pos, z = members[0].download_frame()
members[0].upload_positions(pos)
members[0].step(2)
for c in members:
    c.close()
print("sanitize_run: done")
