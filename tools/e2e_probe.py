#!/usr/bin/env python
"""tools/e2e_probe.py — where the host-buffer path (dml_upload + dml_step + dml_download, bench.py's e2e leg) spends its time."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B
import torch
from din_mol_li_b200 import dml

w = B.workload_brown(int(sys.argv[1]) if len(sys.argv) > 1 else 100000, -104012)
ctx = B.make_ctx(w, 0, 777)
for _ in range(5):
    ctx.step(1)
n = ctx.n_slots(); cap = ctx.cfg.capacity
st = ctx.download(n)
pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True).numpy()
H = {k: pin((cap,) + st[k].shape[1:], torch.float64) for k in ("pos", "vel", "acel", "pos_old", "old_cg", "force", "epot")}
for k in ("z", "flags", "uid", "slot_b"):
    H[k] = pin((cap,), torch.int32)
for k in H:
    H[k][:n] = st[k]
lib = dml.lib(); p = lambda a: a.ctypes.data
T = {"upload": 0.0, "step": 0.0, "download": 0.0}
K = 20
for _ in range(K):
    t0 = time.perf_counter()
    assert lib.dml_upload(ctx.h, n, p(H["pos"]), p(H["vel"]), p(H["acel"]), p(H["pos_old"]), p(H["old_cg"]), p(H["z"]), p(H["flags"]), p(H["uid"]), p(H["slot_b"])) == 0
    t1 = time.perf_counter()
    ctx.step(1); n = ctx.n_slots()
    t2 = time.perf_counter()
    assert lib.dml_download(ctx.h, n, p(H["pos"]), p(H["vel"]), p(H["acel"]), p(H["force"]), p(H["epot"]), p(H["pos_old"]), p(H["old_cg"]), p(H["z"]), p(H["flags"]), p(H["uid"]), p(H["slot_b"])) == 0
    t3 = time.perf_counter()
    T["upload"] += t1 - t0; T["step"] += t2 - t1; T["download"] += t3 - t2
print({k: round(v / K * 1e3, 4) for k, v in T.items()}, "ms per call; n =", n, "H2D MB", n * (5 * 24 + 16) / 1e6, "D2H MB", n * (6 * 24 + 8 + 16) / 1e6)
