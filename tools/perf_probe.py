#!/usr/bin/env python
"""tools/perf_probe.py — per-kernel device timings of the two bench workloads under a list of environment variants.

    python tools/perf_probe.py brown:100000 ermak:1000000 --variants "default;DML_FORCE_MINB=5;DML_COOP_MAX_N=200000"

Every variant creates a fresh ctx (the switches are read by dml_create), runs warm-up steps, then (a) whole steps timed with
CUDA events on the ctx stream with the L2 flushed before each, (b) a second window with one event pair per launch.
Development tool: its numbers are for choosing kernel variants, bench.py produces the reported ones."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402


_WL = {}


def probe(kind, n, variant, steps, out):
    import torch
    from din_mol_li_b200 import dml
    saved = {}
    for kv in [x for x in variant.split(",") if "=" in x]:
        k, v = kv.split("=", 1)
        saved[k] = os.environ.get(k)
        os.environ[k] = v
    try:
        if (kind, n) not in _WL:
            _WL[(kind, n)] = (B.workload_brown(n, -104012) if kind == "brown" else B.workload_gcmc(n, -104012) if kind == "gcmc" else
                             B.workload_ermak(n, -104012, slab=(kind == "ermakslab")))
        w = _WL[(kind, n)]
        ctx = B.make_ctx(w, 0, 4242)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    ext = torch.cuda.ExternalStream(dml.lib().dml_stream(ctx.h), device=0)
    flush = B.L2Flush(ext)
    for _ in range(6):
        ctx.step(1)
    torch.cuda.synchronize()
    evs = []
    for _ in range(steps):
        flush()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(ext)
        ctx.step(1)
        b.record(ext)
        evs.append((a, b))
    torch.cuda.synchronize()
    per = [a.elapsed_time(b) for a, b in evs]
    step_ms = sum(per) / steps
    ctx.profile(True)
    ctx.profile_get(dml.CLS_ALL, reset=True)
    for _ in range(steps):
        flush()
        ctx.step(1)
    torch.cuda.synchronize()
    kern = ctx.profile_kernels()
    ctx.profile(False)
    c = ctx.counters()
    rec = {"workload": kind, "n": int(c.nat_sys), "variant": variant, "ms_per_step": round(step_ms, 4),
           "min_ms": round(min(per), 4), "max_ms": round(max(per), 4), "psteps_per_s": c.nat_sys / (step_ms * 1e-3),
           "nupd": int(c.nupd_vlist), "list_entries": int(c.list_entries),
           "kernels_us_per_step": {k: [round(v[0] / steps * 1e3, 1), round(v[1] / steps, 2)] for k, v in kern.items() if v[1]}}
    if kind.startswith("ermak"):
        # pair force alone on a FIXED state (same rows, same skip bound from call to call): comparable across kernel variants
        ctx.profile(True)
        ctx.profile_get(dml.CLS_ALL, reset=True)
        nf = 20
        for _ in range(nf):
            flush()
            ctx.fuerza()
        torch.cuda.synchronize()
        kf = ctx.profile_kernels()
        ctx.profile(False)
        rec["fuerza_fixed_state_us"] = round(kf["fuerza"][0] / kf["fuerza"][1] * 1e3, 2)
    print(json.dumps(rec), flush=True)
    out.write(json.dumps(rec) + "\n")
    out.flush()
    ctx.close()
    del flush
    torch.cuda.empty_cache()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workloads", nargs="+", help="kind:n, kind in brown|gcmc|ermak|ermakslab")
    ap.add_argument("--variants", default="default")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "probe.jsonl"))
    a = ap.parse_args()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "a") as out:
        for wl in a.workloads:
            kind, n = wl.split(":")
            for v in a.variants.split(";"):
                probe(kind, int(n), v, a.steps, out)


if __name__ == "__main__":
    main()
