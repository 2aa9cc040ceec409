#!/bin/bash
# tools/ensemble_probe.sh — BASELINE config 5 on one GPU: throughput of R concurrent 200k replicas (one ctx + stream + host thread
# each) against the blocks per SM of the cooperative test_update kernel (a full-machine cooperative grid cannot overlap with the
# kernels of another replica).
out=gpurun_out; mkdir -p $out
for bpsm in 4 2 1; do for R in 4 8 16; do
  DML_COOP_TU_BPSM=$bpsm timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu --no-gcmc --ermak-particles 0 --ensemble-replicas $R 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); e=d['ensemble_one_gpu']
print('bpsm=$bpsm R=$R  ensemble %.3e  one alone %.3e  single-box bench %.3e' % (e['particle_steps_per_s'], e['one_replica_alone_particle_steps_per_s'], d['value']))"
done; done
