#!/bin/bash
# tools/gpu_round.sh — what one gpurun call of a round collects: GPU parity tests, the bench line, the ncu launch list of the
# bench command and ncu --set full captures of the dominant kernels.  Everything lands in gpurun_out/ (scratch); the
# summaries committed under profiles/ are made from it with profiles/summarize.py.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh r01c [tests] [bench] [launches] [ncu]'
tag=${1:-rXX}; shift
what=${*:-tests bench launches ncu}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $out/gpu_$tag.txt 2>&1
for w in $what; do
  case $w in
    tests)
      t0=$(date +%s)
      timeout 1200 python -m pytest tests -m gpu -x -q > $out/pytest_gpu_$tag.log 2>&1
      echo "pytest rc=$? in $(( $(date +%s) - t0 )) s"; tail -3 $out/pytest_gpu_$tag.log ;;
    smoke)
      timeout 300 python __graft_entry__.py smoke > $out/smoke_$tag.log 2>&1; echo "smoke rc=$?"; tail -2 $out/smoke_$tag.log ;;
    bench)
      timeout 900 python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err
      echo "bench rc=$?"; cut -c1-600 $out/bench_$tag.json
      timeout 300 python bench.py --impl reference --steps 10 --warmup 1 > $out/bench_ref_$tag.json 2>> $out/bench_$tag.err
      echo "bench ref rc=$?"; cut -c1-300 $out/bench_ref_$tag.json ;;
    launches)
      timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $out/launches_$tag.csv \
        python bench.py --steps 3 --warmup 3 --no-cpu --no-gcmc --ermak-particles 0 --ensemble-replicas 0 > $out/launches_$tag.log 2>&1
      echo "launches rc=$?"
      timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $out/launches_ermak1m_$tag.csv \
        python tools/perf_probe.py ermak:1000000 --steps 4 --out $out/probe_ncu_scratch.jsonl > $out/launches_ermak1m_$tag.log 2>&1
      echo "launches ermak rc=$?" ;;
    ncu)
      timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_fuerza" -s 14 -c 4 -f -o $out/ncu_fuerza1m_$tag \
        python tools/perf_probe.py ermak:1000000 --steps 4 --out $out/probe_ncu_scratch.jsonl > $out/ncu_fuerza1m_$tag.log 2>&1
      echo "ncu fuerza rc=$?"
      timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_rows|k_ov_resolve|k_ov_detect|k_integrate|k_cell_order|k_test_update" -s 40 -c 12 -f \
        -o $out/ncu_brown100k_$tag python bench.py --steps 3 --warmup 3 --no-cpu --no-gcmc --ermak-particles 0 --ensemble-replicas 0 > $out/ncu_brown100k_$tag.log 2>&1
      echo "ncu brown rc=$?" ;;
    probe)
      timeout 900 python tools/perf_probe.py ${PROBE_WL:-brown:100000 ermak:1000000} --steps 30 --variants "${PROBE_VARIANTS:-default}" --out $out/probe_$tag.jsonl > $out/probe_$tag.log 2>&1
      echo "probe rc=$?"; tail -4 $out/probe_$tag.log | cut -c1-900 ;;
  esac
done
ls -la $out | tail -30
