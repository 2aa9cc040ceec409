#!/bin/bash
# tools/sanitize_all.sh <tag> [steps] — compute-sanitizer memcheck / racecheck / initcheck over tools/sanitize_run.py, and memcheck over the
# GPU-backed dana binary drawing the reference's random stream (tests/brown, 500 steps).  Logs land in gpurun_out/.
tag=${1:-rXX}; n=${2:-6}
out=gpurun_out; mkdir -p $out
for tool in memcheck racecheck initcheck; do
  timeout 1500 compute-sanitizer --tool $tool python tools/sanitize_run.py $n > $out/sanitize_${tool}_$tag.log 2>&1
  echo "$tool rc=$?: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $out/sanitize_${tool}_$tag.log | tail -1) $(grep -c 'sanitize_run: done' $out/sanitize_${tool}_$tag.log) run(s) completed"
done
work=$(mktemp -d); cp tests/golden/brown/entrada.ini tests/golden/brown/movedor.ini tests/golden/brown/chunk.xyz $work/
timeout 1200 compute-sanitizer --tool memcheck din_mol_li_b200/dana_b200 $work --rng reference > $out/sanitize_dana_brown_$tag.log 2>&1
echo "dana_b200 --rng reference (brown) memcheck rc=$?: $(grep 'ERROR SUMMARY' $out/sanitize_dana_brown_$tag.log | tail -1)"
n=$(wc -l < tests/golden/brown/ref.xyz)
if diff <(tail -n $n $work/Li.xyz) tests/golden/brown/ref.xyz > /dev/null; then echo "Ok_brown under memcheck"; else echo "FAIL_brown under memcheck"; fi
rm -rf $work
