#!/bin/bash
# tools/test_cases.sh — the reference's tests/test.sh (/root/reference/tests/test.sh:1-41) on the GPU-backed binary: every case runs
# `dana_b200 --rng reference` in a scratch copy of its input files and the last frame of Li.xyz is compared textually with the
# reference's ref.xyz.  Usage: tools/test_cases.sh [cases...]   (cases under tests/golden/, default: ermak brown gcmc)
root=$(cd "$(dirname "$0")/.." && pwd)
exe=$root/din_mol_li_b200/dana_b200
tests=${*:-ermak brown gcmc}
rc=0
for t in $tests; do
  src=$root/tests/golden/$t
  work=$(mktemp -d)
  cp $src/entrada.ini $src/movedor.ini $work/
  [ -f $src/chunk.xyz ] && cp $src/chunk.xyz $work/
  ( time $exe $work --rng reference > $work/data_$t ) 2>&1 | grep real
  n=$(wc -l < $src/ref.xyz)
  if diff <(tail -n $n $work/Li.xyz) $src/ref.xyz > /dev/null; then echo "Ok_$t"; else echo "FAIL_$t"; rc=1; fi
  rm -rf $work
done
exit $rc
