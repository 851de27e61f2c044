#!/bin/bash
# Runs the GPU binary against a checker binary on a list of argv lines; prints SAME/DIFF per line.
# usage: tools/gpu_check.sh <checker-binary> [file-with-argv-lines]
export HOME=/nonexistent
chk=${1:-oracle/_ref/ries}
list=${2:-tools/gpu_check_cases.txt}
gpu=ries_b200/bin/ries
mkdir -p gpurun_out
fail=0
while IFS= read -r line; do
  [ -z "$line" ] && continue
  a=$($chk $line --no-slow-messages 2>&1; echo "exit=$?")
  t0=$(date +%s.%N)
  b=$(timeout 600 $gpu $line --no-slow-messages 2>&1; echo "exit=$?")
  t1=$(date +%s.%N)
  a2=$(echo "$a" | sed -E 's/Time: [0-9.]+/Time: X/' | sed "s#$chk#ries#g")
  b2=$(echo "$b" | sed -E 's/Time: [0-9.]+/Time: X/' | sed "s#$gpu#ries#g")
  dt=$(python3 -c "print(round($t1-$t0,2))")
  if [ "$a2" == "$b2" ]; then echo "SAME  (${dt}s) $line"; else echo "DIFF  (${dt}s) $line"; diff <(echo "$a2") <(echo "$b2") | head -12; fail=1; fi
done < "$list"
exit $fail
