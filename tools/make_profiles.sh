#!/bin/bash
# Turns the raw captures of tools/gpu_profile.sh (gpurun_out/<tag>_*) into the committed summaries under profiles/.
# usage: tools/make_profiles.sh <tag> <round-prefix>      e.g.  tools/make_profiles.sh r1l6 r1_l6
set -eu
tag=$1; name=$2
out=profiles
cp gpurun_out/${tag}_launches.csv $out/${name}_launches.csv
python tools/ncu_summary.py gpurun_out/${tag}_launches.csv > $out/${name}_launch_summary.txt
for rep in gpurun_out/${tag}_k_*.ncu-rep; do
  k=$(basename $rep .ncu-rep); k=${k#${tag}_}
  {
    echo "# ncu --set full --import-source on --clock-control none, the longest launch of $k in the search above"
    python - "$rep" <<'PY'
import sys
sys.path.insert(0, "tools")
import ncu_summary
ncu_summary.report_summary(sys.argv[1])
PY
    echo
    echo "# hottest source lines (warp-stall samples, instruction share, top stall reasons)"
    python tools/ncu_source_hot.py $rep 30
  } > $out/${name}_${k}.txt
done
ls -la $out
