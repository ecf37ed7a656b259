#!/bin/bash
# Summarise ncu captures from gpurun_out/ into profiles/ (tracked): key metrics per kernel launch.
# usage: scripts/ncu_summary.sh <tag>
TAG=$1
mkdir -p profiles
for rep in gpurun_out/${TAG}_prof_*.ncu-rep; do
  [ -f "$rep" ] || continue
  name=$(basename $rep .ncu-rep)
  ncu -i $rep --page raw --csv 2>/dev/null | python3 scripts/ncu_pick.py > profiles/${name}.txt
done
[ -f gpurun_out/${TAG}_launches.csv ] && python3 scripts/ncu_launches.py gpurun_out/${TAG}_launches.csv > profiles/${TAG}_launch_list.txt
ls -la profiles/
