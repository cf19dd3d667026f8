#!/bin/bash
# One `ncu --set full` capture of the step kernel per task (launch 104 of a full-range rollout: stationary regime), summarised
# into gpurun_out/<tag>_<task>.txt by tools/ncu_summary.py.  Each plain run must exit 0 before its ncu run.  usage: ncu_all.sh <tag> "<title suffix>"
set -e
tag=${1:-r02_h}; note=${2:-final round-2 kernel}
declare -A N=( [quadruped_parkour]=4096 [humanoid_dancing]=8192 [humanoid_soccer]=4096 [bipedal_rescue]=2048 [humanoid_construction]=2048 [humanoid_martial_arts]=4096 [robotic_arm_assembly]=2048 )
TASKS=${TASKS:-"quadruped_parkour humanoid_dancing humanoid_soccer bipedal_rescue humanoid_construction humanoid_martial_arts robotic_arm_assembly"}
mkdir -p /tmp/rep gpurun_out
for t in $TASKS; do
  python tools/prof_step.py ${N[$t]} 105 1.0 $t > gpurun_out/plain_$t.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:b2_env_kernel -s 103 -c 1 -o /tmp/rep/${tag}_$t -f python tools/prof_step.py ${N[$t]} 105 1.0 $t > gpurun_out/ncu_${tag}_$t.log 2>&1
  python tools/ncu_summary.py /tmp/rep/${tag}_$t.ncu-rep gpurun_out/${tag}_$t.txt "$tag $t, ${N[$t]} envs, step 104 of a full-range rollout (stationary regime); $note" > /dev/null
  sed -i "s#/tmp/rep/#gpurun_out/#" gpurun_out/${tag}_$t.txt
  grep -m1 "Duration" gpurun_out/${tag}_$t.txt; grep -m1 "Issue Slots Busy" gpurun_out/${tag}_$t.txt
done
