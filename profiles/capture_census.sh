#!/bin/bash
# ncu captures (one stage-B step per model) behind profiles/kernel_census.json; run on the GPU box, then
#   python profiles/ncu_census.py <model> gpurun_out/r02_census_<model>.ncu-rep <units> <hash> on the build box.
set -x
for m in combsubfast combsub sins; do python profiles/prof_stage.py $m 2 | tail -1 > gpurun_out/r02_census_$m.hash; done
ncu --set full --clock-control none --import-source on -k regex:combsubfast_kernel --launch-skip 3 --launch-count 1 \
    -f -o gpurun_out/r02_census_combsubfast python profiles/prof_stage.py combsubfast 4 > gpurun_out/r02_census_combsubfast.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'ltv_|combtooth' --launch-skip 30 --launch-count 10 \
    -f -o gpurun_out/r02_census_combsub python profiles/prof_stage.py combsub 4 > gpurun_out/r02_census_combsub.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'ltv_|sins_osc' --launch-skip 21 --launch-count 7 \
    -f -o gpurun_out/r02_census_sins python profiles/prof_stage.py sins 4 > gpurun_out/r02_census_sins.log 2>&1
tail -2 gpurun_out/r02_census_*.log; cat gpurun_out/r02_census_*.hash
