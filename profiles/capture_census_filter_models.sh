set -x
for m in combsub sins; do python profiles/prof_stage.py $m 2 | tail -1 > gpurun_out/r02_census_$m.hash; done
ncu --set full --clock-control none --import-source on -k regex:'ltv_|combtooth' --launch-skip 30 --launch-count 10 \
    -f -o gpurun_out/r02_census_combsub python profiles/prof_stage.py combsub 4 > gpurun_out/r02_census_combsub.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'ltv_|sins_osc' --launch-skip 21 --launch-count 7 \
    -f -o gpurun_out/r02_census_sins python profiles/prof_stage.py sins 4 > gpurun_out/r02_census_sins.log 2>&1
tail -2 gpurun_out/r02_census_*.log; cat gpurun_out/r02_census_combsub.hash gpurun_out/r02_census_sins.hash
for m in combsub sins; do ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/f_launches_$m.csv python profiles/prof_stage.py $m 3 > /dev/null 2>&1; done
