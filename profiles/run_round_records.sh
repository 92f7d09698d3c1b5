set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/f_pytest.log
bash profiles/capture_census.sh > gpurun_out/f_census.log 2>&1
for m in combsubfast combsub sins; do python bench.py --model $m > gpurun_out/f_bench_$m.json 2> gpurun_out/f_bench_$m.err; done
python bench.py --mode forward > gpurun_out/f_forward.json 2> gpurun_out/f_forward.err
python bench.py --mode train > gpurun_out/f_train.json 2> gpurun_out/f_train.err
python bench.py --mode latency --latency-iters 300 > gpurun_out/f_latency_combsubfast.json 2> gpurun_out/f_latency.err
for m in combsubfast combsub sins; do ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/f_launches_$m.csv python profiles/prof_stage.py $m 3 > /dev/null 2>&1; done
tail -3 gpurun_out/f_pytest.log
