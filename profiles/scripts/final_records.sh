python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --impl reference > gpurun_out/f3_bench_ref.json 2> gpurun_out/f3_bench_ref.err
python bench.py > gpurun_out/f3_bench_combsubfast.json 2> gpurun_out/f3_bench_combsubfast.err
python bench.py --model combsub > gpurun_out/f3_bench_combsub.json 2> gpurun_out/f3_bench_combsub.err
python bench.py --model sins > gpurun_out/f3_bench_sins.json 2> gpurun_out/f3_bench_sins.err
python bench.py --mode forward > gpurun_out/f3_forward.json 2> gpurun_out/f3_forward.err
python bench.py --mode latency > gpurun_out/f3_latency.json 2> gpurun_out/f3_latency.err
python bench.py --mode train > gpurun_out/f3_train.json 2> gpurun_out/f3_train.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
ls -la gpurun_out/f3_*
