set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_default_r1b.json 2> gpurun_out/bench_default_r1b.err; tail -c 400 gpurun_out/bench_default_r1b.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 2 --warmup 3 --no_cpu_baseline --graph 0 > gpurun_out/ncu_bench_b.log 2>&1; tail -2 gpurun_out/ncu_bench_b.log
