set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_default_r1d.json 2> gpurun_out/bench_default_r1d.err; tail -c 400 gpurun_out/bench_default_r1d.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1d.csv python bench.py --steps 2 --warmup 3 --no_cpu_baseline --graph 0 > gpurun_out/ncu_bench_b.log 2>&1; tail -2 gpurun_out/ncu_bench_b.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:lstm2_bwd_wave -c 1 -o gpurun_out/prof_r1_lstm_bwd_wave -f python tools/prof_targets.py > gpurun_out/ncu_bwd.log 2>&1; tail -1 gpurun_out/ncu_bwd.log
