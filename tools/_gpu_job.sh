cd /root/repo
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 2>&1 | tail -60 > gpurun_out/gpu_tests.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2_a.json 2> gpurun_out/bench_r2_a.err
echo done
