cd /root/repo
nproc > gpurun_out/stats_test.log
timeout 1500 python -m pytest tests/test_gpu_window_stats.py -x -q -s --durations=5 2>&1 | tail -60 >> gpurun_out/stats_test.log
echo done
