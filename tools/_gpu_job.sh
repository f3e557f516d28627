cd /root/repo
timeout 900 python -m pytest tests/test_gpu_window_trace.py -x -q 2>&1 | tail -15 > gpurun_out/trace_test.log
( time timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2_b.json 2> gpurun_out/bench_r2_b.err ) 2>> gpurun_out/trace_test.log
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_ref.json 2> gpurun_out/bench_r2_ref.err ) 2>> gpurun_out/trace_test.log
free -g >> gpurun_out/trace_test.log
echo done
