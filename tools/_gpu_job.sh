cd /root/repo
( time timeout 900 python bench.py > gpurun_out/bench_r2_e.json 2> gpurun_out/bench_r2_e.err ) 2> gpurun_out/bench_r2_e.time
echo done
