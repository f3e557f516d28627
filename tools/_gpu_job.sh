cd /root/repo
timeout 900 python -m pytest tests/test_gpu_windows.py tests/test_gpu_window_trace.py tests/test_gpu_multi.py -x -q 2>&1 | tail -8 > gpurun_out/t.log
run() { timeout 300 python bench.py --steps 10 --warmup 3 --no-split --no-tiles --no-cpu-baseline --no-e2e "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', d['value']/1e6, d['ms_per_step'], d['config']['objects_end'])" >> gpurun_out/tune.log; }
run; run --size 4096 --steps 4
echo done
