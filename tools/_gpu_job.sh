cd /root/repo
run() { timeout 300 python bench.py --steps 10 --warmup 3 --no-split --no-tiles --no-cpu-baseline --no-e2e "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', d['value']/1e6, d['ms_per_step'], d['config']['objects_end'])" >> gpurun_out/tune.log; }
run; run --size 4096 --steps 4
echo done
