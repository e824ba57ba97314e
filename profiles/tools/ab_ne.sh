out=gpurun_out/s2f/ab_ne.jsonl; mkdir -p gpurun_out/s2f; : > $out
for v in 2 3 4 2 3; do
ECB_NE_VARIANT=$v python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(json.dumps({'ne_variant':$v,'ms_per_step':d['ms_per_step'],'stages':{k:round(v['ms'],4) for k,v in d['roofline']['kernels'].items()}}))" | tee -a $out
done
python -m pytest tests/test_gpu_cost.py tests/test_gpu_calibrate.py -q -m gpu -x 2>&1 | tail -3
ECB_NE_VARIANT=3 python -m pytest tests/test_gpu_cost.py -q -m gpu -x 2>&1 | tail -3
