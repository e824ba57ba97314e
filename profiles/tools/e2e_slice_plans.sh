mkdir -p gpurun_out/s2m; : > gpurun_out/s2m/plans.txt
run() { ECB_BENCH_STAGGER_US=$1 python bench.py --steps 6 --warmup 3 --no-cpu $2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('stagger $1 plan [$2]: e2e %.3f ms  device %.3f ms' % (d['e2e']['ms_per_step'], d['ms_per_step']))" | tee -a gpurun_out/s2m/plans.txt; }
run 0 ""
run 100 ""
run 100 "--slice-plan 0.4,0.8,1,1,1,1,1,0.8,0.6,0.4"
run 100 "--slice-plan 0.5,1,1,1,1,1,1,0.5"
run 100 "--slice-plan 0.3,0.6,1,1,1,1,1,1,0.7,0.4"
run 0 "--slice-plan 0.4,1,1,0.8,0.8,1,1,0.5"
run 100 "--slice-plan 0.5,1,1.2,1.2,1.2,1.2,1,0.7,0.4"
run 60 "--slice-plan 0.25,0.5,1,1,1,1,1,1,0.75,0.5,0.25"
