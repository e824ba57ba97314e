mkdir -p gpurun_out/s2n
R="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29511"
nvidia-smi topo -m > gpurun_out/s2n/topo.txt 2>&1; nproc >> gpurun_out/s2n/topo.txt; numactl -H >> gpurun_out/s2n/topo.txt 2>&1
$R --nproc-per-node 8 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu > gpurun_out/s2n/c2_n8.json 2> gpurun_out/s2n/c2_n8.err
ECB_BENCH_NO_NUMA=1 $R --nproc-per-node 8 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu > gpurun_out/s2n/c2_n8_nonuma.json 2> gpurun_out/s2n/c2_n8_nonuma.err
$R --nproc-per-node 4 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu > gpurun_out/s2n/c2_n4.json 2> gpurun_out/s2n/c2_n4.err
$R --nproc-per-node 8 bench.py --gpus 8 --config C4 --steps 3 --warmup 2 --no-cpu > gpurun_out/s2n/c4_n8.json 2> gpurun_out/s2n/c4_n8.err
$R --nproc-per-node 4 bench.py --gpus 4 --config C4 --steps 3 --warmup 2 --no-cpu > gpurun_out/s2n/c4_n4.json 2> gpurun_out/s2n/c4_n4.err
$R --nproc-per-node 8 bench.py --gpus 8 --config C3 --steps 5 --warmup 3 --no-cpu > gpurun_out/s2n/c3_n8.json 2> gpurun_out/s2n/c3_n8.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/s2n/*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); e=d["e2e"]
        print(f.split("/")[-1], d["n_gpus"], "ms %.3f"%d["ms_per_step"], "value %.4g"%d["value"], "e2e %.4g"%e["value"], "e2e_ms", e.get("ms_per_step"), e.get("numa"), e.get("frac_of_host_ceiling"))
    except Exception as ex: print(f, "ERR", ex)
PY
