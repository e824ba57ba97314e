#!/bin/bash
# Round-end evidence on one B200: the GPU test suite, the default bench line + the reference arm, the other configurations,
# the ncu launch list of the bench command and one `--set full` capture of every kernel of the step.
o=gpurun_out/$1; mkdir -p $o
python -m pytest tests -x -q -m gpu > $o/pytest.log 2>&1; tail -3 $o/pytest.log
python bench.py --steps 10 --warmup 3 > $o/bench_c2.json 2> $o/bench_c2.err; echo "bench rc=$?"
python bench.py --impl reference --steps 1 --warmup 0 > $o/bench_ref_c2.json 2> $o/bench_ref_c2.err; echo "ref rc=$?"
python bench.py --config C3 --steps 5 --warmup 3 > $o/bench_c3.json 2> $o/bench_c3.err; echo "c3 rc=$?"
python bench.py --config C4 --steps 3 --warmup 3 > $o/bench_c4.json 2> $o/bench_c4.err; echo "c4 rc=$?"
python bench.py --config C5 --steps 2 --warmup 3 > $o/bench_c5.json 2> $o/bench_c5.err; echo "c5 rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $o/launches_bench_c2.csv python bench.py --steps 2 --warmup 1 --no-cpu > $o/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_cluster|k_pair|k_uset_order|k_normal_eq|k_bfs_order|k_assoc_count|k_assoc_write|k_window|k_ingest|k_cost" -s 20 -c 10 -o $o/full python bench.py --steps 1 --warmup 3 --no-cpu > $o/ncu_full.log 2>&1; tail -1 $o/ncu_full.log | cut -c1-120
