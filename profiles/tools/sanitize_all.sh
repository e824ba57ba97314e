o=gpurun_out/s2t; mkdir -p $o
python profiles/tools/sanitize_small.py > $o/plain.log 2>&1; tail -1 $o/plain.log
for t in memcheck synccheck racecheck initcheck; do
  timeout 600 compute-sanitizer --tool $t --print-limit 30 python profiles/tools/sanitize_small.py > $o/$t.log 2>&1
  echo "== $t rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize run ok" $o/$t.log | tail -3
done
