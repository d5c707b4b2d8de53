for w in ns sst spring; do
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python tests/micro/sanitize.py $w > gpurun_out/r02y_memcheck_$w.txt 2>&1; echo "memcheck $w rc=$?"; tail -2 gpurun_out/r02y_memcheck_$w.txt
done
