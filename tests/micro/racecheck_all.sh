for w in spring sst ns; do
  timeout 130 compute-sanitizer --tool racecheck --racecheck-report analysis python tests/micro/sanitize.py $w > gpurun_out/r03e_racecheck_$w.txt 2>&1; echo "racecheck $w rc=$?"; grep -c "Race reported\|WARN\|ERROR" gpurun_out/r03e_racecheck_$w.txt; tail -2 gpurun_out/r03e_racecheck_$w.txt
done
