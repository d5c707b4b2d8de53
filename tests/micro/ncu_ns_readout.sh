for m in cols nocols; do
  if [ $m = nocols ]; then export DYF_DISABLE_READOUT_COLS=1; fi
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_tensor.sum --clock-control none -k regex:"conv_umma|readout|pack_s2d" -s 14 -c 14 --csv --log-file gpurun_out/r02t_$m.csv python tests/micro/prof_forward.py 64 > /dev/null 2>&1
done
python - <<'PY'
import csv
for m in ['cols','nocols']:
    rows=[r for r in csv.reader(open(f'gpurun_out/r02t_{m}.csv')) if len(r)>10]
    hdr=rows[0]
    ki,mi,vi=hdr.index('Kernel Name'),hdr.index('Metric Name'),hdr.index('Metric Value')
    idi=hdr.index('ID')
    d={}
    for r in rows[1:]:
        d.setdefault((int(r[idi]),r[ki][:50]),{})[r[mi]]=r[vi]
    print(m)
    for k in sorted(d): print(k, d[k])
PY
