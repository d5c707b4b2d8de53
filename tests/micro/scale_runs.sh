mkdir -p gpurun_out/r02_scale
for cfg in ns sst spring; do
  for mode in strong weak; do
    for n in 1 2 4; do
      out=gpurun_out/r02_scale/${cfg}_${mode}_n${n}.json
      if [ $n -eq 1 ]; then
        timeout 300 python bench.py --config $cfg --scaling $mode --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline > $out 2> ${out%.json}.err
      else
        timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --config $cfg --scaling $mode --gpus $n --steps 3 --warmup 3 > $out 2> ${out%.json}.err
      fi
      echo "$cfg $mode n=$n rc=$? $(cut -c1-120 $out | head -1)"
    done
  done
done
