mkdir -p gpurun_out/r02_scale_final
for cfg in ns sst spring; do
  out=gpurun_out/r02_scale_final/${cfg}_strong_n2.json
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --config $cfg --scaling strong --gpus 2 --steps 3 --warmup 3 > $out 2> ${out%.json}.err
  echo "$cfg strong n=2 rc=$? $(grep '^{' $out | cut -c1-140)"
done
