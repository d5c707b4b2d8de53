mkdir -p gpurun_out/r02_scale_final
for cfg in ns sst; do
  out=gpurun_out/r02_scale_final/${cfg}_strong_n8.json
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --config $cfg --scaling strong --gpus 8 --steps 3 --warmup 3 > $out 2> ${out%.json}.err
  echo "$cfg strong n=8 rc=$? $(grep '^{' $out | cut -c1-160)"
done
