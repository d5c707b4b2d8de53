mkdir -p gpurun_out/r02_scale_final
for cfg in spring ns sst; do
  out=gpurun_out/r02_scale_final/${cfg}_strong_n4.json
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --config $cfg --scaling strong --gpus 4 --steps 3 --warmup 3 > $out 2> ${out%.json}.err
  echo "$cfg strong n=4 rc=$? $(grep '^{' $out | cut -c1-160)"
done
