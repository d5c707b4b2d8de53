# A-B-A of two builds of the library on the bench (same box, same process layout)
for lib in "" dyffusion_b200/libdyffusion_b200_as3.so "" dyffusion_b200/libdyffusion_b200_as3.so; do
  DYF_LIB=$lib timeout 150 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/ab_lib.json
  python -c "import json; d=json.load(open('gpurun_out/ab_lib.json')); print('lib=[$lib]', round(d['ms_per_step'],2), d['clocks']['sm_mhz'], round(d['kernel_ms_per_step']['conv_up'],2))"
done
