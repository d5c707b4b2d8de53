#!/bin/bash
# Developer tool (GPU): one `ncu --set full` capture of the kernel whose demangled name matches $1 (skip $2 matches),
# running the python script + args that follow; report -> gpurun_out/$NAME.ncu-rep
# usage: NAME=r02x bash tests/micro/ncu_one.sh '<regex>' <skip> tests/micro/prof_forward.py 64
re="$1"; skip="$2"; shift 2
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:"$re" -s "$skip" -c 1 \
  -f -o gpurun_out/${NAME:-ncu_one} python "$@" > gpurun_out/${NAME:-ncu_one}.log 2>&1
tail -2 gpurun_out/${NAME:-ncu_one}.log
