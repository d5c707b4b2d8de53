for m in 1 2 4 8 3 7 15; do
  r=$(DYF_PDL_MASK=$m DYF_DEBUG_GRAPH=1 timeout 300 python bench.py --config sst --rows 38 --steps 2 --warmup 3 2>/tmp/err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],2))")
  echo "mask=$m ms=$r $(grep -c 'capture failed' /tmp/err.txt) failures"
done
