# Developer tool (GPU): which kernel family makes the sampler's graph capture fail (DYF_DEBUG_GRAPH prints the failure)
for e in NONE DYF_DISABLE_GN_FUSE DYF_DISABLE_ATTN_FUSE DYF_DISABLE_ATTN_MMA DYF_DISABLE_STEM_UMMA DYF_DISABLE_HEAD1X1 DYF_DISABLE_UPFUSE DYF_DISABLE_CATFUSE; do
  r=$(env $e=1 DYF_DEBUG_GRAPH=1 timeout 300 python bench.py --config sst --rows 38 --steps 2 --warmup 3 2>/tmp/err.txt | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],2))")
  echo "$e ms=$r failures=$(grep -c 'capture failed' /tmp/err.txt)"
done
