# Developer tool (GPU): ncu launch list (time + DRAM bytes) of one 304-row SST interpolator forward
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"conv_umma_kernel|conv_up_kernel|groupnorm|linattn|attention_mma|pack_x|head1x1|upsample_kernel" -s 78 -c 80 --csv --log-file gpurun_out/r03b_sst_list.csv python tests/micro/prof_forward_sst.py 304 > /dev/null 2>&1
python - <<'PY'
import csv,re,collections
rows=[r for r in csv.reader(open('gpurun_out/r03b_sst_list.csv')) if len(r)>10]
hdr=rows[0]; ki,mi,vi,idi=hdr.index('Kernel Name'),hdr.index('Metric Name'),hdr.index('Metric Value'),hdr.index('ID')
d=collections.OrderedDict()
for r in rows[1:]:
    d.setdefault((int(r[idi]),re.sub(r'\(.*','',r[ki]).replace('dyf::<unnamed>::','').replace('void ','')[:46]),{})[r[mi]]=float(r[vi].replace(',',''))
tot=0
for (i,n),m in d.items():
    t=m['gpu__time_duration.sum']/1000; tot+=t
    print(f"{i:4d} {n:46s} {t:7.1f} us  rd {m['dram__bytes_read.sum']/1e6:7.1f} MB  wr {m['dram__bytes_write.sum']/1e6:7.1f} MB")
print('total', tot)
PY
