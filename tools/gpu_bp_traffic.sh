#!/bin/bash
# one ncu --set full capture of the backward pass at the bench's launch shape (64 problems per launch, one group)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PDDP_GROUPS=1 ncu --set full --clock-control none --import-source on -k regex:bp_kernel -s 2 -c 1 -f -o gpurun_out/prof_bp_g1 python tools/prof_run.py 4 > gpurun_out/prof_bp_g1.log 2>&1
ncu -i gpurun_out/prof_bp_g1.ncu-rep --page raw --csv 2>/dev/null | PYTHONPATH=. python -c "
import csv,sys,json
rows=list(csv.reader(sys.stdin)); h=rows[0]; v=rows[-1]; d=dict(zip(h,v))
rd=float(d['dram__bytes_read.sum'].replace(',','')); wr=float(d['dram__bytes_write.sum'].replace(',',''))
u=rows[1][h.index('dram__bytes_read.sum')]; uw=rows[1][h.index('dram__bytes_write.sum')]
sc={'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9}
out=dict(kernel='bp_kernel<14,7>', problems_per_launch=64, dram_bytes_read=rd*sc[u], dram_bytes_write=wr*sc[uw], dram_bytes_per_launch=rd*sc[u]+wr*sc[uw], time_us=float(d['gpu__time_duration.sum'].replace(',','')), source='ncu --set full --clock-control none, tools/gpu_bp_traffic.sh', kernel_source_sha1=__import__('bench').kernel_source_hash())
json.dump(out, open('gpurun_out/bp_traffic.json','w'), indent=1); print(out)"
