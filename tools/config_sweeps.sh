#!/bin/bash
# SURVEY 8d config 5 (tracer-count scaling of advection + vmix at the tx0.1v3 column depth, reduced horizontal
# extent so that nt = 34 fits one B200) and the gx1v7 shape: one line each into gpurun_out/config_sweeps.log
mkdir -p gpurun_out
OUT=gpurun_out/config_sweeps.log
: > $OUT
for nt in 2 4 8 16 34; do
  python bench.py --workload tx_sample --nt $nt --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); p=d['phases_ms_per_step']
print(json.dumps({'workload':'tx_sample 1200x800x62','nt':$nt,'ms_per_step':round(d['ms_per_step'],3),'phases_ms':{k:round(p[k],3) for k in ('TRACER_UPDATE','VMIX_TRACER_IMPLICIT','STATE','MOMENTUM_COLUMN','MOMENTUM_FINISH','SOLVER','HALO')}}))" >> $OUT
done
python bench.py --workload gx1v7 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(json.dumps({'workload':d['config']['workload'],'ms_per_step':d['ms_per_step'],'value':d['value'],'e2e_ms':d['e2e']['ms_per_step'],'phases_ms':d['phases_ms_per_step'],'iterations':d['config']['solver_iterations_per_step']}))" >> $OUT
cat $OUT
