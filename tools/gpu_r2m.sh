#!/bin/bash
# round-2 GPU call M: Philox interleave depth of the MALA / MH kernels (A/B), then the bench line
mkdir -p gpurun_out
{
for v in "" _ru8 _ru2; do
  echo "== variant '$v'"
  export KLB_LIB_PATH=$PWD/klara.jl_b200/lib/libklara_b200$v.so
  python tools/prof_run.py --sampler MH --nchains 65536 --nsteps 200 --burnin 100 --reps 2 | tail -2 | head -1
  python tools/prof_run.py --sampler MALA --step 0.02 --nchains 65536 --nsteps 200 --burnin 100 --reps 2 | tail -2 | head -1
  python tools/prof_run.py --sampler MH --dim 512 --nchains 65536 --nsteps 200 --burnin 100 --reps 2 | tail -2 | head -1
  python tools/prof_run.py --sampler MALA --target rosen --dim 256 --nchains 32768 --nsteps 2000 --burnin 1000 --step 0.01 --accrate 0.574 --reps 2 | tail -2 | head -1
done
unset KLB_LIB_PATH
} > gpurun_out/r2m_timings.txt 2>&1
cat gpurun_out/r2m_timings.txt
timeout 900 python bench.py > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
echo "bench rc=$?"; tail -2 gpurun_out/r2m_bench.err; grep '^{' gpurun_out/r2m_bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e_full_output'].get('value'), d['parity']['bit_exact'], d['ess']['ess_kernel_ms'])
for k,v in d['configs'].items(): print(k, v['value'], v['ms_per_run'], v['roofline']['frac'])
"
