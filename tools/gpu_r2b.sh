#!/bin/bash
# round-2 GPU call B: control-code variants of the headline kernel, made on the box by patching the shipped library
mkdir -p gpurun_out
L=$PWD/klara.jl_b200/lib/libklara_b200.so
WS=klb_hmc_ws_kernel
run() {  # name, then sass_patch args...
  name=$1; shift
  cp $L /tmp/v.so
  while [ $# -gt 0 ]; do
    IFS=' ' read -r -a args <<< "$1"; shift
    python tools/sass_patch.py /tmp/v.so /tmp/v.so --quiet "${args[@]}" || echo "patch matched nothing: ${args[*]}"
  done
  echo "== $name"
  KLB_LIB_PATH=/tmp/v.so python tools/prof_run.py --nchains 65536 --nsteps 40 --burnin 20 --reps 2 | tail -2
  KLB_LIB_PATH=/tmp/v.so python tools/prof_run.py --nchains 65536 --nsteps 40 --burnin 20 --reps 2 --arith fma | tail -2 | head -1
}
{
run base
run philox_s2 "--select philox --stall 2 --kernel $WS"
run philox_s3 "--select philox --stall 3 --kernel $WS"
run philox_s4 "--select philox --stall 4 --kernel $WS"
run philox_y_s3 "--select philox --yield --stall 3 --kernel $WS"
run int_s2 "--select intalu --stall 2 --kernel $WS"
run int_y "--select intalu --yield --kernel $WS"
run int_y_s2 "--select intalu --yield --stall 2 --kernel $WS"
run fp64_yield "--select fp64 --yield --kernel $WS"
run fp64_hold "--select fp64 --hold --kernel $WS"
run philox_s2_fp64_hold "--select philox --stall 2 --kernel $WS" "--select fp64 --hold --kernel $WS"
run philox_s2_fp64_yield "--select philox --stall 2 --kernel $WS" "--select fp64 --yield --kernel $WS"
echo "== philox-7 (timing only)"
KLB_LIB_PATH=$PWD/klara.jl_b200/lib/libklara_b200_p7.so python tools/prof_run.py --nchains 65536 --nsteps 40 --burnin 20 --reps 2 | tail -2
cp $PWD/klara.jl_b200/lib/libklara_b200_p7.so /tmp/v7.so
python tools/sass_patch.py /tmp/v7.so /tmp/v7.so --quiet --select philox --stall 2 --kernel $WS
echo "== philox-7 + philox_s2"
KLB_LIB_PATH=/tmp/v7.so python tools/prof_run.py --nchains 65536 --nsteps 40 --burnin 20 --reps 2 | tail -2
# the fused single-role kernels (MALA C5 / C2, MH): Philox and fp64 in the same warp
for pa in "" "--select philox --stall 2 --kernel klb_chain_kernel" "--select philox --yield --kernel klb_chain_kernel" "--select intalu --stall 2 --kernel klb_chain_kernel"; do
  cp $L /tmp/v.so
  [ -n "$pa" ] && python tools/sass_patch.py /tmp/v.so /tmp/v.so --quiet $pa
  echo "== chain kernels: '$pa'"
  KLB_LIB_PATH=/tmp/v.so python tools/prof_run.py --sampler MALA --target rosen --dim 256 --nchains 32768 --nsteps 2000 --burnin 1000 --step 0.01 --accrate 0.574 --reps 2 | tail -2
  KLB_LIB_PATH=/tmp/v.so python tools/prof_run.py --sampler MH --nchains 65536 --nsteps 200 --burnin 100 --reps 2 | tail -2
  KLB_LIB_PATH=/tmp/v.so python tools/prof_run.py --sampler HMC --dim 512 --nchains 65536 --nsteps 40 --burnin 20 --reps 2 | tail -2
done
} > gpurun_out/r2b_variants.txt 2>&1
cat gpurun_out/r2b_variants.txt
