#!/bin/bash
# round-2 GPU call 9: flake hunt on the network-level guidance-gradient test, then suite, timeline, bench
mkdir -p gpurun_out
: > gpurun_out/c9_flake.txt
for v in default gemm2off gemm3off; do
  for i in 1 2 3 4 5 6; do
    case $v in
      default) env_opt="";;
      gemm2off) env_opt="B200_OPT_GEMM_V2=0";;
      gemm3off) env_opt="B200_OPT_GEMM_V3=0";;
    esac
    env $env_opt python -m pytest tests/test_unet_gpu.py -q -s -k "guidance_gradient_matches" > gpurun_out/c9_flake_${v}_$i.log 2>&1
    echo "$v run $i rc=$? $(grep -c 'grad rel-L2' gpurun_out/c9_flake_${v}_$i.log) $(grep 'grad rel-L2' gpurun_out/c9_flake_${v}_$i.log | awk '{print $NF}' | tr '\n' ' ')" >> gpurun_out/c9_flake.txt
  done
done
cat gpurun_out/c9_flake.txt
( time python -m pytest tests -m gpu -q -s ) > gpurun_out/c9_tests.log 2>&1
echo "tests rc=$?" | tee gpurun_out/c9_summary.txt
python profiles/timeline_xattn.py > gpurun_out/c9_timeline.txt 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/c9_bench.json 2> gpurun_out/c9_bench.err
echo "bench rc=$?" | tee -a gpurun_out/c9_summary.txt
tail -5 gpurun_out/c9_tests.log; cat gpurun_out/c9_summary.txt; grep with_loss gpurun_out/c9_timeline.txt; tail -c 300 gpurun_out/c9_bench.json
