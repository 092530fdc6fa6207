#!/bin/bash
# end-of-round evidence on one B200: GPU suite, full bench line (+ reference arm), ncu --set full of the tensor kernels (+ the SIMT
# kernels of the step), ncu launch list of the bench command, CUPTI timeline of the replayed update graph
TAG=${1:-final}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gpu_tests.log 2>&1; echo "pytest rc $?" >> gpurun_out/${TAG}_gpu_tests.log; tail -3 gpurun_out/${TAG}_gpu_tests.log
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc $?"
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_bench_ref.err; echo "ref rc $?"
python tools/profile_step.py 24576 1 > /dev/null 2>&1 || { echo "profile_step failed without ncu"; exit 1; }
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"nm_gemm|dw_gemm" -c 7 -f -o gpurun_out/${TAG}_prof python tools/profile_step.py 24576 1 > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc $?"
timeout 300 ncu --set full --clock-control none -k regex:"encode_kernel|top_scan|actor_head|multi_reduce|adam_fused|decode_kernel" -c 6 -f -o gpurun_out/${TAG}_prof_simt python tools/profile_head.py 24576 1 > gpurun_out/${TAG}_ncu_simt.log 2>&1; echo "ncu simt rc $?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --quick --steps 1 --warmup 3 > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "ncu list rc $?"
timeout 200 python tools/step_timeline.py graph gpurun_out/${TAG}_update_timeline.md > gpurun_out/${TAG}_timeline.log 2>&1; echo "timeline rc $?"
ls -la gpurun_out | grep ${TAG}
