#!/bin/bash
# round-2 baseline on one B200: GPU suite, bench line (+ reference arm), ncu launch list, ncu --set full of the tensor kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_gpu_tests.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2a_gpu_tests.log; tail -3 gpurun_out/r2a_gpu_tests.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc $?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err; echo "ref rc $?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --quick --steps 1 --warmup 3 > gpurun_out/r2a_ncu_bench.log 2>&1; echo "ncu list rc $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"nm_gemm|dw_gemm" -c 8 -o gpurun_out/r2a_prof python tools/profile_step.py 24576 1 > gpurun_out/r2a_ncu_full.log 2>&1; echo "ncu full rc $?"
timeout 200 python tools/step_timeline.py graph gpurun_out/r2a_update_timeline.md > gpurun_out/r2a_timeline.log 2>&1; echo "timeline rc $?"
ls -la gpurun_out
