mkdir -p gpurun_out
(timeout 60 tools/ubench/loop > gpurun_out/ubench_loop.log 2>&1; echo "rc=$?" >> gpurun_out/ubench_loop.log)
cat gpurun_out/ubench_loop.log
