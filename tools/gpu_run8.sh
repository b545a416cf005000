mkdir -p gpurun_out
D=gptq-for-llama_b200/dev
(timeout 100 tools/ubench/barrier > gpurun_out/ubench_barrier.log 2>&1; echo "rc=$?" >> gpurun_out/ubench_barrier.log)
python -c "import torch; torch.zeros(1).cuda(); print('warm')"
(timeout 200 python tools/quick_bench.py 7b > gpurun_out/qb_7b.log 2>&1; echo "rc=$?" >> gpurun_out/qb_7b.log)
(GPTQ_B200_LIB=$D/libgptq_b200_trace.so timeout 200 python tools/trace_mega.py 7b > gpurun_out/trace_7b.log 2>&1; echo "rc=$?" >> gpurun_out/trace_7b.log)
cat gpurun_out/ubench_barrier.log gpurun_out/qb_7b.log; tail -24 gpurun_out/trace_7b.log
