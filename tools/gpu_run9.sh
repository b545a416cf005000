mkdir -p gpurun_out
D=gptq-for-llama_b200/dev
python -c "import torch; torch.zeros(1).cuda(); print('warm')"
(timeout 200 python tools/quick_bench.py 7b > gpurun_out/qb_7b.log 2>&1; echo "rc=$?" >> gpurun_out/qb_7b.log)
(GPTQ_B200_LIB=$D/libgptq_b200_trace.so timeout 200 python tools/trace_mega.py 7b > gpurun_out/trace_7b.log 2>&1; echo "rc=$?" >> gpurun_out/trace_7b.log)
(timeout 600 python -m pytest tests/test_gpu_engine.py -x -q > gpurun_out/t_engine.log 2>&1; echo "rc=$?" >> gpurun_out/t_engine.log)
(timeout 1200 python -m pytest tests/test_gpu_engine_fullsize.py -q -s > gpurun_out/t_full.log 2>&1; echo "rc=$?" >> gpurun_out/t_full.log)
(timeout 200 python tools/quick_bench.py 13b 3 act > gpurun_out/qb_13b.log 2>&1; echo "rc=$?" >> gpurun_out/qb_13b.log)
(timeout 300 python tools/quick_bench.py 65b > gpurun_out/qb_65b.log 2>&1; echo "rc=$?" >> gpurun_out/qb_65b.log)
tail -n 3 gpurun_out/qb_7b.log gpurun_out/qb_13b.log gpurun_out/qb_65b.log gpurun_out/t_engine.log; grep -E "passed|failed|Error|spread" gpurun_out/t_full.log | head -20
