mkdir -p gpurun_out
D=gptq-for-llama_b200/dev
python -c "import torch; torch.zeros(1).cuda(); print('warm')"
(timeout 200 python tools/quick_bench.py 7b > gpurun_out/qb_7b.log 2>&1; echo "rc=$?" >> gpurun_out/qb_7b.log)
(GPTQ_B200_LIB=$D/libgptq_b200_trace.so timeout 200 python tools/trace_mega.py 7b > gpurun_out/trace_7b.log 2>&1; echo "rc=$?" >> gpurun_out/trace_7b.log)
(timeout 600 python -m pytest tests/test_gpu_engine.py -x -q > gpurun_out/t_engine.log 2>&1; echo "rc=$?" >> gpurun_out/t_engine.log)
(timeout 1200 python -m pytest tests/test_gpu_engine_fullsize.py -q -s > gpurun_out/t_full.log 2>&1; echo "rc=$?" >> gpurun_out/t_full.log)
(timeout 300 ncu --set full --clock-control none --import-source on -k regex:llama_decode_mega -s 2 -c 1 -f -o gpurun_out/r2_mega_v2b python tools/prof_mega.py 7b 8 > gpurun_out/ncu.log 2>&1; echo "rc=$?" >> gpurun_out/ncu.log)
tail -n 3 gpurun_out/qb_7b.log gpurun_out/t_engine.log gpurun_out/ncu.log; grep -E "passed|failed|Error|spread" gpurun_out/t_full.log | head -20
