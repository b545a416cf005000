mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda(); print('warm', torch.cuda.device_count())"
(timeout 400 python -m pytest tests/test_gpu_tp.py -x -q > gpurun_out/t_tp.log 2>&1; echo "rc=$?" >> gpurun_out/t_tp.log)
tail -n 4 gpurun_out/t_tp.log
(timeout 200 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.log 2>&1; echo "rc=$?" >> gpurun_out/bench_ref.log)
tail -n 2 gpurun_out/bench_ref.log | cut -c1-600
