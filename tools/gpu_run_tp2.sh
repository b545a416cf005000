mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda(); print('warm', torch.cuda.device_count())"
(timeout 400 python -m pytest tests/test_gpu_tp.py -x -q > gpurun_out/t_tp.log 2>&1; echo "rc=$?" >> gpurun_out/t_tp.log)
tail -n 4 gpurun_out/t_tp.log
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/tp_bench.py 65b 30 > gpurun_out/tp2_65b.log 2>&1; echo "rc=$?" >> gpurun_out/tp2_65b.log)
tail -n 3 gpurun_out/tp2_65b.log
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/tp_bench.py 7b 50 > gpurun_out/tp2_7b.log 2>&1; echo "rc=$?" >> gpurun_out/tp2_7b.log)
tail -n 3 gpurun_out/tp2_7b.log
