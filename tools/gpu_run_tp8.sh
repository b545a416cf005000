mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda(); print('warm', torch.cuda.device_count())"
(timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tools/tp_bench.py 65b 30 > gpurun_out/tp8_65b.log 2>&1; echo "rc=$?" >> gpurun_out/tp8_65b.log)
tail -n 3 gpurun_out/tp8_65b.log
(timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 tools/tp_bench.py 65b 30 > gpurun_out/tp4_65b.log 2>&1; echo "rc=$?" >> gpurun_out/tp4_65b.log)
tail -n 3 gpurun_out/tp4_65b.log
