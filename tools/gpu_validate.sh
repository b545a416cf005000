mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda(); print('warm')"
(timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/t_all_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/t_all_gpu.log)
tail -n 8 gpurun_out/t_all_gpu.log
(timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/smoke.log)
tail -n 2 gpurun_out/smoke.log
(timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.log 2>&1; echo "rc=$?" >> gpurun_out/bench_ref.log)
tail -n 2 gpurun_out/bench_ref.log | cut -c1-400
(timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"llama_decode|qmatvec|qgemm|attn_|lm_head|rope_|rmsnorm" -c 40 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 4 --warmup 3 > gpurun_out/ncu_launches.log 2>&1; echo "rc=$?" >> gpurun_out/ncu_launches.log)
tail -n 1 gpurun_out/ncu_launches.log | cut -c1-200
