mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda(); print('warm')"
(timeout 420 ncu --set full --clock-control none --import-source on -k regex:llama_decode_mega -s 2 -c 1 -f -o gpurun_out/r2_mega_v2a python tools/prof_mega.py 7b 8 > gpurun_out/ncu.log 2>&1; echo "rc=$?" >> gpurun_out/ncu.log)
tail -5 gpurun_out/ncu.log; ls -la gpurun_out/*.ncu-rep
