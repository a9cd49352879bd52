#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/test_gpu_multi.log 2>&1
echo "multi rc=$? $(tail -1 gpurun_out/test_gpu_multi.log)"; grep -E "Error|assert" gpurun_out/test_gpu_multi.log | head -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 20 --warmup 3 \
   > gpurun_out/r2_scale_n2_weak_head.json 2> gpurun_out/scale2.err
echo "bench2 rc=$?"; cut -c1-330 gpurun_out/r2_scale_n2_weak_head.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --steps 20 --warmup 3 --key from_label \
   > gpurun_out/r2_scale_n2_weak_from_label.json 2> gpurun_out/scale2b.err
echo "bench2 from_label rc=$?"; cut -c1-330 gpurun_out/r2_scale_n2_weak_from_label.json; tail -3 gpurun_out/scale2b.err
