#!/bin/bash
# two GPUs: multi-GPU tests (partitioned 2-D Euler now runs the triangle kernels), torchrun parity worker, headline bench at N = 2
O=gpurun_out/s4n2; mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_distributed.py -m gpu -x -q ) > $O/multi_tests_n2.log 2>&1
tail -6 $O/multi_tests_n2.log
NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_parity_worker.py > $O/dist_parity_nccl_n2.log 2>&1
tail -6 $O/dist_parity_nccl_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err
tail -3 $O/bench_n2.err; cut -c1-600 $O/bench_n2.json
