#!/bin/bash
mkdir -p gpurun_out/r2c
O=gpurun_out/r2c
nvidia-smi topo -m > $O/topo.txt 2>&1
( time python -m pytest tests/test_gpu_multi.py tests/test_gpu_distributed.py -m gpu -x -q ) > $O/gputests_n2.log 2>&1
tail -12 $O/gputests_n2.log
NCCL_DEBUG=WARN python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_parity_worker.py > $O/dist_parity_n2.log 2>&1
tail -8 $O/dist_parity_n2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err
tail -3 $O/bench_n2.err; cat $O/bench_n2.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $O/bench_ref_n2.json 2> $O/bench_ref_n2.err
cat $O/bench_ref_n2.json | cut -c1-300
