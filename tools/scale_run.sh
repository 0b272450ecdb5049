#!/bin/bash
# One 8-GPU lease: the 1/2/4/8-GPU curve of bench.py (+ the overlap variant at 8), the peer-exchange checks at 8 ranks.
# usage: tools/scale_run.sh <tag>     (writes gpurun_out/<tag>_*.json/.log)
tag=${1:-scale}
run() {  # n extra-env...
  n=$1; shift
  if [ "$n" = 1 ]; then
    env "$@" CPC_B200_PEER_STAMPS=1 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-torch-gpu --no-train-py
  else
    env "$@" CPC_B200_PEER_STAMPS=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) \
      bench.py --gpus $n --steps 30 --warmup 5 --no-cpu-baseline --no-torch-gpu --no-train-py
  fi
}
for n in 1 2 4 8; do
  run $n X=1 > gpurun_out/${tag}_n$n.json 2> gpurun_out/${tag}_n$n.log
  grep "timed\|stamps" gpurun_out/${tag}_n$n.log
done
run 8 CPC_B200_PEER_OVERLAP=1 > gpurun_out/${tag}_n8_overlap.json 2> gpurun_out/${tag}_n8_overlap.log
grep "timed\|stamps" gpurun_out/${tag}_n8_overlap.log
run 8 CPC_B200_FUSED_AR=0 > gpurun_out/${tag}_n8_nccl.json 2> gpurun_out/${tag}_n8_nccl.log
grep "timed" gpurun_out/${tag}_n8_nccl.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29620 tools/peer_adam_check.py --timeout-test > gpurun_out/${tag}_peer_check_n8.log 2>&1
tail -22 gpurun_out/${tag}_peer_check_n8.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 tools/replica_check.py > gpurun_out/${tag}_replicas_n8.log 2>&1
tail -4 gpurun_out/${tag}_replicas_n8.log
python -m pytest tests -m gpu -q -k "peer_adam" -p no:cacheprovider 2>&1 | tail -2
