#!/bin/bash
# weak-scaling run of bench.py on one multi-GPU box: N = 1, 2, 4, 8 (as many as are visible)
tag=$1
out=gpurun_out
mkdir -p $out
ngpu=$(nvidia-smi -L | wc -l)
for n in 1 2 4 8; do
  [ $n -gt $ngpu ] && break
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-gpu-eager > $out/${tag}_scale_n1.json 2> $out/${tag}_scale_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) \
      bench.py --gpus $n --steps 200 --warmup 10 > $out/${tag}_scale_n${n}.json 2> $out/${tag}_scale_n${n}.err
  fi
  echo "rc=$?" >> $out/${tag}_scale_n${n}.err
done
nvidia-smi topo -m > $out/${tag}_topo.txt 2>&1
