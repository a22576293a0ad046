#!/bin/bash
# weak-scaling run of bench.py on one multi-GPU box: N = 1, 2, 4, 8 (as many as are visible; NS overrides the list),
# then the configurations listed in CONFIGS_N at the largest N (config 4: sharded select, config 5: independent shards)
tag=$1
out=gpurun_out
mkdir -p $out
ngpu=$(nvidia-smi -L | wc -l)
for n in ${NS:-1 2 4 8}; do
  [ $n -gt $ngpu ] && continue
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-gpu-eager > $out/${tag}_scale_n1.json 2> $out/${tag}_scale_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) \
      bench.py --gpus $n --steps 200 --warmup 10 > $out/${tag}_scale_n${n}.json 2> $out/${tag}_scale_n${n}.err
  fi
  echo "rc=$?" >> $out/${tag}_scale_n${n}.err
done
for c in ${CONFIGS_N:-}; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $ngpu --master-addr 127.0.0.1 --master-port $((29700+c)) \
    bench.py --config $c --gpus $ngpu --steps 100 --warmup 5 > $out/${tag}_c${c}_n${ngpu}.json 2> $out/${tag}_c${c}_n${ngpu}.err
  echo "rc=$?" >> $out/${tag}_c${c}_n${ngpu}.err
done
nvidia-smi topo -m > $out/${tag}_topo.txt 2>&1
