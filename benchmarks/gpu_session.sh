#!/bin/bash
# Development helper: one gpurun call = a list of stages.  Everything lands in gpurun_out/<tag>_*.
#   gpurun --timeout 1500 -- 'bash benchmarks/gpu_session.sh r2a tests variants bench configs launches'
tag=$1; shift
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > $out/${tag}_gpu.txt 2>&1
for stage in "$@"; do
  case $stage in
    tests)
      timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $out/${tag}_tests.txt ;;
    smoke)
      timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.txt 2>&1 ;;
    variants)
      for v in 2 0; do for h in 1 0; do
        timeout 300 python bench.py --steps 500 --warmup 10 --no-cpu-baseline --no-gpu-eager --row-variant $v --keep-hint $h \
          > $out/${tag}_bench_v${v}_h${h}.json 2> $out/${tag}_bench_v${v}_h${h}.err
      done; done ;;
    bench)
      timeout 600 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err
      timeout 600 python bench.py --steps 20 --warmup 5 > $out/${tag}_bench_s20.json 2>> $out/${tag}_bench.err
      timeout 600 python bench.py --sparsity 0 --no-cpu-baseline > $out/${tag}_bench_dense.json 2>> $out/${tag}_bench.err
      timeout 600 python bench.py --mode eager --no-cpu-baseline --no-gpu-eager > $out/${tag}_bench_eager.json 2>> $out/${tag}_bench.err ;;
    probe)
      timeout 300 python benchmarks/bw_probe.py > $out/${tag}_bw_probe.jsonl 2>&1 ;;
    multi)
      timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -25 > $out/${tag}_tests_multi.txt
      for n in 2; do
        timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
          bench.py --gpus $n --steps 500 --warmup 10 > $out/${tag}_bench_n${n}.json 2> $out/${tag}_bench_n${n}.err
        echo "rc=$?" >> $out/${tag}_bench_n${n}.err
      done ;;
    sanitize)
      for tool in memcheck racecheck synccheck; do
        echo "== compute-sanitizer --tool $tool" >> $out/${tag}_sanitizer.txt
        timeout 900 compute-sanitizer --tool $tool python benchmarks/sanitize_target.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Error|error|sanitize target done" | head -30 >> $out/${tag}_sanitizer.txt
      done ;;
    ref)
      timeout 600 python bench.py --impl reference --steps 10 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err ;;
    configs)
      for c in ${CONFIGS:-1 3 4 5}; do
        timeout 900 python bench.py --config $c --steps 200 > $out/${tag}_bench_c${c}.json 2> $out/${tag}_bench_c${c}.err
      done ;;
    launches)
      timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv \
        --log-file $out/${tag}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-eager > $out/${tag}_launches.log 2>&1 ;;
    launches_nocc)
      timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none --cache-control none -c 120 --csv \
        --log-file $out/${tag}_launches_nocc.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-eager > $out/${tag}_launches_nocc.log 2>&1 ;;
    ncufull)
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:"reduce_rows|map_chan_win" -c 12 \
        -o $out/${tag}_step python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-eager > $out/${tag}_ncufull.log 2>&1 ;;
    *) echo "unknown stage $stage" ;;
  esac
done
ls -la $out | tail -40
