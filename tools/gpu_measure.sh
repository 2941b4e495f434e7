#!/bin/bash
# Round measurement sequence on the GPU box: tests, smoke, bench (both arms), ncu launch list of the bench command,
# one ncu --set full capture per hot kernel.  Outputs go to gpurun_out/ (tag = $1).
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err || tail -5 $OUT/${TAG}_bench_n1.err
tail -c 2500 $OUT/${TAG}_bench_n1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err || tail -5 $OUT/${TAG}_bench_ref.err
tail -c 1200 $OUT/${TAG}_bench_ref.json
# launch list of the same bench command (shorter: 2 steps)
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_plain.json 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_launches.log 2>&1
# full captures of the hot kernels at the bench launch configuration (first launch after the warm-up sweeps)
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:rb_estimator_pair_kernel -s 3 -c 1 \
    -o $OUT/${TAG}_estimator_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_est.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:rb_lu_mw_kernel -s 40 -c 1 \
    -o $OUT/${TAG}_lu_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_lu.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:rb_assemble_rec_kernel -s 40 -c 1 \
    -o $OUT/${TAG}_asm_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_asm.log 2>&1
# offline basis linear algebra: kernel times + launch list + full captures of the two Gram kernels
timeout -s KILL 600 python bench_offline.py > $OUT/${TAG}_offline_kernels.jsonl 2> $OUT/${TAG}_offline.err || tail -5 $OUT/${TAG}_offline.err
cut -c1-260 $OUT/${TAG}_offline_kernels.jsonl
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $OUT/${TAG}_offline_launches.csv \
    python bench_offline.py --reps 1 > $OUT/${TAG}_ncu_offline.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:rb_gram_fused_kernel -c 1 \
    -o $OUT/${TAG}_gram_fused_full -f python tools/gram_probe.py 61 > $OUT/${TAG}_ncu_fused.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:rb_tsmm2_kernel -c 1 \
    -o $OUT/${TAG}_tsmm2_full -f python tools/gram_probe.py 400 > $OUT/${TAG}_ncu_tsmm2.log 2>&1
# POD-Greedy (parabolic) sweep: bench line, launch list, full capture of the fused kernel
timeout -s KILL 300 python bench_parabolic.py > $OUT/${TAG}_parabolic_sweep.json 2> $OUT/${TAG}_parabolic.err || tail -5 $OUT/${TAG}_parabolic.err
cut -c1-400 $OUT/${TAG}_parabolic_sweep.json
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $OUT/${TAG}_parabolic_launches.csv \
    python bench_parabolic.py --cpu-sample 4 > $OUT/${TAG}_ncu_parabolic.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:rb_parabolic_fused -s 1 -c 1 \
    -o $OUT/${TAG}_parabolic_full -f python bench_parabolic.py --cpu-sample 4 > $OUT/${TAG}_ncu_parabolic_full.log 2>&1
ls -la $OUT | tail -12
