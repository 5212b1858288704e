set -x
mkdir -p gpurun_out/final
(time python -m pytest tests -m gpu -q) > gpurun_out/final/tests.log 2>&1; tail -3 gpurun_out/final/tests.log
python __graft_entry__.py smoke > gpurun_out/final/smoke.log 2>&1; tail -1 gpurun_out/final/smoke.log
python bench.py --impl reference 2>gpurun_out/final/ref.err | tail -1 > gpurun_out/final/bench_ref_cfg4.json
python bench.py 2>gpurun_out/final/cfg4.err | tail -1 > gpurun_out/final/bench_cfg4.json
for w in cfg2 cfg3 cfg5; do python bench.py --workload $w 2>gpurun_out/final/$w.err | tail -1 > gpurun_out/final/bench_$w.json; done
head -c 400 gpurun_out/final/bench_cfg4.json; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final/launches_cfg4.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/final/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sampler_kernel -s 3 -c 1 -o gpurun_out/final/prof_hmc_bench -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/final/ncu_full_cfg4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dp_eval_tc_kernel -s 5 -c 1 -o gpurun_out/final/prof_dp_tc_bench -f python bench.py --workload cfg5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/final/ncu_full_cfg5.log 2>&1
ls -la gpurun_out/final
