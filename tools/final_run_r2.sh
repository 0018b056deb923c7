# Round-2 measurement pass on ONE B200 (gpurun): everything lands in gpurun_out/final2/ and is summarised into profiles/r2_*.
set -x
O=gpurun_out/final2
mkdir -p $O
(time timeout 1500 python -m pytest tests -m gpu -q) > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $O/smoke.txt 2>&1; tail -4 $O/smoke.txt
python bench.py --steps 50 --warmup 5 > $O/bench_kitti_b1.json 2> $O/bench_kitti_b1.err
python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
python bench.py --steps 20 --warmup 3 --workload sintel_436x1024_b8 --no-cpu-baseline --no-train > $O/bench_sintel_b8.json 2> $O/bench_sintel_b8.err
python bench.py --steps 10 --warmup 3 --workload hd_1080x1920_b2 --no-cpu-baseline --no-train > $O/bench_hd_b2.json 2> $O/bench_hd_b2.err
python bench.py --steps 20 --warmup 3 --precision tf32x3 --no-cpu-baseline --no-train > $O/bench_kitti_b1_tf32x3.json 2> $O/bench_kitti_b1_tf32x3.err
python bench.py --steps 8 --warmup 3 --workload train_256x832_b4 --no-cpu-baseline > $O/bench_train_b4.json 2> $O/bench_train_b4.err
python bench.py --steps 8 --warmup 3 --workload train_256x832_b4 --train-losses all --no-cpu-baseline > $O/bench_train_b4_all_losses.json 2> $O/bench_train_b4_all_losses.err
python tools/profile_train.py > $O/train_breakdown.txt 2> $O/train_breakdown.err
python tools/profile_step.py > $O/launch_table_kitti_events.txt 2>&1
python tools/time_corr.py > $O/time_corr.txt 2>&1
python tools/time_corr_planar.py > $O/time_corr_planar.txt 2>&1
python tools/dbg_planar_time.py > $O/corr_planar_ablation.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train > $O/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:corr_planar -s 1 -c 1 -o $O/prof_corr_planar -f python tools/run_kernel.py corr_planar > $O/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:corr_pipe -s 1 -c 1 -o $O/prof_corr_pipe -f python tools/run_kernel.py corr > $O/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 4 -c 1 -o $O/prof_wgrad_tc -f python tools/prof_wgrad.py 576 128 > $O/ncu3.log 2>&1
ls -la $O
