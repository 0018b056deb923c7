# Round-2 closing measurement pass on ONE B200 (gpurun): everything lands in gpurun_out/final4/ and is summarised into profiles/r2_*.
set -x
O=gpurun_out/final5
mkdir -p $O
(time timeout 900 python -m pytest tests -m gpu -q) > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $O/smoke.txt 2>&1; tail -4 $O/smoke.txt
python bench.py --steps 50 --warmup 5 > $O/bench_kitti_b1.json 2> $O/bench_kitti_b1.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
python bench.py --steps 20 --warmup 3 --workload sintel_436x1024_b8 --lanes 2 --no-cpu-baseline --no-train > $O/bench_sintel_b8.json 2> $O/bench_sintel_b8.err
python bench.py --steps 10 --warmup 3 --workload hd_1080x1920_b2 --lanes 3 --no-cpu-baseline --no-train > $O/bench_hd_b2.json 2> $O/bench_hd_b2.err
python bench.py --steps 20 --warmup 3 --precision tf32x3 --no-cpu-baseline --no-train > $O/bench_kitti_b1_tf32x3.json 2> $O/bench_kitti_b1_tf32x3.err
python tools/profile_step.py > $O/launch_table_kitti_events.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --lanes 1 --no-cpu-baseline --no-train > $O/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_win -s 1 -c 1 -o $O/prof_conv_win_kxn -f python tools/run_kernel.py conv 544 32 > $O/ncu1.log 2>&1

ls -la $O
python bench.py --steps 8 --warmup 3 --workload train_256x832_b4 --train-losses all --no-cpu-baseline > $O/bench_train_b4_all_losses.json 2> $O/bench_train_b4_all_losses.err
