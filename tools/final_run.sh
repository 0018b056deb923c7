set -x
mkdir -p gpurun_out/final
(time timeout 900 python -m pytest tests -m gpu -q) > gpurun_out/final/pytest_gpu.log 2>&1; tail -3 gpurun_out/final/pytest_gpu.log
python bench.py --steps 50 --warmup 5 > gpurun_out/final/bench_kitti_b1.json 2> gpurun_out/final/bench_kitti_b1.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/final/bench_reference_arm.json 2> gpurun_out/final/bench_reference_arm.err
python bench.py --steps 20 --warmup 3 --workload sintel_436x1024_b8 --no-cpu-baseline > gpurun_out/final/bench_sintel_b8.json 2> gpurun_out/final/bench_sintel_b8.err
python bench.py --steps 10 --warmup 3 --workload hd_1080x1920_b2 --no-cpu-baseline > gpurun_out/final/bench_hd_b2.json 2> gpurun_out/final/bench_hd_b2.err
python bench.py --steps 20 --warmup 3 --precision tf32x3 --no-cpu-baseline > gpurun_out/final/bench_kitti_b1_tf32x3.json 2> gpurun_out/final/bench_kitti_b1_tf32x3.err
python bench.py --steps 8 --warmup 3 --workload train_256x832_b4 > gpurun_out/final/bench_train_b4.json 2> gpurun_out/final/bench_train_b4.err
python bench.py --steps 8 --warmup 3 --workload train_256x832_b4 --train-losses all > gpurun_out/final/bench_train_b4_all_losses.json 2> gpurun_out/final/bench_train_b4_all_losses.err
python bench.py --steps 8 --warmup 3 --workload train_256x832_b4 --train-losses all --torch-losses > gpurun_out/final/bench_train_b4_all_losses_torch.json 2> gpurun_out/final/bench_train_b4_all_losses_torch.err
python tools/profile_train.py > gpurun_out/final/train_breakdown.txt 2> gpurun_out/final/train_breakdown.err
python tools/profile_step.py > gpurun_out/final/launch_table_kitti_events.txt 2>&1
python tools/time_corr.py > gpurun_out/final/time_corr.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/final/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/final/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:corr_pipe -s 1 -c 1 -o gpurun_out/final/prof_corr_pipe -f python tools/run_kernel.py corr > gpurun_out/final/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_halo -s 1 -c 1 -o gpurun_out/final/prof_conv_halo -f python tools/run_kernel.py conv 576 128 > gpurun_out/final/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_win -s 1 -c 1 -o gpurun_out/final/prof_conv_win -f python tools/run_kernel.py conv 544 32 > gpurun_out/final/ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 1 -c 1 -o gpurun_out/final/prof_conv_tc -f python tools/run_kernel.py conv 128 96 8 > gpurun_out/final/ncu4.log 2>&1
ls -la gpurun_out/final
