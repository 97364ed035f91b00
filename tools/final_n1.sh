set -x
mkdir -p gpurun_out/final
timeout 900 python -m pytest tests -q -m gpu -rs > gpurun_out/final/pytest_gpu.log 2>&1; tail -4 gpurun_out/final/pytest_gpu.log
python bench.py > gpurun_out/final/bench.json 2> gpurun_out/final/bench.err; tail -c 600 gpurun_out/final/bench.json
python bench.py --impl reference > gpurun_out/final/reference_arm.json 2> gpurun_out/final/reference_arm.err; tail -c 400 gpurun_out/final/reference_arm.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/final/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:word_encode_fused -s 2 -c 1 -o gpurun_out/final/fused_encode -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --e2e-steps 0 > gpurun_out/final/ncu_full.log 2>&1
ncu -i gpurun_out/final/fused_encode.ncu-rep --page raw --csv > gpurun_out/final/fused_encode_raw.csv 2>/dev/null
ls -la gpurun_out/final
python tools/pcie_probe.py --copies-only > gpurun_out/final/pcie_copies.log 2>&1; tail -2 gpurun_out/final/pcie_copies.log
