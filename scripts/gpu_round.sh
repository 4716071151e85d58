# One GPU-box round: parity tests, per-shape contraction timings, the bench line with phases, and an ncu launch list.
TAG=${1:-x}
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
if [ -z "$SKIP_GEMM" ]; then python scripts/time_gemm.py > gpurun_out/time_gemm_$TAG.txt 2>&1; head -14 gpurun_out/time_gemm_$TAG.txt; fi
python bench.py --steps 10 --warmup 3 --phases > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; cut -c1-330 gpurun_out/bench_$TAG.json; grep phases gpurun_out/bench_$TAG.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_ncu_$TAG.log 2>&1
tail -3 gpurun_out/bench_$TAG.err
