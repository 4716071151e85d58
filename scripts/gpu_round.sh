# One GPU-box round: parity tests, per-shape contraction timings, the bench line with phases, the step timeline,
# an ncu launch list of the bench command and one `ncu --set full` capture of the dominant kernel.
#   SKIP_TESTS / SKIP_GEMM / SKIP_NCU / SKIP_TRACE = 1 skip the respective leg.
TAG=${1:-x}
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then python -m pytest tests -m gpu -x -q 2>&1 | tail -15; fi
if [ -z "$SKIP_GEMM" ]; then python scripts/time_gemm.py > gpurun_out/time_gemm_$TAG.txt 2>&1; head -14 gpurun_out/time_gemm_$TAG.txt; fi
python bench.py --steps 20 --warmup 5 --phases > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; cut -c1-400 gpurun_out/bench_$TAG.json; grep phases gpurun_out/bench_$TAG.err
tail -3 gpurun_out/bench_$TAG.err
if [ -z "$SKIP_TRACE" ]; then python scripts/trace_step.py > gpurun_out/trace_$TAG.txt 2>&1; head -40 gpurun_out/trace_$TAG.txt; fi
if [ -z "$SKIP_NCU" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu_$TAG.log 2>&1
  python scripts/launch_summary.py gpurun_out/launches_$TAG.csv 40 > gpurun_out/launch_summary_$TAG.txt 2>&1; head -45 gpurun_out/launch_summary_$TAG.txt
  ncu --set full --clock-control none --import-source on -k regex:gemm_tn_tc -s 3 -c 2 -o gpurun_out/prof_gemm_tn_tc_$TAG -f python scripts/profile_gemm.py > gpurun_out/prof_gemm_$TAG.log 2>&1
  ncu -i gpurun_out/prof_gemm_tn_tc_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_gemm_tn_tc_$TAG.raw.csv 2>/dev/null
fi
