#!/bin/bash
# The one GPU call of r02w (15 GPU-minutes were left in the round): full GPU test suite, default bench line, smoke(),
# A/B of the previous kernels' library against the new one on the SAME box, ncu --set full of the two dominant
# kernels, launch list of a short bench command.  Everything lands in gpurun_out/ and is summarised under profiles/.
#
#   gpurun --timeout 780 -- 'bash profiles/run_r02w.sh'
#
# build/libcity2ba_cuda_prev.so = the library built from the commit before r02w (CREDUX.F32 packet bounds, alive
# mask, opaque record address, zero-numerator perspective division).  If the new kernels fail a test the script
# falls back to it, so that the round still ends with a verified record of what is then the product.
cd "${GRAFT_REPO_ROOT:-/root/repo}" || exit 1
mkdir -p gpurun_out
OUT=gpurun_out
SO=city2ba_b200/libcity2ba_cuda.so
PREV=build/libcity2ba_cuda_prev.so
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/r02w_gpu.txt 2>&1

TAG=r02w
timeout 420 python -m pytest tests -m gpu -x -q > $OUT/r02w_pytest_gpu.txt 2>&1
RC=$?
tail -4 $OUT/r02w_pytest_gpu.txt
if [ $RC -ne 0 ] && [ -f $PREV ]; then
  echo "NEW KERNELS FAILED A TEST (rc $RC): falling back to the previous library"
  cp $SO build/libcity2ba_cuda_new.so
  cp $PREV $SO
  TAG=r02w_prev
  timeout 300 python -m pytest tests -m gpu -x -q > $OUT/r02w_prev_pytest_gpu.txt 2>&1
  tail -4 $OUT/r02w_prev_pytest_gpu.txt
fi

timeout 400 python bench.py > $OUT/${TAG}_bench_cfg4.json 2> $OUT/${TAG}_bench.err
echo "bench rc $?"; cut -c1-400 $OUT/${TAG}_bench_cfg4.json
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.txt 2>&1
echo "smoke rc $?"; tail -2 $OUT/${TAG}_smoke.txt

# A/B on this box: resident passes (grid rebuilt every pass) with the new and with the previous library
timeout 150 python profiles/shard_probe.py --workload cfg4 --shard 0/1 --steps 10 --drop-grid > $OUT/${TAG}_ab_new.txt 2>&1
tail -1 $OUT/${TAG}_ab_new.txt
if [ "$TAG" = r02w ] && [ -f $PREV ]; then
  cp $SO build/libcity2ba_cuda_new.so
  cp $PREV $SO
  timeout 150 python profiles/shard_probe.py --workload cfg4 --shard 0/1 --steps 10 --drop-grid > $OUT/r02w_ab_prev.txt 2>&1
  tail -1 $OUT/r02w_ab_prev.txt
  cp build/libcity2ba_cuda_new.so $SO
fi

# ncu --set full of the fused kernel and of sort + write (the 4th pass: 3 warm-up passes x 2 kernels skipped)
timeout 240 ncu --set full --import-source on --clock-control none -k regex:'k_visibility_fused|k_sort_write' \
  --launch-skip 6 -c 2 -f -o $OUT/prof_${TAG}_cfg4 python profiles/shard_probe.py --workload cfg4 --shard 0/1 --steps 1 \
  > $OUT/${TAG}_ncu.log 2>&1
echo "ncu rc $?"; ls -la $OUT/prof_${TAG}_cfg4.ncu-rep

# launch list of a short bench command (cold cache, serialised: compare shares)
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_${TAG}_cfg4.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-exhaustive --no-noise --no-secondary --no-extras \
  > $OUT/${TAG}_launch_bench.log 2>&1
echo "launch list rc $?"; wc -l $OUT/launches_${TAG}_cfg4.csv
