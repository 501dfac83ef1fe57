#!/bin/bash
# The GPU call of r02x: same-box A/B of four builds of the library, then the full record for the fastest one.
#
#   gpurun --timeout 600 -- 'bash profiles/run_r02x.sh'
#
# build/lib_<v>.so, each built from a commit of this repository (nvcc flags of __graft_entry__.build()):
#   v0    r02w (90b5930): the kernels the previous call validated
#   v1    + one shared-memory block per warp behind an opaque base in k_visibility_fused      (b481463)
#   v12a  + k_sort_write: OR / AND skip detection, scatter four chunks at a time, __ldg gathers (74ad07b)
#   v12b  + k_sort_write: camera record in shared memory, point gather one iteration ahead      (the commit after)
# The fastest build (resident pass at cfg4, grid rebuilt every pass; a later variant must win by > 0.3 %) becomes
# city2ba_b200/libcity2ba_cuda.so for the test suite, the bench line, smoke() and the ncu capture; the repository's
# sources are then set to that build's commit.  A build that fails a test hands over to the next one down (v12b / v12a -> v1 -> v0).
cd "${GRAFT_REPO_ROOT:-/root/repo}" || exit 1
mkdir -p gpurun_out
OUT=gpurun_out
SO=city2ba_b200/libcity2ba_cuda.so
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > $OUT/r02x_gpu.txt 2>&1

for v in v12b v0 v1 v12a v12b; do   # (the first run also pays the cold import; v12b is measured again at the end)
  cp build/lib_$v.so $SO
  timeout 150 python profiles/shard_probe.py --workload cfg4 --shard 0/1 --steps 10 --drop-grid > $OUT/r02x_ab_$v.txt 2>&1
  echo "$v $(tail -1 $OUT/r02x_ab_$v.txt)"
done

BEST=$(python - <<'EOF'
import json
order = ["v0", "v1", "v12a", "v12b"]
ms = {}
for v in order:
    try:
        ms[v] = json.loads(open(f"gpurun_out/r02x_ab_{v}.txt").read().strip().split("\n")[-1])["ms_total"]
    except Exception:
        pass
best = "v0"
for v in order[1:]:
    if v in ms and best in ms and ms[v] < ms[best] * 0.997:
        best = v
print(best)
EOF
)
echo "fastest build: $BEST"
echo "$BEST" > $OUT/r02x_best.txt
cp build/lib_$BEST.so $SO

# test the fastest build; if it fails a test, the next one down the list (v12b / v12a -> v1 -> v0)
CHAIN="$BEST"
case $BEST in v12b|v12a) CHAIN="$BEST v1 v0";; v1) CHAIN="v1 v0";; esac
for v in $CHAIN; do
  cp build/lib_$v.so $SO
  timeout 420 python -m pytest tests -m gpu -x -q > $OUT/r02x_pytest_gpu_$v.txt 2>&1
  RC=$?
  tail -4 $OUT/r02x_pytest_gpu_$v.txt
  if [ $RC -eq 0 ]; then echo "$v" > $OUT/r02x_final.txt; break; fi
  echo "BUILD $v FAILED A TEST (rc $RC)"
done
echo "final build: $(cat $OUT/r02x_final.txt)"

timeout 400 python bench.py > $OUT/r02x_bench_cfg4.json 2> $OUT/r02x_bench.err
echo "bench rc $?"; cut -c1-330 $OUT/r02x_bench_cfg4.json
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r02x_smoke.txt 2>&1
echo "smoke rc $?"; tail -2 $OUT/r02x_smoke.txt

timeout 240 ncu --set full --import-source on --clock-control none -k regex:'k_visibility_fused|k_sort_write' \
  --launch-skip 6 -c 2 -f -o $OUT/prof_r02x_cfg4 python profiles/shard_probe.py --workload cfg4 --shard 0/1 --steps 1 \
  > $OUT/r02x_ncu.log 2>&1
echo "ncu rc $?"; ls -la $OUT/prof_r02x_cfg4.ncu-rep
