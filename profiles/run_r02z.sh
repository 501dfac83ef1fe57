#!/bin/bash
# The last GPU call of the round (3.9 GPU-minutes were left): the final commit's test suite — now with (u, v) compared
# as raw 64-bit words against the oracle (sign of zero) — a short bench line whose `parity` uses the same bit-level
# comparison, and an `ncu --set full` capture of the final kernels at cfg5 (the BVH-walk variant of the fused kernel).
#
#   gpurun --timeout 230 -- 'bash profiles/run_r02z.sh'
cd "${GRAFT_REPO_ROOT:-/root/repo}" || exit 1
mkdir -p gpurun_out
OUT=gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q > $OUT/r02z_pytest_gpu.txt 2>&1
echo "pytest rc $?"; tail -4 $OUT/r02z_pytest_gpu.txt
timeout 120 python bench.py --steps 5 --warmup 3 --no-exhaustive --no-noise --no-secondary --no-extras \
  > $OUT/r02z_bench_cfg4_short.json 2> $OUT/r02z_bench.err
echo "bench rc $?"; python - <<'EOF'
import json
try:
    d = json.load(open("gpurun_out/r02z_bench_cfg4_short.json"))
    print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["parity"], d["result_hash"]["ok"])
except Exception as e:
    print("no bench line:", e)
EOF
timeout 100 ncu --set full --import-source on --clock-control none -k regex:'k_visibility_fused|k_sort_write' \
  --launch-skip 6 -c 2 -f -o $OUT/prof_r02z_cfg5 python profiles/shard_probe.py --workload cfg5 --shard 0/1 --steps 1 \
  > $OUT/r02z_ncu.log 2>&1
echo "ncu rc $?"; ls -la $OUT/prof_r02z_cfg5.ncu-rep
