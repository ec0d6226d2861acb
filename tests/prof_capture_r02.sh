#!/bin/bash
# round-2 ncu captures of bench.py's own command (run on the GPU box AFTER the plain bench exited 0):
#   launch list (gpu__time_duration.sum) + one --set full capture of each train kernel at steady state + the eval kernel
#   (+ the catalog GEMM unless SKIP_CATALOG=1); summaries: profiles/summarize.py full | traffic
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TAG=${1:-r02}
python bench.py --no-cpu --no-catalog --no-cfg5 --steps 10 --warmup 5 > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches.csv \
  -k regex:"label|seg_|radix|fwd_train|scan|item_catchup|finalize|gcat|prep_rows|series|mean_|write_counters|user_fused|user_commit|shadow_" \
  python bench.py --no-cpu --no-catalog --no-cfg5 --steps 2 --warmup 3 --preroll 6 > gpurun_out/${TAG}_ncu_launches.log 2>&1
# one launch of each train kernel after the pre-roll (40 + 5 steps)
ncu --set full --clock-control none --import-source on -k regex:"user_fused_kernel|seg_tile_kernel|seg_chunk_kernel" \
  --launch-skip 150 -c 4 -f -o gpurun_out/${TAG}_train \
  python bench.py --no-cpu --no-catalog --no-cfg5 --steps 5 --warmup 5 > gpurun_out/${TAG}_ncu_train.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"eval_sampled_kernel" -c 1 -f -o gpurun_out/${TAG}_eval \
  python bench.py --no-cpu --no-catalog --no-cfg5 --steps 2 --warmup 3 --preroll 2 > gpurun_out/${TAG}_ncu_eval.log 2>&1
# the catalog GEMM at cfg2 (first launch after the training legs) -- bench's own command with the catalog leg on
[ -n "$SKIP_CATALOG" ] || ncu --set full --clock-control none --import-source on -k regex:"catalog_gemm_kernel" -c 1 -f -o gpurun_out/${TAG}_catalog \
  python bench.py --no-cpu --steps 2 --warmup 3 --preroll 2 > gpurun_out/${TAG}_ncu_catalog.log 2>&1
ls -la gpurun_out/${TAG}_*
