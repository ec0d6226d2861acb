#!/bin/bash
# ncu: launch list of the final build + one --set full capture of the sampled-evaluation kernel (run on the GPU box)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TAG=${1:-r01c}
python bench.py --no-cpu --no-catalog --steps 3 --warmup 3 --preroll 4 > gpurun_out/${TAG}_plain_small.json 2> gpurun_out/${TAG}_plain_small.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/${TAG}_launches.csv \
  -k regex:"label|seg_|radix|fwd_train|scan|item_catchup|finalize|gcat|prep_rows|series|mean_|write_counters" \
  python bench.py --no-cpu --no-catalog --steps 2 --warmup 3 --preroll 4 > gpurun_out/${TAG}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"eval_sampled_kernel" -c 1 -f -o gpurun_out/${TAG}_eval \
  python bench.py --no-cpu --no-catalog --steps 2 --warmup 3 --preroll 2 > gpurun_out/${TAG}_ncu_eval.log 2>&1
ls -la gpurun_out/${TAG}_*
