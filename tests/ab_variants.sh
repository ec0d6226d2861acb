#!/bin/bash
# A/B harness (run on the GPU box): the GPU test suite, then phases_ms / top-K / e2e per library variant
#   AB_VARIANTS="_pf8 _tc4" bash tests/ab_variants.sh   (variants built with foodrec_b200._build.build_variant)
#   AB_NCU=1 adds an ncu launch list of the same bench command
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
python -m pytest tests -m gpu -x -q > gpurun_out/ab_tests.log 2>&1; tail -3 gpurun_out/ab_tests.log
for v in "" $AB_VARIANTS; do
  FOODREC_B200_LIB=$PWD/foodrec_b200/csrc/libfoodrec_b200$v.so python bench.py --no-cpu --no-catalog --steps 30 --warmup 5 > gpurun_out/ab$v.json 2> gpurun_out/ab$v.err
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/ab$v.json") if l.startswith("{")][-1])
print("variant '$v'", round(d["ms_per_step"],4), {k:round(x,4) for k,x in d["phases_ms"].items()}, "topk ms", round(d["topk"]["ms"],3), "e2e", round(d["e2e"]["value"]/1e6,1), round(d["e2e_compact"]["value"]/1e6,1))
PY
done
if [ -n "$AB_NCU" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/ab_launches.csv \
    -k regex:"label|seg_|radix|fwd_train|scan_|item_catchup|finalize|gcat|prep_rows|series|mean_|write_counters" \
    python bench.py --no-cpu --no-catalog --steps 2 --warmup 1 --preroll 6 > gpurun_out/ab_ncu.log 2>&1
  tail -2 gpurun_out/ab_ncu.log
fi
