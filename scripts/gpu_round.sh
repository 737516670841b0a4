#!/bin/bash
# One GPU-box session: parity tests, bench (both arms), variant timings, ncu launch list + full captures.
# Usage (from the build container): gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh TAG'
TAG=${1:-r1x}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${TAG}_pytest.log
tail -3 $O/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/${TAG}_smoke.log
timeout 400 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"; cat $O/${TAG}_bench.json
timeout 300 python scripts/variant_bench.py C3 0,4,5 > $O/${TAG}_variants.log 2>&1
timeout 300 python scripts/variant_bench.py C2 0,4,5 >> $O/${TAG}_variants.log 2>&1
timeout 300 python scripts/variant_bench.py C4 0,4,5 >> $O/${TAG}_variants.log 2>&1
cat $O/${TAG}_variants.log
timeout 200 python scripts/stage_times.py C3 > $O/${TAG}_stages.log 2>&1; cat $O/${TAG}_stages.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 > $O/${TAG}_ncu_bench.log 2>&1
VBMC_ENTMC_VARIANT=4 timeout 400 ncu --set full --clock-control none --import-source on -k regex:entmc -s 2 -c 1 -f -o $O/${TAG}_entmc_w python scripts/run_entmc.py C3 4 > $O/${TAG}_ncu_w.log 2>&1
VBMC_ENTMC_VARIANT=5 timeout 400 ncu --set full --clock-control none --import-source on -k regex:entmc -s 4 -c 2 -f -o $O/${TAG}_entmc_tc python scripts/run_entmc.py C3 4 > $O/${TAG}_ncu_tc.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tail -s 2 -c 1 -f -o $O/${TAG}_tail python scripts/run_negelcbo.py C3 4 > $O/${TAG}_ncu_tail.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gplj -s 2 -c 1 -f -o $O/${TAG}_gplj python scripts/run_negelcbo.py C3 4 > $O/${TAG}_ncu_gplj.log 2>&1
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err; cat $O/${TAG}_bench_ref.json
ls -la $O
