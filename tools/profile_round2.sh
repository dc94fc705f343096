#!/bin/bash
# Round-2 measurement recipe (run under gpurun, one GPU).  Usage: tools/profile_round2.sh <tag> [stages...]
# stages: tests bench train ncu_gather ncu_bwd ncu_gemm ncu_layer launches   (default: all but ncu_gemm / ncu_layer)
TAG=${1:-r2}; shift
STAGES=${@:-tests bench train ncu_gather ncu_bwd launches}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > $OUT/${TAG}_clocks.csv &
SMI=$!
NCU="ncu --clock-control none"
for st in $STAGES; do
  case $st in
    tests)   timeout 1500 python -m pytest tests -m gpu -q -rP -p no:cacheprovider > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log; tail -5 $OUT/${TAG}_pytest.log ;;
    bench)   timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/${TAG}_bench_n3.json 2> $OUT/${TAG}_bench_n3.err; tail -c 600 $OUT/${TAG}_bench_n3.json ;;
    bench_all) for w in mind_small_dev_n5_L3 wide_n8_L7; do timeout 600 python bench.py --steps 20 --warmup 3 --no-train --workload $w > $OUT/${TAG}_bench_$w.json 2> $OUT/${TAG}_bench_$w.err; done ;;
    train)   timeout 600 python bench.py --mode train --steps 20 --warmup 3 > $OUT/${TAG}_bench_train.json 2> $OUT/${TAG}_bench_train.err; cat $OUT/${TAG}_bench_train.json ;;
    launches) DIGAT_PROFILE_RANGE=1 timeout 900 $NCU --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-train --sustained 0 > $OUT/${TAG}_launches.log 2>&1 ;;
    ncu_gather) timeout 900 $NCU --set full --import-source on -k regex:"gather_rows_kernel|build_user_nodes_kernel|logits_kernel|compact_lists|active_rows" -s 40 -c 14 -o $OUT/${TAG}_ncu_gather -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_gather.log 2>&1 ;;
    ncu_bwd) timeout 900 $NCU --set full --import-source on -k regex:"graph_layer_bwd_kernel|attention_pool_bwd|topic_segment_bwd|gemm_wgrad|colsum|groupsum" -s 60 -c 14 -o $OUT/${TAG}_ncu_bwd -f python bench.py --mode train --eager-train --steps 2 --warmup 3 > $OUT/${TAG}_ncu_bwd.log 2>&1
             timeout 900 $NCU --set full --import-source on -k regex:"gemm_tf32x3_persistent" -s 30 -c 4 -o $OUT/${TAG}_ncu_wgrad -f python bench.py --mode train --eager-train --steps 2 --warmup 3 > $OUT/${TAG}_ncu_wgrad.log 2>&1 ;;
    ncu_layerbwd) timeout 900 $NCU --set full --import-source on -k regex:"graph_layer_bwd_kernel" -s 4 -c 2 -o $OUT/${TAG}_ncu_layerbwd -f python bench.py --mode train --eager-train --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_layerbwd.log 2>&1 ;;
    ncu_gather2) timeout 900 $NCU --set full --import-source on -k regex:"gather_rows_kernel|build_user_nodes_kernel|graph_csr_kernel|compact_lists|active_rows|topic_segment|attention_pool|logits_kernel" -s 160 -c 24 -o $OUT/${TAG}_ncu_gather2 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train --sustained 0 > $OUT/${TAG}_ncu_gather2.log 2>&1 ;;
    ncu_gemm) timeout 900 $NCU --set full --import-source on -k regex:"gemm_tf32x3_persistent" -s 30 -c 3 -o $OUT/${TAG}_ncu_gemm -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train --sustained 0 > $OUT/${TAG}_ncu_gemm.log 2>&1 ;;
    ncu_layer) timeout 900 $NCU --set full --import-source on -k regex:"graph_layer_fwd_sparse" -s 24 -c 6 -o $OUT/${TAG}_ncu_layer -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train --sustained 0 > $OUT/${TAG}_ncu_layer.log 2>&1 ;;
  esac
done
# keep what travels back under gpurun's 64 MiB cap: raw-page CSVs of every capture, the reports themselves only when small
for r in $OUT/${TAG}_*.ncu-rep; do [ -f "$r" ] || continue; ncu -i $r --page raw --csv > ${r%.ncu-rep}.raw.csv 2>/dev/null; sz=$(stat -c %s $r); if [ $sz -gt 12000000 ]; then rm -f $r; fi; done
kill $SMI
