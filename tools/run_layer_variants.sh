echo "== baseline"; timeout 120 python tools/layer_bench.py 2>&1 | grep -E " rows "
for f in build_variants/*.so; do echo "== $f"; timeout 120 python tools/layer_bench.py $f 2>&1 | grep -E " rows "; done
