python -m pytest tests/test_gpr_fit.py -m gpu -x -q 2>&1 | tail -5
STL_K1_CLK=1 python scripts/k1_clk.py 600 0.2 2>&1 | grep "K1 mean" | tail -2
STL_DEBUG_STATS=1 python scripts/paths.py 300 0.2 2>&1 | grep "search paths" | tail -1
for spec in "k1:k_assoc2d:2" "k2:k_nn_knn:2" "lmb:k_lm_knn_b:2" "lin:k_linearize:2"; do
  IFS=: read name pat skip <<< "$spec"
  ncu --set full --clock-control none --import-source on -k regex:$pat -s $skip -c 1 -o gpurun_out/r02c_${name}_full -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02c_ncu_$name.log 2>&1
  ncu -i gpurun_out/r02c_${name}_full.ncu-rep --page source --csv > gpurun_out/r02c_${name}_src.csv 2>/dev/null
done
ls -la gpurun_out/r02c_*
