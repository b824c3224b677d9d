python -m pytest tests -m gpu -x -q 2>&1 | tail -4
bash scripts/ab_bench.sh "STL_LM_UNFUSED=1" "STL_X=0" 2>&1 | tail -8
