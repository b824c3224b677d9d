bash scripts/ab_bench.sh "STL_SUB=4" "STL_SUB=2" "STL_SUB=3" "STL_SUB=6"
