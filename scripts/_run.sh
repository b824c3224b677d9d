bash scripts/ab_lib.sh scripts/ab/lib_minb8.so scripts/ab/lib_minb6.so scripts/ab/lib_minb5.so
