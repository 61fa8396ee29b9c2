for U in 1 2 4; do for R in 1 4; do echo "U=$U ROWS=$R"; FNST_APPLY_U=$U FNST_APPLY_ROWS=$R python tools/bench_apply.py | cut -c1-110; done; done
