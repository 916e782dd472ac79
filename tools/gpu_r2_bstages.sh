for b in 3 4 5 0; do echo "BSTAGES=$b"; TAG_TC_BSTAGES=$b python tools/run_exp.py tools/conv_microbench.py 2>&1 | grep -E "dil (1|8)"; done
