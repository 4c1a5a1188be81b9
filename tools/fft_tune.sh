#!/bin/bash
# sweep of the lines-per-block parameters of the fast Poisson kernels (256^3)
for cfg in "8 8" "8 16" "8 32" "16 16" "4 8"; do set -- $cfg
  echo -n "TL=$1 TX=$2  "; OB200_FFT_TL=$1 OB200_FFT_TX=$2 python tools/poisson_sweep.py 256 2>&1 | grep -o '"ms_per_solve": [0-9.]*\|residual_max_rel": [0-9.e-]*' | tr '\n' ' '; echo
done
