#!/bin/bash
for c in 128 64 32 16; do
  FISS_DYN_MAT=1 FISS_BIG_FRAC=100 FISS_REC_CTAS=$c python tools/chain_probe2.py 2>&1 | tail -2
done
FISS_DYN_MAT=1 FISS_BIG_FRAC=85 FISS_REC_CTAS=32 python tools/chain_probe2.py 2>&1 | tail -2
FISS_DYN_MAT=1 FISS_BIG_FRAC=70 FISS_REC_CTAS=32 python tools/chain_probe2.py 2>&1 | tail -2
FISS_REC_CTAS=32 python tools/chain_probe2.py 2>&1 | tail -2
