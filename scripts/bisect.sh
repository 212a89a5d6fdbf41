#!/bin/bash
# usage: bisect.sh T  -- runs the library variants under scripts/alt/ and the in-tree build under a few environments
T=${1:-4000}
run() { env "$@" timeout 300 python scripts/decode_speed.py $T "$*" 2>&1 | tail -2; }
run ER_DECODE_LL=1 ER_DECODE_HINT=1
run ER_DECODE_LL=1 ER_DECODE_HINT=0
run ER_DECODE_LL=0
run ER_LIB=scripts/alt/lib_attn_noinline.so ER_DECODE_LL=1
