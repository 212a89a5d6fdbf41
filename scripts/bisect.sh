#!/bin/bash
# usage: bisect.sh T  -- runs the library variants under scripts/alt/ and the in-tree build under a few environments
T=${1:-4000}
run() { env "$@" timeout 300 python scripts/decode_speed.py $T "$*" 2>&1 | tail -2; }
run ER_SPLIT_HANDICAP=4
run ER_SPLIT_HANDICAP=7
run ER_SPLIT_HANDICAP=1
