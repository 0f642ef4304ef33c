#!/bin/bash
mkdir -p gpurun_out
for v in none fm bm fmbm fmbmhfm; do
  echo -n "skip=$v: "; TACO_PPO_DEBUG_SKIP=$v timeout 200 python tools/ppo_native_time.py 3 --no-lip 2>/dev/null
done
echo "critic default:"; timeout 100 python tools/critic_bench.py 262144 2>/dev/null | cut -c1-300
echo "critic GATES=2 (no transcendentals):"; TACO_B200_LIB=$PWD/taco_b200/lib/libtaco_b200_mb6.so timeout 100 python tools/critic_bench.py 262144 2>/dev/null | cut -c1-300
