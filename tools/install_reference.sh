#!/bin/bash
# Copies the reference's PPO consumer (IsaacGymEnvs/algorithms/: ppo_asymmetry.py, nets_asymmetry.py, buffer_asymmetry.py)
# UNMODIFIED into the git-ignored baseline/_ref/ so that it travels to the GPU box (gpurun ships baseline/_ref, git does not).
# Used by tools/run_reference_ppo.py and tests/test_reference_consumer_gpu.py: the reference's own PPO.run drives FpvVecTask.
# Build container only (needs /root/reference).
set -e
REF=${TACO_REFERENCE:-/root/reference}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
[ -d "$REF/IsaacGymEnvs/algorithms" ] || { echo "reference tree not found at $REF"; exit 0; }
mkdir -p "$ROOT/baseline/_ref"
rm -rf "$ROOT/baseline/_ref/algorithms"
cp -r "$REF/IsaacGymEnvs/algorithms" "$ROOT/baseline/_ref/algorithms"
find "$ROOT/baseline/_ref/algorithms" -name "__pycache__" -prune -exec rm -rf {} +
[ -f "$ROOT/baseline/_ref/algorithms/__init__.py" ] || touch "$ROOT/baseline/_ref/algorithms/__init__.py"
( cd "$REF/IsaacGymEnvs/algorithms" && sha256sum *.py ) > "$ROOT/baseline/_ref/algorithms.sha256"
echo "installed: $(ls "$ROOT/baseline/_ref/algorithms" | tr '\n' ' ')"
