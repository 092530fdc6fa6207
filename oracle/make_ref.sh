#!/usr/bin/env bash
# oracle/make_ref.sh — stage the UNMODIFIED reference modules of the hot path under oracle/_ref/ so that they travel to
# the GPU box (oracle/_ref/ is git-ignored, NOT gpurun-ignored; /root/reference itself does not exist there).
#
# TEST INFRASTRUCTURE.  Copies, per fork, only the packages SURVEY.md 8(a) cites:
#   rsl_rl/{__init__.py, modules/, algorithms/, storage/, utils/, env/}
# never rsl_rl/runners (wandb.init at import) nor rsl_rl/datasets (needs pybullet_utils + mocap files).
# Nothing is edited; the copies are used (a) as the caller in tests/test_gpu_reference_caller.py (the reference's own
# PPO + RolloutStorage driving this repo's ActorCritic: the plugin contract of SURVEY.md 8b) and (b) as the CPU arm of
# `bench.py --impl reference` (cpu_baseline.kind = "reference").
#
# Usage: bash oracle/make_ref.sh [/root/reference]     (re-run safe; called by __graft_entry__.build() when the
# reference tree is present)
set -euo pipefail
REF="${1:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
declare -A ROOTS=(
  [amp]="AMP/AMP_for_hardware-main/rsl_rl/rsl_rl"
  [cassie]="CASSIE/rsl_rl-master/rsl_rl"
  [humanoid]="HUMANOID/pbrs-humanoid-dev/gpu_rl/rsl_rl"
  [rma]="RMA/quadruped-robot-master/rl-robotics/rsl_rl/rsl_rl"
)
[ -d "$REF" ] || { echo "make_ref: $REF not found (GPU box?): keeping whatever is staged in $OUT"; exit 0; }
rm -rf "$OUT"
for fork in "${!ROOTS[@]}"; do
  src="$REF/${ROOTS[$fork]}"
  dst="$OUT/$fork/rsl_rl"
  mkdir -p "$dst"
  cp "$src/__init__.py" "$dst/"
  for sub in modules algorithms storage utils env; do
    [ -d "$src/$sub" ] && cp -r "$src/$sub" "$dst/$sub"
  done
  find "$dst" -name "__pycache__" -type d -prune -exec rm -rf {} +
done
( cd "$REF" && git rev-parse HEAD 2>/dev/null || echo "no-git" ) > "$OUT/SOURCE_COMMIT" || true
find "$OUT" -name "*.py" | sort | xargs sha256sum > "$OUT/SHA256SUMS"
echo "make_ref: staged $(find "$OUT" -name '*.py' | wc -l) reference files under $OUT"
