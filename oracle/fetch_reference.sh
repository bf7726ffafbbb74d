#!/usr/bin/env bash
# TEST INFRASTRUCTURE.  Copies the UNMODIFIED reference package (cmeraki/vit.triton, /root/reference/vit) into
# baseline/_ref/ (git-ignored, NOT gpurun-ignored: it travels to the GPU box, where /root/reference does not
# exist) so that tools/run_reference_triton.py can run the reference's own Triton forward on the B200 next
# to ours.  Nothing under baseline/_ref/ is imported by the product or committed.
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
SRC="${1:-/root/reference}"
if [ ! -d "$SRC/vit" ]; then
  echo "reference not found at $SRC (this script runs in the build container only)" >&2
  exit 0
fi
mkdir -p "$ROOT/baseline/_ref"
rm -rf "$ROOT/baseline/_ref/vit"
cp -r "$SRC/vit" "$ROOT/baseline/_ref/vit"
find "$ROOT/baseline/_ref" -name __pycache__ -type d -prune -exec rm -rf {} +
echo "reference copied to baseline/_ref/vit"
