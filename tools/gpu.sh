#!/usr/bin/env bash
# Developer wrapper: ALWAYS rebuild the library before a gpurun call (the .so travels with the snapshot; a
# stale one silently tests old kernels), then forward the arguments to gpurun.
set -euo pipefail
cd "$(dirname "$0")/.."
make -C vit.triton_b200 -j"$(nproc)" >/dev/null
exec /usr/local/graft/bin/gpurun "$@"
