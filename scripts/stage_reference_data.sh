#!/bin/bash
# Stage the reference's bundled DATA files (not sources) where scripts/run_driver_e2e.py expects them on the GPU box:
# baseline/_ref is git-ignored but travels with gpurun.  Run in the build container (needs /root/reference).
set -e
REF=${1:-/root/reference}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$ROOT/baseline/_ref/intensity/data" "$ROOT/baseline/_ref/thermo"
cp "$REF"/intensity/data/*.nc "$ROOT/baseline/_ref/intensity/data/"
cp "$REF"/thermo/entropy_table.npz "$ROOT/baseline/_ref/thermo/"
ls -la "$ROOT/baseline/_ref/intensity/data" "$ROOT/baseline/_ref/thermo"
