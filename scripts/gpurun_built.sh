#!/bin/bash
# builds libtophat_b200.so + host binaries HERE (nvcc cross-compiles), then runs the given command on a B200 box
set -e
cd "$(dirname "$0")/.."
python -c "from tophat_b200 import build; build.build_all()" 2>&1 | grep -E "error|Error" && exit 1
exec /usr/local/graft/bin/gpurun "$@"
