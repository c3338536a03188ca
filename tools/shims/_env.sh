# sourced by the shims: the repo root on PYTHONPATH, whichever directory the command runs in
SHIM_DIR=$(cd "$(dirname "$0")" && pwd)
TPP_ROOT=$(cd "$SHIM_DIR/../.." && pwd)
export PYTHONPATH="$TPP_ROOT${PYTHONPATH:+:$PYTHONPATH}"
