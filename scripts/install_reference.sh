#!/usr/bin/env bash
# Installs the UNMODIFIED reference (merrymercy/nums) into the git-ignored baseline/_ref/ so that it
# travels to the GPU box with the repo snapshot (baseline/_ref is not in .gpurunignore):
#   * the `nums` package via pip --target (pure Python; --no-deps because it pins numpy<=1.20 / ray<1.1,
#     --ignore-requires-python because setup.py says python<3.9) -- nums_b200.reference_compat supplies the
#     numpy-2 / no-ray shims at import time, nothing in the installed tree is edited;
#   * the reference's own test files (tests/ is not part of the wheel) into baseline/_ref/reference_tests/,
#     with OUR conftest (scripts/ref_conftest.py: the reference's get_app() plus a "cuda" mode) in place of
#     theirs, which hard-requires a live ray.
# Nothing under baseline/_ref/ is tracked by git.
set -euo pipefail
REPO="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
SRC="${1:-/root/reference}"
DST="$REPO/baseline/_ref"
if [ ! -d "$SRC/nums/core" ]; then
    echo "reference not found at $SRC" >&2
    exit 1
fi
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
cp -r "$SRC" "$TMP/src"            # the build writes egg-info into the source tree; /root/reference is read-only
rm -rf "$DST"
mkdir -p "$DST"
python -m pip install --quiet --no-index --no-build-isolation --no-deps --ignore-requires-python \
    --find-links /opt/wheelhouse --target "$DST" "$TMP/src"
cp -r "$SRC/tests" "$DST/reference_tests"
find "$DST" -name __pycache__ -type d -prune -exec rm -rf {} +
cp "$REPO/scripts/ref_conftest.py" "$DST/reference_tests/conftest.py"
echo "installed reference into $DST"
