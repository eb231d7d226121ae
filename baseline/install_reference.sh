#!/bin/bash
# Install the UNMODIFIED reference (beer-asr/beer) into baseline/_ref/ (git-ignored, NOT gpurun-ignored: it travels
# to the GPU box with the snapshot) so that `bench.py --impl reference` and the in-bench ELBO check can run the real
# reference there.  The source tree is read-only, so the wheel is built from a copy under /tmp.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="${1:-/root/reference}"
[ -d "$SRC/beer" ] || { echo "no reference checkout at $SRC"; exit 0; }
TMP="$(mktemp -d)"
cp -r "$SRC" "$TMP/src"
rm -rf "$HERE/_ref"
python -m pip install --quiet --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$HERE/_ref" "$TMP/src"
rm -rf "$TMP"
echo "installed $(ls "$HERE/_ref" | tr '\n' ' ')"
