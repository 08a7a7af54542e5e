#!/bin/bash
# Installs the UNMODIFIED reference package into baseline/_ref (git-ignored; it travels to the GPU box with the snapshot).
# /root/reference is read-only and the wheel build writes into the source tree, so the install runs from a copy under /tmp;
# --no-deps because the reference's dependency set (ultralytics-yolov5, humanfriendly, jsonpickle, ...) is not in the offline
# wheelhouse.  The two packages it imports at module load time are stubbed in tests/stubs/.
set -e
REPO="$(cd "$(dirname "$0")/.." && pwd)"
SRC="${1:-/root/reference}"
TMP="$(mktemp -d)"
cp -r "$SRC" "$TMP/src"
cd "$TMP/src"
python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --no-deps --target "$REPO/baseline/_ref" .
rm -rf "$TMP"
