#!/bin/bash
# Offline install of the UNMODIFIED reference into oracle/_ref (git-ignored, NOT
# gpurun-ignored: it travels to the GPU box so that bench.py can time the live reference
# beside the engine).  Checker / baseline only -- no product module imports it, and no
# reference source is committed to this repository.  /root/reference is read-only and
# setuptools writes build files into the source tree, hence the copy under /tmp.
set -euo pipefail
cd "$(dirname "$0")"
REF=${SEM_REFERENCE_SRC:-/root/reference}
[ -d "$REF/sem" ] || { echo "install_ref.sh: no reference tree at $REF (nothing to do)"; exit 0; }
TMP=$(mktemp -d /tmp/sem_ref_XXXXXX)
cp -r "$REF/." "$TMP/"
rm -rf _ref
python -m pip install --quiet --no-index --no-build-isolation --no-deps --target _ref "$TMP"
rm -rf "$TMP"
echo "installed the reference into $(pwd)/_ref"
