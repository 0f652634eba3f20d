#!/bin/bash
# Stage the UNMODIFIED reference scripts under the git-ignored baseline/_ref/ (it travels to the GPU box, it never enters
# history).  Only done in the build container, where /root/reference exists.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
SRC=${1:-/root/reference/image_model}
DST="$ROOT/baseline/_ref/image_model"
mkdir -p "$DST"
cp "$SRC"/*.py "$DST"/
rm -rf "$DST/diffusion" && cp -r "$SRC/diffusion" "$DST/diffusion"
(cd "$DST" && sha256sum *.py diffusion/*.py) > "$ROOT/baseline/_ref/image_model.sha256"
echo "staged $(ls "$DST"/*.py | wc -l) scripts into $DST"
