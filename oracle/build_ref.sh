#!/bin/sh
# TEST INFRASTRUCTURE — compiles the REFERENCE's own C++ index builders, from the source where it lies under /root/reference,
# into oracle/_ref/ (git-ignored, travels to the GPU box with the snapshot). The recipe is the reference's Makefile line
# (fengshen/data/megatron_dataloader/Makefile:1-9: g++ -O3 -shared -std=c++11 -fPIC + pybind11 includes), written out here
# because its own build system must not be run and its output must not land in the read-only tree. No source is copied.
set -e
REF="${FSB_REFERENCE_ROOT:-/root/reference}"
SRC="$REF/fengshen/data/megatron_dataloader/helpers.cpp"
HERE="$(cd "$(dirname "$0")" && pwd)"
[ -f "$SRC" ] || { echo "build_ref.sh: $SRC not found (reference tree absent: nothing to build)"; exit 0; }
mkdir -p "$HERE/_ref"
EXT="$(python3 -c 'import sysconfig; print(sysconfig.get_config_var("EXT_SUFFIX"))')"
g++ -O3 -Wall -shared -std=c++11 -fPIC $(python3 -m pybind11 --includes) "$SRC" -o "$HERE/_ref/helpers$EXT"
echo "built $HERE/_ref/helpers$EXT"
