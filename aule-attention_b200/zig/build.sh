#!/bin/sh
# Builds/tests the Zig host mirrors against libaule.so where a Zig toolchain exists.
# (None in the B200 build image: `command -v zig` fails there, so this is skipped.)
set -e
cd "$(dirname "$0")"
if ! command -v zig >/dev/null 2>&1; then
    echo "zig not found: Zig host mirrors not built (C++ host in csrc/host is the tested implementation)"
    exit 0
fi
LIB=../python/aule/lib
zig test attention_gpu.zig -lc -L"$LIB" -laule -rpath "$LIB"
zig test compute_pipeline.zig -lc -L"$LIB" -laule -rpath "$LIB"
