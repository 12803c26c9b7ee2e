#!/bin/sh
# Compiles the part of the reference that CAN be compiled here — its C++ .delta reader and
# writer (lib/profiles_lib) — from the sources where they lie under /root/reference, into
# oracle/_ref/ (git-ignored, travels to the GPU box).  The arithmetic of the nucmer path is
# MUMmer 3.20, which the reference does not vendor, so there is no
# reference implementation of the hot path to build: "unbuildable".
set -e
REF=${PMN_REFERENCE:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
[ -d "$REF/lib/profiles_lib" ] || { echo "build_ref.sh: $REF not present, keeping prebuilt oracle/_ref" >&2; exit 0; }
mkdir -p "$HERE/_ref"
P="$REF/lib/profiles_lib"
g++ -O2 -std=c++11 -I"$P" "$P/m_delta.cc" "$P/m_delta_stream_test.cc" -o "$HERE/_ref/m_delta_stream_test"
g++ -O2 -std=c++11 -I"$P" "$P/m_delta.cc" "$HERE/ref_delta_roundtrip.cc" -o "$HERE/_ref/ref_delta_roundtrip"
