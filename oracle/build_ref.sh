#!/usr/bin/env bash
# TEST INFRASTRUCTURE — builds the UNMODIFIED reference (PDL) into oracle/_ref/.
#
# PDL's hot path is C that PDL::PP *generates* from lib/PDL/{Ops,Ufunc,Primitive}.pd
# at build time, so "gcc on a few files" cannot produce it; this recipe therefore
# drives the reference's own Makefile.PL in a throw-away scratch copy (the
# reference tree is read-only) and keeps only the built blib/ tree.  No reference
# source is copied into the repository: oracle/_ref/ is git-ignored.
#
# Usage: oracle/build_ref.sh [reference_dir]     (default /root/reference)
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${1:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF" ]; then
  echo "build_ref: $REF absent; keeping any prebuilt $OUT" >&2
  exit 0
fi
if [ -f "$OUT/blib/arch/auto/PDL/Ufunc/Ufunc.so" ] && [ -z "${FORCE:-}" ]; then
  echo "build_ref: $OUT already built (FORCE=1 to rebuild)"; exit 0
fi
W="$(mktemp -d /tmp/pdlref.XXXXXX)"
trap 'rm -rf "$W"' EXIT
cp -r "$REF" "$W/src" && chmod -R u+w "$W/src"
cd "$W/src"
export PERL5LIB="$HERE/shim${PERL5LIB:+:$PERL5LIB}"
perl Makefile.PL > "$W/configure.log" 2>&1
make -j"$(nproc)" core > "$W/make.log" 2>&1 || { tail -50 "$W/make.log"; exit 1; }
rm -rf "$OUT"; mkdir -p "$OUT"
cp -r blib "$OUT/blib"
# generated C for the ops on the hot path: kept (ignored by git) so the restatement
# in oracle/pdl_oracle.c can be audited against what PP actually emitted
mkdir -p "$OUT/gen"
cp lib/PDL/Ops-pp-plus.c lib/PDL/Ops-pp-divide.c lib/PDL/Ops-pp-modulo.c lib/PDL/Ops-pp-sqrt.c \
   lib/PDL/Ufunc-pp-sumover.c lib/PDL/Ufunc-pp-average.c lib/PDL/Ufunc-pp-minimum.c \
   lib/PDL/Ufunc-pp-prodover.c lib/PDL/Ufunc-pp-maximum_ind.c \
   lib/PDL/Primitive-pp-matmult.c "$OUT/gen/" 2>/dev/null || true
# the reference's own hot-path tests, so they can be re-run on the GPU box with the shim attached
# (oracle/_ref is git-ignored: nothing from the reference enters the repository history)
mkdir -p "$OUT/t"
for f in ops.t ops-bitwise.t ufunc.t bad.t primitive-matmult.t thread.t slice.t core.t clump.t reduce.t; do
  cp "t/$f" "$OUT/t/" 2>/dev/null || true
done
find "$OUT" -name '*.pod' -delete
perl -I"$OUT/blib/lib" -I"$OUT/blib/arch" -MPDL::LiteF -e \
  'print "oracle/_ref: PDL $PDL::VERSION pthreads=", PDL::Core::pthreads_enabled(), " cpus=", PDL::Core::online_cpus(), "\n"'
