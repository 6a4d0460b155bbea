#!/usr/bin/env bash
# Builds the PDL::B200 XS shim against a built PDL (default: the reference in oracle/_ref).
#   perl/PDL-B200/build.sh [blib_of_pdl]
# Output: perl/PDL-B200/blib/{lib/PDL/B200.pm, arch/auto/PDL/B200/B200.so}  (git-ignored .so)
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
PDLBLIB="${1:-$ROOT/oracle/_ref/blib}"
CORE_INC="$PDLBLIB/lib/PDL/Core"
[ -f "$CORE_INC/pdl.h" ] || { echo "no pdl.h under $CORE_INC (build oracle/_ref first)" >&2; exit 1; }
OUT="$HERE/blib"
mkdir -p "$OUT/lib/PDL" "$OUT/arch/auto/PDL/B200"
cp "$HERE/lib/PDL/B200.pm" "$OUT/lib/PDL/B200.pm"
PRIVLIB="$(perl -MConfig -e 'print $Config{privlibexp}')"
xsubpp -typemap "$PRIVLIB/ExtUtils/typemap" -typemap "$CORE_INC/typemap" "$HERE/B200.xs" > "$OUT/B200.c"
CCOPTS="$(perl -MExtUtils::Embed -e ccopts)"
gcc -O2 -g -fPIC -shared $CCOPTS -I"$CORE_INC" -I"$ROOT/include" -DXS_VERSION=\"0.01\" -DVERSION=\"0.01\" \
    "$OUT/B200.c" -o "$OUT/arch/auto/PDL/B200/B200.so" \
    -L"$ROOT/pdl_b200/lib" -lpdlb200 -Wl,-rpath,"$ROOT/pdl_b200/lib" -Wl,-rpath,'$ORIGIN/../../../../../../../pdl_b200/lib'
echo "built $OUT/arch/auto/PDL/B200/B200.so"
