#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- builds the *unmodified* reference kernels as a CPU oracle.
#
# Cythonises mfrec/lib/{kmf_train,gd_estimator,als_implicit}.pyx straight from the read-only
# reference checkout (no source is copied into this repository) and compiles the
# generated C into oracle/_ref/ (git-ignored, NOT gpurun-ignored: the built
# extension modules travel to the GPU box, where /root/reference does not exist).
#
#   -2 : the .pyx files are Python-2 dialect (print statements, xrange)
#   the shipped *.c / *.so in the reference are Cython 0.19 / Mach-O py2.7 and unusable.
set -euo pipefail
REF="${MFREC_REFERENCE:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
PY="${PYTHON:-python}"
if [ ! -d "$REF/mfrec/lib" ]; then
  echo "build_ref: $REF not present; keeping prebuilt oracle/_ref (if any)" >&2
  exit 0
fi
mkdir -p "$OUT"
PYINC="$($PY -c 'import sysconfig;print(sysconfig.get_paths()["include"])')"
NPINC="$($PY -c 'import numpy;print(numpy.get_include())')"
SUFFIX="$($PY -c 'import sysconfig;print(sysconfig.get_config_var("EXT_SUFFIX"))')"
for m in kmf_train gd_estimator als_implicit; do
  if [ "$OUT/$m$SUFFIX" -nt "$REF/mfrec/lib/$m.pyx" ]; then continue; fi
  $PY -m cython -2 "$REF/mfrec/lib/$m.pyx" -o "$OUT/$m.c" >/dev/null 2>&1 || \
      $PY -m cython -2 "$REF/mfrec/lib/$m.pyx" -o "$OUT/$m.c"
  gcc -O2 -fPIC -shared -w -DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION \
      -I"$PYINC" -I"$NPINC" "$OUT/$m.c" -o "$OUT/$m$SUFFIX" -lm
  rm -f "$OUT/$m.c"
done
echo "build_ref: ok -> $OUT"
