#!/bin/bash
# Builds ahf_b200/host/_build/AHF-b200: the reference's own main() / startrun / readers / tree / writers (compiled from the
# sources where they lie, unmodified, same flags as the reference's "Standard OpenMP" SYSTEM) with the hot-path call sites
# redirected to libahfgpu.so through ahf_glue.c.  Only possible where the reference sources exist (not on the GPU box; the
# binary travels with the repo snapshot).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REPO="$(cd "$HERE/../.." && pwd)"
REF="${AHF_REFERENCE_SRC:-/root/reference/src}"
OUT="$HERE/_build"
[ -d "$REF" ] || { echo "build_dropin.sh: $REF not present - keeping prebuilt $OUT" >&2; exit 0; }
build_one() {
name="$1"; mainflags="$2"; defs="$3"; iofileflags="$4"
mkdir -p "$OUT/obj"
CC="gcc -fopenmp -std=c99 -O2 -DWITH_OPENMP -DAHF $defs -w -I$REF -I$REPO/include"
pids=()
for f in "$REF"/*.c "$REF"/lib*/*.c; do
  base="$(basename "$(dirname "$f")")_$(basename "$f" .c)"
  extra=""
  case "$base" in
    src_main)         case "$mainflags" in *ahfb200_nokeys*) extra="-Dqsort=ahfb200_qsort $mainflags" ;; *) extra="-Dsfc_curve_calcKey=ahfb200_calcKey -Dqsort=ahfb200_qsort $mainflags" ;; esac ;;
    libahf_ahf_halos) extra="-Dahf_halos_sfc_constructHalo=ahfb200_constructHalo" ;;
    libio_io_file)    extra="$iofileflags" ;;
  esac
  $CC $extra -c "$f" -o "$OUT/obj/$base.o" &
  pids+=($!)
  if [ ${#pids[@]} -ge 8 ]; then wait "${pids[0]}"; pids=("${pids[@]:1}"); fi
done
wait
$CC -c "$HERE/ahf_glue.c" -o "$OUT/ahf_glue.o"
mv "$OUT/obj/src_main.o" "$OUT/"
ar rcs "$OUT/libref.a" "$OUT"/obj/*.o
gcc -fopenmp -o "$OUT/$name" "$OUT/src_main.o" "$OUT/ahf_glue.o" "$OUT/libref.a" -L"$REPO/ahf_b200" -lahfgpu -Wl,-rpath,'$ORIGIN/../..' -lm -ldl -lpthread
rm -rf "$OUT/obj" "$OUT/src_main.o" "$OUT/ahf_glue.o" "$OUT/libref.a"
}
# AHF-b200    : key/sort, mesh and halo loop on the GPU
# AHF-b200-kh : key/sort and halo loop on the GPU, mesh on the CPU (reference code)
build_one AHF-b200 "-Dgen_domgrids=ahfb200_gen_domgrids -Dll=ahfb200_ll -Dzero_dens=ahfb200_zero_dens -Dassign_npart=ahfb200_assign_npart -Dgen_AMRhierarchy=ahfb200_gen_AMRhierarchy"
build_one AHF-b200-kh ""
# AHF-b200-mm : the multi-species build (-DMULTIMASS -DGAS_PARTICLES), everything on the GPU
build_one AHF-b200-mm "-Dgen_domgrids=ahfb200_gen_domgrids -Dll=ahfb200_ll -Dzero_dens=ahfb200_zero_dens -Dassign_npart=ahfb200_assign_npart -Dgen_AMRhierarchy=ahfb200_gen_AMRhierarchy" "-DMULTIMASS -DGAS_PARTICLES"
# AHF-b200-full / AHF-b200-mm-full : additionally ahf_gridinfo and ahf_halos themselves (patch tables on the device, tree, halo pass, re-hash,
#                                    ordering and catalogue writers from the library: NEXT-1/2/3 of SURVEY 8f); no quads are rebuilt
# in the -full builds nobody reads the host AoS' keys (the device computes and sorts them): main.c's key loop (:343-350,
#   `part->sfckey = sfc_curve_calcKey(ctype, x, y, z, bits)`) is reduced to a self-assignment the compiler drops, so that the loop does not
#   page in the whole calloc'ed particle array (0.3-0.5 s at 256^3): ahfb200_nokeys.h, pre-included into main.c.  If upstream renames the
#   loop variable the build fails loudly.  The patch a maintainer would make instead is in INTEGRATION.md.
NOKEYS="-include $HERE/ahfb200_nokeys.h"
MESH="-Dgen_domgrids=ahfb200_gen_domgrids -Dll=ahfb200_ll -Dzero_dens=ahfb200_zero_dens -Dassign_npart=ahfb200_assign_npart -Dgen_AMRhierarchy=ahfb200_gen_AMRhierarchy"
# AHF-b200-full additionally reads single-file GADGET snapshots through the bulk ingest (NEXT-4): libio/io_file.c's call of io_gadget_readpart lands in the glue
build_one AHF-b200-full "$MESH -Dahf_gridinfo=ahfb200_gridinfo -Dahf_halos=ahfb200_halos $NOKEYS" "-DAHFB200_FULL" "-Dio_gadget_readpart=ahfb200_gadget_readpart"
build_one AHF-b200-mm-full "$MESH -Dahf_gridinfo=ahfb200_gridinfo -Dahf_halos=ahfb200_halos $NOKEYS" "-DMULTIMASS -DGAS_PARTICLES -DAHFB200_FULL"
ls -la "$OUT"
