#!/bin/bash
# TEST INFRASTRUCTURE.  Builds the UNMODIFIED reference (NegriAndrea/AHF) from the sources where they
# lie under /root/reference/src into oracle/_ref/ (git-ignored, travels to the GPU box as a binary):
#
#   oracle/_ref/ahf_ref        reference main() + oracle/ref_hooks.c pass-through wrappers (timers; state
#                              dumps when $AHF_DUMP_DIR is set).  Arithmetic untouched: the hooks are reached by
#                              compiling the CALLING translation units with -Dcallee=refhook_callee.
#   oracle/_ref/ahf_ref_mm     same, built with -DMULTIMASS -DGAS_PARTICLES (multi-species config)
#   oracle/_ref/ahf_ref_gt     same as ahf_ref plus the reference's own -DAHFgridtreefile option (define.h:108): ahf_halos writes
#                              <prefix>.AHF_gridtree -- centre, node / particle counts and tree links of every isolated refinement
#                              after RefCentre + analyseRef (ahf_halos.c:3066-3120); the pin of the NEXT-2 restatement
#
# Flags follow the reference's default SYSTEM "Standard OpenMP" (Makefile.config:17,200-208):
#   gcc -fopenmp -std=c99 -O2 -DWITH_OPENMP -DAHF
# The reference's own recursive make is NOT run; this is a flat gcc recipe over its .c files.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${AHF_REFERENCE_SRC:-/root/reference/src}"
OUT="$HERE/_ref"
if [ ! -d "$REF" ]; then
  echo "build_ref.sh: $REF not present (GPU box?) - keeping prebuilt $OUT" >&2
  exit 0
fi
mkdir -p "$OUT"

build_variant() {
  local name="$1"; shift
  local defs="$*"
  local obj="$OUT/obj_$name"
  mkdir -p "$obj"
  local CC="gcc -fopenmp -std=c99 -O2 -DWITH_OPENMP -DAHF $defs -w -I$REF"
  local pids=()
  for f in "$REF"/*.c "$REF"/lib*/*.c; do
    local base="$(basename "$(dirname "$f")")_$(basename "$f" .c)"
    local extra=""
    case "$base" in
      src_main)
        extra="-Dgen_domgrids=refhook_gen_domgrids -Dll=refhook_ll -Dzero_dens=refhook_zero_dens -Dassign_npart=refhook_assign_npart -Dahf_gridinfo=refhook_ahf_gridinfo -Dahf_halos=refhook_ahf_halos -Dsfc_curve_calcKey=refhook_calcKey -Dqsort=refhook_qsort" ;;
      libamr_serial_generate_grids)
        extra="-Drefine_grid=refhook_refine_grid -Drelink=refhook_relink -Dzero_dens=refhook_zero_dens -Dassign_npart=refhook_assign_npart" ;;
      libahf_ahf_halos)
        extra="-Dahf_halos_sfc_constructHalo=refhook_constructHalo -Dahf_io_WriteHalos=refhook_WriteHalos" ;;
      libahf_ahf_halos_sfc)
        extra="-Dsort_halo_particles=refhook_sort -Drem_outsideRvir=refhook_rvir -Drem_unbound=refhook_unbound -DHaloProfiles=refhook_profiles" ;;
    esac
    $CC $extra -c "$f" -o "$obj/$base.o" &
    pids+=($!)
    if [ ${#pids[@]} -ge 8 ]; then wait "${pids[0]}"; pids=("${pids[@]:1}"); fi
  done
  wait
  $CC -c "$HERE/ref_hooks.c" -o "$obj/ref_hooks.o"
  # link like the reference does (src/Makefile:56-57): main + static archive, so unused duplicate symbols are never pulled
  mv "$obj/src_main.o" "$obj/ref_hooks.o" "$OUT/"
  ar rcs "$obj/libref.a" "$obj"/*.o
  gcc -fopenmp -o "$OUT/$name" "$OUT/src_main.o" "$OUT/ref_hooks.o" "$obj/libref.a" -lm
  rm -f "$OUT/src_main.o" "$OUT/ref_hooks.o"
  rm -rf "$obj"
}

build_variant ahf_ref
build_variant ahf_ref_mm -DMULTIMASS -DGAS_PARTICLES
build_variant ahf_ref_gt -DAHFgridtreefile
ls -la "$OUT"
