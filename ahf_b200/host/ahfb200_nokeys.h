/* Pre-included (gcc -include) into the reference's main.c by build_dropin.sh for the -full builds.
 * In those builds nobody reads the keys of the host particle array (the device computes and sorts them), so main.c's key loop (:343-350)
 *     part->sfckey = sfc_curve_calcKey(global_info.ctype, x, y, z, BITS_PER_DIMENSION);
 * is reduced to a self-assignment the compiler drops; otherwise the loop pages in the whole calloc'ed array (0.3-0.5 s at 256^3) for
 * nothing.  The real prototype is declared first (include guard), then the call is redefined; if upstream renames the loop variable the
 * build fails loudly.  INTEGRATION.md shows the patch a maintainer would make instead. */
#ifndef AHFB200_NOKEYS_H
#define AHFB200_NOKEYS_H
#include "libsfc/sfc_curve.h"
#define sfc_curve_calcKey(ctype, x, y, z, bits) (part->sfckey)
#endif
