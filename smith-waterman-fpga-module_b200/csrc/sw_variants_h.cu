/* sw_variants_h.cu -- ahead-of-time instances of the strip kernel (one slice of the variant table). */
#include "sw_variants.h"

namespace swk {
static const VariantEntry g_part[] = {
    // interior trips (FL = 31: predicate-free trips over the interior columns of a warp, no L1 prefetch
    // there, first trip included, one column per loop trip in the general path).  Measured on 4 M x 150 nt
    // x 100 queries (profiles/r02_variant_ab_interior.jsonl): 8 973 vs 8 833 (R25x2_G1_U8) vs 8 675 GCUPS
    // (R25x2_G1).  Measured and dropped: the same for R19x2 at four blocks per SM (8 582), for 8-column
    // trips (8 901: spills) and for the long-query instances R38x2 / R32x2 (two warps per scheduler: the
    // boundary row needs its L1 prefetch, 7 150 vs 7 649 GCUPS on a 10 kb query).
    SW_VARIANT_S16F_UF(25, 2, 1, 3, 4, 31),
};
VariantPart sw_variants_part_h() { return {g_part, (int)(sizeof(g_part) / sizeof(g_part[0]))}; }
}  // namespace swk
