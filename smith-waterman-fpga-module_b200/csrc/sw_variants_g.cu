/* sw_variants_g.cu -- ahead-of-time instances of the strip kernel (one slice of the variant table). */
#include "sw_variants.h"

namespace swk {
static const VariantEntry g_part[] = {
    // 8 columns per trip of the step loop
    SW_VARIANT_S16F_U(25, 2, 1, 3, 8),
    SW_VARIANT_S16F_U(25, 3, 1, 2, 8),
    SW_VARIANT_S16F_U(38, 2, 1, 2, 8),
    // measured and dropped (profiles/r02_variant_ab_chains.jsonl): three / four interleaved chains per lane at
    // three resident blocks per SM -- R17x3 8 571, R17x3_U8 8 491, R13x4 8 198 vs R25x2_U8 8 849 GCUPS
};
VariantPart sw_variants_part_g() { return {g_part, (int)(sizeof(g_part) / sizeof(g_part[0]))}; }
}  // namespace swk
