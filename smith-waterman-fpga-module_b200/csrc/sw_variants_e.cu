/* sw_variants_e.cu -- ahead-of-time instances of the strip kernel (one slice of the variant table). */
#include "sw_variants.h"

namespace swk {
static const VariantEntry g_part[] = {
    // G lanes per subject pair (systolic group, shuffles): small databases / few long pairs
    SW_VARIANT_S16(25, 1, 2, 5),
    SW_VARIANT_S16(75, 1, 2, 2),
    SW_VARIANT_S16(25, 3, 2, 2),
    SW_VARIANT_S16F(38, 1, 4, 3),
    SW_VARIANT_S16(19, 2, 4, 3),
    SW_VARIANT_S16(32, 1, 4, 4),
};
VariantPart sw_variants_part_e() { return {g_part, (int)(sizeof(g_part) / sizeof(g_part[0]))}; }
}  // namespace swk
