/* sw_variants_b.cu -- ahead-of-time instances of the strip kernel (one slice of the variant table). */
#include "sw_variants.h"

namespace swk {
static const VariantEntry g_part[] = {
    SW_VARIANT_S16F(50, 1, 1, 3),
    SW_VARIANT_S16F2(25, 2, 1, 3),
    SW_VARIANT_S16(19, 2, 1, 4),
    SW_VARIANT_S16(15, 3, 1, 4),
};
VariantPart sw_variants_part_b() { return {g_part, (int)(sizeof(g_part) / sizeof(g_part[0]))}; }
}  // namespace swk
