/* sw_variants_a.cu -- ahead-of-time instances of the strip kernel (one slice of the variant table). */
#include "sw_variants.h"

namespace swk {
static const VariantEntry g_part[] = {
    SW_VARIANT_S16F(30, 1, 1, 4),
    SW_VARIANT_S16F(38, 1, 1, 4),
    SW_VARIANT_S16F(75, 1, 1, 2),
    SW_VARIANT_S16F(32, 1, 1, 4),
};
VariantPart sw_variants_part_a() { return {g_part, (int)(sizeof(g_part) / sizeof(g_part[0]))}; }
}  // namespace swk
