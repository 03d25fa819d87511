/* sw_variants_c.cu -- ahead-of-time instances of the strip kernel (one slice of the variant table). */
#include "sw_variants.h"

namespace swk {
static const VariantEntry g_part[] = {
    SW_VARIANT_S16(30, 2, 1, 3),
    SW_VARIANT_S16F(64, 1, 1, 2),
    SW_VARIANT_S16F(32, 2, 1, 2),
};
VariantPart sw_variants_part_c() { return {g_part, (int)(sizeof(g_part) / sizeof(g_part[0]))}; }
}  // namespace swk
